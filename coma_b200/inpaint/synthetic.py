"""Synthetic checkpoints: seeded random state dicts with diffusers' key names and the SD-1.5-inpainting / SD-VAE shapes
(`UNet2DConditionModel`, `AutoencoderKL` as loaded by the reference at src/generation/inpaint.py:59-64).

No weights are available offline, so the benchmark (`bench.py` HOI leg), `__graft_entry__.smoke()`, the profiling tools and the
tests instantiate the real architecture with `randn * fan_in^-1/2` weights (SURVEY.md 8d "Synthetic inputs"). This is a data
generator, not an algorithm: the torch restatement of the layers lives in oracle/sd_oracle.py, which re-exports these names so both
sides of a parity test see the same tensors."""
import torch

UNET_CFG = dict(in_channels=9, out_channels=4, block_out_channels=(320, 640, 1280, 1280), layers_per_block=2, heads=8,
                cross_attention_dim=768, groups=32, attn_levels=(True, True, True, False))
VAE_CFG = dict(in_channels=3, latent_channels=4, block_out_channels=(128, 256, 512, 512), layers_per_block=2, groups=32,
               scaling_factor=0.18215)


def tiny_unet_cfg():
    """Same topology at toy width — for fast tests."""
    return dict(in_channels=9, out_channels=4, block_out_channels=(32, 64, 64, 64), layers_per_block=2, heads=2,
                cross_attention_dim=64, groups=8, attn_levels=(True, True, True, False))


def tiny_vae_cfg():
    return dict(in_channels=3, latent_channels=4, block_out_channels=(32, 32, 64, 64), layers_per_block=2, groups=8,
                scaling_factor=0.18215)


# ------------------------------------------------------------------------------------------------ random state dicts
class _Init:
    def __init__(self, seed):
        self.g = torch.Generator().manual_seed(seed)
        self.sd = {}

    def conv(self, name, cin, cout, k):
        self.sd[name + ".weight"] = torch.randn((cout, cin, k, k), generator=self.g) * (cin * k * k) ** -0.5
        self.sd[name + ".bias"] = torch.randn(cout, generator=self.g) * 0.05

    def linear(self, name, cin, cout, bias=True):
        self.sd[name + ".weight"] = torch.randn((cout, cin), generator=self.g) * cin ** -0.5
        if bias:
            self.sd[name + ".bias"] = torch.randn(cout, generator=self.g) * 0.05

    def norm(self, name, c):
        self.sd[name + ".weight"] = 1.0 + 0.1 * torch.randn(c, generator=self.g)
        self.sd[name + ".bias"] = 0.05 * torch.randn(c, generator=self.g)

    def resnet(self, name, cin, cout, temb):
        self.norm(name + ".norm1", cin)
        self.conv(name + ".conv1", cin, cout, 3)
        if temb:
            self.linear(name + ".time_emb_proj", temb, cout)
        self.norm(name + ".norm2", cout)
        self.conv(name + ".conv2", cout, cout, 3)
        if cin != cout:
            self.conv(name + ".conv_shortcut", cin, cout, 1)

    def transformer(self, name, c, ctx):
        self.norm(name + ".norm", c)
        self.conv(name + ".proj_in", c, c, 1)
        t = name + ".transformer_blocks.0"
        for i, kv in ((1, c), (2, ctx)):
            self.norm(f"{t}.norm{i}", c)
            self.linear(f"{t}.attn{i}.to_q", c, c, bias=False)
            self.linear(f"{t}.attn{i}.to_k", kv, c, bias=False)
            self.linear(f"{t}.attn{i}.to_v", kv, c, bias=False)
            self.linear(f"{t}.attn{i}.to_out.0", c, c)
        self.norm(f"{t}.norm3", c)
        self.linear(f"{t}.ff.net.0.proj", c, 8 * c)
        self.linear(f"{t}.ff.net.2", 4 * c, c)
        self.conv(name + ".proj_out", c, c, 1)


def make_unet_state_dict(seed=0, cfg=UNET_CFG):
    it = _Init(seed)
    ch, L, ctx = cfg["block_out_channels"], cfg["layers_per_block"], cfg["cross_attention_dim"]
    temb = ch[0] * 4
    it.linear("time_embedding.linear_1", ch[0], temb)
    it.linear("time_embedding.linear_2", temb, temb)
    it.conv("conv_in", cfg["in_channels"], ch[0], 3)
    cin = ch[0]
    for i, cout in enumerate(ch):
        for j in range(L):
            it.resnet(f"down_blocks.{i}.resnets.{j}", cin, cout, temb)
            if cfg["attn_levels"][i]:
                it.transformer(f"down_blocks.{i}.attentions.{j}", cout, ctx)
            cin = cout
        if i < len(ch) - 1:
            it.conv(f"down_blocks.{i}.downsamplers.0.conv", cout, cout, 3)
    it.resnet("mid_block.resnets.0", ch[-1], ch[-1], temb)
    it.transformer("mid_block.attentions.0", ch[-1], ctx)
    it.resnet("mid_block.resnets.1", ch[-1], ch[-1], temb)
    rev = list(reversed(ch))
    prev = rev[0]
    for i, cout in enumerate(rev):
        skip_in = rev[min(i + 1, len(ch) - 1)]
        for j in range(L + 1):
            skip = skip_in if j == L else cout
            rin = prev if j == 0 else cout
            it.resnet(f"up_blocks.{i}.resnets.{j}", rin + skip, cout, temb)
            if list(reversed(cfg["attn_levels"]))[i]:
                it.transformer(f"up_blocks.{i}.attentions.{j}", cout, ctx)
        if i < len(ch) - 1:
            it.conv(f"up_blocks.{i}.upsamplers.0.conv", cout, cout, 3)
        prev = cout
    it.norm("conv_norm_out", ch[0])
    it.conv("conv_out", ch[0], cfg["out_channels"], 3)
    return it.sd


def make_vae_state_dict(seed=1, cfg=VAE_CFG):
    it = _Init(seed)
    ch, L, lat = cfg["block_out_channels"], cfg["layers_per_block"], cfg["latent_channels"]

    def attn(name, c):
        it.norm(name + ".group_norm", c)
        for k in ("to_q", "to_k", "to_v", "to_out.0"):
            it.linear(f"{name}.{k}", c, c)

    it.conv("encoder.conv_in", cfg["in_channels"], ch[0], 3)
    cin = ch[0]
    for i, cout in enumerate(ch):
        for j in range(L):
            it.resnet(f"encoder.down_blocks.{i}.resnets.{j}", cin, cout, 0)
            cin = cout
        if i < len(ch) - 1:
            it.conv(f"encoder.down_blocks.{i}.downsamplers.0.conv", cout, cout, 3)
    it.resnet("encoder.mid_block.resnets.0", ch[-1], ch[-1], 0)
    attn("encoder.mid_block.attentions.0", ch[-1])
    it.resnet("encoder.mid_block.resnets.1", ch[-1], ch[-1], 0)
    it.norm("encoder.conv_norm_out", ch[-1])
    it.conv("encoder.conv_out", ch[-1], 2 * lat, 3)
    it.conv("quant_conv", 2 * lat, 2 * lat, 1)
    it.conv("post_quant_conv", lat, lat, 1)
    it.conv("decoder.conv_in", lat, ch[-1], 3)
    it.resnet("decoder.mid_block.resnets.0", ch[-1], ch[-1], 0)
    attn("decoder.mid_block.attentions.0", ch[-1])
    it.resnet("decoder.mid_block.resnets.1", ch[-1], ch[-1], 0)
    rev = list(reversed(ch))
    cin = rev[0]
    for i, cout in enumerate(rev):
        for j in range(L + 1):
            it.resnet(f"decoder.up_blocks.{i}.resnets.{j}", cin, cout, 0)
            cin = cout
        if i < len(ch) - 1:
            it.conv(f"decoder.up_blocks.{i}.upsamplers.0.conv", cout, cout, 3)
    it.norm("decoder.conv_norm_out", ch[0])
    it.conv("decoder.conv_out", ch[0], cfg["in_channels"], 3)
    return it.sd
