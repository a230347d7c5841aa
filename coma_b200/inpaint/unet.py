"""SD-1.5 inpainting UNet (diffusers `UNet2DConditionModel`: in 9 / out 4, blocks (320,640,1280,1280), 2 layers per block,
8 heads, cross-attention dim 768 — the model called at utils/adaptive_mask_inpainting.py:1001-1007) on the B200 kernels.
Weights are taken from a diffusers-named state dict; activations are NHWC fp16, accumulation fp32."""
import torch

from . import nn
from .nn import Act, F16, F32

SD15_INPAINT = dict(in_channels=9, out_channels=4, block_out_channels=(320, 640, 1280, 1280), layers_per_block=2, heads=8,
                    cross_attention_dim=768, groups=32, attn_levels=(True, True, True, False))


class UNet:
    def __init__(self, state_dict, cfg=SD15_INPAINT, device="cuda"):
        self.cfg, self.dev = dict(cfg), torch.device(device)
        if self.dev.type != "cuda":
            raise RuntimeError("coma_b200 UNet runs on CUDA (sm_100a) only — there is no CPU fallback")
        self.p = {}
        sd = state_dict
        tproj_w, tproj_b, self.tproj_slices = [], [], {}
        off = 0
        for k, v in sd.items():
            base = k.rsplit(".", 1)[0]
            if k.endswith(".weight"):
                if k.endswith(".ff.net.0.proj.weight"):
                    fused = nn.prep_geglu(v, sd[base + ".bias"], self.dev)
                    if fused is not None:
                        self.p[base.rsplit(".net.0.proj", 1)[0] + ".geglu"] = fused
                        continue
                if ".time_emb_proj" in k:      # all 22 projections run as ONE GEMM per step
                    tproj_w.append(v)
                    tproj_b.append(sd[base + ".bias"])
                    self.tproj_slices[base] = (off, off + v.shape[0])
                    off += v.shape[0]
                elif v.dim() == 4 and v.shape[-1] == 3:
                    self.p[k] = nn.prep_conv3x3(v, self.dev)
                elif v.dim() in (2, 4):
                    self.p[k] = nn.prep_linear(v, self.dev)
                else:
                    self.p[k] = nn.prep_vec(v, self.dev)
            elif ".time_emb_proj" not in k:
                self.p[k] = nn.prep_vec(v, self.dev)
        self.tproj_w = nn.prep_linear(torch.cat(tproj_w, 0), self.dev)
        self.tproj_b = nn.prep_vec(torch.cat(tproj_b, 0), self.dev)
        # row-concatenated projection weights: self-attention q|k|v in one GEMM, cross-attention k|v in one GEMM
        self.cross_layers = []
        for k in [k for k in self.p if k.endswith(".attn1.to_q.weight")]:
            a = k[: -len(".to_q.weight")]
            self.p[a + ".to_qkv.weight"] = torch.cat([self.p.pop(a + ".to_q.weight"), self.p.pop(a + ".to_k.weight"),
                                                      self.p.pop(a + ".to_v.weight")], 0).contiguous()
        for k in [k for k in self.p if k.endswith(".attn2.to_k.weight")]:
            a = k[: -len(".to_k.weight")]
            self.p[a + ".to_kv.weight"] = torch.cat([self.p.pop(a + ".to_k.weight"), self.p.pop(a + ".to_v.weight")], 0).contiguous()
            self.cross_layers.append(a)

    # ------------------------------------------------------------------------------------------------
    def _resnet(self, x: Act, name, tproj):
        p, G = self.p, self.cfg["groups"]
        gn1 = nn.gn_affine(x, p[name + ".norm1.weight"], p[name + ".norm1.bias"], G, 1e-5)
        a, b = self.tproj_slices[name + ".time_emb_proj"]
        h = nn.conv3x3(x, p[name + ".conv1.weight"], p[name + ".conv1.bias"], gn=gn1, act=1, stats=True,
                       bias_rows=tproj[:, a:b])
        gn2 = nn.gn_affine(h, p[name + ".norm2.weight"], p[name + ".norm2.bias"], G, 1e-5)
        if name + ".conv_shortcut.weight" in p:
            sc = nn.gemm(x.t, p[name + ".conv_shortcut.weight"], p[name + ".conv_shortcut.bias"])
        else:
            sc = x.t
        return nn.conv3x3(h, p[name + ".conv2.weight"], p[name + ".conv2.bias"], gn=gn2, act=1, residual=sc, stats=True)

    def context_kv(self, ctx, L, B):
        """Keys / transposed values of every cross-attention layer for a context [B*L, cross_dim] f16 — constant over the
        denoising loop (the prompt embeddings do not change between steps), so the pipeline calls this once per image batch."""
        heads = self.cfg["heads"]
        return {a: nn.project_kv(ctx, B, L, self.p[a + ".to_kv.weight"], heads) for a in self.cross_layers}

    def _transformer(self, x: Act, ctx, L, name, ctx_kv=None):
        p, G, heads = self.p, self.cfg["groups"], self.cfg["heads"]
        B, S = x.B, x.H * x.W
        s, sh = nn.gn_affine(x, p[name + ".norm.weight"], p[name + ".norm.bias"], G, 1e-6)
        h = nn.gemm(nn.affine_act(x, s, sh, 0).t, p[name + ".proj_in.weight"], p[name + ".proj_in.bias"])
        t = name + ".transformer_blocks.0"
        n1 = nn.layernorm(h, p[t + ".norm1.weight"], p[t + ".norm1.bias"])
        h = nn.attention(n1, n1, B, S, S, None, None, None, p[t + ".attn1.to_out.0.weight"], p[t + ".attn1.to_out.0.bias"], heads, h,
                         wqkv=p[t + ".attn1.to_qkv.weight"])
        n2 = nn.layernorm(h, p[t + ".norm2.weight"], p[t + ".norm2.bias"])
        h = nn.attention(n2, ctx, B, S, L, p[t + ".attn2.to_q.weight"], None, None, p[t + ".attn2.to_out.0.weight"],
                         p[t + ".attn2.to_out.0.bias"], heads, h, wkv=p[t + ".attn2.to_kv.weight"],
                         kv=None if ctx_kv is None else ctx_kv[t + ".attn2"])
        n3 = nn.layernorm(h, p[t + ".norm3.weight"], p[t + ".norm3.bias"])
        if t + ".ff.geglu" in p:     # projection + GEGLU in one kernel (interleaved weight rows)
            ff = nn.gemm_geglu(n3, *p[t + ".ff.geglu"])
        else:
            ff = nn.geglu(nn.gemm(n3, p[t + ".ff.net.0.proj.weight"], p[t + ".ff.net.0.proj.bias"]))
        h = nn.gemm(ff, p[t + ".ff.net.2.weight"], p[t + ".ff.net.2.bias"], residual=h)
        out = nn.gemm(h, p[name + ".proj_out.weight"], p[name + ".proj_out.bias"], residual=x.t)
        return Act(out, x.B, x.H, x.W)

    def time_projections(self, t):
        """t [B] f32 -> every ResnetBlock's time_emb_proj(silu(temb)) as one [B, sum Cout] fp32 matrix."""
        p = self.p
        e = nn.timestep_embedding(t, self.cfg["block_out_channels"][0])
        e = nn.gemm(e, p["time_embedding.linear_1.weight"], p["time_embedding.linear_1.bias"], act=1)
        e = nn.gemm(e, p["time_embedding.linear_2.weight"], p["time_embedding.linear_2.bias"])
        return nn.gemm(nn.silu(e), self.tproj_w, self.tproj_b, out_dtype=F32)

    def forward(self, x: Act, t, ctx, L=77, taps=None, ctx_kv=None):
        """x: Act [B, h, w, 9]; t [B] f32 (cuda); ctx [B*L, cross_dim] f16 -> eps: fp32 tensor [B*h*w, 4].
        ctx_kv = context_kv(ctx, L, B) skips the per-step cross-attention key / value projections."""
        cfg, p = self.cfg, self.p
        ch, nl = cfg["block_out_channels"], cfg["layers_per_block"]
        tproj = self.time_projections(t)
        h = nn.conv3x3(x, p["conv_in.weight"], p["conv_in.bias"])
        skips = [h]
        for i in range(len(ch)):
            for j in range(nl):
                h = self._resnet(h, f"down_blocks.{i}.resnets.{j}", tproj)
                if cfg["attn_levels"][i]:
                    h = self._transformer(h, ctx, L, f"down_blocks.{i}.attentions.{j}", ctx_kv)
                skips.append(h)
            if i < len(ch) - 1:
                h = nn.conv3x3(h, p[f"down_blocks.{i}.downsamplers.0.conv.weight"], p[f"down_blocks.{i}.downsamplers.0.conv.bias"], stride=2, stats=True)
                skips.append(h)
        if taps is not None:
            taps["down"] = h
        h = self._resnet(h, "mid_block.resnets.0", tproj)
        h = self._transformer(h, ctx, L, "mid_block.attentions.0", ctx_kv)
        h = self._resnet(h, "mid_block.resnets.1", tproj)
        if taps is not None:
            taps["mid"] = h
        ral = list(reversed(cfg["attn_levels"]))
        for i in range(len(ch)):
            for j in range(nl + 1):
                s = skips.pop()
                cat = torch.empty((h.M, h.C + s.C), dtype=F16, device=self.dev)   # skip concat (channel-wise, NHWC)
                cat[:, : h.C].copy_(h.t)
                cat[:, h.C:].copy_(s.t)
                h = self._resnet(Act(cat, h.B, h.H, h.W), f"up_blocks.{i}.resnets.{j}", tproj)
                if ral[i]:
                    h = self._transformer(h, ctx, L, f"up_blocks.{i}.attentions.{j}", ctx_kv)
            if i < len(ch) - 1:
                h = nn.conv3x3(h, p[f"up_blocks.{i}.upsamplers.0.conv.weight"], p[f"up_blocks.{i}.upsamplers.0.conv.bias"], up=True, stats=True)
        if taps is not None:
            taps["up"] = h
        gn = nn.gn_affine(h, p["conv_norm_out.weight"], p["conv_norm_out.bias"], cfg["groups"], 1e-5)
        out = nn.conv3x3(h, p["conv_out.weight"], p["conv_out.bias"], gn=gn, act=1, out_dtype=F32)
        return out.t
