"""SD-1.x AutoencoderKL (blocks (128,256,512,512), latent 4, scaling 0.18215 — `self.vae` of the pipeline,
utils/adaptive_mask_inpainting.py:680 encode, :1086/:1112 decode) on the B200 kernels."""
import torch

from . import nn
from .nn import Act, F16, F32

SD_VAE = dict(in_channels=3, latent_channels=4, block_out_channels=(128, 256, 512, 512), layers_per_block=2, groups=32,
              scaling_factor=0.18215)


class VAE:
    def __init__(self, state_dict, cfg=SD_VAE, device="cuda"):
        self.cfg, self.dev = dict(cfg), torch.device(device)
        if self.dev.type != "cuda":
            raise RuntimeError("coma_b200 VAE runs on CUDA (sm_100a) only — there is no CPU fallback")
        self.p = {}
        for k, v in state_dict.items():
            if k.endswith(".weight") and v.dim() == 4 and v.shape[-1] == 3:
                self.p[k] = nn.prep_conv3x3(v, self.dev)
            elif k.endswith(".weight") and v.dim() in (2, 4):
                self.p[k] = nn.prep_linear(v, self.dev)
            else:
                self.p[k] = nn.prep_vec(v, self.dev)

    def _resnet(self, x: Act, name):
        p, G = self.p, self.cfg["groups"]
        gn1 = nn.gn_affine(x, p[name + ".norm1.weight"], p[name + ".norm1.bias"], G, 1e-6)
        h = nn.conv3x3(x, p[name + ".conv1.weight"], p[name + ".conv1.bias"], gn=gn1, act=1, stats=True)
        gn2 = nn.gn_affine(h, p[name + ".norm2.weight"], p[name + ".norm2.bias"], G, 1e-6)
        if name + ".conv_shortcut.weight" in p:
            sc = nn.gemm(x.t, p[name + ".conv_shortcut.weight"], p[name + ".conv_shortcut.bias"])
        else:
            sc = x.t
        return nn.conv3x3(h, p[name + ".conv2.weight"], p[name + ".conv2.bias"], gn=gn2, act=1, residual=sc, stats=True)

    def _attn(self, x: Act, name):
        p = self.p
        s, sh = nn.gn_affine(x, p[name + ".group_norm.weight"], p[name + ".group_norm.bias"], self.cfg["groups"], 1e-6)
        hn = nn.affine_act(x, s, sh, 0).t
        S = x.H * x.W
        C = x.C
        q = nn.gemm(hn, p[name + ".to_q.weight"], p[name + ".to_q.bias"])
        k = nn.gemm(hn, p[name + ".to_k.weight"], p[name + ".to_k.bias"])
        v = nn.gemm(hn, p[name + ".to_v.weight"], p[name + ".to_v.bias"])
        scores = torch.empty((x.B, S, S), dtype=F16, device=self.dev)
        nn.gemm_batched(q, C, 0, S * C, k, C, 0, S * C, scores, S, 0, S * S, S, S, C, 1, x.B, alpha=C ** -0.5)
        vt = torch.empty((x.B, C, S), dtype=F16, device=self.dev)
        with torch.cuda.device(self.dev):
            nn.call("coma_softmax_rows_f16", scores.data_ptr(), x.B * S, S, S, nn._stream())
            nn.call("coma_transpose_heads_f16", v.data_ptr(), x.B, S, 1, C, C, vt.data_ptr(), S, nn._stream())
        o = torch.empty((x.B * S, C), dtype=F16, device=self.dev)
        nn.gemm_batched(scores, S, 0, S * S, vt, S, 0, C * S, o, C, 0, S * C, S, C, S, 1, x.B)
        out = nn.gemm(o, p[name + ".to_out.0.weight"], p[name + ".to_out.0.bias"], residual=x.t)
        return Act(out, x.B, x.H, x.W)

    def decode(self, z: Act, out_dtype=F32):
        """z: Act [B,h,w,4] fp16 (latents / scaling_factor) -> image Act [B,8h,8w,3] in [-1,1]."""
        p, ch, nl = self.p, self.cfg["block_out_channels"], self.cfg["layers_per_block"]
        h0 = nn.new_act(z.B, z.H, z.W, self.cfg["latent_channels"], self.dev)
        nn.gemm(z.t, p["post_quant_conv.weight"], p["post_quant_conv.bias"], out=h0.t)
        h = nn.conv3x3(h0, p["decoder.conv_in.weight"], p["decoder.conv_in.bias"])
        h = self._resnet(h, "decoder.mid_block.resnets.0")
        h = self._attn(h, "decoder.mid_block.attentions.0")
        h = self._resnet(h, "decoder.mid_block.resnets.1")
        for i in range(len(ch)):
            for j in range(nl + 1):
                h = self._resnet(h, f"decoder.up_blocks.{i}.resnets.{j}")
            if i < len(ch) - 1:
                h = nn.conv3x3(h, p[f"decoder.up_blocks.{i}.upsamplers.0.conv.weight"], p[f"decoder.up_blocks.{i}.upsamplers.0.conv.bias"], up=True, stats=True)
        gn = nn.gn_affine(h, p["decoder.conv_norm_out.weight"], p["decoder.conv_norm_out.bias"], self.cfg["groups"], 1e-6)
        return nn.conv3x3(h, p["decoder.conv_out.weight"], p["decoder.conv_out.bias"], gn=gn, act=1, out_dtype=out_dtype)

    def encode_moments(self, img: Act):
        """img: Act [B,H,W,3] fp16 in [-1,1] -> (mean, logvar) fp32 tensors [B*h*w, 4] (logvar clamped to [-30, 20])."""
        p, ch, nl = self.p, self.cfg["block_out_channels"], self.cfg["layers_per_block"]
        h = nn.conv3x3(img, p["encoder.conv_in.weight"], p["encoder.conv_in.bias"])
        for i in range(len(ch)):
            for j in range(nl):
                h = self._resnet(h, f"encoder.down_blocks.{i}.resnets.{j}")
            if i < len(ch) - 1:
                h = nn.conv3x3(h, p[f"encoder.down_blocks.{i}.downsamplers.0.conv.weight"],
                               p[f"encoder.down_blocks.{i}.downsamplers.0.conv.bias"], stride=2, pad=0, stats=True)
        h = self._resnet(h, "encoder.mid_block.resnets.0")
        h = self._attn(h, "encoder.mid_block.attentions.0")
        h = self._resnet(h, "encoder.mid_block.resnets.1")
        gn = nn.gn_affine(h, p["encoder.conv_norm_out.weight"], p["encoder.conv_norm_out.bias"], self.cfg["groups"], 1e-6)
        h = nn.conv3x3(h, p["encoder.conv_out.weight"], p["encoder.conv_out.bias"], gn=gn, act=1)
        m = nn.gemm(h.t, p["quant_conv.weight"], p["quant_conv.bias"], out_dtype=F32)
        lat = self.cfg["latent_channels"]
        return m[:, :lat].contiguous(), m[:, lat:].clamp(-30.0, 20.0).contiguous()
