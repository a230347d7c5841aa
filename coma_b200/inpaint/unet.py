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

        # LayerNorm folded into the projection that consumes it (norm1 -> q|k|v, norm2 -> cross-attention q, norm3 -> feed-forward): the
        # weights absorb gamma, the bias absorbs beta, and the GEMM epilogue applies the per-row (rstd, -rstd * mean) — no normalised
        # tensor between LayerNorm and projection (48 passes per forward). Built from the fp32 checkpoint values.
        self.ln_fold = {}
        for k in [k for k in sd if k.endswith(".transformer_blocks.0.norm1.weight") and nn.FUSED_LN]:   # opt-in (see nn.FUSED_LN): no second copy of the weights otherwise
            t = k[: -len(".norm1.weight")]
            f32 = lambda name: sd[name].float()
            w1 = torch.cat([f32(t + ".attn1.to_q.weight"), f32(t + ".attn1.to_k.weight"), f32(t + ".attn1.to_v.weight")], 0)
            for tag, w, b, nrm in (("attn1", w1, None, "norm1"), ("attn2", f32(t + ".attn2.to_q.weight"), None, "norm2"),
                                   ("ff", f32(t + ".ff.net.0.proj.weight"), f32(t + ".ff.net.0.proj.bias"), "norm3")):
                wf, bf = nn.fold_layernorm(w, b, f32(f"{t}.{nrm}.weight"), f32(f"{t}.{nrm}.bias"))
                if tag == "ff":
                    fused = nn.prep_geglu(wf, bf, self.dev)
                    if fused is None:
                        continue
                    wp, bp = fused
                else:
                    wp, bp = nn.prep_linear(wf, self.dev), nn.prep_vec(bf, self.dev)
                self.ln_fold[f"{t}.{tag}"] = (wp, nn.ln_c1(wp), bp)

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
        t = name + ".transformer_blocks.0"
        fold = self.ln_fold if nn.FUSED_LN else {}
        C, M = x.C, x.B * S
        # LayerNorm statistics without a statistics kernel: every GEMM that writes the residual stream leaves per-panel (sum, sumsq) of its
        # rounded output rows (ln_out), the folded projection that follows forms (rstd, -rstd * mean) from them in its epilogue
        part = (lambda key: nn.ln_partials(M, C, x.t.device) if key in fold and C % 32 == 0 else None)
        p1, p2, p3 = part(t + ".attn1"), part(t + ".attn2") if ctx_kv is not None else None, part(t + ".ff")
        h = nn.gemm(nn.affine_act(x, s, sh, 0).t, p[name + ".proj_in.weight"], p[name + ".proj_in.bias"], ln_out=p1)
        if p1 is not None:           # norm1 folded into the q|k|v projection
            w, c1, b = fold[t + ".attn1"]
            h = nn.attention(h, h, B, S, S, None, None, None, p[t + ".attn1.to_out.0.weight"], p[t + ".attn1.to_out.0.bias"], heads, h,
                             wqkv=w, ln=(p1, c1, b), ln_out=p2)
        else:
            n1 = nn.layernorm(h, p[t + ".norm1.weight"], p[t + ".norm1.bias"])
            h = nn.attention(n1, n1, B, S, S, None, None, None, p[t + ".attn1.to_out.0.weight"], p[t + ".attn1.to_out.0.bias"], heads, h,
                             wqkv=p[t + ".attn1.to_qkv.weight"], ln_out=p2)
        if p2 is not None:
            w, c1, b = fold[t + ".attn2"]
            h = nn.attention(h, ctx, B, S, L, w, None, None, p[t + ".attn2.to_out.0.weight"], p[t + ".attn2.to_out.0.bias"], heads, h,
                             wkv=p[t + ".attn2.to_kv.weight"], kv=ctx_kv[t + ".attn2"], ln=(p2, c1, b), ln_out=p3)
        else:
            n2 = nn.layernorm(h, p[t + ".norm2.weight"], p[t + ".norm2.bias"])
            h = nn.attention(n2, ctx, B, S, L, p[t + ".attn2.to_q.weight"], None, None, p[t + ".attn2.to_out.0.weight"],
                             p[t + ".attn2.to_out.0.bias"], heads, h, wkv=p[t + ".attn2.to_kv.weight"],
                             kv=None if ctx_kv is None else ctx_kv[t + ".attn2"], ln_out=p3)
        if p3 is not None:           # norm3 folded into the GEGLU projection
            w, c1, b = fold[t + ".ff"]
            ff = nn.gemm_geglu(h, w, b, ln=(p3, c1))
        else:
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
