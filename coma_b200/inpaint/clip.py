"""CLIP text encoder on the B200 kernels: the `self.text_encoder(text_input_ids)` of `_encode_prompt`
(utils/adaptive_mask_inpainting.py:405-554, calls at :478 and :534) — transformers' CLIPTextModel of the SD-1.5 checkpoint
(ViT-L/14 text tower: 12 layers, width 768, 12 heads, MLP 3072, quick-GELU, causal mask, 77 tokens) — returning
`last_hidden_state` [B, 77, 768], which is what the pipeline feeds the UNet's cross-attention.

Every matrix product runs on the tcgen05 GEMM (coma_gemm_f16_ex: QKV as one GEMM, the quick-GELU folded into fc1's epilogue, the
residual adds folded into out_proj's / fc2's), LayerNorm on coma_layernorm_f16, the 77 x 77 causal attention as two batched GEMMs
around coma_softmax_rows_causal_f16 (77 tokens: a fused kernel would have nothing to hide). fp16 storage, fp32 accumulation, like the
reference's `torch_dtype=torch.float16` pipeline. Tokenisation is string processing on the host (transformers' CLIPTokenizer when a
tokenizer directory exists); the token / position embedding lookup is an index_select. There is no CPU fallback.
"""
import torch

from . import nn
from .._lib import _stream, call

F16, F32 = torch.float16, torch.float32

CLIP_L_TEXT_CFG = dict(vocab_size=49408, hidden_size=768, intermediate_size=3072, num_hidden_layers=12, num_attention_heads=12,
                       max_position_embeddings=77, layer_norm_eps=1e-5, hidden_act="quick_gelu")


class CLIPTextEncoder:
    """state_dict: transformers CLIPTextModel keys (`text_model.embeddings.token_embedding.weight`, `text_model.encoder.layers.N.*`,
    `text_model.final_layer_norm.*`). `encoder(input_ids [B, 77] int64) -> [B, 77, hidden] f16`."""

    def __init__(self, state_dict, cfg=None, device="cuda"):
        self.cfg = dict(CLIP_L_TEXT_CFG if cfg is None else cfg)
        self.dev = torch.device(device)
        if self.dev.type != "cuda":
            raise RuntimeError("coma_b200.inpaint.clip runs on CUDA (sm_100a) only — there is no CPU fallback")
        if self.cfg.get("hidden_act", "quick_gelu") != "quick_gelu":
            raise NotImplementedError("only the quick_gelu text tower of SD-1.x checkpoints is implemented")
        sd = {k[len("text_model."):] if k.startswith("text_model.") else k: v for k, v in state_dict.items()}
        C, H = self.cfg["hidden_size"], self.cfg["num_attention_heads"]
        assert C % H == 0 and (C // H) % 8 == 0
        self.tok = sd["embeddings.token_embedding.weight"].to(self.dev, F16)
        self.pos = sd["embeddings.position_embedding.weight"].to(self.dev, F16)
        self.layers = []
        for i in range(self.cfg["num_hidden_layers"]):
            p = f"encoder.layers.{i}."
            wqkv = torch.cat([sd[p + f"self_attn.{n}_proj.weight"] for n in "qkv"], 0)
            bqkv = torch.cat([sd[p + f"self_attn.{n}_proj.bias"] for n in "qkv"], 0)
            self.layers.append(dict(
                ln1=(nn.prep_vec(sd[p + "layer_norm1.weight"], self.dev), nn.prep_vec(sd[p + "layer_norm1.bias"], self.dev)),
                ln2=(nn.prep_vec(sd[p + "layer_norm2.weight"], self.dev), nn.prep_vec(sd[p + "layer_norm2.bias"], self.dev)),
                wqkv=nn.prep_linear(wqkv, self.dev), bqkv=nn.prep_vec(bqkv, self.dev),
                wo=nn.prep_linear(sd[p + "self_attn.out_proj.weight"], self.dev), bo=nn.prep_vec(sd[p + "self_attn.out_proj.bias"], self.dev),
                w1=nn.prep_linear(sd[p + "mlp.fc1.weight"], self.dev), b1=nn.prep_vec(sd[p + "mlp.fc1.bias"], self.dev),
                w2=nn.prep_linear(sd[p + "mlp.fc2.weight"], self.dev), b2=nn.prep_vec(sd[p + "mlp.fc2.bias"], self.dev)))
        self.lnf = (nn.prep_vec(sd["final_layer_norm.weight"], self.dev), nn.prep_vec(sd["final_layer_norm.bias"], self.dev))

    def _attention(self, x, L, B, S):
        """Causal multi-head self-attention over S tokens + out_proj + residual. x [B*S, C] f16 (the LayerNorm-ed input is computed here)."""
        C, heads = self.cfg["hidden_size"], self.cfg["num_attention_heads"]
        d = C // heads
        Sp = nn.rup(S)
        h = nn.layernorm(x, *L["ln1"], eps=self.cfg["layer_norm_eps"])
        qkv = nn.gemm(h, L["wqkv"], L["bqkv"])                                   # [B*S, 3C]
        q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
        ld = qkv.stride(0)
        scores = torch.empty((B, heads, S, Sp), dtype=F16, device=self.dev)
        nn.gemm_batched(q, ld, d, S * ld, k, ld, d, S * ld, scores, Sp, S * Sp, heads * S * Sp, S, S, d, heads, B, alpha=d ** -0.5)
        vt = torch.empty((B, heads, d, Sp), dtype=F16, device=self.dev)
        with torch.cuda.device(self.dev):
            call("coma_softmax_rows_causal_f16", scores.data_ptr(), B * heads * S, S, S, Sp, _stream())
            call("coma_transpose_heads_f16", v.data_ptr(), B, S, heads, d, ld, vt.data_ptr(), Sp, _stream())
        o = torch.empty((B * S, C), dtype=F16, device=self.dev)
        nn.gemm_batched(scores, Sp, S * Sp, heads * S * Sp, vt, Sp, d * Sp, heads * d * Sp, o, C, d, S * C, S, d, Sp, heads, B)
        return nn.gemm(o, L["wo"], L["bo"], residual=x)

    @torch.no_grad()
    def __call__(self, input_ids):
        ids = torch.as_tensor(input_ids, device=self.dev).long()
        if ids.dim() == 1:
            ids = ids[None]
        B, S = ids.shape
        assert S <= self.cfg["max_position_embeddings"]
        x = (self.tok.index_select(0, ids.reshape(-1)).view(B, S, -1) + self.pos[:S][None]).reshape(B * S, -1).contiguous()
        for L in self.layers:
            x = self._attention(x, L, B, S)
            h = nn.layernorm(x, *L["ln2"], eps=self.cfg["layer_norm_eps"])
            h = nn.gemm(h, L["w1"], L["b1"], act=2)                              # quick-GELU in the epilogue
            x = nn.gemm(h, L["w2"], L["b2"], residual=x)
        return nn.layernorm(x, *self.lnf, eps=self.cfg["layer_norm_eps"]).view(B, S, -1)


def load_text_encoder(model_dir, device="cuda"):
    """<model_dir>/text_encoder/{model.safetensors | pytorch_model.bin} + config.json (diffusers checkpoint layout) -> CLIPTextEncoder."""
    import json
    import os
    d = os.path.join(model_dir, "text_encoder")
    st = os.path.join(d, "model.safetensors")
    if os.path.exists(st):
        from safetensors.torch import load_file
        sd = load_file(st)
    else:
        sd = torch.load(os.path.join(d, "pytorch_model.bin"), map_location="cpu")
    cfg = dict(CLIP_L_TEXT_CFG)
    cj = os.path.join(d, "config.json")
    if os.path.exists(cj):
        with open(cj) as fh:
            cfg.update({k: v for k, v in json.load(fh).items() if k in cfg})
    return CLIPTextEncoder(sd, cfg, device)


def make_embedder(model_dir, device="cuda"):
    """text -> [77, hidden] f16 prompt embeddings: CLIPTokenizer on the host (padding to 77, truncation, as :461-476) + the encoder above."""
    import os
    from transformers import CLIPTokenizer
    tok = CLIPTokenizer.from_pretrained(os.path.join(model_dir, "tokenizer"))
    enc = load_text_encoder(model_dir, device)

    def embed(text):
        ids = tok(text, padding="max_length", max_length=tok.model_max_length, truncation=True, return_tensors="pt").input_ids
        return enc(ids)[0]
    return embed
