"""Times the fused attention kernel alone on the UNet's 64x64-level self-attention shape (B2=8, 8 heads, 4096 tokens, d=40)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coma_b200.inpaint import nn
dev = torch.device("cuda:0"); B, S, heads, d = 8, int(sys.argv[1]) if len(sys.argv) > 1 else 4096, 8, int(os.environ.get("D", 40)); C = heads * d
L = int(sys.argv[2]) if len(sys.argv) > 2 else S
g = torch.Generator(device=dev).manual_seed(0)
q = torch.randn((B * S, C), device=dev, generator=g).half()
k, v = (torch.randn((B * L, C), device=dev, generator=g).half() for _ in range(2))
Lp = (L + 7) // 8 * 8
vt = torch.zeros((B, heads, d, Lp), dtype=torch.float16, device=dev); o = torch.empty((B * S, C), dtype=torch.float16, device=dev)
nn.call("coma_transpose_heads_f16", v.data_ptr(), B, L, heads, d, C, vt.data_ptr(), Lp, nn._stream())
f = lambda: nn.call("coma_attention_fwd_f16", q.data_ptr(), k.data_ptr(), vt.data_ptr(), B, heads, S, L, d, C, C, Lp, float(d ** -0.5), o.data_ptr(), C, nn._stream())
for _ in range(3): f()
torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20): f()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 20
print(f"{os.environ.get('COMA_B200_LIB', 'default'):44s} S={S} L={L} {ms:.4f} ms  {4 * B * heads * S * L * d / ms / 1e9:.0f} TFLOP/s")
