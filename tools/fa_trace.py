"""Timeline of the fused attention kernel (timing build: COMA_NVCC_EXTRA=-DFA_TRACE COMA_B200_LIB=build/libcoma_fa_trace.so python -m coma_b200.build):
per key step, where CTA (0,0,0)'s softmax warp and MMA thread spend their cycles.
    COMA_B200_LIB=/root/repo/build/libcoma_fa_trace.so python tools/fa_trace.py"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coma_b200 import _lib  # noqa: E402
from coma_b200._lib import _stream, call  # noqa: E402

dev = torch.device("cuda:0")
B, heads, S, d = 8, 8, 4096, int(os.environ.get("D", 40))
C = heads * d
g = torch.Generator(device=dev).manual_seed(0)
q = torch.randn((B, S, C), device=dev, generator=g).half()
k = torch.randn((B, S, C), device=dev, generator=g).half()
vt = torch.randn((B, heads, d, S), device=dev, generator=g).half()
out = torch.empty((B, S, C), dtype=torch.float16, device=dev)
for _ in range(3):
    call("coma_attention_fwd_f16", q.data_ptr(), k.data_ptr(), vt.data_ptr(), B, heads, S, S, d, C, C, S, float(d ** -0.5), out.data_ptr(), C, _stream())
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    call("coma_attention_fwd_f16", q.data_ptr(), k.data_ptr(), vt.data_ptr(), B, heads, S, S, d, C, C, S, float(d ** -0.5), out.data_ptr(), C, _stream())
b.record()
torch.cuda.synchronize()
print(f"{a.elapsed_time(b) / 5:.3f} ms per call")
buf = (ctypes.c_longlong * (128 * 16))()
lib = _lib.load()
rc = lib.coma_attention_trace(buf)
t = np.array(buf, dtype=np.int64).reshape(128, 16)
n = S // 64
t = t[:n]
base = t[0, 0]
sm = t[:, :7] - base
mm = t[:, 8:12] - base
names = ["wait s_full", "pass1 (ld+max)", "rescale chk", "wait p_empty", "pass2 (ld+exp+st)", "fence+arrives"]
d_sm = np.diff(sm, axis=1)              # [n, 6]
step = np.diff(sm[:, 0])
print(f"softmax warp: mean step {step[4:].mean():.0f} clk (steps 4..{n})")
for i, nm in enumerate(names):
    print(f"  {nm:20s} mean {d_sm[4:, i].mean():7.0f}  max {d_sm[4:, i].max():6d}")
print(f"  loop tail (6 -> next 0) mean {(sm[1:, 0] - sm[:-1, 6])[4:].mean():7.0f}")
mnames = ["issue_qk(j+1) incl. waits", "wait v_full + p_full", "issue PV + commits"]
d_mm = np.diff(mm, axis=1)
print(f"MMA thread: mean step {np.diff(mm[:, 0])[4:].mean():.0f} clk")
for i, nm in enumerate(mnames):
    print(f"  {nm:28s} mean {d_mm[4:, i].mean():7.0f}  max {d_mm[4:, i].max():6d}")
print("first 6 steps, softmax timestamps:\n", sm[:6])
print("first 6 steps, MMA timestamps:\n", mm[:6])
