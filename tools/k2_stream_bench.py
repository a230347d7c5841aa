"""K2 in its HBM-bound streaming form (S = 1): achieved GB/s with accumulator sets rotating beyond L2."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coma_b200 import ops, synth  # noqa: E402

dev = torch.device("cuda:0")
print("variant", os.environ.get("COMA_B200_K2S", "0"))
H, O = 10475, 1500
hv, hn, ov, on = (torch.from_numpy(a).to(dev) for a in synth.make_sample_arrays(1, H, O, seed=1))
nsets = 8
cs = [torch.zeros((H, O), device=dev) for _ in range(nsets)]
ns = [torch.zeros((H, O), device=dev) for _ in range(nsets)]
for i in range(nsets):
    ops.pair_accumulate(hv, ov, 0.05, 0.15, cs[i], ns[i])
torch.cuda.synchronize()
ev = []
for i in range(4 * nsets):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    ops.pair_accumulate(hv, ov, 0.05, 0.15, cs[i % nsets], ns[i % nsets])
    b.record()
    ev.append((a, b))
torch.cuda.synchronize()
ms = float(np.median([a.elapsed_time(b) for a, b in ev]))
byts = 16.0 * H * O + 12.0 * (H + O)
print(f"K2 stream: {ms * 1e3:.1f} us  {byts / ms / 1e6:.0f} GB/s  ({byts / ms / 1e6 / 6452.5:.3f} of measured HBM peak)")
# a long back-to-back run (no per-launch event overhead): 64 launches over rotating sets
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for i in range(64):
    ops.pair_accumulate(hv, ov, 0.05, 0.15, cs[i % nsets], ns[i % nsets])
b.record()
torch.cuda.synchronize()
ms2 = a.elapsed_time(b) / 64
print(f"K2 stream back-to-back: {ms2 * 1e3:.1f} us/launch  {byts / ms2 / 1e6:.0f} GB/s  ({byts / ms2 / 1e6 / 6452.5:.3f})")
