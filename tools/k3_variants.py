"""A/B timing of the K3 kernel variants on one GPU + cross-variant agreement.
    python tools/k3_variants.py [dense cone cone_noperm ...]      S=<samples> SIGMA=<sigma> as env
`dense` = orient_accumulate_kernel_x2 (all 2 x 250 bins), `cone` = cone-limited kernel with compact patches (ComA's default),
`cone_noperm` = cone-limited with bins grouped in index order, `cone24` = drop_bits 24."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coma_b200 import _lib, ops, synth  # noqa: E402
from coma_b200.misc import get_uniform_points_on_sphere  # noqa: E402

dev = torch.device("cuda:0")
H, O, N, S = int(os.environ.get("H", 10475)), 1500, 250, int(os.environ.get("S", 64))
SIGMA = float(os.environ.get("SIGMA", 0.25))
hv, hn, ov, on = (torch.from_numpy(a).to(dev) for a in synth.make_sample_arrays(S, H, O, seed=1))
gh = np.stack(get_uniform_points_on_sphere(N), -1)
grid = torch.tensor(gh, device=dev)
perm = ops.bin_patches(gh, dev)
PH, PO = torch.zeros((H, O, N), device=dev), torch.zeros((H, O, N), device=dev)
ref = None
KW = {"dense": dict(drop_bits=0), "cone": dict(drop_bits=32, bin_perm=perm), "cone_noperm": dict(drop_bits=32),
      "cone24": dict(drop_bits=24, bin_perm=perm), "cone40": dict(drop_bits=40, bin_perm=perm)}
for variant in sys.argv[1:] or ["dense", "cone", "cone_noperm"]:
    ts = []
    for it in range(4):
        PH.zero_(); PO.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ops.orient_accumulate(hn, on, grid, SIGMA, 1e-10, [0, 0, 1], [0, 1, 0], PH, PO, **KW[variant])
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = float(np.median(ts[1:]))
    evals = 2.0 * N * S * H * O
    clk = 148 * 4 * 1.965e9 / (evals / 32 / (ms * 1e-3))
    sub = torch.cat([PH[:64], PO[:64]]).clone()
    msg = ""
    if ref is None:
        ref = sub
    else:
        err = (sub - ref).abs()
        rel = (err / ref.clamp_min(1e-30))[ref > 1e-6 * ref.amax(-1, keepdim=True)].max().item()
        msg = f"  vs first variant on rows 0..63: max abs diff {err.max().item():.2e} (S*2^-32 = {S * 2.0 ** -32:.2e}), max rel diff on bins > 1e-6 pair max {rel:.2e}"
    print(f"{variant:12s} [{_lib.last_kernel()}]: {ms:8.2f} ms  {evals / ms / 1e9:8.1f} G algorithmic bin-evals/s  {S * H * O / ms / 1e6:7.2f} G pair-samples/s  "
          f"{clk:5.2f} clk/warp-eval/SMSP{msg}", flush=True)
