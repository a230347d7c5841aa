"""A/B timing of the K3 kernel variants (env COMA_B200_K3) on one GPU + cross-variant agreement."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coma_b200 import ops, synth  # noqa: E402
from coma_b200.misc import get_uniform_points_on_sphere  # noqa: E402

dev = torch.device("cuda:0")
H, O, N, S = 10475, 1500, 250, int(os.environ.get("S", 64))
hv, hn, ov, on = (torch.from_numpy(a).to(dev) for a in synth.make_sample_arrays(S, H, O, seed=1))
grid = torch.tensor(np.stack(get_uniform_points_on_sphere(N), -1), device=dev)
PH, PO = torch.zeros((H, O, N), device=dev), torch.zeros((H, O, N), device=dev)
ref = None
for variant in sys.argv[1:] or ["v1", "x2", "x3", "x4"]:
    os.environ["COMA_B200_K3"] = variant
    ts = []
    for it in range(4):
        PH.zero_(); PO.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ops.orient_accumulate(hn, on, grid, 0.25, 1e-10, [0, 0, 1], [0, 1, 0], PH, PO)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = float(np.median(ts[1:]))
    evals = 2.0 * N * S * H * O
    clk = 148 * 4 * 1.965e9 / (evals / 32 / (ms * 1e-3))
    sub = PH[:64].clone()
    msg = ""
    if ref is None:
        ref = sub
    else:
        rel = ((sub - ref).abs() / ref.clamp_min(1e-30)).max().item()
        msg = f"  max rel diff vs first variant on rows 0..63: {rel:.2e}"
    print(f"{variant}: {ms:8.2f} ms  {evals / ms / 1e9:8.1f} G bin-evals/s  {S * H * O / ms / 1e6:7.2f} G pair-samples/s  {clk:5.2f} clk/warp-eval/SMSP{msg}")
