"""Worst error of the K3 grids against the fp64 oracle, as a fraction of the parity bar (rtol 1e-4 + S * 2^-31), per sigma.
Used to A/B builds of the kernel (COMA_B200_LIB=... python tools/k3_err_probe.py): margin, not just pass / fail."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from coma_b200 import ops, synth          # noqa: E402
from oracle import oracle                 # noqa: E402

out = {"lib": os.environ.get("COMA_B200_LIB", "default")}
dev = torch.device("cuda:0")
H, O, N, S = 96, 80, 250, 64
for sigma in (0.1, 0.2, 0.25):
    samples = synth.make_samples(S, H, O, seed=int(sigma * 100)) + synth.make_adversarial_samples(H, O, 0.05, seed=3)
    hn = np.stack([s["human_normals"] for s in samples]).astype(np.float32)
    on = np.stack([s["obj_normals"] for s in samples]).astype(np.float32)
    grid = oracle.fibonacci_sphere(N)
    for order in ("cuda",):
        rPH, rPO = oracle.orient_accumulate(hn, on, grid, sigma, 1e-10, sum_order=order)
        PH, PO = torch.zeros((H, O, N), device=dev), torch.zeros((H, O, N), device=dev)
        ops.orient_accumulate(torch.from_numpy(hn).to(dev), torch.from_numpy(on).to(dev), torch.from_numpy(grid).to(dev), sigma, 1e-10,
                              [0, 0, 1], [0, 1, 0], PH, PO, bin_perm=ops.bin_patches(grid, dev), drop_bits=32, sum_order=order)
        worst = 0.0
        for mine, ref in ((PH.cpu().numpy(), rPH), (PO.cpu().numpy(), rPO)):
            ok = np.isfinite(ref)
            tol = 1e-4 * np.abs(ref[ok]) + len(samples) * 2.0 ** -31
            worst = max(worst, float(np.max(np.abs(mine[ok].astype(np.float64) - ref[ok]) / tol)))
        out[f"sigma{sigma}"] = round(worst, 4)
print(json.dumps(out))
