"""A/B of the occupancy kernels (K4 scatter, K5c read-out) on one B200: round-1 forms against the round-2 forms.

The switches (COMA_B200_OCC_PATH=v1, COMA_B200_K5C=dense) are read once per process, so every variant runs in its own
subprocess:   python tools/occ_ab.py [--rows 1310] [--samples 4096] [--sg 128] [--out gpurun_out/occ_ab.json]
With --one it times the variant selected by the environment and prints one JSON line (also what `ncu` wraps).
"""
import argparse
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def one(rows, samples, sg, iters):
    import numpy as np
    import torch
    from coma_b200 import ops, synth
    from coma_b200.coma_occupancy import load_voxelgrid
    dev = torch.device("cuda:0")
    chunks, obj0 = [], None
    for c0 in range(0, samples, 512):
        hv, _, ov, _ = synth.make_sample_arrays(min(512, samples - c0), rows, 4, seed=900 + c0 // 512, dtype=np.float64)
        obj0 = ov[0, 0] if obj0 is None else obj0
        chunks.append(torch.from_numpy((hv - obj0[None, None]).astype(np.float32)))
    hvc = torch.cat(chunks).to(dev)
    g, _, meta = load_voxelgrid(2.4, sg)
    centers = torch.from_numpy(np.ascontiguousarray(np.stack([g[0, :, 0, 0], g[1, 0, :, 0], g[2, 0, 0, :]]))).to(dev)
    thr = meta["voxel_size"] * 3.0
    grids = torch.zeros((rows, sg, sg, sg), device=dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    k4, k5 = [], []
    hits = field_sum = None
    for it in range(iters + 1):
        grids.zero_()
        torch.cuda.synchronize()
        ev[0].record()
        ops.occupancy_accumulate(hvc, centers, thr, grids)
        ev[1].record()
        if it == 0:
            hits = float(grids.sum(dtype=torch.float64).item())
            nz = float((grids != 0).sum().item()) / grids.numel()
            ev[1].record()
        field = ops.occupancy_readout(grids, None)
        ev[2].record()
        torch.cuda.synchronize()
        if it:   # first pass = warm-up
            k4.append(ev[0].elapsed_time(ev[1]))
            k5.append(ev[1].elapsed_time(ev[2]))
        field_sum = float(field.nan_to_num(0).sum(dtype=torch.float64).item())
    k4m, k5m = sorted(k4)[len(k4) // 2], sorted(k5)[len(k5) // 2]
    V = sg ** 3
    return dict(rows=rows, samples=samples, sg=sg, k4_ms=k4m, k5c_ms=k5m, hits=hits, nonzero_fraction=nz,
                vertex_samples_per_s=rows * samples / k4m * 1e3, hits_per_s=hits / k4m * 1e3,
                k5c_dense_equiv_gbs=12.0 * rows * V / k5m / 1e6, k5c_two_read_gbs=8.0 * rows * V / k5m / 1e6, field_checksum=field_sum,
                k4_all_ms=[round(x, 4) for x in k4], k5c_all_ms=[round(x, 4) for x in k5],
                occ_path=os.environ.get("COMA_B200_OCC_PATH", "default"), k5c=os.environ.get("COMA_B200_K5C", "default"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1310)
    ap.add_argument("--samples", type=int, default=4096)
    ap.add_argument("--sg", type=int, default=128)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--one", action="store_true")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    if a.one:
        print(json.dumps(one(a.rows, a.samples, a.sg, a.iters)))
        return
    res = []
    for occ_path, k5c in (("v1", "dense"), ("default", "default")):
        env = dict(os.environ)
        env.pop("COMA_B200_OCC_PATH", None)
        env.pop("COMA_B200_K5C", None)
        if occ_path != "default":
            env["COMA_B200_OCC_PATH"] = occ_path
        if k5c != "default":
            env["COMA_B200_K5C"] = k5c
        for rows, samples, sg in ((a.rows, a.samples, a.sg), (10475, 256, 30)):
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", "--rows", str(rows), "--samples", str(samples),
                                "--sg", str(sg), "--iters", str(a.iters)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
            line = [l for l in r.stdout.splitlines() if l.startswith("{")]
            res.append(json.loads(line[-1]) if line else dict(error=r.stderr[-2000:], occ_path=occ_path, k5c=k5c, sg=sg))
            print(res[-1], flush=True)
    if a.out:
        with open(a.out, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
