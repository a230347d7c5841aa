"""Which arithmetic does the reference's torch code perform on CUDA for the bit-exact quantities? (SURVEY §7 "which oracle
is the reference": production runs device='cuda', src/coma/extract_coma.py:329.) Prints, for the exact expressions of
utils/coma.py:284-287 and utils/coma_occupancy.py:292-293, which association of the 3-term sum torch's CUDA reduction uses,
whether products/sums are contracted to FMA, and in which type `d < python_float` compares. Run on the GPU box."""
import json
import sys

import numpy as np
import torch

dev = torch.device("cuda:0")
out = {}
g = torch.Generator(device=dev).manual_seed(0)


def assoc_report(d, dtype, shape_desc):
    """d [..., 3] differences -> which scalar formula reproduces sum(square(d), -1) bit for bit."""
    x, y, z = d[..., 0], d[..., 1], d[..., 2]
    s = torch.sum(torch.square(d), dim=-1)
    xx, yy, zz = x * x, y * y, z * z
    cands = {"(xx+yy)+zz": (xx + yy) + zz, "xx+(yy+zz)": xx + (yy + zz), "(xx+zz)+yy": (xx + zz) + yy,
             "fma(z,z,fma(y,y,xx))": torch.addcmul(torch.addcmul(xx, y, y), z, z),
             "fma(x,x,fma(y,y,zz))": torch.addcmul(torch.addcmul(zz, y, y), x, x)}
    return {k: int((v != s).sum().item()) for k, v in cands.items()} | {"n": int(s.numel()), "shape": shape_desc}


for dtype in (torch.float32, torch.float64):
    for (H, O) in ((1000, 180), (10475, 1500), (25, 4), (3, 1), (1003, 181)):
        if dtype == torch.float64 and H * O > 4e6:
            continue
        a = torch.randn((H, 3), device=dev, generator=g, dtype=dtype)
        b = torch.randn((O, 3), device=dev, generator=g, dtype=dtype) * 0.3
        d = a[:, None, :] - b[None, :, :]
        out[f"sum_lastdim_{str(dtype)[6:]}_{H}x{O}"] = assoc_report(d, dtype, [H, O, 3])

# the occupancy expression: (grid[None,3,S,S,S] - hv[H,3,None,None,None]).square().sum(dim=1) in fp64
for (H, S) in ((64, 30), (7, 12), (10, 64)):
    grid = torch.randn((3, S, S, S), device=dev, generator=g, dtype=torch.float64)
    hv = torch.randn((H, 3), device=dev, generator=g, dtype=torch.float32)
    diff = grid[None] - hv[:, :, None, None, None]
    s = diff.square().sum(dim=1)
    x, y, z = diff[:, 0], diff[:, 1], diff[:, 2]
    xx, yy, zz = x * x, y * y, z * z
    out[f"occ_sum_dim1_H{H}_S{S}"] = {"(xx+yy)+zz": int(((xx + yy) + zz != s).sum()), "xx+(yy+zz)": int((xx + (yy + zz) != s).sum()),
                                      "(xx+zz)+yy": int(((xx + zz) + yy != s).sum()), "dtype": str(diff.dtype), "n": int(s.numel())}

# comparison semantics: fp32 tensor < python float
t32 = np.float32(0.03)
vals = torch.tensor([np.nextafter(t32, np.float32(0)), t32, np.nextafter(t32, np.float32(1))], device=dev)
out["lt_python_float_0.03"] = (vals < 0.03).tolist()
out["lt_cpu"] = (vals.cpu() < 0.03).tolist()
v64 = torch.tensor([0.05625 - 1e-17, 0.05625, 0.05625 + 1e-17], device=dev, dtype=torch.float64)
out["lt_f64"] = (v64 < 0.05625).tolist()
# sqrt correctly rounded?
q = torch.rand(1 << 20, device=dev, generator=g) * 0.01
out["sqrt_f32_vs_f64_mismatch"] = int((torch.sqrt(q) != torch.sqrt(q.double()).float()).sum())
# exp(-d/size)
dd = torch.rand(1 << 20, device=dev, generator=g)
e = torch.exp(-dd / 0.15)
out["exp_rel_err_max"] = float(((e.double() - torch.exp(-(dd / 0.15).double())) / torch.exp(-(dd / 0.15).double())).abs().max())
out["div_f32_exact"] = int(((dd / 0.15) != (dd.double() / float(np.float32(0.15))).float()).sum())   # division by fp32(0.15)?
out["div_f64_scalar"] = int(((dd / 0.15) != (dd.double() / 0.15).float()).sum())                     # or by the double scalar?
# CPU for comparison (association on CPU)
a = torch.randn((1000, 3)); b = torch.randn((180, 3))
d = a[:, None, :] - b[None, :, :]
s = torch.sum(torch.square(d), dim=-1)
x, y, z = d[..., 0], d[..., 1], d[..., 2]
out["cpu_sum"] = {"(xx+yy)+zz": int(((x * x + y * y) + z * z != s).sum()), "xx+(yy+zz)": int((x * x + (y * y + z * z) != s).sum())}
# CPU vs CUDA on the same data
out["cpu_vs_cuda_sum_mismatch"] = int((torch.sum(torch.square(d.to(dev)), dim=-1).cpu() != s).sum())
print(json.dumps(out, indent=1))
