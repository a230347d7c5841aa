"""Fit acos(a) ~= sqrt(1-a) * P(a) on a in [0,1] (Abramowitz-Stegun 4.4.45 form) and report the fp32-evaluated error of
acos and of acos^2 (the quantity exp(-acos^2/sigma^2) is sensitive to). Used for the K3 score function."""
import numpy as np
from numpy.polynomial import chebyshev as C, polynomial as P

f32 = np.float32


def target(a):
    a = np.asarray(a, dtype=np.float64)
    w = 1.0 - a
    out = np.empty_like(a)
    small = w < 1e-10
    out[~small] = np.arccos(a[~small]) / np.sqrt(w[~small])
    out[small] = np.sqrt(2.0) * (1 + w[small] / 12)
    return out


def fit(deg, iters=30):
    # Chebyshev interpolation then a few Remez-like reweighting steps (good enough: near-minimax)
    k = np.arange(400)
    x = np.cos(np.pi * (k + 0.5) / 400)
    a = 0.5 * (x + 1)
    c = C.chebfit(x, target(a), deg)
    px = C.cheb2poly(c)
    pa = np.zeros(1)
    for i, co in enumerate(px):
        pa = P.polyadd(pa, co * P.polypow([-1.0, 2.0], i))
    return pa


for deg in (6, 7, 8):
    pa = fit(deg)
    cs = np.linspace(-1, 1, 4000001).astype(f32)
    a = np.minimum(np.abs(cs), f32(1))
    w = (f32(1) - a).astype(f32)
    s = np.sqrt(w).astype(f32)
    r = np.zeros_like(a) + f32(pa[-1])
    for co in pa[-2::-1]:
        r = (r * a + f32(co)).astype(f32)
    ga = (s * r).astype(f32)
    g = np.where(cs >= 0, ga, (f32(np.pi) - ga).astype(f32)).astype(f32)
    ref = np.arccos(np.clip(cs.astype(np.float64), -1, 1))
    e = np.abs(g.astype(np.float64) - ref)
    e2 = np.abs(g.astype(np.float64) ** 2 - ref ** 2)
    rel2 = e2 / np.maximum(ref ** 2, 1e-30)
    print(deg, "max|acos err| %.3e  max|g2 err| %.3e  max rel g2 err (g>1e-3) %.3e" % (e.max(), e2.max(), rel2[ref > 1e-3].max()))
    print("   coeffs low->high:", ", ".join("%.9ef" % f32(v) for v in pa))
