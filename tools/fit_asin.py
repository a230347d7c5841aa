import numpy as np
from numpy.polynomial import chebyshev as C, polynomial as P
# R(z) = (asin(sqrt z)/sqrt z - 1)/z on z in [0, 0.25]
def R(z):
    z = np.asarray(z, dtype=np.float64)
    out = np.empty_like(z)
    small = z < 1e-8
    s = np.sqrt(z[~small])
    out[~small] = (np.arcsin(s)/s - 1)/z[~small]
    out[small] = 1/6 + 3/40*z[small]
    return out
for deg in (3,4,5):
    # chebyshev nodes on [0,0.25]
    k = np.arange(200)
    x = np.cos(np.pi*(k+0.5)/200)
    z = 0.125*(x+1)
    c = C.chebfit(x, R(z), deg)
    # convert to monomial in z: x = 8z - 1
    px = C.cheb2poly(c)
    # substitute x = 8z-1
    pz = np.zeros(1)
    for i, a in enumerate(px):
        pz = P.polyadd(pz, a*P.polypow([-1.0, 8.0], i))
    # evaluate acos in float32 emulation
    cs = np.linspace(-1, 1, 2000001).astype(np.float32)
    f32 = np.float32
    a = np.abs(cs)
    big = a > f32(0.5)
    zz = np.where(big, (f32(1)-a)*f32(0.5), a*a).astype(np.float32)
    s = np.where(big, np.sqrt(zz), a).astype(np.float32)
    r = np.zeros_like(zz) + f32(pz[-1])
    for coef in pz[-2::-1]:
        r = (r*zz + f32(coef)).astype(np.float32)
    asn = (s + s*(zz*r)).astype(np.float32)
    A = np.where(big, np.where(cs>0, f32(0), f32(np.pi)), f32(np.pi/2)).astype(np.float32)
    B = np.where(big, np.where(cs>0, f32(2), f32(-2)), np.where(cs>=0, f32(-1), f32(1))).astype(np.float32)
    g = (A + B*asn).astype(np.float32)
    ref = np.arccos(cs.astype(np.float64))
    err = np.abs(g.astype(np.float64)-ref)
    g2err = np.abs(g.astype(np.float64)**2 - ref**2)
    print(deg, "max abs err acos", err.max(), "max g2 err", g2err.max(), "g2 err for c>-0.7", g2err[cs>-0.7].max())
    print("  coeffs (low->high):", [float(np.float32(v)) for v in pz])
