"""Seeded synthetic 3D-HOI samples: no dataset is needed to exercise the ComA path.

Object = O points on an ellipsoid (semi-axes 0.15, 0.10, 0.22 m) with outward unit normals, FIXED across
samples (the reference asserts this invariant for occupancy, utils/coma_occupancy.py:277-284).
Human per sample = H points on a capsule (r 0.15, half-height 0.75) with outward normals, rotated about z and
translated so that a small fraction of (human, object) pairs falls under the contact threshold.
Everything is float64, exactly what the real pipeline hands to `register_sample_to_cache`.
"""
import numpy as np


def _unit(v):
    return v / np.linalg.norm(v, axis=-1, keepdims=True)


def make_object(O, rng):
    axes = np.array([0.15, 0.10, 0.22])
    d = _unit(rng.standard_normal((O, 3)))
    verts = d * axes
    normals = _unit(d / axes)
    return verts, normals


def make_human(H, rng):
    r, half = 0.15, 0.75
    z = rng.uniform(-half - r, half + r, H)
    phi = rng.uniform(0, 2 * np.pi, H)
    zc = np.clip(z, -half, half)            # nearest point on the capsule axis
    dz = z - zc                             # != 0 only on the caps
    rad = np.sqrt(np.maximum(r * r - dz * dz, 0.0))
    local = np.stack([rad * np.cos(phi), rad * np.sin(phi), z], -1)
    nrm = _unit(np.stack([rad * np.cos(phi), rad * np.sin(phi), dz], -1) + 1e-12)
    return local, nrm


def make_samples(S, H, O, seed=42):
    """-> list of S dicts {human_verts, human_normals, obj_verts, obj_normals} (float64 numpy)."""
    rng = np.random.default_rng(seed)
    ov, on = make_object(O, rng)
    hv0, hn0 = make_human(H, rng)
    out = []
    for _ in range(S):
        a = rng.uniform(0, 2 * np.pi)
        R = np.array([[np.cos(a), -np.sin(a), 0.0], [np.sin(a), np.cos(a), 0.0], [0.0, 0.0, 1.0]])
        t = np.array([rng.uniform(-0.4, 0.4), rng.uniform(-0.4, 0.4), rng.uniform(-0.2, 0.2)])
        jitter = rng.standard_normal((H, 3)) * 0.004
        out.append(dict(human_verts=(hv0 + jitter) @ R.T + t, human_normals=hn0 @ R.T,
                        obj_verts=ov.copy(), obj_normals=on.copy()))
    return out


def make_sample_arrays(S, H, O, seed=42, dtype=np.float32):
    """Batched form for the bench: hv,hn [S,H,3]; ov,on [S,O,3] (object replicated per sample)."""
    ss = make_samples(S, H, O, seed)
    st = lambda k: np.stack([s[k] for s in ss]).astype(dtype)
    return st("human_verts"), st("human_normals"), st("obj_verts"), st("obj_normals")


def make_adversarial_samples(H, O, thres, seed=7):
    """Samples that sit on the decision boundaries the reference's integer outputs depend on:
    pair distances planted at fp32(thres)·(1 ± k·2⁻²⁴), antipodal / degenerate normals, duplicated vertices."""
    rng = np.random.default_rng(seed)
    s = make_samples(2, H, O, seed)
    t32 = np.float32(thres)
    for k, smp in enumerate(s):
        hv, ov = smp["human_verts"], smp["obj_verts"]
        n = min(H, O, 16)
        for i in range(n):
            d = _unit(rng.standard_normal(3))
            scale = float(t32) * (1.0 + (i - n // 2) * 2.0 ** -24)
            hv[i] = ov[i] + d * scale
        hn, on = smp["human_normals"], smp["obj_normals"]
        on[0] = [0.0, 0.0, -1.0]           # exactly antipodal to the principle vector -> reflect branch
        on[1 % O] = [0.0, 1e-12, -1.0]
        hn[0] = [0.0, 0.0, -1.0]
        hn[1 % H] = [0.0, 0.0, 1.0]
        hn[2 % H] = [1.0, 0.0, 0.0]
        hv[3 % H] = hv[4 % H]              # duplicate vertex
    # the object must be identical across samples
    s[1]["obj_verts"] = s[0]["obj_verts"].copy()
    s[1]["obj_normals"] = s[0]["obj_normals"].copy()
    return s
