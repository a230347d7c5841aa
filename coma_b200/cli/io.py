"""File-format helpers of the ComA scripts that the reference delegates to open3d / matplotlib (absent here):
vertex normals, PLY point-cloud output, the 'jet' colour map, seeding."""
import os
import random

import numpy as np


def seed_everything(seed):
    """utils/reproducibility.py:11-20 semantics (python / numpy / torch seeds)."""
    import torch
    random.seed(seed)
    os.environ["PYTHONHASHSEED"] = str(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


def vertex_normals(verts, faces):
    """open3d TriangleMesh.compute_vertex_normals(normalized=True): sum of the (area-weighted, un-normalised) face
    normals (v1-v0)x(v2-v0) over incident faces, then normalised; degenerate sums become (0,0,1). float64."""
    v = np.asarray(verts, dtype=np.float64)
    f = np.asarray(faces, dtype=np.int64)
    fn = np.cross(v[f[:, 1]] - v[f[:, 0]], v[f[:, 2]] - v[f[:, 0]])
    vn = np.zeros_like(v)
    for k in range(3):
        np.add.at(vn, f[:, k], fn)
    n = np.linalg.norm(vn, axis=-1, keepdims=True)
    out = vn / np.where(n > 0, n, 1.0)
    out[(n[:, 0] == 0) | ~np.isfinite(out).all(-1)] = (0.0, 0.0, 1.0)
    return out


def jet_rgb(x):
    """matplotlib's 'jet' (piecewise-linear segment data) evaluated at x in [0,1] -> [n,3] float64."""
    x = np.clip(np.asarray(x, dtype=np.float64), 0.0, 1.0)
    r = np.interp(x, [0.0, 0.35, 0.66, 0.89, 1.0], [0.0, 0.0, 1.0, 1.0, 0.5])
    g = np.interp(x, [0.0, 0.125, 0.375, 0.64, 0.91, 1.0], [0.0, 0.0, 1.0, 1.0, 0.0, 0.0])
    b = np.interp(x, [0.0, 0.11, 0.34, 0.65, 1.0], [0.5, 1.0, 1.0, 0.0, 0.0])
    return np.stack([r, g, b], -1)


def write_point_cloud_ply(path, points, normals=None, colors=None):
    """Binary little-endian PLY with open3d's property names/order: x y z [nx ny nz] [red green blue] (double / uchar)."""
    pts = np.asarray(points, dtype=np.float64)
    n = len(pts)
    fields = [("x", "<f8"), ("y", "<f8"), ("z", "<f8")]
    if normals is not None:
        fields += [("nx", "<f8"), ("ny", "<f8"), ("nz", "<f8")]
    if colors is not None:
        fields += [("red", "u1"), ("green", "u1"), ("blue", "u1")]
    rec = np.zeros(n, dtype=fields)
    rec["x"], rec["y"], rec["z"] = pts.T
    if normals is not None:
        rec["nx"], rec["ny"], rec["nz"] = np.asarray(normals, dtype=np.float64).T
    if colors is not None:
        c = np.clip(np.asarray(colors, dtype=np.float64), 0, 1)
        rec["red"], rec["green"], rec["blue"] = (np.floor(c * 255.0 + 0.0).astype(np.uint8)).T  # open3d: uint8(c*255)
    ptype = {"<f8": "double", "u1": "uchar"}
    header = "ply\nformat binary_little_endian 1.0\ncomment Created by coma_b200\nelement vertex %d\n" % n
    header += "".join(f"property {ptype[t]} {name}\n" for name, t in fields) + "end_header\n"
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "wb") as fh:
        fh.write(header.encode("ascii"))
        fh.write(rec.tobytes())


def read_point_cloud_ply(path):
    """Reader for the files written above (tests / round trips)."""
    with open(path, "rb") as fh:
        fields, n = [], 0
        while True:
            line = fh.readline().decode("ascii").strip()
            if line.startswith("element vertex"):
                n = int(line.split()[-1])
            elif line.startswith("property"):
                _, t, name = line.split()
                fields.append((name, {"double": "<f8", "uchar": "u1", "float": "<f4"}[t]))
            elif line == "end_header":
                break
        return np.frombuffer(fh.read(), dtype=fields, count=n)
