"""Sample ingest for the ComA extraction: per-vertex normals of the fitted SMPL-X meshes on the GPU.

Reference: utils/coma.py:665-686 — every human sample builds an open3d TriangleMesh from (`verts`, `faces`), calls
`compute_vertex_normals()` and passes the result through `normalize_vectors_np(., eps)`. The topology is the same for every
sample, so the incident-face list is built once here and the kernel (K6, csrc/normals.cu) gathers per vertex in the
accumulation order of the numpy restatement (oracle/oracle.py:vertex_normals) — results are bit-identical to it.
There is no CPU fallback.
"""
import numpy as np
import torch

from ._lib import _ptr, _stream, call


def build_corner_list(faces, num_verts):
    """CSR of the faces incident to each vertex, ordered by (vertex, corner index, face index) — the order in which
    `np.add.at(vn, f[:, k], fn)` for k = 0, 1, 2 adds the face normals, which is what makes K6 bit-identical to the numpy
    restatement. Returns (faces int32 [F,3], corner_off int32 [V+1], corner_face int32 [3F])."""
    f = np.ascontiguousarray(np.asarray(faces).reshape(-1, 3).astype(np.int64))
    assert f.size and f.min() >= 0 and f.max() < num_verts, "face indices out of range"
    F = int(f.shape[0])
    corner_vertex = f.T.reshape(-1)                                   # [3F]: corner k of face j at k*F + j
    order = np.argsort(corner_vertex, kind="stable")
    counts = np.bincount(corner_vertex, minlength=num_verts)
    off = np.zeros(num_verts + 1, dtype=np.int32)
    np.cumsum(counts, out=off[1:])
    return f.astype(np.int32), off, (order % F).astype(np.int32)


class MeshNormals:
    """Vertex normals for meshes sharing `faces` [F,3]: `normals = MeshNormals(faces, V)(verts)` with verts [V,3] or [S,V,3]
    (numpy or CUDA tensor, any float dtype; computed in fp64 like the reference) -> same leading shape, fp64 CUDA tensor."""

    def __init__(self, faces, num_verts, device="cuda"):
        self.dev = torch.device(device)
        if self.dev.type != "cuda":
            raise RuntimeError("coma_b200.ingest runs on CUDA only (there is no CPU fallback)")
        f, off, cf = build_corner_list(faces, num_verts)
        self.V, self.F = int(num_verts), int(f.shape[0])
        self.corner_face = torch.from_numpy(cf).to(self.dev)
        self.corner_off = torch.from_numpy(off).to(self.dev)
        self.faces = torch.from_numpy(f).to(self.dev)

    def __call__(self, verts, eps=-1.0):
        v = torch.as_tensor(np.asarray(verts) if not torch.is_tensor(verts) else verts).to(device=self.dev, dtype=torch.float64)
        lead = v.shape[:-2]
        v = v.reshape(-1, self.V, 3).contiguous()
        out = torch.empty_like(v)
        with torch.cuda.device(self.dev):
            call("coma_vertex_normals_f64", _ptr(v), v.shape[0], self.V, _ptr(self.faces), self.F, _ptr(self.corner_off),
                 _ptr(self.corner_face), float(eps), _ptr(out), _stream())
        return out.reshape(*lead, self.V, 3)


_CACHE = {}


def mesh_normals_for(faces, num_verts, device="cuda"):
    """Topology cache (keyed by the face array's bytes and the device): SMPL-X's corner list is built once per process."""
    f = np.ascontiguousarray(np.asarray(faces).reshape(-1, 3).astype(np.int64))
    dev = torch.device(device)
    if dev.type == "cuda" and dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    key = (hash(f.tobytes()), int(num_verts), str(dev))
    mn = _CACHE.get(key)
    if mn is None:
        mn = _CACHE[key] = MeshNormals(f, int(num_verts), dev)
    return mn


def vertex_normals(verts, faces, eps=-1.0, device="cuda"):
    """Convenience wrapper around `mesh_normals_for`: numpy in ([V,3] or [S,V,3]), numpy fp64 out."""
    return mesh_normals_for(faces, int(np.asarray(verts).shape[-2]), device)(verts, eps).cpu().numpy()
