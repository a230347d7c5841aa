"""Differentiable nearest-neighbour distances between point sets on the B200 (K7, csrc/nearest_dist.cu) and the two reference
functions built on them:

  chamfer_distance(A, B)               src/application/optimize.py:155-165  (contact loss of the SMPL-X pose optimiser, autograd)
  minimum_distance(A, B, num_vertices) src/generation/optimize_depth.py:29-44

The reference materialises torch.cdist's [NA, NB] matrix (and its autograd graph) to read one minimum per row; here a kernel
returns (dist, argmin) directly and the backward touches one b per a. torch only carries the autograd plumbing and the O(N)
means / top-k on the resulting vectors. There is no CPU fallback.
"""
import torch

from ._lib import _ptr, _stream, call


class _NearestDistance(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        if not (a.is_cuda and b.is_cuda):
            raise RuntimeError("coma_b200.geometry runs on CUDA (sm_100a) only — there is no CPU fallback")
        a32, b32 = a.detach().to(torch.float32).contiguous(), b.detach().to(torch.float32).contiguous()
        NA, NB = a32.shape[0], b32.shape[0]
        dist = torch.empty(NA, dtype=torch.float32, device=a.device)
        idx = torch.empty(NA, dtype=torch.int32, device=a.device)
        scratch = torch.empty(NA, dtype=torch.int64, device=a.device)
        with torch.cuda.device(a.device):
            call("coma_nearest_distance_f32", _ptr(a32), NA, _ptr(b32), NB, _ptr(dist), _ptr(idx), _ptr(scratch), _stream())
        ctx.save_for_backward(a32, b32, idx, dist)
        ctx.dtypes = (a.dtype, b.dtype)
        ctx.mark_non_differentiable(idx)
        return dist.to(a.dtype), idx

    @staticmethod
    def backward(ctx, grad_dist, _grad_idx):
        a32, b32, idx, dist = ctx.saved_tensors
        need_a, need_b = ctx.needs_input_grad
        ga = torch.empty_like(a32) if need_a else None
        gb = torch.zeros_like(b32) if need_b else None
        if need_a or need_b:
            g = grad_dist.to(torch.float32).contiguous()
            with torch.cuda.device(a32.device):
                call("coma_nearest_distance_backward_f32", _ptr(a32), a32.shape[0], _ptr(b32), b32.shape[0], _ptr(idx), _ptr(dist), _ptr(g),
                     _ptr(ga), _ptr(gb), _stream())
        return (None if ga is None else ga.to(ctx.dtypes[0])), (None if gb is None else gb.to(ctx.dtypes[1]))


def nearest_distance(a, b):
    """a [NA,3], b [NB,3] CUDA tensors -> (dist [NA] = min_j ||a_i - b_j||, idx [NA] int32 first arg-min). Differentiable in a, b."""
    assert a.dim() == 2 and b.dim() == 2 and a.shape[1] == 3 and b.shape[1] == 3, "point sets must be [N,3]"
    return _NearestDistance.apply(a, b)


def chamfer_distance(point_cloud_A, point_cloud_B):
    """src/application/optimize.py:155-165: mean_i min_j ||a_i - b_j|| + mean_j min_i ||b_j - a_i|| (differentiable)."""
    d_ab, _ = nearest_distance(point_cloud_A, point_cloud_B)
    d_ba, _ = nearest_distance(point_cloud_B, point_cloud_A)
    return torch.mean(d_ab) + torch.mean(d_ba)


def minimum_distance(vertsA, vertsB, num_vertices=100):
    """src/generation/optimize_depth.py:29-44: mean of the `num_vertices` smallest nearest distances from A to B.
    vertsA / vertsB: [N,3] or [1,N,3] (the reference's cdist batch of one)."""
    A = vertsA.reshape(-1, 3).float()
    B = vertsB.reshape(-1, 3).float()
    d, _ = nearest_distance(A, B)
    d_sorted, _ = torch.sort(d)
    return torch.mean(d_sorted[:num_vertices])
