"""Multi-GPU plumbing: one process per GPU (torchrun), torch.distributed (NCCL over NVLink; gloo on CPU for tests).

ComA contact (K2/K3)   : shard SAMPLES  -> rank r aggregates samples r::world; one all-reduce(SUM) of the accumulators.
ComA occupancy (K4/K5c): shard HUMAN VERTICES -> every rank sees all samples, owns rows [h0,h1); one all-reduce(MAX) of
                         the [Sg^3] field (a SUM of the 88 GB grid is never needed).
Inpainting work items  : contiguous slices of the sorted work list, exactly the reference's rule
                         (src/generation/inpaint.py:272-278) so that outputs land in the same files.
"""
import os


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init_process_group(backend=None):
    import torch
    import torch.distributed as dist
    rank, world, local = env_rank_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def sample_shard(n_samples, rank, world):
    """Indices of the samples rank `rank` aggregates (strided: r, r+world, ...)."""
    return list(range(rank, n_samples, world))


def human_slice(human_res, rank, world):
    """Contiguous block [h0, h1) of human vertices owned by `rank` (balanced to within one vertex)."""
    base, rem = divmod(human_res, world)
    h0 = rank * base + min(rank, rem)
    return h0, h0 + base + (1 if rank < rem else 0)


def work_item_slice(n_items, parallel_idx, parallel_num):
    """The reference's work-list slice: sub = n//N + 1; items[idx*sub:(idx+1)*sub] (src/generation/inpaint.py:272-278)."""
    sub = n_items // parallel_num + 1
    return parallel_idx * sub, min((parallel_idx + 1) * sub, n_items)


def is_distributed(group=None):
    import torch.distributed as dist
    return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


def all_slices(human_res, world):
    return [human_slice(human_res, r, world) for r in range(world)]


def gather_rows(local, human_res, group=None, dst=None):
    """Concatenate the row blocks of a human-vertex-sharded tensor ([h1-h0, ...] on each rank, `human_slice` rule).
    dst=None: every rank gets the full [human_res, ...] tensor (all-gather of blocks padded to the largest block — meant for
    the small read-out maps). dst=r: only rank r assembles it, block by block through point-to-point transfers into a HOST
    tensor (the 31 GB orientation grids are exported this way without a second device copy); other ranks return None."""
    import torch
    import torch.distributed as dist
    if not is_distributed(group):
        return local
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    slices = all_slices(human_res, world)
    if dst is None:
        rows = max(h1 - h0 for h0, h1 in slices)
        pad = torch.zeros((rows,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        pad[: local.shape[0]] = local
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad, group=group)
        return torch.cat([parts[r][: slices[r][1] - slices[r][0]] for r in range(world)])
    if rank != dst:
        if local.shape[0]:
            dist.send(local.contiguous(), dst=dst, group=group)
        return None
    full = torch.empty((human_res,) + tuple(local.shape[1:]), dtype=local.dtype, device="cpu")
    for r, (h0, h1) in enumerate(slices):
        if h1 == h0:
            continue
        if r == rank:
            full[h0:h1] = local.cpu()
        else:
            buf = torch.empty((h1 - h0,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
            dist.recv(buf, src=r, group=group)
            full[h0:h1] = buf.cpu()
    return full


def all_reduce_max_nan(t, group=None):
    """MAX all-reduce that propagates NaN like torch.max (NCCL / gloo leave NaN ordering undefined): NaN travels as +inf."""
    import torch
    import torch.distributed as dist
    if not is_distributed(group):
        return t
    nan = torch.isnan(t)
    carried = torch.where(nan, torch.full_like(t, float("inf")), t)
    dist.all_reduce(carried, op=dist.ReduceOp.MAX, group=group)
    return torch.where(torch.isinf(carried) & (carried > 0), torch.full_like(carried, float("nan")), carried)
