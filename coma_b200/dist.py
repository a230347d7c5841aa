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
