"""Host->device staging of sample batches: pinned double buffers, copies on a side stream, fp64->fp32 rounding on the
host exactly as the reference's `to_np_torch_recursive` does (utils/misc.py:47-54)."""
import ctypes

import numpy as np
import torch


_from_buffer, _addressof = ctypes.c_char.from_buffer, ctypes.addressof


def _addresses(arrays):
    """Data pointers of a list of numpy arrays. `ndarray.ctypes.data` costs ~1.8 us per array (it builds a helper object); going through
    the buffer protocol costs ~0.6 us — per sample and per array that is a third of the host time of a staged chunk. Read-only arrays
    do not export a writable buffer: they take the slow road. An object that is the same as its predecessor (one shared object mesh for
    all samples, the normal case for `sub`) is not looked up again."""
    out, prev, prev_addr = [], None, 0
    for a in arrays:
        if a is not prev:
            try:
                prev_addr = _addressof(_from_buffer(a))
            except (TypeError, ValueError):
                prev_addr = a.ctypes.data
            prev = a
        out.append(prev_addr)
    return out


def stage_rows_f64(arrays, out, row0=0, sub=None, equal_to=None):
    """out[i] = fp32(arrays[i][row0:row0+rows] - sub[i]) for a list of C-contiguous float64 [*, 3] arrays, through the library's host
    helper `coma_host_stage_rows_f64_f32` (one call for the whole list instead of one numpy call per sample). `out`: writable fp32
    numpy view [>= n, rows, 3] (the pinned staging buffer). Returns the first index whose `sub` row differs bitwise from `equal_to`
    (-1 if none / not requested), or None if some array does not qualify (the caller falls back to numpy)."""
    from . import _lib
    n = len(arrays)
    rows = out.shape[1]
    f64 = np.float64
    for a in arrays:
        if a.dtype != f64 or not a.flags.c_contiguous or a.ndim != 2 or a.shape[1] != 3 or a.shape[0] < row0 + rows:
            return None
    src = (ctypes.c_void_p * n)(*_addresses(arrays))
    subp = None
    if sub is not None:
        for a in sub:
            if a.dtype != f64 or not a.flags.c_contiguous or a.size < 3:
                return None
        subp = (ctypes.c_void_p * n)(*_addresses(sub))
    eq = None if equal_to is None else np.ascontiguousarray(equal_to, dtype=f64)
    mism = ctypes.c_int64(-1)
    _lib.call("coma_host_stage_rows_f64_f32", src, subp, n, row0, rows, out.ctypes.data, None if eq is None else eq.ctypes.data,
              ctypes.addressof(mism))
    return int(mism.value)


def rows_equal_f64(arrays, ref):
    """True iff every C-contiguous float64 array in `arrays` starts with the values of `ref` (one library call); None if an array
    does not qualify."""
    from . import _lib
    ref = np.ascontiguousarray(ref, dtype=np.float64).reshape(-1)
    for a in arrays:
        if a.dtype != np.float64 or not a.flags.c_contiguous or a.size < ref.size:
            return None
    n = len(arrays)
    ptrs = (ctypes.c_void_p * n)(*_addresses(arrays))
    mism = ctypes.c_int64(-1)
    _lib.call("coma_host_rows_equal_f64", ptrs, n, ref.ctypes.data, ref.size, ctypes.addressof(mism))
    return mism.value < 0


class BatchStager:
    """Streams a list of per-sample numpy arrays (each [rows_k, 3]) to the GPU in chunks of `chunk` samples.

    `specs` = {name: rows}. For every chunk it yields {name: cuda fp32 tensor [n, rows, 3]} valid until the next-but-one
    iteration (two buffers in flight). The fp32 rounding happens while filling the pinned buffer (np.copyto), so the
    bytes crossing PCIe/NVLink-C2C are the fp32 ones."""

    def __init__(self, specs, chunk, device):
        self.specs, self.chunk, self.device = dict(specs), int(chunk), torch.device(device)
        self.is_cuda = self.device.type == "cuda"
        self.h2d_bytes = 0
        self.pinned, self.dev = [], []
        for _ in range(2):
            self.pinned.append({k: torch.empty((self.chunk, r, 3), dtype=torch.float32, pin_memory=self.is_cuda)
                                for k, r in self.specs.items()})
            self.dev.append({k: torch.empty((self.chunk, r, 3), dtype=torch.float32, device=self.device)
                             for k, r in self.specs.items()})
        if self.is_cuda:
            self.copy_stream = torch.cuda.Stream(device=self.device)
            self.ready = [torch.cuda.Event() for _ in range(2)]
            self.free = [torch.cuda.Event() for _ in range(2)]
            self.used_once = [False, False]

    def batches(self, getters, S, n_chunks=None, full=False):
        """getters = {name: callable(sample_index) -> array-like [rows,3]}; yields (n, {name: tensor[:n]}).
        `n_chunks` forces that many iterations (trailing ones with n = 0: a rank that ran out of samples still has to
        take part in the exchange collectives); `full` yields the whole [chunk, rows, 3] buffers (fixed-size all-gather)."""
        compute = torch.cuda.current_stream(self.device) if self.is_cuda else None
        total = (S + self.chunk - 1) // self.chunk if n_chunks is None else int(n_chunks)
        for i in range(total):
            s0 = i * self.chunk
            b, n = i & 1, max(0, min(self.chunk, S - s0))
            if self.is_cuda and self.used_once[b]:
                self.free[b].synchronize()       # kernels that read dev[b] are done -> both buffers reusable
            for k in self.specs:
                view = self.pinned[b][k].numpy()
                g = getters[k]
                if hasattr(g, "fill_many"):       # getter that fills rows [0, n) of the pinned buffer in one call (host helper in the library)
                    if n:
                        g.fill_many(view, s0, n)
                elif hasattr(g, "fill"):          # getter that writes straight into the pinned row (no temporary)
                    for j in range(n):
                        g.fill(view[j], s0 + j)
                else:
                    for j in range(n):
                        np.copyto(view[j], g(s0 + j), casting="same_kind")
            if self.is_cuda:
                with torch.cuda.stream(self.copy_stream):
                    for k in self.specs:
                        self.dev[b][k][:n].copy_(self.pinned[b][k][:n], non_blocking=True)
                        self.h2d_bytes += n * self.specs[k] * 12
                    self.ready[b].record(self.copy_stream)
                compute.wait_event(self.ready[b])
            else:
                for k in self.specs:
                    self.dev[b][k][:n].copy_(self.pinned[b][k][:n])
            yield n, {k: (self.dev[b][k] if full else self.dev[b][k][:n]) for k in self.specs}
            if self.is_cuda:
                self.free[b].record(compute)
                self.used_once[b] = True


def exchanged_batches(stager, getters, n_local, group=None):
    """Sample exchange for the ROW-sharded accumulators (ComA / ComA_Occupancy with `human_slice`): every rank stages the
    samples IT loaded (full rows), the staged fp32 chunks are all-gathered over NVLink (a few hundred MB for a whole job)
    and every rank receives all samples. Yields {name: [n_total, rows, 3]} with the ranks' samples concatenated in rank
    order; the caller slices its rows. All ranks must iterate in lock-step (the chunk count is agreed by a MAX all-reduce)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    dev = stager.device
    chunk = stager.chunk
    nmax = torch.tensor([(n_local + chunk - 1) // chunk], dtype=torch.int64, device=dev)
    dist.all_reduce(nmax, op=dist.ReduceOp.MAX, group=group)
    for n, bufs in stager.batches(getters, n_local, n_chunks=int(nmax.item()), full=True):
        counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(counts, torch.tensor([n], dtype=torch.int64, device=dev), group=group)
        counts = [int(c.item()) for c in counts]
        out = {}
        for k, buf in bufs.items():
            parts = [torch.empty_like(buf) for _ in range(world)]
            dist.all_gather(parts, buf, group=group)
            out[k] = torch.cat([parts[r][:counts[r]] for r in range(world) if counts[r] > 0]) if sum(counts) else buf[:0]
        yield sum(counts), out
