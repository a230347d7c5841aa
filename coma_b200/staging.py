"""Host->device staging of sample batches: pinned double buffers, copies on a side stream, fp64->fp32 rounding on the
host exactly as the reference's `to_np_torch_recursive` does (utils/misc.py:47-54)."""
import numpy as np
import torch


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

    def batches(self, getters, S):
        """getters = {name: callable(sample_index) -> array-like [rows,3]}; yields (n, {name: tensor[:n]})."""
        compute = torch.cuda.current_stream(self.device) if self.is_cuda else None
        for i, s0 in enumerate(range(0, S, self.chunk)):
            b, n = i & 1, min(self.chunk, S - s0)
            if self.is_cuda and self.used_once[b]:
                self.free[b].synchronize()       # kernels that read dev[b] are done -> both buffers reusable
            for k in self.specs:
                view = self.pinned[b][k].numpy()
                for j in range(n):
                    np.copyto(view[j], getters[k](s0 + j), casting="same_kind")
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
            yield n, {k: self.dev[b][k][:n] for k in self.specs}
            if self.is_cuda:
                self.free[b].record(compute)
                self.used_once[b] = True
