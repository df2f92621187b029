"""Data-parallel optimizer step fused with the gradient exchange over NVLink peer memory
(`lgs_dp_adam_shard`, csrc/dp_adam.cu): reduce-scatter + Adam-on-shard + all-gather in one launch per
rank, using NVSwitch multicast (multimem.ld_reduce / multimem.st) when the fabric offers it and plain
peer loads/stores otherwise.  Buffers are torch symmetric-memory allocations."""
import ctypes
import os

import torch
import torch.distributed as dist

from . import _lib
from ._lib import check


def symmetric_empty(numel, device):
    import torch.distributed._symmetric_memory as symm_mem
    return symm_mem.empty(numel, dtype=torch.float32, device=device)


class FusedDPAdam:
    """Adam over ONE flat parameter buffer whose tensors start at `seg_sizes` prefix offsets.  Parameters and
    gradients live in symmetric memory on every rank; Adam moments exist only for this rank's shard."""

    def __init__(self, param_flat, grad_flat, seg_sizes, lrs, group=None, betas=(0.9, 0.999), eps=1e-15):
        import torch.distributed._symmetric_memory as symm_mem
        group = group or dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        n = param_flat.numel()
        if n % 4 or any(s % 4 for s in seg_sizes) or grad_flat.numel() != n:
            raise ValueError("fused data-parallel Adam needs tensor sizes that are multiples of 4 floats")
        self.param_flat, self.grad_flat = param_flat, grad_flat
        self.hp = symm_mem.rendezvous(param_flat, group)
        self.hg = symm_mem.rendezvous(grad_flat, group)
        shard = ((n // 4 + self.world - 1) // self.world) * 4
        self.begin = min(self.rank * shard, n)
        self.end = min(self.begin + shard, n)
        dev = param_flat.device
        self.exp_avg = torch.zeros(max(self.end - self.begin, 4), dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros_like(self.exp_avg)
        starts = [0]
        for s in seg_sizes:
            starts.append(starts[-1] + int(s))
        self._seg = (ctypes.c_int64 * len(starts))(*starts)
        self.lrs = [float(x) for x in lrs]
        self.betas, self.eps, self.step_count = betas, eps, 0
        VP = ctypes.c_void_p * self.world
        self._gp = VP(*[int(p) for p in self.hg.buffer_ptrs])
        self._pp = VP(*[int(p) for p in self.hp.buffer_ptrs])
        # multimem.ld_reduce makes every GPU of the group (the requester included) send its copy to the switch: each GPU
        # ships its whole gradient buffer per step whatever G is.  That pays once the switch's reduction saves inbound
        # traffic (G >= 4); with two GPUs plain peer loads / stores move a third less over the links (measured on 2
        # B200s: exchange + Adam 0.35-0.40 ms with P2P, 0.61-0.69 ms with multimem; tools/prof_dp_step.py).
        env = os.environ.get("LGS_DP_MULTIMEM", "")
        use_mc = (self.world >= 4) if env == "" else (env == "1")
        if os.environ.get("LGS_DP_NO_MULTIMEM", "0") == "1":
            use_mc = False
        self.g_mc = int(getattr(self.hg, "multicast_ptr", 0) or 0) if use_mc else 0
        self.p_mc = int(getattr(self.hp, "multicast_ptr", 0) or 0) if use_mc else 0
        if not (self.g_mc and self.p_mc):
            self.g_mc = self.p_mc = 0
        self.uses_multicast = bool(self.g_mc)

    def step(self):
        L = _lib.lib()
        self.step_count += 1
        dev = self.param_flat.device
        n_seg = len(self.lrs)
        lr = (ctypes.c_double * n_seg)(*self.lrs)
        self.hg.barrier(channel=0)  # every rank's gradients are complete
        with torch.cuda.device(dev):
            check(L.lgs_dp_adam_shard(n_seg, self._seg, lr, self.world, self.rank, self._gp, self._pp,
                                      ctypes.c_void_p(self.g_mc or None), ctypes.c_void_p(self.p_mc or None),
                                      self.begin, self.end, self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
                                      float(self.betas[0]), float(self.betas[1]), float(self.eps), self.step_count,
                                      torch.cuda.current_stream(dev).cuda_stream), "lgs_dp_adam_shard")
        self.hp.barrier(channel=1)  # every shard's parameter writes have landed on every rank
