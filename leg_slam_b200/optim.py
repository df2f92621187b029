"""Fused multi-tensor Adam over the Gaussian parameter tensors, mirroring how the reference
sets up `torch::optim::Adam` (src/gaussian_model.cpp:483-518): one single-tensor parameter
group per Gaussian attribute, each with its own learning rate, betas (0.9, 0.999),
eps 1e-15, no weight decay / amsgrad.  `step()` is ONE kernel launch over all groups
(`lgs_adam_multi`, include/lgs.h) instead of ~8 ATen kernels per tensor.

State layout is torch's (`state[p] = {"step", "exp_avg", "exp_avg_sq"}`) so the reference's
densify/prune optimizer-state surgery (src/gaussian_model.cpp:577-727) keeps working on it.
"""
import ctypes

import torch

from . import _lib
from ._lib import check


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=0.0, betas=(0.9, 0.999), eps=1e-15):
        if eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1:
            raise ValueError("invalid Adam hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))

    @torch.no_grad()
    def step(self):
        L = _lib.lib()
        # tensors sharing (betas, eps, step) go into one launch
        buckets = {}
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda:
                    raise _lib.LgsError("FusedAdam has no CPU path")
                if p.dtype != torch.float32 or not p.is_contiguous():
                    raise TypeError("FusedAdam expects contiguous float32 parameters")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                key = (group["betas"][0], group["betas"][1], group["eps"], st["step"], p.device)
                buckets.setdefault(key, []).append((p, g, st["exp_avg"], st["exp_avg_sq"], float(group["lr"])))
        for (b1, b2, eps, step, dev), items in buckets.items():
            for i in range(0, len(items), 16):
                chunk = items[i:i + 16]
                n = len(chunk)
                VP = ctypes.c_void_p * n
                ps = VP(*[t[0].data_ptr() for t in chunk])
                gs = VP(*[t[1].data_ptr() for t in chunk])
                ms = VP(*[t[2].data_ptr() for t in chunk])
                vs = VP(*[t[3].data_ptr() for t in chunk])
                ns = (ctypes.c_int64 * n)(*[t[0].numel() for t in chunk])
                lrs = (ctypes.c_double * n)(*[t[4] for t in chunk])
                with torch.cuda.device(dev):
                    check(L.lgs_adam_multi(n, ps, gs, ms, vs, ns, lrs, float(b1), float(b2), float(eps), int(step),
                                           torch.cuda.current_stream(dev).cuda_stream), "lgs_adam_multi")
