"""Data-parallel optimizer step fused with the gradient exchange over NVLink peer memory
(`lgs_dp_adam_shard`, csrc/dp_adam.cu): reduce-scatter + Adam-on-shard + all-gather in one launch per
rank, using NVSwitch multicast (multimem.ld_reduce / multimem.st) when the fabric offers it and plain
peer loads/stores otherwise.  Buffers are torch symmetric-memory allocations.

With `late_segment` (the language-feature tensor: 64 of the 123 floats per Gaussian) the exchange runs in phases.
Only the render kernels touch that tensor (reference forward.cu:261-392, backward.cu:399-612), so its exchange runs
on a side stream and the NEXT iteration's render forward is the first kernel that waits for it (`lgs_stream_hooks`):
the transfer runs underneath the next preprocess + binning.  By default it starts once the other tensors' exchange
has had the links to itself (measured best on 2 and 4 B200s); LGS_DP_EARLY_FRAC > 0 starts that fraction of it
already when this rank's render backward is done (the backward hook), underneath preprocess backward."""
import ctypes
import os

import torch
import torch.distributed as dist

from . import _lib
from ._lib import check


def symmetric_empty(numel, device, dtype=torch.float32):
    import torch.distributed._symmetric_memory as symm_mem
    return symm_mem.empty(numel, dtype=dtype, device=device)


_hook_owner = None  # the FusedDPAdam whose events are installed as this thread's lgs_stream_hooks


class _Range:
    """One contiguous piece [begin, end) of the flat index space, sharded over the ranks; Adam moments of this rank's
    shard only."""

    def __init__(self, begin, end, phase, world, rank, device):
        self.begin, self.end, self.phase = begin, end, phase
        shard = (((end - begin) // 4 + world - 1) // world) * 4
        self.shard = shard
        self.sb = min(begin + rank * shard, end)
        self.se = min(self.sb + shard, end)
        self.exp_avg = torch.zeros(max(self.se - self.sb, 4), dtype=torch.float32, device=device)
        self.exp_avg_sq = torch.zeros_like(self.exp_avg)


class FusedDPAdam:
    """Adam over ONE flat parameter buffer whose tensors start at `seg_sizes` prefix offsets.  Parameters and
    gradients live in symmetric memory on every rank; Adam moments exist only for this rank's shard."""

    def __init__(self, param_flat, grad_flat, seg_sizes, lrs, group=None, betas=(0.9, 0.999), eps=1e-15,
                 late_segment=None, rows=None):
        """rows = (P, row_lens): every tensor is [P, row_lens[t]] (one row per Gaussian).  Enables the SPARSE exchange: each
        rank reports which Gaussians it rendered this iteration (`mark_rows(radii)` after every local view) and the
        P2P branch loads a rank's copy of a gradient row only where that rank rendered the Gaussian (a culled Gaussian's
        row is exactly zero there).  Bit-identical to the dense exchange; 60 % fewer gradient bytes over NVLink with one
        view per rank at cfgB, which makes the P2P branch the faster one at every G measured, so with `rows` it is the
        default (LGS_DP_MULTIMEM=1 still selects the multicast branch, which cannot skip per rank: the switch adds all copies)."""
        import torch.distributed._symmetric_memory as symm_mem
        group = group or dist.group.WORLD
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        n = param_flat.numel()
        if n % 4 or any(s % 4 for s in seg_sizes) or grad_flat.numel() != n:
            raise ValueError("fused data-parallel Adam needs tensor sizes that are multiples of 4 floats")
        self.param_flat, self.grad_flat = param_flat, grad_flat
        self.hp = symm_mem.rendezvous(param_flat, group)
        self.hg = symm_mem.rendezvous(grad_flat, group)
        dev = param_flat.device
        self.dev = dev
        starts = [0]
        for s in seg_sizes:
            starts.append(starts[-1] + int(s))
        if starts[-1] != n:
            raise ValueError("segment sizes do not add up to the flat buffer")
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
        sparse_ok = rows is not None and os.environ.get("LGS_DP_SPARSE", "1") != "0"
        use_mc = (self.world >= 4 and not sparse_ok) if env == "" else (env == "1")
        if os.environ.get("LGS_DP_NO_MULTIMEM", "0") == "1":
            use_mc = False
        self.g_mc = int(getattr(self.hg, "multicast_ptr", 0) or 0) if use_mc else 0
        self.p_mc = int(getattr(self.hp, "multicast_ptr", 0) or 0) if use_mc else 0
        if not (self.g_mc and self.p_mc):
            self.g_mc = self.p_mc = 0
        self.uses_multicast = bool(self.g_mc)
        # sparse exchange state: local visibility bytes, the symmetric [world][P] table every rank publishes into, the bit mask
        self.sparse = bool(sparse_ok and not self.uses_multicast)
        self._rows_marked = False
        if self.sparse:
            L = _lib.lib()
            self.P = int(rows[0])
            self._row_len = (ctypes.c_int * len(rows[1]))(*[int(x) for x in rows[1]])
            if len(rows[1]) != len(seg_sizes) or any((self.P * int(rl) + 3) & ~3 != int(sz) for rl, sz in zip(rows[1], seg_sizes)):
                raise ValueError("rows = (P, row_lens) must describe the segments: segment size = P * row_len rounded up to 4")
            self.vis = torch.zeros((self.P + 3) & ~3, dtype=torch.uint8, device=param_flat.device)
            self.table = symmetric_empty(int(L.lgs_dp_rows_table_bytes(self.P, self.world)), param_flat.device, torch.uint8)
            self.table.fill_(1)
            self.ht = symm_mem.rendezvous(self.table, group)
            self._tp = (ctypes.c_void_p * self.world)(*[int(p) for p in self.ht.buffer_ptrs])
            self.row_mask = torch.zeros(self.P, dtype=torch.int16, device=param_flat.device)
        # two-phase exchange: phase 1 = the late segment on the side stream, phase 0 = everything else on the caller's
        self.overlap = late_segment is not None and os.environ.get("LGS_DP_OVERLAP", "1") != "0"
        if self.overlap:
            # phase 1: the part of the late segment that fits underneath preprocess backward (starts at the backward hook);
            # phase 2: the rest, launched once phase 0 (the caller's stream) has had the links to itself, so that it runs
            # underneath the next iteration's preprocess + binning instead of competing with phase 0
            a, b = starts[late_segment], starts[late_segment + 1]
            frac = min(max(float(os.environ.get("LGS_DP_EARLY_FRAC", "0.0")), 0.0), 1.0)
            mid = a + (int((b - a) * frac) // 4) * 4
            pieces = [(0, a, 0), (a, mid, 1), (mid, b, 2), (b, n, 0)]
        else:
            pieces = [(0, n, 0)]
        self.ranges = [_Range(b, e, ph, self.world, self.rank, dev) for b, e, ph in pieces if e > b]
        self.side = self.ev_bwd = self.ev_late = self.ev_main = None
        # the side-stream launch leaves SM slots to the kernels it runs underneath: 4 CTAs (1024 threads) per SM
        self.late_ctas = int(os.environ.get("LGS_DP_LATE_CTAS", str(148 * 4)))
        if self.overlap:
            with torch.cuda.device(dev):
                self.side = torch.cuda.Stream(device=dev)
                cur = torch.cuda.current_stream(dev)
                self.ev_bwd, self.ev_late, self.ev_main = torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()
                self.ev_bwd.record(cur)   # creates the handles; both start out complete
                self.ev_late.record(cur)
            self._install_hooks()

    def _install_hooks(self):
        global _hook_owner
        check(_lib.lib().lgs_stream_hooks(ctypes.c_void_p(self.ev_late.cuda_event), ctypes.c_void_p(self.ev_bwd.cuda_event)),
              "lgs_stream_hooks")
        _hook_owner = self

    # shard of the single-range layout (kept for callers / tests that address the moments directly)
    @property
    def begin(self):
        return self.ranges[0].sb

    @property
    def end(self):
        return self.ranges[0].se

    @property
    def exp_avg(self):
        return self.ranges[0].exp_avg

    @property
    def exp_avg_sq(self):
        return self.ranges[0].exp_avg_sq

    def mark_rows(self, radii, first):
        """Report the Gaussians this rank rendered in one of the iteration's local views (radii [P] int32 of that forward);
        `first` = the iteration's first local view.  Without any call in an iteration the exchange of that iteration is dense."""
        if not self.sparse:
            return
        with torch.cuda.device(self.dev):
            check(_lib.lib().lgs_dp_rows_mark(self.P, radii.data_ptr(), self.vis.data_ptr(), 0 if first else 1,
                                              torch.cuda.current_stream(self.dev).cuda_stream), "lgs_dp_rows_mark")
        self._rows_marked = True

    def _publish_rows(self, stream):
        """Before the gradient barrier: this rank's visibility bytes into slot [rank] of every rank's table (peer stores)."""
        if not self._rows_marked:
            self.vis.fill_(1)  # nothing reported: every row may be non-zero
        check(_lib.lib().lgs_dp_rows_publish(self.P, self.world, self.rank, self.vis.data_ptr(), self._tp, stream.cuda_stream),
              "lgs_dp_rows_publish")
        self._rows_marked = False

    def _combine_rows(self, stream):
        """After the gradient barrier: the local table (all ranks' bytes have landed) -> one bit per rank and Gaussian."""
        check(_lib.lib().lgs_dp_rows_combine(self.P, self.world, self.table.data_ptr(), self.row_mask.data_ptr(), stream.cuda_stream),
              "lgs_dp_rows_combine")

    def _launch(self, r, stream, max_ctas=0, sparse=False):
        if r.se <= r.sb:
            return
        n_seg = len(self.lrs)
        lr = (ctypes.c_double * n_seg)(*self.lrs)
        if sparse and self.sparse:
            check(_lib.lib().lgs_dp_adam_shard_sparse(n_seg, self._seg, lr, self._row_len, self.P, self.row_mask.data_ptr(), self.world,
                                                      self.rank, self._gp, self._pp, r.sb, r.se, r.exp_avg.data_ptr(),
                                                      r.exp_avg_sq.data_ptr(), float(self.betas[0]), float(self.betas[1]),
                                                      float(self.eps), self.step_count, int(max_ctas), stream.cuda_stream),
                  "lgs_dp_adam_shard_sparse")
            return
        check(_lib.lib().lgs_dp_adam_shard(n_seg, self._seg, lr, self.world, self.rank, self._gp, self._pp,
                                           ctypes.c_void_p(self.g_mc or None), ctypes.c_void_p(self.p_mc or None),
                                           r.sb, r.se, r.exp_avg.data_ptr(), r.exp_avg_sq.data_ptr(),
                                           float(self.betas[0]), float(self.betas[1]), float(self.eps), self.step_count,
                                           int(max_ctas), stream.cuda_stream), "lgs_dp_adam_shard")

    def step(self, late_ready_at_hook=True):
        """One optimizer step.  `late_ready_at_hook`: the late segment's gradient was final when the backward hook fired
        (one local view written in place); otherwise its exchange starts behind everything queued so far."""
        self.step_count += 1
        dev = self.dev
        with torch.cuda.device(dev):
            cur = torch.cuda.current_stream(dev)
            mine = _hook_owner is self  # another instance may have taken the thread's hooks since the last step
            if self.overlap:
                if late_ready_at_hook and mine:
                    self.side.wait_event(self.ev_bwd)
                else:
                    self.side.wait_stream(cur)
                with torch.cuda.stream(self.side):
                    self.hg.barrier(channel=2)  # every rank's render backward is done: late gradients final, late
                    for r in self.ranges:       # parameters no longer read
                        if r.phase == 1:
                            self._launch(r, self.side, self.late_ctas)
            if self.sparse:
                self._publish_rows(cur)
            self.hg.barrier(channel=0)  # every rank's gradients (and visibility bytes) are complete
            if self.sparse:
                self._combine_rows(cur)
            for r in self.ranges:
                if r.phase == 0:
                    self._launch(r, cur, sparse=True)
            if self.overlap:
                self.ev_main.record(cur)
                with torch.cuda.stream(self.side):
                    self.side.wait_event(self.ev_main)
                    for r in self.ranges:
                        if r.phase == 2:
                            self._launch(r, self.side, self.late_ctas, sparse=True)
                    self.hp.barrier(channel=3)  # late parameters landed everywhere, late gradients read by everyone
                    self.ev_late.record(self.side)  # the next render forward waits for this (lgs_stream_hooks)
            self.hp.barrier(channel=1)  # every shard's parameter writes have landed on every rank
            if self.overlap and not mine:  # the next render forward would not wait by itself: wait here, re-install
                cur.wait_event(self.ev_late)
                self._install_hooks()

    def flush(self):
        """Make the caller's stream wait for the side-stream phase (before anything but the rasterizer reads the late
        segment: checkpoints, densification, evaluation)."""
        if self.overlap:
            torch.cuda.current_stream(self.dev).wait_event(self.ev_late)

    def close(self):
        global _hook_owner
        if self.overlap:
            self.flush()
            if _hook_owner is self:
                check(_lib.lib().lgs_stream_hooks(None, None), "lgs_stream_hooks")
                _hook_owner = None
            self.overlap = False

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- optimizer state as full flat vectors (checkpoints, densification): the shards are disjoint, so a sum gathers them
    def gather_moments(self):
        n = self.param_flat.numel()
        m = torch.zeros(n, dtype=torch.float32, device=self.dev)
        v = torch.zeros(n, dtype=torch.float32, device=self.dev)
        for r in self.ranges:
            if r.se > r.sb:
                m[r.sb:r.se] = r.exp_avg[:r.se - r.sb]
                v[r.sb:r.se] = r.exp_avg_sq[:r.se - r.sb]
        dist.all_reduce(m, op=dist.ReduceOp.SUM, group=self.group)
        dist.all_reduce(v, op=dist.ReduceOp.SUM, group=self.group)
        return m, v

    def load_moments(self, m, v, step_count):
        for r in self.ranges:
            if r.se > r.sb:
                r.exp_avg[:r.se - r.sb] = m[r.sb:r.se]
                r.exp_avg_sq[:r.se - r.sb] = v[r.sb:r.se]
        self.step_count = int(step_count)
