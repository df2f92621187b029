#!/usr/bin/env python
"""bench.py -- mapping iterations/s (fwd+bwd+Adam) of the LEG-SLAM hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1], "cfgB"): Replica-shaped synthetic scene, 500 000 Gaussians,
640x480, SH degree 3, 64-D language feature, ONE keyframe per GPU per mapping iteration.
A "step" is one mapping iteration: forward (preprocess, binning, render), backward (render,
preprocess), Adam over the 123 floats/Gaussian.  With N GPUs every rank renders its own view of
the replicated scene and the flat gradient (492 B/Gaussian) is summed with one NCCL all-reduce
before Adam ("weak" scaling); `value` = views processed per second by the whole job.

  value : the kernel path through the C ABI (include/lgs.h) with every input resident in HBM
          and a fixed seeded upstream gradient.
  e2e   : the same iteration through the public Python API that mirrors the reference's
          (GaussianRasterizer autograd + the reference's loss + FusedAdam, leg_slam_b200.mapper),
          with the step's camera and ground-truth images copied from pinned host memory and the
          loss read back inside the timed region.
  --impl reference : the UNMODIFIED reference rasterizer (oracle/_ref, sm_100 recompile) behind
          the same mapper code, with torch.optim.Adam over the reference's 7 groups; same
          workload, same timing.  Its path is CUDA-only, so this arm also runs on the GPU.
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

P_GAUSS, WIDTH, HEIGHT, SH_DEGREE = 500_000, 640, 480, 3
SCENE_SEED = 2
LF_LOWRES = 37          # encoder feature map is 37x37x64
LR_SCALE = 0.1          # keeps the synthetic scene stationary over the timed window; cost is LR-independent
FLOATS_PER_GAUSSIAN = 123
# strings both arms print verbatim (the driver forms its ratios only between lines whose metric / unit / workload agree)
METRIC = "mapping iters/s (fwd+bwd+Adam), 640x480, 64-D feature, 500k Gaussians"
UNIT = "iters/s"
WORKLOAD = ("cfgB: Replica-shaped 500k Gaussians, 640x480, SH deg 3, 64-D language feature, 1 keyframe per GPU per iteration, "
            "fwd+bwd+Adam (BASELINE.json configs[1])")
ARITH = "fp32; blend products 3xTF32 split on tensor cores (hi/lo operands), fp32 accumulate"
H2D_BYTES_PER_VIEW = (35 + 3 * HEIGHT * WIDTH + HEIGHT * WIDTH + 64 * LF_LOWRES * LF_LOWRES) * 4  # camera + RGB + depth + 37x37x64


# ------------------------------------------------------------------------------------------- utils
def dist_setup(n_gpus):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return world, rank, local


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index, period=0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                 nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_evt.wait(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return dict(sm_mhz=(s[len(s) // 2] if s else None), sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons),
                    samples=len(s))


def timed(fn, steps, warmup, world, device):
    """W untimed + exactly K timed calls of fn(i); barrier + synchronize on both sides; device
    time from CUDA events on the current stream; MAX over ranks.  -> ms per step."""
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize(device)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(warmup + i)
    e1.record()
    torch.cuda.synchronize(device)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(device)
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item()) / steps


def make_workload(rank, world, device):
    """Replicated scene + this rank's view + ground truth (pinned host) rendered-looking images."""
    from leg_slam_b200 import synthetic
    sc = synthetic.make_scene(P_GAUSS, seed=SCENE_SEED)
    cams = synthetic.make_cameras(max(8, world), WIDTH, HEIGHT, seed=SCENE_SEED)
    cam = cams[rank % len(cams)]
    g = torch.Generator().manual_seed(100 + rank)
    HW = HEIGHT * WIDTH
    up = dict(dc=torch.randn(3, HEIGHT, WIDTH, generator=g) / HW, dl=torch.randn(64, HEIGHT, WIDTH, generator=g) / HW,
              dd=torch.randn(1, HEIGHT, WIDTH, generator=g) / HW)
    return sc, cam, up


# ------------------------------------------------------------------------- kernel path (C ABI, HBM)
class KernelPath:
    """forward + backward + (all-reduce) + Adam through liblgs.so with preallocated buffers."""

    def __init__(self, sc, cam, up, device, world):
        from leg_slam_b200 import _lib, synthetic
        self.L = _lib.lib()
        self.check = _lib.check
        self.dev, self.world = device, world
        a = synthetic.activate({k: v.to(device) for k, v in sc.items()})
        self.a = {k: v.contiguous() for k, v in a.items()}
        self.dp = None
        self.cam = cam.to(device)
        self.up = {k: v.to(device) for k, v in up.items()}
        P, W, H = P_GAUSS, WIDTH, HEIGHT
        f32 = dict(dtype=torch.float32, device=device)
        u8 = dict(dtype=torch.uint8, device=device)
        self.bg = torch.zeros(3, **f32)
        self.geom = torch.empty(self.L.lgs_geom_bytes(P), **u8)
        self.img = torch.empty(self.L.lgs_image_bytes(W, H), **u8)
        self.binning = torch.empty(0, **u8)
        self.binning_cap = 0
        self.out_color, self.out_lf, self.out_depth = torch.empty(3, H, W, **f32), torch.empty(64, H, W, **f32), torch.empty(1, H, W, **f32)
        self.radii = torch.empty(P, dtype=torch.int32, device=device)
        # flat gradient buffer in Adam order: means3D 3 | sh 48 | lf 64 | opacity 1 | scales 3 | rot 4
        self.order = [("means3D", 3), ("shs", 48), ("lang_feats", 64), ("opacities", 1), ("scales", 3), ("rotations", 4)]
        lrs = dict(means3D=3.2e-4, shs=2.5e-3, lang_feats=1.5e-3, opacities=0.05, scales=5e-3, rotations=1e-3)
        self.dp_mode = "single GPU"
        if world > 1:
            self.dp_mode = "NCCL all-reduce + fused Adam on every replica"
            try:  # parameters and gradients in symmetric memory: one peer-memory kernel does reduce-scatter + Adam + all-gather
                from leg_slam_b200 import dp as dp_mod
                pflat = dp_mod.symmetric_empty(P * FLOATS_PER_GAUSSIAN, device)
                gflat = dp_mod.symmetric_empty(P * FLOATS_PER_GAUSSIAN, device)
                gflat.zero_()
                off = 0
                for k, n in self.order:
                    view = pflat[off:off + P * n].view_as(self.a[k])
                    view.copy_(self.a[k])
                    self.a[k] = view
                    off += P * n
                self.flat = gflat
                self.dp = dp_mod.FusedDPAdam(pflat, gflat, [P * n for _, n in self.order], [lrs[k] * 1e-3 for k, _ in self.order],
                                            late_segment=[k for k, _ in self.order].index("lang_feats"),
                                            rows=(P, [n for _, n in self.order]))
                self.dp_mode = "fused peer-memory reduce-scatter + Adam + all-gather (lgs_dp_adam_shard, " + \
                               ("NVSwitch multimem" if self.dp.uses_multicast else
                                ("P2P loads/stores, culled Gaussians' zero gradient rows skipped" if self.dp.sparse else "P2P loads/stores")) + \
                               (", language-feature exchange overlapped with preprocess/binning on a side stream)" if self.dp.overlap else ")")
            except Exception as ex:  # symmetric memory unavailable on this box: NCCL path
                self.dp = None
                self.dp_mode += f" (symmetric memory unavailable: {type(ex).__name__})"
        if self.dp is None:
            self.flat = torch.zeros(P * FLOATS_PER_GAUSSIAN, **f32)
        self.g, off = {}, 0
        for k, n in self.order:
            self.g[k] = self.flat[off:off + P * n]
            off += P * n
        self.scratch = dict(m2d=torch.empty(P, 3, **f32), conic=torch.empty(P, 4, **f32), color=torch.empty(P, 3, **f32),
                            cov=torch.empty(P, 6, **f32))
        self.m = {k: torch.zeros_like(self.a[k]) for k, _ in self.order}
        self.v = {k: torch.zeros_like(self.a[k]) for k, _ in self.order}
        n = len(self.order)
        VP = ctypes.c_void_p * n
        self.ad = dict(p=VP(*[self.a[k].data_ptr() for k, _ in self.order]), g=VP(*[self.g[k].data_ptr() for k, _ in self.order]),
                       m=VP(*[self.m[k].data_ptr() for k, _ in self.order]), v=VP(*[self.v[k].data_ptr() for k, _ in self.order]),
                       n=(ctypes.c_int64 * n)(*[self.a[k].numel() for k, _ in self.order]),
                       lr=(ctypes.c_double * n)(*[lrs[k] * 1e-3 for k, _ in self.order]))
        self.R = 0
        self.iteration = 0
        self.stream = torch.cuda.current_stream(device).cuda_stream

    def forward(self):
        """stage1 + stage2.  The first call reads num_rendered back (the reference's protocol, rasterizer_impl.cu:281-282) and
        sizes the work buffers for 1.5 x that; every later call passes NULL for the read-back (include/lgs.h): R stays on the
        device, the kernels read it there, and the host never waits inside a step.  The frame status (true R, overflow flag)
        is fetched asynchronously and checked after the timed region (workload_counts)."""
        L, a, cam, P = self.L, self.a, self.cam, P_GAUSS
        R = ctypes.c_int(0)
        learn = self.binning_cap == 0
        self.check(L.lgs_forward_stage1(P, SH_DEGREE, 16, WIDTH, HEIGHT, a["means3D"].data_ptr(), a["shs"].data_ptr(), None,
                                        a["opacities"].data_ptr(), a["scales"].data_ptr(), 1.0, a["rotations"].data_ptr(), None,
                                        cam.viewmatrix.data_ptr(), cam.projmatrix.data_ptr(), cam.campos.data_ptr(),
                                        cam.tanfovx, cam.tanfovy, 0, self.geom.data_ptr(), self.radii.data_ptr(),
                                        ctypes.byref(R) if learn else None, self.stream), "stage1")
        if learn:
            self.binning_cap = int(R.value * 1.5) + 65536
            self.binning = torch.empty(self.L.lgs_binning_bytes(self.binning_cap), dtype=torch.uint8, device=self.dev)
            self.scratch_bwd = torch.empty(self.L.lgs_backward_scratch_bytes(self.binning_cap, WIDTH, HEIGHT), dtype=torch.uint8, device=self.dev)
            self.status_host = torch.zeros(4, dtype=torch.int32).pin_memory()
        self.R = self.binning_cap  # what stage2 and the backward are told: the capacity the buffers were carved for
        self.check(L.lgs_forward_stage2(P, WIDTH, HEIGHT, self.R, self.bg.data_ptr(), a["lang_feats"].data_ptr(),
                                        self.geom.data_ptr(), self.binning.data_ptr(), self.img.data_ptr(),
                                        self.out_color.data_ptr(), self.out_lf.data_ptr(), self.out_depth.data_ptr(), 1,
                                        self.stream), "stage2")

    def backward(self):
        L, a, cam, g, sc = self.L, self.a, self.cam, self.g, self.scratch
        self.check(L.lgs_backward(P_GAUSS, SH_DEGREE, 16, self.R, WIDTH, HEIGHT, self.bg.data_ptr(), a["means3D"].data_ptr(),
                                  a["shs"].data_ptr(), None, a["lang_feats"].data_ptr(), a["scales"].data_ptr(), 1.0,
                                  a["rotations"].data_ptr(), None, cam.viewmatrix.data_ptr(), cam.projmatrix.data_ptr(),
                                  cam.campos.data_ptr(), cam.tanfovx, cam.tanfovy, self.radii.data_ptr(), self.geom.data_ptr(),
                                  self.binning.data_ptr(), self.img.data_ptr(), self.up["dc"].data_ptr(), self.up["dl"].data_ptr(),
                                  self.up["dd"].data_ptr(), sc["m2d"].data_ptr(), sc["conic"].data_ptr(), g["opacities"].data_ptr(),
                                  sc["color"].data_ptr(), g["lang_feats"].data_ptr(), None, g["means3D"].data_ptr(),
                                  sc["cov"].data_ptr(), g["shs"].data_ptr(), g["scales"].data_ptr(), g["rotations"].data_ptr(),
                                  1, 1, self.scratch_bwd.data_ptr(), self.stream), "backward")

    def adam(self):
        self.iteration += 1
        ad = self.ad
        self.check(self.L.lgs_adam_multi(len(self.order), ad["p"], ad["g"], ad["m"], ad["v"], ad["n"], ad["lr"], 0.9, 0.999,
                                         1e-15, self.iteration, self.stream), "adam")

    def step(self, _i=0):
        self.forward()
        self.backward()
        if self.dp is not None:
            self.dp.mark_rows(self.radii, first=True)
            self.dp.step()
            return
        if self.world > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
        self.adam()

    # kernels per step, every one hand-written (no library launch is left on the path): preprocess, emit_keys, 3 radix passes,
    # tile_ranges_fix, render_fwd, render_bwd_pix (which also clears the accumulated-into gradient arrays), render_bwd_chan,
    # preprocess_bwd, adam
    KERNELS_PER_STEP = 11

    def stage_times(self, reps=20):
        """Per-kernel device time (ms, mean over reps) from CUDA events on the launch stream."""
        L = self.L
        L.lgs_profile_enable(1)
        acc = [0.0] * 10
        buf = (ctypes.c_float * 9)()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(reps):
            self.forward()
            self.backward()
            e0.record()
            self.adam()
            e1.record()
            torch.cuda.synchronize(self.dev)
            L.lgs_profile_read(buf, 9)
            for i in range(9):
                acc[i] += max(buf[i], 0.0)
            acc[9] += e0.elapsed_time(e1)
        L.lgs_profile_enable(0)
        names = ["preprocess", "emit_keys", "sort", "tile_ranges", "render_fwd", "zero_grads", "render_bwd_pix",
                 "render_bwd_chan", "preprocess_bwd", "adam"]
        return {n: acc[i] / reps for i, n in enumerate(names)}

    def workload_counts(self):
        from leg_slam_b200 import debug
        self.check(self.L.lgs_forward_status(self.geom.data_ptr(), P_GAUSS, self.status_host.data_ptr(), self.stream), "status")
        torch.cuda.synchronize(self.dev)
        R_true, overflow, sort_error = int(self.status_host[0]), int(self.status_host[2]), int(self.status_host[3])
        assert overflow == 0 and sort_error == 0 and R_true <= self.binning_cap, (R_true, self.binning_cap, overflow, sort_error)
        iv = debug.image_view(self.img, WIDTH, HEIGHT)
        n_tested = int(iv["n_contrib"].long().sum())
        vis = int((self.radii > 0).sum())
        return dict(R=R_true, P_visible=vis, N_tested=n_tested, capacity=self.binning_cap)


def fma_peak_tflops(device):
    from leg_slam_b200 import _lib
    L = _lib.lib()
    sink = torch.zeros(4, device=device)
    s = torch.cuda.current_stream(device).cuda_stream
    blocks, iters = 148 * 16, 4096
    best = 0.0
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.lgs_bench_fma(blocks, iters, sink.data_ptr(), s)
        e1.record()
        torch.cuda.synchronize(device)
        fl = blocks * 256 * iters * 16 * 8 * 2
        best = max(best, fl / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    return best


# --------------------------------------------------------------------------- e2e (public API + host)
class E2EPath:
    """One mapping iteration through leg_slam_b200.mapper (or the reference rasterizer behind the
    same mapper code) with host-resident inputs: per step the camera (35 floats) and the keyframe's
    ground truth (RGB, depth, 37x37x64 feature map) come from pinned host memory; the loss scalar and
    num_rendered go back to the host."""

    def __init__(self, sc, cam, device, world, impl, n_views_local=1, cams=None):
        from leg_slam_b200 import mapper as M
        self.dev, self.world, self.impl = device, world, impl
        params = {k: v.to(device) for k, v in sc.items()}
        lrs = {k: v * LR_SCALE for k, v in M.DEFAULT_LRS.items()}
        kw = {}
        if impl == "reference":
            # the reference's stock path: eager torch ops for the loss, unfused torch Adam
            kw = dict(optimizer_factory=lambda g: torch.optim.Adam(g, lr=0.0, eps=1e-15), render_fn=self._ref_render_fn(),
                      use_cuda_graph=False)
        if impl == "ours" and world > 1:
            kw["dp_mode"] = "fused"
        try:
            self.mapper = M.Mapper(params, lrs=lrs, sh_degree=SH_DEGREE, **kw)
        except Exception:  # symmetric memory unavailable: NCCL all-reduce path
            kw.pop("dp_mode", None)
            self.mapper = M.Mapper(params, lrs=lrs, sh_degree=SH_DEGREE, **kw)
        if impl == "reference":
            self.mapper.world_size, self.mapper.rank = 1, 0  # the reference is single-GPU: rank 0 does every view
        self.M = M
        self.cams = cams if cams is not None else [cam]
        g = torch.Generator().manual_seed(7)
        pin = lambda t: t.contiguous().pin_memory()  # noqa: E731
        self.host = []
        for c in self.cams:
            self.host.append(dict(view=pin(c.viewmatrix), proj=pin(c.projmatrix), campos=pin(c.campos),
                                  gt_image=pin(torch.rand(3, HEIGHT, WIDTH, generator=g)),
                                  gt_depth=pin(torch.rand(1, HEIGHT, WIDTH, generator=g) * 3.0),
                                  gt_lf=pin(torch.randn(64, LF_LOWRES, LF_LOWRES, generator=g)), cam=c))
        self.h2d_bytes = sum(t.numel() * 4 for k, t in self.host[0].items() if k != "cam") * len(self.host)
        self.loss_host = torch.zeros(64).pin_memory()  # ring of per-step results, written by asynchronous device -> host copies
        self.n_steps = 0
        self.loss_events = [None] * 4   # the host consumes step k's loss before it queues step k + 3: bounded run-ahead
        self.last_loss = None
        self.last_R = 0
        self.copy_stream = None
        self._next = None

    def _ref_render_fn(self):
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import build_ref
        ref = build_ref.load()
        outer = self

        class _RefRasterize(torch.autograd.Function):
            """The autograd glue of src/gaussian_rasterizer.cpp:27-176 around the reference's own
            RasterizeGaussiansCUDA / RasterizeGaussiansBackwardCUDA."""
            @staticmethod
            def forward(ctx, means3D, means2D, sh, lang_feats, opacities, scales, rotations, cam, bg):
                e = torch.empty(0, device=means3D.device)
                R, color, lf, depth, radii, geom, binning, img = ref.rasterize_gaussians(
                    bg, means3D, e, lang_feats, opacities, scales, rotations, 1.0, e, cam.viewmatrix, cam.projmatrix,
                    cam.tanfovx, cam.tanfovy, cam.height, cam.width, sh, SH_DEGREE, cam.campos, False, True)
                ctx.cam, ctx.R = cam, R
                outer.last_R = R
                ctx.save_for_backward(bg, lang_feats, means3D, scales, rotations, radii, sh, geom, binning, img)
                ctx.mark_non_differentiable(radii)
                return color, lf, depth, radii

            @staticmethod
            def backward(ctx, gc, gl, gd, _=None):
                bg, lang_feats, means3D, scales, rotations, radii, sh, geom, binning, img = ctx.saved_tensors
                cam = ctx.cam
                e = torch.empty(0, device=means3D.device)
                (dm2, _dc, dlf, dop, dm3, _dcov, dsh, dsc, drot) = ref.rasterize_gaussians_backward(
                    bg, means3D, radii, e, lang_feats, scales, rotations, 1.0, e, cam.viewmatrix, cam.projmatrix, cam.tanfovx,
                    cam.tanfovy, gc.contiguous(), gl.contiguous(), gd.contiguous(), sh, SH_DEGREE, cam.campos, geom, ctx.R,
                    binning, img, True)
                return dm3, dm2, dsh, dlf, dop, dsc, drot, None, None

        def render(cam, a):
            means2D = torch.zeros_like(a["means3D"], requires_grad=True)
            return _RefRasterize.apply(a["means3D"], means2D, a["shs"], a["lang_feats"], a["opacities"], a["scales"],
                                       a["rotations"], cam, outer.mapper.bg)
        return render

    def _upload(self):
        """Host -> device copies of one step's inputs (camera + ground truth of every local keyframe) from pinned
        memory, issued on the copy stream; returns the device-side window and the event that marks their arrival."""
        from leg_slam_b200.synthetic import Camera
        dev = self.dev
        window = []
        with torch.cuda.stream(self.copy_stream):
            for h in self.host:
                c = h["cam"]
                cam = Camera(c.width, c.height, c.tanfovx, c.tanfovy, h["view"].to(dev, non_blocking=True),
                             h["proj"].to(dev, non_blocking=True), h["campos"].to(dev, non_blocking=True))
                window.append(self.M.Keyframe(cam, h["gt_image"].to(dev, non_blocking=True), h["gt_lf"].to(dev, non_blocking=True),
                                              h["gt_depth"].to(dev, non_blocking=True)))
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        return window, ev

    def step(self, _i=0):
        """One mapping iteration.  The step's inputs were uploaded on the copy stream while the previous
        step computed (the next keyframe of a mapper is known one iteration ahead); this step waits for them,
        starts the upload for the next step, computes, and queues the read-back of the loss."""
        cur = torch.cuda.current_stream(self.dev)
        k = self.n_steps
        if k >= 3 and self.loss_events[(k - 3) % 4] is not None:
            self.loss_events[(k - 3) % 4].synchronize()  # step k-3's loss has landed in pinned memory: read it
            self.last_loss = float(self.loss_host[(k - 3) % 64])
        if self.copy_stream is None:
            self.copy_stream = torch.cuda.Stream(device=self.dev)
            self._next = self._upload()
        window, ev = self._next
        cur.wait_event(ev)
        for kf in window:  # tensors produced on the copy stream, consumed on the compute stream
            for t in (kf.camera.viewmatrix, kf.camera.projmatrix, kf.camera.campos, kf.gt_image, kf.gt_lf, kf.gt_depth):
                t.record_stream(cur)
        loss = self.mapper.train_step(window, presharded=True)
        # the next step's upload is queued while this step's kernels are still running (fresh allocations of the copy
        # stream's own pool; record_stream above keeps this step's inputs alive until the compute stream is done with them)
        self._next = self._upload()
        # device -> host read of the step's result: an asynchronous copy into pinned memory, every step, in stream order.  The
        # mapper does not branch on the loss, so the host reads step k's value (above) only when it is about to queue step
        # k + 3: at most three iterations are in flight, and bench.timed() synchronises at the end of the timed region
        self.loss_host[k % 64:k % 64 + 1].copy_(loss.reshape(1), non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(cur)
        self.loss_events[k % 4] = ev
        self.n_steps += 1
        return self.last_loss


# ------------------------------------------------------------- reference kernel path (oracle/_ref, HBM)
class RefKernelPath:
    """The reference arm's counterpart of KernelPath: the UNMODIFIED reference's RasterizeGaussiansCUDA +
    RasterizeGaussiansBackwardCUDA (oracle/_ref/ref_rasterizer.so; reference src/rasterize_points.cu:37-209) on the same
    HBM-resident activated tensors and the same fixed seeded upstream gradients, then torch.optim.Adam over the
    reference's parameter groups (src/gaussian_model.cpp:483-518) -- no activations, no loss, no host copies.  With n views per iteration
    (the reference is single-GPU) their gradients accumulate before the one Adam step."""

    def __init__(self, sc, cams, ups, device):
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import build_ref
        from leg_slam_b200 import synthetic
        self.ref = build_ref.load()
        self.dev = device
        a = synthetic.activate({k: v.to(device) for k, v in sc.items()})
        P = a["means3D"].shape[0]
        self.a = {k: v.contiguous() for k, v in a.items()}
        self.cams = [c.to(device) for c in cams]
        self.ups = [{k: v.to(device) for k, v in u.items()} for u in ups]
        self.bg = torch.zeros(3, dtype=torch.float32, device=device)
        self.e = torch.empty(0, device=device)
        # Adam updates the rasterized tensors in place, like KernelPath (whose Adam segments are these same 6 tensors): the
        # reference's 7 groups with features_dc + features_rest as one [P,16,3] tensor (same 123 floats per Gaussian)
        lrs = dict(means3D=3.2e-4, shs=2.5e-3, lang_feats=1.5e-3, opacities=0.05, scales=5e-3, rotations=1e-3)
        self.p = {k: torch.nn.Parameter(self.a[k]) for k in lrs}
        self.a = {k: (self.p[k].data if k in self.p else v) for k, v in self.a.items()}
        self.opt = torch.optim.Adam([dict(params=[self.p[k]], lr=lrs[k] * 1e-3, name=k) for k in self.p], lr=0.0, eps=1e-15)
        self.last_R = 0
        assert P == P_GAUSS

    def step(self, _i=0):
        ref, a, e, p = self.ref, self.a, self.e, self.p
        first = True
        for cam, up in zip(self.cams, self.ups):
            R, _c, _l, _d, radii, geom, binning, img = ref.rasterize_gaussians(
                self.bg, a["means3D"], e, a["lang_feats"], a["opacities"], a["scales"], a["rotations"], 1.0, e, cam.viewmatrix,
                cam.projmatrix, cam.tanfovx, cam.tanfovy, HEIGHT, WIDTH, a["shs"], SH_DEGREE, cam.campos, False, True)
            (_dm2, _dc, dlf, dop, dm3, _dcov, dsh, dsc, drot) = ref.rasterize_gaussians_backward(
                self.bg, a["means3D"], radii, e, a["lang_feats"], a["scales"], a["rotations"], 1.0, e, cam.viewmatrix,
                cam.projmatrix, cam.tanfovx, cam.tanfovy, up["dc"], up["dl"], up["dd"], a["shs"], SH_DEGREE, cam.campos, geom, R,
                binning, img, True)
            self.last_R = R
            g = dict(means3D=dm3, shs=dsh, lang_feats=dlf, opacities=dop, scales=dsc, rotations=drot)
            for k in p:
                if first:
                    p[k].grad = g[k]
                else:
                    p[k].grad.add_(g[k])
            first = False
        self.opt.step()


# ------------------------------------------------------------------------------- cfgC (strong scaling)
def cfgc_block(rank, world, device, steps=20, warmup=5):
    """BASELINE.json configs[2]: Replica office0-shaped 1 M Gaussians, a FIXED window of 8 keyframes per mapping iteration,
    data-parallel over the views: each of the N ranks renders and back-propagates 8 / N views against the replicated Gaussian
    set, one exchange + Adam per iteration (leg_slam_b200.mapper.Mapper.train_step: fused activations, rasterizer, fused loss,
    device-resident ground truth).  Total work is fixed, so the driver's per-N values of this block are STRONG scaling."""
    from leg_slam_b200 import mapper as M, synthetic
    P, K = 1_000_000, 8
    sc = synthetic.make_scene(P, seed=3)
    cams = synthetic.make_cameras(K, WIDTH, HEIGHT, seed=3)
    g = torch.Generator().manual_seed(33)
    win = [M.Keyframe(c.to(device), torch.rand(3, HEIGHT, WIDTH, generator=g).to(device),
                      torch.randn(64, LF_LOWRES, LF_LOWRES, generator=g).to(device),
                      (torch.rand(1, HEIGHT, WIDTH, generator=g) * 3.0).to(device), None, i) for i, c in enumerate(cams)]
    params = {k: v.to(device) for k, v in sc.items()}
    lrs = {k: v * LR_SCALE for k, v in M.DEFAULT_LRS.items()}
    kw = dict(dp_mode="fused") if world > 1 else {}
    try:
        m = M.Mapper(params, lrs=lrs, sh_degree=SH_DEGREE, **kw)
    except Exception:
        m = M.Mapper(params, lrs=lrs, sh_degree=SH_DEGREE)
    ms = timed(lambda _i: m.train_step(win), steps, warmup, world, device)
    mode = "single GPU: 8 views accumulate, one fused Adam"
    if world > 1:
        mode = ("fused peer-memory reduce-scatter + Adam + all-gather (lgs_dp_adam_shard, " +
                ("NVSwitch multimem" if m.dp.uses_multicast else ("P2P loads/stores, culled Gaussians' zero gradient rows skipped" if m.dp.sparse
                                                                  else "P2P loads/stores")) + ")") if m.dp is not None else "NCCL all-reduce + fused Adam"
    out = {"workload": "cfgC: Replica office0-shaped 1M Gaussians, 640x480, fixed 8-keyframe window per iteration, data-parallel over "
                       "views (BASELINE.json configs[2])", "P": P, "views_per_iteration": K, "views_per_gpu": -(-K // world),
           "n_gpus": world, "scaling": "strong", "ms_per_iteration": round(ms, 4), "views_per_s": round(K * 1000.0 / ms, 2),
           "iterations_per_s": round(1000.0 / ms, 3), "steps": steps, "warmup": warmup, "exchange": mode,
           "last_num_rendered": m.last_num_rendered, "overflow_steps": m.overflow_steps}
    if m.dp is not None:
        m.dp.close()
    del m
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------ cpu baseline
def cpu_baseline(sc, cam, up, target_seconds=12.0):
    """The CPU oracle (oracle/lgs_oracle.c, OpenMP over all host cores) on full cfgB mapping iterations:
    forward + backward + Adam, repeated until ~target_seconds of CPU work.  A reported baseline, not a target."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import oracle as O
    from leg_slam_b200 import synthetic
    O.build()
    a = synthetic.activate(sc)
    n = lambda t: t.numpy()  # noqa: E731
    bg = np.zeros(3, np.float32)
    pairs = (("means3D", "dL_dmeans3D"), ("shs", "dL_dsh"), ("lang_feats", "dL_dlang_feats"), ("opacities", "dL_dopacity"),
             ("scales", "dL_dscales"), ("rotations", "dL_drotations"))
    params = {k: n(a[k]).copy().reshape(-1) for k, _ in pairs}
    m = {k: np.zeros_like(v) for k, v in params.items()}
    v2 = {k: np.zeros_like(v) for k, v in params.items()}
    iters, t0, f = 0, time.perf_counter(), None
    while True:
        f = O.forward(n(a["means3D"]), n(a["opacities"]), n(cam.viewmatrix), n(cam.projmatrix), n(cam.campos), WIDTH, HEIGHT,
                      cam.tanfovx, cam.tanfovy, bg, shs=n(a["shs"]), degree=SH_DEGREE, lang_feat=n(a["lang_feats"]),
                      scales=n(a["scales"]), rotations=n(a["rotations"]))
        g = O.backward(f, n(a["means3D"]), n(cam.viewmatrix), n(cam.projmatrix), n(cam.campos), cam.tanfovx, cam.tanfovy, bg,
                       n(up["dc"]), n(up["dl"]), n(up["dd"]), shs=n(a["shs"]), degree=SH_DEGREE, lang_feat=n(a["lang_feats"]),
                       scales=n(a["scales"]), rotations=n(a["rotations"]))
        iters += 1
        for k, gk in pairs:
            O.adam(params[k], g[gk].reshape(-1), m[k], v2[k], 1e-9, step=iters)
        dt = time.perf_counter() - t0
        if dt >= target_seconds or iters >= 64:
            break
    return dict(value=iters / dt, unit="iters/s", cores=O.num_threads(), kind="port",
                sample=f"{iters} full cfgB mapping iterations (500k Gaussians, 640x480, R={f['num_rendered']}, "
                       f"N_blend={f['n_blended']}; forward + backward + Adam) on the C oracle with OpenMP, {dt:.1f} s"), f


# ------------------------------------------------------------------------------------------- main
_REAL_STDOUT = None


def _claim_stdout():
    """Route everything native code writes to fd 1 (NCCL prints its version banner there when NCCL_DEBUG is set in the
    environment) to stderr, so that stdout carries exactly the one JSON line of the contract."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, line)


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cfgc", action="store_true", help="skip the cfgC strong-scaling block (1 M Gaussians, 8-view window)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    world, rank, local = dist_setup(args.gpus)
    if not torch.cuda.is_available():
        emit({"error": "no CUDA device: the hot path has no CPU fallback"})
        sys.exit(1)
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)

    if args.impl == "reference":
        return main_reference(args, world, rank, device)

    from leg_slam_b200 import build
    if rank == 0:
        build.build()
    if world > 1:
        dist.barrier()

    sc, cam, up = make_workload(rank, world, device)
    kp = KernelPath(sc, cam, up, device, world)
    sampler = ClockSampler(local)
    sampler.start()
    ms = timed(kp.step, args.steps, args.warmup, world, device)
    clocks = sampler.stop()
    counts = kp.workload_counts()

    # ---- e2e through the public API with host buffers
    e2e = E2EPath(sc, cam, device, world, "ours")
    e2e_steps = max(10, args.steps)
    ms_e2e = timed(e2e.step, e2e_steps, max(3, args.warmup), world, device)
    del e2e
    torch.cuda.empty_cache()
    cfgc = None
    if not args.no_cfgc:
        cfgc = cfgc_block(rank, world, device)

    out = None
    if rank == 0:
        stage = kp.stage_times()
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        fma_peak = fma_peak_tflops(device)
        cpu, f = (None, None)
        if world == 1 and not args.no_cpu_baseline:
            cpu, f = cpu_baseline(sc, cam, up)
        n_blend = f["n_blended"] if f is not None else None
        # dominant kernel: render_bwd.  Algorithmic flops: 604*N_blend + 14*N_tested
        top = max(stage, key=stage.get)
        roof = {}
        if n_blend is not None:
            # render bwd = 604*N_blend + 14*N_tested flop, split here as the pixel kernel's
            # replay + 68-term dot + recurrence (136+60 per blend, 14 per test) and the channel kernel's 74-column
            # reduction (2*74 per blend); the remainder of the reference's 604 is work the restructuring removed
            fl = {"render_bwd_pix": 196.0 * n_blend + 14.0 * counts["N_tested"], "render_bwd_chan": 148.0 * n_blend,
                  "render_fwd": 136.0 * n_blend + 14.0 * counts["N_tested"]}
        else:
            fl = {}
        by = {"adam": 3444.0 * P_GAUSS,
              "preprocess": counts["P_visible"] * (44 + 12 * 16 + 75) + (P_GAUSS - counts["P_visible"]) * 20.0,
              "preprocess_bwd": 536.0 * counts["P_visible"], "sort": 16.0 * counts["R"] * 3}
        kernels = {}
        stage.pop("zero_grads", None)  # no longer a launch: the pixel kernel of the render backward clears the arrays
        for k, t in stage.items():
            ent = dict(ms=round(t, 4), share=round(t / sum(stage.values()), 4))
            if k in fl and t > 0:
                ent.update(bound="fp32_fma", achieved=round(fl[k] / (t * 1e-3) / 1e12, 3), peak=round(fma_peak, 2), unit="TFLOP/s")
                ent["frac"] = round(ent["achieved"] / fma_peak, 4)
            elif k in by and t > 0:
                ent.update(bound="hbm", achieved=round(by[k] / (t * 1e-3) / 1e9, 1), peak=hbm_peak, unit="GB/s")
                ent["frac"] = round(ent["achieved"] / hbm_peak, 4)
            kernels[k] = ent
        r = kernels[top]
        traffic, traffic_src = None, None
        try:  # dram__bytes_read.sum + dram__bytes_write.sum per launch from the newest committed ncu --set full capture
            import glob
            cand = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_dram_traffic.json")))
            if cand:
                traffic = json.load(open(cand[-1])).get(top)
                traffic_src = os.path.relpath(cand[-1], ROOT) + " (ncu --set full: dram__bytes_read.sum + dram__bytes_write.sum per launch)"
        except Exception:
            pass
        roof = dict(kernel=top, bound=r.get("bound"), achieved=r.get("achieved"), peak=r.get("peak"), unit=r.get("unit"),
                    frac=r.get("frac"), traffic=traffic, traffic_source=traffic_src, peak_source=("FP32 FFMA peak measured by lgs_bench_fma in this run"
                                                                   if r.get("bound") == "fp32_fma" else hbm_src),
                    ms=r["ms"], share_of_step=r["share"])
        out = {
            "metric": METRIC,
            "value": round(world * 1000.0 / ms, 3), "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "arith": ARITH,
                       "value_is": "keyframe views per second of the whole job through the kernel path: C ABI calls on HBM-resident "
                                   "inputs with a fixed seeded upstream gradient (no activations, no loss, no host copies)",
                       "P": P_GAUSS, "width": WIDTH, "height": HEIGHT, "views_per_iteration": world,
                       "parallelism": f"data-parallel over views x{world}, 492 B/Gaussian exchanged: {kp.dp_mode}" if world > 1 else "single GPU",
                       "l2": "inputs larger than L2: params+grads+Adam state 984 MB per iteration (L2 126 MB), no flush",
                       "R": counts["R"], "P_visible": counts["P_visible"], "N_tested": counts["N_tested"], "N_blend": n_blend,
                       "adam_lr_scale_kernel_path": 1e-3, "e2e_lr_scale": LR_SCALE},
            "e2e": {"value": round(world * 1000.0 / ms_e2e, 3), "unit": UNIT, "ms_per_step": round(ms_e2e, 4), "steps": e2e_steps,
                    "h2d_bytes_per_step": int(H2D_BYTES_PER_VIEW * world), "d2h_bytes_per_step": 20 * world, "api": "leg_slam_b200.mapper.Mapper.train_step (fused activations + rasterizer + "
                                                    "fused loss + FusedAdam, all liblgs launches), inputs from pinned host memory"},
            "gpu_launches": KernelPath.KERNELS_PER_STEP * args.steps,
            "gpu_launches_note": "kernels per step, all hand-written: preprocess, emit_keys, radix_pass x3, tile_ranges_fix, render_fwd, "
                                 "render_bwd_pix (also clears the gradient arrays), render_bwd_chan, preprocess_bwd, adam (no library "
                                 "launch on the path)",
            "clocks": clocks,
            "roofline": roof,
            "kernels": kernels,
            "blended_mfrag_per_s": (round(2.0 * n_blend / ((stage["render_fwd"] + stage["render_bwd_pix"] + stage["render_bwd_chan"]) * 1e-3) / 1e6, 1)
                                    if n_blend else None),
            "fp32_fma_peak_tflops_measured": round(fma_peak, 2),
        }
        if cfgc is not None:
            out["cfgC"] = cfgc
        if cpu is not None:
            out["cpu_baseline"] = cpu
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        emit(out)


def main_reference(args, world, rank, device):
    """The unmodified reference rasterizer (oracle/_ref, recompiled for sm_100) on the same workload.  `value` = its kernel
    path (RefKernelPath: reference forward + backward on the fixed seeded upstream gradients + torch.optim.Adam, HBM-resident,
    like KernelPath); `e2e` = the same mapper code as ours with the reference rasterizer, its activations and loss as eager
    torch ops through autograd and torch.optim.Adam, with the same pinned-host inputs and the loss read back.  The reference's
    path is CUDA-only (no CPU implementation exists), so this arm runs on the GPU.  Single-GPU code: with N > 1 rank 0 alone
    processes the N views of the iteration by gradient accumulation; the other ranks exit."""
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    so = os.path.join(ROOT, "oracle", "_ref", "ref_rasterizer.so")
    if not os.path.exists(so):
        emit({"impl": "reference", "unavailable": "oracle/_ref/ref_rasterizer.so was not built (needs /root/reference)"})
        return
    from leg_slam_b200 import synthetic
    cams = synthetic.make_cameras(max(8, world), WIDTH, HEIGHT, seed=SCENE_SEED)[:world]
    works = [make_workload(r, world, device) for r in range(world)]  # the views and upstream gradients the N ranks of our arm use
    sc, cam = works[0][0], works[0][1]
    sampler = ClockSampler(device.index or 0)
    sampler.start()
    kp = RefKernelPath(sc, [w[1] for w in works], [w[2] for w in works], device)
    ms = timed(kp.step, args.steps, args.warmup, 1, device)
    clocks = sampler.stop()
    R = kp.last_R
    del kp
    torch.cuda.empty_cache()
    e2e = E2EPath(sc, cam, device, 1, "reference", cams=cams)
    e2e_steps = max(10, args.steps)
    ms_e2e = timed(e2e.step, e2e_steps, max(3, args.warmup), 1, device)
    val = round(world * 1000.0 / ms, 3)
    val_e2e = round(world * 1000.0 / ms_e2e, 3)
    out = {
        "impl": "reference",
        "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms, 4), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "arith": "fp32 (CUDA cores)",
                   "value_is": "keyframe views per second through the reference's kernel path: RasterizeGaussiansCUDA + "
                               "RasterizeGaussiansBackwardCUDA on HBM-resident inputs with the same fixed seeded upstream gradient + "
                               "torch.optim.Adam (eps 1e-15, in place on the same 6 tensors as ours); no activations, no loss, no host copies",
                   "P": P_GAUSS, "width": WIDTH, "height": HEIGHT, "views_per_iteration": world, "R": R,
                   "parallelism": "single GPU" if world == 1 else "reference is single-GPU: rank 0 accumulates the iteration's views",
                   "l2": "inputs larger than L2: params+grads+Adam state 984 MB per iteration (L2 126 MB), no flush",
                   "path": "reference cuda_rasterizer + rasterize_points.cu recompiled for sm_100 (oracle/_ref); its implementation of "
                           "this path is CUDA-only, so the reference arm runs on the same GPU, not on host cores",
                   "adam_lr_scale_kernel_path": 1e-3, "e2e_lr_scale": LR_SCALE},
        "e2e": {"value": val_e2e, "unit": UNIT, "ms_per_step": round(ms_e2e, 4), "steps": e2e_steps,
                "h2d_bytes_per_step": int(H2D_BYTES_PER_VIEW * world), "d2h_bytes_per_step": 4 + 4 * world,
                "api": "leg_slam_b200.mapper.Mapper.train_step with the reference rasterizer behind autograd, activations + reference "
                       "loss as eager torch ops, torch.optim.Adam (7 groups); inputs from pinned host memory, loss read back"},
        "cpu_baseline": {"value": val, "unit": UNIT, "kind": "reference", "cores": 0,
                         "sample": "the reference's implementation of this path is CUDA-only (no CPU code exists); this arm times its "
                                   "unmodified kernels on the same GPU instead of host cores"},
        "clocks": clocks,
    }
    emit(out)


if __name__ == "__main__":
    main()
