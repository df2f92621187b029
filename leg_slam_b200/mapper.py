"""One mapping iteration over a window of keyframes, data-parallel over views.

Host-side mirror of the part of GaussianMapper::trainForOneIteration that is on the hot path
(reference src/gaussian_mapper.cpp:687-796): activations -> GaussianRasterizer (ours) -> loss
-> backward -> Adam, with the reference's 7 single-tensor Adam groups
(src/gaussian_model.cpp:483-518).

New relative to the reference (which renders ONE keyframe per iteration on one GPU,
src/gaussian_mapper.cpp:629): the K views of a window are split over the ranks of a
torch.distributed process group; every rank renders and back-propagates its views against the
replicated Gaussian set; gradients accumulate into ONE flat fp32 buffer [123*P] (the 7 .grad
tensors are views into it), which is summed across ranks with a single NCCL all-reduce over
NVLink (492 B per Gaussian) before the fused Adam step.  Gradients are
summed, not averaged.
"""
from typing import Callable, NamedTuple, Optional, Sequence

import torch
import torch.distributed as dist

from . import fused as fused_mod
from . import loss as loss_mod
from . import rasterize_points as rp
from .optim import FusedAdam
from .rasterizer import GaussianRasterizationSettings, GaussianRasterizer
from .synthetic import Camera

PARAM_ORDER = ("xyz", "features_dc", "features_rest", "lang_feat", "opacity", "scaling", "rotation")
# learning rates of the reference's Replica configuration
DEFAULT_LRS = dict(xyz=3.2e-4, features_dc=2.5e-3, features_rest=2.5e-3 / 20.0, lang_feat=1.5e-3, opacity=0.05,
                   scaling=5e-3, rotation=1e-3)


class ExponLr:
    """GaussianModel::exponLrFunc (reference src/gaussian_model.cpp:1143-1157) with the constants trainingSetup stores for it
    (:513-517: lr_init / lr_final already multiplied by spatial_lr_scale): log-linear interpolation from lr_init at step 0 to
    lr_final at max_steps, times an optional sine-eased delay factor.  Evaluated like the reference: float operands, float
    products and sums, and the C library's own logf / expf / sinf (through ctypes -- numpy's float32 log / exp are a different
    implementation and land one ulp away on some steps), so the rate is the same float the reference's compiled code returns
    (tests/test_reference_model.py holds it to the unmodified class, step for step)."""

    _libm = None

    @classmethod
    def _m(cls):
        if cls._libm is None:
            import ctypes
            import ctypes.util
            m = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
            for name in ("logf", "expf", "sinf"):
                fn = getattr(m, name)
                fn.restype, fn.argtypes = ctypes.c_float, [ctypes.c_float]
            cls._libm = m
        return cls._libm

    def __init__(self, lr_init, lr_final, lr_delay_mult=1.0, max_steps=1_000_000, lr_delay_steps=0):
        import numpy as np
        self._np = np
        f = np.float32
        self.lr_init, self.lr_final, self.lr_delay_mult = f(lr_init), f(lr_final), f(lr_delay_mult)
        self.max_steps, self.lr_delay_steps = int(max_steps), int(lr_delay_steps)

    def __call__(self, step: int) -> float:
        f = self._np.float32
        m = self._m()
        if step < 0 or (self.lr_init == 0 and self.lr_final == 0):
            return 0.0
        if self.lr_delay_steps > 0:
            x = min(max(f(step) / f(self.lr_delay_steps), f(0)), f(1))
            delay = self.lr_delay_mult + (f(1) - self.lr_delay_mult) * f(m.sinf(f(1.57079632679489661923) * x))
        else:
            delay = f(1)
        t = min(max(f(step) / f(self.max_steps), f(0)), f(1))
        log_lerp = f(m.expf(f(m.logf(self.lr_init)) * (f(1) - t) + f(m.logf(self.lr_final)) * t))
        return float(f(delay) * log_lerp)


class DensityControlParams(NamedTuple):
    """The settings GaussianMapper's density control reads (reference src/gaussian_mapper.cpp:338-352; defaults
    include/gaussian_parameters.h:68-75, the Replica configuration overrides them in cfg/gaussian_mapper/RGB-D/Replica/
    replica_rgbd.yaml:69-74)."""
    densification_interval: int = 100
    opacity_reset_interval: int = 3000      # 0: never
    densify_from_iter: int = 500
    densify_until_iter: int = 15_000
    densify_grad_threshold: float = 0.0002
    densify_min_opacity: float = 0.005
    prune_big_point_after_iter: int = 0
    white_background: bool = False


def density_control_actions(iteration: int, p: DensityControlParams):
    """What trainForOneIteration does to the Gaussian set after the backward of iteration `iteration` (reference
    src/gaussian_mapper.cpp:737-761) -> dict(update_stats, densify, size_threshold, reset_opacity).  Pure host logic."""
    out = dict(update_stats=False, densify=False, size_threshold=0, reset_opacity=False)
    if iteration < p.densify_until_iter:
        out["update_stats"] = True
        if iteration > p.densify_from_iter and iteration % p.densification_interval == 0:
            out["densify"] = True
            out["size_threshold"] = 20 if iteration > p.prune_big_point_after_iter else 0
        if p.opacity_reset_interval and (iteration % p.opacity_reset_interval == 0 or
                                         (p.white_background and iteration == p.densify_from_iter)):
            out["reset_opacity"] = True
    return out


def optimizer_step_groups(actions: dict):
    """Which parameter groups the optimizer step at the END of trainForOneIteration (reference src/gaussian_mapper.cpp:793-797)
    still moves after that iteration's density control (`density_control_actions`): a densification rebuilds all seven tensors
    and a resetOpacity the opacity, the rebuilt tensors carry no gradient, and libtorch's Adam skips parameters without one.
    -> tuple of names from PARAM_ORDER.  Pure host logic (held to the reference's own lines on CPU)."""
    if actions.get("densify"):
        return ()
    return tuple(k for k in PARAM_ORDER if not (k == "opacity" and actions.get("reset_opacity")))


class Keyframe(NamedTuple):
    camera: Camera
    gt_image: torch.Tensor   # [3,H,W]
    gt_lf: torch.Tensor      # [64,h,w] encoder resolution (37x37 in the reference), resized per iteration
    gt_depth: torch.Tensor   # [1,H,W]
    mask: Optional[torch.Tensor] = None  # [3,H,W] undistortion mask (ones when absent)
    kf_id: Optional[int] = None          # keyframe identity, when the caller has one (see Mapper: read-back-free forward)


def shard_views(n_views: int, rank: int, world_size: int):
    """Indices of the window's views this rank renders (round-robin, so any K works)."""
    return list(range(rank, n_views, world_size))


class FlatGrads:
    """The 7 parameters' .grad tensors as views into one contiguous buffer (one collective,
    one memset)."""

    def __init__(self, params: dict, flat: Optional[torch.Tensor] = None):
        # every view starts on a 16-byte boundary (the kernels write float4 rows); no padding when P % 4 == 0
        n = sum((params[k].numel() + 3) & ~3 for k in PARAM_ORDER)
        any_p = params[PARAM_ORDER[0]]
        self.flat = torch.zeros(n, dtype=torch.float32, device=any_p.device) if flat is None else flat
        off = 0
        self.views = {}
        for k in PARAM_ORDER:
            p = params[k]
            self.views[k] = self.flat[off:off + p.numel()].view_as(p)
            off += (p.numel() + 3) & ~3

    def attach(self, params: dict):
        for k in PARAM_ORDER:
            params[k].grad = self.views[k]

    def zero_(self):
        self.flat.zero_()


class Mapper:
    """Replicated Gaussian set + fused Adam; `train_step(window)` is one mapping iteration."""

    def __init__(self, params: dict, lrs: Optional[dict] = None, sh_degree: int = 3,
                 process_group=None, optimizer_factory: Optional[Callable] = None,
                 render_fn: Optional[Callable] = None, faithful_loss_sign: bool = True,
                 use_cuda_graph: bool = True, fused: bool = True, dp_mode: str = "allreduce",
                 track_densify_stats: bool = False, async_forward: bool = True):
        """dp_mode: "allreduce" = NCCL all-reduce of the flat gradient, then Adam on every replica;
        "fused" = one peer-memory kernel per rank doing reduce-scatter + Adam-on-shard + all-gather
        (leg_slam_b200.dp.FusedDPAdam; needs world_size > 1, CUDA, P % 4 == 0).
        track_densify_stats: accumulate the densification statistics of every rendered view inside train_step
        (GaussianModel::addDensificationStats, reference src/gaussian_mapper.cpp:737-744) for densify_and_prune.
        async_forward: the fused path renders WITHOUT the reference's blocking read-back of num_rendered
        (rasterizer_impl.cu:281-282): work buffers are sized for 1.5 x the largest instance count seen, R stays on the
        device (include/lgs.h), and the host never waits inside an iteration.  The count is re-learnt with one synchronous
        forward whenever it could jump -- the first step, a new `Keyframe.kf_id`, a changed number of Gaussians, a reported
        overflow -- and every step's status comes back asynchronously (`last_num_rendered`, `overflow_steps`)."""
        lrs = dict(DEFAULT_LRS, **(lrs or {}))
        self._lrs = lrs
        self._optimizer_factory = optimizer_factory
        world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.dp = None
        self._dp_group = process_group
        if dp_mode == "fused" and world > 1 and params["xyz"].is_cuda and render_fn is None and fused:
            self._init_fused_dp({k: params[k].detach() for k in PARAM_ORDER})
        else:
            self.params = {k: torch.nn.Parameter(params[k].detach().clone().contiguous()) for k in PARAM_ORDER}
            groups = [dict(params=[self.params[k]], lr=lrs[k], name=k) for k in PARAM_ORDER]
            self.optimizer = (optimizer_factory or (lambda g: FusedAdam(g, lr=0.0, eps=1e-15)))(groups)
            self.grads = FlatGrads(self.params)
        self.sh_degree = sh_degree          # active_sh_degree_: what the rasterizer evaluates
        self.max_sh_degree = sh_degree      # max_sh_degree_: what features_rest can hold (set_sh_degree clamps to it)
        self.xyz_lr_schedule = None         # ExponLr, installed by set_position_lr_schedule
        self.pg = process_group
        self.world_size = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(process_group) if dist.is_initialized() else 0
        self.render_fn = render_fn or self._render
        self.faithful_loss_sign = faithful_loss_sign
        dev = self.params["xyz"].device
        self.bg = torch.zeros(3, dtype=torch.float32, device=dev)
        self.last_num_views = 0
        self.use_cuda_graph = use_cuda_graph
        self._loss_graphs = {}
        # fused fast path: activations, loss and their backward as liblgs kernels, gradients written
        # straight into the flat buffer (no autograd graph).  Needs our rasterizer and CUDA tensors.
        self.fused = bool(fused and render_fn is None and dev.type == "cuda")
        self._fused_loss = fused_mod.FusedMappingLoss(faithful_sign=faithful_loss_sign) if self.fused else None
        self._fbuf = None
        self.stats = None
        # read-back-free forward: capacity of the work buffers (instances), what justifies it, pending status copies
        self.async_forward = bool(async_forward)
        self._cap, self._cap_P, self._seen_kf = 0, -1, set()
        self._work = {}
        self._pending_status, self._status_pool = [], []
        self.last_num_rendered = 0
        self.overflow_steps = 0
        if track_densify_stats:
            if not self.fused:
                raise ValueError("track_densify_stats needs the fused path (CUDA tensors, our rasterizer)")
            from .densify import DensifyStats
            self.stats = DensifyStats(self.params["xyz"].shape[0], dev)

    # -- fused data-parallel state: parameters + gradients in symmetric memory, Adam moments sharded (leg_slam_b200.dp)
    def _init_fused_dp(self, tensors: dict, moments=None, step_count: int = 0):
        """(Re)build the symmetric-memory parameter / gradient buffers and the sharded optimizer for `tensors`; collective
        (every rank calls it with identical tensors).  Tensors start on 16-byte boundaries of the flat index space, as in
        FlatGrads; `moments` = (exp_avg, exp_avg_sq) dicts of full tensors to carry over."""
        from . import dp as dp_mod
        if self.dp is not None:
            self.dp.close()
        dev0 = tensors["xyz"].device
        sizes = [(tensors[k].numel() + 3) & ~3 for k in PARAM_ORDER]
        n = sum(sizes)
        pflat = dp_mod.symmetric_empty(n, dev0)
        gflat = dp_mod.symmetric_empty(n, dev0)
        pflat.zero_()
        gflat.zero_()
        off, self.params = 0, {}
        for k, sz in zip(PARAM_ORDER, sizes):
            t = tensors[k].contiguous()
            view = pflat[off:off + t.numel()].view_as(t)
            view.copy_(t)
            self.params[k] = torch.nn.Parameter(view, requires_grad=False)
            off += sz
        self.grads = FlatGrads(self.params, flat=gflat)
        torch.cuda.synchronize(dev0)
        P = int(tensors["xyz"].shape[0])
        self.dp = dp_mod.FusedDPAdam(pflat, gflat, sizes, [self._lrs[k] for k in PARAM_ORDER], group=self._dp_group,
                                     late_segment=PARAM_ORDER.index("lang_feat"),
                                     rows=(P, [tensors[k].numel() // max(P, 1) for k in PARAM_ORDER]) if P > 0 else None)
        self.optimizer = None
        if moments is not None:
            m = torch.zeros(n, dtype=torch.float32, device=dev0)
            v = torch.zeros(n, dtype=torch.float32, device=dev0)
            off = 0
            for k, sz in zip(PARAM_ORDER, sizes):
                m[off:off + tensors[k].numel()] = moments[0][k].reshape(-1)
                v[off:off + tensors[k].numel()] = moments[1][k].reshape(-1)
                off += sz
            self.dp.load_moments(m, v, step_count)
        else:
            self.dp.step_count = int(step_count)

    def _fused_dp_moments(self):
        """Full (exp_avg, exp_avg_sq) dicts gathered from the ranks' shards, and the common step count."""
        self.dp.flush()
        m, v = self.dp.gather_moments()
        md, vd, off = {}, {}, 0
        for k in PARAM_ORDER:
            t = self.params[k]
            md[k] = m[off:off + t.numel()].view_as(t)
            vd[k] = v[off:off + t.numel()].view_as(t)
            off += (t.numel() + 3) & ~3
        return md, vd, self.dp.step_count

    # -- activations exactly as the reference applies them each iteration (gaussian_model.cpp:46-68)
    def activated(self):
        p = self.params
        return dict(means3D=p["xyz"], shs=torch.cat([p["features_dc"], p["features_rest"]], dim=1),
                    lang_feats=p["lang_feat"], opacities=torch.sigmoid(p["opacity"]),
                    scales=torch.exp(p["scaling"]), rotations=torch.nn.functional.normalize(p["rotation"]))

    def _render(self, cam: Camera, a: dict):
        rs = GaussianRasterizationSettings(cam.height, cam.width, cam.tanfovx, cam.tanfovy, self.bg, 1.0,
                                           cam.viewmatrix, cam.projmatrix, self.sh_degree, cam.campos, False, True)
        means2D = torch.zeros_like(a["means3D"], requires_grad=True)  # screenspace_points (gaussian_renderer.cpp:41-48)
        return GaussianRasterizer(rs)(a["means3D"], means2D, a["opacities"], shs=a["shs"], lang_feats=a["lang_feats"],
                                      scales=a["scales"], rotations=a["rotations"])

    # -- the loss and its gradient w.r.t. the rendered images as ONE CUDA-graph replay -----------------
    def _graphed_loss(self, image, lf, depth, gt_image, gt_lf, gt_depth, mask):
        """loss and dL/d(image, lf, depth) of `loss_mod.mapping_loss`, captured once per image shape
        into a CUDA graph (static shapes, ~60 small torch kernels forward+backward) and replayed:
        the same stock ops, without their per-launch host overhead."""
        key = (tuple(image.shape), tuple(gt_lf.shape), mask is not None)
        g = self._loss_graphs.get(key)
        if g is None:
            st = dict(image=torch.empty_like(image), lf=torch.empty_like(lf), depth=torch.empty_like(depth),
                      gt_image=torch.empty_like(gt_image), gt_lf=torch.empty_like(gt_lf),
                      gt_depth=torch.empty_like(gt_depth), mask=None if mask is None else torch.empty_like(mask))

            def run():
                im = st["image"].detach().requires_grad_(True)  # fresh leaves aliasing the static buffers
                l_ = st["lf"].detach().requires_grad_(True)
                d_ = st["depth"].detach().requires_grad_(True)
                loss = self._loss_from_images(im, l_, d_, st["gt_image"], st["gt_lf"], st["gt_depth"], st["mask"])
                gi, gl, gd = torch.autograd.grad(loss, [im, l_, d_])
                return loss.detach(), gi, gl, gd

            for k, v in (("image", image), ("lf", lf), ("depth", depth), ("gt_image", gt_image), ("gt_lf", gt_lf),
                         ("gt_depth", gt_depth)):
                st[k].copy_(v.detach())
            if mask is not None:
                st["mask"].copy_(mask)
            side = torch.cuda.Stream(device=image.device)
            side.wait_stream(torch.cuda.current_stream(image.device))
            with torch.cuda.stream(side):  # warm-up off the capture stream (cuDNN/allocator init)
                for _ in range(3):
                    run()
            torch.cuda.current_stream(image.device).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = run()
            g = (graph, st, out)
            self._loss_graphs[key] = g
        graph, st, out = g
        st["image"].copy_(image.detach())
        st["lf"].copy_(lf.detach())
        st["depth"].copy_(depth.detach())
        st["gt_image"].copy_(gt_image)
        st["gt_lf"].copy_(gt_lf)
        st["gt_depth"].copy_(gt_depth)
        if mask is not None:
            st["mask"].copy_(mask)
        graph.replay()
        return out

    def _loss_from_images(self, image, lf, depth, gt_image, gt_lf, gt_depth, mask):
        # gt feature map resized to the render size, nearest (src/gaussian_mapper.cpp:707-708)
        if gt_lf.shape[-2:] != lf.shape[-2:]:
            gt_lf = torch.nn.functional.interpolate(gt_lf.unsqueeze(0), size=tuple(lf.shape[-2:])).squeeze(0)
        if mask is not None:  # :711-713
            image, lf, depth = image * mask, lf * mask[0:1], depth * mask[0:1]
        return loss_mod.mapping_loss(image, lf, depth, gt_image, gt_lf, gt_depth, faithful_sign=self.faithful_loss_sign)

    def _train_views_fused(self, window, mine):
        """The views of this rank without autograd: every op between the raw parameters and the flat
        gradient buffer is a liblgs launch."""
        p = {k: v.data for k, v in self.params.items()}
        P = p["xyz"].shape[0]
        dev = p["xyz"].device
        n_rest = p["features_rest"].shape[1]
        if self._fbuf is None:
            f = dict(dtype=torch.float32, device=dev)
            self._fbuf = dict(act=dict(scales=torch.empty(P, 3, **f), rotations=torch.empty(P, 4, **f),
                                       opacities=torch.empty(P, 1, **f), shs=None),
                              tmp=rp.backward_outputs(P, 0, dev, True, False, True),
                              empty=torch.empty(0, **f))
        fb = self._fbuf
        e = fb["empty"]
        # parameters do not change between the views of a step.  No SH cat: the rasterizer reads features_dc /
        # features_rest in place and writes their gradients straight into the flat buffer (split-SH entry points)
        a = fused_mod.activations_fwd(p, out=fb["act"], cat_sh=False)
        gv = self.grads.views
        total = None
        for n_done, i in enumerate(mine):
            kf = window[i]
            cam = kf.camera
            cap = self._forward_capacity(P, kf)
            R, color, lf, depth, radii, geom, binning, img = rp.rasterize_gaussians(
                self.bg, a["means3D"], e, a["lang_feats"], a["opacities"], a["scales"], a["rotations"], 1.0, e,
                cam.viewmatrix, cam.projmatrix, cam.tanfovx, cam.tanfovy, cam.height, cam.width, p["features_dc"],
                self.sh_degree, cam.campos, False, True, sh_rest=p["features_rest"], capacity=cap, buffers=self._work)
            self._note_forward(P, kf, cap, R, geom)
            if self.dp is not None:  # sparse gradient exchange: which rows of this rank's gradient can be non-zero
                self.dp.mark_rows(radii, first=n_done == 0)
            loss, gi, gl, gd = self._fused_loss(color, lf, depth, kf.gt_image, kf.gt_lf, kf.gt_depth, kf.mask)
            first = n_done == 0
            out = dict(fb["tmp"])
            out["dL_dfeatures_dc"], out["dL_dfeatures_rest"] = gv["features_dc"], gv["features_rest"]
            if first:  # xyz and language-feature gradients need no activation backward: write them in place
                out["dL_dmeans3D"], out["dL_dlang_feats"] = gv["xyz"], gv["lang_feat"]
            rp.rasterize_gaussians_backward_into(
                out, self.bg, a["means3D"], radii, e, a["lang_feats"], a["scales"], a["rotations"], 1.0, e, cam.viewmatrix,
                cam.projmatrix, cam.tanfovx, cam.tanfovy, gi, gl, gd, p["features_dc"], self.sh_degree, cam.campos, geom, R,
                binning, img, True, sh_rest=p["features_rest"], accumulate_sh=not first)
            if self.stats is not None:  # max_radii2D + addDensificationStats of this view (gaussian_mapper.cpp:737-744)
                self.stats.add(radii, out["dL_dmeans2D"])
            if not first:
                gv["xyz"].add_(out["dL_dmeans3D"])
                gv["lang_feat"].add_(out["dL_dlang_feats"])
            fused_mod.activations_bwd(p, a, out["dL_dscales"], out["dL_drotations"], out["dL_dopacity"], None, gv,
                                      accumulate=not first)
            l0 = loss[0:1].clone().reshape(())
            total = l0 if total is None else total + l0
        return total

    # -- read-back-free forward: capacity bookkeeping ------------------------------------------------------------------
    def _forward_capacity(self, P, kf):
        """None = render synchronously (learn R); otherwise the instance capacity to render with."""
        self._drain_status()
        if not self.async_forward or self._cap == 0 or self._cap_P != P or (kf.kf_id is not None and kf.kf_id not in self._seen_kf):
            return None
        return self._cap

    def _note_forward(self, P, kf, cap, R, geom):
        if kf.kf_id is not None:
            self._seen_kf.add(kf.kf_id)
        if cap is None:  # synchronous forward: R is exact
            self.last_num_rendered = R
            if self._cap_P != P:
                self._cap = 0
            self._cap = max(self._cap, int(R * 1.5) + 65536)
            self._cap_P = P
            return
        if P == 0:
            return
        dev = geom.device
        st = self._status_pool.pop() if self._status_pool else torch.zeros(4, dtype=torch.int32).pin_memory()
        from . import _lib
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().lgs_forward_status(geom.data_ptr(), int(P), st.data_ptr(),
                                                     torch.cuda.current_stream(dev).cuda_stream), "lgs_forward_status")
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(dev))
        self._pending_status.append((st, ev, cap))

    def _drain_status(self):
        """Consume the status copies whose events have completed: track R, grow the capacity ahead of need, and re-learn it
        with a synchronous forward after an overflow (that frame's lists were incomplete: its update used partial gradients)."""
        keep = []
        for st, ev, cap in self._pending_status:
            if not ev.query():
                keep.append((st, ev, cap))
                continue
            R, overflow = int(st[0]), int(st[2])
            self._status_pool.append(st)
            self.last_num_rendered = R
            if overflow:
                self.overflow_steps += 1
                self._cap = 0  # next forward is synchronous
            elif self._cap and R > 0.8 * self._cap:
                self._cap = int(R * 1.5) + 65536
        self._pending_status = keep

    # -- .ply checkpoints (reference GaussianModel::savePly / loadPly, src/gaussian_model.cpp:854-1075)
    def save_checkpoint(self, path):
        """The Gaussian set in the reference's .ply layout plus the Adam moments and step counts (leg_slam_b200.ply_io):
        loads in the reference as a plain model, resumes here with `load_checkpoint`."""
        from . import ply_io
        p = {k: v.data for k, v in self.params.items()}
        if self.dp is not None:  # collective: the shards of the Adam state are gathered; rank 0 writes the file
            m, v, step = self._fused_dp_moments()
            if self.rank == 0:
                ply_io.save_ply(path, p, m, v, {k: step for k in PARAM_ORDER})
            dist.barrier(group=self.pg)
            return
        m, v, steps = self._optimizer_moments(p)
        ply_io.save_ply(path, p, m, v, {k: int(s) for k, s in steps.items()})

    def load_checkpoint(self, path):
        """Replace this mapper's Gaussian set and optimizer state by a checkpoint written by `save_checkpoint` (or by a
        plain reference .ply: fresh Adam state)."""
        from . import ply_io
        dev = self.params["xyz"].device
        p2, m2, v2, steps = ply_io.load_ply(path, dev, self.sh_degree)
        if self.dp is not None:  # collective: every rank reads the same file
            if m2 is not None and len({int(steps[k]) for k in PARAM_ORDER}) != 1:
                raise ValueError("dp_mode='fused' keeps one Adam step count for all tensors")
            self._init_fused_dp(p2, None if m2 is None else (m2, v2), 0 if m2 is None else int(steps[PARAM_ORDER[0]]))
            self._fbuf = None
            if self.stats is not None:
                from .densify import DensifyStats
                self.stats = DensifyStats(self.params["xyz"].shape[0], dev)
            return
        self.params = {k: torch.nn.Parameter(p2[k]) for k in PARAM_ORDER}
        groups = [dict(params=[self.params[k]], lr=self._lrs[k], name=k) for k in PARAM_ORDER]
        self.optimizer = (self._optimizer_factory or (lambda g: FusedAdam(g, lr=0.0, eps=1e-15)))(groups)
        if m2 is not None:
            for k in PARAM_ORDER:
                self.optimizer.state[self.params[k]] = dict(step=steps[k], exp_avg=m2[k], exp_avg_sq=v2[k])
        self.grads = FlatGrads(self.params)
        self._fbuf = None
        if self.stats is not None:
            from .densify import DensifyStats
            self.stats = DensifyStats(self.params["xyz"].shape[0], dev)

    def _release_state(self):
        """Drop every reference this mapper holds to the current parameter / gradient / optimizer-state / work tensors, so
        that their blocks return to the caching allocator now (by reference count) rather than at some later garbage
        collection: the buffers of the next, slightly larger P then reuse them instead of costing a cudaMalloc each."""
        if self.optimizer is not None:
            self.optimizer.state.clear()
            for g in self.optimizer.param_groups:
                g["params"] = []
            self.optimizer = None
        for t in self.params.values():
            t.grad = None
        if self.grads is not None:
            self.grads.views.clear()
            self.grads.flat = None
            self.grads = None
        self._fbuf = None
        self.stats = None

    def _optimizer_moments(self, p):
        m, v, steps = {}, {}, {}
        for k in PARAM_ORDER:
            st = self.optimizer.state.get(self.params[k], {})
            m[k] = st["exp_avg"] if "exp_avg" in st else torch.zeros_like(p[k])
            v[k] = st["exp_avg_sq"] if "exp_avg_sq" in st else torch.zeros_like(p[k])
            steps[k] = st.get("step", 0)
        return m, v, steps

    def densify_and_prune(self, max_grad, min_opacity, extent, max_screen_size, percent_dense=0.01, generator=None,
                          empty_cache: bool = False):
        """GaussianModel::densifyAndPrune (reference src/gaussian_model.cpp:806-824) on this mapper's Gaussian set: clone /
        split / prune in one fused gather (leg_slam_b200.densify), carrying the Adam moments and step counts over and
        rebuilding the flat gradient buffer.  With several ranks the statistics are summed (max for the radii) first and
        every rank must pass a generator in the same state, so that the replicas stay identical.
        `empty_cache` repeats the reference's c10::cuda::CUDACachingAllocator::emptyCache() (:823); it is allocator
        hygiene, not semantics, and costs up to 0.5 s of cudaFree / cudaMalloc at 3-6 M Gaussians, so it is off by default."""
        from . import densify as densify_mod
        if self.stats is None:
            raise ValueError("construct the Mapper with track_densify_stats=True")
        if self.world_size > 1:
            dist.all_reduce(self.stats.xyz_gradient_accum, op=dist.ReduceOp.SUM, group=self.pg)
            dist.all_reduce(self.stats.denom, op=dist.ReduceOp.SUM, group=self.pg)
            dist.all_reduce(self.stats.max_radii2D, op=dist.ReduceOp.MAX, group=self.pg)
        p = {k: v.data for k, v in self.params.items()}
        if self.dp is not None:
            m, v, step = self._fused_dp_moments()
            steps = {k: step for k in PARAM_ORDER}
        else:
            m, v, steps = self._optimizer_moments(p)
        with torch.no_grad():
            p2, m2, v2, stats2, info = densify_mod.densify_and_prune(p, m, v, self.stats, max_grad, min_opacity, extent,
                                                                    max_screen_size, percent_dense, generator)
        del p, m, v
        self._rebind(p2, m2, v2, steps, stats2)
        if empty_cache:
            torch.cuda.empty_cache()  # c10::cuda::CUDACachingAllocator::emptyCache(), :823
        return info

    def _rebind(self, p2, m2, v2, steps, stats2):
        """Make (p2, m2, v2, steps) this mapper's parameters and Adam state (the reference keeps step / exp_avg / exp_avg_sq
        per group across its surgery, src/gaussian_model.cpp:683-699) and rebuild the per-P buffers."""
        if self.dp is not None:  # new symmetric buffers for the new P; the moments are re-sharded
            self._init_fused_dp(p2, (m2, v2), steps[PARAM_ORDER[0]])
            self.stats = stats2
            self._fbuf = None
            return
        self._release_state()  # the old tensors go back to the allocator before the new gradient / work buffers are made
        self.params = {k: torch.nn.Parameter(p2[k]) for k in PARAM_ORDER}
        groups = [dict(params=[self.params[k]], lr=self._lrs[k], name=k) for k in PARAM_ORDER]
        self.optimizer = (self._optimizer_factory or (lambda g: FusedAdam(g, lr=0.0, eps=1e-15)))(groups)
        for k in PARAM_ORDER:
            self.optimizer.state[self.params[k]] = dict(step=steps[k], exp_avg=m2[k], exp_avg_sq=v2[k])
        self.grads = FlatGrads(self.params)
        self.stats = stats2
        self._fbuf = None  # per-P work buffers

    def _current_state(self):
        """(params, exp_avg, exp_avg_sq, steps) as dicts of full tensors (the moment shards gathered under dp_mode='fused')."""
        p = {k: v.data for k, v in self.params.items()}
        if self.dp is not None:
            m, v, step = self._fused_dp_moments()
            return p, m, v, {k: step for k in PARAM_ORDER}
        m, v, steps = self._optimizer_moments(p)
        return p, m, v, steps

    def increase_pcd(self, new_points, new_colors, iteration: int):
        """GaussianModel::increasePcd (reference src/gaussian_model.cpp:297-384; called per new keyframe,
        src/gaussian_mapper.cpp:873,974,1482): append Gaussians for `new_points` [n,3] / `new_colors` [n,3] with zero Adam
        moments (leg_slam_b200.densify.increase_pcd).  With several ranks every rank passes the same points."""
        from . import densify as densify_mod
        if self.stats is None:
            raise ValueError("construct the Mapper with track_densify_stats=True")
        if int(new_points.shape[0]) == 0:
            return 0
        p, m, v, steps = self._current_state()
        with torch.no_grad():
            p2, m2, v2, stats2 = densify_mod.increase_pcd(p, m, v, self.stats, new_points, new_colors, iteration, self.sh_degree)
        del p, m, v
        self._rebind(p2, m2, v2, steps, stats2)
        return int(new_points.shape[0])

    def reset_opacity(self):
        """GaussianModel::resetOpacity (:567-595; cadence src/gaussian_mapper.cpp:757-761): opacity through the reference's
        (no-op) clamp, its Adam moments zeroed, step count kept."""
        from . import densify as densify_mod
        p, m, v, steps = self._current_state()
        p, m, v = dict(p), dict(m), dict(v)
        with torch.no_grad():
            densify_mod.reset_opacity(p, m, v)
        if self.dp is not None:
            self._rebind(p, m, v, steps, self.stats)
            return
        self.params["opacity"].data.copy_(p["opacity"])
        st = self.optimizer.state.get(self.params["opacity"])
        if st:
            st["exp_avg"].zero_()
            st["exp_avg_sq"].zero_()

    # -- per-iteration settings the reference applies at the top of trainForOneIteration (src/gaussian_mapper.cpp:662-683) ----
    def _set_lr(self, name, lr):
        self._lrs[name] = float(lr)  # also what a later densify / load rebuilds the optimizer with
        i = PARAM_ORDER.index(name)
        if self.dp is not None:
            self.dp.lrs[i] = float(lr)
        else:
            self.optimizer.param_groups[i]["lr"] = float(lr)

    def learning_rate(self, name) -> float:
        return self._lrs[name]

    def set_position_lr_schedule(self, position_lr_init, position_lr_final, position_lr_delay_mult=0.01,
                                 position_lr_max_steps=30_000, spatial_lr_scale=1.0):
        """The xyz part of GaussianModel::trainingSetup (reference src/gaussian_model.cpp:493,513-517): start at
        position_lr_init * spatial_lr_scale and remember the decay's constants (defaults: include/gaussian_parameters.h:60-63;
        lr_delay_steps_ stays 0 in the reference, so the delay factor is inactive)."""
        self.spatial_lr_scale = float(spatial_lr_scale)
        self.xyz_lr_schedule = ExponLr(float(position_lr_init) * self.spatial_lr_scale, float(position_lr_final) * self.spatial_lr_scale,
                                       position_lr_delay_mult, position_lr_max_steps)
        self._set_lr("xyz", float(position_lr_init) * self.spatial_lr_scale)

    def update_learning_rate(self, step: int) -> float:
        """GaussianModel::updateLearningRate (:520-530): the xyz group's learning rate for this iteration."""
        if self.xyz_lr_schedule is None:
            raise ValueError("call set_position_lr_schedule first")
        lr = self.xyz_lr_schedule(int(step))
        self._set_lr("xyz", lr)
        return lr

    def set_position_learning_rate(self, position_lr):   # :542-544
        self._set_lr("xyz", float(position_lr) * getattr(self, "spatial_lr_scale", 1.0))

    def set_feature_learning_rate(self, feature_lr):     # :546-549 -- the rest coefficients run at a twentieth
        self._set_lr("features_dc", feature_lr)
        self._set_lr("features_rest", float(feature_lr) / 20.0)

    def set_language_feature_learning_rate(self, lr):    # :551-553
        self._set_lr("lang_feat", lr)

    def set_opacity_learning_rate(self, lr):             # :555-557
        self._set_lr("opacity", lr)

    def set_scaling_learning_rate(self, lr):             # :559-561
        self._set_lr("scaling", lr)

    def set_rotation_learning_rate(self, lr):            # :563-565
        self._set_lr("rotation", lr)

    def one_up_sh_degree(self):                          # GaussianModel::oneUpShDegree, :100-103
        if self.sh_degree < self.max_sh_degree:
            self.sh_degree += 1

    def set_sh_degree(self, sh: int):                    # GaussianModel::setShDegree, :105-107
        self.sh_degree = min(int(sh), self.max_sh_degree)

    def density_control(self, iteration: int, p: DensityControlParams, cameras_extent: float, generator=None):
        """The density-control block of trainForOneIteration for iteration `iteration` (reference
        src/gaussian_mapper.cpp:737-761), after `train_step` (which has already accumulated the view's statistics when the
        mapper tracks them): densifyAndPrune on the reference's cadence with its size threshold, then resetOpacity on its own.
        As in the reference this runs BEFORE the iteration's optimizer step would (there the step then finds no gradients on
        the rebuilt tensors); here `train_step` has stepped already, so call it between iterations.  Returns the actions
        taken (`density_control_actions`) with the densification's `info` under "info"."""
        act = density_control_actions(int(iteration), p)
        if act["densify"]:
            act["info"] = self.densify_and_prune(p.densify_grad_threshold, p.densify_min_opacity, cameras_extent,
                                                 act["size_threshold"], generator=generator)
        if act["reset_opacity"]:
            self.reset_opacity()
        return act

    def apply_scaled_transformation(self, s: float, T):
        """GaussianModel::applyScaledTransformation (reference src/gaussian_model.cpp:387-405; the map-wide correction after
        a scale / pose change of the tracker): xyz <- T (s * xyz) with `T` [4,4] stored transposed like the reference's pose
        tensors, and -- as shipped -- the LOG-scale parameter multiplied by s (:403), then xyz and scaling go back into the
        optimizer with zeroed moments and kept step counts (scaledTransformationPostfix, :407-420)."""
        from . import ingest
        p, m, v, steps = self._current_state()
        p, m, v = dict(p), dict(m), dict(v)
        with torch.no_grad():
            p["xyz"] = ingest.transformPoints(p["xyz"].detach() * float(s), T)
            p["scaling"] = p["scaling"].detach() * float(s)
            for k in ("xyz", "scaling"):
                m[k], v[k] = torch.zeros_like(p[k]), torch.zeros_like(p[k])
        if self.dp is not None:
            self._rebind(p, m, v, steps, self.stats)
            return
        for k in ("xyz", "scaling"):
            self.params[k].data.copy_(p[k])
            st = self.optimizer.state.get(self.params[k])
            if st:
                st["exp_avg"].zero_()
                st["exp_avg_sq"].zero_()

    def scaled_transform_visible_points_of_keyframe(self, point_not_transformed_flags, diff_pose, kf_world_view_transform,
                                                    kf_full_proj_transform, kf_creation_iter: int,
                                                    stable_num_iter_existence: int, num_transformed: int = 0, scale: float = 1.0):
        """GaussianModel::scaledTransformVisiblePointsOfKeyframe (reference src/gaussian_model.cpp:422-481; called under the
        render mutex when a loop closure moved a keyframe, src/gaussian_mapper.cpp:924-945): the Gaussians that keyframe sees,
        that are younger than `stable_num_iter_existence` iterations relative to it and not yet transformed, are moved by
        `diff_pose` (one fused kernel, leg_slam_b200.ingest), then xyz and rotation are put back into the optimizer with zeroed
        Adam moments and their step counts kept (replaceTensorToOptimizer, :577-595).  As in the reference the rotation
        parameter becomes the ACTIVATED (normalised) rotation for every Gaussian, and corrected rows carry the shipped
        (w, x, z, 0) layout.  `point_not_transformed_flags` [P] bool is updated in place; returns the updated counter."""
        from . import ingest
        if self.stats is None:
            raise ValueError("construct the Mapper with track_densify_stats=True (exist_since_iter_ lives with the statistics)")
        p, m, v, steps = self._current_state()
        p, m, v = dict(p), dict(m), dict(v)
        with torch.no_grad():
            points = p["xyz"].detach().clone()
            rots = torch.nn.functional.normalize(p["rotation"].detach())  # getRotationActivation, gaussian_model.cpp:50-52
            unstable = torch.abs(self.stats.exist_since_iter - int(kf_creation_iter)) < int(stable_num_iter_existence)
            n = ingest.scaleAndTransformThenMarkVisiblePoints(points, rots, point_not_transformed_flags, unstable, diff_pose,
                                                              kf_world_view_transform, kf_full_proj_transform,
                                                              num_transformed, scale)
            p["xyz"], p["rotation"] = points, rots
            for k in ("xyz", "rotation"):
                m[k], v[k] = torch.zeros_like(p[k]), torch.zeros_like(p[k])
        if self.dp is not None:
            self._rebind(p, m, v, steps, self.stats)
            return n
        for k in ("xyz", "rotation"):
            self.params[k].data.copy_(p[k])
            st = self.optimizer.state.get(self.params[k])
            if st:
                st["exp_avg"].zero_()
                st["exp_avg_sq"].zero_()
        return n

    def train_step(self, window: Sequence[Keyframe], presharded: bool = False):
        """Render + back-propagate this rank's share of `window`, sum gradients over ranks, Adam.
        `window` is the iteration's global list of keyframes (every rank passes the same list and
        takes its round-robin share) unless `presharded`, in which case it already holds only this
        rank's keyframes.  Returns the (detached) sum of this rank's view losses.
        With dp_mode="fused" the language-feature part of the exchange may still be running on its side stream when this
        returns; the next render forward waits for it by itself (lgs_stream_hooks), any OTHER reader of
        `params["lang_feat"]` on the caller's stream calls `self.dp.flush()` first (checkpoints and densification do)."""
        mine = list(range(len(window))) if presharded else shard_views(len(window), self.rank, self.world_size)
        self.last_num_views = len(mine)
        if self.dp is None:
            self.grads.attach(self.params)
        if self.fused and len(mine) > 0:
            with torch.no_grad():
                total = self._train_views_fused(window, mine)  # overwrites the flat buffer: no memset needed
                if self.dp is not None:
                    # reduce-scatter + Adam + all-gather in peer-memory kernels; with one local view the language-feature
                    # gradient was written in place and is final when the render backward ends (exchange starts there)
                    self.dp.step(late_ready_at_hook=len(mine) == 1)
                    return total
                if self.world_size > 1:
                    dist.all_reduce(self.grads.flat, op=dist.ReduceOp.SUM, group=self.pg)
            self.optimizer.step()
            return total
        if self.dp is not None:  # a rank without a view this iteration still takes part in the exchange
            self.grads.zero_()
            self.dp.step(late_ready_at_hook=False)
            return torch.zeros((), device=self.grads.flat.device)
        self.grads.zero_()
        total = None
        for i in mine:
            kf = window[i]
            a = self.activated()
            image, lf, depth, _radii = self.render_fn(kf.camera, a)
            if self.use_cuda_graph and image.is_cuda:
                loss, gi, gl, gd = self._graphed_loss(image, lf, depth, kf.gt_image, kf.gt_lf, kf.gt_depth, kf.mask)
                torch.autograd.backward([image, lf, depth], [gi, gl, gd])
                loss = loss.clone()
            else:
                loss = self._loss_from_images(image, lf, depth, kf.gt_image, kf.gt_lf, kf.gt_depth, kf.mask)
                loss.backward()  # accumulates into the flat buffer through the .grad views
                loss = loss.detach()
            total = loss if total is None else total + loss
        if self.world_size > 1:
            dist.all_reduce(self.grads.flat, op=dist.ReduceOp.SUM, group=self.pg)
        self.optimizer.step()
        return total if total is not None else torch.zeros((), device=self.grads.flat.device)
