"""GaussianRenderer::render (reference src/gaussian_renderer.cpp:24-160) mirrored by leg_slam_b200.renderer: the four input
selections (SHs / override colours / SH->RGB on the host side; scales + rotations / precomputed covariance) against the
unmodified reference rasterizer fed with the same tensors, the 6-tuple it returns, and gradients reaching the raw parameters
and the screen-space leaf."""
import numpy as np
import pytest
import torch

import cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _setup(dev):
    from leg_slam_b200 import synthetic
    from leg_slam_b200.renderer import GaussianModelView, KeyframeView
    W, H, P = 96, 64, 3000
    raw = synthetic.make_scene(P, seed=91, mean_scale=0.06, device=dev)
    raw["rotation"] = raw["rotation"] * 1.3  # un-normalised storage, as during training
    cam = synthetic.make_cameras(1, W, H, seed=91)[0].to(dev)
    return raw, cam, GaussianModelView(raw, sh_degree=3), KeyframeView(cam), W, H, P


def _ref_forward(ref_mod, bg, a, cam, H, W, include_lf, colors=None, cov=None, degree=3):
    e = torch.empty(0, device=bg.device)
    return ref_mod.rasterize_gaussians(bg, a["means3D"], e if colors is None else colors, a["lang_feats"] if include_lf else e,
                                       a["opacities"], e if cov is not None else a["scales"], e if cov is not None else a["rotations"],
                                       1.0, e if cov is None else cov, cam.viewmatrix, cam.projmatrix, cam.tanfovx, cam.tanfovy, H, W,
                                       a["shs"] if colors is None else e, degree, cam.campos, False, include_lf)


def test_render_paths_match_reference_rasterizer(dev, ref_mod):
    from leg_slam_b200 import synthetic
    from leg_slam_b200.renderer import GaussianPipelineParams, GaussianRenderer, eval_sh
    raw, cam, pc, kf, W, H, P = _setup(dev)
    a = synthetic.activate(raw)
    bg = torch.tensor([0.1, 0.2, 0.3], device=dev)
    n = lambda t: t.detach().cpu().numpy()  # noqa: E731
    with torch.no_grad():
        # 1. SHs + scales / rotations + language features (the mapping configuration)
        img, lf, depth, ssp, vis, radii = GaussianRenderer.render(kf, H, W, pc, GaussianPipelineParams(), bg, None, 1.0, False, True)
        Rr, cr, lr, dr, radr, *_ = _ref_forward(ref_mod, bg, a, cam, H, W, True)
        assert torch.equal(radii, radr) and torch.equal(vis, radr > 0) and ssp.shape == (P, 3) and not ssp.any()
        for x, y in ((img, cr), (lf, lr), (depth, dr)):
            assert cases.rel_err(n(x), n(y)) <= 1e-4
        # 2. override colours, no language features: rendered_lf is the reference's zero image (rasterize_points.cu:71)
        oc = torch.rand(P, 3, device=dev)
        img2, lf2, depth2, _s, vis2, radii2 = GaussianRenderer.render(kf, H, W, pc, GaussianPipelineParams(), bg, oc, 1.0, True, False)
        _R, cr2, lr2, dr2, radr2, *_ = _ref_forward(ref_mod, bg, a, cam, H, W, False, colors=oc)
        assert torch.equal(radii2, radr2) and not lf2.any() and lf2.shape == (64, H, W)
        assert cases.rel_err(n(img2), n(cr2)) <= 1e-4 and cases.rel_err(n(depth2), n(dr2)) <= 1e-4
        # 3. precomputed 3D covariance (pipe.compute_cov3D_)
        img3, _l3, depth3, _s3, _v3, radii3 = GaussianRenderer.render(kf, H, W, pc, GaussianPipelineParams(False, True), bg, None, 1.0, False, False)
        _R, cr3, _lr3, dr3, radr3, *_ = _ref_forward(ref_mod, bg, a, cam, H, W, False, cov=pc.getCovarianceActivation())
        assert torch.equal(radii3, radr3)
        assert cases.rel_err(n(img3), n(cr3)) <= 1e-4 and cases.rel_err(n(depth3), n(dr3)) <= 1e-4
        assert cases.rel_err(n(img3), n(_ref_forward(ref_mod, bg, a, cam, H, W, False)[1])) <= 2e-3  # ~ the in-kernel covariance
        # 4. SH -> RGB on the host side (pipe.convert_SHs_): equals the in-kernel conversion
        img4, *_r4 = GaussianRenderer.render(kf, H, W, pc, GaussianPipelineParams(True, False), bg, None, 1.0, False, False)
        assert cases.rel_err(n(img4), n(_ref_forward(ref_mod, bg, a, cam, H, W, False)[1])) <= 1e-4
        d = a["means3D"] - cam.campos[None]
        d = d / d.norm(dim=1, keepdim=True)
        rgb = torch.clamp_min(eval_sh(3, a["shs"].transpose(1, 2), d) + 0.5, 0.0)
        _R, cr4, *_ = _ref_forward(ref_mod, bg, a, cam, H, W, False, colors=rgb)
        assert cases.rel_err(n(img4), n(cr4)) <= 1e-5


def test_render_backward_reaches_raw_parameters(dev, ref_mod):
    """Gradients through render(): raw parameters (through the activations) and the screen-space leaf, against the reference
    rasterizer's backward chained through the same torch activations."""
    from leg_slam_b200 import synthetic
    from leg_slam_b200.renderer import GaussianModelView, GaussianPipelineParams, GaussianRenderer, KeyframeView
    raw, cam, _pc, kf, W, H, P = _setup(dev)
    g = torch.Generator().manual_seed(92)
    up = [torch.randn(c, H, W, generator=g).to(dev) / (H * W) for c in (3, 64, 1)]
    bg = torch.zeros(3, device=dev)
    leaves = {k: v.clone().requires_grad_(True) for k, v in raw.items()}
    img, lf, depth, ssp, vis, radii = GaussianRenderer.render(KeyframeView(cam), H, W, GaussianModelView(leaves), GaussianPipelineParams(), bg,
                                                              None, 1.0, False, True)
    ((img * up[0]).sum() + (lf * up[1]).sum() + (depth * up[2]).sum()).backward()
    assert ssp.grad is not None and ssp.grad.shape == (P, 3) and bool((ssp.grad[vis].abs().sum(1) > 0).any()) and not ssp.grad[~vis].any()
    # reference: its forward + backward on the activated tensors, chained through the activations by autograd
    ref_leaves = {k: v.clone().requires_grad_(True) for k, v in raw.items()}
    a = synthetic.activate(ref_leaves)
    e = torch.empty(0, device=dev)
    Rr, cr, lr, dr, radr, gr, br, ir = ref_mod.rasterize_gaussians(bg, a["means3D"].detach(), e, a["lang_feats"].detach(), a["opacities"].detach(),
                                                                  a["scales"].detach(), a["rotations"].detach(), 1.0, e, cam.viewmatrix,
                                                                  cam.projmatrix, cam.tanfovx, cam.tanfovy, H, W, a["shs"].detach(), 3,
                                                                  cam.campos, False, True)
    (dm2, _dc, dlf, dop, dm3, _dcov, dsh, dsc, drot) = ref_mod.rasterize_gaussians_backward(
        bg, a["means3D"].detach(), radr, e, a["lang_feats"].detach(), a["scales"].detach(), a["rotations"].detach(), 1.0, e, cam.viewmatrix,
        cam.projmatrix, cam.tanfovx, cam.tanfovy, up[0], up[1], up[2], a["shs"].detach(), 3, cam.campos, gr, Rr, br, ir, True)
    torch.autograd.backward([a["means3D"], a["lang_feats"], a["opacities"], a["scales"], a["rotations"], a["shs"]],
                            [dm3, dlf, dop, dsc, drot, dsh])
    assert cases.rel_err(ssp.grad.cpu().numpy(), dm2.cpu().numpy()) <= 1e-3
    for k in raw:
        assert cases.rel_err(leaves[k].grad.cpu().numpy(), ref_leaves[k].grad.cpu().numpy()) <= 1e-3, k
