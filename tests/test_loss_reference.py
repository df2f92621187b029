"""The mapping loss against the UNMODIFIED reference loss (include/loss_utils.h compiled as oracle/_ref/ref_loss.so by
oracle/build_ref.py, chained as src/gaussian_mapper.cpp:707-721): pins SURVEY.md 8f row 2.
  CPU : leg_slam_b200/loss.py (the torch statement the reference arm and the autograd path run) == reference, value and gradients;
  GPU : the fused lgs_mapping_loss kernels == reference on CUDA tensors, value and all three image gradients."""
import pytest
import torch

import cases


@pytest.fixture(scope="module")
def ref_loss():
    import build_ref
    try:
        return build_ref.load_loss()
    except FileNotFoundError as e:
        pytest.skip(str(e))


def _inputs(H, W, seed, dev="cpu", with_mask=True):
    g = torch.Generator().manual_seed(seed)
    t = dict(image=torch.rand(3, H, W, generator=g), lf=torch.randn(64, H, W, generator=g), depth=torch.rand(1, H, W, generator=g) * 3,
             gt_image=torch.rand(3, H, W, generator=g), gt_lf=torch.randn(64, 37, 37, generator=g), gt_depth=torch.rand(1, H, W, generator=g) * 3)
    t["mask"] = ((torch.rand(1, H, W, generator=g) > 0.15).float() if with_mask else torch.ones(1, H, W)).expand(3, H, W).contiguous()
    return {k: v.to(dev) for k, v in t.items()}


def _ref_value_and_grads(ref_loss, t, lam):
    im, lf, d = (t[k].clone().requires_grad_(True) for k in ("image", "lf", "depth"))
    loss = ref_loss.mapping_loss(im, lf, d, t["gt_image"], t["gt_lf"], t["gt_depth"], t["mask"], lam)
    gi, gl, gd = torch.autograd.grad(loss, [im, lf, d])
    return loss.detach(), gi, gl, gd


@pytest.mark.parametrize("H,W,with_mask", [(48, 64, True), (61, 93, False)])
def test_torch_statement_equals_reference_loss_cpu(ref_loss, H, W, with_mask):
    from leg_slam_b200 import loss as loss_mod
    t = _inputs(H, W, 5, "cpu", with_mask)
    lam = 0.2
    ref, gi, gl, gd = _ref_value_and_grads(ref_loss, t, lam)
    im, lf, d = (t[k].clone().requires_grad_(True) for k in ("image", "lf", "depth"))
    up = torch.nn.functional.interpolate(t["gt_lf"].unsqueeze(0), size=(H, W)).squeeze(0)
    ours = loss_mod.mapping_loss(im * t["mask"], lf * t["mask"][0:1], d * t["mask"][0:1], t["gt_image"], up, t["gt_depth"], lam, faithful_sign=True)
    oi, ol, od = torch.autograd.grad(ours, [im, lf, d])
    assert abs(float(ours) - float(ref)) <= 1e-6 * max(1.0, abs(float(ref)))
    for a, b in ((oi, gi), (ol, gl), (od, gd)):
        assert cases.rel_err(a.numpy(), b.numpy()) <= 1e-5
    # the pieces, one by one
    x, y = t["image"], t["gt_image"]
    assert abs(float(loss_mod.l1_loss(x, y)) - float(ref_loss.l1_loss(x, y))) <= 1e-7
    assert abs(float(loss_mod.ssim(x, y)) - float(ref_loss.ssim(x, y))) <= 1e-6
    assert abs(float(loss_mod.psnr(x, y)) - float(ref_loss.psnr(x, y))) <= 1e-5
    assert abs(float(loss_mod.cosine_similarity(t["lf"], up)) - float(ref_loss.cosine_similarity(t["lf"].clone(), up.clone()))) <= 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("H,W,with_mask", [(64, 96, True), (61, 93, False), (480, 640, True)])
def test_fused_loss_equals_reference_loss_gpu(ref_loss, H, W, with_mask):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from leg_slam_b200.fused import FusedMappingLoss
    dev = torch.device("cuda:0")
    t = _inputs(H, W, 6, dev, with_mask)
    ref, gi, gl, gd = _ref_value_and_grads(ref_loss, t, 0.2)
    out, fi, fl, fd = FusedMappingLoss(lambda_dssim=0.2, faithful_sign=True)(t["image"], t["lf"], t["depth"], t["gt_image"], t["gt_lf"],
                                                                            t["gt_depth"], t["mask"] if with_mask else None)
    assert abs(float(out[0]) - float(ref)) <= 2e-5 * max(1.0, abs(float(ref)))
    assert cases.rel_err(fi.cpu().numpy(), gi.cpu().numpy()) <= 1e-4
    assert cases.rel_err(fl.cpu().numpy(), gl.cpu().numpy()) <= 1e-4
    assert cases.rel_err(fd.cpu().numpy(), gd.cpu().numpy()) <= 1e-5
