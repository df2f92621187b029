"""The reference's own helper headers, compiled unmodified (include/general_utils.h, include/sh_utils.h ->
oracle/_ref/ref_utils.so, oracle/ref_utils_wrap.cpp), against every restatement of them in this repo:
  inverse_sigmoid, build_rotation  -> oracle/densify_ref.py, leg_slam_b200.renderer.build_rotation
  eval_sh                          -> leg_slam_b200.renderer.eval_sh (GaussianRenderer::render, convert_SHs path)
  RGB2SH / SH2RGB                  -> leg_slam_b200.densify.increase_pcd's DC coefficient, the 0.5 offset of the SH -> RGB step
"""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


@pytest.fixture(scope="module")
def ref_utils():
    import build_ref
    try:
        if os.path.isdir(build_ref.LOSS_REF_INC):
            build_ref.build_utils(verbose=False)
        return build_ref.load_utils()
    except FileNotFoundError as e:
        pytest.skip(str(e))


def test_inverse_sigmoid_restatement_is_bit_identical_cpu(ref_utils):
    import densify_ref as DR
    g = torch.Generator().manual_seed(3)
    x = torch.cat([torch.rand(4096, 1, generator=g), torch.tensor([[0.01], [0.1], [0.5], [0.99], [1e-6]])])
    assert torch.equal(DR.inverse_sigmoid(x), ref_utils.inverse_sigmoid(x))
    # resetOpacity's argument (gaussian_model.cpp:567-575): min(sigmoid(opacity), 0.01)
    y = torch.minimum(torch.sigmoid(torch.randn(1000, 1, generator=g) * 4), torch.full((1000, 1), 0.01))
    assert torch.equal(DR.inverse_sigmoid(y), ref_utils.inverse_sigmoid(y))


def test_sh_constants_and_rgb2sh_cpu(ref_utils):
    from leg_slam_b200 import densify, renderer, synthetic
    g = torch.Generator().manual_seed(4)
    rgb = torch.rand(1000, 3, generator=g)
    ref = ref_utils.RGB2SH(rgb)
    for c0 in (densify.SH_C0, renderer.SH_C0, synthetic.SH_C0):
        assert torch.equal((rgb - 0.5) / c0, ref)          # the expression increase_pcd uses for features_dc
    assert ref_utils.SH2RGB(0.3) == pytest.approx(float(np.float32(0.3) * np.float32(renderer.SH_C0) + np.float32(0.5)), abs=1e-7)


@pytest.mark.parametrize("deg", [0, 1, 2, 3])
def test_eval_sh_matches_the_reference_header_cpu(ref_utils, deg):
    from leg_slam_b200 import renderer
    g = torch.Generator().manual_seed(10 + deg)
    P = 2000
    sh = torch.randn(P, 3, 16, generator=g)
    dirs = torch.nn.functional.normalize(torch.randn(P, 3, generator=g), dim=1)
    ref = ref_utils.eval_sh(deg, sh.clone(), dirs.clone())
    ours = renderer.eval_sh(deg, sh, dirs)
    assert ours.shape == ref.shape == (P, 3)
    # same polynomial, term order may differ by an ulp or two of the partial sums
    assert float((ours - ref).abs().max()) <= 4e-6 * max(1.0, float(ref.abs().max()))


@pytest.mark.gpu
def test_build_rotation_restatements_match_the_reference_header(ref_utils):
    """general_utils::build_rotation allocates its result on the GPU (general_utils.h:43), so this one needs a device."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import densify_ref as DR
    from leg_slam_b200 import renderer
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(5)
    q = (torch.randn(5000, 4, generator=g) * torch.rand(5000, 1, generator=g) * 3).to(dev)
    ref = ref_utils.build_rotation(q.clone())
    assert torch.equal(DR.build_rotation(q), ref)          # the density-control restatement: same ops, same order
    assert float((renderer.build_rotation(q) - ref).abs().max()) <= 1e-6
    # proper rotations
    eye = torch.eye(3, device=dev).expand(5000, 3, 3)
    assert float((ref @ ref.transpose(1, 2) - eye).abs().max()) <= 1e-5


REF_SH_PY = "/root/reference/eval/sh_utils.py"


@pytest.mark.skipif(not os.path.isfile(REF_SH_PY), reason="reference tree not present")
def test_eval_sh_equals_the_reference_python_module():
    """leg_slam_b200.renderer.eval_sh against the eval package's own eval_sh (reference eval/sh_utils.py:58-113), imported
    unmodified: bit-identical at every degree; RGB2SH / SH2RGB of that module against the constant the package uses."""
    import importlib.util
    from leg_slam_b200 import renderer as RD
    spec = importlib.util.spec_from_file_location("ref_eval_sh_utils", REF_SH_PY)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    g = torch.Generator().manual_seed(0)
    sh = torch.randn(500, 3, 16, generator=g)
    d = torch.nn.functional.normalize(torch.randn(500, 3, generator=g))
    for deg in range(4):
        assert torch.equal(ref.eval_sh(deg, sh, d), RD.eval_sh(deg, sh, d)), deg
    x = torch.rand(100, 3, generator=g)
    assert torch.equal(ref.RGB2SH(x), (x - 0.5) / RD.SH_C0) and torch.equal(ref.SH2RGB(x), x * RD.SH_C0 + 0.5)
