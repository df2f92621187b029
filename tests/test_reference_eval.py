"""The reference's Python evaluation package (route B of INTEGRATION.md), imported UNMODIFIED from /root/reference/eval in this
container and run on CPU: its checkpoint reader (eval/gaussian_model.py:58-111 -- SURVEY.md 8f row 3 names it as the reader
of the .ply format), its camera (eval/utils.py MiniCam) and its render() (eval/render.py) over its own rasterizer wrapper.

What stands in for what this container lacks (none of it computes a compared value):
  * `plyfile` is not installed: a minimal generic reader of binary little-endian PLY (header -> numpy structured array) is
    registered under that name -- element / property access by NAME, as plyfile offers it;
  * `simple_knn._C`, `clip`, `cv2`, `torchvision` are imported by those modules and unused on this path: empty modules;
  * there is no GPU: `device="cuda"` arguments and `.cuda()` calls land on the CPU;
  * the wrapper's compiled `_C` is tests/oracle_l1.py (the CPU oracle behind L1, recording every call).
The reference tree does not travel to the GPU box, where this file skips."""
import importlib.util
import os
import sys
import types

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "oracle"))
import cases  # noqa: E402,F401  (sys.path)
import oracle_l1  # noqa: E402
import ply_ref  # noqa: E402
from leg_slam_b200 import rasterizer as RZ, renderer as RD, synthetic  # noqa: E402

REF_EVAL = "/root/reference/eval"
REF_PY_PKG = os.path.join(REF_EVAL, "submodules/diff-gaussian-rasterization-legs-slam/diff_gaussian_rasterization_legs_slam")
pytestmark = pytest.mark.skipif(not os.path.isdir(REF_PY_PKG), reason="reference tree not present (it does not travel to the GPU box)")
NAMES = ("xyz", "features_dc", "features_rest", "lang_feat", "opacity", "scaling", "rotation")


class _Prop:
    def __init__(self, name):
        self.name = name


class _Element:
    def __init__(self, data):
        self.data = data
        self.properties = [_Prop(n) for n in data.dtype.names]

    def __getitem__(self, name):
        return self.data[name]


class _PlyData:
    """What the reference uses of plyfile.PlyData: read(path).elements[0] with ["name"] and .properties[i].name."""
    TYPES = {"float": "<f4", "float32": "<f4", "double": "<f8", "uchar": "u1", "uint8": "u1", "int": "<i4", "int32": "<i4"}

    def __init__(self, elements):
        self.elements = elements

    @classmethod
    def read(cls, path):
        with open(path, "rb") as f:
            assert f.readline().strip() == b"ply"
            fields, count, in_vertex = [], 0, False
            while True:
                tok = f.readline().decode().split()
                if tok[0] == "format":
                    assert tok[1] == "binary_little_endian"
                elif tok[0] == "element":
                    in_vertex = tok[1] == "vertex"
                    count = int(tok[2]) if in_vertex else count
                elif tok[0] == "property" and in_vertex:
                    fields.append((tok[2], cls.TYPES[tok[1]]))
                elif tok[0] == "end_header":
                    break
            return cls([_Element(np.fromfile(f, dtype=np.dtype(fields), count=count))])


@pytest.fixture()
def ref_eval(monkeypatch):
    """(gaussian_model module, utils module, render module, the wrapper's recording L1) of the reference, on CPU."""
    fake = {"plyfile": dict(PlyData=_PlyData), "simple_knn": {}, "simple_knn._C": dict(distCUDA2=None), "clip": {}, "cv2": {},
            "torchvision": {}, "torchvision.transforms": {}}
    for name, attrs in fake.items():
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__dict__.update(attrs)
            monkeypatch.setitem(sys.modules, name, m)
    on_cpu = lambda kw: {k: v for k, v in kw.items() if k != "device"}  # noqa: E731
    real_tensor, real_zeros_like = torch.tensor, torch.zeros_like
    monkeypatch.setattr(torch, "tensor", lambda *a, **kw: real_tensor(*a, **on_cpu(kw)))
    monkeypatch.setattr(torch, "zeros_like", lambda *a, **kw: real_zeros_like(*a, **on_cpu(kw)))
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    l1 = oracle_l1.RecordingL1()
    c = types.ModuleType("diff_gaussian_rasterization_legs_slam._C")
    c.rasterize_gaussians, c.rasterize_gaussians_backward, c.mark_visible = (
        l1.rasterize_gaussians, l1.rasterize_gaussians_backward, l1.mark_visible)
    monkeypatch.setitem(sys.modules, "diff_gaussian_rasterization_legs_slam._C", c)
    spec = importlib.util.spec_from_file_location("diff_gaussian_rasterization_legs_slam", os.path.join(REF_PY_PKG, "__init__.py"),
                                                  submodule_search_locations=[REF_PY_PKG])
    wrap = importlib.util.module_from_spec(spec)
    monkeypatch.setitem(sys.modules, "diff_gaussian_rasterization_legs_slam", wrap)
    spec.loader.exec_module(wrap)
    mods = {}
    for name in ("gaussian_model", "utils", "render"):     # eval/ scripts import each other by these top-level names
        spec = importlib.util.spec_from_file_location(name, os.path.join(REF_EVAL, name + ".py"))
        m = importlib.util.module_from_spec(spec)
        monkeypatch.setitem(sys.modules, name, m)
        spec.loader.exec_module(m)
        mods[name] = m
    return mods["gaussian_model"], mods["utils"], mods["render"], l1


def test_reference_python_reader_reads_the_checkpoint_format(ref_eval, tmp_path):
    """eval/gaussian_model.py load_ply on the file oracle/ply_ref.py writes (byte-identical to GaussianModel::savePly's and to
    ply_io.save_ply's: tests/test_reference_model.py, tests/test_ply_io.py): the seven tensors come back with the shapes and
    values that went in -- language features included -- and equal what oracle/ply_ref.py read_ply (the restatement of this
    reader) returns; the model's activated accessors equal the mapper's activations."""
    gm, _, _, _ = ref_eval
    P = 300
    sc = synthetic.make_scene(P, seed=9)
    path = str(tmp_path / "point_cloud.ply")
    ply_ref.write_ply(path, *[sc[k].numpy() for k in NAMES])
    m = gm.GaussianModel(sh_degree=3)
    m.load_ply(path)
    got = dict(xyz=m._xyz, features_dc=m._features_dc, features_rest=m._features_rest, lang_feat=m._language_features,
               opacity=m._opacity, scaling=m._scaling, rotation=m._rotation)
    back = ply_ref.read_ply(path, max_sh_degree=3)
    for k in NAMES:
        assert got[k].shape == sc[k].shape and got[k].dtype == torch.float32, k
        assert torch.equal(got[k].detach(), sc[k]), k
        assert np.array_equal(np.asarray(back[k], np.float32).reshape(sc[k].shape), sc[k].numpy()), k
    assert m.active_sh_degree == 3
    view = RD.GaussianModelView({k: sc[k] for k in NAMES})
    assert torch.equal(m.scaling, view.getScalingActivation()) and torch.equal(m.rotation, view.getRotationActivation())
    assert torch.equal(m.opacity, view.getOpacityActivation()) and torch.equal(m.features, view.getFeatures())
    assert torch.equal(m.language_features, view.getLanguageFeatures())


@pytest.mark.parametrize("use_override", [False, True])
def test_reference_python_render_makes_the_same_calls(ref_eval, monkeypatch, tmp_path, use_override):
    """eval/render.py render() -- the eval scripts' renderer, with an override colour for the query's heat map
    (eval/find_objects_gaussians.py) -- from a checkpoint it loaded itself and its own MiniCam, against
    leg_slam_b200.renderer.GaussianRenderer.render on the synthetic camera: the same L1 call, the same six results."""
    import oracle as O
    gm, utils, render, la = ref_eval
    lb = oracle_l1.RecordingL1()
    monkeypatch.setattr(RZ, "_C", lb)
    n = O.num_threads()
    O.lib().omp_set_num_threads(1)
    try:
        P, W, H = 300, 48, 32
        sc = synthetic.make_scene(P, seed=5, mean_scale=0.08)
        path = str(tmp_path / "point_cloud.ply")
        ply_ref.write_ply(path, *[sc[k].numpy() for k in NAMES])
        pc = gm.GaussianModel(sh_degree=3)
        pc.load_ply(path)
        g = torch.Generator().manual_seed(3)
        c = torch.tensor([2.0, 1.5, 1.4], dtype=torch.float64)
        R = synthetic.look_at(tuple(c.tolist()), (4.0, 3.0, 1.2))
        fx = fy = W / 2.0
        cam = synthetic.camera_from_pose(R, c, W, H, fx, fy)
        mini = utils.MiniCam(W, H, utils.focal2fov(fx, W), utils.focal2fov(fy, H), utils.get_world2view(R.numpy(), c.numpy()))
        # one camera under both renderers, so that the calls can be compared bit for bit (the two constructions agree to 1e-6:
        # tests/test_host_logic.py::test_synthetic_cameras_equal_the_reference_minicam)
        kv = RD.KeyframeView(cam)
        kv.FoVx_, kv.FoVy_ = mini.FoVx, mini.FoVy
        kv.world_view_transform_, kv.full_proj_transform_, kv.camera_center_ = (mini.world_view_transform, mini.full_proj_transform,
                                                                                mini.camera_center)
        bg = torch.zeros(3)
        heat = torch.rand(P, 3, generator=g) if use_override else None
        with torch.no_grad():
            r = render.render(mini, pc, bg, override_color=heat)
            o = RD.GaussianRenderer.render(kv, H, W, RD.GaussianModelView({k: sc[k] for k in NAMES}, 3, pc.active_sh_degree),
                                           RD.GaussianPipelineParams(), bg, heat, 1.0, use_override, True)
        assert oracle_l1.same_calls(la.calls, lb.calls) and [x[0] for x in la.calls] == ["rasterize_gaussians"]
        for key, ours in zip(("rendered_image", "rendered_lf", "rendered_depth", "viewspace_points", "visibility_filter", "radii"), o):
            assert torch.equal(r[key], ours), key
        assert 0 < int(r["visibility_filter"].sum()) < P and r["rendered_lf"].abs().sum() > 0
    finally:
        O.lib().omp_set_num_threads(n)


def test_semantic_query_lines_of_the_reference_equal_the_oracle(monkeypatch):
    """The semantic query as the reference computes it -- the statements of eval/find_objects_gaussians.py between "Compute
    similarity between each Gaussian and text embedding" and the relevance normalisation (:164-175), and the per-pixel line
    (:323) -- cut out of the file as they stand and executed on CPU tensors, against the C oracle's cosine (what the CUDA
    kernels are held to at 2e-6, tests/test_gpu_parity.py::test_cosine_query) and the relevance formula."""
    import textwrap
    import oracle as O
    import torch.nn.functional as F
    src = open(os.path.join(REF_EVAL, "find_objects_gaussians.py")).read().splitlines()
    a = next(i for i, ln in enumerate(src) if "Compute similarity between each Gaussian and text embedding" in ln)
    b = next(i for i, ln in enumerate(src) if ln.strip().startswith("similarities = 1 - ("))
    block = textwrap.dedent("\n".join(src[a:b + 1]))
    assert "F.normalize(gaussian_features, dim=1)" in block and "torch.matmul" in block
    pix = next(ln.strip() for ln in src if ln.strip().startswith("dist = F.cosine_similarity(rendered_lf, text_embed, dim=0)"))
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *x, **k: self)
    feats, text = cases.cosine_case()                    # [P,64], [Q,64]
    ns = dict(F=F, torch=torch, gaussian_features=feats, text_emb_compressed=text[0:1].clone())   # the first request (:167)
    exec(block, ns)
    sim0 = O.cosine(feats.numpy(), text[0:1].numpy())[:, 0].astype(np.float64)
    want = 1 - (sim0 - sim0.min()) / (sim0.max() - sim0.min())
    assert ns["similarities"].shape == (feats.shape[0], 1)    # text_embed[:, 0] of the [64,1,1] embedding is [64,1]
    assert float(np.abs(ns["similarities"].numpy()[:, 0] - want).max()) <= 5e-6
    # per pixel: text_embed is [64,1,1] after the block above (:167), rendered_lf a [64,H,W] feature image
    H, W = 24, 40
    lf_img = feats[:H * W].t().reshape(64, H, W).contiguous()
    ns2 = dict(F=F, rendered_lf=lf_img, text_embed=ns["text_embed"])
    exec(pix.replace(".detach()", ""), ns2)
    want_pix = O.cosine(feats[:H * W].numpy(), text[0:1].numpy())[:, 0].reshape(H, W)
    assert ns2["dist"].shape == (H, W) and float(np.abs(ns2["dist"].numpy() - want_pix).max()) <= 2e-6
