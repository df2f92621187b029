"""GPU parity tests: the CUDA path (through the reference-shaped API on the C ABI) against
(1) the CPU oracle on the same seeded inputs, (2) the golden fixtures produced by the
UNMODIFIED reference (tests/golden/), (3) the compiled reference itself when oracle/_ref was
shipped.  Gates: keys / sorted order / tile ranges / radii bit-exact;
images <= 1e-4 relative; gradients <= 1e-3 relative."""
import os

import numpy as np
import pytest
import torch

import cases
from conftest import golden

pytestmark = pytest.mark.gpu

IMG_TOL = 1e-4
GRAD_TOL = 1e-3


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def run_ours(cs):
    from leg_slam_b200 import rasterize_points as rp, debug
    R, color, lf, depth, radii, geom, binning, img = rp.rasterize_gaussians(*cases.fwd_args(cs))
    grads = rp.rasterize_gaussians_backward(*cases.bwd_args(cs, radii, geom, R, binning, img))
    torch.cuda.synchronize()
    P, W, H = cs["P"], cs["W"], cs["H"]
    n = lambda t: t.detach().cpu().numpy()  # noqa: E731
    # the reference's key arrays, re-expressed from what the production path (32-bit keys, own radix sort, run repair) computed
    gv, bv, iv = debug.geom_view(geom, P), debug.reference_keys(geom, binning, img, P, R, W, H), debug.image_view(img, W, H)
    out = dict(num_rendered=R, radii=n(radii), out_color=n(color), out_lf=n(lf), out_depth=n(depth),
               records=n(gv["records"]), cov3D=n(gv["cov3D"]), tiles_touched=n(gv["tiles_touched"]),
               keys_unsorted=n(bv["keys_unsorted"]), values_unsorted=n(bv["values_unsorted"]),
               keys_sorted=n(bv["keys_sorted"]), point_list=n(bv["point_list"]), ranges=n(iv["ranges"]),
               n_contrib=n(iv["n_contrib"]), final_T=n(iv["final_T"]))
    for name, g in zip(cases.GRAD_NAMES, grads):
        out[name] = n(g)
    return out


@pytest.mark.parametrize("name", list(cases.CASES))
def test_forward_backward_vs_oracle(name, dev, oracle_mod):
    cs = cases.make_case(name, dev)
    ours = run_ours(cs)
    f = cases.oracle_forward(cases.make_case(name), oracle_mod)
    g = cases.oracle_backward(cases.make_case(name), f, oracle_mod)
    vis = f["radii"] > 0
    # ---- integer / bit-exact gates
    assert ours["num_rendered"] == f["num_rendered"]
    np.testing.assert_array_equal(ours["radii"], f["radii"])
    np.testing.assert_array_equal(ours["tiles_touched"].view(np.uint32), f["tiles_touched"])
    np.testing.assert_array_equal(ours["records"][vis, 2].view(np.uint32), f["depths"][vis].view(np.uint32))
    np.testing.assert_array_equal(ours["records"][vis, 0:2].copy().view(np.uint32), f["means2D"][vis].view(np.uint32))
    np.testing.assert_array_equal(ours["keys_unsorted"].view(np.uint64), np.asarray(f["keys_unsorted"]).view(np.uint64))
    np.testing.assert_array_equal(ours["values_unsorted"].view(np.uint32), np.asarray(f["values_unsorted"]).view(np.uint32))
    np.testing.assert_array_equal(ours["keys_sorted"].view(np.uint64), f["keys_sorted"])
    np.testing.assert_array_equal(ours["point_list"].view(np.uint32), f["point_list"])
    np.testing.assert_array_equal(ours["ranges"].view(np.uint32), f["ranges"])
    # conic: same float sequence on both sides (division / fma are IEEE on CPU and GPU)
    np.testing.assert_array_equal(ours["records"][vis, 4:8].copy().view(np.uint32), f["conic_opacity"][vis].view(np.uint32))
    # ---- images (expf differs by ulps between libm and CUDA: tolerance, and the per-pixel
    #      contributor count may differ for a vanishing fraction of threshold-straddling fragments)
    assert cases.rel_err(ours["out_color"], f["out_color"]) <= IMG_TOL
    assert cases.rel_err(ours["out_depth"], f["out_depth"]) <= IMG_TOL
    if cs["include_lf"]:
        assert cases.rel_err(ours["out_lf"], f["out_lf"]) <= IMG_TOL
    else:
        assert not ours["out_lf"].any()  # allocated as zeros, like the reference (rasterize_points.cu:71)
    assert (ours["n_contrib"].view(np.uint32) != f["n_contrib"]).mean() <= 1e-3
    # ---- gradients
    for gname in cases.GRAD_NAMES:
        assert ours[gname].shape == g[gname].shape, gname
        assert cases.rel_err(ours[gname], g[gname]) <= GRAD_TOL, gname


@pytest.mark.parametrize("name", list(cases.CASES))
def test_vs_reference_golden(name, dev):
    gd = golden(name)
    cs = cases.make_case(name, dev)
    ours = run_ours(cs)
    vis = gd["visible"]
    assert ours["num_rendered"] == int(gd["num_rendered"])
    for k in ("radii", "tiles_touched", "keys_sorted", "point_list", "ranges", "n_contrib"):
        np.testing.assert_array_equal(ours[k].view(gd[k].dtype), gd[k], err_msg=k)
    np.testing.assert_array_equal(ours["keys_unsorted"].view(np.uint64), gd["keys_unsorted"].view(np.uint64))
    np.testing.assert_array_equal(ours["values_unsorted"].view(np.uint32), gd["values_unsorted"].view(np.uint32))
    np.testing.assert_array_equal(ours["records"][vis, 2].view(np.uint32), gd["depths"][vis].view(np.uint32))
    np.testing.assert_array_equal(ours["records"][vis, 0:2].copy().view(np.uint32), gd["means2D"][vis].view(np.uint32))
    np.testing.assert_array_equal(ours["final_T"].view(np.uint32), gd["final_T"].view(np.uint32))
    assert cases.rel_err(ours["out_color"], gd["out_color"]) <= IMG_TOL
    assert cases.rel_err(ours["out_depth"], gd["out_depth"]) <= IMG_TOL
    assert cases.rel_err(ours["out_lf"][cases.LF_GOLDEN_CH], gd["out_lf_sub"]) <= IMG_TOL
    for gname in cases.GRAD_NAMES:
        o = ours[gname][:, cases.LF_GOLDEN_CH] if gname == "dL_dlang_feats" else ours[gname]
        assert cases.rel_err(o, gd[gname]) <= GRAD_TOL, gname


def _compare_with_reference_live(cs, ref_mod, oracle_mod=None):
    """Same call, same tensors, reference .so vs ours, side by side on this GPU: point_list, ranges, radii, n_contrib, final_T
    and both 64-bit key arrays bit-equal; images <= 1e-4; the 9 gradients <= 1e-3."""
    import refbuf
    from leg_slam_b200 import rasterize_points as rp, debug
    P, W, H = cs["P"], cs["W"], cs["H"]
    Rr, cr, lr, dr, radr, gr, br, ir = ref_mod.rasterize_gaussians(*cases.fwd_args(cs))
    Ro, co, lo, do, rado, go, bo, io = rp.rasterize_gaussians(*cases.fwd_args(cs))
    assert Rr == Ro
    assert torch.equal(radr, rado)
    rb, ob = refbuf.ref_binning_view(br, Rr), debug.reference_keys(go, bo, io, P, Ro, W, H)
    assert torch.equal(rb["point_list"], ob["point_list"])
    assert torch.equal(rb["keys_sorted"], ob["keys_sorted"])
    assert torch.equal(rb["keys_unsorted"], ob["keys_unsorted"]) and torch.equal(rb["point_list_unsorted"], ob["values_unsorted"])
    ri, oi = refbuf.ref_image_view(ir, W, H), debug.image_view(io, W, H)
    assert torch.equal(ri["ranges"], oi["ranges"])
    assert torch.equal(ri["n_contrib"], oi["n_contrib"])
    assert torch.equal(ri["final_T"].view(torch.int32), oi["final_T"].view(torch.int32))
    n = lambda t: t.cpu().numpy()  # noqa: E731
    assert cases.rel_err(n(co), n(cr)) <= IMG_TOL and cases.rel_err(n(lo), n(lr)) <= IMG_TOL
    assert cases.rel_err(n(do), n(dr)) <= IMG_TOL
    gref = ref_mod.rasterize_gaussians_backward(*cases.bwd_args(cs, radr, gr, Rr, br, ir))
    gour = rp.rasterize_gaussians_backward(*cases.bwd_args(cs, rado, go, Ro, bo, io))
    for gname, a, b in zip(cases.GRAD_NAMES, gour, gref):
        assert cases.rel_err(n(a), n(b)) <= GRAD_TOL, gname
    return dict(R=Ro, color=co, lf=lo, depth=do, radii=rado, n_contrib=oi["n_contrib"], grads=gour)


@pytest.mark.parametrize("name", ["sh3_lf", "dense_opaque"])
def test_vs_compiled_reference_live(name, dev, ref_mod):
    _compare_with_reference_live(cases.make_case(name, dev), ref_mod)


@pytest.mark.parametrize("name", list(cases.BASELINE_CASES))
def test_vs_compiled_reference_live_at_baseline_sizes(name, dev, ref_mod):
    """BASELINE.json configs at their stated sizes against the unmodified reference on the same GPU: cfgA (10 k, 320x240),
    cfgB (500 k, 640x480), a ragged cfgB (637x475: partial edge tiles) and cfgD (2 M, 1296x968: 19 602 tiles, 15-bit tile ids,
    R ~ 3 M) -- the production binning path (32-bit keys, own 3-pass radix sort, run repair) and both blend kernels."""
    _compare_with_reference_live(cases.make_baseline_case(name, dev), ref_mod)


def test_cfgA_vs_cpu_oracle(dev, oracle_mod):
    """BASELINE.json configs[0]: 10 k Gaussians, RGB + depth + 64-D feature forward / backward at 320x240, one view, against the
    CPU golden (the C oracle)."""
    cs = cases.make_baseline_case("cfgA", dev)
    cpu = cases.make_baseline_case("cfgA")
    ours = run_ours(cs)
    f = cases.oracle_forward(cpu, oracle_mod)
    g = cases.oracle_backward(cpu, f, oracle_mod)
    assert ours["num_rendered"] == f["num_rendered"]
    np.testing.assert_array_equal(ours["radii"], f["radii"])
    np.testing.assert_array_equal(ours["keys_sorted"].view(np.uint64), f["keys_sorted"])
    np.testing.assert_array_equal(ours["point_list"].view(np.uint32), f["point_list"])
    np.testing.assert_array_equal(ours["ranges"].view(np.uint32), f["ranges"])
    for k in ("out_color", "out_depth", "out_lf"):
        assert cases.rel_err(ours[k], f[k]) <= IMG_TOL, k
    assert (ours["n_contrib"].view(np.uint32) != f["n_contrib"]).mean() <= 1e-3
    for gname in cases.GRAD_NAMES:
        assert cases.rel_err(ours[gname], g[gname]) <= GRAD_TOL, gname


def test_quantised_depths_long_runs_of_equal_keys(dev, ref_mod):
    """Thousands of Gaussians at EXACTLY the same depth in one tile (what increasePcd from a quantised depth image of a
    fronto-parallel wall produces when re-rendered from the same keyframe): the 32-bit sort keys are all equal inside a tile,
    so the order comes entirely from the run repair (warp-cooperative for runs > 32).  Lists bit-equal to the reference."""
    cs = cases.make_case("sh3_lf", dev)
    P = cs["P"]
    # put every Gaussian on three planes of constant view-space depth: depth = z_view, camera looks along view[:,2]
    V = cs["viewmatrix"]  # column-major W2C: p_view = p_world @ V[:3,:3] + V[3,:3]
    pv = cs["means3D"] @ V[:3, :3] + V[3, :3]
    planes = torch.tensor([1.5, 2.25, 3.0], device=dev)
    pv[:, 2] = planes[torch.arange(P, device=dev) % 3]
    cs["means3D"] = ((pv - V[3, :3]) @ torch.linalg.inv(V[:3, :3])).contiguous()
    out = _compare_with_reference_live(cs, ref_mod)
    from leg_slam_b200 import debug
    assert out["R"] > 0


@pytest.mark.parametrize("n", [1, 63, 64, 65, 128, 129, 257, 700])
def test_translucent_stack_hits_the_channel_kernels_batch_boundaries(n, dev, ref_mod):
    """n translucent Gaussians stacked over the same few tiles, none terminating the pixel: every (tile, half) work item of
    the backward's channel kernel then holds about n half-records -- below, at and just above its 64-record batches, up to a
    dozen batches in one item -- and a 24 x 16 image gives fewer work items than resident CTAs.  Whole forward + backward
    against the compiled reference."""
    cs = cases.make_case("sh3_lf", dev)
    g = torch.Generator().manual_seed(900 + n)
    V = cs["viewmatrix"]  # p_view = p_world @ V[:3,:3] + V[3,:3]
    W, H = 24, 16
    pv = torch.zeros(n, 3)
    pv[:, 0] = (torch.rand(n, generator=g) - 0.5) * 0.6          # spread over the middle tiles (fx = W / 2 = 12 px per unit at z = 1)
    pv[:, 1] = (torch.rand(n, generator=g) - 0.5) * 0.4
    pv[:, 2] = 2.0 + torch.rand(n, generator=g)                  # depths 2 .. 3
    pv = pv.to(dev)
    cs.update(P=n, W=W, H=H, means3D=((pv - V[3, :3]) @ torch.linalg.inv(V[:3, :3])).contiguous(),
              opacities=torch.full((n, 1), 0.012, device=dev),  # alpha <= 0.012 >= 1/255 near the centre: 0.988^700 > 1e-4, nobody stops
              scales=(0.25 + 0.2 * torch.rand(n, 3, generator=g)).to(dev), rotations=torch.nn.functional.normalize(
                  torch.randn(n, 4, generator=g), dim=1).to(dev),
              shs=(torch.randn(n, 16, 3, generator=g) * 0.3).to(dev), lang_feats=torch.randn(n, 64, generator=g).to(dev),
              dL_dcolor=(torch.randn(3, H, W, generator=g) / (H * W)).to(dev), dL_dlf=(torch.randn(64, H, W, generator=g) / (H * W)).to(dev),
              dL_ddepth=(torch.randn(1, H, W, generator=g) / (H * W)).to(dev))
    # same intrinsics family as the case (FoV 90 degrees): only the image size changes
    out = _compare_with_reference_live(cs, ref_mod)
    assert out["R"] >= n  # every Gaussian lands in at least one tile
    assert int(out["n_contrib"].max()) >= min(n, 60)


# ---------------------------------------------------------------------------- edge cases
def test_empty_and_all_culled(dev):
    from leg_slam_b200 import rasterize_points as rp
    cs = cases.make_case("sh3_lf", dev)
    # P == 0: zero-filled outputs, nothing rendered (rasterize_points.cu:85: `if (P != 0)`)
    e = dict(cs)
    for k in ("means3D", "opacities", "lang_feats", "shs", "scales", "rotations"):
        e[k] = cs[k][:0]
    R, color, lf, depth, radii, *_ = rp.rasterize_gaussians(*cases.fwd_args(e))
    assert R == 0 and radii.numel() == 0 and not color.any() and not lf.any() and not depth.any()
    # everything behind the camera: R == 0, the image is the background, gradients are all zero
    b = dict(cs)
    b["means3D"] = cs["means3D"] - 1000.0 * cs["viewmatrix"][:3, 2]  # push along -view_z
    R, color, lf, depth, radii, geom, binning, img = rp.rasterize_gaussians(*cases.fwd_args(b))
    assert R == 0 and not radii.any()
    assert torch.allclose(color, cs["bg"][:, None, None].expand_as(color))
    assert not lf.any() and not depth.any()
    grads = rp.rasterize_gaussians_backward(*cases.bwd_args(b, radii, geom, R, binning, img))
    assert all(not g.any() for g in grads)


def test_single_gaussian_and_huge_radius(dev, oracle_mod):
    """One Gaussian covering the whole image: its rectangle is clamped to the tile grid."""
    from leg_slam_b200 import rasterize_points as rp
    cs = cases.make_case("sh3_lf", dev)
    cpu = cases.make_case("sh3_lf")
    for d in (cs, cpu):
        idx = int((cpu["means3D"] @ cpu["viewmatrix"][:3, 2] + cpu["viewmatrix"][3, 2]).argmax())  # farthest in front
        for k in ("means3D", "opacities", "lang_feats", "shs", "scales", "rotations"):
            d[k] = d[k][idx:idx + 1].contiguous()
        d["scales"] = d["scales"] * 0 + 5.0
        d["P"] = 1
    R, color, lf, depth, radii, *_ = rp.rasterize_gaussians(*cases.fwd_args(cs))
    f = cases.oracle_forward(cpu, oracle_mod)
    assert R == f["num_rendered"] == ((cs["W"] + 7) // 8) * ((cs["H"] + 7) // 8)
    assert int(radii[0]) == int(f["radii"][0])
    assert cases.rel_err(color.cpu().numpy(), f["out_color"]) <= IMG_TOL


def test_radii_pointer_optional_and_mark_visible(dev, oracle_mod):
    from leg_slam_b200 import rasterize_points as rp
    cs = cases.make_case("ragged_sh1", dev)
    vis = rp.mark_visible(cs["means3D"], cs["viewmatrix"], cs["projmatrix"])
    ref = oracle_mod.mark_visible(cs["means3D"].cpu().numpy(), cs["viewmatrix"].cpu().numpy())
    assert vis.dtype == torch.bool and np.array_equal(vis.cpu().numpy(), ref)
    assert rp.mark_visible(cs["means3D"][:0], cs["viewmatrix"], cs["projmatrix"]).numel() == 0


def test_no_writes_outside_caller_buffers(dev):
    """compute-sanitizer is closed on this pool, so: every buffer handed to the C ABI sits between two 64 KB
    guard regions filled with a pattern; after forward + backward the guards must be untouched, and buffers
    sized exactly by lgs_*_bytes must suffice."""
    import ctypes
    from leg_slam_b200 import _lib
    L = _lib.lib()
    G = 1 << 16
    held = []

    def guarded(nbytes, dtype=torch.uint8):
        esz = torch.empty((), dtype=dtype).element_size()
        raw = torch.full((G + ((nbytes + 255) // 256) * 256 + G,), 0xA5, dtype=torch.uint8, device=dev)
        held.append((raw, nbytes))
        return raw[G:G + nbytes].view(dtype) if nbytes % esz == 0 else raw[G:G + nbytes]

    for name in ("ragged_sh1", "dense_opaque"):
        held.clear()
        cs = cases.make_case(name, dev)
        P, W, H, M_ = cs["P"], cs["W"], cs["H"], 16
        s = torch.cuda.current_stream(dev).cuda_stream
        geom, img = guarded(L.lgs_geom_bytes(P)), guarded(L.lgs_image_bytes(W, H))
        radii = guarded(4 * P, torch.int32)
        color, lf, depth = guarded(12 * H * W, torch.float32), guarded(256 * H * W, torch.float32), guarded(4 * H * W, torch.float32)
        R = ctypes.c_int(0)
        t = {k: cs[k].contiguous() for k in ("means3D", "shs", "opacities", "scales", "rotations", "viewmatrix", "projmatrix",
                                              "campos", "bg", "lang_feats", "dL_dcolor", "dL_dlf", "dL_ddepth")}
        _lib.check(L.lgs_forward_stage1(P, cs["degree"], M_, W, H, t["means3D"].data_ptr(), t["shs"].data_ptr(), None,
                                        t["opacities"].data_ptr(), t["scales"].data_ptr(), 1.0, t["rotations"].data_ptr(), None,
                                        t["viewmatrix"].data_ptr(), t["projmatrix"].data_ptr(), t["campos"].data_ptr(),
                                        cs["tanfovx"], cs["tanfovy"], 0, geom.data_ptr(), radii.data_ptr(), ctypes.byref(R), s), "s1")
        binning = guarded(L.lgs_binning_bytes(R.value))
        _lib.check(L.lgs_forward_stage2(P, W, H, R.value, t["bg"].data_ptr(), t["lang_feats"].data_ptr(), geom.data_ptr(),
                                        binning.data_ptr(), img.data_ptr(), color.data_ptr(), lf.data_ptr(), depth.data_ptr(), 1, s), "s2")
        scratch = guarded(L.lgs_backward_scratch_bytes(R.value, W, H))
        g = {k: guarded(4 * n, torch.float32) for k, n in (("m2d", 3 * P), ("conic", 4 * P), ("op", P), ("col", 3 * P), ("lf", 64 * P),
                                                            ("dep", P), ("m3d", 3 * P), ("cov", 6 * P), ("sh", 48 * P), ("sc", 3 * P),
                                                            ("rot", 4 * P))}
        _lib.check(L.lgs_backward(P, cs["degree"], M_, R.value, W, H, t["bg"].data_ptr(), t["means3D"].data_ptr(), t["shs"].data_ptr(),
                                  None, t["lang_feats"].data_ptr(), t["scales"].data_ptr(), 1.0, t["rotations"].data_ptr(), None,
                                  t["viewmatrix"].data_ptr(), t["projmatrix"].data_ptr(), t["campos"].data_ptr(), cs["tanfovx"],
                                  cs["tanfovy"], radii.data_ptr(), geom.data_ptr(), binning.data_ptr(), img.data_ptr(),
                                  t["dL_dcolor"].data_ptr(), t["dL_dlf"].data_ptr(), t["dL_ddepth"].data_ptr(), g["m2d"].data_ptr(),
                                  g["conic"].data_ptr(), g["op"].data_ptr(), g["col"].data_ptr(), g["lf"].data_ptr(), g["dep"].data_ptr(),
                                  g["m3d"].data_ptr(), g["cov"].data_ptr(), g["sh"].data_ptr(), g["sc"].data_ptr(), g["rot"].data_ptr(),
                                  1, 1, scratch.data_ptr(), s), "bwd")
        torch.cuda.synchronize()
        for raw, nbytes in held:
            assert bool((raw[:G] == 0xA5).all()), "write before a buffer"
            assert bool((raw[G + ((nbytes + 255) // 256) * 256:] == 0xA5).all()), "write past a buffer"
        assert bool(torch.isfinite(g["lf"]).all() and torch.isfinite(color).all())


# ---------------------------------------------------------------------------- full-size properties
@pytest.fixture(scope="module")
def cfgB(dev):
    """BASELINE.json configs[1]: 500k Gaussians, 640x480."""
    from leg_slam_b200 import synthetic, rasterize_points as rp
    sc = synthetic.make_scene(500_000, seed=2, device=dev)
    cam = synthetic.make_cameras(1, 640, 480, seed=2)[0].to(dev)
    a = synthetic.activate(sc)
    bg = torch.zeros(3, device=dev)
    e = torch.empty(0, device=dev)
    args = (bg, a["means3D"], e, a["lang_feats"], a["opacities"], a["scales"], a["rotations"], 1.0, e, cam.viewmatrix,
            cam.projmatrix, cam.tanfovx, cam.tanfovy, 480, 640, a["shs"], 3, cam.campos, False, True)
    out = rp.rasterize_gaussians(*args)
    return dict(a=a, cam=cam, bg=bg, args=args, out=out)


def test_fullsize_binning_properties(cfgB, dev):
    from leg_slam_b200 import debug
    R, color, lf, depth, radii, geom, binning, img = cfgB["out"]
    P, W, H = 500_000, 640, 480
    gv, bv, iv = debug.geom_view(geom, P), debug.reference_keys(geom, binning, img, P, R, W, H), debug.image_view(img, W, H)
    tt = gv["tiles_touched"].long()
    assert int(tt.sum()) == R and torch.equal(bv["point_offsets"].long(), tt.cumsum(0))
    assert torch.equal(tt > 0, radii > 0)
    ks = bv["keys_sorted"]
    assert bool((ks[1:] >= ks[:-1]).all())  # sorted (keys are < 2^63, signed compare is fine)
    # the production 32-bit keys are sorted too, and agree with the 64-bit ones on the tile
    k32 = debug.binning_view(binning, R)["keys_sorted32"].long() & 0xffffffff
    assert bool((k32[1:] >= k32[:-1]).all())
    # the sorted list is a permutation of the emitted list: checksum of checksums
    assert int(ks.sum()) == int(bv["keys_unsorted"].sum())
    assert int(bv["point_list"].long().sum()) == int(bv["values_unsorted"].long().sum())
    # equal keys are in ascending Gaussian index (what the reference's stable sort of a Gaussian-major emission gives)
    same = ks[1:] == ks[:-1]
    pl = bv["point_list"].long()
    assert bool((pl[1:][same] > pl[:-1][same]).all())
    # ranges tile the list: every tile's range holds exactly its tile id, lengths add up to R
    rg = iv["ranges"].long()
    lens = rg[:, 1] - rg[:, 0]
    assert int(lens.sum()) == R
    tile_of = (ks >> 32)
    nz = lens > 0
    assert torch.equal(tile_of[rg[nz, 0]], torch.nonzero(nz).flatten())
    assert torch.equal(tile_of[rg[nz, 1] - 1], torch.nonzero(nz).flatten())
    # depth bits in keys equal the stored depths of the listed Gaussians
    assert torch.equal((ks & 0xffffffff).int(), gv["records"][pl, 2].contiguous().view(torch.int32))
    assert bool((iv["n_contrib"].view(H, W).long() <= lens.view(H // 8, W // 8).repeat_interleave(8, 0).repeat_interleave(8, 1)).all())


def test_forward_without_readback_gives_the_same_frame(cfgB, dev):
    """lgs_forward_stage1 with num_rendered_host == NULL leaves R on the device; stage2 and the backward then get the CAPACITY of
    the binning buffer.  Same lists, same images, same gradients as the synchronous call; lgs_forward_status reports R; a
    capacity below R sets the overflow flag and stays inside the buffers."""
    from leg_slam_b200 import rasterize_points as rp, debug
    R, color, lf, depth, radii, geom, binning, img = cfgB["out"]
    pl = debug.binning_view(binning, R)["point_list"].clone()
    iv = debug.image_view(img, 640, 480)
    rg, nc = iv["ranges"].clone(), iv["n_contrib"].clone()
    cap = int(R * 1.37) + 999
    R2, c2, l2, d2, radii2, geom2, binning2, img2 = rp.rasterize_gaussians(*cfgB["args"], capacity=cap)
    assert R2 == cap
    st = rp.forward_status(geom2, 500_000)
    torch.cuda.synchronize()
    assert int(st[0]) == R and int(st[2]) == 0 and int(st[3]) == 0
    assert torch.equal(radii2, radii) and torch.equal(c2, color) and torch.equal(l2, lf) and torch.equal(d2, depth)
    assert torch.equal(debug.binning_view(binning2, cap)["point_list"][:R], pl)
    iv2 = debug.image_view(img2, 640, 480)
    assert torch.equal(iv2["ranges"], rg) and torch.equal(iv2["n_contrib"], nc)
    a, cam, bg = cfgB["a"], cfgB["cam"], cfgB["bg"]
    e = torch.empty(0, device=dev)
    g = torch.Generator(device="cpu").manual_seed(6)
    dc, dl, dd = ((torch.randn(c, 480, 640, generator=g) / 307200).to(dev) for c in (3, 64, 1))

    def bwd(rad, ge, Rx, bi, im):
        return rp.rasterize_gaussians_backward(bg, a["means3D"], rad, e, a["lang_feats"], a["scales"], a["rotations"], 1.0, e,
                                               cam.viewmatrix, cam.projmatrix, cam.tanfovx, cam.tanfovy, dc, dl, dd, a["shs"], 3,
                                               cam.campos, ge, Rx, bi, im, True)
    for name, x, y in zip(cases.GRAD_NAMES, bwd(radii, geom, R, binning, img), bwd(radii2, geom2, cap, binning2, img2)):
        assert cases.rel_err(y.cpu().numpy(), x.cpu().numpy()) <= 1e-4, name  # same kernels, atomics reorder
    # capacity too small: flagged, finite, nothing written outside (the guard test covers the buffers themselves)
    small = R // 2
    R3, c3, *_rest = rp.rasterize_gaussians(*cfgB["args"], capacity=small)
    st3 = rp.forward_status(_rest[3], 500_000)
    torch.cuda.synchronize()
    assert int(st3[0]) == R and int(st3[2]) == 1
    assert bool(torch.isfinite(c3).all())


def test_fullsize_forward_idempotent_backward_linear(cfgB, dev):
    from leg_slam_b200 import rasterize_points as rp
    R, color, lf, depth, radii, geom, binning, img = cfgB["out"]
    R2, color2, lf2, depth2, radii2, *_ = rp.rasterize_gaussians(*cfgB["args"])
    assert R2 == R and torch.equal(color, color2) and torch.equal(lf, lf2) and torch.equal(depth, depth2)
    assert torch.equal(radii, radii2)
    assert bool(torch.isfinite(color).all() and torch.isfinite(lf).all() and torch.isfinite(depth).all())
    a, cam, bg = cfgB["a"], cfgB["cam"], cfgB["bg"]
    e = torch.empty(0, device=dev)
    g = torch.Generator(device="cpu").manual_seed(5)
    dc = (torch.randn(3, 480, 640, generator=g) / 307200).to(dev)
    dl = (torch.randn(64, 480, 640, generator=g) / 307200).to(dev)
    dd = (torch.randn(1, 480, 640, generator=g) / 307200).to(dev)

    def bwd(s):
        return rp.rasterize_gaussians_backward(bg, a["means3D"], radii, e, a["lang_feats"], a["scales"], a["rotations"],
                                               1.0, e, cam.viewmatrix, cam.projmatrix, cam.tanfovx, cam.tanfovy,
                                               dc * s, dl * s, dd * s, a["shs"], 3, cam.campos, geom, R, binning, img, True)
    g1, g2 = bwd(1.0), bwd(-2.0)
    for name, x, y in zip(cases.GRAD_NAMES, g1, g2):
        assert bool(torch.isfinite(x).all()), name
        # linear in the upstream gradient (two runs differ by atomic summation order, which the
        # scale/rotation chain amplifies: same 1e-3 gate as against the reference)
        assert cases.rel_err((y / -2.0).cpu().numpy(), x.cpu().numpy()) <= GRAD_TOL, name
    # culled Gaussians receive exactly zero gradient
    inv = radii == 0
    for x in g1:
        assert not x[inv].any()


def test_autograd_wrapper_matches_direct_call(dev):
    from leg_slam_b200 import GaussianRasterizationSettings, GaussianRasterizer, rasterize_points as rp
    cs = cases.make_case("sh3_lf", dev)
    rs = GaussianRasterizationSettings(cs["H"], cs["W"], cs["tanfovx"], cs["tanfovy"], cs["bg"], 1.0, cs["viewmatrix"],
                                       cs["projmatrix"], cs["degree"], cs["campos"], False, True)
    leaves = {k: cs[k].clone().requires_grad_(True) for k in ("means3D", "opacities", "shs", "lang_feats", "scales", "rotations")}
    means2D = torch.zeros_like(cs["means3D"], requires_grad=True)
    color, lf, depth, radii = GaussianRasterizer(rs)(leaves["means3D"], means2D, leaves["opacities"], shs=leaves["shs"],
                                                     lang_feats=leaves["lang_feats"], scales=leaves["scales"],
                                                     rotations=leaves["rotations"])
    ((color * cs["dL_dcolor"]).sum() + (lf * cs["dL_dlf"]).sum() + (depth * cs["dL_ddepth"]).sum()).backward()
    R, c2, l2, d2, rad2, geom, binning, img = rp.rasterize_gaussians(*cases.fwd_args(cs))
    direct = dict(zip(cases.GRAD_NAMES, rp.rasterize_gaussians_backward(*cases.bwd_args(cs, rad2, geom, R, binning, img))))
    assert torch.equal(color, c2) and torch.equal(radii, rad2)
    pairs = [(leaves["means3D"].grad, direct["dL_dmeans3D"]), (means2D.grad, direct["dL_dmeans2D"]),
             (leaves["shs"].grad, direct["dL_dsh"]), (leaves["lang_feats"].grad, direct["dL_dlang_feats"]),
             (leaves["opacities"].grad, direct["dL_dopacity"]), (leaves["scales"].grad, direct["dL_dscales"]),
             (leaves["rotations"].grad, direct["dL_drotations"])]
    for a, b in pairs:
        assert cases.rel_err(a.cpu().numpy(), b.cpu().numpy()) <= 1e-4  # atomics reorder between runs


def test_split_sh_entry_points_match_concatenated(dev):
    """lgs_forward_stage1_split_sh / lgs_backward_split_sh (features_dc + features_rest read and written in
    place) vs the reference-layout calls on cat(features_dc, features_rest): same images, same gradients;
    accumulate_sh adds a second backward on top."""
    from leg_slam_b200 import rasterize_points as rp
    cs = cases.make_case("sh3_lf", dev)
    R, c1, l1, d1, rad1, geom, binning, img = rp.rasterize_gaussians(*cases.fwd_args(cs))
    g1 = dict(zip(cases.GRAD_NAMES, rp.rasterize_gaussians_backward(*cases.bwd_args(cs, rad1, geom, R, binning, img))))
    dc, rest = cs["shs"][:, :1, :].contiguous(), cs["shs"][:, 1:, :].contiguous()
    fa = list(cases.fwd_args(cs))
    fa[15] = dc
    R2, c2, l2, d2, rad2, geom2, binning2, img2 = rp.rasterize_gaussians(*fa, sh_rest=rest)
    assert R2 == R and torch.equal(rad1, rad2)
    for a, b in ((c1, c2), (l1, l2), (d1, d2)):
        assert cases.rel_err(a.cpu().numpy(), b.cpu().numpy()) <= 1e-6
    P = cs["means3D"].shape[0]
    out = rp.backward_outputs(P, 0, dev, True, False, True)
    out["dL_dfeatures_dc"], out["dL_dfeatures_rest"] = torch.full_like(dc, 7.0), torch.full_like(rest, 7.0)
    ba = list(cases.bwd_args(cs, rad2, geom2, R2, binning2, img2))
    ba[16] = dc
    rp.rasterize_gaussians_backward_into(out, *ba, sh_rest=rest)
    got = torch.cat([out["dL_dfeatures_dc"], out["dL_dfeatures_rest"]], dim=1)
    assert cases.rel_err(got.cpu().numpy(), g1["dL_dsh"].cpu().numpy()) <= 1e-4
    assert cases.rel_err(out["dL_dmeans3D"].cpu().numpy(), g1["dL_dmeans3D"].cpu().numpy()) <= 1e-4
    rp.rasterize_gaussians_backward_into(out, *ba, sh_rest=rest, accumulate_sh=True)
    got2 = torch.cat([out["dL_dfeatures_dc"], out["dL_dfeatures_rest"]], dim=1)
    assert cases.rel_err(got2.cpu().numpy(), 2 * g1["dL_dsh"].cpu().numpy()) <= 1e-4


# ---------------------------------------------------------------------------- mapper
def _mapper_window(dev, n_views=2):
    from leg_slam_b200 import mapper as M, synthetic
    W, H = 96, 64
    sc = synthetic.make_scene(4000, seed=51, mean_scale=0.06, device=dev)
    cams = synthetic.make_cameras(n_views, W, H, seed=51)
    g = torch.Generator().manual_seed(52)
    win = [M.Keyframe(c.to(dev), torch.rand(3, H, W, generator=g).to(dev), torch.randn(64, 37, 37, generator=g).to(dev),
                      (torch.rand(1, H, W, generator=g) * 3).to(dev)) for c in cams]
    return sc, win


def test_mapper_step_matches_reference_rasterizer_and_torch_adam(dev, ref_mod):
    """One mapping iteration (activations -> rasterize -> loss -> backward -> Adam) through our
    rasterizer + FusedAdam (+ CUDA-graphed loss) vs the compiled reference rasterizer + torch Adam
    behind the same mapper code."""
    import bench
    from leg_slam_b200 import mapper as M
    sc, win = _mapper_window(dev)
    ours = M.Mapper(sc, sh_degree=3)                                       # fused, autograd-free path
    eager = M.Mapper(sc, sh_degree=3, use_cuda_graph=False, fused=False)   # autograd + eager torch loss
    graphed = M.Mapper(sc, sh_degree=3, fused=False)                       # autograd + CUDA-graphed loss
    assert ours.fused and not eager.fused

    class _Shim(bench.E2EPath):  # borrow the reference autograd glue of the bench
        def __init__(self):
            self.mapper = None
    shim = _Shim()
    bench.SH_DEGREE = 3
    refm = M.Mapper(sc, sh_degree=3, use_cuda_graph=False,
                    optimizer_factory=lambda g: torch.optim.Adam(g, lr=0.0, eps=1e-15))
    shim.mapper = refm
    refm.render_fn = shim._ref_render_fn()
    for _ in range(3):
        lo, le, lg, lr = ours.train_step(win), eager.train_step(win), graphed.train_step(win), refm.train_step(win)
        for l_ in (lo, le, lg):
            assert abs(float(l_) - float(lr)) <= 1e-4 * abs(float(lr))
    for k in M.PARAM_ORDER:
        r = refm.params[k].detach().cpu().numpy()
        # Adam's sign-like first steps amplify gradient noise near zero gradients; compare updates to the lr scale
        step = 3 * M.DEFAULT_LRS[k]
        for m in (ours, eager, graphed):
            d = np.abs(m.params[k].detach().cpu().numpy() - r)
            assert (d > 0.05 * step).mean() <= 2e-3, (k, float(d.max()), step)


def test_mapper_sh_degree_and_lr_schedule_follow_reference(dev, ref_mod):
    """The per-iteration settings of trainForOneIteration (reference src/gaussian_mapper.cpp:662-683): the active SH degree
    grows from 0 (setShDegree / oneUpShDegree) and the xyz learning rate follows exponLrFunc; ours (fused path + FusedAdam)
    against the compiled reference rasterizer + torch.optim.Adam under the same schedule.  Coefficients above the active
    degree receive no gradient, so Adam leaves them exactly where they were."""
    import bench
    from leg_slam_b200 import mapper as M
    sc, win = _mapper_window(dev)

    class _Shim(bench.E2EPath):
        def __init__(self):
            self.mapper = None
    shim = _Shim()
    refm = M.Mapper(sc, sh_degree=3, use_cuda_graph=False, optimizer_factory=lambda g: torch.optim.Adam(g, lr=0.0, eps=1e-15))
    shim.mapper = refm
    refm.render_fn = shim._ref_render_fn()
    ours = M.Mapper(sc, sh_degree=3)
    assert ours.fused
    rest0 = sc["features_rest"].clone()
    xyz_lrs = []
    try:
        for m in (ours, refm):
            m.set_position_lr_schedule(3.2e-4, 3.2e-6, 0.01, 6, spatial_lr_scale=2.0)
            m.set_sh_degree(0)
        for it in range(6):
            if it > 0 and it % 2 == 0:
                ours.one_up_sh_degree(), refm.one_up_sh_degree()
            assert ours.sh_degree == it // 2
            bench.SH_DEGREE = refm.sh_degree  # the reference arm's autograd glue reads the degree from the bench module
            xyz_lrs.append(ours.update_learning_rate(it))
            assert refm.update_learning_rate(it) == xyz_lrs[-1]
            lo, lr = ours.train_step(win), refm.train_step(win)
            assert abs(float(lo) - float(lr)) <= 2e-4 * abs(float(lr)), (it, float(lo), float(lr))
            n_active = (ours.sh_degree + 1) ** 2 - 1  # rest coefficients the rasterizer reads
            for m in (ours, refm):
                assert torch.equal(m.params["features_rest"].detach()[:, n_active:], rest0[:, n_active:]), (it, n_active)
            if n_active:
                assert not torch.equal(ours.params["features_rest"].detach()[:, :n_active], rest0[:, :n_active])
    finally:
        bench.SH_DEGREE = 3
    assert xyz_lrs[0] > xyz_lrs[-1] > 6.4e-6 and abs(xyz_lrs[0] - 6.4e-4) < 1e-9
    for k in M.PARAM_ORDER:
        r = refm.params[k].detach().cpu().numpy()
        step = sum(xyz_lrs) if k == "xyz" else 6 * M.DEFAULT_LRS[k]
        d = np.abs(ours.params[k].detach().cpu().numpy() - r)
        assert (d > 0.05 * step).mean() <= 3e-3, (k, float(d.max()), step)


def test_psnr_after_n_iterations_matches_reference(dev, ref_mod):
    """north_star gate: PSNR after N mapping iterations within 0.05 dB of the reference trajectory.  Same seeded
    scene, same ground truth (rendered by the reference from a perturbed copy of the scene), same loss; ours =
    fused path + FusedAdam, reference = compiled reference rasterizer + eager torch loss + torch.optim.Adam."""
    import bench
    from leg_slam_b200 import loss as loss_mod, mapper as M, synthetic
    W, H, P, N_IT = 160, 120, 20000, int(os.environ.get("LGS_PSNR_ITERS", "200"))
    truth = synthetic.make_scene(P, seed=81, mean_scale=0.05, device=dev)
    cams = [c.to(dev) for c in synthetic.make_cameras(3, W, H, seed=81)]
    g = torch.Generator().manual_seed(82)
    start = {k: v.clone() for k, v in truth.items()}
    start["xyz"] = start["xyz"] + 0.01 * torch.randn(P, 3, generator=g).to(dev)
    start["features_dc"] = start["features_dc"] + 0.3 * torch.randn(P, 1, 3, generator=g).to(dev)
    start["lang_feat"] = start["lang_feat"] + 0.1 * torch.randn(P, 64, generator=g).to(dev)
    start["opacity"] = start["opacity"] - 0.5
    bench.SH_DEGREE = 3

    class _Shim(bench.E2EPath):
        def __init__(self):
            self.mapper = None
    shim = _Shim()
    refm = M.Mapper(start, sh_degree=3, use_cuda_graph=False, optimizer_factory=lambda gr: torch.optim.Adam(gr, lr=0.0, eps=1e-15))
    shim.mapper = refm
    refm.render_fn = shim._ref_render_fn()
    ours = M.Mapper(start, sh_degree=3)
    assert ours.fused
    # ground truth from the TRUE scene through the reference rasterizer
    gt_m = M.Mapper(truth, sh_degree=3, use_cuda_graph=False, optimizer_factory=lambda gr: torch.optim.Adam(gr, lr=0.0))
    shim2 = _Shim()
    shim2.mapper = gt_m
    render_ref = shim2._ref_render_fn()
    window = []
    with torch.no_grad():
        for c in cams:
            img, lf, dep, _ = render_ref(c, gt_m.activated())
            lf_low = torch.nn.functional.interpolate(lf.unsqueeze(0), size=(37, 37)).squeeze(0)
            window.append(M.Keyframe(c, img.clone(), lf_low.contiguous(), dep.clone()))

    def psnr_of(m):
        vals = []
        with torch.no_grad():
            for kf in window:
                img, _, _, _ = render_ref(kf.camera, m.activated())
                vals.append(float(loss_mod.psnr(img, kf.gt_image)))
        return sum(vals) / len(vals)

    p0 = psnr_of(ours)
    for it in range(N_IT):  # one keyframe per iteration, like the reference mapper
        kf = [window[it % len(window)]]
        ours.train_step(kf)
        refm.train_step(kf)
    po, pr = psnr_of(ours), psnr_of(refm)
    assert pr > p0 + 0.5, (p0, pr)          # the optimisation actually moved the image
    print(f"psnr gate: N={N_IT} start {p0:.3f} dB, ours {po:.3f} dB, reference {pr:.3f} dB")
    assert abs(po - pr) <= 0.05, (p0, po, pr)


def test_fused_loss_matches_torch_loss(dev):
    """lgs_mapping_loss vs the stock-torch statement of include/loss_utils.h + gaussian_mapper.cpp:707-721."""
    from leg_slam_b200 import loss as loss_mod
    from leg_slam_b200.fused import FusedMappingLoss
    g = torch.Generator().manual_seed(61)
    for (H, W, with_mask, faithful) in ((64, 96, False, True), (61, 93, True, False), (480, 640, False, True)):
        image = torch.rand(3, H, W, generator=g).to(dev)
        lf = torch.randn(64, H, W, generator=g).to(dev)
        depth = (torch.rand(1, H, W, generator=g) * 3).to(dev)
        gt_image = torch.rand(3, H, W, generator=g).to(dev)
        gt_lf = torch.randn(64, 37, 37, generator=g).to(dev)
        gt_depth = (torch.rand(1, H, W, generator=g) * 3).to(dev)
        mask = None
        if with_mask:
            mask = (torch.rand(1, H, W, generator=g) > 0.2).float().expand(3, H, W).contiguous().to(dev)
        lf[:, 3, 5] = 0.0  # a zero feature vector exercises the eps clamp of cosine_similarity
        im, l_, d_ = image.clone().requires_grad_(True), lf.clone().requires_grad_(True), depth.clone().requires_grad_(True)
        up = torch.nn.functional.interpolate(gt_lf.unsqueeze(0), size=(H, W)).squeeze(0)
        a, b, c = (im * mask, l_ * mask[0:1], d_ * mask[0:1]) if with_mask else (im, l_, d_)
        ref = loss_mod.mapping_loss(a, b, c, gt_image, up, gt_depth, faithful_sign=faithful)
        gi, gl, gd = torch.autograd.grad(ref, [im, l_, d_])
        out, fi, fl, fd = FusedMappingLoss(faithful_sign=faithful)(image, lf, depth, gt_image, gt_lf, gt_depth, mask)
        assert abs(float(out[0]) - float(ref)) <= 2e-5 * max(1.0, abs(float(ref)))
        assert abs(float(out[2]) - float(loss_mod.ssim(a.detach(), gt_image))) <= 2e-5
        assert cases.rel_err(fi.cpu().numpy(), gi.cpu().numpy()) <= 1e-4
        assert cases.rel_err(fd.cpu().numpy(), gd.cpu().numpy()) <= 1e-5
        # the eps-clamped pixel has an enormous (1/eps) gradient in both; compare the rest at 1e-4
        fl2, gl2 = fl.clone(), gl.clone()
        fl2[:, 3, 5] = 0
        gl2[:, 3, 5] = 0
        assert cases.rel_err(fl2.cpu().numpy(), gl2.cpu().numpy()) <= 1e-4


def test_fused_activations_match_torch(dev):
    from leg_slam_b200 import fused, synthetic
    sc = synthetic.make_scene(5000, seed=71, device=dev)
    sc["rotation"] = sc["rotation"] * 1.7  # un-normalised storage, as during training
    leaves = {k: v.clone().requires_grad_(True) for k, v in sc.items()}
    ref = synthetic.activate(leaves)
    act = fused.activations_fwd(sc)
    for k in ("shs", "opacities", "scales", "rotations"):
        assert cases.rel_err(act[k].cpu().numpy(), ref[k].detach().cpu().numpy()) <= 2e-6, k
    g = torch.Generator().manual_seed(72)
    up = {k: torch.randn(ref[k].shape, generator=g).to(dev) for k in ("shs", "opacities", "scales", "rotations")}
    torch.autograd.backward([ref[k] for k in up], [up[k] for k in up])
    out = {k: torch.full_like(sc[k], 7.0) for k in ("scaling", "rotation", "opacity", "features_dc", "features_rest")}
    fused.activations_bwd(sc, act, up["scales"], up["rotations"], up["opacities"], up["shs"], out, accumulate=False)
    for k in out:
        assert cases.rel_err(out[k].cpu().numpy(), leaves[k].grad.cpu().numpy()) <= 5e-6, k
    fused.activations_bwd(sc, act, up["scales"], up["rotations"], up["opacities"], up["shs"], out, accumulate=True)
    for k in out:
        assert cases.rel_err(out[k].cpu().numpy(), 2 * leaves[k].grad.cpu().numpy()) <= 5e-6, k


# ---------------------------------------------------------------------------- Adam / cosine
def test_fused_adam_vs_oracle_and_torch_golden(dev, oracle_mod):
    from leg_slam_b200 import FusedAdam
    params, grads = cases.adam_case()
    tp = {k: torch.nn.Parameter(v.clone().to(dev)) for k, v in params.items()}
    opt = FusedAdam([dict(params=[tp[k]], lr=cases.ADAM_LRS[k], name=k) for k in tp], lr=0.0, eps=1e-15)
    op = {k: v.clone().numpy() for k, v in params.items()}
    om = {k: np.zeros_like(v) for k, v in op.items()}
    ov = {k: np.zeros_like(v) for k, v in op.items()}
    try:
        gd = golden("adam")
    except pytest.skip.Exception:
        gd = None
    for step, gr in enumerate(grads):
        for k in tp:
            tp[k].grad = gr[k].to(dev)
            oracle_mod.adam(op[k].reshape(-1), gr[k].numpy().reshape(-1), om[k].reshape(-1), ov[k].reshape(-1),
                            cases.ADAM_LRS[k], step=step + 1)
        opt.step()
        for k in tp:
            got = tp[k].detach().cpu().numpy()
            np.testing.assert_array_equal(got.view(np.uint32), op[k].view(np.uint32), err_msg=f"{k} step {step}")  # bit-exact vs oracle
            if gd is not None:
                assert cases.rel_err(got, gd[f"p{step}_{k}"]) <= 1e-6, k  # Adam-updated params <= 1e-6 rel
    if gd is not None:
        for k in tp:
            assert cases.rel_err(opt.state[tp[k]]["exp_avg"].cpu().numpy(), gd[f"m_{k}"]) <= 1e-6
            assert cases.rel_err(opt.state[tp[k]]["exp_avg_sq"].cpu().numpy(), gd[f"v_{k}"]) <= 1e-6


def test_adam_odd_sizes_and_unaligned(dev, oracle_mod):
    """Ragged tails, a tensor smaller than one vector, an unaligned view."""
    from leg_slam_b200 import FusedAdam
    g = torch.Generator().manual_seed(3)
    shapes = [(1,), (3,), (4097,), (8192,), (12291,)]
    base = [torch.randn(s[0] + 1, generator=g) for s in shapes]
    ps = [torch.nn.Parameter(b.to(dev)[1:]) for b in base]  # +4 bytes: not 16-byte aligned
    gs = [torch.randn(s, generator=g) for s in shapes]
    opt = FusedAdam([dict(params=[p], lr=1e-2) for p in ps], eps=1e-15)
    for p, gr in zip(ps, gs):
        p.grad = gr.to(dev)
    opt.step()
    for b, p, gr in zip(base, ps, gs):
        ref = b[1:].clone().numpy()
        oracle_mod.adam(ref, gr.numpy(), np.zeros_like(ref), np.zeros_like(ref), 1e-2, step=1)
        np.testing.assert_array_equal(p.detach().cpu().numpy().view(np.uint32), ref.view(np.uint32))


def test_dp_adam_shard_kernel_single_rank(dev, oracle_mod):
    """lgs_dp_adam_shard with world = 1 (no peers, no multicast) is exactly the oracle's Adam over the flat
    buffer, per-tensor learning rates resolved from the segment table, two shards covering the index space.
    The multi-rank exchange itself is exercised by tools/check_dp_fused.py under torchrun (needs >= 2 GPUs)."""
    import ctypes
    from leg_slam_b200 import _lib
    L = _lib.lib()
    g = torch.Generator().manual_seed(91)
    sizes = [1200, 400, 6000, 8192, 400, 1200, 1600]
    lrs = [3.2e-4, 2.5e-3, 1.25e-4, 1.5e-3, 0.05, 5e-3, 1e-3]
    n = sum(sizes)
    p0 = torch.randn(n, generator=g)
    grads = [torch.randn(n, generator=g) * 1e-3 for _ in range(2)]
    p = p0.clone().to(dev)
    m = torch.zeros(n, device=dev)
    v = torch.zeros(n, device=dev)
    ref_p, ref_m, ref_v = p0.clone().numpy(), np.zeros(n, np.float32), np.zeros(n, np.float32)
    starts = np.cumsum([0] + sizes)
    seg = (ctypes.c_int64 * (len(sizes) + 1))(*[int(x) for x in starts])
    lr = (ctypes.c_double * len(sizes))(*lrs)
    half = (n // 8) * 4
    for step, gr in enumerate(grads, start=1):
        gd = gr.to(dev)
        gp = (ctypes.c_void_p * 1)(gd.data_ptr())
        pp = (ctypes.c_void_p * 1)(p.data_ptr())
        for b, e in ((0, half), (half, n)):
            _lib.check(L.lgs_dp_adam_shard(len(sizes), seg, lr, 1, 0, gp, pp, None, None, b, e, m[b:].data_ptr(), v[b:].data_ptr(),
                                           0.9, 0.999, 1e-15, step, 0, torch.cuda.current_stream(dev).cuda_stream), "dp_adam")
        for t in range(len(sizes)):
            a, z = int(starts[t]), int(starts[t + 1])
            oracle_mod.adam(ref_p[a:z], gr.numpy()[a:z], ref_m[a:z], ref_v[a:z], lrs[t], step=step)
        np.testing.assert_array_equal(p.cpu().numpy().view(np.uint32), ref_p.view(np.uint32))
    np.testing.assert_array_equal(m.cpu().numpy().view(np.uint32), ref_m.view(np.uint32))
    np.testing.assert_array_equal(v.cpu().numpy().view(np.uint32), ref_v.view(np.uint32))


def test_dp_adam_shard_grid_bound_is_bit_identical(dev):
    """max_ctas only bounds the grid of lgs_dp_adam_shard (a launch that runs underneath other kernels): same bits."""
    import ctypes
    from leg_slam_b200 import _lib
    L = _lib.lib()
    g = torch.Generator().manual_seed(92)
    n = 4 * 70001
    p0, gr = torch.randn(n, generator=g), torch.randn(n, generator=g) * 1e-3
    seg = (ctypes.c_int64 * 3)(0, 4 * 30000, n)
    lr = (ctypes.c_double * 2)(1e-3, 5e-2)
    outs = []
    for cap in (0, 1, 148 * 4):
        p, m, v, gd = p0.clone().to(dev), torch.zeros(n, device=dev), torch.zeros(n, device=dev), gr.to(dev)
        gp, pp = (ctypes.c_void_p * 1)(gd.data_ptr()), (ctypes.c_void_p * 1)(p.data_ptr())
        for step in (1, 2):
            _lib.check(L.lgs_dp_adam_shard(2, seg, lr, 1, 0, gp, pp, None, None, 0, n, m.data_ptr(), v.data_ptr(), 0.9, 0.999, 1e-15,
                                           step, cap, torch.cuda.current_stream(dev).cuda_stream), "dp_adam")
        outs.append((p.cpu(), m.cpu(), v.cpu()))
    for o in outs[1:]:
        for a, b in zip(outs[0], o):
            assert torch.equal(a, b)


def test_stream_hooks(dev):
    """lgs_stream_hooks: the forward waits for the caller's event between binning and the render kernel, the backward records
    the caller's event after the render backward; results are unchanged and NULL clears the hooks."""
    import ctypes
    from leg_slam_b200 import _lib, rasterize_points as rp
    L = _lib.lib()
    cs = cases.make_case("sh3_lf", dev)
    base = rp.rasterize_gaussians(*cases.fwd_args(cs))
    gbase = rp.rasterize_gaussians_backward(*cases.bwd_args(cs, base[4], base[5], base[0], base[6], base[7]))
    side = torch.cuda.Stream(device=dev)
    cur = torch.cuda.current_stream(dev)
    ev_wait, ev_rec = torch.cuda.Event(), torch.cuda.Event()
    ev_wait.record(cur)
    ev_rec.record(cur)
    torch.cuda.synchronize(dev)
    try:
        _lib.check(L.lgs_stream_hooks(ctypes.c_void_p(ev_wait.cuda_event), ctypes.c_void_p(ev_rec.cuda_event)), "hooks")
        # the event the render kernel waits for is recorded behind a long busy kernel on another stream
        with torch.cuda.stream(side):
            torch.cuda._sleep(200_000_000)  # ~0.1 s
            ev_wait.record(side)
        out = rp.rasterize_gaussians(*cases.fwd_args(cs))
        grads = rp.rasterize_gaussians_backward(*cases.bwd_args(cs, out[4], out[5], out[0], out[6], out[7]))
        side.wait_event(ev_rec)  # the backward recorded it: waiting must not dead-lock and must complete
        torch.cuda.synchronize(dev)
        assert ev_rec.query()
    finally:
        _lib.check(L.lgs_stream_hooks(None, None), "hooks")
    assert out[0] == base[0]
    for a, b in zip(out[1:5], base[1:5]):
        assert torch.equal(a, b)
    for n, a, b in zip(cases.GRAD_NAMES, grads, gbase):
        assert cases.rel_err(a.cpu().numpy(), b.cpu().numpy()) <= 1e-4, n


def test_cosine_query(dev, oracle_mod):
    """Tensor-core kernel (tcgen05, 3xTF32) and the SIMT cross-check kernel against the fp64-accumulating oracle."""
    from leg_slam_b200 import cosine_query, relevance_scores
    feats, text = cases.cosine_case()
    sim = cosine_query(feats.to(dev), text.to(dev)).cpu().numpy()
    ref = oracle_mod.cosine(feats.numpy(), text.numpy())
    assert sim.shape == ref.shape and np.abs(sim - ref).max() <= 2e-6
    simt = cosine_query(feats.to(dev), text.to(dev), simt=True).cpu().numpy()
    assert np.abs(simt - ref).max() <= 1e-6
    # shapes that exercise the tile / column-chunk edges of the tensor-core kernel: rows not a multiple of 128,
    # > 148 tiles (persistent loop, TMEM double buffering wraps), Q not a multiple of 4 / 16 / 256
    g = torch.Generator().manual_seed(33)
    for P, Q in ((1, 1), (127, 3), (129, 16), (40001, 17), (5000, 256), (1000, 300)):
        f = (torch.randn(P, 64, generator=g) * (0.1 + torch.rand(P, 1, generator=g)))
        t = torch.randn(Q, 64, generator=g)
        got = cosine_query(f.to(dev), t.to(dev)).cpu().numpy()
        assert np.abs(got - oracle_mod.cosine(f.numpy(), t.numpy())).max() <= 2e-6, (P, Q)
    one = cosine_query(feats.to(dev), text[0].to(dev)).cpu().numpy()
    np.testing.assert_array_equal(one, sim[:, 0])
    rel = relevance_scores(feats.to(dev), text[0].to(dev)).cpu().numpy()
    s0 = ref[:, 0].astype(np.float64)
    np.testing.assert_allclose(rel, 1 - (s0 - s0.min()) / (s0.max() - s0.min()), atol=5e-6)
    try:
        gd = golden("cosine")
    except pytest.skip.Exception:
        return
    assert np.abs(sim - gd["sim"]).max() <= 2e-6
    assert np.abs(rel - gd["relevance0"]).max() <= 5e-6
    # many queries (one CTA column sweep is 256 wide)
    g = torch.Generator().manual_seed(9)
    text2 = torch.randn(300, 64, generator=g)
    sim2 = cosine_query(feats[:1000].to(dev), text2.to(dev)).cpu().numpy()
    assert np.abs(sim2 - oracle_mod.cosine(feats[:1000].numpy(), text2.numpy())).max() <= 2e-6


def test_per_pixel_cosine_query_and_heatmap_render(dev, ref_mod):
    """The per-pixel query of the reference (eval/find_objects_gaussians.py:323) on a rendered feature image, against torch's
    F.cosine_similarity in fp32 (tolerance 2e-6: same formula, different summation order) and a float64 numpy restatement; and
    BASELINE.json configs[4] "query, then heat-map render": per-Gaussian relevance -> heat colours -> a forward through the
    colors_precomp path, against the unmodified reference rasterizer given the same colours (1e-4)."""
    from leg_slam_b200 import cosine_image, heat_colors, heatmap_render, rasterize_points as rp, relevance_scores, synthetic
    cs = cases.make_case("ragged_sh1", dev)
    _R, _c, lf, _d, *_ = rp.rasterize_gaussians(*cases.fwd_args(cs))
    g = torch.Generator().manual_seed(77)
    text = torch.randn(11, 64, generator=g).to(dev)
    lf[:, 2, 3] = 0.0  # a zero feature vector exercises the eps clamp
    got = cosine_image(lf, text)
    ref = torch.stack([torch.nn.functional.cosine_similarity(lf, t[:, None, None], dim=0) for t in text])
    assert got.shape == ref.shape == (11, cs["H"], cs["W"])
    assert float((got - ref).abs().max()) <= 2e-6
    a, t = lf.double().cpu().numpy().reshape(64, -1), text.double().cpu().numpy()
    exact = (t @ a) / (np.maximum(np.linalg.norm(a, axis=0), 1e-8)[None] * np.maximum(np.linalg.norm(t, axis=1), 1e-8)[:, None])
    assert np.abs(got.cpu().numpy().reshape(11, -1) - exact).max() <= 2e-6
    assert torch.equal(cosine_image(lf, text[3]), got[3])
    # heat-map render
    cam = synthetic.make_cameras(1, cs["W"], cs["H"], seed=12)[0].to(dev)
    heat, depth, radii, scores = heatmap_render(cs["means3D"], cs["opacities"], cs["scales"], cs["rotations"], cs["lang_feats"], text[0], cam)
    assert torch.equal(scores, relevance_scores(cs["lang_feats"], text[0]))
    colors = heat_colors(scores)
    s = scores.clamp(0, 1)
    assert torch.allclose(colors, torch.stack([s, 1 - (2 * s - 1).abs(), 1 - s], 1), atol=1e-6)
    e = torch.empty(0, device=dev)
    bg = torch.zeros(3, device=dev)
    Rr, cr, _lr, dr, radr, *_ = ref_mod.rasterize_gaussians(bg, cs["means3D"], colors, e, cs["opacities"], cs["scales"], cs["rotations"], 1.0, e,
                                                          cam.viewmatrix, cam.projmatrix, cam.tanfovx, cam.tanfovy, cs["H"], cs["W"], e, 0,
                                                          cam.campos, False, False)
    assert torch.equal(radr, radii)
    assert cases.rel_err(heat.cpu().numpy(), cr.cpu().numpy()) <= IMG_TOL and cases.rel_err(depth.cpu().numpy(), dr.cpu().numpy()) <= IMG_TOL
