"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/lgs.h
declares (no compute calls -- there is no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    from leg_slam_b200 import build
    return build.build()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "lgs.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lgs_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_whole_path():
    syms = declared_symbols()
    for needed in ("lgs_forward_stage1", "lgs_forward_stage2", "lgs_backward", "lgs_mark_visible", "lgs_adam_multi",
                   "lgs_cosine_query", "lgs_geom_bytes", "lgs_image_bytes", "lgs_binning_bytes"):
        assert needed in syms


def test_library_exports_every_declared_symbol(built):
    L = ctypes.CDLL(built)
    for s in declared_symbols():
        assert hasattr(L, s), f"{s} declared in include/lgs.h but not exported by liblgs.so"


def test_python_binding_covers_the_header(built):
    from leg_slam_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    L = _lib.lib()
    assert L.lgs_abi_version() == 1
    assert L.lgs_status_string(0) == b"ok"
    assert b"shs" in L.lgs_status_string(2)
    assert L.lgs_last_cuda_error() == 0


def test_argument_validation_needs_no_gpu(built):
    """Bad arguments are rejected before any CUDA call: status codes, never exceptions or crashes."""
    from leg_slam_b200 import _lib
    L = _lib.lib()
    R = ctypes.c_int(-1)
    # P < 0
    assert L.lgs_forward_stage1(-1, 0, 0, 8, 8, None, None, None, None, None, 1.0, None, None, None, None, None,
                                1.0, 1.0, 0, None, None, ctypes.byref(R), None) == 1
    # P == 0 is a no-op that reports zero instances
    assert L.lgs_forward_stage1(0, 0, 0, 8, 8, None, None, None, None, None, 1.0, None, None, None, None, None,
                                1.0, 1.0, 0, None, None, ctypes.byref(R), None) == 0 and R.value == 0
    buf = (ctypes.c_float * 64)()
    p = ctypes.cast(buf, ctypes.c_void_p)
    # neither SHs nor precomputed colours -> LGS_ERR_NO_COLOR (the reference throws, rasterizer_impl.cu:243-245)
    assert L.lgs_forward_stage1(1, 0, 0, 8, 8, p, None, None, p, p, 1.0, p, None, p, p, p, 1.0, 1.0, 0, p, None,
                                ctypes.byref(R), None) == 2
    # neither scale/rotation nor precomputed covariance -> LGS_ERR_NO_COV
    assert L.lgs_forward_stage1(1, 0, 0, 8, 8, p, None, p, p, None, 1.0, None, None, p, p, p, 1.0, 1.0, 0, p, None,
                                ctypes.byref(R), None) == 3
    assert L.lgs_adam_multi(17, None, None, None, None, None, None, 0.9, 0.999, 1e-15, 1, None) == 1
    assert L.lgs_adam_multi(0, None, None, None, None, None, None, 0.9, 0.999, 1e-15, 1, None) == 0
    assert L.lgs_cosine_query(0, 4, None, None, None, None) == 0
    assert L.lgs_cosine_query(4, 4, None, None, None, None) == 1
    assert L.lgs_mark_visible(0, None, None, None, None, None) == 0
    # data-parallel pair: the hooks only store two handles; the exchange checks its table before touching a device
    assert L.lgs_stream_hooks(None, None) == 0
    seg = (ctypes.c_int64 * 2)(0, 8)
    lr = (ctypes.c_double * 1)(1e-3)
    ptrs = (ctypes.c_void_p * 1)(p.value)
    common = (0.9, 0.999, 1e-15)
    assert L.lgs_dp_adam_shard(0, seg, lr, 1, 0, ptrs, ptrs, None, None, 0, 8, p, p, *common, 1, 0, None) == 1      # n_seg < 1
    assert L.lgs_dp_adam_shard(1, seg, lr, 1, 1, ptrs, ptrs, None, None, 0, 8, p, p, *common, 1, 0, None) == 1      # rank >= world
    assert L.lgs_dp_adam_shard(1, seg, lr, 1, 0, ptrs, ptrs, None, None, 0, 8, p, p, *common, 0, 0, None) == 1      # step < 1
    assert L.lgs_dp_adam_shard(1, seg, lr, 1, 0, ptrs, ptrs, None, None, 2, 8, p, p, *common, 1, 0, None) == 1      # shard not on 16 B
    assert L.lgs_dp_adam_shard(1, seg, lr, 1, 0, ptrs, ptrs, None, None, 4, 4, p, p, *common, 1, 0, None) == 0      # empty shard
    # the operators beside the path: counts, null pointers and alignment are checked before any launch
    assert L.lgs_scale_transform_mark_visible(-1, 1.0, p, p, p, p, p, p, 1, p, None) == 1
    assert L.lgs_scale_transform_mark_visible(0, 1.0, None, None, None, None, None, None, 1, None, None) == 0
    assert L.lgs_scale_transform_mark_visible(4, 1.0, p, p, p, p, p, p, 1, None, None) == 1                         # no counter
    assert L.lgs_scale_transform_mark_visible(4, 1.0, p, ctypes.c_void_p(p.value + 4), p, p, p, p, 1, p, None) == 1  # rots off 16 B
    assert L.lgs_inactive_geo_scratch_bytes(0) == 0 and L.lgs_inactive_geo_scratch_bytes(100) >= 100 * 25
    assert L.lgs_inactive_geo_densify(-1, 8, 1.0, 1.0, 0.0, 0.0, 1.0, p, p, p, p, 64, p, p, p, p, None) == 1
    assert L.lgs_inactive_geo_densify(4, 0, 1.0, 1.0, 0.0, 0.0, 1.0, p, p, p, p, 64, p, p, p, p, None) == 1          # width <= 0
    assert L.lgs_inactive_geo_densify(4, 8, 1.0, 1.0, 0.0, 0.0, 1.0, p, p, p, p, 64, p, p, None, p, None) == 1       # no counter
    assert L.lgs_inactive_geo_densify(4, 8, 1.0, 1.0, 0.0, 0.0, 1.0, p, p, p, None, 64, p, p, p, p, None) == 1       # no colours
    assert L.lgs_inactive_geo_densify(4, 8, 1.0, 1.0, 0.0, 0.0, 1.0, ctypes.c_void_p(p.value + 4), p, p, p, 64, p, p, p, p, None) == 1


def test_sass_is_sm100a_with_tma(built):
    """The shipped cubin targets sm_100a and the blend kernels stage through TMA bulk copies."""
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    elf = subprocess.run(["cuobjdump", "-lelf", built], capture_output=True, text=True).stdout
    assert "sm_100a" in elf
    sass = subprocess.run(["cuobjdump", "-sass", built], capture_output=True, text=True).stdout
    assert "UBLKCP" in sass and "SYNCS.ARRIVE.TRANS64" in sass  # cp.async.bulk + mbarrier expect_tx
