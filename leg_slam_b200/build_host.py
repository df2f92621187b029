#!/usr/bin/env python
"""Build the C++ host layers above the C ABI, in-tree (g++ only, no nvcc):

  leg_slam_b200/liblgs_host.so : CudaRasterizer::Rasterizer (include/cuda_rasterizer/rasterizer.h), no torch
  leg_slam_b200/_C.so          : libtorch RasterizeGaussiansCUDA / ...BackwardCUDA / markVisible
                                 (include/rasterize_points.h), the geometry operators transformPoints /
                                 scaleAndTransformThenMarkVisiblePoints / reprojectDepthPinhole /
                                 monocularPinholeInactiveGeoDensify... / distCUDA2 (include/operate_points.h,
                                 stereo_vision.h, spatial.h) + the pybind module `_C`
  leg_slam_b200/_L2.so         : GaussianRasterizationSettings / GaussianRasterizerFunction / GaussianRasterizer
                                 (include/gaussian_rasterizer.h), LgsFusedAdam (include/lgs_adam.h), GaussianModel
                                 (include/gaussian_model.h) + the pybind module `_L2` for the tests
  leg_slam_b200/liblgs_torch.so: the same libtorch layers (L1 functions, geometry operators, L2 rasterizer, LgsFusedAdam,
                                 GaussianModel, GaussianRenderer + mapping-iteration functions) as a plain C++ library with
                                 exported symbols and no pybind module: what a C++ consumer such as the reference's
                                 gaussian_mapper links (tests/cpp/mapper_driver.cpp does)

    python -m leg_slam_b200.build_host [--force]
"""
import os
import subprocess
import sys
import sysconfig

from . import build as lgs_build

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
HOST = os.path.join(PKG, "csrc", "host")
LIB_HOST = os.path.join(PKG, "liblgs_host.so")
LIB_C = os.path.join(PKG, "_C.so")
LIB_L2 = os.path.join(PKG, "_L2.so")
LIB_TORCH = os.path.join(PKG, "liblgs_torch.so")
DRIVER_SRC = os.path.join(ROOT, "tests", "cpp", "mapper_driver.cpp")
DRIVER_EXE = os.path.join(PKG, "build", "host", "mapper_driver")


def build_mapper_driver(force=False, verbose=False):
    """tests/cpp/mapper_driver.cpp: a plain C++ program (no Python in the process) on liblgs_torch.so, built here so that the
    GPU box does not spend its time compiling libtorch headers.  Links libpython only because libtorch_python, which
    <torch/extension.h> (include/rasterize_points.h, like the reference's) pulls in, needs its symbols."""
    if not force and not _stale(DRIVER_EXE, [DRIVER_SRC, LIB_TORCH]):
        return DRIVER_EXE
    import torch  # noqa: F401
    from torch.utils import cpp_extension as ce
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    os.makedirs(os.path.dirname(DRIVER_EXE), exist_ok=True)
    pyver = "python%d.%d" % sys.version_info[:2]
    cmd = ["g++", "-O1", "-std=c++17", "-D_GLIBCXX_USE_CXX11_ABI=1", DRIVER_SRC, "-I" + os.path.join(ROOT, "include"),
           "-I" + os.path.join(cuda, "include"), "-I" + sysconfig.get_paths()["include"]] + ["-I" + p for p in ce.include_paths()]
    cmd += ["-o", DRIVER_EXE, "-L" + PKG, "-llgs_torch", "-llgs"] + ["-L" + p for p in ce.library_paths()]
    cmd += ["-L" + os.path.join(cuda, "lib64"), "-lc10", "-ltorch_cpu", "-ltorch", "-lc10_cuda", "-Wl,--no-as-needed", "-ltorch_cuda",
            "-Wl,--as-needed", "-lcudart", "-L" + (sysconfig.get_config_var("LIBDIR") or "/usr/lib"), "-l" + pyver,
            "-Wl,-rpath," + PKG]
    cmd += ["-Wl,-rpath," + p for p in ce.library_paths()]
    if verbose:
        print("[lgs host]", " ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return DRIVER_EXE


def _stale(target, deps):
    return (not os.path.exists(target)) or any(os.path.getmtime(d) > os.path.getmtime(target) for d in deps)


def build(force=False, verbose=False):
    lgs_build.build()
    inc = os.path.join(ROOT, "include")
    hdrs = [os.path.join(inc, "lgs.h"), os.path.join(inc, "cuda_rasterizer", "rasterizer.h"),
            os.path.join(inc, "rasterize_points.h")]
    geo_hdrs = [os.path.join(inc, h) for h in ("operate_points.h", "stereo_vision.h", "spatial.h")]

    def run(cmd):
        if verbose:
            print("[lgs host]", " ".join(cmd), flush=True)
        subprocess.check_call(cmd)

    src0 = os.path.join(HOST, "rasterizer.cpp")
    if force or _stale(LIB_HOST, [src0] + hdrs):
        run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I" + inc, src0, "-o", LIB_HOST, "-L" + PKG, "-llgs",
             "-Wl,-rpath,$ORIGIN"])
    all_hdrs = hdrs + geo_hdrs + [os.path.join(inc, h) for h in ("gaussian_rasterizer.h", "lgs_adam.h", "gaussian_model.h",
                                                                   "gaussian_renderer.h", "gaussian_keyframe.h")]
    objdir = os.path.join(PKG, "build", "host")
    os.makedirs(objdir, exist_ok=True)
    from concurrent.futures import ThreadPoolExecutor

    def torch_flags():
        import torch  # noqa: F401
        from torch.utils import cpp_extension as ce
        cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
        incs = ["-I" + inc, "-I" + os.path.join(cuda, "include"), "-I" + sysconfig.get_paths()["include"]] + ["-I" + p for p in ce.include_paths()]
        libdirs = ["-L" + p for p in ce.library_paths()] + ["-L" + os.path.join(cuda, "lib64")]
        rpaths = ["-Wl,-rpath," + p for p in ce.library_paths()]
        return incs, libdirs, rpaths

    def target(lib, names, hdr_deps, cflags, suffix, libs):
        """One shared library from libtorch-heavy translation units (~1 min each): objects in parallel, then the link."""
        srcs = [os.path.join(HOST, n) for n in names]
        if not (force or _stale(lib, srcs + hdr_deps)):
            return
        incs, libdirs, rpaths = torch_flags()
        common = ["g++", "-O2", "-std=c++17", "-fPIC", "-D_GLIBCXX_USE_CXX11_ABI=1"] + cflags + incs
        objs = [os.path.join(objdir, n[:-4] + suffix) for n in names]
        with ThreadPoolExecutor(max_workers=len(srcs)) as ex:
            list(ex.map(lambda so: run(common + ["-c", so[0], "-o", so[1]]), zip(srcs, objs)))
        run(["g++", "-shared"] + objs + ["-o", lib, "-L" + PKG, "-llgs", "-Wl,-rpath,$ORIGIN"] + libdirs + libs + rpaths)

    py_libs = ["-lc10", "-ltorch_cpu", "-ltorch", "-ltorch_python", "-lc10_cuda", "-ltorch_cuda", "-lcudart"]
    # libtorch_cuda registers the CUDA backend from static initialisers: nothing references it by symbol, so the C++ library
    # keeps it explicitly, or a pure C++ process has no CUDA tensors (inside Python, torch has loaded it already)
    cpp_libs = ["-lc10", "-ltorch_cpu", "-ltorch", "-lc10_cuda", "-Wl,--no-as-needed", "-ltorch_cuda", "-Wl,--as-needed", "-lcudart"]
    jobs = [
        # L1 (+ the geometry operators) and the pybind module `_C`
        (LIB_C, ["rasterize_points.cpp", "geometry_ops.cpp", "ext.cpp"], hdrs + geo_hdrs,
         ["-DTORCH_EXTENSION_NAME=_C", "-DTORCH_API_INCLUDE_EXTENSION_H"], ".c.o", py_libs),
        # L2 in C++: autograd node + module, optimizer, model, renderer, with its own pybind module for the tests
        (LIB_L2, ["gaussian_rasterizer.cpp", "rasterize_points.cpp", "fused_adam.cpp", "geometry_ops.cpp", "gaussian_model.cpp",
                  "gaussian_renderer.cpp", "l2_ext.cpp"], all_hdrs,
         ["-fvisibility=hidden", "-DTORCH_EXTENSION_NAME=_L2", "-DTORCH_API_INCLUDE_EXTENSION_H"], ".l2.o", py_libs),
        # the libtorch layers as a C++ library (default symbol visibility, no Python module in it)
        (LIB_TORCH, ["rasterize_points.cpp", "geometry_ops.cpp", "gaussian_rasterizer.cpp", "fused_adam.cpp", "gaussian_model.cpp",
                     "gaussian_renderer.cpp"], all_hdrs, [], ".lib.o", cpp_libs),
    ]
    with ThreadPoolExecutor(max_workers=len(jobs)) as ex:  # the three targets side by side
        list(ex.map(lambda j: target(*j), jobs))
    build_mapper_driver(force, verbose)
    return LIB_HOST, LIB_C


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
