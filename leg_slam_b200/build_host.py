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
    srcs = [os.path.join(HOST, "rasterize_points.cpp"), os.path.join(HOST, "geometry_ops.cpp"), os.path.join(HOST, "ext.cpp")]
    if force or _stale(LIB_C, srcs + hdrs + geo_hdrs):
        import torch  # noqa: F401
        from torch.utils import cpp_extension as ce
        cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
        cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-DTORCH_EXTENSION_NAME=_C",
               "-DTORCH_API_INCLUDE_EXTENSION_H", "-D_GLIBCXX_USE_CXX11_ABI=1", "-I" + inc,
               "-I" + os.path.join(cuda, "include"), "-I" + sysconfig.get_paths()["include"]]
        cmd += ["-I" + p for p in ce.include_paths()] + srcs + ["-o", LIB_C, "-L" + PKG, "-llgs", "-Wl,-rpath,$ORIGIN"]
        cmd += ["-L" + p for p in ce.library_paths()] + ["-L" + os.path.join(cuda, "lib64"), "-lc10", "-ltorch_cpu",
                                                          "-ltorch", "-ltorch_python", "-lc10_cuda", "-ltorch_cuda",
                                                          "-lcudart"]
        cmd += ["-Wl,-rpath," + p for p in ce.library_paths()]
        run(cmd)
    # L2 in C++ (include/gaussian_rasterizer.h): autograd node + module, with its own pybind module for the tests
    srcs2 = [os.path.join(HOST, "gaussian_rasterizer.cpp"), os.path.join(HOST, "rasterize_points.cpp"),
             os.path.join(HOST, "fused_adam.cpp"), os.path.join(HOST, "geometry_ops.cpp"), os.path.join(HOST, "gaussian_model.cpp"),
             os.path.join(HOST, "gaussian_renderer.cpp"), os.path.join(HOST, "l2_ext.cpp")]
    if force or _stale(LIB_L2, srcs2 + hdrs + geo_hdrs + [os.path.join(inc, "gaussian_rasterizer.h"), os.path.join(inc, "lgs_adam.h"),
                                                          os.path.join(inc, "gaussian_model.h"), os.path.join(inc, "gaussian_renderer.h"),
                                                          os.path.join(inc, "gaussian_keyframe.h")]):
        import torch  # noqa: F401
        from torch.utils import cpp_extension as ce
        cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
        common = ["g++", "-O2", "-std=c++17", "-fPIC", "-fvisibility=hidden", "-DTORCH_EXTENSION_NAME=_L2",
                  "-DTORCH_API_INCLUDE_EXTENSION_H", "-D_GLIBCXX_USE_CXX11_ABI=1", "-I" + inc,
                  "-I" + os.path.join(cuda, "include"), "-I" + sysconfig.get_paths()["include"]]
        common += ["-I" + p for p in ce.include_paths()]
        objdir = os.path.join(PKG, "build", "host")
        os.makedirs(objdir, exist_ok=True)
        objs = [os.path.join(objdir, os.path.basename(x)[:-4] + ".l2.o") for x in srcs2]
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=len(srcs2)) as ex:  # libtorch-heavy translation units: ~1 min each
            list(ex.map(lambda so: run(common + ["-c", so[0], "-o", so[1]]), zip(srcs2, objs)))
        cmd = ["g++", "-shared"] + objs + ["-o", LIB_L2, "-L" + PKG, "-llgs", "-Wl,-rpath,$ORIGIN"]
        cmd += ["-L" + p for p in ce.library_paths()] + ["-L" + os.path.join(cuda, "lib64"), "-lc10", "-ltorch_cpu",
                                                          "-ltorch", "-ltorch_python", "-lc10_cuda", "-ltorch_cuda",
                                                          "-lcudart"]
        cmd += ["-Wl,-rpath," + p for p in ce.library_paths()]
        run(cmd)
    return LIB_HOST, LIB_C


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
