#!/usr/bin/env python
"""Build the UNMODIFIED reference rasterizer for sm_100 into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Nothing under leg_slam_b200/ may import or link what
this script produces; only tests/, __graft_entry__.smoke() and bench.py's
reference / cpu_baseline legs do.

The reference ships a self-contained copy of its CUDA rasterizer + libtorch
binding + vendored glm under
    /root/reference/eval/submodules/diff-gaussian-rasterization-legs-slam/
(byte-identical to /root/reference/cuda_rasterizer + src/rasterize_points.cu
except LF_NUM_CHANNELS is hard-coded to 64).  We compile
those sources *where they lie* (no copy into this repo) with our own short
recipe -- not the reference's setup.py/CMake:

  nvcc  -gencode arch=compute_100,code=sm_100 -include cstdint   forward.cu backward.cu rasterizer_impl.cu
  g++   -x c++                                                   rasterize_points.cu ext.cpp   (no kernels in them)
  g++   -shared  ->  oracle/_ref/ref_rasterizer.so   (python module `ref_rasterizer`)

`-include cstdint` is needed because gcc-13 no longer leaks std::uintptr_t
through <iostream> (rasterizer_impl.h:24).  No other change.

oracle/_ref/ is git-ignored (binary) but NOT gpurun-ignored: the .so travels to
the GPU box, where /root/reference does not exist.
"""
import os
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = "/root/reference/eval/submodules/diff-gaussian-rasterization-legs-slam"
MODNAME = "ref_rasterizer"


def _torch_paths():
    import torch  # noqa: F401
    from torch.utils import cpp_extension as ce
    return ce.include_paths(), ce.library_paths()


def build(force=False, verbose=True):
    so = os.path.join(OUT, MODNAME + ".so")
    if os.path.exists(so) and not force:
        return so
    if not os.path.isdir(REF):
        raise RuntimeError("reference tree not present (GPU box?) and no prebuilt " + so)
    os.makedirs(OUT, exist_ok=True)
    inc, lib = _torch_paths()
    pyinc = sysconfig.get_paths()["include"]
    glm = os.path.join(REF, "third_party", "glm")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")

    nvcc = [os.path.join(cuda, "bin", "nvcc"), "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
            "-gencode", "arch=compute_100,code=sm_100", "-include", "cstdint",
            "-I" + glm, "-I" + REF, "-c"]
    cxx = ["g++", "-O2", "-std=c++17", "-fPIC", "-x", "c++",
           "-DTORCH_EXTENSION_NAME=" + MODNAME, "-DTORCH_API_INCLUDE_EXTENSION_H",
           "-D_GLIBCXX_USE_CXX11_ABI=1",
           "-I" + REF, "-I" + os.path.join(cuda, "include"), "-I" + pyinc] + ["-I" + p for p in inc] + ["-c"]

    jobs = []
    objs = []
    for f in ("cuda_rasterizer/forward.cu", "cuda_rasterizer/backward.cu", "cuda_rasterizer/rasterizer_impl.cu"):
        o = os.path.join(OUT, os.path.basename(f) + ".o")
        objs.append(o)
        jobs.append(nvcc + [os.path.join(REF, f), "-o", o])
    for f in ("rasterize_points.cu", "ext.cpp"):
        o = os.path.join(OUT, os.path.basename(f) + ".o")
        objs.append(o)
        jobs.append(cxx + [os.path.join(REF, f), "-o", o])

    def run(cmd):
        if verbose:
            print("[build_ref]", " ".join(cmd), flush=True)
        subprocess.check_call(cmd)

    with ThreadPoolExecutor(max_workers=5) as ex:
        list(ex.map(run, jobs))

    link = ["g++", "-shared", "-o", so] + objs + ["-L" + p for p in lib] + \
           ["-L" + os.path.join(cuda, "lib64"), "-lc10", "-ltorch_cpu", "-ltorch", "-ltorch_python",
            "-lc10_cuda", "-ltorch_cuda", "-lcudart"] + ["-Wl,-rpath," + p for p in lib]
    run(link)
    return so


LOSS_REF_INC = "/root/reference/include"
LOSS_MOD = "ref_loss"
LOSS_SO = os.path.join(OUT, LOSS_MOD + ".so")


def build_loss(force=False, verbose=True):
    """The UNMODIFIED reference loss (include/loss_utils.h: l1_loss, cosine_similarity, ssim, psnr -- header-only libtorch) behind
    oracle/ref_loss_wrap.cpp, which also chains them as src/gaussian_mapper.cpp:707-721 does -> oracle/_ref/ref_loss.so (python
    module `ref_loss`).  Oracle for SURVEY.md 8f row 2 (the fused loss); works on CPU and CUDA tensors.  Only flag added:
    -DLANGUAGE_FEATURES_DIM=64 (the reference's CMakeLists.txt:4 defines it)."""
    if os.path.exists(LOSS_SO) and not force:
        return LOSS_SO
    if not os.path.isdir(LOSS_REF_INC):
        raise RuntimeError("reference tree not present (GPU box?) and no prebuilt " + LOSS_SO)
    os.makedirs(OUT, exist_ok=True)
    inc, lib = _torch_paths()
    pyinc = sysconfig.get_paths()["include"]
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-DTORCH_EXTENSION_NAME=" + LOSS_MOD, "-DTORCH_API_INCLUDE_EXTENSION_H",
           "-D_GLIBCXX_USE_CXX11_ABI=1", "-DLANGUAGE_FEATURES_DIM=64", "-I" + LOSS_REF_INC, "-I" + pyinc] + ["-I" + p for p in inc] + \
          [os.path.join(HERE, "ref_loss_wrap.cpp"), "-o", LOSS_SO] + ["-L" + p for p in lib] + \
          ["-lc10", "-ltorch_cpu", "-ltorch", "-ltorch_python"] + ["-Wl,-rpath," + p for p in lib]
    if verbose:
        print("[build_ref]", " ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return LOSS_SO


def load_loss():
    import importlib.util
    import torch  # noqa: F401
    if not os.path.exists(LOSS_SO):
        raise FileNotFoundError(LOSS_SO + " missing: run `python oracle/build_ref.py` where /root/reference exists")
    spec = importlib.util.spec_from_file_location(LOSS_MOD, LOSS_SO)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


UTILS_MOD = "ref_utils"
UTILS_SO = os.path.join(OUT, UTILS_MOD + ".so")


def build_utils(force=False, verbose=True):
    """The UNMODIFIED reference helper headers include/general_utils.h (inverse_sigmoid, build_rotation) and include/sh_utils.h
    (eval_sh, RGB2SH, SH2RGB) -- header-only libtorch -- behind oracle/ref_utils_wrap.cpp -> oracle/_ref/ref_utils.so (python
    module `ref_utils`).  Pins the numeric helpers of SURVEY.md 8f rows 1 and 4 and of the GaussianRenderer counterpart."""
    if os.path.exists(UTILS_SO) and not force:
        return UTILS_SO
    if not os.path.isdir(LOSS_REF_INC):
        raise RuntimeError("reference tree not present (GPU box?) and no prebuilt " + UTILS_SO)
    os.makedirs(OUT, exist_ok=True)
    inc, lib = _torch_paths()
    pyinc = sysconfig.get_paths()["include"]
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-DTORCH_EXTENSION_NAME=" + UTILS_MOD, "-DTORCH_API_INCLUDE_EXTENSION_H",
           "-D_GLIBCXX_USE_CXX11_ABI=1", "-I" + LOSS_REF_INC, "-I" + pyinc] + ["-I" + p for p in inc] + \
          [os.path.join(HERE, "ref_utils_wrap.cpp"), "-o", UTILS_SO] + ["-L" + p for p in lib] + \
          ["-lc10", "-ltorch_cpu", "-ltorch", "-ltorch_python"] + ["-Wl,-rpath," + p for p in lib]
    if verbose:
        print("[build_ref]", " ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return UTILS_SO


def load_utils():
    import importlib.util
    import torch  # noqa: F401
    if not os.path.exists(UTILS_SO):
        raise FileNotFoundError(UTILS_SO + " missing: run `python oracle/build_ref.py` where /root/reference exists")
    spec = importlib.util.spec_from_file_location(UTILS_MOD, UTILS_SO)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


PLY_REF = "/root/reference/third_party/tinyply"
PLY_SO = os.path.join(OUT, "ref_ply.so")


def build_ply(force=False, verbose=True):
    """The UNMODIFIED reference .ply library (third_party/tinyply/tinyply.cpp + tinyply.h, plain C++) + our C entry points
    oracle/ref_ply_wrap.cpp, which issue savePly's / loadPly's tinyply calls -> oracle/_ref/ref_ply.so (ctypes, CPU only).
    Oracle for SURVEY.md 8f row 3 (.ply checkpoints)."""
    if os.path.exists(PLY_SO) and not force:
        return PLY_SO
    if not os.path.isdir(PLY_REF):
        raise RuntimeError("reference tree not present (GPU box?) and no prebuilt " + PLY_SO)
    os.makedirs(OUT, exist_ok=True)
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I" + PLY_REF, os.path.join(PLY_REF, "tinyply.cpp"),
           os.path.join(HERE, "ref_ply_wrap.cpp"), "-o", PLY_SO]
    if verbose:
        print("[build_ref]", " ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return PLY_SO


def load_ply():
    import ctypes
    if not os.path.exists(PLY_SO):
        raise FileNotFoundError(PLY_SO + " missing: run `python oracle/build_ref.py` where /root/reference exists")
    lib = ctypes.CDLL(PLY_SO)
    fp = ctypes.c_void_p
    lib.ref_ply_write.restype = ctypes.c_int
    lib.ref_ply_write.argtypes = [ctypes.c_char_p] + [ctypes.c_int] * 4 + [fp] * 8
    lib.ref_ply_count.restype = ctypes.c_longlong
    lib.ref_ply_count.argtypes = [ctypes.c_char_p]
    lib.ref_ply_read.restype = ctypes.c_int
    lib.ref_ply_read.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int] + [fp] * 7
    return lib


KNN_REF = "/root/reference/third_party/simple-knn"
KNN_SO = os.path.join(OUT, "ref_simple_knn.so")


def build_knn(force=False, verbose=True):
    """The UNMODIFIED reference simple-knn (third_party/simple-knn/simple_knn.cu, no torch in it) + our C entry point
    oracle/ref_knn_wrap.cu -> oracle/_ref/ref_simple_knn.so (ctypes).  Oracle for SURVEY.md 8f row 4 (distCUDA2)."""
    if os.path.exists(KNN_SO) and not force:
        return KNN_SO
    if not os.path.isdir(KNN_REF):
        raise RuntimeError("reference tree not present (GPU box?) and no prebuilt " + KNN_SO)
    os.makedirs(OUT, exist_ok=True)
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    cmd = [os.path.join(cuda, "bin", "nvcc"), "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-shared",
           "-gencode", "arch=compute_100,code=sm_100", "-include", "cfloat", "-I" + KNN_REF,
           os.path.join(KNN_REF, "simple_knn.cu"), os.path.join(HERE, "ref_knn_wrap.cu"), "-o", KNN_SO, "-lcudart"]
    if verbose:
        print("[build_ref]", " ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return KNN_SO


def load_knn():
    import ctypes
    if not os.path.exists(KNN_SO):
        raise FileNotFoundError(KNN_SO + " missing: run `python oracle/build_ref.py` where /root/reference exists")
    lib = ctypes.CDLL(KNN_SO)
    lib.ref_simple_knn.restype = ctypes.c_int
    lib.ref_simple_knn.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    return lib


GEO_REF = "/root/reference"
GEO_MOD = "ref_geometry"
GEO_SO = os.path.join(OUT, GEO_MOD + ".so")


def _geometry_cmds(force):
    inc, _ = _torch_paths()
    pyinc = sysconfig.get_paths()["include"]
    glm = os.path.join(REF, "third_party", "glm")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    common = ["-D_GLIBCXX_USE_CXX11_ABI=1", "-DLANGUAGE_FEATURES_DIM=64", "-I" + GEO_REF, "-I" + os.path.join(GEO_REF, "include"),
              "-I" + glm, "-I" + pyinc] + ["-I" + p for p in inc]
    nvcc = [os.path.join(cuda, "bin", "nvcc"), "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-gencode",
            "arch=compute_100,code=sm_100", "-include", "cstdint"] + common + ["-c"]
    cxx = ["g++", "-O2", "-std=c++17", "-fPIC", "-DTORCH_EXTENSION_NAME=" + GEO_MOD, "-DTORCH_API_INCLUDE_EXTENSION_H",
           "-I" + os.path.join(cuda, "include")] + common + ["-c"]
    jobs, objs = [], []
    for f in ("stereo_vision.cu", "operate_points.cu"):
        o = os.path.join(OUT, "geo_" + f + ".o")
        objs.append(o)
        src = os.path.join(GEO_REF, "src", f)
        if force or not os.path.exists(o) or os.path.getmtime(o) < os.path.getmtime(src):
            jobs.append(nvcc + [src, "-o", o])
    o = os.path.join(OUT, "ref_geometry_wrap.cpp.o")
    objs.append(o)
    jobs.append(cxx + [os.path.join(HERE, "ref_geometry_wrap.cpp"), "-o", o])
    return jobs, objs


def build_geometry_objects(force=False, verbose=True):
    """Compile step of build_geometry (independent of build(): can run beside it)."""
    if os.path.exists(GEO_SO) and not force:
        return
    if not os.path.isdir(os.path.join(GEO_REF, "src")):
        raise RuntimeError("reference tree not present (GPU box?) and no prebuilt " + GEO_SO)
    os.makedirs(OUT, exist_ok=True)
    jobs, _ = _geometry_cmds(force)

    def run(cmd):
        if verbose:
            print("[build_ref]", " ".join(cmd), flush=True)
        subprocess.check_call(cmd)

    with ThreadPoolExecutor(max_workers=3) as ex:
        list(ex.map(run, jobs))


def build_geometry(force=False, verbose=True):
    """The UNMODIFIED reference geometry operators src/stereo_vision.cu (reprojectDepthPinhole,
    monocularPinholeInactiveGeoDensifyBySearchingNeighborhoodKeypoints) and src/operate_points.cu (transformPoints,
    scaleAndTransformThenMarkVisiblePoints), compiled where they lie for sm_100 and linked with the reference rasterizer
    objects of build() (markVisible) behind oracle/ref_geometry_wrap.cpp -> oracle/_ref/ref_geometry.so (python module
    `ref_geometry`).  Oracle for SURVEY.md 8f row 4 and the two neighbouring operators (tests/test_ingest.py).  Both .cu
    files include <torch/torch.h>, so nvcc needs ~5 minutes for them (compiled in parallel)."""
    if os.path.exists(GEO_SO) and not force:
        return GEO_SO
    build_geometry_objects(force, verbose)
    ras_objs = [os.path.join(OUT, f) for f in ("forward.cu.o", "backward.cu.o", "rasterizer_impl.cu.o", "rasterize_points.cu.o")]
    if not all(os.path.exists(o) for o in ras_objs):
        build(force=True, verbose=verbose)
    _, lib = _torch_paths()
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    _, objs = _geometry_cmds(False)
    cmd = ["g++", "-shared", "-o", GEO_SO] + objs + ras_objs + ["-L" + p for p in lib] + \
          ["-L" + os.path.join(cuda, "lib64"), "-lc10", "-ltorch_cpu", "-ltorch", "-ltorch_python", "-lc10_cuda", "-ltorch_cuda",
           "-lcudart"] + ["-Wl,-rpath," + p for p in lib]
    if verbose:
        print("[build_ref]", " ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return GEO_SO


def load_geometry():
    import importlib.util
    import torch  # noqa: F401
    if not os.path.exists(GEO_SO):
        raise FileNotFoundError(GEO_SO + " missing: run `python oracle/build_ref.py` where /root/reference exists")
    spec = importlib.util.spec_from_file_location(GEO_MOD, GEO_SO)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


MODEL_REF = "/root/reference"
MODEL_MOD = "ref_model"
MODEL_SO = os.path.join(OUT, MODEL_MOD + ".so")
STUBS = os.path.join(HERE, "ref_stubs")


def build_model(force=False, verbose=True):
    """The UNMODIFIED reference GaussianModel (src/gaussian_model.cpp + src/gaussian_parameters.cpp, libtorch; tinyply.cpp for
    savePly / loadPly), its autograd glue and renderer (src/gaussian_rasterizer.cpp, src/gaussian_renderer.cpp) behind oracle/ref_model_wrap.cpp -> oracle/_ref/ref_model.so (python module `ref_model`, runs on CPU
    tensors).  Eigen / OpenCV / Sophus are absent from this image: oracle/ref_stubs/ (first on the include path) holds type-only
    stand-ins for the headers gaussian_model.h pulls in, and ref_model_prelude.h (force-included) re-points three names for a
    driverless machine with libtorch 2.11: the literal torch::kCUDA in build_rotation, emptyCache(), the optimizer-state key.  The CUDA operators the class calls are supplied by the test as Python
    callables (see the wrapper's header).  Pins SURVEY.md 8f rows 1-3 and a18's Adam formula by the reference itself."""
    if os.path.exists(MODEL_SO) and not force:
        return MODEL_SO
    if not os.path.isdir(os.path.join(MODEL_REF, "src")):
        raise RuntimeError("reference tree not present (GPU box?) and no prebuilt " + MODEL_SO)
    os.makedirs(OUT, exist_ok=True)
    objdir = os.path.join(OUT, "model_obj")
    os.makedirs(objdir, exist_ok=True)
    inc, lib = _torch_paths()
    pyinc = sysconfig.get_paths()["include"]
    common = ["g++", "-O1", "-std=c++17", "-fPIC", "-DLANGUAGE_FEATURES_DIM=64", "-D_GLIBCXX_USE_CXX11_ABI=1",
              "-DTORCH_EXTENSION_NAME=" + MODEL_MOD, "-DTORCH_API_INCLUDE_EXTENSION_H",
              "-include", os.path.join(STUBS, "ref_model_prelude.h"),
              "-I" + STUBS, "-I" + MODEL_REF, "-I" + os.path.join(MODEL_REF, "include"), "-I/usr/local/cuda/include",
              "-I" + pyinc] + ["-I" + p for p in inc]
    srcs = [os.path.join(MODEL_REF, "src", "gaussian_model.cpp"), os.path.join(MODEL_REF, "src", "gaussian_parameters.cpp"),
            os.path.join(MODEL_REF, "src", "gaussian_rasterizer.cpp"), os.path.join(MODEL_REF, "src", "gaussian_renderer.cpp"),
            os.path.join(MODEL_REF, "third_party", "tinyply", "tinyply.cpp"), os.path.join(HERE, "ref_model_wrap.cpp")]
    objs = [os.path.join(objdir, os.path.basename(s) + ".o") for s in srcs]
    cmds = [common + ["-c", s, "-o", o] for s, o in zip(srcs, objs)]

    def run(cmd):
        if verbose:
            print("[build_ref]", " ".join(cmd), flush=True)
        subprocess.check_call(cmd)

    # Render -> loss -> backward (:686-724), the density-control block (:737-761) and the optimizer step (:793-797) of
    # GaussianMapper::trainForOneIteration (src/gaussian_mapper.cpp) cannot be compiled with their file (ORB-SLAM3, OpenCV, jsoncpp ...).  The wrapper runs those very lines inside a
    # harness whose members carry the mapper's names: they are cut out of the file HERE, at build time, into three .inc files
    # next to the objects (git-ignored), #included by the wrapper, and removed again after the compile.
    mapper_src = open(os.path.join(MODEL_REF, "src", "gaussian_mapper.cpp")).read().splitlines()
    a = next(i for i, ln in enumerate(mapper_src) if ln.strip() == "// Densification")
    b = next(i for i in range(a, len(mapper_src)) if mapper_src[i].strip().startswith("auto iter_end_timing"))
    c = next(i for i in range(b, len(mapper_src)) if mapper_src[i].strip() == "// Optimizer step")
    r0 = next(i for i, ln in enumerate(mapper_src) if ln.strip() == "// Render")
    r1 = next(i for i in range(r0, len(mapper_src)) if mapper_src[i].strip() == "loss.backward();")
    incs = {"density_control_block.inc": mapper_src[a:b], "optimizer_step_block.inc": mapper_src[c:c + 5],
            "render_loss_backward_block.inc": mapper_src[r0:r1 + 1]}
    assert r1 < a and "GaussianRenderer::render" in "\n".join(incs["render_loss_backward_block.inc"])
    assert "resetOpacity" in "\n".join(incs["density_control_block.inc"]) and "zero_grad" in "\n".join(incs["optimizer_step_block.inc"])
    for name, lines in incs.items():
        with open(os.path.join(objdir, name), "w") as f:
            f.write("\n".join(lines) + "\n")
    cmds = [cmd + ["-I" + objdir] for cmd in cmds]
    try:
        with ThreadPoolExecutor(max_workers=len(cmds)) as ex:
            list(ex.map(run, cmds))
    finally:
        for name in incs:
            os.remove(os.path.join(objdir, name))
    run(["g++", "-shared"] + objs + ["-o", MODEL_SO] + ["-L" + p for p in lib] +
        ["-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch", "-ltorch_python"] + ["-Wl,-rpath," + p for p in lib])
    return MODEL_SO


def load_model():
    import importlib.util
    import torch  # noqa: F401
    if not os.path.exists(MODEL_SO):
        raise FileNotFoundError(MODEL_SO + " missing: run `python oracle/build_ref.py` where /root/reference exists")
    spec = importlib.util.spec_from_file_location(MODEL_MOD, MODEL_SO)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load():
    """Import the prebuilt module (after `import torch`)."""
    import importlib.util
    import torch  # noqa: F401  (libtorch must be loaded first)
    so = os.path.join(OUT, MODNAME + ".so")
    if not os.path.exists(so):
        raise FileNotFoundError(so + " missing: run `python oracle/build_ref.py` where /root/reference exists")
    spec = importlib.util.spec_from_file_location(MODNAME, so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
    print(build_knn(force="--force" in sys.argv))
    print(build_loss(force="--force" in sys.argv))
    print(build_ply(force="--force" in sys.argv))
    print(build_utils(force="--force" in sys.argv))
    print(build_model(force="--force" in sys.argv))
    print(build_geometry(force="--force" in sys.argv))
