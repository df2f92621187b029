// ref_utils_wrap.cpp -- pybind entry points around the UNMODIFIED reference helper headers (TEST INFRASTRUCTURE,
// oracle/_ref/ref_utils.so): include/general_utils.h (inverse_sigmoid :25-27, build_rotation :29-60) and include/sh_utils.h
// (eval_sh :63-131, RGB2SH :133-135, SH2RGB :137-139).  Both are header-only libtorch code and are included from where they lie
// (-I/root/reference/include on the command line of oracle/build_ref.py build_utils()); nothing of them is copied.
// They carry the numeric part of the reference's density control (split: build_rotation of the parents, resetOpacity / split
// opacity: inverse_sigmoid; src/gaussian_model.cpp:567-575, 665-700), of increasePcd (RGB2SH, :301-357) and of
// GaussianRenderer::render's convert_SHs path (eval_sh, src/gaussian_renderer.cpp:95-116).
#include <torch/extension.h>

#include "general_utils.h"
#include "sh_utils.h"

static torch::Tensor ref_inverse_sigmoid(torch::Tensor x) { return general_utils::inverse_sigmoid(x); }
static torch::Tensor ref_build_rotation(torch::Tensor r) { return general_utils::build_rotation(r); }
static torch::Tensor ref_eval_sh(int64_t deg, torch::Tensor sh, torch::Tensor dirs) { return sh_utils::eval_sh((int)deg, sh, dirs); }
static torch::Tensor ref_rgb2sh(torch::Tensor rgb) { return sh_utils::RGB2SH(rgb); }
static double ref_sh2rgb(double sh) { return (double)sh_utils::SH2RGB((float)sh); }

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.def("inverse_sigmoid", &ref_inverse_sigmoid);
    m.def("build_rotation", &ref_build_rotation);
    m.def("eval_sh", &ref_eval_sh);
    m.def("RGB2SH", &ref_rgb2sh);
    m.def("SH2RGB", &ref_sh2rgb);
}
