// l2_ext.cpp -- pybind module `_L2`: the C++ L2 layer (include/gaussian_rasterizer.h) made callable from Python so that the
// test-suite can hold it to the Python twin (leg_slam_b200/rasterizer.py).  Not part of the reference's Python package.
#include <torch/extension.h>

#include "gaussian_rasterizer.h"
#include "lgs_adam.h"

namespace {
// Test hook: the reference's optimizer layout (one single-tensor group per parameter, src/gaussian_model.cpp:488-511) on
// LgsFusedAdam; grads[s][i] is the gradient of parameter i at step s.  Parameters are updated in place; returns
// {exp_avg..., exp_avg_sq..., step counts as a tensor}.
std::vector<torch::Tensor> fused_adam_run(std::vector<torch::Tensor> params, std::vector<std::vector<torch::Tensor>> grads,
                                          std::vector<double> lrs, double eps) {
    TORCH_CHECK(params.size() == lrs.size(), "one learning rate per parameter");
    std::vector<torch::optim::OptimizerParamGroup> groups;
    for (size_t i = 0; i < params.size(); ++i) {
        params[i].set_requires_grad(true);
        groups.emplace_back(std::vector<torch::Tensor>{params[i]},
                            std::make_unique<torch::optim::AdamOptions>(torch::optim::AdamOptions(lrs[i]).eps(eps)));
    }
    LgsFusedAdam opt(groups, torch::optim::AdamOptions(0.0).eps(eps));
    for (auto& g : grads) {
        TORCH_CHECK(g.size() == params.size(), "one gradient per parameter and step");
        for (size_t i = 0; i < params.size(); ++i) params[i].mutable_grad() = g[i];
        opt.step();
    }
    std::vector<torch::Tensor> out;
    std::vector<int64_t> steps;
    for (int which = 0; which < 2; ++which)
        for (auto& p : params) {
            auto& st = static_cast<torch::optim::AdamParamState&>(*opt.state().at(p.unsafeGetTensorImpl()));
            out.push_back(which == 0 ? st.exp_avg() : st.exp_avg_sq());
            if (which == 0) steps.push_back(st.step());
        }
    out.push_back(torch::tensor(steps));
    return out;
}
}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    pybind11::class_<GaussianRasterizationSettings>(m, "GaussianRasterizationSettings")
        .def(pybind11::init<int, int, float, float, torch::Tensor&, float, torch::Tensor&, torch::Tensor&, int, torch::Tensor&,
                            bool, bool>());
    pybind11::class_<GaussianRasterizer, std::shared_ptr<GaussianRasterizer>>(m, "GaussianRasterizer")
        .def(pybind11::init<GaussianRasterizationSettings&>())
        .def("forward", &GaussianRasterizer::forward)
        .def("markVisibleGaussians", &GaussianRasterizer::markVisibleGaussians);
    m.def("fused_adam_run", &fused_adam_run);
}
