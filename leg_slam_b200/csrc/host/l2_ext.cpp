// l2_ext.cpp -- pybind module `_L2`: the C++ L2 layer (include/gaussian_rasterizer.h) made callable from Python so that the
// test-suite can hold it to the Python twin (leg_slam_b200/rasterizer.py).  Not part of the reference's Python package.
#include <torch/extension.h>

#include "gaussian_model.h"
#include "gaussian_rasterizer.h"
#include "gaussian_renderer.h"
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

// Test hooks on GaussianModel (include/gaussian_model.h): what a C++ caller reaches through optimizer_ directly
std::tuple<int64_t, torch::Tensor, torch::Tensor> model_adam_state(GaussianModel& g, int idx) {
    auto& p = g.optimizer_->param_groups()[idx].params()[0];
    auto& states = g.optimizer_->state();
    auto it = states.find(p.unsafeGetTensorImpl());
    if (it == states.end()) return std::make_tuple((int64_t)-1, torch::Tensor(), torch::Tensor());
    auto& st = static_cast<torch::optim::AdamParamState&>(*it->second);
    return std::make_tuple(st.step(), st.exp_avg(), st.exp_avg_sq());
}
double model_learning_rate(GaussianModel& g, int idx) {
    return static_cast<torch::optim::AdamOptions&>(g.optimizer_->param_groups()[idx].options()).lr();
}
// one optimizer step on given gradients, in the optimizer's group order (src/gaussian_mapper.cpp:793-796)
void model_step(GaussianModel& g, std::vector<torch::Tensor> grads) {
    TORCH_CHECK(grads.size() == 7, "seven gradients");
    for (int i = 0; i < 7; ++i) g.optimizer_->param_groups()[i].params()[0].mutable_grad() = grads[i];
    g.optimizer_->step();
}
bool model_params_are_the_optimizers(GaussianModel& g) {
    torch::Tensor* mine[7] = {&g.xyz_, &g.features_dc_, &g.features_rest_, &g.language_features_, &g.opacity_, &g.scaling_, &g.rotation_};
    for (int i = 0; i < 7; ++i)
        if (!mine[i]->is_same(g.optimizer_->param_groups()[i].params()[0]) || !mine[i]->requires_grad()) return false;
    return g.Tensor_vec_xyz_[0].is_same(g.xyz_) && g.Tensor_vec_rotation_[0].is_same(g.rotation_);
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    pybind11::class_<GaussianOptimizationParams>(m, "GaussianOptimizationParams")
        .def(pybind11::init<>())
        .def_readwrite("position_lr_init_", &GaussianOptimizationParams::position_lr_init_)
        .def_readwrite("position_lr_final_", &GaussianOptimizationParams::position_lr_final_)
        .def_readwrite("position_lr_delay_mult_", &GaussianOptimizationParams::position_lr_delay_mult_)
        .def_readwrite("position_lr_max_steps_", &GaussianOptimizationParams::position_lr_max_steps_)
        .def_readwrite("feature_lr_", &GaussianOptimizationParams::feature_lr_)
        .def_readwrite("language_feature_lr_", &GaussianOptimizationParams::language_feature_lr_)
        .def_readwrite("opacity_lr_", &GaussianOptimizationParams::opacity_lr_)
        .def_readwrite("scaling_lr_", &GaussianOptimizationParams::scaling_lr_)
        .def_readwrite("rotation_lr_", &GaussianOptimizationParams::rotation_lr_)
        .def_readwrite("percent_dense_", &GaussianOptimizationParams::percent_dense_);
    pybind11::class_<GaussianKeyframe, std::shared_ptr<GaussianKeyframe>>(m, "GaussianKeyframe")
        .def(pybind11::init<>())
        .def_readwrite("FoVx_", &GaussianKeyframe::FoVx_)
        .def_readwrite("FoVy_", &GaussianKeyframe::FoVy_)
        .def_readwrite("image_height_", &GaussianKeyframe::image_height_)
        .def_readwrite("image_width_", &GaussianKeyframe::image_width_)
        .def_readwrite("world_view_transform_", &GaussianKeyframe::world_view_transform_)
        .def_readwrite("full_proj_transform_", &GaussianKeyframe::full_proj_transform_)
        .def_readwrite("camera_center_", &GaussianKeyframe::camera_center_)
        .def_readwrite("language_features_", &GaussianKeyframe::language_features_);
    pybind11::class_<GaussianPipelineParams>(m, "GaussianPipelineParams")
        .def(pybind11::init<bool, bool>(), pybind11::arg("convert_SHs") = false, pybind11::arg("compute_cov3D") = false);
    m.def("render", [](std::shared_ptr<GaussianKeyframe> cam, int H, int W, std::shared_ptr<GaussianModel> g, GaussianPipelineParams pipe,
                       torch::Tensor bg, torch::Tensor override_color, float scaling_modifier, bool use_override, bool include_lf) {
        return GaussianRenderer::render(cam, H, W, g, pipe, bg, override_color, scaling_modifier, use_override, include_lf);
    });
    m.def("mapping_iteration_backward", [](std::shared_ptr<GaussianModel> g, std::shared_ptr<GaussianKeyframe> cam,
                                           GaussianPipelineParams pipe, torch::Tensor bg, torch::Tensor gt_image, torch::Tensor gt_depth,
                                           torch::Tensor mask, float lambda_dssim, bool stats) {
        pybind11::gil_scoped_release no_gil;  // the autograd engine must not be entered with the GIL held
        return mappingIterationBackward(g, cam, pipe, bg, gt_image, gt_depth, mask, lambda_dssim, stats);
    });
    m.def("mapping_iteration_step", &mappingIterationStep);
    pybind11::class_<GaussianModel, std::shared_ptr<GaussianModel>>(m, "GaussianModel")
        .def(pybind11::init<int>())
        .def("getScalingActivation", &GaussianModel::getScalingActivation)
        .def("getRotationActivation", &GaussianModel::getRotationActivation)
        .def("getXYZ", &GaussianModel::getXYZ)
        .def("getFeatures", &GaussianModel::getFeatures)
        .def("getLanguageFeatures", &GaussianModel::getLanguageFeatures)
        .def("getOpacityActivation", &GaussianModel::getOpacityActivation)
        .def("getCovarianceActivation", &GaussianModel::getCovarianceActivation)
        .def("oneUpShDegree", &GaussianModel::oneUpShDegree)
        .def("setShDegree", &GaussianModel::setShDegree)
        .def("createFromPcd", [](GaussianModel& g, torch::Tensor p, torch::Tensor c, torch::Tensor l, float s) { g.createFromPcd(p, c, l, s); })
        .def("increasePcd", [](GaussianModel& g, torch::Tensor p, torch::Tensor c, int it) { g.increasePcd(p, c, it); })
        .def("applyScaledTransformation", [](GaussianModel& g, float s, torch::Tensor T) { g.applyScaledTransformation(s, T); })
        .def("scaledTransformVisiblePointsOfKeyframe",
             [](GaussianModel& g, torch::Tensor flags, torch::Tensor diff, torch::Tensor view, torch::Tensor proj, int kf_iter, int stable,
                int num, float scale) {
                 g.scaledTransformVisiblePointsOfKeyframe(flags, diff, view, proj, kf_iter, stable, num, scale);
                 return num;
             })
        .def("trainingSetup", &GaussianModel::trainingSetup)
        .def("updateLearningRate", &GaussianModel::updateLearningRate)
        .def("setPositionLearningRate", &GaussianModel::setPositionLearningRate)
        .def("setFeatureLearningRate", &GaussianModel::setFeatureLearningRate)
        .def("setLanguageFeatureLearningRate", &GaussianModel::setLanguageFeatureLearningRate)
        .def("setOpacityLearningRate", &GaussianModel::setOpacityLearningRate)
        .def("setScalingLearningRate", &GaussianModel::setScalingLearningRate)
        .def("setRotationLearningRate", &GaussianModel::setRotationLearningRate)
        .def("resetOpacity", &GaussianModel::resetOpacity)
        .def("prunePoints", [](GaussianModel& g, torch::Tensor mask) { g.prunePoints(mask); })
        .def("densifyAndPrune", &GaussianModel::densifyAndPrune)
        .def("addDensificationStats", [](GaussianModel& g, torch::Tensor v, torch::Tensor f) { g.addDensificationStats(v, f); })
        .def("savePly", [](GaussianModel& g, std::string path, bool with_state) { g.savePly(path, with_state); })
        .def("loadPly", [](GaussianModel& g, std::string path) { g.loadPly(path); })
        .def("percentDense", &GaussianModel::percentDense)
        .def("setPercentDense", &GaussianModel::setPercentDense)
        .def("adam_state", &model_adam_state)
        .def("learning_rate", &model_learning_rate)
        .def("step", &model_step)
        .def("params_are_the_optimizers", &model_params_are_the_optimizers)
        .def_readwrite("active_sh_degree_", &GaussianModel::active_sh_degree_)
        .def_readwrite("max_sh_degree_", &GaussianModel::max_sh_degree_)
        .def_readwrite("xyz_", &GaussianModel::xyz_)
        .def_readwrite("features_dc_", &GaussianModel::features_dc_)
        .def_readwrite("features_rest_", &GaussianModel::features_rest_)
        .def_readwrite("language_features_", &GaussianModel::language_features_)
        .def_readwrite("scaling_", &GaussianModel::scaling_)
        .def_readwrite("rotation_", &GaussianModel::rotation_)
        .def_readwrite("opacity_", &GaussianModel::opacity_)
        .def_readwrite("max_radii2D_", &GaussianModel::max_radii2D_)
        .def_readwrite("xyz_gradient_accum_", &GaussianModel::xyz_gradient_accum_)
        .def_readwrite("denom_", &GaussianModel::denom_)
        .def_readwrite("exist_since_iter_", &GaussianModel::exist_since_iter_)
        .def_readwrite("spatial_lr_scale_", &GaussianModel::spatial_lr_scale_);
    pybind11::class_<GaussianRasterizationSettings>(m, "GaussianRasterizationSettings")
        .def(pybind11::init<int, int, float, float, torch::Tensor&, float, torch::Tensor&, torch::Tensor&, int, torch::Tensor&,
                            bool, bool>());
    pybind11::class_<GaussianRasterizer, std::shared_ptr<GaussianRasterizer>>(m, "GaussianRasterizer")
        .def(pybind11::init<GaussianRasterizationSettings&>())
        .def("forward", &GaussianRasterizer::forward)
        .def("markVisibleGaussians", &GaussianRasterizer::markVisibleGaussians);
    m.def("fused_adam_run", &fused_adam_run);
}
