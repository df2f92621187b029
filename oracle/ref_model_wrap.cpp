// ref_model_wrap.cpp -- pybind entry points around the UNMODIFIED reference GaussianModel, GaussianRasterizer(Function) and
// GaussianRenderer (TEST INFRASTRUCTURE, oracle/_ref/ref_model.so, built by oracle/build_ref.py build_model()).
//
// /root/reference/src/gaussian_model.cpp, gaussian_parameters.cpp, gaussian_rasterizer.cpp and gaussian_renderer.cpp are compiled
// from where they lie; nothing of them is copied into the repository.  Three passages of src/gaussian_mapper.cpp (a file that
// needs ORB-SLAM3 / OpenCV / jsoncpp) -- render -> loss -> backward, density control, optimizer step of trainForOneIteration --
// are #included as they stand from .inc files that build_model() cuts out at build time and removes after the compile
// (RefDensityControl below).  The class is libtorch code and runs on CPU tensors when its parameters say data_device != "cuda"
// (gaussian_model.cpp:37-41), so the reference's density control (addDensificationStats, densifyAndClone / Split / Prune,
// densificationPostfix, prunePoints), its optimizer-state surgery (replaceTensorToOptimizer, resetOpacity), trainingSetup's
// seven Adam groups with libtorch's own Adam::step, the learning-rate schedule (exponLrFunc), the activations, createFromPcd /
// increasePcd and savePly / loadPly all execute HERE, without a GPU.  tests/test_reference_model.py holds
// oracle/densify_ref.py, oracle/ply_ref.py, the C oracle's Adam and leg_slam_b200's host logic to them (SURVEY.md 8f rows 1-3,
// row a18).
//
// What stands in for absent pieces (none of it is arithmetic a pinned result depends on, except where noted):
//   * Eigen / OpenCV / Sophus do not exist in this image; oracle/ref_stubs/ holds type-only stand-ins so that the headers
//     gaussian_model.h pulls in (point3d.h, tensor_utils.h, se3.hpp) compile.  applyScaledTransformation's 4x4 goes through the
//     stand-in's matrix() -> EigenMatrix2TorchTensor (a transpose and a copy).
//   * oracle/ref_stubs/ref_model_prelude.h (force-included in front of the reference source and of this file) re-points three
//     names for a driverless machine with libtorch 2.11: the literal torch::kCUDA inside general_utils::build_rotation, the
//     closing emptyCache() of increasePcd / densifyAndPrune, and the optimizer-state key (std::string in libtorch 2.0.1, the
//     pointer now).  Its header says why each is needed.
//   * The three CUDA operators the class calls -- distCUDA2 (simple-knn), transformPoints and
//     scaleAndTransformThenMarkVisiblePoints (src/operate_points.cu) -- are CUDA-only.  They are defined below as calls into
//     Python callables the test supplies (the numpy / torch restatements in oracle/ingest_ref.py, which are themselves held
//     bit-identical to the compiled reference operators on a B200: tests/test_ingest.py, tests/golden/geometry.npz).
#include <torch/extension.h>

#include <map>
#include <memory>
#include <stdexcept>

#include "include/gaussian_model.h"
#include "include/gaussian_rasterizer.h"
#include "include/gaussian_renderer.h"
#include "include/loss_utils.h"

namespace py = pybind11;

// leaked on purpose: destroying a py::object after interpreter shutdown is undefined
static py::object* g_dist2 = nullptr;
static py::object* g_transform = nullptr;
static py::object* g_scale_transform = nullptr;

static py::object* g_rasterize = nullptr;
static py::object* g_rasterize_backward = nullptr;
static py::object* g_mark_visible = nullptr;

static void set_cb(py::object*& slot, py::object f) {
    if (slot) { delete slot; slot = nullptr; }
    if (!f.is_none()) slot = new py::object(std::move(f));
}

// third_party/simple-knn/spatial.h:14
torch::Tensor distCUDA2(const torch::Tensor& points) {
    if (!g_dist2) throw std::runtime_error("ref_model: distCUDA2 called and no callable set (set_dist2)");
    py::gil_scoped_acquire gil;
    return (*g_dist2)(points).cast<torch::Tensor>();
}

// include/operate_points.h:27-29 -- in place on `points`
void transformPoints(torch::Tensor& points, torch::Tensor& transformmatrix) {
    if (!g_transform) throw std::runtime_error("ref_model: transformPoints called and no callable set (set_transform_points)");
    py::gil_scoped_acquire gil;
    (*g_transform)(points, transformmatrix);
}

// include/operate_points.h:31-40 -- in place on points, rots, point_not_transformed_mask; the callable returns num_transformed
void scaleAndTransformThenMarkVisiblePoints(torch::Tensor& points, torch::Tensor& rots, torch::Tensor& point_not_transformed_mask,
                                            torch::Tensor& point_unstable_mask, torch::Tensor& transformmatrix,
                                            torch::Tensor& viewmatrix, torch::Tensor& projmatrix, int& num_transformed,
                                            const float scale) {
    if (!g_scale_transform)
        throw std::runtime_error("ref_model: scaleAndTransformThenMarkVisiblePoints called and no callable set");
    py::gil_scoped_acquire gil;
    num_transformed = (*g_scale_transform)(points, rots, point_not_transformed_mask, point_unstable_mask, transformmatrix,
                                           viewmatrix, projmatrix, scale).cast<int>();
}

// include/rasterize_points.h:18-39 -- L1 of the boundary.  Under GaussianRasterizerFunction / GaussianRasterizer /
// GaussianRenderer (src/gaussian_rasterizer.cpp, src/gaussian_renderer.cpp, compiled unmodified into this module) the CUDA
// entry points are Python callables: tests/oracle_l1.py runs the CPU oracle behind them and records every argument, so that
// what the reference's glue hands to L1 -- and what it does with the results -- can be compared with the package's glue
// call for call.
std::tuple<int, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor>
RasterizeGaussiansCUDA(const torch::Tensor& background, const torch::Tensor& means3D, const torch::Tensor& colors,
                       const torch::Tensor& lang_feat, const torch::Tensor& opacity, const torch::Tensor& scales,
                       const torch::Tensor& rotations, const float scale_modifier, const torch::Tensor& cov3D_precomp,
                       const torch::Tensor& viewmatrix, const torch::Tensor& projmatrix, const float tan_fovx,
                       const float tan_fovy, const int image_height, const int image_width, const torch::Tensor& sh,
                       const int degree, const torch::Tensor& campos, const bool prefiltered, const bool include_lang_feat) {
    if (!g_rasterize) throw std::runtime_error("ref_model: RasterizeGaussiansCUDA called and no callable set (set_rasterizer)");
    py::gil_scoped_acquire gil;
    py::tuple r = (*g_rasterize)(background, means3D, colors, lang_feat, opacity, scales, rotations, scale_modifier,
                                 cov3D_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy, image_height, image_width, sh, degree,
                                 campos, prefiltered, include_lang_feat);
    auto T = [&](int i) { return r[i].cast<torch::Tensor>(); };
    return std::make_tuple(r[0].cast<int>(), T(1), T(2), T(3), T(4), T(5), T(6), T(7));
}

// include/rasterize_points.h:41-66
std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor,
           torch::Tensor>
RasterizeGaussiansBackwardCUDA(const torch::Tensor& background, const torch::Tensor& means3D, const torch::Tensor& radii,
                               const torch::Tensor& colors, const torch::Tensor& lang_feat, const torch::Tensor& scales,
                               const torch::Tensor& rotations, const float scale_modifier, const torch::Tensor& cov3D_precomp,
                               const torch::Tensor& viewmatrix, const torch::Tensor& projmatrix, const float tan_fovx,
                               const float tan_fovy, const torch::Tensor& dL_dout_color, const torch::Tensor& dL_dout_lang_feat,
                               const torch::Tensor& dL_dout_depth, const torch::Tensor& sh, const int degree,
                               const torch::Tensor& campos, const torch::Tensor& geomBuffer, const int R,
                               const torch::Tensor& binningBuffer, const torch::Tensor& imageBuffer,
                               const bool include_lang_feat) {
    if (!g_rasterize_backward) throw std::runtime_error("ref_model: RasterizeGaussiansBackwardCUDA called and no callable set");
    py::gil_scoped_acquire gil;
    py::tuple r = (*g_rasterize_backward)(background, means3D, radii, colors, lang_feat, scales, rotations, scale_modifier,
                                          cov3D_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy, dL_dout_color,
                                          dL_dout_lang_feat, dL_dout_depth, sh, degree, campos, geomBuffer, R, binningBuffer,
                                          imageBuffer, include_lang_feat);
    auto T = [&](int i) { return r[i].cast<torch::Tensor>(); };
    return std::make_tuple(T(0), T(1), T(2), T(3), T(4), T(5), T(6), T(7), T(8));
}

// include/rasterize_points.h:68-71
torch::Tensor markVisible(torch::Tensor& means3D, torch::Tensor& viewmatrix, torch::Tensor& projmatrix) {
    if (!g_mark_visible) throw std::runtime_error("ref_model: markVisible called and no callable set");
    py::gil_scoped_acquire gil;
    return (*g_mark_visible)(means3D, viewmatrix, projmatrix).cast<torch::Tensor>();
}

namespace {

struct RefModel {
    std::shared_ptr<GaussianModel> m;

    explicit RefModel(int sh_degree) {
        GaussianModelParams p("", "", "", sh_degree, "images", -1.0f, false, /*data_device=*/"cpu", false);
        m = std::make_shared<GaussianModel>(p);
    }

    // the seven leaves + exist_since_iter_ of an existing map (what createFromPcd / loadPly leave behind)
    void set_state(std::vector<torch::Tensor> t, torch::Tensor exist_since_iter, double spatial_lr_scale) {
        TORCH_CHECK(t.size() == 7, "seven tensors: xyz, f_dc, f_rest, lang, opacity, scaling, rotation");
        auto leaf = [](const torch::Tensor& x) { return x.detach().clone().requires_grad_(); };
        m->xyz_ = leaf(t[0]);
        m->features_dc_ = leaf(t[1]);
        m->features_rest_ = leaf(t[2]);
        m->language_features_ = leaf(t[3]);
        m->opacity_ = leaf(t[4]);
        m->scaling_ = leaf(t[5]);
        m->rotation_ = leaf(t[6]);
        m->Tensor_vec_xyz_ = {m->xyz_};
        m->Tensor_vec_feature_dc_ = {m->features_dc_};
        m->Tensor_vec_feature_rest_ = {m->features_rest_};
        m->Tensor_vec_language_feature_ = {m->language_features_};
        m->Tensor_vec_opacity_ = {m->opacity_};
        m->Tensor_vec_scaling_ = {m->scaling_};
        m->Tensor_vec_rotation_ = {m->rotation_};
        m->exist_since_iter_ = exist_since_iter.detach().clone();
        m->max_radii2D_ = torch::zeros({m->xyz_.size(0)});
        m->spatial_lr_scale_ = (float)spatial_lr_scale;
    }

    void training_setup(double position_lr_init, double position_lr_final, double position_lr_delay_mult,
                        int position_lr_max_steps, double feature_lr, double language_feature_lr, double opacity_lr,
                        double scaling_lr, double rotation_lr, double percent_dense) {
        GaussianOptimizationParams o(30000, (float)position_lr_init, (float)position_lr_final, (float)position_lr_delay_mult,
                                     position_lr_max_steps, (float)feature_lr, (float)language_feature_lr, (float)opacity_lr,
                                     (float)scaling_lr, (float)rotation_lr, (float)percent_dense);
        m->trainingSetup(o);
    }

    std::vector<torch::Tensor> params() {
        return {m->xyz_, m->features_dc_, m->features_rest_, m->language_features_, m->opacity_, m->scaling_, m->rotation_};
    }

    // the tensors the optimizer itself holds per group (must be the same objects as params())
    std::vector<torch::Tensor> optimizer_params() {
        std::vector<torch::Tensor> out;
        for (auto& g : m->optimizer_->param_groups()) out.push_back(g.params()[0]);
        return out;
    }

    void set_grads(std::vector<torch::Tensor> g) {
        auto p = params();
        TORCH_CHECK(g.size() == 7);
        for (int i = 0; i < 7; ++i) p[i].mutable_grad() = g[i].detach().clone();
    }

    void step() { m->optimizer_->step(); }
    void zero_grad() { m->optimizer_->zero_grad(); }

    // (step, exp_avg, exp_avg_sq) of group i, or None when the optimizer holds no state for its tensor
    py::object moments(int i) {
        auto& param = m->optimizer_->param_groups()[i].params()[0];
        auto& state = m->optimizer_->state();
        auto it = state.find(param.unsafeGetTensorImpl());
        if (it == state.end()) return py::none();
        auto& s = static_cast<torch::optim::AdamParamState&>(*it->second);
        return py::make_tuple((int64_t)s.step(), s.exp_avg(), s.exp_avg_sq());
    }

    int64_t state_size() { return (int64_t)m->optimizer_->state().size(); }

    std::vector<double> lrs() {
        std::vector<double> out;
        for (auto& g : m->optimizer_->param_groups()) out.push_back(static_cast<torch::optim::AdamOptions&>(g.options()).lr());
        return out;
    }

    void add_densification_stats(torch::Tensor viewspace_grad, torch::Tensor update_filter) {
        torch::Tensor vs = torch::zeros_like(viewspace_grad).requires_grad_();
        vs.mutable_grad() = viewspace_grad;
        m->addDensificationStats(vs, update_filter);
    }

    void create_from_pcd(torch::Tensor xyz, torch::Tensor color, torch::Tensor lang, double spatial_lr_scale) {
        std::map<point3D_id_t, Point3D> pcd;
        auto x = xyz.to(torch::kDouble).contiguous(), c = color.to(torch::kFloat).contiguous(),
             l = lang.to(torch::kFloat).contiguous();
        for (int64_t i = 0; i < x.size(0); ++i) {
            Point3D p;
            for (int k = 0; k < 3; ++k) {
                p.xyz_(k) = x.data_ptr<double>()[3 * i + k];
                p.color_(k) = c.data_ptr<float>()[3 * i + k];
            }
            for (int k = 0; k < LANGUAGE_FEATURES_DIM; ++k) p.lang_features_(k) = l.data_ptr<float>()[LANGUAGE_FEATURES_DIM * i + k];
            pcd[(point3D_id_t)i] = p;
        }
        m->createFromPcd(pcd, (float)spatial_lr_scale);
    }

    void apply_scaled_transformation(double s, torch::Tensor R, torch::Tensor t) {
        Eigen::Matrix3f Re;
        Eigen::Vector3f te;
        auto Rc = R.to(torch::kFloat).contiguous(), tc = t.to(torch::kFloat).contiguous();
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j) Re(i, j) = Rc.data_ptr<float>()[3 * i + j];
            te(i) = tc.data_ptr<float>()[i];
        }
        m->applyScaledTransformation((float)s, Sophus::SE3f(Re, te));
    }

    int scaled_transform_visible_points_of_keyframe(torch::Tensor flags, torch::Tensor diff_pose, torch::Tensor view,
                                                    torch::Tensor proj, int kf_creation_iter, int stable_num_iter_existence,
                                                    double scale) {
        int n = 0;
        m->scaledTransformVisiblePointsOfKeyframe(flags, diff_pose, view, proj, kf_creation_iter, stable_num_iter_existence, n,
                                                  (float)scale);
        return n;
    }
};

static std::shared_ptr<GaussianKeyframe> make_keyframe(double FoVx, double FoVy, torch::Tensor view, torch::Tensor proj,
                                                       torch::Tensor campos) {
    auto kf = std::make_shared<GaussianKeyframe>();   // only the fields render() reads (src/gaussian_renderer.cpp:51-66)
    kf->FoVx_ = (float)FoVx;
    kf->FoVy_ = (float)FoVy;
    kf->world_view_transform_ = view;
    kf->full_proj_transform_ = proj;
    kf->camera_center_ = campos;
    return kf;
}

// GaussianRenderer::render (src/gaussian_renderer.cpp:24-160) on a RefModel
static std::vector<torch::Tensor> ref_render(RefModel& model, double FoVx, double FoVy, torch::Tensor view, torch::Tensor proj,
                                             torch::Tensor campos, int image_height, int image_width, bool convert_SHs,
                                             bool compute_cov3D, torch::Tensor bg, torch::Tensor override_color,
                                             double scaling_modifier, bool use_override_color, bool include_language_features) {
    auto kf = make_keyframe(FoVx, FoVy, view, proj, campos);
    GaussianPipelineParams pipe(convert_SHs, compute_cov3D);
    auto r = GaussianRenderer::render(kf, image_height, image_width, model.m, pipe, bg, override_color, (float)scaling_modifier,
                                      use_override_color, include_language_features);
    return {std::get<0>(r), std::get<1>(r), std::get<2>(r), std::get<3>(r), std::get<4>(r), std::get<5>(r)};
}

// GaussianRasterizer::forward (src/gaussian_rasterizer.cpp:178-236) with its own settings object
static std::vector<torch::Tensor> ref_rasterizer_forward(int image_height, int image_width, double tanfovx, double tanfovy,
                                                         torch::Tensor bg, double scale_modifier, torch::Tensor view,
                                                         torch::Tensor proj, int sh_degree, torch::Tensor campos, bool prefiltered,
                                                         bool include_language_features, torch::Tensor means3D,
                                                         torch::Tensor means2D, torch::Tensor opacities, bool has_shs,
                                                         bool has_colors_precomp, bool has_lang_feat, bool has_scales,
                                                         bool has_rotations, bool has_cov3D_precomp, torch::Tensor shs,
                                                         torch::Tensor colors_precomp, torch::Tensor lang_feat,
                                                         torch::Tensor scales, torch::Tensor rotations, torch::Tensor cov3D_precomp) {
    GaussianRasterizationSettings rs(image_height, image_width, (float)tanfovx, (float)tanfovy, bg, (float)scale_modifier, view, proj,
                                     sh_degree, campos, prefiltered, include_language_features);
    GaussianRasterizer rasterizer(rs);
    auto r = rasterizer.forward(means3D, means2D, opacities, has_shs, has_colors_precomp, has_lang_feat, has_scales, has_rotations,
                                has_cov3D_precomp, shs, colors_precomp, lang_feat, scales, rotations, cov3D_precomp);
    return {std::get<0>(r), std::get<1>(r), std::get<2>(r), std::get<3>(r)};
}

static torch::Tensor ref_mark_visible_gaussians(torch::Tensor view, torch::Tensor proj, torch::Tensor campos, torch::Tensor bg,
                                                torch::Tensor positions) {
    GaussianRasterizationSettings rs(1, 1, 1.0f, 1.0f, bg, 1.0f, view, proj, 0, campos, false, false);
    GaussianRasterizer rasterizer(rs);
    return rasterizer.markVisibleGaussians(positions);
}

// The density-control block and the optimizer step of GaussianMapper::trainForOneIteration, run as they stand
// (src/gaussian_mapper.cpp:737-761 and 793-797, cut out of the file at build time by oracle/build_ref.py build_model() --
// gaussian_mapper.cpp itself needs ORB-SLAM3 / OpenCV / jsoncpp).  The members below carry the names those lines use, with
// the meaning include/gaussian_mapper.h and src/gaussian_mapper.cpp:338-352,1841-1854 give them; `gaussians_` is the
// reference's own GaussianModel.
struct RefDensityControl {
    std::shared_ptr<GaussianModel> gaussians_;
    GaussianOptimizationParams opt_params_;
    GaussianModelParams model_params_;
    struct Scene { float cameras_extent_ = 0.0f; } scene_storage_;
    Scene* scene_ = &scene_storage_;
    int prune_big_point_after_iter_ = 0;
    float densify_min_opacity_ = 0.005f;
    int iteration_ = 0;

    int getIteration() { return iteration_; }
    int opacityResetInterval() { return opt_params_.opacity_reset_interval_; }
    float densifyGradThreshold() { return opt_params_.densify_grad_threshold_; }
    int densifyInterval() { return opt_params_.densification_interval_; }

    GaussianPipelineParams pipe_params_;
    torch::Tensor background_ = torch::zeros({3});
    torch::Tensor override_color_ = torch::empty(0);
    torch::DeviceType device_type_ = torch::kCPU;
    float lambdaDssim() { return opt_params_.lambda_dssim_; }

    // one iteration: render -> loss -> backward (:686-724), then, under no_grad as at :729, density control (:737-761) and the
    // optimizer step (:793-797) -- all three pieces are the reference's lines; what lies between them in the file (timing,
    // logging, keyframe recording) is left out.  Returns (loss, radii, visibility_filter).
    std::vector<torch::Tensor> train_iteration(int iteration, double FoVx, double FoVy, torch::Tensor view, torch::Tensor proj,
                                               torch::Tensor campos, torch::Tensor kf_language_features, int image_height,
                                               int image_width, torch::Tensor gt_image, torch::Tensor gt_depth, torch::Tensor mask) {
        iteration_ = iteration;
        std::shared_ptr<GaussianKeyframe> viewpoint_cam = make_keyframe(FoVx, FoVy, view, proj, campos);
        viewpoint_cam->language_features_ = kf_language_features;
#include "render_loss_backward_block.inc"
        {
            torch::NoGradGuard no_grad;
#include "density_control_block.inc"
#include "optimizer_step_block.inc"
        }
        return {loss.detach(), radii, visibility_filter};
    }

    void run(int iteration, torch::Tensor viewspace_grad, torch::Tensor visibility_filter, torch::Tensor radii) {
        iteration_ = iteration;
        torch::Tensor viewspace_point_tensor = torch::zeros_like(viewspace_grad).requires_grad_();
        viewspace_point_tensor.mutable_grad() = viewspace_grad;
        {
            torch::NoGradGuard no_grad;   // the block sits inside the no_grad scope opened at :729
#include "density_control_block.inc"
#include "optimizer_step_block.inc"
        }
    }
};

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, mod) {
    mod.def("set_dist2", [](py::object f) { set_cb(g_dist2, f); });
    mod.def("set_transform_points", [](py::object f) { set_cb(g_transform, f); });
    mod.def("set_scale_and_transform", [](py::object f) { set_cb(g_scale_transform, f); });
    mod.def("set_rasterizer", [](py::object fwd, py::object bwd, py::object mark) {
        set_cb(g_rasterize, fwd);
        set_cb(g_rasterize_backward, bwd);
        set_cb(g_mark_visible, mark);
    });
    py::class_<RefDensityControl>(mod, "DensityControl")
        .def(py::init([](RefModel& model, int iterations, int densification_interval, int opacity_reset_interval, int densify_from_iter,
                         int densify_until_iter, double densify_grad_threshold, double densify_min_opacity,
                         int prune_big_point_after_iter, bool white_background, double cameras_extent) {
                 auto d = std::make_unique<RefDensityControl>();
                 d->gaussians_ = model.m;
                 d->opt_params_.iterations_ = iterations;
                 d->opt_params_.densification_interval_ = densification_interval;
                 d->opt_params_.opacity_reset_interval_ = opacity_reset_interval;
                 d->opt_params_.densify_from_iter_ = densify_from_iter;
                 d->opt_params_.densify_until_iter_ = densify_until_iter;
                 d->opt_params_.densify_grad_threshold_ = (float)densify_grad_threshold;
                 d->densify_min_opacity_ = (float)densify_min_opacity;
                 d->prune_big_point_after_iter_ = prune_big_point_after_iter;
                 d->model_params_.white_background_ = white_background;
                 d->scene_storage_.cameras_extent_ = (float)cameras_extent;
                 return d;
             }),
             py::arg("model"), py::arg("iterations"), py::arg("densification_interval"), py::arg("opacity_reset_interval"),
             py::arg("densify_from_iter"), py::arg("densify_until_iter"), py::arg("densify_grad_threshold"),
             py::arg("densify_min_opacity"), py::arg("prune_big_point_after_iter"), py::arg("white_background"),
             py::arg("cameras_extent"))
        .def("run", &RefDensityControl::run)
        .def("train_iteration", &RefDensityControl::train_iteration, py::call_guard<py::gil_scoped_release>())   // loss.backward() inside
        .def("set_lambda_dssim", [](RefDensityControl& d, double v) { d.opt_params_.lambda_dssim_ = (float)v; })
        .def("set_background", [](RefDensityControl& d, torch::Tensor bg) { d.background_ = bg; });
    mod.def("render", &ref_render);
    mod.def("rasterizer_forward", &ref_rasterizer_forward);
    mod.def("mark_visible_gaussians", &ref_mark_visible_gaussians);
    py::class_<RefModel>(mod, "GaussianModel")
        .def(py::init<int>())
        .def("set_state", &RefModel::set_state)
        .def("training_setup", &RefModel::training_setup, py::arg("position_lr_init"), py::arg("position_lr_final"),
             py::arg("position_lr_delay_mult"), py::arg("position_lr_max_steps"), py::arg("feature_lr"),
             py::arg("language_feature_lr"), py::arg("opacity_lr"), py::arg("scaling_lr"), py::arg("rotation_lr"),
             py::arg("percent_dense"))
        .def("params", &RefModel::params)
        .def("optimizer_params", &RefModel::optimizer_params)
        .def("set_grads", &RefModel::set_grads)
        .def("step", &RefModel::step)
        .def("zero_grad", &RefModel::zero_grad)
        .def("moments", &RefModel::moments)
        .def("state_size", &RefModel::state_size)
        .def("lrs", &RefModel::lrs)
        .def("add_densification_stats", &RefModel::add_densification_stats)
        .def("create_from_pcd", &RefModel::create_from_pcd)
        .def("apply_scaled_transformation", &RefModel::apply_scaled_transformation)
        .def("scaled_transform_visible_points_of_keyframe", &RefModel::scaled_transform_visible_points_of_keyframe)
        .def("increase_pcd", [](RefModel& r, torch::Tensor pts, torch::Tensor cols, int it) {
            r.m->increasePcd(pts, cols, it);
        })
        .def("increase_pcd_vec", [](RefModel& r, std::vector<float> pts, std::vector<float> cols, int it) {
            r.m->increasePcd(pts, cols, it);
        })
        .def("densify_and_prune", [](RefModel& r, double max_grad, double min_opacity, double extent, int max_screen_size) {
            r.m->densifyAndPrune((float)max_grad, (float)min_opacity, (float)extent, max_screen_size);
        })
        .def("densify_and_clone", [](RefModel& r, torch::Tensor g, double thr, double extent) {
            r.m->densifyAndClone(g, (float)thr, (float)extent);
        })
        .def("densify_and_split", [](RefModel& r, torch::Tensor g, double thr, double extent, int N) {
            r.m->densifyAndSplit(g, (float)thr, (float)extent, N);
        })
        .def("prune_points", [](RefModel& r, torch::Tensor mask) { r.m->prunePoints(mask); })
        .def("reset_opacity", [](RefModel& r) { r.m->resetOpacity(); })
        .def("update_learning_rate", [](RefModel& r, int step) { return (double)r.m->updateLearningRate(step); })
        .def("set_position_learning_rate", [](RefModel& r, double v) { r.m->setPositionLearningRate((float)v); })
        .def("set_feature_learning_rate", [](RefModel& r, double v) { r.m->setFeatureLearningRate((float)v); })
        .def("set_language_feature_learning_rate", [](RefModel& r, double v) { r.m->setLanguageFeatureLearningRate((float)v); })
        .def("set_opacity_learning_rate", [](RefModel& r, double v) { r.m->setOpacityLearningRate((float)v); })
        .def("set_scaling_learning_rate", [](RefModel& r, double v) { r.m->setScalingLearningRate((float)v); })
        .def("set_rotation_learning_rate", [](RefModel& r, double v) { r.m->setRotationLearningRate((float)v); })
        .def("one_up_sh_degree", [](RefModel& r) { r.m->oneUpShDegree(); })
        .def("set_sh_degree", [](RefModel& r, int sh) { r.m->setShDegree(sh); })
        .def("active_sh_degree", [](RefModel& r) { return r.m->active_sh_degree_; })
        .def("get_scaling_activation", [](RefModel& r) { return r.m->getScalingActivation(); })
        .def("get_rotation_activation", [](RefModel& r) { return r.m->getRotationActivation(); })
        .def("get_opacity_activation", [](RefModel& r) { return r.m->getOpacityActivation(); })
        .def("get_features", [](RefModel& r) { return r.m->getFeatures(); })
        .def("get_language_features", [](RefModel& r) { return r.m->getLanguageFeatures(); })
        .def("get_covariance_activation", [](RefModel& r, int s) { return r.m->getCovarianceActivation(s); })
        .def("save_ply", [](RefModel& r, std::string p) { r.m->savePly(p); })
        .def("load_ply", [](RefModel& r, std::string p) { r.m->loadPly(p); })
        .def("save_sparse_points_ply", [](RefModel& r, std::string p) { r.m->saveSparsePointsPly(p); })
        .def("percent_dense", [](RefModel& r) { return (double)r.m->percentDense(); })
        .def_property("xyz_gradient_accum", [](RefModel& r) { return r.m->xyz_gradient_accum_; },
                      [](RefModel& r, torch::Tensor t) { r.m->xyz_gradient_accum_ = t.detach().clone(); })
        .def_property("denom", [](RefModel& r) { return r.m->denom_; },
                      [](RefModel& r, torch::Tensor t) { r.m->denom_ = t.detach().clone(); })
        .def_property("max_radii2D", [](RefModel& r) { return r.m->max_radii2D_; },
                      [](RefModel& r, torch::Tensor t) { r.m->max_radii2D_ = t.detach().clone(); })
        .def_property("exist_since_iter", [](RefModel& r) { return r.m->exist_since_iter_; },
                      [](RefModel& r, torch::Tensor t) { r.m->exist_since_iter_ = t.detach().clone(); })
        .def_property_readonly("spatial_lr_scale", [](RefModel& r) { return (double)r.m->spatial_lr_scale_; })
        .def_property_readonly("sparse_points_xyz", [](RefModel& r) { return r.m->sparse_points_xyz_; })
        .def_property_readonly("sparse_points_color", [](RefModel& r) { return r.m->sparse_points_color_; });
}
