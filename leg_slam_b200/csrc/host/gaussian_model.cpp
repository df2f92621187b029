// gaussian_model.cpp -- GaussianModel (include/gaussian_model.h) with the reference's member names and method semantics
// (reference src/gaussian_model.cpp) on liblgs and the libtorch operators of geometry_ops.cpp.  What differs from the
// reference's implementation, unobservably for its callers:
//   * densifyAndPrune is one classification pass and one gather (lgs_densify_plan / lgs_densify_apply) instead of
//     densifyAndClone + densifyAndSplit + prunePoints (~150 libtorch kernels); the result -- parameters, Adam moments and
//     step counts, exist_since_iter_, zeroed statistics -- is the reference sequence's;
//   * the optimizer trainingSetup builds is an LgsFusedAdam (a torch::optim::Adam whose step is one launch);
//   * the allocator cache is not emptied after growth (the reference calls emptyCache(), :383,823: allocator hygiene).
#include "gaussian_model.h"

#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>

#include <algorithm>
#include <cmath>
#include <fstream>
#include <map>
#include <sstream>
#include <string>

#include "lgs.h"
#include "operate_points.h"
#include "spatial.h"

namespace {
const float SH_C0 = 0.28209479177387814f;  // include/sh_utils.h:32

void check(int status, const char* what) {
    TORCH_CHECK(status == LGS_OK, what, ": ", lgs_status_string(status), " (cudaError ", lgs_last_cuda_error(), ")");
}
void* stream() { return (void*)at::cuda::getCurrentCUDAStream().stream(); }

// libtorch 2.0/2.1 key the optimizer's state map by the printed TensorImpl address, later versions by the pointer itself
template <typename Map>
auto state_key(const Map&, const torch::Tensor& t) {
    using K = typename Map::key_type;
    if constexpr (std::is_same_v<K, std::string>) {
        std::ostringstream ss;
        ss << t.unsafeGetTensorImpl();
        return ss.str();
    } else {
        return static_cast<K>(t.unsafeGetTensorImpl());
    }
}
torch::optim::AdamParamState* find_state(torch::optim::Adam& opt, const torch::Tensor& p) {
    auto& states = opt.state();
    auto it = states.find(state_key(states, p));
    return it == states.end() ? nullptr : static_cast<torch::optim::AdamParamState*>(it->second.get());
}
// group `idx` of the optimizer gets `fresh` as its parameter; its Adam state becomes (step, m, v) when `with_state`
torch::Tensor swap_param(torch::optim::Adam& opt, int idx, torch::Tensor fresh, bool with_state, int64_t step, torch::Tensor m,
                         torch::Tensor v) {
    auto& p = opt.param_groups()[idx].params()[0];
    auto& states = opt.state();
    states.erase(state_key(states, p));
    p = fresh.detach().requires_grad_();
    if (with_state) {
        auto st = std::make_unique<torch::optim::AdamParamState>();
        st->step(step);
        st->exp_avg(m);
        st->exp_avg_sq(v);
        states[state_key(states, p)] = std::move(st);
    }
    return p;
}
torch::Tensor inverse_sigmoid(const torch::Tensor& x) { return torch::log(x / (1 - x)); }  // include/general_utils.h:30-33
}  // namespace

GaussianModel::GaussianModel(const int sh_degree)
    : active_sh_degree_(0), spatial_lr_scale_(0.0f), lr_delay_steps_(0), lr_delay_mult_(1.0f), max_steps_(1000000) {
    max_sh_degree_ = sh_degree;
    percent_dense_ = 0.01f;
    lr_init_ = lr_final_ = 0.0f;
    device_type_ = torch::kCUDA;  // no CPU path
    auto o = torch::TensorOptions().device(torch::kCPU);  // placeholders; every tensor is replaced by a CUDA tensor on first use
    xyz_ = torch::empty(0, o); features_dc_ = torch::empty(0, o); features_rest_ = torch::empty(0, o);
    language_features_ = torch::empty(0, o); scaling_ = torch::empty(0, o); rotation_ = torch::empty(0, o);
    opacity_ = torch::empty(0, o); max_radii2D_ = torch::empty(0, o); xyz_gradient_accum_ = torch::empty(0, o);
    denom_ = torch::empty(0, o); exist_since_iter_ = torch::empty(0, o.dtype(torch::kInt32));
    tensorsToVec();
}

void GaussianModel::tensorsToVec() {
    Tensor_vec_xyz_ = {xyz_};
    Tensor_vec_feature_dc_ = {features_dc_};
    Tensor_vec_feature_rest_ = {features_rest_};
    Tensor_vec_language_feature_ = {language_features_};
    Tensor_vec_opacity_ = {opacity_};
    Tensor_vec_scaling_ = {scaling_};
    Tensor_vec_rotation_ = {rotation_};
}

torch::Tensor& GaussianModel::param(int i) {
    switch (i) {
        case 0: return xyz_;
        case 1: return features_dc_;
        case 2: return features_rest_;
        case 3: return language_features_;
        case 4: return opacity_;
        case 5: return scaling_;
        case 6: return rotation_;
    }
    TORCH_CHECK(false, "tensor_idx must be 0 ... 6");
}

// ---- activations (:46-98)
torch::Tensor GaussianModel::getScalingActivation() { return torch::exp(scaling_); }
torch::Tensor GaussianModel::getRotationActivation() { return torch::nn::functional::normalize(rotation_); }
torch::Tensor GaussianModel::getXYZ() { return xyz_; }
torch::Tensor GaussianModel::getFeatures() { return torch::cat({features_dc_.clone(), features_rest_.clone()}, 1); }
torch::Tensor GaussianModel::getLanguageFeatures() { return language_features_.clone(); }
torch::Tensor GaussianModel::getOpacityActivation() { return torch::sigmoid(opacity_); }

torch::Tensor GaussianModel::getCovarianceActivation(int scaling_modifier) {
    // Sigma = (R S)(R S)^T with R from the UN-normalised quaternion divided by its norm (general_utils::build_rotation),
    // upper triangle in the order xx, xy, xz, yy, yz, zz
    auto q = rotation_ / torch::sqrt((rotation_ * rotation_).sum(1, /*keepdim=*/true));
    auto r = q.select(1, 0), x = q.select(1, 1), y = q.select(1, 2), z = q.select(1, 3);
    auto R = torch::stack({1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                           2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                           2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)}, 1).view({-1, 3, 3});
    auto L = R * (scaling_modifier * getScalingActivation()).unsqueeze(1);  // R @ diag(s): column j scaled by s_j
    auto cov = L.matmul(L.transpose(1, 2));
    return torch::stack({cov.select(1, 0).select(1, 0), cov.select(1, 0).select(1, 1), cov.select(1, 0).select(1, 2),
                         cov.select(1, 1).select(1, 1), cov.select(1, 1).select(1, 2), cov.select(1, 2).select(1, 2)}, 1);
}

void GaussianModel::oneUpShDegree() {
    if (active_sh_degree_ < max_sh_degree_) active_sh_degree_ += 1;
}
void GaussianModel::setShDegree(const int sh) { active_sh_degree_ = (sh > max_sh_degree_ ? max_sh_degree_ : sh); }

// ---- growth (:109-384)
void GaussianModel::createFromPcd(torch::Tensor& points, torch::Tensor& colors, torch::Tensor& lang_feats,
                                  const float spatial_lr_scale) {
    torch::NoGradGuard no_grad;
    TORCH_CHECK(points.is_cuda(), "leg_slam_b200 has no CPU path: tensors must live on a CUDA device");
    TORCH_CHECK(points.ndimension() == 2 && points.size(1) == 3 && colors.sizes() == points.sizes(),
                "points and colors must have dimensions (num_points, 3)");
    spatial_lr_scale_ = spatial_lr_scale;
    const int64_t n = points.size(0);
    auto f = torch::TensorOptions().dtype(torch::kFloat32).device(points.device());
    TORCH_CHECK(lang_feats.numel() == 0 || (lang_feats.ndimension() == 2 && lang_feats.size(0) == n && lang_feats.size(1) == LGS_LF_DIM),
                "lang_feats must have dimensions (num_points, ", LGS_LF_DIM, ")");
    torch::Tensor fused_point_cloud = points.to(f).contiguous().clone();
    torch::Tensor fused_color = (colors.to(f) - 0.5f) / SH_C0;  // sh_utils::RGB2SH
    const int64_t coef = (int64_t)(max_sh_degree_ + 1) * (max_sh_degree_ + 1);
    torch::Tensor dist2 = torch::clamp_min(distCUDA2(fused_point_cloud.clone()), 0.0000001);
    torch::Tensor scales = torch::log(torch::sqrt(dist2)).unsqueeze(1).repeat({1, 3});
    torch::Tensor rots = torch::zeros({n, 4}, f);
    rots.select(1, 0).fill_(1);
    xyz_ = fused_point_cloud.requires_grad_();
    features_dc_ = fused_color.unsqueeze(1).contiguous().requires_grad_();
    features_rest_ = torch::zeros({n, coef - 1, 3}, f).requires_grad_();
    language_features_ = (lang_feats.numel() == 0 ? torch::zeros({n, LGS_LF_DIM}, f) : lang_feats.to(f).contiguous().clone()).requires_grad_();
    scaling_ = scales.requires_grad_();
    rotation_ = rots.requires_grad_();
    opacity_ = inverse_sigmoid(0.1f * torch::ones({n, 1}, f)).requires_grad_();
    exist_since_iter_ = torch::zeros({n}, f.dtype(torch::kInt32));
    tensorsToVec();
    max_radii2D_ = torch::zeros({n}, f);
}

void GaussianModel::increasePcd(torch::Tensor& new_point_cloud, torch::Tensor& new_colors, const int iteration) {
    torch::NoGradGuard no_grad;
    const int64_t n = new_point_cloud.size(0);
    if (n == 0) return;
    TORCH_CHECK(new_point_cloud.is_cuda(), "leg_slam_b200 has no CPU path: tensors must live on a CUDA device");
    auto f = torch::TensorOptions().dtype(torch::kFloat32).device(new_point_cloud.device());
    const int64_t coef = (int64_t)(max_sh_degree_ + 1) * (max_sh_degree_ + 1);
    torch::Tensor new_xyz = new_point_cloud.to(f).contiguous();
    torch::Tensor new_features_dc = ((new_colors.to(f) - 0.5f) / SH_C0).unsqueeze(1).contiguous();
    torch::Tensor new_features_rest = torch::zeros({n, coef - 1, 3}, f);
    torch::Tensor new_language_features = torch::zeros({n, LGS_LF_DIM}, f);
    torch::Tensor dist2 = torch::clamp_min(distCUDA2(new_xyz.clone()), 0.0000001);
    torch::Tensor new_scaling = torch::log(torch::sqrt(dist2)).unsqueeze(1).repeat({1, 3});
    torch::Tensor new_rotation = torch::zeros({n, 4}, f);
    new_rotation.select(1, 0).fill_(1);
    torch::Tensor new_opacities = inverse_sigmoid(0.1f * torch::ones({n, 1}, f));
    torch::Tensor new_exist_since_iter = torch::full({n}, iteration, f.dtype(torch::kInt32));
    densificationPostfix(new_xyz, new_features_dc, new_features_rest, new_language_features, new_opacities, new_scaling,
                         new_rotation, new_exist_since_iter);
}

// ---- corrections (:387-481)
void GaussianModel::applyScaledTransformation(const float s, torch::Tensor& T) {
    torch::NoGradGuard no_grad;
    torch::Tensor new_xyz = xyz_.detach() * s;  // pt <- (s * R pt + t)
    transformPoints(new_xyz, T);
    torch::Tensor new_scaling = scaling_.detach() * s;  // as shipped: the LOG-scale parameter times s (:403)
    scaledTransformationPostfix(new_xyz, new_scaling);
}

void GaussianModel::scaledTransformationPostfix(torch::Tensor& new_xyz, torch::Tensor& new_scaling) {
    xyz_ = replaceTensorToOptimizer(new_xyz, 0);
    scaling_ = replaceTensorToOptimizer(new_scaling, 5);
    tensorsToVec();
}

void GaussianModel::scaledTransformVisiblePointsOfKeyframe(torch::Tensor& point_not_transformed_flags, torch::Tensor& diff_pose,
                                                           torch::Tensor& kf_world_view_transform,
                                                           torch::Tensor& kf_full_proj_transform, const int kf_creation_iter,
                                                           const int stable_num_iter_existence, int& num_transformed,
                                                           const float scale) {
    torch::NoGradGuard no_grad;
    torch::Tensor points = getXYZ().detach().clone();
    torch::Tensor rots = getRotationActivation().detach();
    torch::Tensor point_unstable_flags = torch::abs(exist_since_iter_ - kf_creation_iter) < stable_num_iter_existence;
    scaleAndTransformThenMarkVisiblePoints(points, rots, point_not_transformed_flags, point_unstable_flags, diff_pose,
                                           kf_world_view_transform, kf_full_proj_transform, num_transformed, scale);
    xyz_ = replaceTensorToOptimizer(points, 0);
    rotation_ = replaceTensorToOptimizer(rots, 6);
    tensorsToVec();
}

// ---- optimizer (:483-565, 1143-1157)
void GaussianModel::trainingSetup(const GaussianOptimizationParams& a) {
    setPercentDense(a.percent_dense_);
    auto f = torch::TensorOptions().dtype(torch::kFloat32).device(xyz_.device());
    xyz_gradient_accum_ = torch::zeros({xyz_.size(0), 1}, f);
    denom_ = torch::zeros({xyz_.size(0), 1}, f);
    tensorsToVec();
    auto group = [](std::vector<torch::Tensor>& v, double lr) {
        return torch::optim::OptimizerParamGroup(v, std::make_unique<torch::optim::AdamOptions>(torch::optim::AdamOptions(lr).eps(1e-15)));
    };
    std::vector<torch::optim::OptimizerParamGroup> groups;
    groups.push_back(group(Tensor_vec_xyz_, a.position_lr_init_ * spatial_lr_scale_));
    groups.push_back(group(Tensor_vec_feature_dc_, a.feature_lr_));
    groups.push_back(group(Tensor_vec_feature_rest_, a.feature_lr_ / 20.0));
    groups.push_back(group(Tensor_vec_language_feature_, a.language_feature_lr_));
    groups.push_back(group(Tensor_vec_opacity_, a.opacity_lr_));
    groups.push_back(group(Tensor_vec_scaling_, a.scaling_lr_));
    groups.push_back(group(Tensor_vec_rotation_, a.rotation_lr_));
    optimizer_ = std::make_shared<LgsFusedAdam>(groups, torch::optim::AdamOptions(0.0).eps(1e-15));
    lr_init_ = a.position_lr_init_ * spatial_lr_scale_;
    lr_final_ = a.position_lr_final_ * spatial_lr_scale_;
    lr_delay_mult_ = a.position_lr_delay_mult_;
    max_steps_ = a.position_lr_max_steps_;
}

namespace {
void set_lr(torch::optim::Adam& opt, int idx, double lr) {
    static_cast<torch::optim::AdamOptions&>(opt.param_groups()[idx].options()).lr(lr);
}
}  // namespace

float GaussianModel::updateLearningRate(int step) {
    const float lr = exponLrFunc(step);
    set_lr(*optimizer_, 0, lr);
    return lr;
}
void GaussianModel::setPositionLearningRate(float position_lr) { set_lr(*optimizer_, 0, position_lr * spatial_lr_scale_); }
void GaussianModel::setFeatureLearningRate(float feature_lr) {
    set_lr(*optimizer_, 1, feature_lr);
    set_lr(*optimizer_, 2, feature_lr / 20.0);
}
void GaussianModel::setLanguageFeatureLearningRate(float lang_feat_lr) { set_lr(*optimizer_, 3, lang_feat_lr); }
void GaussianModel::setOpacityLearningRate(float opacity_lr) { set_lr(*optimizer_, 4, opacity_lr); }
void GaussianModel::setScalingLearningRate(float scaling_lr) { set_lr(*optimizer_, 5, scaling_lr); }
void GaussianModel::setRotationLearningRate(float rot_lr) { set_lr(*optimizer_, 6, rot_lr); }

float GaussianModel::exponLrFunc(int step) {
    if (step < 0 || (lr_init_ == 0.0f && lr_final_ == 0.0f)) return 0.0f;
    float delay_rate = 1.0f;
    if (lr_delay_steps_ > 0)
        delay_rate = lr_delay_mult_ + (1.0f - lr_delay_mult_) *
                                          std::sin(1.57079632679489661923f * std::clamp(static_cast<float>(step) / lr_delay_steps_, 0.0f, 1.0f));
    const float t = std::clamp(static_cast<float>(step) / max_steps_, 0.0f, 1.0f);
    const float log_lerp = std::exp(std::log(lr_init_) * (1 - t) + std::log(lr_final_) * t);
    return delay_rate * log_lerp;
}

// ---- optimizer-state surgery (:567-727)
void GaussianModel::resetOpacity() {
    torch::NoGradGuard no_grad;
    // min(sigmoid(opacity), ones): the reference clamps against ones_like(x * 0.01) = 1, i.e. not at all (:567-571)
    torch::Tensor act = getOpacityActivation();
    torch::Tensor opacities_new = inverse_sigmoid(torch::min(act, torch::ones_like(act * 0.01)));
    opacity_ = replaceTensorToOptimizer(opacities_new, 4);
    tensorsToVec();
}

torch::Tensor GaussianModel::replaceTensorToOptimizer(torch::Tensor& t, int tensor_idx) {
    torch::NoGradGuard no_grad;
    TORCH_CHECK(optimizer_ != nullptr, "call trainingSetup first");
    auto& old = optimizer_->param_groups()[tensor_idx].params()[0];
    auto* st = find_state(*optimizer_, old);
    const int64_t step = st ? st->step() : 0;  // (the reference dereferences a missing state; here the group simply has none)
    return swap_param(*optimizer_, tensor_idx, t, st != nullptr, step, torch::zeros_like(t), torch::zeros_like(t));
}

void GaussianModel::prunePoints(torch::Tensor& mask) {
    torch::NoGradGuard no_grad;
    torch::Tensor valid = ~mask;
    for (int i = 0; i < 7; ++i) {
        auto& old = optimizer_->param_groups()[i].params()[0];
        auto* st = find_state(*optimizer_, old);
        torch::Tensor m, v;
        int64_t step = 0;
        if (st) {
            step = st->step();
            m = st->exp_avg().index({valid});
            v = st->exp_avg_sq().index({valid});
        }
        torch::Tensor kept = old.detach().index({valid});
        param(i) = swap_param(*optimizer_, i, kept, st != nullptr, step, m, v);
    }
    tensorsToVec();
    exist_since_iter_ = exist_since_iter_.index({valid});
    xyz_gradient_accum_ = xyz_gradient_accum_.index({valid});
    denom_ = denom_.index({valid});
    max_radii2D_ = max_radii2D_.index({valid});
}

void GaussianModel::densificationPostfix(torch::Tensor& new_xyz, torch::Tensor& new_features_dc, torch::Tensor& new_features_rest,
                                         torch::Tensor& new_language_features, torch::Tensor& new_opacities,
                                         torch::Tensor& new_scaling, torch::Tensor& new_rotation,
                                         torch::Tensor& new_exist_since_iter) {
    torch::NoGradGuard no_grad;
    std::vector<torch::Tensor> ext = {new_xyz, new_features_dc, new_features_rest, new_language_features, new_opacities,
                                      new_scaling, new_rotation};
    for (int i = 0; i < 7; ++i) {
        auto& old = optimizer_->param_groups()[i].params()[0];
        auto* st = find_state(*optimizer_, old);
        torch::Tensor m, v;
        int64_t step = 0;
        if (st) {
            step = st->step();
            m = torch::cat({st->exp_avg(), torch::zeros_like(ext[i])}, 0);
            v = torch::cat({st->exp_avg_sq(), torch::zeros_like(ext[i])}, 0);
        }
        torch::Tensor grown = torch::cat({old.detach(), ext[i]}, 0);
        param(i) = swap_param(*optimizer_, i, grown, st != nullptr, step, m, v);
    }
    tensorsToVec();
    exist_since_iter_ = torch::cat({exist_since_iter_, new_exist_since_iter}, 0);
    auto f = torch::TensorOptions().dtype(torch::kFloat32).device(xyz_.device());
    xyz_gradient_accum_ = torch::zeros({xyz_.size(0), 1}, f);
    denom_ = torch::zeros({xyz_.size(0), 1}, f);
    max_radii2D_ = torch::zeros({xyz_.size(0)}, f);
}

// ---- density control (:729-847)
void GaussianModel::densifyAndPrune(float max_grad, float min_opacity, float extent, int max_screen_size) {
    torch::NoGradGuard no_grad;
    TORCH_CHECK(optimizer_ != nullptr, "call trainingSetup first");
    TORCH_CHECK(xyz_.is_cuda(), "leg_slam_b200 has no CPU path: tensors must live on a CUDA device");
    const c10::cuda::CUDAGuard guard(xyz_.device());
    const int P = (int)xyz_.size(0);
    auto f = torch::TensorOptions().dtype(torch::kFloat32).device(xyz_.device());
    // current tensors: parameters and their moments (zeros where the optimizer has not stepped yet)
    std::vector<torch::Tensor> cur_p(7), cur_m(7), cur_v(7);
    std::vector<int64_t> steps(7, 0);
    std::vector<bool> has_state(7, false);
    for (int i = 0; i < 7; ++i) {
        cur_p[i] = param(i).detach().contiguous();
        if (auto* st = find_state(*optimizer_, optimizer_->param_groups()[i].params()[0])) {
            has_state[i] = true;
            steps[i] = st->step();
            cur_m[i] = st->exp_avg().contiguous();
            cur_v[i] = st->exp_avg_sq().contiguous();
        } else {
            cur_m[i] = torch::zeros_like(cur_p[i]);
            cur_v[i] = torch::zeros_like(cur_p[i]);
        }
    }
    torch::Tensor accum = xyz_gradient_accum_.contiguous(), den = denom_.contiguous();
    torch::Tensor plan = torch::empty({(int64_t)lgs_densify_plan_bytes(P)}, f.dtype(torch::kByte));
    int totals[4] = {0, 0, 0, 0};
    check(lgs_densify_plan(P, accum.data_ptr<float>(), den.data_ptr<float>(), cur_p[5].data_ptr<float>(), cur_p[4].data_ptr<float>(),
                           max_grad, min_opacity, extent, percent_dense_, max_screen_size, (char*)plan.data_ptr<uint8_t>(), totals,
                           stream()),
          "lgs_densify_plan");
    const int64_t nS = totals[3], newP = (int64_t)totals[0] + totals[1] + 2 * (int64_t)totals[2];
    torch::Tensor samples = nS > 0 ? torch::randn({2 * nS, 3}, f) : torch::empty({0, 3}, f);  // densifyAndSplit's normal draws
    std::vector<const float*> src;
    std::vector<float*> dst;
    std::vector<int> rows, modes;
    std::vector<torch::Tensor> out_p(7), out_m(7), out_v(7);
    for (int i = 0; i < 7; ++i) {
        auto shape = cur_p[i].sizes().vec();
        shape[0] = newP;
        out_p[i] = torch::empty(shape, f);
        out_m[i] = torch::empty(shape, f);
        out_v[i] = torch::empty(shape, f);
        const int row = P > 0 ? (int)(cur_p[i].numel() / P) : 1;
        const int mode = i == 0 ? 2 : (i == 5 ? 3 : 0);  // xyz: children are sampled; scaling: children shrink by 1.6
        src.push_back(cur_p[i].data_ptr<float>()); dst.push_back(out_p[i].data_ptr<float>()); rows.push_back(row); modes.push_back(mode);
        src.push_back(cur_m[i].data_ptr<float>()); dst.push_back(out_m[i].data_ptr<float>()); rows.push_back(row); modes.push_back(1);
        src.push_back(cur_v[i].data_ptr<float>()); dst.push_back(out_v[i].data_ptr<float>()); rows.push_back(row); modes.push_back(1);
    }
    torch::Tensor old_exist = exist_since_iter_.contiguous();
    torch::Tensor new_exist = torch::empty({newP}, f.dtype(torch::kInt32));
    src.push_back(reinterpret_cast<const float*>(old_exist.data_ptr<int>()));
    dst.push_back(reinterpret_cast<float*>(new_exist.data_ptr<int>()));
    rows.push_back(1);
    modes.push_back(0);
    if (newP > 0) {
        torch::Tensor scratch = torch::empty({2 * newP}, f.dtype(torch::kInt32));
        check(lgs_densify_apply(P, (const char*)plan.data_ptr<uint8_t>(), totals, (int)src.size(), src.data(), dst.data(), rows.data(),
                                modes.data(), cur_p[5].data_ptr<float>(), cur_p[6].data_ptr<float>(),
                                nS > 0 ? samples.data_ptr<float>() : nullptr, reinterpret_cast<uint32_t*>(scratch.data_ptr<int>()),
                                stream()),
              "lgs_densify_apply");
    }
    for (int i = 0; i < 7; ++i) param(i) = swap_param(*optimizer_, i, out_p[i], has_state[i], steps[i], out_m[i], out_v[i]);
    tensorsToVec();
    exist_since_iter_ = new_exist;
    xyz_gradient_accum_ = torch::zeros({newP, 1}, f);  // densificationPostfix zeroes the statistics at the new size (:723-725)
    denom_ = torch::zeros({newP, 1}, f);
    max_radii2D_ = torch::zeros({newP}, f);
}

void GaussianModel::addDensificationStats(torch::Tensor& viewspace_point_tensor, torch::Tensor& update_filter) {
    torch::NoGradGuard no_grad;
    xyz_gradient_accum_.index_put_({update_filter},
                                   viewspace_point_tensor.grad().index({update_filter, torch::indexing::Slice(0, 2)})
                                       .norm(2, std::vector<int64_t>{-1}, /*keepdim=*/true),
                                   /*accumulate=*/true);
    denom_.index_put_({update_filter}, denom_.index({update_filter}) + 1);
}

float GaussianModel::percentDense() { return percent_dense_; }
void GaussianModel::setPercentDense(const float percent_dense) { percent_dense_ = percent_dense; }

// ---- checkpoints (:854-1075; file format of leg_slam_b200/ply_io.py = what tinyply writes for the reference)
namespace {
struct PlyColumn {
    std::string name;
    int tensor;  // index into the tensor list handed to the kernel, -1 = no tensor (normals: zeros / skipped)
    int elem;
};
const char* const kGroupNames[7] = {"xyz", "features_dc", "features_rest", "lang_feat", "opacity", "scaling", "rotation"};

// tensors 0-6 = the parameters in optimizer order, 7-13 = exp_avg, 14-20 = exp_avg_sq
std::vector<PlyColumn> ply_columns(int n_rest, int n_lf, bool with_adam) {
    std::vector<PlyColumn> c = {{"x", 0, 0}, {"y", 0, 1}, {"z", 0, 2}, {"nx", -1, 0}, {"ny", -1, 0}, {"nz", -1, 0}};
    for (int k = 0; k < 3; ++k) c.push_back({"f_dc_" + std::to_string(k), 1, k});
    for (int i = 0; i < 3 * n_rest; ++i) c.push_back({"f_rest_" + std::to_string(i), 2, 3 * (i % n_rest) + i / n_rest});  // channel-major
    for (int i = 0; i < n_lf; ++i) c.push_back({"lf_" + std::to_string(i), 3, i});
    c.push_back({"opacity", 4, 0});
    for (int i = 0; i < 3; ++i) c.push_back({"scale_" + std::to_string(i), 5, i});
    for (int i = 0; i < 4; ++i) c.push_back({"rot_" + std::to_string(i), 6, i});
    if (with_adam) {
        const int rows[7] = {3, 3, 3 * n_rest, n_lf, 1, 3, 4};
        for (int which = 0; which < 2; ++which) {
            int i = 0;
            for (int t = 0; t < 7; ++t)
                for (int e = 0; e < rows[t]; ++e) c.push_back({std::string(which ? "adam_v_" : "adam_m_") + std::to_string(i++), 7 + 7 * which + t, e});
        }
    }
    return c;
}

void ply_run(bool pack, int64_t P, const std::vector<PlyColumn>& cols, std::vector<torch::Tensor>& tensors, torch::Tensor& block) {
    const c10::cuda::CUDAGuard guard(block.device());
    std::vector<int> col_t, col_e, rows;
    for (auto& c : cols) {
        col_t.push_back(c.tensor >= 0 && c.tensor < (int)tensors.size() && tensors[c.tensor].defined() ? c.tensor : -1);
        col_e.push_back(c.elem);
    }
    std::vector<float*> ptrs;
    for (auto& t : tensors) {
        ptrs.push_back(t.defined() ? t.data_ptr<float>() : nullptr);
        rows.push_back(t.defined() && P > 0 ? (int)(t.numel() / P) : 1);
    }
    auto i32 = torch::TensorOptions().dtype(torch::kInt32);
    torch::Tensor d_t = torch::tensor(col_t, i32).to(block.device()), d_e = torch::tensor(col_e, i32).to(block.device());
    const int st = pack ? lgs_ply_pack(P, (int)cols.size(), d_t.data_ptr<int>(), d_e.data_ptr<int>(), (int)ptrs.size(), ptrs.data(),
                                       rows.data(), block.data_ptr<float>(), stream())
                        : lgs_ply_unpack(P, (int)cols.size(), d_t.data_ptr<int>(), d_e.data_ptr<int>(), (int)ptrs.size(), ptrs.data(),
                                         rows.data(), block.data_ptr<float>(), stream());
    check(st, pack ? "lgs_ply_pack" : "lgs_ply_unpack");
}
}  // namespace

void GaussianModel::savePly(std::filesystem::path result_path, bool with_optimizer_state) {
    torch::NoGradGuard no_grad;
    TORCH_CHECK(xyz_.is_cuda(), "leg_slam_b200 has no CPU path: tensors must live on a CUDA device");
    const int64_t P = xyz_.size(0);
    const int n_rest = (int)features_rest_.size(1), n_lf = (int)language_features_.size(1);
    std::vector<torch::Tensor> tensors;
    for (int i = 0; i < 7; ++i) tensors.push_back(param(i).detach().contiguous());
    std::vector<int64_t> steps(7, 0);
    const bool with_adam = with_optimizer_state && optimizer_ != nullptr;
    if (with_adam) {
        for (int which = 0; which < 2; ++which)
            for (int i = 0; i < 7; ++i) {
                auto* st = find_state(*optimizer_, optimizer_->param_groups()[i].params()[0]);
                steps[i] = st ? st->step() : 0;
                tensors.push_back(st ? (which ? st->exp_avg_sq() : st->exp_avg()).contiguous() : torch::zeros_like(tensors[i]));
            }
    }
    auto cols = ply_columns(n_rest, n_lf, with_adam);
    torch::Tensor block = torch::empty({P, (int64_t)cols.size()}, xyz_.options().dtype(torch::kFloat32));
    ply_run(true, P, cols, tensors, block);
    torch::Tensor host = block.cpu();
    std::ofstream f(result_path, std::ios::binary);
    TORCH_CHECK(f.good(), "cannot open ", result_path.string(), " for writing");
    std::ostringstream h;
    h << "ply\nformat binary_little_endian 1.0\n";
    if (with_adam)
        for (int i = 0; i < 7; ++i) h << "comment lgs_adam_step " << kGroupNames[i] << " " << steps[i] << "\n";
    h << "element vertex " << P << "\n";
    for (auto& c : cols) h << "property float " << c.name << "\n";
    h << "end_header\n";
    const std::string hs = h.str();
    f.write(hs.data(), (std::streamsize)hs.size());
    f.write(reinterpret_cast<const char*>(host.data_ptr<float>()), (std::streamsize)(host.numel() * sizeof(float)));
    TORCH_CHECK(f.good(), "write to ", result_path.string(), " failed");
}

void GaussianModel::loadPly(std::filesystem::path ply_path) {
    torch::NoGradGuard no_grad;
    std::ifstream f(ply_path, std::ios::binary);
    TORCH_CHECK(f.good(), "cannot open ", ply_path.string());
    std::string line;
    std::getline(f, line);
    TORCH_CHECK(line.rfind("ply", 0) == 0, "not a ply file");
    int64_t P = -1;
    bool in_vertex = false;
    std::vector<std::string> props;
    std::map<std::string, int64_t> step_of;
    for (;;) {
        TORCH_CHECK((bool)std::getline(f, line), "ply header not terminated");
        std::istringstream ss(line);
        std::string tok, a, b, c;
        ss >> tok;
        if (tok == "format") {
            ss >> a;
            TORCH_CHECK(a == "binary_little_endian", "only binary_little_endian ply files are supported");
        } else if (tok == "comment") {
            ss >> a >> b >> c;
            if (a == "lgs_adam_step" && !c.empty()) step_of[b] = std::stoll(c);
        } else if (tok == "element") {
            ss >> a >> b;
            in_vertex = a == "vertex";
            if (in_vertex) P = std::stoll(b);
        } else if (tok == "property" && in_vertex) {
            ss >> a >> b;
            TORCH_CHECK(a == "float" || a == "float32", "property ", b, ": only float32 vertex properties are supported");
            props.push_back(b);
        } else if (tok == "end_header") {
            break;
        }
    }
    TORCH_CHECK(P >= 0, "ply file has no vertex element");
    const int n_rest = (max_sh_degree_ + 1) * (max_sh_degree_ + 1) - 1;
    int n_have = 0, n_lf = 0;
    bool with_adam = false;
    for (auto& p : props) {
        n_have += p.rfind("f_rest_", 0) == 0;
        n_lf += p.rfind("lf_", 0) == 0;
        with_adam = with_adam || p.rfind("adam_m_", 0) == 0;
    }
    TORCH_CHECK(n_have == 3 * n_rest, "file holds ", n_have, " f_rest properties, max_sh_degree=", max_sh_degree_, " needs ", 3 * n_rest);
    torch::Tensor host = torch::empty({P, (int64_t)props.size()}, torch::kFloat32);
    f.read(reinterpret_cast<char*>(host.data_ptr<float>()), (std::streamsize)(host.numel() * sizeof(float)));
    TORCH_CHECK(f.gcount() == (std::streamsize)(host.numel() * sizeof(float)), "ply file is shorter than its header says");
    std::map<std::string, PlyColumn> want;
    for (auto& c : ply_columns(n_rest, n_lf, with_adam)) want[c.name] = c;
    std::vector<PlyColumn> cols;
    for (auto& p : props) {  // unknown properties are skipped, like the reference does
        auto it = want.find(p);
        cols.push_back(it == want.end() ? PlyColumn{p, -1, 0} : it->second);
    }
    auto f32 = torch::TensorOptions().dtype(torch::kFloat32).device(torch::kCUDA);
    const std::vector<std::vector<int64_t>> shapes = {{P, 3}, {P, 1, 3}, {P, n_rest, 3}, {P, n_lf}, {P, 1}, {P, 3}, {P, 4}};
    std::vector<torch::Tensor> tensors;
    for (int rep = 0; rep < (with_adam ? 3 : 1); ++rep)
        for (int i = 0; i < 7; ++i) tensors.push_back(torch::zeros(shapes[i], f32));
    torch::Tensor block = host.to(torch::kCUDA);
    ply_run(false, P, cols, tensors, block);
    if (optimizer_ != nullptr) {  // a model that is being trained: the groups take the loaded tensors (and state)
        for (int i = 0; i < 7; ++i) {
            auto it = step_of.find(kGroupNames[i]);
            param(i) = swap_param(*optimizer_, i, tensors[i], with_adam, it == step_of.end() ? 0 : it->second,
                                  with_adam ? tensors[7 + i] : torch::Tensor(), with_adam ? tensors[14 + i] : torch::Tensor());
        }
    } else {
        for (int i = 0; i < 7; ++i) param(i) = tensors[i].requires_grad_();
    }
    tensorsToVec();
    active_sh_degree_ = max_sh_degree_;  // :967
    exist_since_iter_ = torch::zeros({P}, f32.dtype(torch::kInt32));
    max_radii2D_ = torch::zeros({P}, f32);
    xyz_gradient_accum_ = torch::zeros({P, 1}, f32);
    denom_ = torch::zeros({P, 1}, f32);
}
