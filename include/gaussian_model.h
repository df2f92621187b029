/*
 * gaussian_model.h -- the reference's GaussianModel (include/gaussian_model.h:62-236, src/gaussian_model.cpp) for the part of
 * it that sits on and beside the mapping path, re-declared with the reference's member names and method signatures and
 * implemented on liblgs (include/lgs.h) in leg_slam_b200/csrc/host/gaussian_model.cpp:
 *
 *   activations              getXYZ / getFeatures / getLanguageFeatures / getOpacityActivation / getScalingActivation /
 *                            getRotationActivation / getCovarianceActivation                      (:46-98)
 *   per-iteration settings   oneUpShDegree / setShDegree, trainingSetup, updateLearningRate (exponLrFunc), set*LearningRate
 *                                                                                                  (:100-107,483-565,1143-1157)
 *   the optimizer            optimizer_ is a torch::optim::Adam -- here an LgsFusedAdam (include/lgs_adam.h): one launch per step
 *   optimizer-state surgery  replaceTensorToOptimizer, resetOpacity, prunePoints, densificationPostfix  (:567-727)
 *   density control          addDensificationStats, densifyAndPrune -- the latter as ONE classification pass + ONE gather of
 *                            the 7 parameters, their 14 Adam moments and exist_since_iter_ (lgs_densify_plan / _apply)
 *                            instead of densifyAndClone + densifyAndSplit + prunePoints                (:729-847)
 *   growth and correction    createFromPcd, increasePcd, applyScaledTransformation, scaledTransformationPostfix,
 *                            scaledTransformVisiblePointsOfKeyframe                                  (:109-481)
 *
 * Signatures that name Eigen / Sophus / colmap-style types (absent from this image) take tensors instead and say so:
 * createFromPcd(points, colors, lang_feats, spatial_lr_scale) for the std::map<point3D_id_t, Point3D> overload,
 * applyScaledTransformation(s, T) with T the pose tensor stored transposed for the Sophus::SE3f overload.
 *   checkpoints              savePly / loadPly (:854-1075) without tinyply: the same binary_little_endian file (property
 *                            order x y z nx ny nz f_dc_* f_rest_* lf_* opacity scale_* rot_*, SH stored channel-major), the
 *                            interleaving on the GPU (lgs_ply_pack / lgs_ply_unpack); loadPly also restores the language
 *                            features (the reference's loader drops them) and, with savePly(path, true), the Adam state rides
 *                            along as extra properties (adam_m_* / adam_v_*, step counts as header comments) -- the format of
 *                            leg_slam_b200.ply_io, byte for byte.
 * Not here: densifyAndClone / densifyAndSplit as separate steps (fused into densifyAndPrune) and saveSparsePointsPly.
 */
#pragma once
#include <torch/torch.h>

#include <filesystem>
#include <memory>
#include <vector>

#include "lgs_adam.h"

/* reference include/gaussian_parameters.h:55-95, the fields GaussianModel reads */
struct GaussianOptimizationParams {
    int iterations_ = 30'000;
    float position_lr_init_ = 0.00016f;
    float position_lr_final_ = 0.0000016f;
    float position_lr_delay_mult_ = 0.01f;
    int position_lr_max_steps_ = 30'000;
    float feature_lr_ = 0.0025f;
    float language_feature_lr_ = 0.0015f;
    float opacity_lr_ = 0.05f;
    float scaling_lr_ = 0.005f;
    float rotation_lr_ = 0.001f;
    float percent_dense_ = 0.01f;
    float lambda_dssim_ = 0.2f;
    int densification_interval_ = 100;
    int opacity_reset_interval_ = 3000;
    int densify_from_iter_ = 500;
    int densify_until_iter_ = 15'000;
    float densify_grad_threshold_ = 0.0002f;
};

class GaussianModel {
public:
    explicit GaussianModel(const int sh_degree);

    torch::Tensor getScalingActivation();
    torch::Tensor getRotationActivation();
    torch::Tensor getXYZ();
    torch::Tensor getFeatures();
    torch::Tensor getLanguageFeatures();
    torch::Tensor getOpacityActivation();
    torch::Tensor getCovarianceActivation(int scaling_modifier = 1);

    void oneUpShDegree();
    void setShDegree(const int sh);

    /* points / colors [n,3], lang_feats [n,64] or empty (zeros) */
    void createFromPcd(torch::Tensor &points, torch::Tensor &colors, torch::Tensor &lang_feats, const float spatial_lr_scale);
    void increasePcd(torch::Tensor &new_point_cloud, torch::Tensor &new_colors, const int iteration);

    void applyScaledTransformation(const float s, torch::Tensor &T);
    void scaledTransformationPostfix(torch::Tensor &new_xyz, torch::Tensor &new_scaling);
    void scaledTransformVisiblePointsOfKeyframe(torch::Tensor &point_not_transformed_flags, torch::Tensor &diff_pose,
                                                torch::Tensor &kf_world_view_transform, torch::Tensor &kf_full_proj_transform,
                                                const int kf_creation_iter, const int stable_num_iter_existence,
                                                int &num_transformed, const float scale = 1.0f);

    void trainingSetup(const GaussianOptimizationParams &training_args);
    float updateLearningRate(int step);
    void setPositionLearningRate(float position_lr);
    void setFeatureLearningRate(float feature_lr);
    void setLanguageFeatureLearningRate(float lang_feat_lr);
    void setOpacityLearningRate(float opacity_lr);
    void setScalingLearningRate(float scaling_lr);
    void setRotationLearningRate(float rot_lr);

    void resetOpacity();
    torch::Tensor replaceTensorToOptimizer(torch::Tensor &t, int tensor_idx);
    void prunePoints(torch::Tensor &mask);
    void densificationPostfix(torch::Tensor &new_xyz, torch::Tensor &new_features_dc, torch::Tensor &new_features_rest,
                              torch::Tensor &new_language_features, torch::Tensor &new_opacities, torch::Tensor &new_scaling,
                              torch::Tensor &new_rotation, torch::Tensor &new_exist_since_iter);
    void densifyAndPrune(float max_grad, float min_opacity, float extent, int max_screen_size);
    void addDensificationStats(torch::Tensor &viewspace_point_tensor, torch::Tensor &update_filter);

    void loadPly(std::filesystem::path ply_path);
    void savePly(std::filesystem::path result_path, bool with_optimizer_state = false);

    float percentDense();
    void setPercentDense(const float percent_dense);

protected:
    float exponLrFunc(int step);
    torch::Tensor &param(int tensor_idx);  // xyz_, features_dc_, ... in the optimizer's group order
    void tensorsToVec();

public:
    torch::DeviceType device_type_;

    int active_sh_degree_;
    int max_sh_degree_;

    torch::Tensor xyz_;
    torch::Tensor features_dc_;
    torch::Tensor features_rest_;
    torch::Tensor language_features_;
    torch::Tensor scaling_;
    torch::Tensor rotation_;
    torch::Tensor opacity_;
    torch::Tensor max_radii2D_;
    torch::Tensor xyz_gradient_accum_;
    torch::Tensor denom_;
    torch::Tensor exist_since_iter_;

    std::vector<torch::Tensor> Tensor_vec_xyz_, Tensor_vec_feature_dc_, Tensor_vec_feature_rest_, Tensor_vec_language_feature_,
        Tensor_vec_opacity_, Tensor_vec_scaling_, Tensor_vec_rotation_;

    std::shared_ptr<torch::optim::Adam> optimizer_;
    float percent_dense_;
    float spatial_lr_scale_;

protected:
    float lr_init_;
    float lr_final_;
    int lr_delay_steps_;
    float lr_delay_mult_;
    int max_steps_;
};
