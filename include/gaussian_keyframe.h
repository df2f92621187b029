/*
 * gaussian_keyframe.h -- the fields of the reference's GaussianKeyframe (include/gaussian_keyframe.h, src/gaussian_keyframe.cpp)
 * that the mapping path reads: the field of view, the three camera tensors computeTransformTensors leaves
 * (src/gaussian_keyframe.cpp:111-137: world_view_transform_ and full_proj_transform_ stored TRANSPOSED, camera_center_), the
 * image size and the keyframe's low-resolution language-feature map (src/gaussian_mapper.cpp:707-708).  The pose / image
 * bookkeeping of the reference class needs Eigen, Sophus and OpenCV (absent here) and is not on the path.
 */
#pragma once
#include <torch/torch.h>

class GaussianKeyframe {
public:
    float FoVx_ = 0.0f, FoVy_ = 0.0f;
    int image_height_ = 0, image_width_ = 0;
    torch::Tensor world_view_transform_, full_proj_transform_, camera_center_;
    torch::Tensor language_features_;  // [64, h, w], resized to the render size per iteration
    int creation_iter_ = 0;
};
