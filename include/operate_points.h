/*
 * operate_points.h -- the reference's libtorch point operators, re-declared with identical signatures (reference
 * include/operate_points.h:27-40; implementation src/operate_points.cu:72-140) and implemented on liblgs (include/lgs.h) in
 * leg_slam_b200/csrc/host/geometry_ops.cpp.  Callers -- GaussianModel::scaledTransformVisiblePointsOfKeyframe
 * (src/gaussian_model.cpp:420-452), the keyframe-ingest code of GaussianMapper (src/gaussian_mapper.cpp:1289-1300) --
 * compile unchanged.
 */
#pragma once
#include <torch/torch.h>

/* points [P,3] <- transformPoint4x3(points, transformmatrix); `points` is rebound to the transformed copy. */
void transformPoints(torch::Tensor &points, torch::Tensor &transformmatrix);

/* In place: the rows visible from `viewmatrix` (markVisible), flagged in both masks, get
 * points <- T (scale * points), rots <- quaternion of T[:3,:3] R(rots), and their not-transformed flag cleared;
 * num_transformed is incremented by their number. */
void scaleAndTransformThenMarkVisiblePoints(torch::Tensor &points, torch::Tensor &rots,
                                            torch::Tensor &point_not_transformed_mask, torch::Tensor &point_unstable_mask,
                                            torch::Tensor &transformmatrix, torch::Tensor &viewmatrix,
                                            torch::Tensor &projmatrix, int &num_transformed, const float scale = 1.0f);
