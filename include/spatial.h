/*
 * spatial.h -- the reference's simple-knn entry point, re-declared with the identical signature (reference
 * third_party/simple-knn/spatial.h:14; implementation spatial.cu:15-27 on simple_knn.cu:185-220) and implemented on liblgs
 * (lgs_knn_mean_dist2, include/lgs.h) in leg_slam_b200/csrc/host/geometry_ops.cpp.  Callers -- GaussianModel::createFromPcd /
 * increasePcd (src/gaussian_model.cpp:157,242,331) -- compile unchanged.
 */
#pragma once
#include <torch/torch.h>

/* points [P,3] -> [P]: mean squared distance of every point to its three nearest neighbours. */
torch::Tensor distCUDA2(const torch::Tensor &points);
