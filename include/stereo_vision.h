/*
 * stereo_vision.h -- the reference's libtorch depth / keypoint operators, re-declared with identical signatures (reference
 * include/stereo_vision.h:25-40; implementation src/stereo_vision.cu:135-212) and implemented on liblgs (include/lgs.h) in
 * leg_slam_b200/csrc/host/geometry_ops.cpp.  Callers -- GaussianMapper's keyframe ingest
 * (src/gaussian_mapper.cpp:1265-1300,1430-1456) -- compile unchanged.
 */
#pragma once
#include <torch/torch.h>

#include <tuple>
#include <vector>

/* depth [P] (row-major image), mask [P] bool, intr = {fx, fy, cx, cy} -> points [P,3], zeros where the mask is false. */
torch::Tensor reprojectDepthPinhole(torch::Tensor &depth, torch::Tensor &mask, std::vector<float> &intr, int width);

/* <0> points, <1> colours of the keypoints that end with a positive depth, in keypoint order. */
std::tuple<torch::Tensor, torch::Tensor>
monocularPinholeInactiveGeoDensifyBySearchingNeighborhoodKeypoints(torch::Tensor &kps_pixel, torch::Tensor &kps_has3D,
                                                                   torch::Tensor &kps_point_local, torch::Tensor &colors,
                                                                   float max_pixel_dist, std::vector<float> &intr, int width);
