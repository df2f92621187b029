/*
 * lgs_adam.h -- drop-in for the optimizer GaussianModel::trainingSetup builds (reference src/gaussian_model.cpp:483-518:
 * `optimizer_.reset(new torch::optim::Adam(...))` with 7 single-tensor groups, stepped at src/gaussian_mapper.cpp:793-796).
 *
 * LgsFusedAdam IS a torch::optim::Adam -- same param_groups(), same AdamParamState (step, exp_avg, exp_avg_sq) per
 * parameter, same options -- so the reference's optimizer-state surgery in densificationPostfix / prunePoints /
 * replaceTensorToOptimizer (src/gaussian_model.cpp:577-727) keeps working on it unchanged; only step() differs: ONE
 * lgs_adam_multi launch (include/lgs.h) over all groups that share (betas, eps, step) instead of ~8 ATen kernels per
 * tensor.  weight_decay and amsgrad are not used by the reference and are refused.
 */
#pragma once
#include <torch/torch.h>

class LgsFusedAdam : public torch::optim::Adam {
public:
    using torch::optim::Adam::Adam;
    torch::Tensor step(LossClosure closure = nullptr) override;
};
