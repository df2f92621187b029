// fused_adam.cpp -- LgsFusedAdam::step (include/lgs_adam.h): libtorch's Adam bookkeeping, liblgs' fused update.
#include "lgs_adam.h"

#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>

#include <map>
#include <sstream>
#include <string>
#include <tuple>
#include <vector>

#include "lgs.h"

namespace {
struct Batch {  // tensors that share (beta1, beta2, eps, step, device): one launch (chunks of LGS_ADAM_MAX_TENSORS)
    std::vector<float*> p, m, v;
    std::vector<const float*> g;
    std::vector<int64_t> n;
    std::vector<double> lr;
};
// libtorch 2.0/2.1 key the state map by the printed TensorImpl address, later versions by the pointer itself
template <typename Map>
auto state_key(const Map&, const torch::Tensor& t) {
    using K = typename Map::key_type;
    if constexpr (std::is_same_v<K, std::string>) {
        std::ostringstream ss;  // what c10::guts::to_string(TensorImpl*) printed
        ss << t.unsafeGetTensorImpl();
        return ss.str();
    } else {
        return static_cast<K>(t.unsafeGetTensorImpl());
    }
}
}  // namespace

torch::Tensor LgsFusedAdam::step(LossClosure closure) {
    torch::NoGradGuard no_grad;
    torch::Tensor loss;
    if (closure != nullptr) {
        at::AutoGradMode enable_grad(true);
        loss = closure();
    }
    std::map<std::tuple<double, double, double, int64_t, int>, Batch> batches;
    for (auto& group : param_groups()) {
        auto& opt = static_cast<torch::optim::AdamOptions&>(group.options());
        TORCH_CHECK(opt.weight_decay() == 0 && !opt.amsgrad(), "LgsFusedAdam: weight_decay / amsgrad are not supported");
        for (auto& p : group.params()) {
            if (!p.grad().defined()) continue;
            TORCH_CHECK(p.is_cuda(), "leg_slam_b200 has no CPU path: tensors must live on a CUDA device");
            TORCH_CHECK(p.scalar_type() == torch::kFloat32 && p.is_contiguous(), "LgsFusedAdam expects contiguous float32 parameters");
            auto& states = state();
            auto key = state_key(states, p);
            auto it = states.find(key);
            if (it == states.end()) {
                auto fresh = std::make_unique<torch::optim::AdamParamState>();
                fresh->step(0);
                fresh->exp_avg(torch::zeros_like(p, torch::MemoryFormat::Preserve));
                fresh->exp_avg_sq(torch::zeros_like(p, torch::MemoryFormat::Preserve));
                it = states.emplace(key, std::move(fresh)).first;
            }
            auto& st = static_cast<torch::optim::AdamParamState&>(*it->second);
            st.step(st.step() + 1);
            const torch::Tensor g = p.grad().contiguous();
            TORCH_CHECK(g.scalar_type() == torch::kFloat32 && g.numel() == p.numel(), "LgsFusedAdam: gradient does not match its parameter");
            auto& b = batches[{std::get<0>(opt.betas()), std::get<1>(opt.betas()), opt.eps(), st.step(), p.get_device()}];
            b.p.push_back(p.data_ptr<float>());
            b.g.push_back(g.data_ptr<float>());
            b.m.push_back(st.exp_avg().data_ptr<float>());
            b.v.push_back(st.exp_avg_sq().data_ptr<float>());
            b.n.push_back(p.numel());
            b.lr.push_back(opt.lr());
            if (!g.is_same(p.grad())) p.mutable_grad() = g;  // keep the contiguous copy alive until the launch is queued
        }
    }
    for (auto& [key, b] : batches) {
        const auto& [beta1, beta2, eps, step_no, device] = key;
        const c10::cuda::CUDAGuard guard(device);
        const int total = (int)b.p.size();
        for (int i = 0; i < total; i += 16) {
            const int k = std::min(16, total - i);
            const int st = lgs_adam_multi(k, b.p.data() + i, b.g.data() + i, b.m.data() + i, b.v.data() + i, b.n.data() + i,
                                          b.lr.data() + i, beta1, beta2, eps, (int)step_no,
                                          (void*)at::cuda::getCurrentCUDAStream(device).stream());
            TORCH_CHECK(st == LGS_OK, "lgs_adam_multi: ", lgs_status_string(st), " (cudaError ", lgs_last_cuda_error(), ")");
        }
    }
    return loss;
}
