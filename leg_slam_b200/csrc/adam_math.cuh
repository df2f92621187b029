// adam_math.cuh -- the per-element Adam update shared by adam.cu (single GPU / after NCCL all-reduce) and
// dp_adam.cu (fused peer-memory reduce + update + broadcast).  Operation order and contractions follow
// libtorch 2.0.1's torch::optim::Adam (see adam.cu's header) so both agree with the reference to the last bits.
#pragma once

namespace lgs {

__device__ __forceinline__ void adam_elem(float& p, float g, float& m, float& v, float b1, float omb1, float b2,
                                          float omb2, float inv_bc2_sqrt, float eps, float neg_step) {
    m = __fmaf_rn(omb1, g, __fmul_rn(m, b1));
    v = __fmaf_rn(__fmul_rn(omb2, g), g, __fmul_rn(v, b2));
    const float dn = __fadd_rn(__fmul_rn(__fsqrt_rn(v), inv_bc2_sqrt), eps);
    p = __fmaf_rn(neg_step, __fdiv_rn(m, dn), p);
}

}  // namespace lgs
