// adam_math.cuh -- the per-element Adam update shared by adam.cu (single GPU / after NCCL all-reduce) and
// dp_adam.cu (fused peer-memory reduce + update + broadcast).  Operation order and contractions follow
// libtorch 2.0.1's torch::optim::Adam (see adam.cu's header) so both agree with the reference to the last bits.
#pragma once

namespace lgs {

__device__ __forceinline__ void adam_elem(float& p, float g, float& m, float& v, float b1, float omb1, float b2,
                                          float omb2, float inv_bc2_sqrt, float eps, float neg_step) {
    m = __fmaf_rn(omb1, g, __fmul_rn(m, b1));
    v = __fmaf_rn(__fmul_rn(omb2, g), g, __fmul_rn(v, b2));
    // Culled Gaussians have g = 0 and (until they are first rendered) m = v = 0: 60 % of the elements at cfgB.  The IEEE
    // sqrt / division sequences send zero operands through their out-of-line special-case paths (a CALL per warp that
    // holds one), which made this bandwidth-bound kernel instruction-bound (0.339 ms; 0.269 ms on dense random data).
    // sqrt(+-0) = +-0 and +-0 / dn = +-0 for dn > 0, so the zeros are passed through -- bit-identical results.
    float sq = v;
    if (v != 0.0f) sq = __fsqrt_rn(v);
    const float dn = __fadd_rn(__fmul_rn(sq, inv_bc2_sqrt), eps);
    float q = m;
    if (m != 0.0f) q = __fdiv_rn(m, dn);
    p = __fmaf_rn(neg_step, q, p);
}

}  // namespace lgs
