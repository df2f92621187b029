// TEST INFRASTRUCTURE -- force-included (-include) in front of the reference's src/gaussian_model.cpp and of
// oracle/ref_model_wrap.cpp when oracle/build_ref.py build_model() compiles them for a machine WITHOUT a GPU and against this
// image's libtorch 2.11 (the reference targets libtorch 2.0.1 + CUDA, setup.sh).  The reference sources stay untouched; three
// names they use are re-pointed, after every libtorch header has been read (so libtorch itself is not affected):
//
//  1. torch::kCUDA -> torch::kCPU.  The class takes its device from its parameters (src/gaussian_model.cpp:37-41) and we ask
//     for "cpu", but general_utils::build_rotation allocates with a literal torch::kCUDA (include/general_utils.h:42), so
//     densifyAndSplit would stop at its first statement here.  Same arithmetic, on the only device this container has.
//  2. c10::cuda::CUDACachingAllocator::emptyCache() -- the closing statement of increasePcd and densifyAndPrune
//     (src/gaussian_model.cpp:291,380,831) -- raises "Found no NVIDIA driver" on a driverless machine; it is routed to a
//     function that calls it only when a CUDA device exists.
//  3. libtorch 2.0.1 -> 2.11: Optimizer::state() was keyed by std::string (c10::guts::to_string of the TensorImpl pointer) and
//     is keyed by the pointer itself (void*) now, and c10::guts::to_string is gone.  The reference only uses the result as that
//     map's key (src/gaussian_model.cpp:580,590,608,619,680,693), so returning the pointer keeps every line meaning the same.
#pragma once
#include <torch/torch.h>
#include <c10/cuda/CUDACachingAllocator.h>
#include <c10/cuda/CUDAFunctions.h>
#ifdef TORCH_EXTENSION_NAME
#include <torch/extension.h>
#endif

namespace c10 {
namespace guts {
inline void* to_string(c10::TensorImpl* p) { return p; }
}  // namespace guts
namespace cuda {
namespace CUDACachingAllocator {
inline void lgs_empty_cache_if_cuda() {
    if (c10::cuda::device_count() > 0) emptyCache();
}
}  // namespace CUDACachingAllocator
}  // namespace cuda
}  // namespace c10

#define emptyCache lgs_empty_cache_if_cuda
#define kCUDA kCPU
