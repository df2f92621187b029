// TEST INFRASTRUCTURE -- C entry point around the UNMODIFIED reference simple-knn
// (/root/reference/third_party/simple-knn/simple_knn.cu, compiled from where it lies by oracle/build_ref.py into
// oracle/_ref/ref_simple_knn.so).  Device pointers in, device pointer out; synchronises.
#include <cuda_runtime.h>
#include "simple_knn.h"

extern "C" int ref_simple_knn(int P, float* points_dev, float* mean_dists_dev) {
    SimpleKNN::knn(P, reinterpret_cast<float3*>(points_dev), mean_dists_dev);
    return (int)cudaDeviceSynchronize();
}
