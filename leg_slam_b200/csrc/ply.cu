// ply.cu -- packing of the Gaussian set into / out of the interleaved per-vertex record of the reference's .ply
// checkpoints, on the GPU.  The reference copies 8 tensors to the host one by one
// (with two transposes) and lets tinyply interleave them on one CPU thread (src/gaussian_model.cpp:972-1075); loading
// goes the other way and drops the language features (:854-970).  Here ONE kernel builds (or consumes) the [P][C]
// float32 block exactly as it lies in the file, so the host side is a header plus one bulk copy.
//
// The column layout is data: `col_tensor[c]` / `col_elem[c]` say which tensor and which element of its row column c
// holds (tensor -1 = a zero column on packing / ignored on unpacking: the reference's normals), built on the host from
// the property names (leg_slam_b200/ply_io.py).  That covers the reference's own files, files with the optimizer
// extension (Adam moments as extra properties) and files whose properties come in another order.
#include "common.cuh"

namespace lgs {

constexpr int PLY_MAX_TENSORS = 24;
struct PlyTable {
    float* ptr[PLY_MAX_TENSORS];
    int row[PLY_MAX_TENSORS];
};

template <bool PACK>
__global__ void __launch_bounds__(256)
ply_pack_kernel(long long P, int C, const int* __restrict__ col_tensor, const int* __restrict__ col_elem, PlyTable tab,
                float* __restrict__ block) {
    const long long n = P * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long v = i / C;
        const int c = (int)(i - v * C);
        const int t = col_tensor[c];
        if (PACK) {
            block[i] = t < 0 ? 0.0f : tab.ptr[t][v * tab.row[t] + col_elem[c]];
        } else if (t >= 0) {
            tab.ptr[t][v * tab.row[t] + col_elem[c]] = block[i];
        }
    }
}

}  // namespace lgs

using namespace lgs;

static int ply_run(bool pack, long long P, int C, const int* col_tensor, const int* col_elem, int n_tensors, float* const* tensors,
                   const int* row_floats, float* block, void* stream) {
    if (P < 0 || C <= 0 || n_tensors <= 0 || n_tensors > PLY_MAX_TENSORS) return LGS_ERR_INVALID_ARG;
    if (P == 0) return LGS_OK;
    if (!col_tensor || !col_elem || !tensors || !row_floats || !block) return LGS_ERR_INVALID_ARG;
    PlyTable tab;
    for (int t = 0; t < n_tensors; ++t) {
        if (!tensors[t] || row_floats[t] <= 0) return LGS_ERR_INVALID_ARG;
        tab.ptr[t] = tensors[t];
        tab.row[t] = row_floats[t];
    }
    const long long n = P * C;
    const int grid = (int)((n + 255) / 256 < 148LL * 16 ? (n + 255) / 256 : 148LL * 16);
    if (pack)
        ply_pack_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(P, C, col_tensor, col_elem, tab, block);
    else
        ply_pack_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(P, C, col_tensor, col_elem, tab, block);
    LGS_LAUNCH_CHECK();
    return LGS_OK;
}

extern "C" int lgs_ply_pack(long long P, int C, const int* col_tensor, const int* col_elem, int n_tensors,
                            const float* const* tensors, const int* row_floats, float* block, void* stream) {
    return ply_run(true, P, C, col_tensor, col_elem, n_tensors, const_cast<float* const*>(tensors), row_floats, block, stream);
}
extern "C" int lgs_ply_unpack(long long P, int C, const int* col_tensor, const int* col_elem, int n_tensors, float* const* tensors,
                              const int* row_floats, const float* block, void* stream) {
    return ply_run(false, P, C, col_tensor, col_elem, n_tensors, tensors, row_floats, const_cast<float*>(block), stream);
}
