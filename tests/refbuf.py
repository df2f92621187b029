"""Typed views of the REFERENCE rasterizer's opaque buffers (test infrastructure).

Layouts restated from the reference's fromChunk routines (cuda_rasterizer/rasterizer_impl.cu:
155-194, obtain() rasterizer_impl.h:21-28: each array starts at the next 128-byte aligned
ADDRESS).  Only arrays that precede the CUB scratch space are exposed, so no CUB size query
is needed."""
import torch


def _carve(buf, specs):
    base = buf.data_ptr()
    cur = base
    out = {}
    for name, count, dtype in specs:
        cur = (cur + 127) & ~127
        nbytes = count * torch.empty((), dtype=dtype).element_size()
        off = cur - base
        out[name] = buf[off:off + nbytes].view(dtype)
        cur += nbytes
    return out


def ref_geom_view(buf, P):
    v = _carve(buf, [("depths", P, torch.float32), ("clamped", 3 * P, torch.uint8),
                     ("internal_radii", P, torch.int32), ("means2D", 2 * P, torch.float32),
                     ("cov3D", 6 * P, torch.float32), ("conic_opacity", 4 * P, torch.float32),
                     ("rgb", 3 * P, torch.float32), ("tiles_touched", P, torch.int32)])
    v["means2D"] = v["means2D"].view(P, 2)
    v["cov3D"] = v["cov3D"].view(P, 6)
    v["conic_opacity"] = v["conic_opacity"].view(P, 4)
    v["rgb"] = v["rgb"].view(P, 3)
    v["clamped"] = v["clamped"].view(P, 3)
    return v


def ref_binning_view(buf, R):
    return _carve(buf, [("point_list", R, torch.int32), ("point_list_unsorted", R, torch.int32),
                        ("keys_sorted", R, torch.int64), ("keys_unsorted", R, torch.int64)])


def ref_image_view(buf, W, H):
    N = W * H
    v = _carve(buf, [("final_T", N, torch.float32), ("n_contrib", N, torch.int32), ("ranges", 2 * N, torch.int32)])
    tiles = ((W + 7) // 8) * ((H + 7) // 8)
    v["ranges"] = v["ranges"][:2 * tiles].view(tiles, 2)
    return v
