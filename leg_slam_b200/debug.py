"""Typed views of the three opaque work buffers (parity tests compare keys / sorted order /
tile ranges bit-exactly).  Layout knowledge stays in the library: offsets come from
lgs_view_* (include/lgs.h), nothing is re-derived here."""
import ctypes

import torch

from . import _lib
from ._lib import check


def _view(buf, addr, count, dtype):
    off = addr - buf.data_ptr()
    nbytes = count * torch.empty((), dtype=dtype).element_size()
    assert 0 <= off and off + nbytes <= buf.numel(), "view outside buffer"
    return buf[off:off + nbytes].view(dtype)


def binning_view(binning_buffer, R):
    v = _lib.BinningView()
    check(_lib.lib().lgs_view_binning(binning_buffer.data_ptr(), R, ctypes.byref(v)), "lgs_view_binning")
    return dict(point_list=_view(binning_buffer, v.point_list, R, torch.int32),
                keys_sorted32=_view(binning_buffer, v.keys_sorted32, R, torch.int32))


def reference_keys(geom_buffer, binning_buffer, image_buffer, P, R, W, H):
    """The reference's BinningState arrays (rasterizer_impl.h:50-63) for a finished forward: the pairs duplicateWithKeys would
    have emitted, in its order, and the sorted 64-bit keys that correspond to this library's point_list and ranges
    (lgs_debug_reference_keys).  Parity tests compare them bit for bit with the reference's / the oracle's / the fixtures'."""
    L = _lib.lib()
    buf = torch.empty(L.lgs_debug_keys_bytes(P, R), dtype=torch.uint8, device=geom_buffer.device)
    v = _lib.ReferenceKeysView()
    with torch.cuda.device(geom_buffer.device):
        s = torch.cuda.current_stream(geom_buffer.device).cuda_stream
        check(L.lgs_debug_reference_keys(P, R, W, H, _lib.ptr(geom_buffer), _lib.ptr(binning_buffer), image_buffer.data_ptr(),
                                         buf.data_ptr(), ctypes.byref(v), s), "lgs_debug_reference_keys")
    return dict(keys_unsorted=_view(buf, v.keys_unsorted, R, torch.int64), values_unsorted=_view(buf, v.values_unsorted, R, torch.int32),
                keys_sorted=_view(buf, v.keys_sorted, R, torch.int64), point_offsets=_view(buf, v.point_offsets, P, torch.int32),
                point_list=binning_view(binning_buffer, R)["point_list"])


def image_view(image_buffer, W, H):
    v = _lib.ImageView()
    check(_lib.lib().lgs_view_image(image_buffer.data_ptr(), W, H, ctypes.byref(v)), "lgs_view_image")
    tiles = ((W + 7) // 8) * ((H + 7) // 8)
    return dict(ranges=_view(image_buffer, v.ranges, tiles * 2, torch.int32).view(tiles, 2),
                final_T=_view(image_buffer, v.final_T, W * H, torch.float32),
                n_contrib=_view(image_buffer, v.n_contrib, W * H, torch.int32))


def geom_view(geom_buffer, P):
    v = _lib.GeomView()
    check(_lib.lib().lgs_view_geom(geom_buffer.data_ptr(), P, ctypes.byref(v)), "lgs_view_geom")
    return dict(records=_view(geom_buffer, v.records, P * 12, torch.float32).view(P, 12),
                cov3D=_view(geom_buffer, v.cov3D, P * 6, torch.float32).view(P, 6),
                tiles_touched=_view(geom_buffer, v.tiles_touched, P, torch.int32),
                internal_radii=_view(geom_buffer, v.internal_radii, P, torch.int32),
                clamped=_view(geom_buffer, v.clamped, P, torch.uint8))
