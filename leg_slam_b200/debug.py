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
    return dict(keys_unsorted=_view(binning_buffer, v.keys_unsorted, R, torch.int64),
                values_unsorted=_view(binning_buffer, v.values_unsorted, R, torch.int32),
                keys_sorted=_view(binning_buffer, v.keys_sorted, R, torch.int64),
                point_list=_view(binning_buffer, v.point_list, R, torch.int32))


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
                point_offsets=_view(geom_buffer, v.point_offsets, P, torch.int32),
                internal_radii=_view(geom_buffer, v.internal_radii, P, torch.int32),
                clamped=_view(geom_buffer, v.clamped, P, torch.uint8))


def binning_mode(mode):
    """1 = the reference's single radix sort (default), 0 = tile-local binning (include/lgs.h)."""
    check(_lib.lib().lgs_binning_mode(int(mode)), "lgs_binning_mode")


def debug_keys(on):
    """Keep / materialise the reference's exact 64-bit key arrays for binning_view (include/lgs.h)."""
    check(_lib.lib().lgs_debug_keys(int(bool(on))), "lgs_debug_keys")
