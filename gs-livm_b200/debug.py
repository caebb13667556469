"""Debug / parity accessors: typed torch views of the three opaque scratch buffers, through the C ABI's
gsr_view_* (the counterpart of the reference's GeometryState/BinningState/ImageState::fromChunk,
src/cuda_rasterizer/rasterizer_impl.cu:137-177). Zero-copy: the returned tensors alias the buffers."""
from __future__ import annotations

import ctypes as C
from typing import Dict

import torch

from . import _lib


def _typed(buf: torch.Tensor, view: "_lib.ArrayView", dtype: torch.dtype, width: int) -> torch.Tensor:
    """[count, width] strided view of `buf` (uint8) starting at view.ptr with view.stride bytes per row."""
    count, stride = int(view.count), int(view.stride)
    esz = torch.empty((), dtype=dtype).element_size()
    if count == 0:
        return torch.empty((0, width), dtype=dtype, device=buf.device)
    off = int(view.ptr) - buf.data_ptr()
    assert off >= 0 and off % esz == 0 and stride % esz == 0
    nbytes = (count - 1) * stride + width * esz
    flat = buf[off:off + nbytes].view(dtype)
    return flat.as_strided((count, width), (stride // esz, 1))


def geom_state(geomBuffer: torch.Tensor, P: int) -> Dict[str, torch.Tensor]:
    v = _lib.GeomView()
    _lib.check(_lib.load().gsr_view_geom(geomBuffer.data_ptr(), int(P), C.byref(v)), "gsr_view_geom")
    f32, i32, u8 = torch.float32, torch.int32, torch.uint8
    return dict(
        means2D=_typed(geomBuffer, v.means2D, f32, 2),
        extent=_typed(geomBuffer, v.extent, f32, 2),
        conic_opacity=_typed(geomBuffer, v.conic_opacity, f32, 4),
        rgb=_typed(geomBuffer, v.rgb, f32, 3),
        depths=_typed(geomBuffer, v.depths, f32, 1)[:, 0],
        clamped=_typed(geomBuffer, v.clamped, u8, 1)[:, 0],
        internal_radii=_typed(geomBuffer, v.internal_radii, i32, 1)[:, 0],
        tiles_touched=_typed(geomBuffer, v.tiles_touched, i32, 1)[:, 0],
        point_offsets=_typed(geomBuffer, v.point_offsets, i32, 1)[:, 0],
    )


def binning_state(binningBuffer: torch.Tensor, R: int) -> Dict[str, torch.Tensor]:
    v = _lib.BinningView()
    _lib.check(_lib.load().gsr_view_binning(binningBuffer.data_ptr(), int(R), C.byref(v)), "gsr_view_binning")
    i32, i64 = torch.int32, torch.int64
    return dict(
        point_list_keys_unsorted=_typed(binningBuffer, v.point_list_keys_unsorted, i64, 1)[:, 0],
        point_list_unsorted=_typed(binningBuffer, v.point_list_unsorted, i32, 1)[:, 0],
        point_list_keys=_typed(binningBuffer, v.point_list_keys, i64, 1)[:, 0],
        point_list=_typed(binningBuffer, v.point_list, i32, 1)[:, 0],
        instance_records=_typed(binningBuffer, v.instance_records, torch.float32, 12),
        block_candidates=_typed(binningBuffer, v.block_candidates, i32, 2),
    )


def image_state(imgBuffer: torch.Tensor, width: int, height: int) -> Dict[str, torch.Tensor]:
    v = _lib.ImageView()
    _lib.check(_lib.load().gsr_view_image(imgBuffer.data_ptr(), int(width), int(height), C.byref(v)), "gsr_view_image")
    return dict(
        accum_alpha=_typed(imgBuffer, v.accum_alpha, torch.float32, 1)[:, 0],
        n_contrib=_typed(imgBuffer, v.n_contrib, torch.int32, 1)[:, 0],
        ranges=_typed(imgBuffer, v.ranges, torch.int32, 2),
        tile_last=_typed(imgBuffer, v.tile_last, torch.int32, 1)[:, 0],
        tile_order=_typed(imgBuffer, v.tile_order, torch.int32, 1)[:, 0],
        block_count=_typed(imgBuffer, v.block_count, torch.int32, 1)[:, 0],
        last_group=_typed(imgBuffer, v.last_group, torch.int32, 1)[:, 0],
        summary=_typed(imgBuffer, v.summary, torch.int32, 1)[:, 0],
    )


def frame_counters(imgBuffer: torch.Tensor, width: int, height: int) -> Dict[str, int]:
    """The per-frame counters the kernels leave in the image state: instance count, list statistics and the blend's
    executed work (pairs inside the candidates' boxes = alphas evaluated; pairs blended = what the backward visits)."""
    w = [int(x) & 0xffffffff for x in image_state(imgBuffer, width, height)["summary"].cpu().tolist()]
    return dict(num_rendered=w[0], longest_list=w[1], long_lists=w[2], overflow=w[3], depth_buckets=w[4],
                pairs_in_boxes=w[6] | (w[7] << 32), pairs_blended=w[8] | (w[9] << 32))
