"""Python mirror of the reference's LibTorch operator surface, driving libgsr_b200.so through its C ABI.

Same names, argument meaning, return tuples and error behaviour as the reference, so the parity tests
read like the reference's own call sites:

  GaussianRasterizationSettings   <- include/gs/gs/rasterizer.cuh:8-20
  rasterize_gaussians_cuda        <- RasterizeGaussiansCUDA          (src/gs/rasterize_points.cu:46-130)
  rasterize_gaussians_backward_cuda <- RasterizeGaussiansBackwardCUDA (src/gs/rasterize_points.cu:132-224)
  mark_visible                    <- markVisible                     (src/gs/rasterize_points.cu:226-241)
  _RasterizeGaussians             <- torch::autograd::Function       (src/gs/rasterizer.cu:6-149)
  GaussianRasterizer              <- src/gs/rasterizer.cu:152-288

torch is plumbing only (device memory, streams, autograd bookkeeping); all arithmetic happens in the
hand-written CUDA kernels. No CPU path exists: tensors must be CUDA tensors.
"""
from __future__ import annotations

import ctypes as C
from typing import NamedTuple, Optional, Tuple

import torch

from . import _lib

NUM_CHANNELS = 3


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    camera_center: torch.Tensor
    prefiltered: bool = False


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    """Device pointer of a contiguous tensor; None for undefined / empty tensors (the reference passes
    .data<float>() of 0-element tensors, i.e. nullptr, to select the SH / scale-rotation paths)."""
    if t is None or t.numel() == 0:
        return None
    return t.data_ptr()


def _f32c(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    if t is None:
        return None
    if t.dtype != torch.float32:
        raise TypeError("rasterizer tensors must be float32")
    return t.contiguous()


class _ScratchAllocator:
    """resizeFunctional (rasterize_points.cu:36-44): a byte tensor the C ABI sizes through a callback.

    The ctypes callback closes over a one-element list, not over `self`: a bound method would put the buffer
    into a reference cycle (object -> callback -> method -> object) and hand its lifetime to Python's cyclic
    collector, i.e. hundreds of MB of scratch per render would pile up until the next collection and the
    caching allocator would answer with cudaMalloc instead of reusing the previous render's blocks."""

    def __init__(self, device: torch.device):
        cell = [torch.empty(0, dtype=torch.uint8, device=device)]

        def _alloc(_ctx, nbytes, _cell=cell, _device=device):
            _cell[0] = torch.empty(int(nbytes), dtype=torch.uint8, device=_device)
            return _cell[0].data_ptr()

        self._cell = cell
        self.fn = _lib.ALLOC_FN(_alloc)

    @property
    def tensor(self) -> torch.Tensor:
        return self._cell[0]


def _stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def rasterize_gaussians_cuda(background, means3D, colors, opacity, scales, rotations, scale_modifier, cov3D_precomp,
                             viewmatrix, projmatrix, tan_fovx, tan_fovy, image_height, image_width, sh, degree,
                             campos, prefiltered, debug):
    """-> (num_rendered, out_color[3,H,W], out_depth[1,H,W], out_acc[1,H,W], radii[P], geomBuffer,
    binningBuffer, imgBuffer), exactly the reference's tuple."""
    if means3D.dim() != 2 or means3D.size(1) != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")
    if not means3D.is_cuda:
        raise RuntimeError("gs_livm_b200 has no CPU path: tensors must live on a CUDA device")
    lib = _lib.load()
    dev = means3D.device
    P, H, W = means3D.size(0), int(image_height), int(image_width)

    out_color = torch.zeros((NUM_CHANNELS, H, W), dtype=torch.float32, device=dev)
    out_depth = torch.zeros((1, H, W), dtype=torch.float32, device=dev)
    out_acc = torch.zeros((1, H, W), dtype=torch.float32, device=dev)
    radii = torch.zeros((P,), dtype=torch.int32, device=dev)
    geom, binning, img = _ScratchAllocator(dev), _ScratchAllocator(dev), _ScratchAllocator(dev)

    rendered = 0
    if P != 0:
        M = sh.size(1) if (sh is not None and sh.numel() != 0) else 0
        keep = [_f32c(t) for t in (background, means3D, sh, colors, opacity, scales, rotations, cov3D_precomp,
                                   viewmatrix, projmatrix, campos)]
        bg_c, m3_c, sh_c, col_c, op_c, sc_c, rot_c, cov_c, view_c, proj_c, cam_c = keep
        a = _lib.FwdArgs()
        a.P, a.D, a.M, a.width, a.height = P, int(degree), int(M), W, H
        a.background, a.means3D, a.shs, a.colors_precomp = _ptr(bg_c), _ptr(m3_c), _ptr(sh_c), _ptr(col_c)
        a.opacities, a.scales, a.scale_modifier = _ptr(op_c), _ptr(sc_c), float(scale_modifier)
        a.rotations, a.cov3D_precomp = _ptr(rot_c), _ptr(cov_c)
        a.viewmatrix, a.projmatrix, a.cam_pos = _ptr(view_c), _ptr(proj_c), _ptr(cam_c)
        a.tan_fovx, a.tan_fovy, a.prefiltered = float(tan_fovx), float(tan_fovy), int(bool(prefiltered))
        a.out_color, a.out_depth, a.out_acc = out_color.data_ptr(), out_depth.data_ptr(), out_acc.data_ptr()
        a.radii, a.debug = radii.data_ptr(), int(bool(debug))
        n = C.c_int(0)
        with torch.cuda.device(dev):
            st = lib.gsr_forward(C.byref(a), geom.fn, None, binning.fn, None, img.fn, None, _stream_ptr(dev),
                                 C.byref(n))
        _lib.check(st, "gsr_forward")
        rendered = n.value
    return rendered, out_color, out_depth, out_acc, radii, geom.tensor, binning.tensor, img.tensor


def rasterize_gaussians_backward_cuda(background, means3D, radii, colors, scales, rotations, scale_modifier,
                                      cov3D_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy, dL_dout_color,
                                      dL_dout_acc, sh, degree, campos, geomBuffer, R, binningBuffer, imageBuffer,
                                      debug):
    """-> (dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations)."""
    lib = _lib.load()
    dev = means3D.device
    P, H, W = means3D.size(0), dL_dout_color.size(1), dL_dout_color.size(2)
    M = sh.size(1) if (sh is not None and sh.numel() != 0) else 0
    opts = dict(dtype=torch.float32, device=dev)
    # The reference zero-fills all nine tensors here (rasterize_points.cu:173-181). The library is told the
    # outputs are uninitialised instead (GSR_BWD_OUTPUTS_UNINITIALIZED) and writes every element of what it
    # computes; only what it does not touch on this argument combination is zeroed here.
    have_cov = cov3D_precomp is not None and cov3D_precomp.numel() != 0
    fresh = torch.empty if P != 0 else torch.zeros
    dL_dmeans3D = fresh((P, 3), **opts)
    dL_dmeans2D = fresh((P, 3), **opts)
    dL_dcolors = fresh((P, NUM_CHANNELS), **opts)
    dL_dconic = fresh((P, 2, 2), **opts)
    dL_dopacity = fresh((P, 1), **opts)
    dL_dcov3D = fresh((P, 6), **opts)
    dL_dsh = fresh((P, M, 3), **opts)
    dL_dscales = (torch.zeros if have_cov else fresh)((P, 3), **opts)
    dL_drotations = (torch.zeros if have_cov else fresh)((P, 4), **opts)

    if P != 0:
        keep = [_f32c(t) for t in (background, means3D, sh, colors, scales, rotations, cov3D_precomp, viewmatrix,
                                   projmatrix, campos, dL_dout_color, dL_dout_acc)]
        bg_c, m3_c, sh_c, col_c, sc_c, rot_c, cov_c, view_c, proj_c, cam_c, dpix_c, dacc_c = keep
        radii_c = radii.contiguous()
        a = _lib.BwdArgs()
        a.P, a.D, a.M, a.R, a.width, a.height = P, int(degree), int(M), int(R), W, H
        a.background, a.means3D, a.shs, a.colors_precomp = _ptr(bg_c), _ptr(m3_c), _ptr(sh_c), _ptr(col_c)
        a.scales, a.scale_modifier, a.rotations, a.cov3D_precomp = _ptr(sc_c), float(scale_modifier), _ptr(rot_c), _ptr(cov_c)
        a.viewmatrix, a.projmatrix, a.cam_pos = _ptr(view_c), _ptr(proj_c), _ptr(cam_c)
        a.tan_fovx, a.tan_fovy = float(tan_fovx), float(tan_fovy)
        a.radii = _ptr(radii_c)
        a.geom_buffer, a.binning_buffer, a.image_buffer = geomBuffer.data_ptr(), binningBuffer.data_ptr(), imageBuffer.data_ptr()
        a.dL_dpix, a.dL_dacc = _ptr(dpix_c), _ptr(dacc_c)
        a.dL_dmean2D, a.dL_dconic, a.dL_dopacity = dL_dmeans2D.data_ptr(), dL_dconic.data_ptr(), dL_dopacity.data_ptr()
        a.dL_dcolor, a.dL_dmean3D, a.dL_dcov3D = dL_dcolors.data_ptr(), dL_dmeans3D.data_ptr(), dL_dcov3D.data_ptr()
        a.dL_dsh, a.dL_dscale, a.dL_drot = _ptr(dL_dsh), dL_dscales.data_ptr(), dL_drotations.data_ptr()
        a.debug, a.flags = int(bool(debug)), _lib.BWD_OUTPUTS_UNINITIALIZED
        with torch.cuda.device(dev):
            st = lib.gsr_backward(C.byref(a), _stream_ptr(dev))
        _lib.check(st, "gsr_backward")
    return dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations


def mark_visible(means3D: torch.Tensor, viewmatrix: torch.Tensor, projmatrix: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    P = means3D.size(0)
    present = torch.zeros((P,), dtype=torch.bool, device=means3D.device)
    if P != 0:
        m, v, p = _f32c(means3D), _f32c(viewmatrix), _f32c(projmatrix)
        with torch.cuda.device(means3D.device):
            st = lib.gsr_mark_visible(P, m.data_ptr(), v.data_ptr(), p.data_ptr(), present.data_ptr(),
                                      _stream_ptr(means3D.device))
        _lib.check(st, "gsr_mark_visible")
    return present


class _RasterizeGaussians(torch.autograd.Function):
    """Reference: src/gs/rasterizer.cu:6-149. Returns (color, radii, depth, acc); the backward consumes only
    the colour and silhouette gradients — the depth gradient is dropped exactly like the reference does
    (rasterizer.cu:77-80,117-118)."""

    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                raster_settings: GaussianRasterizationSettings):
        rs = raster_settings
        (num_rendered, color, out_depth, out_acc, radii, geomBuffer, binningBuffer, imgBuffer) = \
            rasterize_gaussians_cuda(rs.bg, means3D, colors_precomp, opacities, scales, rotations, rs.scale_modifier,
                                     cov3Ds_precomp, rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy,
                                     rs.image_height, rs.image_width, sh, rs.sh_degree, rs.camera_center,
                                     rs.prefiltered, False)
        ctx.raster_settings = rs
        ctx.num_rendered = num_rendered
        ctx.save_for_backward(colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geomBuffer,
                              binningBuffer, imgBuffer)
        ctx.mark_non_differentiable(radii)
        return color, radii, out_depth, out_acc

    @staticmethod
    def backward(ctx, grad_out_color, _grad_radii, _grad_out_depth, grad_acc):
        rs = ctx.raster_settings
        (colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geomBuffer, binningBuffer,
         imgBuffer) = ctx.saved_tensors
        if grad_out_color is None:
            grad_out_color = torch.zeros((NUM_CHANNELS, rs.image_height, rs.image_width), dtype=torch.float32,
                                         device=means3D.device)
        if grad_acc is None:
            grad_acc = torch.zeros((1, rs.image_height, rs.image_width), dtype=torch.float32, device=means3D.device)
        (grad_means2D, grad_colors_precomp, grad_opacities, grad_means3D, grad_cov3Ds_precomp, grad_sh, grad_scales,
         grad_rotations) = rasterize_gaussians_backward_cuda(
            rs.bg, means3D, radii, colors_precomp, scales, rotations, rs.scale_modifier, cov3Ds_precomp,
            rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, grad_out_color, grad_acc, sh, rs.sh_degree,
            rs.camera_center, geomBuffer, ctx.num_rendered, binningBuffer, imgBuffer, False)
        return (grad_means3D, grad_means2D, grad_sh, grad_colors_precomp, grad_opacities, grad_scales, grad_rotations,
                grad_cov3Ds_precomp, None)


class GaussianRasterizer(torch.nn.Module):
    """Reference: GaussianRasterizer (include/gs/gs/rasterizer.cuh:52-80, src/gs/rasterizer.cu:152-288)."""

    def __init__(self, raster_settings: GaussianRasterizationSettings):
        super().__init__()
        self.raster_settings = raster_settings

    def mark_visible(self, positions: torch.Tensor) -> torch.Tensor:
        with torch.no_grad():
            return mark_visible(positions, self.raster_settings.viewmatrix, self.raster_settings.projmatrix)

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
        if (shs is None) == (colors_precomp is None):
            raise ValueError("Please provide exactly one of either SHs or precomputed colors!")
        if ((scales is not None or rotations is not None) and cov3D_precomp is not None) or \
                (scales is None and rotations is None and cov3D_precomp is None):
            raise ValueError("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")
        dev = means3D.device
        empty = torch.empty(0, dtype=torch.float32, device=dev)
        shs = empty if shs is None else shs
        colors_precomp = empty if colors_precomp is None else colors_precomp
        scales = empty if scales is None else scales
        rotations = empty if rotations is None else rotations
        cov3D_precomp = empty if cov3D_precomp is None else cov3D_precomp
        rs = self.raster_settings
        rs = rs._replace(bg=rs.bg.to(dev), viewmatrix=rs.viewmatrix.to(dev), projmatrix=rs.projmatrix.to(dev),
                         camera_center=rs.camera_center.to(dev))
        return _RasterizeGaussians.apply(means3D, means2D, shs, colors_precomp, opacities, scales, rotations,
                                         cov3D_precomp, rs)
