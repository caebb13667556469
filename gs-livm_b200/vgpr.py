"""Producer side of the map on the GPU: one LiDAR sweep -> voxel buckets -> per-voxel Gaussian-process
regression (VGPR) -> variance-centred Gaussian initialisation -> the parameters GaussianModel::Create_from_pcd would hold.

Reference: GpMap::dividePointsIntoCellInitMap (src/gp3d/map.cpp:7-37,51-111: voxel hash on floor(p / grid), at most
2 x min_points points per voxel, voxels with >= min_points points are processed, the LAST min_points of a bucket are the
training set, gpprocess.cu:251-262), Cell / PointMatrix (src/gp3d/cell.cpp:5-30, include/gs/gp3d/pointMatrix.h:105-125:
PCA, the GP regresses along the axis closest to the smallest-eigenvalue direction; surface iff lambda_max / lambda_mid >
eigen_1), gpProcess::forward_gp3d + fastInitial3DGS (src/gp3d/gpprocess.cu:504-915, :420-458) and
GaussianModel::Create_from_pcd (src/gs/gaussian.cu:325-386). The reference allocates and frees per frame, creates a
cuBLAS handle per call, round-trips to the host and builds tensors point by point; here the bucketing and the 3x3 PCA
are a handful of torch sort / scatter ops on the device (plumbing) and the regression + initialisation is ONE launch of
libgsr_b200's vgpr kernel (csrc/vgpr.cu, one CTA per voxel). No CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Dict, Optional

import torch

from . import _lib

SH_C0 = 0.28209479177387814
_OFF = 1 << 20  # voxel coordinates are packed as 21-bit fields of one int64 key


def bucket_points(points: torch.Tensor, grid: float = 0.2, min_points: int = 10):
    """points [N,3] (CUDA, float32 or float64, arrival order) -> (voxel [V,3] int32, train [V,min_points,3] float32,
    member index [K] into `points` of every kept point, voxel id [K] of every kept point, kept count [V]). Voxels come
    out in ascending (x, y, z) order of their integer coordinates."""
    if not points.is_cuda:
        raise RuntimeError("gs_livm_b200 has no CPU path: points must live on a CUDA device")
    key3 = torch.floor(points.double() / grid).to(torch.int64)
    if int(key3.abs().max()) >= _OFF:
        raise ValueError("voxel coordinates beyond +-2^20")
    packed = ((key3[:, 0] + _OFF) << 42) | ((key3[:, 1] + _OFF) << 21) | (key3[:, 2] + _OFF)
    skey, order = torch.sort(packed, stable=True)  # arrival order kept inside a voxel
    ukey, count = torch.unique_consecutive(skey, return_counts=True)
    start = torch.cumsum(count, 0) - count
    kept = torch.clamp(count, max=2 * min_points)  # "do not add point" beyond 2 x min_points (map.cpp:30-33)
    ok = kept >= min_points
    ukey, start, kept = ukey[ok], start[ok], kept[ok]
    V = int(ukey.numel())
    voxel = torch.stack([(ukey >> 42) - _OFF, ((ukey >> 21) & (2 * _OFF - 1)) - _OFF, (ukey & (2 * _OFF - 1)) - _OFF], 1)
    ar = torch.arange(min_points, device=points.device)
    train = points[order[(start + kept - min_points)[:, None] + ar[None, :]]].float().contiguous()
    # every kept point with its voxel id (for the PCA)
    vid = torch.repeat_interleave(torch.arange(V, device=points.device), kept)
    first = torch.cumsum(kept, 0) - kept
    member = order[start[vid] + (torch.arange(int(kept.sum()), device=points.device) - first[vid])]
    return voxel.to(torch.int32).contiguous(), train, member, vid, kept


def surface_direction(points: torch.Tensor, member: torch.Tensor, vid: torch.Tensor, kept: torch.Tensor,
                      eigen_1: float = 1.0):
    """-> (direction [V] int32 in {0,1,2}, is_surface [V] bool): PCA of every voxel's kept points in float64."""
    V = int(kept.numel())
    p = points.double()[member]
    n = kept.double()[:, None]
    mean = torch.zeros((V, 3), dtype=torch.float64, device=points.device).index_add_(0, vid, p) / n
    d = p - mean[vid]
    outer = (d[:, :, None] * d[:, None, :]).reshape(-1, 9)
    cov = (torch.zeros((V, 9), dtype=torch.float64, device=points.device).index_add_(0, vid, outer) / n).reshape(V, 3, 3)
    lam, vec = torch.linalg.eigh(cov)  # ascending
    direction = vec[:, :, 0].abs().argmax(dim=1)  # the axis with the smallest angle to the smallest-eigenvalue direction
    is_surface = lam[:, 2] / lam[:, 1] > eigen_1
    return direction.to(torch.int32).contiguous(), is_surface


def gp_gaussians(voxel: torch.Tensor, train: torch.Tensor, direction: torch.Tensor, grid: float = 0.2,
                 num_gp_side: int = 12, neighbour: int = 3, kernel_size: float = 1.0, variance_sensor: float = 0.05,
                 scale_factor: float = 3.0, full_cover: bool = False, keep_lattice: bool = False,
                 ) -> Dict[str, torch.Tensor]:
    """libgsr_b200's vgpr kernel: per voxel the GP posterior on the lattice and the g x g Gaussians of its blocks.
    -> means [V g g,3], covs [V g g,3,3], scales [V g g,3], var_mean [V] (+ test_points, kvar with keep_lattice)."""
    lib = _lib.load()
    V, T = int(train.shape[0]), int(train.shape[1])
    dev = train.device
    g = num_gp_side // neighbour
    f32 = dict(dtype=torch.float32, device=dev)
    out = dict(means=torch.empty((V * g * g, 3), **f32), covs=torch.empty((V * g * g, 3, 3), **f32),
               scales=torch.empty((V * g * g, 3), **f32), var_mean=torch.empty((V,), **f32))
    if keep_lattice:
        out["test_points"] = torch.empty((V, num_gp_side * num_gp_side, 3), **f32)
        out["kvar"] = torch.empty((V, num_gp_side * num_gp_side), **f32)
    prm = _lib.VgprParams(float(grid), int(num_gp_side), int(neighbour), float(kernel_size), float(variance_sensor),
                          float(scale_factor), int(bool(full_cover)))
    st = lib.gsr_vgpr_gaussians(V, T, train.contiguous().data_ptr(), voxel.contiguous().data_ptr(),
                                direction.contiguous().data_ptr(), C.byref(prm), out["means"].data_ptr(),
                                out["covs"].data_ptr(), out["scales"].data_ptr(), out["var_mean"].data_ptr(),
                                out["test_points"].data_ptr() if keep_lattice else None,
                                out["kvar"].data_ptr() if keep_lattice else None,
                                torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(st, "gsr_vgpr_gaussians")
    return out


def project_colours(xyz: torch.Tensor, viewmatrix: torch.Tensor, fx: float, fy: float, width: int, height: int,
                    image_u8: torch.Tensor):
    """gpProcess::getColors (src/gp3d/gpprocess.cu:917-969, without lens distortion): project into the keyframe image
    [H,W,3] uint8, drop what falls outside. -> (rgb uint8 [n,3], valid [n])."""
    pc = torch.cat([xyz.double(), torch.ones((xyz.shape[0], 1), dtype=torch.float64, device=xyz.device)], 1) @ viewmatrix.double()
    z = pc[:, 2]
    u = (fx * pc[:, 0] / z + 0.5 * width).to(torch.int64)
    v = (fy * pc[:, 1] / z + 0.5 * height).to(torch.int64)
    valid = (z > 0.2) & (u >= 0) & (u < width) & (v >= 0) & (v < height)
    rgb = image_u8[v.clamp(0, height - 1), u.clamp(0, width - 1)]
    return rgb, valid


def vgpr_frame(points: torch.Tensor, colour_of: Callable, grid: float = 0.2, min_points: int = 10, num_gp_side: int = 12,
               neighbour: int = 3, kernel_size: float = 1.0, variance_sensor: float = 0.05, eigen_1: float = 1.0,
               max_var_mean: float = 0.30, scale_factor: float = 3.0, sh_degree: int = 3,
               raw: bool = False) -> Dict[str, torch.Tensor]:
    """One frame of the reference's producer side (lioOptimization.cpp:1201 gsPointCloudUpdate ->
    GpMap::dividePointsIntoCellInitMap -> gpProcess::forward_gp3d -> Create_from_pcd / addNewPointcloud) on the device.
    colour_of(xyz [n,3] CUDA) -> (rgb uint8 [n,3], valid bool [n]). Returns the ACTIVATED parameters the rasterizer
    consumes (means3D, scales, rotations, opacities, shs); raw=True: GaussianModel's raw parameters instead (log scales,
    logit opacity), for the GSR_ACT_* entry point. Also "needs_more_points" (voxels over the variance gate) and "voxels"."""
    voxel, train, member, vid, kept = bucket_points(points, grid, min_points)
    direction, is_surface = surface_direction(points, member, vid, kept, eigen_1)
    voxel, train, direction = voxel[is_surface].contiguous(), train[is_surface].contiguous(), direction[is_surface].contiguous()
    gp = gp_gaussians(voxel, train, direction, grid, num_gp_side, neighbour, kernel_size, variance_sensor, scale_factor)
    rgb, valid = colour_of(gp["means"])
    means, scales, rgb = gp["means"][valid], gp["scales"][valid], rgb[valid]
    n, dev = int(means.shape[0]), means.device
    M = (sh_degree + 1) ** 2
    shs = torch.zeros((n, M, 3), dtype=torch.float32, device=dev)
    shs[:, 0, :] = (rgb.float() / 255.0 - 0.5) / SH_C0
    rots = torch.zeros((n, 4), dtype=torch.float32, device=dev)
    rots[:, 0] = 1.0
    if raw:
        scales_out, opac = torch.log(scales.clamp_min(1e-12)), torch.zeros((n, 1), dtype=torch.float32, device=dev)  # logit(0.5)
    else:
        scales_out, opac = scales, torch.full((n, 1), 0.5, dtype=torch.float32, device=dev)
    return dict(means3D=means.contiguous(), scales=scales_out.contiguous(), rotations=rots, opacities=opac, shs=shs,
                voxels=int(voxel.shape[0]), needs_more_points=int((gp["var_mean"] > max_var_mean).sum()))
