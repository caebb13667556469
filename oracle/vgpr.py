"""TEST INFRASTRUCTURE / CPU BASELINE — numpy restatement of the reference's CPU-side mapping path around the
rasterizer: voxel-GP regression (VGPR), variance-centred Gaussian initialisation and the image loss.

Only tests/ and bench.py's cpu_baseline leg may import this module (see oracle/gsr_oracle.h). It exists because
BASELINE.json's north star reports the GPU rasterizer "alongside the reference's CPU-side mapping path" and because
configs[0] (one NTU-VIRAL-shaped frame: VGPR + Gaussian init of ~200 K Gaussians on the host, then one 752x480
render checked against the scalar oracle) needs Gaussians produced the way the reference produces them. The
reference's own implementation needs ROS/Eigen/OpenCV/cuBLAS and cannot be built here, so this is a
restatement ("port"), batched over voxels with numpy. Paths are relative to /root/reference.

  bucket_points        src/gp3d/map.cpp:7-37 (voxel hash on floor(p / grid), at most 2 x min_points per voxel),
                       :70-84 (voxels with >= min_points are processed), gpprocess.cu:251-262 (the LAST
                       min_points points of a bucket are the training set)
  surface_direction    src/gp3d/cell.cpp:5-30 + include/gs/gp3d/pointMatrix.h:105-125 (PCA: the GP regresses along
                       the axis closest to the smallest-eigenvalue direction; surface iff lambda_max / lambda_mid >
                       eigen_1)
  gp_regress           src/gp3d/gpprocess.cu:8-46 (test lattice, k = exp(-kernel_size * |d|), + sigma^2 on the
                       diagonal), :48-60 (mean-centred targets), :602-668 (Ky^-1, Ky^-1 K*, f*, K*^T Ky^-1 K*),
                       :461-501 (test points + their variance entry), :63-122 (variance gate max_var_mean)
  variance_centred_init  src/gp3d/gpprocess.cu:420-458 (fastInitial3DGS: inverse-variance weighted mean / covariance
                       of every neighbour x neighbour block of the lattice)
  gaussians_from_init  src/gs/gaussian.cu:10-11,325-386 (scale = sqrt(diag(cov) * scale_factor) after activation,
                       identity rotation, opacity 0.5, SH DC = (rgb / 255 - 0.5) / C0, higher SH = 0)
  image_loss           include/gs/gs/loss_utils.cuh:11-70 (L1, SSIM with its own 11-tap window), combined as
                       (1 - l) L1 + l (1 - SSIM), src/liw/lioOptimization.cpp:1729-1736
Parameters default to config/basic_common.yaml:55,69-83.
"""
from __future__ import annotations

import time
from typing import Callable, Dict

import numpy as np

SH_C0 = 0.28209479177387814


def bucket_points(points: np.ndarray, grid: float = 0.2, min_points: int = 10):
    """-> (voxel index triples [V,3], training points [V,min_points,3], all kept points as a list of arrays)."""
    key = np.floor(points / grid).astype(np.int64)
    order = np.lexsort((np.arange(len(points)), key[:, 2], key[:, 1], key[:, 0]))  # stable: arrival order kept
    ks = key[order]
    new = np.ones(len(ks), bool)
    new[1:] = (ks[1:] != ks[:-1]).any(axis=1)
    start = np.flatnonzero(new)
    count = np.diff(np.append(start, len(ks)))
    kept = np.minimum(count, 2 * min_points)  # "do not add point" beyond 2 x min_points (map.cpp:30-33)
    ok = kept >= min_points
    start, kept = start[ok], kept[ok]
    idx = (start + kept)[:, None] - min_points + np.arange(min_points)[None, :]  # the last min_points kept points
    train = points[order[idx]]
    allpts = [points[order[s:s + k]] for s, k in zip(start, kept)]
    return ks[start], train, allpts


def surface_direction(allpts, eigen_1: float = 1.0):
    """-> (direction [V] in {0,1,2}, is_surface [V])."""
    V = len(allpts)
    cov = np.empty((V, 3, 3))
    for i, p in enumerate(allpts):
        d = p - p.mean(axis=0)
        cov[i] = d.T @ d / len(p)
    lam, vec = np.linalg.eigh(cov)  # ascending
    vmin = vec[:, :, 0]
    ang = np.arccos(np.clip(vmin, -1.0, 1.0))
    ang = np.where(ang > np.pi / 2, np.pi - ang, ang)
    direction = ang.argmin(axis=1)
    with np.errstate(divide="ignore", invalid="ignore"):
        is_surface = lam[:, 2] / lam[:, 1] > eigen_1
    return direction, is_surface


_AXES = {0: (1, 2, 0), 1: (2, 0, 1), 2: (0, 1, 2)}  # direction -> (train x axis, train y axis, regressed axis)


def gp_regress(voxel_idx, train, direction, grid=0.2, num_gp_side=12, kernel_size=1.0, variance_sensor=0.05,
               full_cover=False):
    """Per voxel: GP posterior mean on the num_gp_side^2 lattice and the diagonal of K*^T Ky^-1 K*.
    -> (test points [V,n^2,3] float32, variance entry [V,n^2] float32)."""
    V, T, _ = train.shape
    n2 = num_gp_side * num_gp_side
    tr = train.astype(np.float32)
    ax = np.array([_AXES[int(d)] for d in direction])
    rows = np.arange(V)[:, None]
    tx, ty, tf = tr[rows, np.arange(T)[None, :], ax[:, 0:1]], tr[rows, np.arange(T)[None, :], ax[:, 1:2]], \
        tr[rows, np.arange(T)[None, :], ax[:, 2:3]]
    mean = tf.mean(axis=1, keepdims=True, dtype=np.float32)
    tf = tf - mean
    lo = (voxel_idx * grid).astype(np.float32)  # region minimum corner
    interval = np.float32(grid / (num_gp_side - 1 if full_cover else num_gp_side))
    i = np.arange(n2)
    off = np.float32(0.0 if full_cover else 0.5)
    gx = lo[rows, ax[:, 0:1]] + interval * ((i // num_gp_side)[None, :] + off)
    gy = lo[rows, ax[:, 1:2]] + interval * ((i % num_gp_side)[None, :] + off)
    d_tt = np.sqrt((tx[:, :, None] - tx[:, None, :]) ** 2 + (ty[:, :, None] - ty[:, None, :]) ** 2)
    ky = np.exp(-np.float32(kernel_size) * d_tt) + np.eye(T, dtype=np.float32)[None] * np.float32(variance_sensor) ** 2
    d_st = np.sqrt((gx[:, :, None] - tx[:, None, :]) ** 2 + (gy[:, :, None] - ty[:, None, :]) ** 2)
    ks = np.exp(-np.float32(kernel_size) * d_st)                       # [V, n2, T]
    kky = np.linalg.solve(ky, ks.transpose(0, 2, 1)).astype(np.float32)  # Ky^-1 K*   [V, T, n2]
    fstar = np.einsum("vt,vtn->vn", tf, kky) + mean                    # posterior mean + removed mean
    kvar = np.einsum("vnt,vtn->vn", ks, kky)                           # diag(K*^T Ky^-1 K*)
    pts = np.empty((V, n2, 3), np.float32)
    pts[rows, i[None, :], ax[:, 0:1]] = gx
    pts[rows, i[None, :], ax[:, 1:2]] = gy
    pts[rows, i[None, :], ax[:, 2:3]] = fstar
    return pts, kvar.astype(np.float32)


def variance_centred_init(pts, kvar, num_gp_side=12, neighbour=3):
    """fastInitial3DGS -> (means [V*g*g,3], covariances [V*g*g,3,3]), g = num_gp_side / neighbour."""
    V = pts.shape[0]
    g = num_gp_side // neighbour
    P = pts.reshape(V, g, neighbour, g, neighbour, 3).transpose(0, 1, 3, 2, 4, 5).reshape(V, g * g, neighbour ** 2, 3)
    W = (1.0 / kvar).reshape(V, g, neighbour, g, neighbour).transpose(0, 1, 3, 2, 4).reshape(V, g * g, neighbour ** 2)
    wsum = W.sum(axis=2, keepdims=True)
    mean = (P * W[..., None]).sum(axis=2) / wsum
    c = P - mean[:, :, None, :]
    cov = np.einsum("vbki,vbkj->vbij", c * W[..., None], c) / wsum[..., None]
    return mean.reshape(-1, 3).astype(np.float32), cov.reshape(-1, 3, 3).astype(np.float32)


def gaussians_from_init(means, covs, rgb_u8, scale_factor=3.0, sh_degree=3):
    """Activated parameters the rasterizer consumes (GaussianModel::Create_from_pcd + the activations of
    include/gs/gs/gaussian.cuh:42-54): scales = exp(log(sqrt(diag * scale_factor))), normalised identity quaternion,
    sigmoid(logit 0.5)."""
    n = means.shape[0]
    diag = np.stack([covs[:, 0, 0], covs[:, 1, 1], covs[:, 2, 2]], axis=1)
    scales = np.sqrt(np.maximum(diag, 0.0) * scale_factor).astype(np.float32)
    rots = np.zeros((n, 4), np.float32)
    rots[:, 0] = 1.0
    opac = np.full((n, 1), 0.5, np.float32)
    M = (sh_degree + 1) ** 2
    shs = np.zeros((n, M, 3), np.float32)
    shs[:, 0, :] = (rgb_u8.astype(np.float32) / 255.0 - 0.5) / SH_C0
    return dict(means3D=means.astype(np.float32), scales=scales, rotations=rots, opacities=opac, shs=shs)


def vgpr_frame(points: np.ndarray, colour_of: Callable[[np.ndarray], np.ndarray], grid=0.2, min_points=10,
               num_gp_side=12, neighbour=3, kernel_size=1.0, variance_sensor=0.05, eigen_1=1.0, max_var_mean=0.30,
               scale_factor=3.0, sh_degree=3) -> Dict[str, np.ndarray]:
    """One frame of the reference's producer side (lioOptimization.cpp:1201 gsPointCloudUpdate ->
    GpMap::dividePointsIntoCellInitMap -> gpProcess::forward_gp3d), on the host. `colour_of(xyz[n,3])` returns
    uint8 rgb [n,3] and a validity mask [n] (the reference projects into the current image and drops points that
    fall outside it, gpprocess.cu:834-837,936-969)."""
    t0 = time.perf_counter()
    vidx, train, allpts = bucket_points(points, grid, min_points)
    direction, is_surface = surface_direction(allpts, eigen_1)
    keep = is_surface
    vidx, train, direction = vidx[keep], train[keep], direction[keep]
    t1 = time.perf_counter()
    pts, kvar = gp_regress(vidx, train, direction, grid, num_gp_side, kernel_size, variance_sensor)
    var_mean = 1.0 - kvar.mean(axis=1)
    t2 = time.perf_counter()
    means, covs = variance_centred_init(pts, kvar, num_gp_side, neighbour)
    rgb, valid = colour_of(means)
    g = gaussians_from_init(means[valid], covs[valid], rgb[valid], scale_factor, sh_degree)
    t3 = time.perf_counter()
    g["timing"] = dict(bucket_pca_s=t1 - t0, gp_s=t2 - t1, init_s=t3 - t2, voxels=int(len(vidx)),
                       needs_more_points=int((var_mean > max_var_mean).sum()))
    return g


# ------------------------------------------------------------------------------------------------- image loss
def _ssim_window(window_size=11, sigma=1.5):
    """loss_utils.cuh:24-37 — note the reference's own (non-standard) tap positions floor((x - window) / 2)."""
    import torch
    x = torch.arange(window_size, dtype=torch.float32)
    g = torch.exp(-(torch.floor((x - window_size) / 2.0) ** 2) / (2.0 * sigma * sigma))
    g = (g / g.sum()).unsqueeze(1)
    return g.mm(g.t()).unsqueeze(0).unsqueeze(0)


def image_loss(render, target, lambda_dssim=0.2, window_size=11):
    """(1 - l) * L1 + l * (1 - SSIM) on [3,H,W] torch CPU tensors; returns the scalar loss tensor."""
    import torch
    import torch.nn.functional as F
    c1, c2 = 0.01 ** 2, 0.03 ** 2
    w = _ssim_window(window_size).expand(3, 1, window_size, window_size).contiguous()
    a, b = render.unsqueeze(0), target.unsqueeze(0)
    conv = lambda t: F.conv2d(t, w, padding=window_size // 2, groups=3)  # noqa: E731
    mu1, mu2 = conv(a), conv(b)
    s1, s2, s12 = conv(a * a) - mu1 * mu1, conv(b * b) - mu2 * mu2, conv(a * b) - mu1 * mu2
    ssim = (((2 * mu1 * mu2 + c1) * (2 * s12 + c2)) / ((mu1 * mu1 + mu2 * mu2 + c1) * (s1 + s2 + c2))).mean()
    l1 = (render - target).abs().mean()
    return (1.0 - lambda_dssim) * l1 + lambda_dssim * (1.0 - ssim)
