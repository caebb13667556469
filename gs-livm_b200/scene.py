"""Synthetic NTU-VIRAL-shaped scenes and GS-LIVM camera conventions.

Everything is generated on the host with numpy's PCG64 from an explicit seed, so the same (P, seed)
gives the same scene on every box; tensors are float32 torch tensors the caller moves to the GPU.

Geometry mimics what GS-LIVM's voxel-GP front end emits for an outdoor sequence: a ground plane and
vertical facades, 0.2 m voxels (config/basic_common.yaml:72), 16 Gaussians per voxel on a 4x4 lattice
(0.05 m pitch), flat splats (thin along the surface normal), a few percent oversized or behind the camera
so that both of the reference's culls (forward.cu:223-229) are exercised.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List

import numpy as np
import torch

SH_C0 = 0.28209479177387814

# config/ntu.yaml:25-27 (sensor-native 752x480)
NTU = dict(width=752, height=480, fx=425.0259, fy=426.7976)
# config/botanic_garden.yaml:22-27 intrinsics x2 (sensor-native 1920x1200)
BOTANIC = dict(width=1920, height=1200, fx=1285.84, fy=1283.84)


@dataclass
class Camera:
    """What GS-LIVM's Camera hands to render() (src/gs/camera.cu:36-48,57-82)."""
    width: int
    height: int
    tanfovx: float
    tanfovy: float
    viewmatrix: torch.Tensor   # [4,4] = Tcw^T  (Eigen column-major blob read row-major)
    projmatrix: torch.Tensor   # [4,4] = (P @ Tcw)^T
    camera_center: torch.Tensor  # [3]

    def to(self, device) -> "Camera":
        return Camera(self.width, self.height, self.tanfovx, self.tanfovy, self.viewmatrix.to(device),
                      self.projmatrix.to(device), self.camera_center.to(device))


def projection_matrix(znear: float, zfar: float, fovx: float, fovy: float) -> np.ndarray:
    """getProjectionMatrix (camera.cu:57-82): symmetric frustum from the FoV only (cx, cy ignored)."""
    tan_y, tan_x = math.tan(fovy / 2.0), math.tan(fovx / 2.0)
    top, right = tan_y * znear, tan_x * znear
    bottom, left = -top, -right
    P = np.zeros((4, 4), dtype=np.float32)
    P[0, 0] = 2.0 * znear / (right - left)
    P[1, 1] = 2.0 * znear / (top - bottom)
    P[0, 2] = (right + left) / (right - left)
    P[1, 2] = (top + bottom) / (top - bottom)
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


def make_camera(R_wc: np.ndarray, t_wc: np.ndarray, width: int, height: int, fx: float, fy: float,
                znear: float = 0.01, zfar: float = 100.0) -> Camera:
    """R_wc, t_wc: camera-to-world rotation / camera position (Camera::_R, _T)."""
    fovx = 2.0 * math.atan(width / (2.0 * fx))   # focal2fov, camera.cu:88-90
    fovy = 2.0 * math.atan(height / (2.0 * fy))
    Tcw = np.eye(4, dtype=np.float32)
    Tcw[:3, :3] = R_wc.T
    Tcw[:3, 3] = -R_wc.T @ t_wc
    view = torch.from_numpy(np.ascontiguousarray(Tcw.T))
    proj = torch.from_numpy(np.ascontiguousarray(projection_matrix(znear, zfar, fovx, fovy).T))
    full = view @ proj
    center = torch.linalg.inv(view)[3, :3].contiguous()
    return Camera(width, height, math.tan(fovx * 0.5), math.tan(fovy * 0.5), view.contiguous(), full.contiguous(),
                  center)


def track_camera(step: int, intr: Dict = NTU, spacing: float = 0.3, yaw_deg: float = 3.0, height: float = 1.5,
                 start: np.ndarray | None = None) -> Camera:
    """Pose `step` of the keyframe window: `spacing` m and `yaw_deg` degrees apart along +x, z up."""
    yaw = math.radians(yaw_deg) * step
    fwd = np.array([math.cos(yaw), math.sin(yaw), 0.0])
    down = np.array([0.0, 0.0, -1.0])
    right = np.cross(down, fwd)  # x_cam = y_cam x z_cam
    R_wc = np.stack([right, down, fwd], axis=1).astype(np.float32)
    pos = (np.array([0.0, 0.0, height]) if start is None else start) + np.array([spacing * step, 0.0, 0.0])
    return make_camera(R_wc, pos.astype(np.float32), intr["width"], intr["height"], intr["fx"], intr["fy"])


def window_cameras(n: int = 8, intr: Dict = NTU, **kw) -> List[Camera]:
    return [track_camera(i, intr, **kw) for i in range(n)]


def _planes(length: float, width: float, height: float):
    """(origin, u axis, v axis, normal axis index, u extent, v extent) of the ground and facades of a street
    canyon of `length` x `width` with `height`-tall facades; everything lies within ~3-70 m of the track."""
    L, Wd, Ht = length, width, height
    y0, y1 = -0.45 * Wd, 0.55 * Wd
    return [
        (np.array([-0.15 * L, y0, 0.0]), 0, 1, 2, 1.15 * L, Wd),        # ground z = 0
        (np.array([-0.15 * L, y1, 0.0]), 0, 2, 1, 1.15 * L, Ht),        # left facade
        (np.array([-0.15 * L, y0, 0.0]), 0, 2, 1, 1.15 * L, Ht),        # right facade
        (np.array([0.55 * L, 1.0, 0.0]), 1, 2, 0, 0.4 * Wd, 0.7 * Ht),  # kiosk facing the track
        (np.array([0.25 * L + 3.0, y0, 0.0]), 1, 2, 0, 0.3 * Wd, 0.6 * Ht),  # near wall stub, >= 3 m ahead
        (np.array([L, y0, 0.0]), 1, 2, 0, Wd, 1.3 * Ht),                # end facade x = L
    ]


LIDAR_FULL_DENSITY_RANGE = 25.0  # m; beyond it the map thins out like LiDAR returns do, ~ 1/d^2


def _keep_probability(points: np.ndarray) -> np.ndarray:
    """GS-LIVM only spawns Gaussians in voxels that collected enough LiDAR points (src/gp3d/map.cpp:51-111),
    so map density falls off with range from the sensor track; modelled as min(1, (25 m / d)^2)."""
    d = np.linalg.norm(points - np.array([0.0, 0.0, 1.5], dtype=points.dtype), axis=1)
    return np.minimum(1.0, (LIDAR_FULL_DENSITY_RANGE / np.maximum(d, 1e-3)) ** 2)


def _expected_count(length: float, width: float, height: float, pitch: float) -> float:
    total = 0.0
    cell = 0.5
    for (org, ua, va, na, ue, ve) in _planes(length, width, height):
        nu, nv = max(1, int(ue / cell)), max(1, int(ve / cell))
        uu, vv = np.meshgrid((np.arange(nu) + 0.5) * (ue / nu), (np.arange(nv) + 0.5) * (ve / nv), indexing="ij")
        p = np.zeros((nu * nv, 3))
        p[:, ua], p[:, va] = uu.ravel(), vv.ravel()
        p += org
        total += float(_keep_probability(p).sum()) * (ue / nu) * (ve / nv) / (pitch * pitch)
    return total


def _fit_geometry(P: int, pitch: float):
    """Grow the canyon (length first, capped at 70 m; then width/height) until it holds ~P Gaussians."""
    length, width, height = 12.0, 11.0, 6.0
    for _ in range(200):
        if _expected_count(length, width, height, pitch) >= P * 1.03:
            return length, width, height
        if length < 70.0:
            length = min(70.0, length * 1.1)
        elif width < 40.0:
            width, height = width * 1.08, min(height * 1.08, 16.0)
        else:
            return None
    return None


def make_scene(P: int, seed: int, sh_degree: int = 3, heavy_overlap: bool = False) -> Dict[str, torch.Tensor]:
    """P Gaussians: activated scales/opacities, unit quaternions, SH coefficients [P, (deg+1)^2, 3]."""
    rng = np.random.default_rng(seed)
    # 0.05 m lattice pitch (16 Gaussians per 0.2 m voxel) while the canyon can hold P points within ~70 m of the
    # track; beyond that (multi-million maps) the lattice gets denser instead of the scene reaching the horizon
    pitch = 0.05
    fit = _fit_geometry(P, pitch)
    while fit is None:
        pitch *= 0.9
        fit = _fit_geometry(P, pitch)
    length, width, height = fit
    pts, normal_axis = [], []
    for (org, ua, va, na, ue, ve) in _planes(length, width, height):
        nu, nv = int(ue / pitch), int(ve / pitch)
        uu, vv = np.meshgrid(np.arange(nu, dtype=np.float32) * pitch, np.arange(nv, dtype=np.float32) * pitch,
                             indexing="ij")
        p = np.zeros((nu * nv, 3), dtype=np.float32)
        p[:, ua] = uu.ravel()
        p[:, va] = vv.ravel()
        p += org.astype(np.float32)
        pts.append(p)
        normal_axis.append(np.full((nu * nv,), na, dtype=np.int64))
    pts = np.concatenate(pts)
    normal_axis = np.concatenate(normal_axis)
    kept = np.nonzero(rng.random(pts.shape[0]) < _keep_probability(pts))[0]
    if kept.shape[0] < P:
        raise RuntimeError("scene generator: lattice too small (%d < %d)" % (kept.shape[0], P))
    sel = kept[rng.choice(kept.shape[0], size=P, replace=False)]
    sel.sort()
    means = pts[sel]
    naxis = normal_axis[sel]
    rows = np.arange(P)
    means[rows, naxis] += rng.normal(0.0, 0.005, size=P).astype(np.float32)  # 5 mm jitter along the normal

    scales = rng.uniform(0.01, 0.05, size=(P, 3)).astype(np.float32)
    scales[rows, naxis] = rng.uniform(0.002, 0.01, size=P).astype(np.float32)
    if heavy_overlap:
        scales = np.minimum(scales * 4.0, 0.29).astype(np.float32)
    big = rng.random(P) < 0.005  # 0.5 % trip scale_abnormal (> 0.3 m)
    scales[big, 0] = rng.uniform(0.3001, 0.5, size=int(big.sum())).astype(np.float32)

    rot = np.zeros((P, 4), dtype=np.float32)
    rot[:, 0] = 1.0
    rot += rng.normal(0.0, 0.1, size=(P, 4)).astype(np.float32)
    rot /= np.linalg.norm(rot, axis=1, keepdims=True)

    opacity = (1.0 / (1.0 + np.exp(-rng.normal(0.0, 1.5, size=(P, 1))))).astype(np.float32)

    M = (sh_degree + 1) ** 2
    shs = np.zeros((P, M, 3), dtype=np.float32)
    shs[:, 0, :] = ((rng.random((P, 3)) - 0.5) / SH_C0).astype(np.float32)
    if M > 1:
        shs[:, 1:, :] = rng.normal(0.0, 0.05, size=(P, M - 1, 3)).astype(np.float32)

    out = dict(means3D=means, scales=scales, rotations=rot.astype(np.float32), opacities=opacity, shs=shs)
    t = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in out.items()}
    t["background"] = torch.ones(3, dtype=torch.float32)  # white_background = true (parameters.cuh:43)
    t["track_start"] = torch.tensor([0.0, 0.0, 1.5])
    return t


def scene_camera(scene: Dict[str, torch.Tensor], step: int = 0, intr: Dict = NTU, close_up: bool = False) -> Camera:
    """Pose `step` on the track of `scene`; close_up=True parks the camera 2 m in front of a facade
    (the 'heavy overlap' configuration)."""
    if close_up:
        fwd = np.array([0.0, 1.0, 0.0])
        down = np.array([0.0, 0.0, -1.0])
        right = np.cross(down, fwd)
        R_wc = np.stack([right, down, fwd], axis=1).astype(np.float32)
        pos = np.array([3.0 + 0.3 * step, 4.0, 1.5], dtype=np.float32)
        return make_camera(R_wc, pos, intr["width"], intr["height"], intr["fx"], intr["fy"])
    return track_camera(step, intr, start=scene["track_start"].cpu().numpy().astype(np.float64))


def settings_for(cam: Camera, background: torch.Tensor, sh_degree: int, scale_modifier: float = 1.0):
    from .rasterizer import GaussianRasterizationSettings
    return GaussianRasterizationSettings(cam.height, cam.width, cam.tanfovx, cam.tanfovy, background, scale_modifier,
                                         cam.viewmatrix, cam.projmatrix, sh_degree, cam.camera_center, False)


# ---------------------------------------------------------------------------------------------------------
# configs[0] of BASELINE.json: one NTU-VIRAL-shaped frame as the producer side sees it — a LiDAR sweep of the
# street canyon plus the colour image of the keyframe camera. The VGPR / Gaussian-initialisation restatement that
# turns it into Gaussians is test infrastructure (oracle/vgpr.py); this only makes the synthetic sensor data.
def lidar_frame(n_points: int, seed: int, length: float = 30.0, width: float = 11.0, height: float = 6.0,
                noise: float = 0.005) -> np.ndarray:
    """n_points returns on the canyon's surfaces, area-uniform, `noise` m of range noise along the normal."""
    rng = np.random.default_rng(seed)
    planes = _planes(length, width, height)
    area = np.array([ue * ve for (_, _, _, _, ue, ve) in planes])
    counts = rng.multinomial(n_points, area / area.sum())
    out = []
    for (org, ua, va, na, ue, ve), c in zip(planes, counts):
        p = np.zeros((c, 3))
        p[:, ua] = rng.random(c) * ue
        p[:, va] = rng.random(c) * ve
        p += org
        p[:, na] += rng.normal(0.0, noise, size=c)
        out.append(p)
    pts = np.concatenate(out)
    return pts[rng.permutation(pts.shape[0])]


def frame_colour(cam: Camera, intr: Dict = NTU):
    """Colour look-up the way gpProcess::getColors does it (src/gp3d/gpprocess.cu:917-969, without lens
    distortion): project into the keyframe image, drop what falls outside. The image is procedural."""
    view = cam.viewmatrix.cpu().numpy().astype(np.float64)  # stored transposed: row-vector convention

    def colour_of(xyz: np.ndarray):
        pc = np.concatenate([xyz.astype(np.float64), np.ones((xyz.shape[0], 1))], axis=1) @ view
        z = pc[:, 2]
        with np.errstate(divide="ignore", invalid="ignore"):
            u = (intr["fx"] * pc[:, 0] / z + 0.5 * intr["width"]).astype(np.int64)
            v = (intr["fy"] * pc[:, 1] / z + 0.5 * intr["height"]).astype(np.int64)
        valid = (z > 0.2) & (u >= 0) & (u < intr["width"]) & (v >= 0) & (v < intr["height"])
        r = 127.5 + 100.0 * np.sin(0.045 * u) * np.cos(0.03 * v)
        g = 127.5 + 100.0 * np.sin(0.02 * u + 0.05 * v)
        b = 127.5 + 100.0 * np.cos(0.06 * v)
        rgb = np.clip(np.stack([r, g, b], axis=1), 0, 255).astype(np.uint8)
        return rgb, valid

    return colour_of


def procedural_image(intr: Dict = NTU) -> np.ndarray:
    """The keyframe image frame_colour() looks colours up in, as an array [H,W,3] uint8."""
    u, v = np.meshgrid(np.arange(intr["width"]), np.arange(intr["height"]))
    r = 127.5 + 100.0 * np.sin(0.045 * u) * np.cos(0.03 * v)
    g = 127.5 + 100.0 * np.sin(0.02 * u + 0.05 * v)
    b = 127.5 + 100.0 * np.cos(0.06 * v)
    return np.clip(np.stack([r, g, b], axis=2), 0, 255).astype(np.uint8)
