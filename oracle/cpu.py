"""TEST INFRASTRUCTURE — numpy/ctypes driver of Oracle B (oracle/gsr_oracle.c, the scalar CPU restatement).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgsr_oracle_cpu.so")
_lib = None


class _Inputs(C.Structure):
    _fields_ = [("P", C.c_int), ("D", C.c_int), ("M", C.c_int), ("W", C.c_int), ("H", C.c_int),
                ("bg", C.c_void_p), ("means3D", C.c_void_p), ("shs", C.c_void_p), ("colors_precomp", C.c_void_p),
                ("opacities", C.c_void_p), ("scales", C.c_void_p), ("rotations", C.c_void_p),
                ("cov3D_precomp", C.c_void_p), ("view", C.c_void_p), ("proj", C.c_void_p), ("campos", C.c_void_p),
                ("scale_modifier", C.c_float), ("tan_fovx", C.c_float), ("tan_fovy", C.c_float)]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "gsr_oracle.c")
    hdr = os.path.join(_HERE, "gsr_oracle.h")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < max(os.path.getmtime(src),
                                                                               os.path.getmtime(hdr)):
        subprocess.check_call(["make", "-C", _HERE, "cpu"], stdout=subprocess.DEVNULL)
    return LIB_PATH


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB_PATH)
        _lib.gsro_scan.restype = C.c_int64
    return _lib


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a) -> Optional[np.ndarray]:
    if a is None:
        return None
    if hasattr(a, "detach"):
        a = a.detach().cpu().numpy()
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a if a.size else None


class OracleScene:
    """Inputs of one render in the reference's layout (numpy float32, contiguous)."""

    def __init__(self, means3D, opacities, background, viewmatrix, projmatrix, campos, tan_fovx, tan_fovy, width,
                 height, shs=None, sh_degree=0, colors_precomp=None, scales=None, rotations=None, cov3D_precomp=None,
                 scale_modifier=1.0):
        self.means3D, self.opacities, self.bg = _f32(means3D), _f32(opacities), _f32(background)
        self.view, self.proj, self.campos = _f32(viewmatrix), _f32(projmatrix), _f32(campos)
        self.shs, self.colors_precomp = _f32(shs), _f32(colors_precomp)
        self.scales, self.rotations, self.cov3D_precomp = _f32(scales), _f32(rotations), _f32(cov3D_precomp)
        self.P = int(self.means3D.shape[0])
        self.D = int(sh_degree)
        self.M = int(self.shs.shape[1]) if self.shs is not None else 0
        self.W, self.H = int(width), int(height)
        self.tan_fovx, self.tan_fovy, self.scale_modifier = float(tan_fovx), float(tan_fovy), float(scale_modifier)

    def c_inputs(self) -> _Inputs:
        s = _Inputs()
        s.P, s.D, s.M, s.W, s.H = self.P, self.D, self.M, self.W, self.H
        s.bg, s.means3D, s.shs, s.colors_precomp = _p(self.bg), _p(self.means3D), _p(self.shs), _p(self.colors_precomp)
        s.opacities, s.scales, s.rotations = _p(self.opacities), _p(self.scales), _p(self.rotations)
        s.cov3D_precomp, s.view, s.proj, s.campos = _p(self.cov3D_precomp), _p(self.view), _p(self.proj), _p(self.campos)
        s.scale_modifier, s.tan_fovx, s.tan_fovy = self.scale_modifier, self.tan_fovx, self.tan_fovy
        return s


def forward(sc: OracleScene, nthreads: int = 1, stages: str = "all") -> Dict[str, np.ndarray]:
    """Run preprocess -> scan -> bin -> blend. Returns every intermediate array, reference names."""
    lib = load()
    P, W, H = sc.P, sc.W, sc.H
    gx, gy = (W + 15) // 16, (H + 15) // 16
    o: Dict[str, np.ndarray] = {}
    o["radii"] = np.zeros(P, np.int32)
    o["means2D"] = np.zeros((P, 2), np.float32)
    o["depths"] = np.zeros(P, np.float32)
    o["cov3D"] = np.zeros((P, 6), np.float32)
    o["conic_opacity"] = np.zeros((P, 4), np.float32)
    o["rgb"] = np.zeros((P, 3), np.float32)
    o["clamped"] = np.zeros((P, 3), np.uint8)
    o["tiles_touched"] = np.zeros(P, np.uint32)
    inp = sc.c_inputs()
    lib.gsro_preprocess(C.byref(inp), _p(o["radii"]), _p(o["means2D"]), _p(o["depths"]), _p(o["cov3D"]),
                        _p(o["conic_opacity"]), _p(o["rgb"]), _p(o["clamped"]), _p(o["tiles_touched"]))
    o["point_offsets"] = np.zeros(P, np.uint32)
    R = int(lib.gsro_scan(P, _p(o["tiles_touched"]), _p(o["point_offsets"])))
    o["num_rendered"] = R
    if stages == "preprocess":
        return o
    n = max(R, 1)
    o["point_list_keys_unsorted"] = np.zeros(n, np.uint64)
    o["point_list_unsorted"] = np.zeros(n, np.uint32)
    o["point_list_keys"] = np.zeros(n, np.uint64)
    o["point_list"] = np.zeros(n, np.uint32)
    o["ranges"] = np.zeros((gx * gy, 2), np.uint32)
    rc = lib.gsro_bin(P, C.c_int64(R), W, H, _p(o["radii"]), _p(o["means2D"]), _p(o["depths"]), _p(o["point_offsets"]),
                      _p(o["point_list_keys_unsorted"]), _p(o["point_list_unsorted"]), _p(o["point_list_keys"]),
                      _p(o["point_list"]), _p(o["ranges"]))
    if rc != 0:
        raise MemoryError("gsro_bin")
    for k in ("point_list_keys_unsorted", "point_list_unsorted", "point_list_keys", "point_list"):
        o[k] = o[k][:R]
    if stages == "bin":
        return o
    feats = sc.colors_precomp if sc.colors_precomp is not None else o["rgb"]
    o["out_color"] = np.zeros((3, H, W), np.float32)
    o["out_depth"] = np.zeros((1, H, W), np.float32)
    o["out_acc"] = np.zeros((1, H, W), np.float32)
    o["final_T"] = np.zeros(H * W, np.float32)
    o["n_contrib"] = np.zeros(H * W, np.uint32)
    o["fragile"] = np.zeros(H * W, np.uint8)
    o["counters"] = np.zeros(4, np.uint64)
    pl = o["point_list"] if R > 0 else np.zeros(1, np.uint32)
    lib.gsro_blend_fwd(W, H, _p(o["ranges"]), _p(pl), _p(o["means2D"]), _p(o["conic_opacity"]), _p(feats),
                       _p(o["depths"]), _p(sc.bg), _p(o["out_color"]), _p(o["out_depth"]), _p(o["out_acc"]),
                       _p(o["final_T"]), _p(o["n_contrib"]), _p(o["fragile"]), _p(o["counters"]), int(nthreads))
    return o


def backward(sc: OracleScene, fwd: Dict[str, np.ndarray], dL_dpix, dL_dacc, nthreads: int = 1) -> Dict[str, np.ndarray]:
    """float64 gradients in the reference's output shapes."""
    lib = load()
    P, W, H, M = sc.P, sc.W, sc.H, sc.M
    dpix, dacc = _f32(dL_dpix), _f32(dL_dacc)
    g = dict(dL_dmean2D=np.zeros((P, 3)), dL_dconic=np.zeros((P, 2, 2)), dL_dopacity=np.zeros((P, 1)),
             dL_dcolor=np.zeros((P, 3)), dL_dmean3D=np.zeros((P, 3)), dL_dcov3D=np.zeros((P, 6)),
             dL_dsh=np.zeros((P, max(M, 1), 3)), dL_dscale=np.zeros((P, 3)), dL_drot=np.zeros((P, 4)))
    feats = sc.colors_precomp if sc.colors_precomp is not None else fwd["rgb"]
    pl = fwd["point_list"] if fwd["num_rendered"] > 0 else np.zeros(1, np.uint32)
    counters = np.zeros(3, np.uint64)
    lib.gsro_blend_bwd(W, H, _p(fwd["ranges"]), _p(pl), _p(fwd["means2D"]), _p(fwd["conic_opacity"]), _p(feats),
                       _p(sc.bg), _p(fwd["final_T"]), _p(fwd["n_contrib"]), _p(dpix), _p(dacc), _p(g["dL_dmean2D"]),
                       _p(g["dL_dconic"]), _p(g["dL_dopacity"]), _p(g["dL_dcolor"]), _p(counters), int(nthreads))
    inp = sc.c_inputs()
    lib.gsro_geom_bwd(C.byref(inp), _p(fwd["radii"]), _p(fwd["cov3D"]), _p(fwd["clamped"]), _p(g["dL_dmean2D"]),
                      _p(g["dL_dconic"]), _p(g["dL_dcolor"]), _p(g["dL_dmean3D"]), _p(g["dL_dcov3D"]), _p(g["dL_dsh"]),
                      _p(g["dL_dscale"]), _p(g["dL_drot"]))
    g["dL_dsh"] = g["dL_dsh"][:, :M, :]
    g["bwd_counters"] = counters
    return g


def mark_visible(means3D, viewmatrix) -> np.ndarray:
    lib = load()
    m, v = _f32(means3D), _f32(viewmatrix)
    out = np.zeros(m.shape[0], np.uint8)
    lib.gsro_mark_visible(m.shape[0], _p(m), _p(v), _p(out))
    return out.astype(bool)
