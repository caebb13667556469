"""TEST INFRASTRUCTURE — loader/driver of Oracle A: the UNMODIFIED reference rasterizer compiled for
sm_100a into oracle/_ref/gsr_ref_ext.so by oracle/Makefile (`make ref`, needs /root/reference; the
GPU box only sees the prebuilt .so). Only tests/, smoke() and bench.py's reference arm import this.
"""
from __future__ import annotations

import importlib.util
import os
from typing import Dict, Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
EXT_PATH = os.path.join(_HERE, "_ref", "gsr_ref_ext.so")
_ext = None


def available() -> bool:
    return os.path.exists(EXT_PATH)


def load():
    global _ext
    if _ext is None:
        if not available():
            raise ImportError(f"{EXT_PATH} missing: run `make -C oracle ref` where /root/reference exists")
        spec = importlib.util.spec_from_file_location("gsr_ref_ext", EXT_PATH)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _ext = mod
    return _ext


def _e(t: Optional[torch.Tensor], dev) -> torch.Tensor:
    return torch.empty(0, dtype=torch.float32, device=dev) if t is None else t


def forward(bg, means3D, opacities, scales, rotations, shs, sh_degree, viewmatrix, projmatrix, campos, tanfovx,
            tanfovy, height, width, colors_precomp=None, cov3D_precomp=None, scale_modifier=1.0,
            with_state: bool = True) -> Dict[str, torch.Tensor]:
    """Reference RasterizeGaussiansCUDA + (optionally) every intermediate array of its three buffers."""
    ext = load()
    dev = means3D.device
    R, color, depth, acc, radii, geom, binning, img = ext.forward(
        bg, means3D, _e(colors_precomp, dev), opacities, _e(scales, dev), _e(rotations, dev), float(scale_modifier),
        _e(cov3D_precomp, dev), viewmatrix, projmatrix, float(tanfovx), float(tanfovy), int(height), int(width),
        _e(shs, dev), int(sh_degree), campos, False, False)
    out = dict(num_rendered=R, out_color=color, out_depth=depth, out_acc=acc, radii=radii, geomBuffer=geom,
               binningBuffer=binning, imgBuffer=img)
    if with_state:
        P = means3D.size(0)
        tiles = ((width + 15) // 16) * ((height + 15) // 16)
        out.update(ext.geom_state(geom, P))
        out.update(ext.binning_state(binning, R))
        out.update(ext.image_state(img, width * height, tiles))
    return out


def backward(fwd: Dict[str, torch.Tensor], bg, means3D, scales, rotations, shs, sh_degree, viewmatrix, projmatrix,
             campos, tanfovx, tanfovy, dL_dcolor, dL_dacc, colors_precomp=None, cov3D_precomp=None,
             scale_modifier=1.0) -> Dict[str, torch.Tensor]:
    ext = load()
    dev = means3D.device
    names = ("dL_dmean2D", "dL_dcolor", "dL_dopacity", "dL_dmean3D", "dL_dcov3D", "dL_dsh", "dL_dscale", "dL_drot")
    res = ext.backward(bg, means3D, fwd["radii"], _e(colors_precomp, dev), _e(scales, dev), _e(rotations, dev),
                       float(scale_modifier), _e(cov3D_precomp, dev), viewmatrix, projmatrix, float(tanfovx),
                       float(tanfovy), dL_dcolor, dL_dacc, _e(shs, dev), int(sh_degree), campos, fwd["geomBuffer"],
                       int(fwd["num_rendered"]), fwd["binningBuffer"], fwd["imgBuffer"], False)
    return dict(zip(names, res))
