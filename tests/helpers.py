"""Shared drivers for the parity tests: run our C-ABI path, the CPU oracle (B) and the reference CUDA
oracle (A) on the same inputs and hand back numpy dictionaries with the reference's names."""
from __future__ import annotations

import os
from typing import Dict, Optional

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = ("ntu_deg3", "partial_tiles_deg0", "precomp_colors")
GRAD_NAMES = ("dL_dmean2D", "dL_dcolor", "dL_dopacity", "dL_dmean3D", "dL_dcov3D", "dL_dsh", "dL_dscale", "dL_drot")


class Case:
    """Inputs of one render in the reference's layout (CPU torch tensors)."""

    def __init__(self, means3D, scales, rotations, opacities, shs, background, viewmatrix, projmatrix, camera_center,
                 tanfovx, tanfovy, width, height, sh_degree, colors_precomp=None, dL_dcolor=None, dL_dacc=None,
                 cov3D_precomp=None):
        t = lambda a: None if a is None else torch.as_tensor(np.asarray(a)).contiguous()  # noqa: E731
        self.means3D, self.scales, self.rotations, self.opacities = t(means3D), t(scales), t(rotations), t(opacities)
        self.shs, self.background = t(shs), t(background)
        self.viewmatrix, self.projmatrix, self.camera_center = t(viewmatrix), t(projmatrix), t(camera_center)
        self.colors_precomp, self.dL_dcolor, self.dL_dacc = t(colors_precomp), t(dL_dcolor), t(dL_dacc)
        self.cov3D_precomp = t(cov3D_precomp)  # given -> scales / rotations must be None (reference rasterizer.cu:165-169)
        self.tanfovx, self.tanfovy = float(tanfovx), float(tanfovy)
        self.width, self.height, self.sh_degree = int(width), int(height), int(sh_degree)
        self.P = int(self.means3D.shape[0])

    @property
    def use_colors(self):
        return self.colors_precomp is not None

    def with_upstream(self, seed: int = 0) -> "Case":
        g = torch.Generator().manual_seed(seed)
        n = self.height * self.width
        self.dL_dcolor = torch.randn((3, self.height, self.width), generator=g) / n
        self.dL_dacc = torch.randn((1, self.height, self.width), generator=g) / n
        return self


def load_golden(name: str):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    P, deg, W, H, R = (int(v) for v in z["meta"])
    tanx, tany = (float(v) for v in z["tanfov"])
    colors = z["in_colors_precomp"] if "in_colors_precomp" in z.files else None
    case = Case(z["in_means3D"], z["in_scales"], z["in_rotations"], z["in_opacities"], z["in_shs"], z["in_background"],
                z["in_viewmatrix"], z["in_projmatrix"], z["in_camera_center"], tanx, tany, W, H, deg,
                colors_precomp=colors, dL_dcolor=z["in_dL_dcolor"], dL_dacc=z["in_dL_dacc"])
    ref = {k[4:]: z[k] for k in z.files if k.startswith("ref_")}
    ref["num_rendered"] = R
    return case, ref


def scene_case(P: int, seed: int, sh_degree: int = 3, intr: Optional[dict] = None, step: int = 0, heavy=False,
               close_up=False) -> Case:
    import gs_livm_b200  # noqa: F401
    from gs_livm_b200 import scene
    sc = scene.make_scene(P, seed, sh_degree, heavy_overlap=heavy)
    cam = scene.scene_camera(sc, step, intr or scene.NTU, close_up=close_up)
    return Case(sc["means3D"], sc["scales"], sc["rotations"], sc["opacities"], sc["shs"], sc["background"],
                cam.viewmatrix, cam.projmatrix, cam.camera_center, cam.tanfovx, cam.tanfovy, cam.width, cam.height,
                sh_degree).with_upstream(seed + 1)


# ------------------------------------------------------------------------------------------------ oracle B
def run_oracle_cpu(case: Case, backward: bool = True, nthreads: int = 4) -> Dict[str, np.ndarray]:
    from oracle import cpu
    sc = cpu.OracleScene(case.means3D, case.opacities, case.background, case.viewmatrix, case.projmatrix,
                         case.camera_center, case.tanfovx, case.tanfovy, case.width, case.height,
                         shs=None if case.use_colors else case.shs, sh_degree=case.sh_degree,
                         colors_precomp=case.colors_precomp, scales=case.scales, rotations=case.rotations,
                         cov3D_precomp=case.cov3D_precomp)
    out = cpu.forward(sc, nthreads=nthreads)
    if backward and case.dL_dcolor is not None:
        out.update(cpu.backward(sc, out, case.dL_dcolor, case.dL_dacc, nthreads=nthreads))
    return out


# ------------------------------------------------------------------------------------------------ ours
SORT_PATHS = ("tile", "device")
# blend / sort variants of the library (gsr_configure knobs): the pair-sparse kernels with their default warp shapes,
# the same with the FINE shape (four lanes per pixel) forced on every tile, and the defaults with every list longer
# than 255 instances cut into small depth buckets (the long-list path of tile_sort.cu on small scenes)
BLENDS = ("sparse", "sparse-fine", "small-buckets")
_BLEND_KNOBS = {"sparse": {}, "sparse-fine": {"heavy_fwd": 1, "heavy_bwd": 1},
                "small-buckets": {"split_len": 255, "bucket_target": 256}}
_ALL_KNOBS = ("sort_device", "heavy_fwd", "heavy_bwd", "split_len", "bucket_target")


def run_ours(case: Case, dev, backward: bool = True, sort_path: str = "tile", blend: str = "sparse") -> Dict[str, np.ndarray]:
    """Through the tensor-level mirror of the reference operator -> C ABI -> CUDA kernels.

    sort_path "tile": per-tile binning (tile_sort.cu, the default; falls back by itself for very long lists);
    "device": the reference's own sequence scan -> duplicateWithKeys -> device-wide radix sort -> ranges, which
    also leaves the reference's intermediate arrays (point_offsets, unsorted keys/values) behind."""
    import gs_livm_b200  # noqa: F401
    from gs_livm_b200 import _lib, debug, rasterizer
    _lib.configure(**{k: -1 for k in _ALL_KNOBS})
    _lib.configure(sort_device=1 if sort_path == "device" else 0, **_BLEND_KNOBS[blend])
    d = lambda t: torch.empty(0, device=dev) if t is None else t.to(dev)  # noqa: E731
    shs = None if case.use_colors else case.shs
    bg, m3, sc, rot, op = d(case.background), d(case.means3D), d(case.scales), d(case.rotations), d(case.opacities)
    sh, col, view, proj, cam = d(shs), d(case.colors_precomp), d(case.viewmatrix), d(case.projmatrix), d(case.camera_center)
    cov = d(case.cov3D_precomp)
    R, color, depth, acc, radii, gb, bb, ib = rasterizer.rasterize_gaussians_cuda(
        bg, m3, col, op, sc, rot, 1.0, cov, view, proj, case.tanfovx, case.tanfovy,
        case.height, case.width, sh, case.sh_degree, cam, False, False)
    g, b, im = debug.geom_state(gb, case.P), debug.binning_state(bb, R), debug.image_state(ib, case.width, case.height)
    out = dict(sort_path=sort_path, num_rendered=R, out_color=color, out_depth=depth, out_acc=acc, radii=radii, depths=g["depths"],
               means2D=g["means2D"], conic_opacity=g["conic_opacity"], rgb=g["rgb"], clamped_bits=g["clamped"],
               tiles_touched=g["tiles_touched"], point_offsets=g["point_offsets"], extent=g["extent"],
               point_list_keys=b["point_list_keys"], point_list=b["point_list"],
               point_list_keys_unsorted=b["point_list_keys_unsorted"], point_list_unsorted=b["point_list_unsorted"],
               ranges=im["ranges"], accum_alpha=im["accum_alpha"], n_contrib=im["n_contrib"], tile_last=im["tile_last"])
    if backward and case.dL_dcolor is not None:
        grads = rasterizer.rasterize_gaussians_backward_cuda(
            bg, m3, radii, col, sc, rot, 1.0, cov, view, proj, case.tanfovx, case.tanfovy,
            d(case.dL_dcolor), d(case.dL_dacc), sh, case.sh_degree, cam, gb, R, bb, ib, False)
        out.update(dict(zip(GRAD_NAMES, grads)))
    torch.cuda.synchronize()
    _lib.configure(**{k: -1 for k in _ALL_KNOBS})
    return {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else v) for k, v in out.items()}


# ------------------------------------------------------------------------------------------------ oracle A
def ref_available() -> bool:
    from oracle import ref
    return ref.available()


def run_ref_cuda(case: Case, dev, backward: bool = True) -> Dict[str, np.ndarray]:
    from oracle import ref
    d = lambda t: None if t is None else t.to(dev)  # noqa: E731
    shs = None if case.use_colors else d(case.shs)
    fwd = ref.forward(d(case.background), d(case.means3D), d(case.opacities), d(case.scales), d(case.rotations), shs,
                      case.sh_degree, d(case.viewmatrix), d(case.projmatrix), d(case.camera_center), case.tanfovx,
                      case.tanfovy, case.height, case.width, colors_precomp=d(case.colors_precomp))
    out = dict(fwd)
    if backward and case.dL_dcolor is not None:
        out.update(ref.backward(fwd, d(case.background), d(case.means3D), d(case.scales), d(case.rotations), shs,
                                case.sh_degree, d(case.viewmatrix), d(case.projmatrix), d(case.camera_center),
                                case.tanfovx, case.tanfovy, d(case.dL_dcolor), d(case.dL_dacc),
                                colors_precomp=d(case.colors_precomp)))
    torch.cuda.synchronize()
    return {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else v) for k, v in out.items()
            if k not in ("geomBuffer", "binningBuffer", "imgBuffer")}


# ------------------------------------------------------------------------------------------------ comparisons
def bits(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32)).view(np.uint32)


def clamped_bits_to_bool3(bits_u8) -> np.ndarray:
    b = np.asarray(bits_u8).astype(np.uint8)
    return np.stack([(b >> c) & 1 for c in range(3)], axis=1).astype(bool)


def assert_integer_stages_equal(ours: Dict, ref: Dict, check_unsorted: bool = False):
    """Everything that must be bit-exact: radii, tiles, offsets, depth / mean2D bit patterns, sorted keys,
    point list (incl. tie order), tile ranges."""
    assert int(ours["num_rendered"]) == int(ref["num_rendered"]), "num_rendered"
    np.testing.assert_array_equal(ours["radii"], ref["radii"], err_msg="radii")
    vis = np.asarray(ref["radii"]) > 0
    np.testing.assert_array_equal(np.asarray(ours["tiles_touched"]).astype(np.int64),
                                  np.asarray(ref["tiles_touched"]).astype(np.int64), err_msg="tiles_touched")
    device_path = ours.get("sort_path", "device") == "device"  # only that path materialises the reference's intermediates
    if device_path:
        np.testing.assert_array_equal(np.asarray(ours["point_offsets"]).astype(np.int64),
                                      np.asarray(ref["point_offsets"]).astype(np.int64), err_msg="point_offsets")
    np.testing.assert_array_equal(bits(ours["depths"])[vis], bits(ref["depths"])[vis], err_msg="depths (bits)")
    np.testing.assert_array_equal(bits(ours["means2D"])[vis], bits(ref["means2D"])[vis], err_msg="means2D (bits)")
    np.testing.assert_array_equal(np.asarray(ours["point_list_keys"]).astype(np.uint64),
                                  np.asarray(ref["point_list_keys"]).astype(np.uint64), err_msg="sorted keys")
    np.testing.assert_array_equal(np.asarray(ours["point_list"]).astype(np.int64),
                                  np.asarray(ref["point_list"]).astype(np.int64), err_msg="point_list")
    np.testing.assert_array_equal(np.asarray(ours["ranges"]).astype(np.int64),
                                  np.asarray(ref["ranges"]).astype(np.int64), err_msg="tile ranges")
    if check_unsorted and device_path:
        np.testing.assert_array_equal(np.asarray(ours["point_list_keys_unsorted"]).astype(np.uint64),
                                      np.asarray(ref["point_list_keys_unsorted"]).astype(np.uint64))
        np.testing.assert_array_equal(np.asarray(ours["point_list_unsorted"]).astype(np.int64),
                                      np.asarray(ref["point_list_unsorted"]).astype(np.int64))


# ------------------------------------------------------------------------------------------------ fragile pixels
FRAGILE_STATS = {"cases": 0, "cases_with_fragile_pixels": 0, "fragile_pixels": 0, "gaussians_masked": 0, "gaussians": 0}


def fragile_gaussian_rows(ref: Dict, case: "Case") -> np.ndarray:
    """Bool [P]: Gaussians whose gradients the CPU oracle cannot vouch for — those in the list of a tile that contains
    a pixel the oracle flagged fragile (a discrete decision within rounding distance of its threshold, where two
    correct expf implementations may differ). Everything else is compared; the counts are accumulated in
    FRAGILE_STATS and printed by the tests instead of the whole gradient check being skipped."""
    P = case.P
    mask = np.zeros(P, dtype=bool)
    frag = (np.asarray(ref["fragile"]) != 0).reshape(case.height, case.width)
    FRAGILE_STATS["cases"] += 1
    FRAGILE_STATS["gaussians"] += P
    if not frag.any():
        return mask
    FRAGILE_STATS["cases_with_fragile_pixels"] += 1
    FRAGILE_STATS["fragile_pixels"] += int(frag.sum())
    grid_x = (case.width + 15) // 16
    ys, xs = np.nonzero(frag)
    tiles = np.unique((ys // 16) * grid_x + xs // 16)
    ranges, plist = np.asarray(ref["ranges"]).astype(np.int64), np.asarray(ref["point_list"]).astype(np.int64)
    for t in tiles:
        mask[plist[ranges[t, 0]:ranges[t, 1]]] = True
    FRAGILE_STATS["gaussians_masked"] += int(mask.sum())
    return mask
