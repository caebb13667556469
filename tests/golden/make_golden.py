"""Generate the golden fixtures in this directory by running the UNMODIFIED reference rasterizer
(Oracle A = oracle/_ref/gsr_ref_ext.so, built by `make -C oracle ref` from /root/reference) on a B200.

The reference ships no golden vectors for this path, so these are outputs of the
reference itself on small seeded inputs. Inputs are stored next to the outputs, so the tests do not depend
on the scene generator staying bit-stable.

    gpurun -- python tests/golden/make_golden.py gpurun_out/golden     # then copy *.npz here

Gradients come from float atomics, so two runs of the reference differ in the last bits; the stored
gradients are one such run (tests compare at 1e-4 relative, see tests/tolerances.py).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import gs_livm_b200  # noqa: E402,F401
from gs_livm_b200 import scene  # noqa: E402
from oracle import ref  # noqa: E402

CASES = {
    # name: (P, seed, sh_degree, intrinsics, mode); scales x4 ("heavy overlap") so that lists span several
    # staging batches and pixels saturate (the T < 1e-4 termination is exercised)
    "ntu_deg3": (4000, 20241017 + 11, 3, dict(width=160, height=96, fx=90.43, fy=90.81), "sh"),
    "partial_tiles_deg0": (6000, 20241017 + 12, 0, dict(width=100, height=50, fx=56.5, fy=56.7), "sh"),
    "precomp_colors": (3000, 20241017 + 13, 1, dict(width=96, height=64, fx=54.2, fy=54.4), "colors"),
}


def run_case(name, P, seed, deg, intr, mode, dev):
    sc = scene.make_scene(P, seed, deg, heavy_overlap=True)
    cam = scene.scene_camera(sc, 1, intr)
    g = torch.Generator().manual_seed(seed)
    H, W = cam.height, cam.width
    dC = torch.randn((3, H, W), generator=g) / (H * W)
    dS = torch.randn((1, H, W), generator=g) / (H * W)
    colors = torch.rand((P, 3), generator=g) if mode == "colors" else None
    inputs = dict(means3D=sc["means3D"], scales=sc["scales"], rotations=sc["rotations"], opacities=sc["opacities"],
                  shs=sc["shs"], background=sc["background"], viewmatrix=cam.viewmatrix, projmatrix=cam.projmatrix,
                  camera_center=cam.camera_center, dL_dcolor=dC, dL_dacc=dS)
    if colors is not None:
        inputs["colors_precomp"] = colors
    d = {k: v.to(dev) for k, v in inputs.items()}
    shs = None if mode == "colors" else d["shs"]
    fwd = ref.forward(d["background"], d["means3D"], d["opacities"], d["scales"], d["rotations"], shs, deg,
                      d["viewmatrix"], d["projmatrix"], d["camera_center"], cam.tanfovx, cam.tanfovy, H, W,
                      colors_precomp=d.get("colors_precomp"))
    bwd = ref.backward(fwd, d["background"], d["means3D"], d["scales"], d["rotations"], shs, deg, d["viewmatrix"],
                       d["projmatrix"], d["camera_center"], cam.tanfovx, cam.tanfovy, d["dL_dcolor"], d["dL_dacc"],
                       colors_precomp=d.get("colors_precomp"))
    torch.cuda.synchronize()
    out = {"in_" + k: v.numpy() for k, v in inputs.items()}
    out["meta"] = np.array([P, deg, W, H, fwd["num_rendered"]], dtype=np.int64)
    out["tanfov"] = np.array([cam.tanfovx, cam.tanfovy], dtype=np.float64)
    keep = ("out_color", "out_depth", "out_acc", "radii", "depths", "means2D", "conic_opacity", "rgb", "clamped",
            "tiles_touched", "point_offsets", "point_list_keys", "point_list", "ranges", "accum_alpha", "n_contrib",
            "cov3D")
    for k in keep:
        out["ref_" + k] = fwd[k].cpu().numpy()
    for k, v in bwd.items():
        out["ref_" + k] = v.cpu().numpy()
    return out


def main():
    outdir = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(outdir, exist_ok=True)
    dev = torch.device("cuda:0")
    for name, (P, seed, deg, intr, mode) in CASES.items():
        out = run_case(name, P, seed, deg, intr, mode, dev)
        path = os.path.join(outdir, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, "R =", int(out["meta"][4]), "visible =", int((out["ref_radii"] > 0).sum()), "->", path,
              os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
