"""Developer diagnostic (not a test): stage-by-stage comparison of libgsr_b200 with Oracle A (reference CUDA)
and coarse timings of both. Run on a GPU box:  python tools/dev_compare.py [P ...]"""
from __future__ import annotations

import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import gs_livm_b200  # noqa: E402
from gs_livm_b200 import debug, rasterizer, scene  # noqa: E402
from oracle import ref  # noqa: E402


def cuda_time(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def compare(P, intr, seed, sh_degree=3, heavy=False, close_up=False, timing=True):
    dev = torch.device("cuda:0")
    sc = {k: v.to(dev) for k, v in scene.make_scene(P, seed, sh_degree, heavy_overlap=heavy).items()}
    cam = scene.scene_camera(sc, 0, intr, close_up=close_up).to(dev)
    W, H = cam.width, cam.height
    print(f"=== P={P} {W}x{H} deg={sh_degree} heavy={heavy} close={close_up}")

    def ours_fwd():
        return rasterizer.rasterize_gaussians_cuda(sc["background"], sc["means3D"], torch.empty(0, device=dev),
                                                   sc["opacities"], sc["scales"], sc["rotations"], 1.0,
                                                   torch.empty(0, device=dev), cam.viewmatrix, cam.projmatrix,
                                                   cam.tanfovx, cam.tanfovy, H, W, sc["shs"], sh_degree,
                                                   cam.camera_center, False, False)

    def ref_fwd(state=False):
        return ref.forward(sc["background"], sc["means3D"], sc["opacities"], sc["scales"], sc["rotations"], sc["shs"],
                           sh_degree, cam.viewmatrix, cam.projmatrix, cam.camera_center, cam.tanfovx, cam.tanfovy, H,
                           W, with_state=state)

    R, color, depth, acc, radii, gb, bb, ib = ours_fwd()
    torch.cuda.synchronize()
    r = ref_fwd(True)
    torch.cuda.synchronize()
    g, b, im = debug.geom_state(gb, P), debug.binning_state(bb, R), debug.image_state(ib, W, H)
    vis = r["radii"] > 0
    print(f"R ours={R} ref={r['num_rendered']}  visible ref={int(vis.sum())}")
    print("radii mismatches:", int((radii != r["radii"]).sum()))
    myvis = radii > 0
    both = vis & myvis
    print("tiles_touched mismatches:", int((g["tiles_touched"] != r["tiles_touched"]).sum()))
    print("depth bit mismatches:", int((g["depths"][both].view(torch.int32) != r["depths"][both].view(torch.int32)).sum()))
    print("means2D bit mismatches:", int((g["means2D"][both].view(torch.int32) != r["means2D"][both].view(torch.int32)).sum()))
    d = (g["conic_opacity"][both] - r["conic_opacity"][both]).abs()
    print("conic_opacity max abs diff:", float(d.max()) if d.numel() else 0.0,
          " bit mismatches:", int((g["conic_opacity"][both].view(torch.int32) != r["conic_opacity"][both].view(torch.int32)).sum()))
    d = (g["rgb"][both] - r["rgb"][both]).abs()
    print("rgb max abs diff:", float(d.max()) if d.numel() else 0.0)
    rc = r["clamped"][both].to(torch.uint8)
    mine = g["clamped"][both]
    mine3 = torch.stack([(mine >> c) & 1 for c in range(3)], dim=1)
    print("clamped mismatches:", int((mine3 != rc).sum()))
    if R == r["num_rendered"]:
        print("keys_unsorted mismatches:", int((b["point_list_keys_unsorted"] != r["point_list_keys_unsorted"]).sum()))
        print("keys sorted mismatches:", int((b["point_list_keys"] != r["point_list_keys"]).sum()))
        print("point_list mismatches:", int((b["point_list"] != r["point_list"]).sum()))
        print("ranges mismatches:", int((im["ranges"] != r["ranges"]).sum()))
    for name, mine_t, ref_t in (("C", color, r["out_color"]), ("D", depth, r["out_depth"]), ("S", acc, r["out_acc"])):
        dd = (mine_t - ref_t).abs()
        print(f"{name}: max abs diff {float(dd.max()):.3e}  bit mismatches {int((mine_t.view(torch.int32) != ref_t.view(torch.int32)).sum())}"
              f"  >1e-5: {int((dd > 1e-5).sum())}")
    print("final_T bit mismatches:", int((im["accum_alpha"].view(torch.int32) != r["accum_alpha"].view(torch.int32)).sum()))
    print("n_contrib mismatches:", int((im["n_contrib"] != r["n_contrib"]).sum()))

    # backward
    gen = torch.Generator(device="cpu").manual_seed(seed + 1)
    dC = (torch.randn((3, H, W), generator=gen) / (H * W)).to(dev)
    dS = (torch.randn((1, H, W), generator=gen) / (H * W)).to(dev)

    def ours_bwd():
        return rasterizer.rasterize_gaussians_backward_cuda(
            sc["background"], sc["means3D"], radii, torch.empty(0, device=dev), sc["scales"], sc["rotations"], 1.0,
            torch.empty(0, device=dev), cam.viewmatrix, cam.projmatrix, cam.tanfovx, cam.tanfovy, dC, dS, sc["shs"],
            sh_degree, cam.camera_center, gb, R, bb, ib, False)

    def ref_bwd():
        return ref.backward(r, sc["background"], sc["means3D"], sc["scales"], sc["rotations"], sc["shs"], sh_degree,
                            cam.viewmatrix, cam.projmatrix, cam.camera_center, cam.tanfovx, cam.tanfovy, dC, dS)

    mine_g = dict(zip(("dL_dmean2D", "dL_dcolor", "dL_dopacity", "dL_dmean3D", "dL_dcov3D", "dL_dsh", "dL_dscale",
                       "dL_drot"), ours_bwd()))
    ref_g = ref_bwd()
    ref_g2 = ref_bwd()
    torch.cuda.synchronize()
    for k in mine_g:
        a, bref, b2 = mine_g[k].double(), ref_g[k].double(), ref_g2[k].double()
        scale = float(bref.abs().max()) + 1e-30
        err = float((a - bref).abs().max()) / scale
        self_err = float((b2 - bref).abs().max()) / scale
        rel = ((a - bref).abs() / (bref.abs() + 1e-3 * scale)).max()
        print(f"{k}: max|ours-ref|/max|ref| = {err:.3e}   ref-vs-ref rerun = {self_err:.3e}   max elem-rel(+1e-3 floor) = {float(rel):.3e}"
              f"  finite={bool(torch.isfinite(a).all())}")

    if timing:
        t_of = cuda_time(ours_fwd)
        t_rf = cuda_time(ref_fwd)
        t_ob = cuda_time(ours_bwd)
        t_rb = cuda_time(ref_bwd)
        print(f"TIMING ms  ours fwd {t_of:.3f} bwd {t_ob:.3f} total {t_of + t_ob:.3f} | ref fwd {t_rf:.3f} bwd {t_rb:.3f} "
              f"total {t_rf + t_rb:.3f} | speedup {(t_rf + t_rb) / (t_of + t_ob):.2f}x")


if __name__ == "__main__":
    torch.cuda.init()
    print(torch.cuda.get_device_name(0))
    sizes = [int(a) for a in sys.argv[1:]] or [20000, 200000, 1000000]
    for P in sizes:
        compare(P, scene.NTU, 20241017 + 1)
    compare(50000, dict(width=376, height=240, fx=212.513, fy=213.399), 20241017 + 7, sh_degree=0)
    compare(200000, scene.NTU, 20241017 + 4, heavy=True, close_up=True)
