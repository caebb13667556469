"""Developer diagnostic: per-stage device times of libgsr_b200 (CUDA events inside the C ABI) next to the
reference CUDA rasterizer's forward / backward totals, same scene.  python tools/stage_profile.py [P] [iters]"""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import gs_livm_b200  # noqa: E402
from gs_livm_b200 import _lib, engine, scene  # noqa: E402


def main():
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    which = sys.argv[3] if len(sys.argv) > 3 else "both"
    intr = scene.BOTANIC if (len(sys.argv) > 4 and sys.argv[4] == "botanic") else scene.NTU
    pose = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    heavy = len(sys.argv) > 6 and sys.argv[6] == "heavy"
    close = len(sys.argv) > 7 and sys.argv[7] == "close"
    dev = torch.device("cuda:0")
    seed = 20241017 + (4 if heavy else 2)
    sc = {k: v.to(dev) for k, v in scene.make_scene(P, seed, 3, heavy_overlap=heavy).items()}
    cam = scene.scene_camera(sc, pose, intr, close_up=close).to(dev)
    W, H = cam.width, cam.height
    gen = torch.Generator(device="cpu").manual_seed(7)
    dC = (torch.randn((3, H, W), generator=gen) / (H * W)).to(dev)
    dS = (torch.randn((1, H, W), generator=gen) / (H * W)).to(dev)

    if which in ("both", "ours"):
        ctx = engine.RasterContext(P, 16, W, H, dev, 3)
        ctx.stage_timing(True)
        acc = {n: 0.0 for n in _lib.STAGES}
        tot = 0.0
        for it in range(iters + 3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ctx.forward(sc, cam)
            ctx.backward(dC, dS)
            e1.record()
            torch.cuda.synchronize()
            if it >= 3:
                for n, v in ctx.stage_times().items():
                    acc[n] += max(v, 0.0)  # -1: the stage did not run on this path
                tot += e0.elapsed_time(e1)
        ctx.check()
        print(f"ours P={P} {W}x{H} pose={pose} heavy={heavy} close={close} R={getattr(ctx, 'last_rendered', ctx.num_rendered)} visible={int((ctx.radii > 0).sum())}")
        for n in _lib.STAGES:
            print(f"  {n:12s} {acc[n] / iters * 1000:9.1f} us")
        print(f"  sum stages   {sum(acc.values()) / iters * 1000:9.1f} us ; fwd+bwd wall (events) {tot / iters * 1000:9.1f} us")

    if which in ("both", "ref"):
        from oracle import ref
        e = torch.empty(0, device=dev)

        def rf():
            return ref.forward(sc["background"], sc["means3D"], sc["opacities"], sc["scales"], sc["rotations"],
                               sc["shs"], 3, cam.viewmatrix, cam.projmatrix, cam.camera_center, cam.tanfovx,
                               cam.tanfovy, H, W, with_state=False)

        tf = tb = 0.0
        for it in range(iters + 3):
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record()
            r = rf()
            e1.record()
            ref.backward(r, sc["background"], sc["means3D"], sc["scales"], sc["rotations"], sc["shs"], 3,
                         cam.viewmatrix, cam.projmatrix, cam.camera_center, cam.tanfovx, cam.tanfovy, dC, dS)
            e2.record()
            torch.cuda.synchronize()
            if it >= 3:
                tf += e0.elapsed_time(e1)
                tb += e1.elapsed_time(e2)
        print(f"reference CUDA: fwd {tf / iters * 1000:.1f} us  bwd {tb / iters * 1000:.1f} us  total {(tf + tb) / iters * 1000:.1f} us")


if __name__ == "__main__":
    main()
