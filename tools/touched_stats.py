"""How sparse are a view's gradients? Per pose of the bench window: Gaussians rendered (radii > 0) vs Gaussians the blend
backward actually touched (any of dL_dcolor / dL_dopacity / dL_dmean2D / dL_dconic non-zero), and the unions over the
window. Everything the per-Gaussian backward computes is linear in those four, so an untouched Gaussian has all-zero
parameter gradients and need not be exchanged."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gs_livm_b200  # noqa
from gs_livm_b200 import engine, scene
dev = torch.device("cuda:0")
P, SEED = 1_000_000, 20241017 + 2
sc_cpu = scene.make_scene(P, SEED, 3)
sc = {k: v.to(dev) for k, v in sc_cpu.items()}
cams = [scene.scene_camera(sc_cpu, i, scene.NTU).to(dev) for i in range(8)]
H, W = cams[0].height, cams[0].width
g = torch.Generator().manual_seed(SEED + 1)
dC = (torch.randn((3, H, W), generator=g) / (H * W)).to(dev)
dS = (torch.randn((1, H, W), generator=g) / (H * W)).to(dev)
ctx = engine.RasterContext(P, 16, W, H, dev, 3)
uv = torch.zeros(P, dtype=torch.bool, device=dev); ut = torch.zeros(P, dtype=torch.bool, device=dev)
for k, cam in enumerate(cams):
    ctx.forward(sc, cam); ctx.backward(dC, dS); torch.cuda.synchronize()
    vis = ctx.radii > 0
    gr = ctx.grads
    touched = (gr["dL_dcolor"].abs().sum(1) != 0) | (gr["dL_dopacity"].abs().sum(1) != 0) | (gr["dL_dmean2D"].abs().sum(1) != 0) | (gr["dL_dconic"].abs().sum(1) != 0)
    rgbnz = gr["dL_dcolor"].abs().sum(1) != 0
    param_nz = (gr["dL_dmean3D"].abs().sum(1) != 0) | (gr["dL_dsh"].abs().sum(1) != 0)
    # granularity of 4 consecutive Gaussians (what the exchange kernels skip by)
    t4 = touched.view(-1, 4).any(1).float().mean().item(); v4 = vis.view(-1, 4).any(1).float().mean().item()
    print(f"pose {k}: rendered {vis.float().mean().item():.3f} touched {touched.float().mean().item():.3f} dRGB!=0 {rgbnz.float().mean().item():.3f} "
          f"param grads != 0 {param_nz.float().mean().item():.3f} | groups of 4: rendered {v4:.3f} touched {t4:.3f}; touched & ~rendered {int((touched & ~vis).sum())}")
    uv |= vis; ut |= touched
print(f"union over 8 poses: rendered {uv.float().mean().item():.3f} touched {ut.float().mean().item():.3f}; groups of 4: rendered {uv.view(-1,4).any(1).float().mean().item():.3f} touched {ut.view(-1,4).any(1).float().mean().item():.3f}")
