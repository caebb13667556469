"""Developer diagnostic: where the operator-level (e2e) step spends its time."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gs_livm_b200
from gs_livm_b200 import rasterizer, scene
dev = torch.device("cuda:0")
P = 1_000_000
sc_cpu = scene.make_scene(P, 20241017 + 2, 3)
sc = {k: v.to(dev) for k, v in sc_cpu.items()}
cam_cpu = scene.scene_camera(sc_cpu, 0, scene.NTU)
cam = cam_cpu.to(dev)
H, W = cam.height, cam.width
leaves = {k: sc[k].clone().requires_grad_(True) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
means2D = torch.zeros_like(leaves["means3D"], requires_grad=True)
target = torch.rand((3, H, W), device=dev)
rs = rasterizer.GaussianRasterizationSettings(H, W, cam_cpu.tanfovx, cam_cpu.tanfovy, sc["background"], 1.0,
                                              cam.viewmatrix, cam.projmatrix, 3, cam.camera_center, False)
def sync(): torch.cuda.synchronize()
def timed(name, fn, n=20):
    for _ in range(3): fn()
    sync(); t0 = time.perf_counter()
    for _ in range(n): fn()
    sync(); print(f"{name:40s} {(time.perf_counter() - t0) / n * 1e3:8.3f} ms")
def fwd_only():
    with torch.no_grad():
        rasterizer.GaussianRasterizer(rs)(leaves["means3D"], means2D, leaves["opacities"], shs=leaves["shs"], scales=leaves["scales"], rotations=leaves["rotations"])
def fwd_bwd():
    for t in list(leaves.values()) + [means2D]: t.grad = None
    color, radii, depth, acc = rasterizer.GaussianRasterizer(rs)(leaves["means3D"], means2D, leaves["opacities"], shs=leaves["shs"], scales=leaves["scales"], rotations=leaves["rotations"])
    loss = (color - target).abs().mean() + 0.1 * acc.mean()
    loss.backward()
def raw_fwd_bwd():
    R, color, depth, acc, radii, gb, bb, ib = rasterizer.rasterize_gaussians_cuda(sc["background"], sc["means3D"], torch.empty(0, device=dev), sc["opacities"], sc["scales"], sc["rotations"], 1.0, torch.empty(0, device=dev), cam.viewmatrix, cam.projmatrix, cam_cpu.tanfovx, cam_cpu.tanfovy, H, W, sc["shs"], 3, cam.camera_center, False, False)
    rasterizer.rasterize_gaussians_backward_cuda(sc["background"], sc["means3D"], radii, torch.empty(0, device=dev), sc["scales"], sc["rotations"], 1.0, torch.empty(0, device=dev), cam.viewmatrix, cam.projmatrix, cam_cpu.tanfovx, cam_cpu.tanfovy, target, acc, sc["shs"], 3, cam.camera_center, gb, R, bb, ib, False)
timed("operator forward (no grad)", fwd_only)
timed("raw fwd+bwd (tensor-level, no autograd)", raw_fwd_bwd)
timed("operator fwd + loss + autograd bwd", fwd_bwd)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3): fwd_bwd()
    sync()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
