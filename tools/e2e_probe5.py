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
cam_cpu = scene.scene_camera(sc_cpu, 0, scene.NTU); cam = cam_cpu.to(dev)
H, W = cam.height, cam.width
leaves = {k: sc[k].clone().requires_grad_(True) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
means2D = torch.zeros_like(leaves["means3D"], requires_grad=True)
target = torch.rand((3, H, W), device=dev); loss_host = torch.zeros(1).pin_memory()
rs = rasterizer.GaussianRasterizationSettings(H, W, cam_cpu.tanfovx, cam_cpu.tanfovy, sc["background"], 1.0, cam.viewmatrix, cam.projmatrix, 3, cam.camera_center, False)
arena = rasterizer.RenderArena(dev)
acc_t = [0.0] * 6
def step(record):
    t0 = time.perf_counter()
    for t in list(leaves.values()) + [means2D]: t.grad = None
    t1 = time.perf_counter()
    color, radii, depth, acc = rasterizer.GaussianRasterizer(rs, arena=arena)(leaves["means3D"], means2D, leaves["opacities"], shs=leaves["shs"], scales=leaves["scales"], rotations=leaves["rotations"])
    t2 = time.perf_counter()
    loss = (color - target).abs().mean() + 0.1 * acc.mean()
    t3 = time.perf_counter()
    loss.backward()
    t4 = time.perf_counter()
    loss_host.copy_(loss.detach().reshape(1), non_blocking=True); torch.cuda.current_stream().synchronize()
    t5 = time.perf_counter()
    if record:
        for i, (a, b) in enumerate(((t0, t1), (t1, t2), (t2, t3), (t3, t4), (t4, t5), (t0, t5))): acc_t[i] += b - a
for _ in range(8): step(False)
N = 30
for _ in range(N): step(True)
names = ["clear grads", "forward call (CPU)", "loss ops (CPU)", "backward call (CPU)", "D2H + sync wait", "total"]
for n, v in zip(names, acc_t): print(f"{n:24s} {v / N * 1e6:8.1f} us")
