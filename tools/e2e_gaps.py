"""Developer diagnostic: GPU timeline of the operator-level step (kernel start/duration, idle gaps)."""
import os, sys, json, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gs_livm_b200
from gs_livm_b200 import rasterizer, scene
from torch.profiler import profile, ProfilerActivity
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
def step():
    for t in list(leaves.values()) + [means2D]: t.grad = None
    color, radii, depth, acc = rasterizer.GaussianRasterizer(rs)(leaves["means3D"], means2D, leaves["opacities"], shs=leaves["shs"], scales=leaves["scales"], rotations=leaves["rotations"])
    loss = (color - target).abs().mean() + 0.1 * acc.mean(); loss.backward()
    loss_host.copy_(loss.detach().reshape(1), non_blocking=True); torch.cuda.current_stream().synchronize()
for _ in range(5): step()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3): step()
prof.export_chrome_trace("/tmp/trace.json")
ev = json.load(open("/tmp/trace.json"))["traceEvents"]
k = sorted([e for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")], key=lambda e: e["ts"])
t0 = k[0]["ts"]; prev_end = None
# print the middle step
for e in k:
    gap = (e["ts"] - prev_end) if prev_end is not None else 0
    print("%9.1f  dur %7.1f  gap %7.1f  %s" % (e["ts"] - t0, e["dur"], gap, e["name"][:70]))
    prev_end = e["ts"] + e["dur"]
