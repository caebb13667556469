import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gs_livm_b200
from gs_livm_b200 import rasterizer, scene, engine
dev = torch.device("cuda:0")
P = 1_000_000
sc_cpu = scene.make_scene(P, 20241017 + 2, 3)
sc = {k: v.to(dev) for k, v in sc_cpu.items()}
cam_cpu = scene.scene_camera(sc_cpu, 0, scene.NTU)
cam = cam_cpu.to(dev)
H, W = cam.height, cam.width
e = torch.empty(0, device=dev)
def raw_fwd():
    return rasterizer.rasterize_gaussians_cuda(sc["background"], sc["means3D"], e, sc["opacities"], sc["scales"], sc["rotations"], 1.0, e, cam.viewmatrix, cam.projmatrix, cam_cpu.tanfovx, cam_cpu.tanfovy, H, W, sc["shs"], 3, cam.camera_center, False, False)
for mode in ("tile", "device", "tile"):
    if mode == "device": os.environ["GSR_SORT"] = "device"
    else: os.environ.pop("GSR_SORT", None)
    for _ in range(3): raw_fwd()
    torch.cuda.synchronize()
    s0 = torch.cuda.memory_stats()
    ts = []
    t0 = time.perf_counter()
    for _ in range(20):
        a = time.perf_counter(); raw_fwd(); ts.append((time.perf_counter() - a) * 1e3)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    s1 = torch.cuda.memory_stats()
    print(mode, "avg %.3f ms" % ((t1 - t0) / 20 * 1e3), "cudaMallocs", s1["num_device_alloc"] - s0["num_device_alloc"], "frees", s1["num_device_free"] - s0["num_device_free"],
          "retries", s1["num_alloc_retries"] - s0["num_alloc_retries"], "per-call ms", [round(x, 2) for x in ts], flush=True)
