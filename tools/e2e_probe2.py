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
def sync(): torch.cuda.synchronize()
def timed(name, fn, n=20):
    for _ in range(3): fn()
    sync(); t0 = time.perf_counter()
    for _ in range(n): fn()
    sync(); print(f"{name:50s} {(time.perf_counter() - t0) / n * 1e3:8.3f} ms", flush=True)
def raw_fwd():
    return rasterizer.rasterize_gaussians_cuda(sc["background"], sc["means3D"], e, sc["opacities"], sc["scales"], sc["rotations"], 1.0, e, cam.viewmatrix, cam.projmatrix, cam_cpu.tanfovx, cam_cpu.tanfovy, H, W, sc["shs"], 3, cam.camera_center, False, False)
ctx = engine.RasterContext(P, 16, W, H, dev, 3)
timed("RasterContext.forward", lambda: ctx.forward(sc, cam))
timed("rasterize_gaussians_cuda (tile path)", raw_fwd)
os.environ["GSR_SORT"] = "device"
timed("rasterize_gaussians_cuda (device sort path)", raw_fwd)
timed("RasterContext.forward (device sort path)", lambda: ctx.forward(sc, cam))
os.environ.pop("GSR_SORT")
def allocs():
    a = torch.zeros((3, H, W), device=dev); b = torch.zeros((P,), dtype=torch.int32, device=dev)
    c = torch.empty(80_000_000, dtype=torch.uint8, device=dev); d = torch.empty(160_000_000, dtype=torch.uint8, device=dev)
timed("allocations only", allocs)
import ctypes as C
from gs_livm_b200 import _lib
lib = _lib.load()
lib.gsr_stage_timing_enable(1)
for it in range(3):
    sync(); t0 = time.perf_counter()
    out = raw_fwd()
    sync(); t1 = time.perf_counter()
    arr = (C.c_float * len(_lib.STAGES))()
    lib.gsr_stage_times(arr, len(_lib.STAGES))
    print("wall %.3f ms" % ((t1 - t0) * 1e3), {n: round(float(arr[i]) * 1e3, 1) for i, n in enumerate(_lib.STAGES)}, flush=True)
lib.gsr_stage_timing_enable(0)
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(5): raw_fwd()
sync(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(12)
