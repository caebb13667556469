"""Sweep of the blend knobs on the bench scene: per configuration the captured fwd+bwd step averaged over the window's
eight poses (ms).  python tools/knob_sweep.py"""
import itertools, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gs_livm_b200  # noqa
from gs_livm_b200 import _lib, engine, scene
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
for c in cams:
    ctx.forward(sc, c); ctx.backward(dC, dS)
    if not ctx.check():
        ctx.forward(sc, c, redo=True); ctx.check()
def measure():
    frames = [ctx.capture_step(sc, c, dC, dS) for c in cams]
    for f in frames: f.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        for f in frames: f.replay()
    e1.record(); torch.cuda.synchronize(); ctx.check()
    return e0.elapsed_time(e1) / 24
base = measure()
print("defaults: %.4f ms" % base, flush=True)
for hf, hb, fi, bi in itertools.product((2048, 3072, 4096), (2048, 3072, 4096, 6144), (4,), (1, 2)):
    _lib.configure(heavy_fwd=hf, heavy_bwd=hb, fwd_ilp=fi, bwd_ilp=bi)
    print("heavy_fwd %5d heavy_bwd %5d fwd_ilp %d bwd_ilp %d: %.4f ms" % (hf, hb, fi, bi, measure()), flush=True)
_lib.configure(heavy_fwd=-1, heavy_bwd=-1, fwd_ilp=-1, bwd_ilp=-1)
for sp, bt in ((2047, 512), (2047, 640), (2047, 896), (2047, 1280), (1023, 640)):
    _lib.configure(split_len=sp, bucket_target=bt)
    print("split_len %d bucket_target %d: %.4f ms" % (sp, bt, measure()), flush=True)
