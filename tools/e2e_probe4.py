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
leaves = {k: sc[k].clone().requires_grad_(True) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
means2D = torch.zeros_like(leaves["means3D"], requires_grad=True)
target = torch.rand((3, H, W), device=dev)
host_target = torch.rand((3, H, W)).pin_memory()
loss_host = torch.zeros(1).pin_memory()
rs = rasterizer.GaussianRasterizationSettings(H, W, cam_cpu.tanfovx, cam_cpu.tanfovy, sc["background"], 1.0,
                                              cam.viewmatrix, cam.projmatrix, 3, cam.camera_center, False)
def timed(name, fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); print(f"{name:60s} {(time.perf_counter() - t0) / n * 1e3:8.3f} ms", flush=True)
def render():
    return rasterizer.GaussianRasterizer(rs)(leaves["means3D"], means2D, leaves["opacities"], shs=leaves["shs"], scales=leaves["scales"], rotations=leaves["rotations"])
def clear():
    for t in list(leaves.values()) + [means2D]: t.grad = None
def a_full():
    clear(); color, radii, depth, acc = render()
    loss = (color - target).abs().mean() + 0.1 * acc.mean(); loss.backward()
    loss_host.copy_(loss.detach().reshape(1), non_blocking=True); torch.cuda.current_stream().synchronize()
def b_nosync():
    clear(); color, radii, depth, acc = render()
    loss = (color - target).abs().mean() + 0.1 * acc.mean(); loss.backward()
def c_simple_loss():
    clear(); color, radii, depth, acc = render()
    loss = color.sum(); loss.backward()
    torch.cuda.current_stream().synchronize()
dC = torch.randn((3, H, W), device=dev) / (H * W); dS = torch.randn((1, H, W), device=dev) / (H * W)
def d_grad_direct():
    clear(); color, radii, depth, acc = render()
    torch.autograd.backward([color, acc], [dC, dS])
    torch.cuda.current_stream().synchronize()
ctx = engine.RasterContext(P, 16, W, H, dev, 3)
def e_ctx_sync():
    ctx.forward(sc, cam); ctx.backward(dC, dS); torch.cuda.current_stream().synchronize()
def f_ctx_nosync():
    ctx.forward(sc, cam); ctx.backward(dC, dS)
def g_raw():
    e = torch.empty(0, device=dev)
    R, color, depth, acc, radii, gb, bb, ib = rasterizer.rasterize_gaussians_cuda(sc["background"], sc["means3D"], e, sc["opacities"], sc["scales"], sc["rotations"], 1.0, e, cam.viewmatrix, cam.projmatrix, cam_cpu.tanfovx, cam_cpu.tanfovy, H, W, sc["shs"], 3, cam.camera_center, False, False)
    rasterizer.rasterize_gaussians_backward_cuda(sc["background"], sc["means3D"], radii, e, sc["scales"], sc["rotations"], 1.0, e, cam.viewmatrix, cam.projmatrix, cam_cpu.tanfovx, cam_cpu.tanfovy, dC, dS, sc["shs"], 3, cam.camera_center, gb, R, bb, ib, False)
    torch.cuda.current_stream().synchronize()
timed("A operator + L1 loss + autograd + D2H + sync per step", a_full)
timed("B same without per-step sync / D2H", b_nosync)
timed("C operator + color.sum() loss + sync", c_simple_loss)
timed("D operator + autograd.backward(given grads) + sync", d_grad_direct)
timed("G raw tensor-level fwd+bwd (no autograd) + sync", g_raw)
timed("E RasterContext fwd+bwd + sync per step", e_ctx_sync)
timed("F RasterContext fwd+bwd, no per-step sync", f_ctx_nosync)
