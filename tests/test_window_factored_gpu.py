"""The factored exchange of the SH gradients (csrc/window_factored.cu, window.FactoredExchange) on ONE GPU: the rows it
rebuilds from the pushed dRGB must be the rows geom_bwd writes — bit for bit (up to the sign of a zero) with one view
(same basis, same product), and the sum of the two views' rows with two, to within the rounding of one FMA per term. The multi-rank side
(peer stores / multicast, barriers, the 11-float all-reduce) is covered by tests/test_window_nccl_gpu.py on >= 2 GPUs
and by bench.py's exchange_max_err."""
import numpy as np
import pytest
import torch

import helpers

pytestmark = pytest.mark.gpu


def _setup(P, dev, n_views, seed=20241017 + 5, deg=3, chunks=4, track=False):
    import gs_livm_b200  # noqa: F401
    from gs_livm_b200 import engine, scene, window
    sc_cpu = scene.make_scene(P, seed, deg)
    sc = {k: v.to(dev) for k, v in sc_cpu.items()}
    cams = [scene.scene_camera(sc_cpu, i, scene.NTU).to(dev) for i in (0, 3)]
    H, W = cams[0].height, cams[0].width
    ex = window.FactoredExchange(P, 16, dev, n_views=n_views, n_chunks=chunks, track_zero_rows=track)
    ctx = engine.RasterContext(P, 16, W, H, dev, deg, grad_flat=ex.flat)
    g = torch.Generator().manual_seed(seed + 1)
    dC = (torch.randn((3, H, W), generator=g) / (H * W)).to(dev)
    dS = (torch.randn((1, H, W), generator=g) / (H * W)).to(dev)
    return sc, cams, ex, ctx, dC, dS


@pytest.mark.parametrize("P,chunks", [(50_000, 4), (49_999, 3), (130, 4), (50_000, 1), (1_000, 32)])
def test_one_view_rows_equal_geom_bwd_bit_for_bit(P, chunks, built_library, cuda_device):
    """chunks: the exchange pipelines push / per-Gaussian backward / expansion over that many ranges of Gaussians."""
    dev = cuda_device
    sc, cams, ex, ctx, dC, dS = _setup(P, dev, 1, chunks=chunks, track=(P == 50_000))
    ex.cam_centres[0].copy_(cams[0].camera_center)
    ctx.grad_flat.fill_(float("nan"))
    ctx.invalidate_grads()
    ex.mark_rows_dirty()
    ex.render_step(ctx, sc, cams[0], dC, dS, parity=0)
    torch.cuda.synchronize()
    got = {n: ctx.grads[n].clone() for n in ("dL_dsh", "dL_dmean3D", "dL_dscale", "dL_drot", "dL_dopacity", "dL_dcov3D")}
    assert int((ctx.radii > 0).sum()) > P // 10 or P < 1000
    # the dense per-Gaussian half on the SAME accumulated inputs (no second blend backward: float atomics differ per run)
    ctx.backward(dC, dS, part="geom")
    torch.cuda.synchronize()
    for n, t in got.items():
        if n not in ("dL_dsh", "dL_dopacity"):  # the two template instances of geom_bwd (with / without the SH rows) are
            # compiled separately: the covariance and direction-gradient sums contract differently -- last-place
            # differences of the terms (visible in elements that are tiny through cancellation), not bits
            scale = float(ctx.grads[n].abs().max())
            assert float((t - ctx.grads[n]).abs().max()) <= 1e-5 * scale, n
            continue
        np.testing.assert_array_equal(t.cpu().numpy(), ctx.grads[n].cpu().numpy(), err_msg=n)  # (-0 == +0)
    assert float(ctx.grads["dL_dsh"].abs().max()) > 0
    # ... and the whole thing replayed from a CUDA graph, on the other parity
    frame = ctx.capture(lambda: ex.render_step(ctx, sc, cams[0], dC, dS, parity=1))
    ctx.grad_flat.fill_(float("nan"))
    ctx.invalidate_grads()
    ex.mark_rows_dirty()
    frame.replay()
    torch.cuda.synchronize()
    assert ctx.check()
    rows = ctx.grads["dL_dsh"].clone()
    ctx.backward(dC, dS, part="geom")
    torch.cuda.synchronize()
    np.testing.assert_array_equal(rows.cpu().numpy(), ctx.grads["dL_dsh"].cpu().numpy())


def test_two_views_rows_are_the_sum_of_the_views_rows(built_library, cuda_device):
    dev = cuda_device
    P = 60_000
    sc, cams, ex, ctx, dC, dS = _setup(P, dev, 2)
    dense, vis = [], []
    for v, cam in enumerate(cams):
        ex.cam_centres[v].copy_(cam.camera_center)
        ctx.forward(sc, cam)
        ctx.backward(dC, dS, part="blend")
        ex.push(ctx, parity=0, view=v)
        ctx.backward(dC, dS, part="geom")  # dense rows of this view, from the same accumulators
        torch.cuda.synchronize()
        dense.append(ctx.grads["dL_dsh"].clone())
        vis.append((ctx.radii > 0).clone())
    assert int((vis[0] & vis[1]).sum()) > 1000 and int((vis[0] ^ vis[1]).sum()) > 1000, "views must overlap partly"
    ctx.grads["dL_dsh"].fill_(float("nan"))
    ex.mark_rows_dirty()
    ex.expand(ctx, sc["means3D"], parity=0)
    torch.cuda.synchronize()
    # one FMA per term: fma(b1, c1, round(b0 c0)) against round(round(b0 c0) + round(b1 c1)) -- half an ulp of a term
    want, bound = dense[0] + dense[1], 1.2e-7 * (dense[0].abs() + dense[1].abs())
    assert bool(((ctx.grads["dL_dsh"] - want).abs() <= bound).all())
    only0 = vis[0] & ~vis[1]  # rendered by one view only: still exactly that view's rows
    np.testing.assert_array_equal(ctx.grads["dL_dsh"][only0].cpu().numpy(), dense[0][only0].cpu().numpy())
    none = ~(vis[0] | vis[1])
    assert int(none.sum()) > 0 and float(ctx.grads["dL_dsh"][none].abs().max()) == 0.0


def test_operator_arena_with_an_exchange_returns_the_same_gradients(built_library, cuda_device):
    """GaussianRasterizer(settings, arena=RenderArena(exchange=FactoredExchange)): the operator's backward runs in halves
    around the factored exchange and hands out its gradients from the exchange's buffer; on one rank they must be the
    gradients of the plain operator."""
    import gs_livm_b200  # noqa: F401
    from gs_livm_b200 import rasterizer, scene, window
    from tolerances import assert_grad
    dev = cuda_device
    P = 40_000
    sc_cpu = scene.make_scene(P, 20241017 + 9, 3)
    cam_cpu = scene.scene_camera(sc_cpu, 1, scene.NTU)
    cam = cam_cpu.to(dev)
    H, W = cam.height, cam.width
    g = torch.Generator().manual_seed(3)
    wC = (torch.randn((3, H, W), generator=g) / (H * W)).to(dev)
    wS = (torch.randn((1, H, W), generator=g) / (H * W)).to(dev)
    rs = scene.settings_for(cam, sc_cpu["background"].to(dev), 3)

    def run(arena):
        leaves = {k: sc_cpu[k].to(dev).clone().requires_grad_(True) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
        m2 = torch.zeros_like(leaves["means3D"], requires_grad=True)
        out = None
        for _ in range(2):  # the second render goes through the sync-free path when there is an arena
            for t in leaves.values():
                t.grad = None
            color, radii, depth, acc = rasterizer.GaussianRasterizer(rs, arena=arena)(
                leaves["means3D"], m2, leaves["opacities"], shs=leaves["shs"], scales=leaves["scales"],
                rotations=leaves["rotations"])
            ((color * wC).sum() + (acc * wS).sum()).backward()
            torch.cuda.synchronize()
            if arena is not None:
                assert arena.check()
            out = {k: v.grad.detach().cpu().numpy().copy() for k, v in leaves.items()}
        return out

    want = run(None)
    ex = window.FactoredExchange(P, 16, dev)
    ex.cam_centres[0].copy_(cam.camera_center)
    got = run(rasterizer.RenderArena(dev, exchange=ex))
    for k in want:
        assert_grad(k, got[k], want[k])
