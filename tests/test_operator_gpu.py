"""Operator-level tests on the GPU: the Python mirror of GaussianRasterizer/_RasterizeGaussians under autograd,
markVisible, the allocation-free RasterContext, and — when built — the reference's UNCHANGED operator files
(src/gs/rasterize_points.cu + rasterizer.cu) linked against our library (oracle/_ref/gsr_dropin_ext.so)."""
import importlib.util
import os

import numpy as np
import pytest
import torch

import helpers
from tolerances import assert_grad

pytestmark = pytest.mark.gpu
SMALL = dict(width=160, height=96, fx=90.43, fy=90.81)
DROPIN = os.path.join(helpers.ROOT, "oracle", "_ref", "gsr_dropin_ext.so")


def _leaf(t, dev):
    return t.to(dev).clone().requires_grad_(True)


def test_autograd_through_gaussian_rasterizer(built_library, cuda_device):
    """loss.backward() through GaussianRasterizer.forward gives the same gradients as the raw backward entry point;
    a loss that only uses D contributes nothing (the reference drops dL/dD, rasterizer.cu:77-80)."""
    import gs_livm_b200  # noqa: F401
    from gs_livm_b200 import rasterizer, scene
    case = helpers.scene_case(4000, 501, 3, intr=SMALL, heavy=True)
    dev = cuda_device
    leaves = {k: _leaf(getattr(case, k), dev) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
    means2D = torch.zeros_like(leaves["means3D"], requires_grad=True)
    cam = scene.Camera(case.width, case.height, case.tanfovx, case.tanfovy, case.viewmatrix, case.projmatrix,
                       case.camera_center)
    rs = scene.settings_for(cam, case.background, case.sh_degree)
    rast = rasterizer.GaussianRasterizer(rs)
    color, radii, depth, acc = rast(leaves["means3D"], means2D, leaves["opacities"], shs=leaves["shs"],
                                    scales=leaves["scales"], rotations=leaves["rotations"])
    assert color.shape == (3, case.height, case.width) and depth.shape == (1, case.height, case.width)
    assert radii.dtype == torch.int32 and not radii.requires_grad
    wC, wS = case.dL_dcolor.to(dev), case.dL_dacc.to(dev)
    loss = (color * wC).sum() + (acc * wS).sum() + 123.0 * depth.sum()
    loss.backward()
    raw = helpers.run_ours(case, dev)
    for leaf, name in (("means3D", "dL_dmean3D"), ("shs", "dL_dsh"), ("opacities", "dL_dopacity"),
                       ("scales", "dL_dscale"), ("rotations", "dL_drot")):
        assert_grad(name, leaves[leaf].grad.cpu().numpy().reshape(raw[name].shape), raw[name])
    assert_grad("dL_dmean2D", means2D.grad.cpu().numpy(), raw["dL_dmean2D"])


def test_render_arena_gives_the_same_gradients_without_waiting(built_library, cuda_device):
    """GaussianRasterizer(settings, arena=RenderArena): first render synchronous (learns the instance count), the
    following ones through gsr_forward_presized; images bit-identical, gradients as without an arena."""
    import gs_livm_b200  # noqa: F401
    from gs_livm_b200 import rasterizer, scene
    case = helpers.scene_case(6000, 502, 3, intr=SMALL, heavy=True)
    dev = cuda_device
    cam = scene.Camera(case.width, case.height, case.tanfovx, case.tanfovy, case.viewmatrix, case.projmatrix,
                       case.camera_center)
    rs = scene.settings_for(cam, case.background, case.sh_degree)
    raw = helpers.run_ours(case, dev)
    arena = rasterizer.RenderArena(dev)
    wC, wS = case.dL_dcolor.to(dev), case.dL_dacc.to(dev)
    for it in range(3):
        leaves = {k: _leaf(getattr(case, k), dev) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
        means2D = torch.zeros_like(leaves["means3D"], requires_grad=True)
        color, radii, depth, acc = rasterizer.GaussianRasterizer(rs, arena=arena)(
            leaves["means3D"], means2D, leaves["opacities"], shs=leaves["shs"], scales=leaves["scales"],
            rotations=leaves["rotations"])
        ((color * wC).sum() + (acc * wS).sum()).backward()
        torch.cuda.synchronize()
        assert arena.check() and (it == 0 or arena.capacity > raw["num_rendered"])
        np.testing.assert_array_equal(helpers.bits(color.detach().cpu().numpy()), helpers.bits(raw["out_color"]))
        np.testing.assert_array_equal(radii.cpu().numpy(), raw["radii"])
        for leaf, name in (("means3D", "dL_dmean3D"), ("shs", "dL_dsh"), ("opacities", "dL_dopacity"),
                           ("scales", "dL_dscale"), ("rotations", "dL_drot")):
            assert_grad(name, leaves[leaf].grad.cpu().numpy().reshape(raw[name].shape), raw[name])


def test_captured_step_replays_the_operator_step(built_library, cuda_device):
    """rasterizer.CapturedStep: the step (camera H2D from a pinned staging tensor, render through the arena, loss,
    backward, loss D2H) captured as one CUDA graph; replays with a NEW camera written into the staging tensor give
    the images and gradients of eager renders with that camera."""
    import gs_livm_b200  # noqa: F401
    from gs_livm_b200 import rasterizer, scene
    case = helpers.scene_case(6000, 503, 3, intr=SMALL, heavy=True)
    dev = cuda_device
    leaves = {k: _leaf(getattr(case, k), dev) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
    means2D = torch.zeros_like(leaves["means3D"], requires_grad=True)
    wC, wS = case.dL_dcolor.to(dev), case.dL_dacc.to(dev)
    bg = case.background.to(dev)
    cams = [torch.cat([case.viewmatrix.reshape(-1), case.projmatrix.reshape(-1), case.camera_center.reshape(-1)])]
    shifted = case.viewmatrix.clone()
    shifted[3, 0] += 0.05  # the reference's matrices are stored transposed: row 3 is the translation
    cams.append(torch.cat([shifted.reshape(-1), (shifted @ (torch.linalg.inv(case.viewmatrix) @ case.projmatrix)).reshape(-1),
                           case.camera_center.reshape(-1)]))
    stage = torch.zeros(35).pin_memory()
    cam_dev = torch.zeros(35, device=dev)
    loss_host = torch.zeros(1).pin_memory()
    arena = rasterizer.RenderArena(dev)
    out = {}

    def settings():
        return rasterizer.GaussianRasterizationSettings(case.height, case.width, case.tanfovx, case.tanfovy, bg, 1.0,
                                                        cam_dev[:16].view(4, 4), cam_dev[16:32].view(4, 4),
                                                        case.sh_degree, cam_dev[32:35], False)

    def step():
        cam_dev.copy_(stage, non_blocking=True)
        for t in list(leaves.values()) + [means2D]:
            t.grad = None
        color, radii, depth, acc = rasterizer.GaussianRasterizer(settings(), arena=arena)(
            leaves["means3D"], means2D, leaves["opacities"], shs=leaves["shs"], scales=leaves["scales"],
            rotations=leaves["rotations"])
        loss = (color * wC).sum() + (acc * wS).sum()
        loss.backward()
        loss_host.copy_(loss.detach().reshape(1), non_blocking=True)
        out["color"] = color.detach()  # (nothing that keeps the autograd graph alive may outlive the step)

    stage.copy_(cams[0])
    with torch.no_grad():  # eager: the arena learns its capacity
        cam_dev.copy_(stage)
        rasterizer.GaussianRasterizer(settings(), arena=arena)(
            leaves["means3D"], means2D, leaves["opacities"], shs=leaves["shs"], scales=leaves["scales"],
            rotations=leaves["rotations"])
    torch.cuda.synchronize()
    assert arena.check()
    captured = rasterizer.CapturedStep(step, arena)
    for which in (1, 0, 1):
        stage.copy_(cams[which])
        captured.replay()
        torch.cuda.synchronize()
        assert arena.check()
        got_color = out["color"].detach().cpu().numpy().copy()
        got_loss = float(loss_host[0])
        got = {k: v.grad.detach().cpu().numpy().copy() for k, v in leaves.items()}
        # the same step, eagerly, without an arena
        cam_dev.copy_(stage)
        fresh = {k: v.detach().clone().requires_grad_(True) for k, v in leaves.items()}
        m2 = torch.zeros_like(fresh["means3D"], requires_grad=True)
        color, radii, depth, acc = rasterizer.GaussianRasterizer(settings())(
            fresh["means3D"], m2, fresh["opacities"], shs=fresh["shs"], scales=fresh["scales"],
            rotations=fresh["rotations"])
        loss = (color * wC).sum() + (acc * wS).sum()
        loss.backward()
        np.testing.assert_array_equal(helpers.bits(got_color), helpers.bits(color.detach().cpu().numpy()))
        ref_loss = float(loss.detach())
        assert abs(got_loss - ref_loss) <= 1e-5 * max(1.0, abs(ref_loss))
        for k in leaves:
            assert_grad(k, got[k], fresh[k].grad.cpu().numpy())
    # the two cameras really differ
    assert not np.array_equal(cams[0].numpy(), cams[1].numpy())


def test_mark_visible_matches_oracle(built_library, cuda_device):
    import gs_livm_b200  # noqa: F401
    from gs_livm_b200 import rasterizer
    from oracle import cpu
    case = helpers.scene_case(50_000, 502, 0)
    got = rasterizer.mark_visible(case.means3D.to(cuda_device), case.viewmatrix.to(cuda_device),
                                  case.projmatrix.to(cuda_device)).cpu().numpy()
    want = cpu.mark_visible(case.means3D, case.viewmatrix)
    np.testing.assert_array_equal(got, want)
    assert 0 < got.sum() < got.size


def test_raster_context_matches_operator_and_reuses_memory(built_library, cuda_device):
    import gs_livm_b200  # noqa: F401
    from gs_livm_b200 import engine, scene
    case = helpers.scene_case(20_000, 503, 3, intr=SMALL, heavy=True)
    dev = cuda_device
    sc = dict(means3D=case.means3D.to(dev), scales=case.scales.to(dev), rotations=case.rotations.to(dev),
              opacities=case.opacities.to(dev), shs=case.shs.to(dev), background=case.background.to(dev))
    cam = scene.Camera(case.width, case.height, case.tanfovx, case.tanfovy, case.viewmatrix, case.projmatrix,
                       case.camera_center).to(dev)
    ctx = engine.RasterContext(case.P, 16, case.width, case.height, dev, 3)
    raw = helpers.run_ours(case, dev)
    ptrs = None
    for it in range(4):  # the first forward is synchronous and sizes the binning arena; the others never wait
        color, depth, acc = ctx.forward(sc, cam)
        g = ctx.backward(case.dL_dcolor.to(dev), case.dL_dacc.to(dev))
        torch.cuda.synchronize()
        assert ctx.check(), "the frame must fit the presized binning arena"
        now = (ctx.geom.buf.data_ptr(), ctx.binning.buf.data_ptr(), ctx.image.buf.data_ptr())
        assert it < 2 or ptrs == now, "scratch arenas must be reused, not re-allocated"
        ptrs = now
        np.testing.assert_array_equal(helpers.bits(color.cpu().numpy()), helpers.bits(raw["out_color"]))
        np.testing.assert_array_equal(helpers.bits(depth.cpu().numpy()), helpers.bits(raw["out_depth"]))
        for n in helpers.GRAD_NAMES:
            assert_grad(n, g[n].cpu().numpy().reshape(raw[n].shape), raw[n])
    flat = ctx.param_grads()
    assert flat.numel() == 59 * case.P and flat.data_ptr() == ctx.grads["dL_dsh"].data_ptr()
    ctx.stage_timing(True)
    ctx.forward(sc, cam)
    ctx.backward(case.dL_dcolor.to(dev), case.dL_dacc.to(dev))
    times = ctx.stage_times()
    ctx.stage_timing(False)
    assert all(times[k] > 0 for k in ("preprocess", "sort", "blend_fwd", "blend_bwd", "geom_bwd"))
    # the same forward + backward as ONE captured graph (the gradient memsets run beside the forward): replays give
    # the eager results, also after the upstream gradients changed in place
    dC, dS = case.dL_dcolor.to(dev).clone(), case.dL_dacc.to(dev).clone()
    frame = ctx.capture_step(sc, cam, dC, dS)
    for scale in (1.0, -0.5):
        dC.copy_(case.dL_dcolor.to(dev) * scale)
        dS.copy_(case.dL_dacc.to(dev) * scale)
        ctx.grad_flat.fill_(float("nan"))  # every element must be rewritten by the replay
        ctx.invalidate_grads()
        frame.replay()
        torch.cuda.synchronize()
        assert ctx.check()
        np.testing.assert_array_equal(helpers.bits(ctx.out_color.cpu().numpy()), helpers.bits(raw["out_color"]))
        for n in helpers.GRAD_NAMES:
            assert_grad(n, ctx.grads[n].cpu().numpy().reshape(raw[n].shape), raw[n] * scale)


def test_sync_free_forward_reports_a_frame_that_does_not_fit(built_library, cuda_device):
    """gsr_forward_presized with a binning capacity below R: nothing is written out of bounds, the outputs are
    those of an empty scene, the status says so, and redoing the frame synchronously gives the right image."""
    import gs_livm_b200  # noqa: F401
    from gs_livm_b200 import engine, scene
    case = helpers.scene_case(20_000, 504, 3, intr=SMALL, heavy=True)
    dev = cuda_device
    sc = dict(means3D=case.means3D.to(dev), scales=case.scales.to(dev), rotations=case.rotations.to(dev),
              opacities=case.opacities.to(dev), shs=case.shs.to(dev), background=case.background.to(dev))
    cam = scene.Camera(case.width, case.height, case.tanfovx, case.tanfovy, case.viewmatrix, case.projmatrix,
                       case.camera_center).to(dev)
    raw = helpers.run_ours(case, dev, backward=False)
    ctx = engine.RasterContext(case.P, 16, case.width, case.height, dev, 3)
    ctx.forward(sc, cam)
    assert ctx.num_rendered == raw["num_rendered"] and ctx._capacity > ctx.num_rendered
    ctx._capacity = raw["num_rendered"] // 2  # pretend the arena was sized for a much emptier view
    color, _, acc = ctx.forward(sc, cam)
    g = ctx.backward(case.dL_dcolor.to(dev), case.dL_dacc.to(dev))
    assert not ctx.check()
    assert float(acc.abs().max()) == 0.0 and all(float(v.abs().max()) == 0.0 for v in g.values())
    color, depth, acc = ctx.forward(sc, cam, redo=True)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(helpers.bits(color.cpu().numpy()), helpers.bits(raw["out_color"]))
    color, depth, acc = ctx.forward(sc, cam)  # and the arena has been re-learned
    assert ctx.check()
    np.testing.assert_array_equal(helpers.bits(color.cpu().numpy()), helpers.bits(raw["out_color"]))


@pytest.mark.skipif(not os.path.exists(DROPIN), reason="oracle/_ref/gsr_dropin_ext.so not built (make -C oracle dropin)")
def test_unchanged_reference_operator_links_and_runs_on_our_library(built_library, cuda_device):
    """The reference's own GaussianRasterizer (unchanged C++), compiled against include/cuda_rasterizer/*.h and linked
    to libgsr_b200.so, produces the same bits as the Python mirror and back-propagates through LibTorch autograd."""
    spec = importlib.util.spec_from_file_location("gsr_dropin_ext", DROPIN)
    ext = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ext)
    case = helpers.scene_case(6000, 504, 3, intr=SMALL, heavy=True)
    dev = cuda_device
    d = lambda t: t.to(dev)  # noqa: E731
    out = ext.render_and_backward(d(case.means3D), d(case.opacities), d(case.shs), d(case.scales), d(case.rotations),
                                  d(case.background), d(case.viewmatrix), d(case.projmatrix), d(case.camera_center),
                                  case.tanfovx, case.tanfovy, case.height, case.width, case.sh_degree,
                                  d(case.dL_dcolor), d(case.dL_dacc))
    color, radii, depth, acc, g_means, g_sh, g_op, g_sc, g_rot = [t.cpu().numpy() for t in out]
    raw = helpers.run_ours(case, dev)
    np.testing.assert_array_equal(helpers.bits(color), helpers.bits(raw["out_color"]))
    np.testing.assert_array_equal(helpers.bits(depth), helpers.bits(raw["out_depth"]))
    np.testing.assert_array_equal(helpers.bits(acc), helpers.bits(raw["out_acc"]))
    np.testing.assert_array_equal(radii, raw["radii"])
    for got, name in ((g_means, "dL_dmean3D"), (g_sh, "dL_dsh"), (g_op, "dL_dopacity"), (g_sc, "dL_dscale"),
                      (g_rot, "dL_drot")):
        assert_grad(name, got.reshape(raw[name].shape), raw[name])


def test_overflowing_frame_from_a_cold_state_is_an_empty_render(built_library, cuda_device):
    """gsr_forward_presized as the FIRST call on fresh (garbage-filled) buffers, with a capacity the frame does not
    fit: the launch order must still be a permutation of the tiles and nothing may be read through stale tile ids --
    the image is exactly the background, every gradient zero, and a different camera afterwards behaves the same."""
    import ctypes as C
    import gs_livm_b200  # noqa: F401
    from gs_livm_b200 import _lib, debug, scene
    case = helpers.scene_case(20_000, 505, 3, intr=SMALL, heavy=True)
    dev = cuda_device
    lib = _lib.load()
    d = lambda t: t.to(dev).contiguous()  # noqa: E731
    t = dict(bg=d(case.background), m=d(case.means3D), sh=d(case.shs), op=d(case.opacities), sc=d(case.scales),
             rot=d(case.rotations))
    cams = [scene.Camera(case.width, case.height, case.tanfovx, case.tanfovy, case.viewmatrix, case.projmatrix,
                         case.camera_center).to(dev)]
    sc_cpu = scene.make_scene(20_000, 505, 3, heavy_overlap=True)
    cams.append(scene.scene_camera(sc_cpu, 5, SMALL).to(dev))
    W, H, P = case.width, case.height, case.P
    cap = 1000  # far below R
    junk = lambda n: torch.randint(0, 255, (int(n),), dtype=torch.uint8, device=dev)  # noqa: E731
    geom, img, binning = junk(lib.gsr_required_geom(P)), junk(lib.gsr_required_image(W, H)), junk(lib.gsr_required_binning(cap))
    out = torch.full((5, H, W), float("nan"), device=dev)
    radii = torch.zeros(P, dtype=torch.int32, device=dev)
    status = torch.zeros(4, dtype=torch.int32).pin_memory()
    for cam in cams:
        a = _lib.FwdArgs()
        a.P, a.D, a.M, a.width, a.height = P, 3, 16, W, H
        a.background, a.means3D, a.shs, a.colors_precomp = t["bg"].data_ptr(), t["m"].data_ptr(), t["sh"].data_ptr(), None
        a.opacities, a.scales, a.scale_modifier = t["op"].data_ptr(), t["sc"].data_ptr(), 1.0
        a.rotations, a.cov3D_precomp = t["rot"].data_ptr(), None
        a.viewmatrix, a.projmatrix, a.cam_pos = cam.viewmatrix.data_ptr(), cam.projmatrix.data_ptr(), cam.camera_center.data_ptr()
        a.tan_fovx, a.tan_fovy, a.prefiltered = cam.tanfovx, cam.tanfovy, 0
        a.out_color, a.out_depth, a.out_acc = out[:3].data_ptr(), out[3:4].data_ptr(), out[4:5].data_ptr()
        a.radii, a.debug = radii.data_ptr(), 0
        st = lib.gsr_forward_presized(C.byref(a), geom.data_ptr(), binning.data_ptr(), cap, img.data_ptr(),
                                      torch.cuda.current_stream(dev).cuda_stream, status.data_ptr())
        _lib.check(st, "gsr_forward_presized")
        torch.cuda.synchronize()
        assert int(status[3]) == 1 and int(status[0]) > cap
        assert torch.equal(out[:3], torch.ones_like(out[:3])), "background only"
        assert float(out[3:].abs().max()) == 0.0
        order = debug.image_state(img, W, H)["tile_order"].cpu().numpy()
        np.testing.assert_array_equal(np.sort(order), np.arange(order.size))
        out.fill_(float("nan"))
