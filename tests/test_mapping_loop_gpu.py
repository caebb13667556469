"""BASELINE.json configs[4] in the small: the mapping loop (gs_livm_b200/mapping.py) on a replayed synthetic stream —
VGPR + insertion per keyframe, optimisation iterations on raw parameters with the fused loss and fused Adam. The map
grows inside its capacity, the loss falls, every frame fits its arena, and the optimised map renders the same image
through the activated-value entry point."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_mapping_loop_inserts_and_optimises(built_library, cuda_device):
    import gs_livm_b200  # noqa: F401
    from gs_livm_b200 import _lib, engine, mapping, scene, vgpr
    dev = cuda_device
    intr = dict(width=376, height=240, fx=212.513, fy=213.399)
    W, H = intr["width"], intr["height"]
    cloud = torch.from_numpy(scene.lidar_frame(400_000, 3, length=30.0)).to(dev)
    image_u8 = torch.from_numpy(scene.procedural_image(intr)).to(dev)
    target = (image_u8.permute(2, 0, 1).float() / 255.0).contiguous()
    loop = mapping.MappingLoop(300_000, W, H, dev)
    counts, window = [], []
    for kf in range(3):
        cam = scene.track_camera(kf * 4, intr, start=np.array([2.0, 0.0, 1.5])).to(dev)
        near = (cloud - cam.camera_center.double()[None, :]).norm(dim=1) < 8.0
        view = cam.viewmatrix
        added = loop.insert(cloud[near], lambda xyz: vgpr.project_colours(xyz, view, intr["fx"], intr["fy"], W, H, image_u8))
        counts.append(loop.count)
        assert added > 0 or kf > 0
        window = (window + [(cam, target)])[-8:]
        losses = []
        for it in range(25):
            c, t = window[it % len(window)]
            losses.append(loop.iterate([c], [t]).clone())
        torch.cuda.synchronize()
        assert loop.check(), "every frame must fit its presized arena"
        vals = [float(x[0]) for x in losses]
        assert np.isfinite(vals).all()
        if kf == 0:
            assert vals[-1] < 0.9 * vals[0], "the loss must fall: %.4f -> %.4f" % (vals[0], vals[-1])
    assert counts[0] > 10_000 and counts[-1] > counts[0] and counts[-1] <= 300_000
    assert np.isfinite(loop.params.cpu().numpy()).all()
    # the optimised map through the activated-value entry point: the same image as through the raw-parameter one
    cam = window[-1][0]
    raw_img = loop.ctx.forward(dict(loop.raw, background=loop.background), cam, activations=_lib.ACT_ALL)[0].clone()
    act = loop.activated()
    ctx = engine.RasterContext(loop.count, 16, W, H, dev, 3)
    img = ctx.forward(act, cam)[0]
    torch.cuda.synchronize()
    assert float((img - raw_img).abs().max()) <= 1e-5
    assert float(raw_img.mean()) > 0.05  # something was rendered
