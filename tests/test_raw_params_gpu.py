"""GaussianModel's activations (sigmoid / exp / normalize, include/gs/gs/gaussian.cuh:40-54) fused into
the kernels. Oracle = LibTorch's own activations (and their autograd backwards) around the existing activated-value
path: the raw-parameter path must render the same image and return d loss / d RAW parameter."""
import numpy as np
import pytest
import torch

import helpers
from tolerances import assert_cds, assert_grad

pytestmark = pytest.mark.gpu


def _raw_scene(P, dev, seed=20241017 + 6, deg=3):
    import gs_livm_b200  # noqa: F401
    from gs_livm_b200 import scene
    sc_cpu = scene.make_scene(P, seed, deg)
    sc = {k: v.to(dev) for k, v in sc_cpu.items()}
    g = torch.Generator().manual_seed(seed)
    raw = dict(sc)
    # raw parameters whose activations are the scene's values, up to rounding; quaternions get arbitrary lengths
    raw["scales"] = torch.log(sc["scales"])
    op = sc["opacities"].clamp(1e-4, 1 - 1e-4)
    raw["opacities"] = torch.log(op / (1 - op))
    raw["rotations"] = sc["rotations"] * (0.25 + 3.0 * torch.rand((P, 1), generator=g)).to(dev)
    cam = scene.scene_camera(sc_cpu, 2, scene.NTU).to(dev)
    return raw, cam


def test_raw_parameter_path_matches_libtorch_activations(built_library, cuda_device):
    from gs_livm_b200 import _lib, engine
    dev = cuda_device
    P = 60_000
    raw, cam = _raw_scene(P, dev)
    H, W = cam.height, cam.width
    g = torch.Generator().manual_seed(11)
    dC = (torch.randn((3, H, W), generator=g) / (H * W)).to(dev)
    dS = (torch.randn((1, H, W), generator=g) / (H * W)).to(dev)

    # oracle: LibTorch activations, the activated-value path, autograd through the activations
    leaves = {k: raw[k].clone().requires_grad_(True) for k in ("scales", "opacities", "rotations")}
    act = dict(raw)
    act["scales"] = torch.exp(leaves["scales"])
    act["opacities"] = torch.sigmoid(leaves["opacities"])
    act["rotations"] = torch.nn.functional.normalize(leaves["rotations"], dim=1)
    ctx_a = engine.RasterContext(P, 16, W, H, dev, 3)
    col_a, dep_a, acc_a = (t.clone() for t in ctx_a.forward({k: v.detach() for k, v in act.items()}, cam))
    ga = {k: v.clone() for k, v in ctx_a.backward(dC, dS).items()}
    torch.cuda.synchronize()
    radii_a = ctx_a.radii.clone()
    torch.autograd.backward([act["scales"], act["opacities"], act["rotations"]],
                            [ga["dL_dscale"], ga["dL_dopacity"].reshape(act["opacities"].shape), ga["dL_drot"]])

    # ours: raw parameters in, gradients w.r.t. raw parameters out
    ctx_r = engine.RasterContext(P, 16, W, H, dev, 3)
    col_r, dep_r, acc_r = ctx_r.forward(raw, cam, activations=_lib.ACT_ALL)
    gr = ctx_r.backward(dC, dS)
    torch.cuda.synchronize()

    # the activations are a last-place matter (exp / the quaternion's norm may round differently from LibTorch's
    # kernels), so a radius may flip for a handful of Gaussians whose 3-sigma extent sits on an integer: not bit-exact
    same = float((ctx_r.radii == radii_a).float().mean())
    assert same > 0.9995, "radii agree for %.5f of the Gaussians" % same
    assert_cds("C", col_r.cpu().numpy(), col_a.cpu().numpy())
    assert_cds("D", dep_r.cpu().numpy(), dep_a.cpu().numpy())
    assert_cds("S", acc_r.cpu().numpy(), acc_a.cpu().numpy())
    assert_grad("dL_d_scaling (raw)", gr["dL_dscale"].cpu().numpy(), leaves["scales"].grad.cpu().numpy())
    assert_grad("dL_d_opacity (raw)", gr["dL_dopacity"].reshape(-1).cpu().numpy(), leaves["opacities"].grad.reshape(-1).cpu().numpy())
    assert_grad("dL_d_rotation (raw)", gr["dL_drot"].cpu().numpy(), leaves["rotations"].grad.cpu().numpy())
    for n in ("dL_dmean3D", "dL_dsh"):
        assert_grad(n, gr[n].cpu().numpy(), ga[n].cpu().numpy())
    # a raw quaternion's gradient is orthogonal to it (the length does not matter)
    radial = (gr["dL_drot"] * raw["rotations"]).sum(1).abs().max() / (gr["dL_drot"].abs().max() * raw["rotations"].abs().max())
    assert float(radial) < 1e-4


def test_adam_row_pattern_equals_two_parameter_groups(built_library, cuda_device):
    """ONE [P,16,3] SH array with the learning rates as a pattern over the 48-float row == torch.optim.Adam over the
    reference's separate _features_dc / _features_rest tensors (src/gs/gaussian.cu:406-412)."""
    import gs_livm_b200  # noqa: F401
    from gs_livm_b200 import train
    P, lr_dc = 3001, 2.5e-3
    g = torch.Generator().manual_seed(5)
    sh = torch.randn((P, 16, 3), generator=g).to(cuda_device)
    extra = torch.randn(7, generator=g).to(cuda_device)  # a second, plain group behind a range that is not a multiple of 4
    flat = torch.cat([sh.reshape(-1), extra]).contiguous()
    dc = sh[:, :1, :].clone().requires_grad_(True)
    rest = sh[:, 1:, :].clone().requires_grad_(True)
    ex = extra.clone().requires_grad_(True)
    ref = torch.optim.Adam([dict(params=[dc], lr=lr_dc), dict(params=[rest], lr=lr_dc / 20), dict(params=[ex], lr=1e-2)],
                           lr=0.0, eps=1e-15)
    ours = train.FusedAdam(flat, [(0, 48 * P, lr_dc, 48, 0, 3, lr_dc / 20), (48 * P, 7, 1e-2)])
    for _ in range(10):
        grads = torch.randn(flat.numel(), generator=g).to(cuda_device) * 1e-3
        gs = grads[:48 * P].view(P, 16, 3)
        dc.grad, rest.grad, ex.grad = gs[:, :1, :].clone(), gs[:, 1:, :].clone(), grads[48 * P:].clone()
        ref.step()
        ours.step(grads)
    torch.cuda.synchronize()
    want = torch.cat([torch.cat([dc, rest], dim=1).detach().reshape(-1), ex.detach()])
    assert float((flat - want).abs().max()) <= 1e-6
