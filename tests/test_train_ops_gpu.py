"""The two fused steps behind the rasterizer against plain PyTorch fp32 restatements of the
reference's LibTorch code: the image loss of include/gs/gs/loss_utils.cuh:11-70 combined as in
src/liw/lioOptimization.cpp:1729-1734 (forward value AND gradient through autograd), and torch.optim.Adam configured
like GaussianModel::Training_setup (src/gs/gaussian.cu:396-428)."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def reference_window(window_size=11, sigma=1.5):
    """loss_utils.cuh:24-36: exp(-floor((x - window_size) / 2)^2 / (2 sigma^2)), normalised; outer product."""
    g = torch.tensor([math.exp(-(math.floor((x - window_size) / 2.0) ** 2) / (2.0 * sigma * sigma))
                      for x in range(window_size)], dtype=torch.float32)
    g = g / g.sum()
    w2 = g.unsqueeze(1).mm(g.unsqueeze(0))
    return w2.expand(3, 1, window_size, window_size).contiguous()


def reference_loss(image, target, lam):
    """(1 - lam) * l1_loss + lam * (1 - ssim), loss_utils.cuh:11-13,43-70 with conv2d(padding = 5, groups = 3)."""
    F = torch.nn.functional
    l1 = (image - target).abs().mean()
    if lam == 0.0:
        return l1, l1, torch.zeros(())
    win = reference_window().to(image.device)
    a, b = image.unsqueeze(0), target.unsqueeze(0)
    conv = lambda t: F.conv2d(t, win, padding=5, groups=3)  # noqa: E731
    mu1, mu2 = conv(a), conv(b)
    s1, s2, s12 = conv(a * a) - mu1.pow(2), conv(b * b) - mu2.pow(2), conv(a * b) - mu1 * mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    ssim_map = ((2 * mu1 * mu2 + C1) * (2 * s12 + C2)) / ((mu1.pow(2) + mu2.pow(2) + C1) * (s1 + s2 + C2))
    ss = ssim_map.mean()
    return (1 - lam) * l1 + lam * (1 - ss), l1, ss


@pytest.mark.parametrize("W,H", [(752, 480), (376, 240), (100, 50), (17, 9)])
@pytest.mark.parametrize("lam", [0.2, 0.0, 1.0])
def test_image_loss_matches_the_reference_formula(W, H, lam, built_library, cuda_device):
    import gs_livm_b200  # noqa: F401
    from gs_livm_b200 import train
    g = torch.Generator().manual_seed(W * 7 + H)
    target = torch.rand((3, H, W), generator=g).to(cuda_device)
    image = (target + 0.2 * torch.randn((3, H, W), generator=g).to(cuda_device)).clamp(0, 1.3)
    image[:, : H // 3, : W // 3] = target[:, : H // 3, : W // 3]  # a region of exact zeros in |x - y|: sign(0) = 0
    ref_in = image.clone().requires_grad_(True)
    ref_loss, ref_l1, ref_ss = reference_loss(ref_in, target, lam)
    ref_loss.backward()
    fused = train.ImageLoss(W, H, cuda_device, lambda_dssim=lam)
    loss, grad = fused(image, target)
    torch.cuda.synchronize()
    got = loss.cpu().numpy()
    assert abs(got[0] - float(ref_loss)) <= 2e-6 * max(1.0, abs(float(ref_loss)))
    assert abs(got[1] - float(ref_l1)) <= 2e-6
    if lam != 0.0:
        assert abs(got[2] - float(ref_ss)) <= 5e-6
    scale = float(ref_in.grad.abs().max())
    err = float((grad - ref_in.grad).abs().max())
    assert err <= 2e-5 * scale, "dL/dimage: max abs err %.3e vs scale %.3e" % (err, scale)


def test_adam_step_matches_torch_adam_with_the_reference_groups(built_library, cuda_device):
    """Six groups with the reference's relative learning rates, eps 1e-15, 25 steps, against torch.optim.Adam."""
    import gs_livm_b200  # noqa: F401
    from gs_livm_b200 import train
    P = 5000
    widths = dict(xyz=3, f_dc=3, f_rest=45, scaling=3, rotation=4, opacity=1)
    lrs = dict(xyz=1.6e-4, f_dc=2.5e-3, f_rest=2.5e-3 / 20, scaling=5e-3, rotation=1e-3, opacity=5e-2)
    g = torch.Generator().manual_seed(3)
    flat = torch.randn(sum(widths.values()) * P, generator=g).to(cuda_device)
    groups, off, ref_params = [], 0, []
    for name, w in widths.items():
        groups.append((off, w * P, lrs[name]))
        ref_params.append(flat[off:off + w * P].clone().requires_grad_(True))
        off += w * P
    ref_opt = torch.optim.Adam([dict(params=[p], lr=lrs[n]) for p, n in zip(ref_params, widths)], lr=0.0, eps=1e-15)
    ours = train.FusedAdam(flat, groups)
    for step in range(25):
        grads = (torch.randn(flat.numel(), generator=g) * (10.0 ** float(torch.randint(-6, 1, (1,), generator=g)))).to(cuda_device)
        if step % 5 == 4:
            grads[::7] = 0.0  # Gaussians no view rendered: zero gradient, the moments still decay
        o = 0
        for p in ref_params:
            p.grad = grads[o:o + p.numel()].clone()
            o += p.numel()
        ref_opt.step()
        ours.step(grads)
    torch.cuda.synchronize()
    want = torch.cat([p.detach() for p in ref_params])
    err = float((flat - want).abs().max())
    assert err <= 1e-5, "parameters after 25 Adam steps differ by %.3e" % err
    assert np.isfinite(flat.cpu().numpy()).all()
