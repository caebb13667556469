"""Host side of the two fused steps behind the rasterizer in GS-LIVM's mapping iteration:

  image_loss   <- (1 - lambda) l1_loss + lambda (1 - ssim)  (include/gs/gs/loss_utils.cuh:11-70,
                  src/liw/lioOptimization.cpp:1729-1734), loss AND dL/dimage from one C-ABI call (gsr_image_loss)
  FusedAdam    <- torch::optim::Adam as GaussianModel::Training_setup configures it (src/gs/gaussian.cu:396-428), over
                  flat parameter / gradient buffers (gsr_adam_step)

torch is plumbing (device memory, streams); the arithmetic is in csrc/train_ops.cu.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence, Tuple

import torch

from . import _lib


class ImageLoss:
    """Persistent buffers for one image size; `__call__(image, target)` -> (loss[3] device tensor, dL_dimage[3,H,W]).
    loss = {total, mean|image - target|, mean SSIM}. Nothing in here waits for the GPU (graph-capturable)."""

    def __init__(self, width: int, height: int, device, lambda_dssim: float = 0.2):
        self.lib = _lib.load()
        self.W, self.H, self.lam = int(width), int(height), float(lambda_dssim)
        self.device = torch.device(device)
        self.scratch = torch.empty(self.lib.gsr_image_loss_scratch_floats(self.W, self.H), dtype=torch.float32,
                                   device=self.device)
        self.grad = torch.empty((3, self.H, self.W), dtype=torch.float32, device=self.device)
        self.loss = torch.zeros(3, dtype=torch.float32, device=self.device)

    def __call__(self, image: torch.Tensor, target: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        if image.shape != (3, self.H, self.W) or target.shape != image.shape:
            raise ValueError("image and target must be [3, %d, %d]" % (self.H, self.W))
        if image.dtype != torch.float32 or target.dtype != torch.float32 or not image.is_cuda or not target.is_cuda:
            raise TypeError("image_loss needs float32 CUDA tensors (no CPU path)")
        image, target = image.contiguous(), target.contiguous()
        st = self.lib.gsr_image_loss(image.data_ptr(), target.data_ptr(), self.W, self.H, self.lam, self.grad.data_ptr(),
                                     self.loss.data_ptr(), self.scratch.data_ptr(),
                                     torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(st, "gsr_image_loss")
        return self.loss, self.grad


class FusedAdam:
    """Adam over ONE flat float32 parameter buffer with per-group learning rates, eps 1e-15 and the default betas like
    the reference (gaussian.cu:396-428). groups: sequence of (first_element, count, lr) or
    (first_element, count, lr, period, phase_lo, phase_hi, lr_other): rows of `period` elements whose elements
    [phase_lo, phase_hi) take lr and the others lr_other — e.g. ONE [P,16,3] SH array instead of the reference's
    _features_dc / _features_rest pair: (first, 48 P, lr_dc, 48, 0, 3, lr_dc / 20)."""

    def __init__(self, params: torch.Tensor, groups: Sequence[Tuple[int, int, float]], betas=(0.9, 0.999),
                 eps: float = 1e-15):
        if params.dtype != torch.float32 or not params.is_cuda or not params.is_contiguous() or params.dim() != 1:
            raise TypeError("FusedAdam needs a flat contiguous float32 CUDA buffer")
        if not 1 <= len(groups) <= 8:
            raise ValueError("1..8 parameter groups")
        groups = [tuple(g) + (0, 0, 0, 0.0) * (len(g) == 3) for g in groups]
        for first, count, *_ in groups:
            if first < 0 or count < 0 or first + count > params.numel():
                raise ValueError("group outside the parameter buffer")
        self.lib = _lib.load()
        self.params = params
        self.exp_avg = torch.zeros_like(params)
        self.exp_avg_sq = torch.zeros_like(params)
        self.betas, self.eps, self.steps = (float(betas[0]), float(betas[1])), float(eps), 0
        self._set(groups)

    def _set(self, groups) -> None:
        self._groups = (_lib.AdamGroup * len(groups))(*[
            _lib.AdamGroup(int(f), int(c), float(lr), int(per), int(lo), int(hi), float(lro))
            for f, c, lr, per, lo, hi, lro in groups])

    def set_groups(self, groups: Sequence[tuple]) -> None:
        """New element ranges / learning rates (a map that grew inside its capacity); the moments are kept."""
        if not 1 <= len(groups) <= 8:
            raise ValueError("1..8 parameter groups")
        groups = [tuple(g) + (0, 0, 0, 0.0) * (len(g) == 3) for g in groups]
        for first, count, *_ in groups:
            if first < 0 or count < 0 or first + count > self.params.numel():
                raise ValueError("group outside the parameter buffer")
        self._set(groups)

    def step(self, grads: torch.Tensor) -> None:
        if grads.dtype != torch.float32 or grads.numel() < self.params.numel() or not grads.is_contiguous():
            raise TypeError("grads must be a contiguous float32 buffer at least as long as the parameters")
        self.steps += 1
        st = self.lib.gsr_adam_step(self.params.data_ptr(), grads.data_ptr(), self.exp_avg.data_ptr(),
                                    self.exp_avg_sq.data_ptr(), self._groups, len(self._groups), self.betas[0],
                                    self.betas[1], self.eps, self.steps,
                                    torch.cuda.current_stream(self.params.device).cuda_stream)
        _lib.check(st, "gsr_adam_step")
