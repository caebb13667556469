"""The window's gradient exchange on real GPUs (needs >= 2 of them; the single-GPU box skips it): every mode of
window.GradExchange — libgsr_b200's NVSwitch-multicast kernel, its two-shot peer-to-peer kernels, NCCL — gives the
NCCL sum, leaves the rest of the buffer alone and ends with identical bits on every rank; the row-sparse exchange
gives exactly the dense result."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.timeout(600)
def test_grad_exchange_modes_agree_across_ranks(built_library):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533",
                          os.path.join(ROOT, "tools", "allreduce_check.py")], capture_output=True, text=True, timeout=500)
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("requested")]
    assert out.returncode == 0 and len(lines) == 3, out.stdout + out.stderr
    for ln in lines:
        assert "tail untouched True" in ln and "ranks identical True" in ln, ln
        err = float(ln.split("max|err vs NCCL|")[1].split()[0])
        assert err <= 1e-4, ln
    # the sparse exchange (rows of Gaussians no rank rendered are skipped) must give the dense sum exactly
    sparse = [ln for ln in out.stdout.splitlines() if ln.startswith("sparse rows")]
    assert sparse, out.stdout
    for ln in sparse:
        assert float(ln.split("max|err vs NCCL|")[1].split(",")[0]) == 0.0, ln


@pytest.mark.gpu
@pytest.mark.timeout(600)
def test_factored_exchange_gives_the_dense_sum_on_every_rank(built_library):
    """window.FactoredExchange on two GPUs (tools/factored_check.py): pushed dRGB + sparse 11-float all-reduce + local SH
    expansion == NCCL's all-reduce of the dense gradients from the same accumulators, identical bits on both ranks, in
    every mode combination the system offers, eagerly and from a captured graph."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29534",
                          os.path.join(ROOT, "tools", "factored_check.py"), "200000", "0145"],
                         capture_output=True, text=True, timeout=500)
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("factored reduce=")]
    assert out.returncode == 0 and len(lines) >= 2, out.stdout[-3000:] + out.stderr[-3000:]
    for ln in lines:
        assert "ranks identical True" in ln, ln
        assert float(ln.split("vs dense NCCL")[1].split(",")[0]) <= 1e-5, ln
        assert "graph nan" not in ln, ln  # the step must be capturable
