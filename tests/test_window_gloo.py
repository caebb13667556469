"""The N>1 host path on CPU: two gloo ranks shard a window of views, sum their flat parameter gradients with one
all-reduce and must end with bit-identical parameters. The per-view 'renderer' here is the CPU oracle (tests may use
it); on the GPU box the same WindowStep drives RasterContext (tests/test_multi_gpu_step.py, bench.py --gpus N)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _view_grad(view: int):
    """Flat gradient (mean3D | scale | rot | opacity) of one window view from the CPU oracle."""
    intr = dict(width=48, height=32, fx=36.0, fy=36.0)
    case = helpers.scene_case(400, 991, 0, intr=intr, step=view, heavy=True)
    out = helpers.run_oracle_cpu(case, nthreads=1)
    flat = np.concatenate([out[k].ravel() for k in ("dL_dmean3D", "dL_dscale", "dL_drot", "dL_dopacity")])
    return case, torch.from_numpy(flat.astype(np.float32))


def _worker(rank, world, port, n_views, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import gs_livm_b200  # noqa: F401
    from gs_livm_b200 import window
    torch.manual_seed(0)
    case, g0 = _view_grad(0)
    params = [case.means3D.clone(), case.scales.clone(), case.rotations.clone(), case.opacities.clone()]

    def render_view(v, acc):
        _, g = _view_grad(v)
        acc += g
        return float(g.abs().sum())

    step = window.WindowStep(n_views, g0.numel(), "cpu", render_view)
    stats = step.run()
    window.sgd_update_(params, step.flat, lr=0.1)
    # numpy copies travel by value; torch tensors would be handed over as shared-memory descriptors that the
    # parent can only open while this process is still alive
    q.put((rank, step.views, step.flat.numpy().copy(), [p.numpy().copy() for p in params], stats))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_ranks_shard_views_and_stay_identical():
    n_views, world = 4, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_views, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=240) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (r0, v0, f0, p0, s0), (r1, v1, f1, p1, s1) = res
    f0, f1 = torch.from_numpy(f0), torch.from_numpy(f1)
    p0, p1 = [torch.from_numpy(a) for a in p0], [torch.from_numpy(a) for a in p1]
    assert v0 == [0, 2] and v1 == [1, 3]
    assert torch.equal(f0, f1), "all-reduced gradient must be bit-identical on every rank"
    for a, b in zip(p0, p1):
        assert torch.equal(a, b), "replicated parameters diverged"
    # and it is the SUM over all four views (single-process reference)
    total = sum(_view_grad(v)[1] for v in range(n_views))
    assert torch.allclose(f0, total, rtol=1e-5, atol=1e-9)
    assert s0["local_views"] == 2 and s1["local_views"] == 2
