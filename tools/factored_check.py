"""Multi-GPU check of window.FactoredExchange (run under torchrun, one rank per GPU): every rank renders its own view of
the bench scene; the factored exchange (pushed dRGB + 11-float all-reduce + local SH expansion) must give the dense sum
of the ranks' parameter gradients (NCCL all-reduce of what the dense backward writes from the SAME accumulators) and
leave identical bits on every rank. Prints one line per mode combination; then times the step eagerly and from a graph.

    torchrun --nproc-per-node N tools/factored_check.py [P]"""
import os, sys, time
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gs_livm_b200  # noqa
from gs_livm_b200 import engine, scene, window

local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
P = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
SEED = 20241017 + 2
sc_cpu = scene.make_scene(P, SEED, 3)
sc = {k: v.to(dev) for k, v in sc_cpu.items()}
cams = [scene.scene_camera(sc_cpu, i, scene.NTU).to(dev) for i in range(8)]
H, W = cams[0].height, cams[0].width
g = torch.Generator().manual_seed(SEED + 1)
dC = (torch.randn((3, H, W), generator=g) / (H * W)).to(dev)
dS = (torch.randn((1, H, W), generator=g) / (H * W)).to(dev)

combos = [("p2p", "p2p", True, 1), ("multimem", "multimem", True, 1), ("multimem", "multimem", True, 2),
          ("multimem", "multimem", True, 4), ("multimem", "multimem", False, 1), ("auto", "auto", True, 4),
          ("multimem", "multimem", True, -1), ("p2p", "p2p", True, -1)]  # chunks -1: one chunk, expansion behind the reduce
if len(sys.argv) > 2:
    combos = [c for i, c in enumerate(combos) if str(i) in sys.argv[2]]
for reduce_mode, push_mode, sparse, track in combos:
    try:
        ex = window.FactoredExchange(P, 16, dev, reduce_mode=reduce_mode, push_mode=push_mode, sparse_reduce=sparse,
                                     n_chunks=max(track, 1), expand_after_reduce=track < 0)
    except RuntimeError as e:
        if rank == 0:
            print(f"factored reduce={reduce_mode} push={push_mode}: skipped ({e})", flush=True)
        continue
    ctx = engine.RasterContext(P, 16, W, H, dev, 3, grad_flat=ex.flat, track_rows=False)
    n_param = ctx.param_grad_floats

    def cams_of(i):
        return [cams[(r + i) % 8] for r in range(world)]

    def set_centres(i):
        ex.cam_centres.copy_(torch.stack([c.camera_center for c in cams_of(i)]))

    # learn the arena on every pose
    for k in range(8):
        ctx.forward(sc, cams[k]); ctx.backward(dC, dS)
        if not ctx.check():
            ctx.forward(sc, cams[k], redo=True); ctx.check()
    worst, same_all = 0.0, True
    for i in range(3):
        cam = cams_of(i)[rank]
        set_centres(i)
        # dense reference from the same accumulators: blend half once, dense geom half, NCCL sum
        ctx.forward(sc, cam)
        ctx.backward(dC, dS, part="blend")
        ctx.backward(dC, dS, part="geom")
        want = ctx.param_grads().clone()
        dist.all_reduce(want)
        # factored: push + geom half without the SH rows + exchange, from the same accumulators
        ctx.grads["dL_dsh"].fill_(float("nan"))
        ex.mark_rows_dirty()
        ev = torch.cuda.Event(); ev.record()
        ex.exchange(ctx, sc["means3D"], i % 2, ev,
                    lambda: ctx.backward(dC, dS, part="geom", sh_factored=True))
        torch.cuda.synchronize()
        got = ctx.param_grads()
        scale = float(want.abs().max())
        err = float((got - want).abs().max()) / scale
        worst = max(worst, err)
        chk = torch.tensor([got.double().sum().item(), got.double().abs().sum().item()], device=dev, dtype=torch.float64)
        lst = [torch.empty_like(chk) for _ in range(world)]; dist.all_gather(lst, chk)
        same_all &= all(torch.equal(x, lst[0]) for x in lst)
    # the segments of real (eager) steps, from events inside FactoredExchange.exchange: us after the blend backward
    ex.segment_events = {}
    seg = {}
    for i in range(4, 12):
        set_centres(i); ex.render_step(ctx, sc, cams_of(i)[rank], dC, dS, parity=i % 2)
        torch.cuda.synchronize()
        E = ex.segment_events
        for name, (a, b) in {"push": ("push_begin", "push_end"), "geom": ("geom_begin", "geom_end"),
                             "geom_end->reduce_end": ("geom_end", "reduce_end"), "blend_end->expand_end": ("geom_begin", "expand_end"),
                             "blend_end->step_end": ("geom_begin", "step_end")}.items():
            seg[name] = seg.get(name, 0.0) + E[a].elapsed_time(E[b]) * 1e3 / 8
    ex.segment_events = None
    t = torch.tensor(list(seg.values()), device=dev, dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("   segments, max over ranks (us): " + ", ".join(f"{k} {float(v):.1f}" for k, v in zip(seg, t)), flush=True)
    # timing: the full step, eager and captured
    def step(i):
        ex.render_step(ctx, sc, cams_of(i)[rank], dC, dS, parity=i % 2)
    for i in range(4):
        set_centres(i); step(i)
    assert ctx.check()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(16):
        step(i)
    e1.record(); torch.cuda.synchronize()
    eager_us = e0.elapsed_time(e1) / 16 * 1e3
    graph_us = float("nan")
    try:
        frames = []
        for i in range(8):
            set_centres(i)
            frames.append(ctx.capture(lambda i=i: step(i)))
        centres = [torch.stack([c.camera_center for c in cams_of(i)]) for i in range(8)]
        for i in range(8):
            ex.cam_centres.copy_(centres[i]); frames[i].replay()
        torch.cuda.synchronize(); dist.barrier()
        e0.record()
        for i in range(32):
            ex.cam_centres.copy_(centres[i % 8], non_blocking=True); frames[i % 8].replay()
        e1.record(); torch.cuda.synchronize()
        graph_us = e0.elapsed_time(e1) / 32 * 1e3
        assert ctx.check()
    except Exception as e:  # noqa
        if rank == 0:
            print("graph capture failed:", repr(e)[:300], flush=True)
    t = torch.tensor([eager_us, graph_us], device=dev, dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"factored reduce={ex.reduce_mode} push={ex.push_mode} sparse={sparse} chunks={track} (asked {reduce_mode}/{push_mode}): max|err|/max|ref| vs dense NCCL "
              f"{worst:.3e}, ranks identical {same_all}, step eager {float(t[0]):.1f} us, graph {float(t[1]):.1f} us", flush=True)
    del ctx, ex
    torch.cuda.synchronize(); dist.barrier()

if len(sys.argv) > 2:
    dist.barrier(); dist.destroy_process_group(); sys.exit(0)
# the dense exchange for comparison (window.GradExchange, as bench.py used it in round 1)
grad_floats = 75 * P
exd = window.GradExchange(grad_floats, 59 * P, dev, sparse_rows=(0, P, 48))
ctx = engine.RasterContext(P, 16, W, H, dev, 3, grad_flat=exd.flat, track_rows=False)
for k in range(8):
    ctx.forward(sc, cams[k]); ctx.backward(dC, dS)
    if not ctx.check():
        ctx.forward(sc, cams[k], redo=True); ctx.check()
def dstep(i):
    ctx.forward(sc, cams[(rank + i) % 8]); ctx.backward(dC, dS); exd.allreduce_(radii=ctx.radii)
for i in range(4): dstep(i)
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(16): dstep(i)
e1.record(); torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / 16 * 1e3], device=dev, dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"dense exchange mode {exd.mode}: step eager {float(t[0]):.1f} us", flush=True)
dist.barrier(); dist.destroy_process_group()
