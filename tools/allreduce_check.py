"""Multi-GPU check of window.GradExchange (run under torchrun): every mode gives the same sum on every rank as
NCCL's all-reduce, bit-identical across ranks, and how long the 236 MB exchange of the bench takes in each."""
import os, sys, time
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gs_livm_b200
from gs_livm_b200 import window
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
n = 59 * 1_000_000
g = torch.Generator(device=dev).manual_seed(100 + rank)
src = torch.randn(n, device=dev, generator=g)
ref = src.clone(); dist.all_reduce(ref)
for mode in ("multimem", "p2p", "nccl"):
    ex = window.GradExchange(n + 1024, n, dev, prefer=mode)
    ex.flat[:n].copy_(src); ex.flat[n:].fill_(7.0)
    ex.allreduce_(); torch.cuda.synchronize()
    err = (ex.flat[:n] - ref).abs().max().item(); tail_ok = bool((ex.flat[n:] == 7.0).all())
    gathered = [torch.empty(4096, device=dev) for _ in range(world)]
    dist.all_gather(gathered, ex.flat[12345:12345 + 4096].contiguous())
    same = all(torch.equal(gathered[0], t) for t in gathered)
    chk = torch.tensor([ex.flat[:n].double().sum().item()], device=dev, dtype=torch.float64)
    lst = [torch.empty_like(chk) for _ in range(world)]; dist.all_gather(lst, chk)
    same_sum = all(float(x) == float(lst[0]) for x in lst)
    for _ in range(3): ex.allreduce_()
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ex.allreduce_()
    e1.record(); torch.cuda.synchronize()
    if rank == 0:
        print(f"requested {mode:9s} -> mode {ex.mode:9s} {ex.why} max|err vs NCCL| {err:.3e} tail untouched {tail_ok} "
              f"ranks identical {same and same_sum}  {e0.elapsed_time(e1) / 10 * 1e3:8.1f} us per 236 MB exchange", flush=True)
    del ex
# sparse rows: the layout of the bench (dL_dsh | dL_dmean3D | ...), every rank "renders" a different 55 % of the Gaussians
# and the rows it did not render are zero, as geom_bwd leaves them
P = 1_000_000
for mode in ("multimem", "p2p"):
    ex = window.GradExchange(n + 1024, n, dev, prefer=mode, sparse_rows=(0, P, 48))
    if ex.mode == "nccl":
        continue
    vis = torch.rand(P, device=dev, generator=g) < 0.55
    radii = vis.to(torch.int32) * 3
    mine = src.clone()
    mine[:48 * P].view(P, 48)[~vis] = 0.0
    want = mine.clone(); dist.all_reduce(want)
    ex.flat[:n].copy_(mine)
    ex.allreduce_(radii=radii); torch.cuda.synchronize()
    err = (ex.flat[:n] - want).abs().max().item()
    for _ in range(3): ex.allreduce_(radii=radii)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ex.allreduce_(radii=radii)
    e1.record(); torch.cuda.synchronize()
    uni = torch.zeros(P, device=dev); uni[vis] = 1; dist.all_reduce(uni, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"sparse rows, mode {ex.mode:9s}: max|err vs NCCL| {err:.3e}, union of rendered Gaussians {float(uni.mean()):.2f}, "
              f"{e0.elapsed_time(e1) / 10 * 1e3:8.1f} us per exchange", flush=True)
    del ex
dist.barrier(); dist.destroy_process_group()
