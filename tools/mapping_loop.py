"""BASELINE.json configs[4]: the mapping thread's loop on a replayed synthetic stream — per keyframe a LiDAR sweep
(the returns within `range` m of the pose), VGPR + Gaussian insertion on the GPU, then `iters` optimisation iterations
over the sliding window of the last 8 keyframes (one view per rank and iteration, fused loss, factored gradient
exchange, fused Adam). Single process, or one rank per GPU under torchrun. Prints ONE JSON line (rank 0).

    python tools/mapping_loop.py [--keyframes K] [--iters I] [--points N] [--capacity P]
    torchrun --nproc-per-node 8 tools/mapping_loop.py ..."""
import argparse, json, os, sys, time
import numpy as np
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gs_livm_b200  # noqa
from gs_livm_b200 import mapping, scene, vgpr


run = mapping.run_synthetic_stream


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--keyframes", type=int, default=6)
    ap.add_argument("--iters", type=int, default=100)
    ap.add_argument("--points", type=int, default=2_000_000)
    ap.add_argument("--capacity", type=int, default=1_500_000)
    a = ap.parse_args()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        dist.init_process_group("nccl", device_id=dev)
    res = run(a.keyframes, a.iters, a.points, a.capacity, dev=dev)
    if res is not None:
        print(json.dumps(res))
    if dist.is_initialized():
        dist.barrier(); dist.destroy_process_group()
