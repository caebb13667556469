"""The mapping thread's GPU work as one loop (BASELINE.json configs[4]; src/liw/lioOptimization.cpp:1492-1847): for every
keyframe, VGPR + variance-centred initialisation of the voxels the sweep newly covers -> Gaussians appended to the map
(GaussianModel::addNewPointcloud, src/gs/gaussian.cu:241-313) -> optimisation iterations over the sliding keyframe
window: render on the RAW parameters (activations fused, GSR_ACT_*), fused image loss (L1 + lambda DSSIM) -> dL/dC,
backward, gradient exchange across the ranks (one view per rank and iteration, window.FactoredExchange), fused Adam with
the reference's six learning rates. Poses are given (the odometry is outside the replaced path; the synthetic stream
replays them).

Everything lives in buffers laid out ONCE for the map's capacity: parameters, Adam moments and gradients are flat
arrays [capacity, ...] whose first `count` rows are in use (engine.RasterContext.set_active), so inserting Gaussians
never re-allocates or copies the map (the reference torch::cat's every parameter and optimiser-state tensor per
insertion). Every rank runs the same insertion on the same sweep, so the replicas stay identical without a broadcast.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence

import torch
import torch.distributed as dist

from . import _lib, engine, train, vgpr, window

# config/basic_common.yaml:56-62 (position_lr_init x spatial_lr_scale 1, feature_lr, feature_lr / 20, scaling, rotation, opacity)
LEARNING_RATES = dict(xyz=0.0005, f_dc=0.001, f_rest=0.001 / 20.0, scaling=0.0025, rotation=0.0025, opacity=0.025)


class MappingLoop:
    def __init__(self, capacity: int, width: int, height: int, device, sh_degree: int = 3, lambda_dssim: float = 0.2,
                 lrs: Optional[Dict[str, float]] = None, group=None, grid: float = 0.2):
        self.cap, self.W, self.H, self.D = int(capacity), int(width), int(height), int(sh_degree)
        self.device = torch.device(device)
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        self.lrs = dict(LEARNING_RATES, **(lrs or {}))
        self.grid = float(grid)
        at, _, self.param_floats, _ = engine.grad_layout(self.cap, 16)
        self._at = at
        self.params = torch.zeros(self.param_floats, dtype=torch.float32, device=self.device)
        P = self.cap
        v = lambda n, shape: self.params[at[n][0]:at[n][0] + at[n][1] * P].view(shape)  # noqa: E731
        # views of the flat buffer in the gradient buffer's layout, so that gradient element i belongs to parameter i
        self.raw = {"shs": v("dL_dsh", (P, 16, 3)), "means3D": v("dL_dmean3D", (P, 3)), "scales": v("dL_dscale", (P, 3)),
                    "rotations": v("dL_drot", (P, 4)), "opacities": v("dL_dopacity", (P, 1))}
        self.raw["rotations"][:, 0] = 1.0  # (unused rows stay valid quaternions)
        self.background = torch.zeros(3, dtype=torch.float32, device=self.device)
        self.exchange = window.FactoredExchange(self.cap, 16, self.device, group=group) if self.world > 1 else None
        self.ctx = engine.RasterContext(self.cap, 16, self.W, self.H, self.device, self.D,
                                        grad_flat=self.exchange.flat if self.exchange else None,
                                        track_rows=self.exchange is None)
        self.loss_fn = train.ImageLoss(self.W, self.H, self.device, lambda_dssim=lambda_dssim)
        self.adam = train.FusedAdam(self.params, [(0, 0, 0.0)])
        self.count = 0
        self._seen = torch.empty(0, dtype=torch.int64, device=self.device)  # packed keys of the voxels already in the map
        self._zero_acc = torch.zeros((1, self.H, self.W), dtype=torch.float32, device=self.device)
        self._step = 0
        self._sized = False

    # ------------------------------------------------------------------------------------------------ insertion
    def insert(self, sweep: torch.Tensor, colour_of: Callable) -> int:
        """VGPR + initialisation of the sweep's voxels that are not in the map yet; the new Gaussians are appended.
        -> how many were added (0 when the capacity is exhausted)."""
        key3 = torch.floor(sweep.double() / self.grid).to(torch.int64)
        packed = ((key3[:, 0] + vgpr._OFF) << 42) | ((key3[:, 1] + vgpr._OFF) << 21) | (key3[:, 2] + vgpr._OFF)
        fresh = ~torch.isin(packed, self._seen)
        pts = sweep[fresh]
        if pts.shape[0] == 0:
            return 0
        g = vgpr.vgpr_frame(pts, colour_of, grid=self.grid, sh_degree=self.D, raw=True)
        n = min(int(g["means3D"].shape[0]), self.cap - self.count)
        self._seen = torch.unique(torch.cat([self._seen, packed[fresh]]))
        if n <= 0:
            return 0
        lo, hi = self.count, self.count + n
        for name in ("means3D", "scales", "rotations", "opacities", "shs"):
            self.raw[name][lo:hi].copy_(g[name][:n].reshape(self.raw[name][lo:hi].shape))
        self.count = hi
        self.ctx.set_active(hi)
        if self.exchange is not None:
            self.exchange.set_active(hi)
        at, c, lr = self._at, self.count, self.lrs
        self.adam.set_groups([
            (at["dL_dsh"][0], 48 * c, lr["f_dc"], 48, 0, 3, lr["f_rest"]),
            (at["dL_dmean3D"][0], 3 * c, lr["xyz"]), (at["dL_dscale"][0], 3 * c, lr["scaling"]),
            (at["dL_drot"][0], 4 * c, lr["rotation"]), (at["dL_dopacity"][0], c, lr["opacity"])])
        self.ctx.invalidate_grads()
        self._sized = False
        return n

    # ------------------------------------------------------------------------------------------------ optimisation
    def iterate(self, cams: Sequence, targets: Sequence[torch.Tensor]) -> torch.Tensor:
        """One optimisation iteration: this rank renders cams[rank] against targets[rank] ([3,H,W] on the device); with
        several ranks `cams` holds the iteration's views in rank order. -> the device tensor {loss, L1, SSIM} of this
        rank's view (read it whenever convenient: nothing in here waits for the GPU, except a first frame after an
        insertion, which sizes the binning arena)."""
        if self.count == 0:
            raise RuntimeError("the map is empty: insert() first")
        ctx, cam, target = self.ctx, cams[self.rank], targets[self.rank]
        main = torch.cuda.current_stream(self.device)
        scene = dict(self.raw, background=self.background)
        if not self._sized:  # the instance count of the grown map is not known yet: one synchronous forward
            ctx._capacity = 0
            ctx.forward(scene, cam, activations=_lib.ACT_ALL)
            self._sized = True
        if self.exchange is None:
            ctx.forward(scene, cam, activations=_lib.ACT_ALL)
            loss, dimg = self.loss_fn(ctx.out_color, target)
            ctx.backward(dimg, self._zero_acc)
        else:
            ex = self.exchange
            ex.cam_centres.copy_(torch.stack([c.camera_center for c in cams]), non_blocking=True)
            ex.side.wait_stream(main)
            with torch.cuda.stream(ex.side):
                ctx.prepare_backward()
            ctx.forward(scene, cam, activations=_lib.ACT_ALL)
            loss, dimg = self.loss_fn(ctx.out_color, target)
            main.wait_stream(ex.side)
            ctx.backward(dimg, self._zero_acc, accumulators_zeroed=True, part="blend")
            ev = torch.cuda.Event()
            ev.record(main)
            ex.exchange(ctx, self.raw["means3D"], self._step % 2, ev,
                        lambda: ctx.backward(dimg, self._zero_acc, part="geom", sh_factored=True))
        self._step += 1
        self.adam.step(ctx.param_grads())
        return loss

    def check(self) -> bool:
        """True if every frame since the last call fitted its presized binning arena (synchronises); after a False the
        next iterate() re-sizes it."""
        ok = self.ctx.check()
        if not ok:
            self._sized = False
        return ok

    def activated(self) -> Dict[str, torch.Tensor]:
        """The map as the rasterizer's activated-value entry point would take it (a copy)."""
        n = self.count
        return dict(means3D=self.raw["means3D"][:n].clone(), shs=self.raw["shs"][:n].clone(),
                    scales=torch.exp(self.raw["scales"][:n]), opacities=torch.sigmoid(self.raw["opacities"][:n]),
                    rotations=torch.nn.functional.normalize(self.raw["rotations"][:n], dim=1), background=self.background)


def run_synthetic_stream(keyframes=6, iters=100, points=2_000_000, capacity=1_500_000, sensor_range=12.0, intr=None,
                         seed=11, dev=None):
    """The loop on a replayed synthetic stream: poses along scene.track_camera, per keyframe the returns of a fixed
    canyon point cloud within `sensor_range` m of the pose, the procedural keyframe image as colour source and target.
    (The image is fixed in image space, i.e. not multi-view consistent: the loss values say the loop optimises, not that
    it converges to a scene.) -> a dict of timings (rank 0; None elsewhere)."""
    import time
    import numpy as np
    from . import scene
    intr = intr or scene.NTU
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    W, H = intr["width"], intr["height"]
    cloud = torch.from_numpy(scene.lidar_frame(points, seed, length=30.0)).to(dev)  # the canyon's surfaces, float64
    image_u8 = torch.from_numpy(scene.procedural_image(intr)).to(dev)
    target = (image_u8.permute(2, 0, 1).float() / 255.0).contiguous()
    loop = MappingLoop(capacity, W, H, dev)
    torch.linalg.eigh(torch.eye(3, dtype=torch.float64, device=dev).expand(4, 3, 3).contiguous())  # (solver start-up is not the loop's)
    window, log = [], []
    torch.cuda.synchronize()
    t_all = time.perf_counter()
    for kf in range(keyframes):
        cam_cpu = scene.track_camera(kf, intr, start=np.array([2.0, 0.0, 1.5]))
        cam = cam_cpu.to(dev)
        t0 = time.perf_counter()
        near = ((cloud - cam.camera_center.double()[None, :]).norm(dim=1) < sensor_range)
        sweep = cloud[near]
        view = cam.viewmatrix
        added = loop.insert(sweep, lambda xyz: vgpr.project_colours(xyz, view, intr["fx"], intr["fy"], W, H, image_u8))
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        window = (window + [(cam, target)])[-8:]
        first = last = None
        for it in range(iters):
            views = [window[(it * world + r) % len(window)] for r in range(world)]
            loss = loop.iterate([v[0] for v in views], [v[1] for v in views])
            if it == 0:
                first = float(loss[0])  # (synchronises: the first iteration after an insertion does anyway)
        torch.cuda.synchronize()
        last = float(loss[0])
        fits = loop.check()
        t2 = time.perf_counter()
        log.append(dict(keyframe=kf, sweep_points=int(sweep.shape[0]), added=added, gaussians=loop.count,
                        insert_ms=round((t1 - t0) * 1e3, 2), iters_ms=round((t2 - t1) * 1e3, 2),
                        ms_per_iter=round((t2 - t1) * 1e3 / iters, 4), loss_first=round(first, 5), loss_last=round(last, 5),
                        fits=bool(fits)))
    torch.cuda.synchronize()
    total = time.perf_counter() - t_all
    out = dict(config="BASELINE.json configs[4]: synthetic stream, VGPR -> insertion -> %d iterations per keyframe" % iters,
               n_gpus=world, keyframes=keyframes, iters_per_keyframe=iters, gaussians=loop.count, total_s=round(total, 3),
               keyframes_per_s=round(keyframes / total, 3), views_per_s=round(keyframes * iters * world / total, 1),
               ms_per_iter_last=log[-1]["ms_per_iter"], insert_ms_last=log[-1]["insert_ms"], per_keyframe=log)
    if world > 1:  # replicas must have stayed identical
        chk = torch.tensor([loop.params.double().sum().item(), loop.params.double().abs().sum().item()], device=dev, dtype=torch.float64)
        lst = [torch.empty_like(chk) for _ in range(world)]
        dist.all_gather(lst, chk)
        out["replicas_identical"] = bool(all(torch.equal(x, lst[0]) for x in lst))
    return out if rank == 0 else None
