"""Allocation-free forward/backward driver over the C ABI: persistent scratch arenas, outputs and gradient
buffers, so that a mapping iteration issues no allocator traffic (the reference re-`resize_`s its three
scratch tensors from 0 bytes on every render, src/gs/rasterize_points.cu:83-90).

This is the C-ABI-level path bench.py times with device-resident inputs; the tensor-level operator that
mirrors the reference one-to-one lives in rasterizer.py.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from . import _lib
from .scene import Camera

GRAD_LAYOUT = (  # name, floats per Gaussian — one flat buffer, so a window's all-reduce is ONE collective.
    # The SH rows come first and the other 11 parameter-gradient floats behind them: the factored window exchange
    # (window.FactoredExchange) all-reduces only those 11; dL_dopacity sits next to the other three arrays the blend
    # backward accumulates into, so one memset clears all four.
    ("dL_dsh", None), ("dL_dmean3D", 3), ("dL_dscale", 3), ("dL_drot", 4), ("dL_dopacity", 1),
    ("dL_dmean2D", 3), ("dL_dconic", 4), ("dL_dcolor", 3), ("dL_dcov3D", 6),
)


def grad_layout(P: int, M: int):
    """-> ({name: (first float, floats per Gaussian)}, total floats, floats of the parameter gradients, floats of the SH
    rows). Every array starts on a 16-byte boundary (float4 accesses in geom_bwd and the exchange kernels), so for P
    not a multiple of 4 there are up to 3 unused floats behind an array."""
    at, off = {}, 0
    for name, w in GRAD_LAYOUT:
        w = int(M) * 3 if w is None else w
        at[name] = (off, w)
        off += (w * int(P) + 3) // 4 * 4
    return at, off, at["dL_dmean2D"][0], at["dL_dmean3D"][0]


class _Arena:
    """Grow-only device byte arena handed to the C ABI's allocator callback."""

    def __init__(self, device):
        self.device = device
        self.buf = torch.empty(0, dtype=torch.uint8, device=device)
        self.fn = _lib.ALLOC_FN(self._alloc)
        self.requested = 0

    def _alloc(self, _ctx, nbytes):
        self.requested = int(nbytes)
        if self.buf.numel() < nbytes:
            self.buf = torch.empty(int(nbytes * 1.25) + 256, dtype=torch.uint8, device=self.device)
        return self.buf.data_ptr()


class CapturedFrame:
    """One captured forward+backward of a RasterContext (RasterContext.capture_step)."""

    def __init__(self, ctx, graph, slots, fwd_args, keep):
        self.ctx, self.graph, self._slots, self._fwd_args, self._keep = ctx, graph, slots, fwd_args, keep

    def replay(self) -> None:
        ctx = self.ctx
        if len(ctx._unchecked) + len(self._slots) >= ctx._status.shape[0] - 1:
            ctx.check()  # status ring about to wrap (synchronises)
        self.graph.replay()
        ctx._unchecked.extend(s for s in self._slots if s not in ctx._unchecked)
        ctx._fwd_args, ctx.num_rendered = self._fwd_args, ctx._capacity


class RasterContext:
    """One (Gaussian set, image size) rasterization context on one device."""

    def __init__(self, P: int, M: int, width: int, height: int, device, sh_degree: int = 3, sync_free: bool = True,
                 grad_flat: Optional[torch.Tensor] = None, track_rows: bool = True):
        """sync_free: after a first frame has told how many (Gaussian, tile) instances the scene produces, later
        forwards go through gsr_forward_presized — the binning arena is kept 30 % larger than the largest R seen
        and the host never waits for the GPU inside a forward. `check()` (any time after a synchronisation the
        caller does anyway) tells whether every frame since the last check fitted; a frame that did not is
        redone through the synchronous path by `forward(..., redo=True)`."""
        self.lib = _lib.load()
        self.grad_numel = grad_layout(P, M)[1]
        self.P, self.M, self.W, self.H, self.D = int(P), int(M), int(width), int(height), int(sh_degree)
        self.active = self.P  # Gaussians in use: set_active() (a growing map renders a prefix of its capacity)
        self.device = torch.device(device)
        f32 = dict(dtype=torch.float32, device=self.device)
        self.out_color = torch.zeros((3, self.H, self.W), **f32)
        self.out_depth = torch.zeros((1, self.H, self.W), **f32)
        self.out_acc = torch.zeros((1, self.H, self.W), **f32)
        self.radii = torch.zeros((self.P,), dtype=torch.int32, device=self.device)
        self.geom, self.binning, self.image = _Arena(self.device), _Arena(self.device), _Arena(self.device)
        # flat gradient buffer; the leading 59 floats per Gaussian (sh, mean3D, scale, rot, opacity) are what a
        # multi-GPU window all-reduces (236 B per Gaussian at M = 16)
        at, total, param_floats, sh_floats = grad_layout(self.P, self.M)
        # (a multi-GPU window passes its symmetric gradient buffer here: window.GradExchange.flat)
        self.grad_flat = torch.zeros((total,), **f32) if grad_flat is None else grad_flat
        if self.grad_flat.numel() < total or self.grad_flat.dtype != torch.float32:
            raise ValueError("grad_flat must hold %d float32 values" % total)
        self.grads: Dict[str, torch.Tensor] = {n: self.grad_flat[off:off + w * self.P].view(self.P, w)
                                               for n, (off, w) in at.items()}
        self.param_grad_floats = param_floats
        self.sh_floats = sh_floats  # [0, sh_floats): dL_dsh; [sh_floats, param_grad_floats): the other 11 per Gaussian
        # which Gaussians' assigned gradient outputs may be non-zero (gsr_bwd_args.row_state): culled Gaussians whose
        # outputs are still zero from an earlier frame are not zero-filled again. 1 = unknown (a caller's own buffer).
        self.track_rows = bool(track_rows)
        self.row_state = torch.ones((self.P,), dtype=torch.uint8, device=self.device)
        self.num_rendered = 0
        self._fwd_args: Optional[_lib.FwdArgs] = None
        self._keep = None
        self.sync_free = bool(sync_free)
        self._capacity = 0  # instances the binning arena is carved for (0: not known yet)
        self._status = torch.zeros((16, 4), dtype=torch.int32).pin_memory()  # ring of gsr_async_status
        self._status_np = self._status.numpy()  # same memory
        self._status_head = 0
        self._unchecked = []  # ring slots written since the last check()

    # ------------------------------------------------------------------------------------------------
    def forward(self, scene: Dict[str, torch.Tensor], cam: Camera, scale_modifier: float = 1.0, debug: bool = False,
                redo: bool = False, activations: int = 0):
        """activations (_lib.ACT_* bits): scene["opacities"] / ["scales"] / ["rotations"] are GaussianModel's RAW
        parameters and the kernels apply sigmoid / exp / normalize themselves; the matching
        backward then returns gradients w.r.t. the raw parameters."""
        a = _lib.FwdArgs()
        a.activations = int(activations)
        self._activations = int(activations)
        a.P, a.D, a.M, a.width, a.height = self.active, self.D, self.M, self.W, self.H
        a.background, a.means3D, a.shs = scene["background"].data_ptr(), scene["means3D"].data_ptr(), scene["shs"].data_ptr()
        a.colors_precomp, a.opacities, a.scales = None, scene["opacities"].data_ptr(), scene["scales"].data_ptr()
        a.scale_modifier, a.rotations, a.cov3D_precomp = float(scale_modifier), scene["rotations"].data_ptr(), None
        a.viewmatrix, a.projmatrix, a.cam_pos = cam.viewmatrix.data_ptr(), cam.projmatrix.data_ptr(), cam.camera_center.data_ptr()
        a.tan_fovx, a.tan_fovy, a.prefiltered = float(cam.tanfovx), float(cam.tanfovy), 0
        a.out_color, a.out_depth, a.out_acc = self.out_color.data_ptr(), self.out_depth.data_ptr(), self.out_acc.data_ptr()
        a.radii, a.debug = self.radii.data_ptr(), int(debug)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        self._fwd_args, self._keep = a, (scene, cam)
        if self.sync_free and self._capacity > 0 and not redo:
            slot = self._status_head % self._status.shape[0]
            self._status_head += 1
            if len(self._unchecked) >= self._status.shape[0] - 1:
                self.check()  # the ring is about to wrap: look at what is pending (synchronises)
            self._unchecked.append(slot)
            need = self.lib.gsr_required_binning(self._capacity)
            if self.binning.buf.numel() < need:  # (the previous frame's backward is over: same stream)
                self.binning._alloc(None, need)
            st = self.lib.gsr_forward_presized(C.byref(a), self.geom.buf.data_ptr(), self.binning.buf.data_ptr(),
                                               self._capacity, self.image.buf.data_ptr(), stream,
                                               self._status[slot].data_ptr())
            _lib.check(st, "gsr_forward_presized")
            self.num_rendered = self._capacity  # what the matching backward carves the binning arena with
            return self.out_color, self.out_depth, self.out_acc
        n = C.c_int(0)
        st = self.lib.gsr_forward(C.byref(a), self.geom.fn, None, self.binning.fn, None, self.image.fn, None, stream,
                                  C.byref(n))
        _lib.check(st, "gsr_forward")
        self.num_rendered = n.value
        if self.sync_free:  # the frames to come get a binning arena with 30 % head-room over this R
            self._capacity = max(self._capacity, int(n.value * 1.3) + 65536)
        return self.out_color, self.out_depth, self.out_acc

    def check(self) -> bool:
        """True if every sync-free forward since the last call fitted its presized binning arena. Synchronises
        the current stream. After a False the arena is re-learned by the next forward (synchronous path)."""
        if not self._unchecked:
            return True
        torch.cuda.current_stream(self.device).synchronize()
        words = self._status_np  # numpy view of the pinned ring: no torch CPU ops between two GPU launches
        ok, worst = True, 0
        for slot in self._unchecked:
            if int(words[slot, 3]) != 0:  # the frame had more instances than the arena was carved for
                ok = False
                worst = max(worst, int(words[slot, 0]))
        if ok:
            self.last_rendered = int(words[self._unchecked[-1], 0])
        else:
            self._capacity = int(worst * 1.3) + 65536
        self._unchecked = []
        return ok

    def prepare_backward(self) -> None:
        """Clears the four gradient arrays the blend backward accumulates into (gsr_backward_prepare) on the current
        stream. Independent of the forward: issue it on a side stream beside the forward, then call
        backward(..., accumulators_zeroed=True)."""
        g = self.grads
        a = _lib.BwdArgs()
        a.P = self.active
        a.dL_dmean2D, a.dL_dconic, a.dL_dopacity = g["dL_dmean2D"].data_ptr(), g["dL_dconic"].data_ptr(), g["dL_dopacity"].data_ptr()
        a.dL_dcolor = g["dL_dcolor"].data_ptr()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self.lib.gsr_backward_prepare(C.byref(a), stream), "gsr_backward_prepare")

    def backward(self, dL_dcolor: torch.Tensor, dL_dacc: torch.Tensor, debug: bool = False,
                 accumulators_zeroed: bool = False, part: str = "all", sh_factored: bool = False,
                 gaussian_range: Optional[tuple] = None) -> Dict[str, torch.Tensor]:
        """Gradients of the last forward. The reference's caller zero-fills all nine outputs before every
        backward (rasterize_points.cu:173-181); here the library is told they are uninitialised
        (GSR_BWD_OUTPUTS_UNINITIALIZED) and writes every element itself, inside whatever region the caller is
        timing. part: "all", or one half per call — "blend" (accumulator clearing + blend backward; dL_dmean2D,
        dL_dconic, dL_dopacity, dL_dcolor are complete after it) and "geom" (the per-Gaussian kernel).
        sh_factored: dL_dsh is not written (GSR_BWD_SH_FACTORED; window.FactoredExchange rebuilds the window's sum).
        gaussian_range = (first, count): with part == "geom", only those Gaussians (a window exchange pipelines the
        per-Gaussian kernel in chunks)."""
        f = self._fwd_args
        if f is None:
            raise RuntimeError("backward() before forward()")
        g = self.grads
        a = _lib.BwdArgs()
        a.P, a.D, a.M, a.R, a.width, a.height = self.active, self.D, self.M, self.num_rendered, self.W, self.H
        a.background, a.means3D, a.shs, a.colors_precomp = f.background, f.means3D, f.shs, None
        a.scales, a.scale_modifier, a.rotations, a.cov3D_precomp = f.scales, f.scale_modifier, f.rotations, None
        a.viewmatrix, a.projmatrix, a.cam_pos = f.viewmatrix, f.projmatrix, f.cam_pos
        a.tan_fovx, a.tan_fovy, a.radii = f.tan_fovx, f.tan_fovy, self.radii.data_ptr()
        a.geom_buffer, a.binning_buffer, a.image_buffer = (self.geom.buf.data_ptr(), self.binning.buf.data_ptr(),
                                                           self.image.buf.data_ptr())
        a.dL_dpix, a.dL_dacc = dL_dcolor.data_ptr(), dL_dacc.data_ptr()
        a.dL_dmean2D, a.dL_dconic, a.dL_dopacity = g["dL_dmean2D"].data_ptr(), g["dL_dconic"].data_ptr(), g["dL_dopacity"].data_ptr()
        a.dL_dcolor, a.dL_dmean3D, a.dL_dcov3D = g["dL_dcolor"].data_ptr(), g["dL_dmean3D"].data_ptr(), g["dL_dcov3D"].data_ptr()
        a.dL_dsh, a.dL_dscale, a.dL_drot = g["dL_dsh"].data_ptr(), g["dL_dscale"].data_ptr(), g["dL_drot"].data_ptr()
        a.debug = int(debug)
        a.flags = _lib.BWD_OUTPUTS_UNINITIALIZED | (_lib.BWD_ACCUMULATORS_ZEROED if accumulators_zeroed else 0)
        a.flags |= {"all": 0, "blend": _lib.BWD_ONLY_BLEND, "geom": _lib.BWD_ONLY_GEOM}[part]
        a.flags |= _lib.BWD_SH_FACTORED if sh_factored else 0
        a.activations, a.opacities = getattr(self, "_activations", 0), f.opacities
        # (the state describes mean3D / cov3D / scale / rot and, unless factored, the SH rows: a context that mixes
        # factored and dense backwards must not skip SH rows on the strength of a factored frame)
        if self.track_rows and part in ("all", "geom"):
            if sh_factored != getattr(self, "_rows_factored", sh_factored):
                self.row_state.fill_(1)
            self._rows_factored = sh_factored
            a.row_state = self.row_state.data_ptr()
        if gaussian_range is not None:
            if part != "geom":
                raise ValueError("gaussian_range needs part == 'geom'")
            a.range_first, a.range_count = int(gaussian_range[0]), int(gaussian_range[1])
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self.lib.gsr_backward(C.byref(a), stream), "gsr_backward")
        return g

    def capture_step(self, scene: Dict[str, torch.Tensor], cam: Camera, dL_dcolor: torch.Tensor,
                     dL_dacc: torch.Tensor) -> "CapturedFrame":
        """forward(scene, cam) + backward(dL_dcolor, dL_dacc) recorded once as a CUDA graph (ours; the reference
        launches every kernel of every iteration from the host). All tensors are used by address: new parameter
        values, a new camera or new upstream gradients are written INTO them before `replay()`. Needs the sync-free
        path with a known capacity, i.e. one synchronous frame before; `check()` works as after eager frames."""
        if not (self.sync_free and self._capacity > 0):
            raise RuntimeError("capture_step: render one frame eagerly first (the binning arena must be sized)")
        self.forward(scene, cam)  # eager once: one-time kernel attributes, arena growth
        self.backward(dL_dcolor, dL_dacc)
        if not self.check():
            raise RuntimeError("capture_step: the frame does not fit the presized binning arena")
        before = len(self._unchecked)
        graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(device=self.device)
        with torch.cuda.graph(graph):
            main = torch.cuda.current_stream(self.device)
            side.wait_stream(main)
            with torch.cuda.stream(side):  # the 44 B/Gaussian of memsets the backward starts with run beside the forward
                self.prepare_backward()
            self.forward(scene, cam)
            main.wait_stream(side)
            self.backward(dL_dcolor, dL_dacc, accumulators_zeroed=True)
        slots = self._unchecked[before:]
        del self._unchecked[before:]
        return CapturedFrame(self, graph, slots, self._fwd_args, (scene, cam, dL_dcolor, dL_dacc))

    def capture(self, step_fn) -> "CapturedFrame":
        """Any sequence of this context's sync-free forwards / backwards and other stream work (e.g. a window exchange:
        window.FactoredExchange.render_step) recorded once as a CUDA graph. step_fn() must not wait for the GPU. It is
        run eagerly once first (one-time kernel attributes, arena growth), then captured."""
        if not (self.sync_free and self._capacity > 0):
            raise RuntimeError("capture: render one frame eagerly first (the binning arena must be sized)")
        step_fn()
        if not self.check():
            raise RuntimeError("capture: the frame does not fit the presized binning arena")
        before = len(self._unchecked)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            step_fn()
        slots = self._unchecked[before:]
        del self._unchecked[before:]
        return CapturedFrame(self, graph, slots, self._fwd_args, (step_fn, self._keep))

    def set_active(self, n: int) -> None:
        """Render / differentiate only the first n <= P Gaussians from now on (a map that grows inside buffers laid out
        for its capacity P: every per-Gaussian array a caller passes is then a [P, ...] array whose first n rows are in
        use; the gradient arrays keep their capacity layout and rows >= n are left alone). Frames captured before the
        change keep their old count."""
        if not 0 < int(n) <= self.P:
            raise ValueError("active count must be in (0, P]")
        self.active = int(n)

    def invalidate_grads(self) -> None:
        """Call after anything but this context's backward wrote into the gradient buffer (a fill, an in-place
        collective over it): the next backward rewrites every Gaussian's outputs, culled ones included."""
        self.row_state.fill_(1)

    def param_grads(self) -> torch.Tensor:
        """Flat view of the gradients of the optimised parameters (what a window all-reduces)."""
        return self.grad_flat[:self.param_grad_floats]

    # ------------------------------------------------------------------------------------------------
    def stage_timing(self, on: bool) -> None:
        self.lib.gsr_stage_timing_enable(int(on))

    def stage_times(self) -> Dict[str, float]:
        arr = (C.c_float * len(_lib.STAGES))()
        _lib.check(self.lib.gsr_stage_times(arr, len(_lib.STAGES)), "gsr_stage_times")
        return {n: float(arr[i]) for i, n in enumerate(_lib.STAGES)}
