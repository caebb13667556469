"""Python mirror of the reference's LibTorch operator surface, driving libgsr_b200.so through its C ABI.

Same names, argument meaning, return tuples and error behaviour as the reference, so the parity tests
read like the reference's own call sites:

  GaussianRasterizationSettings   <- include/gs/gs/rasterizer.cuh:8-20
  rasterize_gaussians_cuda        <- RasterizeGaussiansCUDA          (src/gs/rasterize_points.cu:46-130)
  rasterize_gaussians_backward_cuda <- RasterizeGaussiansBackwardCUDA (src/gs/rasterize_points.cu:132-224)
  mark_visible                    <- markVisible                     (src/gs/rasterize_points.cu:226-241)
  _RasterizeGaussians             <- torch::autograd::Function       (src/gs/rasterizer.cu:6-149)
  GaussianRasterizer              <- src/gs/rasterizer.cu:152-288

torch is plumbing only (device memory, streams, autograd bookkeeping); all arithmetic happens in the
hand-written CUDA kernels. No CPU path exists: tensors must be CUDA tensors.
"""
from __future__ import annotations

import ctypes as C
from typing import NamedTuple, Optional, Tuple

import torch

from . import _lib

NUM_CHANNELS = 3


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    camera_center: torch.Tensor
    prefiltered: bool = False


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    """Device pointer of a contiguous tensor; None for undefined / empty tensors (the reference passes
    .data<float>() of 0-element tensors, i.e. nullptr, to select the SH / scale-rotation paths)."""
    if t is None or t.numel() == 0:
        return None
    return t.data_ptr()


def _f32c(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    if t is None:
        return None
    if t.dtype != torch.float32:
        raise TypeError("rasterizer tensors must be float32")
    return t.contiguous()


class _ScratchAllocator:
    """resizeFunctional (rasterize_points.cu:36-44): a byte tensor the C ABI sizes through a callback.

    The ctypes callback closes over a one-element list, not over `self`: a bound method would put the buffer
    into a reference cycle (object -> callback -> method -> object) and hand its lifetime to Python's cyclic
    collector, i.e. hundreds of MB of scratch per render would pile up until the next collection and the
    caching allocator would answer with cudaMalloc instead of reusing the previous render's blocks."""

    def __init__(self, device: torch.device):
        cell = [torch.empty(0, dtype=torch.uint8, device=device)]

        def _alloc(_ctx, nbytes, _cell=cell, _device=device):
            _cell[0] = torch.empty(int(nbytes), dtype=torch.uint8, device=_device)
            return _cell[0].data_ptr()

        self._cell = cell
        self.fn = _lib.ALLOC_FN(_alloc)

    @property
    def tensor(self) -> torch.Tensor:
        return self._cell[0]


class RenderArena:
    """Persistent scratch for a training loop: grow-only geometry / image / binning buffers
    that successive renders reuse, and — once a first render has told how many (Gaussian, tile) instances the
    scene produces — the sync-free forward (gsr_forward_presized): the binning buffer is kept 30 % larger than
    the largest instance count seen and the host never waits for the GPU inside a render.

    Pass it as `GaussianRasterizer(settings, arena=...)`. One arena serves ONE render at a time: the buffers of
    a render are reused by the next one, so a render's backward must run before the arena renders again (use one
    arena per view when several views are rendered before a joint backward). `check()` — after any
    synchronisation the loop does anyway, e.g. reading the loss — tells whether every render since the last call
    fitted; a render that did not produced an empty image and has to be repeated."""

    def __init__(self, device, exchange=None):
        """exchange (optional): a window.FactoredExchange. The backward through this arena then writes its gradients into
        the exchange's buffer and sums them over the ranks itself (one view per rank: every rank's backward takes part),
        so the parameter gradients autograd hands out are the WINDOW's: no all-reduce in the caller's step. Before every
        step the caller puts the camera centres of the step's views (rank order) into exchange.cam_centres and sets
        `exchange_parity` to the step's parity (0 / 1 alternating)."""
        self.device = torch.device(device)
        self.exchange = exchange
        self.exchange_parity = 0
        self._bufs = {k: torch.empty(0, dtype=torch.uint8, device=self.device) for k in ("geom", "binning", "image")}
        self._fns = {k: _lib.ALLOC_FN(lambda _ctx, n, _k=k: self._grow(_k, n)) for k in self._bufs}
        self.capacity = 0
        self.param_grad_span = None  # after a backward: the contiguous span that holds the parameter gradients
        self._status = torch.zeros((16, 4), dtype=torch.int32).pin_memory()
        self._status_np = self._status.numpy()  # same memory
        self._head = 0
        self._unchecked = []
        # the backward's outputs live in the arena too: ONE persistent buffer (the gradients a backward returns are views
        # of it and are overwritten by the next backward through this arena, like the scratch buffers), which lets the
        # library skip re-zeroing the outputs of Gaussians that were culled last time as well (gsr_bwd_args.row_state)
        # and clear the four accumulated arrays on a side stream while the forward still runs
        self._grad_key, self._grad_flat, self._grad_start, self.row_state = None, None, None, None
        self._side = None
        self._accumulators_zeroed = False

    GRAD_ORDER = ("means3D", "sh", "scales", "rotations", "opacity", "means2D", "colors", "conic", "cov3D")

    def grad_buffer(self, P: int, M: int):
        """-> (flat float32 buffer, {name: first float}, {name: floats per Gaussian}); 256-byte aligned arrays, the
        optimised parameters' gradients first (param_grad_span), the four accumulated arrays adjacent."""
        width = dict(means3D=3, sh=M * 3, scales=3, rotations=4, opacity=1, means2D=3, colors=NUM_CHANNELS, conic=4, cov3D=6)
        if self.exchange is not None and self._grad_key != (P, M):
            # the exchange's symmetric buffer in engine.grad_layout's order; no zero-row tracking (sums land in the rows
            # of Gaussians this rank culled)
            from . import engine
            if self.exchange.P != P or M != 16:
                raise ValueError("the arena's exchange was built for %d Gaussians with 16 SH coefficients" % self.exchange.P)
            at = engine.grad_layout(P, M)[0]
            names = dict(means3D="dL_dmean3D", sh="dL_dsh", scales="dL_dscale", rotations="dL_drot", opacity="dL_dopacity",
                         means2D="dL_dmean2D", colors="dL_dcolor", conic="dL_dconic", cov3D="dL_dcov3D")
            self._grad_flat, self._grad_start = self.exchange.flat, {k: at[v][0] for k, v in names.items()}
            self._grad_key, self.row_state = (P, M), None
            self._accumulators_zeroed = False
        if self._grad_key != (P, M):
            start, off = {}, 0
            for k in self.GRAD_ORDER:
                start[k] = off
                off += (P * width[k] + 63) // 64 * 64
            self._grad_flat = torch.zeros((off,), dtype=torch.float32, device=self.device)
            self._grad_start, self._grad_key = start, (P, M)
            self.row_state = torch.ones((P,), dtype=torch.uint8, device=self.device)
            self._accumulators_zeroed = False
        return self._grad_flat, self._grad_start, width

    def invalidate_grads(self) -> None:
        """Call after anything but a backward through this arena wrote into the gradient buffer (e.g. an in-place
        all-reduce over param_grad_span: it leaves sums in the rows of Gaussians this rank culled): the next backward
        rewrites every Gaussian's outputs."""
        if self.row_state is not None:
            self.row_state.fill_(1)

    def prepare_backward(self, lib, P: int, M: int) -> None:
        """Called by the forward when a backward will follow: clears the four arrays the blend backward accumulates into
        (44 B per Gaussian) on a side stream, beside the forward."""
        flat, start, width = self.grad_buffer(P, M)
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device)
        main = torch.cuda.current_stream(self.device)
        self._side.wait_stream(main)
        a = _lib.BwdArgs()
        a.P = P
        base = flat.data_ptr()
        a.dL_dmean2D, a.dL_dconic = base + 4 * start["means2D"], base + 4 * start["conic"]
        a.dL_dopacity, a.dL_dcolor = base + 4 * start["opacity"], base + 4 * start["colors"]
        with torch.cuda.stream(self._side):
            _lib.check(lib.gsr_backward_prepare(C.byref(a), self._side.cuda_stream), "gsr_backward_prepare")
        self._accumulators_zeroed = True

    def _grow(self, kind: str, nbytes: int) -> int:
        if self._bufs[kind].numel() < nbytes:
            self._bufs[kind] = torch.empty(int(nbytes * 1.25) + 256, dtype=torch.uint8, device=self.device)
        return self._bufs[kind].data_ptr()

    def buffer(self, kind: str) -> torch.Tensor:
        return self._bufs[kind]

    def reserve(self, lib, P: int, W: int, H: int) -> None:
        key = (P, W, H, self.capacity)
        if key != getattr(self, "_reserved", None):
            self._grow("geom", lib.gsr_required_geom(P))
            self._grow("image", lib.gsr_required_image(W, H))
            self._grow("binning", lib.gsr_required_binning(self.capacity))
            self._reserved = key

    def learn(self, num_rendered: int) -> None:
        self.capacity = max(self.capacity, int(num_rendered * 1.3) + 65536)

    def next_status(self) -> int:
        if len(self._unchecked) >= self._status.shape[0] - 1:
            self.check()
        slot = self._head % self._status.shape[0]
        self._head += 1
        self._unchecked.append(slot)
        return self._status[slot].data_ptr()

    def check(self, sync: bool = True, keep_last: int = 0) -> bool:
        """sync=False: the caller has already waited for the renders in question (e.g. on an event behind a step whose
        successor is already in flight); keep_last: the newest `keep_last` renders are still running and stay pending."""
        n = len(self._unchecked) - max(0, int(keep_last))
        if n <= 0:
            return True
        if sync:
            torch.cuda.current_stream(self.device).synchronize()
        # (numpy view of the pinned words: this runs once per step between two GPU launches, and a handful of torch
        # CPU ops here cost more than the camera upload)
        words = self._status_np
        ok, worst = True, 0
        for slot in self._unchecked[:n]:
            if int(words[slot, 3]) != 0:  # the frame had more instances than the arena was carved for
                ok = False
                worst = max(worst, int(words[slot, 0]))
        del self._unchecked[:n]
        if not ok:
            self.capacity = int(worst * 1.3) + 65536
        return ok


class CapturedStep:
    """CUDA-graph replay of one optimisation step built on `GaussianRasterizer(settings, arena=RenderArena)`
    (ours; the reference's mapping loop launches every kernel of every iteration from the host,
    src/liw/lioOptimization.cpp:1693-1760).

    `step_fn()` is the step as the caller would write it eagerly — host-to-device copies from PINNED staging
    tensors, the render, the loss, `loss.backward()`, the device-to-host copy of the loss — with three rules:
    no host synchronisation inside, every tensor it reads from the outside keeps its address (write new camera /
    image data INTO the staging tensors before `replay()`), the arena already knows its capacity (one eager
    render did `arena.learn`), and nothing that keeps an autograd graph alive (the loss, an output that
    requires grad) survives a call -- keep `.detach()`ed views instead. It runs `warmup` times eagerly on a side stream, is captured once, and
    `replay()` then launches the whole step — ~20 kernels and copies — with one host call. Tensors created
    inside `step_fn` (outputs, `.grad`s, the loss) live in the graph's memory pool and are overwritten by every
    replay; `result` is whatever `step_fn` returned at capture time.

    After `replay()` and whatever synchronisation the loop does anyway, `arena.check()` tells whether the frame
    fitted the presized binning buffer, exactly as on the eager path."""

    def __init__(self, step_fn, arena: RenderArena, warmup: int = 3):
        if arena.capacity <= 0:
            raise ValueError("CapturedStep: the arena has no capacity yet; render once eagerly first")
        self.arena = arena
        dev = arena.device
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                step_fn()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        if not arena.check():
            raise RuntimeError("CapturedStep: a warm-up render did not fit the arena")
        before = len(arena._unchecked)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.result = step_fn()
        self._slots = arena._unchecked[before:]  # status words the captured renders report into
        del arena._unchecked[before:]

    def replay(self):
        self.graph.replay()
        self.arena._unchecked.extend(self._slots)
        return self.result


def _stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def rasterize_gaussians_cuda(background, means3D, colors, opacity, scales, rotations, scale_modifier, cov3D_precomp,
                             viewmatrix, projmatrix, tan_fovx, tan_fovy, image_height, image_width, sh, degree,
                             campos, prefiltered, debug, arena: Optional[RenderArena] = None):
    """-> (num_rendered, out_color[3,H,W], out_depth[1,H,W], out_acc[1,H,W], radii[P], geomBuffer,
    binningBuffer, imgBuffer), exactly the reference's tuple. With an `arena` the three buffers are the arena's
    and, on the sync-free path, num_rendered is the arena's capacity (what the backward must be given)."""
    if means3D.dim() != 2 or means3D.size(1) != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")
    if not means3D.is_cuda:
        raise RuntimeError("gs_livm_b200 has no CPU path: tensors must live on a CUDA device")
    lib = _lib.load()
    dev = means3D.device
    P, H, W = means3D.size(0), int(image_height), int(image_width)

    # the reference fills these with zeros (rasterize_points.cu:78-81); our kernels write every pixel of every
    # tile and every Gaussian's radius, so only the P == 0 case (no kernel runs) needs the fill
    fresh = torch.empty if P != 0 else torch.zeros
    planes = fresh((NUM_CHANNELS + 2, H, W), dtype=torch.float32, device=dev)  # one allocation, three views
    out_color, out_depth, out_acc = planes[:NUM_CHANNELS], planes[NUM_CHANNELS:NUM_CHANNELS + 1], planes[NUM_CHANNELS + 1:]
    radii = fresh((P,), dtype=torch.int32, device=dev)
    if arena is None:
        geom, binning, img = _ScratchAllocator(dev), _ScratchAllocator(dev), _ScratchAllocator(dev)

    rendered = 0
    if P != 0:
        M = sh.size(1) if (sh is not None and sh.numel() != 0) else 0
        keep = [_f32c(t) for t in (background, means3D, sh, colors, opacity, scales, rotations, cov3D_precomp,
                                   viewmatrix, projmatrix, campos)]
        bg_c, m3_c, sh_c, col_c, op_c, sc_c, rot_c, cov_c, view_c, proj_c, cam_c = keep
        a = _lib.FwdArgs()
        a.P, a.D, a.M, a.width, a.height = P, int(degree), int(M), W, H
        a.background, a.means3D, a.shs, a.colors_precomp = _ptr(bg_c), _ptr(m3_c), _ptr(sh_c), _ptr(col_c)
        a.opacities, a.scales, a.scale_modifier = _ptr(op_c), _ptr(sc_c), float(scale_modifier)
        a.rotations, a.cov3D_precomp = _ptr(rot_c), _ptr(cov_c)
        a.viewmatrix, a.projmatrix, a.cam_pos = _ptr(view_c), _ptr(proj_c), _ptr(cam_c)
        a.tan_fovx, a.tan_fovy, a.prefiltered = float(tan_fovx), float(tan_fovy), int(bool(prefiltered))
        a.out_color, a.out_depth, a.out_acc = out_color.data_ptr(), out_depth.data_ptr(), out_acc.data_ptr()
        a.radii, a.debug = radii.data_ptr(), int(bool(debug))
        n = C.c_int(0)
        with torch.cuda.device(dev):
            if arena is None:
                st = lib.gsr_forward(C.byref(a), geom.fn, None, binning.fn, None, img.fn, None, _stream_ptr(dev),
                                     C.byref(n))
                _lib.check(st, "gsr_forward")
                rendered = n.value
            elif arena.capacity > 0:
                arena.reserve(lib, P, W, H)
                st = lib.gsr_forward_presized(C.byref(a), arena.buffer("geom").data_ptr(),
                                              arena.buffer("binning").data_ptr(), arena.capacity,
                                              arena.buffer("image").data_ptr(), _stream_ptr(dev), arena.next_status())
                _lib.check(st, "gsr_forward_presized")
                rendered = arena.capacity
            else:
                st = lib.gsr_forward(C.byref(a), arena._fns["geom"], None, arena._fns["binning"], None,
                                     arena._fns["image"], None, _stream_ptr(dev), C.byref(n))
                _lib.check(st, "gsr_forward")
                rendered = n.value
                arena.learn(rendered)
    if arena is not None:
        return (rendered, out_color, out_depth, out_acc, radii, arena.buffer("geom"), arena.buffer("binning"),
                arena.buffer("image"))
    return rendered, out_color, out_depth, out_acc, radii, geom.tensor, binning.tensor, img.tensor


def rasterize_gaussians_backward_cuda(background, means3D, radii, colors, scales, rotations, scale_modifier,
                                      cov3D_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy, dL_dout_color,
                                      dL_dout_acc, sh, degree, campos, geomBuffer, R, binningBuffer, imageBuffer,
                                      debug, arena: Optional["RenderArena"] = None):
    """-> (dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations)."""
    lib = _lib.load()
    dev = means3D.device
    P, H, W = means3D.size(0), dL_dout_color.size(1), dL_dout_color.size(2)
    M = sh.size(1) if (sh is not None and sh.numel() != 0) else 0
    opts = dict(dtype=torch.float32, device=dev)
    # The reference zero-fills all nine tensors here (rasterize_points.cu:173-181). The library is told the
    # outputs are uninitialised instead (GSR_BWD_OUTPUTS_UNINITIALIZED) and writes every element of what it
    # computes; only what it does not touch on this argument combination is zeroed here.
    have_cov = cov3D_precomp is not None and cov3D_precomp.numel() != 0
    # one allocation, nine 256-byte aligned views; the gradients of the optimised parameters (mean3D, sh, scale,
    # rotation, opacity) come first, so that a view-sharded window can sum them with ONE collective over a contiguous
    # span (arena.param_grad_span) instead of one per tensor
    order = RenderArena.GRAD_ORDER
    width = dict(means3D=3, sh=M * 3, scales=3, rotations=4, opacity=1, means2D=3, colors=NUM_CHANNELS, conic=4, cov3D=6)
    if arena is not None and P != 0:
        flat, start, _ = arena.grad_buffer(P, M)  # persistent: overwritten by the next backward through this arena
    else:
        start, off = {}, 0
        for k in order:
            start[k] = off
            off += (P * width[k] + 63) // 64 * 64
        flat = (torch.empty if P != 0 else torch.zeros)((off,), **opts)
    view = {k: flat[start[k]:start[k] + P * width[k]] for k in order}
    dL_dmeans3D, dL_dmeans2D, dL_dcolors = view["means3D"].view(P, 3), view["means2D"].view(P, 3), view["colors"].view(P, NUM_CHANNELS)
    dL_dconic, dL_dopacity, dL_dcov3D = view["conic"].view(P, 2, 2), view["opacity"].view(P, 1), view["cov3D"].view(P, 6)
    dL_dsh, dL_dscales, dL_drotations = view["sh"].view(P, M, 3), view["scales"].view(P, 3), view["rotations"].view(P, 4)
    if arena is not None:
        arena.param_grad_span = flat[:start["means2D"]]  # (with an exchange: engine.grad_layout's order, sh first)
    if have_cov and P != 0:
        dL_dscales.zero_()
        dL_drotations.zero_()

    if P != 0:
        keep = [_f32c(t) for t in (background, means3D, sh, colors, scales, rotations, cov3D_precomp, viewmatrix,
                                   projmatrix, campos, dL_dout_color, dL_dout_acc)]
        bg_c, m3_c, sh_c, col_c, sc_c, rot_c, cov_c, view_c, proj_c, cam_c, dpix_c, dacc_c = keep
        radii_c = radii.contiguous()
        a = _lib.BwdArgs()
        a.P, a.D, a.M, a.R, a.width, a.height = P, int(degree), int(M), int(R), W, H
        a.background, a.means3D, a.shs, a.colors_precomp = _ptr(bg_c), _ptr(m3_c), _ptr(sh_c), _ptr(col_c)
        a.scales, a.scale_modifier, a.rotations, a.cov3D_precomp = _ptr(sc_c), float(scale_modifier), _ptr(rot_c), _ptr(cov_c)
        a.viewmatrix, a.projmatrix, a.cam_pos = _ptr(view_c), _ptr(proj_c), _ptr(cam_c)
        a.tan_fovx, a.tan_fovy = float(tan_fovx), float(tan_fovy)
        a.radii = _ptr(radii_c)
        a.geom_buffer, a.binning_buffer, a.image_buffer = geomBuffer.data_ptr(), binningBuffer.data_ptr(), imageBuffer.data_ptr()
        a.dL_dpix, a.dL_dacc = _ptr(dpix_c), _ptr(dacc_c)
        a.dL_dmean2D, a.dL_dconic, a.dL_dopacity = dL_dmeans2D.data_ptr(), dL_dconic.data_ptr(), dL_dopacity.data_ptr()
        a.dL_dcolor, a.dL_dmean3D, a.dL_dcov3D = dL_dcolors.data_ptr(), dL_dmeans3D.data_ptr(), dL_dcov3D.data_ptr()
        a.dL_dsh, a.dL_dscale, a.dL_drot = _ptr(dL_dsh), dL_dscales.data_ptr(), dL_drotations.data_ptr()
        a.debug, a.flags = int(bool(debug)), _lib.BWD_OUTPUTS_UNINITIALIZED
        if arena is not None:
            a.row_state = arena.row_state.data_ptr() if arena.row_state is not None else None
            if arena._accumulators_zeroed:  # the forward cleared them on the arena's side stream
                torch.cuda.current_stream(dev).wait_stream(arena._side)
                a.flags |= _lib.BWD_ACCUMULATORS_ZEROED
                arena._accumulators_zeroed = False
        ex = arena.exchange if arena is not None else None
        with torch.cuda.device(dev):
            if ex is None:
                _lib.check(lib.gsr_backward(C.byref(a), _stream_ptr(dev)), "gsr_backward")
            else:
                # the backward in halves around the window's factored exchange (window.FactoredExchange.exchange)
                if M != 16 or have_cov:
                    raise ValueError("the window exchange needs SH colours (M = 16) and scale / rotation inputs")
                base_flags = a.flags
                a.flags = base_flags | _lib.BWD_ONLY_BLEND
                _lib.check(lib.gsr_backward(C.byref(a), _stream_ptr(dev)), "gsr_backward")
                blend_done = torch.cuda.Event()
                blend_done.record(torch.cuda.current_stream(dev))

                def geom_half():
                    a.flags = (base_flags | _lib.BWD_ONLY_GEOM | _lib.BWD_SH_FACTORED) & ~_lib.BWD_ACCUMULATORS_ZEROED
                    _lib.check(lib.gsr_backward(C.byref(a), _stream_ptr(dev)), "gsr_backward")

                class _Ctx:  # what FactoredExchange reads of an engine.RasterContext
                    pass
                cx = _Ctx()
                cx.geom, cx.radii, cx.D = _Ctx(), radii_c, int(degree)
                cx.geom.buf = geomBuffer
                cx.grads = {"dL_dcolor": dL_dcolors, "dL_dsh": dL_dsh}
                ex.exchange(cx, m3_c, int(arena.exchange_parity), blend_done, geom_half)
    return dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations


def mark_visible(means3D: torch.Tensor, viewmatrix: torch.Tensor, projmatrix: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    P = means3D.size(0)
    present = torch.zeros((P,), dtype=torch.bool, device=means3D.device)
    if P != 0:
        m, v, p = _f32c(means3D), _f32c(viewmatrix), _f32c(projmatrix)
        with torch.cuda.device(means3D.device):
            st = lib.gsr_mark_visible(P, m.data_ptr(), v.data_ptr(), p.data_ptr(), present.data_ptr(),
                                      _stream_ptr(means3D.device))
        _lib.check(st, "gsr_mark_visible")
    return present


class _RasterizeGaussians(torch.autograd.Function):
    """Reference: src/gs/rasterizer.cu:6-149. Returns (color, radii, depth, acc); the backward consumes only
    the colour and silhouette gradients — the depth gradient is dropped exactly like the reference does
    (rasterizer.cu:77-80,117-118)."""

    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                raster_settings: GaussianRasterizationSettings, arena: Optional[RenderArena] = None):
        rs = raster_settings
        (num_rendered, color, out_depth, out_acc, radii, geomBuffer, binningBuffer, imgBuffer) = \
            rasterize_gaussians_cuda(rs.bg, means3D, colors_precomp, opacities, scales, rotations, rs.scale_modifier,
                                     cov3Ds_precomp, rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy,
                                     rs.image_height, rs.image_width, sh, rs.sh_degree, rs.camera_center,
                                     rs.prefiltered, False, arena)
        if arena is not None and means3D.size(0) != 0 and any(ctx.needs_input_grad):
            arena.prepare_backward(_lib.load(), means3D.size(0), sh.size(1) if sh.numel() != 0 else 0)
        ctx.raster_settings = rs
        ctx.arena = arena
        ctx.num_rendered = num_rendered
        ctx.save_for_backward(colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geomBuffer,
                              binningBuffer, imgBuffer)
        ctx.mark_non_differentiable(radii)
        return color, radii, out_depth, out_acc

    @staticmethod
    def backward(ctx, grad_out_color, _grad_radii, _grad_out_depth, grad_acc):
        rs = ctx.raster_settings
        (colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geomBuffer, binningBuffer,
         imgBuffer) = ctx.saved_tensors
        if grad_out_color is None:
            grad_out_color = torch.zeros((NUM_CHANNELS, rs.image_height, rs.image_width), dtype=torch.float32,
                                         device=means3D.device)
        if grad_acc is None:
            grad_acc = torch.zeros((1, rs.image_height, rs.image_width), dtype=torch.float32, device=means3D.device)
        (grad_means2D, grad_colors_precomp, grad_opacities, grad_means3D, grad_cov3Ds_precomp, grad_sh, grad_scales,
         grad_rotations) = rasterize_gaussians_backward_cuda(
            rs.bg, means3D, radii, colors_precomp, scales, rotations, rs.scale_modifier, cov3Ds_precomp,
            rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, grad_out_color, grad_acc, sh, rs.sh_degree,
            rs.camera_center, geomBuffer, ctx.num_rendered, binningBuffer, imgBuffer, False, ctx.arena)
        return (grad_means3D, grad_means2D, grad_sh, grad_colors_precomp, grad_opacities, grad_scales, grad_rotations,
                grad_cov3Ds_precomp, None, None)


class _FusedImageLoss(torch.autograd.Function):
    """(1 - lambda) mean|image - target| + lambda (1 - mean SSIM) as ONE differentiable op over libgsr_b200's fused loss
    kernels (gsr_image_loss; include/gs/gs/loss_utils.cuh:11-70, src/liw/lioOptimization.cpp:1729-1734): the forward
    already produces dL/dimage, the backward scales it."""

    @staticmethod
    def forward(ctx, image, target, fused):
        loss, grad = fused(image, target)
        ctx.save_for_backward(grad)
        return loss[0].reshape(())

    @staticmethod
    def backward(ctx, grad_out):
        (grad,) = ctx.saved_tensors
        return grad * grad_out, None, None


def fused_image_loss(image: torch.Tensor, target: torch.Tensor, fused) -> torch.Tensor:
    """`fused`: a train.ImageLoss for the image size (it owns the scratch and the gradient buffer: one loss in flight
    per instance). lambda_dssim = 0 gives the plain L1 loss in one pass."""
    return _FusedImageLoss.apply(image, target, fused)


class GaussianRasterizer(torch.nn.Module):
    """Reference: GaussianRasterizer (include/gs/gs/rasterizer.cuh:52-80, src/gs/rasterizer.cu:152-288)."""

    def __init__(self, raster_settings: GaussianRasterizationSettings, arena: Optional[RenderArena] = None):
        """arena (ours, optional): persistent scratch + sync-free forward for training loops, see RenderArena."""
        super().__init__()
        self.raster_settings = raster_settings
        self.arena = arena

    def mark_visible(self, positions: torch.Tensor) -> torch.Tensor:
        with torch.no_grad():
            return mark_visible(positions, self.raster_settings.viewmatrix, self.raster_settings.projmatrix)

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
        if (shs is None) == (colors_precomp is None):
            raise ValueError("Please provide exactly one of either SHs or precomputed colors!")
        if ((scales is not None or rotations is not None) and cov3D_precomp is not None) or \
                (scales is None and rotations is None and cov3D_precomp is None):
            raise ValueError("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")
        dev = means3D.device
        empty = torch.empty(0, dtype=torch.float32, device=dev)
        shs = empty if shs is None else shs
        colors_precomp = empty if colors_precomp is None else colors_precomp
        scales = empty if scales is None else scales
        rotations = empty if rotations is None else rotations
        cov3D_precomp = empty if cov3D_precomp is None else cov3D_precomp
        rs = self.raster_settings
        if not (rs.bg.device == dev and rs.viewmatrix.device == dev and rs.projmatrix.device == dev
                and rs.camera_center.device == dev):  # reference: .to(torch::kCUDA) on every call (rasterizer.cu:219-241)
            rs = rs._replace(bg=rs.bg.to(dev), viewmatrix=rs.viewmatrix.to(dev), projmatrix=rs.projmatrix.to(dev),
                             camera_center=rs.camera_center.to(dev))
        return _RasterizeGaussians.apply(means3D, means2D, shs, colors_precomp, opacities, scales, rotations,
                                         cov3D_precomp, rs, self.arena)
