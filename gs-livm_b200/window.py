"""View-sharded keyframe window (multi-GPU), SURVEY.md section 8e.

GS-LIVM's mapping iteration renders several keyframes of the sliding window against the SAME parameter
snapshot, sums their losses and takes one optimiser step (src/liw/lioOptimization.cpp:1691-1736,
1822-1832). The renders are independent, so the window shards by view: Gaussians (and optimiser state)
are replicated on every GPU, each rank renders its share of the views forward+backward with the
single-GPU kernels, and the per-Gaussian parameter gradients (59 floats = 236 B per Gaussian at SH degree
3) are summed over the ranks before the identical update is applied everywhere. There is no other cross-rank
traffic on the data path.

FactoredExchange (default, csrc/window_factored.cu): the SH rows — 81 % of those bytes — are rank one per view, so the
ranks exchange 3 dRGB floats per rendered Gaussian and view (pushed into every rank's memory under geom_bwd), all-reduce
only the other 11 floats for the Gaussians some view rendered, and rebuild the SH rows locally; the cross-rank flags
live inside the kernels, so a whole step is one CUDA graph. GradExchange (round 1, csrc/window_allreduce.cu): dense
all-reduce of the flat buffer by the library's NVSwitch-multicast / peer-to-peer kernels, NCCL as last resort; kept as
the fallback where no peer-mapped memory exists. Both need torch's symmetric memory for the peer mappings; the gloo /
CPU tests exercise the host logic with NCCL-style all-reduce only.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence

import torch
import torch.distributed as dist


def views_for_rank(n_views: int, rank: int, world_size: int) -> List[int]:
    """Round-robin assignment of window views to ranks (view v -> rank v % world_size)."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    return [v for v in range(n_views) if v % world_size == rank]


def allreduce_sum_(flat: torch.Tensor, group=None) -> torch.Tensor:
    """Sum (not mean: the reference sums the view losses, lioOptimization.cpp:1823-1825) in place."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


class GradExchange:
    """Flat FP32 gradient buffer of one rank + the all-reduce (sum) of its first `numel_reduce` elements.

    mode "multimem": NVSwitch multicast (gsr_window_allreduce_multimem); "p2p": two-shot over peer loads
    (gsr_window_allreduce_p2p); "nccl": torch.distributed.all_reduce. `prefer` forces one of them ("auto": the
    best the system offers). Every rank ends up with the same bits in all modes (each slice is summed once, in
    one place, then distributed)."""

    def __init__(self, numel_total: int, numel_reduce: int, device, group=None, prefer: str = "auto",
                 sparse_rows: Optional[tuple] = None):
        """sparse_rows = (first_float, n_rows, floats_per_row): floats [first_float, first_float + n_rows *
        floats_per_row) of the buffer are one row per Gaussian (dL_dsh) that is all zero on every rank whose view did
        not render the Gaussian; `allreduce_(radii=...)` then exchanges only the rows rendered by some view."""
        self.group = group
        self.n = int(numel_reduce)
        self.rows = None
        if sparse_rows is not None:
            first, n_rows, fpr = (int(v) for v in sparse_rows)
            if first % 4 == 0 and fpr % 4 == 0 and fpr > 0 and first + n_rows * fpr <= self.n:
                self.rows = (first, n_rows, fpr)
        n_words = (self.rows[1] + 31) // 32 if self.rows else 0
        words_at = (int(numel_total) + 3) // 4 * 4  # visibility words live behind the gradients, 16-byte aligned
        self.device = torch.device(device)
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        self.mode, self.why, self.hdl, self.flat = "nccl", "", None, None
        if self.world > 1 and self.device.type == "cuda" and self.n % 4 == 0 and prefer in ("auto", "multimem", "p2p"):
            try:
                import torch.distributed._symmetric_memory as symm_mem
                grp = group if group is not None else dist.group.WORLD
                flat = symm_mem.empty(words_at + n_words, dtype=torch.float32, device=self.device)
                hdl = symm_mem.rendezvous(flat, grp)
                flat.zero_()
                self.flat, self.hdl = flat, hdl
                mc = int(getattr(hdl, "multicast_ptr", 0) or 0)
                self.mode = "multimem" if (mc != 0 and prefer != "p2p") else "p2p"
                if prefer == "multimem" and mc == 0:
                    self.why = "no multicast support: using p2p"
            except Exception as e:  # no peer access / symmetric memory unavailable
                self.why = "symmetric memory unavailable (%s)" % (repr(e)[:200],)
                self.flat, self.hdl = None, None
        if self.world > 1 and self.device.type == "cuda" and prefer in ("auto", "multimem", "p2p"):
            # the ranks must agree: a rank that fell back on its own would meet its peers in different collectives
            ok = torch.tensor([1 if self.flat is not None else 0], dtype=torch.int32, device=self.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if int(ok.item()) == 0 and self.flat is not None:
                self.flat, self.hdl, self.mode = None, None, "nccl"
                self.why = "symmetric memory unavailable on another rank"
        if self.flat is None:
            self.mode = "nccl"
            self.flat = torch.zeros(int(numel_total), dtype=torch.float32, device=self.device)
        if self.mode != "nccl":
            from . import _lib
            self._lib = _lib
            self.lib = _lib.load()
            import ctypes as C
            off = int(getattr(self.hdl, "offset", 0) or 0)
            self._mc = int(getattr(self.hdl, "multicast_ptr", 0) or 0) + (off if self.mode == "multimem" else 0)
            ptrs = [int(p) + off for p in self.hdl.buffer_ptrs]
            self._peers = (C.c_void_p * self.world)(*ptrs)
            if ptrs[self.rank] != self.flat.data_ptr():
                raise RuntimeError("symmetric buffer pointer mismatch")
            if self.rows is not None:
                self._my_words = self.flat[words_at:words_at + n_words].view(torch.int32)
                self._peer_words = (C.c_void_p * self.world)(*[p + 4 * words_at for p in ptrs])
                self._union = torch.zeros(n_words, dtype=torch.int32, device=self.device)
                self._rows_arg = _lib.WindowRows(self._union.data_ptr(), self.rows[0], self.rows[1], self.rows[2])
            if prefer == "auto":
                self._calibrate()

    def _calibrate(self) -> None:
        """auto: time what the system offers on the real buffer (contents are zeros at this point) and keep the
        fastest; the ranks agree on the choice through a MAX all-reduce of the timings."""
        modes = (["multimem"] if self._mc else []) + ["p2p", "nccl"]
        times = []
        for m in modes:
            self.mode = m
            for _ in range(2):
                self.allreduce_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                self.allreduce_()
            e1.record()
            torch.cuda.synchronize(self.device)
            times.append(e0.elapsed_time(e1) / 3)
        t = torch.tensor(times, dtype=torch.float64, device=self.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        self.calibration_ms = dict(zip(modes, (float(v) for v in t.tolist())))
        # the timings are of the DENSE exchange; only the peer-memory kernels can skip the rows no view rendered
        # (typically a quarter to a third of the traffic), so with sparse rows they are credited 20 %
        score = [v * (0.8 if (self.rows is not None and m != "nccl") else 1.0) for m, v in zip(modes, t.tolist())]
        self.mode = modes[min(range(len(modes)), key=lambda i: score[i])]

    def allreduce_(self, radii: Optional[torch.Tensor] = None) -> torch.Tensor:
        """radii (int32 [n_rows], this rank's view: > 0 = rendered): exchange the sparse rows only where some rank
        rendered the Gaussian (needs sparse_rows at construction and one of the peer-memory modes)."""
        if self.world <= 1:
            return self.flat
        if self.mode == "nccl":
            dist.all_reduce(self.flat[:self.n], op=dist.ReduceOp.SUM, group=self.group)
            return self.flat
        import ctypes as C
        stream = torch.cuda.current_stream(self.device).cuda_stream
        rows = None
        if radii is not None and self.rows is not None:
            if radii.dtype != torch.int32 or radii.numel() != self.rows[1] or not radii.is_contiguous():
                raise ValueError("radii must be a contiguous int32 tensor with one entry per sparse row")
            st = self.lib.gsr_window_visibility_words(radii.data_ptr(), self.rows[1], self._my_words.data_ptr(), stream)
            self._lib.check(st, "gsr_window_visibility_words")
            rows = C.byref(self._rows_arg)
        self.hdl.barrier(channel=0)  # every rank's gradients (and visibility words) are complete
        if rows is not None:
            st = self.lib.gsr_window_union_words(self._peer_words, self._union.numel(), self.world,
                                                 self._union.data_ptr(), stream)
            self._lib.check(st, "gsr_window_union_words")
        if self.mode == "multimem":
            st = self.lib.gsr_window_allreduce_multimem_rows(self._mc, self.n, self.rank, self.world, rows, stream)
            self._lib.check(st, "gsr_window_allreduce_multimem")
        else:
            st = self.lib.gsr_window_allreduce_p2p_rows(self._peers, self.n, self.rank, self.world, 1, rows, stream)
            self._lib.check(st, "gsr_window_allreduce_p2p")
            self.hdl.barrier(channel=1)  # every slice is reduced
            st = self.lib.gsr_window_allreduce_p2p_rows(self._peers, self.n, self.rank, self.world, 2, rows, stream)
            self._lib.check(st, "gsr_window_allreduce_p2p")
        self.hdl.barrier(channel=2)  # every rank holds the sum; buffers may be rewritten
        return self.flat


class FactoredExchange:
    """The window's gradient exchange with the SH rows factored (csrc/window_factored.cu).

    Per view the 48-float dL_dsh row of a Gaussian is basis(dir_v) (x) dRGB_v, so the ranks exchange the 3 dRGB floats
    of the Gaussians their view rendered (pushed into every rank's slot of that view as soon as the blend backward is
    done), all-reduce only the other 11 parameter-gradient floats — and only for the Gaussians some view of the step
    rendered — and rebuild  dL_dsh = sum_v basis(dir_v) (x) dRGB_v  locally: same order on every rank, hence identical
    bits. The push and the expansion are pipelined over `n_chunks` ranges of Gaussians (the expansion of a range starts
    as soon as its dRGB has arrived from every rank); the per-Gaussian backward, which writes no SH rows at all, and
    the all-reduce behind it run once over all Gaussians. The cross-rank signalling lives inside
    libgsr_b200's kernels (flag words in the symmetric allocation), so `render_step` can be captured in ONE CUDA graph.
    One view per rank and step; `flat` is the gradient buffer an engine.RasterContext must be built on
    (`grad_flat=exchange.flat`).

    world == 1 works without any peer memory (the slots are local): that is how the single-GPU tests check the
    factorisation against the rows geom_bwd writes."""

    def __init__(self, P: int, M: int, device, group=None, reduce_mode: str = "auto", push_mode: str = "auto",
                 n_views: Optional[int] = None, sparse_reduce: bool = True, n_chunks: int = 1, track_zero_rows: bool = False,
                 expand_after_reduce: bool = False):
        """n_views: gather slots = views summed per step (default: one per rank; more only make sense where the caller
        pushes several views itself, as the single-GPU tests do). sparse_reduce: the 11-float all-reduce skips the
        Gaussians no view rendered. n_chunks: ranges of Gaussians the push -> expansion pipeline works on. (Cutting the
        per-Gaussian backward and the all-reduce into the same chunks as well was measured SLOWER at N = 2: 623 us per
        step with 1 chunk, 697 with 4, 804 with 8 -- every chunk boundary drains geom_bwd's CTAs, and a chunk's reducing
        kernel holds SM slots while it waits for the slowest rank.)"""
        import ctypes as C
        from . import _lib, engine
        self._lib, self.lib, self._C = _lib, _lib.load(), C
        self.group = group
        self.P, self.M = int(P), int(M)
        if self.M != 16:
            raise ValueError("FactoredExchange needs M == 16 SH coefficients (degree-3 storage)")
        self.device = torch.device(device)
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        self.n_views = self.world if n_views is None else int(n_views)
        if self.world > 16 or not (self.world <= self.n_views <= 16):
            raise ValueError("at most 16 ranks and 16 views, at least one view per rank")
        # chunk boundaries at multiples of 128 Gaussians (whole visibility words per warp of the push kernel)
        n_chunks = max(1, min(int(n_chunks), 32, (self.P + 127) // 128))
        per = ((self.P + n_chunks - 1) // n_chunks + 127) // 128 * 128
        self.chunks = [(f, min(per, self.P - f)) for f in range(0, self.P, per)]
        self._chunk_per, self.active = per, self.P
        at, self.grad_floats, param_floats, self.sh_floats = engine.grad_layout(self.P, self.M)
        self._rest_first = (C.c_size_t * 4)(*[at[n][0] for n in ("dL_dmean3D", "dL_dscale", "dL_drot", "dL_dopacity")])
        self.rest_floats = param_floats - self.sh_floats  # mean3D, scale, rotation, opacity: 11 floats per Gaussian
        self.slot_floats = int(self.lib.gsr_window_slot_floats(self.P))
        self.vis_words = ((self.P + 31) // 32 + 3) // 4 * 4
        r32 = lambda n: (n + 31) // 32 * 32  # noqa: E731
        self._pad_at = r32(self.grad_floats)
        self._slots_at = self._pad_at + r32(int(self.lib.gsr_window_pad_words()))
        self._slot_stride = r32(self.slot_floats + self.vis_words)
        total = self._slots_at + 2 * self.n_views * self._slot_stride  # two parities x one slot per view
        self.hdl, mc = None, 0
        if self.world > 1:
            err, ptrs = "", []
            try:
                import torch.distributed._symmetric_memory as symm_mem
                grp = group if group is not None else dist.group.WORLD
                self.buf = symm_mem.empty(total, dtype=torch.float32, device=self.device)
                self.hdl = symm_mem.rendezvous(self.buf, grp)
                off = int(getattr(self.hdl, "offset", 0) or 0)
                ptrs = [int(p) + off for p in self.hdl.buffer_ptrs]
                if ptrs[self.rank] != self.buf.data_ptr():
                    err = "symmetric buffer pointer mismatch"
                mc = int(getattr(self.hdl, "multicast_ptr", 0) or 0)
                mc = mc + off if mc else 0
            except Exception as e:  # no peer access / symmetric memory unavailable
                err = "symmetric memory unavailable (%s)" % (repr(e)[:200],)
            # every rank raises or none does (a rank that gave up alone would leave its peers inside a collective)
            ok = torch.tensor([0 if err else 1], dtype=torch.int32, device=self.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if int(ok.item()) == 0:
                raise RuntimeError("FactoredExchange: " + (err or "symmetric memory unavailable on another rank"))
        else:
            self.buf = torch.empty(total, dtype=torch.float32, device=self.device)
            ptrs = [self.buf.data_ptr()]
        self.buf.zero_()
        torch.cuda.synchronize(self.device)
        if self.world > 1:
            dist.barrier(group=group)  # every rank's flag words and epoch are zero before anybody signals
        self._ptrs, self._mc = ptrs, mc
        self.flat = self.buf[:self.grad_floats]
        self.reduce_mode = reduce_mode if reduce_mode != "auto" else ("multimem" if (mc and self.world > 2) else "p2p")
        self.push_mode = push_mode if push_mode != "auto" else ("multimem" if (mc and self.world > 2) else "p2p")
        if not mc and "multimem" in (self.reduce_mode, self.push_mode):
            raise RuntimeError("no multicast mapping on this system")
        self.sparse_reduce = bool(sparse_reduce)
        self.expand_after_reduce = bool(expand_after_reduce)  # the expansion behind the all-reduce instead of beside it
        # track_zero_rows: dL_dsh rows no view rendered are rewritten only if they still hold an earlier step's data (the
        # buffer starts zeroed and state 'unknown'); after anything but expand() wrote dL_dsh call mark_rows_dirty().
        # Off by default: 71 MB fewer writes per step, but no measurable gain at N = 8 (668 vs 665 us per step) -- the
        # expansion is latency-bound, not write-bound
        self.row_state = torch.full(((self.P + 31) // 32,), -1, dtype=torch.int32, device=self.device) if track_zero_rows else None
        self._comm = None
        if self.world > 1:
            self._comm = _lib.WindowComm()
            self._comm.rank, self._comm.world = self.rank, self.world
            for r in range(self.world):
                self._comm.pads[r] = ptrs[r] + 4 * self._pad_at
        self._reduce_bufs = (C.c_void_p * self.world)(*([mc] if self.reduce_mode == "multimem" else ptrs))
        # Streams. The pushes and the all-reduces are short on SM work but sit on the step's critical path and run beside
        # kernels that fill every SM (geom_bwd, the expansion): on high-priority streams their CTAs are placed as soon as
        # a slot frees up instead of queueing behind thousands of pending CTAs.
        self.side = torch.cuda.Stream(device=self.device)
        self.hi_push = torch.cuda.Stream(device=self.device, priority=-1)
        self.hi_reduce = torch.cuda.Stream(device=self.device, priority=-1)
        self.segment_events = None
        self.cam_centres = torch.zeros((self.n_views, 3), dtype=torch.float32, device=self.device)

    def set_active(self, n: int) -> None:
        """Exchange only the first n <= P Gaussians from now on (see engine.RasterContext.set_active; the buffers keep
        the layout of the capacity P). Every rank must set the same count."""
        if not 0 < int(n) <= self.P:
            raise ValueError("active count must be in (0, P]")
        self.active = int(n)
        per = self._chunk_per
        self.chunks = [(f, min(per, self.active - f)) for f in range(0, self.active, per)]

    # slot of view v, parity q: floats [at, at + slot_floats) then the visibility words
    def _slot_at(self, v: int, q: int) -> int:
        return self._slots_at + (q * self.n_views + v) * self._slot_stride

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _comm_ref(self):
        return self._C.byref(self._comm) if self._comm is not None else None

    def _local_slots(self, parity: int):
        C, base, V = self._C, self.buf.data_ptr(), self.n_views
        rgb = (C.c_void_p * V)(*[base + 4 * self._slot_at(v, parity) for v in range(V)])
        vis = (C.c_void_p * V)(*[base + 4 * (self._slot_at(v, parity) + self.slot_floats) for v in range(V)])
        return rgb, vis

    def mark_rows_dirty(self) -> None:
        """Something other than expand() wrote dL_dsh (a dense backward, a fill): the next expand rewrites every row."""
        if self.row_state is not None:
            self.row_state.fill_(-1)

    def timed_out(self) -> int:
        """0, or 1 + the channel on which one of this rank's waits gave up (~10 s: a peer died or the ranks diverged).
        Synchronises. The waits never hang the GPU; after a time-out the exchanged gradients are void."""
        if self._comm is None:
            return 0
        torch.cuda.synchronize(self.device)
        v = self._C.c_uint32(0)
        self._lib.check(self.lib.gsr_window_status(self._comm_ref(), self._C.byref(v)), "gsr_window_status")
        return int(v.value)

    def push(self, ctx, parity: int, view: Optional[int] = None, chunk: Optional[int] = None) -> None:
        """This rank's masked dRGB + visibility words (of one chunk, or all of them one after the other) into slot `view`
        (default: this rank's) of every rank, each followed by its PUSHED signal (current stream). Needs the blend half of
        ctx's backward."""
        C = self._C
        at = 4 * self._slot_at(self.rank if view is None else int(view), parity)
        bases = [self._mc] if self.push_mode == "multimem" else self._ptrs
        rgb = (C.c_void_p * len(bases))(*[b + at for b in bases])
        vis = (C.c_void_p * len(bases))(*[b + at + 4 * self.slot_floats for b in bases])
        for c in (range(len(self.chunks)) if chunk is None else [chunk]):
            first, count = self.chunks[c]
            st = self.lib.gsr_window_push_rgb(ctx.geom.buf.data_ptr(), self.active, first, count, c, ctx.radii.data_ptr(),
                                              ctx.grads["dL_dcolor"].data_ptr(), rgb, vis, len(bases),
                                              1 if self.push_mode == "multimem" else 0, self._comm_ref(), self._stream())
            self._lib.check(st, "gsr_window_push_rgb")

    def wait_pushed(self, chunk: int) -> None:
        """Every peer's push of the chunk has landed here (current stream)."""
        if self._comm is not None:
            self._lib.check(self.lib.gsr_window_wait(self._comm_ref(), self._lib.WINDOW_CH_PUSHED, chunk, 0, self._stream()),
                            "gsr_window_wait")

    def expand(self, ctx, means3D: torch.Tensor, parity: int, chunk: Optional[int] = None) -> None:
        """dL_dsh of the whole window from the gathered slots (current stream); one chunk, or all."""
        rgb, vis = self._local_slots(parity)
        for c in (range(len(self.chunks)) if chunk is None else [chunk]):
            first, count = self.chunks[c]
            st = self.lib.gsr_window_expand_sh(self.active, first, count, ctx.D, self.M, self.n_views, means3D.data_ptr(),
                                               self.cam_centres.data_ptr(), rgb, vis, ctx.grads["dL_dsh"].data_ptr(),
                                               self.row_state.data_ptr() if self.row_state is not None else None,
                                               self._stream())
            self._lib.check(st, "gsr_window_expand_sh")

    def allreduce_rest_(self, parity: int) -> None:
        """Sum over the ranks of the 11 non-SH parameter-gradient floats per Gaussian (current stream). The kernels wait
        for every rank's READY (and, sparse, the last chunk's PUSHED) themselves and signal DONE; THIS rank's own pushes
        must be ordered before by the caller."""
        if self._comm is None:
            return
        vis = self._local_slots(parity)[1] if self.sparse_reduce else None
        st = self.lib.gsr_window_allreduce_rest(self._comm_ref(), self._reduce_bufs, 1 if self.reduce_mode == "multimem" else 0,
                                                self.active, 0, self.active, 0, len(self.chunks) - 1, self._rest_first,
                                                self.n_views if vis is not None else 0, vis, self._stream())
        self._lib.check(st, "gsr_window_allreduce_rest")

    def step_end(self) -> None:
        """LAST call of a step (current stream, after the other streams have been joined): every rank's part of the
        all-reduce has landed here; advances the step epoch."""
        if self._comm is not None:
            self._lib.check(self.lib.gsr_window_wait(self._comm_ref(), self._lib.WINDOW_CH_DONE, 0, 1, self._stream()),
                            "gsr_window_wait")

    def exchange(self, ctx, means3D: torch.Tensor, parity: int, blend_done: torch.cuda.Event, geom) -> None:
        """Everything after the blend half of the backward. `blend_done`: recorded on the current stream right after the
        blend half; `geom()`: issues the per-Gaussian half (GSR_BWD_SH_FACTORED) on the current stream. Per chunk c:
        push(c) on hi_push, PUSHED wait + expansion(c) on side; geom on the current stream, the all-reduce on hi_reduce
        behind geom and the last push. Joined at the end, then the DONE wait."""
        main = torch.cuda.current_stream(self.device)
        ev = self.segment_events  # None, or a dict that receives timing events of this step's segments (diagnostics)
        mark = (lambda n, s: ev.setdefault(n, torch.cuda.Event(enable_timing=True)).record(s)) if ev is not None else (lambda n, s: None)
        n = len(self.chunks)
        pushed = [torch.cuda.Event() for _ in range(n)]
        geom_done, reduced = torch.cuda.Event(), torch.cuda.Event()
        with torch.cuda.stream(self.hi_push):
            self.hi_push.wait_event(blend_done)
            mark("push_begin", self.hi_push)
            for c in range(n):
                self.push(ctx, parity, chunk=c)
                pushed[c].record(self.hi_push)
            mark("push_end", self.hi_push)
        def expansion():
            with torch.cuda.stream(self.side):  # (normal priority: as a high-priority stream it starves geom_bwd, 761 vs 685 us)
                for c in range(n):
                    self.side.wait_event(pushed[c])
                    self.wait_pushed(c)
                    self.expand(ctx, means3D, parity, chunk=c)
                mark("expand_end", self.side)

        if not (self.expand_after_reduce and self._comm is not None):
            expansion()
        mark("geom_begin", main)
        geom()
        geom_done.record(main)
        mark("geom_end", main)
        if self._comm is not None:
            with torch.cuda.stream(self.hi_reduce):
                self.hi_reduce.wait_event(geom_done)
                self.hi_reduce.wait_event(pushed[n - 1])
                self.allreduce_rest_(parity)
                mark("reduce_end", self.hi_reduce)
                reduced.record(self.hi_reduce)
            if self.expand_after_reduce:
                self.side.wait_event(reduced)
                expansion()
            main.wait_stream(self.hi_reduce)
        main.wait_stream(self.hi_push)
        main.wait_stream(self.side)
        self.step_end()
        mark("step_end", main)

    def render_step(self, ctx, scene: Dict[str, torch.Tensor], cam, dL_dcolor: torch.Tensor, dL_dacc: torch.Tensor,
                    parity: int) -> None:
        """One view per rank: forward + backward of `cam` on ctx, then the factored exchange; afterwards ctx's
        parameter gradients (ctx.param_grads()) hold the window's sum on every rank. `self.cam_centres` [world,3] must
        hold the camera centres of the step's views in rank order before the step runs (it is read on the device, so a
        captured step can be replayed with new cameras). Nothing in here waits for the GPU. Steps alternate `parity`
        (the gather slots are double-buffered: a rank may be pushing step k+1 while a peer still expands step k)."""
        main = torch.cuda.current_stream(self.device)
        self.side.wait_stream(main)
        with torch.cuda.stream(self.side):
            ctx.prepare_backward()
        ctx.forward(scene, cam)
        main.wait_stream(self.side)
        ctx.backward(dL_dcolor, dL_dacc, accumulators_zeroed=True, part="blend")
        blend_done = torch.cuda.Event()
        blend_done.record(main)
        self.exchange(ctx, scene["means3D"], parity, blend_done,
                      lambda: ctx.backward(dL_dcolor, dL_dacc, part="geom", sh_factored=True))


class WindowStep:
    """One optimisation step over a window of views, sharded over the ranks of `group`.

    render_view(view_index, accumulate_into: Tensor) must run forward + loss + backward for that view on this
    rank's device and ADD the view's flat parameter gradient into `accumulate_into` (same layout on all ranks).
    """

    def __init__(self, n_views: int, grad_numel: int, device, render_view: Callable[[int, torch.Tensor], float],
                 group=None, dtype=torch.float32):
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.group = group
        self.views = views_for_rank(n_views, self.rank, self.world)
        self.flat = torch.zeros(grad_numel, dtype=dtype, device=device)
        self.render_view = render_view

    def run(self) -> Dict[str, float]:
        self.flat.zero_()
        loss = 0.0
        for v in self.views:
            loss += float(self.render_view(v, self.flat))
        allreduce_sum_(self.flat, self.group)
        return {"local_loss": loss, "local_views": float(len(self.views))}


def sgd_update_(params: Sequence[torch.Tensor], flat_grad: torch.Tensor, lr: float) -> None:
    """Tiny deterministic update used by the tests to check that replicas stay bit-identical. (The product keeps
    LibTorch's Adam, src/gs/gaussian.cu:396-428, which is outside the replaced path.)"""
    off = 0
    for p in params:
        n = p.numel()
        p.add_(flat_grad[off:off + n].view_as(p), alpha=-lr)
        off += n
    if off != flat_grad.numel():
        raise ValueError("flat gradient layout does not match the parameter list")
