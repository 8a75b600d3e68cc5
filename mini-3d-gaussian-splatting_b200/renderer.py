"""Drop-in for the reference's differentiable renderer (src/core/renderer.py).

Same public surface -- ``RenderSettings`` (renderer.py:13-20), ``GaussianRenderer(tile_size,
radius_min, radius_max)`` (renderer.py:24-28) and ``render(camera, gaussians, settings)``
(renderer.py:31-114) returning ``image, alpha, depth, viewspace_points, visibility_filter,
radii, conics`` with autograd intact -- but every stage runs in hand-written sm_100a kernels
behind the C ABI of ``include/gsplat_b200.h``.  Host code here is plumbing only: two
``torch.autograd.Function``s around the projection and compositing kernels with the
non-differentiable depth sort / tile binning between them.

CUDA only.  There is deliberately no CPU path: CPU tensors raise.
"""
from __future__ import annotations

import collections
import ctypes
import math
from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr

RASTER_BLOCK = 16                 # csrc/raster.cu: a warp composites a 16x16 pixel block; a tile is ceil(T/16)^2 of them
MAX_COUNTING_TILES = 200000       # csrc/binsort.cu kMaxCountingTiles: one byte of shared memory per tile
_U8 = torch.uint8
_I32 = torch.int32
_I64 = torch.int64
_F32 = torch.float32


@dataclass
class RenderSettings:
    """Render settings; field for field the reference's dataclass (renderer.py:13-20)."""
    image_height: int
    image_width: int
    bg_color: torch.Tensor
    scale_modifier: float = 1.0   # never read by the reference renderer either
    debug: bool = False


def next_list_cap(cap: int, flagged: int, deepest: int, recent_deepest: int, auto: bool):
    """List-cap policy of the truncated tile lists: (new cap, new memory of the deepest recent walk).

    `flagged` tiles of an earlier frame needed more than their stored prefix (they were completed and composited again
    inside that frame: any cap is safe); `deepest` is the deepest walk any tile made (-1 when not measured).
    With `auto` the cap follows what the tiles actually walk -- the deepest walk of the recent frames, forgotten by 2 % per
    report, plus an eighth, in steps of 64 entries and never below 128: a tight cap saves scattered stores (8.4 M
    four-byte stores at cap 1 024 on config[1], whose tiles never walk past entry 492).  Without it, or when a frame had
    to complete tiles although the target says the cap should have sufficed, the cap doubles."""
    hi = recent_deepest
    if auto and deepest >= 0:
        hi = max(deepest, int(recent_deepest * 0.98))
        target = max(128, -(-int(hi * 1.125 + 16) // 64) * 64)
        cap = target if not flagged else max(target, cap * 2)
    elif flagged > 0:
        cap = cap * 2
    return min(cap, 1 << 30), hi


class _ViewMeta:
    """Per-call constants shared by forward and backward of the two Functions."""

    __slots__ = ("cam", "W", "H", "tile", "rmin", "rmax", "param_mode", "opacity_is_logit",
                 "feat_stride", "sh_degree", "rest_in_feat", "sink")

    def __init__(self):
        self.cam = None
        self.sink = None


def _camera_block(camera) -> ctypes.Array:
    """renderer.py:140-152: intrinsics in python float64 rounded once to fp32; W2C rotation and
    translation, then the camera centre -R^T t (used by the SH option only).  20 host floats
    (GS_CAMERA_FLOATS).  Plain Python arithmetic on the 16 matrix entries: this runs once per frame on
    the launch path, ahead of the first kernel."""
    W, H = camera._width, camera._height
    fx = float(np.float32(0.5 * W / math.tan(camera._FoVx * 0.5)))
    fy = float(np.float32(0.5 * H / math.tan(camera._FoVy * 0.5)))
    wv = camera.world_view_transform()
    if wv.device.type != "cpu" or wv.dtype != _F32:
        wv = wv.detach().to(device="cpu", dtype=_F32)
    m = wv.tolist()
    r0, r1, r2 = m[0], m[1], m[2]
    tx, ty, tz = r0[3], r1[3], r2[3]
    return (ctypes.c_float * 20)(r0[0], r0[1], r0[2], r1[0], r1[1], r1[2], r2[0], r2[1], r2[2], tx, ty, tz,
                                 fx, fy, float(np.float32(W * 0.5)), float(np.float32(H * 0.5)),
                                 -(r0[0] * tx + r1[0] * ty + r2[0] * tz), -(r0[1] * tx + r1[1] * ty + r2[1] * tz),
                                 -(r0[2] * tx + r1[2] * ty + r2[2] * tz), 0.0)


def _stream(device) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class StageTimer:
    """Optional CUDA-event timing of each ABI call, on the stream the kernels are launched on.
    Assign an instance to ``stage_timer.active`` (module level) to collect; bench.py does, tests
    and normal use leave it None so no events are recorded."""

    def __init__(self):
        self.events = []          # (name, start_event, end_event)

    def summary_ms(self):
        torch.cuda.synchronize()
        out = {}
        for name, a, b in self.events:
            out.setdefault(name, []).append(a.elapsed_time(b))
        return out


class _TimerSlot:
    active: Optional[StageTimer] = None


stage_timer = _TimerSlot()


class _timed:
    def __init__(self, name, device):
        self.name, self.device, self.t = name, device, stage_timer.active

    def __enter__(self):
        if self.t is not None:
            self.a = torch.cuda.Event(enable_timing=True)
            self.a.record(torch.cuda.current_stream(self.device))

    def __exit__(self, *exc):
        if self.t is not None:
            b = torch.cuda.Event(enable_timing=True)
            b.record(torch.cuda.current_stream(self.device))
            self.t.events.append((self.name, self.a, b))
        return False


class _ProjectFn(torch.autograd.Function):
    """Stage P+M+C (gs_project_fwd / gs_project_bwd)."""

    @staticmethod
    def forward(ctx, meta, xyz, scaling, rotation, cov3d, opacity, feat_src, features_rest):
        lib = _lib.load()
        n = xyz.shape[0]
        dev = xyz.device
        new = lambda *shape, dtype=_F32: torch.empty(shape, dtype=dtype, device=dev)  # noqa: E731
        means2d, depths, conics, radii = new(n, 2), new(n), new(n, 2, 2), new(n)
        colors, opac = new(n, 3), new(n)
        vis = new(n, dtype=_U8)
        tiles_touched = new(n, dtype=_I32)
        tile_rect = new(n, 4, dtype=torch.int16)
        depth_keys = new(n, dtype=_I32)
        rec = new(n, 12)
        with _timed("project_fwd", dev):
            check(lib.gs_project_fwd(
                n, ptr(xyz), ptr(scaling), ptr(rotation), ptr(cov3d), ptr(opacity), int(meta.opacity_is_logit),
                ptr(feat_src), meta.feat_stride, *_sh_args(meta, feat_src, features_rest), meta.cam,
                meta.W, meta.H, meta.tile, meta.rmin, meta.rmax,
                ptr(means2d), ptr(depths), ptr(conics), ptr(radii), ptr(colors), ptr(opac), ptr(vis),
                ptr(tiles_touched), ptr(tile_rect), ptr(depth_keys), ptr(rec), _stream(dev)), "gs_project_fwd")
        ctx.meta = meta
        ctx.set_materialize_grads(False)       # unused outputs arrive as None instead of freshly zeroed tensors
        ctx.save_for_backward(xyz, scaling, rotation, cov3d, opacity, feat_src, features_rest)
        if meta.sink is not None:
            ctx.stat_inputs = (radii, vis)        # plain (non-graph) outputs: read by the fused statistics
        vis_b = vis.view(torch.bool)
        ctx.mark_non_differentiable(radii, vis_b, tiles_touched, tile_rect, depth_keys, rec)
        return means2d, conics, depths, colors, opac, radii, vis_b, tiles_touched, tile_rect, depth_keys, rec

    @staticmethod
    def backward(ctx, g_means2d, g_conics, g_depths, g_colors, g_opac, *_unused):
        lib = _lib.load()
        meta = ctx.meta
        xyz, scaling, rotation, cov3d, opacity, feat_src, features_rest = ctx.saved_tensors
        n = xyz.shape[0]
        dev = xyz.device

        def dense(g, *shape):
            if g is None:
                return torch.zeros(shape, dtype=_F32, device=dev)
            return g.contiguous()

        g_means2d = dense(g_means2d, n, 2)
        g_conics = dense(g_conics, n, 2, 2)
        g_depths = dense(g_depths, n)
        g_colors = dense(g_colors, n, 3)
        g_opac = dense(g_opac, n)
        sink = meta.sink
        if sink is not None:
            # multi-view accumulation: add straight into the caller's flat gradient buffer and fill the
            # densification statistics in the same pass; autograd then has nothing left to accumulate
            g_xyz, g_feat, g_scaling, g_rotation, g_opacity = sink.views
            radii, vis = ctx.stat_inputs
            # the first view of a step WRITES gradients and statistics (the buffer needs no zero-fill), later views add
            accumulate = 0 if getattr(sink, "fresh", False) else 1
            sink.fresh = False
            with _timed("project_bwd", dev):
                check(lib.gs_project_bwd(
                    n, ptr(xyz), ptr(scaling), ptr(rotation), None, ptr(opacity), 1,
                    ptr(feat_src), meta.feat_stride, None, 0, 0, meta.cam,
                    ptr(g_means2d), ptr(g_conics), ptr(g_depths), ptr(g_colors), ptr(g_opac),
                    ptr(g_xyz), ptr(g_scaling), ptr(g_rotation), None, ptr(g_opacity),
                    ptr(g_feat), 3, None, 0, accumulate,
                    ptr(radii), ptr(vis), ptr(sink.grad_norm_sum), ptr(sink.vis_count), ptr(sink.max_radii),
                    _stream(dev)), "gs_project_bwd")
            return (None,) * 8
        g_xyz = torch.empty_like(xyz)
        g_scaling = torch.empty_like(scaling) if meta.param_mode else None
        g_rotation = torch.empty_like(rotation) if meta.param_mode else None
        g_cov3d = None if meta.param_mode else torch.empty((n, 3, 3), dtype=_F32, device=dev)
        g_opacity = torch.empty_like(opacity)
        # only features[:,0,:] feeds the colour; every other SH row gets the dense zeros the
        # reference's autograd produces
        sh_on = meta.sh_degree > 0
        if feat_src.shape[1] == 1 or (sh_on and meta.rest_in_feat):
            g_feat = torch.empty_like(feat_src)            # every row is written by the kernel
        else:
            g_feat = torch.zeros_like(feat_src)
        g_rest = None
        if sh_on and not meta.rest_in_feat:
            g_rest = torch.empty_like(features_rest)
        elif features_rest is not None and ctx.needs_input_grad[7]:
            g_rest = torch.zeros_like(features_rest)
        if sh_on:
            g_sh = (ctypes.c_void_p(g_feat.data_ptr() + 12), g_feat.stride(0)) if meta.rest_in_feat else (ptr(g_rest), g_rest.stride(0))
        else:
            g_sh = (None, 0)
        with _timed("project_bwd", dev):
            check(lib.gs_project_bwd(
                n, ptr(xyz), ptr(scaling), ptr(rotation), ptr(cov3d), ptr(opacity), int(meta.opacity_is_logit),
                ptr(feat_src), meta.feat_stride, *_sh_args(meta, feat_src, features_rest), meta.cam,
                ptr(g_means2d), ptr(g_conics), ptr(g_depths), ptr(g_colors), ptr(g_opac),
                ptr(g_xyz), ptr(g_scaling), ptr(g_rotation), ptr(g_cov3d), ptr(g_opacity),
                ptr(g_feat), g_feat.stride(0), g_sh[0], g_sh[1], 0, None, None, None, None, None,
                _stream(dev)), "gs_project_bwd")
        return None, g_xyz, g_scaling, g_rotation, g_cov3d, g_opacity, g_feat, g_rest


def _sh_args(meta, feat_src, features_rest):
    """(sh_rest pointer, row stride in floats, degree) for the ABI.  The higher-order rows live either
    in a separate `_features_rest [N,15,3]` (GaussianModel layout) or behind row 0 of `features [N,K,3]`."""
    if meta.sh_degree <= 0:
        return None, 0, 0
    if meta.rest_in_feat:
        return ctypes.c_void_p(feat_src.data_ptr() + 12), feat_src.stride(0), meta.sh_degree
    return ptr(features_rest), features_rest.stride(0), meta.sh_degree


class _RasterizeFn(torch.autograd.Function):
    """Stage B+R: tile binning (gs_bin_sort, non-differentiable) and compositing (gs_raster_fwd /
    gs_raster_bwd).  `means2d` is the tensor returned to the caller as ``viewspace_points``; its
    ``.grad`` after ``retain_grad()`` is what this backward emits.

    The binning needs the frame's pair count D, which only the device knows.  With a capacity learnt
    from earlier frames (`bins.d_cap`) everything is enqueued at once with device-side sizes and the
    counters are read back afterwards (no bubble on the GPU); the first frame, and a frame whose D
    outgrows the capacity, go through the exact path: read D, then enqueue."""

    @staticmethod
    def forward(ctx, meta, bins, means2d, conics, depths, colors, opac, rec, bg, track):
        lib = _lib.load()
        dev = means2d.device
        H, W = meta.H, meta.W
        T = meta.tile
        n = means2d.shape[0]
        tiles_x, tiles_y = (W + T - 1) // T, (H + T - 1) // T
        tiles = tiles_x * tiles_y
        sub = (T + RASTER_BLOCK - 1) // RASTER_BLOCK
        slots = tiles * sub * sub             # launch slots of the compositing kernels: (tile, 16x16 block) pairs
        stream = _stream(dev)
        image = torch.empty((3, H, W), dtype=_F32, device=dev)
        alpha = torch.empty((1, H, W), dtype=_F32, device=dev)
        depth = torch.empty((1, H, W), dtype=_F32, device=dev)
        pix_state = torch.empty((H * W, 4), dtype=_F32, device=dev)
        # per-pixel consumed-entry counts: debug / parity output only (RenderSettings.debug)
        n_consumed = torch.empty((H, W), dtype=_I32, device=dev) if track else None
        tile_consumed = torch.empty((slots,), dtype=_I32, device=dev)
        tile_ranges = torch.empty((tiles, 2), dtype=_I32, device=dev)
        # launch order of the tiles: heaviest first, by the work they had in the previous frame of this size
        # (any permutation gives the same image; a good one keeps full-size tiles out of the last wave)
        cached_order = bins.cached_order          # this camera's order by exact work from its previous visit ("camera" mode)
        if cached_order is not None and (cached_order.numel() != slots or cached_order.device != dev):
            cached_order = None
        prev = bins.prev_consumed if (bins.fwd_order == "previous" and cached_order is None) else None
        if prev is not None and (prev.numel() != slots or prev.device != dev):
            prev = None
        # tiles larger than a block: the list lengths are per tile, not per slot -- without an estimate per slot (this
        # camera's previous visit, or the previous frame) the slots are taken in natural order
        ranges_order = sub == 1
        if cached_order is not None:
            bins.tile_order = cached_order
        elif prev is not None or ranges_order:
            bins.tile_order = torch.empty(slots, dtype=_I32, device=dev)
        else:
            bins.tile_order = None
        if prev is not None:
            check(lib.gs_tile_order(slots, ptr(prev), None, ptr(bins.tile_order), stream), "gs_tile_order")

        # truncated lists (flat counting sort, not in debug mode): store and composite only each tile's first
        # `cap` list entries; a tile that needs more is flagged, completed and composited again by the two
        # self-skipping completion launches, so the result never depends on `cap`
        cap = int(bins.list_cap or 0) if (bins.algo in (0, 1) and not track and sub == 1) else 0
        flag_bytes = (tiles + 3) // 4 * 4
        # no zero-fill: gs_bin_sort zeroes the count, the first compositing pass writes every tile's flag
        flagbuf = torch.empty(flag_bytes + 4, dtype=_U8, device=dev) if cap else None
        tile_flags = flagbuf[:tiles] if cap else None
        flag_count = flagbuf[flag_bytes:].view(_I32) if cap else None

        # What the backward pass needs besides the forward's outputs -- the zeroed gradient slab its atomics add into and
        # the tiles' launch order by exact work -- is produced on a side stream while this stream composites / evaluates
        # the loss: the 44 MB fill has no dependency at all, the order needs only tile_consumed.
        need_bwd = any(ctx.needs_input_grad) and n > 0
        want_order = need_bwd or bins.fwd_order == "camera"       # the same order serves this camera's next forward
        side = bins.side_stream() if want_order else None
        slab = bwd_order = None
        if side is not None:
            main = torch.cuda.current_stream(dev)
            ev = bins.side_events()
            bwd_order = torch.empty(slots, dtype=_I32, device=dev)
            bwd_order.record_stream(side)
            if need_bwd:
                slab = torch.empty(n * 11, dtype=_F32, device=dev)
                ev[0].record(main)               # the slab's memory may still be in use by kernels enqueued so far
                side.wait_event(ev[0])
                with torch.cuda.stream(side):
                    slab.zero_()
                slab.record_stream(side)

        def enqueue(num_sorted, d_size, counters_dev):
            entry_ids = torch.empty(max(d_size, 1), dtype=_I32, device=dev)
            ws_bytes = int(lib.gs_bin_workspace_bytes(num_sorted, d_size, tiles))
            ws = torch.empty(ws_bytes, dtype=_U8, device=dev)
            # this frame's list lengths give the forward's tile order: the flat counting sort emits it from the launch
            # that scans the tiles; the other algorithms need the separate kernel
            fused_order = prev is None and cached_order is None and ranges_order and int(bins.algo) == 1
            with _timed("bin_sort", dev):
                check(lib.gs_bin_sort(n, num_sorted, d_size, ptr(bins.sorted_ids), ptr(bins.offsets), ptr(bins.tile_rect),
                                      ptr(bins.depth_keys), tiles_x, tiles, int(bins.algo), ptr(ws), ws.numel(),
                                      ptr(entry_ids), ptr(tile_ranges), None, counters_dev, cap,
                                      ptr(bins.tile_order) if fused_order else None, ptr(flag_count), stream), "gs_bin_sort")
            if prev is None and cached_order is None and ranges_order and not fused_order:
                check(lib.gs_tile_order(tiles, None, ptr(tile_ranges), ptr(bins.tile_order), stream), "gs_tile_order")
            vis_host = int(bins.num_vis > 0) if counters_dev is None else 0

            def composite(rerun):
                check(lib.gs_raster_fwd(W, H, T, ptr(entry_ids), ptr(tile_ranges), ptr(rec), ptr(bg), vis_host, counters_dev,
                                        ptr(bins.tile_order), cap, ptr(tile_flags), ptr(flag_count), rerun,
                                        ptr(image), ptr(alpha), ptr(depth), ptr(pix_state),
                                        ptr(n_consumed), ptr(tile_consumed), stream), "gs_raster_fwd")
            with _timed("raster_fwd", dev):
                composite(0)
                if cap:
                    check(lib.gs_bin_complete(n, num_sorted, d_size, ptr(bins.sorted_ids), ptr(bins.offsets), ptr(bins.tile_rect),
                                              tiles_x, tiles, ptr(ws), ws.numel(), cap, ptr(tile_flags), ptr(flag_count),
                                              ptr(entry_ids), counters_dev, stream), "gs_bin_complete")
                    composite(1)
            return entry_ids

        entry_ids = None
        if bins.d_cap is not None and bins.algo != 2 and n > 0:
            entry_ids = enqueue(n, bins.d_cap, ptr(bins.counters))
            bins.read_counters()
            if bins.D > bins.d_cap:
                entry_ids = None                        # did not fit: nothing was written, repeat with exact sizes
        else:
            bins.read_counters()
        if entry_ids is None:
            entry_ids = enqueue(bins.num_sorted, bins.D, None)
        entry_ids = entry_ids[:bins.D]
        bins.entry_ids, bins.tile_ranges = entry_ids, tile_ranges
        bins.renderer_consumed[(dev.index, W, H, T)] = tile_consumed

        ctx.side = None
        if cap and side is None:
            bins.report_flagged(flag_count)      # asynchronous: a later frame doubles the cap if tiles were flagged
        if side is not None:
            ev[1].record(main)                   # tile_consumed and the flag count are final
            side.wait_event(ev[1])
            if cap:
                bins.report_flagged(flag_count, tile_consumed, side)
            check(lib.gs_tile_order(slots, ptr(tile_consumed), None, ptr(bwd_order),
                                    ctypes.c_void_p(side.cuda_stream)), "gs_tile_order")
            ev[2].record(side)
            bins.new_order = (bwd_order, ev[2])
            if need_bwd:
                ctx.side = (ev[2], slab, bwd_order)
        ctx.meta = meta
        ctx.set_materialize_grads(False)
        ctx.any_visible = bins.num_vis > 0
        ctx.n = n
        ctx.save_for_backward(rec, entry_ids, tile_ranges, bg, alpha, pix_state, tile_consumed)
        if n_consumed is None:
            n_consumed = torch.empty(0, dtype=_I32, device=dev)
        ctx.mark_non_differentiable(n_consumed, tile_consumed)
        return image, alpha, depth, n_consumed, tile_consumed

    @staticmethod
    def backward(ctx, g_image, g_alpha, g_depth, *_unused):
        lib = _lib.load()
        meta = ctx.meta
        rec, entry_ids, tile_ranges, bg, alpha, pix_state, tile_consumed = ctx.saved_tensors
        n = ctx.n
        dev = rec.device
        H, W = meta.H, meta.W
        # one zeroed slab carved into the five contiguous gradient tensors (atomics accumulate into it); zero-filled, like
        # the tile order below, on the side stream during the forward pass
        order = None
        if ctx.side is not None:
            side_done, slab, order = ctx.side
            ctx.side = None                      # a second backward through the same graph gets a fresh slab
            torch.cuda.current_stream(dev).wait_event(side_done)
        else:
            slab = torch.zeros(n * 11, dtype=_F32, device=dev)
        # float4-read array first, then the float2 one: both stay naturally aligned for any n
        g_conics = slab[0:4 * n].view(n, 2, 2)
        g_means2d = slab[4 * n:6 * n].view(n, 2)
        g_depths = slab[6 * n:7 * n]
        g_colors = slab[7 * n:10 * n].view(n, 3)
        g_opac = slab[10 * n:11 * n]
        if ctx.any_visible and entry_ids.numel() > 0:
            # outputs no gradient flows into: image needs dense zeros, alpha / depth are passed as NULL
            gi = g_image.contiguous() if g_image is not None else torch.zeros((3, H, W), dtype=_F32, device=dev)
            ga = g_alpha.contiguous() if g_alpha is not None else None
            gd = g_depth.contiguous() if g_depth is not None else None
            if gd is not None and ga is None:
                ga = torch.zeros((1, H, W), dtype=_F32, device=dev)
            scratch = None
            if order is None:                    # heaviest tiles first (exact work): ordered inside gs_raster_bwd
                order = scratch = torch.empty(tile_consumed.numel(), dtype=_I32, device=dev)
            with _timed("raster_bwd", dev):
                check(lib.gs_raster_bwd(W, H, meta.tile, ptr(entry_ids), ptr(tile_ranges), ptr(rec), ptr(bg), ptr(alpha),
                                        ptr(pix_state), ptr(tile_consumed), ptr(order), int(scratch is None), ptr(gi), ptr(ga), ptr(gd),
                                        ptr(g_means2d), ptr(g_conics), ptr(g_depths), ptr(g_colors), ptr(g_opac),
                                        _stream(dev)), "gs_raster_bwd")
        return None, None, g_means2d, g_conics, g_depths, g_colors, g_opac, None, None, None


class _FrameBins:
    """Per-frame binning state handed to `_RasterizeFn`: the depth order, the device counters
    {splats with tiles, tile pairs D, visible splats} and their asynchronous host copy."""

    def __init__(self, renderer, device):
        self.device = device
        self.d_cap = renderer._d_cap.get(device.index) if renderer.optimistic_binning else None
        slot = renderer._readback.get(device.index)
        if slot is None:
            slot = (torch.empty(3, dtype=_I64).pin_memory(), torch.cuda.Event())
            renderer._readback[device.index] = slot
        self._host, self._event = slot
        self.renderer_consumed = renderer._tile_consumed
        self.prev_consumed = None
        self.cached_order = None
        self.new_order = None
        self.tile_order = None
        self.fwd_order = renderer.fwd_tile_order
        self._renderer = renderer
        fb = renderer._cap_feedback.get(device.index)
        if fb is not None and fb[1].query():          # an earlier frame's report {flagged tiles, deepest walk} has arrived
            flagged, deepest = int(fb[0][0]), int(fb[0][1])
            renderer._cap_feedback[device.index] = None
            if renderer.list_cap:
                renderer.list_cap, renderer._deepest_walk[device.index] = next_list_cap(
                    int(renderer.list_cap), flagged, deepest, renderer._deepest_walk.get(device.index, 0), renderer.list_cap_auto)
        self.list_cap = renderer.list_cap
        self.num_sorted = self.D = self.num_vis = None
        self.entry_ids = self.tile_ranges = None

    def report_flagged(self, flag_count, tile_consumed=None, stream=None):
        """Asynchronous report for the list-cap policy: {tiles the first pass flagged, deepest walk of any tile (-1 when
        not measured)} copied to pinned memory on `stream` (the side stream when there is one: off the frame's critical
        path) and read by a later frame once it has arrived."""
        slot = self._renderer._cap_pinned.get(self.device.index)
        if slot is None:
            slot = self._renderer._cap_pinned[self.device.index] = torch.zeros(2, dtype=_I32).pin_memory()
        if self._renderer._cap_feedback.get(self.device.index) is None:      # one report in flight at a time
            stream = stream or torch.cuda.current_stream(self.device)
            with torch.cuda.stream(stream):
                slot[0:1].copy_(flag_count, non_blocking=True)
                if tile_consumed is not None and self._renderer.list_cap_auto:
                    slot[1:2].copy_(tile_consumed.max().to(_I32).reshape(1), non_blocking=True)
                else:
                    slot[1] = -1
                ev = torch.cuda.Event()
                ev.record(stream)
            self._renderer._cap_feedback[self.device.index] = (slot, ev)

    def side_stream(self):
        """Per-device side stream for work the backward pass will need (None when disabled)."""
        r = self._renderer
        if not r.side_stream_prep:
            return None
        s_ = r._side.get(self.device.index)
        if s_ is None:
            s_ = r._side[self.device.index] = (torch.cuda.Stream(device=self.device),
                                               [torch.cuda.Event() for _ in range(3)])
        return s_[0]

    def side_events(self):
        return self._renderer._side[self.device.index][1]

    def start_readback(self):
        self._host.copy_(self.counters, non_blocking=True)
        self._event.record(torch.cuda.current_stream(self.device))

    def read_counters(self):
        """The frame's one host wait: for the 24-byte counter copy, not for the kernels enqueued after it
        (the reference syncs at the same point, on vis_mask.sum(), renderer.py:74)."""
        self._event.synchronize()
        self.num_sorted, self.D, self.num_vis = (int(v) for v in self._host.tolist())


def _is_parameter_model(g) -> bool:
    """True for objects laid out like the reference's GaussianModel (gaussian_model.py:21-40):
    raw parameters plus the exp / sigmoid / normalize activations, which the projection kernel
    can fuse.  Anything else goes through the duck-typed accessors."""
    names = ("_xyz", "_scaling", "_rotation", "_opacity", "_features_dc")
    if not all(isinstance(getattr(g, a, None), torch.Tensor) for a in names):
        return False
    return (getattr(g, "scaling_activation", None) is torch.exp
            and getattr(g, "opacity_activation", None) is torch.sigmoid
            and getattr(g, "rotation_activation", None) is torch.nn.functional.normalize)


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != _F32:
        raise TypeError(f"expected float32 tensors, got {t.dtype}")
    return t.contiguous()


class GaussianRenderer:
    """3D Gaussian differentiable renderer -- B200 implementation of renderer.py:22-367."""

    def __init__(self, tile_size: int = 16, radius_min: float = 0.01, radius_max: float = 50.0, sh_degree: int = 0):
        """`sh_degree` is an extension: 0 (default) is the reference's DC-only colour
        sigmoid(features[:,0,:]); 1..3 add view-dependent real-SH terms from the higher feature rows
        (identical output while those rows are zero)."""
        if not 0 <= int(sh_degree) <= 3:
            raise ValueError("sh_degree must be 0..3")
        self.sh_degree = int(sh_degree)
        # any tile size, as the reference (renderer.py:24): the kernels composite 16x16 pixel blocks, a tile is covered by
        # ceil(T/16)^2 of them (16 is the fast case: one block per tile, nothing masked; truncated lists need T <= 16)
        if not 1 <= int(tile_size) <= 4096:
            raise ValueError(f"tile_size must be in 1..4096, got {tile_size}")
        self.tile_size = int(tile_size)
        self.radius_min = radius_min
        self.radius_max = radius_max
        self.device = torch.device("cuda")
        self.last_stats: Dict[str, int] = {}
        self.bin_algo = 0          # 0 = choose; 1 = flat counting sort, 2 = library radix sort (cross-check), 3 = blocked counting sort
        # Optional gradient sink (multiview.FlatGradBuffer): while set, the projection backward adds the
        # parameter gradients and the densification statistics straight into it (see `accumulate_into`).
        self.grad_sink = None
        # Optimistic binning: enqueue tile binning + compositing with a pair capacity learnt from the
        # previous frame instead of waiting for this frame's count (falls back to the exact path on
        # the first frame and whenever the count outgrows the capacity).
        self.optimistic_binning = True
        self._d_cap: Dict[int, Optional[int]] = {}
        self._readback: Dict[int, tuple] = {}
        self._tile_consumed: Dict[tuple, torch.Tensor] = {}     # last frame's per-tile work per (device, W, H)
        # work estimate behind the forward's tile launch order:
        #   "camera"   (default) the order by exact work (tile_consumed) from the last time THIS camera was rendered -- a
        #              training loop visits its cameras once per epoch, the scene moves little in between; the order is the
        #              one that frame's backward pass used, kept per camera (LRU of `camera_cache_size` poses, 32 KB each
        #              at 1080p), so a hit costs no kernel at all; a camera seen for the first time falls back to "ranges";
        #   "ranges"   this frame's list lengths (stateless; emitted by the binning's tile scan);
        #   "previous" what the tiles consumed in the previous frame of the same size, whatever its camera.
        self.fwd_tile_order = "camera"
        self.camera_cache_size = 1024
        self._order_by_camera = collections.OrderedDict()
        # truncated tile lists: entries stored / composited per tile before the completion path kicks in (0 = complete
        # lists).  Doubled automatically when a frame had to complete tiles.
        self.list_cap = 1024
        self.list_cap_auto = True       # let the cap follow the deepest walk the tiles actually make (see _FrameBins)
        self._deepest_walk: Dict[int, int] = {}
        # slab zero-fill and backward tile order on a side stream during the forward pass (see _RasterizeFn.forward)
        self.side_stream_prep = True
        self._side: Dict[int, tuple] = {}
        self._cap_feedback: Dict[int, Optional[tuple]] = {}
        self._cap_pinned: Dict[int, torch.Tensor] = {}
        _lib.load()   # fail at construction, not at first render, if the extension is missing

    def accumulate_into(self, sink):
        """Context manager: renders made inside it back-propagate by ADDING into `sink` -- an object
        with `n`, `views` (gradient buffers for _xyz, _features_dc, _scaling, _rotation, _opacity, in
        that order) and `grad_norm_sum / vis_count / max_radii` [N] -- instead of through autograd's
        per-tensor `.grad` accumulation.  Only GaussianModel-shaped inputs with DC-only colour take this
        path; anything else back-propagates as usual."""
        renderer = self

        class _Scope:
            def __enter__(self_inner):
                self_inner.prev = renderer.grad_sink
                renderer.grad_sink = sink
                return sink

            def __exit__(self_inner, *exc):
                renderer.grad_sink = self_inner.prev
                return False
        return _Scope()

    # ------------------------------------------------------------------------------------
    def render(self, camera, gaussians, settings: RenderSettings) -> Dict[str, torch.Tensor]:
        """Main render method; same contract as the reference (renderer.py:31-114).

        camera:    object with ``_width, _height, _FoVx, _FoVy`` and a callable
                   ``world_view_transform()`` -> 4x4 world-to-camera tensor.
        gaussians: either a GaussianModel-shaped object (raw ``_xyz/_scaling/_rotation/_opacity/
                   _features_dc`` + the reference's activations), or any object exposing
                   ``get_xyz, get_covariance, get_features, get_opacity``.
        """
        xyz = gaussians.get_xyz
        device = xyz.device
        if device.type != "cuda":
            raise RuntimeError("GaussianRenderer (B200) renders CUDA tensors only; there is no CPU fallback "
                               f"(gaussians.get_xyz is on {device})")
        H, W = int(settings.image_height), int(settings.image_width)
        with torch.cuda.device(device):
            return self._render(camera, gaussians, settings, device, H, W)

    def _render(self, camera, gaussians, settings, device, H, W):
        lib = _lib.load()
        T = int(self.tile_size)
        bg = settings.bg_color.detach().to(device=device, dtype=_F32).reshape(3).contiguous()

        meta = _ViewMeta()
        meta.cam = _camera_block(camera)
        meta.W, meta.H, meta.tile = W, H, T
        meta.rmin, meta.rmax = float(self.radius_min), float(self.radius_max)

        # ---- stage P+M+C -------------------------------------------------------------------
        if _is_parameter_model(gaussians):
            meta.param_mode, meta.opacity_is_logit = True, True
            xyz = _f32c(gaussians._xyz)
            scaling, rotation = _f32c(gaussians._scaling), _f32c(gaussians._rotation)
            opacity = _f32c(gaussians._opacity)
            feat_src = _f32c(gaussians._features_dc)
            rest = getattr(gaussians, "_features_rest", None)
            if rest is not None and rest.numel() == 0:
                rest = None
            cov3d = None
        else:
            meta.param_mode, meta.opacity_is_logit = False, False
            xyz = _f32c(gaussians.get_xyz)
            cov3d = _f32c(gaussians.get_covariance)
            feats = gaussians.get_features
            if not (feats.dim() == 3 and feats.shape[1] >= 1):          # renderer.py:89-92
                feats = gaussians._features_dc
            feat_src = _f32c(feats)
            opacity = _f32c(gaussians.get_opacity)
            scaling = rotation = rest = None
        n = xyz.shape[0]
        if feat_src.dim() != 3 or feat_src.shape[0] != n or feat_src.shape[2] != 3:
            raise ValueError(f"features must be [N,K,3], got {tuple(feat_src.shape)}")
        if opacity.numel() != n:
            raise ValueError(f"opacity must have N={n} elements, got {tuple(opacity.shape)}")
        meta.feat_stride = feat_src.stride(0)
        meta.sh_degree = self.sh_degree
        meta.rest_in_feat = False
        if self.sh_degree > 0:
            need = (self.sh_degree + 1) ** 2
            if rest is not None and rest.shape[1] >= need - 1:
                rest = _f32c(rest)
                if rest.shape[1] != 15:
                    raise ValueError("_features_rest must be [N,15,3]")
            elif feat_src.shape[1] >= 16:
                meta.rest_in_feat, rest = True, None
            else:
                raise ValueError(f"sh_degree={self.sh_degree} needs 15 higher-order feature rows "
                                 f"(_features_rest [N,15,3] or get_features [N,16,3])")

        sink = self.grad_sink
        if (sink is not None and meta.param_mode and self.sh_degree == 0 and sink.n == n and feat_src.shape[1] == 1
                and torch.is_grad_enabled()):
            meta.sink = sink
        elif sink is not None and getattr(sink, "fresh", False):
            # the caller skipped the zero-fill expecting this view to overwrite the buffer, but the view cannot take
            # the sink path: gradients will arrive through autograd's `+=`, which needs zeros
            sink.storage.zero_()
            sink.fresh = False

        (means2d, conics, depths, colors, opac, radii, vis, tiles_touched, tile_rect, depth_keys,
         rec) = _ProjectFn.apply(meta, xyz, scaling, rotation, cov3d, opacity, feat_src, rest)

        # ---- stage S (depth order, non-differentiable) ---------------------------------------------
        tiles_x, tiles_y = (W + T - 1) // T, (H + T - 1) // T
        num_tiles = tiles_x * tiles_y
        stream = _stream(device)
        bins = _FrameBins(self, device)
        bins.prev_consumed = self._tile_consumed.get((device.index, W, H, T))
        cam_key = (device.index, W, H, T, bytes(meta.cam))
        if self.fwd_tile_order == "camera" and self.side_stream_prep:
            hit = self._order_by_camera.get(cam_key)
            if hit is not None:
                self._order_by_camera.move_to_end(cam_key)
                torch.cuda.current_stream(device).wait_event(hit[1])      # written on the side stream, long since
                bins.cached_order = hit[0]
        # 0 = the flat counting sort.  3 (blocked two-level sort, coalesced final stores) is bit-identical and
        # measured no faster (385 vs 376 us at config[1]); it needs rectangles of at most 8 tiles per side
        algo = self.bin_algo if self.bin_algo else (1 if num_tiles <= MAX_COUNTING_TILES else 2)
        if algo == 3 and self.radius_max > 50.0:
            raise ValueError("bin_algo=3 (blocked) needs radius_max <= 50")
        bins.tile_rect, bins.depth_keys, bins.algo = tile_rect, depth_keys, algo
        if algo == 2:
            bins.d_cap = None            # the radix path needs exact sizes (and stores complete lists: list_cap is ignored)
        bins.counters = torch.empty(3, dtype=_I64, device=device)
        bins.sorted_ids = torch.empty(n, dtype=_I32, device=device)
        bins.offsets = torch.empty(n, dtype=_I64, device=device)
        ws_bytes = int(lib.gs_bin_workspace_bytes(n, 0, num_tiles))
        ws = torch.empty(ws_bytes, dtype=_U8, device=device)
        with _timed("bin_prepare", device):
            check(lib.gs_bin_prepare(n, ptr(depth_keys), ptr(tiles_touched), ptr(ws), ws_bytes, ptr(bins.sorted_ids),
                                     ptr(bins.offsets), ptr(bins.counters), stream), "gs_bin_prepare")
        bins.start_readback()

        # ---- stage B+R -----------------------------------------------------------------------
        image, alpha, depth, n_consumed, tile_consumed = _RasterizeFn.apply(
            meta, bins, means2d, conics, depths, colors, opac, rec, bg, bool(getattr(settings, "debug", False)))
        num_sorted, D, num_vis = bins.num_sorted, bins.D, bins.num_vis
        entry_ids, tile_ranges, sorted_ids = bins.entry_ids, bins.tile_ranges, bins.sorted_ids
        # capacity for the next frame's optimistic binning: this frame's pair count plus a quarter
        # (gs_bin_sort addresses pairs with int32: a capacity beyond that falls back to exact sizes)
        if bins.new_order is not None and self.fwd_tile_order == "camera":
            self._order_by_camera[cam_key] = bins.new_order
            while len(self._order_by_camera) > self.camera_cache_size:
                self._order_by_camera.popitem(last=False)
        cap_next = int(D * 1.25) + 4096
        self._d_cap[device.index] = cap_next if (self.optimistic_binning and cap_next < (1 << 31)) else None

        self.last_stats = {"num_visible": num_vis, "num_binned": num_sorted, "tile_pairs": D}
        self._last_debug = {"tile_consumed": tile_consumed, "n_consumed": n_consumed, "entry_ids": entry_ids,
                            "tile_ranges": tile_ranges, "depths": depths, "tiles_touched": tiles_touched,
                            "tile_rect": tile_rect, "depth_keys": depth_keys, "sorted_ids": sorted_ids[:num_sorted]}
        return {
            "image": image,
            "alpha": alpha,
            "depth": depth,
            "viewspace_points": means2d,
            "visibility_filter": vis,
            "radii": radii,
            "conics": conics,
        }
