"""Multi-view step: views sharded over ranks, Gaussians replicated, one gradient/statistics
reduction per step.

The reference has no multi-GPU code; this is the caller the render path gets when a batch of
cameras is split over the GPUs of one box.  Each rank renders its views through the ordinary
single-GPU pipeline.  Parameter gradients and densification statistics accumulate directly into ONE
flat fp32 buffer laid out as

    [ xyz 3N | features_dc 3N | scaling 3N | rotation 4N | opacity N | grad-norm sum N | visible count N ][ max radii N ]

(every segment 16-byte aligned) -- written by gs_project_bwd itself when the renderer is the B200 one
(`GaussianRenderer.accumulate_into`), through autograd's in-place `.grad` accumulation otherwise -- so the
exchange is one SUM over the first 16 floats per splat plus one MAX over the screen radii, with no packing
kernel: a single peer-memory kernel over NVLink on a CUDA process group (`gs_peer_allreduce`), two
all-reduces on any other torch.distributed backend (gloo in the CPU tests, where a stand-in renderer
supplies the per-view outputs).
"""
from __future__ import annotations

import os
from typing import Callable, Dict, List, Optional, Sequence

import torch
import torch.distributed as dist

PARAM_ORDER = ("_xyz", "_features_dc", "_scaling", "_rotation", "_opacity")


def shard_views(num_views: int, rank: int, world_size: int) -> List[int]:
    """Contiguous block partition of view indices; the first `num_views % world_size` ranks take one extra."""
    base, extra = divmod(num_views, world_size)
    start = rank * base + min(rank, extra)
    return list(range(start, start + base + (1 if rank < extra else 0)))


class FlatGradBuffer:
    """One flat fp32 buffer whose slices are installed as the parameters' `.grad`.

    `peer=True` (default when a CUDA process group with more than one rank is up; `GSPLAT_B200_PEER=0`
    turns it off) places the buffer -- and the max-radii vector behind it -- in symmetric memory that
    every rank of the node maps, and `all_reduce` then runs as ONE hand-written kernel per rank over NVLink
    peer loads/stores (`gs_peer_allreduce`) between two device-side barriers, instead of two NCCL
    collectives.  If symmetric memory cannot be set up the buffer falls back to NCCL and records why in
    `peer_error`."""

    def __init__(self, model, peer: Optional[bool] = None, group=None):
        self.model = model
        params = [getattr(model, name) for name in PARAM_ORDER]
        n = params[0].shape[0]
        sizes = [p.numel() for p in params]
        device = params[0].device
        self.n = n
        # every segment starts on a 16-byte boundary (the kernels accumulate into them with float4 accesses)
        pad4 = lambda k: (k + 3) // 4 * 4  # noqa: E731
        offs, off = [], 0
        for sz in sizes:
            offs.append(off)
            off = pad4(off + sz)
        self.param_elems = off
        self.sum_elems = self.param_elems + 2 * pad4(n)
        self.max_elems = pad4(n)
        self.peer = None
        self.peer_error = None
        self.fresh = False
        want_peer = peer if peer is not None else os.environ.get("GSPLAT_B200_PEER", "1") != "0"
        storage = None
        if (want_peer and device.type == "cuda" and dist.is_available() and dist.is_initialized()
                and dist.get_world_size(group) > 1 and dist.get_backend(group) == "nccl"):
            try:
                storage = self._symmetric_storage(device, group)
            except Exception as e:                      # NCCL remains available: not a silent CPU path
                self.peer_error = f"{type(e).__name__}: {e}"
                if peer:
                    raise
        if storage is not None:
            # One exchange of the still-zero buffer now, at construction: it goes through torch's private symmetric-memory
            # API (barrier signature, buffer_ptrs) and our kernel, so an API drift or a launch failure shows up here --
            # where falling back to NCCL is clean -- and not in the middle of a training step.  All ranks agree on the
            # outcome with one MIN all-reduce.
            self.storage = storage
            ok = 1
            try:
                self._peer_exchange()
                torch.cuda.current_stream(device).synchronize()
                if float(storage.abs().max()) != 0.0:
                    raise RuntimeError("probe exchange of a zero buffer returned non-zero values")
            except Exception as e:                               # noqa: BLE001 -- recorded in peer_error, NCCL takes over
                ok, self.peer_error = 0, f"{type(e).__name__}: {e}"
            flag = torch.tensor([ok], dtype=torch.int32, device=device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            if int(flag.item()) != 1:
                if self.peer_error is None:
                    self.peer_error = "peer exchange probe failed on another rank"
                if peer:
                    raise RuntimeError(self.peer_error)
                self.peer, storage = None, None
        if storage is None:
            storage = torch.zeros(self.sum_elems + self.max_elems, dtype=torch.float32, device=device)
        self.storage = storage
        self.flat = storage[:self.sum_elems]
        self.max_radii = storage[self.sum_elems:self.sum_elems + n]
        self.views = [self.flat[o:o + sz].view_as(p) for p, o, sz in zip(params, offs, sizes)]
        self.grad_norm_sum = self.flat[off:off + n]
        self.vis_count = self.flat[off + pad4(n):off + pad4(n) + n]

    def _symmetric_storage(self, device, group):
        import torch.distributed._symmetric_memory as symm_mem
        t = symm_mem.empty(self.sum_elems + self.max_elems, dtype=torch.float32, device=device)
        t.zero_()
        handle = symm_mem.rendezvous(t, group if group is not None else dist.group.WORLD)
        ptrs = [int(p) for p in handle.buffer_ptrs]
        if len(ptrs) != dist.get_world_size(group) or any(p == 0 for p in ptrs):
            raise RuntimeError(f"symmetric memory rendezvous returned {ptrs}")
        import ctypes
        # NVSwitch multicast (multimem.ld_reduce / multimem.st: the switch reduces and replicates, ~36 % fewer bytes per link
        # at 8 GPUs).  Measured (profiles/r2_multigpu.md): 0.148 ms against 0.187 ms (TMA) / 0.208 ms (loads/stores) at 8 GPUs
        # with 16 reductions in flight per thread, but 0.19 against 0.108 ms at 2 GPUs, where it saves no bytes -- default
        # from 8 ranks up; GSPLAT_B200_MULTICAST=0/1 overrides
        mc = 0
        mc_env = os.environ.get("GSPLAT_B200_MULTICAST", "")
        if mc_env == "1" or (mc_env != "0" and dist.get_world_size(group) >= 8):
            try:
                mc = int(handle.multicast_ptr or 0)
            except Exception:
                mc = 0
        # GS_PEER_TMA: bulk asynchronous copies (TMA) instead of per-thread loads/stores.  Measured on this pool's boxes
        # (profiles/r2_multigpu.md): identical at 2 GPUs (both sit at the links' practical rate, ~630 GB/s per direction), 10 %
        # faster at 8 (0.188 vs 0.210 ms) -- default from 4 ranks up; GSPLAT_B200_PEER_TMA=0/1 overrides
        tma_env = os.environ.get("GSPLAT_B200_PEER_TMA", "")
        flags = int(tma_env == "1") if tma_env in ("0", "1") else int(len(ptrs) >= 4)
        self.peer = {"handle": handle, "rank": int(handle.rank), "world": int(handle.world_size),
                     "ptrs": (ctypes.c_uint64 * len(ptrs))(*ptrs), "multicast": mc, "flags": flags}
        return t

    def install(self, zero: bool = True) -> None:
        """Point the parameters' `.grad` at the buffer.  `zero=False` when the first view of the step is going to
        overwrite every element anyway (gs_project_bwd in write mode: `fresh`)."""
        if zero:
            self.storage.zero_()
        self.fresh = not zero
        for name, v in zip(PARAM_ORDER, self.views):
            getattr(self.model, name).grad = v
        rest = getattr(self.model, "_features_rest", None)
        if rest is not None:
            rest.grad = None           # DC-only colour: identically zero, never exchanged

    def add_view_stats(self, viewspace_grad: torch.Tensor, visibility: torch.Tensor, radii: torch.Tensor) -> None:
        vis_f = visibility.to(torch.float32)
        self.grad_norm_sum += viewspace_grad.norm(dim=-1) * vis_f
        self.vis_count += vis_f
        torch.maximum(self.max_radii, radii * vis_f, out=self.max_radii)

    def _peer_exchange(self) -> None:
        import ctypes
        from . import _lib
        h = self.peer["handle"]
        stream = ctypes.c_void_p(torch.cuda.current_stream(self.storage.device).cuda_stream)
        h.barrier(channel=0)                 # every rank's buffer is complete (device-side, on this stream)
        _lib.check(_lib.load().gs_peer_allreduce(self.peer["ptrs"], self.peer["multicast"], self.peer["world"], self.peer["rank"],
                                                 0, self.sum_elems, self.sum_elems, self.max_elems, self.peer["flags"], stream),
                   "gs_peer_allreduce")
        h.barrier(channel=1)                 # every rank's slice has landed everywhere

    def all_reduce(self, group=None) -> None:
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1):
            return
        if self.fresh:
            # the write-mode projection backward that was to overwrite the buffer never ran (loss_fn raised, or no
            # view took the fused path): what the buffer holds is the previous step's result, not this step's
            raise RuntimeError("FlatGradBuffer.all_reduce: no backward pass has filled the buffer since install(zero=False)")
        if self.peer is not None:
            self._peer_exchange()
            return
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(self.max_radii, op=dist.ReduceOp.MAX, group=group)


def multiview_step(model, renderer, cameras: Sequence, settings, loss_fn: Callable[[Dict[str, torch.Tensor], int], torch.Tensor],
                   view_ids: Optional[Sequence[int]] = None, buffer: Optional[FlatGradBuffer] = None,
                   group=None, reduce: bool = True) -> Dict[str, object]:
    """Render this rank's views, back-propagate `loss_fn(out, view_id)` for each, reduce.

    After the call every rank holds, in `model.<param>.grad`, the sum over ALL views of all ranks,
    and in `buffer.grad_norm_sum / vis_count / max_radii` the densification statistics the
    reference allocates but never fills (gaussian_model.py:29-31).
    """
    buf = buffer if buffer is not None else FlatGradBuffer(model)
    if hasattr(model, "get_num_points") and buf.n != model.get_num_points():
        raise ValueError(f"FlatGradBuffer was built for {buf.n} splats but the model now holds {model.get_num_points()} "
                         "(densification changes the row count): build a new FlatGradBuffer(model)")
    ids = list(view_ids) if view_ids is not None else list(range(len(cameras)))
    losses = []
    # B200 renderer: the projection backward adds gradients and statistics into `buf` itself
    # (gs_project_bwd accumulate + stat_*); any other renderer goes through autograd accumulation.
    fused = (hasattr(renderer, "accumulate_into") and getattr(renderer, "sh_degree", 0) == 0
             and buf.flat.is_cuda and getattr(model, "get_num_points", lambda: 0)() == buf.n)
    # fused path with at least one view: the first view's backward writes the whole buffer, so it is not zeroed
    buf.install(zero=not (fused and len(ids) > 0))
    if fused:
        with renderer.accumulate_into(buf):
            for cam, vid in zip(cameras, ids):
                out = renderer.render(cam, model, settings)
                loss = loss_fn(out, vid)
                loss.backward()
                losses.append(loss.detach())
    else:
        for cam, vid in zip(cameras, ids):
            out = renderer.render(cam, model, settings)
            out["viewspace_points"].retain_grad()
            loss = loss_fn(out, vid)
            loss.backward()
            with torch.no_grad():
                buf.add_view_stats(out["viewspace_points"].grad, out["visibility_filter"], out["radii"].detach())
            losses.append(loss.detach())
    if reduce:
        buf.all_reduce(group)
    return {"buffer": buf, "losses": losses}
