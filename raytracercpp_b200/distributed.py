"""Screen-tile sharding of one frame over the GPUs of a box.

Every pixel's ray tree is independent, so the path shards with no data-path exchange: the scene is replicated on
every GPU, the frame is cut into `tile_size`^2 final-resolution tiles dealt round-robin over the ranks
(rtb::owned_tiles), each rank traces, shades and RESOLVES its own tiles.  One process per GPU, `torch.distributed`
is the plumbing; frames are device memory handed to the C ABI as raw pointers.

Gathering the frame (mode "peer", the default for world > 1): rank 0 owns the frame buffers (rt_frame_alloc, two of
them, used alternately) and every other rank maps them over CUDA IPC (rt_frame_open).  A rank renders with the mapped
pointer as its output, so the kernel that produces a final pixel -- k_resolve, or the shading kernels when there is no
SSAA -- stores it straight into rank 0's memory over NVLink: no pack, no collective on the pixel data, no unpack, and no
copy of the frame to the seven ranks that never read it.  What remains is ONE small cross-rank step per frame: an
all-reduce of a single word on the same stream, which orders "every rank's stores have landed" before anything rank 0
enqueues next (its device-to-host copy, the next frame).  Two buffers, because a fast rank may start storing frame k+1
while rank 0 still copies frame k to the host; it cannot get further ahead than that (the barrier of frame k+1 needs
rank 0).

Mode "nccl" is the round-1 path, kept as the fallback when IPC mapping is refused: pack (kernel) -> all_gather over
NCCL -> unpack (one kernel for all peers); every rank ends up with the whole frame.

With backend "gloo" and CPU tensors the same code runs against the host kernel emulation in the CPU tests (its
rt_frame_alloc hands out POSIX shared memory).
"""
from __future__ import annotations

import contextlib
import ctypes

import numpy as np
import torch
import torch.distributed as dist

from . import api

TILE = 64


class _DevicePtr:
    """A raw device allocation as something torch.as_tensor understands (read side of rank 0's frame)."""

    def __init__(self, ptr: int, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<i4", "data": (ptr, False), "version": 3, "strides": None}


def _view(ptr: int, h: int, w: int, device: torch.device) -> torch.Tensor:
    if device.type == "cuda":
        return torch.as_tensor(_DevicePtr(ptr, (h, w)), device=device)
    buf = (ctypes.c_int32 * (h * w)).from_address(ptr)
    return torch.from_numpy(np.ctypeslib.as_array(buf).reshape(h, w))


class ShardedFrame:
    def __init__(self, ctx: api.Context, settings: api.RtSettings, rank: int, world: int, tile_size: int = TILE, device=None,
                 gather: str = "peer"):
        self.ctx, self.s, self.rank, self.world, self.tile = ctx, settings, rank, world, tile_size
        self.device = device if device is not None else torch.device("cuda", ctx.device)
        h, w = settings.image_height, settings.image_width
        self.h, self.w = h, w
        self.mode = gather if world > 1 else "single"
        self._owned, self._mapped = [], []
        self.why_not_peer = None
        if self.mode == "peer":
            self._setup_peer()
        if self.mode != "peer":
            self.frame = torch.zeros((h, w), dtype=torch.int32, device=self.device)      # ARGB32 words
            self.targets = [self.frame.data_ptr()]
        self.flip = 0
        if self.mode == "nccl":
            self.counts = [ctx.tile_count(settings, tile_size, world, r) for r in range(world)]
            self.slot = max(self.counts) * tile_size * tile_size                         # equal-size all_gather slots
            self.staging = torch.zeros(self.slot, dtype=torch.int32, device=self.device)
            self.gathered = torch.zeros(world * self.slot, dtype=torch.int32, device=self.device)
        if world > 1:
            self.sync = torch.zeros(1, dtype=torch.int32, device=self.device)            # operand of the per-frame barrier

    # ---- set-up of the shared frames ---------------------------------------------------------------------------------
    def _setup_peer(self):
        nbytes = self.h * self.w * 4
        handles, err = [None, None], None
        if self.rank == 0:
            try:
                for i in range(2):
                    ptr, hd = self.ctx.frame_alloc(nbytes)
                    self._owned.append(ptr)
                    handles[i] = hd
            except api.RtError as e:
                err = str(e)
        box = [handles, err]
        dist.broadcast_object_list(box, src=0)
        handles, err = box
        ptrs = list(self._owned)
        if err is None and self.rank != 0:
            try:
                for hd in handles:
                    ptrs.append(self.ctx.frame_open(hd))
                self._mapped = list(ptrs)
            except api.RtError as e:
                err = str(e)
        errs = [None] * self.world
        dist.all_gather_object(errs, err)
        bad = [e for e in errs if e]
        if bad:                                                   # some rank cannot map the frame: everybody falls back
            self.close()
            self.mode, self.why_not_peer = "nccl", bad[0]
            return
        self.targets = ptrs
        self.views = [_view(p, self.h, self.w, self.device) for p in ptrs] if self.rank == 0 else None
        self.frame = self.views[0] if self.rank == 0 else None

    def close(self):
        for p in self._mapped:
            self.ctx.frame_close(p)
        for p in self._owned:
            self.ctx.frame_free(p)
        self._mapped, self._owned = [], []

    # ---- one frame ---------------------------------------------------------------------------------------------------
    def render(self) -> api.RtRenderStats:
        """Trace + shade + resolve this rank's tiles (into rank 0's frame in peer mode, else into self.frame)."""
        return self.ctx.render_device(self.s, self.targets[self.flip], self.tile, self.world, self.rank)

    def begin(self) -> int:
        """Enqueues this rank's part of the next frame and the cross-rank step; returns the pointer of the frame buffer that
        will hold the complete frame on rank 0 (peer mode) / on every rank (nccl mode) once the stream gets there."""
        tgt = self.targets[self.flip]
        self.ctx.render_device_begin(self.s, tgt, self.tile, self.world, self.rank)
        if self.mode == "peer":
            dist.all_reduce(self.sync)                            # every rank's stores into rank 0's frame have landed
            if self.rank == 0:
                self.frame = self.views[self.flip]
            self.flip ^= 1
        elif self.mode == "nccl":
            self.gather()
        return tgt

    def end(self) -> api.RtRenderStats:
        return self.ctx.render_device_end()

    def render_and_gather(self) -> api.RtRenderStats:
        """One frame: kernels and the cross-rank step are enqueued back to back on the context's stream; the host waits once."""
        self.begin()
        return self.end()

    def render_to_host(self, out: np.ndarray) -> api.RtRenderStats:
        """The sharded counterpart of rt_render: the complete frame ends up in `out`, an ordinary (pageable) host array of
        rank 0; the other ranks only render.  Everything is stream-ordered; rank 0 returns when `out` is complete."""
        tgt = self.begin()
        if self.rank == 0:
            self.ctx.frame_to_host(tgt, out)
        return self.end()

    def gather(self):
        """nccl mode: all ranks end up with the complete frame."""
        if self.world == 1:
            return self.frame
        self.ctx.pack_tiles(self.s, self.frame.data_ptr(), self.staging.data_ptr(), self.tile, self.world, self.rank)
        dist.all_gather_into_tensor(self.gathered, self.staging)
        # one launch scatters the peers' tiles; pack, collective and unpack are stream-ordered, no host synchronisation
        self.ctx.unpack_gathered(self.s, self.frame.data_ptr(), self.gathered.data_ptr(), self.tile, self.world, self.rank)
        return self.frame


class FramePipeline:
    """Several frames in flight on one GPU: each ShardedFrame has its own context (the scene is resident once per context) on
    its own stream; frame k+1 is enqueued before the host waits for frame k, so the tail of one frame's launches -- a
    persistent kernel's last, longest packets, the short item passes of split packets -- runs beside the next frame's first
    kernels instead of leaving SMs idle.  Throughput of a sequence of frames (an animation, a turntable), not the latency of
    one; the frames are complete and in order.  `streams` are torch streams, streams[i] the one frames[i].ctx runs on."""

    def __init__(self, frames, streams):
        assert len(frames) == len(streams) and frames
        self.frames, self.streams = list(frames), list(streams)
        self.inflight = []
        self.next = 0

    def submit(self):
        """Enqueues the next frame; returns the stats of the frame that had to be waited for to make room (or None)."""
        done = None
        if len(self.inflight) == len(self.frames):
            done = self.frames[self.inflight.pop(0)].end()
        i = self.next
        self.next = (self.next + 1) % len(self.frames)
        with (torch.cuda.stream(self.streams[i]) if self.streams[i] is not None else contextlib.nullcontext()):   # None: CPU tests
            self.frames[i].begin()
        self.inflight.append(i)
        return done

    def drain(self):
        """Waits for the frames still in flight; returns their stats in order."""
        out = [self.frames[i].end() for i in self.inflight]
        self.inflight = []
        return out
