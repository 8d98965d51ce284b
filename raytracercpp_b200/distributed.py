"""Screen-tile sharding of one frame over the GPUs of a box.

Every pixel's ray tree is independent, so the path shards with no data-path exchange: the scene is replicated on
every GPU, the frame is cut into `tile_size`^2 final-resolution tiles dealt round-robin over the ranks
(rtb::owned_tiles), each rank traces, shades and RESOLVES its own tiles, and the only collective is the gather of
the final ARGB32 tiles: pack (kernel) -> all_gather over NCCL/NVLink -> unpack (one kernel for all peers).  One process per GPU,
`torch.distributed` is the plumbing; the tensors are only device memory handed to the C ABI as raw pointers.

With backend "gloo" and CPU tensors the same code runs against the host kernel emulation in the CPU tests.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import api

TILE = 64


class ShardedFrame:
    def __init__(self, ctx: api.Context, settings: api.RtSettings, rank: int, world: int, tile_size: int = TILE, device=None):
        self.ctx, self.s, self.rank, self.world, self.tile = ctx, settings, rank, world, tile_size
        self.device = device if device is not None else torch.device("cuda", ctx.device)
        h, w = settings.image_height, settings.image_width
        self.frame = torch.zeros((h, w), dtype=torch.int32, device=self.device)          # ARGB32 words
        self.counts = [ctx.tile_count(settings, tile_size, world, r) for r in range(world)]
        per = tile_size * tile_size
        self.slot = max(self.counts) * per                                               # equal-size all_gather slots
        self.staging = torch.zeros(self.slot, dtype=torch.int32, device=self.device)
        self.gathered = torch.zeros(world * self.slot, dtype=torch.int32, device=self.device) if world > 1 else None

    def render(self) -> api.RtRenderStats:
        """Trace + shade + resolve this rank's tiles into self.frame."""
        return self.ctx.render_device(self.s, self.frame.data_ptr(), self.tile, self.world, self.rank)

    def render_and_gather(self) -> api.RtRenderStats:
        """One frame, all ranks end up with all of it: the frame's kernels, pack, all-gather and unpack are enqueued back
        to back on the context's stream; the host waits once, at the end."""
        self.ctx.render_device_begin(self.s, self.frame.data_ptr(), self.tile, self.world, self.rank)
        self.gather()
        return self.ctx.render_device_end()

    def gather(self):
        """All ranks end up with the complete frame."""
        if self.world == 1:
            return self.frame
        self.ctx.pack_tiles(self.s, self.frame.data_ptr(), self.staging.data_ptr(), self.tile, self.world, self.rank)
        dist.all_gather_into_tensor(self.gathered, self.staging)
        # one launch scatters the peers' tiles; pack, collective and unpack are stream-ordered, no host synchronisation
        self.ctx.unpack_gathered(self.s, self.frame.data_ptr(), self.gathered.data_ptr(), self.tile, self.world, self.rank)
        return self.frame
