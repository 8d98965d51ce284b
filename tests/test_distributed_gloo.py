"""The N>1 host logic on the CPU: gloo ranks, CPU tensors, the host kernel emulation standing in for the device.  Both
gather modes of ShardedFrame: "peer" (every rank stores its tiles straight into rank 0's shared frame -- CUDA IPC on the
GPU, POSIX shared memory here -- plus one barrier per frame) and "nccl" (pack -> all-gather -> unpack).  The same class
runs on NCCL in bench.py."""
import ctypes
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, str(ROOT))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from raytracercpp_b200 import api
    from raytracercpp_b200.distributed import ShardedFrame
    from tests import common
    lib = api.bind(ctypes.CDLL(str(ROOT / "tests" / "hostsim" / "librtb200_hostsim.so")))
    z = np.load(ROOT / "tests" / "golden" / "robot_scene.npz")
    rows = z["materials"]
    mats = [dict(ambient_coeff=tuple(r[0:3]), diffuse=tuple(r[3:6]), specular=tuple(r[6:9]), emission=tuple(r[9:12]),
                 reflection=float(r[12]), roughness=float(r[13]), ns=float(r[14]), specular_threshold=float(r[15])) for r in rows]
    robot = dict(xyz9=z["xyz9"], uv6=z["uv6"], mat=z["mat"])
    kw, m, tex = common.config_table(mats)["cfg2"]
    kw = dict(kw, image_width=150, image_height=84)            # not a multiple of the tile size: ragged edge tiles
    r = common.product_renderer(lib, robot, kw, m, tex)
    # round-1 path, kept as the fallback: pack -> all_gather -> unpack, every rank ends up with the frame
    frame = ShardedFrame(r.ctx, r.render_settings(), rank, world, tile_size=16, device=torch.device("cpu"), gather="nccl")
    stats = frame.render()
    full = frame.gather().numpy().view(np.uint32).copy()
    frame.frame.zero_()
    stats2 = frame.render_and_gather()                       # the begin / cross-rank step / end form used by bench.py
    assert np.array_equal(frame.frame.numpy().view(np.uint32), full) and stats2.primary_hits == stats.primary_hits
    np.save(Path(out_dir) / f"frame_{rank}.npy", full)
    np.save(Path(out_dir) / f"rays_{rank}.npy", np.array([stats.primary_rays, stats.primary_hits, stats.traced_primary_rays]))
    # the fused path: every rank stores its tiles straight into rank 0's (shared) frame, one barrier per frame, two
    # frames used alternately; rank 0 ends up with the complete frame, through render_to_host in a pageable array
    peer = ShardedFrame(r.ctx, r.render_settings(), rank, world, tile_size=16, device=torch.device("cpu"), gather="peer")
    assert peer.mode == "peer", peer.why_not_peer
    host = np.zeros((84, 150), np.uint32)
    for it in range(3):                                      # three frames: both buffers, and the first one again
        if rank == 0:
            peer.views[peer.flip].zero_()
        dist.barrier()
        st3 = peer.render_to_host(host)
        assert st3.primary_hits == stats.primary_hits
        if rank == 0:
            assert np.array_equal(host, full), f"peer-store frame {it} differs"
            assert np.array_equal(peer.frame.numpy().view(np.uint32), full)
        dist.barrier()
    peer.close()
    # two frames in flight (FramePipeline, as bench.py's timed loop): two contexts, each with its own shared frames; every
    # frame that comes out is the complete frame
    from raytracercpp_b200.distributed import FramePipeline
    r2 = common.product_renderer(lib, robot, kw, m, tex)
    fa = ShardedFrame(r.ctx, r.render_settings(), rank, world, tile_size=16, device=torch.device("cpu"), gather="peer")
    fb = ShardedFrame(r2.ctx, r2.render_settings(), rank, world, tile_size=16, device=torch.device("cpu"), gather="peer")
    pipe = FramePipeline([fa, fb], [None, None])
    done = [x for x in (pipe.submit() for _ in range(5)) if x is not None] + pipe.drain()
    assert len(done) == 5 and all(d.primary_hits == stats.primary_hits for d in done)
    dist.barrier()
    if rank == 0:
        for f in (fa, fb):
            for v in f.views:
                assert np.array_equal(v.numpy().view(np.uint32), full)
    dist.barrier()
    fa.close(); fb.close()
    r2.close()
    if rank == 0:
        r.ray_trace()
        np.save(Path(out_dir) / "single.npy", r.get_image())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_frame_equals_single_frame(hostsim_lib, tmp_path, world):
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    single = np.load(tmp_path / "single.npy")
    rays = 0
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"frame_{r}.npy"), single), f"rank {r} frame differs from the 1-rank frame"
        rays += int(np.load(tmp_path / f"rays_{r}.npy")[0])
    assert rays == 150 * 84 * 4
