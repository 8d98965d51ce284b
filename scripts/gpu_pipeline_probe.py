#!/usr/bin/env python
"""One GPU: frames per second with ONE frame at a time against TWO frames in flight (two contexts on two streams, frame k+1
enqueued before frame k is waited for), for the whole frame and for one rank's share of an N-way tile-sharded frame.

    python scripts/gpu_pipeline_probe.py --mod 8 --tile 32
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg4_sphere10M_4k_16spp")
    ap.add_argument("--mod", type=int, default=8)
    ap.add_argument("--tile", type=int, default=32)
    ap.add_argument("--frames", type=int, default=24)
    ap.add_argument("--contexts", type=int, default=2)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import torch
    from raytracercpp_b200 import api
    lib = api.load_library()
    lib.rt_set_host_threads(0)
    scene = bench.make_scene(args.workload)
    s = api.default_settings(lib, **scene["kw"])

    class A:
        opt = []
    ctxs, frames = [], []
    for i in range(args.contexts):
        ctx = api.Context(0, lib)
        ctx.set_triangles(scene["xyz9"], scene["uv6"], scene["mat"])
        bench.setup_context(ctx, api, scene, A, leaf_split=8)
        ctxs.append(ctx)
        frames.append(torch.zeros((s.image_height, s.image_width), dtype=torch.int32, device="cuda"))

    def run(n_ctx, tile, mod, rem):
        def one_pass(count):
            started = []
            for k in range(count):
                i = k % n_ctx
                if len(started) == n_ctx:
                    ctxs[started.pop(0)].render_device_end()
                ctxs[i].render_device_begin(s, frames[i].data_ptr(), tile, mod, rem)
                started.append(i)
            for i in started:
                ctxs[i].render_device_end()
        one_pass(4)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        one_pass(args.frames)
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) * 1e3 / args.frames

    rows = []
    for label, tile, mod, rems in (("whole", 64, 1, [0]), ("shard", args.tile, args.mod, list(range(args.mod)))):
        for n_ctx in range(1, args.contexts + 1):
            ms = [run(n_ctx, tile, mod, r) for r in rems]
            row = {"what": label, "frames_in_flight": n_ctx, "ms_per_frame": [round(x, 3) for x in ms], "max": max(ms), "mean": float(np.mean(ms))}
            rows.append(row)
            print(json.dumps(row), flush=True)
    same = bool(torch.equal(frames[0], frames[-1]))
    print(json.dumps({"frames_identical": same}))
    if args.out:
        Path(args.out).write_text(json.dumps(rows, indent=1))


if __name__ == "__main__":
    main()
