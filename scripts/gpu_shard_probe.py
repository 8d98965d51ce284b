#!/usr/bin/env python
"""One GPU, the headline workload: what ONE rank of an N-way tile-sharded frame costs (tile_mod = N, each tile_rem in
turn), against the whole frame on the same GPU -- the kernel side of the strong-scaling efficiency without needing N GPUs.
Sweeps library options given as "id=value,id=value" strings.

    python scripts/gpu_shard_probe.py --mod 8 --tile 32 --sets "" "9=3" "9=4" "7=64,10=16,11=64"
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
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--sets", nargs="*", default=[""])
    ap.add_argument("--lib", default=None)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import torch
    from raytracercpp_b200 import api
    lib = api.load_library(args.lib)
    lib.rt_set_host_threads(0)
    scene = bench.make_scene(args.workload)
    ctx = api.Context(0, lib)
    ctx.set_triangles(scene["xyz9"], scene["uv6"], scene["mat"])
    s = api.default_settings(lib, **scene["kw"])

    class A:
        opt = []
    bench.setup_context(ctx, api, scene, A, leaf_split=8)
    frame = torch.zeros((s.image_height, s.image_width), dtype=torch.int32, device="cuda")
    results = []
    defaults = {api.RT_OPT_LANES: 1, api.RT_OPT_PACKET_ROUNDS: -256, api.RT_OPT_ITEM_ROUNDS: -16, api.RT_OPT_PRIMARY_ROUNDS: -256, api.RT_OPT_ITEM_PASSES: 6, api.RT_OPT_GRAPH: 1}
    for optset in args.sets:
        for k, v in defaults.items():
            ctx.set_option(k, v)
        for kv in filter(None, optset.split(",")):
            k, v = kv.split("=")
            ctx.set_option(int(k), int(v))

        def timed(tile, mod, rem):
            for _ in range(2):
                ctx.render_device(s, frame.data_ptr(), tile, mod, rem)
            ms = []
            for _ in range(args.reps):
                st = ctx.render_device(s, frame.data_ptr(), tile, mod, rem)
                ms.append(st.device_ms)
            return float(np.median(ms)), st
        whole, st = timed(64, 1, 0)
        whole32, _ = timed(args.tile, 1, 0)
        shards = [timed(args.tile, args.mod, r)[0] for r in range(args.mod)]
        eff = whole / (args.mod * max(shards))
        row = {"opts": optset, "whole_ms": whole, "whole_tile%d_ms" % args.tile: whole32, "shard_ms": [round(x, 3) for x in shards],
               "shard_max_ms": max(shards), "shard_mean_ms": float(np.mean(shards)), "kernel_side_efficiency": eff, "launches": st.kernel_launches}
        results.append(row)
        print(json.dumps(row), flush=True)
    if args.out:
        Path(args.out).write_text(json.dumps(results, indent=1))


if __name__ == "__main__":
    main()
