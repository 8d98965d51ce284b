#!/usr/bin/env python
"""Per-launch CUDA-event times (RTB200_TRACE) of ONE tile shard of the headline frame on one GPU, kernels back to back
(RT_OPT_LANES 0), plus the cost of a frame in which nothing is traced (camera looking away: the launch chain alone)."""
import os
import sys
from pathlib import Path

os.environ["RTB200_TRACE"] = "1"
import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402


def main():
    import torch
    from raytracercpp_b200 import api
    mod = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    tile = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    opts = sys.argv[3] if len(sys.argv) > 3 else ""
    lib = api.load_library()
    lib.rt_set_host_threads(0)
    scene = bench.make_scene("cfg4_sphere10M_4k_16spp")
    ctx = api.Context(0, lib)
    ctx.set_triangles(scene["xyz9"], scene["uv6"], scene["mat"])
    s = api.default_settings(lib, **scene["kw"])

    class A:
        opt = []
    _, (proj_inv, cam, pos) = bench.setup_context(ctx, api, scene, A, leaf_split=8)
    for kv in filter(None, opts.split(",")):
        k, v = kv.split("=")
        ctx.set_option(int(k), int(v))
    frame = torch.zeros((s.image_height, s.image_width), dtype=torch.int32, device="cuda")
    for lanes in (0, 1):
        ctx.set_option(api.RT_OPT_LANES, lanes)
        for _ in range(3):
            st = ctx.render_device(s, frame.data_ptr(), tile, mod, 0)
        print(f"==== shard 0 of {mod}, tile {tile}, lanes {lanes}: device {st.device_ms:.3f} ms, launches {st.kernel_launches}, traced primary {st.traced_primary_rays}, hits {st.primary_hits}", file=sys.stderr, flush=True)
    ctx.set_option(api.RT_OPT_COUNT_WORK, 1)
    ctx.render_device(s, frame.data_ptr(), tile, mod, 0)
    ctx.set_option(api.RT_OPT_COUNT_WORK, 0)
    print("==== whole frame, lanes 0", file=sys.stderr, flush=True)
    ctx.set_option(api.RT_OPT_LANES, 0)
    for _ in range(2):
        st = ctx.render_device(s, frame.data_ptr(), 64, 1, 0)
    print(f"==== whole: device {st.device_ms:.3f} ms", file=sys.stderr, flush=True)
    away = np.diag([-1, 1, -1, 1]).astype(np.float32)
    ctx.set_camera(proj_inv, away, (0, 0, 0))
    ctx.set_option(api.RT_OPT_LANES, 1)
    for _ in range(3):
        st = ctx.render_device(s, frame.data_ptr(), tile, mod, 0)
    print(f"==== nothing on screen (shard 0 of {mod}): device {st.device_ms:.3f} ms, launches {st.kernel_launches}", file=sys.stderr, flush=True)
    ctx.set_option(api.RT_OPT_SCREEN_CULL, 0)
    for _ in range(3):
        st = ctx.render_device(s, frame.data_ptr(), tile, mod, 0)
    print(f"==== nothing on screen, no cull (every packet traced against the root): device {st.device_ms:.3f} ms, launches {st.kernel_launches}", file=sys.stderr, flush=True)


if __name__ == "__main__":
    main()
