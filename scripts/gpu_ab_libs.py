#!/usr/bin/env python
"""A/B of library builds on the headline frame (one GPU): whole-frame device time and stage times (kernels back to back)
for every build given.  usage: python scripts/gpu_ab_libs.py [--workload W] name=path.so ..."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402


def main():
    import torch
    from raytracercpp_b200 import api
    args = sys.argv[1:]
    workload = "cfg4_sphere10M_4k_16spp"
    if args and args[0] == "--workload":
        workload, args = args[1], args[2:]
    scene = bench.make_scene(workload)
    for spec in args:
        name, path = spec.split("=")
        lib = api.load_library(path if path != "default" else None)
        lib.rt_set_host_threads(0)
        ctx = api.Context(0, lib)
        ctx.set_triangles(scene["xyz9"], scene["uv6"], scene["mat"])
        s = api.default_settings(lib, **scene["kw"])

        class A:
            opt = []
        bench.setup_context(ctx, api, scene, A, leaf_split=8)
        frame = torch.zeros((s.image_height, s.image_width), dtype=torch.int32, device="cuda")
        out = {"lib": name}
        for lanes in (1, 0):
            ctx.set_option(api.RT_OPT_LANES, lanes)
            for _ in range(3):
                ctx.render_device(s, frame.data_ptr(), 64, 1, 0)
            ms, st = [], None
            for _ in range(7):
                st = ctx.render_device(s, frame.data_ptr(), 64, 1, 0)
                ms.append(st.device_ms)
            out["lanes%d_ms" % lanes] = round(float(np.median(ms)), 3)
            if lanes == 0:
                out.update(primary=round(st.trace_primary_ms, 3), shade=round(st.shade_ms, 3), compact=round(st.compact_ms, 3), resolve=round(st.resolve_ms, 3))
        out["crc"] = int(frame.sum().item()) & 0xffffffff
        print(json.dumps(out), flush=True)
        ctx.close()


if __name__ == "__main__":
    main()
