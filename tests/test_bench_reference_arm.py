"""bench.py --impl reference (the reference's own CPU tracer on the host cores) on a small workload: exactly one JSON
line on stdout with the keys of the bench contract."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def test_reference_arm_prints_one_json_line():
    import os
    # as under torchrun with nproc-per-node > 1: OMP_NUM_THREADS=1 in the environment must not starve the CPU arm
    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE="2")
    p = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--workload", "sphere1M_1080p_4spp", "--steps", "1",
                        "--warmup", "1", "--gpus", "2"], capture_output=True, text=True, timeout=600, cwd=str(ROOT), env=env)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, p.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Mrays/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["config"]["workload"] == "sphere1M_1080p_4spp" and d["n_gpus"] == 2 and d["steps"] == 1
    cores = len(os.sched_getaffinity(0))
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] == cores and d["cpu_baseline"]["value"] == d["value"]
    if cores > 1:
        assert d["ms_per_step"] * 1e-3 < 0.6 * single_thread_seconds(d), "the reference arm ran on one thread"
    assert d["e2e"] == {"value": d["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def single_thread_seconds(d):
    """What the same sample costs on ONE thread, from the line's own numbers and a one-thread run of a slice of it."""
    import numpy as np
    sys.path.insert(0, str(ROOT))
    import bench
    scene = bench.make_scene("sphere1M_1080p_4spp")
    _, _, rays, ms, _, _ = bench.reference_sample(scene, row_step=64, threads=1)
    rays_step = int(d["cpu_baseline"]["sample"].split("(")[1].split()[0])
    return ms * 1e-3 * rays_step / rays


def test_reference_arm_other_ranks_do_nothing():
    import os
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    p = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True, text=True, timeout=120,
                       cwd=str(ROOT), env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""
