#!/usr/bin/env python
"""Turns one gpurun's ncu output (gpurun_out/) into the tracked summaries under profiles/ (run here, no GPU needed):

    python scripts/make_profiles.py <tag> [gpurun_out/launches.csv] [gpurun_out/prof.ncu-rep]

  profiles/<tag>_launches.csv      per-kernel totals of the `--metrics gpu__time_duration.sum` launch list
  profiles/<tag>_launches_raw.csv  the launch list itself (this library's kernels only)
  profiles/<tag>_ncu_full.txt      the metrics DESIGN.md quotes from the `--set full` capture + hottest SASS lines
  profiles/traffic.json            dram bytes per launch of the packet kernels (bench.py's roofline.traffic)
"""
import csv, io, json, subprocess, sys
from collections import OrderedDict
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
tag = sys.argv[1]
launches = Path(sys.argv[2] if len(sys.argv) > 2 else ROOT / "gpurun_out" / "launches.csv")
rep = Path(sys.argv[3] if len(sys.argv) > 3 else ROOT / "gpurun_out" / "prof.ncu-rep")
out = ROOT / "profiles"
out.mkdir(exist_ok=True)

if launches.exists():
    rows = [r for r in csv.reader(l for l in open(launches) if l.startswith('"'))]
    hdr, body = rows[0], rows[1:]
    k, v = hdr.index("Kernel Name"), hdr.index("Metric Value")
    per = OrderedDict()
    raw = []
    for r in body:
        name = r[k].split("(")[0].replace("void ", "")
        ns = float(r[v].replace(",", ""))
        ours = name.startswith("rtb::") or name.startswith("k_")
        if ours:
            raw.append((r[hdr.index("ID")], name, r[hdr.index("Grid Size")], r[hdr.index("Block Size")], ns))
        d = per.setdefault(name if ours else "(torch / NCCL / memset kernels)", [0, 0.0])
        d[0] += 1; d[1] += ns
    total = sum(d[1] for d in per.values())
    with open(out / f"{tag}_launches.csv", "w") as f:
        f.write("kernel,launches,total_ms,mean_ms,share_of_all_launches\n")
        for name, (n, ns) in sorted(per.items(), key=lambda kv: -kv[1][1]):
            f.write(f"\"{name}\",{n},{ns / 1e6:.3f},{ns / n / 1e6:.3f},{ns / total:.4f}\n")
    with open(out / f"{tag}_launches_raw.csv", "w") as f:
        f.write("id,kernel,grid,block,duration_ns\n")
        for r in raw:
            f.write(",".join(f'"{x}"' if isinstance(x, str) and "," in x else str(x) for x in r) + "\n")
    print(open(out / f"{tag}_launches.csv").read())

WANT = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum']
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
if rep.exists():
    txt = subprocess.run(['ncu', '-i', str(rep), '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    lines, traffic = [], {}                      # traffic: per kernel family, summed over the captured launches (= one frame)
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        kname = d['Kernel Name'].split("(")[0].replace("void ", "")
        lines.append(f"===== {kname}   (ncu --set full --clock-control none, launch id {d.get('ID')})")
        for key in WANT:
            if key in d:
                lines.append(f"  {key} = {d[key]} {units[hdr.index(key)]}")
        for key in hdr:
            if 'warp_issue_stalled' in key and key.endswith('_per_warp_active.pct'):
                try:
                    val = float(d[key])
                except ValueError:
                    continue
                if val > 2:
                    lines.append(f"  stall {key.replace('smsp__warp_issue_stalled_', '').replace('_per_warp_active.pct', '')} = {val:.1f} % of active warps")
        try:
            b = sum(float(d[m].replace(",", "")) * UNIT[units[hdr.index(m)]] for m in ('dram__bytes_read.sum', 'dram__bytes_write.sum'))
            short = "k_primary" if "k_primary" in kname else "k_shade" if "k_shade" in kname else "k_reflect" if "k_reflect" in kname else kname
            t = traffic.setdefault(short, {"dram_bytes": 0, "warp_inst": 0, "launches": 0, "kernel_ms": 0.0})
            t["dram_bytes"] += int(b)
            t["warp_inst"] += int(float(d['smsp__inst_executed.sum'].replace(",", "")))
            t["launches"] += 1
            t["kernel_ms"] += float(d['gpu__time_duration.sum'].replace(",", "")) * {"msecond": 1.0, "usecond": 1e-3, "nsecond": 1e-6, "second": 1e3}.get(units[hdr.index('gpu__time_duration.sum')], 1.0)
        except Exception:
            pass
    # hottest SASS lines per kernel
    src = subprocess.run(['ncu', '-i', str(rep), '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(src)))
    starts = [i for i, r in enumerate(srows) if r and r[0] == 'Address']
    for si, s0 in enumerate(starts):
        h = srows[s0]
        ix = {kk: h.index(kk) for kk in ('Source', '# Samples', 'Instructions Executed', 'Avg. Threads Executed')}
        end = starts[si + 1] if si + 1 < len(starts) else len(srows)
        body = [r for r in srows[s0 + 1:end] if len(r) > ix['# Samples'] and r[ix['# Samples']].isdigit()]
        tot_s = sum(int(r[ix['# Samples']]) for r in body) or 1
        tot_i = sum(int(r[ix['Instructions Executed']]) for r in body) or 1
        lines.append(f"----- source page, kernel #{si}: {len(body)} SASS lines, {tot_s} stall samples, {tot_i} warp instructions; hottest by samples:")
        for r in sorted(body, key=lambda r: -int(r[ix['# Samples']]))[:14]:
            lines.append(f"   {100 * int(r[ix['# Samples']]) / tot_s:5.1f}% smp  {int(r[ix['Instructions Executed']]):11d} exe  {r[ix['Avg. Threads Executed']]:>5s} thr  {r[ix['Source']].strip()[:64]}")
    (out / f"{tag}_ncu_full.txt").write_text("\n".join(lines) + "\n")
    print("\n".join(lines[:70]))
    tf = out / "traffic.json"
    cur = json.loads(tf.read_text()) if tf.exists() else {}
    for t in traffic.values():
        t["source"] = (f"profiles/{tag}_ncu_full.txt: dram__bytes_read.sum + dram__bytes_write.sum and smsp__inst_executed.sum of the {t['launches']} packet-kernel "
                       "launches of one frame (the item passes and finish kernels of the stage are not in the capture)")
    cur["cfg4_sphere10M_4k_16spp"] = traffic
    cur["_source"] = f"profiles/{tag}_ncu_full.txt"
    tf.write_text(json.dumps(cur, indent=1) + "\n")
