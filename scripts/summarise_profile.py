#!/usr/bin/env python
"""Summarise an ncu report (run here, no GPU needed): key raw metrics, stall breakdown and the hottest source lines ->
profiles/<tag>_summary.txt, and DRAM traffic per launch -> profiles/ncu_traffic.json (read by bench.py).
Usage: scripts/summarise_profile.py gpurun_out/prof_X.ncu-rep <tag> <kernel-key e.g. 'mpc_admm_kernel<64,60>'> <source.cu>"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

rep, tag, key, srcfile = sys.argv[1:5]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw[raw.index('"ID"'):])))
hdr, units, vals = rows[0], rows[1], rows[2]
m = dict(zip(hdr, vals))
u = dict(zip(hdr, units))
want = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_static",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum"]
out = [f"ncu summary {tag}  ({os.path.basename(rep)})", ""]
for k in want:
    if k in m:
        out.append(f"{k:70s} {m[k]} {u.get(k, '')}")
out.append("")
out.append("stall reasons (warps stalled per issue-active cycle):")
st = sorted(((float(v or 0), k) for k, v in m.items() if "average_warps_issue_stalled" in k and "per_issue_active" in k), reverse=True)
for v, k in st[:8]:
    out.append(f"  {k.split('issue_stalled_')[1].split('_per_issue')[0]:24s} {v:.3f}")


def to_bytes(v, unit):
    f = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(v) * f.get(unit, 1)


traffic = to_bytes(m["dram__bytes_read.sum"], u["dram__bytes_read.sum"]) + to_bytes(m["dram__bytes_write.sum"], u["dram__bytes_write.sum"])
out.append("")
out.append(f"DRAM traffic per launch: {traffic / 1e6:.2f} MB")
both = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
cur, h = None, None
for r in csv.reader(io.StringIO(both)):
    if not r:
        continue
    if r[0] == "File Path":
        cur = os.path.basename(r[1]); continue
    if r[0] == "Line No":
        h = r; continue
    if r[0] == "Function Name" or h is None:
        continue
    try:
        line = int(r[0])
    except ValueError:
        continue
    if r[2] == "":
        continue
    d = dict(zip(h[2:], r[2:]))
    try:
        s, ie = int(d.get("# Samples") or 0), int(d.get("Instructions Executed") or 0)
    except ValueError:
        continue
    a = agg[(cur, line)]
    a[0] += s; a[1] += ie
    for k, v in d.items():
        if k.startswith("stall_") and "Not Issued" not in k and v not in ("", "0"):
            a[2][k] += int(v)
ts, ti = sum(a[0] for a in agg.values()) or 1, sum(a[1] for a in agg.values()) or 1
src = open(srcfile).read().split("\n")
out.append("")
out.append(f"hottest source lines (warp-state samples, {os.path.basename(srcfile)}):")
for (f, l), a in sorted(agg.items(), key=lambda x: -x[1][0])[:20]:
    top = ", ".join(f"{k[6:]}:{v}" for k, v in a[2].most_common(2))
    txt = src[l - 1].strip()[:70] if f == os.path.basename(srcfile) and l <= len(src) else ""
    out.append(f"  {f}:{l:<5d} samples {100 * a[0] / ts:5.1f}%  inst {100 * a[1] / ti:5.1f}%  [{top}]  {txt}")
path = os.path.join(ROOT, "profiles", f"{tag}_summary.txt")
open(path, "w").write("\n".join(out) + "\n")
tj = os.path.join(ROOT, "profiles", "ncu_traffic.json")
cur_t = json.load(open(tj)) if os.path.exists(tj) else {}
cur_t[key] = {"dram_bytes_per_launch": traffic, "source": os.path.basename(path), "kernel_ms_under_ncu": m.get("gpu__time_duration.sum")}
json.dump(cur_t, open(tj, "w"), indent=1)
print("\n".join(out[:40]))
