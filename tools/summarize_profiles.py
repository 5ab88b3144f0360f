#!/usr/bin/env python
"""Turns the raw ncu output in gpurun_out/ into the tracked summaries under profiles/:
   profiles/<tag>_launches*.csv  (copied), profiles/<tag>_kernels.csv (key metrics per captured kernel),
   profiles/<tag>_summary.md, profiles/hidden_kernel_traffic.json (read by bench.py for roofline.traffic)."""
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
OUT, PROF = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(PROF, exist_ok=True)

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
]


def raw(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


kernels = []
for name in ("fused", "hidden_throughput", "hidden", "input", "fixup", "fixup_stream", "hidden_stream", "output_stream"):
    rep = os.path.join(OUT, f"{tag}_{name}.ncu-rep")
    if not os.path.exists(rep):
        continue
    hdr, units, rows = raw(rep)
    for r in rows:
        d = {"capture": name, "kernel": r[hdr.index("Kernel Name")][:80]}
        for k in KEYS:
            if k in hdr:
                d[k] = r[hdr.index(k)] + " " + units[hdr.index(k)]
        kernels.append(d)

with open(os.path.join(PROF, f"{tag}_kernels.csv"), "w", newline="") as f:
    w = csv.DictWriter(f, fieldnames=["capture", "kernel"] + KEYS)
    w.writeheader()
    for d in kernels:
        w.writerow(d)

for suffix in ("launches.csv", "launches_warm.csv", "launches_latency.csv", "launches_throughput.csv", "launches_throughput_warm.csv", "launches_stream.csv",
               "bench.json", "bench_reference.json", "stage_times.log"):
    src = os.path.join(OUT, f"{tag}_{suffix}")
    if os.path.exists(src):
        shutil.copy(src, os.path.join(PROF, f"{tag}_{suffix}"))


def launch_table(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = {}
    for r in rows[1:]:
        k = r[ki].split("(")[0].replace("void ", "").replace("fdnn::<unnamed>::", "")
        agg.setdefault(k, []).append(float(r[vi].replace(",", "")))
    return agg


lines = [f"# ncu summary `{tag}` (B200, `tools/make_profiles.sh`; all kernels of two passes at batch 512; one context, one stream)\n"]
for suffix, title in (("launches.csv", "cold caches (ncu flushes L2 before every kernel; serialised)"),
                      ("launches_warm.csv", "`--cache-control none` (weights L2-resident, as in the real pass)")):
    p = os.path.join(OUT, f"{tag}_{suffix}")
    if not os.path.exists(p):
        continue
    agg = launch_table(p)
    total = sum(sum(v) for v in agg.values())
    lines.append(f"\n## Launch list, {title}\n\n| kernel | launches | mean us | share of pass |\n|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        lines.append(f"| `{k}` | {len(v)} | {sum(v) / len(v) / 1e3:.2f} | {100 * sum(v) / total:.1f} % |")
    lines.append(f"\npass total (sum of kernel durations): {total / 2 / 1e3:.1f} us")

lines.append("\n## Full captures (`ncu --set full --clock-control none`), key metrics\n")
for d in kernels:
    lines.append(f"\n### {d['capture']}: `{d['kernel']}`\n")
    for k in KEYS:
        if k in d:
            lines.append(f"* `{k}` = {d[k]}")

hid = [d for d in kernels if d["capture"] == "hidden"]
if hid:
    def num(s):
        v, u = s.split(" ", 1)
        v = float(v.replace(",", ""))
        return v * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1}.get(u.strip(), 1)
    t = num(hid[-1]["dram__bytes_read.sum"]) + num(hid[-1]["dram__bytes_write.sum"])
    json.dump({"dram_bytes_per_launch": t, "source": f"profiles/{tag}_kernels.csv (ncu --set full, cold L2: weights + activations come from HBM once)",
               "algorithmic_bytes_per_launch": 2048 * 2048 + 2 * 512 * 2048}, open(os.path.join(PROF, "hidden_kernel_traffic.json"), "w"), indent=1)
open(os.path.join(PROF, f"{tag}_summary.md"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:40]))
