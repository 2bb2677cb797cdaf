#!/usr/bin/env python3
"""Summarise ncu outputs (run here, no GPU needed) into profiles/:
   tools/ncu_summary.py launches gpurun_out/launches.csv profiles/NAME.md     per-kernel share of the ncu launch list
   tools/ncu_summary.py full gpurun_out/prof.ncu-rep profiles/NAME.md         key metrics of an `ncu --set full` capture
"""
import csv, io, re, subprocess, sys
from collections import OrderedDict

def launches(src, dst):
    rows = [r for r in csv.reader(open(src, errors="ignore")) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    ui = hdr.index("Metric Unit")
    agg = OrderedDict(); total = 0.0; n = 0
    for r in rows:
        if r is hdr or len(r) <= vi or r[mi] != "gpu__time_duration.sum": continue
        v = float(r[vi].replace(",", "")); u = r[ui]
        v = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)   # -> us
        name = re.sub(r"\(.*", "", r[ki]); name = re.sub(r"lf::|<lf::GoldilocksRing(, )?|>$", "", name)
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v; total += v; n += 1
    with open(dst, "w") as f:
        f.write(f"# ncu launch list summary ({src}): {n} launches, {total/1e3:.2f} ms of kernel time (cold-cache, serialised: compare shares)\n\n")
        f.write("| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {k} | {c} | {t/1e3:.3f} | {100*t/total:.1f}% |\n")
    print(open(dst).read())

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_selected",
        "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "smsp__pcsamp_warps_issue_stalled_not_selected", "smsp__pcsamp_warps_issue_stalled_dispatch_stall",
        "smsp__pcsamp_warps_issue_stalled_short_scoreboard", "smsp__pcsamp_warps_issue_stalled_barrier", "smsp__pcsamp_warps_issue_stalled_lg_throttle"]

def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full capture: {src}\n")
        for r in rows[2:]:
            f.write(f"\n## {r[hdr.index('Kernel Name')][:120]}  (launch id {r[hdr.index('ID')]})\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in KEYS:
                if k in hdr: f.write(f"| {k} | {r[hdr.index(k)]} | {units[hdr.index(k)]} |\n")
    print(open(dst).read())

if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
