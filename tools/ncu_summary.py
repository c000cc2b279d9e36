#!/usr/bin/env python
"""Condense an .ncu-rep (ncu --set full --import-source on) into the text summary kept under profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep "title" units_per_launch > profiles/xxx.txt
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.avg", "smsp__thread_inst_executed_per_inst_executed.ratio",
]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, title = sys.argv[1], sys.argv[2]
    units = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
    raw = page(rep, "raw")
    hdr, unit, val = raw[0], raw[1], raw[2]
    d = {h: (v, u) for h, v, u in zip(hdr, val, unit)}
    print(f"# {title}")
    print(f"# source: {rep}  (ncu --set full --clock-control none --import-source on; one launch, replayed)")
    print(f"kernel: {d.get('Kernel Name', ('?', ''))[0]}")
    for k in KEYS:
        if k in d:
            print(f"{k:72s} {d[k][0]:>18s} {d[k][1]}")
    stalls = sorted(((float(v), h) for h, (v, u) in d.items() if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")), reverse=True)
    print("\nwarp stall reasons (avg warps stalled per issued instruction):")
    for v, h in stalls[:8]:
        print(f"  {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:28s} {v:6.3f}")
    src = page(rep, "source")
    if len(src) > 2:
        h2 = src[1]
        ia, isrc = h2.index("Instructions Executed"), h2.index("Source")
        byop, tot = collections.Counter(), 0
        for r in src[2:]:
            try:
                n = int(r[ia])
            except Exception:
                continue
            t = r[isrc].split()
            op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
            byop[op] += n
            tot += n
        print(f"\nexecuted warp instructions: {tot}" + (f"  ({tot / units:.1f} per unit, unit = one warp of 32 pixels)" if units else ""))
        for op, n in byop.most_common(14):
            print(f"  {op:10s} {n:14d}" + (f" {n / units:9.1f}/unit" if units else ""))


if __name__ == "__main__":
    main()
