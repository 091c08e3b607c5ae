"""Turn ncu CSV exports (brought back in gpurun_out/) into the small summaries kept under profiles/.

    python profiles/summarize.py launches gpurun_out/launches_fast.csv  > profiles/r01_launches_fast.md
    python profiles/summarize.py raw gpurun_out/raw_decim.csv           > profiles/r01_ncu_decim_fast.md
"""
import collections
import csv
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size",
        "launch__block_size", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        v = v / 1000 if r[ui] == "ns" else v * 1000 if r[ui] == "ms" else v
        agg.setdefault(r[ki].split("(")[0], []).append(v)
    tot = sum(sum(v) for k, v in agg.items() if "tb200" in k)
    print("| kernel | launches | mean us | share of tb200 time |\n|---|---|---|---|")
    for k, v in agg.items():
        share = f"{100 * sum(v) / tot:.1f} %" if "tb200" in k else "(torch fill)"
        print(f"| `{k[:80]}` | {len(v)} | {sum(v) / len(v):.1f} | {share} |")


def raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    for d in rows[2:]:
        print(f"### `{d[ki].split('(')[0]}`\n\n| metric | value | unit |\n|---|---|---|")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"| {k} | {d[i]} | {units[i]} |")
        print()


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2])
