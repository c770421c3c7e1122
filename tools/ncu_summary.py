#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU): key counters and the hottest source lines.

usage: tools/ncu_summary.py gpurun_out/x.ncu-rep [--top 25]
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
    "sm__cycles_elapsed.avg", "smsp__cycles_active.avg",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__inst_executed.sum", "smsp__inst_executed_pipe_fp64.sum",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 25
    raw = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = raw[0], raw[1]
    for row in raw[2:]:
        d = dict(zip(hdr, row))
        print("## kernel:", d.get("Kernel Name", "?")[:100])
        for k in KEYS:
            if k in d:
                print(f"{k:85s} {d[k]:>16s} {units[hdr.index(k)]}")
    src = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv", "--print-source", "cuda"]))))
    if len(src) > 2:
        h = src[0]
        try:
            i_src = h.index("Source")
            i_smp = h.index("# Samples") if "# Samples" in h else h.index("Sampling Data (All)")
        except ValueError:
            print("source columns:", h[:12])
            return
        rows = []
        for r in src[1:]:
            try:
                rows.append((float(r[i_smp]), r[0], r[i_src].strip()))
            except (ValueError, IndexError):
                continue
        tot = sum(x[0] for x in rows) or 1
        print(f"\n## hottest source lines (samples, total {tot:.0f})")
        for smp, ln, text in sorted(rows, reverse=True)[:top]:
            print(f"{100 * smp / tot:5.1f}%  L{ln:>4s}  {text[:110]}")


if __name__ == "__main__":
    main()
