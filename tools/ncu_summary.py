#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here, no GPU needed): headline metrics + hottest source lines.

usage: tools/ncu_summary.py gpurun_out/x.ncu-rep [--top 30]
"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_bytes.sum", "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 30
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "raw", "--csv"))))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print(f"== launch {r[hdr.index('ID')]}: {r[hdr.index('Kernel Name')]}  grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}")
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"{w:86s} {r[i]:>16s} {units[i]}")
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "source", "--csv", "--print-source", "cuda,sass"))))
    hdr, out, ops = None, [], {}
    for r in rows:
        if len(r) > 5 and r[0] == "Line No":
            hdr = r
            continue
        if not hdr or len(r) != len(hdr):
            continue
        try:
            n = int(r[hdr.index("Instructions Executed")])
        except ValueError:
            continue
        if r[0] != "":
            out.append((n, int(r[6]), r[0], r[1][:120]))
        else:
            toks = r[3].split()
            if toks:
                op = (toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]).split(".")[0]
                ops[op] = ops.get(op, 0) + n
    tot = sum(o[0] for o in out) or 1
    print(f"\n== hottest source lines (share of {tot} executed warp instructions; samples)")
    for n, s, l, src in sorted(out, reverse=True)[:top]:
        print(f"{100 * n / tot:5.1f}%  smp {s:5d}  L{l:>4s}  {src}")
    print("\n== opcode mix")
    for op, n in sorted(ops.items(), key=lambda kv: -kv[1])[:16]:
        print(f"{op:10s} {100 * n / tot:5.1f}%")


if __name__ == "__main__":
    main()
