#!/usr/bin/env python
"""Summarise an `ncu --set full` report into the handful of numbers DESIGN.md / bench.py quote.

    python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep > profiles/<name>.md
Needs the `ncu` CLI (no GPU required to read a report).
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe % (active)"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe % (active)"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps / cycle / SMSP"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
]
STALLS = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio"
STALL_NAMES = ["long_scoreboard", "short_scoreboard", "wait", "dispatch_stall", "mio_throttle", "barrier", "math_pipe_throttle", "not_selected",
               "no_instruction", "lg_throttle", "branch_resolving"]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    print("# ncu --set full summary: `%s`\n" % path.split("/")[-1])
    print("Per-launch values under the profiler (cold cache, serialised): compare shares, not absolutes.\n")
    for r in data:
        print("## %s\n" % r[col["Kernel Name"]])
        print("| metric | value |\n|---|---|")
        for k, label in KEYS:
            if k in col and r[col[k]] != "":
                print("| %s (`%s`) | %s %s |" % (label, k, r[col[k]], units[col[k]]))
        st = [(n, float(r[col[STALLS % n]])) for n in STALL_NAMES if (STALLS % n) in col and r[col[STALLS % n]] != ""]
        print("\nWarp stall cycles per issued instruction: " + ", ".join("%s %.3f" % s for s in sorted(st, key=lambda x: -x[1])) + "\n")


if __name__ == "__main__":
    main(sys.argv[1])
