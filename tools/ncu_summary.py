#!/usr/bin/env python
"""Key metrics of every kernel launch in an .ncu-rep (ncu --set full), as the text summary kept under profiles/.
usage: python tools/ncu_summary.py x.ncu-rep > profiles/x_ncu.txt"""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.sum",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
]
STALLS = "smsp__average_warps_issue_stalled_"


def main():
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# {sys.argv[1]}: ncu --set full --clock-control none; one block per captured launch")
    for r in rows[2:]:
        print(f"\n== {r[col['Kernel Name']]}  grid {r[col['Grid Size']]} block {r[col['Block Size']]}")
        for w in WANT:
            if w in col:
                print(f"  {w:78s} {r[col[w]]} {units[col[w]]}")
        st = sorted(((float(r[i] or 0), h) for h, i in col.items() if h.startswith(STALLS) and h.endswith("per_issue_active.ratio")),
                    reverse=True)
        print("  stall reasons (warps per issue-active cycle): " +
              ", ".join(f"{h[len(STALLS):-len('_per_issue_active.ratio')]} {v:.2f}" for v, h in st[:7]))


if __name__ == "__main__":
    main()
