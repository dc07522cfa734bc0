#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into the few numbers the design decisions rest on.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [more.ncu-rep ...] > profiles/x.txt"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"), ("launch__shared_mem_per_block_static", "static smem/block"),
    ("launch__occupancy_limit_registers", "occ limit regs (blocks)"), ("launch__occupancy_limit_shared_mem", "occ limit smem (blocks)"),
    ("launch__occupancy_limit_warps", "occ limit warps (blocks)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe % (inst)"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe cycles active %"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe cycles active %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("lts__t_sectors_op_atom.sum", "L2 atomic sectors"), ("lts__t_sectors_op_red.sum", "L2 red sectors"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / warp inst"),
]


def raw_rows(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    i0 = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    return rows[i0], rows[i0 + 1], rows[i0 + 2:]


def main():
    for path in sys.argv[1:]:
        hdr, units, rows = raw_rows(path)
        col = {h: i for i, h in enumerate(hdr)}
        for r in rows:
            print(f"== {path.split('/')[-1]} :: {r[col['Kernel Name']][:110]}")
            for k, label in KEYS:
                if k in col and r[col[k]] != "":
                    print(f"   {label:32s} {r[col[k]]:>16s} {units[col[k]]}")
            stalls = []
            for h, i in col.items():
                if h.startswith("smsp__average_warp") and h.endswith("_per_issue_active.ratio") and "latency_issue_stalled" in h:
                    try:
                        stalls.append((float(r[i]), h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")))
                    except ValueError:
                        pass
                elif h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio"):
                    try:
                        stalls.append((float(r[i]), h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")))
                    except ValueError:
                        pass
            stalls.sort(reverse=True)
            if stalls:
                print("   warp stall reasons (warps per issue-active cycle): " +
                      ", ".join(f"{n}={v:.2f}" for v, n in stalls[:7]))
            dr, dw = col.get("dram__bytes_read.sum"), col.get("dram__bytes_write.sum")
            if dr is not None and r[dr]:
                mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                tot = float(r[dr]) * mult.get(units[dr], 1.0) + float(r[dw]) * mult.get(units[dw], 1.0)
                print(f"   traffic (DRAM read+write)        {tot / 1e6:16.3f} MB per launch")
            print()


if __name__ == "__main__":
    main()
