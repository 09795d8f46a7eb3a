#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): one block of key metrics per captured launch.
usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.txt"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__occupancy_limit_warps",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# {rep}: {len(rows) - 2} captured launch(es); ncu --set full --clock-control none")
    for r in rows[2:]:
        print(f"\n== {r[idx['Kernel Name']]}  grid {r[idx['Grid Size']]} block {r[idx['Block Size']]}")
        for k in KEYS:
            if k in idx and r[idx[k]] != "":
                print(f"  {k:88s} {r[idx[k]]:>18s} {units[idx[k]]}")
        stalls = [(float(r[i]), h) for h, i in idx.items()
                  if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and r[i] not in ("", "n/a")]
        for v, h in sorted(stalls, reverse=True)[:7]:
            print(f"  stall {h[len(STALL):-len('_per_issue_active.ratio')]:82s} {v:18.3f} warps/issue")


if __name__ == "__main__":
    main()
