#!/usr/bin/env python
"""Summarise an .ncu-rep (one kernel) into the text kept under profiles/.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.txt"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__warps_active.avg.per_cycle_active", "sm__cycles_active.avg", "gpc__cycles_elapsed.max",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "lts__t_bytes.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    print(f"# ncu summary of {path}  ({len(data)} launch(es) captured; ncu --set full --clock-control none)")
    for n, r in enumerate(data):
        print(f"\n## launch {n}: {r[ix['Kernel Name']] if 'Kernel Name' in ix else ''}")
        for k in KEYS:
            if k in ix and r[ix[k]] not in ("", "n/a"):
                print(f"{k:70s} {r[ix[k]]:>18s} {units[ix[k]]}")
        st = {}
        for h, i in ix.items():
            if "smsp__pcsamp_warps_issue_stalled" in h and not h.endswith("not_issued"):
                try:
                    st[h.replace("smsp__pcsamp_warps_issue_stalled_", "")] = float(r[i])
                except ValueError:
                    pass
        tot = sum(st.values()) or 1.0
        print("warp stall samples (all):")
        for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]:
            print(f"    {k:26s} {v:9.0f}  {100 * v / tot:5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
