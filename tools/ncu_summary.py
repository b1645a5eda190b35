#!/usr/bin/env python
"""Summarise one kernel of an .ncu-rep (read here with `ncu -i ... --page raw --csv`) into a small JSON/markdown pair
under profiles/.  Usage: tools/ncu_summary.py gpurun_out/prof_k_step.ncu-rep profiles/r01_k_step [launch_index]"""
import csv
import json
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "duration",
    "launch__grid_size": "grid", "launch__block_size": "block", "launch__registers_per_thread": "regs_per_thread",
    "launch__stack_size": "stack_bytes_per_thread", "launch__shared_mem_per_block_static": "smem_static_per_block",
    "launch__shared_mem_per_block_dynamic": "smem_dynamic_per_block", "launch__waves_per_multiprocessor": "waves_per_sm",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "smsp__inst_executed.sum": "warp_instructions",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "avg_active_threads_per_inst",
    "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "lts__t_sectors_srcunit_tex_op_read.sum": "l2_read_sectors", "lts__t_sectors_srcunit_tex_op_write.sum": "l2_write_sectors",
    "sass__inst_executed_local_loads": "local_load_insts", "sass__inst_executed_local_stores": "local_store_insts",
    "l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct": "local_ld_l1_hit_pct",
    "l1tex__t_sector_pipe_lsu_mem_local_op_st_hit_rate.pct": "local_st_l1_hit_pct",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio": "stall_no_instruction",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio": "stall_branch_resolving",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_scoreboard",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio": "stall_lg_throttle",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "stall_barrier",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio": "stall_math_pipe",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio": "stall_mio_throttle",
    "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active": "hmma_pct",
}
UNIT_SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    idx = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    r = data[idx]
    res = {"report": rep, "launch_index": idx, "launches_in_report": len(data)}
    for i, h in enumerate(hdr):
        if h == "Kernel Name":
            res["kernel"] = r[i].split("(")[0]
        if h in KEYS:
            try:
                v = float(r[i].replace(",", ""))
            except ValueError:
                continue
            u = units[i]
            if u in UNIT_SCALE and KEYS[h] in ("duration", "dram_read", "dram_write"):
                v *= UNIT_SCALE[u]
                u = "s" if KEYS[h] == "duration" else "byte"
            res[KEYS[h]] = v
            res.setdefault("_units", {})[KEYS[h]] = u
    if "dram_read" in res and "dram_write" in res:
        res["dram_bytes_per_launch"] = res["dram_read"] + res["dram_write"]
    json.dump(res, open(out + ".json", "w"), indent=1, sort_keys=True)
    with open(out + ".md", "w") as f:
        f.write(f"# ncu --set full summary: {res.get('kernel')} (launch {idx} of {len(data)} in {rep})\n\n| metric | value | unit |\n|---|---|---|\n")
        for k in sorted(res):
            if k.startswith("_") or k in ("report", "kernel"):
                continue
            f.write(f"| {k} | {res[k]} | {res.get('_units', {}).get(k, '')} |\n")
    print(json.dumps(res, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
