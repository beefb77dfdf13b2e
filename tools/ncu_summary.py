#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into a small JSON: per kernel launch, the counters the DESIGN/roofline sections quote.
    python tools/ncu_summary.py report.ncu-rep out.json"""
import csv
import json
import subprocess
import sys

KEEP = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed": "l1tex_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "smsp__inst_executed.sum": "warp_instructions",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_wavefronts",
    "smsp__inst_executed_op_shared_atom.sum": "smem_atomic_instructions",
    "lts__t_sectors_op_atom.sum": "l2_atom_sectors",
    "lts__t_sectors_op_red.sum": "l2_red_sectors",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__occupancy_limit_shared_mem": "occ_limit_smem",
    "launch__occupancy_limit_registers": "occ_limit_regs",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "stall_barrier",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_scoreboard",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio": "stall_mio_throttle",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio": "stall_lg_throttle",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio": "stall_math_pipe",
}


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for i, h in enumerate(hdr):      # some columns carry a section prefix ("FBSP.TriageCompute.dram__throughput…")
        for k in KEEP:
            if h.endswith("." + k) and k not in idx:
                idx[k] = i
    res = []
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        d = {"kernel": r[idx["Kernel Name"]].split("(")[0]}
        for k, name in KEEP.items():
            if k in idx:
                v = r[idx[k]].replace(",", "")
                try:
                    d[name] = float(v)
                except ValueError:
                    d[name] = v
                if name in ("duration", "dram_read", "dram_write"):
                    d[name + "_unit"] = units[idx[k]]
        res.append(d)
    json.dump({"source": rep, "how": "ncu --set full --clock-control none (cold cache, serialised launches)", "launches": res}, open(out, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
