#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU): one line per captured kernel with the metrics the roofline needs."""
import csv, subprocess, sys, io, re
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[0]
def col(name):
    return hdr.index(name) if name in hdr else None
want = {
 "kernel": "Kernel Name", "grid": "launch__grid_size", "block": "launch__block_size", "regs": "launch__registers_per_thread",
 "ms": "gpu__time_duration.sum", "fp64_pct": "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
 "fp64_pipe_active": "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
 "issue_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active", "warps_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
 "dram_rd": "dram__bytes_read.sum", "dram_wr": "dram__bytes_write.sum", "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
 "dfma_thread_x2": "derived__smsp__sass_thread_inst_executed_op_dfma_pred_on_x2",
 "inst": "smsp__inst_executed.sum", "local_ld": "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum", "local_st": "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum",
 "red": "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum",
 "stall_long_sb": "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
 "stall_no_inst": "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
 "stall_wait": "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
 "stall_math": "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
 "stall_lg": "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
 "stall_short_sb": "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
 "stall_mio": "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
}
units = rows[1]
for r in rows[2:]:
    d = {}
    for k, name in want.items():
        c = col(name)
        if c is None:
            # fuzzy
            cand = [i for i, h in enumerate(hdr) if name.split(".")[0] in h and (len(name.split(".")) < 2 or name.split(".", 1)[1] in h)]
            c = cand[0] if cand else None
        d[k] = (r[c] + (" " + units[c] if units[c] and k in ("ms", "dram_rd", "dram_wr") else "")) if c is not None else "n/a"
    d["kernel"] = re.sub(r"\(qcf::PairGroup.*", "", d["kernel"])
    print(" | ".join(f"{k}={v}" for k, v in d.items()))
