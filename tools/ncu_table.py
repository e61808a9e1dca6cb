"""Key metrics of an `ncu --page raw --csv` export as a markdown table (one row per captured launch)."""
import csv
import sys

WANT = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "DRAM rd"),
    ("dram__bytes_write.sum", "DRAM wr"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__registers_per_thread", "regs"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
    ("smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "stall long sb"),
    ("smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio", "stall short sb"),
    ("smsp__average_warp_latency_issue_stalled_barrier.ratio", "stall barrier"),
    ("smsp__average_warp_latency_issue_stalled_membar.ratio", "stall membar"),
    ("smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio", "stall math"),
    ("smsp__average_warp_latency_issue_stalled_wait.ratio", "stall wait"),
    ("smsp__inst_executed_op_local_ld.sum", "local ld"),
    ("smsp__inst_executed_op_local_st.sum", "local st"),
    ("smsp__inst_executed.sum", "warp insts"),
]
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
cols = [(m, n) for m, n in WANT if m in idx]
print("| kernel | grid | " + " | ".join(n for _, n in cols) + " |")
print("|---|---|" + "---|" * len(cols))
for r in rows[2:]:
    out = []
    for m, _ in cols:
        v, u = r[idx[m]], units[idx[m]]
        try:
            x = float(v.replace(",", ""))
            v = "%.4g" % x
        except ValueError:
            pass
        out.append("%s %s" % (v, u) if u and u not in ("%", "inst", "register/thread", "ratio") else v)
    print("| %s | %s | %s |" % (r[idx["Kernel Name"]].split("(")[0][:28], r[idx["Grid Size"]], " | ".join(out)))
