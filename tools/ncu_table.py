"""Summarise an `ncu --page raw --csv` export: one line per launch with the metrics we steer by."""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
h = rows[0]
pat = sys.argv[2] if len(sys.argv) > 2 else "."
cols = {"t_us": "gpu__time_duration.sum", "dramR": "dram__bytes_read.sum", "dramW": "dram__bytes_write.sum",
        "dram%": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts%": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1%": "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1hit": "l1tex__t_sector_hit_rate.pct",
        "fma%": "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "alu%": "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "xu%": "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "lsu%": "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "warps%": "sm__warps_active.avg.pct_of_peak_sustained_active", "regs": "launch__registers_per_thread",
        "inst_M": "smsp__inst_executed.sum", "stall_lsb": "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "stall_math": "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "stall_wait": "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "stall_mio": "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "stall_lg": "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "ipc": "sm__inst_executed.avg.per_cycle_active", "tensor%": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"}
idx = {k: h.index(v) for k, v in cols.items() if v in h}
ki, gi = h.index("Kernel Name"), h.index("Grid Size")
units = rows[1]                      # ncu scales each column: take the unit of the byte columns from the unit row
label = lambda k: (k + "_" + units[idx[k]].replace("byte", "B")) if k.startswith("dram") and "byte" in units[idx[k]] else k
print("kernel".ljust(34), "grid".rjust(8), " ".join(label(k).rjust(9) for k in idx))
for r in rows[2:]:
    name = re.sub(r"\(.*", "", r[ki]).replace("void dfd::", "").replace("void ", "")
    if not re.search(pat, name):
        continue
    vals = []
    for k, i in idx.items():
        try:
            v = float(r[i].replace(",", ""))
            if k == "inst_M":
                v /= 1e6
            vals.append(f"{v:9.2f}")
        except Exception:
            vals.append("      n/a")
    print(name[:34].ljust(34), r[gi].split(",")[0].strip("(").rjust(8), " ".join(vals))
