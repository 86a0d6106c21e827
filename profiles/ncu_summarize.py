"""Summarises an ncu report (ncu -i X.ncu-rep --page raw --csv) into the handful of numbers the roofline accounting uses.
    ncu -i gpurun_out/r2_sweep.ncu-rep --page raw --csv > /tmp/raw.csv; python profiles/ncu_summarize.py /tmp/raw.csv [labels...]"""
import csv
import json
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput % of peak"),
    ("dram__bytes.sum.per_second", "dram bytes/s"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput % of peak"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe active %"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short scoreboard"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not selected"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("sm__cycles_active.min", "SM active cycles min"),
    ("sm__cycles_active.avg", "SM active cycles avg"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long scoreboard"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
]


def num(s):
    try:
        return float(s.replace(",", ""))
    except Exception:
        return None


rows = list(csv.reader(open(sys.argv[1])))
labels = sys.argv[2:]
hdr = rows[0]
units = rows[1]
col = {h: i for i, h in enumerate(hdr)}
out = []
for li, r in enumerate(rows[2:]):
    name = r[col["Kernel Name"]]
    rec = {"launch": li, "label": labels[li] if li < len(labels) else "", "kernel": name[:90]}
    for k, nice in KEYS:
        if k in col:
            v = num(r[col[k]])
            rec[nice] = v
            rec[nice + " unit"] = units[col[k]]
    out.append(rec)
for rec in out:
    print(f"--- launch {rec['launch']} {rec['label']}  {rec['kernel']}")
    for k, nice in KEYS:
        if nice in rec and rec[nice] is not None:
            print(f"    {nice:32s} {rec[nice]:>18,.3f} {rec[nice + ' unit']}")
    if rec.get("dram read") is not None and rec.get("dram write") is not None:
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        tot = rec["dram read"] * scale[rec["dram read unit"]] + rec["dram write"] * scale[rec["dram write unit"]]
        rec["dram bytes"] = tot
        print(f"    {'dram read + write':32s} {tot:>18,.0f} byte")
json.dump(out, open(sys.argv[1] + ".json", "w"), indent=1)
