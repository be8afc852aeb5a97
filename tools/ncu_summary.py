"""Summarise an .ncu-rep (ncu --set full) into the text form kept under profiles/: one block per kernel launch with the
duration, DRAM bytes and throughput, L2 hit rate, pipe utilisation and the leading stall reasons.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [kernel-name regex] > profiles/rNN_ncu_x.txt
       python tools/ncu_summary.py --traffic gpurun_out/x.ncu-rep [kernel-name regex]
           merges dram__bytes_read.sum + dram__bytes_write.sum of the LARGEST launch of every kernel of the capture into
           profiles/ncu_traffic.json, keyed "name<template args>" (what bench.py's roofline.traffic reads)"""
import csv
import io
import re
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes.sum.per_second", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__ops_path_tensor_src_fp64.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
]
argv = [a for a in sys.argv[1:] if a != "--traffic"]
traffic_mode = "--traffic" in sys.argv[1:]
rep = argv[0]
pat = re.compile(argv[1]) if len(argv) > 1 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
if traffic_mode:
    import json
    from pathlib import Path

    def to_bytes(v, unit):
        mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
        return float(v.replace(",", "")) * mult[unit]
    path = Path(__file__).resolve().parent.parent / "profiles" / "ncu_traffic.json"
    table = json.loads(path.read_text()) if path.exists() else {}
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        if pat and not pat.search(name):
            continue
        m = re.match(r"(?:void )?(?:\w+::)*(\w+)(?:<([^>]*)>)?", name)
        targs = ",".join(re.sub(r"\(\w+\)", "", a).strip() for a in (m.group(2) or "").split(",")) if m.group(2) else ""
        key = m.group(1) + (f"<{targs}>" if targs else "")
        rd = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
        wr = to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
        us = r[col["gpu__time_duration.sum"]] + " " + units[col["gpu__time_duration.sum"]]
        if key not in table or rd + wr > table[key]["dram_bytes_per_launch"] or table[key].get("capture") != rep:
            if key in table and table[key].get("capture") == rep and rd + wr <= table[key]["dram_bytes_per_launch"]:
                continue
            table[key] = {"dram_bytes_per_launch": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr, "grid": r[col["Grid Size"]],
                          "duration_under_ncu": us, "capture": rep}
    path.write_text(json.dumps(table, indent=1) + "\n")
    print(json.dumps(table, indent=1))
    sys.exit(0)
print(f"# ncu --set full --clock-control none: {rep}")
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    if pat and not pat.search(name):
        continue
    print(f"== {name}  grid={r[col['Grid Size']]} block={r[col['Block Size']]}")
    for m in METRICS:
        if m in col:
            print(f"   {m} = {r[col[m]]} {units[col[m]]}")
