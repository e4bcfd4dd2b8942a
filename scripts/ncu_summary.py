#!/usr/bin/env python
"""Summarise one kernel of an ncu --set full report into a text file for profiles/.
usage: ncu_summary.py <report.ncu-rep> <out.txt> <rows> <description>"""
import csv, subprocess, sys
rep, out, rows, desc = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rs = list(csv.reader(raw.splitlines()))
hdr, units, d = rs[0], rs[1], dict(zip(rs[0], rs[2]))
keys = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__block_size", "launch__grid_size",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__ops_path_tensor_op_utchmma_src_tf32_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "smsp__inst_executed.sum"]
mult = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1}
lines = [f"ncu --set full --clock-control none, kernel {d['Kernel Name']}", desc,
         "numbers under a profiler are evidence for pipe shares and traffic, not bench values", ""]
for k in keys:
    if k in d:
        lines.append(f"{k:100s} {d[k]:>18s} {units[hdr.index(k)]}")
rd = float(d["dram__bytes_read.sum"]) * mult[units[hdr.index("dram__bytes_read.sum")]]
wr = float(d["dram__bytes_write.sum"]) * mult[units[hdr.index("dram__bytes_write.sum")]]
lines += ["", f"DRAM traffic per launch: {rd + wr:.0f} B = {(rd + wr) / rows:.1f} B/row at {rows} rows"]
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[4:]))
