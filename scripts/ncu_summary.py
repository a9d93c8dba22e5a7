"""Selected metrics of one kernel out of an `ncu --set full` report, in the plain-text form kept under profiles/.

usage: python scripts/ncu_summary.py report.ncu-rep "header comment (the command that produced the report)" > out.txt"""
import csv
import io
import subprocess
import sys

KEEP = ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct", "gpu__time_duration.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "launch__block_size", "launch__grid_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__cycles_elapsed.max", "sm__pipe_tensor_cycles_active.avg.pct", "sm__throughput.avg.pct",
        "sm__warps_active.avg.pct", "sm__inst_executed.sum.per_cycle", "smsp__average_warp_latency_per_inst_issued",
        "smsp__average_warps_issue_stalled", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct",
        "smsp__thread_inst_executed_per_inst_executed.ratio")
rep, header = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
names, units, vals = rows[0], rows[1], rows[2]
print("# " + header)
for n, u, v in sorted(zip(names, units, vals)):
    if any(n.startswith(k) for k in KEEP) and "not_issued" not in n:
        print(f"{n} [{u}] = {v}")
