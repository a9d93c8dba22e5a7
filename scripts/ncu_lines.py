"""Rank CUDA source lines of an ncu report by executed instructions / stall samples.

usage: python scripts/ncu_lines.py report.ncu-rep [top_n]
(reads `ncu -i report --page source --print-source cuda,sass --csv`; needs -lineinfo at compile time)
"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))


def ok(r):
    try:
        int(r[7]); int(r[4])
        return r[0].isdigit()
    except (ValueError, IndexError):
        return False


data = [r for r in rows if len(r) > 10 and ok(r)]
tot_i = sum(int(r[7]) for r in data)
tot_s = sum(int(r[4]) for r in data)
print(f"total warp instructions {tot_i}, stall samples {tot_s}")
print("line samples  smp%   warp-inst  inst% thr/inst source")
data.sort(key=lambda r: -int(r[7]))
for r in data[:top]:
    print(r[0].rjust(4), r[4].rjust(7), ("%.1f%%" % (100 * int(r[4]) / max(tot_s, 1))).rjust(6), r[7].rjust(11),
          ("%.1f%%" % (100 * int(r[7]) / max(tot_i, 1))).rjust(6), r[10].rjust(3), " ", r[1][:110])
