"""Summarise one hybrid step out of an ncu launch list (gpu__time_duration.sum CSV).

usage: python scripts/launch_list.py launches.csv [step_index]
A step ends with rrf_pair_kernel (one GPU) or hybrid_merge_kernel (row-sharded).  Capture plain steps (bench.py runs K of
them after the submitted loop): inside the submitted loop two batches interleave."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hi]
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
L = [(r[ki], float(r[vi].replace(",", ""))) for r in rows[hi + 1:] if len(r) > vi]
ends = [i for i, (n, _) in enumerate(L) if "rrf_kernel" in n or "rrf_pair_kernel" in n or "hybrid_merge" in n]
which = int(sys.argv[2]) if len(sys.argv) > 2 else len(ends) - 1
s, e = ends[which - 1] + 1, ends[which] + 1
tot = sum(v for _, v in L[s:e]) / 1e6
print("kernel,duration_ms,share")
for n, v in L[s:e]:
    print(f"{n.split('(')[0]},{v / 1e6:.4f},{v / 1e6 / tot:.3f}")
print(f"TOTAL,{tot:.4f},1.000")
