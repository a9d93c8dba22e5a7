"""Summarise one hybrid step out of an ncu launch list (gpu__time_duration.sum CSV).

usage: python scripts/launch_list.py launches.csv [step_index]
A step starts at a query_sq_kernel launch (first kernel of orag_cosine_topk)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hi]
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
L = [(r[ki], float(r[vi].replace(",", ""))) for r in rows[hi + 1:] if len(r) > vi]
idx = [i for i, (n, _) in enumerate(L) if "query_sq" in n]
which = int(sys.argv[2]) if len(sys.argv) > 2 else len(idx) - 2
s, e = idx[which], idx[which + 1] if which + 1 < len(idx) else len(L)
tot = sum(v for _, v in L[s:e]) / 1e6
print("kernel,duration_ms,share")
for n, v in L[s:e]:
    print(f"{n.split('(')[0]},{v / 1e6:.4f},{v / 1e6 / tot:.3f}")
print(f"TOTAL,{tot:.4f},1.000")
