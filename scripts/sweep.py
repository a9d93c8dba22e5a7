"""Batch-size sweep (B in {1, 16, 64, 256, 1024}) and BASELINE configs 2 / 4 / 5 through bench.py, one
process per line (each run rebuilds its corpus).  Writes the contract lines to profiles/r02_sweep.json.

usage: python scripts/sweep.py [out.json]"""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
out = Path(sys.argv[1]) if len(sys.argv) > 1 else ROOT / "profiles" / "r02_sweep.json"
runs = [("config3_B%d" % b, ["--config", "3", "--batch", str(b), "--verify-queries", "0"]) for b in (1, 16, 64, 256, 1024)]
runs += [("config2", ["--config", "2", "--verify-queries", "8"]), ("config4", ["--config", "4", "--verify-queries", "8"]),
         ("config2_B1024", ["--config", "2", "--batch", "1024", "--verify-queries", "0"]),
         ("config4_B256", ["--config", "4", "--batch", "256", "--verify-queries", "0"]),
         ("pairwise", ["--config", "pairwise", "--verify-queries", "0"])]
result = {}
for name, extra in runs:
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "20", "--warmup", "3", *extra],
                       capture_output=True, text=True, cwd=str(ROOT))
    line = next((ln for ln in r.stdout.splitlines() if ln.startswith("{")), None)
    result[name] = json.loads(line) if line else {"error": (r.stderr or r.stdout)[-2000:]}
    d = result[name]
    if "error" in d:
        print(name, "FAILED", d["error"][-300:], flush=True)
    else:
        rf = d.get("roofline") or {}
        print(f"{name}: {d['value']:.4g} {d['unit']}, {d['ms_per_step']:.3f} ms/step, e2e {d['e2e']['value']:.4g}, roofline "
              f"{rf.get('bound')} {rf.get('frac', 0):.3f}", flush=True)
out.write_text(json.dumps(result, indent=1))
print("wrote", out)
