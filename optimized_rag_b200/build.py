"""Build recipe for the C-ABI library (csrc/*.cu -> csrc/liborag.so), sm_100a only.

nvcc cross-compiles without a GPU, so this runs in the build container; the resulting .so is
git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

CSRC = Path(__file__).resolve().parent / "csrc"
LIB = CSRC / "liborag.so"
SOURCES = ["api.cu", "gen.cu", "cosine_exact.cu", "cosine_tc.cu", "bm25.cu", "bm25_ms.cu", "bm25_build.cu", "rrf.cu", "pairwise.cu", "exchange.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
# -fmad=false: the float64 paths must perform one IEEE rounding per Python-level operation
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-fmad=false",
         "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _stale(out: Path, deps) -> bool:
    if not out.exists():
        return True
    t = out.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    headers = list(CSRC.glob("*.cuh")) + [CSRC.parent.parent / "include" / "orag.h"]
    objs = []
    jobs = []
    for src in SOURCES:
        s = CSRC / src
        o = CSRC / (s.stem + ".o")
        objs.append(o)
        if force or _stale(o, [s] + headers):
            jobs.append((s, o))

    def compile_one(job):
        s, o = job
        r = subprocess.run([NVCC, *FLAGS, "-c", str(s), "-o", str(o)], capture_output=True, text=True)
        return s, r

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for s, r in ex.map(compile_one, jobs):
                if verbose or r.returncode != 0:
                    sys.stderr.write(f"--- nvcc {s.name}\n{r.stdout}{r.stderr}\n")
                if r.returncode != 0:
                    raise RuntimeError(f"nvcc failed on {s.name}")
    if force or jobs or _stale(LIB, objs):
        r = subprocess.run([NVCC, "-shared", "-o", str(LIB), *map(str, objs), "-gencode",
                            "arch=compute_100a,code=sm_100a"], capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
