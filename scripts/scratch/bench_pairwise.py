"""BASELINE config 5: consistency-checker pairwise claim cosine, M claims x 1536-d, one B200.

    python scripts/bench_pairwise.py [M]

Times the tensor-core path (tcgen05 tf32 first pass with a fixed threshold + float64 re-score) with
CUDA events and prints one JSON line; the pair set is checked against the exact float64 sweep on a
sub-block (the full exact sweep at 64k is O(M^2 D) float64 and takes seconds)."""
import json
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from optimized_rag_b200 import engine, synthetic as syn  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
DIM, THR = 1536, 0.85
dev = "cuda:0"
emb = engine.gen_embeddings(M, DIM, 0, syn.SEED_CORPUS, 0, device=dev)
# plant near-duplicates: every 64th claim is a noisy copy of another claim (cosine ~0.85-0.97)
g = torch.Generator(device="cpu").manual_seed(5)
src = torch.randint(0, M, (M // 64,), generator=g)
dst = torch.arange(0, M, 64)[: src.numel()]
w = (0.25 + 0.4 * torch.rand(src.numel(), generator=g)).to(dev)[:, None]
emb[dst.to(dev)] = emb[src.to(dev)] + w * emb[dst.to(dev)]
doc = (torch.arange(M, device=dev) // 16).to(torch.int32)

res = engine.pairwise_cosine_threshold(emb, doc, THR, mode="tc")  # warm-up
torch.cuda.synchronize()
times = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res = engine.pairwise_cosine_threshold(emb, doc, THR, mode="tc")
    e1.record()
    torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1))
ms = float(np.median(times))
n_pairs = int(res[0].numel())
# check a 4096-row sub-block against the exact sweep
sub = 4096
a = engine.pairwise_cosine_threshold(emb[:sub].contiguous(), doc[:sub].contiguous(), THR, mode="exact")
b = engine.pairwise_cosine_threshold(emb[:sub].contiguous(), doc[:sub].contiguous(), THR, mode="tc")
same = all(torch.equal(x, y) for x, y in zip(a, b))
flops_tri = float(M) * (M - 1) * DIM                      # upper triangle actually required
blocks = (M + 255) // 256
flops_done = sum(2.0 * 256 * min(M, (j + 1) * 256) * DIM for j in range(blocks))  # rows [0, j0+256) per query block
peaks = json.loads((Path(__file__).resolve().parents[2] / "MEASURED_PEAKS.json").read_text()) \
    if (Path(__file__).resolve().parents[2] / "MEASURED_PEAKS.json").exists() else {"bf16_tflops_sustained": 1400.0}
print(json.dumps({"config": f"pairwise cosine {M} x {DIM}, threshold {THR}, doc_idx = i // 16", "ms": ms,
                  "pairs_found": n_pairs, "subblock_equals_exact_sweep": bool(same),
                  "tflops_required_triangle": flops_tri / ms / 1e9,
                  "tflops_executed": flops_done / ms / 1e9,
                  "frac_of_sustained_bf16_peak_executed": flops_done / ms / 1e9 / peaks["bf16_tflops_sustained"],
                  "note": "tf32 first pass: the tf32 tensor ceiling is half the bf16 peak"}))
