"""BASELINE.json configs 2 and 4 on one GPU (the bench.py line is config 3 sized for one GPU):
  config 2: 1M chunks x 1536-d fp32 exact cosine top-10, query batch 256
  config 4: BM25-only over a 10M-chunk Zipf corpus (50k vocab), query batch 1024
usage: python scripts/bench_configs.py [2|4|both]   -> one JSON line per config (CUDA events, 20 calls after 3 warm-ups)"""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from optimized_rag_b200 import engine, synthetic as syn  # noqa: E402
from optimized_rag_b200.bm25_index import Bm25Index  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "both"
dev = torch.device("cuda:0")
peaks = json.loads((Path(__file__).resolve().parents[2] / "MEASURED_PEAKS.json").read_text()) \
    if (Path(__file__).resolve().parents[2] / "MEASURED_PEAKS.json").exists() else {"hbm_gbs": 6544.3}


def timed(fn, steps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


if which in ("2", "both"):
    N, B, D = 1_000_000, 256, 1536
    corpus = engine.gen_embeddings(N, D, 0, syn.SEED_CORPUS, 0, device=dev)
    q = torch.from_numpy(syn.query_embeddings(B, N, D)).to(dev)
    out = {}
    for mode in ("f16", "tf32"):
        idx = engine.CosineIndex(corpus, mode=mode)
        ms = timed(lambda: idx.topk(q, 10, check_overflow=False))
        out[mode] = {"ms_per_batch": ms, "queries_per_s": B / ms * 1e3,
                     "fp32_corpus_gbs": N * D * 4 / ms / 1e6, "frac_of_hbm_peak_fp32_bytes": N * D * 4 / ms / 1e6 / peaks["hbm_gbs"]}
        del idx
    ex = engine.CosineIndex(corpus, mode="exact")
    ids_e, sc_e = ex.topk(q[:8].contiguous(), 10)
    ids_f, sc_f = engine.CosineIndex(corpus, mode="f16").topk(q[:8].contiguous(), 10)
    print(json.dumps({"config": "2: 1M x 1536 exact cosine top-10, batch 256, 1 GPU", "modes": out,
                      "first_pass_equals_exhaustive_float64_scan": bool(torch.equal(ids_e, ids_f) and torch.equal(sc_e, sc_f))}))
    del corpus

if which in ("4", "both"):
    N, B, V = 10_000_000, 1024, 50000
    thr = syn.zipf_thresholds(V)
    off, tok = engine.gen_token_corpus(N, 0, syn.SEED_TOKENS, thr, V, 100, 300, device=dev)
    ix = Bm25Index(off, tok, V, tile_docs=2048)
    del tok
    qt, ql = syn.keyword_queries(B, V, thresholds=thr)
    qt, ql = torch.from_numpy(qt).to(dev), torch.from_numpy(ql).to(dev)
    st = []
    ms = timed(lambda: ix.topk(qt, ql, 10, check_overflow=False, status_out=st))
    by = ix.posting_bytes(qt, ql)
    a = ix.topk(qt[:16].contiguous(), ql[:16].contiguous(), 10, force="exact_tiles")
    b = ix.topk(qt[:16].contiguous(), ql[:16].contiguous(), 10)
    print(json.dumps({"config": "4: BM25-only, 10M-chunk Zipf corpus, V=50k, batch 1024, 1 GPU", "ms_per_batch": ms,
                      "queries_per_s": B / ms * 1e3, "algorithmic_gbs": by / ms / 1e6,
                      "frac_of_hbm_peak": by / ms / 1e6 / peaks["hbm_gbs"], "overflowed_queries": int((st[-1] != 0).sum()),
                      "first_pass_equals_float64_scatter": bool(all(torch.equal(x, y) for x, y in zip(a, b)))}))
