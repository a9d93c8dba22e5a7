"""GPU debug helper: host-side timeline of one end-to-end step (pinned H2D -> search -> D2H -> sync)."""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from optimized_rag_b200 import engine, synthetic as syn  # noqa: E402
from optimized_rag_b200.bm25_index import Bm25Index  # noqa: E402
from optimized_rag_b200.dist import ShardedHybrid  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
B, V, DIM, K = 256, 50000, 1536, 10
dev = torch.device("cuda:0")
thr = syn.zipf_thresholds(V)
corpus = engine.gen_embeddings(N, DIM, 0, syn.SEED_CORPUS, 0, device=dev)
cos = engine.CosineIndex(corpus, mode="f16")
off, tok = engine.gen_token_corpus(N, 0, syn.SEED_TOKENS, thr, V, 100, 300, device=dev)
bm = Bm25Index(off, tok, V, tile_docs=2048)
del tok
sh = ShardedHybrid(engine.HybridShard(cos, bm))
qe_h = torch.from_numpy(syn.query_embeddings(B, N, DIM)).pin_memory()
qt_np, ql_np = syn.keyword_queries(B, V, thresholds=thr)
qt_h, ql_h = torch.from_numpy(qt_np).pin_memory(), torch.from_numpy(ql_np).pin_memory()
qe, qt, ql = qe_h.to(dev), qt_h.to(dev), ql_h.to(dev)
ids_h = torch.empty((B, K), dtype=torch.int64).pin_memory()
sc_h = torch.empty((B, K), dtype=torch.float64).pin_memory()
st_h = torch.empty(B, dtype=torch.int32).pin_memory()
for _ in range(3):
    sh.search(qe, qt, ql, K)
torch.cuda.synchronize()
for it in range(6):
    t0 = time.perf_counter()
    qe.copy_(qe_h, non_blocking=True); qt.copy_(qt_h, non_blocking=True); ql.copy_(ql_h, non_blocking=True)
    t1 = time.perf_counter()
    r = sh.search(qe, qt, ql, K, check_overflow=False)
    t2 = time.perf_counter()
    ids_h.copy_(r["ids"], non_blocking=True); sc_h.copy_(r["rrf_scores"], non_blocking=True)
    st_h.copy_(r["status"], non_blocking=True)
    t3 = time.perf_counter()
    torch.cuda.current_stream().synchronize()
    t4 = time.perf_counter()
    print(f"iter {it}: H2D enqueue {1e3 * (t1 - t0):.3f}  search enqueue {1e3 * (t2 - t1):.3f}  D2H enqueue "
          f"{1e3 * (t3 - t2):.3f}  wait {1e3 * (t4 - t3):.3f}  total {1e3 * (t4 - t0):.3f} ms")
