"""Stand-alone duration of every tagged launch of one hybrid search (serial: one stream, nothing else on the GPU).

usage: python scripts/timeline_serial.py [ROWS] [QUERIES]   -- mean / min over 10 searches, per tag (csrc/common.cuh)"""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
from bench import TIMELINE_TAGS, VOCAB, LMIN, LMAX, DIM  # noqa: E402
from optimized_rag_b200 import _ffi, engine, synthetic as syn  # noqa: E402
from optimized_rag_b200.bm25_index import Bm25Index, Bm25Plan  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_250_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dev = torch.device("cuda:0")
L = _ffi.lib()
thr = syn.zipf_thresholds(VOCAB)
cos = engine.CosineIndex(engine.gen_embeddings(rows, DIM, 0, syn.SEED_CORPUS, 0, device=dev), mode="f16")
off, tok = engine.gen_token_corpus(rows, 0, syn.SEED_TOKENS, thr, VOCAB, LMIN, LMAX, device=dev)
shard = engine.HybridShard(cos, Bm25Index.from_plan(Bm25Plan(off, tok, VOCAB)))
del tok
shard.coschedule, shard.serial = False, True
q = torch.from_numpy(syn.query_embeddings(nq, rows, DIM)).to(dev)
qt, ql = syn.keyword_queries(nq, VOCAB, thresholds=thr)
qt, ql = torch.from_numpy(qt).to(dev), torch.from_numpy(ql).to(dev)
for _ in range(3):
    shard.search(q, qt, ql, 10, check_overflow=False)
torch.cuda.synchronize()
L.orag_timeline_enable(1)
n_steps = 10
for _ in range(n_steps):
    shard.search(q, qt, ql, 10, check_overflow=False)
cap = 4096
tags, t0, t1 = (ctypes.c_int * cap)(), (ctypes.c_float * cap)(), (ctypes.c_float * cap)()
n = int(L.orag_timeline_read(tags, t0, t1, cap))
L.orag_timeline_enable(0)
by = {}
for i in range(n):
    by.setdefault(TIMELINE_TAGS[tags[i]], []).append(t1[i] - t0[i])
span = (max(t1[i] for i in range(n)) - min(t0[i] for i in range(n))) / n_steps
print(f"# {rows} rows x {nq} queries, serial searches: {span:.4f} ms per search")
print("# tag mean_ms min_ms launches_per_search")
tot = 0.0
for tag, xs in sorted(by.items(), key=lambda kv: -sum(kv[1])):
    print(f"{tag:16s} {sum(xs) / len(xs):8.4f} {min(xs):8.4f} {len(xs) / n_steps:4.1f}")
    tot += sum(xs) / n_steps
print(f"# sum of the brackets {tot:.4f} ms per search")
