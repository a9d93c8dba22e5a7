"""GPU debug helper: timeline of one hybrid step (two streams) from CUDA events.
usage: python scripts/debug_step.py [rows] [queries]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from optimized_rag_b200 import engine, synthetic as syn  # noqa: E402
from optimized_rag_b200.bm25_index import Bm25Index  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dev = torch.device("cuda:0")
V, DIM, K = 50000, 1536, 10
thr = syn.zipf_thresholds(V)
corpus = engine.gen_embeddings(N, DIM, 0, syn.SEED_CORPUS, 0, device=dev)
cos = engine.CosineIndex(corpus, mode="f16")
off, tok = engine.gen_token_corpus(N, 0, syn.SEED_TOKENS, thr, V, 100, 300, device=dev)
bm = Bm25Index(off, tok, V, tile_docs=2048)
del tok
shard = engine.HybridShard(cos, bm)
q_emb = torch.from_numpy(syn.query_embeddings(B, N, DIM)).to(dev)
qt, ql = syn.keyword_queries(B, V, thresholds=thr)
qt, ql = torch.from_numpy(qt).to(dev), torch.from_numpy(ql).to(dev)
for _ in range(3):
    shard.search(q_emb, qt, ql, K)
torch.cuda.synchronize()


def ev():
    return torch.cuda.Event(enable_timing=True)


import ctypes
from optimized_rag_b200 import _ffi
L = _ffi.lib()
L.orag_profile_enable(1)
fa, fb = ctypes.c_float(), ctypes.c_float()
for it in range(3):
    cur = torch.cuda.current_stream(dev)
    side = shard._side
    e = {k: ev() for k in ("t0", "b0", "b1", "c0", "c1", "end")}
    e["t0"].record(cur)
    side.wait_stream(cur)
    e["c0"].record(cur)
    ci, cs = cos.topk(q_emb, K, check_overflow=False)
    e["c1"].record(cur)
    with torch.cuda.stream(side):
        L.orag_stream_wait_prescan(side.cuda_stream)
        e["b0"].record(side)
        bi, bs, bmax = bm.topk(qt, ql, K, normalize=True, check_overflow=False, background=True)
        e["b1"].record(side)
    cur.wait_stream(side)
    lists = torch.stack([ci, bi], dim=1).contiguous()
    fi, fs, src = engine.rrf_fuse(lists, 60, K, want_src=True)
    e["end"].record(cur)
    torch.cuda.synchronize()
    L.orag_profile_read(ctypes.byref(fa), ctypes.byref(fb))
    t = {k: e["t0"].elapsed_time(v) for k, v in e.items()}
    print(f"iter {it}: bm25 [{t['b0']:.3f} .. {t['b1']:.3f}] (first-pass kernel {fb.value:.3f})  cosine [{t['c0']:.3f} .. "
          f"{t['c1']:.3f}] (scan kernel {fa.value:.3f})  end {t['end']:.3f} ms")
# each pipeline alone
for name, fn in (("bm25 alone", lambda: bm.topk(qt, ql, K, normalize=True, check_overflow=False)),
                 ("bm25 alone, background shape", lambda: bm.topk(qt, ql, K, normalize=True, check_overflow=False,
                                                                  background=True)),
                 ("cosine alone", lambda: cos.topk(q_emb, K, check_overflow=False))):
    a, b = ev(), ev()
    torch.cuda.synchronize()
    a.record()
    fn()
    b.record()
    torch.cuda.synchronize()
    print(f"{name}: {a.elapsed_time(b):.3f} ms")
