"""Small, fast exercise of every hand-written kernel family for compute-sanitizer: the tcgen05 scan in its
three flavours (plain; ORAG_SCAN_CLUSTER=2 forces CTA pairs on small inputs: 2-SM MMA, or multicast with ORAG_SCAN_2SM=0)
incl. pair mode, the re-score chain, the BM25 builder + first pass + finalize, RRF, and the peer exchange with two virtual
ranks in one process (pushes complete before the waits are launched).  Results are checked against the exhaustive kernels.

usage (one tool per gpurun call):
    compute-sanitizer --tool memcheck  python scripts/sanitize_target.py
    ORAG_SCAN_CLUSTER=2 compute-sanitizer --tool racecheck python scripts/sanitize_target.py scan"""
import ctypes
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from optimized_rag_b200 import _ffi, engine, synthetic as syn  # noqa: E402
from optimized_rag_b200.bm25_index import Bm25Index  # noqa: E402
from optimized_rag_b200.dist import hybrid_merge, pack_local  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "all"
dev = "cuda:0"
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
ok = True

if what in ("all", "scan"):
    n, dim, nq, k = 6000, 128, 20, 10
    corpus = t(syn.embeddings(syn.SEED_CORPUS, 0, n, dim, 5))
    q = t(syn.query_embeddings(nq, n, dim, dup_per_mille=5))
    want = engine.CosineIndex(corpus, mode="exact").topk(q, k)
    for mode in ("f16", "tf32"):
        got = engine.CosineIndex(corpus, mode=mode).topk(q, k)
        same = torch.equal(got[0], want[0]) and torch.equal(got[1], want[1])
        print(f"scan {mode}: {'ok' if same else 'MISMATCH'}", flush=True)
        ok &= same
    m = 2048
    emb = syn.embeddings(syn.SEED_CORPUS, 0, m, dim)
    emb[::64] = emb[9::64] + np.float32(0.3) * emb[::64]      # near-duplicates of a row of ANOTHER document
    doc = t((np.arange(m) // 4).astype(np.int32))
    a = engine.pairwise_cosine_threshold(t(emb), doc, 0.85, mode="exact")
    b = engine.pairwise_cosine_threshold(t(emb), doc, 0.85, mode="tc")
    same = all(torch.equal(x, y) for x, y in zip(a, b)) and a[0].numel() > 0
    print(f"pair mode: {'ok' if same else 'MISMATCH'} ({a[0].numel()} pairs)", flush=True)
    ok &= same

if what in ("all", "bm25"):
    n, vocab, nq, k = 3000, 500, 16, 10
    thr = syn.zipf_thresholds(vocab)
    doc_off, tok = syn.token_corpus(syn.SEED_TOKENS, 0, n, vocab, 10, 60, thr)
    ix = Bm25Index(t(doc_off), t(tok), vocab, tile_docs=256, fp_tile_docs=512)
    qt, ql = syn.keyword_queries(nq, vocab, min_rank=5, thresholds=thr)
    a = ix.topk(t(qt), t(ql), k, force="sparse")
    b = ix.topk(t(qt), t(ql), k, force="dense")
    same = all(torch.equal(x, y) for x, y in zip(a, b))
    print(f"bm25 builder + first pass: {'ok' if same else 'MISMATCH'}", flush=True)
    ok &= same
    shard = engine.HybridShard(engine.CosineIndex(t(syn.embeddings(syn.SEED_CORPUS, 0, n, 64)), mode="exact"), ix)
    res = shard.search(t(syn.query_embeddings(nq, n, 64)), t(qt), t(ql), k)
    ok &= int(res["status"].max().item()) == 0

if what in ("all", "exchange"):
    L = _ffi.lib()
    G, maxq, fk, kk, k, B = 2, 32, 10, 16, 10, 24
    rng = np.random.default_rng(3)
    bufs = []
    for _ in range(G):
        p = ctypes.c_void_p()
        _ffi.check(L.orag_exchange_alloc(int(L.orag_exchange_bytes(G, maxq, fk, kk)), ctypes.byref(p)), "alloc")
        bufs.append(p.value)
    d_peers = torch.tensor(bufs, dtype=torch.int64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    x_ok = True
    for seq in (1, 2, 3, 4, 5):
        lists = []
        for g in range(G):
            ci = np.stack([rng.permutation(500)[:fk] + 1000 * g for _ in range(B)]).astype(np.int64)
            cs = rng.integers(0, 6, (B, fk)).astype(np.float64) / 8.0
            bi = np.stack([rng.permutation(500)[:kk] + 1000 * g for _ in range(B)]).astype(np.int64)
            bs = rng.integers(1, 9, (B, kk)).astype(np.float64) / 3.0
            lists.append([t(x) for x in (ci, cs, bi, bs, bs.max(1), np.zeros(B, np.int32))])
        for g in range(G):
            _ffi.check(L.orag_hybrid_push(*[x.data_ptr() for x in lists[g]], None, B, fk, kk, g, G, maxq,
                                          d_peers.data_ptr(), seq, st), "push")
        want, _ = hybrid_merge(torch.stack([pack_local(*lists[g]) for g in range(G)]).contiguous(), fk, kk, 60, k)
        for g in range(G):
            out = ctypes.c_void_p()
            _ffi.check(L.orag_hybrid_wait(bufs[g], G, maxq, B, fk, kk, seq, 2000, ctypes.byref(out), st), "wait")
            got, _ = hybrid_merge(out.value, fk, kk, 60, k, shape=(G, B, 2 * fk + 2 * kk + 2), device=torch.device(dev))
            x_ok &= all(torch.equal(got[key], want[key]) for key in want)
    torch.cuda.synchronize()
    for b in bufs:
        L.orag_exchange_free(b)
    print(f"exchange (two virtual ranks): {'ok' if x_ok else 'MISMATCH'}", flush=True)
    ok &= x_ok

torch.cuda.synchronize()
print("SANITIZE TARGET", "PASSED" if ok else "FAILED", flush=True)
sys.exit(0 if ok else 1)
