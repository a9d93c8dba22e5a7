"""GPU parity tests proper (-m gpu): every kernel, called through the C ABI, against the CPU oracle
on the same seeded inputs and against the committed golden vectors.

Bars (BASELINE.json north_star): BM25 / RRF bit-exact ids, ranks and scores; cosine exact ids with
scores within 1e-5 relative (we additionally assert bitwise equality, which the float64 re-score
achieves), ties broken by chunk id.
"""
import os
import time
from pathlib import Path

import numpy as np
import pytest
import torch

import oracle
from optimized_rag_b200 import synthetic as syn
from conftest import fromhex

pytestmark = pytest.mark.gpu

DEV = "cuda:0"
ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def eng():
    from optimized_rag_b200 import engine
    assert torch.cuda.is_available()
    return engine


def _t(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    return t if dtype is None else t.to(dtype)


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


# ------------------------------------------------------------------------------------------------ generators
def test_gen_embeddings_bit_identical(eng):
    for n, dim, start, dup in [(300, 1536, 0, 0), (257, 64, 1000, 50), (5, 96, 123456789, 0)]:
        got = eng.gen_embeddings(n, dim, start, syn.SEED_CORPUS, dup, device=DEV).cpu().numpy()
        want = syn.embeddings(syn.SEED_CORPUS, start, n, dim, dup)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_gen_tokens_bit_identical(eng):
    for n, start, vocab, lmin, lmax in [(500, 0, 50000, 100, 300), (64, 777, 50, 1, 9)]:
        thr = syn.zipf_thresholds(vocab)
        off, tok = eng.gen_token_corpus(n, start, syn.SEED_TOKENS, thr, vocab, lmin, lmax, device=DEV)
        off_w, tok_w = syn.token_corpus(syn.SEED_TOKENS, start, n, vocab, lmin, lmax, thr)
        assert np.array_equal(off.cpu().numpy(), off_w)
        assert np.array_equal(tok.cpu().numpy(), tok_w)


# ------------------------------------------------------------------------------------------------ cosine, exact path
def _cos_inputs(case):
    corpus = syn.embeddings(syn.SEED_CORPUS, 0, case["n"], case["dim"], case["dup_per_mille"])
    queries = syn.query_embeddings(case["n_queries"], case["n"], case["dim"], dup_per_mille=case["dup_per_mille"])
    if case.get("zero_row") is not None:
        corpus[case["zero_row"], :] = 0.0
    return corpus, queries


def test_cosine_dense_golden_bit_exact(eng, golden):
    for case in golden["cosine"]:
        corpus, queries = _cos_inputs(case)
        idx = eng.CosineIndex(_t(corpus), mode="exact")
        if case["name"] == "zero_query":
            got = idx.dense(torch.zeros((1, case["dim"]), dtype=torch.float32, device=DEV)).cpu().numpy()[0]
            assert np.array_equal(got, np.array([fromhex(x) for x in case["zero_query_scores"]]))
            continue
        got = idx.dense(_t(queries)).cpu().numpy()
        want = np.array([[fromhex(x) for x in row] for row in case["scores"]])
        assert np.array_equal(_bits(got), _bits(want)), case["name"]


@pytest.mark.parametrize("n,dim,nq,dup", [(1000, 1536, 9, 0), (3000, 128, 17, 30), (37, 100, 3, 0), (5, 64, 2, 0)])
def test_cosine_exact_topk_vs_oracle(eng, n, dim, nq, dup):
    corpus = syn.embeddings(syn.SEED_CORPUS, 0, n, dim, dup)
    queries = syn.query_embeddings(nq, n, dim, dup_per_mille=dup)
    idx = eng.CosineIndex(_t(corpus), row_id_base=1000, mode="exact")
    ids, sc = idx.topk(_t(queries), 10)
    wi, ws = oracle.cosine_topk(corpus, queries, 10, id_base=1000)
    assert np.array_equal(ids.cpu().numpy(), wi)
    valid = wi >= 0
    assert np.array_equal(_bits(sc.cpu().numpy()[valid]), _bits(ws[valid]))


# ------------------------------------------------------------------------------------------------ cosine, tensor-core path
@pytest.mark.parametrize("mode,eps", [("tf32", 2.3e-3), ("bf16", 8.1e-3), ("f16", 1.3e-3)])
def test_firstpass_dense_within_bound(eng, mode, eps):
    n, dim, nq = 1000, 1536, 70
    corpus = syn.embeddings(syn.SEED_CORPUS, 0, n, dim)
    queries = syn.query_embeddings(nq, n, dim)
    idx = eng.CosineIndex(_t(corpus), mode=mode)
    got = idx.firstpass_dense(_t(queries), mode).cpu().numpy().astype(np.float64)  # [n, nq] = cos * |q|
    c64, q64 = corpus.astype(np.float64), queries.astype(np.float64)
    want = (c64 @ q64.T) / np.linalg.norm(c64, axis=1)[:, None]
    err = np.abs(got - want) / np.linalg.norm(q64, axis=1)[None, :]
    assert err.max() < eps, (mode, err.max())
    # and it is a real low-precision product, not garbage that happens to be small
    assert np.corrcoef(got.ravel(), want.ravel())[0, 1] > 0.999


@pytest.mark.parametrize("mode", ["tf32", "bf16", "f16"])
@pytest.mark.parametrize("n,dim,nq,dup", [(20000, 1536, 40, 0), (9000, 256, 256, 20), (4500, 64, 3, 0)])
def test_cosine_tc_topk_vs_oracle(eng, mode, n, dim, nq, dup):
    corpus = syn.embeddings(syn.SEED_CORPUS, 0, n, dim, dup)
    queries = syn.query_embeddings(nq, n, dim, dup_per_mille=dup)
    if n == 9000:
        corpus[4000:4040] = corpus[17]  # a block of exact duplicates straddling the top-k boundary
        queries[5] = 0.0               # zero query -> all cosines 0.0 -> ids 0..k-1
    idx = eng.CosineIndex(_t(corpus), row_id_base=7, mode=mode)
    ids, sc = idx.topk(_t(queries), 10)
    sub = list(range(0, nq, max(1, nq // 12)))
    if n == 9000:
        sub = sorted(set(sub + [5, 17 * 0 + 2]))
    wi, ws = oracle.cosine_topk(corpus, queries[sub], 10, id_base=7)
    assert np.array_equal(ids.cpu().numpy()[sub], wi), mode
    assert np.array_equal(_bits(sc.cpu().numpy()[sub]), _bits(ws)), mode


def test_cosine_tc_matches_exact_path_all_queries(eng):
    n, dim, nq = 30000, 1536, 256
    corpus = eng.gen_embeddings(n, dim, 0, syn.SEED_CORPUS, 1, device=DEV)
    queries = _t(syn.query_embeddings(nq, n, dim, dup_per_mille=1))
    a = eng.CosineIndex(corpus, mode="exact").topk(queries, 10)
    for mode in ("tf32", "bf16", "f16"):
        b = eng.CosineIndex(corpus, mode=mode).topk(queries, 10)
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]), mode


def test_cosine_f16_rows_of_wildly_different_scale(eng):
    """The fp16 shadow scales every row (and every query) by its own power of two: rows and queries whose
    magnitudes span 60 binary orders, far outside fp16's range, must still give the exact top-k (kernels must not
    assume unit norm: the reference divides by both magnitudes, rag/retrieval.py:365-371)."""
    n, dim, nq = 12000, 256, 48
    rng = np.random.default_rng(8)
    corpus = syn.embeddings(syn.SEED_CORPUS, 0, n, dim)
    corpus *= (2.0 ** rng.integers(-30, 31, n)).astype(np.float32)[:, None]
    corpus[100] = 0.0
    queries = syn.query_embeddings(nq, n, dim)
    queries *= (2.0 ** rng.integers(-30, 31, nq)).astype(np.float32)[:, None]
    queries[3] = 0.0
    idx = eng.CosineIndex(_t(corpus), mode="f16")
    ids, sc = idx.topk(_t(queries), 10)
    wi, ws = oracle.cosine_topk(corpus, queries, 10)
    assert np.array_equal(ids.cpu().numpy(), wi)
    assert np.array_equal(_bits(sc.cpu().numpy()), _bits(ws))


# ------------------------------------------------------------------------------------------------ BM25
def _bm25_case(n, vocab, lmin, lmax, nq, min_rank, tile_docs, eng):
    thr = syn.zipf_thresholds(vocab)
    doc_off, tok = syn.token_corpus(syn.SEED_TOKENS, 0, n, vocab, lmin, lmax, thr)
    qtok, qlen = syn.keyword_queries(nq, vocab, min_rank=min_rank, thresholds=thr)
    from optimized_rag_b200.bm25_index import Bm25Index
    ix = Bm25Index(_t(doc_off), _t(tok), vocab, tile_docs=tile_docs, doc_id_base=100)
    orc = oracle.BM25Index(doc_off, tok, vocab)
    return ix, orc, qtok, qlen


def test_bm25_golden_bit_exact(eng, golden):
    from optimized_rag_b200.bm25_index import Bm25Index
    for case in golden["bm25"]["cases"]:
        thr = syn.zipf_thresholds(case["vocab"])
        doc_off, tok = syn.token_corpus(syn.SEED_TOKENS, 0, case["n"], case["vocab"], case["lmin"], case["lmax"], thr)
        ix = Bm25Index(_t(doc_off), _t(tok), case["vocab"], tile_docs=32)
        mt = max(len(q["terms"]) for q in case["queries"])
        qt = np.full((len(case["queries"]), mt), -2, dtype=np.int32)
        ql = np.zeros(len(case["queries"]), dtype=np.int32)
        for i, q in enumerate(case["queries"]):
            qt[i, :len(q["terms"])] = q["terms"]
            ql[i] = len(q["terms"])
        raw = ix.dense_scores(_t(qt), _t(ql)).cpu().numpy()
        for i, q in enumerate(case["queries"]):
            want = np.array([fromhex(x) for x in q["normalized"]])
            m = raw[i].max() if raw[i].max() > 0 else 1.0
            assert np.array_equal(_bits(raw[i] / m), _bits(want)), (case["name"], i)
        # and the top-k entry point, both paths, equals the oracle's ranking of those scores
        for force in ("dense", "sparse", "exact_tiles"):
            if force != "dense" and ix.has_negative_idf:
                continue
            ids, sc, mx = ix.topk(_t(qt), _t(ql), 10, normalize=True, force=force)
            for i, q in enumerate(case["queries"]):
                want = np.array([fromhex(x) for x in q["normalized"]])
                wi, wv = oracle.topk(want, 10)
                got_i = ids[i].cpu().numpy()
                assert np.array_equal(got_i[:len(wi)], wi), (case["name"], force, i)
                assert np.array_equal(_bits(sc[i].cpu().numpy()[:len(wi)]), _bits(wv)), (case["name"], force, i)


@pytest.mark.parametrize("force", ["sparse", "exact_tiles", "dense"])
@pytest.mark.parametrize("n,vocab,lmin,lmax,nq,min_rank,tile", [
    (20000, 5000, 20, 120, 64, 20, 2048),
    (3000, 300, 5, 60, 40, 3, 256),
    (70000, 50000, 100, 300, 32, 100, 1024),
])
def test_bm25_topk_vs_oracle(eng, force, n, vocab, lmin, lmax, nq, min_rank, tile):
    ix, orc, qtok, qlen = _bm25_case(n, vocab, lmin, lmax, nq, min_rank, tile, eng)
    assert ix.avgdl == orc.avgdl and ix.eps == orc.eps
    assert np.array_equal(_bits(ix.idf.cpu().numpy()), _bits(orc.idf))
    if force != "dense" and ix.has_negative_idf:
        pytest.skip("negative idf -> dense path only")
    if force == "sparse":
        assert ix.postings_r16 is not None  # the MaxScore first pass is the path under test
    for normalize in (True, False):
        ids, sc, mx = ix.topk(_t(qtok), _t(qlen), 10, normalize=normalize, force=force)
        for b in range(nq):
            raw = orc.scores_raw(qtok[b, :qlen[b]])
            m = raw.max() if raw.max() > 0 else 1.0
            want = raw / m if normalize else raw
            wi, wv = oracle.topk(want, 10, id_base=100)
            assert np.array_equal(ids[b].cpu().numpy(), wi), (force, normalize, b)
            assert np.array_equal(_bits(sc[b].cpu().numpy()), _bits(wv)), (force, normalize, b)
            assert float(mx[b]) == (m if normalize else max(raw.max(), 0.0))


def test_bm25_rare_terms_zero_fill(eng):
    """Queries matching fewer than k docs: the list continues with zero-score docs in id order."""
    ix, orc, _, _ = _bm25_case(5000, 5000, 20, 120, 1, 20, 1024, eng)
    df = orc.df
    rare = [int(t) for t in np.nonzero((df > 0) & (df < 4))[0][:3]]
    qt = np.array([rare + [-1], [rare[0], rare[0], -2, -2], [-1, -1, -2, -2]], dtype=np.int32)
    ql = np.array([4, 2, 2], dtype=np.int32)
    for force in ("sparse", "exact_tiles", "dense"):
        ids, sc, _ = ix.topk(_t(qt), _t(ql), 10, force=force)
        for b in range(3):
            norm, _ = orc.scores(qt[b, :ql[b]])
            wi, wv = oracle.topk(norm, 10, id_base=100)
            assert np.array_equal(ids[b].cpu().numpy(), wi), (force, b)
            assert np.array_equal(_bits(sc[b].cpu().numpy()), _bits(wv)), (force, b)


# ------------------------------------------------------------------------------------------------ RRF / merge
def test_rrf_golden_bit_exact(eng, golden):
    for case in golden["rrf"]:
        L = len(case["lists"])
        n = max(1, max((len(l) for l in case["lists"]), default=1))
        arr = np.full((1, L, n), -1, dtype=np.int64)
        for i, l in enumerate(case["lists"]):
            arr[0, i, :len(l)] = l
        ids, sc = eng.rrf_fuse(_t(arr), case["k"], case["top_k"])
        got_ids = [int(x) for x in ids[0].cpu().numpy() if x >= 0]
        assert got_ids == case["ids"], case["name"]
        assert sc[0].cpu().numpy()[:len(got_ids)].tolist() == [fromhex(x) for x in case["scores"]], case["name"]


def test_rrf_batch_vs_oracle(eng):
    rng = np.random.default_rng(3)
    B, L, n = 300, 3, 12
    arr = np.stack([np.stack([rng.permutation(30)[:n] for _ in range(L)]) for _ in range(B)]).astype(np.int64)
    arr[5, 1, 7:] = -1
    for tie in ("reference", "chunk_id"):
        ids, sc, src = eng.rrf_fuse(_t(arr), 60, 10, tie=tie, want_src=True)
        for b in range(B):
            lists = [[int(x) for x in arr[b, l] if x >= 0] for l in range(L)]
            wi, ws = oracle.rrf_fuse(lists, 60, 10, tie=tie)
            assert np.array_equal(ids[b].cpu().numpy(), wi) and np.array_equal(sc[b].cpu().numpy(), ws)
            for r in range(10):
                for l in range(L):
                    want_rank = lists[l].index(int(wi[r])) + 1 if int(wi[r]) in lists[l] else 0
                    assert int(src[b, r, l]) == want_rank


def test_topk_merge(eng):
    rng = np.random.default_rng(5)
    B, m = 50, 48
    ids = np.stack([rng.permutation(1000)[:m] for _ in range(B)]).astype(np.int64)
    sc = rng.integers(0, 8, (B, m)).astype(np.float64) / 4.0  # many ties
    ids[3, 10:] = -1
    gi, gs, _ = eng.topk_merge(_t(ids), _t(sc), 10)
    for b in range(B):
        ok = ids[b] >= 0
        order = sorted(np.nonzero(ok)[0], key=lambda i: (-sc[b, i], ids[b, i]))[:10]
        assert gi[b].cpu().numpy()[:len(order)].tolist() == ids[b, order].tolist()
        assert gs[b].cpu().numpy()[:len(order)].tolist() == sc[b, order].tolist()
    smax = rng.random((B, 4)) * 3
    gi, gs, gm = eng.topk_merge(_t(ids), _t(sc), 10, shard_max=_t(smax))
    for b in range(B):
        m_ = max(smax[b].max(), sc[b][ids[b] >= 0].max())
        assert float(gm[b]) == m_
        ok = ids[b] >= 0
        norm = sc[b] / m_
        order = sorted(np.nonzero(ok)[0], key=lambda i: (-norm[i], ids[b, i]))[:10]
        assert gs[b].cpu().numpy()[:len(order)].tolist() == norm[order].tolist()


def test_hybrid_merge_equals_merge_then_rrf(eng):
    """The single post-all-gather launch (orag_hybrid_merge) == the step-by-step path it replaces, which is
    itself pinned to the oracle above (topk_merge with global-max normalisation, then rrf_fuse)."""
    from optimized_rag_b200.dist import hybrid_merge, pack_local, unpack_gathered
    rng = np.random.default_rng(11)
    G, B, fk, kk, k = 4, 37, 10, 16, 10
    shards = []
    for g in range(G):
        ci = np.stack([rng.permutation(500)[:fk] + 1000 * g for _ in range(B)]).astype(np.int64)
        cs = rng.integers(0, 6, (B, fk)).astype(np.float64) / 8.0          # many exact ties across shards
        bi = np.stack([rng.permutation(500)[:kk] + 1000 * g for _ in range(B)]).astype(np.int64)
        bs = rng.integers(1, 9, (B, kk)).astype(np.float64) / 3.0
        if g == 1:
            ci[:, fk - 2:] = -1   # a shard with fewer than fetch_k rows
            bi[5, :] = -1         # a shard where the query matched nothing
            bs[5, :] = 0.0
        bm = np.where((bi >= 0).any(1), np.where(bi >= 0, bs, 0).max(1), 0.0)
        st = np.zeros(B, np.int32)
        st[7] = 1 if g == 2 else 0
        shards.append(pack_local(_t(ci), _t(cs), _t(bi), _t(bs), _t(bm), _t(st)))
    buf = torch.stack(shards).contiguous()
    out, status = hybrid_merge(buf, fk, kk, 60, k)
    gci, gcs, gbi, gbs, gbm, gst = unpack_gathered(buf, fk, kk)
    ci, cs, _ = eng.topk_merge(gci, gcs, fk)
    bi, bs, bmax = eng.topk_merge(gbi, gbs, fk, shard_max=gbm)
    fi, fs, src = eng.rrf_fuse(torch.stack([ci, bi], dim=1).contiguous(), 60, k, want_src=True)
    for key, want in (("cos_ids", ci), ("cos_scores", cs), ("bm25_ids", bi), ("bm25_scores", bs), ("bm25_max", bmax),
                      ("ids", fi), ("rrf_scores", fs), ("src_ranks", src)):
        assert torch.equal(out[key], want), key
    # status = OR of the shards' words, plus the truncation rule of the BM25 lists (csrc/rrf.cu): a full shard list whose
    # last entry normalises to the k-th merged value while a different raw value (an entry of the list, or the double just
    # below the last one) lands on the same normalised double
    want_status = []
    gbs_np, gbi_np = gbs.cpu().numpy().reshape(B, G, kk), gbi.cpu().numpy().reshape(B, G, kk)
    for b in range(B):
        flag = 1 if b == 7 else 0
        M = float(bmax[b].item())
        vk = float(bs[b, fk - 1].item())
        if int(bi[b, fk - 1].item()) >= 0:
            for g in range(G):
                raws = gbs_np[b, g]
                if gbi_np[b, g, kk - 1] < 0 or raws[-1] / M != vk:
                    continue
                if (raws[-1] > 0 and np.nextafter(raws[-1], 0.0) / M == vk) or \
                        any(r != raws[-1] and r / M == vk for r in raws[:-1]):
                    flag |= 1
        want_status.append(flag)
    assert status.cpu().tolist() == want_status and sum(want_status) > 1


def _random_lists(rng, g, B, fk, kk):
    ci = np.stack([rng.permutation(500)[:fk] + 1000 * g for _ in range(B)]).astype(np.int64)
    cs = rng.integers(0, 6, (B, fk)).astype(np.float64) / 8.0
    bi = np.stack([rng.permutation(500)[:kk] + 1000 * g for _ in range(B)]).astype(np.int64)
    bs = rng.integers(1, 9, (B, kk)).astype(np.float64) / 3.0
    bm = bs.max(1)
    st = (rng.integers(0, 20, B) == 0).astype(np.int32)
    return [_t(x) for x in (ci, cs, bi, bs, bm, st)]


def test_peer_exchange_kernels_two_virtual_ranks(eng):
    """exchange_push_kernel + exchange_wait_kernel (csrc/exchange.cu) with both 'ranks' in this process (two local
    buffers instead of IPC-mapped ones): after the pushes every rank's slot must be exactly the array
    pack + all-gather builds, search after search (slot alternation, shrinking batch), and the merge over it
    must equal the merge over the packed tensor."""
    import ctypes
    from optimized_rag_b200 import _ffi
    from optimized_rag_b200.dist import hybrid_merge, pack_local
    L = _ffi.lib()
    G, maxq, fk, kk, k = 2, 64, 10, 16, 10
    W = 2 * fk + 2 * kk + 2
    nbytes = int(L.orag_exchange_bytes(G, maxq, fk, kk))
    assert nbytes == 256 + 6 * G * maxq * W * 8   # flags (padded) + six slots
    bufs = []
    for _ in range(G):
        p = ctypes.c_void_p()
        _ffi.check(L.orag_exchange_alloc(nbytes, ctypes.byref(p)), "alloc")
        bufs.append(p.value)
    try:
        d_peers = torch.tensor(bufs, dtype=torch.int64, device=DEV)
        st = torch.cuda.current_stream().cuda_stream
        rng = np.random.default_rng(5)
        for seq, B in enumerate([64, 64, 37, 64, 1, 64, 5, 64], start=1):   # wraps around the six slots
            lists = [_random_lists(rng, g, B, fk, kk) for g in range(G)]
            for g in range(G):
                _ffi.check(L.orag_hybrid_push(*[t.data_ptr() for t in lists[g]], None, B, fk, kk, g, G, maxq,
                                              d_peers.data_ptr(), seq, st), "push")
            want = torch.stack([pack_local(*lists[g]) for g in range(G)]).contiguous()
            ref, ref_status = hybrid_merge(want, fk, kk, 60, k)
            for g in range(G):
                out = ctypes.c_void_p()
                _ffi.check(L.orag_hybrid_wait(bufs[g], G, maxq, B, fk, kk, seq, 2000, ctypes.byref(out), st), "wait")
                assert out.value == bufs[g] + 256 + (seq % 6) * G * maxq * W * 8
                got, status = hybrid_merge(out.value, fk, kk, 60, k, shape=(G, B, W), device=torch.device(DEV))
                for key in ref:
                    assert torch.equal(got[key], ref[key]), (seq, g, key)
                assert torch.equal(status, ref_status)
        # a block that never arrives: the wait gives up after the timeout and flags every query
        out = ctypes.c_void_p()
        t0 = time.perf_counter()
        _ffi.check(L.orag_hybrid_wait(bufs[0], G, maxq, 8, fk, kk, 99, 50, ctypes.byref(out), st), "wait")
        _, status = hybrid_merge(out.value, fk, kk, 60, k, shape=(G, 8, W), device=torch.device(DEV))
        assert bool(((status & _ffi.ORAG_STATUS_EXCHANGE_TIMEOUT) != 0).all())
        assert time.perf_counter() - t0 < 5.0
    finally:
        torch.cuda.synchronize()
        for b in bufs:
            _ffi.check(L.orag_exchange_free(b), "free")


_PEER_WORKER = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["ORAG_ROOT"])
from optimized_rag_b200.dist import PeerExchange, hybrid_merge, pack_local
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
dev = torch.device("cuda", 0)            # both ranks on one GPU: CUDA IPC works between processes of one device
torch.cuda.set_device(dev)
fk, kk, k, B = 10, 16, 10, 48
def lists(seed, g, n):
    rng = np.random.default_rng(seed * 100 + g)
    ci = np.stack([rng.permutation(500)[:fk] + 1000 * g for _ in range(n)]).astype(np.int64)
    cs = rng.integers(0, 6, (n, fk)).astype(np.float64) / 8.0
    bi = np.stack([rng.permutation(500)[:kk] + 1000 * g for _ in range(n)]).astype(np.int64)
    bs = rng.integers(1, 9, (n, kk)).astype(np.float64) / 3.0
    st = (rng.integers(0, 20, n) == 0).astype(np.int32)
    return [torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (ci, cs, bi, bs, bs.max(1), st)]
# setup failure on ONE rank (its IPC open is made to fail) must surface as the same exception on EVERY rank
from optimized_rag_b200 import _ffi
L = _ffi.lib()
real_open = L.orag_exchange_open
if rank == 1:
    L.orag_exchange_open = lambda *a: -2
try:
    PeerExchange(dev, 64, fk, kk)
    ok = False
except _ffi.OragError:
    ok = True
L.orag_exchange_open = real_open
px = PeerExchange(dev, 64, fk, kk)
for step, n in enumerate([B, B, 7, 64, B, B]):
    # Both ranks share ONE GPU here, and kernels of different processes that wait on one another are not guaranteed
    # to run at the same time (B200_PROFILING.md): push, make sure on the HOST that every rank's push has finished,
    # and only then launch the wait kernel -- it finds all sequence numbers published and never spins.
    px.push(*lists(step, rank, n))
    torch.cuda.synchronize()
    dist.barrier()
    ptr, shape = px.wait(n)
    got, status = hybrid_merge(ptr, fk, kk, 60, k, shape=shape, device=dev)
    want = torch.stack([pack_local(*lists(step, g, n)) for g in range(world)]).contiguous()
    ref, ref_status = hybrid_merge(want, fk, kk, 60, k)
    ok &= all(torch.equal(got[key], ref[key]) for key in ref) and torch.equal(status, ref_status)
    torch.cuda.synchronize()
    dist.barrier()          # nobody pushes search s+1 into a slot a peer may still be reading
torch.cuda.synchronize()
px.close()
dist.destroy_process_group()
print("PEER_OK" if ok else "PEER_MISMATCH", flush=True)
sys.exit(0 if ok else 1)
"""


def test_peer_exchange_across_processes_via_cuda_ipc(tmp_path):
    """PeerExchange end to end between two PROCESSES (IPC export/open, remote stores, sequence numbers, slot re-use,
    close; a setup failure on one rank raises on both): both ranks share this box's GPU, the host rendezvous is gloo."""
    import socket
    import subprocess
    import sys
    script = tmp_path / "peer_worker.py"
    script.write_text(_PEER_WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   ORAG_ROOT=str(Path(__file__).resolve().parents[1]))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        try:
            o, _ = p.communicate(timeout=240)
        except subprocess.TimeoutExpired:
            p.kill()
            o, _ = p.communicate()
        outs.append(o)
    assert all(p.returncode == 0 for p in procs) and all("PEER_OK" in o for o in outs), "\n".join(outs)


# ------------------------------------------------------------------------------------------------ pairwise / config 1 / hybrid
def test_pairwise_golden(eng, golden):
    g = golden["pairwise"]
    m, d = g["m"], g["dim"]
    emb = syn.embeddings(syn.SEED_CORPUS, 0, m, d)
    for i in range(0, m, 4):
        emb[i + 1] = (emb[i] + np.float32(0.25) * emb[i + 1]).astype(np.float32)
        emb[i + 3] = (emb[i] + np.float32(0.5) * emb[i + 3]).astype(np.float32)
    oi, oj, sim = eng.pairwise_cosine_threshold(_t(emb), _t(np.array(g["doc_idx"], dtype=np.int32)), g["threshold"])
    got = [[int(a), int(b), round(float(s), 3)] for a, b, s in zip(oi.cpu(), oj.cpu(), sim.cpu())]
    assert got == g["pairs"]
    wi, wj, ws = oracle.pairwise_candidates(emb, g["doc_idx"], g["threshold"])
    assert np.array_equal(_bits(sim.cpu().numpy()), _bits(ws))


def test_config1_golden(eng, golden):
    from optimized_rag_b200.bm25_index import Bm25Index
    g = golden["config1"]
    dim = g["dim"]
    emb = np.concatenate([syn.embeddings(s, 0, 1, dim) for s in g["chunk_seeds"]], axis=0)
    toks = np.concatenate([np.array(t, dtype=np.int32) for t in g["chunk_tokens"]])
    off = np.cumsum([0] + [len(t) for t in g["chunk_tokens"]]).astype(np.int64)
    shard = eng.HybridShard(eng.CosineIndex(_t(emb), mode="exact"),
                            Bm25Index(_t(off), _t(toks), g["vocab_size"], tile_docs=32))
    nq = len(g["queries"])
    mt = max(len(q["query_terms"]) for q in g["queries"])
    qt = np.full((nq, mt), -2, dtype=np.int32)
    ql = np.zeros(nq, dtype=np.int32)
    qe = np.zeros((nq, dim), dtype=np.float32)
    for i, q in enumerate(g["queries"]):
        qt[i, :len(q["query_terms"])] = q["query_terms"]
        ql[i] = len(q["query_terms"])
        qe[i] = (syn.embeddings(q["query_noise_seed"], 0, 1, dim)[0]
                 + np.float32(0.75) * emb[q["query_base_chunk"]]).astype(np.float32)
    res = shard.search(_t(qe), _t(qt), _t(ql), k=10)
    for i, q in enumerate(g["queries"]):
        assert res["cos_ids"][i].cpu().tolist() == q["cos_ids"]
        assert res["cos_scores"][i].cpu().tolist() == [fromhex(x) for x in q["cos_scores"]]
        assert res["bm25_ids"][i].cpu().tolist() == q["bm25_ids"]
        assert res["bm25_scores"][i].cpu().tolist() == [fromhex(x) for x in q["bm25_scores"]]
        assert res["ids"][i].cpu().tolist() == q["rrf_ids"]
        assert res["rrf_scores"][i].cpu().tolist() == [fromhex(x) for x in q["rrf_scores"]]


def test_hybrid_vs_oracle_mid_size(eng):
    from optimized_rag_b200.bm25_index import Bm25Index
    n, dim, vocab, nq = 12000, 256, 3000, 48
    thr = syn.zipf_thresholds(vocab)
    corpus = syn.embeddings(syn.SEED_CORPUS, 0, n, dim, 2)
    queries = syn.query_embeddings(nq, n, dim, dup_per_mille=2)
    doc_off, tok = syn.token_corpus(syn.SEED_TOKENS, 0, n, vocab, 20, 100, thr)
    qtok, qlen = syn.keyword_queries(nq, vocab, min_rank=10, thresholds=thr)
    shard = eng.HybridShard(eng.CosineIndex(_t(corpus), mode="tf32"),
                            Bm25Index(_t(doc_off), _t(tok), vocab, tile_docs=2048))
    res = shard.search(_t(queries), _t(qtok), _t(qlen), k=10)
    orc = oracle.BM25Index(doc_off, tok, vocab)
    want = oracle.hybrid_topk(corpus, queries, orc, [qtok[b, :qlen[b]] for b in range(nq)], k=10)
    for b in range(nq):
        assert res["ids"][b].cpu().tolist() == want[b]["ids"].tolist(), b
        assert res["rrf_scores"][b].cpu().tolist() == want[b]["rrf_scores"].tolist(), b
        assert res["cos_ids"][b].cpu().tolist() == want[b]["cos_ids"].tolist(), b
        assert res["bm25_ids"][b].cpu().tolist() == want[b]["bm25_ids"].tolist(), b


def test_pairwise_tensor_core_equals_exact(eng):
    """Config-5 path: tcgen05 first pass + float64 re-score finds exactly the pairs of the exact sweep."""
    m, d = 3000, 256
    emb = syn.embeddings(syn.SEED_CORPUS, 0, m, d)
    rng = np.random.default_rng(11)
    for i in range(0, m, 7):  # planted near-duplicates at assorted similarities around the threshold
        j = int(rng.integers(0, m))
        w = np.float32(rng.choice([0.3, 0.55, 0.6, 0.62, 0.65, 0.8]))
        emb[i] = (emb[j] + w * emb[i]).astype(np.float32)
    emb[5] = 0.0
    doc = (np.arange(m) // 4).astype(np.int32)
    a = eng.pairwise_cosine_threshold(_t(emb), _t(doc), 0.85, mode="exact")
    b = eng.pairwise_cosine_threshold(_t(emb), _t(doc), 0.85, mode="tc")
    assert len(a[0]) > 50
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    wi, wj, ws = oracle.pairwise_candidates(emb[:400], doc[:400], 0.85)
    c = eng.pairwise_cosine_threshold(_t(emb[:400]), _t(doc[:400]), 0.85, mode="tc")
    assert np.array_equal(c[0].cpu().numpy(), wi) and np.array_equal(c[1].cpu().numpy(), wj)
    assert np.array_equal(_bits(c[2].cpu().numpy()), _bits(ws))


# ------------------------------------------------------------------------------------------------ edge cases
@pytest.mark.parametrize("mode", ["tf32", "bf16", "f16"])
def test_cosine_tc_ragged_sizes_and_batches(eng, mode):
    """N not a multiple of the 128-row tile, B > 256 (two query groups), B = 1, k = 1, 64 and 128, id base."""
    n, dim = 4097 + 128 * 3 + 5, 192
    corpus = syn.embeddings(syn.SEED_CORPUS, 0, n, dim, 3)
    idx = eng.CosineIndex(_t(corpus), row_id_base=10_000_000_000, mode=mode)
    for nq, k in [(300, 10), (1, 1), (7, 64), (5, 128)]:
        queries = syn.query_embeddings(nq, n, dim, dup_per_mille=3)
        ids, sc = idx.topk(_t(queries), k)
        sub = sorted(set([0, nq - 1, nq // 2, min(nq - 1, 257)]))
        wi, ws = oracle.cosine_topk(corpus, queries[sub], k, id_base=10_000_000_000)
        assert np.array_equal(ids.cpu().numpy()[sub], wi), (mode, nq, k)
        assert np.array_equal(_bits(sc.cpu().numpy()[sub]), _bits(ws)), (mode, nq, k)


def test_cosine_candidate_overflow_falls_back_to_exact_scan(eng):
    """6000 identical rows tie for the top: more near-ties than candidate slots -> status bit -> exact scan."""
    n, dim = 12000, 64
    corpus = syn.embeddings(syn.SEED_CORPUS, 0, n, dim)
    corpus[3000:9000] = corpus[11]
    queries = np.stack([corpus[11], syn.query_embeddings(1, n, dim)[0]]).astype(np.float32)
    idx = eng.CosineIndex(_t(corpus), mode="tf32")
    ids, sc = idx.topk(_t(queries), 10)
    wi, ws = oracle.cosine_topk(corpus, queries, 10)
    assert np.array_equal(ids.cpu().numpy(), wi) and np.array_equal(_bits(sc.cpu().numpy()), _bits(ws))
    assert wi[0].tolist() == [11] + list(range(3000, 3009))  # ties broken by chunk id
    raw_ids, _ = idx.topk(_t(queries), 10, check_overflow=False)  # without the fallback the flag must be set
    st = []
    idx.topk(_t(queries), 10, check_overflow=False, status_out=st)
    assert int(st[0][0]) == 1 and int(st[0][1]) == 0


def test_cosine_k_larger_than_corpus_and_auto_mode(eng):
    corpus = syn.embeddings(syn.SEED_CORPUS, 0, 6, 100)  # dim % 32 != 0 -> auto picks the exact path
    idx = eng.CosineIndex(_t(corpus))
    assert idx.mode == "exact"
    ids, sc = idx.topk(_t(syn.query_embeddings(2, 6, 100)), 10)
    assert (ids[:, 6:] == -1).all() and (ids[:, :6] >= 0).all() and (sc[:, 6:] == 0).all()
    assert sorted(ids[0, :6].cpu().tolist()) == list(range(6))


def test_bm25_long_and_degenerate_queries(eng):
    """Queries longer than 32 terms (chunked path), all-OOV, empty, single query, duplicate-heavy."""
    ix, orc, _, _ = _bm25_case(6000, 800, 10, 80, 1, 5, 512, eng)
    rng = np.random.default_rng(4)
    mt = 48
    rows = [rng.integers(0, 800, 48), rng.integers(300, 800, 33), np.full(5, -1), np.zeros(0, int),
            np.array([7, 7, 7, 7, 9]), rng.integers(0, 20, 40)]
    qt = np.full((len(rows), mt), -2, dtype=np.int32)
    ql = np.zeros(len(rows), dtype=np.int32)
    for i, r in enumerate(rows):
        qt[i, :len(r)] = r
        ql[i] = len(r)
    for force in ("sparse", "exact_tiles", "dense"):
        if force != "dense" and ix.has_negative_idf:
            continue
        ids, sc, mx = ix.topk(_t(qt), _t(ql), 10, force=force)
        for b in range(len(rows)):
            norm, m = orc.scores(qt[b, :ql[b]])
            wi, wv = oracle.topk(norm, 10, id_base=100)
            assert np.array_equal(ids[b].cpu().numpy(), wi), (force, b)
            assert np.array_equal(_bits(sc[b].cpu().numpy()), _bits(wv)), (force, b)
            assert float(mx[b]) == m
    one = ix.topk(_t(qt[:1]), _t(ql[:1]), 10, force="sparse" if not ix.has_negative_idf else "dense")
    assert torch.equal(one[0][0], ids[0])


def _check_bm25_vs_oracle(ix, orc, qt, ql, force, k=10, **kw):
    ids, sc, mx = ix.topk(_t(qt), _t(ql), k, force=force, **kw)
    for b in range(qt.shape[0]):
        norm, m = orc.scores(qt[b, :ql[b]])
        wi, wv = oracle.topk(norm, k, id_base=ix.doc_id_base)
        assert np.array_equal(ids[b].cpu().numpy(), wi), (force, kw, b)
        assert np.array_equal(_bits(sc[b].cpu().numpy()), _bits(wv)), (force, kw, b)
        assert float(mx[b]) == m


@pytest.mark.parametrize("fp_tile", [32, 256, 4096, 16384])
def test_bm25_first_pass_tile_sizes_and_shapes(eng, fp_tile):
    """MaxScore first pass (csrc/bm25_ms.cu) over every first-pass tile size, in the stand-alone and in the
    background working-set shape, with queries of up to 32 tokens incl. repeated and OOV tokens."""
    from optimized_rag_b200.bm25_index import Bm25Index
    vocab, n = 3000, 40000
    thr = syn.zipf_thresholds(vocab)
    doc_off, tok = syn.token_corpus(syn.SEED_TOKENS, 0, n, vocab, 30, 90, thr)
    ix = Bm25Index(_t(doc_off), _t(tok), vocab, tile_docs=1024, fp_tile_docs=fp_tile, doc_id_base=7)
    assert ix.postings_r16 is not None and ix.fp_tile_docs == fp_tile
    orc = oracle.BM25Index(doc_off, tok, vocab)
    rng = np.random.default_rng(fp_tile)
    qt, ql = syn.keyword_queries(24, vocab, min_rank=30, thresholds=thr)
    long_q = np.full((8, 32), -2, dtype=np.int32)
    long_len = np.zeros(8, dtype=np.int32)
    for i in range(8):
        m = int(rng.integers(9, 33))
        row = rng.integers(30, vocab, m)
        row[rng.integers(0, m, 3)] = row[0]          # repeated tokens count twice in the reference
        row[rng.integers(0, m)] = vocab + 5           # OOV
        long_q[i, :m], long_len[i] = row, m
    wide = np.full((qt.shape[0], 32), -2, dtype=np.int32)
    wide[:, :qt.shape[1]] = qt
    qt2, ql2 = np.concatenate([wide, long_q]), np.concatenate([ql, long_len])
    for background in (False, True):
        _check_bm25_vs_oracle(ix, orc, qt2, ql2, "sparse", background=background)
    for k in (1, 32, 50, 100):   # k > 32 switches the in-tile local threshold off; 10 / 16 / 32 / 64 / 128 = warm-start levels
        _check_bm25_vs_oracle(ix, orc, qt2[:6], ql2[:6], "sparse", k=k)


def test_bm25_first_pass_skewed_docs_overflow_and_repair(eng):
    """All postings of the query's term sit in one narrow doc range: the sub-range splitter (which assumes
    roughly uniform docs) marks more docs than the compact accumulator holds -> the per-query overflow flag is
    raised and the caller's repair path (dense kernel) returns the exact list."""
    from optimized_rag_b200.bm25_index import Bm25Index
    vocab, n, hot = 50, 20000, 1500
    rng = np.random.default_rng(3)
    docs = []
    for d in range(n):
        body = rng.integers(2, vocab, int(rng.integers(5, 12))).tolist()
        if d < hot:
            body += [0] * int(rng.integers(1, 4))     # term 0 only in docs [0, hot)
        if d % 7 == 0:
            body.append(1)
        docs.append(body)
    doc_off = np.zeros(n + 1, dtype=np.int64)
    doc_off[1:] = np.cumsum([len(x) for x in docs])
    tok = np.concatenate([np.asarray(x, dtype=np.int32) for x in docs])
    ix = Bm25Index(_t(doc_off), _t(tok), vocab, tile_docs=2048, fp_tile_docs=8192)
    orc = oracle.BM25Index(doc_off, tok, vocab)
    assert ix.postings_r16 is not None
    qt = np.array([[0, -2], [0, 1]], dtype=np.int32)
    ql = np.array([1, 2], dtype=np.int32)
    st = []
    ix.topk(_t(qt), _t(ql), 10, force="sparse", check_overflow=False, status_out=st)
    assert int(st[0][0]) & 1 == 1, "the skewed query must raise the overflow flag"
    _check_bm25_vs_oracle(ix, orc, qt, ql, "sparse")      # with the repair path: exact


def test_dense_topk_normalisation_edges(eng):
    s = torch.tensor([[0.0, 0.0, 0.0, 0.0], [-1.0, -2.0, -0.5, -3.0], [2.0, 4.0, 4.0, 1.0]], dtype=torch.float64,
                     device=DEV)
    ids, sc, mx = eng.dense_topk(s, 3, normalize=True)
    assert ids.cpu().tolist() == [[0, 1, 2], [2, 0, 1], [1, 2, 0]]
    assert mx.cpu().tolist() == [1.0, 1.0, 4.0]  # max <= 0 -> divisor 1.0 (rag/retrieval.py:344)
    assert sc.cpu().tolist() == [[0.0, 0.0, 0.0], [-0.5, -1.0, -2.0], [1.0, 1.0, 0.5]]


def test_cosine_scan_cluster_pairs_forced_on_small_inputs():
    """The main scan runs as clusters of two CTAs (one cta_group::2 MMA over both SMs, or -- ORAG_SCAN_2SM=0 -- two
    1-SM MMAs that multicast the query slabs to each other) once a shard has >= 4 tiles per SM; ORAG_SCAN_CLUSTER=2
    (read once per process, hence the subprocess) forces that path on small inputs: odd tile counts (one CTA of the
    last pair gets an all-zero tile), ragged tails, few queries."""
    import subprocess
    import sys
    code = r'''
import numpy as np, torch, sys
sys.path.insert(0, %r)
from optimized_rag_b200 import engine, synthetic as syn
for n, dim, nq in [(128 * 5 + 77, 256, 70), (128 * 301, 1536, 256), (4500, 64, 3)]:
    corpus = torch.from_numpy(syn.embeddings(syn.SEED_CORPUS, 0, n, dim, 2)).cuda()
    q = torch.from_numpy(syn.query_embeddings(nq, n, dim, dup_per_mille=2)).cuda()
    want = engine.CosineIndex(corpus, mode="exact").topk(q, 10)
    for mode in ("f16", "tf32", "bf16"):
        got = engine.CosineIndex(corpus, mode=mode).topk(q, 10)
        assert torch.equal(want[0], got[0]) and torch.equal(want[1], got[1]), (n, dim, nq, mode)
print("cluster ok")
''' % str(ROOT)
    for two_sm in ("1", "0"):   # tcgen05.mma.cta_group::2 over the pair (default) / multicast pairs of 1-SM MMAs
        env = dict(os.environ, ORAG_SCAN_CLUSTER="2", ORAG_SCAN_2SM=two_sm)
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0 and "cluster ok" in r.stdout, (two_sm, r.stdout + r.stderr)


@pytest.mark.parametrize("head_stream,coschedule", [(False, True), (True, True), (False, False)])
def test_submit_keeps_batches_in_flight_and_equals_search(eng, head_stream, coschedule):
    """ShardedHybrid.submit (three lanes: own streams, workspaces, exchange slots) over a sequence of DIFFERENT batches:
    every ticket's result equals the plain search of its batch, bit for bit, whatever order the lanes finish in.
    head_stream=True drives the cosine search through the split entry point (orag_cosine_topk_phase: preparation on the
    lane's stream, every scan on one shared stream, re-scores and selection back on the lane's stream)."""
    from optimized_rag_b200.bm25_index import Bm25Index
    from optimized_rag_b200.dist import ShardedHybrid
    n, dim, vocab, k = 30000, 256, 3000, 10
    thr = syn.zipf_thresholds(vocab)
    corpus = syn.embeddings(syn.SEED_CORPUS, 0, n, dim, 5)
    doc_off, tok = syn.token_corpus(syn.SEED_TOKENS, 0, n, vocab, 20, 80, thr)
    sh = ShardedHybrid(eng.HybridShard(eng.CosineIndex(_t(corpus), mode="f16"),
                                       Bm25Index(_t(doc_off), _t(tok), vocab, tile_docs=1024)))
    sh.shard.head_stream, sh.shard.coschedule = head_stream, coschedule   # (default: co-scheduled for large shards only)
    batches = []
    for i, nq in enumerate([64, 17, 64, 1, 40, 300, 64]):   # 300: split into two 256-query groups on two lanes
        q = syn.query_embeddings(nq, n, dim, query_seed=syn.SEED_QUERIES + i, dup_per_mille=5)
        qt, ql = syn.keyword_queries(nq, vocab, seed=syn.SEED_KWQUERIES + 10 * i, min_rank=10, thresholds=thr)
        batches.append((_t(q), _t(qt), _t(ql)))
    tickets = [sh.submit(*b, k) for b in batches]          # nothing waits in between
    got = [t.wait() for t in tickets]
    torch.cuda.synchronize()
    for b, g in zip(batches, got):
        want = sh.search(*b, k)
        assert int(g["status"].max().item()) == 0
        for key in ("ids", "rrf_scores", "cos_ids", "cos_scores", "bm25_ids", "bm25_scores", "bm25_max"):
            assert torch.equal(g[key], want[key]), key


def test_hybrid_merge_flags_lists_cut_inside_a_collapsed_tie_group(eng):
    """Shards rank BM25 by raw score; dividing by the global max can fold two neighbouring raw values into one
    normalised double, inside which the order is by id.  orag_hybrid_merge must flag a query when a shard's FULL list
    ends inside the k-th value's group and that group mixes raw values (something may have been cut off), and must
    not flag equal-raw ties or lists that end below the k-th value."""
    from optimized_rag_b200.dist import hybrid_merge, pack_local
    M = 3.0
    a = None
    x = 1.9
    for _ in range(10000):     # a raw value whose lower neighbour normalises to the same double
        x = np.nextafter(x, 0.0)
        if x / M == np.nextafter(x, 0.0) / M:
            a = float(x)
            break
    assert a is not None
    b = float(np.nextafter(a, 0.0))
    lone = a
    while lone / M == np.nextafter(lone, 0.0) / M or lone / M == np.nextafter(lone, 4.0) / M:
        lone = float(np.nextafter(lone, 0.0))   # a raw value with no collapsing neighbour
    fk, kk, k = 2, 3, 2

    def shard(ids, raws, smax):
        ci = np.array([[100 + ids[0], 101 + ids[0]]], dtype=np.int64)
        cs = np.array([[0.5, 0.25]])
        return pack_local(_t(ci), _t(cs), _t(np.array([ids], dtype=np.int64)), _t(np.array([raws])),
                          _t(np.array([smax])), _t(np.zeros(1, dtype=np.int32)))

    other = shard([50, -1, -1], [0.5, 0.0, 0.0], 0.5)                 # second shard: short list, far below
    cases = [
        ([1, 9, 3], [M, a, b], 1, [1, 3]),      # full list ends in the k-th value's group, raws a != b collapse: flag
        ([1, 9, 12], [M, a, a], 1, [1, 9]),     # equal raws, but a's lower neighbour would collapse too: flag
        ([1, 9, 12], [M, lone, lone], 0, [1, 9]),   # equal raws, nothing else can collapse: the lowest ids were kept
        ([1, 9, 3], [M, a, 0.7], 0, [1, 9]),    # the list ends below the k-th value
        ([1, 9, -1], [M, a, 0.0], 0, [1, 9]),   # the list is not full: nothing was cut off
    ]
    for ids, raws, flag, want_ids in cases:
        g = torch.stack([shard(ids, raws, M), other]).contiguous()
        out, status = hybrid_merge(g, fk, kk, 60, k)
        assert int(status.item()) == flag, (ids, raws)
        assert out["bm25_ids"][0].cpu().tolist() == want_ids
        assert out["bm25_scores"][0].cpu().tolist() == [1.0, raws[1] / M]


@pytest.mark.parametrize("mode", ["f16", "tf32", "bf16"])
def test_cosine_dim_3072_with_planted_near_ties(eng, mode):
    """dim = 3072 (text-embedding-3-large): the first-pass margin is a function of the vector length (accumulation
    error grows with dim: api.cu first_pass_eps), and the final order of rows whose cosines differ by 1e-7 .. 1e-6 --
    far below what the 16-bit / tf32 first pass can resolve -- must still be the oracle's, bit for bit."""
    n, dim, nq, k = 12000, 3072, 5, 10
    corpus = syn.embeddings(syn.SEED_CORPUS, 0, n, dim, 0)
    queries = syn.query_embeddings(nq, n, dim)
    rng = np.random.default_rng(11)
    for b in range(nq):
        q = queries[b].astype(np.float64)
        u = rng.standard_normal(dim)
        u -= q * (u @ q) / (q @ q)                       # orthogonal to the query
        u *= np.linalg.norm(q) / np.linalg.norm(u)
        for j in range(24):                              # cos = 1 / sqrt(1 + eps^2): 24 rows within ~2e-5 of each other
            eps = 0.30 + 2.5e-6 * j
            corpus[(b * 997 + j * 31 + 5) % n] = ((q + eps * u) * (0.5 + 0.25 * (j % 3))).astype(np.float32)
        corpus[(b * 997 + 3) % n] = corpus[(b * 997 + 5) % n]      # exact duplicate: tie broken by row id
    index = eng.CosineIndex(_t(corpus), mode=mode)
    ids, sc = index.topk(_t(queries), k)
    want_i, want_s = oracle.cosine_topk(corpus, queries, k)
    assert np.array_equal(ids.cpu().numpy(), want_i)
    assert np.array_equal(_bits(sc.cpu().numpy()), _bits(want_s))
    gaps = np.diff(want_s, axis=1)
    assert (np.abs(gaps) < 1e-5).sum() >= nq * 5         # the planted rows really are the ones being ranked


@pytest.mark.parametrize("n_docs", [8192 + 2000, 3 * 8192 + 700])
def test_bm25_first_pass_on_a_mostly_empty_last_tile(n_docs):
    """A shard whose last 8192-doc first-pass tile holds a fraction of the docs: the doc sub-ranges of a cold, dense
    (query, tile) pair must be cut from the tile's own doc range -- equal slices of the full tile put every posting
    into the first few sub-ranges and overflowed the per-warp accumulator (flagged queries, repaired, but on every
    batch).  No query may be flagged, in the stand-alone and in the background configuration, and the lists must equal
    the dense path's."""
    from optimized_rag_b200 import engine
    from optimized_rag_b200.bm25_index import Bm25Index, Bm25Plan
    vocab = 50000
    thr = syn.zipf_thresholds(vocab)
    off, tok = engine.gen_token_corpus(n_docs, 0, syn.SEED_TOKENS, thr, vocab, 100, 300, device=torch.device(DEV))
    ix = Bm25Index.from_plan(Bm25Plan(off, tok, vocab, fp_tile_docs=8192))
    assert ix.struct.fp_tile_docs == 8192
    qt, ql = syn.keyword_queries(256, vocab, thresholds=thr)
    qt, ql = torch.from_numpy(qt).to(DEV), torch.from_numpy(ql).to(DEV)
    for k in (10, 16):
        want = ix.topk(qt, ql, k, normalize=False, force="dense", check_overflow=False)
        for background in (False, True):
            st = []
            got = ix.topk(qt, ql, k, normalize=False, check_overflow=False, status_out=st, background=background)
            assert int((st[0] != 0).sum()) == 0, (k, background, sorted(set(st[0].tolist())))
            for g, w in zip(got, want):
                assert torch.equal(g, w), (k, background)


def test_timeline_and_candidate_count_hooks(eng):
    """Diagnostic entry points (bench.py --timeline, DESIGN.md candidate statistics): after orag_timeline_enable(1) a
    search leaves one (tag, begin, end) record per tagged launch, begin <= end, the main scan among them; the candidate
    counts of the last search are at least k per query and the fp32 survivors at most the candidates."""
    import ctypes
    from optimized_rag_b200 import _ffi
    L = _ffi.lib()
    n, dim, k = 40000, 256, 10
    corpus = syn.embeddings(syn.SEED_CORPUS, 0, n, dim, 0)
    ix = eng.CosineIndex(_t(corpus), mode="f16")
    q = _t(syn.query_embeddings(64, n, dim))
    ix.topk(q, k)
    torch.cuda.synchronize()
    assert L.orag_timeline_enable(1) == 0
    ids, _ = ix.topk(q, k)
    cap = 64
    tags, t0, t1 = (ctypes.c_int * cap)(), (ctypes.c_float * cap)(), (ctypes.c_float * cap)()
    got = int(L.orag_timeline_read(tags, t0, t1, cap))
    assert L.orag_timeline_enable(0) == 0
    assert 6 <= got <= cap
    seen = {int(tags[i]) for i in range(got)}
    assert {2, 3, 4, 5, 6, 7} <= seen                      # seed scan, seed finalize, main scan, prefilter, rescore, select
    assert all(0.0 <= t0[i] <= t1[i] for i in range(got))
    cand, surv = ix.last_counts(64, k)
    assert (cand >= k).all() and (surv >= k).all() and (surv <= cand).all() and int(cand.max()) <= 4096
    # disabled again: nothing is recorded
    ix.topk(q, k)
    assert int(L.orag_timeline_read(tags, t0, t1, cap)) == got
