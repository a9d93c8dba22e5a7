"""CPU: the streamed oracle (oracle.c "Streamed oracle" section: inputs regenerated from the seeds block by block,
queries evaluated side by side in SIMD lanes) equals the in-memory oracle bit for bit, and its C generators equal the
numpy generators of optimized_rag_b200/synthetic.py.  These are the functions the BASELINE-size parity gates use
(tests/test_zz_gpu_parity_scale.py, bench.py `verified_against_oracle`)."""
import numpy as np

import oracle
from optimized_rag_b200 import synthetic as syn


def test_c_generators_equal_numpy_generators():
    for dup in (0, 7):
        a = syn.embeddings(syn.SEED_CORPUS, 12345, 300, 96, dup)
        b = oracle.gen_embeddings(syn.SEED_CORPUS, 12345, 300, 96, dup)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    for vocab, lmin, lmax, start in ((50000, 100, 300, 0), (2000, 1, 40, 777), (50, 3, 9, 5)):
        thr = syn.zipf_thresholds(vocab)
        off_a, tok_a = syn.token_corpus(syn.SEED_TOKENS, start, 400, vocab, lmin, lmax, thr)
        off_b, tok_b = oracle.gen_token_corpus(syn.SEED_TOKENS, start, 400, vocab, lmin, lmax, thr)
        assert np.array_equal(off_a, off_b) and np.array_equal(tok_a, tok_b)


def test_streamed_cosine_topk_equals_in_memory_oracle():
    n, dim, nq, k = 3000, 192, 37, 10   # 37 queries: two full SIMD blocks + a ragged one
    for dup in (0, 20):                 # planted duplicate rows: exact ties -> lower id first
        corpus = syn.embeddings(syn.SEED_CORPUS, 0, n, dim, dup)
        queries = syn.query_embeddings(nq, n, dim, dup_per_mille=dup)
        queries[5] = 0.0                # zero query: every cosine is 0.0 -> ids 0..k-1
        want_i, want_s = oracle.cosine_topk(corpus, queries, k)
        got_i, got_s = oracle.cosine_topk_stream(queries, k, n, seed=syn.SEED_CORPUS, dim=dim, dup_per_mille=dup)
        assert np.array_equal(got_i, want_i)
        assert np.array_equal(got_s.view(np.uint64), want_s.view(np.uint64))
        # the same through an in-memory block with a row offset (the sharded / sub-range use)
        g2_i, g2_s = oracle.cosine_topk_stream(queries, k, 1000, row_start=500, dim=dim, corpus=corpus[500:1500])
        w2_i, w2_s = oracle.cosine_topk(corpus[500:1500], queries, k, id_base=500)
        assert np.array_equal(g2_i, w2_i) and np.array_equal(g2_s.view(np.uint64), w2_s.view(np.uint64))
    # plain-sum mode (CPython < 3.12) is carried through as well
    a = oracle.cosine_topk_stream(queries[:3], k, n, seed=syn.SEED_CORPUS, dim=dim, dup_per_mille=20, neumaier=False)
    b = oracle.cosine_topk(corpus, queries[:3], k, neumaier=False)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint64), b[1].view(np.uint64))


def test_streamed_cosine_fewer_rows_than_k():
    q = syn.query_embeddings(2, 4, 64)
    ids, sc = oracle.cosine_topk_stream(q, 10, 4, seed=syn.SEED_CORPUS, dim=64)
    want_i, want_s = oracle.cosine_topk(syn.embeddings(syn.SEED_CORPUS, 0, 4, 64), q, 10)
    assert np.array_equal(ids, want_i) and np.array_equal(sc.view(np.uint64), want_s.view(np.uint64))


def test_streamed_bm25_equals_in_memory_oracle():
    for n, vocab, lmin, lmax, min_rank in ((6000, 3000, 20, 80, 10), (9000, 50000, 100, 300, 100)):
        thr = syn.zipf_thresholds(vocab)
        doc_off, tok = syn.token_corpus(syn.SEED_TOKENS, 0, n, vocab, lmin, lmax, thr)
        ref = oracle.BM25Index(doc_off, tok, vocab)
        st = oracle.StreamedBM25(syn.SEED_TOKENS, n, vocab, lmin, lmax, thr)
        assert np.array_equal(st.df, ref.df) and st.total_len == int(doc_off[-1])
        assert st.order.tolist() == ref.first_seen.tolist()      # dict insertion order of the vocabulary
        assert st.avgdl == ref.avgdl and st.average_idf == ref.average_idf and st.eps == ref.eps
        assert np.array_equal(st.idf.view(np.uint64), ref.idf.view(np.uint64))
        qt, ql = syn.keyword_queries(48, vocab, min_rank=min_rank, thresholds=thr)   # incl. OOV / duplicate / common terms
        raw = st.scores_raw(qt, ql)
        ids, sc, mx = st.topk(qt, ql, 10)
        for b in range(qt.shape[0]):
            want = ref.scores_raw(qt[b, :ql[b]])
            assert np.array_equal(raw[b].view(np.uint64), want.view(np.uint64)), b
            wi, ws, wm = ref.topk(qt[b, :ql[b]], 10)
            assert ids[b].tolist() == wi.tolist() and sc[b].tolist() == ws.tolist() and mx[b] == wm


def test_scale_check_reference_lists_and_compare():
    """The wiring the BASELINE-size gates use (oracle/scale_check.py), at a size the in-memory oracle also handles."""
    from oracle import scale_check
    n, dim, vocab, nq, k = 2500, 128, 3000, 20, 10
    thr = syn.zipf_thresholds(vocab)
    q = syn.query_embeddings(nq, n, dim)
    qt, ql = syn.keyword_queries(nq, vocab, min_rank=10, thresholds=thr)
    st = oracle.StreamedBM25(syn.SEED_TOKENS, n, vocab, 20, 80, thr)
    want, secs = scale_check.reference_lists(n, dim, q, qt, ql, k, seed_corpus=syn.SEED_CORPUS, bm25=st)
    assert set(secs) == {"cosine", "bm25", "rrf"}
    doc_off, tok = syn.token_corpus(syn.SEED_TOKENS, 0, n, vocab, 20, 80, thr)
    full = oracle.hybrid_topk(syn.embeddings(syn.SEED_CORPUS, 0, n, dim), q, oracle.BM25Index(doc_off, tok, vocab),
                              [qt[b, :ql[b]] for b in range(nq)], k=k)
    got = {key: np.stack([np.asarray(full[b][key]) for b in range(nq)]) for key in
           ("cos_ids", "cos_scores", "bm25_ids", "bm25_scores", "ids", "rrf_scores")}
    got["bm25_max"] = np.array([full[b]["bm25_max"] for b in range(nq)])
    assert scale_check.compare(got, want) == []
    # a subset of a larger batch, one flipped score bit, one swapped id
    rows = np.array([1, 7, 19])
    sub = {key: val[rows] for key, val in want.items()}
    assert scale_check.compare(got, sub, rows) == []
    got["cos_scores"] = got["cos_scores"].copy()
    got["cos_scores"][7, 3] = np.nextafter(got["cos_scores"][7, 3], 2.0)
    got["ids"] = got["ids"].copy()
    got["ids"][19, 0] += 1
    bad = scale_check.compare(got, sub, rows)
    assert len(bad) == 2 and bad[0].startswith("cos_scores: 1 of") and bad[1].startswith("ids: 1 of")


def test_blocked_pairwise_equals_scalar_pairwise():
    m, dim = 300, 160
    emb = syn.embeddings(syn.SEED_CORPUS, 0, m, dim, 0)
    emb[10] = emb[3] + np.float32(0.3) * emb[10]      # planted similar claims
    emb[200] = emb[77] * np.float32(2.0)              # cosine exactly 1 up to rounding
    emb[5] = 0.0                                      # zero vector -> cosine 0.0
    doc = (np.arange(m) // 4).astype(np.int32)
    for thr in (0.85, 0.05, -1.0):
        a = oracle.pairwise_candidates(emb, doc, thr)
        b = oracle.pairwise_candidates_parallel(emb, doc, thr)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
        assert np.array_equal(a[2].view(np.uint64), b[2].view(np.uint64))
    assert len(oracle.pairwise_candidates_parallel(emb, doc, 0.85)[0]) >= 2


def test_pick_queries_includes_special_keyword_queries():
    from oracle import scale_check
    thr = syn.zipf_thresholds(50000)
    qt, ql = syn.keyword_queries(1024, 50000, thresholds=thr)
    rows = scale_check.pick_queries(qt, ql, 1024, 32)
    assert len(rows) == 32 and len(set(rows.tolist())) == 32 and rows.min() >= 0 and rows.max() < 1024
    assert any(((qt[b, :ql[b]] >= 0) & (qt[b, :ql[b]] < 10)).any() for b in rows) and any((qt[b, :ql[b]] == -1).any() for b in rows)
    assert len(scale_check.pick_queries(qt, ql, 4, 32)) == 4
