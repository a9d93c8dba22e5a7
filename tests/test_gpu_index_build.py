"""GPU (-m gpu): the index builder (csrc/bm25_build.cu: orag_bm25_index_plan / orag_bm25_index_fill) against the oracle.

The exact view must decode to the oracle's postings; the first-pass view must hold, for exactly those postings,
fp16(r) with r = tf*(k1+1)/(tf + t4[dl]), every run 16-byte aligned and padded to four postings with zero-impact
copies of its last doc; statistics (df, first-seen order, doc lengths) must be the oracle's; the on-disk format must
round-trip."""
import numpy as np
import pytest
import torch

import oracle
from optimized_rag_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _build(doc_off, tok, vocab, **kw):
    from optimized_rag_b200.bm25_index import Bm25Index
    return Bm25Index(torch.from_numpy(doc_off).to(DEV), torch.from_numpy(tok).to(DEV), vocab, **kw)


def _decode(post, base, off, tile_docs, n_tiles, vocab, padded):
    """{(term, doc): low half-word}; checks doc-sorted runs and, for the padded layout, alignment + filler postings."""
    post = post.cpu().numpy().view(np.uint32)
    base, off = base.cpu().numpy(), off.cpu().numpy()
    out = {}
    for tl in range(n_tiles):
        if padded:
            assert base[tl] % 4 == 0 and (off[tl] % 4 == 0).all()
        for t in range(vocab):
            p = post[base[tl] + off[tl, t]:base[tl] + off[tl, t + 1]]
            if padded and len(p):
                real = len(p)
                while real > 1 and (p[real - 1] & 0xFFFF) == 0:
                    real -= 1
                assert len(p) - real < 4 and (p[real:] == (p[real - 1] & 0xFFFF0000)).all()   # filler = last doc, impact 0
                p = p[:real]
            d = tl * tile_docs + (p >> 16).astype(np.int64)
            assert (np.diff(d) > 0).all()  # doc-sorted runs: the kernels binary-search and merge them
            for di, lo in zip(d.tolist(), (p & 0xFFFF).tolist()):
                out[(t, di)] = lo
    return out


@pytest.mark.parametrize("n,vocab,lmin,lmax,tile,fp_tile", [(300, 200, 3, 40, 64, None), (1000, 5000, 20, 60, 256, 128),
                                                            (5, 8, 1, 6, 32, 32), (700, 300, 3, 40, 64, 256),
                                                            (1500, 2000, 20, 60, 128, None), (90, 40, 1, 9, 32, 32)])
def test_index_builder_layout_matches_oracle(n, vocab, lmin, lmax, tile, fp_tile):
    thr = syn.zipf_thresholds(vocab)
    doc_off, tok = syn.token_corpus(syn.SEED_TOKENS, 0, n, vocab, lmin, lmax, thr)
    ix = _build(doc_off, tok, vocab, tile_docs=tile, fp_tile_docs=fp_tile)
    orc = oracle.BM25Index(doc_off, tok, vocab)
    assert ix.avgdl == orc.avgdl and ix.average_idf == orc.average_idf and ix.eps == orc.eps
    assert np.array_equal(ix.idf.cpu().numpy().view(np.uint64), orc.idf.view(np.uint64))
    assert np.array_equal(ix.dl.cpu().numpy(), orc.dl) and ix.max_dl == int(orc.dl.max())
    assert np.array_equal(ix.stats.df, orc.df) and np.array_equal(ix.df_local, orc.df)
    seen = np.nonzero(ix.stats.df > 0)[0]
    assert seen[np.argsort(ix.stats.first_seen[seen], kind="stable")].tolist() == orc.first_seen.tolist()
    assert ix.has_negative_idf == bool((orc.idf < 0).any())
    off, pdoc, ptf = orc.postings()
    assert ix.n_postings == len(pdoc)
    exact = _decode(ix.postings, ix.tile_base, ix.tile_term_off, ix.tile_docs, ix.n_tiles, vocab, padded=False)
    want = {(t, int(d)): int(f) for t in range(vocab) for d, f in zip(pdoc[off[t]:off[t + 1]], ptf[off[t]:off[t + 1]])}
    assert exact == want
    if fp_tile is None:
        assert ix.fp_tile_docs == min(8192, max(32, 1 << ((n // 16 - 1).bit_length())))
    if ix.has_negative_idf:
        assert ix.postings_r16 is None
        return
    assert ix.postings_r16 is not None and ix.postings_r16.data_ptr() % 16 == 0 and ix.struct.reserved == 1
    first = _decode(ix.postings_r16, ix.fp_tile_base, ix.fp_tile_term_off, ix.fp_tile_docs, ix.fp_n_tiles, vocab,
                    padded=True)
    assert first.keys() == exact.keys()
    t4 = ix.t4_table.cpu().numpy()
    tmax = np.zeros(vocab, dtype=np.float32)
    for (t, d), tf in exact.items():
        r = tf * 2.5 / (tf + t4[orc.dl[d]])
        r16 = np.float16(np.float32(r))
        assert first[(t, d)] == int(r16.view(np.uint16)), (t, d)
        assert abs(float(r16) - r) <= r * 2.0 ** -11 and float(r16) >= 6.2e-5
        tmax[t] = max(tmax[t], np.float32(r16))
    assert np.array_equal(ix.term_max_r.cpu().numpy(), tmax)
    # threshold warm start: the K-th largest fp16 impact of every term (0 when the term has fewer postings)
    per_term = [[] for _ in range(vocab)]
    for (t, d), bits in first.items():
        per_term[t].append(bits)
    kth = ix.term_kth_r.cpu().numpy()
    for level, K in enumerate((10, 16, 32, 64, 128)):
        want_k = np.zeros(vocab, dtype=np.float32)
        for t in range(vocab):
            if len(per_term[t]) >= K:
                want_k[t] = np.float32(np.uint16(sorted(per_term[t], reverse=True)[K - 1]).view(np.float16))
        assert np.array_equal(kth[level], want_k), K


def test_long_and_repetitive_documents():
    """Documents far longer than one de-duplication pass (768 tokens), one of them a single term repeated 40 000 times,
    an empty document in between; doc length limit and token range are enforced."""
    rng = np.random.default_rng(3)
    vocab = 5000
    docs = [rng.integers(0, vocab, 3000), np.full(40000, 7), np.zeros(0, np.int64), rng.integers(0, 50, 2500),
            rng.integers(0, vocab, 10), np.arange(vocab)]
    doc_off = np.cumsum([0] + [len(d) for d in docs]).astype(np.int64)
    tok = np.concatenate(docs).astype(np.int32)
    ix = _build(doc_off, tok, vocab, tile_docs=32)
    orc = oracle.BM25Index(doc_off, tok, vocab)
    off, pdoc, ptf = orc.postings()
    exact = _decode(ix.postings, ix.tile_base, ix.tile_term_off, 32, ix.n_tiles, vocab, padded=False)
    want = {(t, int(d)): int(f) for t in range(vocab) for d, f in zip(pdoc[off[t]:off[t + 1]], ptf[off[t]:off[t + 1]])}
    assert exact == want and exact[(7, 1)] >= 40000 and ix.max_dl == 40000
    with pytest.raises(ValueError, match="65535"):
        _build(np.array([0, 70000], dtype=np.int64), np.zeros(70000, dtype=np.int32), 4, tile_docs=32)
    with pytest.raises(ValueError, match="vocab"):
        _build(np.array([0, 3], dtype=np.int64), np.array([0, 9, 1], dtype=np.int32), 4, tile_docs=32)


def test_first_pass_view_is_dropped_when_idf_goes_negative():
    """When the final idf table keeps a negative entry (common terms whose replacement eps*average_idf is itself
    negative, rank_bm25's behaviour), pruning by upper bounds is unsound: the index must not build the first-pass
    view, and the exact tile kernel serves every query."""
    doc_off = np.arange(0, 4 * 40 + 1, 4, dtype=np.int64)
    tok = np.tile(np.array([0, 1, 2, 3], dtype=np.int32), 40)
    tok[3::8] = 5
    ix = _build(doc_off, tok, 8, tile_docs=32)
    assert ix.has_negative_idf and ix.postings_r16 is None and ix.struct.d_postings_r16 is None and ix.struct.reserved == 0


def test_plan_statistics_of_shards_merge_into_the_global_ones():
    from optimized_rag_b200.bm25_index import Bm25Index, Bm25Plan, idf_table
    vocab, n = 400, 900
    thr = syn.zipf_thresholds(vocab)
    doc_off, tok = syn.token_corpus(syn.SEED_TOKENS, 0, n, vocab, 5, 50, thr)
    orc = oracle.BM25Index(doc_off, tok, vocab)
    parts, plans = None, []
    for s, e in [(0, 300), (300, 650), (650, 900)]:
        plan = Bm25Plan(torch.from_numpy(doc_off[s:e + 1] - doc_off[s]).to(DEV),
                        torch.from_numpy(tok[doc_off[s]:doc_off[e]]).to(DEV), vocab, tile_docs=64,
                        token_pos_base=int(doc_off[s]))
        plans.append((plan, s))
        parts = plan.local_stats if parts is None else parts.merged(plan.local_stats)
    idf, avg, eps = idf_table(parts)
    assert np.array_equal(idf.view(np.uint64), orc.idf.view(np.uint64)) and eps == orc.eps and parts.avgdl == orc.avgdl
    # every shard's index, filled with the global statistics, scores like the oracle restricted to its docs
    qt, ql = syn.keyword_queries(6, vocab, min_rank=3, thresholds=thr)
    for plan, s in plans:
        ix = Bm25Index.from_plan(plan, parts, doc_id_base=s)
        dense = ix.dense_scores(torch.from_numpy(qt).to(DEV), torch.from_numpy(ql).to(DEV)).cpu().numpy()
        for b in range(6):
            want = orc.scores_raw(qt[b, :ql[b]])[s:s + ix.n_docs]
            assert np.array_equal(dense[b].view(np.uint64), want.view(np.uint64))


@pytest.mark.parametrize("n,vocab,tile,negative", [(700, 300, 64, False), (40, 8, 32, True), (0, 5, 32, False)])
def test_bm25_index_save_load_round_trip(tmp_path, n, vocab, tile, negative):
    """On-disk format of the keyword index: every array and scalar survives save -> load bit for
    bit, and the struct handed to the kernels is field-for-field the one the builder makes (pointers aside)."""
    from optimized_rag_b200 import _ffi
    from optimized_rag_b200.bm25_index import Bm25Index
    if negative:
        doc_off = np.arange(0, 4 * n + 1, 4, dtype=np.int64)
        tok = np.tile(np.array([0, 1, 2, 3], dtype=np.int32), n)
        tok[3::8] = 5
    else:
        thr = syn.zipf_thresholds(vocab)
        doc_off, tok = syn.token_corpus(syn.SEED_TOKENS, 0, n, vocab, 3, 40, thr)
    ix = _build(doc_off, tok, vocab, tile_docs=tile, doc_id_base=1234)
    assert ix.has_negative_idf == negative and (ix.postings_r16 is None) == (negative or n == 0)
    ix.save(tmp_path / "kw")
    back = Bm25Index.load(tmp_path / "kw", device="cpu")
    for name in Bm25Index._ARRAYS:
        a, b = getattr(ix, name), getattr(back, name)
        assert (a is None) == (b is None), name
        if a is not None:
            assert a.dtype == b.dtype and a.shape == b.shape and torch.equal(a.cpu(), b), name
    for key in ("n_docs", "vocab", "tile_docs", "n_tiles", "fp_tile_docs", "fp_n_tiles", "doc_id_base", "n_postings",
                "n_postings_fp", "max_dl", "has_negative_idf", "avgdl", "average_idf", "eps"):
        assert getattr(ix, key) == getattr(back, key), key
    assert back.stats.n_docs == ix.stats.n_docs and back.stats.total_len == ix.stats.total_len
    assert np.array_equal(back.stats.df, ix.stats.df) and np.array_equal(back.stats.first_seen, ix.stats.first_seen)
    assert np.array_equal(back.df_local, ix.df_local)
    for field, ctype in _ffi.Bm25IndexStruct._fields_:
        va, vb = getattr(ix.struct, field), getattr(back.struct, field)
        if field == "d_term_kth_r":
            assert vb is None   # derived on the GPU from the first-pass view (orag_bm25_term_kth): not for a host copy
        elif field.startswith("d_"):
            assert (va is None) == (vb is None), field   # same arrays present, each pointing at its own copy
        else:
            assert va == vb, field
    # a reloaded index answers like the one that was saved
    if n:
        thr = syn.zipf_thresholds(vocab)
        qt, ql = syn.keyword_queries(5, vocab, min_rank=min(3, vocab - 1), thresholds=thr)
        dev_ix = Bm25Index.load(tmp_path / "kw", device=DEV)
        a = ix.topk(torch.from_numpy(qt).to(DEV), torch.from_numpy(ql).to(DEV), 5, force="sparse")
        b = dev_ix.topk(torch.from_numpy(qt).to(DEV), torch.from_numpy(ql).to(DEV), 5, force="sparse")
        assert all(torch.equal(x, y) for x, y in zip(a, b))
        assert (dev_ix.term_kth_r is None) == (ix.term_kth_r is None)
        if ix.term_kth_r is not None:
            assert torch.equal(dev_ix.term_kth_r, ix.term_kth_r)   # recomputed on load, not stored
    # a damaged file is refused
    data = (tmp_path / "kw.bin").read_bytes()
    if len(data) > 16:
        (tmp_path / "kw.bin").write_bytes(data[:-8])
        with pytest.raises(ValueError):
            Bm25Index.load(tmp_path / "kw", device="cpu")


def test_term_kth_scans_a_frequent_term_only_until_it_has_enough_postings():
    """orag_bm25_term_kth: a term with >= 16384 postings inside the first block of 256 first-pass tiles is bounded from
    that block alone (any subset of the postings gives a valid K-th largest impact); a term that only occurs beyond
    that block is still scanned in full.  The searches stay exact with the weaker bound."""
    rng = np.random.default_rng(11)
    n, vocab, fp_tile = 40000, 60, 128
    cut = 256 * fp_tile                       # docs of the first block
    docs = []
    for d in range(n):
        body = [0] * int(rng.integers(1, 5)) + rng.integers(2, vocab, int(rng.integers(3, 30))).tolist()
        if d >= cut + 100 and d % 3 == 0:
            body += [1] * int(rng.integers(1, 4))
        docs.append(body)
    doc_off = np.zeros(n + 1, dtype=np.int64)
    doc_off[1:] = np.cumsum([len(x) for x in docs])
    tok = np.concatenate([np.asarray(x, dtype=np.int32) for x in docs])
    ix = _build(doc_off, tok, vocab, tile_docs=128, fp_tile_docs=fp_tile)
    assert ix.fp_n_tiles > 256 and ix.term_kth_r is not None
    first = _decode(ix.postings_r16, ix.fp_tile_base, ix.fp_tile_term_off, fp_tile, ix.fp_n_tiles, vocab, padded=True)
    kth = ix.term_kth_r.cpu().numpy()
    for term, limit in ((0, cut), (1, n)):
        bits = sorted((b for (t, d), b in first.items() if t == term and d < limit), reverse=True)
        assert len(bits) >= 128
        for level, K in enumerate((10, 16, 32, 64, 128)):
            assert kth[level, term] == np.float32(np.uint16(bits[K - 1]).view(np.float16)), (term, K)
    qt = torch.tensor([[0, 1, 5], [1, 7, -2], [0, -2, -2]], dtype=torch.int32, device=DEV)
    ql = torch.tensor([3, 2, 1], dtype=torch.int32, device=DEV)
    got = ix.topk(qt, ql, 10, normalize=False, force="sparse")
    dense = ix.topk(qt, ql, 10, normalize=False, force="dense")
    assert all(torch.equal(g, w) for g, w in zip(got, dense))
