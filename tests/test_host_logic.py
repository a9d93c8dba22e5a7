"""CPU tests (-m "not gpu"): host-side logic, the C-ABI library's exports, the index builder's layout.
No compute call into the CUDA library is made here (there is no GPU in the build container)."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest
import torch

import oracle
from optimized_rag_b200 import synthetic as syn

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def built_lib():
    from optimized_rag_b200 import build
    return build.build()


def test_library_exports_every_declared_symbol(built_lib):
    from optimized_rag_b200 import _ffi
    header = (ROOT / "include" / "orag.h").read_text()
    declared = set(re.findall(r"\b(orag_[a-z0-9_]+)\s*\(", header))
    declared -= {"orag_bm25_index"}
    assert declared == set(_ffi.SYMBOLS), declared ^ set(_ffi.SYMBOLS)
    L = ctypes.CDLL(str(built_lib))
    for name in declared:
        assert hasattr(L, name), name
    # argument-free calls are safe without a GPU
    assert _ffi.lib().orag_version() == 1
    assert _ffi.lib().orag_last_error() is not None


def test_library_is_sm100a_with_tcgen05_and_tma(built_lib):
    import subprocess
    sass = subprocess.run(["cuobjdump", "-sass", str(built_lib)], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert mnemonic in sass, mnemonic


def test_product_never_imports_oracle():
    for p in (ROOT / "optimized_rag_b200").rglob("*.py"):
        src = p.read_text()
        assert "import oracle" not in src and "from oracle" not in src, p


def test_no_cpu_path_for_cpu_tensors(built_lib):
    from optimized_rag_b200 import _ffi, engine
    with pytest.raises(_ffi.OragError):
        engine.CosineIndex(torch.zeros((4, 32), dtype=torch.float32))
    with pytest.raises(_ffi.OragError):
        engine.rrf_fuse(torch.zeros((1, 2, 3), dtype=torch.int64))


def test_synthetic_generators_are_shard_consistent():
    a = syn.embeddings(syn.SEED_CORPUS, 0, 300, 128, 20)
    b = np.concatenate([syn.embeddings(syn.SEED_CORPUS, s, 100, 128, 20) for s in (0, 100, 200)])
    assert np.array_equal(a, b)
    # exact representability: every value is a multiple of 2**-28 below 2**-5
    assert np.all(np.abs(a) < 2.0 ** -5) and np.array_equal(a * 2.0 ** 28, np.round(a * 2.0 ** 28))
    src = syn.source_rows(syn.SEED_CORPUS, np.arange(300), 20)
    dup = np.nonzero(src != np.arange(300))[0]
    assert len(dup) > 0 and all(np.array_equal(a[r], syn.embeddings(syn.SEED_CORPUS, int(src[r]), 1, 128, 0)[0])
                                for r in dup)
    thr = syn.zipf_thresholds(1000)
    o1, t1 = syn.token_corpus(syn.SEED_TOKENS, 0, 50, 1000, 5, 20, thr)
    o2, t2 = syn.token_corpus(syn.SEED_TOKENS, 30, 20, 1000, 5, 20, thr)
    assert np.array_equal(t1[o1[30]:], t2)
    q, ql = syn.keyword_queries(500, 1000, thresholds=thr)
    assert ql.min() >= 3 and ql.max() <= 8 and ((q == -1).sum(axis=1) > 0).sum() > 0


@pytest.mark.parametrize("n,vocab,lmin,lmax,tile", [(300, 200, 3, 40, 64), (1000, 5000, 20, 60, 256), (5, 8, 1, 6, 32)])
def test_index_builder_layout_matches_oracle(built_lib, n, vocab, lmin, lmax, tile):
    """The torch-built tiled index decodes to exactly the oracle's postings / idf / t4."""
    from optimized_rag_b200.bm25_index import Bm25Index
    thr = syn.zipf_thresholds(vocab)
    doc_off, tok = syn.token_corpus(syn.SEED_TOKENS, 0, n, vocab, lmin, lmax, thr)
    ix = Bm25Index(torch.from_numpy(doc_off), torch.from_numpy(tok), vocab, tile_docs=tile)
    orc = oracle.BM25Index(doc_off, tok, vocab)
    assert ix.avgdl == orc.avgdl and ix.average_idf == orc.average_idf and ix.eps == orc.eps
    assert np.array_equal(ix.idf.numpy().view(np.uint64), orc.idf.view(np.uint64))
    assert ix.has_negative_idf == bool((orc.idf < 0).any())
    # t4 = k1 * (1 - b + b * dl / avgdl) in the oracle's operation order
    dl = orc.dl.astype(np.float64)
    assert np.array_equal(ix.doc_t4.numpy(), 1.5 * (0.25 + (0.75 * dl) / orc.avgdl))
    off, pdoc, ptf = orc.postings()
    post = ix.postings.numpy().view(np.uint32)
    tbase = ix.tile_base.numpy()
    toff = ix.tile_term_off.numpy()
    assert ix.n_postings == len(pdoc)
    for t in range(vocab):
        docs, tfs = [], []
        for tl in range(ix.n_tiles):
            s, e = tbase[tl] + toff[tl, t], tbase[tl] + toff[tl, t + 1]
            p = post[s:e]
            docs.extend((tl * tile + (p >> 16)).tolist())
            tfs.extend((p & 0xFFFF).tolist())
        assert docs == pdoc[off[t]:off[t + 1]].tolist(), t
        assert tfs == ptf[off[t]:off[t + 1]].tolist(), t


def test_sharded_stats_merge_equals_global(built_lib):
    from optimized_rag_b200.bm25_index import idf_table, local_stats
    vocab, n = 400, 900
    thr = syn.zipf_thresholds(vocab)
    doc_off, tok = syn.token_corpus(syn.SEED_TOKENS, 0, n, vocab, 5, 50, thr)
    whole = local_stats(torch.from_numpy(doc_off), torch.from_numpy(tok), vocab)
    parts = None
    for s, e in [(0, 300), (300, 650), (650, 900)]:
        off = doc_off[s:e + 1] - doc_off[s]
        st = local_stats(torch.from_numpy(off), torch.from_numpy(tok[doc_off[s]:doc_off[e]]), vocab,
                         token_pos_base=int(doc_off[s]))
        parts = st if parts is None else parts.merged(st)
    assert parts.n_docs == whole.n_docs and parts.total_len == whole.total_len
    assert np.array_equal(parts.df, whole.df) and np.array_equal(parts.first_seen, whole.first_seen)
    a, b = idf_table(parts), idf_table(whole)
    assert np.array_equal(a[0], b[0]) and a[1:] == b[1:]
    orc = oracle.BM25Index(doc_off, tok, vocab)
    assert np.array_equal(a[0].view(np.uint64), orc.idf.view(np.uint64))
    order = np.argsort(whole.first_seen[whole.df > 0], kind="stable")
    assert np.array_equal(np.nonzero(whole.df > 0)[0][order], orc.first_seen)


@pytest.mark.parametrize("n,vocab,lmin,lmax,tile,fp_tile", [(700, 300, 3, 40, 64, 256), (90, 40, 1, 9, 32, 32),
                                                            (1500, 2000, 20, 60, 128, None)])
def test_first_pass_view_is_a_rounded_copy_of_the_exact_view(built_lib, n, vocab, lmin, lmax, tile, fp_tile):
    """The MaxScore first-pass view (csrc/bm25_ms.cu) must hold, for exactly the postings of the exact view,
    fp16(r) with r = tf*(k1+1)/(tf + t4[dl]) in float64, and term_max_r must bound every r16 of its term: the
    superset guarantee of the first pass rests on |r16 - r| <= 2^-11 r and on those upper bounds."""
    from optimized_rag_b200.bm25_index import Bm25Index
    thr = syn.zipf_thresholds(vocab)
    doc_off, tok = syn.token_corpus(syn.SEED_TOKENS, 0, n, vocab, lmin, lmax, thr)
    ix = Bm25Index(torch.from_numpy(doc_off), torch.from_numpy(tok), vocab, tile_docs=tile, fp_tile_docs=fp_tile)
    assert ix.postings_r16 is not None and not ix.has_negative_idf
    T, F = ix.tile_docs, ix.fp_tile_docs
    if fp_tile is None:
        assert F == min(4096, max(32, 1 << ((n // 16 - 1).bit_length())))

    def decode(post, base, off, tile_docs, n_tiles):
        post = post.numpy().view(np.uint32)
        base, off = base.numpy(), off.numpy()
        out = {}
        for tl in range(n_tiles):
            for t in range(vocab):
                p = post[base[tl] + off[tl, t]:base[tl] + off[tl, t + 1]]
                d = tl * tile_docs + (p >> 16).astype(np.int64)
                assert (np.diff(d) > 0).all()  # doc-sorted runs: the kernels binary-search and merge them
                for di, lo in zip(d.tolist(), (p & 0xFFFF).tolist()):
                    out[(t, di)] = lo
        return out

    exact = decode(ix.postings, ix.tile_base, ix.tile_term_off, T, ix.n_tiles)
    first = decode(ix.postings_r16, ix.fp_tile_base, ix.fp_tile_term_off, F, ix.fp_n_tiles)
    assert exact.keys() == first.keys() and len(exact) == ix.n_postings
    t4 = ix.t4_table.numpy()
    dl = ix.dl.numpy()
    tmax = np.zeros(vocab, dtype=np.float32)
    for (t, d), tf in exact.items():
        r = tf * 2.5 / (tf + t4[dl[d]])
        r16 = np.float16(np.float32(r))
        assert first[(t, d)] == int(r16.view(np.uint16)), (t, d)
        assert abs(float(r16) - r) <= r * 2.0 ** -11 and float(r16) >= 6.2e-5
        tmax[t] = max(tmax[t], np.float32(r16))
    assert np.array_equal(ix.term_max_r.numpy(), tmax)


def test_first_pass_view_is_dropped_when_idf_goes_negative(built_lib):
    """When the final idf table keeps a negative entry (common terms whose replacement eps*average_idf is itself
    negative, rank_bm25's behaviour), pruning by upper bounds is unsound: the index must not build the first-pass
    view, and the exact tile kernel serves every query."""
    from optimized_rag_b200.bm25_index import Bm25Index
    doc_off = np.arange(0, 4 * 40 + 1, 4, dtype=np.int64)
    tok = np.tile(np.array([0, 1, 2, 3], dtype=np.int32), 40)
    tok[3::8] = 5
    ix = Bm25Index(torch.from_numpy(doc_off), torch.from_numpy(tok), 8, tile_docs=32)
    assert ix.has_negative_idf and ix.postings_r16 is None and ix.struct.d_postings_r16 is None


def test_semantic_dedup_host_logic_with_a_stand_in_cosine_matrix(built_lib, monkeypatch):
    """Every line of Deduplicator.semantic_dedup except the kernel launch, on the CPU: the cosine index is replaced
    by a stand-in that serves the oracle's float64 cosines, blocks are forced to be small, and the survivors must be
    the ones the reference kept (golden vectors).  The real launch is covered by the -m gpu test."""
    import json
    from pathlib import Path
    from optimized_rag_b200 import data_wrangler, engine
    from test_oracle_golden import dedup_inputs
    golden = json.loads((Path(__file__).parent / "golden" / "golden.json").read_text())

    monkeypatch.setattr(engine, "CosineIndex", _OracleCosineIndex)
    monkeypatch.setattr(data_wrangler, "BLOCK_ROWS", 7)
    for case in golden["dedup"]["cases"]:
        emb = dedup_inputs(case)
        chunks = [{"content": f"c{i}", "n": i} for i in range(case["m"])]
        out = data_wrangler.Deduplicator.semantic_dedup(chunks, [[float(x) for x in e] for e in emb],
                                                        threshold=case["threshold"], device="cpu")
        assert [c["n"] for c in out] == case["kept"], case["name"]
        assert all(o is chunks[o["n"]] for o in out)          # the reference returns the same dict objects
    assert data_wrangler.Deduplicator.semantic_dedup([], [], 0.95, device="cpu") == []
    one = [{"content": "a"}, {"content": "b"}]
    assert data_wrangler.Deduplicator.semantic_dedup(one, [[1.0, 0.0]], 0.95, device="cpu") == one[:1]   # zip semantics


@pytest.mark.parametrize("n,vocab,tile,negative", [(700, 300, 64, False), (40, 8, 32, True), (0, 5, 32, False)])
def test_bm25_index_save_load_round_trip(built_lib, tmp_path, n, vocab, tile, negative):
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
    ix = Bm25Index(torch.from_numpy(doc_off), torch.from_numpy(tok), vocab, tile_docs=tile, doc_id_base=1234)
    assert ix.has_negative_idf == negative and (ix.postings_r16 is None) == (negative or n == 0)
    ix.save(tmp_path / "kw")
    assert (tmp_path / "kw.bin").stat().st_size % 1 == 0 and (tmp_path / "kw.json").exists()
    back = Bm25Index.load(tmp_path / "kw", device="cpu")
    for name in Bm25Index._ARRAYS:
        a, b = getattr(ix, name), getattr(back, name)
        assert (a is None) == (b is None), name
        if a is not None:
            assert a.dtype == b.dtype and a.shape == b.shape and torch.equal(a, b), name
    for key in ("n_docs", "vocab", "tile_docs", "n_tiles", "fp_tile_docs", "fp_n_tiles", "doc_id_base", "n_postings",
                "max_dl", "has_negative_idf", "avgdl", "average_idf", "eps"):
        assert getattr(ix, key) == getattr(back, key), key
    assert back.stats.n_docs == ix.stats.n_docs and back.stats.total_len == ix.stats.total_len
    assert np.array_equal(back.stats.df, ix.stats.df) and np.array_equal(back.stats.first_seen, ix.stats.first_seen)
    for field, ctype in _ffi.Bm25IndexStruct._fields_:
        va, vb = getattr(ix.struct, field), getattr(back.struct, field)
        if field.startswith("d_"):
            assert (va is None) == (vb is None), field   # same arrays present, each pointing at its own copy
        else:
            assert va == vb, field
    # a damaged file is refused
    data = (tmp_path / "kw.bin").read_bytes()
    if len(data) > 16:
        (tmp_path / "kw.bin").write_bytes(data[:-8])
        with pytest.raises(ValueError):
            Bm25Index.load(tmp_path / "kw", device="cpu")


class _OracleCosineIndex:
    """Stand-in for engine.CosineIndex on the CPU: serves the oracle's float64 cosines (host-logic tests only)."""

    def __init__(self, emb, mode="exact"):
        assert mode == "exact" and emb.dtype == torch.float32
        self.rows = emb.numpy()

    def dense(self, q):
        return torch.from_numpy(np.stack([oracle.cosine_scores(self.rows, r) for r in q.numpy()]))


def test_mmr_host_logic_replays_the_reference_golden(built_lib, monkeypatch):
    """MMRDiversifier.diversify with the kernel launch replaced by the oracle's cosines: the greedy loop, tie rule,
    in-place `mmr_score` and top_k > m behaviour must reproduce what the reference recorded (golden.json "mmr")."""
    import json
    from pathlib import Path
    from optimized_rag_b200 import engine, reranker
    from test_oracle_golden import mmr_inputs
    golden = json.loads((Path(__file__).parent / "golden" / "golden.json").read_text())
    monkeypatch.setattr(engine, "CosineIndex", _OracleCosineIndex)
    for case in golden["mmr"]["cases"]:
        emb, q = mmr_inputs(case)
        for run in case["runs"]:
            docs = [{"content": f"d{i}", "embedding": [float(x) for x in emb[i]]} for i in range(case["m"])]
            out = reranker.MMRDiversifier(lambda_param=run["lambda"], device="cpu").diversify(
                [float(x) for x in q], docs, top_k=run["top_k"])
            assert [int(d["content"][1:]) for d in out] == run["picked"], (case["name"], run["lambda"])
            assert [d["mmr_score"].hex() for d in out] == run["mmr_scores"], (case["name"], run["lambda"])


def test_c_abi_rejects_bad_arguments_before_touching_the_gpu(built_lib):
    """Error behaviour of the C ABI that needs no device: every entry point validates its arguments first and
    returns ORAG_EINVAL (-1) / ORAG_EWORKSPACE (-3) with a message in orag_last_error(); size queries are pure host
    arithmetic."""
    import ctypes
    from optimized_rag_b200 import _ffi
    L = _ffi.lib()
    assert L.orag_version() == 1
    # size queries
    W = 2 * 10 + 2 * 16 + 2
    assert L.orag_exchange_bytes(8, 256, 10, 16) == 256 + 2 * 8 * 256 * W * 8
    assert L.orag_exchange_bytes(0, 256, 10, 16) == 0
    assert L.orag_cosine_workspace_bytes(1000, 1536, 0, 10, _ffi.ORAG_COS_F16) == 0
    small = L.orag_cosine_workspace_bytes(1000, 1536, 4, 10, _ffi.ORAG_COS_EXACT)
    assert small >= 4 * 8 + 4 * 1000 * 8
    assert L.orag_cosine_workspace_bytes(10_000_000, 1536, 256, 10, _ffi.ORAG_COS_F16) < 16 << 20
    # argument validation
    rc = L.orag_cosine_topk(None, None, None, None, 10, 1536, 0, None, 4, 10, _ffi.ORAG_COS_F16, None, None, None, None, 0,
                            None)
    assert rc == -1 and b"cosine_topk" in L.orag_last_error()
    with pytest.raises(_ffi.OragError, match="code -1"):
        _ffi.check(rc, "orag_cosine_topk")
    assert L.orag_rrf_fuse(None, 1, 2, 10, 60, 10, 0, None, None, None, None) == -1
    buf = (ctypes.c_int64 * 64)()
    out_i, out_s = (ctypes.c_int64 * 16)(), (ctypes.c_double * 16)()
    p = lambda a: ctypes.cast(a, ctypes.c_void_p)
    assert L.orag_rrf_fuse(p(buf), 1, 9, 10, 60, 10, 0, p(out_i), p(out_s), None, None) == -1       # > 8 lists
    assert b"rrf_fuse sizes" in L.orag_last_error()
    assert L.orag_rrf_fuse(p(buf), 1, 2, 100, 60, 10, 0, p(out_i), p(out_s), None, None) == -1      # union > 128
    assert L.orag_rrf_fuse(p(buf), 0, 2, 10, 60, 10, 0, p(out_i), p(out_s), None, None) == 0        # empty batch: no launch
    assert L.orag_hybrid_merge(p(buf), 40, 1, 10, 16, 60, 10, 0, *([p(buf)] * 9), None) == -1         # 40 * 16 > 256
    assert L.orag_hybrid_push(*([p(buf)] * 6), None, 4, 10, 16, 2, 2, 256, p(buf), 1, None) == -1          # rank >= n_shards
    assert L.orag_hybrid_push(*([p(buf)] * 6), None, 300, 10, 16, 0, 2, 256, p(buf), 1, None) == -1        # batch > max_queries
    assert L.orag_hybrid_push(*([p(buf)] * 6), None, 4, 10, 16, 0, 2, 256, p(buf), 0, None) == -1          # seq starts at 1
    got = ctypes.c_void_p()
    assert L.orag_hybrid_wait(p(buf), 2, 256, 4, 10, 16, 1, 0, ctypes.byref(got), None) == -1         # timeout_ms > 0
    assert L.orag_exchange_alloc(0, ctypes.byref(got)) == -1
    assert L.orag_exchange_free(None) == 0 and L.orag_exchange_close(None) == 0
    assert L.orag_weighted_sum3(None, None, None, 5, 0.5, 0.3, 0.2, None, None) == -1
    assert L.orag_div_scalar(p(buf), 0, 2.0, p(buf), None) == 0
