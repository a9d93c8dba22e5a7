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


def _numpy_stats(doc_off, tok, vocab, pos_base=0):
    """Bm25Stats of a token corpus with plain numpy (what csrc/bm25_build.cu's counting pass produces on the GPU)."""
    from optimized_rag_b200.bm25_index import Bm25Stats
    n = len(doc_off) - 1
    df = np.zeros(vocab, dtype=np.int64)
    first = np.full(vocab, np.iinfo(np.int64).max, dtype=np.int64)
    for d in range(n):
        df[np.unique(tok[doc_off[d]:doc_off[d + 1]])] += 1
    for pos in range(len(tok) - 1, -1, -1):
        first[tok[pos]] = pos + pos_base
    return Bm25Stats(n, int(doc_off[-1]), df, first)


def test_sharded_stats_merge_equals_global(built_lib):
    """Host half of the multi-shard index build: per-shard (df, first-seen position) merge into the global statistics,
    and `idf_table` turns those into the oracle's idf / epsilon bit for bit (dict-insertion order included)."""
    from optimized_rag_b200.bm25_index import idf_table, t4_table
    vocab, n = 400, 900
    thr = syn.zipf_thresholds(vocab)
    doc_off, tok = syn.token_corpus(syn.SEED_TOKENS, 0, n, vocab, 5, 50, thr)
    whole = _numpy_stats(doc_off, tok, vocab)
    parts = None
    for s, e in [(0, 300), (300, 650), (650, 900)]:
        st = _numpy_stats(doc_off[s:e + 1] - doc_off[s], tok[doc_off[s]:doc_off[e]], vocab, pos_base=int(doc_off[s]))
        parts = st if parts is None else parts.merged(st)
    assert parts.n_docs == whole.n_docs and parts.total_len == whole.total_len
    assert np.array_equal(parts.df, whole.df) and np.array_equal(parts.first_seen, whole.first_seen)
    a, b = idf_table(parts), idf_table(whole)
    assert np.array_equal(a[0], b[0]) and a[1:] == b[1:]
    orc = oracle.BM25Index(doc_off, tok, vocab)
    assert np.array_equal(a[0].view(np.uint64), orc.idf.view(np.uint64)) and a[1] == orc.average_idf and a[2] == orc.eps
    assert whole.avgdl == orc.avgdl
    order = np.argsort(whole.first_seen[whole.df > 0], kind="stable")
    assert np.array_equal(np.nonzero(whole.df > 0)[0][order], orc.first_seen)
    # t4 = k1 * (1 - b + b * dl / avgdl) in the oracle's operation order
    dl = orc.dl.astype(np.float64)
    assert np.array_equal(t4_table(int(dl.max()), orc.avgdl)[orc.dl], 1.5 * (0.25 + (0.75 * dl) / orc.avgdl))


def test_index_build_refuses_host_tensors(built_lib):
    """The tiled index is built by the library's CUDA builder only (csrc/bm25_build.cu): no CPU / torch stand-in."""
    from optimized_rag_b200 import _ffi
    from optimized_rag_b200.bm25_index import Bm25Index, default_fp_tile_docs
    doc_off = torch.tensor([0, 2, 5], dtype=torch.int64)
    tok = torch.tensor([0, 1, 1, 2, 3], dtype=torch.int32)
    with pytest.raises(_ffi.OragError, match="no CPU builder"):
        Bm25Index(doc_off, tok, 8, tile_docs=32)
    assert default_fp_tile_docs(10_000_000) == 8192 and default_fp_tile_docs(700) == 64 and default_fp_tile_docs(0) == 32
    L = _ffi.lib()
    assert L.orag_bm25_build_workspace_bytes(1000, 50, 64, 256) > 0
    assert L.orag_bm25_build_workspace_bytes(1000, 50, 48, 256) == 0       # tile sizes are powers of two
    assert L.orag_bm25_index_plan(None, None, 10, 50, 64, 256, 0, None, None, None, None, None, None, None, None, None,
                                  0, None, None) == -1


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
    assert L.orag_exchange_bytes(8, 256, 10, 16) == 512 + 6 * 8 * 256 * W * 8   # flags (6 slots x 8 shards x 8 B, padded to 256) + six slots
    assert L.orag_exchange_bytes(0, 256, 10, 16) == 0
    assert L.orag_cosine_workspace_bytes(1000, 1536, 0, 10, _ffi.ORAG_COS_F16) == 0
    small = L.orag_cosine_workspace_bytes(1000, 1536, 4, 10, _ffi.ORAG_COS_EXACT)
    assert small >= 4 * 8 + 4 * 1000 * 8
    assert L.orag_cosine_workspace_bytes(10_000_000, 1536, 256, 10, _ffi.ORAG_COS_F16) < 24 << 20
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


def test_c_abi_rejects_bad_arguments_of_the_round_2_entry_points(built_lib):
    """Same for the entry points added in the second half of round 2: the split-phase cosine search, the per-term
    K-th impacts, the diagnostic hooks."""
    import ctypes
    from optimized_rag_b200 import _ffi
    L = _ffi.lib()
    buf = (ctypes.c_int64 * 64)()
    p = lambda a: ctypes.cast(a, ctypes.c_void_p)
    args = (p(buf), p(buf), p(buf), p(buf), 1000, 1536, 0, p(buf), 4, 10, _ffi.ORAG_COS_F16, p(buf), p(buf), p(buf), p(buf),
            1 << 30)
    assert L.orag_cosine_topk_phase(*args, 0, None) == -1 and b"phases" in L.orag_last_error()
    assert L.orag_cosine_topk_phase(*args, 8, None) == -1
    many = args[:8] + (300,) + args[9:]                     # split phases hold one query group
    assert L.orag_cosine_topk_phase(*many, _ffi.ORAG_PHASE_SCAN, None) == -1 and b"split phases" in L.orag_last_error()
    exact = args[:10] + (_ffi.ORAG_COS_EXACT,) + args[11:]  # ... and exist for the tensor-core modes only
    assert L.orag_cosine_topk_phase(*exact, _ffi.ORAG_PHASE_FINISH, None) == -1
    assert _ffi.ORAG_PHASE_PREP | _ffi.ORAG_PHASE_SCAN | _ffi.ORAG_PHASE_FINISH == _ffi.ORAG_PHASE_ALL
    ix = _ffi.Bm25IndexStruct(n_docs=10, vocab=5)
    assert L.orag_bm25_term_kth(ctypes.byref(ix), p(buf), None) == -1 and b"first-pass view" in L.orag_last_error()
    assert L.orag_bm25_term_kth(None, p(buf), None) == -1
    assert L.orag_cosine_last_counts(None, 1536, 4, p(buf), p(buf), None) == -1
    assert L.orag_cosine_last_counts(p(buf), 1536, 300, p(buf), p(buf), None) == -1
    assert L.orag_timeline_read(None, None, None, 4) == -1
    assert L.orag_timeline_enable(0) == 0


def test_ticket_group_concatenates_the_groups_of_a_split_submission(built_lib):
    """dist._TicketGroup (a submission of more than 256 queries is split into 256-query groups over the lanes): `wait`
    returns one dict with every tensor concatenated in submission order."""
    import torch
    from optimized_rag_b200.dist import _TicketGroup

    class Fake:
        def __init__(self, lo, n):
            self.out = {"ids": torch.arange(lo, lo + n).view(n, 1), "status": torch.zeros(n, dtype=torch.int32),
                        "src": torch.full((n, 2, 2), lo)}

        def wait(self):
            return self.out

    got = _TicketGroup([Fake(0, 256), Fake(256, 256), Fake(512, 88)]).wait()
    assert got["ids"].shape == (600, 1) and got["ids"][:, 0].tolist() == list(range(600))
    assert got["status"].shape == (600,) and got["src"].shape == (600, 2, 2) and int(got["src"][599, 0, 0]) == 512
