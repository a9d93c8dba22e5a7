"""CPU oracle for the hybrid-retrieval hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product
(``optimized_rag_b200``) never does: it fails loudly when its CUDA library is
missing instead of falling back to anything in here.

Contents
  oracle.c        plain-C restatement (cosine / BM25 / RRF / weighted hybrid /
                  pairwise), each function citing the reference file:line.
  rank_bm25.py    restatement of the third-party ``rank_bm25.BM25Okapi``
                  (rank-bm25 0.2.2; requirements.txt:22) so that the reference's
                  own ``HybridRetriever._bm25_scores`` glue can run here.
  ref_loader.py   imports the reference's own modules by path (only in the
                  build container, where /root/reference exists).
  pyref.py        pure-Python literal restatement for tiny cases.

Parity pin: see the header of oracle.c and tests/golden/README.md.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "liboracle.so"
_SRC = _HERE / "oracle.c"

_c_i64p = ctypes.POINTER(ctypes.c_int64)
_c_i32p = ctypes.POINTER(ctypes.c_int32)
_c_f32p = ctypes.POINTER(ctypes.c_float)
_c_f64p = ctypes.POINTER(ctypes.c_double)


def build(force: bool = False) -> Path:
    """Compile oracle.c with gcc (no FMA contraction, OpenMP for the row loop)."""
    if not force and _LIB_PATH.exists() and _LIB_PATH.stat().st_mtime >= _SRC.stat().st_mtime:
        return _LIB_PATH
    cmd = ["gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC",
           "-o", str(_LIB_PATH), str(_SRC), "-lm"]
    subprocess.run(cmd, check=True)
    return _LIB_PATH


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(str(_LIB_PATH))
        L.orc_set_threads.restype = ctypes.c_int
        L.orc_set_threads.argtypes = [ctypes.c_int]
        L.orc_cosine.restype = ctypes.c_double
        L.orc_cosine.argtypes = [_c_f32p, _c_f32p, ctypes.c_int, ctypes.c_int]
        L.orc_cosine_scores.restype = None
        L.orc_cosine_scores.argtypes = [_c_f32p, ctypes.c_int64, ctypes.c_int, _c_f32p, ctypes.c_int, _c_f64p]
        L.orc_topk.restype = ctypes.c_int
        L.orc_topk.argtypes = [_c_f64p, ctypes.c_int64, ctypes.c_int, ctypes.c_int64, _c_i64p, _c_f64p]
        L.orc_bm25_build.restype = ctypes.c_void_p
        L.orc_bm25_build.argtypes = [_c_i64p, _c_i32p, ctypes.c_int64, ctypes.c_int32]
        L.orc_bm25_free.restype = None
        L.orc_bm25_free.argtypes = [ctypes.c_void_p]
        for name, rt in [("avgdl", ctypes.c_double), ("average_idf", ctypes.c_double), ("eps", ctypes.c_double),
                         ("n_terms", ctypes.c_int32), ("idf", _c_f64p), ("df", _c_i64p), ("dl", _c_i32p),
                         ("first_seen", _c_i32p), ("num_postings", ctypes.c_int64), ("post_off", _c_i64p),
                         ("post_doc", _c_i32p), ("post_tf", _c_i32p)]:
            f = getattr(L, "orc_bm25_" + name)
            f.restype = rt
            f.argtypes = [ctypes.c_void_p]
        L.orc_bm25_scores_raw.restype = None
        L.orc_bm25_scores_raw.argtypes = [ctypes.c_void_p, _c_i32p, ctypes.c_int, _c_f64p]
        L.orc_bm25_normalize.restype = ctypes.c_double
        L.orc_bm25_normalize.argtypes = [_c_f64p, ctypes.c_int64, _c_f64p]
        L.orc_rrf_fuse.restype = ctypes.c_int
        L.orc_rrf_fuse.argtypes = [_c_i64p, _c_i32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                   _c_i64p, _c_f64p]
        L.orc_weighted_hybrid.restype = None
        L.orc_weighted_hybrid.argtypes = [_c_f64p, _c_f64p, _c_f64p, ctypes.c_int64, ctypes.c_double,
                                          ctypes.c_double, ctypes.c_double, _c_f64p]
        L.orc_pairwise_candidates.restype = ctypes.c_int64
        L.orc_pairwise_candidates.argtypes = [_c_f32p, ctypes.c_int64, ctypes.c_int, _c_i32p, ctypes.c_double,
                                              ctypes.c_int, ctypes.c_int64, _c_i32p, _c_i32p, _c_f64p]
        _lib = L
    return _lib


def _p(a: np.ndarray, typ):
    return a.ctypes.data_as(typ)


def set_threads(n: int | None) -> None:
    """Threads of the OpenMP row loop (overrides an OMP_NUM_THREADS=1 exported by a launcher such as torchrun)."""
    if n:
        os.environ["OMP_NUM_THREADS"] = str(n)
        lib().orc_set_threads(int(n))


# --------------------------------------------------------------------------- cosine
def cosine(a, b, neumaier: bool = True) -> float:
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    assert a.shape == b.shape and a.ndim == 1
    return float(lib().orc_cosine(_p(a, _c_f32p), _p(b, _c_f32p), a.shape[0], int(neumaier)))


def dot(a, b, neumaier: bool = True) -> float:
    """sum(x*y for x, y in zip(a, b)) in the reference's float64 arithmetic (Neumaier `sum` under CPython >= 3.12)."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    L = lib()
    L.orc_dot.restype = ctypes.c_double
    L.orc_dot.argtypes = [_c_f32p, _c_f32p, ctypes.c_int, ctypes.c_int]
    return float(L.orc_dot(_p(a, _c_f32p), _p(b, _c_f32p), min(a.shape[0], b.shape[0]), int(neumaier)))


def cosine_scores(corpus, query, neumaier: bool = True) -> np.ndarray:
    """float64 cosine of `query` against every row of `corpus` (fp32 [n, d])."""
    corpus = np.ascontiguousarray(corpus, dtype=np.float32)
    query = np.ascontiguousarray(query, dtype=np.float32)
    n, d = corpus.shape
    out = np.empty(n, dtype=np.float64)
    lib().orc_cosine_scores(_p(corpus, _c_f32p), n, d, _p(query, _c_f32p), int(neumaier), _p(out, _c_f64p))
    return out


def topk(scores, k: int, id_base: int = 0):
    """(ids, scores) of the first k under (score desc, index asc)."""
    scores = np.ascontiguousarray(scores, dtype=np.float64)
    ids = np.full(k, -1, dtype=np.int64)
    vals = np.zeros(k, dtype=np.float64)
    cnt = lib().orc_topk(_p(scores, _c_f64p), scores.shape[0], k, id_base, _p(ids, _c_i64p), _p(vals, _c_f64p))
    return ids[:cnt].copy(), vals[:cnt].copy()


def cosine_topk(corpus, queries, k: int, id_base: int = 0, neumaier: bool = True):
    """Exact cosine top-k for a batch.  Returns ids int64 [B,k] (-1 padded), scores f64 [B,k]."""
    corpus = np.ascontiguousarray(corpus, dtype=np.float32)
    queries = np.ascontiguousarray(queries, dtype=np.float32).reshape(-1, corpus.shape[1])
    B = queries.shape[0]
    ids = np.full((B, k), -1, dtype=np.int64)
    vals = np.zeros((B, k), dtype=np.float64)
    for b in range(B):
        i, v = topk(cosine_scores(corpus, queries[b], neumaier), k, id_base)
        ids[b, : len(i)] = i
        vals[b, : len(v)] = v
    return ids, vals


# --------------------------------------------------------------------------- BM25
class BM25Index:
    """C restatement of rank_bm25.BM25Okapi over integer token ids."""

    def __init__(self, doc_off, tokens, vocab: int):
        self.doc_off = np.ascontiguousarray(doc_off, dtype=np.int64)
        self.tokens = np.ascontiguousarray(tokens, dtype=np.int32)
        self.n_docs = int(self.doc_off.shape[0] - 1)
        self.vocab = int(vocab)
        if self.tokens.size:
            assert self.tokens.min() >= 0 and self.tokens.max() < vocab
        self._h = lib().orc_bm25_build(_p(self.doc_off, _c_i64p), _p(self.tokens, _c_i32p), self.n_docs, self.vocab)
        L = lib()
        self.avgdl = float(L.orc_bm25_avgdl(self._h))
        self.average_idf = float(L.orc_bm25_average_idf(self._h))
        self.eps = float(L.orc_bm25_eps(self._h))
        self.n_terms = int(L.orc_bm25_n_terms(self._h))

    def _arr(self, name, n, dtype):
        ptr = getattr(lib(), "orc_bm25_" + name)(self._h)
        return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)

    @property
    def idf(self):
        return self._arr("idf", self.vocab, np.float64)

    @property
    def df(self):
        return self._arr("df", self.vocab, np.int64)

    @property
    def dl(self):
        return self._arr("dl", max(self.n_docs, 1), np.int32)[: self.n_docs]

    @property
    def first_seen(self):
        return self._arr("first_seen", max(self.n_terms, 1), np.int32)[: self.n_terms]

    def postings(self):
        P = int(lib().orc_bm25_num_postings(self._h))
        off = self._arr("post_off", self.vocab + 1, np.int64)
        doc = self._arr("post_doc", max(P, 1), np.int32)[:P]
        tf = self._arr("post_tf", max(P, 1), np.int32)[:P]
        return off, doc, tf

    def scores_raw(self, query) -> np.ndarray:
        q = np.ascontiguousarray(query, dtype=np.int32)
        out = np.empty(max(self.n_docs, 1), dtype=np.float64)
        lib().orc_bm25_scores_raw(self._h, _p(q, _c_i32p), int(q.shape[0]), _p(out, _c_f64p))
        return out[: self.n_docs]

    def scores(self, query):
        """Normalised scores (rag/retrieval.py:343-345) and the max used."""
        raw = self.scores_raw(query)
        norm = np.empty_like(raw)
        m = lib().orc_bm25_normalize(_p(raw, _c_f64p), raw.shape[0], _p(norm, _c_f64p)) if raw.size else 1.0
        return norm, float(m)

    def topk(self, query, k: int, id_base: int = 0):
        norm, m = self.scores(query)
        ids, vals = topk(norm, k, id_base)
        return ids, vals, m

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().orc_bm25_free(self._h)
                self._h = None
        except Exception:
            pass


# --------------------------------------------------------------------------- RRF
def rrf_fuse(lists, rrf_k: int = 60, top_k: int = 10, tie: str = "reference"):
    """lists: sequence of 1-D int64 id sequences (rank order).  Returns (ids, scores)."""
    lens = np.array([len(l) for l in lists], dtype=np.int32)
    flat = np.ascontiguousarray(np.concatenate([np.asarray(l, dtype=np.int64).reshape(-1) for l in lists])
                                if len(lists) else np.zeros(0, np.int64))
    if flat.size == 0:
        flat = np.zeros(1, dtype=np.int64)
    out_ids = np.full(max(top_k, 1), -1, dtype=np.int64)
    out_sc = np.zeros(max(top_k, 1), dtype=np.float64)
    cnt = lib().orc_rrf_fuse(_p(flat, _c_i64p), _p(lens, _c_i32p), len(lists), rrf_k, top_k,
                             0 if tie == "reference" else 1, _p(out_ids, _c_i64p), _p(out_sc, _c_f64p))
    return out_ids[:cnt].copy(), out_sc[:cnt].copy()


def weighted_hybrid(sem, kw, temp, alpha, beta, gamma):
    sem = np.ascontiguousarray(sem, dtype=np.float64)
    kw = np.ascontiguousarray(kw, dtype=np.float64)
    out = np.empty_like(sem)
    tp = None
    if temp is not None:
        temp = np.ascontiguousarray(temp, dtype=np.float64)
        tp = _p(temp, _c_f64p)
    lib().orc_weighted_hybrid(_p(sem, _c_f64p), _p(kw, _c_f64p), tp, sem.shape[0], alpha, beta, gamma,
                              _p(out, _c_f64p))
    return out


def pairwise_candidates(emb, doc_idx, thr: float = 0.85, neumaier: bool = True, cap: int = 1 << 20):
    emb = np.ascontiguousarray(emb, dtype=np.float32)
    doc_idx = np.ascontiguousarray(doc_idx, dtype=np.int32)
    m, d = emb.shape
    oi = np.empty(cap, dtype=np.int32)
    oj = np.empty(cap, dtype=np.int32)
    os_ = np.empty(cap, dtype=np.float64)
    cnt = lib().orc_pairwise_candidates(_p(emb, _c_f32p), m, d, _p(doc_idx, _c_i32p), thr, int(neumaier), cap,
                                        _p(oi, _c_i32p), _p(oj, _c_i32p), _p(os_, _c_f64p))
    assert cnt <= cap, "pair cap exceeded"
    return oi[:cnt].copy(), oj[:cnt].copy(), os_[:cnt].copy()


def pairwise_candidates_parallel(emb, doc_idx, thr: float = 0.85, neumaier: bool = True, cap: int = 1 << 22):
    """`pairwise_candidates` on all host cores with 16 partner rows per SIMD block: same pairs, same float64 bits
    (tests/test_oracle_stream.py), fast enough for an 8k-claim sub-block of BASELINE config 5."""
    emb = np.ascontiguousarray(emb, dtype=np.float32)
    doc_idx = np.ascontiguousarray(doc_idx, dtype=np.int32)
    m, d = emb.shape
    L = lib()
    L.orc_pairwise_candidates_blocked.restype = ctypes.c_int64
    L.orc_pairwise_candidates_blocked.argtypes = L.orc_pairwise_candidates.argtypes
    oi = np.empty(cap, dtype=np.int32)
    oj = np.empty(cap, dtype=np.int32)
    os_ = np.empty(cap, dtype=np.float64)
    cnt = L.orc_pairwise_candidates_blocked(_p(emb, _c_f32p), m, d, _p(doc_idx, _c_i32p), thr, int(neumaier), cap,
                                            _p(oi, _c_i32p), _p(oj, _c_i32p), _p(os_, _c_f64p))
    assert cnt <= cap, "pair cap exceeded"
    return oi[:cnt].copy(), oj[:cnt].copy(), os_[:cnt].copy()


# --------------------------------------------------------------------------- hybrid (composition)
def hybrid_topk(corpus, queries, bm25: BM25Index, query_tokens, k: int = 10, rrf_k: int = 60,
                fetch_k: int | None = None, id_base: int = 0):
    """Cosine list + BM25 list -> RRF, per query (the composition README.md:215-220 describes;
    SURVEY.md 'three facts' #1).  Returns per-query dict of the three lists."""
    fetch_k = fetch_k or k
    out = []
    for b in range(len(queries)):
        ci, cv = topk(cosine_scores(corpus, queries[b]), fetch_k, id_base)
        bi, bv, m = bm25.topk(query_tokens[b], fetch_k, id_base)
        fi, fv = rrf_fuse([ci, bi], rrf_k, k)
        out.append({"cos_ids": ci, "cos_scores": cv, "bm25_ids": bi, "bm25_scores": bv, "bm25_max": m,
                    "ids": fi, "rrf_scores": fv})
    return out


# --------------------------------------------------------------------------- MMR (rag/reranker.py:116-193)
def mmr_select(query, embeddings, lambda_param: float, top_k: int):
    """Indices and MMR scores picked by the reference's greedy loop (first maximum wins), cosines from `cosine`."""
    m = len(embeddings)
    selected, scores, remaining = [], [], list(range(m))
    while len(selected) < top_k and remaining:
        best = None
        for j in remaining:
            relevance = cosine(query, embeddings[j])
            diversity = 1 - max(cosine(embeddings[j], embeddings[s]) for s in selected) if selected else 1.0
            score = lambda_param * relevance + (1 - lambda_param) * diversity
            if best is None or score > best[0]:
                best = (score, j)
        selected.append(best[1])
        scores.append(best[0])
        remaining.remove(best[1])
    return selected, scores


# --------------------------------------------------------------------------- semantic dedup (rag/data_wrangler.py:295-326)
def semantic_dedup_keep(embeddings, threshold: float = 0.95):
    """Indices the reference's greedy loop keeps: chunk i survives unless its cosine with an EARLIER SURVIVOR is
    >= threshold (same float64 cosine as everywhere else: Neumaier `sum`, dot / (|a| * |b|), 0 for a zero vector)."""
    emb = np.ascontiguousarray(embeddings, dtype=np.float32)
    kept: list[int] = []
    for i in range(len(emb)):
        if not any(cosine(emb[i], emb[j]) >= threshold for j in kept):
            kept.append(i)
    return kept


# --------------------------------------------------------------------------- streamed oracle (BASELINE-size parity gates)
_stream_ready = False


def _stream_lib():
    global _stream_ready
    L = lib()
    if not _stream_ready:
        u64, i64, i32 = ctypes.c_uint64, ctypes.c_int64, ctypes.c_int
        _c_u64p = ctypes.POINTER(ctypes.c_uint64)
        L.orc_gen_embeddings.restype = None
        L.orc_gen_embeddings.argtypes = [u64, i64, i64, i32, i32, _c_f32p]
        L.orc_cosine_topk_stream.restype = ctypes.c_int
        L.orc_cosine_topk_stream.argtypes = [u64, i64, i64, i32, i32, _c_f32p, _c_f32p, i32, i32, i32, _c_i64p, _c_f64p]
        L.orc_gen_token_corpus.restype = i64
        L.orc_gen_token_corpus.argtypes = [u64, i64, i64, i32, i32, i32, _c_u64p, _c_i64p, _c_i32p]
        L.orc_bm25_stream_stats.restype = ctypes.c_int
        L.orc_bm25_stream_stats.argtypes = [u64, i64, i64, i32, i32, i32, _c_u64p, _c_i64p, _c_i64p, _c_i64p]
        L.orc_bm25_idf_from_df.restype = ctypes.c_double
        L.orc_bm25_idf_from_df.argtypes = [i64, i32, _c_i64p, _c_i32p, i32, _c_f64p, _c_f64p]
        L.orc_bm25_stream_scores.restype = ctypes.c_int
        L.orc_bm25_stream_scores.argtypes = [u64, i64, i64, i32, i32, i32, _c_u64p, ctypes.c_double, _c_f64p, _c_i32p,
                                             _c_i32p, i32, i32, _c_f64p]
        _stream_ready = True
    return L


def gen_embeddings(seed: int, row_start: int, n_rows: int, dim: int = 1536, dup_per_mille: int = 0) -> np.ndarray:
    """C twin of optimized_rag_b200.synthetic.embeddings (tests check the two agree bit for bit)."""
    out = np.empty((n_rows, dim), dtype=np.float32)
    _stream_lib().orc_gen_embeddings(seed, row_start, n_rows, dim, dup_per_mille, _p(out, _c_f32p))
    return out


def gen_token_corpus(seed: int, doc_start: int, n_docs: int, vocab: int, lmin: int, lmax: int, thresholds):
    """C twin of optimized_rag_b200.synthetic.token_corpus -> (doc_off int64 [n+1], tokens int32 [total])."""
    thr = np.ascontiguousarray(thresholds, dtype=np.uint64)
    L = _stream_lib()
    doc_off = np.zeros(n_docs + 1, dtype=np.int64)
    up = ctypes.POINTER(ctypes.c_uint64)
    total = L.orc_gen_token_corpus(seed, doc_start, n_docs, vocab, lmin, lmax, _p(thr, up), _p(doc_off, _c_i64p), None)
    tok = np.empty(max(int(total), 1), dtype=np.int32)
    L.orc_gen_token_corpus(seed, doc_start, n_docs, vocab, lmin, lmax, _p(thr, up), _p(doc_off, _c_i64p), _p(tok, _c_i32p))
    return doc_off, tok[:int(total)]


def cosine_topk_stream(queries, k: int, n_rows: int, row_start: int = 0, seed: int | None = None, dim: int = 1536,
                       dup_per_mille: int = 0, corpus=None, neumaier: bool = True):
    """Exact cosine top-k (reference arithmetic, rag/retrieval.py:362-371 + :320) of a query batch against the
    synthetic rows [row_start, row_start + n_rows) REGENERATED from `seed` block by block (no corpus in memory), or
    against the fp32 block `corpus` whose row r is global row row_start + r.  Queries run side by side (SIMD lanes);
    every lane performs the scalar code's operations, so scores equal `cosine_scores` bit for bit.
    Returns ids int64 [B, k] (global rows, -1 padded), scores float64 [B, k]."""
    queries = np.ascontiguousarray(queries, dtype=np.float32).reshape(-1, dim)
    B = queries.shape[0]
    ids = np.full((B, k), -1, dtype=np.int64)
    sc = np.zeros((B, k), dtype=np.float64)
    cp = None
    if corpus is not None:
        corpus = np.ascontiguousarray(corpus, dtype=np.float32)
        assert corpus.shape == (n_rows, dim)
        cp = _p(corpus, _c_f32p)
    elif seed is None:
        raise ValueError("either a seed or an in-memory corpus block")
    _stream_lib().orc_cosine_topk_stream(seed or 0, row_start, n_rows, dim, dup_per_mille, cp, _p(queries, _c_f32p), B,
                                         k, int(neumaier), _p(ids, _c_i64p), _p(sc, _c_f64p))
    return ids, sc


class StreamedBM25:
    """rank_bm25.BM25Okapi over the synthetic token corpus of `n_docs` docs, WITHOUT holding the corpus: the
    statistics pass (BM25._initialize + BM25Okapi._calc_idf: df, first-seen order, avgdl, idf, epsilon floor) and the
    scoring pass (get_scores, rag/retrieval.py:341-345) both regenerate the tokens doc by doc from the seed.  Same
    arithmetic as `BM25Index`; tests/test_oracle_stream.py requires the two to agree bit for bit."""

    def __init__(self, seed: int, n_docs: int, vocab: int, lmin: int, lmax: int, thresholds):
        self.seed, self.n_docs, self.vocab, self.lmin, self.lmax = seed, int(n_docs), int(vocab), lmin, lmax
        self.thr = np.ascontiguousarray(thresholds, dtype=np.uint64)
        L = _stream_lib()
        up = ctypes.POINTER(ctypes.c_uint64)
        self.df = np.zeros(vocab, dtype=np.int64)
        first = np.zeros(vocab, dtype=np.int64)
        total = ctypes.c_int64(0)
        L.orc_bm25_stream_stats(seed, 0, self.n_docs, vocab, lmin, lmax, _p(self.thr, up), _p(self.df, _c_i64p),
                                _p(first, _c_i64p), ctypes.byref(total))
        self.total_len = int(total.value)
        self.avgdl = self.total_len / self.n_docs if self.n_docs else 0.0
        seen = np.nonzero(self.df > 0)[0]
        self.order = np.ascontiguousarray(seen[np.argsort(first[seen], kind="stable")], dtype=np.int32)
        self.idf = np.zeros(vocab, dtype=np.float64)
        avg = ctypes.c_double(0.0)
        self.eps = float(L.orc_bm25_idf_from_df(self.n_docs, vocab, _p(self.df, _c_i64p), _p(self.order, _c_i32p),
                                                len(self.order), _p(self.idf, _c_f64p), ctypes.byref(avg)))
        self.average_idf = float(avg.value)

    def scores_raw(self, q_terms, q_lens) -> np.ndarray:
        """Raw float64 scores [B, n_docs] of a padded query batch (int32 [B, lq_max], lens int32 [B])."""
        qt = np.ascontiguousarray(q_terms, dtype=np.int32)
        ql = np.ascontiguousarray(q_lens, dtype=np.int32)
        B, lq = qt.shape
        raw = np.empty((B, max(self.n_docs, 1)), dtype=np.float64)
        up = ctypes.POINTER(ctypes.c_uint64)
        _stream_lib().orc_bm25_stream_scores(self.seed, 0, self.n_docs, self.vocab, self.lmin, self.lmax,
                                             _p(self.thr, up), self.avgdl, _p(self.idf, _c_f64p), _p(qt, _c_i32p),
                                             _p(ql, _c_i32p), B, lq, _p(raw, _c_f64p))
        return raw[:, :self.n_docs]

    def topk(self, q_terms, q_lens, k: int):
        """(ids int64 [B,k], normalised scores float64 [B,k], divisors [B]): scores / max (rag/retrieval.py:343-345),
        ranked by (score desc, id asc) as `sorted(..., reverse=True)` over index-ordered input does."""
        raw = self.scores_raw(q_terms, q_lens)
        B = raw.shape[0]
        ids = np.full((B, k), -1, dtype=np.int64)
        sc = np.zeros((B, k), dtype=np.float64)
        mx = np.ones(B, dtype=np.float64)
        norm = np.empty(self.n_docs, dtype=np.float64)
        for b in range(B):
            row = np.ascontiguousarray(raw[b])
            mx[b] = lib().orc_bm25_normalize(_p(row, _c_f64p), self.n_docs, _p(norm, _c_f64p)) if self.n_docs else 1.0
            i, v = topk(norm, k)
            ids[b, :len(i)], sc[b, :len(v)] = i, v
        return ids, sc, mx
