"""BASELINE-size parity gate -- TEST INFRASTRUCTURE (tests/test_zz_gpu_parity_scale.py and bench.py's
`verified_against_oracle` / `cpu_baseline` leg; never imported by the product).

SURVEY.md §8(d) "Parity gates": the CUDA path must equal the oracle at N <= 1M against the full oracle and at 10M
against the streamed oracle on a query subset (>= 32 queries).  The 10M-row inputs do not fit host memory, so the
oracle side regenerates them from the seeds (oracle.c "Streamed oracle"); this module only wires the pieces
together and compares, array by array, bit by bit:

    cosine   ids + float64 score bits      rag/retrieval.py:362-371, ranking rule :320
    BM25     ids + normalised score bits + divisor   rank_bm25 0.2.2 get_scores + rag/retrieval.py:324-347
    RRF      ids + float64 score bits      rag/reranker.py:224-271 (ties: insertion order)
"""
from __future__ import annotations

import time

import numpy as np

import oracle


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def pick_queries(qt, ql, n_total: int, n_pick: int = 32) -> np.ndarray:
    """Query subset for a full-size check: evenly spaced over the batch plus the special keyword queries (a rank<10
    term -> epsilon idf, df ~ N posting list; an OOV token + a duplicated term).  Sorted indices, len <= n_pick."""
    n_pick = min(n_pick, n_total)
    special = [b for b in range(n_total) if ((qt[b, :ql[b]] >= 0) & (qt[b, :ql[b]] < 10)).any()][:5] + \
              [b for b in range(n_total) if (qt[b, :ql[b]] == -1).any()][:3]
    special = special[:max(n_pick // 4, 1)]
    even = [int(x) for x in np.linspace(0, n_total - 1, n_pick)]
    picked = sorted(set(special + even))
    while len(picked) > n_pick:
        drop = next((b for b in picked[1:-1] if b not in special), picked[1])
        picked.remove(drop)
    return np.array(picked, dtype=np.int64)


def reference_lists(n_rows: int, dim: int, q_emb, q_terms, q_lens, k: int, fetch_k: int | None = None, rrf_k: int = 60,
                    *, seed_corpus: int, bm25: "oracle.StreamedBM25 | None" = None, want_cosine: bool = True):
    """Oracle results for a query subset at full corpus size.  Returns (dict of arrays, dict of seconds)."""
    fetch_k = fetch_k or k
    out, secs = {}, {}
    B = len(q_emb) if want_cosine else len(q_terms)
    if want_cosine:
        t0 = time.perf_counter()
        out["cos_ids"], out["cos_scores"] = oracle.cosine_topk_stream(q_emb, fetch_k, n_rows, seed=seed_corpus, dim=dim)
        secs["cosine"] = time.perf_counter() - t0
    if bm25 is not None:
        t0 = time.perf_counter()
        out["bm25_ids"], out["bm25_scores"], out["bm25_max"] = bm25.topk(q_terms, q_lens, fetch_k)
        secs["bm25"] = time.perf_counter() - t0
    if want_cosine and bm25 is not None:
        t0 = time.perf_counter()
        ids = np.full((B, k), -1, dtype=np.int64)
        sc = np.zeros((B, k), dtype=np.float64)
        for b in range(B):
            ci = out["cos_ids"][b][out["cos_ids"][b] >= 0]
            bi = out["bm25_ids"][b][out["bm25_ids"][b] >= 0]
            fi, fs = oracle.rrf_fuse([ci, bi], rrf_k, k)
            ids[b, :len(fi)], sc[b, :len(fs)] = fi, fs
        out["ids"], out["rrf_scores"] = ids, sc
        secs["rrf"] = time.perf_counter() - t0
    return out, secs


def compare(got: dict, want: dict, rows=None) -> list[str]:
    """Mismatches between the CUDA result arrays `got` (numpy, one row per query of the subset, or the full batch with
    `rows` selecting the subset) and the oracle's `want`.  Integer arrays compare exactly, float64 arrays bitwise."""
    bad = []
    for key, w in want.items():
        if key not in got:
            bad.append(f"{key}: missing from the CUDA result")
            continue
        g = np.asarray(got[key])
        if rows is not None:
            g = g[rows]
        if w.dtype.kind == "f":
            same = g.shape == w.shape and np.array_equal(_bits(g), _bits(w))
        else:
            same = g.shape == w.shape and np.array_equal(g, w)
        if not same:
            where = np.argwhere(_bits(g) != _bits(w)) if (w.dtype.kind == "f" and g.shape == w.shape) else \
                (np.argwhere(g != w) if g.shape == w.shape else [])
            first = tuple(where[0]) if len(where) else None
            detail = f" first at {first}: got {g[first]!r} want {w[first]!r}" if first is not None else ""
            bad.append(f"{key}: {len(where)} of {w.size} entries differ{detail}")
    return bad
