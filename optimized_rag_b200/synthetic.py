"""Deterministic synthetic inputs (host side, numpy) -- SURVEY.md §8(d).

Counter-based integer hashing only, so these numpy generators and the CUDA fill
kernels in csrc/gen.cu produce BIT-IDENTICAL data for any shard layout:

  mix(z)          splitmix64 finaliser
  row_key(s, r)   mix(s ^ (r * 0xD1B54A32D192ED03))
  emb(s, r, c)    (int(mix(row_key + c) >> 40) - 2**23) * 2**-28   -> fp32, exact,
                  uniform in [-2**-5, 2**-5); row norm ~ 0.71 (NOT unit: the
                  reference divides by both magnitudes, rag/retrieval.py:365-371)
  duplicates      with dup_per_mille > 0, row r > 0 is an exact copy of an earlier
                  row with probability dup_per_mille/1000 (forces exact score ties)
  doc_len(s, d)   Lmin + mix(s ^ (d * 0x9FB21C651E98DF25)) % (Lmax - Lmin + 1)
  token(s, d, j)  Zipf(s=1) rank via inverse CDF on a u63 threshold table

Seeds: corpus 0x5EED0001, queries 0x5EED0002, tokens 0x5EED0003,
keyword queries 0x5EED0004.
"""
from __future__ import annotations

import numpy as np

SEED_CORPUS = 0x5EED0001
SEED_QUERIES = 0x5EED0002
SEED_TOKENS = 0x5EED0003
SEED_KWQUERIES = 0x5EED0004

_K_ROW = np.uint64(0xD1B54A32D192ED03)
_K_DOC = np.uint64(0x9FB21C651E98DF25)
_K_DUP = np.uint64(0xA24BAED4963EE407)
_G = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)

EMB_SCALE = 2.0 ** -28
QUERY_STRIDE = 7919


def mix(z):
    z = np.asarray(z, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = z + _G
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        return z ^ (z >> np.uint64(31))


def row_key(seed: int, rows):
    rows = np.asarray(rows, dtype=np.uint64)
    with np.errstate(over="ignore"):
        return mix(np.uint64(seed) ^ (rows * _K_ROW))


def source_rows(seed: int, rows, dup_per_mille: int = 0):
    """Row whose values row r carries (r itself unless it is a planted duplicate)."""
    rows = np.asarray(rows, dtype=np.uint64)
    if dup_per_mille <= 0:
        return rows
    with np.errstate(over="ignore"):
        h = mix(np.uint64(seed) ^ _K_DUP ^ (rows * _K_DOC))
        is_dup = ((h % np.uint64(1000)) < np.uint64(dup_per_mille)) & (rows > 0)
        src = mix(h) % np.maximum(rows, np.uint64(1))
    return np.where(is_dup, src, rows)


def embeddings(seed: int, row_start: int, n_rows: int, dim: int = 1536, dup_per_mille: int = 0) -> np.ndarray:
    """fp32 [n_rows, dim] for global rows [row_start, row_start + n_rows)."""
    rows = np.arange(row_start, row_start + n_rows, dtype=np.uint64)
    src = source_rows(seed, rows, dup_per_mille)
    key = row_key(seed, src)[:, None]
    cols = np.arange(dim, dtype=np.uint64)[None, :]
    with np.errstate(over="ignore"):
        h = mix(key + cols)
    v = (h >> np.uint64(40)).astype(np.int64) - (1 << 23)
    return (v.astype(np.float32) * np.float32(EMB_SCALE)).astype(np.float32)


def query_embeddings(n_queries: int, n_corpus: int, dim: int = 1536, corpus_seed: int = SEED_CORPUS,
                     query_seed: int = SEED_QUERIES, dup_per_mille: int = 0) -> np.ndarray:
    """Query b = corpus row (b*7919) mod N + 0.5 * fresh noise (fp32 round-to-nearest add)."""
    b = np.arange(n_queries, dtype=np.uint64)
    target = (b * np.uint64(QUERY_STRIDE)) % np.uint64(max(n_corpus, 1))
    base = np.concatenate([embeddings(corpus_seed, int(t), 1, dim, dup_per_mille) for t in target], axis=0) \
        if n_queries else np.zeros((0, dim), np.float32)
    noise = embeddings(query_seed, 0, n_queries, dim, 0)
    return (base + np.float32(0.5) * noise).astype(np.float32)


# --------------------------------------------------------------------------- Zipf token corpus
def zipf_thresholds(vocab: int = 50000) -> np.ndarray:
    """u63 thresholds T[i] = floor(CDF(i) * 2**63); token rank = first i with (u >> 1) < T[i]."""
    w = 1.0 / np.arange(1, vocab + 1, dtype=np.float64)
    cdf = np.cumsum(w)
    cdf = cdf / cdf[-1]
    t = np.floor(cdf * float(1 << 63))
    t = np.minimum(t, float(1 << 63))
    out = t.astype(np.uint64)
    out[-1] = np.uint64(1 << 63)
    return out


def doc_lengths(seed: int, doc_start: int, n_docs: int, lmin: int = 100, lmax: int = 300) -> np.ndarray:
    d = np.arange(doc_start, doc_start + n_docs, dtype=np.uint64)
    with np.errstate(over="ignore"):
        h = mix(np.uint64(seed) ^ (d * _K_DOC))
    return (lmin + (h % np.uint64(lmax - lmin + 1))).astype(np.int32)


def token_corpus(seed: int, doc_start: int, n_docs: int, vocab: int = 50000, lmin: int = 100, lmax: int = 300,
                 thresholds: np.ndarray | None = None):
    """Returns (doc_off int64 [n_docs+1], tokens int32 [total]) for global docs [doc_start, +n_docs)."""
    if thresholds is None:
        thresholds = zipf_thresholds(vocab)
    lens = doc_lengths(seed, doc_start, n_docs, lmin, lmax).astype(np.int64)
    doc_off = np.zeros(n_docs + 1, dtype=np.int64)
    np.cumsum(lens, out=doc_off[1:])
    total = int(doc_off[-1])
    docs = np.repeat(np.arange(doc_start, doc_start + n_docs, dtype=np.uint64), lens)
    pos = (np.arange(total, dtype=np.int64) - np.repeat(doc_off[:-1], lens)).astype(np.uint64)
    with np.errstate(over="ignore"):
        u = mix(row_key(seed + 1, docs) + pos) >> np.uint64(1)
    tok = np.searchsorted(thresholds, u, side="right")
    tok = np.minimum(tok, vocab - 1).astype(np.int32)
    return doc_off, tok


def keyword_queries(n_queries: int, vocab: int = 50000, seed: int = SEED_KWQUERIES, lq_min: int = 3,
                    lq_max: int = 8, min_rank: int = 100, thresholds: np.ndarray | None = None):
    """int32 [B, lq_max] padded with -2, lens int32 [B].

    Terms ~ Zipf over ranks >= min_rank; 5 % of queries get a rank<10 term in slot 0
    (exercises the negative-idf -> epsilon branch); 2 % get an OOV term (-1) in slot 1
    and a duplicate of slot 0 in slot 2."""
    if thresholds is None:
        thresholds = zipf_thresholds(vocab)
    min_rank = min(min_rank, max(vocab - 1, 0))
    b = np.arange(n_queries, dtype=np.uint64)
    with np.errstate(over="ignore"):
        hl = mix(np.uint64(seed) ^ (b * _K_DOC))
        lens = (lq_min + (hl % np.uint64(lq_max - lq_min + 1))).astype(np.int32)
        key = row_key(seed + 1, b)[:, None]
        j = np.arange(lq_max, dtype=np.uint64)[None, :]
        u = mix(key + j) >> np.uint64(1)
        lo = thresholds[min_rank - 1] if min_rank > 0 else np.uint64(0)
        span = np.uint64(1 << 63) - lo
        u = lo + (u % span)
        tok = np.minimum(np.searchsorted(thresholds, u, side="right"), vocab - 1).astype(np.int32)
        h5 = mix(np.uint64(seed + 2) ^ (b * _K_ROW))
        common = (h5 % np.uint64(100)) < np.uint64(5)
        tok[common, 0] = (mix(h5[common]) % np.uint64(min(10, vocab))).astype(np.int32)
        h2 = mix(np.uint64(seed + 3) ^ (b * _K_ROW))
        odd = (h2 % np.uint64(100)) < np.uint64(2)
    tok[odd, 1] = -1
    tok[odd, 2] = tok[odd, 0]
    mask = np.arange(lq_max)[None, :] >= lens[:, None]
    tok[mask] = -2
    return tok, lens


def tokens_to_text(tokens) -> str:
    """Chunk text for the Python reference: ' '.join(f't{id}') so lower().split() is exact."""
    return " ".join(f"t{int(t)}" if t >= 0 else "oov" for t in tokens)
