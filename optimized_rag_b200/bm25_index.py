"""GPU-resident, doc-range-tiled BM25 index (ingest side; torch is used for the sort/scan plumbing).

Layout consumed by csrc/bm25.cu (see include/orag.h `orag_bm25_index_t`):
  tile t = docs [t*T, (t+1)*T) of the shard
  postings      uint32 [(doc_in_tile << 16) | tf], grouped by (tile, term), ascending doc
  tile_base     int64 [n_tiles+1]   first posting of each tile
  tile_term_off int32 [n_tiles, V+1] offsets of each term's run inside its tile
  doc_len       int32 [n_docs]; t4_table float64 [max_dl+1] = k1*(1 - b + b*dl/avgdl) (global avgdl)
  idf           float64 [V]         global idf with the epsilon floor (rank_bm25 0.2.2 BM25Okapi._calc_idf)
  first-pass view (csrc/bm25_ms.cu), a second tiling with tiles of fp_tile_docs (up to 16384) docs:
  postings_r16  uint32 [(doc_in_tile << 16) | fp16(tf*(k1+1)/(tf + t4[dl]))], fp_tile_base, fp_tile_term_off
  term_max_r    float32 [V]         max fp16 r over this shard's postings of each term (MaxScore upper bounds)

Global statistics (N, avgdl, df, first-seen order -> idf, eps) follow the reference's
`BM25Okapi(tokenized_corpus)` construction at rag/retrieval.py:338; they are computed over the WHOLE
corpus (all shards) so that sharded and single-GPU results are identical.
"""
from __future__ import annotations

import ctypes
import math
from dataclasses import dataclass

import numpy as np
import torch

from . import _ffi

K1 = 1.5
B = 0.75
EPSILON = 0.25


@dataclass
class Bm25Stats:
    """Corpus-wide statistics; `first_seen` is the global token position of each term's first occurrence
    (int64 max for unseen terms) -- its argsort is the dict-insertion order rank_bm25 sums idf in."""
    n_docs: int
    total_len: int
    df: np.ndarray          # int64 [V]
    first_seen: np.ndarray  # int64 [V]

    @property
    def avgdl(self) -> float:
        return self.total_len / self.n_docs if self.n_docs else 0.0

    def merged(self, other: "Bm25Stats") -> "Bm25Stats":
        return Bm25Stats(self.n_docs + other.n_docs, self.total_len + other.total_len, self.df + other.df,
                         np.minimum(self.first_seen, other.first_seen))


def idf_table(stats: Bm25Stats):
    """rank_bm25 BM25Okapi._calc_idf: idf = log(N - df + .5) - log(df + .5) walked in first-seen order with a
    running `+=` sum; negatives replaced by epsilon * average_idf.  Returns (idf float64 [V], average_idf, eps)."""
    V = stats.df.shape[0]
    idf = np.zeros(V, dtype=np.float64)
    seen = np.nonzero(stats.df > 0)[0]
    order = seen[np.argsort(stats.first_seen[seen], kind="stable")]
    n = stats.n_docs
    running = 0
    negatives = []
    df = stats.df
    for t in order.tolist():
        f = int(df[t])
        v = math.log(n - f + 0.5) - math.log(f + 0.5)
        idf[t] = v
        running += v
        if v < 0:
            negatives.append(t)
    average_idf = running / len(order) if len(order) else 0.0
    eps = EPSILON * average_idf
    for t in negatives:
        idf[t] = eps
    return idf, average_idf, eps


def t4_table(max_dl: int, avgdl: float) -> np.ndarray:
    """k1 * (1 - b + b * dl / avgdl) for dl = 0..max_dl with numpy float64 in rank_bm25's operation order."""
    dl = np.arange(max_dl + 1)
    if avgdl == 0:
        return np.full(max_dl + 1, K1 * (1 - B), dtype=np.float64)
    return K1 * (1 - B + B * dl / avgdl)


def local_stats(doc_off: torch.Tensor, tokens: torch.Tensor, vocab: int, token_pos_base: int = 0,
                chunk_docs: int = 1 << 20) -> Bm25Stats:
    """df / first-seen / lengths of one shard.  doc_off int64 [n+1], tokens int32 [total] (same device)."""
    dev = tokens.device
    n = doc_off.numel() - 1
    df = torch.zeros(vocab, dtype=torch.int64, device=dev)
    big = torch.iinfo(torch.int64).max
    first = torch.full((vocab,), big, dtype=torch.int64, device=dev)
    total = int(doc_off[-1].item()) if n > 0 else 0
    for d0 in range(0, n, chunk_docs):
        d1 = min(n, d0 + chunk_docs)
        lo, hi = int(doc_off[d0].item()), int(doc_off[d1].item())
        if hi == lo:
            continue
        tok = tokens[lo:hi].long()
        pos = torch.arange(lo, hi, dtype=torch.int64, device=dev) + token_pos_base
        first.scatter_reduce_(0, tok, pos, reduce="amin", include_self=True)
        lens = (doc_off[d0 + 1:d1 + 1] - doc_off[d0:d1])
        doc = torch.repeat_interleave(torch.arange(d1 - d0, dtype=torch.int64, device=dev), lens)
        pair = torch.unique(doc * vocab + tok)
        df += torch.bincount(pair % vocab, minlength=vocab)
        del tok, pos, doc, pair
    return Bm25Stats(n, total, df.cpu().numpy(), first.cpu().numpy())


class Bm25Index:
    """One shard of the inverted index, resident on `device`."""

    def __init__(self, doc_off: torch.Tensor, tokens: torch.Tensor, vocab: int, tile_docs: int = 1024,
                 stats: Bm25Stats | None = None, doc_id_base: int = 0, chunk_docs: int = 1 << 20,
                 first_pass: bool = True, fp_tile_docs: int | None = None):
        assert doc_off.dtype == torch.int64 and tokens.dtype == torch.int32
        assert tile_docs & (tile_docs - 1) == 0 and 32 <= tile_docs <= 2048
        dev = tokens.device
        self.device = dev
        self.vocab = int(vocab)
        self.tile_docs = int(tile_docs)
        self.n_docs = int(doc_off.numel() - 1)
        self.doc_id_base = int(doc_id_base)
        self.n_tiles = (self.n_docs + tile_docs - 1) // tile_docs
        if stats is None:
            stats = local_stats(doc_off, tokens, vocab, chunk_docs=chunk_docs)
        self.stats = stats
        idf, self.average_idf, self.eps = idf_table(stats)
        self.avgdl = stats.avgdl
        self.has_negative_idf = bool((idf < 0).any())
        self.idf = torch.from_numpy(idf).to(dev)

        dl = (doc_off[1:] - doc_off[:-1])
        self.dl = dl.to(torch.int32)
        max_dl = int(dl.max().item()) if self.n_docs else 0
        if max_dl > 0xFFFF:
            raise ValueError("documents longer than 65535 tokens are not representable in the tile layout")
        self.max_dl = max_dl
        t4_np = t4_table(max_dl, self.avgdl)
        self.t4_table = torch.from_numpy(t4_np).to(dev)
        # r[dl, tf-1] = tf*(k1+1) / (tf + t4[dl]) for tf = 1..4, numpy float64 in rank_bm25's operation order
        tf_np = np.arange(1, 5)[None, :]
        self.r_table = torch.from_numpy(np.ascontiguousarray(tf_np * (K1 + 1) / (tf_np + t4_np[:, None]))).to(dev)

        # exact view: (doc_in_tile << 16) | tf over tiles of `tile_docs` docs
        self.postings, self.tile_base, self.tile_term_off, _ = self._tiled_view(doc_off, tokens, dl, self.tile_docs,
                                                                               chunk_docs, want_r16=False)
        self.n_postings = int(self.tile_base[-1].item())
        # first-pass view: (doc_in_tile << 16) | fp16(r) over (much larger) tiles of `fp_tile_docs` docs
        self.postings_r16 = self.term_max_r = self.fp_tile_base = self.fp_tile_term_off = None
        if fp_tile_docs is None:
            # measured on B200 (scripts/debug_bm25.py): 4096..8192-doc tiles are within 5 % of each other stand-alone;
            # 4096 keeps the per-warp bitmap small enough for the background configuration (csrc/bm25_ms.cu)
            fp_tile_docs = min(4096, max(32, 1 << max(0, (max(self.n_docs, 1) // 16 - 1).bit_length())))
        assert fp_tile_docs & (fp_tile_docs - 1) == 0 and 32 <= fp_tile_docs <= 16384
        self.fp_tile_docs = int(fp_tile_docs)
        self.fp_n_tiles = (self.n_docs + self.fp_tile_docs - 1) // self.fp_tile_docs
        if first_pass and not self.has_negative_idf and self.n_docs > 0:
            post, base, off, tmax = self._tiled_view(doc_off, tokens, dl, self.fp_tile_docs, chunk_docs, want_r16=True)
            if post is not None:
                self.postings_r16, self.fp_tile_base, self.fp_tile_term_off, self.term_max_r = post, base, off, tmax
                assert self.postings_r16.data_ptr() % 16 == 0

        self.struct = _ffi.Bm25IndexStruct(
            n_docs=self.n_docs, vocab=self.vocab, tile_docs=self.tile_docs, n_tiles=self.n_tiles,
            has_negative_idf=int(self.has_negative_idf),
            d_tile_base=self.tile_base.data_ptr(), d_tile_term_off=self.tile_term_off.data_ptr(),
            max_doc_len=self.max_dl, reserved=0, d_postings=self.postings.data_ptr(), d_doc_len=self.dl.data_ptr(),
            d_t4_table=self.t4_table.data_ptr(), d_r_table=self.r_table.data_ptr(), d_idf=self.idf.data_ptr(),
            d_postings_r16=self.postings_r16.data_ptr() if self.postings_r16 is not None else None,
            d_term_max_r=self.term_max_r.data_ptr() if self.term_max_r is not None else None,
            fp_tile_docs=self.fp_tile_docs, fp_n_tiles=self.fp_n_tiles,
            d_fp_tile_base=self.fp_tile_base.data_ptr() if self.fp_tile_base is not None else None,
            d_fp_tile_term_off=self.fp_tile_term_off.data_ptr() if self.fp_tile_term_off is not None else None)
        self._ws = None

    def _tiled_view(self, doc_off, tokens, dl, T: int, chunk_docs: int, want_r16: bool):
        """Postings sorted by (tile, term, doc_in_tile) for tiles of T docs -> (postings int32, tile_base int64
        [n_tiles+1], tile_term_off int32 [n_tiles, V+1], term_max_r float32 [V] | None).  With want_r16 the low
        half-word is fp16(tf*(k1+1)/(tf + t4[dl])) instead of tf; returns (None,)*4 if an r is not a normal fp16."""
        dev = self.device
        V1 = self.vocab + 1
        n_tiles = (self.n_docs + T - 1) // T
        chunk_docs = max(T, chunk_docs // T * T)
        post_chunks, cnt_chunks = [], []
        term_max = torch.zeros(self.vocab, dtype=torch.float32, device=dev) if want_r16 else None
        for d0 in range(0, self.n_docs, chunk_docs):
            d1 = min(self.n_docs, d0 + chunk_docs)
            lo, hi = int(doc_off[d0].item()), int(doc_off[d1].item())
            n_t = (d1 - d0 + T - 1) // T
            if hi == lo:
                cnt_chunks.append(torch.zeros(n_t * self.vocab, dtype=torch.int64, device=dev))
                continue
            tok = tokens[lo:hi].long()
            doc = torch.repeat_interleave(torch.arange(d1 - d0, dtype=torch.int64, device=dev), dl[d0:d1])
            # key orders postings by (tile, term, doc_in_tile)
            key = ((doc // T) * self.vocab + tok) * T + (doc % T)
            del tok, doc
            key, tf = torch.unique(key, return_counts=True)  # sorted
            if int(tf.max().item()) > 0xFFFF:
                raise ValueError("term frequency above 65535 is not representable in the posting format")
            tt = key // T
            cnt_chunks.append(torch.bincount(tt, minlength=n_t * self.vocab))
            if not want_r16:
                post_chunks.append((((key % T) << 16) | tf).to(torch.int32))
            else:
                doc_abs = (tt // self.vocab) * T + (key % T) + d0
                tf_f = tf.double()
                r = tf_f * (K1 + 1) / (tf_f + self.t4_table[self.dl[doc_abs].long()])
                r16 = r.to(torch.float32).to(torch.float16)
                if float(r16.min().item()) < 6.2e-5 or not bool(torch.isfinite(r16).all()):
                    return None, None, None, None
                term_max.scatter_reduce_(0, tt % self.vocab, r16.float(), reduce="amax", include_self=True)
                post_chunks.append((((key % T) << 16) | (r16.view(torch.int16).long() & 0xFFFF)).to(torch.int32))
                del doc_abs, tf_f, r, r16
            del key, tf, tt
        post_chunks.append(torch.zeros(4, dtype=torch.int32, device=dev))  # padding (16-byte reads never fault)
        postings = torch.cat(post_chunks)
        del post_chunks
        if cnt_chunks:
            counts = torch.cat(cnt_chunks).view(n_tiles, self.vocab)
        else:
            counts = torch.zeros((0, self.vocab), dtype=torch.int64, device=dev)
        off = torch.zeros((n_tiles, V1), dtype=torch.int64, device=dev)
        torch.cumsum(counts, dim=1, out=off[:, 1:])
        per_tile = off[:, -1] if n_tiles else torch.zeros(0, dtype=torch.int64, device=dev)
        tile_base = torch.zeros(n_tiles + 1, dtype=torch.int64, device=dev)
        if n_tiles:
            torch.cumsum(per_tile, dim=0, out=tile_base[1:])
            assert int(per_tile.max().item()) < 2 ** 31
        return postings, tile_base, off.to(torch.int32).contiguous(), term_max

    @property
    def doc_t4(self) -> torch.Tensor:
        return self.t4_table[self.dl.long()]

    # ------------------------------------------------------------------ queries
    def _workspace(self, n_queries: int, k: int, flags: int) -> torch.Tensor:
        need = int(_ffi.lib().orag_bm25_workspace_bytes(ctypes.byref(self.struct), n_queries, k, flags))
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(max(need, 256), dtype=torch.uint8, device=self.device)
        return self._ws

    def topk(self, query_terms: torch.Tensor, query_lens: torch.Tensor, k: int, normalize: bool = True,
             force: str | None = None, check_overflow: bool = True, status_out: list | None = None,
             background: bool = False):
        """query_terms int32 [B, max_terms] (negative = OOV/padding), query_lens int32 [B].
        `status_out` (a list) receives the per-query status tensor so that a caller can defer the overflow
        check and pay one host sync for several calls (see engine.HybridShard.local_lists).
        Returns ids int64 [B,k], scores f64 [B,k], max f64 [B] (divisor if normalize else shard max raw)."""
        assert query_terms.dtype == torch.int32 and query_lens.dtype == torch.int32
        assert query_terms.is_cuda and query_terms.is_contiguous() and query_lens.is_contiguous()
        Bq, mt = query_terms.shape
        flags = (_ffi.ORAG_BM25_NORMALIZE if normalize else 0)
        flags |= {None: 0, "sparse": _ffi.ORAG_BM25_FORCE_SPARSE, "dense": _ffi.ORAG_BM25_FORCE_DENSE,
                  "exact_tiles": _ffi.ORAG_BM25_FORCE_SPARSE | _ffi.ORAG_BM25_EXACT_TILES}[force]
        if background:
            flags |= _ffi.ORAG_BM25_BACKGROUND
        ids = torch.empty((Bq, k), dtype=torch.int64, device=self.device)
        sc = torch.empty((Bq, k), dtype=torch.float64, device=self.device)
        mx = torch.empty(Bq, dtype=torch.float64, device=self.device)
        status = torch.empty(Bq, dtype=torch.int32, device=self.device)
        ws = self._workspace(Bq, k, flags)
        st = torch.cuda.current_stream(self.device).cuda_stream
        _ffi.check(_ffi.lib().orag_bm25_topk(ctypes.byref(self.struct), self.doc_id_base, query_terms.data_ptr(),
                                             query_lens.data_ptr(), Bq, mt, k, flags, ids.data_ptr(), sc.data_ptr(),
                                             mx.data_ptr(), status.data_ptr(), ws.data_ptr(), ws.numel(), st),
                   "orag_bm25_topk")
        if status_out is not None:
            status_out.append(status)
        if check_overflow and force != "dense":
            bad = torch.nonzero(status != 0).flatten()
            if bad.numel():
                # candidate buffer overflowed for these queries: exact dense path (never a CPU fallback)
                i2, s2, m2 = self.topk(query_terms[bad].contiguous(), query_lens[bad].contiguous(), k, normalize,
                                       force="dense", check_overflow=False)
                ids[bad], sc[bad], mx[bad] = i2, s2, m2
        return ids, sc, mx

    def dense_scores(self, query_terms: torch.Tensor, query_lens: torch.Tensor) -> torch.Tensor:
        """Raw float64 scores [B, n_docs] (tests / small corpora / weighted hybrid)."""
        Bq, mt = query_terms.shape
        out = torch.empty((Bq, max(self.n_docs, 1)), dtype=torch.float64, device=self.device)[:, :self.n_docs]
        out = out.contiguous()
        st = torch.cuda.current_stream(self.device).cuda_stream
        _ffi.check(_ffi.lib().orag_bm25_dense(ctypes.byref(self.struct), query_terms.data_ptr(),
                                              query_lens.data_ptr(), Bq, mt, out.data_ptr(), st), "orag_bm25_dense")
        return out

    def posting_bytes(self, query_terms: torch.Tensor, query_lens: torch.Tensor) -> int:
        """Algorithmic bytes of a batch: sum over query tokens of df_shard(t) * 6."""
        off = self.tile_term_off.long()
        df_shard = (off[:, 1:] - off[:, :-1]).sum(dim=0)
        idf_nz = self.idf != 0
        t = query_terms.long()
        mask = (t >= 0) & (t < self.vocab) & (torch.arange(t.shape[1], device=t.device)[None, :] < query_lens[:, None])
        tc = t.clamp(0, self.vocab - 1)
        per = torch.where(mask & idf_nz[tc], df_shard[tc], torch.zeros_like(tc))
        return int(per.sum().item()) * 6

    # ------------------------------------------------------------------ on-disk format
    # The reference's "format" for the keyword side is nothing at all: `BM25Okapi(tokenized_corpus)` is rebuilt from
    # the Postgres rows on every call (rag/retrieval.py:334-338).  Here one index = two files:
    #   <stem>.json   scalars (sizes, avgdl, eps, ...) + a manifest {name: dtype, shape, byte offset}
    #   <stem>.bin    the arrays of the layout described at the top of this module, little-endian, each aligned to
    #                 256 bytes, exactly as they sit in HBM -> loading is read (or mmap) + one copy per array
    FORMAT_VERSION = 1
    _ARRAYS = ("idf", "dl", "t4_table", "r_table", "postings", "tile_base", "tile_term_off",
               "postings_r16", "fp_tile_base", "fp_tile_term_off", "term_max_r")

    def save(self, stem) -> None:
        import json
        from pathlib import Path
        stem = Path(stem)
        arrays = {name: getattr(self, name) for name in self._ARRAYS if getattr(self, name) is not None}
        arrays["stats_df"] = torch.from_numpy(np.ascontiguousarray(self.stats.df, dtype=np.int64))
        arrays["stats_first_seen"] = torch.from_numpy(np.ascontiguousarray(self.stats.first_seen, dtype=np.int64))
        manifest, pos = {}, 0
        with open(stem.with_suffix(".bin"), "wb") as fh:
            for name, t in arrays.items():
                a = np.ascontiguousarray(t.detach().cpu().numpy())
                a = a.astype(a.dtype.newbyteorder("<"), copy=False)
                pad = (-pos) % 256
                fh.write(b"\0" * pad)
                pos += pad
                manifest[name] = {"dtype": a.dtype.str, "shape": list(a.shape), "offset": pos}
                fh.write(a.tobytes())
                pos += a.nbytes
        meta = {"format": "orag-bm25-index", "version": self.FORMAT_VERSION, "n_docs": self.n_docs, "vocab": self.vocab,
                "tile_docs": self.tile_docs, "n_tiles": self.n_tiles, "fp_tile_docs": self.fp_tile_docs,
                "fp_n_tiles": self.fp_n_tiles, "doc_id_base": self.doc_id_base, "n_postings": self.n_postings,
                "max_dl": self.max_dl, "has_negative_idf": self.has_negative_idf, "avgdl": float(self.avgdl).hex(),
                "average_idf": float(self.average_idf).hex(), "eps": float(self.eps).hex(),
                "stats_n_docs": self.stats.n_docs, "stats_total_len": self.stats.total_len, "bytes": pos,
                "arrays": manifest}
        stem.with_suffix(".json").write_text(json.dumps(meta))

    @classmethod
    def load(cls, stem, device="cuda") -> "Bm25Index":
        """Re-creates a saved index on `device` without touching the token corpus (no sort, no statistics pass)."""
        import json
        from pathlib import Path
        stem = Path(stem)
        meta = json.loads(stem.with_suffix(".json").read_text())
        if meta.get("format") != "orag-bm25-index" or meta.get("version") != cls.FORMAT_VERSION:
            raise ValueError(f"{stem}: not a version-{cls.FORMAT_VERSION} orag BM25 index")
        blob = np.memmap(stem.with_suffix(".bin"), dtype=np.uint8, mode="r")
        if blob.shape[0] != meta["bytes"]:
            raise ValueError(f"{stem}.bin: truncated ({blob.shape[0]} of {meta['bytes']} bytes)")
        dev = torch.device(device)

        def arr(name):
            m = meta["arrays"].get(name)
            if m is None:
                return None
            dt = np.dtype(m["dtype"])
            count = int(np.prod(m["shape"])) if m["shape"] else 1
            a = np.frombuffer(blob, dtype=dt, count=count, offset=m["offset"]).reshape(m["shape"])
            return np.array(a, dtype=dt.newbyteorder("="))  # private, native-endian copy

        self = cls.__new__(cls)
        self.device = dev
        for key in ("n_docs", "vocab", "tile_docs", "n_tiles", "fp_tile_docs", "fp_n_tiles", "doc_id_base",
                    "n_postings", "max_dl"):
            setattr(self, key, int(meta[key]))
        self.has_negative_idf = bool(meta["has_negative_idf"])
        self.avgdl, self.average_idf, self.eps = (float.fromhex(meta[k]) for k in ("avgdl", "average_idf", "eps"))
        self.stats = Bm25Stats(int(meta["stats_n_docs"]), int(meta["stats_total_len"]), arr("stats_df"),
                               arr("stats_first_seen"))
        for name in cls._ARRAYS:
            a = arr(name)
            setattr(self, name, None if a is None else torch.from_numpy(a).to(dev))
        ptr = lambda t: t.data_ptr() if t is not None else None
        # same field-by-field construction as __init__ (tests/test_host_logic.py compares the two structs)
        self.struct = _ffi.Bm25IndexStruct(
            n_docs=self.n_docs, vocab=self.vocab, tile_docs=self.tile_docs, n_tiles=self.n_tiles,
            has_negative_idf=int(self.has_negative_idf),
            d_tile_base=ptr(self.tile_base), d_tile_term_off=ptr(self.tile_term_off),
            max_doc_len=self.max_dl, reserved=0, d_postings=ptr(self.postings), d_doc_len=ptr(self.dl),
            d_t4_table=ptr(self.t4_table), d_r_table=ptr(self.r_table), d_idf=ptr(self.idf),
            d_postings_r16=ptr(self.postings_r16), d_term_max_r=ptr(self.term_max_r),
            fp_tile_docs=self.fp_tile_docs, fp_n_tiles=self.fp_n_tiles,
            d_fp_tile_base=ptr(self.fp_tile_base), d_fp_tile_term_off=ptr(self.fp_tile_term_off))
        self._ws = None
        return self
