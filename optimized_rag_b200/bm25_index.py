"""GPU-resident, doc-range-tiled BM25 index (ingest side).

The token corpus is turned into the layout of include/orag.h `orag_bm25_index_t` by the library's own builder
(csrc/bm25_build.cu: orag_bm25_index_plan / orag_bm25_index_fill); this module only owns the arrays, reduces the
statistics and evaluates the handful of scalars that must come out of CPython's `math.log` (idf):
  tile t = docs [t*T, (t+1)*T) of the shard
  postings      uint32 [(doc_in_tile << 16) | tf], grouped by (tile, term), ascending doc
  tile_base     int64 [n_tiles+1]   first posting of each tile
  tile_term_off int32 [n_tiles, V+1] offsets of each term's run inside its tile
  doc_len       int32 [n_docs]; t4_table float64 [max_dl+1] = k1*(1 - b + b*dl/avgdl) (global avgdl)
  idf           float64 [V]         global idf with the epsilon floor (rank_bm25 0.2.2 BM25Okapi._calc_idf)
  first-pass view (csrc/bm25_ms.cu), a second tiling with tiles of fp_tile_docs (up to 16384) docs:
  postings_r16  uint32 [(doc_in_tile << 16) | fp16(tf*(k1+1)/(tf + t4[dl]))], fp_tile_base, fp_tile_term_off; every run
                starts on a 16-byte boundary and is padded to four postings (copies of its last doc, impact 0)
  term_max_r    float32 [V]         max fp16 r over this shard's postings of each term (MaxScore upper bounds)

Global statistics (N, avgdl, df, first-seen order -> idf, eps) follow the reference's
`BM25Okapi(tokenized_corpus)` construction at rag/retrieval.py:338; they are computed over the WHOLE
corpus (all shards) so that sharded and single-GPU results are identical.
"""
from __future__ import annotations

import ctypes
import math
import os
from dataclasses import dataclass

import numpy as np
import torch

from . import _ffi

K1 = 1.5
B = 0.75
EPSILON = 0.25
INT64_MAX = np.iinfo(np.int64).max


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


def default_fp_tile_docs(n_docs: int) -> int:
    # measured on B200 (scripts/debug_bm25.py, 10M docs x 256 queries): 4096-doc tiles 1.27 ms, 8192 0.93 ms, 16384
    # 1.34 ms (the per-(query, tile) fixed cost halves with every doubling until the docs a pair marks outgrow the
    # per-warp accumulator); small corpora keep >= 16 tiles for the work hand-out
    return min(8192, max(32, 1 << max(0, (max(n_docs, 1) // 16 - 1).bit_length())))


class Bm25Plan:
    """Phase 1 of the index build (orag_bm25_index_plan): one counting pass over the shard's token corpus gives the
    doc lengths, the local df / first-seen statistics, and the final run offsets + tile bases of both tilings.
    `local_stats` is what a multi-GPU build all-reduces (dist.reduce_stats) before `Bm25Index.from_plan`."""

    def __init__(self, doc_off: torch.Tensor, tokens: torch.Tensor, vocab: int, tile_docs: int = 1024,
                 fp_tile_docs: int | None = None, token_pos_base: int = 0):
        assert doc_off.dtype == torch.int64 and tokens.dtype == torch.int32
        if not tokens.is_cuda or not doc_off.is_cuda:
            raise _ffi.OragError("the BM25 index is built on the GPU (csrc/bm25_build.cu): doc_off / tokens must be CUDA "
                                 "tensors -- there is no CPU builder")
        assert tile_docs & (tile_docs - 1) == 0 and 32 <= tile_docs <= 2048
        dev = tokens.device
        self.device, self.vocab, self.tile_docs = dev, int(vocab), int(tile_docs)
        self.doc_off, self.tokens = doc_off.contiguous(), tokens.contiguous()
        self.n_docs = int(doc_off.numel() - 1)
        self.fp_tile_docs = int(fp_tile_docs or os.environ.get("ORAG_FP_TILE_DOCS") or default_fp_tile_docs(self.n_docs))
        assert self.fp_tile_docs & (self.fp_tile_docs - 1) == 0 and 32 <= self.fp_tile_docs <= 16384
        self.n_tiles = (self.n_docs + self.tile_docs - 1) // self.tile_docs
        self.fp_n_tiles = (self.n_docs + self.fp_tile_docs - 1) // self.fp_tile_docs
        L = _ffi.lib()
        V1 = self.vocab + 1
        i32 = lambda *shape: torch.empty(shape, dtype=torch.int32, device=dev)
        i64 = lambda *shape: torch.empty(shape, dtype=torch.int64, device=dev)
        self.dl = i32(max(self.n_docs, 1))[:self.n_docs]
        self.tile_base, self.fp_tile_base = i64(self.n_tiles + 1), i64(self.fp_n_tiles + 1)
        self.tile_term_off, self.fp_tile_term_off = i32(self.n_tiles, V1), i32(self.fp_n_tiles, V1)
        df = torch.zeros(self.vocab, dtype=torch.int64, device=dev)
        first = torch.full((self.vocab,), INT64_MAX, dtype=torch.int64, device=dev)
        self.info = torch.zeros(4, dtype=torch.int32, device=dev)
        need = int(L.orag_bm25_build_workspace_bytes(self.n_docs, self.vocab, self.tile_docs, self.fp_tile_docs))
        self.ws = torch.empty(max(need, 256), dtype=torch.uint8, device=dev)
        totals = (ctypes.c_int64 * 2)()
        st = torch.cuda.current_stream(dev).cuda_stream
        _ffi.check(L.orag_bm25_index_plan(self.doc_off.data_ptr(), self.tokens.data_ptr(), self.n_docs, self.vocab,
                                          self.tile_docs, self.fp_tile_docs, int(token_pos_base), self.dl.data_ptr(),
                                          df.data_ptr(), first.data_ptr(), self.tile_base.data_ptr(),
                                          self.tile_term_off.data_ptr(), self.fp_tile_base.data_ptr(),
                                          self.fp_tile_term_off.data_ptr(), self.info.data_ptr(), self.ws.data_ptr(),
                                          self.ws.numel(), totals, st), "orag_bm25_index_plan")
        self.n_postings, self.n_postings_fp = int(totals[0]), int(totals[1])
        max_tf, max_dl, err, _ = self.info.tolist()
        if err & 1:
            raise ValueError("token ids must lie in [0, vocab)")
        if err & 2:
            raise ValueError("documents longer than 65535 tokens are not representable in the tile layout")
        if err & 4:
            raise _ffi.OragError("bm25 build: term de-duplication table overflow")
        if err & 8:
            raise ValueError("a tile holds 2^31 or more postings")
        if max_tf > 0xFFFF:
            raise ValueError("term frequency above 65535 is not representable in the posting format")
        self.max_dl = int(max_dl)
        total_len = int(doc_off[-1].item()) if self.n_docs > 0 else 0
        self.df_local = df.cpu().numpy()
        self.local_stats = Bm25Stats(self.n_docs, total_len, self.df_local, first.cpu().numpy())


def local_stats(doc_off: torch.Tensor, tokens: torch.Tensor, vocab: int, token_pos_base: int = 0) -> Bm25Stats:
    """df / first-seen / lengths of one shard (a by-product of `Bm25Plan`; callers that go on to build the index should
    create the plan themselves and keep it)."""
    return Bm25Plan(doc_off, tokens, vocab, token_pos_base=token_pos_base).local_stats


class Bm25Index:
    """One shard of the inverted index, resident on `device`."""

    def __init__(self, doc_off: torch.Tensor, tokens: torch.Tensor, vocab: int, tile_docs: int = 1024,
                 stats: Bm25Stats | None = None, doc_id_base: int = 0, first_pass: bool = True,
                 fp_tile_docs: int | None = None, plan: Bm25Plan | None = None):
        if plan is None:
            plan = Bm25Plan(doc_off, tokens, vocab, tile_docs, fp_tile_docs)
        self._from_plan(plan, stats or plan.local_stats, doc_id_base, first_pass)

    @classmethod
    def from_plan(cls, plan: Bm25Plan, stats: Bm25Stats | None = None, doc_id_base: int = 0,
                  first_pass: bool = True) -> "Bm25Index":
        self = cls.__new__(cls)
        self._from_plan(plan, stats or plan.local_stats, doc_id_base, first_pass)
        return self

    def _from_plan(self, plan: Bm25Plan, stats: Bm25Stats, doc_id_base: int, first_pass: bool):
        dev = plan.device
        L = _ffi.lib()
        self.device, self.vocab, self.tile_docs = dev, plan.vocab, plan.tile_docs
        self.n_docs, self.doc_id_base = plan.n_docs, int(doc_id_base)
        self.n_tiles, self.fp_tile_docs, self.fp_n_tiles = plan.n_tiles, plan.fp_tile_docs, plan.fp_n_tiles
        self.stats = stats
        self.df_local = plan.df_local
        idf, self.average_idf, self.eps = idf_table(stats)
        self.avgdl = stats.avgdl
        self.has_negative_idf = bool((idf < 0).any())
        self.idf = torch.from_numpy(idf).to(dev)
        self.dl, self.max_dl = plan.dl, plan.max_dl
        t4_np = t4_table(self.max_dl, self.avgdl)
        self.t4_table = torch.from_numpy(t4_np).to(dev)
        # r[dl, tf-1] = tf*(k1+1) / (tf + t4[dl]) for tf = 1..4, numpy float64 in rank_bm25's operation order
        tf_np = np.arange(1, 5)[None, :]
        self.r_table = torch.from_numpy(np.ascontiguousarray(tf_np * (K1 + 1) / (tf_np + t4_np[:, None]))).to(dev)
        self.tile_base, self.tile_term_off = plan.tile_base, plan.tile_term_off
        self.fp_tile_base, self.fp_tile_term_off = plan.fp_tile_base, plan.fp_tile_term_off
        self.n_postings, self.n_postings_fp = plan.n_postings, plan.n_postings_fp
        # (+4 entries: 16-byte reads of the last run never leave the allocation)
        self.postings = torch.empty(self.n_postings + 4, dtype=torch.int32, device=dev)
        self.postings[self.n_postings:].zero_()
        want_fp = first_pass and not self.has_negative_idf and self.n_docs > 0
        self.postings_r16 = torch.empty(self.n_postings_fp + 4, dtype=torch.int32, device=dev) if want_fp else None
        if want_fp:
            self.postings_r16[self.n_postings_fp:].zero_()
        self.term_max_r = torch.zeros(self.vocab, dtype=torch.float32, device=dev) if want_fp else None
        st = torch.cuda.current_stream(dev).cuda_stream
        _ffi.check(L.orag_bm25_index_fill(plan.doc_off.data_ptr(), plan.tokens.data_ptr(), self.n_docs, self.vocab,
                                          self.tile_docs, self.fp_tile_docs, self.t4_table.data_ptr(), self.max_dl,
                                          self.tile_base.data_ptr(), self.tile_term_off.data_ptr(),
                                          self.postings.data_ptr(), self.fp_tile_base.data_ptr(),
                                          self.fp_tile_term_off.data_ptr(),
                                          self.postings_r16.data_ptr() if want_fp else None,
                                          self.term_max_r.data_ptr() if want_fp else None, plan.info.data_ptr(),
                                          plan.ws.data_ptr(), plan.ws.numel(), st), "orag_bm25_index_fill")
        err = int(plan.info[2].item())
        if err & 4:
            raise _ffi.OragError("bm25 build: term de-duplication table overflow")
        if err & 16:  # an impact that is not a normal fp16 number: no first-pass view (the float64 kernel serves alone)
            self.postings_r16 = self.term_max_r = None
        if self.postings_r16 is None:
            self.fp_tile_base = self.fp_tile_term_off = None
        plan.ws = plan.doc_off = plan.tokens = None  # the cursors are spent: a plan fills one index
        self._make_struct()
        self._ws = self._lane_ws = None

    def _make_struct(self):
        ptr = lambda t: t.data_ptr() if t is not None else None
        self.term_kth_r = None
        self._fill_struct(ptr)
        if (self.postings_r16 is not None and self.postings_r16.is_cuda
                and os.environ.get("ORAG_BM25_WARM_START", "1") == "1"):
            # threshold warm start of the first pass: K-th largest impact of every term, derived from the first-pass
            # view itself (a few ms at 10M docs; recomputed on load rather than stored)
            self.term_kth_r = torch.empty((_ffi.ORAG_BM25_KTH_LEVELS, self.vocab), dtype=torch.float32, device=self.device)
            _ffi.check(_ffi.lib().orag_bm25_term_kth(ctypes.byref(self.struct), self.term_kth_r.data_ptr(),
                                                     torch.cuda.current_stream(self.device).cuda_stream),
                       "orag_bm25_term_kth")
            self._fill_struct(ptr)

    def _fill_struct(self, ptr):
        # reserved bit 0: first-pass runs are 16-byte aligned and padded to four postings (csrc/bm25_build.cu)
        self.struct = _ffi.Bm25IndexStruct(
            n_docs=self.n_docs, vocab=self.vocab, tile_docs=self.tile_docs, n_tiles=self.n_tiles,
            has_negative_idf=int(self.has_negative_idf), max_doc_len=self.max_dl,
            reserved=1 if self.postings_r16 is not None else 0,
            d_tile_base=ptr(self.tile_base), d_tile_term_off=ptr(self.tile_term_off), d_postings=ptr(self.postings),
            d_doc_len=ptr(self.dl), d_t4_table=ptr(self.t4_table), d_r_table=ptr(self.r_table), d_idf=ptr(self.idf),
            d_postings_r16=ptr(self.postings_r16), d_term_max_r=ptr(self.term_max_r),
            fp_tile_docs=self.fp_tile_docs, fp_n_tiles=self.fp_n_tiles,
            d_fp_tile_base=ptr(self.fp_tile_base), d_fp_tile_term_off=ptr(self.fp_tile_term_off),
            d_term_kth_r=ptr(self.term_kth_r))

    @property
    def doc_t4(self) -> torch.Tensor:
        return self.t4_table[self.dl.long()]

    # ------------------------------------------------------------------ queries
    def _workspace(self, n_queries: int, k: int, flags: int, lane: int = 0) -> torch.Tensor:
        """One workspace per lane (calls of different lanes may be in flight at the same time); `_ws` = lane 0's."""
        need = int(_ffi.lib().orag_bm25_workspace_bytes(ctypes.byref(self.struct), n_queries, k, flags))
        if self._lane_ws is None:
            self._lane_ws = {}
        ws = self._lane_ws.get(lane)
        if ws is None or ws.numel() < need:
            ws = self._lane_ws[lane] = torch.empty(max(need, 256), dtype=torch.uint8, device=self.device)
        if lane == 0:
            self._ws = ws
        return ws

    def topk(self, query_terms: torch.Tensor, query_lens: torch.Tensor, k: int, normalize: bool = True,
             force: str | None = None, check_overflow: bool = True, status_out: list | None = None,
             background: bool = False, lane: int = 0):
        """query_terms int32 [B, max_terms] (negative = OOV/padding), query_lens int32 [B].
        `status_out` (a list) receives the per-query status tensor so that a caller can defer the overflow
        check and pay one host sync for several calls (see engine.HybridShard.local_lists).
        Returns ids int64 [B,k], scores f64 [B,k], max f64 [B] (divisor if normalize else shard max raw)."""
        assert query_terms.dtype == torch.int32 and query_lens.dtype == torch.int32
        assert query_terms.is_cuda and query_terms.is_contiguous() and query_lens.is_contiguous()
        Bq, mt = query_terms.shape
        if mt > 64 and force != "dense":
            force = "dense"   # the candidate kernels hold <= 64 terms per query; the dense path takes any length
        flags = (_ffi.ORAG_BM25_NORMALIZE if normalize else 0)
        flags |= {None: 0, "sparse": _ffi.ORAG_BM25_FORCE_SPARSE, "dense": _ffi.ORAG_BM25_FORCE_DENSE,
                  "exact_tiles": _ffi.ORAG_BM25_FORCE_SPARSE | _ffi.ORAG_BM25_EXACT_TILES}[force]
        if background:
            flags |= _ffi.ORAG_BM25_BACKGROUND
        ids = torch.empty((Bq, k), dtype=torch.int64, device=self.device)
        sc = torch.empty((Bq, k), dtype=torch.float64, device=self.device)
        mx = torch.empty(Bq, dtype=torch.float64, device=self.device)
        status = torch.empty(Bq, dtype=torch.int32, device=self.device)
        ws = self._workspace(Bq, k, flags, lane)
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
                                       force="dense", check_overflow=False, lane=lane)
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

    def posting_bytes(self, query_terms, query_lens) -> int:
        """Algorithmic bytes of a batch: sum over query tokens of df_shard(t) * 6 (host arithmetic on
        the shard's document frequencies; a measurement helper, not part of the search)."""
        t = query_terms.cpu().numpy() if isinstance(query_terms, torch.Tensor) else np.asarray(query_terms)
        n = query_lens.cpu().numpy() if isinstance(query_lens, torch.Tensor) else np.asarray(query_lens)
        mask = (t >= 0) & (t < self.vocab) & (np.arange(t.shape[1])[None, :] < n[:, None])
        tc = np.clip(t, 0, self.vocab - 1)
        idf_nz = self.idf.cpu().numpy() != 0
        return int(np.where(mask & idf_nz[tc], self.df_local[tc], 0).sum()) * 6

    # ------------------------------------------------------------------ on-disk format
    # The reference's "format" for the keyword side is nothing at all: `BM25Okapi(tokenized_corpus)` is rebuilt from
    # the Postgres rows on every call (rag/retrieval.py:334-338).  Here one index = two files:
    #   <stem>.json   scalars (sizes, avgdl, eps, ...) + a manifest {name: dtype, shape, byte offset}
    #   <stem>.bin    the arrays of the layout described at the top of this module, little-endian, each aligned to
    #                 256 bytes, exactly as they sit in HBM -> loading is read (or mmap) + one copy per array
    FORMAT_VERSION = 2  # 2: first-pass runs 16-byte aligned and padded to four postings
    _ARRAYS = ("idf", "dl", "t4_table", "r_table", "postings", "tile_base", "tile_term_off",
               "postings_r16", "fp_tile_base", "fp_tile_term_off", "term_max_r")

    def save(self, stem) -> None:
        import json
        from pathlib import Path
        stem = Path(stem)
        arrays = {name: getattr(self, name) for name in self._ARRAYS if getattr(self, name) is not None}
        arrays["df_local"] = torch.from_numpy(np.ascontiguousarray(self.df_local, dtype=np.int64))
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
                "n_postings_fp": self.n_postings_fp, "max_dl": self.max_dl, "has_negative_idf": self.has_negative_idf, "avgdl": float(self.avgdl).hex(),
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
                    "n_postings", "n_postings_fp", "max_dl"):
            setattr(self, key, int(meta[key]))
        self.has_negative_idf = bool(meta["has_negative_idf"])
        self.avgdl, self.average_idf, self.eps = (float.fromhex(meta[k]) for k in ("avgdl", "average_idf", "eps"))
        self.stats = Bm25Stats(int(meta["stats_n_docs"]), int(meta["stats_total_len"]), arr("stats_df"),
                               arr("stats_first_seen"))
        for name in cls._ARRAYS:
            a = arr(name)
            setattr(self, name, None if a is None else torch.from_numpy(a).to(dev))
        self.df_local = arr("df_local")
        self._make_struct()
        self._ws = self._lane_ws = None
        return self
