"""Tensor-level Python surface over the C ABI (include/orag.h).

PyTorch is plumbing here: it owns device memory and streams; every computation on the retrieval
hot path is a call into csrc/liborag.so.  Nothing in this module falls back to torch or CPU math.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np
import torch

from . import _ffi
from .bm25_index import Bm25Index

MODE = {"exact": _ffi.ORAG_COS_EXACT, "tf32": _ffi.ORAG_COS_TF32, "bf16": _ffi.ORAG_COS_BF16, "f16": _ffi.ORAG_COS_F16}
SMALL_N = 4096  # below this the exact CUDA-core scan is used directly


def _stream(dev) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def _require_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise _ffi.OragError(f"{name} must live on a CUDA device (no CPU path exists)")


# --------------------------------------------------------------------------- synthetic fills
def gen_embeddings(n_rows: int, dim: int, row_start: int, seed: int, dup_per_mille: int = 0,
                   device="cuda", out: torch.Tensor | None = None) -> torch.Tensor:
    if out is None:
        out = torch.empty((n_rows, dim), dtype=torch.float32, device=device)
    _ffi.check(_ffi.lib().orag_gen_embeddings(out.data_ptr(), n_rows, dim, row_start, seed, dup_per_mille,
                                              _stream(out.device)), "orag_gen_embeddings")
    return out


def gen_token_corpus(n_docs: int, doc_start: int, seed: int, thresholds: np.ndarray, vocab: int, lmin: int = 100,
                     lmax: int = 300, device="cuda"):
    """(doc_off int64 [n+1], tokens int32 [total]) on `device`, bit-identical to synthetic.token_corpus."""
    dev = torch.device(device)
    lens = torch.empty(n_docs, dtype=torch.int32, device=dev)
    _ffi.check(_ffi.lib().orag_gen_doc_lengths(lens.data_ptr(), n_docs, doc_start, seed, lmin, lmax, _stream(dev)),
               "orag_gen_doc_lengths")
    doc_off = torch.zeros(n_docs + 1, dtype=torch.int64, device=dev)
    torch.cumsum(lens, dim=0, out=doc_off[1:])
    total = int(doc_off[-1].item())
    thr = torch.from_numpy(thresholds.view(np.int64)).to(dev)
    tokens = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
    _ffi.check(_ffi.lib().orag_gen_tokens(tokens.data_ptr(), doc_off.data_ptr(), n_docs, doc_start, seed,
                                          thr.data_ptr(), vocab, _stream(dev)), "orag_gen_tokens")
    return doc_off, tokens[:total]


# --------------------------------------------------------------------------- cosine
class CosineIndex:
    """One shard of chunk embeddings resident in HBM: fp32 rows (+ fp32 inverse norms, + optional 16-bit shadow copy
    for the first pass: "f16" = IEEE fp16 rows scaled by powers of two (preferred), "bf16").  `topk` = exact float64
    cosine top-k (see orag_cosine_topk)."""

    def __init__(self, corpus: torch.Tensor, row_id_base: int = 0, mode: str = "auto", shadow: bool | None = None):
        _require_cuda(corpus, "corpus")
        assert corpus.dtype == torch.float32 and corpus.dim() == 2 and corpus.is_contiguous()
        self.corpus = corpus
        self.device = corpus.device
        self.n_rows, self.dim = corpus.shape
        self.row_id_base = int(row_id_base)
        if mode == "auto":
            if self.n_rows < SMALL_N or self.dim % 32 != 0:
                mode = "exact"
            else:
                mode = "f16" if (shadow and self.dim % 64 == 0) else "tf32"
        self.mode = mode
        self.inv_norm = None
        self.shadow = None
        if mode in ("tf32", "bf16"):
            self.inv_norm = torch.empty(self.n_rows, dtype=torch.float32, device=self.device)
            _ffi.check(_ffi.lib().orag_row_inv_norms(corpus.data_ptr(), self.n_rows, self.dim,
                                                     self.inv_norm.data_ptr(), _stream(self.device)),
                       "orag_row_inv_norms")
        if mode == "bf16":
            self.shadow = torch.empty((self.n_rows, self.dim), dtype=torch.bfloat16, device=self.device)
            _ffi.check(_ffi.lib().orag_f32_to_bf16(corpus.data_ptr(), self.shadow.data_ptr(),
                                                   self.n_rows * self.dim, _stream(self.device)), "orag_f32_to_bf16")
        if mode == "f16":
            # fp16 rows scaled by a per-row power of two; inv_norm carries the same scale
            self.shadow = torch.empty((self.n_rows, self.dim), dtype=torch.float16, device=self.device)
            self.inv_norm = torch.empty(self.n_rows, dtype=torch.float32, device=self.device)
            _ffi.check(_ffi.lib().orag_f32_to_f16_rows(corpus.data_ptr(), self.n_rows, self.dim, self.shadow.data_ptr(),
                                                       self.inv_norm.data_ptr(), None, _stream(self.device)),
                       "orag_f32_to_f16_rows")
        self.row_sq = None
        if mode != "exact":
            # sum(a*a) of every row in the reference's float64 arithmetic: a per-row constant of the final re-score
            self.row_sq = torch.empty(self.n_rows, dtype=torch.float64, device=self.device)
            _ffi.check(_ffi.lib().orag_row_sq(corpus.data_ptr(), self.n_rows, self.dim, self.row_sq.data_ptr(),
                                              _stream(self.device)), "orag_row_sq")
        self._ws = {}

    @classmethod
    def from_arrays(cls, corpus: torch.Tensor, inv_norm: torch.Tensor | None, shadow: torch.Tensor | None,
                    row_sq: torch.Tensor | None, mode: str, row_id_base: int = 0) -> "CosineIndex":
        """A view over arrays the caller maintains itself (e.g. row by row as chunks arrive: `derive_rows`) -- nothing
        is recomputed here.  `mode` "exact" needs only the corpus; "f16" / "bf16" the 16-bit shadow + scaled inverse
        norms + row_sq; "tf32" the inverse norms + row_sq."""
        _require_cuda(corpus, "corpus")
        assert corpus.dtype == torch.float32 and corpus.dim() == 2 and corpus.is_contiguous() and mode in MODE
        self = cls.__new__(cls)
        self.corpus, self.device = corpus, corpus.device
        self.n_rows, self.dim = corpus.shape
        self.row_id_base, self.mode = int(row_id_base), mode
        self.inv_norm, self.shadow, self.row_sq = inv_norm, shadow, row_sq
        if mode != "exact":
            assert inv_norm is not None and row_sq is not None and inv_norm.shape[0] == self.n_rows
            assert mode == "tf32" or (shadow is not None and shadow.shape == corpus.shape)
        self._ws = {}
        return self

    @staticmethod
    def derive_rows(corpus: torch.Tensor, shadow_f16: torch.Tensor, inv_norm: torch.Tensor, row_sq: torch.Tensor):
        """Fill the derived arrays of the "f16" mode for the given rows (all four tensors are row-aligned slices):
        fp16 shadow scaled by powers of two + its inverse norms (orag_f32_to_f16_rows) and the float64 sum(a*a)."""
        n, dim = corpus.shape
        if n == 0:
            return
        st = _stream(corpus.device)
        _ffi.check(_ffi.lib().orag_f32_to_f16_rows(corpus.data_ptr(), n, dim, shadow_f16.data_ptr(), inv_norm.data_ptr(),
                                                   None, st), "orag_f32_to_f16_rows")
        _ffi.check(_ffi.lib().orag_row_sq(corpus.data_ptr(), n, dim, row_sq.data_ptr(), st), "orag_row_sq")

    def _workspace(self, n_queries: int, k: int, mode: int, lane: int = 0) -> torch.Tensor:
        """One workspace per (mode, lane): calls of different lanes may be in flight at the same time (see
        dist.ShardedHybrid.submit), a workspace belongs to one call until its work has finished."""
        need = int(_ffi.lib().orag_cosine_workspace_bytes(self.n_rows, self.dim, n_queries, k, mode))
        ws = self._ws.get((mode, lane))
        if ws is None or ws.numel() < need:
            ws = torch.empty(max(need, 256), dtype=torch.uint8, device=self.device)
            self._ws[(mode, lane)] = ws
        return ws

    def topk(self, queries: torch.Tensor, k: int, mode: str | None = None, check_overflow: bool = True,
             status_out: list | None = None, lane: int = 0):
        """queries fp32 [B, dim] on the same device -> (ids int64 [B,k], scores float64 [B,k]).
        `status_out` (a list) receives the per-query status tensor for a deferred overflow check."""
        _require_cuda(queries, "queries")
        assert queries.dtype == torch.float32 and queries.is_contiguous() and queries.shape[1] == self.dim
        m = MODE[mode or self.mode]
        if m != _ffi.ORAG_COS_EXACT and self.inv_norm is None:
            raise _ffi.OragError("index was built for the exact path only")
        if m != _ffi.ORAG_COS_EXACT and m != MODE[self.mode]:
            raise _ffi.OragError(f"index was built for the {self.mode} first pass")
        Bq = queries.shape[0]
        ids = torch.empty((Bq, k), dtype=torch.int64, device=self.device)
        sc = torch.empty((Bq, k), dtype=torch.float64, device=self.device)
        status = torch.empty(Bq, dtype=torch.int32, device=self.device)
        self._call(queries, k, m, ids, sc, status, lane, _ffi.ORAG_PHASE_ALL)
        if status_out is not None:
            status_out.append(status)
        if check_overflow and m != _ffi.ORAG_COS_EXACT:
            bad = torch.nonzero(status != 0).flatten()
            if bad.numel():
                # more near-ties than candidate slots: exact scan for these queries (still on the GPU)
                i2, s2 = self.topk(queries[bad].contiguous(), k, mode="exact", check_overflow=False, lane=lane)
                ids[bad], sc[bad] = i2, s2
        return ids, sc

    def _call(self, queries, k, m, ids, sc, status, lane, phases):
        Bq = queries.shape[0]
        ws = self._workspace(Bq, k, m, lane)
        _ffi.check(_ffi.lib().orag_cosine_topk_phase(
            self.corpus.data_ptr(), self.inv_norm.data_ptr() if self.inv_norm is not None else None,
            self.shadow.data_ptr() if self.shadow is not None else None,
            self.row_sq.data_ptr() if (self.row_sq is not None and m != _ffi.ORAG_COS_EXACT) else None,
            self.n_rows, self.dim, self.row_id_base, queries.data_ptr(), Bq, k, m, ids.data_ptr(), sc.data_ptr(),
            status.data_ptr(), ws.data_ptr(), ws.numel(), phases, _stream(self.device)), "orag_cosine_topk_phase")

    def topk_scan(self, queries: torch.Tensor, k: int, lane: int = 0, stream: "torch.cuda.Stream | None" = None):
        """First half of `topk` (tensor-core modes, <= 256 queries; orag_cosine_topk_phase): query preparation on the
        current stream, then seed pass and main scan on `stream` (default: the current one), ordered after everything
        the current stream holds; the candidates stay in the lane's workspace.  The output tensors are allocated here, on
        the CURRENT stream -- the one `topk_finish` will run on.  Returns the handle `topk_finish` takes; the caller
        orders that stream after `stream`."""
        _require_cuda(queries, "queries")
        assert queries.dtype == torch.float32 and queries.is_contiguous() and queries.shape[1] == self.dim
        m = MODE[self.mode]
        if m == _ffi.ORAG_COS_EXACT or queries.shape[0] > 256:
            raise _ffi.OragError("topk_scan: tensor-core modes and at most 256 queries")
        Bq = queries.shape[0]
        ids = torch.empty((Bq, k), dtype=torch.int64, device=self.device)
        sc = torch.empty((Bq, k), dtype=torch.float64, device=self.device)
        status = torch.empty(Bq, dtype=torch.int32, device=self.device)
        # query conversion / norms on the CURRENT stream: they run as soon as the inputs are there, next to whatever scan
        # occupies the SMs, not in the gap between two scans
        self._call(queries, k, m, ids, sc, status, lane, _ffi.ORAG_PHASE_PREP)
        if stream is None:
            self._call(queries, k, m, ids, sc, status, lane, _ffi.ORAG_PHASE_SCAN)
        else:
            stream.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(stream):
                self._call(queries, k, m, ids, sc, status, lane, _ffi.ORAG_PHASE_SCAN)
        return queries, k, m, ids, sc, status, lane

    def last_counts(self, n_queries: int, k: int, lane: int = 0):
        """(first-pass candidates, fp32 survivors) per query of the last tensor-core search of `lane` (diagnostic)."""
        import numpy as np
        ws = self._workspace(n_queries, k, MODE[self.mode], lane)
        c = np.zeros(n_queries, dtype=np.uint32)
        v = np.zeros(n_queries, dtype=np.uint32)
        _ffi.check(_ffi.lib().orag_cosine_last_counts(ws.data_ptr(), self.dim, n_queries, c.ctypes.data, v.ctypes.data,
                                                      _stream(self.device)), "orag_cosine_last_counts")
        return c, v

    def topk_finish(self, handle):
        """Second half: candidate re-scores and selection on the current stream -> (ids, scores, status)."""
        queries, k, m, ids, sc, status, lane = handle
        self._call(queries, k, m, ids, sc, status, lane, _ffi.ORAG_PHASE_FINISH)
        return ids, sc, status

    def dense(self, queries: torch.Tensor) -> torch.Tensor:
        """float64 cosine matrix [B, n_rows] (reference arithmetic; tests, small corpora, weighted hybrid)."""
        Bq = queries.shape[0]
        buf = torch.empty(Bq * self.n_rows + Bq, dtype=torch.float64, device=self.device)
        _ffi.check(_ffi.lib().orag_cosine_dense(self.corpus.data_ptr(), self.n_rows, self.dim, queries.data_ptr(), Bq,
                                                buf.data_ptr(), _stream(self.device)), "orag_cosine_dense")
        return buf[:Bq * self.n_rows].view(Bq, self.n_rows)

    def dots(self, queries: torch.Tensor):
        """(dots float64 [B, n_rows], row_sq float64 [n_rows], query_sq float64 [B]): the reference's three sums without
        the final sqrt / divide (orag_dot_dense)."""
        Bq = queries.shape[0]
        dots = torch.empty((Bq, max(self.n_rows, 1)), dtype=torch.float64, device=self.device)[:, :self.n_rows].contiguous()
        row_sq = torch.empty(max(self.n_rows, 1), dtype=torch.float64, device=self.device)[:self.n_rows]
        q_sq = torch.empty(Bq, dtype=torch.float64, device=self.device)
        _ffi.check(_ffi.lib().orag_dot_dense(self.corpus.data_ptr(), self.n_rows, self.dim, queries.data_ptr(), Bq,
                                             dots.data_ptr(), row_sq.data_ptr(), q_sq.data_ptr(), _stream(self.device)),
                   "orag_dot_dense")
        return dots, row_sq, q_sq

    def firstpass_dense(self, queries: torch.Tensor, mode: str) -> torch.Tensor:
        """Raw tensor-core first-pass values (dot * inv_norm[row]) fp32 [n_rows, B] -- test hook."""
        Bq = queries.shape[0]
        rows_pad = (self.n_rows + 127) // 128 * 128
        out = torch.zeros((rows_pad, 256), dtype=torch.float32, device=self.device)
        ws = torch.empty(Bq * self.dim * 2 + 4096, dtype=torch.uint8, device=self.device)
        _ffi.check(_ffi.lib().orag_cosine_firstpass_dense(
            self.corpus.data_ptr(), self.inv_norm.data_ptr(),
            self.shadow.data_ptr() if self.shadow is not None else None, self.n_rows, self.dim, queries.data_ptr(),
            Bq, MODE[mode], out.data_ptr(), ws.data_ptr(), ws.numel(), _stream(self.device)),
            "orag_cosine_firstpass_dense")
        return out[:self.n_rows, :Bq]

    def stream_bytes(self, n_queries: int, mode: str | None = None) -> int:
        """Bytes of corpus the first pass streams per batch of <= 256 queries."""
        m = mode or self.mode
        per = 2 if m in ("bf16", "f16") else 4
        groups = (n_queries + 255) // 256
        return self.n_rows * self.dim * per * groups


# --------------------------------------------------------------------------- selection / fusion
def dense_topk(scores: torch.Tensor, k: int, id_base: int = 0, normalize: bool = False):
    """scores float64 [B, n] -> ids, scores, max (exact (score desc, id asc) selection)."""
    _require_cuda(scores, "scores")
    assert scores.dtype == torch.float64 and scores.dim() == 2 and scores.stride(1) == 1
    Bq, n = scores.shape
    ids = torch.empty((Bq, k), dtype=torch.int64, device=scores.device)
    sc = torch.empty((Bq, k), dtype=torch.float64, device=scores.device)
    mx = torch.ones(Bq, dtype=torch.float64, device=scores.device)
    _ffi.check(_ffi.lib().orag_dense_topk(scores.data_ptr(), n, scores.stride(0) if n else 0, Bq, k, id_base,
                                          int(normalize), ids.data_ptr(), sc.data_ptr(), mx.data_ptr(),
                                          _stream(scores.device)), "orag_dense_topk")
    return ids, sc, mx


def topk_merge(cand_ids: torch.Tensor, cand_scores: torch.Tensor, k: int, shard_max: torch.Tensor | None = None):
    """cand_* [B, m] gathered from all shards -> global top-k.  With shard_max [B, G] (raw BM25 maxima) the
    scores are normalised by the global max first (rag/retrieval.py:343-345)."""
    _require_cuda(cand_ids, "cand_ids")
    assert cand_ids.dtype == torch.int64 and cand_scores.dtype == torch.float64
    assert cand_ids.is_contiguous() and cand_scores.is_contiguous()
    Bq, m = cand_ids.shape
    ids = torch.empty((Bq, k), dtype=torch.int64, device=cand_ids.device)
    sc = torch.empty((Bq, k), dtype=torch.float64, device=cand_ids.device)
    mx = torch.ones(Bq, dtype=torch.float64, device=cand_ids.device)
    g = 0
    if shard_max is not None:
        assert shard_max.dtype == torch.float64 and shard_max.is_contiguous()
        g = shard_max.shape[1]
    _ffi.check(_ffi.lib().orag_topk_merge(cand_ids.data_ptr(), cand_scores.data_ptr(), m, Bq, k,
                                          shard_max.data_ptr() if shard_max is not None else None, g,
                                          ids.data_ptr(), sc.data_ptr(), mx.data_ptr(), _stream(cand_ids.device)),
               "orag_topk_merge")
    return ids, sc, mx


def rrf_fuse(list_ids: torch.Tensor, rrf_k: int = 60, top_k: int = 10, tie: str = "reference",
             want_src: bool = False):
    """list_ids int64 [B, L, len] (-1 tail padding) -> fused ids [B, top_k], rrf scores, (src ranks [B,top_k,L])."""
    _require_cuda(list_ids, "list_ids")
    assert list_ids.dtype == torch.int64 and list_ids.dim() == 3 and list_ids.is_contiguous()
    Bq, L, n = list_ids.shape
    ids = torch.empty((Bq, top_k), dtype=torch.int64, device=list_ids.device)
    sc = torch.empty((Bq, top_k), dtype=torch.float64, device=list_ids.device)
    src = torch.empty((Bq, top_k, L), dtype=torch.int32, device=list_ids.device) if want_src else None
    _ffi.check(_ffi.lib().orag_rrf_fuse(list_ids.data_ptr(), Bq, L, n, rrf_k, top_k, 0 if tie == "reference" else 1,
                                        ids.data_ptr(), sc.data_ptr(), src.data_ptr() if want_src else None,
                                        _stream(list_ids.device)), "orag_rrf_fuse")
    return (ids, sc, src) if want_src else (ids, sc)


def rrf_fuse_pair(ids_a: torch.Tensor, ids_b: torch.Tensor, rrf_k: int = 60, top_k: int = 10,
                  status_a: torch.Tensor | None = None, status_b: torch.Tensor | None = None, tie: str = "reference"):
    """RRF of two ranked lists held in separate [B, len] tensors (no packing launch) -> fused ids, scores, src ranks
    [B, top_k, 2], status int32 [B] = status_a | status_b."""
    _require_cuda(ids_a, "ids_a")
    assert ids_a.dtype == torch.int64 and ids_b.dtype == torch.int64 and ids_a.shape == ids_b.shape
    assert ids_a.is_contiguous() and ids_b.is_contiguous()
    Bq, n = ids_a.shape
    dev = ids_a.device
    ids = torch.empty((Bq, top_k), dtype=torch.int64, device=dev)
    sc = torch.empty((Bq, top_k), dtype=torch.float64, device=dev)
    src = torch.empty((Bq, top_k, 2), dtype=torch.int32, device=dev)
    status = torch.empty(Bq, dtype=torch.int32, device=dev)
    _ffi.check(_ffi.lib().orag_rrf_fuse_pair(ids_a.data_ptr(), ids_b.data_ptr(), Bq, n, rrf_k, top_k,
                                             0 if tie == "reference" else 1,
                                             status_a.data_ptr() if status_a is not None else None,
                                             status_b.data_ptr() if status_b is not None else None, ids.data_ptr(),
                                             sc.data_ptr(), src.data_ptr(), status.data_ptr(), _stream(dev)),
               "orag_rrf_fuse_pair")
    return ids, sc, src, status


def weighted_sum3(sem: torch.Tensor, kw: torch.Tensor, temp: torch.Tensor | None, alpha: float, beta: float,
                  gamma: float) -> torch.Tensor:
    """(alpha*sem + beta*kw) + gamma*temp in float64 without contraction (rag/retrieval.py:302)."""
    _require_cuda(sem, "sem")
    assert sem.dtype == torch.float64 and kw.dtype == torch.float64 and sem.is_contiguous() and kw.is_contiguous()
    out = torch.empty_like(sem)
    _ffi.check(_ffi.lib().orag_weighted_sum3(sem.data_ptr(), kw.data_ptr(),
                                             temp.data_ptr() if temp is not None else None, sem.numel(), alpha, beta,
                                             gamma, out.data_ptr(), _stream(sem.device)), "orag_weighted_sum3")
    return out


def div_scalar(x: torch.Tensor, divisor: float) -> torch.Tensor:
    """x / divisor in float64, one IEEE division per element (rag/retrieval.py:345)."""
    _require_cuda(x, "x")
    assert x.dtype == torch.float64 and x.is_contiguous()
    out = torch.empty_like(x)
    _ffi.check(_ffi.lib().orag_div_scalar(x.data_ptr(), x.numel(), divisor, out.data_ptr(), _stream(x.device)),
               "orag_div_scalar")
    return out


def pairwise_cosine_threshold(emb: torch.Tensor, doc_idx: torch.Tensor, threshold: float = 0.85,
                              cap: int = 1 << 20, mode: str = "auto"):
    """All i<j, doc_idx differ, float64 cosine >= threshold (rag/consistency_checker.py:169-189).
    Returns (i, j, sim) sorted by (i, j).  mode: "exact" = float64 CUDA-core sweep of every pair,
    "tc" = tcgen05 first pass + float64 re-score (same pair set), "auto" = tc for large inputs."""
    _require_cuda(emb, "emb")
    assert emb.dtype == torch.float32 and emb.is_contiguous() and doc_idx.dtype == torch.int32
    m, dim = emb.shape
    dev = emb.device
    if mode == "auto":
        mode = "tc" if (m >= 2048 and dim % 32 == 0 and threshold > 0.01) else "exact"
    oi = torch.empty(cap, dtype=torch.int32, device=dev)
    oj = torch.empty(cap, dtype=torch.int32, device=dev)
    osim = torch.empty(cap, dtype=torch.float64, device=dev)
    cnt = torch.zeros(2, dtype=torch.int64, device=dev)
    L = _ffi.lib()
    if mode == "tc":
        need = int(L.orag_pairwise_tc_workspace_bytes(m, dim))
        ws = torch.empty(max(need, 256), dtype=torch.uint8, device=dev)
        _ffi.check(L.orag_pairwise_cosine_threshold_tc(emb.data_ptr(), m, dim, doc_idx.data_ptr(), threshold, cap,
                                                       oi.data_ptr(), oj.data_ptr(), osim.data_ptr(), cnt.data_ptr(),
                                                       ws.data_ptr(), ws.numel(), _stream(dev)),
                   "orag_pairwise_cosine_threshold_tc")
        if int(cnt[1].item()) != 0:  # a row had more first-pass candidates than slots: exact sweep instead
            return pairwise_cosine_threshold(emb, doc_idx, threshold, cap, mode="exact")
    else:
        need = int(L.orag_pairwise_workspace_bytes(m, dim))
        ws = torch.empty(max(need, 256), dtype=torch.uint8, device=dev)
        _ffi.check(L.orag_pairwise_cosine_threshold(emb.data_ptr(), m, dim, doc_idx.data_ptr(), threshold, cap,
                                                    oi.data_ptr(), oj.data_ptr(), osim.data_ptr(), cnt.data_ptr(),
                                                    ws.data_ptr(), ws.numel(), _stream(dev)),
                   "orag_pairwise_cosine_threshold")
    n = int(cnt[0].item())
    if n > cap:
        raise _ffi.OragError(f"pair capacity exceeded ({n} > {cap})")
    key = oi[:n].long() * m + oj[:n].long()
    order = torch.argsort(key)
    return oi[:n][order], oj[:n][order], osim[:n][order]


class PairwiseIndex:
    """A claim-embedding matrix prepared for the all-pairs search of `ConsistencyChecker._find_contradictions`
    (rag/consistency_checker.py:163-189; BASELINE config 5): fp32 rows + doc_idx, and -- derived once at construction by
    orag_pairwise_prepare -- the fp16 shadow (rows scaled by powers of two), first-pass norms and float64 sum(a*a).
    `pairs(threshold)` = every i < j with doc_idx[i] != doc_idx[j] and float64 cosine >= threshold."""

    def __init__(self, emb: torch.Tensor, doc_idx: torch.Tensor):
        _require_cuda(emb, "emb")
        assert emb.dtype == torch.float32 and emb.dim() == 2 and emb.is_contiguous()
        assert doc_idx.dtype == torch.int32 and doc_idx.is_contiguous() and doc_idx.shape[0] == emb.shape[0]
        self.emb, self.doc_idx = emb, doc_idx
        self.m, self.dim = emb.shape
        self.device = emb.device
        self.tensor_cores = self.m >= 2048 and self.dim % 32 == 0
        self._ws = None
        if self.tensor_cores:
            L = _ffi.lib()
            self.prepared = torch.empty(max(int(L.orag_pairwise_prepared_bytes(self.m, self.dim)), 256), dtype=torch.uint8,
                                        device=self.device)
            _ffi.check(L.orag_pairwise_prepare(emb.data_ptr(), self.m, self.dim, self.prepared.data_ptr(),
                                               self.prepared.numel(), _stream(self.device)), "orag_pairwise_prepare")

    def pairs(self, threshold: float = 0.85, cap: int = 1 << 20, sync: bool = True):
        """(i int32, j int32, sim float64) sorted by (i, j).  With sync=False nothing synchronises with the host: the raw
        output buffers and the device-side count are returned instead ((i, j, sim, count[2]); count[1] != 0 = a per-row
        candidate buffer overflowed and the caller must use the exact sweep)."""
        if not self.tensor_cores or threshold <= 0.01:
            assert sync, "the exact sweep is the small-input path"
            return pairwise_cosine_threshold(self.emb, self.doc_idx, threshold, cap, mode="exact")
        L = _ffi.lib()
        dev = self.device
        if self._ws is None:
            self._ws = torch.empty(max(int(L.orag_pairwise_pairs_workspace_bytes(self.m)), 256), dtype=torch.uint8, device=dev)
        oi = torch.empty(cap, dtype=torch.int32, device=dev)
        oj = torch.empty(cap, dtype=torch.int32, device=dev)
        osim = torch.empty(cap, dtype=torch.float64, device=dev)
        cnt = torch.empty(2, dtype=torch.int64, device=dev)
        _ffi.check(L.orag_pairwise_pairs(self.emb.data_ptr(), self.prepared.data_ptr(), self.m, self.dim,
                                         self.doc_idx.data_ptr(), threshold, cap, oi.data_ptr(), oj.data_ptr(),
                                         osim.data_ptr(), cnt.data_ptr(), self._ws.data_ptr(), self._ws.numel(),
                                         _stream(dev)), "orag_pairwise_pairs")
        if not sync:
            return oi, oj, osim, cnt
        n, overflow = (int(x) for x in cnt.tolist())
        if overflow:  # a row had more first-pass candidates than slots: exact sweep instead
            return pairwise_cosine_threshold(self.emb, self.doc_idx, threshold, cap, mode="exact")
        if n > cap:
            raise _ffi.OragError(f"pair capacity exceeded ({n} > {cap})")
        order = torch.argsort(oi[:n].long() * self.m + oj[:n].long())
        return oi[:n][order], oj[:n][order], osim[:n][order]


# --------------------------------------------------------------------------- hybrid (one shard)
class HybridShard:
    """Cosine list + BM25 list -> RRF for the rows/docs of one GPU (SURVEY.md 'three facts' #1:
    the composition the README describes, each piece in the reference's own arithmetic)."""

    def __init__(self, cosine: CosineIndex, bm25: Bm25Index, rrf_k: int = 60):
        assert cosine.n_rows == bm25.n_docs and cosine.row_id_base == bm25.doc_id_base
        self.cosine = cosine
        self.bm25 = bm25
        self.rrf_k = rrf_k
        self._side: dict = {}   # lane -> side stream of the BM25 pipeline
        self._head = None       # the stream every scan of this shard is enqueued on (co-scheduled searches)
        self.head_stream = os.environ.get("ORAG_HEAD_STREAM", "0") == "1"
        self.serial = False     # no side stream at all (measurement: see local_lists)
        self._marked = False
        # BM25 first pass NEXT TO the scan (see local_lists): ORAG_COSCHEDULE=0/1 forces it; by default only for large
        # shards.  Measured on B200, three batches in flight (bench.py --rows R, ms per batch, on / off): 10M rows 8.11 /
        # 8.33, 5M 4.12 / 4.12, 2.5M 2.15 / 2.12, 1.25M 1.21 / 1.16 -- next to a short scan the background CTAs mostly
        # keep the tail kernels of the previous batch off the SMs.
        env = os.environ.get("ORAG_COSCHEDULE")
        self.coschedule = (env == "1") if env in ("0", "1") else cosine.n_rows >= 6_000_000

    def local_lists(self, query_emb: torch.Tensor, query_terms: torch.Tensor, query_lens: torch.Tensor, fetch_k: int,
                    bm25_k: int, normalize: bool, lane: int = 0):
        """Cosine top-fetch_k and BM25 top-bm25_k of this shard, enqueued WITHOUT any host synchronisation.
        The BM25 pipeline runs on a side stream next to the cosine pipeline, so the small latency-bound kernels
        of either (query norms, candidate re-score, selection, finalize) overlap the other's main kernel.
        Returns (cos ids, cos scores, bm25 ids, bm25 scores, bm25 max, cosine status, BM25 status): the two int32 [B]
        candidate-overflow words are OR-ed by the kernel that consumes the lists (orag_rrf_fuse_pair / orag_hybrid_push);
        callers check the result once, at the end of the whole step (`repair`)."""
        dev = query_emb.device
        cur = torch.cuda.current_stream(dev)
        L = _ffi.lib()
        if lane not in self._side:
            self._side[lane] = torch.cuda.Stream(dev)
        if self.coschedule and not self._marked:
            L.orag_cosine_mark_prescan(1)   # (process-wide; the cosine calls record one more event each from now on)
            self._marked = True
        side = self._side[lane]
        # inputs are ready; also what makes the caching allocator's per-stream pools safe without record_stream():
        # every side-stream block (BM25 outputs, status) is only ever re-issued to side-stream work of a LATER call,
        # which this wait orders after everything the current stream has enqueued on those blocks
        side.wait_stream(cur)
        st_c: list = []
        st_b: list = []
        if self.coschedule and self.cosine.mode != "exact":
            # the scan takes the SMs first; the BM25 first pass (8-warp CTAs, < 31 KB smem) then runs NEXT TO the
            # resident scan CTA of every SM (see __init__ for when that pays: long scans only).
            # head_stream (off by default -- measured: no gain, the tails cannot run next to a resident scan either
            # way): all scans back to back on ONE high-priority stream (`_head`) through the split entry point; the
            # tails -- candidate re-scores, selection, and whatever the caller appends: fusion, exchange, merge -- stay
            # on the caller's (lane) stream.
            if query_emb.shape[0] <= 256 and self.head_stream:
                if self._head is None:
                    self._head = torch.cuda.Stream(dev, priority=-1)
                # (topk_scan orders _head after cur: inputs ready, the lane's workspace free -- tails of its last batch)
                handle = self.cosine.topk_scan(query_emb, fetch_k, lane=lane, stream=self._head)
                scanned = torch.cuda.Event()
                scanned.record(self._head)
            else:
                handle = None
                ci, cs = self.cosine.topk(query_emb, fetch_k, check_overflow=False, status_out=st_c, lane=lane)
            with torch.cuda.stream(side):
                # (ORAG_BM25_BACKGROUND: the first-pass launch waits for the pre-scan event recorded just above)
                bi, bs, bmax = self.bm25.topk(query_terms, query_lens, bm25_k, normalize=normalize,
                                              check_overflow=False, status_out=st_b, background=True, lane=lane)
            if handle is not None:
                cur.wait_event(scanned)
                ci, cs, st = self.cosine.topk_finish(handle)
                st_c.append(st)
        else:
            # (`serial`: both pipelines on the caller's stream, one after the other -- what bench.py brackets the two
            # dominant kernels in, each with the whole GPU)
            with torch.cuda.stream(cur if self.serial else side):
                bi, bs, bmax = self.bm25.topk(query_terms, query_lens, bm25_k, normalize=normalize,
                                              check_overflow=False, status_out=st_b, lane=lane)
            ci, cs = self.cosine.topk(query_emb, fetch_k, check_overflow=False, status_out=st_c, lane=lane)
        cur.wait_stream(side)
        return ci, cs, bi, bs, bmax, st_c[0], st_b[0]

    def exact_lists(self, query_emb: torch.Tensor, query_terms: torch.Tensor, query_lens: torch.Tensor, fetch_k: int,
                    bm25_k: int, normalize: bool):
        """The same lists through the exhaustive kernels (exact float64 scan, dense BM25): the repair path for
        queries whose candidate buffers overflowed (thousands of duplicate rows / near-ties).  Still on the GPU."""
        ci, cs = self.cosine.topk(query_emb, fetch_k, mode="exact", check_overflow=False)
        bi, bs, bmax = self.bm25.topk(query_terms, query_lens, bm25_k, normalize, force="dense", check_overflow=False)
        return ci, cs, bi, bs, bmax

    def search(self, query_emb: torch.Tensor, query_terms: torch.Tensor, query_lens: torch.Tensor, k: int = 10,
               fetch_k: int | None = None, check_overflow: bool = True, lane: int = 0):
        fetch_k = fetch_k or k
        ci, cs, bi, bs, bmax, st_c, st_b = self.local_lists(query_emb, query_terms, query_lens, fetch_k, fetch_k, True,
                                                            lane=lane)
        fi, fs, src, status = rrf_fuse_pair(ci, bi, self.rrf_k, k, st_c, st_b)
        out = {"ids": fi, "rrf_scores": fs, "scores": fs, "src_ranks": src, "cos_ids": ci, "cos_scores": cs, "bm25_ids": bi,
               "bm25_scores": bs, "bm25_max": bmax, "status": status}
        # one host round trip per step, after everything has been enqueued (the caller reads the result anyway).
        # With check_overflow=False nothing synchronises: a caller that pipelines batches checks out["status"]
        # (non-zero = that query's lists are invalid and must be repaired) whenever it reads the results.
        if check_overflow and bool(status.any()):
            bad = torch.nonzero(status).flatten()
            ci2, cs2, bi2, bs2, bm2 = self.exact_lists(query_emb[bad].contiguous(), query_terms[bad].contiguous(),
                                                       query_lens[bad].contiguous(), fetch_k, fetch_k, True)
            f2, s2, r2 = rrf_fuse(torch.stack([ci2, bi2], dim=1).contiguous(), self.rrf_k, k, want_src=True)
            for key, val in (("ids", f2), ("rrf_scores", s2), ("src_ranks", r2), ("cos_ids", ci2), ("cos_scores", cs2),  # "scores" aliases "rrf_scores"
                             ("bm25_ids", bi2), ("bm25_scores", bs2), ("bm25_max", bm2)):
                out[key][bad] = val
            out["status"] = torch.zeros_like(status)
        return out
