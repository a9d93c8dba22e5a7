"""Row-sharded hybrid retrieval across the GPUs of one box.

One process per GPU (torch.distributed, NCCL over NVLink/NVSwitch).  Rank r owns the contiguous
chunk-id range [r*N/G, (r+1)*N/G): its fp32 rows (+ shadow), and the BM25 postings of the same
docs.  Queries are replicated.  Per batch the only exchange is ONE all-gather of each rank's
local winners (cosine top-k, BM25 raw top-(k+guard) and the shard's max raw BM25 score), packed
into a single int64 buffer (~B * (4k + 2*guard + 1) * 8 bytes per rank, tens of KB: latency-bound);
every rank then runs the G*k -> k merges and RRF itself, so all ranks hold the result.

BM25 statistics (N, avgdl, df, first-seen order -> idf, eps) are GLOBAL: `sharded_stats` all-reduces
them once at index build so that sharded results equal the single-GPU / oracle results bit for bit.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import _ffi, engine
from .bm25_index import Bm25Stats, local_stats

BM25_GUARD = 6  # extra raw-score entries per shard: distinct raw scores that collapse to one normalised double


def shard_range(n_total: int, rank: int, world: int):
    """Contiguous, balanced split; boundaries are multiples of nothing in particular (the kernels mask tails)."""
    lo = (n_total * rank) // world
    hi = (n_total * (rank + 1)) // world
    return lo, hi


def sharded_stats(doc_off: torch.Tensor, tokens: torch.Tensor, vocab: int, group=None) -> Bm25Stats:
    """Global Bm25Stats from each rank's local docs (ranks hold consecutive doc ranges in rank order)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local_stats(doc_off, tokens, vocab)
    rank = dist.get_rank(group)
    dev = tokens.device
    n_local = doc_off.numel() - 1
    total_local = int(doc_off[-1].item()) if n_local > 0 else 0
    sizes = torch.zeros(world, 2, dtype=torch.int64, device=dev)
    mine = torch.tensor([n_local, total_local], dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(sizes.view(-1), mine, group=group)
    token_base = int(sizes[:rank, 1].sum().item())
    st = local_stats(doc_off, tokens, vocab, token_pos_base=token_base)
    df = torch.from_numpy(st.df).to(dev)
    first = torch.from_numpy(st.first_seen).to(dev)
    dist.all_reduce(df, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(first, op=dist.ReduceOp.MIN, group=group)
    return Bm25Stats(int(sizes[:, 0].sum().item()), int(sizes[:, 1].sum().item()), df.cpu().numpy(),
                     first.cpu().numpy())


def pack_local(cos_ids, cos_scores, bm_ids, bm_scores, bm_max, status):
    """[B, W] int64 buffer: cosine ids | cosine score bits | bm25 ids | bm25 raw score bits | bm25 max bits |
    status bits (the layout orag_hybrid_merge reads, include/orag.h)."""
    return torch.cat([cos_ids, cos_scores.view(torch.int64), bm_ids, bm_scores.view(torch.int64),
                      bm_max.view(torch.int64)[:, None], status.long()[:, None]], dim=1).contiguous()


def unpack_gathered(buf: torch.Tensor, fetch_k: int, kk: int):
    """Host-side decoder of the gathered layout (tests / debugging; the product path reads the buffer directly
    in orag_hybrid_merge): buf [G, B, W] -> cosine ids/scores [B, G*fetch_k], bm25 ids/raw [B, G*kk],
    shard max [B, G], status [B, G]."""
    G, Bq, W = buf.shape
    assert W == 2 * fetch_k + 2 * kk + 2
    b = buf.permute(1, 0, 2)  # [B, G, W]
    o = 0
    ci = b[:, :, o:o + fetch_k].reshape(Bq, G * fetch_k).contiguous(); o += fetch_k
    cs = b[:, :, o:o + fetch_k].reshape(Bq, G * fetch_k).contiguous().view(torch.float64); o += fetch_k
    bi = b[:, :, o:o + kk].reshape(Bq, G * kk).contiguous(); o += kk
    bs = b[:, :, o:o + kk].reshape(Bq, G * kk).contiguous().view(torch.float64); o += kk
    bm = b[:, :, o].contiguous().view(torch.float64)
    st = b[:, :, o + 1].contiguous()
    return ci, cs, bi, bs, bm, st


def hybrid_merge(gathered: torch.Tensor, fetch_k: int, kk: int, rrf_k: int, k: int):
    """gathered int64 [G, B, W] -> the result dict of a hybrid search (one launch: csrc/rrf.cu
    hybrid_merge_kernel) plus the OR of the shards' overflow flags."""
    G, Bq, W = gathered.shape
    assert W == 2 * fetch_k + 2 * kk + 2 and gathered.is_contiguous()
    dev = gathered.device
    i64 = lambda *shape: torch.empty(shape, dtype=torch.int64, device=dev)
    f64 = lambda *shape: torch.empty(shape, dtype=torch.float64, device=dev)
    fi, fs, src = i64(Bq, k), f64(Bq, k), torch.empty((Bq, k, 2), dtype=torch.int32, device=dev)
    ci, cs, bi, bs, bmax = i64(Bq, fetch_k), f64(Bq, fetch_k), i64(Bq, fetch_k), f64(Bq, fetch_k), f64(Bq)
    status = torch.empty(Bq, dtype=torch.int32, device=dev)
    _ffi.check(_ffi.lib().orag_hybrid_merge(gathered.data_ptr(), G, Bq, fetch_k, kk, rrf_k, k, 0, fi.data_ptr(),
                                            fs.data_ptr(), src.data_ptr(), ci.data_ptr(), cs.data_ptr(), bi.data_ptr(),
                                            bs.data_ptr(), bmax.data_ptr(), status.data_ptr(),
                                            torch.cuda.current_stream(dev).cuda_stream), "orag_hybrid_merge")
    return {"ids": fi, "rrf_scores": fs, "src_ranks": src, "cos_ids": ci, "cos_scores": cs, "bm25_ids": bi,
            "bm25_scores": bs, "bm25_max": bmax}, status


class ShardedHybrid:
    """Hybrid (cosine + BM25 -> RRF) search over a corpus row-sharded across the ranks of `group`."""

    def __init__(self, shard: engine.HybridShard, group=None):
        self.shard = shard
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._gather_buf = None

    def _exchange(self, mine: torch.Tensor, fetch_k: int, kk: int, k: int):
        Bq, W = mine.shape
        if self._gather_buf is None or self._gather_buf.shape != (self.world, Bq, W):
            self._gather_buf = torch.empty((self.world, Bq, W), dtype=torch.int64, device=mine.device)
        dist.all_gather_into_tensor(self._gather_buf.view(-1), mine.view(-1), group=self.group)
        return hybrid_merge(self._gather_buf, fetch_k, kk, self.shard.rrf_k, k)

    def search(self, query_emb, query_terms, query_lens, k: int = 10, fetch_k: int | None = None,
               check_overflow: bool = True):
        """Per batch: local lists (no host sync) -> ONE all-gather of the packed winners -> ONE merge+RRF launch.
        Candidate-buffer overflow on any rank is seen by every rank in the gathered status column; the affected
        queries (rare: thousands of duplicates / near-ties) are then repaired by all ranks together through the
        exhaustive kernels and a second, small exchange -- every rank takes the same branch."""
        fetch_k = fetch_k or k
        if self.world == 1:
            return self.shard.search(query_emb, query_terms, query_lens, k, fetch_k, check_overflow)
        kk = fetch_k + BM25_GUARD
        ci, cs, bi, bs, bm, st = self.shard.local_lists(query_emb, query_terms, query_lens, fetch_k, kk, False)
        out, status = self._exchange(pack_local(ci, cs, bi, bs, bm, st), fetch_k, kk, k)
        out["status"] = status  # check_overflow=False: no host sync at all; the caller checks it with the results
        if check_overflow and bool(status.any()):
            bad = torch.nonzero(status).flatten()
            lists = self.shard.exact_lists(query_emb[bad].contiguous(), query_terms[bad].contiguous(),
                                           query_lens[bad].contiguous(), fetch_k, kk, False)
            zero = torch.zeros(bad.numel(), dtype=torch.int32, device=bad.device)
            buf = self._gather_buf
            self._gather_buf = None
            fixed, _ = self._exchange(pack_local(*lists, zero), fetch_k, kk, k)
            self._gather_buf = buf
            for key, val in fixed.items():
                out[key][bad] = val
            out["status"] = torch.zeros_like(status)
        return out
