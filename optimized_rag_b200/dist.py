"""Row-sharded hybrid retrieval across the GPUs of one box.

One process per GPU (torch.distributed, NCCL over NVLink/NVSwitch).  Rank r owns the contiguous
chunk-id range [r*N/G, (r+1)*N/G): its fp32 rows (+ shadow), and the BM25 postings of the same
docs.  Queries are replicated.  Per batch the only exchange is ONE all-gather of each rank's
local winners (cosine top-k, BM25 raw top-(k+guard) and the shard's max raw BM25 score), packed
into a single int64 buffer (~B * (4k + 2*guard + 1) * 8 bytes per rank, tens of KB: latency-bound);
every rank then runs the G*k -> k merges and RRF itself, so all ranks hold the result.
Two transports for that exchange: NCCL (`pack_local` + `all_gather_into_tensor`) and `PeerExchange`
(csrc/exchange.cu): one kernel packs the winners and stores them straight into every peer's buffer over
NVLink, one tiny kernel waits for the peers' sequence numbers -- no packing kernels, no collective call.

BM25 statistics (N, avgdl, df, first-seen order -> idf, eps) are GLOBAL: `sharded_stats` all-reduces
them once at index build so that sharded results equal the single-GPU / oracle results bit for bit.
"""
from __future__ import annotations

import ctypes
import logging
import os

import numpy as np
import torch
import torch.distributed as dist

from . import _ffi, engine
from .bm25_index import Bm25Plan, Bm25Stats

logger = logging.getLogger(__name__)

# Extra raw-score entries per shard beyond fetch_k.  Shards rank their docs by RAW score (the divisor -- the global max --
# is only known after the exchange) and x -> x / max can fold neighbouring raw values into one normalised double, inside
# which the order is by id: a shard list that ends in such a group may have cut off a doc it should have kept.  The
# guard makes that rare; it is NOT what correctness rests on: orag_hybrid_merge detects every list that ends inside
# the k-th value's group with mixed raw values and flags the query, and the repair path (`_exact_lists_global_max`)
# exchanges the maxima first and ranks by the normalised value on every shard.
BM25_GUARD = 6
# Searches in flight per ShardedHybrid (`submit`): lane = sequence number % LANES, each lane with its own stream and
# workspaces.  The tails of the batches cannot run while a scan and the background BM25 CTA fill the SMs, so the steady
# state is LANES scans back to back, then the tails of all of them together: three lanes amortise that phase better than
# two (1.25M rows per GPU: 1.24 -> 1.21 ms per batch).  The peer exchange has 2 x LANES slots (csrc/exchange.cu).
LANES = 3


def shard_range(n_total: int, rank: int, world: int):
    """Contiguous, balanced split; boundaries are multiples of nothing in particular (the kernels mask tails)."""
    lo = (n_total * rank) // world
    hi = (n_total * (rank + 1)) // world
    return lo, hi


def _coll_device(group, device):
    """Tensors of a collective live where the backend wants them (NCCL: the GPU; gloo: the host)."""
    return device if dist.get_backend(group) == "nccl" else torch.device("cpu")


def token_position_base(n_local: int, total_local: int, group=None, device=None):
    """(global position of this rank's first token, docs of all ranks, tokens of all ranks): ranks hold consecutive doc
    ranges in rank order, so a term's first-seen position is comparable across shards."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return 0, n_local, total_local
    rank = dist.get_rank(group)
    dev = _coll_device(group, device)
    sizes = torch.zeros(world, 2, dtype=torch.int64, device=dev)
    mine = torch.tensor([n_local, total_local], dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(sizes.view(-1), mine, group=group)
    sizes = sizes.cpu()
    return int(sizes[:rank, 1].sum()), int(sizes[:, 0].sum()), int(sizes[:, 1].sum())


def reduce_stats(local: Bm25Stats, n_total: int, total_len: int, group=None, device=None) -> Bm25Stats:
    """Global Bm25Stats from every rank's local ones (df summed, first-seen positions -- already global -- min-ed)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    dev = _coll_device(group, device)
    df = torch.from_numpy(np.ascontiguousarray(local.df)).to(dev)
    first = torch.from_numpy(np.ascontiguousarray(local.first_seen)).to(dev)
    dist.all_reduce(df, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(first, op=dist.ReduceOp.MIN, group=group)
    return Bm25Stats(n_total, total_len, df.cpu().numpy(), first.cpu().numpy())


def sharded_plan(doc_off: torch.Tensor, tokens: torch.Tensor, vocab: int, tile_docs: int = 1024,
                 fp_tile_docs: int | None = None, group=None):
    """(Bm25Plan of this rank's docs, GLOBAL Bm25Stats): the counting pass of the index build runs once per shard and
    its by-products (df, first-seen positions) are all-reduced; `Bm25Index.from_plan(plan, stats, doc_id_base)` then
    fills the shard's index with the global idf / avgdl."""
    n_local = doc_off.numel() - 1
    total_local = int(doc_off[-1].item()) if n_local > 0 else 0
    base, n_total, total_len = token_position_base(n_local, total_local, group, tokens.device)
    plan = Bm25Plan(doc_off, tokens, vocab, tile_docs, fp_tile_docs, token_pos_base=base)
    return plan, reduce_stats(plan.local_stats, n_total, total_len, group, tokens.device)


def sharded_stats(doc_off: torch.Tensor, tokens: torch.Tensor, vocab: int, group=None) -> Bm25Stats:
    """Global statistics only (callers that go on to build the index should use `sharded_plan` and keep the plan)."""
    return sharded_plan(doc_off, tokens, vocab, group=group)[1]


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


def hybrid_merge(gathered, fetch_k: int, kk: int, rrf_k: int, k: int, shape=None, device=None):
    """gathered int64 [G, B, W] (a tensor, or a raw device pointer with `shape` and `device`) -> the result dict
    of a hybrid search (one launch: csrc/rrf.cu hybrid_merge_kernel) plus the OR of the shards' status flags."""
    if isinstance(gathered, torch.Tensor):
        G, Bq, W = gathered.shape
        assert gathered.is_contiguous()
        dev, ptr = gathered.device, gathered.data_ptr()
    else:
        (G, Bq, W), dev, ptr = shape, device, int(gathered)
    assert W == 2 * fetch_k + 2 * kk + 2
    i64 = lambda *shape: torch.empty(shape, dtype=torch.int64, device=dev)
    f64 = lambda *shape: torch.empty(shape, dtype=torch.float64, device=dev)
    fi, fs, src = i64(Bq, k), f64(Bq, k), torch.empty((Bq, k, 2), dtype=torch.int32, device=dev)
    ci, cs, bi, bs, bmax = i64(Bq, fetch_k), f64(Bq, fetch_k), i64(Bq, fetch_k), f64(Bq, fetch_k), f64(Bq)
    status = torch.empty(Bq, dtype=torch.int32, device=dev)
    _ffi.check(_ffi.lib().orag_hybrid_merge(ptr, G, Bq, fetch_k, kk, rrf_k, k, 0, fi.data_ptr(),
                                            fs.data_ptr(), src.data_ptr(), ci.data_ptr(), cs.data_ptr(), bi.data_ptr(),
                                            bs.data_ptr(), bmax.data_ptr(), status.data_ptr(),
                                            torch.cuda.current_stream(dev).cuda_stream), "orag_hybrid_merge")
    return {"ids": fi, "rrf_scores": fs, "scores": fs, "src_ranks": src, "cos_ids": ci, "cos_scores": cs, "bm25_ids": bi,
            "bm25_scores": bs, "bm25_max": bmax}, status


class PeerExchange:
    """Exchange buffers of all ranks of `group`, mapped into this process through CUDA IPC (csrc/exchange.cu).

    Collective constructor (every rank, same arguments).  The 64-byte IPC handles travel through
    `all_gather_object` once; after that an exchange is two kernel launches and no host communication.
    Every rank must call `exchange` the same number of times with the same shapes."""

    TIMEOUT_MS = int(os.environ.get("ORAG_EXCHANGE_TIMEOUT_MS", "30000"))

    def __init__(self, device: torch.device, max_queries: int, fetch_k: int, kk: int, group=None):
        L = _ffi.lib()
        self.device = torch.device(device)
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.max_queries, self.fetch_k, self.kk = int(max_queries), int(fetch_k), int(kk)
        self.W = 2 * fetch_k + 2 * kk + 2
        self.bytes = int(L.orag_exchange_bytes(self.world, self.max_queries, fetch_k, kk))
        self._mine, self._opened, self.seq = None, [], 0
        error = None
        with torch.cuda.device(self.device):
            handle = ctypes.create_string_buffer(64)
            try:
                mine = ctypes.c_void_p()
                _ffi.check(L.orag_exchange_alloc(self.bytes, ctypes.byref(mine)), "orag_exchange_alloc")
                self._mine = mine.value
                _ffi.check(L.orag_exchange_export(self._mine, handle), "orag_exchange_export")
            except _ffi.OragError as e:
                error = e
            handles = [None] * self.world
            dist.all_gather_object(handles, None if error else handle.raw, group=group)
            ptrs = []
            if error is None and all(h is not None for h in handles):
                try:
                    for r, h in enumerate(handles):
                        if r == self.rank:
                            ptrs.append(self._mine)
                            continue
                        p = ctypes.c_void_p()
                        _ffi.check(L.orag_exchange_open(h, ctypes.byref(p)), "orag_exchange_open")
                        self._opened.append(p.value)
                        ptrs.append(p.value)
                except _ffi.OragError as e:
                    error = e
            torch.cuda.synchronize(self.device)
            # agreement + barrier in one collective: every buffer is zeroed and mapped everywhere before the first
            # push, or every rank gives up together
            on_dev = dist.get_backend(group) == "nccl"
            ok = torch.tensor([0 if (error or len(ptrs) != self.world) else 1], dtype=torch.int32,
                              device=self.device if on_dev else "cpu")
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if int(ok.item()) == 0:
                for p in self._opened:
                    L.orag_exchange_close(p)
                dist.barrier(group=group)
                if self._mine is not None:
                    L.orag_exchange_free(self._mine)
                self._mine, self._opened = None, []
                raise _ffi.OragError(f"peer exchange setup failed on at least one rank (this rank: {error or 'ok'})")
        self.peer_ptrs = ptrs
        self._d_peers = torch.tensor(ptrs, dtype=torch.int64, device=self.device)

    def fits(self, n_queries: int, fetch_k: int, kk: int) -> bool:
        return n_queries <= self.max_queries and fetch_k == self.fetch_k and kk == self.kk

    def push(self, cos_ids, cos_scores, bm_ids, bm_scores, bm_max, status, status2=None, seq: int | None = None) -> int:
        """Pack this rank's lists and store them into every peer's slot of search `seq` (default: one more than the
        last; same on every rank, strictly increasing), then publish the sequence number.  Returns seq."""
        L = _ffi.lib()
        Bq = cos_ids.shape[0]
        assert self.fits(Bq, cos_ids.shape[1], bm_ids.shape[1])
        for t in (cos_ids, cos_scores, bm_ids, bm_scores, bm_max, status):
            assert t.is_contiguous() and t.device == self.device
        assert status.dtype == torch.int32 and bm_max.dtype == torch.float64
        assert status2 is None or (status2.dtype == torch.int32 and status2.is_contiguous())
        assert seq is None or seq > self.seq
        self.seq = self.seq + 1 if seq is None else int(seq)
        st = torch.cuda.current_stream(self.device).cuda_stream
        _ffi.check(L.orag_hybrid_push(cos_ids.data_ptr(), cos_scores.data_ptr(), bm_ids.data_ptr(), bm_scores.data_ptr(),
                                      bm_max.data_ptr(), status.data_ptr(),
                                      status2.data_ptr() if status2 is not None else None, Bq, self.fetch_k, self.kk, self.rank,
                                      self.world, self.max_queries, self._d_peers.data_ptr(), self.seq, st),
                   "orag_hybrid_push")
        return self.seq

    def wait(self, n_queries: int, seq: int | None = None):
        """Acquire every rank's block of search `seq` (default: the last pushed) -> (device pointer of [G, B, W], shape)."""
        out = ctypes.c_void_p()
        st = torch.cuda.current_stream(self.device).cuda_stream
        _ffi.check(_ffi.lib().orag_hybrid_wait(self._mine, self.world, self.max_queries, n_queries, self.fetch_k, self.kk,
                                               self.seq if seq is None else int(seq), self.TIMEOUT_MS, ctypes.byref(out),
                                               st), "orag_hybrid_wait")
        return out.value, (self.world, n_queries, self.W)

    def exchange(self, cos_ids, cos_scores, bm_ids, bm_scores, bm_max, status, status2=None, seq: int | None = None):
        """push + wait: this rank's lists to every peer, theirs back -> (device pointer of [G, B, W], shape).  `status`
        and the optional `status2` (the cosine and the BM25 call's overflow words) are OR-ed by the push kernel."""
        self.push(cos_ids, cos_scores, bm_ids, bm_scores, bm_max, status, status2, seq)
        return self.wait(cos_ids.shape[0])

    def close(self):
        """Collective: unmap the peers' buffers, then free the own one (after everybody has unmapped it).  Never
        raises between the two barriers (a rank that bailed out would leave the others waiting): failures are logged."""
        if self._mine is None:
            return
        L = _ffi.lib()
        errors = []
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)
        with torch.cuda.device(self.device):
            for p in self._opened:
                if L.orag_exchange_close(p) != 0:
                    errors.append(L.orag_last_error())
            dist.barrier(group=self.group)
            if L.orag_exchange_free(self._mine) != 0:
                errors.append(L.orag_last_error())
        self._mine, self._opened = None, []
        if errors:
            logger.warning("peer exchange teardown: %s", errors)


class ShardedHybrid:
    """Hybrid (cosine + BM25 -> RRF) search over a corpus row-sharded across the ranks of `group`.
    `exchange`: "peer" (PeerExchange: pack + NVLink peer stores in one kernel) or "nccl" (pack + all-gather);
    default from ORAG_EXCHANGE, else "peer".  If the peer buffers cannot be set up (no CUDA IPC / peer access between
    the ranks' devices) ALL ranks switch to "nccl" together and say so in `exchange_note` -- both feed the same merge
    kernel with the same bytes."""

    def __init__(self, shard: engine.HybridShard, group=None, exchange: str | None = None):
        self.shard = shard
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.exchange = exchange or os.environ.get("ORAG_EXCHANGE", "peer")
        assert self.exchange in ("nccl", "peer")
        self.exchange_note = ""
        self._gather_buf: dict = {}   # lane -> NCCL all-gather buffer
        self._peer: PeerExchange | None = None      # the exchange of the most recent search
        self._peers: dict = {}                      # (fetch_k, kk) -> PeerExchange
        self._poisoned = ""
        self._retired: list[PeerExchange] = []  # outgrown exchanges: peers may still have them mapped until close()
        self._searches = 0            # searches issued so far (all ranks count alike): lane = parity of the next one
        self._lane_streams: dict = {}
        self.lanes = LANES

    def _exchange(self, lists, fetch_k: int, kk: int, k: int, lane: int = 0):
        ci, cs, bi, bs, bm, st, st2 = lists
        if self.exchange == "peer":
            Bq = ci.shape[0]
            peer = self._peers.get((fetch_k, kk))
            if peer is None or not peer.fits(Bq, fetch_k, kk):
                # collective (every rank sees the same shapes).  One exchange per (fetch_k, kk), kept for the life of the
                # object: a serving loop that alternates top_k re-uses them instead of allocating and IPC-mapping a new
                # buffer per switch; only a larger batch replaces one (the outgrown buffer stays mapped until close()).
                if peer is not None:
                    self._retired.append(peer)
                try:
                    peer = self._peers[(fetch_k, kk)] = PeerExchange(ci.device, max(Bq, 256), fetch_k, kk, group=self.group)
                except _ffi.OragError as e:  # raised on every rank together (see PeerExchange.__init__)
                    logger.warning("peer exchange unavailable, using the NCCL all-gather: %s", e)
                    self.exchange, self.exchange_note = "nccl", f"peer setup failed: {e}"
            self._peer = peer if self.exchange == "peer" else None
        if self.exchange == "peer":
            ptr, shape = self._peer.exchange(ci, cs, bi, bs, bm, st, st2, seq=self._searches)
            return hybrid_merge(ptr, fetch_k, kk, self.shard.rrf_k, k, shape=shape, device=ci.device)
        mine = pack_local(ci, cs, bi, bs, bm, st if st2 is None else st | st2)
        Bq, W = mine.shape
        buf = self._gather_buf.get(lane)
        if buf is None or buf.shape != (self.world, Bq, W):
            buf = self._gather_buf[lane] = torch.empty((self.world, Bq, W), dtype=torch.int64, device=mine.device)
        dist.all_gather_into_tensor(buf.view(-1), mine.view(-1), group=self.group)
        return hybrid_merge(buf, fetch_k, kk, self.shard.rrf_k, k)

    def close(self):
        """Collective: unmap / free the peer exchange buffers (call on every rank before the process group goes away,
        so that no rank frees a buffer another one still has mapped)."""
        for px in self._retired + list(self._peers.values()):
            px.close()
        self._retired, self._peer, self._peers = [], None, {}

    def search(self, query_emb, query_terms, query_lens, k: int = 10, fetch_k: int | None = None,
               check_overflow: bool = True, lane: int | None = None):
        """Per batch: local lists (no host sync) -> ONE exchange of the packed winners -> ONE merge+RRF launch.
        Candidate-buffer overflow on any rank is seen by every rank in the gathered status column; the affected
        queries (rare: thousands of duplicates / near-ties) are then repaired by all ranks together through the
        exhaustive kernels and a second, small exchange -- every rank takes the same branch."""
        fetch_k = fetch_k or k
        if self._poisoned:
            raise _ffi.OragError(f"sharded search is shut down: {self._poisoned}")
        if lane is None:  # the lane of a search follows from its exchange sequence number (PeerExchange slots)
            lane = (self._searches + 1) % LANES
        self._searches += 1
        if self.world == 1:
            return self.shard.search(query_emb, query_terms, query_lens, k, fetch_k, check_overflow, lane=lane)
        kk = fetch_k + BM25_GUARD
        if self.world * kk > 256 or fetch_k > 64:
            raise _ffi.OragError(f"sharded search: n_shards * (fetch_k + {BM25_GUARD}) must be <= 256 and fetch_k <= 64 "
                                 f"(orag_hybrid_merge); got {self.world} shards, fetch_k {fetch_k}")
        lists = self.shard.local_lists(query_emb, query_terms, query_lens, fetch_k, kk, False, lane=lane)
        out, status = self._exchange(lists, fetch_k, kk, k, lane)
        if os.environ.get("ORAG_TEST_FORCE_REPAIR") == "1":   # test hook (scripts/check_dist.py): every query repaired
            status = status | _ffi.ORAG_STATUS_OVERFLOW
        out["status"] = status  # check_overflow=False: no host sync at all; the caller checks it with the results
        if check_overflow:
            # A timeout is only visible to the rank that waited in vain: agree on it (one tiny all-reduce; this checked
            # path synchronises with the host anyway) so that every rank takes the same branch.  After a timeout the
            # slot re-use invariant of the exchange no longer holds: the object is shut down on all ranks together.
            # Callers that pass check_overflow=False (batches in flight) must look at out["status"] themselves before the
            # next search: bit 2 set = stop using this object.
            flags = torch.stack([(status & _ffi.ORAG_STATUS_EXCHANGE_TIMEOUT).any(), status.any()]).to(torch.int32)
            dist.all_reduce(flags, op=dist.ReduceOp.MAX, group=self.group)
            timed_out, any_bad = (bool(x) for x in flags.tolist())
            if timed_out:
                self._poisoned = "a peer's block did not arrive within the exchange timeout"
                raise _ffi.OragError("sharded search: " + self._poisoned)
        if check_overflow and any_bad:
            bad = torch.nonzero(status).flatten()
            lists, gmax = self._exact_lists_global_max(query_emb[bad].contiguous(), query_terms[bad].contiguous(),
                                                       query_lens[bad].contiguous(), fetch_k, kk)
            zero = torch.zeros(bad.numel(), dtype=torch.int32, device=bad.device)
            # the repair exchange takes a sequence number (and an all-gather buffer) of its own; it is a synchronous
            # path: callers that keep batches in flight (`submit`) drain them before repairing
            self._searches += 1
            fixed, _ = self._exchange((*lists, zero, None), fetch_k, kk, k, lane=LANES)
            fixed["bm25_max"] = gmax
            for key, val in fixed.items():
                out[key][bad] = val
            out["status"] = torch.zeros_like(status)
        return out

    def _exact_lists_global_max(self, query_emb, query_terms, query_lens, fetch_k: int, kk: int):
        """The repair path's lists: exhaustive kernels, and -- unlike the fast path, which ranks a shard's docs by RAW
        score because the divisor is only known after the exchange -- the BM25 divisor FIRST (an all-reduce of the
        shards' maxima), then every shard ranks its docs by (score / global max desc, id asc) exactly as the
        single-GPU path does.  No truncation hazard remains: the merge of per-shard exact top lists is exact.
        Returns ((cos ids, cos scores, bm25 ids, NORMALISED bm25 scores, ones), global max)."""
        shard = self.shard
        ci, cs = shard.cosine.topk(query_emb, fetch_k, mode="exact", check_overflow=False)
        ix = shard.bm25
        nb = query_terms.shape[0]
        dev = query_terms.device
        bi = torch.empty((nb, kk), dtype=torch.int64, device=dev)
        bs = torch.empty((nb, kk), dtype=torch.float64, device=dev)
        gmax = torch.empty(nb, dtype=torch.float64, device=dev)
        ids_row = torch.arange(ix.n_docs, dtype=torch.int64, device=dev) + ix.doc_id_base
        step = max(1, (256 << 20) // max(8 * ix.n_docs, 1))
        for lo in range(0, nb, step):
            hi = min(nb, lo + step)
            raw = ix.dense_scores(query_terms[lo:hi].contiguous(), query_lens[lo:hi].contiguous())
            m = raw.amax(dim=1).clamp_min(0.0) if ix.n_docs else torch.zeros(hi - lo, dtype=torch.float64, device=dev)
            dist.all_reduce(m, op=dist.ReduceOp.MAX, group=self.group)
            gmax[lo:hi] = torch.where(m > 0, m, torch.ones_like(m))
            i2, s2, _ = engine.topk_merge(ids_row.expand(hi - lo, -1).contiguous(), raw.contiguous(), kk,
                                          shard_max=m[:, None].contiguous())
            bi[lo:hi], bs[lo:hi] = i2, s2
        ones = torch.ones(nb, dtype=torch.float64, device=dev)
        return (ci, cs, bi, bs, ones), gmax

    # ------------------------------------------------------------------ two batches in flight
    def submit(self, query_emb, query_terms, query_lens, k: int = 10, fetch_k: int | None = None) -> "Ticket":
        """Enqueue one search WITHOUT waiting for the one before it: consecutive submissions alternate between two
        lanes (streams with their own workspaces and exchange slots), so the latency-bound tail of batch i -- candidate
        re-scores, selection, exchange, merge -- runs under the scan of batch i+1.  Nothing synchronises with the host.
        `Ticket.wait()` orders the caller's stream after the search and hands out the result dict (with the per-query
        `status` word: non-zero = that query must be repeated through `search(check_overflow=True)`).  Every rank
        must submit the same sequence of batches."""
        dev = query_emb.device
        if query_emb.shape[0] > 256:
            # one search per 256-query group (the tensor-core scan's block), each on the next lane: the groups of one
            # call overlap like separate batches instead of running scan -> tail -> scan -> tail on one stream
            # (batch 1024 on one GPU, 10M rows: 38.7 -> 35.9 ms per batch)
            parts = [self.submit(query_emb[i:i + 256], query_terms[i:i + 256], query_lens[i:i + 256], k, fetch_k)
                     for i in range(0, query_emb.shape[0], 256)]
            return _TicketGroup(parts)
        lane = (self._searches + 1) % LANES
        if lane not in self._lane_streams:
            self._lane_streams[lane] = torch.cuda.Stream(dev)
        stream = self._lane_streams[lane]
        stream.wait_stream(torch.cuda.current_stream(dev))   # the inputs are ready
        with torch.cuda.stream(stream):
            out = self.search(query_emb, query_terms, query_lens, k, fetch_k, check_overflow=False, lane=lane)
            done = torch.cuda.Event()
            done.record(stream)
        return Ticket(out, done, dev, (query_emb, query_terms, query_lens))


class Ticket:
    """Handle of a submitted search (ShardedHybrid.submit)."""

    def __init__(self, out: dict, done: "torch.cuda.Event", device, inputs):
        self._out, self._done, self._device, self._inputs = out, done, device, inputs

    def wait(self) -> dict:
        """The result dict; the CURRENT stream is ordered after the search (no host synchronisation)."""
        cur = torch.cuda.current_stream(self._device)
        cur.wait_event(self._done)
        for t in self._out.values():
            if isinstance(t, torch.Tensor):
                t.record_stream(cur)   # produced on the lane's stream, consumed on the caller's
        self._inputs = None
        return self._out


class _TicketGroup:
    """The tickets of one call that was split into 256-query groups: `wait` concatenates their results in order."""

    def __init__(self, parts):
        self._parts = parts

    def wait(self) -> dict:
        outs = [t.wait() for t in self._parts]
        return {key: (torch.cat([o[key] for o in outs], dim=0) if isinstance(outs[0][key], torch.Tensor) else outs[0][key])
                for key in outs[0]}


class _ShardedList:
    """Shared plumbing of the single-list sharded searches below: one NCCL all-gather of each rank's packed winners
    ([B, width] int64 words) per batch, then one merge launch.  (The hybrid search, the BASELINE metric, has the fused
    peer-memory exchange above; these two serve BASELINE configs 2 and 4.)"""

    exchange = "nccl"
    exchange_note = ""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._bufs: dict = {}          # lane -> all-gather buffer
        self._lane_streams: dict = {}
        self._submitted = 0
        self.lanes = LANES

    def _gather(self, mine: torch.Tensor, lane: int = 0) -> torch.Tensor:
        """mine int64 [B, w] -> [B, G, w] (shard-major inside a query)."""
        Bq, w = mine.shape
        buf = self._bufs.get(lane)
        if buf is None or buf.shape != (self.world, Bq, w):
            buf = self._bufs[lane] = torch.empty((self.world, Bq, w), dtype=torch.int64, device=mine.device)
        dist.all_gather_into_tensor(buf.view(-1), mine.contiguous().view(-1), group=self.group)
        return buf.permute(1, 0, 2)

    def close(self):
        self._bufs = {}

    def submit(self, *args, **kw) -> "Ticket":
        """Same calling convention as ShardedHybrid.submit: consecutive submissions rotate over LANES streams (own
        workspaces and gather buffers), so the re-scores / selection / merge of one batch run under the first pass of
        the next.  Every rank must submit the same sequence of batches."""
        dev = args[0].device
        lane = self._submitted % LANES
        self._submitted += 1
        if lane not in self._lane_streams:
            self._lane_streams[lane] = torch.cuda.Stream(dev)
        stream = self._lane_streams[lane]
        stream.wait_stream(torch.cuda.current_stream(dev))   # the inputs are ready
        with torch.cuda.stream(stream):
            out = self.search(*args, check_overflow=False, lane=lane, **kw)
            done = torch.cuda.Event()
            done.record(stream)
        return Ticket(out, done, dev, args)


class ShardedCosine(_ShardedList):
    """Exact cosine top-k over a row-sharded corpus (BASELINE config 2 at N GPUs): local `CosineIndex.topk`, all-gather
    of the G * k winners, `orag_topk_merge` by (score desc, id asc)."""

    def __init__(self, index: engine.CosineIndex, group=None):
        super().__init__(group)
        self.index = index

    def search(self, query_emb, k: int = 10, check_overflow: bool = True, lane: int = 0):
        st: list = []
        ids, sc = self.index.topk(query_emb, k, check_overflow=check_overflow, status_out=st, lane=lane)
        status = st[0] if not check_overflow else torch.zeros_like(st[0])
        if self.world > 1:
            mine = torch.cat([ids, sc.view(torch.int64), status.long()[:, None]], dim=1)
            g = self._gather(mine, lane)
            Bq = ids.shape[0]
            ids, sc, _ = engine.topk_merge(g[:, :, :k].reshape(Bq, -1).contiguous(),
                                           g[:, :, k:2 * k].reshape(Bq, -1).contiguous().view(torch.float64), k)
            status = g[:, :, 2 * k].amax(dim=1).to(torch.int32)
            if check_overflow and bool(status.any()):
                raise _ffi.OragError("sharded cosine search: a shard reported an unrepaired candidate overflow")
        return {"ids": ids, "scores": sc, "cos_ids": ids, "cos_scores": sc, "status": status}


class ShardedBm25(_ShardedList):
    """BM25 top-k over a doc-sharded inverted index with GLOBAL statistics (BASELINE config 4 at N GPUs): local RAW
    top-(k + guard) and the shard's max raw score, all-gather, division by the global max and merge in one launch
    (`orag_topk_merge` with shard maxima: rag/retrieval.py:343-345 over the whole corpus)."""

    def __init__(self, index, group=None):
        super().__init__(group)
        self.index = index

    def search(self, query_terms, query_lens, k: int = 10, check_overflow: bool = True, lane: int = 0):
        st: list = []
        if self.world == 1:
            ids, sc, mx = self.index.topk(query_terms, query_lens, k, normalize=True, check_overflow=check_overflow,
                                          status_out=st, lane=lane)
            status = st[0] if not check_overflow else torch.zeros_like(st[0])
        else:
            kk = k + BM25_GUARD
            ids, raw, mx = self.index.topk(query_terms, query_lens, kk, normalize=False, check_overflow=check_overflow,
                                           status_out=st, lane=lane)
            status = st[0] if not check_overflow else torch.zeros_like(st[0])
            mine = torch.cat([ids, raw.view(torch.int64), mx.view(torch.int64)[:, None], status.long()[:, None]], dim=1)
            g = self._gather(mine, lane)
            Bq = ids.shape[0]
            ids, sc, mx = engine.topk_merge(g[:, :, :kk].reshape(Bq, -1).contiguous(),
                                            g[:, :, kk:2 * kk].reshape(Bq, -1).contiguous().view(torch.float64), k,
                                            shard_max=g[:, :, 2 * kk].contiguous().view(torch.float64))
            status = g[:, :, 2 * kk + 1].amax(dim=1).to(torch.int32)
            if check_overflow and bool(status.any()):
                raise _ffi.OragError("sharded BM25 search: a shard reported an unrepaired candidate overflow")
        return {"ids": ids, "scores": sc, "bm25_ids": ids, "bm25_scores": sc, "bm25_max": mx, "status": status}
