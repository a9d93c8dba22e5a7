"""CPU, world_size=2, gloo: the host-side logic of the row-sharded path (optimized_rag_b200/dist.py):
global BM25 statistics by all-reduce, and the single all-gather exchange of local winners + merge.
The per-shard kernels are GPU-only; here the oracle stands in for them so that only the sharding /
packing / merge protocol is under test."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from optimized_rag_b200 import synthetic as syn

import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent))

N, DIM, VOCAB, NQ, K = 600, 64, 300, 6, 10


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from optimized_rag_b200.bm25_index import idf_table
        from optimized_rag_b200.dist import (BM25_GUARD, pack_local, reduce_stats, shard_range, token_position_base,
                                             unpack_gathered)
        from test_host_logic import _numpy_stats
        thr = syn.zipf_thresholds(VOCAB)
        lo, hi = shard_range(N, rank, world)
        # --- global statistics from local shards
        off, tok = syn.token_corpus(syn.SEED_TOKENS, lo, hi - lo, VOCAB, 5, 40, thr)
        # (the per-shard counting pass is a CUDA kernel; numpy stands in for it here: the collective half is under test)
        base, n_total, total_len = token_position_base(hi - lo, int(off[-1]))
        st = reduce_stats(_numpy_stats(off, tok, VOCAB, pos_base=base), n_total, total_len)
        g_off, g_tok = syn.token_corpus(syn.SEED_TOKENS, 0, N, VOCAB, 5, 40, thr)
        orc = oracle.BM25Index(g_off, g_tok, VOCAB)
        idf, avg_idf, eps = idf_table(st)
        assert st.n_docs == N and st.total_len == int(g_off[-1]) and st.avgdl == orc.avgdl
        assert np.array_equal(idf.view(np.uint64), orc.idf.view(np.uint64)) and eps == orc.eps
        # --- exchange: local winners (oracle restricted to this shard's rows) -> one all-gather -> merge
        corpus = syn.embeddings(syn.SEED_CORPUS, 0, N, DIM, 20)
        queries = syn.query_embeddings(NQ, N, DIM, dup_per_mille=20)
        qtok, qlen = syn.keyword_queries(NQ, VOCAB, min_rank=3, thresholds=thr)
        kk = K + BM25_GUARD
        ci = np.full((NQ, K), -1, np.int64); cs = np.zeros((NQ, K))
        bi = np.full((NQ, kk), -1, np.int64); bs = np.zeros((NQ, kk)); bm = np.zeros(NQ)
        want = []
        for b in range(NQ):
            cos = oracle.cosine_scores(corpus, queries[b])
            raw = orc.scores_raw(qtok[b, :qlen[b]])
            i, v = oracle.topk(cos[lo:hi], K, id_base=lo)
            ci[b, :len(i)], cs[b, :len(i)] = i, v
            i, v = oracle.topk(raw[lo:hi], kk, id_base=lo)
            bi[b, :len(i)], bs[b, :len(i)] = i, v
            bm[b] = max(raw[lo:hi].max(), 0.0)
            m = raw.max() if raw.max() > 0 else 1.0
            want.append((oracle.topk(cos, K), oracle.topk(raw / m, K), m))
        mine = pack_local(torch.from_numpy(ci), torch.from_numpy(cs), torch.from_numpy(bi), torch.from_numpy(bs),
                          torch.from_numpy(bm), torch.full((NQ,), rank, dtype=torch.int32))
        buf = torch.empty((world,) + tuple(mine.shape), dtype=torch.int64)
        dist.all_gather_into_tensor(buf.view(-1), mine.view(-1))
        gci, gcs, gbi, gbs, gbm, gst = unpack_gathered(buf, K, kk)
        assert gci.shape == (NQ, world * K) and gbm.shape == (NQ, world)
        assert gst.tolist() == [list(range(world))] * NQ  # status column travels with the winners
        for b in range(NQ):
            order = sorted(np.nonzero(gci[b].numpy() >= 0)[0], key=lambda j: (-gcs[b, j].item(), gci[b, j].item()))[:K]
            assert gci[b, order].tolist() == want[b][0][0].tolist()
            assert gcs[b, order].tolist() == want[b][0][1].tolist()
            m = max(gbm[b].max().item(), 0.0) or 1.0
            assert m == want[b][2]
            norm = gbs[b].numpy() / m
            order = sorted(np.nonzero(gbi[b].numpy() >= 0)[0], key=lambda j: (-norm[j], gbi[b, j].item()))[:K]
            assert gbi[b, order].tolist() == want[b][1][0].tolist()
            assert norm[order].tolist() == want[b][1][1].tolist()
        out_q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        import traceback
        out_q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_sharded_protocol_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in results:
        assert msg == "ok", f"rank {rank}: {msg}"


def test_shard_range_covers_everything():
    from optimized_rag_b200.dist import shard_range
    for n, w in [(10, 3), (10_000_000, 8), (7, 8), (0, 2)]:
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
