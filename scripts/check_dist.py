"""torchrun --nproc-per-node G scripts/check_dist.py [rows] [modes, comma separated: tf32,f16]

Row-sharded hybrid search over G GPUs must equal the single-shard search over the same corpus,
bit for bit (ids, cosine / BM25 / RRF scores).  Rank 0 additionally builds the whole corpus on its
own GPU as the comparison point."""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from optimized_rag_b200 import engine, synthetic as syn  # noqa: E402
from optimized_rag_b200.bm25_index import Bm25Index  # noqa: E402
from optimized_rag_b200.dist import ShardedBm25, ShardedCosine, ShardedHybrid, shard_range, sharded_stats  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 300_000
DIM, VOCAB, B, K = 1536, 50000, 64, 10
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
thr = syn.zipf_thresholds(VOCAB)


def build(lo, hi, stats=None, mode="tf32"):
    corpus = engine.gen_embeddings(hi - lo, DIM, lo, syn.SEED_CORPUS, 1, device=dev)
    off, tok = engine.gen_token_corpus(hi - lo, lo, syn.SEED_TOKENS, thr, VOCAB, 100, 300, device=dev)
    st = stats(off, tok) if stats else None
    return engine.HybridShard(engine.CosineIndex(corpus, row_id_base=lo, mode=mode),
                              Bm25Index(off, tok, VOCAB, stats=st, doc_id_base=lo))


lo, hi = shard_range(N, rank, world)
q_emb = torch.from_numpy(syn.query_embeddings(B, N, DIM, dup_per_mille=1)).to(dev)
qt, ql = syn.keyword_queries(B, VOCAB, thresholds=thr)
qt, ql = torch.from_numpy(qt).to(dev), torch.from_numpy(ql).to(dev)
ok = True
KEYS = ("ids", "rrf_scores", "cos_ids", "cos_scores", "bm25_ids", "bm25_scores", "bm25_max")
for mode in (sys.argv[2].split(",") if len(sys.argv) > 2 else ("tf32", "f16")):
    shard = build(lo, hi, lambda o, t: sharded_stats(o, t, VOCAB), mode)
    ref = build(0, N, None, mode).search(q_emb, qt, ql, K) if rank == 0 else None
    for exchange in ("nccl", "peer"):
        sh = ShardedHybrid(shard, exchange=exchange)
        res = sh.search(q_emb, qt, ql, K)
        torch.cuda.synchronize()
        if rank == 0:
            for key in KEYS:
                same = torch.equal(res[key], ref[key])
                ok &= same
                print(f"[{mode}/{exchange}] world={world} rows={N}: {key:12s} {'identical' if same else 'DIFFERENT'}",
                      flush=True)
        # which of the two first passes flags candidate overflow on this shard (flagged queries are repaired: correct, slower)
        for cos_on in (True, False):
            shard.coschedule = cos_on
            for lane in range(3):
                ll = shard.local_lists(q_emb, qt, ql, K, K + 6, False, lane=lane)
                torch.cuda.synchronize()
                print(f"[{mode}/{exchange}] rank {rank} coschedule {cos_on} lane {lane}: cosine flagged "
                      f"{int((ll[5] != 0).sum())}, bm25 flagged {int((ll[6] != 0).sum())}", flush=True)
        shard.coschedule = True
        # the repair path (every query forced through it): maxima first, ranking by the normalised value on every shard
        import os
        os.environ["ORAG_TEST_FORCE_REPAIR"] = "1"
        rep = sh.search(q_emb, qt, ql, K)
        os.environ.pop("ORAG_TEST_FORCE_REPAIR")
        torch.cuda.synchronize()
        if rank == 0:
            same = all(torch.equal(rep[key], ref[key]) for key in KEYS) and not bool(rep["status"].any())
            ok &= same
            print(f"[{mode}/{exchange}] repair path (all queries): {'identical' if same else 'DIFFERENT'}", flush=True)
        # two batches in flight: tickets of interleaved submissions
        tickets = [sh.submit(q_emb, qt, ql, K) for _ in range(6)]
        outs = [tk.wait() for tk in tickets]
        torch.cuda.synchronize()
        if rank == 0:
            same = all(torch.equal(o[key], ref[key]) for o in outs for key in KEYS)
            flagged = [int((o["status"] != 0).sum()) for o in outs]
            same &= not any(flagged)
            ok &= same
            print(f"[{mode}/{exchange}] 6 submitted batches (three in flight): {'identical' if same else 'DIFFERENT'}", flush=True)
            if not same:
                for i, o in enumerate(outs):
                    diff = [key for key in KEYS if not torch.equal(o[key], ref[key])]
                    print(f"    batch {i}: differing {diff}; flagged queries {flagged[i]}; status bits "
                          f"{sorted(set(o['status'].tolist()))}", flush=True)
        # device time and host time per step of back-to-back searches (no host sync inside)
        for _ in range(5):
            sh.search(q_emb, qt, ql, K, check_overflow=False)
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        steps = 100
        import time
        e0.record(); t0 = time.perf_counter()
        for _ in range(steps):
            last = sh.search(q_emb, qt, ql, K, check_overflow=False)
        t_host = time.perf_counter() - t0
        e1.record(); torch.cuda.synchronize()
        same = all(torch.equal(last[key], res[key]) for key in KEYS) and not bool(last["status"].any())
        ok &= same
        t = torch.tensor([e0.elapsed_time(e1) / steps, t_host * 1e3 / steps], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"[{mode}/{exchange}] {steps} back-to-back searches: {t[0]:.3f} ms/step on the device, "
                  f"{t[1]:.3f} ms/step of host enqueue time; stable={same}", flush=True)
        sh.close()
        dist.barrier()
    # the single-list searches (BASELINE configs 2 and 4 under torchrun): plain and with six batches submitted on the
    # lanes, against the matching lists of the single-shard hybrid result
    for name, obj, args, keys in (("cosine", ShardedCosine(shard.cosine), (q_emb,), ("cos_ids", "cos_scores")),
                                  ("bm25", ShardedBm25(shard.bm25), (qt, ql), ("bm25_ids", "bm25_scores", "bm25_max"))):
        plain = obj.search(*args, K)
        outs = [tk.wait() for tk in [obj.submit(*args, K) for _ in range(6)]]
        torch.cuda.synchronize()
        if rank == 0:
            same = all(torch.equal(o[key], ref[key]) for o in [plain] + outs for key in keys)
            same &= not any(bool(o["status"].any()) for o in outs)
            ok &= same
            print(f"[{mode}] sharded {name}-only search, plain + 6 submitted: {'identical' if same else 'DIFFERENT'}",
                  flush=True)
        obj.close()
        dist.barrier()
if rank == 0:
    print("DIST CHECK", "PASSED" if ok else "FAILED", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
