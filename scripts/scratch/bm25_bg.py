"""BM25 first pass, background configuration vs the dense path, at the shapes the sharded search uses."""
import sys, torch, numpy as np
sys.path.insert(0, ".")
from optimized_rag_b200 import engine, synthetic as syn
from optimized_rag_b200.bm25_index import Bm25Index, Bm25Plan
dev = torch.device("cuda:0")
V = 50000
thr = syn.zipf_thresholds(V)
qt, ql = syn.keyword_queries(256, V, thresholds=thr)
qt, ql = torch.from_numpy(qt).to(dev), torch.from_numpy(ql).to(dev)
for n in (150000, 400000):
    off, tok = engine.gen_token_corpus(n, 0, syn.SEED_TOKENS, thr, V, 100, 300, device=dev)
    plan = Bm25Plan(off, tok, V)
    ix = Bm25Index.from_plan(plan)
    for k in (10, 16):
        want = ix.topk(qt, ql, k, normalize=False, force="dense", check_overflow=False)
        for bg in (False, True):
            bad = flagged = 0
            bits = {}
            for it in range(20):
                st = []
                got = ix.topk(qt, ql, k, normalize=False, check_overflow=False, status_out=st, background=bg)
                torch.cuda.synchronize()
                flagged += int((st[0] != 0).sum())
                for v in st[0][st[0] != 0].tolist():
                    bits[v] = bits.get(v, 0) + 1
                ok_rows = (st[0] == 0)
                bad += int((~((got[0] == want[0]).all(1) & (got[1] == want[1]).all(1)) & ok_rows).sum())
            print(f"docs {n} tile {ix.struct.fp_tile_docs} k {k} background {bg}: flagged {flagged} unflagged-but-different {bad} (20 runs x 256 queries) status words {bits}", flush=True)
