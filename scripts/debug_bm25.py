"""GPU debug helper: BM25 sparse path statistics (emission counts, overflow, kernel time)."""
import ctypes
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from optimized_rag_b200 import _ffi, engine, synthetic as syn  # noqa: E402
from optimized_rag_b200.bm25_index import Bm25Index  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
tile = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
dev = "cuda:0"
V = 50000
thr = syn.zipf_thresholds(V)
doc_off, tokens = engine.gen_token_corpus(n, 0, syn.SEED_TOKENS, thr, V, 100, 300, device=dev)
ix = Bm25Index(doc_off, tokens, V, tile_docs=tile)
del tokens
qt, ql = syn.keyword_queries(B, V, thresholds=thr)
qt_d, ql_d = torch.from_numpy(qt).to(dev), torch.from_numpy(ql).to(dev)
L = _ffi.lib()
L.orag_profile_enable(1)
a, b = ctypes.c_float(), ctypes.c_float()
res = {}
for force in ("exact_tiles", "sparse"):
    for it in range(3):
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        ids, sc, mx = ix.topk(qt_d, ql_d, 10, force=force, check_overflow=False, status_out=(st := []))
        t1.record()
        torch.cuda.synchronize()
        L.orag_profile_read(ctypes.byref(a), ctypes.byref(b))
        ws = ix._ws
        off = 0
        al = lambda x: (x + 255) // 256 * 256
        off += al(B * 8)
        cnt = ws[off:off + B * 4].view(torch.int32).cpu().numpy()
        print(f"{force} iter {it}: tile kernel {b.value:.3f} ms, whole call {t0.elapsed_time(t1):.3f} ms; emissions per "
              f"query: min {cnt.min()} mean {cnt.mean():.0f} max {cnt.max()} total {cnt.sum()}; overflow "
              f"{int((st[0] != 0).sum())}; postings(6B) {ix.posting_bytes(qt_d, ql_d) / 6:.3e}", flush=True)
    res[force] = (ids, sc, mx)
print("ms == exact_tiles:", all(torch.equal(x, y) for x, y in zip(res["sparse"], res["exact_tiles"])))
if "--no-dense" in sys.argv:
    sys.exit(0)
ids2, sc2, mx2 = ix.topk(qt_d, ql_d, 10, force="dense")
torch.cuda.synchronize()
L.orag_profile_read(ctypes.byref(a), ctypes.byref(b))
print(f"dense tile kernel (last chunk) {b.value:.3f} ms; equal ids {torch.equal(ids, ids2)} scores {torch.equal(sc, sc2)}")
