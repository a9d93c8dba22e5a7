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
fp_tile = int(sys.argv[4]) if len(sys.argv) > 4 and sys.argv[4].isdigit() else None
import time
torch.cuda.synchronize(); _t0 = time.perf_counter()
ix = Bm25Index(doc_off, tokens, V, tile_docs=tile, fp_tile_docs=fp_tile)
torch.cuda.synchronize()
print(f"index build ({n} docs, {ix.n_postings} postings, fp view {ix.n_postings_fp} incl. padding): {time.perf_counter() - _t0:.3f} s", flush=True)
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
# MaxScore statistics from the workspace of the last "sparse" call (layout: bm25_ms.cu ms_carve)
if ix.postings_r16 is not None:
    ws = ix._ws
    al = lambda x: (x + 255) // 256 * 256
    o_thr = 0
    o_qd = al(B * 8) + al(B * 4) + al(B * 4) + al(B * 8192 * 4)
    thr_f = ws[o_thr:o_thr + B * 8].view(torch.float64).float() * (1 - 1 / 512)
    qd_raw = ws[o_qd:o_qd + B * 32 * 16].view(torch.int32).view(B, 32, 4)
    term = qd_raw[:, :, 0].long()
    pre = qd_raw[:, :, 2].contiguous().view(torch.float32)
    nn = qd_raw[:, 0, 3]
    off = ix.fp_tile_term_off.long()
    df = (off[:, 1:] - off[:, :-1]).sum(0)
    valid = torch.arange(32, device=dev)[None, :] < nn[:, None]
    dfq = torch.where(valid, df[term.clamp(min=0)], torch.zeros_like(term))
    ne = valid & (pre < thr_f[:, None])
    tot = dfq.sum().item(); ess = dfq[valid & ~ne].sum().item()
    print(f"MaxScore at final thresholds: essential postings {ess:.3e} of {tot:.3e} = {ess / tot:.1%}; "
          f"queries with no non-essential term: {int((ne.sum(1) == 0).sum())}/{B}; mean terms {nn.float().mean():.2f}, "
          f"mean non-essential {ne.sum(1).float().mean():.2f}")
    frac = (dfq * (valid & ~ne)).sum(1).float() / dfq.sum(1).float().clamp(min=1)
    print("per-query essential fraction deciles:", [round(float(x), 3) for x in torch.quantile(frac, torch.linspace(0, 1, 11, device=dev))])
    heavy = torch.argsort(-(dfq * (valid & ~ne)).sum(1))[:5]
    for h in heavy.tolist():
        print("  heavy query", h, "n", int(nn[h]), "thr", float(thr_f[h]), "pre", [round(float(x), 2) for x in pre[h, :nn[h]]],
              "df/N", [round(float(x) / n, 4) for x in dfq[h, :nn[h]]])
print("ms == exact_tiles:", all(torch.equal(x, y) for x, y in zip(res["sparse"], res["exact_tiles"])))
if "--no-dense" in sys.argv:
    sys.exit(0)
ids2, sc2, mx2 = ix.topk(qt_d, ql_d, 10, force="dense")
torch.cuda.synchronize()
L.orag_profile_read(ctypes.byref(a), ctypes.byref(b))
print(f"dense tile kernel (last chunk) {b.value:.3f} ms; equal ids {torch.equal(ids, ids2)} scores {torch.equal(sc, sc2)}")
