"""Experiment: how much would per-(tile, term) upper bounds ("block-max") tighten the MaxScore partition?
Compares essential-posting counts under global per-term bounds (what bm25_ms.cu uses) and tile-level bounds,
at the final thresholds of a batch."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from optimized_rag_b200 import engine, synthetic as syn  # noqa: E402
from optimized_rag_b200.bm25_index import Bm25Index  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
B, V = 256, 50000
dev = "cuda:0"
thr = syn.zipf_thresholds(V)
doc_off, tokens = engine.gen_token_corpus(n, 0, syn.SEED_TOKENS, thr, V, 100, 300, device=dev)
ix = Bm25Index(doc_off, tokens, V, tile_docs=2048)
del tokens
qt, ql = syn.keyword_queries(B, V, thresholds=thr)
qt_d, ql_d = torch.from_numpy(qt).to(dev), torch.from_numpy(ql).to(dev)
ix.topk(qt_d, ql_d, 10, force="sparse", check_overflow=False)
torch.cuda.synchronize()
ws = ix._ws
al = lambda x: (x + 255) // 256 * 256
o_qd = al(B * 8) + al(B * 4) + al(B * 4) + al(B * 8192 * 4)
thr_f = ws[0:B * 8].view(torch.float64).float() * (1 - 1 / 512)
qd = ws[o_qd:o_qd + B * 32 * 16].view(torch.int32).view(B, 32, 4)
term = qd[:, :, 0].long()
w = qd[:, :, 1].contiguous().view(torch.float32)
pre = qd[:, :, 2].contiguous().view(torch.float32)
nn = qd[:, 0, 3]
valid = torch.arange(32, device=dev)[None, :] < nn[:, None]
off = ix.fp_tile_term_off.long()                        # [tiles, V+1]
n_tiles = off.shape[0]
# per-(tile, term) max r
r16 = (ix.postings_r16[:-4] & 0xFFFF).to(torch.int16).view(torch.float16).float()
lens = (off[:, 1:] - off[:, :-1]).reshape(-1)
tmax = torch.segment_reduce(r16, "max", lengths=lens, unsafe=True).nan_to_num(0.0, neginf=0.0).view(n_tiles, V)
tc = term.clamp(min=0)
ess_g = ess_t = tot = 0
skip_pairs = 0
for t0 in range(0, n_tiles, 256):
    t1 = min(n_tiles, t0 + 256)
    ln = (off[t0:t1, 1:] - off[t0:t1, :-1])[:, tc]       # [tiles, B, 32] run lengths
    ln = torch.where(valid[None], ln, torch.zeros_like(ln))
    ub_t = w[None] * tmax[t0:t1][:, tc] * valid[None]      # tile-level bounds
    pre_t = torch.cumsum(ub_t, dim=2)
    ne_g = (pre < thr_f[:, None])[None] & valid[None]
    ne_t = (pre_t < thr_f[None, :, None]) & valid[None]
    # non-essential sets are prefixes: enforce prefix property for the tile-level one
    ne_t = torch.cumprod(ne_t.int(), dim=2).bool()
    tot += int(ln.sum())
    ess_g += int((ln * (~ne_g & valid[None])).sum())
    ess_t += int((ln * (~ne_t & valid[None])).sum())
    skip_pairs += int(((ln * (~ne_t & valid[None])).sum(2) == 0).sum())
print(f"postings {tot:.3e}: essential under global bounds {ess_g / tot:.1%}, under tile bounds {ess_t / tot:.1%}; "
      f"pairs with no essential posting under tile bounds {skip_pairs / (n_tiles * B):.1%}")
