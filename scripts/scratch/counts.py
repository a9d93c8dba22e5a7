"""Candidate-set sizes of the cosine first pass (diagnostic): python scripts/scratch/counts.py ROWS"""
import sys, torch, numpy as np
sys.path.insert(0, ".")
from optimized_rag_b200 import engine, synthetic as syn
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_250_000
dev = torch.device("cuda:0")
emb = engine.gen_embeddings(rows, 1536, 0, syn.SEED_CORPUS, 0, device=dev)
q = torch.from_numpy(syn.query_embeddings(256, rows, 1536)).to(dev)
ix = engine.CosineIndex(emb, mode="f16")
for k in (10, 64):
    ix.topk(q, k)
    torch.cuda.synchronize()
    c, v = ix.last_counts(256, k)
    print(f"rows {rows} k {k}: candidates mean {c.mean():.1f} max {c.max()} | survivors mean {v.mean():.1f} max {v.max()}")
