import numpy as np, torch, sys
sys.path.insert(0, "/root/repo")
from optimized_rag_b200 import engine, synthetic as syn
for n, dim, nq in [(128 * 6, 256, 64), (128 * 5 + 77, 256, 70), (128 * 301, 1536, 256), (4500, 64, 3)]:
    corpus = torch.from_numpy(syn.embeddings(syn.SEED_CORPUS, 0, n, dim, 2)).cuda()
    q = torch.from_numpy(syn.query_embeddings(nq, n, dim, dup_per_mille=2)).cuda()
    want = engine.CosineIndex(corpus, mode="exact").topk(q, 10)
    for mode in ("f16", "tf32"):
        got = engine.CosineIndex(corpus, mode=mode).topk(q, 10)
        ok = torch.equal(want[0], got[0]) and torch.equal(want[1], got[1])
        print(n, dim, nq, mode, "OK" if ok else "MISMATCH", flush=True)
        assert ok
print("2sm ok")
