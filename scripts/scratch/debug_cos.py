"""GPU debug helper: candidate statistics of the cosine first pass (workspace layout: csrc/api.cu carve_tc)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from optimized_rag_b200 import engine, synthetic as syn  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
B = 256
dev = torch.device("cuda:0")
corpus = engine.gen_embeddings(N, 1536, 0, syn.SEED_CORPUS, 0, device=dev)
cos = engine.CosineIndex(corpus, mode="f16")
q = torch.from_numpy(syn.query_embeddings(B, N, 1536)).to(dev)
for _ in range(2):
    cos.topk(q, 10, check_overflow=False)
torch.cuda.synchronize()
ws = cos._ws[engine.MODE["f16"]]
al = lambda x: (x + 255) // 256 * 256
o = al(256 * 8) + al(256 * 4) * 3
cnt = ws[o:o + 1024].view(torch.int32).cpu().numpy()
o2 = o + al(1024) + al(256 * 512 * 4) + al(256 * 4096 * 4) * 2 + al(256 * 512 * 4)
surv = ws[o2:o2 + 1024].view(torch.int32).cpu().numpy()
print(f"first-pass candidates per query: min {cnt.min()} mean {cnt.mean():.0f} max {cnt.max()}; "
      f"fp32 survivors: min {surv.min()} mean {surv.mean():.1f} max {surv.max()}")
