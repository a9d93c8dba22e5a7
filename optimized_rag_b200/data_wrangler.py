"""Embedding-based de-duplication of the ingest path.

Mirrors `Deduplicator.semantic_dedup` (rag/data_wrangler.py:295-326), which `DataWrangler.process_chunks`
(rag/data_wrangler.py:529-535) runs over the chunks of an upload: a chunk is dropped when its cosine with an earlier
SURVIVING chunk reaches the threshold.  The reference evaluates O(m^2) Python cosines one pair at a time; here all
pairs of a block of chunks come from one launch of the float64 cosine kernel (orag_cosine_dense, the reference's
arithmetic) and the greedy scan walks the matrix.  Embeddings are cast to fp32 (what the store / pgvector hold).
There is no CPU path: the matrix comes from the GPU or the call raises.
"""
from __future__ import annotations

import logging
from typing import Any, Dict, List

import numpy as np
import torch

from . import _ffi, engine

logger = logging.getLogger(__name__)

BLOCK_ROWS = 2048  # chunks scored per launch: the float64 block [BLOCK_ROWS, m] stays at 16 KB per chunk of the upload


def _cosine_block(index: "engine.CosineIndex", emb: torch.Tensor, lo: int, hi: int) -> np.ndarray:
    """float64 cosines of chunks lo..hi-1 against ALL chunks -> host array [hi-lo, m]."""
    return index.dense(emb[lo:hi].contiguous()).cpu().numpy()


class Deduplicator:
    """Drop-in for the embedding half of `Deduplicator` (rag/data_wrangler.py:252-326)."""

    device = "cuda"

    @staticmethod
    def semantic_dedup(chunks: List[Dict[str, Any]], embeddings: List[List[float]], threshold: float = 0.95,
                       device: str | torch.device | None = None) -> List[Dict[str, Any]]:
        """Remove semantically similar chunks (same signature and result as the reference; `zip` semantics: the
        shorter of the two lists decides how many chunks are looked at)."""
        m = min(len(chunks), len(embeddings))
        if m == 0:
            logger.info(f"Semantic dedup: {len(chunks)} → 0")
            return []
        dev = torch.device(device or Deduplicator.device)
        kept: List[int] = []
        with _ffi.GPU_LOCK:
            emb = torch.tensor([list(e) for e in embeddings[:m]], dtype=torch.float32, device=dev)
            index = engine.CosineIndex(emb, mode="exact")
            for lo in range(0, m, BLOCK_ROWS):
                hi = min(m, lo + BLOCK_ROWS)
                sim = _cosine_block(index, emb, lo, hi)
                for i in range(lo, hi):
                    row = sim[i - lo]
                    if not kept or not bool((row[kept] >= threshold).any()):
                        kept.append(i)
        logger.info(f"Semantic dedup: {len(chunks)} → {len(kept)}")
        return [chunks[i] for i in kept]
