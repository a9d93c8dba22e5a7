"""Small batched-cosine consumers of the RAG graph: drop-ins for `apply_mmr` and
`cosine_similarity` of rag/nodes/helpers.py:183-290.

The reference evaluates O(k * m) Python cosines inside the greedy MMR loop; here every dot product and sum of squares
of a call -- query x documents and documents x documents -- comes from ONE launch of the float64 kernel
(orag_dot_dense, the reference's summation order), and the host only finishes each cosine the way this particular
reference function does: magnitudes as `sum ** 0.5` (libm pow, not sqrt), `dot / (mag1 * mag2)`, 0.0 for a zero
magnitude.  The greedy selection then replays the reference's loop over that matrix (`max` keeps the first maximum).
Embeddings are cast to fp32 (what the store / pgvector hold); vectors of different lengths are zero-padded, which is
bit-identical to the reference's `zip` truncation.  No CPU path: without the CUDA library the call raises inside and,
like the reference (:262-264), `apply_mmr` then returns `documents[:k]`.
"""
from __future__ import annotations

import logging
from typing import Any, Dict, List

import torch

from . import _ffi, engine

logger = logging.getLogger(__name__)
DEVICE = "cuda"


def _sums(rows: List[List[float]], device) -> tuple:
    """(dots [m][m], sq [m]) as Python floats for the given vectors (row i against row j)."""
    dim = max((len(r) for r in rows), default=0)
    dim = max(dim, 1)
    dim += (-dim) % 4   # rows are read in 16-byte pieces: zero columns change nothing
    mat = torch.tensor([list(r) + [0.0] * (dim - len(r)) for r in rows], dtype=torch.float32, device=device)
    dots, row_sq, _ = engine.CosineIndex(mat, mode="exact").dots(mat)
    return dots.cpu().tolist(), row_sq.cpu().tolist()


def _finish(dot: float, sq1: float, sq2: float) -> float:
    mag1 = sq1 ** 0.5
    mag2 = sq2 ** 0.5
    if mag1 == 0 or mag2 == 0:
        return 0.0
    return dot / (mag1 * mag2)


def cosine_similarity(vec1: List[float], vec2: List[float]) -> float:
    """Cosine similarity of two vectors (rag/nodes/helpers.py:266-290); 0.0 on any failure, like the reference."""
    try:
        if min(len(vec1), len(vec2)) == 0:
            return 0.0
        with _ffi.GPU_LOCK:
            dots, sq = _sums([vec1, vec2], torch.device(DEVICE))
        return _finish(dots[0][1], sq[0], sq[1])
    except Exception as e:  # noqa: BLE001
        logger.error(f"Cosine similarity calculation failed: {e}", exc_info=True)
        return 0.0


def apply_mmr(query: str, documents: List[Dict[str, Any]], lambda_: float, k: int, embedding_service
              ) -> List[Dict[str, Any]]:
    """Maximal Marginal Relevance (rag/nodes/helpers.py:183-264): MMR = lambda * relevance - (1 - lambda) * max
    similarity to the documents already selected; documents without an embedding get one (stored in place)."""
    if len(documents) <= k:
        return documents
    try:
        query_embedding = embedding_service.generate_embedding(query)
        doc_embeddings = []
        for doc in documents:
            if "embedding" in doc and doc["embedding"]:
                doc_embeddings.append(doc["embedding"])
            else:
                emb = embedding_service.generate_embedding(doc.get("content", doc.get("text", "")))
                doc["embedding"] = emb
                doc_embeddings.append(emb)
        with _ffi.GPU_LOCK:
            dots, sq = _sums([query_embedding] + doc_embeddings, torch.device(DEVICE))
        m = len(documents)
        relevance = [_finish(dots[0][1 + i], sq[0], sq[1 + i]) for i in range(m)]

        def sim(i, j):
            return _finish(dots[1 + i][1 + j], sq[1 + i], sq[1 + j])

        selected: List[int] = []
        remaining = list(range(m))
        while len(selected) < k and remaining:
            best_idx, best_score = None, None
            for idx in remaining:
                max_sim = max(sim(idx, s) for s in selected) if selected else 0.0
                mmr = lambda_ * relevance[idx] - (1 - lambda_) * max_sim
                if best_score is None or mmr > best_score:   # max(): the first maximum wins
                    best_idx, best_score = idx, mmr
            selected.append(best_idx)
            remaining.remove(best_idx)
            logger.debug(f"MMR selected doc {best_idx} with score {best_score:.3f}")
        return [documents[i] for i in selected]
    except Exception as e:  # noqa: BLE001
        logger.error(f"MMR calculation failed: {e}", exc_info=True)
        return documents[:k]
