"""Drop-in for `ReciprocalRankFusion` (rag/reranker.py:212-271), fused on the GPU.

Same constructor, same `fuse(result_lists, top_k=10)` signature and semantics: results are keyed by
their `content` string (`doc.get('content', '')`), a key's first sighting supplies the returned dict,
later sightings only add 1/(k + rank); the union is ordered by RRF score with ties in first-sighting
order (stable sort over dict insertion order); the first `top_k` dicts get `doc['rrf_score']` set IN
PLACE and are returned.  The arithmetic runs in csrc/rrf.cu (orag_rrf_fuse) -- there is no CPU path.
"""
from __future__ import annotations

import logging
from typing import Any, Dict, List

import torch

from . import _ffi, engine

logger = logging.getLogger(__name__)

_MAX_LISTS = 8
_MAX_UNION = 128


class ReciprocalRankFusion:
    """Combine rankings from multiple sources using RRF"""

    def __init__(self, k: int = 60, device: str | torch.device = "cuda"):
        self.k = k
        self.device = torch.device(device)

    def fuse(self, result_lists: List[List[Dict[str, Any]]], top_k: int = 10) -> List[Dict[str, Any]]:
        key_of: Dict[str, int] = {}
        first_doc: List[Dict[str, Any]] = []
        id_lists: List[List[int]] = []
        for result_list in result_lists:
            ids = []
            for doc in result_list:
                content = doc.get('content', '')
                kid = key_of.get(content)
                if kid is None:
                    kid = len(first_doc)
                    key_of[content] = kid
                    first_doc.append(doc)
                ids.append(kid)
            id_lists.append(ids)
        if not first_doc or top_k <= 0:
            logger.info(f"RRF fused {len(result_lists)} lists into 0 results")
            return []
        n_lists = len(id_lists)
        list_len = max(1, max(len(l) for l in id_lists))
        if n_lists > _MAX_LISTS or n_lists * list_len > _MAX_UNION or list_len > 255:
            raise _ffi.OragError(
                f"orag_rrf_fuse supports up to {_MAX_LISTS} lists with n_lists*len <= {_MAX_UNION}; got {n_lists} x "
                f"{list_len} (the reference's callers pass <= 2 x 15, rag/nodes/rerank_and_eval.py:229-242)")
        arr = torch.full((1, n_lists, list_len), -1, dtype=torch.int64)
        for i, ids in enumerate(id_lists):
            if ids:
                arr[0, i, :len(ids)] = torch.tensor(ids, dtype=torch.int64)
        k_out = min(top_k, len(first_doc))
        ids, scores = engine.rrf_fuse(arr.to(self.device), self.k, k_out)
        ids = ids[0].cpu().tolist()
        scores = scores[0].cpu().tolist()
        fused_results = []
        for kid, sc in zip(ids, scores):
            if kid < 0:
                break
            doc = first_doc[kid]
            doc['rrf_score'] = sc
            fused_results.append(doc)
        logger.info(f"RRF fused {len(result_lists)} lists into {len(fused_results)} results")
        return fused_results


class MMRDiversifier:
    """Drop-in for `MMRDiversifier` (rag/reranker.py:104-195): Maximal Marginal Relevance over results that carry
    an `embedding`.  All cosines of a call -- query x documents and documents x documents -- come from ONE launch of
    the float64 cosine kernel (orag_cosine_dense, the reference's arithmetic: rag/reranker.py:196-208); the greedy
    selection then replays the reference's loop over that matrix (first maximum wins, `mmr_score` written in place).
    Embeddings are cast to fp32 (what the store / pgvector hold); results whose embedding is empty, non-numeric,
    non-finite or of another dimension than the query's are filtered out like the reference filters invalid ones.
    """

    def __init__(self, lambda_param: float = 0.7, device: str | torch.device = "cuda"):
        self.lambda_param = lambda_param
        self.device = torch.device(device)

    def diversify(self, query_embedding: List[float], results: List[Dict[str, Any]], top_k: int = 5
                  ) -> List[Dict[str, Any]]:
        import math
        if not results:
            return []
        dim = len(query_embedding) if query_embedding else 0
        valid = [r for r in results
                 if r.get('embedding') and isinstance(r['embedding'], list) and len(r['embedding']) == dim
                 and all(isinstance(v, (int, float)) and not math.isnan(v) and not math.isinf(v)
                         for v in r['embedding'])]
        if not valid or dim == 0:
            logger.warning("MMR: No valid embeddings found, returning original results")
            return results[:top_k]
        if len(valid) < len(results):
            logger.warning(f"MMR: Filtered {len(results) - len(valid)} results with invalid embeddings")
        m = len(valid)
        with _ffi.GPU_LOCK:
            emb = torch.tensor([r['embedding'] for r in valid], dtype=torch.float32, device=self.device)
            q = torch.tensor([query_embedding], dtype=torch.float32, device=self.device)
            cos = engine.CosineIndex(emb, mode="exact").dense(torch.cat([q, emb]).contiguous()).cpu().tolist()
        rel, sim = cos[0], cos[1:]          # rel[j] = cos(query, doc j); sim[i][j] = cos(doc i, doc j)
        selected: List[int] = []
        remaining = list(range(m))
        while len(selected) < top_k and remaining:
            best_j, best_score = None, None
            for j in remaining:
                diversity = 1 - max(sim[j][s] for s in selected) if selected else 1.0
                score = self.lambda_param * rel[j] + (1 - self.lambda_param) * diversity
                if best_score is None or score > best_score:   # max(): the first maximum wins
                    best_j, best_score = j, score
            valid[best_j]['mmr_score'] = best_score
            selected.append(best_j)
            remaining.remove(best_j)
        logger.info(f"MMR diversified to {len(selected)} results")
        return [valid[j] for j in selected]
