"""Drop-in for `HybridRetriever` (rag/retrieval.py:13-371) backed by the GPU kernels.

Same constructor, attributes and method signatures as the reference:
  retrieve(query, sources, top_k=20)        dispatcher, concatenates per-source lists, swallows errors
  hybrid_search(query, corpus, embeddings, query_embedding, top_k=10, documents_metadata=None,
                query_intent=None)          alpha*semantic + beta*keyword + gamma*temporal, stable sort
  get_weights_for_intent(intent)            INTENT_WEIGHTS table (rag/retrieval.py:22-47)
  _cosine_similarity / _bm25_scores / _simple_keyword_scores    the per-piece helpers

The arithmetic (float64 cosine, rank_bm25 BM25Okapi scores normalised by their max, the weighted sum
and the (score desc, index asc) ordering) runs on the GPU through include/orag.h; results are
bit-identical to the reference's CPU path on the same inputs (tests/test_gpu_boundary.py).
`datetime.now()` is injectable (`now=`) so that the temporal term is testable.
"""
from __future__ import annotations

import logging
from datetime import datetime
from typing import Any, Callable, Dict, List, Optional

import numpy as np
import torch

from . import engine
from .bm25_index import Bm25Index
from .document_store import TextVocab, _gpu_locked

logger = logging.getLogger(__name__)

# config.py:38-40 defaults of the reference (ENABLE_TEMPORAL_BOOST / RECENCY_WEIGHT / RECENCY_HALF_LIFE_DAYS)
ENABLE_TEMPORAL_BOOST = True
RECENCY_WEIGHT = 0.1
RECENCY_HALF_LIFE_DAYS = 30


class HybridRetriever:
    """Hybrid retrieval with adaptive weights by intent (GPU-backed)."""

    INTENT_WEIGHTS = {
        'question_answering': {'alpha': 0.55, 'beta': 0.40, 'gamma': 0.05},
        'fact_checking': {'alpha': 0.50, 'beta': 0.45, 'gamma': 0.05},
        'multi_hop_reasoning': {'alpha': 0.60, 'beta': 0.30, 'gamma': 0.10},
        'comparison': {'alpha': 0.50, 'beta': 0.45, 'gamma': 0.05},
        'summarization': {'alpha': 0.65, 'beta': 0.25, 'gamma': 0.10},
        'search': {'alpha': 0.45, 'beta': 0.50, 'gamma': 0.05},
        'clarification': {'alpha': 0.70, 'beta': 0.20, 'gamma': 0.10},
        'conversational': {'alpha': 0.70, 'beta': 0.20, 'gamma': 0.10},
        'default': {'alpha': 0.55, 'beta': 0.35, 'gamma': 0.10},
    }

    def __init__(self, memory_manager, document_store, agent_id: str, alpha: float = 0.55, beta: float = 0.35,
                 gamma: float = 0.10, weight_manager=None, use_adaptive_weights: bool = True,
                 device: str | torch.device = "cuda", now: Optional[Callable[[], datetime]] = None,
                 enable_temporal_boost: bool = ENABLE_TEMPORAL_BOOST, recency_weight: float = RECENCY_WEIGHT,
                 recency_half_life_days: float = RECENCY_HALF_LIFE_DAYS):
        self.memory_manager = memory_manager
        self.document_store = document_store
        self.agent_id = agent_id
        self.alpha = alpha
        self.beta = beta
        self.gamma = gamma
        self.weight_manager = weight_manager
        self.use_adaptive_weights = use_adaptive_weights
        self.device = torch.device(device)
        self._now = now or datetime.now
        self.enable_temporal_boost = enable_temporal_boost
        self.recency_weight = recency_weight
        self.recency_half_life_days = recency_half_life_days
        self.bm25_available = self._check_bm25()

    def get_weights_for_intent(self, intent: str) -> tuple:
        intent_key = intent.lower().replace(' ', '_') if intent else 'default'
        w = self.INTENT_WEIGHTS.get(intent_key, self.INTENT_WEIGHTS['default'])
        return w['alpha'], w['beta'], w['gamma']

    def _check_bm25(self) -> bool:
        """BM25 is built into the CUDA library; there is nothing optional to import."""
        from . import _ffi
        _ffi.lib()
        return True

    # ------------------------------------------------------------------ multi-source dispatch (:122-212)
    def retrieve(self, query: str, sources: List[str], top_k: int = 20) -> List[Dict[str, Any]]:
        all_results = []
        if 'archival' in sources or 'archival_memory' in sources:
            all_results.extend(self._retrieve_archival(query, top_k))
        if 'documents' in sources:
            all_results.extend(self._retrieve_documents(query, top_k))
        if 'conversation' in sources or 'conversation_history' in sources:
            all_results.extend(self._retrieve_conversation(query, top_k))
        logger.info(f"Retrieved {len(all_results)} total results from {len(sources)} sources")
        return all_results

    def _retrieve_archival(self, query: str, top_k: int) -> List[Dict[str, Any]]:
        try:
            results = self.memory_manager.archival_memory_search(query, top_k=top_k)
            for result in results:
                result['source'] = 'archival_memory'
            return results
        except Exception as e:  # noqa: BLE001
            logger.error(f"Archival retrieval failed: {e}")
            return []

    def _retrieve_documents(self, query: str, top_k: int) -> List[Dict[str, Any]]:
        try:
            results = self.document_store.search(agent_id=self.agent_id, query=query, top_k=top_k)
            for result in results:
                result['source'] = 'documents'
            return results
        except Exception as e:  # noqa: BLE001
            logger.error(f"Document retrieval failed: {e}")
            return []

    def _retrieve_conversation(self, query: str, top_k: int) -> List[Dict[str, Any]]:
        try:
            conversation_id = self.memory_manager.agent_id
            results = self.memory_manager.conversation_search(conversation_id, query, limit=top_k)
            return [{'content': m['content'], 'source': 'conversation_history',
                     'metadata': {'role': m['role'], 'timestamp': m.get('created_at', '')}, 'similarity': 0.5}
                    for m in results]
        except Exception as e:  # noqa: BLE001
            logger.error(f"Conversation retrieval failed: {e}")
            return []

    # ------------------------------------------------------------------ the reference's in-memory hybrid (:214-322)
    def _semantic_scores_gpu(self, embeddings, query_embedding) -> torch.Tensor:
        emb = torch.as_tensor(np.asarray(embeddings, dtype=np.float32)).to(self.device).contiguous()
        q = torch.as_tensor(np.asarray([query_embedding], dtype=np.float32)).to(self.device).contiguous()
        return engine.CosineIndex(emb, mode="exact").dense(q)[0].contiguous()

    def _bm25_index(self, corpus: List[str]):
        vocab = TextVocab()
        toks = [vocab.encode_doc(doc) for doc in corpus]
        off = np.zeros(len(corpus) + 1, dtype=np.int64)
        np.cumsum([len(t) for t in toks], out=off[1:])
        flat = np.concatenate(toks).astype(np.int32) if off[-1] > 0 else np.zeros(0, dtype=np.int32)
        ix = Bm25Index(torch.from_numpy(off).to(self.device), torch.from_numpy(flat).to(self.device),
                       max(len(vocab), 1), tile_docs=1024)
        return ix, vocab

    def _keyword_scores_gpu(self, query: str, corpus: List[str]) -> torch.Tensor:
        n = len(corpus)
        if not corpus or all(len(doc.split()) == 0 for doc in corpus):
            logger.warning("BM25: Empty or whitespace-only corpus, returning zeros")
            return torch.zeros(n, dtype=torch.float64, device=self.device)
        ix, vocab = self._bm25_index(corpus)
        terms = vocab.encode_query(query)
        terms = terms[terms >= 0]   # out-of-vocabulary tokens add an exact 0.0; any length is fine (orag_bm25_dense)
        qt = torch.from_numpy(terms if len(terms) else np.full(1, -1, np.int32)).to(self.device)[None, :].contiguous()
        ql = torch.tensor([len(terms)], dtype=torch.int32, device=self.device)
        raw = ix.dense_scores(qt, ql)[0].contiguous()
        _, top, _ = engine.dense_topk(raw[None, :], 1)  # max(scores), on the GPU
        m = float(top[0, 0].item()) if n > 0 else 0.0
        max_score = m if m > 0 else 1.0
        return engine.div_scalar(raw, max_score)  # `s / max_score`, one IEEE division per element

    def _temporal_scores(self, n: int, documents_metadata) -> Optional[np.ndarray]:
        if not (documents_metadata and self.enable_temporal_boost):
            return None
        current_time = self._now()
        out = []
        for metadata in documents_metadata:
            timestamp = metadata.get('created_at') or metadata.get('uploaded_at')
            score = 0.0
            if timestamp:
                if isinstance(timestamp, str):
                    try:
                        timestamp = datetime.fromisoformat(timestamp.replace('Z', '+00:00'))
                    except ValueError:
                        timestamp = None
                if timestamp:
                    days_old = (current_time - timestamp).total_seconds() / 86400
                    score = self.recency_weight * (0.5 ** (days_old / self.recency_half_life_days))
            out.append(score)
        return np.asarray(out, dtype=np.float64)

    @_gpu_locked
    def hybrid_search(self, query: str, corpus: List[str], embeddings: List[List[float]],
                      query_embedding: List[float], top_k: int = 10,
                      documents_metadata: Optional[List[Dict[str, Any]]] = None,
                      query_intent: Optional[str] = None) -> List[Dict[str, Any]]:
        if self.use_adaptive_weights and query_intent:
            alpha, beta, gamma = self.get_weights_for_intent(query_intent)
        else:
            alpha, beta, gamma = self.alpha, self.beta, self.gamma
        n = len(corpus)
        if n == 0:
            return []
        sem = self._semantic_scores_gpu(embeddings, query_embedding)
        kw = self._keyword_scores_gpu(query, corpus)
        temp_np = self._temporal_scores(n, documents_metadata)
        temp = torch.from_numpy(temp_np).to(self.device) if temp_np is not None else None
        if temp is not None and temp.numel() != n:
            raise ValueError("documents_metadata must have one entry per corpus item")
        hyb = engine.weighted_sum3(sem, kw, temp, alpha, beta, gamma)
        k = min(top_k, n)
        ids, vals, _ = engine.dense_topk(hyb[None, :].contiguous(), k)
        ids = ids[0].cpu().tolist()
        vals = vals[0].cpu().tolist()
        sem_h, kw_h = sem.cpu().tolist(), kw.cpu().tolist()
        ranked = []
        for i, h in zip(ids, vals):
            if i < 0:
                continue
            result = {'content': corpus[i], 'hybrid_score': h, 'semantic_score': sem_h[i], 'keyword_score': kw_h[i],
                      'temporal_score': float(temp_np[i]) if temp_np is not None else 0.0,
                      'embedding': embeddings[i]}
            if documents_metadata and i < len(documents_metadata):
                result['metadata'] = documents_metadata[i]
            ranked.append(result)
        return ranked

    # ------------------------------------------------------------------ per-piece helpers (:324-371)
    def _bm25_scores(self, query: str, corpus: List[str]) -> List[float]:
        return self._keyword_scores_gpu(query, corpus).cpu().tolist()

    def _simple_keyword_scores(self, query: str, corpus: List[str]) -> List[float]:
        """Host-side set overlap, kept for interface parity (the reference's no-rank_bm25 fallback)."""
        query_terms = set(query.lower().split())
        return [len(query_terms & set(doc.lower().split())) / len(query_terms) if query_terms else 0.0
                for doc in corpus]

    def _cosine_similarity(self, vec1: List[float], vec2: List[float]) -> float:
        return float(self._semantic_scores_gpu([vec2], vec1)[0].item())
