"""GPU-resident archival memory: the cosine top-k kernel over a second table.

Mirrors the archival half of `MemoryManager` (memory/manager.py:196-295) and the SQL it runs
(`DatabaseOperations.search_archival_memory`, database/operations.py:109-159:
`SELECT id, content, metadata, 1 - (embedding <=> q) AS similarity, created_at ... ORDER BY embedding <=> q LIMIT k`),
so that `HybridRetriever.retrieve(sources=["archival", "documents"])` (rag/retrieval.py:142-144, 158-170) stays on
the GPU path end to end.  Same signatures, result keys and error behaviour (insert/search raise; the retriever
catches).  Similarities are the reference's float64 cosine (exact: not pgvector's HNSW approximation), ordered by
(similarity desc, id asc).
"""
from __future__ import annotations

import logging
from datetime import datetime, timezone
from typing import Any, Callable, Dict, List, Optional

import numpy as np
import torch

from . import engine
from .document_store import _gpu_locked

logger = logging.getLogger(__name__)

ARCHIVAL_SEARCH_RESULTS = 5  # memory/manager.py default (config.ARCHIVAL_SEARCH_RESULTS)


class GpuArchivalMemory:
    """`memory_manager` stand-in for the archival tier of one agent."""

    def __init__(self, agent_id: str, embedding_service, device: str | torch.device = "cuda",
                 now: Callable[[], datetime] = lambda: datetime.now(timezone.utc)):
        self.agent_id = agent_id
        self.embeddings = embedding_service
        self.device = torch.device(device)
        self.dim = self.embeddings.get_embedding_dimension()
        self._now = now
        self._records: List[Dict[str, Any]] = []
        self._emb = torch.empty((0, self.dim), dtype=torch.float32, device=self.device)
        self._n = 0
        self._next_id = 1
        self._index: Optional[engine.CosineIndex] = None

    def __len__(self):
        return self._n

    # ------------------------------------------------------------------ memory/manager.py:196-247
    @_gpu_locked
    def archival_memory_insert(self, content: str, metadata: Optional[Dict[str, Any]] = None) -> int:
        embedding = self.embeddings.generate_embedding(content)  # raises ValueError on empty text, like the reference
        vec = np.ascontiguousarray(embedding, dtype=np.float32)
        if vec.shape != (self.dim,) or not np.isfinite(vec).all():
            raise ValueError("invalid embedding for archival memory")
        if self._n == self._emb.shape[0]:
            grown = torch.empty((max(256, 2 * self._emb.shape[0]), self.dim), dtype=torch.float32, device=self.device)
            grown[:self._n] = self._emb[:self._n]
            self._emb = grown
        self._emb[self._n] = torch.from_numpy(vec).to(self.device)
        record_id = self._next_id
        self._next_id += 1
        self._records.append({"id": record_id, "content": content, "metadata": dict(metadata or {}),
                              "created_at": self._now()})
        self._n += 1
        self._index = None
        logger.info(f"Inserted archival memory {record_id} for agent {self.agent_id}")
        return record_id

    # ------------------------------------------------------------------ memory/manager.py:249-295
    @_gpu_locked
    def archival_memory_search(self, query: str, top_k: int = ARCHIVAL_SEARCH_RESULTS) -> List[Dict[str, Any]]:
        query_embedding = self.embeddings.generate_embedding(query)
        if self._n == 0 or top_k <= 0:
            return []
        if self._index is None:
            self._index = engine.CosineIndex(self._emb[:self._n].contiguous(), mode="auto", shadow=True)
        q = torch.tensor([query_embedding], dtype=torch.float32, device=self.device)
        ids, scores = self._index.topk(q, min(top_k, self._n))
        out = []
        for i, s in zip(ids[0].cpu().tolist(), scores[0].cpu().tolist()):
            if i < 0:
                continue
            r = self._records[i]
            # fresh dicts: the retriever mutates results in place (rag/retrieval.py:164-165)
            out.append({"id": r["id"], "content": r["content"], "metadata": dict(r["metadata"]),
                        "similarity": float(s), "created_at": r["created_at"]})
        logger.info(f"Archival search returned {len(out)} results for agent {self.agent_id}")
        return out

    # ------------------------------------------------------------------ database/operations.py:161-175
    @_gpu_locked
    def delete_archival_memory(self, memory_id: int) -> bool:
        keep = [i for i, r in enumerate(self._records) if r["id"] != memory_id]
        if len(keep) == self._n:
            return False
        idx = torch.tensor(keep, dtype=torch.int64, device=self.device)
        self._emb = self._emb[:self._n][idx].contiguous() if keep else \
            torch.empty((0, self.dim), dtype=torch.float32, device=self.device)
        self._records = [self._records[i] for i in keep]
        self._n = len(keep)
        self._index = None
        return True
