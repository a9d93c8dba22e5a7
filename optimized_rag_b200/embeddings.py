"""Deterministic synthetic stand-in for the reference's EmbeddingService (memory/embeddings.py:31-332).

The reference calls OpenAI `text-embedding-3-small` over HTTPS; there is no network here, so the same
three-method surface the retrieval path uses is provided over a hash of the text
(SURVEY.md §2 row 9, §8b):
    generate_embedding(text) -> List[float]        ValueError on empty text (memory/embeddings.py:77-78)
    generate_embeddings_batch(texts) -> List[List[float]]   empty text -> [] entry (memory/embeddings.py:165-169)
    get_embedding_dimension() -> int               1536 (memory/embeddings.py:324-325)
"""
from __future__ import annotations

import hashlib
import threading
from typing import List

import numpy as np

from . import synthetic


class SyntheticEmbeddingService:
    def __init__(self, dimensions: int = 1536, model: str = "synthetic-text-embedding-3-small", cache_size: int = 4096):
        self.dimensions = int(dimensions)
        self.model = model
        self._cache: dict = {}
        self._cache_size = cache_size
        self._cache_lock = threading.Lock()

    def get_embedding_dimension(self) -> int:
        return self.dimensions

    def embed_array(self, text: str) -> np.ndarray:
        """fp32 [dim]; a pure function of the text (sha256 -> counter-based hash row)."""
        seed = int.from_bytes(hashlib.sha256(text.encode("utf-8")).digest()[:8], "little") & 0x7FFFFFFFFFFFFFFF
        return synthetic.embeddings(seed, 0, 1, self.dimensions)[0]

    def generate_embedding(self, text: str, use_cache: bool = True) -> List[float]:
        if not text or not text.strip():
            raise ValueError("Text cannot be empty or whitespace-only")
        if use_cache:
            with self._cache_lock:
                hit = self._cache.get(text)
            if hit is not None:
                return list(hit)
        v = tuple(float(x) for x in self.embed_array(text))
        if use_cache:
            with self._cache_lock:
                if len(self._cache) >= self._cache_size:
                    self._cache.pop(next(iter(self._cache)))
                self._cache[text] = v
        return list(v)

    def generate_embeddings_batch(self, texts: List[str], use_cache: bool = True) -> List[List[float]]:
        if not texts:
            return []
        out: List[List[float]] = []
        for t in texts:
            if not t or not t.strip():
                out.append([])
            else:
                out.append(self.generate_embedding(t, use_cache))
        return out
