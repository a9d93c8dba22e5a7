"""Sentence scoring of the context compressor on the GPU.

`ContextCompressor._score_sentences_hybrid` (rag/context_compressor.py:217-241) embeds the query and every sentence
of a document and scores each sentence as 0.7 * cosine + 0.3 * keyword overlap; the cosines are one 1 x m row.  This
module provides that method (and the three helpers it leans on: `_split_sentences` :206-215, `_cosine_similarity`
:243-263, `_score_sentence_lexical` :265-289) as a mixin with the reference's names and signatures, the cosine row
coming from one launch of the float64 kernel (orag_cosine_dense: Neumaier sums, sqrt, one multiply, one divide -- the
arithmetic of :255-263).  The compression POLICY around it (confidence bands, intent thresholds from config.py,
rag/context_compressor.py:56-204) is control plane and stays the reference's: a maintainer mixes this class in,

    class ContextCompressor(GpuSentenceScoring, rag.context_compressor.ContextCompressor): pass

and nothing else changes.  Like the reference, a failure of the semantic part falls back to the lexical score.
"""
from __future__ import annotations

import logging
import re
from typing import List, Tuple

import torch

from . import _ffi, engine

logger = logging.getLogger(__name__)

_STOP_WORDS = frozenset({'the', 'a', 'an', 'and', 'or', 'but', 'in', 'on', 'at', 'to', 'for', 'of', 'with', 'by', 'from',
                         'is', 'was', 'are', 'were', 'be', 'been', 'being'})
_WORD = re.compile(r'\b\w+\b')
_SENTENCE_BREAK = re.compile(r'[.!?]+\s+')


class GpuSentenceScoring:
    """Mixin: needs `self.embedding_service`, `self.semantic_weight`, `self.lexical_weight` (set by the reference's
    constructor, rag/context_compressor.py:41-48)."""

    device = "cuda"
    semantic_weight = 0.7
    lexical_weight = 0.3

    def _split_sentences(self, text: str) -> List[str]:
        if not text:
            return []
        return [s.strip() for s in _SENTENCE_BREAK.split(text) if len(s.strip()) > 20]

    def _cosine_row(self, query_embedding, sentence_embeddings) -> List[float]:
        dim = max([len(query_embedding)] + [len(e) for e in sentence_embeddings])
        if dim == 0:
            return [0.0] * len(sentence_embeddings)
        dim += (-dim) % 4
        pad = lambda v: list(v) + [0.0] * (dim - len(v))   # zip truncation == zero padding, bit for bit
        with _ffi.GPU_LOCK:
            dev = torch.device(self.device)
            rows = torch.tensor([pad(e) for e in sentence_embeddings], dtype=torch.float32, device=dev)
            q = torch.tensor([pad(query_embedding)], dtype=torch.float32, device=dev)
            return engine.CosineIndex(rows, mode="exact").dense(q)[0].cpu().tolist()

    def _score_sentences_hybrid(self, query: str, sentences: List[str]) -> List[Tuple[str, float]]:
        try:
            if not self.embedding_service:
                raise ValueError("Embedding service not available")
            query_embedding = self.embedding_service.generate_embedding(query)
            sentence_embeddings = self.embedding_service.generate_embeddings_batch(sentences)
            m = min(len(sentences), len(sentence_embeddings))
            semantic = self._cosine_row(query_embedding, sentence_embeddings[:m]) if m else []
            scored = []
            for sent, semantic_score in zip(sentences, semantic):
                lexical_score = self._score_sentence_lexical(query, sent)
                scored.append((sent, self.semantic_weight * semantic_score + self.lexical_weight * lexical_score))
            return scored
        except Exception as e:  # noqa: BLE001
            logger.error(f"Semantic scoring failed, falling back to lexical: {e}")
            return [(sent, self._score_sentence_lexical(query, sent)) for sent in sentences]

    def _cosine_similarity(self, vec1: List[float], vec2: List[float]) -> float:
        if min(len(vec1), len(vec2)) == 0:
            return 0.0
        return self._cosine_row(vec1, [vec2])[0]

    def _score_sentence_lexical(self, query: str, sentence: str) -> float:
        query_lower, sentence_lower = query.lower(), sentence.lower()
        query_words = set(_WORD.findall(query_lower)) - _STOP_WORDS
        sent_words = set(_WORD.findall(sentence_lower)) - _STOP_WORDS
        if not query_words:
            return 0.0
        score = len(query_words & sent_words) / len(query_words)
        if query_lower in sentence_lower:
            score += 0.2
        return min(score, 1.0)


class SentenceScorer(GpuSentenceScoring):
    """Stand-alone use of the mixin (tests, callers without the reference's compressor)."""

    def __init__(self, embedding_service, device: str | torch.device = "cuda"):
        self.embedding_service = embedding_service
        self.device = str(device)
