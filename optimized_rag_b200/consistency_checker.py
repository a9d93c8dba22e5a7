"""Cross-document consistency check on the GPU (BASELINE config 5).

Drop-in for `ConsistencyChecker` (rag/consistency_checker.py:15-280): same constructor, `check_consistency`,
`_extract_claims`, `_find_contradictions`, `_is_contradiction`, `_cosine_similarity`, `_generate_warning`, same
result dicts.  What moves to the GPU is the O(M^2 D) part of `_find_contradictions` (:168-176): the Python double loop
that evaluates one float64 cosine per claim pair becomes ONE all-pairs search
(`engine.PairwiseIndex.pairs`: exact float64 sweep for the handful of claims of a chat turn, tcgen05 first pass +
float64 re-score from 2048 claims up -- 65 536 claims in BASELINE config 5), which returns exactly the pairs the loop
would have passed to `_is_contradiction`, in the loop's (i, j) order, with the loop's float64 similarity.  The string
heuristics (:193-239) stay on the host and only ever see those pairs.

Embeddings are cast to fp32 at the boundary (what the document store / pgvector hold); rows of different lengths are
zero-padded, which is bit-identical to the reference's `zip` truncation (the extra products and squares are exact
zeros).  There is no CPU path for the pair search: without the CUDA library the call raises inside, and -- like the
reference, which fails open (:103-112) -- `check_consistency` then reports the error in `warning`.
"""
from __future__ import annotations

import logging
import re
from typing import Any, Dict, List

import torch

from . import _ffi, engine

logger = logging.getLogger(__name__)

_SENTENCE_END = re.compile(r'[.!?]+')
_META = tuple(re.compile(p) for p in (
    r'^(this|that|these|those|it|they)\s+(is|are|was|were)',
    r'^(here|there)\s+(is|are)',
    r'^(in conclusion|in summary|overall|finally)',
))
_NUMBER = re.compile(r'\b\d+\.?\d*\b')
# (negated form, plain form): one claim holding the first and the other the second reads as a contradiction
_NEGATIONS = (("is not", "is"), ("are not", "are"), ("was not", "was"), ("were not", "were"), ("does not", "does"),
              ("do not", "do"), ("did not", "did"), ("cannot", "can"), ("will not", "will"), ("should not", "should"),
              ("no", "yes"), ("false", "true"), ("incorrect", "correct"), ("never", "always"))


class ConsistencyChecker:
    """Checks retrieved documents for contradicting claims (reference: rag/consistency_checker.py:15)."""

    def __init__(self, embedding_service, similarity_threshold: float = 0.85, device: str | torch.device = "cuda"):
        self.embedding_service = embedding_service
        self.similarity_threshold = similarity_threshold
        self.device = torch.device(device)

    # ------------------------------------------------------------------ public entry (reference :33-112)
    def check_consistency(self, documents: List[Dict[str, Any]], query: str) -> Dict[str, Any]:
        if len(documents) < 2:
            return {"consistent": True, "contradictions": [], "confidence": 1.0, "warning": None}
        try:
            all_claims = []
            for idx, doc in enumerate(documents):
                for claim in self._extract_claims(doc.get("content", "")):
                    all_claims.append({"text": claim, "doc_idx": idx, "source": doc.get("source", f"doc_{idx}")})
            if len(all_claims) < 2:
                return {"consistent": True, "contradictions": [], "confidence": 1.0,
                        "warning": "Too few claims to check consistency"}
            contradictions = self._find_contradictions(all_claims)
            total_pairs = len(all_claims) * (len(all_claims) - 1) / 2
            consistency_score = 1.0 - min(len(contradictions) / max(total_pairs, 1), 1.0)
            result = {
                "consistent": len(contradictions) == 0 or consistency_score >= 0.8,
                "contradictions": contradictions[:5],
                "contradiction_count": len(contradictions),
                "confidence": consistency_score,
                "total_claims": len(all_claims),
                "warning": self._generate_warning(contradictions) if contradictions else None,
            }
            if contradictions:
                logger.warning(f"Consistency check found {len(contradictions)} contradictions "
                               f"(score: {consistency_score:.2f})")
            return result
        except Exception as e:  # fail open, like the reference
            logger.error(f"Consistency check failed: {e}")
            return {"consistent": True, "contradictions": [], "confidence": 0.5,
                    "warning": f"Consistency check error: {str(e)}"}

    # ------------------------------------------------------------------ host text logic (reference :114-146, 193-239)
    def _extract_claims(self, text: str) -> List[str]:
        claims = []
        for sent in _SENTENCE_END.split(text):
            sent = sent.strip()
            if len(sent) < 20:
                continue
            low = sent.lower()
            if any(p.match(low) for p in _META):
                continue
            claims.append(sent)
        return claims

    def _is_contradiction(self, text1: str, text2: str) -> bool:
        a, b = text1.lower(), text2.lower()
        for neg, pos in _NEGATIONS:
            if (neg in a and pos in b) or (pos in a and neg in b):
                return True
        n1, n2 = _NUMBER.findall(text1), _NUMBER.findall(text2)
        return bool(n1 and n2 and set(n1) != set(n2))

    # ------------------------------------------------------------------ the pair search (reference :148-191)
    def candidate_pairs(self, embeddings, doc_idx) -> List[tuple]:
        """[(i, j, float64 cosine)] for every i < j with doc_idx[i] != doc_idx[j] and cosine >= similarity_threshold,
        in (i, j) order -- the pairs the reference's double loop hands to `_is_contradiction`."""
        m = len(embeddings)
        if m < 2:
            return []
        dim = max((len(e) for e in embeddings), default=0)
        if dim == 0:
            # every vector empty: every cosine is 0.0 (rag/consistency_checker.py:257-259)
            sims = [(i, j, 0.0) for i in range(m) for j in range(i + 1, m) if doc_idx[i] != doc_idx[j]]
            return [p for p in sims if p[2] >= self.similarity_threshold]
        with _ffi.GPU_LOCK:
            if isinstance(embeddings, torch.Tensor):
                emb = embeddings.to(device=self.device, dtype=torch.float32).contiguous()
            else:
                rows = [list(e) + [0.0] * (dim - len(e)) for e in embeddings]
                emb = torch.tensor(rows, dtype=torch.float32, device=self.device)
            if emb.shape[1] % 4:  # the kernels read rows in 16-byte pieces: zero columns change nothing
                emb = torch.nn.functional.pad(emb, (0, 4 - emb.shape[1] % 4)).contiguous()
            doc = torch.as_tensor(list(doc_idx) if not isinstance(doc_idx, torch.Tensor) else doc_idx,
                                  dtype=torch.int32).to(self.device).contiguous()
            i, j, sim = engine.PairwiseIndex(emb, doc).pairs(float(self.similarity_threshold),
                                                             cap=max(1 << 20, 4 * m))
            return list(zip(i.cpu().tolist(), j.cpu().tolist(), sim.cpu().tolist()))

    def _find_contradictions(self, claims: List[Dict[str, Any]]) -> List[Dict[str, Any]]:
        claim_texts = [c["text"] for c in claims]
        try:
            embeddings = self.embedding_service.generate_embeddings_batch(claim_texts)
        except Exception as e:
            logger.error(f"Failed to compute embeddings: {e}")
            return []
        contradictions = []
        for i, j, similarity in self.candidate_pairs(embeddings, [c["doc_idx"] for c in claims]):
            if self._is_contradiction(claims[i]["text"], claims[j]["text"]):
                contradictions.append({
                    "claim_1": claims[i]["text"][:200],
                    "claim_2": claims[j]["text"][:200],
                    "source_1": claims[i]["source"],
                    "source_2": claims[j]["source"],
                    "similarity": round(similarity, 3),
                    "type": "semantic_contradiction",
                })
        return contradictions

    def _cosine_similarity(self, vec1: List[float], vec2: List[float]) -> float:
        """One pair through the same kernel (API completeness; the pair search above never calls it)."""
        n = min(len(vec1), len(vec2))
        if n == 0:
            return 0.0
        dim = max(len(vec1), len(vec2))
        with _ffi.GPU_LOCK:
            rows = torch.tensor([list(vec1) + [0.0] * (dim - len(vec1)), list(vec2) + [0.0] * (dim - len(vec2))],
                                dtype=torch.float32, device=self.device)
            if dim % 4:
                rows = torch.nn.functional.pad(rows, (0, 4 - dim % 4)).contiguous()
            return float(engine.CosineIndex(rows[1:].contiguous(), mode="exact").dense(rows[:1].contiguous())[0, 0].item())

    def _generate_warning(self, contradictions: List[Dict[str, Any]]) -> str:
        count = len(contradictions)
        if count == 1:
            return "Warning: Found 1 potential contradiction in sources. Response may be unreliable."
        if count <= 3:
            return f"Warning: Found {count} contradictions in sources. Please verify information."
        return f"Warning: Found {count} contradictions in sources. High uncertainty in response."

