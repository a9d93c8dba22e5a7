"""Inputs of the ConsistencyChecker golden case (tests/golden/make_golden.py `golden_consistency`): a handful of short
documents written for this test, and a deterministic embedder under which claims that open with the same word are
similar (cosine ~0.92 for the 0.3 noise weight, ~0.79 for 0.52) and all others are near-orthogonal."""
import hashlib

import numpy as np

from optimized_rag_b200 import synthetic as syn

DIM = 256

DOCUMENTS = [
    {"content": "Alpha reactor output is 40 megawatts at full load. Bravo pipeline is not pressurised during the night "
                "shift! This is a note. Charlie depot always ships on mondays and thursdays. Short one.",
     "source": "ops-manual"},
    {"content": "Alpha reactor output is 55 megawatts at full load? Bravo pipeline is pressurised during the night shift. "
                "In conclusion nothing else matters here. Charlie depot never ships on mondays and thursdays.",
     "source": "audit-report"},
    {"content": "Alpha reactor output was measured twice by the crew. Delta valve can be opened by hand in an emergency. "
                "There are several valves in the hall. Echo ledger records every transfer of the quarter.",
     "metadata": {"kind": "no source key"}},
    {"content": "Delta valve cannot be opened by hand in an emergency. Echo ledger records every transfer of the quarter. "
                "Foxtrot gauge reads 12.5 bar when idle",
     "source": "field-notes"},
]


def _row(tag: str) -> np.ndarray:
    seed = int.from_bytes(hashlib.sha256(tag.encode("utf-8")).digest()[:8], "little") & 0x7FFFFFFFFFFFFFFF
    return syn.embeddings(seed, 0, 1, DIM)[0]


class TopicEmbedder:
    """generate_embeddings_batch(texts) -> list of Python-float lists: base(first word) + w(text) * noise(text), all in
    fp32 (so the reference's float64 cosine over these lists is the cosine of fp32 inputs)."""

    def generate_embeddings_batch(self, texts):
        out = []
        for t in texts:
            w = np.float32(0.3 if len(t) % 3 else 0.52)
            v = (_row("topic:" + t.split()[0].lower()) + w * _row("text:" + t)).astype(np.float32)
            out.append([float(x) for x in v])
        return out
