"""Drop-ins for the small batched-cosine consumers against golden vectors recorded from the live
reference: `apply_mmr` / `cosine_similarity` (rag/nodes/helpers.py:183-290) and the context compressor's sentence
scoring (rag/context_compressor.py:206-289).  Host logic on CPU (the kernel launch replaced by the oracle's sums), the
real launches under -m gpu."""
import hashlib
import sys
from pathlib import Path

import numpy as np
import pytest

import oracle
from optimized_rag_b200 import synthetic as syn
from conftest import fromhex

sys.path.insert(0, str(Path(__file__).resolve().parent))
import consistency_fixture as fx  # noqa: E402


class HashEmbedder:
    """The embedding stand-in of tests/golden/make_golden.py (`_HashEmbedder`)."""

    def __init__(self, dim):
        self.dim = dim

    def generate_embedding(self, text):
        seed = int.from_bytes(hashlib.sha256(text.encode("utf-8")).digest()[:8], "little") & 0x7FFFFFFFFFFFFFFF
        return [float(v) for v in syn.embeddings(seed, 0, 1, self.dim)[0]]

    def generate_embeddings_batch(self, texts):
        return [self.generate_embedding(t) if t.strip() else [] for t in texts]


def _mmr_docs(case):
    emb = syn.embeddings(syn.SEED_CORPUS, 0, case["m"], case["dim"], case["dup_per_mille"])
    docs = [{"content": f"doc {i} text", "embedding": [float(v) for v in emb[i]]} for i in range(case["m"])]
    for i in case["missing"]:
        del docs[i]["embedding"]
    return docs


def _cosine_pairs():
    a, b = syn.embeddings(syn.SEED_CORPUS, 0, 2, 200)
    f = lambda v: [float(x) for x in v]
    return [(f(a), f(b)), (f(a), f(a)), (f(a[:50]), f(b)), ([0.0] * 8, f(b[:8])), ([3.0, 4.0], [4.0, 3.0])]


def _oracle_sums(rows, device=None):
    m = len(rows)
    dots = [[oracle.dot(rows[i], rows[j]) for j in range(m)] for i in range(m)]
    return dots, [dots[i][i] for i in range(m)]


def test_apply_mmr_host_logic_replays_the_reference_golden(golden, monkeypatch):
    from optimized_rag_b200 import helpers
    monkeypatch.setattr(helpers, "_sums", _oracle_sums)
    for case in golden["helpers"]["cases"]:
        docs = _mmr_docs(case)
        out = helpers.apply_mmr("which doc is it", docs, case["lambda"], case["k"], HashEmbedder(case["dim"]))
        assert [int(x["content"].split()[1]) for x in out] == case["picked"], case["name"]
        assert all(o is docs[int(o["content"].split()[1])] for o in out)
        assert all("embedding" in docs[i] for i in case["missing"]) or len(docs) <= case["k"]
    got = [helpers.cosine_similarity(x, y) for x, y in _cosine_pairs()]
    assert got == [fromhex(v) for v in golden["helpers"]["cosine"]]
    # failure of the sums -> the reference's fallbacks
    monkeypatch.setattr(helpers, "_sums", lambda *a, **k: (_ for _ in ()).throw(RuntimeError("no device")))
    docs = _mmr_docs(golden["helpers"]["cases"][0])
    assert helpers.apply_mmr("q", docs, 0.7, 3, HashEmbedder(96)) == docs[:3]
    assert helpers.cosine_similarity([1.0], [1.0]) == 0.0


def test_sentence_split_and_lexical_score_match_the_reference(golden):
    from optimized_rag_b200.context_compressor import SentenceScorer
    sc = SentenceScorer(None, device="cpu")
    queries = ("Alpha reactor output megawatts", "the and of", "bravo pipeline is pressurised during the night shift")
    cases = iter(golden["compressor"]["cases"])
    for query in queries:
        for doc in fx.DOCUMENTS:
            case = next(cases)
            sents = sc._split_sentences(doc["content"] + " Tail sentence without a final stop that is long enough")
            assert case["query"] == query and sents == case["sentences"]
            assert [sc._score_sentence_lexical(query, s) for s in sents] == [fromhex(v) for v in case["lexical"]]
            # no embedding service -> the reference's lexical fallback
            assert sc._score_sentences_hybrid(query, sents) == [(s, fromhex(v)) for s, v in zip(sents, case["lexical"])]


@pytest.mark.gpu
def test_apply_mmr_and_cosine_similarity_golden(golden):
    from optimized_rag_b200 import helpers
    for case in golden["helpers"]["cases"]:
        docs = _mmr_docs(case)
        out = helpers.apply_mmr("which doc is it", docs, case["lambda"], case["k"], HashEmbedder(case["dim"]))
        assert [int(x["content"].split()[1]) for x in out] == case["picked"], case["name"]
    got = [helpers.cosine_similarity(x, y) for x, y in _cosine_pairs()]
    assert got == [fromhex(v) for v in golden["helpers"]["cosine"]]


@pytest.mark.gpu
def test_sentence_scoring_golden(golden):
    from optimized_rag_b200.context_compressor import SentenceScorer
    sc = SentenceScorer(HashEmbedder(golden["compressor"]["dim"]))
    for case in golden["compressor"]["cases"]:
        scored = sc._score_sentences_hybrid(case["query"], case["sentences"])
        assert [s for s, _ in scored] == case["sentences"]
        assert [v for _, v in scored] == [fromhex(x) for x in case["hybrid"]]
    a = HashEmbedder(64).generate_embedding("one")
    b = HashEmbedder(64).generate_embedding("two")
    assert sc._cosine_similarity(a, b) == oracle.cosine(np.float32(a), np.float32(b))
    assert sc._cosine_similarity([], a) == 0.0
