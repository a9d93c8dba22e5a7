"""GPU tests added after the last GPU session of round 1: semantic de-duplication, the
on-disk BM25 index, MMR against the reference's golden vectors.  Their host logic is pinned on the CPU
(tests/test_host_logic.py) and the kernels they launch by the parity tests; this file runs the two together.
It sorts last on purpose."""
import json
from pathlib import Path

import numpy as np
import pytest
import torch

import oracle
from test_oracle_golden import dedup_inputs

pytestmark = pytest.mark.gpu


def test_semantic_dedup_matches_reference_golden_and_oracle(monkeypatch):
    from optimized_rag_b200 import data_wrangler
    assert torch.cuda.is_available()
    golden = json.loads((Path(__file__).parent / "golden" / "golden.json").read_text())
    for block in (2048, 16):
        monkeypatch.setattr(data_wrangler, "BLOCK_ROWS", block)
        for case in golden["dedup"]["cases"]:
            emb = dedup_inputs(case)
            chunks = [{"content": f"c{i}", "n": i} for i in range(case["m"])]
            out = data_wrangler.Deduplicator.semantic_dedup(chunks, [[float(x) for x in e] for e in emb],
                                                            threshold=case["threshold"], device="cuda:0")
            assert [c["n"] for c in out] == case["kept"], (case["name"], block)
    # a larger seeded case against the oracle: 300 chunks, every third one a near-duplicate of its predecessor
    rng = np.random.default_rng(31)
    emb = rng.standard_normal((300, 128)).astype(np.float32)
    for i in range(0, 299, 3):
        emb[i + 1] = (emb[i] + np.float32(0.1 * (1 + i % 5)) * emb[i + 1]).astype(np.float32)
    monkeypatch.setattr(data_wrangler, "BLOCK_ROWS", 64)
    chunks = [{"content": f"c{i}", "n": i} for i in range(300)]
    out = data_wrangler.Deduplicator.semantic_dedup(chunks, [[float(x) for x in e] for e in emb], 0.95, device="cuda:0")
    assert [c["n"] for c in out] == oracle.semantic_dedup_keep(emb, 0.95)


def test_bm25_index_loaded_from_disk_answers_like_the_built_one(tmp_path):
    """Bm25Index.save / load (on-disk format): a loaded index must return bit-identical
    top-k lists through both kernels (MaxScore first pass and dense)."""
    from optimized_rag_b200 import synthetic as syn
    from optimized_rag_b200.bm25_index import Bm25Index
    dev = "cuda:0"
    n, vocab = 20000, 3000
    thr = syn.zipf_thresholds(vocab)
    doc_off, tok = syn.token_corpus(syn.SEED_TOKENS, 0, n, vocab, 20, 60, thr)
    ix = Bm25Index(torch.from_numpy(doc_off).to(dev), torch.from_numpy(tok).to(dev), vocab, tile_docs=256, doc_id_base=77)
    ix.save(tmp_path / "kw")
    back = Bm25Index.load(tmp_path / "kw", device=dev)
    qt, ql = syn.keyword_queries(32, vocab, thresholds=thr)
    qt, ql = torch.from_numpy(qt).to(dev), torch.from_numpy(ql).to(dev)
    for force in (None, "dense"):
        a = ix.topk(qt, ql, 10, force=force)
        b = back.topk(qt, ql, 10, force=force)
        assert all(torch.equal(x, y) for x, y in zip(a, b)), force


def test_mmr_diversifier_reproduces_the_reference_golden():
    """MMRDiversifier on the GPU against what the reference's MMRDiversifier.diversify recorded (golden.json "mmr":
    rag/reranker.py:104-195 run on seeded embeddings, incl. duplicates, a zero row, lambda 0 / 1, top_k > m)."""
    from optimized_rag_b200.reranker import MMRDiversifier
    from test_oracle_golden import mmr_inputs
    golden = json.loads((Path(__file__).parent / "golden" / "golden.json").read_text())
    for case in golden["mmr"]["cases"]:
        emb, q = mmr_inputs(case)
        for run in case["runs"]:
            docs = [{"content": f"d{i}", "embedding": [float(x) for x in emb[i]]} for i in range(case["m"])]
            out = MMRDiversifier(lambda_param=run["lambda"], device="cuda:0").diversify(
                [float(x) for x in q], docs, top_k=run["top_k"])
            assert [int(d["content"][1:]) for d in out] == run["picked"], (case["name"], run["lambda"])
            assert [d["mmr_score"].hex() for d in out] == run["mmr_scores"], (case["name"], run["lambda"])
