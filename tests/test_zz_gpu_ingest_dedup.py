"""GPU test of the ingest-side semantic de-duplication, added after the last GPU session of
round 1: its host logic is pinned on the CPU (tests/test_host_logic.py), the kernel it launches (orag_cosine_dense)
by the parity tests; this file runs the two together.  It sorts last on purpose."""
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
