"""ConsistencyChecker drop-in (optimized_rag_b200/consistency_checker.py) against golden vectors recorded from the
live reference (rag/consistency_checker.py via tests/golden/make_golden.py `golden_consistency`).  The host text logic
and the early exits run on CPU; the pair search itself needs the GPU (-m gpu)."""
import sys
from pathlib import Path

import numpy as np
import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent))
import consistency_fixture as fx  # noqa: E402


def _checker(thr=0.85, device="cuda"):
    from optimized_rag_b200.consistency_checker import ConsistencyChecker
    return ConsistencyChecker(fx.TopicEmbedder(), similarity_threshold=thr, device=device)


def test_host_text_logic_matches_reference(golden):
    g = golden["consistency"]
    chk = _checker()
    for case in g["cases"]:
        claims = [chk._extract_claims(d["content"]) for d in fx.DOCUMENTS]
        assert claims == case["claims"]
        flat = [c for cl in claims for c in cl]
        assert [[int(chk._is_contradiction(a, b)) for b in flat] for a in flat] == case["is_contradiction"]
    # early exits never reach the GPU (rag/consistency_checker.py:47-53, 68-74)
    assert chk.check_consistency(fx.DOCUMENTS[:1], "q") == g["single_doc"]
    assert chk.check_consistency([{"content": "Tiny."}, {"content": "Alpha reactor output is 40 megawatts"}], "q") \
        == g["few_claims"]
    assert chk._generate_warning([1]) .startswith("Warning: Found 1 potential")
    assert "Please verify" in chk._generate_warning([1, 2, 3]) and "High uncertainty" in chk._generate_warning([1] * 4)


def test_fails_open_like_the_reference_when_the_pair_search_raises():
    class Broken:
        def generate_embeddings_batch(self, texts):
            return [[float("nan")] * 8 for _ in texts]

    from optimized_rag_b200.consistency_checker import ConsistencyChecker
    chk = ConsistencyChecker(Broken(), device="cpu")      # tensors on the CPU: the CUDA-only search must refuse them
    res = chk.check_consistency(fx.DOCUMENTS, "q")
    assert res["consistent"] is True and res["confidence"] == 0.5 and res["warning"].startswith("Consistency check error")


@pytest.mark.gpu
def test_check_consistency_golden(golden):
    g = golden["consistency"]
    for case in g["cases"]:
        chk = _checker(case["threshold"])
        assert chk.check_consistency(fx.DOCUMENTS, "what does the plant do") == case["result"]
        assert chk.check_consistency(fx.DOCUMENTS[:2], "q") == case["result_first_two"]
    a, b = fx.TopicEmbedder().generate_embeddings_batch(["Alpha one two three", "Alpha four five six seven"])
    import oracle
    assert _checker()._cosine_similarity(a, b) == oracle.cosine(np.float32(a), np.float32(b))
    assert _checker()._cosine_similarity([], a) == 0.0


@pytest.mark.gpu
def test_candidate_pairs_tensor_core_size_vs_oracle():
    """4096 claims: the tcgen05 first pass + float64 re-score returns the oracle's pairs, order and float64 bits; ragged
    rows (zip truncation in the reference) and an all-zero row included."""
    import oracle
    from optimized_rag_b200 import synthetic as syn
    m, dim = 4096, 192
    emb = syn.embeddings(syn.SEED_CORPUS, 0, m, dim)
    for t in range(0, m, 37):
        emb[t] = emb[(t * 7 + 3) % m] + np.float32(0.2 + 0.05 * (t % 9)) * emb[t]
    emb[11] = 0.0
    doc = (np.arange(m) // 5).astype(np.int32)
    rows = [[float(x) for x in r] for r in emb]
    rows[20] = rows[20][:100]                       # shorter row: the reference's zip stops at 100
    padded = emb.copy()
    padded[20, 100:] = 0.0
    wi, wj, ws = oracle.pairwise_candidates_parallel(padded, doc, 0.85)
    got = _checker().candidate_pairs(rows, doc.tolist())
    assert len(got) == len(wi) > 50
    assert [p[0] for p in got] == wi.tolist() and [p[1] for p in got] == wj.tolist()
    assert [p[2] for p in got] == ws.tolist()
