"""CPU: the oracle (oracle/oracle.c via ctypes) reproduces every golden vector recorded
from the reference's own functions (tests/golden/make_golden.py), bit for bit."""
import numpy as np
import pytest

import oracle
from optimized_rag_b200 import synthetic as syn

from conftest import fromhex


def _cos_inputs(case):
    corpus = syn.embeddings(syn.SEED_CORPUS, 0, case["n"], case["dim"], case["dup_per_mille"])
    queries = syn.query_embeddings(case["n_queries"], case["n"], case["dim"], dup_per_mille=case["dup_per_mille"])
    if case.get("zero_row") is not None:
        corpus[case["zero_row"], :] = 0.0
    return corpus, queries


def test_cosine_bit_exact(golden):
    for case in golden["cosine"]:
        corpus, queries = _cos_inputs(case)
        if case["name"] == "zero_query":
            got = oracle.cosine_scores(corpus, np.zeros(case["dim"], np.float32))
            want = np.array([fromhex(x) for x in case["zero_query_scores"]])
            assert np.array_equal(got, want) and np.all(got == 0.0)
            continue
        for b, row in enumerate(case["scores"]):
            want = np.array([fromhex(x) for x in row])
            got = oracle.cosine_scores(corpus, queries[b])
            assert np.array_equal(got.view(np.uint64), want.view(np.uint64)), case["name"]
            if case.get("zero_row") is not None:
                assert got[case["zero_row"]] == 0.0


def test_cosine_plain_sum_mode_differs_only_in_last_bits(golden):
    case = golden["cosine"][0]
    corpus, queries = _cos_inputs(case)
    a = oracle.cosine_scores(corpus, queries[0], neumaier=True)
    b = oracle.cosine_scores(corpus, queries[0], neumaier=False)
    assert np.allclose(a, b, rtol=1e-12, atol=0)


def test_topk_tie_rule():
    s = np.array([0.5, 0.9, 0.5, 0.9, 0.1, 0.5])
    ids, vals = oracle.topk(s, 4)
    assert ids.tolist() == [1, 3, 0, 2] and vals.tolist() == [0.9, 0.9, 0.5, 0.5]
    ids, _ = oracle.topk(s, 10, id_base=100)
    assert ids.tolist() == [101, 103, 100, 102, 105, 104]
    # equals Python's stable sorted(reverse=True)
    rng = np.random.default_rng(0)
    s = rng.integers(0, 5, 200).astype(np.float64)
    want = sorted(range(200), key=lambda i: s[i], reverse=True)[:17]
    assert oracle.topk(s, 17)[0].tolist() == want


def test_bm25_bit_exact(golden):
    for case in golden["bm25"]["cases"]:
        thr = syn.zipf_thresholds(case["vocab"])
        doc_off, tok = syn.token_corpus(syn.SEED_TOKENS, 0, case["n"], case["vocab"], case["lmin"], case["lmax"], thr)
        ix = oracle.BM25Index(doc_off, tok, case["vocab"])
        for q in case["queries"]:
            want = np.array([fromhex(x) for x in q["normalized"]])
            got, _ = ix.scores(q["terms"])
            assert np.array_equal(got.view(np.uint64), want.view(np.uint64)), (case["name"], q["terms"])


def test_bm25_glue_edges(golden):
    e = golden["bm25"]["edge"]
    assert e["empty_corpus"] == [] and e["whitespace_corpus"] == [0.0, 0.0, 0.0]
    # no_match: all-zero raw scores -> max_score falls back to 1.0
    corpus = [[0, 1, 2], [1, 2], [3]]
    off = np.cumsum([0] + [len(c) for c in corpus])
    ix = oracle.BM25Index(off, np.concatenate(corpus), 5)
    got, m = ix.scores([-1])
    assert m == 1.0 and got.tolist() == [fromhex(x) for x in e["no_match"]]
    # case_fold: "T1 t4" over ["t1 T2 t3","t2 t3 t9 t9","T4 t1 t1","t5"] lower-cased
    vocab = {"t1": 0, "t2": 1, "t3": 2, "t9": 3, "t4": 4, "t5": 5}
    docs = [["t1", "t2", "t3"], ["t2", "t3", "t9", "t9"], ["t4", "t1", "t1"], ["t5"]]
    off = np.cumsum([0] + [len(c) for c in docs])
    ix = oracle.BM25Index(off, np.array([vocab[w] for d in docs for w in d]), 6)
    got, _ = ix.scores([vocab["t1"], vocab["t4"]])
    assert got.tolist() == [fromhex(x) for x in e["case_fold"]]


def test_rrf_bit_exact(golden):
    for case in golden["rrf"]:
        ids, sc = oracle.rrf_fuse(case["lists"], case["k"], case["top_k"])
        assert ids.tolist() == case["ids"], case["name"]
        assert sc.tolist() == [fromhex(x) for x in case["scores"]], case["name"]


def test_rrf_survey_kat():
    ids, sc = oracle.rrf_fuse([[0, 1, 2], [3, 1]])
    assert ids.tolist() == [1, 0, 3, 2]
    assert sc.tolist() == [0.03225806451612903, 0.01639344262295082, 0.01639344262295082, 0.015873015873015872]
    ids2, _ = oracle.rrf_fuse([[5, 1, 2], [3, 1]], tie="chunk_id")
    assert ids2.tolist() == [1, 3, 5, 2]


def test_weighted_hybrid(golden):
    g = golden["weighted"]
    thr = syn.zipf_thresholds(g["vocab"])
    corpus = syn.embeddings(syn.SEED_CORPUS, 0, g["n"], g["dim"])
    queries = syn.query_embeddings(3, g["n"], g["dim"])
    doc_off, tok = syn.token_corpus(syn.SEED_TOKENS, 0, g["n"], g["vocab"], g["lmin"], g["lmax"], thr)
    # make_golden appended a unique token " u{i}" to every chunk: vocab ids vocab+i
    lens = np.diff(doc_off)
    toks2, off2 = [], [0]
    for i in range(g["n"]):
        toks2.extend(tok[doc_off[i]:doc_off[i + 1]].tolist() + [g["vocab"] + i])
        off2.append(off2[-1] + int(lens[i]) + 1)
    ix = oracle.BM25Index(np.array(off2), np.array(toks2), g["vocab"] + g["n"])
    weights = {None: (0.55, 0.35, 0.10), "search": (0.45, 0.50, 0.05), "Multi Hop Reasoning": (0.60, 0.30, 0.10)}
    for b, case in enumerate(g["cases"]):
        sem = oracle.cosine_scores(corpus, queries[b])
        kw, _ = ix.scores(case["terms"])
        a, be, ga = weights[case["intent"]]
        hyb = oracle.weighted_hybrid(sem, kw, None, a, be, ga)
        ids, vals = oracle.topk(hyb, 10)
        assert ids.tolist() == case["ids"]
        assert vals.tolist() == [fromhex(x) for x in case["hybrid"]]
        assert sem[ids].tolist() == [fromhex(x) for x in case["semantic"]]
        assert kw[ids].tolist() == [fromhex(x) for x in case["keyword"]]


def _pairwise_emb(g):
    m, d = g["m"], g["dim"]
    emb = syn.embeddings(syn.SEED_CORPUS, 0, m, d)
    for i in range(0, m, 4):
        emb[i + 1] = (emb[i] + np.float32(0.25) * emb[i + 1]).astype(np.float32)
        emb[i + 3] = (emb[i] + np.float32(0.5) * emb[i + 3]).astype(np.float32)
    return emb


def test_pairwise(golden):
    import hashlib
    g = golden["pairwise"]
    emb = _pairwise_emb(g)
    assert hashlib.sha256(emb.tobytes()).hexdigest() == g["emb_sha256"]
    oi, oj, sim = oracle.pairwise_candidates(emb, g["doc_idx"], g["threshold"])
    got = [[int(i), int(j), round(float(s), 3)] for i, j, s in zip(oi, oj, sim)]
    assert got == g["pairs"] and len(got) > 0


def mmr_inputs(case):
    """Embeddings + query of one golden MMR case (shared with the GPU boundary test)."""
    emb = syn.embeddings(syn.SEED_CORPUS, 0, case["m"], case["dim"], case["dup_per_mille"])
    if case["zero_row"] is not None:
        emb[case["zero_row"], :] = 0.0
    q = syn.query_embeddings(1, case["m"], case["dim"], dup_per_mille=case["dup_per_mille"])[0]
    return emb, q


def test_mmr_golden(golden):
    """oracle.mmr_select replays what the reference's MMRDiversifier.diversify
    (rag/reranker.py:104-195) picked and scored, bit for bit (duplicates, a zero row, lambda 0 and 1, top_k > m)."""
    for case in golden["mmr"]["cases"]:
        emb, q = mmr_inputs(case)
        for run in case["runs"]:
            sel, sc = oracle.mmr_select(q, emb, run["lambda"], run["top_k"])
            assert sel == run["picked"], (case["name"], run["lambda"])
            assert sc == [fromhex(x) for x in run["mmr_scores"]], (case["name"], run["lambda"])


def dedup_inputs(case):
    """Embeddings of one golden semantic-dedup case (shared with the host-logic and GPU tests)."""
    m = case["m"]
    emb = syn.embeddings(syn.SEED_CORPUS, 0, m, case["dim"], case["dup_per_mille"])
    if case["near"]:
        for i in range(0, m - 1, 3):
            emb[i + 1] = (emb[i] + np.float32(case["near"] * (1 + i % 4)) * emb[i + 1]).astype(np.float32)
    if case["name"] == "m40_d96_near":
        emb[7, :] = 0.0
    return emb


def test_dedup_golden(golden):
    """oracle.semantic_dedup_keep keeps exactly the chunks the reference's
    Deduplicator.semantic_dedup (rag/data_wrangler.py:295-326) kept (exact duplicates, near-duplicates on both sides
    of the threshold, a zero vector, a low threshold with chains of drops)."""
    for case in golden["dedup"]["cases"]:
        kept = oracle.semantic_dedup_keep(dedup_inputs(case), case["threshold"])
        assert kept == case["kept"], case["name"]
        assert 0 < len(kept) < case["m"]


def test_config1(golden):
    g = golden["config1"]
    n, dim = g["n_chunks"], g["dim"]
    emb = np.concatenate([syn.embeddings(s, 0, 1, dim) for s in g["chunk_seeds"]], axis=0)
    toks = np.concatenate([np.array(t, dtype=np.int32) for t in g["chunk_tokens"]])
    off = np.cumsum([0] + [len(t) for t in g["chunk_tokens"]])
    ix = oracle.BM25Index(off, toks, g["vocab_size"])
    for q in g["queries"]:
        qemb = (syn.embeddings(q["query_noise_seed"], 0, 1, dim)[0]
                + np.float32(0.75) * emb[q["query_base_chunk"]]).astype(np.float32)
        ci, cv = oracle.topk(oracle.cosine_scores(emb, qemb), 10)
        bi, bv, _ = ix.topk(q["query_terms"], 10)
        fi, fv = oracle.rrf_fuse([ci, bi], 60, 10)
        assert ci.tolist() == q["cos_ids"] and cv.tolist() == [fromhex(x) for x in q["cos_scores"]]
        assert bi.tolist() == q["bm25_ids"] and bv.tolist() == [fromhex(x) for x in q["bm25_scores"]]
        assert fi.tolist() == q["rrf_ids"] and fv.tolist() == [fromhex(x) for x in q["rrf_scores"]]


@pytest.mark.skipif(not __import__("oracle.ref_loader", fromlist=["x"]).available(),
                    reason="reference tree not present (GPU box)")
def test_live_reference_agrees_with_oracle():
    """Build container only: fresh random cases straight through the reference's functions."""
    from oracle import ref_loader
    h = ref_loader.hybrid_retriever()
    rng = np.random.default_rng(123)
    a = rng.standard_normal((6, 257)).astype(np.float32)
    for i in range(5):
        want = h._cosine_similarity([float(x) for x in a[i]], [float(x) for x in a[i + 1]])
        assert oracle.cosine(a[i], a[i + 1]) == want
    # bm25 on a tiny vocabulary (many negative idfs -> epsilon branch)
    docs = [rng.integers(0, 6, rng.integers(1, 9)).tolist() for _ in range(25)]
    texts = [" ".join(f"w{t}" for t in d) for d in docs]
    off = np.cumsum([0] + [len(d) for d in docs])
    ix = oracle.BM25Index(off, np.concatenate(docs), 6)
    for q in ([0, 1], [5, 5, 2], [3], [4, 0, 4, 1]):
        want = h._bm25_scores(" ".join(f"w{t}" for t in q), texts)
        got, _ = ix.scores(q)
        assert got.tolist() == want


def test_live_reference_mmr_agrees_with_oracle():
    """Build container only: the reference's MMRDiversifier (rag/reranker.py:104-195) vs oracle.mmr_select."""
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("reference tree not present")
    mod = ref_loader.load("reranker")
    rng = np.random.default_rng(21)
    emb = rng.standard_normal((14, 48)).astype(np.float32)
    emb[5] = emb[2]                      # exact duplicate -> tie handling (first maximum wins)
    q = rng.standard_normal(48).astype(np.float32)
    for lam, k in ((0.7, 5), (0.3, 14), (1.0, 3)):
        docs = [{"content": f"d{i}", "embedding": [float(x) for x in emb[i]]} for i in range(14)]
        out = mod.MMRDiversifier(lambda_param=lam).diversify([float(x) for x in q], docs, top_k=k)
        sel, sc = oracle.mmr_select(q, emb, lam, k)
        assert [d["content"] for d in out] == [f"d{i}" for i in sel]
        assert [d["mmr_score"] for d in out] == sc
