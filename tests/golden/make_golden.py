#!/usr/bin/env python
"""Generate golden vectors by running the REFERENCE'S OWN functions.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference has no tests or fixtures of its own, so parity is
pinned to the outputs recorded here: every expected value below is produced by
code imported from /root/reference (rag/retrieval.py, rag/reranker.py,
rag/consistency_checker.py, rag/chunking.py), under this container's CPython
(3.12: ``sum()`` is Neumaier-compensated).  BM25 goes through the reference's
``HybridRetriever._bm25_scores`` glue with oracle/rank_bm25.py standing in for
the absent third-party wheel (see that file's header).

Inputs are regenerated from seeds by optimized_rag_b200/synthetic.py, so the
fixtures hold parameters + expected outputs only (floats as C99 hex strings for
bit-exactness).  No reference source or document text is copied into the repo:
the config-1 fixture stores token ids and hash seeds of the chunks, not text.
"""
from __future__ import annotations

import hashlib
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle import ref_loader  # noqa: E402
from optimized_rag_b200 import synthetic as syn  # noqa: E402

OUT = Path(__file__).resolve().parent


def hx(x) -> str:
    return float(x).hex()


def py_list(a):
    return [float(v) for v in a]


def golden_cosine():
    h = ref_loader.hybrid_retriever()
    cases = []
    for name, n, d, nq, dup in [("d1536_n48", 48, 1536, 3, 0), ("d64_n200_dups", 200, 64, 4, 50),
                                ("d96_n33", 33, 96, 2, 0)]:
        corpus = syn.embeddings(syn.SEED_CORPUS, 0, n, d, dup)
        queries = syn.query_embeddings(nq, n, d, dup_per_mille=dup)
        if name == "d96_n33":
            corpus[5, :] = 0.0  # zero-norm row -> 0.0 (rag/retrieval.py:368-369)
        scores = []
        for q in queries:
            ql = py_list(q)
            scores.append([hx(h._cosine_similarity(ql, py_list(r))) for r in corpus])
        cases.append({"name": name, "n": n, "dim": d, "n_queries": nq, "dup_per_mille": dup,
                      "zero_row": 5 if name == "d96_n33" else None, "scores": scores})
    # zero query
    corpus = syn.embeddings(syn.SEED_CORPUS, 0, 8, 32)
    z = [0.0] * 32
    cases.append({"name": "zero_query", "n": 8, "dim": 32, "n_queries": 0, "dup_per_mille": 0, "zero_row": None,
                  "zero_query_scores": [hx(h._cosine_similarity(z, py_list(r))) for r in corpus]})
    return cases


def golden_bm25():
    h = ref_loader.hybrid_retriever()
    assert h.bm25_available
    cases = []
    for name, n, vocab, lmin, lmax, nq in [("v50_n60", 60, 50, 5, 40, 12), ("v2000_n300", 300, 2000, 20, 120, 16),
                                           ("v8_n5", 5, 8, 1, 6, 6)]:
        thr = syn.zipf_thresholds(vocab)
        doc_off, tok = syn.token_corpus(syn.SEED_TOKENS, 0, n, vocab, lmin, lmax, thr)
        texts = [syn.tokens_to_text(tok[doc_off[i]:doc_off[i + 1]]) for i in range(n)]
        qtok, qlen = syn.keyword_queries(nq, vocab, min_rank=min(3, vocab - 1), thresholds=thr)
        qs = []
        for b in range(nq):
            terms = [int(t) for t in qtok[b, :qlen[b]]]
            text = syn.tokens_to_text(terms)
            norm = h._bm25_scores(text, texts)
            qs.append({"terms": terms, "normalized": [hx(v) for v in norm]})
        cases.append({"name": name, "n": n, "vocab": vocab, "lmin": lmin, "lmax": lmax, "queries": qs})
    # glue edge cases (rag/retrieval.py:329-331, 344)
    edge = {
        "empty_corpus": h._bm25_scores("t1 t2", []),
        "whitespace_corpus": h._bm25_scores("t1", ["  ", "\t\n", ""]),
        "no_match": [hx(v) for v in h._bm25_scores("zzz", ["t1 t2 t3", "t2 t3", "t4"])],
        "case_fold": [hx(v) for v in h._bm25_scores("T1 t4", ["t1 T2 t3", "t2 t3 t9 t9", "T4 t1 t1", "t5"])],
    }
    return {"cases": cases, "edge": edge}


def golden_rrf():
    rr = ref_loader.rrf
    out = []

    def run(name, lists, k=60, top_k=10):
        dl = [[{"content": f"c{i}"} for i in l] for l in lists]
        res = rr(k).fuse(dl, top_k=top_k)
        out.append({"name": name, "lists": lists, "k": k, "top_k": top_k,
                    "ids": [int(d["content"][1:]) for d in res], "scores": [hx(d["rrf_score"]) for d in res]})

    run("survey_kat", [[0, 1, 2], [3, 1]])
    run("disjoint", [[1, 2, 3, 4], [5, 6, 7, 8]])
    run("identical", [[4, 3, 2, 1], [4, 3, 2, 1]])
    run("dup_inside_list", [[7, 7, 8], [8, 9, 7]])
    run("three_lists", [[1, 2, 3, 4, 5], [5, 4, 3, 2, 1], [3, 9, 1]])
    run("truncate", [list(range(15)), list(range(14, -1, -1))], top_k=5)
    run("k1", [[1, 2, 3], [3, 2, 1]], k=1)
    run("empty_second", [[1, 2, 3], []])
    run("all_empty", [[], []])
    run("single", [[9, 8, 7]])
    rng = np.random.default_rng(7)
    for i in range(6):
        a = rng.permutation(40)[:10].tolist()
        b = rng.permutation(40)[:10].tolist()
        run(f"rand{i}", [a, b])
    return out


def golden_weighted():
    """hybrid_search (rag/retrieval.py:214-322): weighted sum + stable sort; temporal term 0."""
    h = ref_loader.hybrid_retriever()
    n, d, vocab = 40, 64, 30
    thr = syn.zipf_thresholds(vocab)
    corpus = syn.embeddings(syn.SEED_CORPUS, 0, n, d)
    queries = syn.query_embeddings(3, n, d)
    doc_off, tok = syn.token_corpus(syn.SEED_TOKENS, 0, n, vocab, 5, 30, thr)
    texts = [syn.tokens_to_text(tok[doc_off[i]:doc_off[i + 1]]) + f" u{i}" for i in range(n)]
    qtok, qlen = syn.keyword_queries(3, vocab, min_rank=2, thresholds=thr)
    embs = [py_list(r) for r in corpus]
    cases = []
    for b, intent in enumerate([None, "search", "Multi Hop Reasoning"]):
        terms = [int(t) for t in qtok[b, :qlen[b]]]
        res = h.hybrid_search(syn.tokens_to_text(terms), texts, embs, py_list(queries[b]), top_k=10,
                              query_intent=intent)
        cases.append({"intent": intent, "terms": terms,
                      "ids": [int(r["content"].rsplit(" u", 1)[1]) for r in res],
                      "hybrid": [hx(r["hybrid_score"]) for r in res],
                      "semantic": [hx(r["semantic_score"]) for r in res],
                      "keyword": [hx(r["keyword_score"]) for r in res]})
    return {"n": n, "dim": d, "vocab": vocab, "lmin": 5, "lmax": 30, "cases": cases}


def golden_pairwise():
    cc = ref_loader.load("consistency_checker")
    m, d = 40, 48
    emb = syn.embeddings(syn.SEED_CORPUS, 0, m, d, dup_per_mille=0)
    # plant near-duplicates so that some pairs clear 0.85
    for i in range(0, m, 4):
        emb[i + 1] = (emb[i] + np.float32(0.25) * emb[i + 1]).astype(np.float32)
        emb[i + 3] = (emb[i] + np.float32(0.5) * emb[i + 3]).astype(np.float32)
    doc_idx = [i // 3 for i in range(m)]

    class FakeEmb:
        def generate_embeddings_batch(self, texts):
            return [py_list(e) for e in emb]

    chk = cc.ConsistencyChecker(FakeEmb(), similarity_threshold=0.85)
    # every text holds "is not" and "is" -> _is_contradiction is always True, so the output is
    # exactly the candidate pair set of the loop at rag/consistency_checker.py:169-189
    claims = [{"text": f"claim {i} is not so", "doc_idx": doc_idx[i], "source": str(i)} for i in range(m)]
    res = chk._find_contradictions(claims)
    pairs = [[int(r["source_1"]), int(r["source_2"]), r["similarity"]] for r in res]
    return {"m": m, "dim": d, "threshold": 0.85, "doc_idx": doc_idx, "pairs": pairs,
            "emb_sha256": hashlib.sha256(emb.tobytes()).hexdigest()}


def golden_config1():
    """BASELINE config 1: reference text chunked by the reference's FixedSizeChunker(1200,150),
    deterministic synthetic embeddings, cosine list + BM25 list -> RRF top-10, all through the
    reference's own functions.  Stores token ids + hash seeds, not text."""
    ch = ref_loader.load("chunking")
    h = ref_loader.hybrid_retriever()
    rr = ref_loader.rrf(60)
    text = (ref_loader.REF_ROOT / "README.md").read_text(encoding="utf-8")
    chunks = [c["content"] for c in ch.FixedSizeChunker(1200, 150).chunk(text)]
    chunks = [c for c in chunks if c.strip()]
    # unique contents (RRF keys on content)
    assert len(set(chunks)) == len(chunks)
    vocab = {}
    tok_lists = []
    for c in chunks:
        ids = []
        for w in c.lower().split():
            if w not in vocab:
                vocab[w] = len(vocab)
            ids.append(vocab[w])
        tok_lists.append(ids)
    seeds = [int.from_bytes(hashlib.sha256(c.encode("utf-8")).digest()[:8], "little") for c in chunks]
    dim = 1536
    emb = np.concatenate([syn.embeddings(s & 0x7FFFFFFFFFFFFFFF, 0, 1, dim) for s in seeds], axis=0)
    embs = [py_list(e) for e in emb]
    query_texts = ["hybrid retrieval semantic keyword", "how does the agent manage memory",
                   "reciprocal rank fusion reranking", "install postgresql pgvector extension",
                   "the and of to a", "zzzunknownzzz hallucination"]
    out_q = []
    for qi, qt in enumerate(query_texts):
        q_ids = [vocab.get(w, -1) for w in qt.lower().split()]
        qemb = syn.embeddings(0x1234 + qi, 0, 1, dim)[0] + np.float32(0.75) * emb[(qi * 5) % len(chunks)]
        qemb = qemb.astype(np.float32)
        ql = py_list(qemb)
        cos = [h._cosine_similarity(ql, e) for e in embs]
        kw = h._bm25_scores(qt, chunks)
        cos_rank = sorted(range(len(chunks)), key=lambda i: cos[i], reverse=True)[:10]
        kw_rank = sorted(range(len(chunks)), key=lambda i: kw[i], reverse=True)[:10]
        fused = rr.fuse([[{"content": chunks[i], "id": i} for i in cos_rank],
                         [{"content": chunks[i], "id": i} for i in kw_rank]], top_k=10)
        out_q.append({"query_terms": q_ids, "query_noise_seed": 0x1234 + qi, "query_base_chunk": (qi * 5) % len(chunks),
                      "cos_ids": cos_rank, "cos_scores": [hx(cos[i]) for i in cos_rank],
                      "bm25_ids": kw_rank, "bm25_scores": [hx(kw[i]) for i in kw_rank],
                      "rrf_ids": [d["id"] for d in fused], "rrf_scores": [hx(d["rrf_score"]) for d in fused]})
    return {"n_chunks": len(chunks), "dim": dim, "vocab_size": len(vocab), "chunk_tokens": tok_lists,
            "chunk_seeds": [s & 0x7FFFFFFFFFFFFFFF for s in seeds], "queries": out_q,
            "source": "reference README.md via rag/chunking.py FixedSizeChunker(1200,150)"}


def golden_mmr():
    """MMRDiversifier.diversify (rag/reranker.py:104-195) on seeded embeddings: picked indices + mmr scores."""
    mod = ref_loader.load("reranker")
    cases = []
    for name, m, d, dup, zero_row, params in [
            ("m24_d96", 24, 96, 0, None, [(0.7, 5), (0.3, 24), (1.0, 4)]),
            ("m40_d64_dups_zero", 40, 64, 100, 11, [(0.7, 10), (0.0, 6), (0.5, 40)]),
            ("m3_d1536", 3, 1536, 0, None, [(0.7, 5)])]:
        emb = syn.embeddings(syn.SEED_CORPUS, 0, m, d, dup)
        if zero_row is not None:
            emb[zero_row, :] = 0.0
        q = syn.query_embeddings(1, m, d, dup_per_mille=dup)[0]
        runs = []
        for lam, k in params:
            docs = [{"content": f"d{i}", "embedding": py_list(emb[i])} for i in range(m)]
            out = mod.MMRDiversifier(lambda_param=lam).diversify(py_list(q), docs, top_k=k)
            runs.append({"lambda": lam, "top_k": k, "picked": [int(x["content"][1:]) for x in out],
                         "mmr_scores": [hx(x["mmr_score"]) for x in out]})
        cases.append({"name": name, "m": m, "dim": d, "dup_per_mille": dup, "zero_row": zero_row, "runs": runs})
    return {"cases": cases}


def golden_dedup():
    """Deduplicator.semantic_dedup (rag/data_wrangler.py:295-326) on seeded embeddings: indices of the survivors."""
    mod = ref_loader.load("data_wrangler")
    cases = []
    for name, m, d, dup, near, thr in [("m60_d64_dups", 60, 64, 200, 0.0, 0.95), ("m40_d96_near", 40, 96, 0, 0.2, 0.95),
                                       ("m30_d32_low_threshold", 30, 32, 0, 0.0, 0.2), ("m5_d1536", 5, 1536, 0, 0.05, 0.95)]:
        emb = syn.embeddings(syn.SEED_CORPUS, 0, m, d, dup)
        if near:
            for i in range(0, m - 1, 3):   # near-duplicates of an earlier row on both sides of the threshold
                emb[i + 1] = (emb[i] + np.float32(near * (1 + i % 4)) * emb[i + 1]).astype(np.float32)
        if name == "m40_d96_near":
            emb[7, :] = 0.0               # zero vector: cosine 0 with everything, always kept
        chunks = [{"content": f"c{i}"} for i in range(m)]
        out = mod.Deduplicator.semantic_dedup(chunks, [py_list(e) for e in emb], threshold=thr)
        cases.append({"name": name, "m": m, "dim": d, "dup_per_mille": dup, "near": near, "threshold": thr,
                      "kept": [int(c["content"][1:]) for c in out]})
    return {"cases": cases}


def golden_consistency():
    """ConsistencyChecker.check_consistency (rag/consistency_checker.py:33-112) end to end on the documents and the
    embedder of tests/consistency_fixture.py: extracted claims, per-pair heuristic verdicts, the result dicts."""
    sys.path.insert(0, str(ROOT / "tests"))
    import consistency_fixture as fx
    cc = ref_loader.load("consistency_checker")
    out = {"cases": []}
    for thr in (0.85, 0.5):
        chk = cc.ConsistencyChecker(fx.TopicEmbedder(), similarity_threshold=thr)
        claims = [chk._extract_claims(d["content"]) for d in fx.DOCUMENTS]
        res = chk.check_consistency(fx.DOCUMENTS, "what does the plant do")
        res_two = chk.check_consistency(fx.DOCUMENTS[:2], "q")
        flat = [c for cl in claims for c in cl]
        heur = [[int(chk._is_contradiction(a, b)) for b in flat] for a in flat]
        out["cases"].append({"threshold": thr, "claims": claims, "result": res, "result_first_two": res_two,
                             "is_contradiction": heur})
    chk = cc.ConsistencyChecker(fx.TopicEmbedder())
    out["single_doc"] = chk.check_consistency(fx.DOCUMENTS[:1], "q")
    out["few_claims"] = chk.check_consistency([{"content": "Tiny."}, {"content": "Alpha reactor output is 40 megawatts"}], "q")
    return out


class _HashEmbedder:
    """Embedding service stand-in shared by the helper goldens: fp32 values of a counter-based hash row of the text,
    with generate_embedding / generate_embeddings_batch returning Python-float lists (sha256 -> synthetic.embeddings)."""

    def __init__(self, dim):
        self.dim = dim

    def generate_embedding(self, text):
        seed = int.from_bytes(hashlib.sha256(text.encode("utf-8")).digest()[:8], "little") & 0x7FFFFFFFFFFFFFFF
        return py_list(syn.embeddings(seed, 0, 1, self.dim)[0])

    def generate_embeddings_batch(self, texts):
        return [self.generate_embedding(t) if t.strip() else [] for t in texts]


def golden_helpers():
    """apply_mmr / cosine_similarity of rag/nodes/helpers.py:183-290 (magnitudes via `** 0.5`)."""
    h = ref_loader.load("helpers")
    cases = []
    for name, m, d, dup, lam, k, missing in [("m12_d96", 12, 96, 0, 0.7, 5, [2, 7]), ("m30_d64_dups", 30, 64, 150, 0.3, 8, []),
                                             ("m6_d1536", 6, 1536, 0, 1.0, 3, [0]), ("m4_k9", 4, 32, 0, 0.5, 9, [])]:
        emb = syn.embeddings(syn.SEED_CORPUS, 0, m, d, dup)
        docs = [{"content": f"doc {i} text", "embedding": py_list(emb[i])} for i in range(m)]
        for i in missing:
            del docs[i]["embedding"]          # generated by the service, stored in place (:215-223)
        out = h.apply_mmr("which doc is it", docs, lam, k, _HashEmbedder(d))
        cases.append({"name": name, "m": m, "dim": d, "dup_per_mille": dup, "lambda": lam, "k": k, "missing": missing,
                      "picked": [int(x["content"].split()[1]) for x in out]})
    a, b = syn.embeddings(syn.SEED_CORPUS, 0, 2, 200)
    pairs = [(py_list(a), py_list(b)), (py_list(a), py_list(a)), (py_list(a[:50]), py_list(b)), ([0.0] * 8, py_list(b[:8])),
             ([3.0, 4.0], [4.0, 3.0])]
    return {"cases": cases, "cosine": [hx(h.cosine_similarity(x, y)) for x, y in pairs],
            "cosine_inputs": "rows 0/1 of synthetic.embeddings(SEED_CORPUS, 0, 2, 200): (a,b) (a,a) (a[:50],b) (zeros8,b[:8]) ([3,4],[4,3])"}


def golden_compressor():
    """ContextCompressor._score_sentences_hybrid / _split_sentences / _score_sentence_lexical
    (rag/context_compressor.py:206-289) on the documents of tests/consistency_fixture.py."""
    sys.path.insert(0, str(ROOT / "tests"))
    import consistency_fixture as fx
    cc = ref_loader.load("context_compressor")
    comp = cc.ContextCompressor(embedding_service=_HashEmbedder(128))
    out = []
    for query in ("Alpha reactor output megawatts", "the and of", "bravo pipeline is pressurised during the night shift"):
        for doc in fx.DOCUMENTS:
            text = doc["content"] + " Tail sentence without a final stop that is long enough"
            sents = comp._split_sentences(text)
            scored = comp._score_sentences_hybrid(query, sents)
            out.append({"query": query, "sentences": sents, "hybrid": [hx(s) for _, s in scored],
                        "lexical": [hx(comp._score_sentence_lexical(query, s)) for s in sents]})
    return {"dim": 128, "cases": out}


def main():
    assert ref_loader.available(), "needs /root/reference"
    data = {
        "python": sys.version.split()[0],
        "numpy": np.__version__,
        "cosine": golden_cosine(),
        "bm25": golden_bm25(),
        "rrf": golden_rrf(),
        "weighted": golden_weighted(),
        "pairwise": golden_pairwise(),
        "config1": golden_config1(),
        "mmr": golden_mmr(),
        "dedup": golden_dedup(),
        "consistency": golden_consistency(),
        "helpers": golden_helpers(),
        "compressor": golden_compressor(),
    }
    p = OUT / "golden.json"
    p.write_text(json.dumps(data, separators=(",", ":")))
    print(f"wrote {p} ({p.stat().st_size / 1024:.1f} KiB)")


if __name__ == "__main__":
    main()
