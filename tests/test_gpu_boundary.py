"""GPU tests of the drop-in boundary (-m gpu): the reference-facing Python classes
(DocumentStore / HybridRetriever / ReciprocalRankFusion / embedding provider) read like the
reference's own call sites and reproduce its outputs (golden vectors recorded from the reference)."""
import datetime as dt

import numpy as np
import pytest
import torch

import oracle
from optimized_rag_b200 import synthetic as syn
from conftest import fromhex

pytestmark = pytest.mark.gpu


class WordChunker:
    """Minimal chunking strategy with the reference's interface: chunk(text) -> [{'content', 'metadata'}]."""

    def __init__(self, words=12):
        self.words = words

    def chunk(self, text, metadata=None):
        w = text.split()
        return [{"content": " ".join(w[i:i + self.words]), "metadata": {"chunk_id": i // self.words}}
                for i in range(0, len(w), self.words)]


def _corpus_text(n_docs=40, seed=7):
    rng = np.random.default_rng(seed)
    vocab = [f"w{i}" for i in range(60)]
    p = 1.0 / np.arange(1, 61)
    p /= p.sum()
    return [" ".join(rng.choice(vocab, size=rng.integers(30, 90), p=p)) + f" Doc{d} unique{d}" for d in range(n_docs)]


@pytest.fixture(scope="module")
def store():
    from optimized_rag_b200.document_store import DocumentStore
    from optimized_rag_b200.embeddings import SyntheticEmbeddingService
    emb = SyntheticEmbeddingService()
    st = DocumentStore(None, emb, WordChunker(), device="cuda:0")
    for d, text in enumerate(_corpus_text()):
        r = st.upload_and_index("agent-a", f"/tmp/doc{d}.txt", file_content=text, metadata={"d": d})
        assert r["success"] and r["chunks_created"] == r["chunk_count"] > 0 and r["chunks_skipped"] == 0
    st.upload_and_index("agent-b", "/tmp/other.txt", file_content="totally different tenant text here")
    return st


def test_embedding_provider_surface():
    from optimized_rag_b200.embeddings import SyntheticEmbeddingService
    e = SyntheticEmbeddingService()
    assert e.get_embedding_dimension() == 1536
    v = e.generate_embedding("hello world")
    assert len(v) == 1536 and v == e.generate_embedding("hello world")
    with pytest.raises(ValueError):
        e.generate_embedding("   ")
    assert e.generate_embeddings_batch(["a", "", "b"])[1] == [] and e.generate_embeddings_batch([]) == []


def test_document_store_search_matches_oracle(store):
    table = store._tables["agent-a"]
    emb = table._emb[:len(table)].cpu().numpy()
    for query in ["w3 w7 unique5", "Doc11 w1", "w0 w0 w2"]:
        res = store.search("agent-a", query, top_k=5)
        q = np.asarray(store.embeddings.generate_embedding(query), dtype=np.float32)
        wi, wv = oracle.topk(oracle.cosine_scores(emb, q), 5)
        assert [r["content"] for r in res] == [table.records[i]["content"] for i in wi]
        assert [r["score"] for r in res] == wv.tolist()
        assert set(res[0]) == {"content", "filename", "file_type", "score", "metadata"}
        assert res[0]["file_type"] == ".txt" and "chunk_id" in res[0]["metadata"] and "d" in res[0]["metadata"]
    # multi-tenant isolation (WHERE dc.agent_id = %s), fresh dicts, never raises
    assert all("different tenant" not in r["content"] for r in store.search("agent-a", "tenant", 50))
    assert store.search("nobody", "w1", 5) == [] and store.search("agent-a", "", 5) == []
    a = store.search("agent-a", "w1", 2)
    a[0]["source"] = "documents"
    assert "source" not in store.search("agent-a", "w1", 2)[0]


def test_document_store_hybrid_matches_oracle(store):
    table = store._tables["agent-a"]
    n = len(table)
    emb = table._emb[:n].cpu().numpy()
    off = np.cumsum([0] + [len(t) for t in table.tokens])
    orc = oracle.BM25Index(off, np.concatenate(table.tokens), len(table.vocab))
    for query in ["w3 w7 unique5", "w1 w2 w3 w4 nosuchword", "Doc3 doc3 W9"]:
        res = store.hybrid_search("agent-a", query, top_k=5, fetch_k=10)
        q = np.asarray(store.embeddings.generate_embedding(query), dtype=np.float32)
        want = oracle.hybrid_topk(emb, q[None, :], orc, [table.vocab.encode_query(query)], k=5, fetch_k=10)[0]
        assert [r["content"] for r in res] == [table.records[i]["content"] for i in want["ids"]]
        assert [r["rrf_score"] for r in res] == want["rrf_scores"].tolist()
        cos = oracle.cosine_scores(emb, q)
        assert [r["score"] for r in res] == [cos[i] for i in want["ids"]]  # `score` stays a cosine


def test_document_store_is_thread_safe(store):
    """One store shared by several threads (the reference shares one agent across a connection pool): concurrent
    searches return exactly what serial searches return."""
    import threading
    queries = ["w3 w10 w25", "w1 w7 unique12", "w40 w41 w2 w2", "Doc5 w9"]
    want = {q: (store.search("agent-a", q, top_k=5), store.hybrid_search("agent-a", q, top_k=5)) for q in queries}
    errors = []

    def worker(q):
        try:
            for _ in range(5):
                assert store.search("agent-a", q, top_k=5) == want[q][0]
                assert store.hybrid_search("agent-a", q, top_k=5) == want[q][1]
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=(q,)) for q in queries for _ in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    assert all(len(v[0]) > 0 for v in want.values())


def test_document_store_bookkeeping(store):
    docs = store.list_documents("agent-a")
    assert len(docs) == 40 and {"id", "filename", "file_type", "quality_score", "chunk_count", "uploaded_at"} <= set(docs[0])
    before = len(store._tables["agent-a"])
    victim = docs[0]["id"]
    assert store.delete_document("agent-a", victim) is True
    assert len(store._tables["agent-a"]) == before - docs[0]["chunk_count"]
    assert all(d["id"] != victim for d in store.list_documents("agent-a"))
    assert store.search("agent-a", "w1 w2", 3)  # indices rebuilt lazily after the delete
    bad = store.upload_and_index("agent-a", "/nonexistent/file.txt")
    assert bad["success"] is False and "error" in bad


def test_rrf_dropin_golden(golden):
    from optimized_rag_b200.reranker import ReciprocalRankFusion
    for case in golden["rrf"]:
        lists = [[{"content": f"c{i}", "tag": (li, r)} for r, i in enumerate(l)] for li, l in enumerate(case["lists"])]
        res = ReciprocalRankFusion(k=case["k"], device="cuda:0").fuse(lists, top_k=case["top_k"])
        assert [int(d["content"][1:]) for d in res] == case["ids"], case["name"]
        assert [d["rrf_score"] for d in res] == [fromhex(x) for x in case["scores"]], case["name"]
        # first sighting supplies the returned dict object, mutated in place
        for d in res:
            assert any(d is x for l in lists for x in l)


def _weighted_inputs(g):
    thr = syn.zipf_thresholds(g["vocab"])
    corpus = syn.embeddings(syn.SEED_CORPUS, 0, g["n"], g["dim"])
    queries = syn.query_embeddings(3, g["n"], g["dim"])
    doc_off, tok = syn.token_corpus(syn.SEED_TOKENS, 0, g["n"], g["vocab"], g["lmin"], g["lmax"], thr)
    texts = [syn.tokens_to_text(tok[doc_off[i]:doc_off[i + 1]]) + f" u{i}" for i in range(g["n"])]
    return corpus, queries, texts


def test_hybrid_retriever_weighted_golden(golden):
    """HybridRetriever.hybrid_search == the reference's own hybrid_search output, bit for bit."""
    from optimized_rag_b200.retrieval import HybridRetriever
    g = golden["weighted"]
    corpus, queries, texts = _weighted_inputs(g)
    hr = HybridRetriever(None, None, "a", device="cuda:0")
    embs = [[float(x) for x in r] for r in corpus]
    for b, case in enumerate(g["cases"]):
        res = hr.hybrid_search(syn.tokens_to_text(case["terms"]), texts, embs, [float(x) for x in queries[b]],
                               top_k=10, query_intent=case["intent"])
        assert [int(r["content"].rsplit(" u", 1)[1]) for r in res] == case["ids"]
        assert [r["hybrid_score"] for r in res] == [fromhex(x) for x in case["hybrid"]]
        assert [r["semantic_score"] for r in res] == [fromhex(x) for x in case["semantic"]]
        assert [r["keyword_score"] for r in res] == [fromhex(x) for x in case["keyword"]]
        assert set(res[0]) == {"content", "hybrid_score", "semantic_score", "keyword_score", "temporal_score",
                               "embedding"}


def test_hybrid_retriever_temporal_and_helpers(golden):
    from optimized_rag_b200.retrieval import HybridRetriever
    g = golden["weighted"]
    corpus, queries, texts = _weighted_inputs(g)
    now = dt.datetime(2026, 1, 31, 12, 0, 0)
    hr = HybridRetriever(None, None, "a", device="cuda:0", now=lambda: now)
    meta = [{"created_at": (now - dt.timedelta(days=3 * i)).isoformat()} if i % 3 else {"x": 1} for i in range(g["n"])]
    embs = [[float(x) for x in r] for r in corpus]
    case = g["cases"][0]
    res = hr.hybrid_search(syn.tokens_to_text(case["terms"]), texts, embs, [float(x) for x in queries[0]], top_k=7,
                           documents_metadata=meta)
    # re-derive with the oracle: temporal = 0.1 * 0.5 ** (days / 30), weights 0.55 / 0.35 / 0.10
    sem = oracle.cosine_scores(corpus, queries[0])
    kw = np.array(hr._bm25_scores(syn.tokens_to_text(case["terms"]), texts))
    temp = np.array([0.1 * 0.5 ** ((3 * i) / 30) if i % 3 else 0.0 for i in range(g["n"])])
    hyb = oracle.weighted_hybrid(sem, kw, temp, 0.55, 0.35, 0.10)
    wi, wv = oracle.topk(hyb, 7)
    assert [int(r["content"].rsplit(" u", 1)[1]) for r in res] == wi.tolist()
    assert [r["hybrid_score"] for r in res] == wv.tolist()
    assert res[0]["metadata"] is meta[wi[0]]
    # glue edge cases recorded from the reference (rag/retrieval.py:329-331, 344)
    e = golden["bm25"]["edge"]
    assert hr._bm25_scores("t1 t2", []) == e["empty_corpus"]
    assert hr._bm25_scores("t1", ["  ", "\t\n", ""]) == e["whitespace_corpus"]
    assert hr._bm25_scores("zzz", ["t1 t2 t3", "t2 t3", "t4"]) == [fromhex(x) for x in e["no_match"]]
    assert hr._bm25_scores("T1 t4", ["t1 T2 t3", "t2 t3 t9 t9", "T4 t1 t1", "t5"]) == [fromhex(x) for x in e["case_fold"]]
    a, b = corpus[0], corpus[1]
    assert hr._cosine_similarity([float(x) for x in a], [float(x) for x in b]) == oracle.cosine(a, b)
    assert hr.get_weights_for_intent("Multi Hop Reasoning") == (0.60, 0.30, 0.10)
    assert hr.get_weights_for_intent("nope") == (0.55, 0.35, 0.10) and hr.bm25_available is True


def test_hybrid_retriever_dispatch(store):
    from optimized_rag_b200.retrieval import HybridRetriever

    class Mem:
        agent_id = "conv-1"

        def archival_memory_search(self, query, top_k):
            return [{"id": 1, "content": "archived", "similarity": 0.9}]

        def conversation_search(self, cid, query, limit):
            raise RuntimeError("db down")

    hr = HybridRetriever(Mem(), store, "agent-a", device="cuda:0")
    out = hr.retrieve("w1 w2", ["documents", "archival", "conversation"], top_k=3)
    assert [r["source"] for r in out] == ["archival_memory"] + ["documents"] * 3
    assert hr.retrieve("w1", [], 3) == []


def test_gpu_archival_memory_matches_oracle(store):
    """the archival tier on the same cosine kernel, result keys of
    database/operations.py:147-156, exact float64 similarities ordered (similarity desc, id asc)."""
    from datetime import datetime, timezone
    from optimized_rag_b200.archival import GpuArchivalMemory
    from optimized_rag_b200.embeddings import SyntheticEmbeddingService
    from optimized_rag_b200.retrieval import HybridRetriever
    emb = SyntheticEmbeddingService()
    t0 = datetime(2026, 1, 1, tzinfo=timezone.utc)
    mem = GpuArchivalMemory("agent-a", emb, device="cuda:0", now=lambda: t0)
    assert mem.archival_memory_search("anything", top_k=3) == []
    texts = [f"memory number {i} about topic t{i % 17} and w{i % 5}" for i in range(300)]
    ids = [mem.archival_memory_insert(t, {"i": i}) for i, t in enumerate(texts)]
    assert ids == list(range(1, 301)) and len(mem) == 300
    with pytest.raises(ValueError):
        mem.archival_memory_insert("   ")
    corpus = np.asarray([emb.generate_embedding(t) for t in texts], dtype=np.float32)
    for query in ("memory number 7 about topic t7 and w2", "something else entirely"):
        got = mem.archival_memory_search(query, top_k=6)
        q = np.asarray(emb.generate_embedding(query), dtype=np.float32)
        wi, ws = oracle.topk(oracle.cosine_scores(corpus, q), 6)
        assert [g["id"] for g in got] == [int(i) + 1 for i in wi]
        assert [g["similarity"] for g in got] == ws.tolist()
        assert set(got[0]) == {"id", "content", "metadata", "similarity", "created_at"}
        assert got[0]["created_at"] == t0 and got[0]["metadata"] == {"i": int(wi[0])}
    got[0]["metadata"]["mutated"] = True  # results are fresh dicts
    assert "mutated" not in mem.archival_memory_search("something else entirely", top_k=1)[0]["metadata"]
    # the retriever dispatch stays on the GPU path for both sources
    hr = HybridRetriever(mem, store, "agent-a", device="cuda:0")
    out = hr.retrieve("memory number 7 about topic t7 and w2", ["archival", "documents"], top_k=4)
    assert [r["source"] for r in out] == ["archival_memory"] * 4 + ["documents"] * 4
    assert out[0]["content"] == texts[7]
    assert mem.delete_archival_memory(8) and not mem.delete_archival_memory(8) and len(mem) == 299
    assert mem.archival_memory_search("memory number 7 about topic t7 and w2", top_k=1)[0]["id"] != 8


def test_document_store_save_and_load_round_trip(store, tmp_path):
    """SURVEY 8f row f2, on-disk format: a saved store reloads into identical search results (cosine + hybrid)."""
    from optimized_rag_b200.document_store import DocumentStore
    from optimized_rag_b200.embeddings import SyntheticEmbeddingService
    store.save(str(tmp_path))
    assert (tmp_path / "store.json").exists() and any(p.suffix == ".f32" for p in tmp_path.iterdir())
    st2 = DocumentStore(None, SyntheticEmbeddingService(), WordChunker(), device="cuda:0")
    st2.load(str(tmp_path))
    for q in ("w3 w10 w25", "Doc5 w9"):
        assert st2.search("agent-a", q, top_k=5) == store.search("agent-a", q, top_k=5)
        assert st2.hybrid_search("agent-a", q, top_k=5) == store.hybrid_search("agent-a", q, top_k=5)
    assert st2.list_documents("agent-a") == store.list_documents("agent-a")
    r = st2.upload_and_index("agent-a", "/tmp/new.txt", file_content="w1 w2 w3 brand new text")
    assert r["success"] and r["document_id"] == store._next_doc_id   # ids continue after a reload


def test_mmr_diversifier_matches_oracle():
    """MMR over result embeddings, one cosine-matrix launch + the reference's greedy loop."""
    from optimized_rag_b200.reranker import MMRDiversifier
    rng = np.random.default_rng(22)
    emb = rng.standard_normal((40, 96)).astype(np.float32)
    emb[7] = emb[3]
    emb[11] = 0.0                                       # zero vector -> cosine 0.0
    q = rng.standard_normal(96).astype(np.float32)
    for lam, k in ((0.7, 5), (0.25, 40), (1.0, 8)):
        docs = [{"content": f"d{i}", "embedding": [float(x) for x in emb[i]]} for i in range(40)]
        docs.insert(4, {"content": "bad", "embedding": [float("nan")] * 96})
        docs.insert(9, {"content": "none"})
        out = MMRDiversifier(lambda_param=lam, device="cuda:0").diversify([float(x) for x in q], docs, top_k=k)
        sel, sc = oracle.mmr_select(q, emb, lam, k)
        assert [d["content"] for d in out] == [f"d{i}" for i in sel]
        assert [d["mmr_score"] for d in out] == sc
    assert MMRDiversifier(device="cuda:0").diversify([1.0, 2.0], [], 3) == []
    only_bad = [{"content": "x"}, {"content": "y", "embedding": []}]
    assert MMRDiversifier(device="cuda:0").diversify([1.0, 2.0], only_bad, 1) == only_bad[:1]


# ------------------------------------------------------------------------------------------------ ingest robustness
class _FlakyEmbeddings:
    """SyntheticEmbeddingService that can be told to fail (a network error in the reference's real service)."""

    def __init__(self):
        from optimized_rag_b200.embeddings import SyntheticEmbeddingService
        self._inner = SyntheticEmbeddingService()
        self.fail = False

    def get_embedding_dimension(self):
        return self._inner.get_embedding_dimension()

    def generate_embedding(self, text):
        return self._inner.generate_embedding(text)

    def generate_embeddings_batch(self, texts):
        if self.fail:
            raise RuntimeError("embedding service unavailable")
        return self._inner.generate_embeddings_batch(texts)


def test_failed_reupload_keeps_the_old_chunks_searchable():
    """rag/document_store.py:317-389 embeds before it opens the DELETE + INSERT transaction: a re-upload that fails
    (embedding service down, wrong dimension) or that yields no chunks must leave the earlier upload intact."""
    from optimized_rag_b200.document_store import DocumentStore
    emb = _FlakyEmbeddings()
    st = DocumentStore(None, emb, WordChunker(), device="cuda:0")
    first = st.upload_and_index("a", "/tmp/report.txt", file_content="alpha beta gamma delta " * 10)
    assert first["success"] and first["chunks_created"] > 0
    before = st.search("a", "alpha beta", 3)
    assert before
    emb.fail = True
    bad = st.upload_and_index("a", "/tmp/report.txt", file_content="completely new words here " * 10)
    assert bad["success"] is False and "unavailable" in bad["error"]
    emb.fail = False
    assert st.search("a", "alpha beta", 3) == before and len(st.list_documents("a")) == 1
    empty = st.upload_and_index("a", "/tmp/report.txt", file_content="   ")
    assert empty["success"] and empty["chunk_count"] == 0 and empty["document_id"] == first["document_id"]
    assert st.search("a", "alpha beta", 3) == before              # "No chunks generated": nothing was deleted
    # a successful re-upload replaces the chunks in one step
    good = st.upload_and_index("a", "/tmp/report.txt", file_content="completely new words here " * 10)
    assert good["success"] and good["document_id"] == first["document_id"]
    after = st.search("a", "alpha beta", 3)
    assert after and all("alpha" not in r["content"] for r in after)


def test_incremental_table_equals_a_fresh_index_and_large_k(store):
    """Chunks arrive upload by upload (only new rows are converted), one document is deleted in between: the table's
    cosine view must answer exactly like an index built from scratch over the same rows -- on the tensor-core path
    (>= 4096 chunks) as well -- and top_k beyond the kernels' list limits is still served."""
    from optimized_rag_b200.document_store import DocumentStore
    from optimized_rag_b200.embeddings import SyntheticEmbeddingService
    from optimized_rag_b200 import engine
    st = DocumentStore(None, SyntheticEmbeddingService(dimensions=128), WordChunker(words=3), device="cuda:0")
    texts = _corpus_text(n_docs=300, seed=11)
    for d, text in enumerate(texts):
        assert st.upload_and_index("big", f"/tmp/f{d}.txt", file_content=text)["success"]
    assert st.delete_document("big", 7)
    table = st._tables["big"]
    assert len(table) >= engine.SMALL_N and table.cosine().mode == "f16"
    fresh = engine.CosineIndex(table._emb[:len(table)].clone(), mode="f16")
    q = torch.tensor([st.embeddings.generate_embedding("w3 w7 unique5"), st.embeddings.generate_embedding("w1 w2")],
                     dtype=torch.float32, device="cuda:0")
    a, b = table.cosine().topk(q, 10), fresh.topk(q, 10)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    emb = table._emb[:len(table)].cpu().numpy()
    wi, wv = oracle.topk(oracle.cosine_scores(emb, q[0].cpu().numpy()), 200)
    res = st.search("big", "w3 w7 unique5", top_k=200)            # > 128: the exact scan takes over
    assert [r["score"] for r in res] == wv.tolist()
    assert len(st.hybrid_search("big", "w3 w7 unique5", top_k=100)) == 100   # > 64: semantic ranking, not []


def test_queries_longer_than_64_words_through_the_drop_ins(store):
    """The reference handles any query length (rag/retrieval.py:334-341); the candidate kernels hold 64 terms, longer
    queries take the per-document dense kernel -- same float64 scores as the oracle."""
    from optimized_rag_b200.retrieval import HybridRetriever
    table = store._tables["agent-a"]
    words = [f"w{i % 37}" for i in range(90)] + ["unique5", "notinvocab"]
    query = " ".join(words)
    res = store.hybrid_search("agent-a", query, top_k=5)
    assert len(res) == 5 and all(r["keyword_score"] is None or 0.0 <= r["keyword_score"] <= 1.0 for r in res)
    off, toks = table.token_arrays()
    orc = oracle.BM25Index(off, toks, len(table.vocab))
    q_ids = table.vocab.encode_query(query)
    want, m = orc.scores(q_ids[q_ids >= 0])
    wi, wv = oracle.topk(want, 10)
    kw = {r["content"]: r["keyword_score"] for r in res if r["keyword_rank"]}
    for i, v in zip(wi, wv):
        c = table.records[i]["content"]
        if c in kw:
            assert kw[c] == v
    hr = HybridRetriever(None, store, "agent-a")
    corpus = [r["content"] for r in table.records[:50]]
    got = hr._bm25_scores(query, corpus)
    import oracle.rank_bm25 as rb
    ref = rb.BM25Okapi([d.lower().split() for d in corpus]).get_scores(query.lower().split())
    mx = max(ref) if len(ref) and max(ref) > 0 else 1.0
    assert got == [float(s / mx) for s in ref]


def test_load_adopts_the_saved_keyword_index(store, tmp_path):
    from optimized_rag_b200.document_store import DocumentStore
    store.hybrid_search("agent-a", "w3 w7", 5)      # makes sure the index exists
    store.save(str(tmp_path))
    back = DocumentStore(None, store.embeddings, WordChunker(), device="cuda:0", retrieval_mode="hybrid")
    back.load(str(tmp_path))
    t = back._tables["agent-a"]
    assert t._bm25 is not None and t._bm25.n_docs == len(t)      # adopted, not rebuilt
    a = store.hybrid_search("agent-a", "w3 w7 unique5", 5)
    b = back.hybrid_search("agent-a", "w3 w7 unique5", 5)
    assert a == b and a
