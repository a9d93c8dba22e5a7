"""GPU parity at the BASELINE sizes (-m gpu; runs last: it wants most of the GPU's memory).

SURVEY.md §8(d) "Parity gates": full oracle at N <= 1M, streamed oracle at 10M on >= 32 queries.
  config 2   1M x 1536 exact cosine top-10, batch 256: ALL queries, fp16 and tf32 first pass
  config 3   10M x 1536 hybrid (cosine + BM25 + RRF) top-10, batch 256: 32 queries
  config 4   BM25-only over the 10M-doc Zipf corpus, batch 1024: 32 queries
Everything is compared bit for bit (ids, ranks, float64 score bits) with the oracle's reference arithmetic
(rag/retrieval.py:362-371, 324-347, 320; rag/reranker.py:224-271), which regenerates the synthetic inputs from the
seeds (oracle.c "Streamed oracle"; pinned to the in-memory oracle by tests/test_oracle_stream.py).

ORAG_SCALE_ROWS overrides the 10M (e.g. a smaller GPU); ORAG_SCALE_TEST=0 skips the module.
"""
import os

import numpy as np
import pytest
import torch

import oracle
from oracle import scale_check
from optimized_rag_b200 import synthetic as syn

pytestmark = pytest.mark.gpu

DEV = "cuda:0"
DIM, VOCAB, K = 1536, 50000, 10
N_BIG = int(os.environ.get("ORAG_SCALE_ROWS", 10_000_000))

if os.environ.get("ORAG_SCALE_TEST", "1") == "0":
    pytest.skip("ORAG_SCALE_TEST=0", allow_module_level=True)


_subset = scale_check.pick_queries


def _np(res, keys):
    return {key: res[key].cpu().numpy() for key in keys}


@pytest.fixture(scope="module", autouse=True)
def _all_host_cores():
    oracle.set_threads(os.cpu_count() or 1)
    torch.cuda.empty_cache()
    yield
    torch.cuda.empty_cache()


@pytest.mark.parametrize("mode", ["f16", "tf32"])
def test_config2_cosine_1m_all_queries(mode):
    from optimized_rag_b200 import engine
    n, nq = 1_000_000, 256
    corpus = engine.gen_embeddings(n, DIM, 0, syn.SEED_CORPUS, 0, device=DEV)
    index = engine.CosineIndex(corpus, mode=mode)
    q = syn.query_embeddings(nq, n, DIM)
    status: list = []
    ids, sc = index.topk(torch.from_numpy(q).to(DEV), K, check_overflow=False, status_out=status)
    assert int(status[0].max().item()) == 0          # no candidate buffer overflowed: the fast path itself is on trial
    want, _ = scale_check.reference_lists(n, DIM, q, None, None, K, seed_corpus=syn.SEED_CORPUS)
    bad = scale_check.compare({"cos_ids": ids.cpu().numpy(), "cos_scores": sc.cpu().numpy()}, want)
    assert not bad, bad
    assert (want["cos_ids"][:, 0] == (np.arange(nq) * syn.QUERY_STRIDE) % n).all()   # the planted neighbour wins
    del index, corpus


@pytest.fixture(scope="module")
def big():
    """The bench workload: 10M x 1536 fp32 rows (+ fp16 shadow) and the 10M-doc Zipf token corpus on one GPU."""
    from optimized_rag_b200 import engine
    from optimized_rag_b200.bm25_index import Bm25Index
    torch.cuda.empty_cache()
    thr = syn.zipf_thresholds(VOCAB)
    corpus = engine.gen_embeddings(N_BIG, DIM, 0, syn.SEED_CORPUS, 0, device=DEV)
    cos = engine.CosineIndex(corpus, mode="f16")
    doc_off, tokens = engine.gen_token_corpus(N_BIG, 0, syn.SEED_TOKENS, thr, VOCAB, 100, 300, device=DEV)
    bm25 = Bm25Index(doc_off, tokens, VOCAB, tile_docs=2048)
    del tokens, doc_off
    torch.cuda.empty_cache()
    ref = oracle.StreamedBM25(syn.SEED_TOKENS, N_BIG, VOCAB, 100, 300, thr)
    # the global statistics the index was built from are the oracle's (dict order, epsilon floor included)
    assert bm25.avgdl == ref.avgdl and bm25.eps == ref.eps and bm25.average_idf == ref.average_idf
    assert np.array_equal(bm25.idf.cpu().numpy().view(np.uint64), ref.idf.view(np.uint64))
    yield {"shard": engine.HybridShard(cos, bm25), "ref": ref, "thr": thr}
    del cos, bm25, corpus
    torch.cuda.empty_cache()


def test_config3_hybrid_10m_32_queries(big):
    nq = 256
    q = syn.query_embeddings(nq, N_BIG, DIM)
    qt, ql = syn.keyword_queries(nq, VOCAB, thresholds=big["thr"])
    res = big["shard"].search(torch.from_numpy(q).to(DEV), torch.from_numpy(qt).to(DEV), torch.from_numpy(ql).to(DEV),
                              K, check_overflow=False)
    assert int(res["status"].max().item()) == 0      # fast path, no repair through the exhaustive kernels
    rows = _subset(qt, ql, nq)
    assert len(rows) >= 32
    want, _ = scale_check.reference_lists(N_BIG, DIM, q[rows], qt[rows], ql[rows], K, seed_corpus=syn.SEED_CORPUS,
                                          bm25=big["ref"])
    got = _np(res, ["cos_ids", "cos_scores", "bm25_ids", "bm25_scores", "bm25_max", "ids", "rrf_scores"])
    bad = scale_check.compare(got, want, rows)
    assert not bad, bad


def test_config4_bm25_10m_batch_1024(big):
    nq = 1024
    qt, ql = syn.keyword_queries(nq, VOCAB, thresholds=big["thr"])
    status: list = []
    ids, sc, mx = big["shard"].bm25.topk(torch.from_numpy(qt).to(DEV), torch.from_numpy(ql).to(DEV), K,
                                         check_overflow=False, status_out=status)
    assert int(status[0].max().item()) == 0
    rows = _subset(qt, ql, nq)
    want, _ = scale_check.reference_lists(N_BIG, DIM, None, qt[rows], ql[rows], K, seed_corpus=syn.SEED_CORPUS,
                                          bm25=big["ref"], want_cosine=False)
    got = {"bm25_ids": ids.cpu().numpy(), "bm25_scores": sc.cpu().numpy(), "bm25_max": mx.cpu().numpy()}
    bad = scale_check.compare(got, want, rows)
    assert not bad, bad
