"""Restatement of the third-party package ``rank-bm25`` 0.2.2 (``BM25Okapi`` only).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference depends on ``rank-bm25>=0.2.2`` (requirements.txt:22, no lock file)
and calls it at rag/retrieval.py:116, 326, 338, 341.  The package is not vendored
under /root/reference and not installed in this image (no network), so its
published algorithm is restated here from the description in SURVEY.md §8(a3) /
Appendix A.2.  Putting this directory on ``sys.path`` under the module name
``rank_bm25`` makes the reference's ``HybridRetriever.bm25_available`` True and
lets its own ``_bm25_scores`` glue run unmodified -- that is how the golden
vectors for BM25 are produced.  BM25 parity is therefore pinned to THIS
restatement, not to the upstream wheel ("parity unpinned" w.r.t. upstream).

Arithmetic that matters for bit-exactness:
  * nd (word -> document frequency) is filled in first-seen order; idf is
    computed walking that dict; ``idf_sum`` is a plain running ``+=``.
  * idf = log(N - df + 0.5) - log(df + 0.5) with math.log; negatives are replaced
    by epsilon * average_idf where average_idf = idf_sum / len(idf) is formed
    BEFORE the replacement.
  * get_scores: numpy float64, one query token at a time in query order:
    idf * (f * (k1 + 1) / (f + k1 * (1 - b + b * doc_len / avgdl))).
"""
import math

import numpy as np


class BM25Okapi:
    def __init__(self, corpus, tokenizer=None, k1=1.5, b=0.75, epsilon=0.25):
        self.k1 = k1
        self.b = b
        self.epsilon = epsilon
        self.tokenizer = tokenizer
        self.corpus_size = 0
        self.avgdl = 0
        self.doc_freqs = []
        self.idf = {}
        self.doc_len = []
        self.average_idf = 0.0

        if tokenizer:
            corpus = [tokenizer(doc) for doc in corpus]
        containing = self._scan(corpus)
        self._fill_idf(containing)

    def _scan(self, corpus):
        containing = {}  # word -> number of documents holding it (first-seen order)
        total_len = 0
        for tokens in corpus:
            self.doc_len.append(len(tokens))
            total_len += len(tokens)
            counts = {}
            for w in tokens:
                counts[w] = counts.get(w, 0) + 1
            self.doc_freqs.append(counts)
            for w in counts:
                if w in containing:
                    containing[w] += 1
                else:
                    containing[w] = 1
            self.corpus_size += 1
        self.avgdl = total_len / self.corpus_size
        return containing

    def _fill_idf(self, containing):
        running = 0
        below_zero = []
        for w, df in containing.items():
            v = math.log(self.corpus_size - df + 0.5) - math.log(df + 0.5)
            self.idf[w] = v
            running += v
            if v < 0:
                below_zero.append(w)
        self.average_idf = running / len(self.idf)
        floor = self.epsilon * self.average_idf
        for w in below_zero:
            self.idf[w] = floor

    def get_scores(self, query):
        score = np.zeros(self.corpus_size)
        doc_len = np.array(self.doc_len)
        for q in query:
            f = np.array([(d.get(q) or 0) for d in self.doc_freqs])
            score += (self.idf.get(q) or 0) * (f * (self.k1 + 1) /
                                               (f + self.k1 * (1 - self.b + self.b * doc_len / self.avgdl)))
        return score

    def get_top_n(self, query, documents, n=5):
        scores = self.get_scores(query)
        order = np.argsort(scores)[::-1][:n]
        return [documents[i] for i in order]
