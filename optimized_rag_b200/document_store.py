"""Drop-in for `DocumentStore` (rag/document_store.py:14-542) with the chunks resident in HBM.

What the reference keeps in Postgres tables `documents` / `document_chunks(vector(1536))`
(rag/document_store.py:190-221) lives here in a per-agent `ChunkTable`: fp32 embeddings in one device
buffer (row = chunk id), the chunk texts/metadata on the host, and a tiled BM25 index over the
lower-cased, whitespace-split chunk texts.  `search` keeps the reference's signature, result shape
(`content, filename, file_type, score, metadata`), ordering (descending cosine) and error convention
(never raises: log + `[]`, rag/document_store.py:483-485) but is EXACT instead of HNSW-approximate.

`hybrid_search` is the composition the reference's README describes (cosine list + BM25 list -> RRF):
`score` stays the chunk's cosine (downstream thresholds treat it as a cosine-like value in [0, 1],
SURVEY.md §8b "score-scale hazard"); fusion data goes to extra keys.
"""
from __future__ import annotations

import logging
import math
from datetime import datetime, timezone
from pathlib import Path
from typing import Any, Dict, List, Optional

import numpy as np
import torch

import functools

from . import _ffi, engine
from .bm25_index import Bm25Index

logger = logging.getLogger(__name__)


def _gpu_locked(fn):
    """Serialise the GPU section of a public method on the process-wide lock (see _ffi.GPU_LOCK)."""
    @functools.wraps(fn)
    def wrapper(*a, **kw):
        with _ffi.GPU_LOCK:
            return fn(*a, **kw)
    return wrapper


class TextVocab:
    """word -> term id in first-seen order over `doc.lower().split()` (rag/retrieval.py:334-335)."""

    def __init__(self):
        self.ids: Dict[str, int] = {}

    def encode_doc(self, text: str) -> np.ndarray:
        out = []
        ids = self.ids
        for w in text.lower().split():
            t = ids.get(w)
            if t is None:
                t = len(ids)
                ids[w] = t
            out.append(t)
        return np.asarray(out, dtype=np.int32)

    def encode_query(self, text: str) -> np.ndarray:
        return np.asarray([self.ids.get(w, -1) for w in text.lower().split()], dtype=np.int32)

    def __len__(self):
        return len(self.ids)


class ChunkTable:
    """All chunks of one agent: host records + device embeddings + GPU indices.

    The cosine side is maintained INCREMENTALLY: the fp32 rows, their fp16 shadow, inverse norms and float64 sum(a*a)
    live in capacity-doubling device buffers and only the rows that arrive are converted (`extend`); deleting a
    document compacts all four with one gather each.  The BM25 side cannot be appended to -- every impact depends on
    the corpus-wide avgdl and every idf on N and df -- so it is rebuilt by the library's builder (csrc/bm25_build.cu,
    ~0.05 s per million chunks) the first time a search needs it after a change."""

    def __init__(self, dim: int, device: torch.device, tile_docs: int = 1024):
        self.dim = dim
        self.device = device
        self.tile_docs = tile_docs
        self.records: List[Dict[str, Any]] = []     # content, metadata, filename, file_type, document_id, chunk_index
        self.tokens: List[np.ndarray] = []
        self.vocab = TextVocab()
        self._n = 0
        self._alloc(0)
        self._cosine: Optional[engine.CosineIndex] = None
        self._bm25: Optional[Bm25Index] = None

    def _alloc(self, cap: int):
        dev = self.device
        self._emb = torch.empty((cap, self.dim), dtype=torch.float32, device=dev)
        self._shadow = torch.empty((cap, self.dim), dtype=torch.float16, device=dev)
        self._inv_norm = torch.empty(cap, dtype=torch.float32, device=dev)
        self._row_sq = torch.empty(cap, dtype=torch.float64, device=dev)

    def _arrays(self):
        return self._emb, self._shadow, self._inv_norm, self._row_sq

    def __len__(self):
        return self._n

    @property
    def shadow_usable(self) -> bool:
        return self.dim % 64 == 0

    def extend(self, records: List[Dict[str, Any]], embeddings: np.ndarray):
        """Append chunks (records[i] <-> embeddings[i], fp32 [m, dim]); converts only the new rows."""
        m = len(records)
        if m == 0:
            return
        embeddings = np.ascontiguousarray(embeddings, dtype=np.float32).reshape(m, self.dim)
        n0, n1 = self._n, self._n + m
        if n1 > self._emb.shape[0]:
            old = self._arrays()
            self._alloc(max(256, 2 * self._emb.shape[0], n1))
            for new, prev in zip(self._arrays(), old):
                new[:n0] = prev[:n0]
        self._emb[n0:n1] = torch.from_numpy(embeddings).to(self.device)
        if self.shadow_usable:
            engine.CosineIndex.derive_rows(self._emb[n0:n1], self._shadow[n0:n1], self._inv_norm[n0:n1],
                                           self._row_sq[n0:n1])
        self.records.extend(records)
        self.tokens.extend(self.vocab.encode_doc(r["content"]) for r in records)
        self._n = n1
        self._cosine = self._bm25 = None

    def append(self, record: Dict[str, Any], embedding: np.ndarray):
        self.extend([record], np.asarray(embedding, dtype=np.float32)[None, :])

    def remove_document(self, document_id: int) -> int:
        keep = [i for i, r in enumerate(self.records) if r["document_id"] != document_id]
        removed = self._n - len(keep)
        if removed:
            idx = torch.tensor(keep, dtype=torch.int64, device=self.device)
            old = self._arrays()
            self._alloc(max(256, len(keep)))
            for new, prev in zip(self._arrays(), old):
                new[:len(keep)] = prev[:self._n][idx]
            self.records = [self.records[i] for i in keep]
            self.tokens = [self.tokens[i] for i in keep]
            self._n = len(keep)
            self._cosine = self._bm25 = None
        return removed

    def cosine(self) -> engine.CosineIndex:
        """Exact scan for small tables, fp16 tensor-core first pass over the incrementally kept shadow otherwise (a
        view: nothing is recomputed)."""
        if self._cosine is None:
            n = self._n
            if n < engine.SMALL_N or not self.shadow_usable:
                self._cosine = engine.CosineIndex.from_arrays(self._emb[:n], None, None, None, "exact")
            else:
                self._cosine = engine.CosineIndex.from_arrays(self._emb[:n], self._inv_norm[:n], self._shadow[:n],
                                                              self._row_sq[:n], "f16")
        return self._cosine

    def token_arrays(self):
        lens = np.asarray([len(t) for t in self.tokens], dtype=np.int64)
        off = np.zeros(self._n + 1, dtype=np.int64)
        np.cumsum(lens, out=off[1:])
        toks = np.concatenate(self.tokens) if self._n and off[-1] > 0 else np.zeros(0, dtype=np.int32)
        return off, toks.astype(np.int32)

    def bm25(self) -> Bm25Index:
        if self._bm25 is None:
            off, toks = self.token_arrays()
            self._bm25 = Bm25Index(torch.from_numpy(off).to(self.device), torch.from_numpy(toks).to(self.device),
                                   max(len(self.vocab), 1), tile_docs=self.tile_docs)
        return self._bm25

    def adopt_bm25(self, index: Bm25Index) -> bool:
        """Use a saved index instead of rebuilding, if it describes exactly this table's chunks."""
        off, _ = self.token_arrays()
        if index.n_docs != self._n or index.vocab != max(len(self.vocab), 1) or index.tile_docs != self.tile_docs or \
                index.stats.total_len != int(off[-1]):
            return False
        self._bm25 = index
        return True


class DocumentStore:
    """GPU-resident document storage with exact cosine / hybrid retrieval."""

    def __init__(
        self,
        database_ops,
        embedding_service,
        chunking_strategy,
        data_wrangler=None,
        kg_extractor=None,
        index_type: str = "hnsw",   # accepted for signature compatibility; retrieval here is exact
        ivfflat_lists: int = 100,
        device: str | torch.device = "cuda",
        retrieval_mode: str = "semantic",  # "semantic" = the reference's live path, "hybrid" = cosine+BM25->RRF
        rrf_k: int = 60,
    ):
        self.db = database_ops
        self.embeddings = embedding_service
        self.chunker = chunking_strategy
        self.wrangler = data_wrangler
        self.kg_extractor = kg_extractor
        self.index_type = index_type.lower()
        self.ivfflat_lists = ivfflat_lists
        self.embedding_dim = self.embeddings.get_embedding_dimension()
        self.device = torch.device(device)
        self.retrieval_mode = retrieval_mode
        self.rrf_k = rrf_k
        self._tables: Dict[str, ChunkTable] = {}
        self._documents: Dict[int, Dict[str, Any]] = {}
        self._next_doc_id = 1
        logger.info(f"GPU DocumentStore initialized: device={self.device}, embedding_dim={self.embedding_dim}, "
                    f"retrieval_mode={retrieval_mode}")

    # ------------------------------------------------------------------ ingest (rag/document_store.py:238-422)
    def _table(self, agent_id: str) -> ChunkTable:
        t = self._tables.get(agent_id)
        if t is None:
            t = self._tables[agent_id] = ChunkTable(self.embedding_dim, self.device)
        return t

    def upload_and_index(self, agent_id: str, file_path: str, file_content: Optional[str] = None,
                         metadata: Optional[Dict[str, Any]] = None) -> Dict[str, Any]:
        """Wrangle, chunk, embed and validate FIRST (no lock, nothing changed yet -- these are the slow, fallible steps:
        the reference generates its embeddings before it opens the transaction that deletes and inserts,
        rag/document_store.py:317-389); only then, under the lock, replace the document's chunks in one step.  A failure
        anywhere before that leaves an earlier upload of the same file fully searchable."""
        try:
            file_path_obj = Path(file_path)
            filename = file_path_obj.name
            file_type = file_path_obj.suffix.lower()
            raw = file_content if file_content else file_path_obj.read_text(encoding="utf-8", errors="replace")
            if self.wrangler:
                wr = self.wrangler.process(raw)
                content, quality_score, extracted = wr['cleaned_text'], wr['quality_score'], wr.get('metadata', {})
            else:
                content, quality_score, extracted = raw, None, {}
            full_metadata = {**extracted, **(metadata or {})}
            content = content.replace('\x00', '')
            chunks = self.chunker.chunk(content)
            records, rows, skipped = [], [], 0
            if len(chunks) > 0:
                embeddings = self.embeddings.generate_embeddings_batch([c['content'] for c in chunks])
                for i, (chunk, emb) in enumerate(zip(chunks, embeddings)):
                    if emb is None or len(emb) == 0:
                        logger.warning(f"Skipping chunk {i}: empty embedding")
                        skipped += 1
                        continue
                    if any(math.isnan(v) or math.isinf(v) for v in emb):
                        logger.warning(f"Skipping chunk {i}: embedding contains NaN or Inf")
                        skipped += 1
                        continue
                    if len(emb) != self.embedding_dim:
                        raise ValueError(f"embedding dimension {len(emb)} != {self.embedding_dim}")
                    records.append({"content": chunk['content'].replace('\x00', ''),
                                    "metadata": {**full_metadata, **(chunk.get('metadata', {}))},
                                    "filename": filename, "file_type": file_type, "chunk_index": i})
                    rows.append(np.asarray(emb, dtype=np.float32))
            with _ffi.GPU_LOCK:
                # same (agent, filename) replaces the earlier upload (ON CONFLICT ... DO UPDATE in the reference)
                document_id = next((d for d, r in self._documents.items()
                                    if r["agent_id"] == agent_id and r["filename"] == filename), None)
                if document_id is None:
                    document_id = self._next_doc_id
                    self._next_doc_id += 1
                self._documents[document_id] = {"agent_id": agent_id, "filename": filename, "file_type": file_type,
                                                "quality_score": quality_score, "metadata": full_metadata,
                                                "uploaded_at": datetime.now(timezone.utc)}
                if len(chunks) == 0:   # the reference returns before it touches any chunk row
                    return {"document_id": document_id, "filename": filename, "chunk_count": 0, "chunks_created": 0,
                            "chunks_skipped": 0, "quality_score": quality_score, "success": True,
                            "error": "No chunks generated from document"}
                table = self._table(agent_id)
                table.remove_document(document_id)
                for r in records:
                    r["document_id"] = document_id
                table.extend(records, np.stack(rows) if rows else np.zeros((0, self.embedding_dim), np.float32))
            if self.kg_extractor:
                try:
                    triples = self.kg_extractor.extract_triples(text=content, source_doc_id=document_id, max_triples=20)
                    self.kg_extractor.store_triples(triples, agent_id)
                except Exception as e:  # noqa: BLE001 - mirrors the reference's blanket handler
                    logger.warning(f"KG extraction failed: {e}")
            return {"document_id": document_id, "filename": filename, "chunk_count": len(chunks),
                    "chunks_created": len(records), "chunks_skipped": skipped, "quality_score": quality_score,
                    "success": True}
        except Exception as e:  # noqa: BLE001
            logger.error(f"Upload and index failed: {e}")
            return {"success": False, "error": str(e)}

    # ------------------------------------------------------------------ retrieval
    def _result(self, rec: Dict[str, Any], score: float) -> Dict[str, Any]:
        # fresh dicts: callers mutate results in place (rag/retrieval.py:182-183)
        return {"content": rec["content"], "filename": rec["filename"], "file_type": rec["file_type"],
                "score": float(score), "metadata": dict(rec["metadata"])}

    def search(self, agent_id: str, query: str, top_k: int = 5) -> List[Dict[str, Any]]:
        """Search document chunks (rag/document_store.py:424-485)."""
        try:
            if self.retrieval_mode == "hybrid":
                return self.hybrid_search(agent_id, query, top_k)
            query_embedding = self.embeddings.generate_embedding(query)   # (a network call in the reference: no lock)
            with _ffi.GPU_LOCK:
                return self._semantic(agent_id, query_embedding, top_k)
        except Exception as e:  # noqa: BLE001
            logger.error(f"Search failed: {e}")
            return []

    def _semantic(self, agent_id: str, query_embedding, top_k: int) -> List[Dict[str, Any]]:
        table = self._tables.get(agent_id)
        if table is None or len(table) == 0 or top_k <= 0:
            return []
        q = torch.tensor([query_embedding], dtype=torch.float32, device=self.device)
        k = min(top_k, len(table))
        index = table.cosine()
        # the tensor-core first pass keeps at most 128 winners per query: larger requests take the exact scan
        ids, scores = index.topk(q, k, mode="exact" if k > 128 else None)
        return [self._result(table.records[i], s)
                for i, s in zip(ids[0].cpu().tolist(), scores[0].cpu().tolist()) if i >= 0]

    def hybrid_search(self, agent_id: str, query: str, top_k: int = 5, fetch_k: Optional[int] = None
                      ) -> List[Dict[str, Any]]:
        """Cosine top-fetch_k + BM25 top-fetch_k -> RRF top_k, all on the GPU."""
        try:
            query_embedding = self.embeddings.generate_embedding(query)
            with _ffi.GPU_LOCK:
                return self._hybrid(agent_id, query, query_embedding, top_k, fetch_k)
        except Exception as e:  # noqa: BLE001
            logger.error(f"Hybrid search failed: {e}")
            return []

    def _hybrid(self, agent_id, query, query_embedding, top_k, fetch_k):
        table = self._tables.get(agent_id)
        if table is None or len(table) == 0 or top_k <= 0:
            return []
        n = len(table)
        fetch_k = min(fetch_k or max(top_k, 10), n)
        if fetch_k > 64 or min(top_k, n) > 64:
            # the fusion kernel holds two lists of at most 64: beyond that the semantic ranking alone is returned
            logger.warning("hybrid search: top_k above 64 is served by the semantic ranking only")
            return self._semantic(agent_id, query_embedding, top_k)
        q = torch.tensor([query_embedding], dtype=torch.float32, device=self.device)
        terms = table.vocab.encode_query(query)
        terms = terms[terms >= 0]   # out-of-vocabulary tokens score an exact 0.0 (rank_bm25: `idf.get(q) or 0`)
        qt = torch.from_numpy(terms if len(terms) else np.full(1, -1, np.int32)).to(self.device)[None, :].contiguous()
        ql = torch.tensor([len(terms)], dtype=torch.int32, device=self.device)
        shard = engine.HybridShard(table.cosine(), table.bm25(), self.rrf_k)
        res = shard.search(q, qt, ql, k=min(top_k, n), fetch_k=fetch_k)
        ids = res["ids"][0].cpu().tolist()
        rrf = res["rrf_scores"][0].cpu().tolist()
        src = res["src_ranks"][0].cpu().tolist()
        cos_of = dict(zip(res["cos_ids"][0].cpu().tolist(), res["cos_scores"][0].cpu().tolist()))
        kw_of = dict(zip(res["bm25_ids"][0].cpu().tolist(), res["bm25_scores"][0].cpu().tolist()))
        missing = [i for i in ids if i >= 0 and i not in cos_of]
        if missing:  # fused items that came from the BM25 list only: their cosine, same float64 arithmetic
            sub = engine.CosineIndex(table._emb[torch.tensor(missing, device=self.device)].contiguous(), mode="exact")
            for i, s in zip(missing, sub.dense(q)[0].cpu().tolist()):
                cos_of[i] = s
        out = []
        for i, r, sr in zip(ids, rrf, src):
            if i < 0:
                continue
            d = self._result(table.records[i], cos_of[i])
            d["rrf_score"] = r
            d["keyword_score"] = kw_of.get(i)
            d["semantic_rank"] = sr[0] or None
            d["keyword_rank"] = sr[1] or None
            out.append(d)
        return out

    # ------------------------------------------------------------------ persistence
    # The reference's "format" is Postgres rows (DDL rag/document_store.py:190-221).  Here: one directory per store,
    #   store.json                 documents, per-agent chunk records (content, metadata, filename, document_id, ...)
    #   <agent>.emb.f32            raw little-endian fp32 [n_chunks, dim] row-major: the packed corpus, memory-mappable
    # Derived structures (fp16 shadow, norms, BM25 postings) are rebuilt from these on first use after `load`.
    @_gpu_locked
    def save(self, directory: str) -> None:
        import json
        d = Path(directory)
        d.mkdir(parents=True, exist_ok=True)
        agents = {}
        for n, (agent_id, t) in enumerate(sorted(self._tables.items())):
            fname = f"agent{n}.emb.f32"
            t._emb[:len(t)].cpu().numpy().astype("<f4").tofile(d / fname)
            # the vocabulary in id order: ids are assigned in first-seen order and survive deletions, so the saved
            # keyword index is only meaningful together with the very mapping it was built under
            agents[agent_id] = {"file": fname, "n": len(t), "records": t.records, "bm25": None,
                                "vocab": list(t.vocab.ids)}
            if len(t) > 0:   # the keyword index as it sits in HBM: adopted by `load` instead of a rebuild
                t.bm25().save(d / f"agent{n}.bm25")
                agents[agent_id]["bm25"] = f"agent{n}.bm25"
        docs = {str(k): {**v, "uploaded_at": v["uploaded_at"].isoformat()} for k, v in self._documents.items()}
        (d / "store.json").write_text(json.dumps({"version": 1, "dim": self.embedding_dim, "next_doc_id": self._next_doc_id,
                                                  "documents": docs, "agents": agents}))

    @_gpu_locked
    def load(self, directory: str) -> None:
        import json
        d = Path(directory)
        meta = json.loads((d / "store.json").read_text())
        if meta.get("version") != 1 or meta["dim"] != self.embedding_dim:
            raise ValueError("incompatible store directory")
        self._tables, self._documents = {}, {}
        self._next_doc_id = int(meta["next_doc_id"])
        for k, v in meta["documents"].items():
            self._documents[int(k)] = {**v, "uploaded_at": datetime.fromisoformat(v["uploaded_at"])}
        for agent_id, a in meta["agents"].items():
            t = self._table(agent_id)
            emb = np.fromfile(d / a["file"], dtype="<f4").reshape(a["n"], self.embedding_dim)
            t.vocab.ids = {w: i for i, w in enumerate(a.get("vocab", []))}
            t.extend(a["records"], emb)   # one copy + one conversion launch for the whole table
            if a.get("bm25") and "vocab" in a:   # (term ids are only stable under the saved mapping)
                try:
                    if not t.adopt_bm25(Bm25Index.load(d / a["bm25"], device=self.device)):
                        logger.warning(f"saved keyword index of {agent_id} does not match its chunks: it will be rebuilt")
                except (OSError, ValueError) as e:
                    logger.warning(f"saved keyword index of {agent_id} unusable ({e}): it will be rebuilt")

    # ------------------------------------------------------------------ bookkeeping (rag/document_store.py:487-542)
    @_gpu_locked
    def list_documents(self, agent_id: str) -> List[Dict[str, Any]]:
        try:
            table = self._tables.get(agent_id)
            docs = []
            for doc_id, r in self._documents.items():
                if r["agent_id"] != agent_id:
                    continue
                chunks = sum(1 for c in (table.records if table else []) if c["document_id"] == doc_id)
                docs.append({"id": doc_id, "filename": r["filename"], "file_type": r["file_type"],
                             "quality_score": float(r["quality_score"]) if r["quality_score"] else None,
                             "chunk_count": chunks, "uploaded_at": r["uploaded_at"].isoformat()})
            docs.sort(key=lambda d: d["uploaded_at"], reverse=True)
            return docs
        except Exception as e:  # noqa: BLE001
            logger.error(f"List documents failed: {e}")
            return []

    @_gpu_locked
    def delete_document(self, agent_id: str, document_id: int) -> bool:
        try:
            r = self._documents.get(document_id)
            if r is not None and r["agent_id"] == agent_id:
                del self._documents[document_id]
                table = self._tables.get(agent_id)
                if table:
                    table.remove_document(document_id)
            logger.info(f"Document deleted: {document_id}")
            return True
        except Exception as e:  # noqa: BLE001
            logger.error(f"Delete failed: {e}")
            return False
