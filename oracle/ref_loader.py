"""Load the reference's own hot-path modules by file path (build container only).

TEST INFRASTRUCTURE ONLY.  /root/reference does not exist on the GPU box, so
nothing that runs there (``-m gpu`` tests, smoke(), bench.py) may call this;
it is used by tests/golden/make_golden.py and by the CPU tests that re-validate
the restatement when the reference tree is present (skipped otherwise).

Recipe follows SURVEY.md §8(c): ``import rag`` fails (rag/__init__.py pulls
langdetect), so retrieval.py / reranker.py / chunking.py / consistency_checker.py
are loaded with importlib by path, with dummy OPENAI_API_KEY / POSTGRES_URI so
``import config`` inside hybrid_search (rag/retrieval.py:240) succeeds, and with
oracle/rank_bm25.py on sys.path so ``bm25_available`` is True.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types
from pathlib import Path

REF_ROOT = Path(os.environ.get("ORAG_REFERENCE_ROOT", "/root/reference"))
_ORACLE_DIR = Path(__file__).resolve().parent
_cache: dict = {}


def available() -> bool:
    return (REF_ROOT / "rag" / "retrieval.py").exists()


def _prepare():
    os.environ.setdefault("OPENAI_API_KEY", "dummy")
    os.environ.setdefault("POSTGRES_URI", "postgresql://x:y@localhost/z")
    if str(REF_ROOT) not in sys.path:
        sys.path.insert(0, str(REF_ROOT))
    if str(_ORACLE_DIR) not in sys.path:
        sys.path.append(str(_ORACLE_DIR))  # exposes rank_bm25 restatement


def load(name: str):
    """name in {'retrieval','reranker','chunking','consistency_checker','data_wrangler','context_compressor','helpers'}."""
    if name in _cache:
        return _cache[name]
    if not available():
        raise FileNotFoundError(f"reference tree not present at {REF_ROOT}")
    _prepare()
    if name == "consistency_checker":
        # its `from memory.embeddings import EmbeddingService` would drag in psycopg2
        if "memory" not in sys.modules:
            mem = types.ModuleType("memory")
            emb = types.ModuleType("memory.embeddings")
            emb.EmbeddingService = type("EmbeddingService", (), {})
            mem.embeddings = emb
            sys.modules["memory"] = mem
            sys.modules["memory.embeddings"] = emb
    if name in ("context_compressor", "helpers"):
        # `from rag.models.intent_analysis import QueryIntent` would import the rag package (langdetect, langgraph, ...):
        # register the package shells and load the one light module by path; langdetect itself is only used by
        # helpers.detect_language, never on the paths recorded here
        for pkg in ("rag", "rag.models"):
            if pkg not in sys.modules:
                sys.modules[pkg] = types.ModuleType(pkg)
                sys.modules[pkg].__path__ = []
        if "rag.models.intent_analysis" not in sys.modules:
            spec = importlib.util.spec_from_file_location("rag.models.intent_analysis",
                                                          REF_ROOT / "rag" / "models" / "intent_analysis.py")
            mod = importlib.util.module_from_spec(spec)
            sys.modules["rag.models.intent_analysis"] = mod
            spec.loader.exec_module(mod)
        if "langdetect" not in sys.modules:
            ld = types.ModuleType("langdetect")
            ld.detect = lambda text: "en"
            lde = types.ModuleType("langdetect.lang_detect_exception")
            lde.LangDetectException = type("LangDetectException", (Exception,), {})
            sys.modules["langdetect"], sys.modules["langdetect.lang_detect_exception"] = ld, lde
        if "memory" not in sys.modules:
            mem = types.ModuleType("memory")
            emb = types.ModuleType("memory.embeddings")
            emb.EmbeddingService = type("EmbeddingService", (), {})
            mem.embeddings = emb
            sys.modules["memory"] = mem
            sys.modules["memory.embeddings"] = emb
    path = REF_ROOT / "rag" / ("nodes/helpers.py" if name == "helpers" else f"{name}.py")
    spec = importlib.util.spec_from_file_location(f"_orag_ref_{name}", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _cache[name] = mod
    return mod


def hybrid_retriever(**kw):
    """A reference HybridRetriever with no stores attached (enough for hybrid_search,
    _bm25_scores, _cosine_similarity)."""
    return load("retrieval").HybridRetriever(None, None, "oracle", **kw)


def rrf(k: int = 60):
    return load("reranker").ReciprocalRankFusion(k=k)
