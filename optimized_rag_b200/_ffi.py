"""ctypes binding of include/orag.h -- the only door from Python into the CUDA kernels.

There is NO fallback: if csrc/liborag.so is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes
from ctypes import POINTER, Structure, c_char_p, c_double, c_int, c_int32, c_int64, c_size_t, c_uint64, c_void_p
from pathlib import Path

LIB_PATH = Path(__file__).resolve().parent / "csrc" / "liborag.so"

ORAG_COS_EXACT, ORAG_COS_TF32, ORAG_COS_BF16, ORAG_COS_F16 = 0, 1, 2, 3
ORAG_STATUS_OVERFLOW = 1
ORAG_STATUS_EXCHANGE_TIMEOUT = 2
ORAG_BM25_NORMALIZE, ORAG_BM25_FORCE_SPARSE, ORAG_BM25_FORCE_DENSE, ORAG_BM25_EXACT_TILES = 1, 2, 4, 8
ORAG_BM25_BACKGROUND = 16
ORAG_BM25_KTH_LEVELS = 5
ORAG_PHASE_SCAN, ORAG_PHASE_FINISH, ORAG_PHASE_PREP, ORAG_PHASE_ALL = 1, 2, 4, 7

# every symbol include/orag.h declares (tests check the .so exports each one)
SYMBOLS = [
    "orag_version", "orag_last_error", "orag_device_info",
    "orag_launch_count", "orag_profile_enable", "orag_profile_read", "orag_profile_read_all", "orag_timeline_enable", "orag_timeline_read",
    "orag_gen_embeddings", "orag_gen_doc_lengths", "orag_gen_tokens",
    "orag_row_inv_norms", "orag_row_sq", "orag_f32_to_bf16", "orag_f32_to_f16_rows",
    "orag_cosine_mark_prescan", "orag_stream_wait_prescan",
    "orag_cosine_workspace_bytes", "orag_cosine_topk", "orag_cosine_topk_phase", "orag_cosine_last_counts", "orag_cosine_dense", "orag_dot_dense", "orag_cosine_firstpass_dense",
    "orag_bm25_build_workspace_bytes", "orag_bm25_index_plan", "orag_bm25_index_fill", "orag_bm25_term_kth",
    "orag_bm25_workspace_bytes", "orag_bm25_topk", "orag_bm25_dense", "orag_dense_topk",
    "orag_topk_merge", "orag_rrf_fuse", "orag_rrf_fuse_pair", "orag_hybrid_merge", "orag_weighted_sum3", "orag_div_scalar",
    "orag_pairwise_workspace_bytes", "orag_pairwise_cosine_threshold",
    "orag_pairwise_tc_workspace_bytes", "orag_pairwise_cosine_threshold_tc",
    "orag_pairwise_prepared_bytes", "orag_pairwise_pairs_workspace_bytes", "orag_pairwise_prepare", "orag_pairwise_pairs",
    "orag_exchange_bytes", "orag_exchange_alloc", "orag_exchange_free", "orag_exchange_export", "orag_exchange_open",
    "orag_exchange_close", "orag_hybrid_push", "orag_hybrid_wait",
]


class OragError(RuntimeError):
    pass


class Bm25IndexStruct(Structure):
    _fields_ = [
        ("n_docs", c_int64),
        ("vocab", c_int32),
        ("tile_docs", c_int32),
        ("n_tiles", c_int32),
        ("has_negative_idf", c_int32),
        ("max_doc_len", c_int32),
        ("reserved", c_int32),
        ("d_tile_base", c_void_p),
        ("d_tile_term_off", c_void_p),
        ("d_postings", c_void_p),
        ("d_doc_len", c_void_p),
        ("d_t4_table", c_void_p),
        ("d_r_table", c_void_p),
        ("d_idf", c_void_p),
        ("d_postings_r16", c_void_p),
        ("d_term_max_r", c_void_p),
        ("fp_tile_docs", c_int32),
        ("fp_n_tiles", c_int32),
        ("d_fp_tile_base", c_void_p),
        ("d_fp_tile_term_off", c_void_p),
        ("d_term_kth_r", c_void_p),
    ]


_lib = None
# The library keeps per-process state (profiling brackets, auxiliary stream, co-scheduling event) and the Python
# wrappers cache workspaces per index: the drop-in classes serialise their GPU sections on this lock, which is what
# the reference's own sharing model needs (one agent used from several threads: ThreadedConnectionPool /
# EmbeddingService lock, SURVEY.md §8b "Threading").
import threading  # noqa: E402
GPU_LOCK = threading.RLock()


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise OragError(
            f"{LIB_PATH} is missing: build it with `python -m optimized_rag_b200.build` "
            "(there is no CPU or PyTorch fallback for the retrieval hot path)")
    L = ctypes.CDLL(str(LIB_PATH))
    vp = c_void_p
    L.orag_version.restype = c_int
    L.orag_last_error.restype = c_char_p
    L.orag_device_info.argtypes = [POINTER(c_int), POINTER(c_int), POINTER(c_int)]
    L.orag_launch_count.restype = ctypes.c_ulonglong
    L.orag_profile_enable.argtypes = [c_int]
    L.orag_profile_read.argtypes = [POINTER(ctypes.c_float), POINTER(ctypes.c_float)]
    L.orag_profile_read_all.argtypes = [c_int, POINTER(ctypes.c_float), c_int]
    L.orag_timeline_enable.argtypes = [c_int]
    L.orag_timeline_read.argtypes = [POINTER(c_int), POINTER(ctypes.c_float), POINTER(ctypes.c_float), c_int]
    L.orag_gen_embeddings.argtypes = [vp, c_int64, c_int, c_int64, c_uint64, c_int, vp]
    L.orag_gen_doc_lengths.argtypes = [vp, c_int64, c_int64, c_uint64, c_int, c_int, vp]
    L.orag_gen_tokens.argtypes = [vp, vp, c_int64, c_int64, c_uint64, vp, c_int, vp]
    L.orag_row_inv_norms.argtypes = [vp, c_int64, c_int, vp, vp]
    L.orag_row_sq.argtypes = [vp, c_int64, c_int, vp, vp]
    L.orag_f32_to_bf16.argtypes = [vp, vp, c_int64, vp]
    L.orag_f32_to_f16_rows.argtypes = [vp, c_int64, c_int, vp, vp, vp, vp]
    L.orag_cosine_mark_prescan.argtypes = [c_int]
    L.orag_stream_wait_prescan.argtypes = [vp]
    L.orag_cosine_workspace_bytes.restype = c_size_t
    L.orag_cosine_workspace_bytes.argtypes = [c_int64, c_int, c_int, c_int, c_int]
    L.orag_cosine_topk.argtypes = [vp, vp, vp, vp, c_int64, c_int, c_int64, vp, c_int, c_int, c_int, vp, vp, vp, vp,
                                   c_size_t, vp]
    L.orag_cosine_topk_phase.argtypes = [vp, vp, vp, vp, c_int64, c_int, c_int64, vp, c_int, c_int, c_int, vp, vp, vp, vp,
                                         c_size_t, c_int, vp]
    L.orag_cosine_last_counts.argtypes = [vp, c_int, c_int, vp, vp, vp]
    L.orag_cosine_dense.argtypes = [vp, c_int64, c_int, vp, c_int, vp, vp]
    L.orag_dot_dense.argtypes = [vp, c_int64, c_int, vp, c_int, vp, vp, vp, vp]
    L.orag_cosine_firstpass_dense.argtypes = [vp, vp, vp, c_int64, c_int, vp, c_int, c_int, vp, vp, c_size_t, vp]
    L.orag_bm25_build_workspace_bytes.restype = c_size_t
    L.orag_bm25_build_workspace_bytes.argtypes = [c_int64, c_int, c_int, c_int]
    L.orag_bm25_index_plan.argtypes = [vp, vp, c_int64, c_int, c_int, c_int, c_int64, vp, vp, vp, vp, vp, vp, vp, vp, vp,
                                       c_size_t, POINTER(c_int64), vp]
    L.orag_bm25_index_fill.argtypes = [vp, vp, c_int64, c_int, c_int, c_int, vp, c_int, vp, vp, vp, vp, vp, vp, vp, vp,
                                       vp, c_size_t, vp]
    L.orag_bm25_term_kth.argtypes = [POINTER(Bm25IndexStruct), vp, vp]
    L.orag_bm25_workspace_bytes.restype = c_size_t
    L.orag_bm25_workspace_bytes.argtypes = [POINTER(Bm25IndexStruct), c_int, c_int, c_int]
    L.orag_bm25_topk.argtypes = [POINTER(Bm25IndexStruct), c_int64, vp, vp, c_int, c_int, c_int, c_int, vp, vp, vp, vp,
                                 vp, c_size_t, vp]
    L.orag_bm25_dense.argtypes = [POINTER(Bm25IndexStruct), vp, vp, c_int, c_int, vp, vp]
    L.orag_dense_topk.argtypes = [vp, c_int64, c_int64, c_int, c_int, c_int64, c_int, vp, vp, vp, vp]
    L.orag_topk_merge.argtypes = [vp, vp, c_int, c_int, c_int, vp, c_int, vp, vp, vp, vp]
    L.orag_rrf_fuse.argtypes = [vp, c_int, c_int, c_int, c_int, c_int, c_int, vp, vp, vp, vp]
    L.orag_rrf_fuse_pair.argtypes = [vp, vp, c_int, c_int, c_int, c_int, c_int, vp, vp, vp, vp, vp, vp, vp]
    L.orag_hybrid_merge.argtypes = [vp, c_int, c_int, c_int, c_int, c_int, c_int, c_int, vp, vp, vp, vp, vp, vp, vp, vp,
                                    vp, vp]
    L.orag_weighted_sum3.argtypes = [vp, vp, vp, c_int64, c_double, c_double, c_double, vp, vp]
    L.orag_div_scalar.argtypes = [vp, c_int64, c_double, vp, vp]
    L.orag_pairwise_workspace_bytes.restype = c_size_t
    L.orag_pairwise_workspace_bytes.argtypes = [c_int64, c_int]
    L.orag_pairwise_cosine_threshold.argtypes = [vp, c_int64, c_int, vp, c_double, c_int64, vp, vp, vp, vp, vp,
                                                 c_size_t, vp]
    L.orag_pairwise_tc_workspace_bytes.restype = c_size_t
    L.orag_pairwise_tc_workspace_bytes.argtypes = [c_int64, c_int]
    L.orag_pairwise_cosine_threshold_tc.argtypes = L.orag_pairwise_cosine_threshold.argtypes
    L.orag_pairwise_prepared_bytes.restype = c_size_t
    L.orag_pairwise_prepared_bytes.argtypes = [c_int64, c_int]
    L.orag_pairwise_pairs_workspace_bytes.restype = c_size_t
    L.orag_pairwise_pairs_workspace_bytes.argtypes = [c_int64]
    L.orag_pairwise_prepare.argtypes = [vp, c_int64, c_int, vp, c_size_t, vp]
    L.orag_pairwise_pairs.argtypes = [vp, vp, c_int64, c_int, vp, c_double, c_int64, vp, vp, vp, vp, vp, c_size_t, vp]
    L.orag_exchange_bytes.restype = c_size_t
    L.orag_exchange_bytes.argtypes = [c_int, c_int, c_int, c_int]
    L.orag_exchange_alloc.argtypes = [c_size_t, POINTER(c_void_p)]
    L.orag_exchange_free.argtypes = [vp]
    L.orag_exchange_export.argtypes = [vp, ctypes.c_char_p]
    L.orag_exchange_open.argtypes = [ctypes.c_char_p, POINTER(c_void_p)]
    L.orag_exchange_close.argtypes = [vp]
    L.orag_hybrid_push.argtypes = [vp, vp, vp, vp, vp, vp, vp, c_int, c_int, c_int, c_int, c_int, c_int, vp, c_uint64, vp]
    L.orag_hybrid_wait.argtypes = [vp, c_int, c_int, c_int, c_int, c_int, c_uint64, c_int, POINTER(c_void_p), vp]
    for name in SYMBOLS:
        f = getattr(L, name)
        if name not in ("orag_last_error", "orag_launch_count", "orag_cosine_workspace_bytes", "orag_bm25_workspace_bytes",
                        "orag_pairwise_workspace_bytes", "orag_pairwise_tc_workspace_bytes", "orag_exchange_bytes",
                        "orag_pairwise_prepared_bytes", "orag_pairwise_pairs_workspace_bytes",
                        "orag_bm25_build_workspace_bytes"):
            f.restype = c_int
    _lib = L
    return L


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().orag_last_error()
        raise OragError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")
