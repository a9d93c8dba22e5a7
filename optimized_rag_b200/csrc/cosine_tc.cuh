// cosine_tc.cuh -- shared declarations of the tcgen05 similarity scan (cosine_tc.cu) for api.cu / pairwise.cu
#pragma once
#include "common.cuh"

namespace orag {
namespace tc {

constexpr int kTileM = 128;
constexpr int kMaxN = 256;
constexpr int kStages = 4;
constexpr int kABytes = kTileM * 128;  // 16 KiB
constexpr int kBBytes = kMaxN * 128;   // 32 KiB
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kThreads = 384;
constexpr int kEpiWarp0 = 4;
constexpr int kEpiThreads = 256;
constexpr int kTmemCols = 512;
constexpr int kHistBins = 512;  // bins of width 2/512 over cos in [-1, 1]
constexpr int kStages2 = 6;     // ring of the 2-SM flavour (32 KiB stages: each CTA stages half of the query slab)
// dynamic shared memory of one scan CTA: alignment slack + ring + barriers + two threshold vectors
__host__ __device__ constexpr size_t scan_smem_bytes(bool mma2, int stages)
{
    return 1024 + (size_t)stages * (kABytes + (mma2 ? kBBytes / 2 : kBBytes)) + 256 + 2 * kMaxN * sizeof(float);
}

struct ScanParams {
    int64_t row_begin;      // first corpus row of this launch (multiple of 128 not required)
    int64_t row_end;        // one past the last valid row
    int num_tiles;
    int k_chunks;           // dim * elem_bytes / 128
    int chunk_elems;        // 32 (tf32) or 64 (bf16 / fp16)
    int f16;                // 16-bit operands are IEEE fp16 (1) or bf16 (0)
    int n_queries;
    int umma_n;             // round_up(n_queries, 16)
    uint32_t idesc;
    const float *inv_norm;  // [rows]
    int dense;              // 1: store every v to dense_out[(row - row_begin) * 256 + q]; 2: transposed,
                            //    dense_out[q * dense_ld + (row - row_begin)] (seed pass)
    float *dense_out;
    int dense_ld;
    // scan-mode state (per query)
    uint32_t *thr_key;      // ordered-uint of the threshold cosine
    uint32_t *cnt;
    uint32_t *hist;         // [n_queries, kHistBins]
    int32_t *cand;          // [n_queries, cap] local row ids
    float *cand_v;          // [n_queries, cap] first-pass value of each candidate (may be null): lets the fp32 re-score
                            //    skip the candidates that the FINAL threshold would no longer admit
    int cap;
    const float *qnorm;     // |q| as fp32
    const float *inv_qnorm;
    float margin;           // 2 * eps (cosine units)
    int k;
    int fixed_thr;          // 1: thresholds are given (pairwise >= t search): no histogram, no tightening
    int pair_mode;          // 1: all-pairs search, queries = corpus rows, triangular (query block, row tile) tiles
    int mma2;               // 1: the CTA pair issues ONE tcgen05.mma.cta_group::2 (M = 256 over both SMs): each CTA holds its
                            //    128 rows and HALF of the query slab (six 32 KiB stages instead of four 48 KiB ones);
                            //    default for main scans (ORAG_SCAN_2SM=0 falls back to the multicast pairs)
    int cluster2;           // 1: launched as clusters of two CTAs that share every query slab: each CTA fetches half of
                            //    it from L2 and multicasts it into both CTAs' shared memory (halves the L2 -> SM traffic
                            //    of the B operand); set by launch_scan
};

int launch_scan(bool bf16, const void *a_base, int64_t a_rows, const void *q_base, int dim, ScanParams p,
                cudaStream_t st);
// |first-pass cosine - cosine| bound of a tensor-core mode (ORAG_COS_TF32 / BF16 / F16) at vector length `dim` (api.cu)
float first_pass_eps(int mode, int dim);
int launch_seed_finalize(const float *seed, int seed_ld, int n_seed, int n_queries, int k, float margin, const float *qnorm,
                         const float *inv_qnorm, uint32_t *thr_key, uint32_t *cnt, uint32_t *hist, int32_t *cand,
                         float *cand_v,
                         int cap, cudaStream_t st);
int launch_query_norms(const double *sq, int n, float *qnorm, float *inv_qnorm, cudaStream_t st);

}  // namespace tc
}  // namespace orag
