// pairwise.cu -- all-pairs cosine candidates for the consistency checker
// (rag/consistency_checker.py:169-189): every i < j with doc_idx[i] != doc_idx[j] and
// float64 cosine >= threshold.  Exact CUDA-core version: lane <-> row j, a block of 4 "query"
// rows i per pass, same float64 arithmetic as the retrieval cosine (exact.cuh).
#include "common.cuh"
#include "exact.cuh"

namespace orag {

constexpr int kPairQB = 4;

// sq[r] = Neumaier sum of emb[r]^2
__global__ void __launch_bounds__(256) pair_sq_kernel(const float *__restrict__ emb, int64_t m, int dim,
                                                     double *__restrict__ sq)
{
    __shared__ float stage_all[8][32 * 33];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    int64_t warp = blockIdx.x * (int64_t)(blockDim.x >> 5) + wib;
    int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t rb = warp; rb < (m + 31) / 32; rb += n_warps) {
        int64_t r = rb * 32 + lane;
        const float *rowptr = r < m ? emb + r * (int64_t)dim : nullptr;
        const float *qp[1] = {nullptr};
        NeuSum dot[1], s;
        warp_score_rows<1>(rowptr, qp, dim, stage_all[wib], dot, s, true);
        if (r < m) sq[r] = s.result();
    }
}

__global__ void __launch_bounds__(256) pairwise_kernel(const float *__restrict__ emb, int64_t m, int dim,
                                                      const int32_t *__restrict__ doc_idx, const double *__restrict__ sq,
                                                      double thr, int64_t cap, int32_t *__restrict__ out_i,
                                                      int32_t *__restrict__ out_j, double *__restrict__ out_sim,
                                                      unsigned long long *__restrict__ out_count)
{
    __shared__ float stage_all[8][32 * 33];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t n_jb = (m + 31) / 32;
    const int64_t n_ib = (m + kPairQB - 1) / kPairQB;
    int64_t warp = blockIdx.x * (int64_t)(blockDim.x >> 5) + wib;
    int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t w = warp; w < n_jb * n_ib; w += n_warps) {
        const int64_t jb = w % n_jb, ib = w / n_jb;
        const int64_t i0 = ib * kPairQB;
        if (jb * 32 + 31 <= i0) continue;  // whole j block is <= every i of the group: no i < j pair
        const int64_t j = jb * 32 + lane;
        const float *rowptr = j < m ? emb + j * (int64_t)dim : nullptr;
        const float *qp[kPairQB];
#pragma unroll
        for (int b = 0; b < kPairQB; ++b) qp[b] = (i0 + b < m) ? emb + (i0 + b) * (int64_t)dim : nullptr;
        NeuSum dot[kPairQB], s;
        warp_score_rows<kPairQB>(rowptr, qp, dim, stage_all[wib], dot, s, false);
        if (j >= m) continue;
#pragma unroll
        for (int b = 0; b < kPairQB; ++b) {
            const int64_t i = i0 + b;
            if (i >= m || i >= j || doc_idx[i] == doc_idx[j]) continue;
            // argument order of the reference: cosine(emb[i], emb[j]) -> products emb[i][c] * emb[j][c]
            const double c = cosine_from_sums(dot[b].result(), sq[i], sq[j]);
            if (c >= thr) {
                unsigned long long slot = atomicAdd(out_count, 1ull);
                if ((int64_t)slot < cap) {
                    out_i[slot] = (int32_t)i;
                    out_j[slot] = (int32_t)j;
                    out_sim[slot] = c;
                }
            }
        }
    }
}

}  // namespace orag

extern "C" int orag_pairwise_cosine_threshold(const float *d_emb, int64_t m, int dim, const int32_t *d_doc_idx,
                                              double threshold, int64_t cap, int32_t *d_out_i, int32_t *d_out_j,
                                              double *d_out_sim, unsigned long long *d_out_count, void *d_workspace,
                                              size_t workspace_bytes, void *stream)
{
    ORAG_REQUIRE(d_emb && d_doc_idx && d_out_i && d_out_j && d_out_sim && d_out_count && m >= 0 && dim > 0 && cap >= 0,
                 "pairwise");
    ORAG_REQUIRE(m < ((int64_t)1 << 31), "m must fit int32");
    if (workspace_bytes < orag_pairwise_workspace_bytes(m, dim) || !d_workspace) {
        orag::set_error("pairwise: workspace too small");
        return ORAG_EWORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    ORAG_CUDA_CHECK(cudaMemsetAsync(d_out_count, 0, sizeof(unsigned long long), st));
    if (m < 2) return ORAG_OK;
    double *sq = (double *)d_workspace;
    int64_t blocks = ((m + 31) / 32 + 7) / 8;
    orag::pair_sq_kernel<<<(unsigned)(blocks < 1 ? 1 : blocks), 256, 0, st>>>(d_emb, m, dim, sq);
    ORAG_LAUNCH_CHECK();
    int64_t warps = ((m + 31) / 32) * ((m + orag::kPairQB - 1) / orag::kPairQB);
    int64_t grid = (warps + 7) / 8;
    int64_t lim = (int64_t)orag::sm_count() * 8;
    if (grid > lim) grid = lim;
    orag::pairwise_kernel<<<(unsigned)grid, 256, 0, st>>>(d_emb, m, dim, d_doc_idx, sq, threshold, cap, d_out_i, d_out_j,
                                                          d_out_sim, d_out_count);
    ORAG_LAUNCH_CHECK();
    return ORAG_OK;
}
