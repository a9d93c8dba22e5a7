// gen.cu -- synthetic input fill kernels, bit-identical to optimized_rag_b200/synthetic.py
// (counter-based integer hashing so host and device agree for any shard layout).
#include "common.cuh"

namespace orag {

__device__ __forceinline__ uint64_t source_row(uint64_t seed, uint64_t row, int dup_per_mille)
{
    if (dup_per_mille <= 0 || row == 0) return row;
    uint64_t h = mix64(seed ^ ORAG_K_DUP ^ (row * ORAG_K_DOC));
    if ((h % 1000ull) < (uint64_t)dup_per_mille) return mix64(h) % row;
    return row;
}

// one CTA per row, threads stride over 4-column groups (16-byte stores, coalesced)
__global__ void __launch_bounds__(256) gen_embeddings_kernel(float *__restrict__ out, int64_t n_rows, int dim,
                                                            int64_t row_start, uint64_t seed, int dup_per_mille)
{
    for (int64_t r = blockIdx.x; r < n_rows; r += gridDim.x) {
        uint64_t key = row_key(seed, source_row(seed, (uint64_t)(row_start + r), dup_per_mille));
        float *dst = out + r * (int64_t)dim;
        int dim4 = dim & ~3;
        for (int c = threadIdx.x * 4; c < dim4; c += blockDim.x * 4) {
            float4 v;
            v.x = (float)((int)(mix64(key + (uint64_t)c) >> 40) - (1 << 23)) * 3.7252902984619140625e-09f;
            v.y = (float)((int)(mix64(key + (uint64_t)(c + 1)) >> 40) - (1 << 23)) * 3.7252902984619140625e-09f;
            v.z = (float)((int)(mix64(key + (uint64_t)(c + 2)) >> 40) - (1 << 23)) * 3.7252902984619140625e-09f;
            v.w = (float)((int)(mix64(key + (uint64_t)(c + 3)) >> 40) - (1 << 23)) * 3.7252902984619140625e-09f;
            *reinterpret_cast<float4 *>(dst + c) = v;
        }
        for (int c = dim4 + threadIdx.x; c < dim; c += blockDim.x)
            dst[c] = (float)((int)(mix64(key + (uint64_t)c) >> 40) - (1 << 23)) * 3.7252902984619140625e-09f;
    }
}

__global__ void gen_doc_lengths_kernel(int32_t *__restrict__ out, int64_t n_docs, int64_t doc_start, uint64_t seed,
                                       int lmin, int lmax)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_docs) return;
    uint64_t h = mix64(seed ^ ((uint64_t)(doc_start + i) * ORAG_K_DOC));
    out[i] = lmin + (int32_t)(h % (uint64_t)(lmax - lmin + 1));
}

// one warp per document; lanes stride over token positions; inverse-CDF lookup by binary search
__global__ void __launch_bounds__(256) gen_tokens_kernel(int32_t *__restrict__ out, const int64_t *__restrict__ doc_off,
                                                        int64_t n_docs, int64_t doc_start, uint64_t seed,
                                                        const uint64_t *__restrict__ thr, int vocab)
{
    int lane = threadIdx.x & 31;
    int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t d = warp; d < n_docs; d += n_warps) {
        int64_t lo = doc_off[d], hi = doc_off[d + 1];
        uint64_t key = row_key(seed + 1, (uint64_t)(doc_start + d));
        for (int64_t p = lo + lane; p < hi; p += 32) {
            uint64_t u = mix64(key + (uint64_t)(p - lo)) >> 1;
            // first index with thr[idx] > u  (numpy searchsorted side='right')
            int a = 0, b = vocab;
            while (a < b) {
                int m = (a + b) >> 1;
                if (thr[m] > u) b = m; else a = m + 1;
            }
            out[p] = a < vocab ? a : vocab - 1;
        }
    }
}

}  // namespace orag

extern "C" int orag_gen_embeddings(float *d_out, int64_t n_rows, int dim, int64_t row_start, uint64_t seed,
                                   int dup_per_mille, void *stream)
{
    ORAG_REQUIRE(d_out && n_rows >= 0 && dim > 0, "gen_embeddings");
    ORAG_REQUIRE((reinterpret_cast<uintptr_t>(d_out) & 15) == 0 && (dim % 4 == 0), "16-byte aligned rows");
    if (n_rows == 0) return ORAG_OK;
    int64_t grid = n_rows < (int64_t)orag::sm_count() * 16 ? n_rows : (int64_t)orag::sm_count() * 16;
    orag::gen_embeddings_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(d_out, n_rows, dim, row_start, seed,
                                                                                  dup_per_mille);
    ORAG_LAUNCH_CHECK();
    return ORAG_OK;
}

extern "C" int orag_gen_doc_lengths(int32_t *d_out, int64_t n_docs, int64_t doc_start, uint64_t seed, int lmin,
                                    int lmax, void *stream)
{
    ORAG_REQUIRE(d_out && n_docs >= 0 && lmin >= 0 && lmax >= lmin, "gen_doc_lengths");
    if (n_docs == 0) return ORAG_OK;
    orag::gen_doc_lengths_kernel<<<(unsigned)((n_docs + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        d_out, n_docs, doc_start, seed, lmin, lmax);
    ORAG_LAUNCH_CHECK();
    return ORAG_OK;
}

extern "C" int orag_gen_tokens(int32_t *d_out, const int64_t *d_doc_off, int64_t n_docs, int64_t doc_start,
                               uint64_t seed, const uint64_t *d_thresholds, int vocab, void *stream)
{
    ORAG_REQUIRE(d_out && d_doc_off && d_thresholds && n_docs >= 0 && vocab > 0, "gen_tokens");
    if (n_docs == 0) return ORAG_OK;
    int64_t warps = n_docs;
    int64_t blocks = (warps + 7) / 8;
    int64_t cap = (int64_t)orag::sm_count() * 32;
    if (blocks > cap) blocks = cap;
    orag::gen_tokens_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_out, d_doc_off, n_docs, doc_start,
                                                                               seed, d_thresholds, vocab);
    ORAG_LAUNCH_CHECK();
    return ORAG_OK;
}
