// exchange.cu -- the one data-path exchange of a row-sharded hybrid search, over NVLink peer memory.
//
// After its local kernels every rank holds, per query, its shard's cosine top-fetch_k, BM25 raw top-kk,
// the shard's max raw BM25 score and the overflow flags: W = 2*fetch_k + 2*kk + 2 int64 words per query,
// ~110 KB per rank at 256 queries.  Instead of packing them (torch.cat) and calling an NCCL all-gather,
// `exchange_push_kernel` packs AND stores the block straight into every peer's exchange buffer (peer
// pointers obtained once through CUDA IPC; NVSwitch gives every pair its own full-bandwidth path) and then
// publishes a sequence number with a system-scope release store; `exchange_wait_kernel` (one tiny CTA)
// acquires the G sequence numbers in the rank's OWN buffer, after which the slot is exactly the
// [G, B, W] array orag_hybrid_merge reads.  Two launches replace ~6 packing kernels + the collective.
//
// Buffer of one rank (orag_exchange_bytes):   uint64 seq[6][G] (padded to 256 B) | int64 slot[6][G * max_queries * W]
// A search with sequence number s uses slot s % 6, and callers may keep three searches in flight on three streams
// ("lanes": lane = s % 3, each lane in stream order).  Re-use is safe with 2 x lanes slots: rank r writes slot s % 6 of
// peer p for search s+6 after -- same lane, stream order -- its own wait for search s+3 returned, i.e. after p
// published s+3, which p does (its lane of s, stream order) after its merge of search s has finished reading that slot.
//
// A wait that does not see a peer's sequence number within the timeout does not hang the GPU: it sets
// ORAG_STATUS_EXCHANGE_TIMEOUT in the status word of every query of that shard's block, which the merge ORs
// into the per-query status the caller checks.
#include <string.h>

#include "common.cuh"

namespace orag {

constexpr int kXSlots = 6;  // 2 x lanes (see above)
constexpr int kXMaxShards = 64;

static inline size_t x_flag_bytes(int n_shards) { return align_up((size_t)kXSlots * n_shards * 8, 256); }
static inline size_t x_slot_words(int n_shards, int max_queries, int fetch_k, int kk)
{
    return (size_t)n_shards * max_queries * (2 * fetch_k + 2 * kk + 2);
}

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// CTA p packs this rank's block and stores it into peer p's slot (p == rank: its own buffer).
__global__ void __launch_bounds__(256) exchange_push_kernel(
    const int64_t *__restrict__ cos_ids, const double *__restrict__ cos_scores, const int64_t *__restrict__ bm_ids,
    const double *__restrict__ bm_scores, const double *__restrict__ bm_max, const int32_t *__restrict__ status,
    const int32_t *__restrict__ status2, int B,
    int fetch_k, int kk, int rank, int G, void *const *__restrict__ peers, size_t flag_bytes, size_t slot_words,
    unsigned long long seq)
{
    const int p = blockIdx.x;
    const int slot = (int)(seq % (unsigned long long)kXSlots);
    const int W = 2 * fetch_k + 2 * kk + 2;
    uint8_t *base = (uint8_t *)peers[p];
    int64_t *dst = (int64_t *)(base + flag_bytes) + (size_t)slot * slot_words + (size_t)rank * B * W;
    const int total = B * W;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int q = i / W, c = i - q * W;
        int64_t v;
        if (c < fetch_k) v = cos_ids[q * fetch_k + c];
        else if (c < 2 * fetch_k) v = __double_as_longlong(cos_scores[q * fetch_k + c - fetch_k]);
        else if (c < 2 * fetch_k + kk) v = bm_ids[q * kk + c - 2 * fetch_k];
        else if (c < 2 * fetch_k + 2 * kk) v = __double_as_longlong(bm_scores[q * kk + c - 2 * fetch_k - kk]);
        else if (c == W - 2) v = __double_as_longlong(bm_max[q]);
        else v = (int64_t)((status ? status[q] : 0) | (status2 ? status2[q] : 0));
        dst[i] = v;
    }
    // every thread's stores are ordered before the flag: fence (system scope), CTA barrier, release store
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0)
        st_release_sys((unsigned long long *)base + (size_t)slot * G + rank, seq);
}

// thread g waits until shard g's block of search `seq` has landed in this rank's buffer
__global__ void __launch_bounds__(kXMaxShards) exchange_wait_kernel(uint8_t *mine, int G, int B, int W, size_t flag_bytes,
                                                                   size_t slot_words, unsigned long long seq,
                                                                   unsigned long long timeout_ns)
{
    const int g = threadIdx.x;
    if (g >= G) return;
    const int slot = (int)(seq % (unsigned long long)kXSlots);
    const unsigned long long *flag = (const unsigned long long *)mine + (size_t)slot * G + g;
    const unsigned long long t0 = global_ns();
    while (ld_acquire_sys(flag) < seq) {
        if (global_ns() - t0 > timeout_ns) {
            int64_t *blk = (int64_t *)(mine + flag_bytes) + (size_t)slot * slot_words + (size_t)g * B * W;
            for (int q = 0; q < B; ++q) blk[(size_t)q * W + W - 1] |= (int64_t)ORAG_STATUS_EXCHANGE_TIMEOUT;
            break;
        }
        __nanosleep(100);
    }
}

}  // namespace orag

using namespace orag;

extern "C" size_t orag_exchange_bytes(int n_shards, int max_queries, int fetch_k, int kk)
{
    if (n_shards < 1 || max_queries < 1 || fetch_k < 1 || kk < 1) return 0;
    return x_flag_bytes(n_shards) + (size_t)kXSlots * x_slot_words(n_shards, max_queries, fetch_k, kk) * 8;
}

extern "C" int orag_exchange_alloc(size_t bytes, void **d_buf)
{
    ORAG_REQUIRE(d_buf && bytes > 0, "exchange_alloc");
    // cudaMalloc (not a sub-block of a caching allocator): the IPC handle of the allocation is the buffer itself
    ORAG_CUDA_CHECK(cudaMalloc(d_buf, bytes));
    ORAG_CUDA_CHECK(cudaMemset(*d_buf, 0, bytes));
    ORAG_CUDA_CHECK(cudaDeviceSynchronize());
    return ORAG_OK;
}

extern "C" int orag_exchange_free(void *d_buf)
{
    if (d_buf) ORAG_CUDA_CHECK(cudaFree(d_buf));
    return ORAG_OK;
}

extern "C" int orag_exchange_export(void *d_buf, unsigned char handle[64])
{
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    ORAG_REQUIRE(d_buf && handle, "exchange_export");
    cudaIpcMemHandle_t h;
    ORAG_CUDA_CHECK(cudaIpcGetMemHandle(&h, d_buf));
    memcpy(handle, &h, sizeof(h));
    return ORAG_OK;
}

extern "C" int orag_exchange_open(const unsigned char handle[64], void **d_peer_buf)
{
    ORAG_REQUIRE(handle && d_peer_buf, "exchange_open");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    // maps the peer's allocation into this process and enables peer access between the two devices if needed
    ORAG_CUDA_CHECK(cudaIpcOpenMemHandle(d_peer_buf, h, cudaIpcMemLazyEnablePeerAccess));
    return ORAG_OK;
}

extern "C" int orag_exchange_close(void *d_peer_buf)
{
    if (d_peer_buf) ORAG_CUDA_CHECK(cudaIpcCloseMemHandle(d_peer_buf));
    return ORAG_OK;
}

extern "C" int orag_hybrid_push(const int64_t *d_cos_ids, const double *d_cos_scores, const int64_t *d_bm25_ids,
                                const double *d_bm25_scores, const double *d_bm25_max, const int32_t *d_status,
                                const int32_t *d_status2, int n_queries, int fetch_k, int kk, int rank, int n_shards,
                                int max_queries,
                                void *const *d_peer_bufs, uint64_t seq, void *stream)
{
    ORAG_REQUIRE(d_cos_ids && d_cos_scores && d_bm25_ids && d_bm25_scores && d_bm25_max && d_peer_bufs, "hybrid_push pointers");
    ORAG_REQUIRE(n_shards >= 1 && n_shards <= kXMaxShards && rank >= 0 && rank < n_shards, "hybrid_push shards");
    ORAG_REQUIRE(n_queries >= 1 && n_queries <= max_queries && fetch_k >= 1 && kk >= fetch_k && seq >= 1,
                 "hybrid_push sizes");
    orag::TimelineScope tl(orag::TL_PUSH, (cudaStream_t)stream);
    exchange_push_kernel<<<n_shards, 256, 0, (cudaStream_t)stream>>>(
        d_cos_ids, d_cos_scores, d_bm25_ids, d_bm25_scores, d_bm25_max, d_status, d_status2, n_queries, fetch_k, kk, rank,
        n_shards,
        d_peer_bufs, x_flag_bytes(n_shards), x_slot_words(n_shards, max_queries, fetch_k, kk), (unsigned long long)seq);
    ORAG_LAUNCH_CHECK();
    return ORAG_OK;
}

extern "C" int orag_hybrid_wait(void *d_buf, int n_shards, int max_queries, int n_queries, int fetch_k, int kk,
                                uint64_t seq, int timeout_ms, const int64_t **d_gathered, void *stream)
{
    ORAG_REQUIRE(d_buf && d_gathered, "hybrid_wait pointers");
    ORAG_REQUIRE(n_shards >= 1 && n_shards <= kXMaxShards && n_queries >= 1 && n_queries <= max_queries && fetch_k >= 1 &&
                     kk >= fetch_k && seq >= 1 && timeout_ms > 0,
                 "hybrid_wait sizes");
    const size_t fb = x_flag_bytes(n_shards), sw = x_slot_words(n_shards, max_queries, fetch_k, kk);
    const int W = 2 * fetch_k + 2 * kk + 2;
    orag::TimelineScope tl(orag::TL_WAIT, (cudaStream_t)stream);
    exchange_wait_kernel<<<1, kXMaxShards, 0, (cudaStream_t)stream>>>((uint8_t *)d_buf, n_shards, n_queries, W, fb, sw,
                                                                      (unsigned long long)seq,
                                                                      (unsigned long long)timeout_ms * 1000000ull);
    ORAG_LAUNCH_CHECK();
    *d_gathered = (const int64_t *)((uint8_t *)d_buf + fb) + (size_t)(seq % (uint64_t)kXSlots) * sw;
    return ORAG_OK;
}
