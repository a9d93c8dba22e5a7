// cosine_tc.cu -- tcgen05 / TMA / TMEM similarity scan with fused threshold top-k candidate emission.
//
// One persistent CTA per SM.  D[128 rows x N queries] += A[128 x K] * B[N x K]^T with
//   A = corpus tile  (fp32 straight from HBM, kind::tf32  -- or fp16 / bf16 shadow copy, kind::f16)
//   B = query block  (L2-resident, re-streamed per tile)
// streamed K-chunk by K-chunk (128 bytes of K per stage = one SWIZZLE_128B row) through a
// TMA->smem ring; accumulators live in TMEM (2 stages x 256 columns) so the epilogue of
// tile i overlaps the MMAs of tile i+1.
//
// Three flavours of the same kernel (launch_scan picks):
//   <.., kMma2 = true>  main scans: clusters of two CTAs on adjacent row tiles, ONE tcgen05.mma.cta_group::2
//                       (M = 256 over both SMs) per K step issued by the pair's leader; every CTA loads its rows
//                       and half of the query slab (cta_group::2 TMA, completing on the leader's mbarrier);
//                       six 32 KiB stages
//   cluster2 (runtime)  pairs of 1-SM MMAs, each CTA multicasts half of the query slab to both (ORAG_SCAN_2SM=0)
//   plain               one CTA per tile, four 48 KiB stages: seed / dense passes, small inputs
//
// Warp roles (384 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane),
// warp 2 = TMEM allocator, warp 3 idle, warps 4-11 = epilogue (two warps per TMEM lane quarter,
// each taking half of the query columns).
//
// Epilogue, scan mode: v = acc * inv_norm[row] (= cos * |q|) is compared with a per-query running
// threshold; survivors are appended to the query's candidate list and binned into a per-query
// histogram of approximate cosines from which the threshold (a lower bound of the k-th best
// approximate cosine, minus the first-pass error margin) is tightened with atomicMax.  Every true
// top-k row provably survives (DESIGN.md "exactness of the first pass"); candidates are re-scored
// in float64 by cosine_exact.cu.  Dense mode stores v for every (row, query) instead (seed pass
// over the first rows, tests).
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"
#include "cosine_tc.cuh"

namespace orag {
namespace tc {

// ---- PTX wrappers ----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok = 0;
    const long long t0 = clock64();
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (ok) break;
        if (clock64() - t0 > 4000000000ll) {  // ~2 s: a protocol bug must fault, not hang the GPU
            printf("orag cosine_tc: mbarrier timeout block %d thread %d bar %u parity %u\n", blockIdx.x, threadIdx.x, bar,
                   parity);
            __trap();
        }
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar,
                                            uint64_t policy)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(dst), "l"((uint64_t)(uintptr_t)map), "r"(bar), "r"(c0), "r"(c1), "l"(policy)
        : "memory");
}
// same load, delivered to the same shared-memory offset of every CTA in `cta_mask` (and signalling the mbarrier at the
// same offset in each of them)
__device__ __forceinline__ void tma_load_2d_multicast(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar,
                                                      uint16_t cta_mask, uint64_t policy)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster.L2::cache_hint"
        " [%0], [%1, {%4, %5}], [%2], %3, %6;"
        ::"r"(dst), "l"((uint64_t)(uintptr_t)map), "r"(bar), "h"(cta_mask), "r"(c0), "r"(c1), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void tc_commit_multicast(uint32_t bar, uint16_t cta_mask)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(cta_mask) : "memory");
}
// 2-SM variants: the load lands in the executing CTA's shared memory but signals the mbarrier of the pair's leader
// (even CTA: the peer bit of the shared::cluster address is cleared)
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar,
                                                uint64_t policy)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(dst), "l"((uint64_t)(uintptr_t)map), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void tc_commit2_multicast(uint32_t bar, uint16_t cta_mask)
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(cta_mask) : "memory");
}
template <bool kBf16>
__device__ __forceinline__ void tc_mma2(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum)
{
    if constexpr (kBf16) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
            : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
            : "memory");
    }
}
// arrive on the barrier at the same shared-memory offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t cta)
{
    asm volatile(
        "{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
        ::"r"(bar), "r"(cta) : "memory");
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
template <bool kBf16>
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum)
{
    if constexpr (kBf16) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
            : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
            : "memory");
    }
}
// K-major, SWIZZLE_128B operand tile: rows at 128-byte pitch, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);  // start address
    d |= (uint64_t)1 << 16;                        // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;              // stride byte offset
    d |= (uint64_t)1 << 46;                        // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                        // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ int cos_bin(float c)
{
    int b = (int)floorf((c + 1.0f) * (kHistBins / 2));
    return b < 0 ? 0 : (b > kHistBins - 1 ? kHistBins - 1 : b);
}
// lower edge of bin b: a valid lower bound of every value binned into b (float rounding at the
// edge is covered by the slack inside the first-pass margin)
__device__ __forceinline__ float bin_floor(int b) { return (float)b * (2.0f / kHistBins) - 1.0f; }

// Candidate emission + threshold tightening (rare path).  Every 8th emission of a query re-derives
// its threshold from the histogram: largest bin b with count(bins >= b) >= k.  The scan walks down
// from the top in batches of 16 bins fetched with four independent 16-byte loads.
__device__ __noinline__ void emit_candidate(const ScanParams &p, int q, int32_t local_row, float v)
{
    const uint32_t slot = atomicAdd(p.cnt + q, 1u);
    if (slot < (uint32_t)p.cap) {
        p.cand[(int64_t)q * p.cap + slot] = local_row;
        if (p.cand_v) p.cand_v[(int64_t)q * p.cap + slot] = v;
    }
    if (p.fixed_thr) return;
    const int bin = cos_bin(v * p.inv_qnorm[q]);
    uint32_t *h = p.hist + (int64_t)q * kHistBins;
    atomicAdd(h + bin, 1u);
    if ((slot & 7u) != 7u) return;
    const uint4 *h4 = reinterpret_cast<const uint4 *>(h);
    uint32_t acc = 0;
    int found = -1;
    for (int c = kHistBins / 4 - 1; c >= 3 && found < 0; c -= 4) {
        const uint4 w0 = __ldcg(h4 + c), w1 = __ldcg(h4 + c - 1), w2 = __ldcg(h4 + c - 2), w3 = __ldcg(h4 + c - 3);
        const uint32_t vals[16] = {w0.w, w0.z, w0.y, w0.x, w1.w, w1.z, w1.y, w1.x,
                                   w2.w, w2.z, w2.y, w2.x, w3.w, w3.z, w3.y, w3.x};
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            acc += vals[j];
            if (found < 0 && acc >= (uint32_t)p.k) found = 4 * c + 3 - j;
        }
    }
    if (found >= 1) atomicMax(p.thr_key + q, float_to_ordered(bin_floor(found)));
}

// tile index -> (first corpus row, first query row).  Scan mode: tiles walk the row range, one query
// block.  Pair mode (all-pairs search over one matrix): query block qb = rows [256 qb, 256 qb + 256) is
// matched against the row tiles below its end, tiles [qb (qb+1), (qb+1)(qb+2)) of a triangular enumeration.
__device__ __forceinline__ void decode_tile(const ScanParams &p, int tile, int64_t &row0, int &qoff)
{
    if (!p.pair_mode) {
        row0 = p.row_begin + (int64_t)tile * kTileM;
        qoff = 0;
        return;
    }
    int qb = (int)((sqrtf(1.0f + 4.0f * (float)tile) - 1.0f) * 0.5f);
    while ((qb + 1) * (qb + 2) <= tile) ++qb;
    while (qb * (qb + 1) > tile) --qb;
    row0 = (int64_t)(tile - qb * (qb + 1)) * kTileM;
    qoff = qb * kMaxN;
}

// kMma2: the cta_group::2 flavour is a separate instantiation -- a kernel that contains cta_group::2 instructions can
// only be launched as clusters of an even size
// (Tried and dropped: capping the registers at 80 and a five-stage ring for the 2-SM flavour, to leave room on every SM for
// the tail kernels of the previous batch next to the scan and the background BM25 CTA -- the main scan lost 5 % and 16 %.)
template <bool kBf16, bool kMma2, int kSt>
__global__ void __launch_bounds__(kThreads, 1)
cosine_scan_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                   const __grid_constant__ ScanParams p)
{
    // 2-SM flavour: every CTA stages only half of each query slab (32 KiB stages instead of 48 KiB)
    constexpr int kBB = kMma2 ? kBBytes / 2 : kBBytes;
    constexpr int kRing = kSt * (kABytes + kBB);
    static_assert(kSt <= 8 && scan_smem_bytes(kMma2, kSt) <= 227 * 1024, "ring");
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *smem_a = smem;
    uint8_t *smem_b = smem + kSt * kABytes;
    uint64_t *bars = (uint64_t *)(smem + kRing);
    uint64_t *full_bar = bars;                    // [kSt] (room for 8 stages)
    uint64_t *empty_bar = bars + 8;               // [kSt]
    uint64_t *tmem_full = bars + 16;              // [2]
    uint64_t *tmem_empty = bars + 18;             // [2]
    uint32_t *tmem_ptr = (uint32_t *)(bars + 20);
    float *thrv_s = (float *)(smem + kRing + 256);  // [2][kMaxN]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)(uintptr_t)&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)(uintptr_t)&map_b) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kSt; ++s) {
            mbar_init(smem_u32(full_bar + s), 1);
            mbar_init(smem_u32(empty_bar + s), (p.cluster2 && !p.mma2) ? 2 : 1);  // multicast pair: both MMAs release a slot
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(smem_u32(tmem_full + s), 1);
            mbar_init(smem_u32(tmem_empty + s), (p.mma2 ? 2 : 1) * (kEpiThreads / 32));  // mma2: both CTAs' epilogues
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        if constexpr (kMma2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)),
                         "r"(kTmemCols)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)),
                         "r"(kTmemCols)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    if (p.cluster2) cluster_sync_all();  // the peer's barriers exist before anything is multicast to them
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    // cluster2: both CTAs of a pair run the same number of tiles (they share every query slab); tiles past the
    // end are all-zero (TMA out-of-bounds fill) and emit nothing
    const int tile_limit = p.cluster2 ? (int)((p.num_tiles + gridDim.x - 1) / gridDim.x) * (int)gridDim.x : p.num_tiles;
    const uint32_t cta_rank = p.cluster2 ? cluster_ctarank() : 0u;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            const uint64_t pol_stream = 0x12F0000000000000ull;  // evict-first: corpus is read once
            const uint64_t pol_keep = 0x14F0000000000000ull;    // evict-last: queries are re-read per tile
            const uint32_t tx = kABytes + (uint32_t)p.umma_n * 128u;
            int stage = 0;
            uint32_t phase = 0;
            const int half_rows = p.umma_n >> 1;  // cluster2: this CTA fetches query rows [rank * half, +half)
            for (int tile = blockIdx.x; tile < tile_limit; tile += gridDim.x) {
                int64_t row0;
                int qoff;
                decode_tile(p, tile, row0, qoff);
                const int row = (int)row0;
                for (int kc = 0; kc < p.k_chunks; ++kc) {
                    mbar_wait(smem_u32(empty_bar + stage), phase ^ 1u);
                    const uint32_t fb = smem_u32(full_bar + stage);
                    if constexpr (kMma2) {
                        // both CTAs load their own rows and their half of the query slab (same offsets in both), all
                        // four loads complete on the LEADER's barrier
                        if (cta_rank == 0) mbar_expect_tx(fb, 2u * (kABytes + (uint32_t)half_rows * 128u));
                        tma_load_2d_2sm(smem_u32(smem_a + stage * kABytes), &map_a, kc * p.chunk_elems, row, fb, pol_stream);
                        tma_load_2d_2sm(smem_u32(smem_b + stage * kBB), &map_b, kc * p.chunk_elems,
                                        qoff + (int)cta_rank * half_rows, fb, pol_keep);
                        if (++stage == kSt) { stage = 0; phase ^= 1u; }
                        continue;
                    }
                    mbar_expect_tx(fb, tx);
                    tma_load_2d(smem_u32(smem_a + stage * kABytes), &map_a, kc * p.chunk_elems, row, fb, pol_stream);
                    if (p.cluster2)
                        tma_load_2d_multicast(smem_u32(smem_b + stage * kBB) + cta_rank * (uint32_t)half_rows * 128u,
                                              &map_b, kc * p.chunk_elems, qoff + (int)cta_rank * half_rows, fb,
                                              (uint16_t)3, pol_keep);
                    else
                        tma_load_2d(smem_u32(smem_b + stage * kBB), &map_b, kc * p.chunk_elems, qoff, fb, pol_keep);
                    if (++stage == kSt) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (mma2: the pair's leader only) =====================
        if (lane == 0 && !(p.mma2 && cta_rank != 0)) {
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < tile_limit; tile += gridDim.x, ++it) {
                const int as = it & 1;
                mbar_wait(smem_u32(tmem_empty + as), (((uint32_t)it >> 1) & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * kMaxN);
                for (int kc = 0; kc < p.k_chunks; ++kc) {
                    mbar_wait(smem_u32(full_bar + stage), phase);
                    tc_fence_after();
                    const uint64_t adesc = umma_desc(smem_u32(smem_a + stage * kABytes));
                    const uint64_t bdesc = umma_desc(smem_u32(smem_b + stage * kBB));
                    if constexpr (kMma2) {
#pragma unroll
                        for (int k4 = 0; k4 < 4; ++k4)
                            tc_mma2<kBf16>(d_tmem, adesc + (uint64_t)(k4 * 2), bdesc + (uint64_t)(k4 * 2), p.idesc,
                                           (uint32_t)((kc | k4) != 0));
                        tc_commit2_multicast(smem_u32(empty_bar + stage), (uint16_t)3);  // both CTAs' slots
                    } else {
#pragma unroll
                        for (int k4 = 0; k4 < 4; ++k4)  // 4 x 32 bytes of K per 128-byte chunk
                            tc_mma<kBf16>(d_tmem, adesc + (uint64_t)(k4 * 2), bdesc + (uint64_t)(k4 * 2), p.idesc,
                                          (uint32_t)((kc | k4) != 0));
                        // frees the smem slot when these MMAs retire (multicast pair: in both CTAs)
                        if (p.cluster2) tc_commit_multicast(smem_u32(empty_bar + stage), (uint16_t)3);
                        else tc_commit(smem_u32(empty_bar + stage));
                    }
                    if (++stage == kSt) { stage = 0; phase ^= 1u; }
                }
                // accumulator ready for the epilogue (mma2: of both CTAs)
                if constexpr (kMma2) tc_commit2_multicast(smem_u32(tmem_full + as), (uint16_t)3);
                else tc_commit(smem_u32(tmem_full + as));
            }
        }
    } else if (warp >= kEpiWarp0) {
        // ===================== epilogue =====================
        const int ew = warp - kEpiWarp0;
        const int quarter = ew & 3;  // == warp % 4: the TMEM lane quarter this warp may read
        const int half = ew >> 2;    // which 128 query columns
        const int et = threadIdx.x - kEpiWarp0 * 32;  // 0..255
        int it = 0;
        for (int tile = blockIdx.x; tile < tile_limit; tile += gridDim.x, ++it) {
            const int as = it & 1;
            float *thrv = thrv_s + as * kMaxN;
            int64_t row0;
            int qoff;
            decode_tile(p, tile, row0, qoff);
            if (!p.dense) {
                float t = INFINITY;
                if (qoff + et < p.n_queries) {
                    float thr = ordered_to_float(__ldcg(p.thr_key + qoff + et));
                    t = (thr - p.margin) * p.qnorm[qoff + et];
                    if (!(t == t)) t = INFINITY;  // |q| == 0: nothing passes (handled by the seed pass)
                }
                thrv[et] = t;
                asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
            }
            mbar_wait(smem_u32(tmem_full + as), ((uint32_t)it >> 1) & 1u);
            tc_fence_after();
            const int64_t row = row0 + quarter * 32 + lane;
            const bool valid = row < p.row_end;
            const float scale = valid ? __ldg(p.inv_norm + row) : 0.f;
            const int32_t local_row = (int32_t)row;
            const uint32_t taddr0 = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * kMaxN + half * 128);
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                const int col0 = half * 128 + c * 32;
                if (col0 >= p.umma_n) break;
                uint32_t r[32];
                tmem_ld32(taddr0 + (uint32_t)(c * 32), r);
                tmem_ld_wait();
                if (p.dense == 2) {
                    // transposed (seed pass): dense_out[q * dense_ld + row] -- the 32 lanes of a warp hold 32 consecutive
                    // rows, so every store instruction writes one 128-byte line
                    if (valid) {
                        float *dst = p.dense_out + (int64_t)col0 * p.dense_ld + (row - p.row_begin);
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (col0 + j < p.n_queries) dst[(int64_t)j * p.dense_ld] = __uint_as_float(r[j]) * scale;
                    }
                } else if (p.dense) {
                    if (valid) {
                        float4 *dst = reinterpret_cast<float4 *>(p.dense_out + (row - p.row_begin) * kMaxN + col0);
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            dst[j] = make_float4(__uint_as_float(r[4 * j]) * scale, __uint_as_float(r[4 * j + 1]) * scale,
                                                 __uint_as_float(r[4 * j + 2]) * scale,
                                                 __uint_as_float(r[4 * j + 3]) * scale);
                    }
                } else {
                    bool any = false;
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 t4 = *reinterpret_cast<const float4 *>(thrv + col0 + j);
                        any |= (__uint_as_float(r[j]) * scale >= t4.x);
                        any |= (__uint_as_float(r[j + 1]) * scale >= t4.y);
                        any |= (__uint_as_float(r[j + 2]) * scale >= t4.z);
                        any |= (__uint_as_float(r[j + 3]) * scale >= t4.w);
                    }
                    if (any && valid) {
#pragma unroll  // static register indices: a dynamic r[j] would push the whole tile row to local memory
                        for (int j = 0; j < 32; ++j) {
                            const int q = qoff + col0 + j;
                            const float v = __uint_as_float(r[j]) * scale;
                            // pair mode keeps i < j only (each unordered pair once, no self pairs)
                            if (q < p.n_queries && v >= thrv[col0 + j] && (!p.pair_mode || local_row < q))
                                emit_candidate(p, q, local_row, v);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (p.mma2 && cta_rank != 0) mbar_arrive_remote(smem_u32(tmem_empty + as), 0u);  // the leader issues the MMAs
                else mbar_arrive(smem_u32(tmem_empty + as));
            }
        }
    }

    tc_fence_before();
    if (p.cluster2) cluster_sync_all();  // no CTA leaves while its peer can still write its shared memory / barriers
    else __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        if constexpr (kMma2)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// ---- seed finalisation ----------------------------------------------------------------------
constexpr int kSeedPerThread = 32;  // seed values per thread of seed_finalize_kernel: seed_ld = 256 * 32 rows

// One CTA per query over the dense first-pass values of the first `n_seed` rows (transposed: seed_t[q * seed_ld + row]):
// histogram -> threshold; survivors -> candidate list.
__global__ void __launch_bounds__(256) seed_finalize_kernel(const float *__restrict__ seed_t, int seed_ld, int n_seed,
                                                           int n_queries,
                                                           int k, float margin, const float *__restrict__ qnorm,
                                                           const float *__restrict__ inv_qnorm,
                                                           uint32_t *__restrict__ thr_key, uint32_t *__restrict__ cnt,
                                                           uint32_t *__restrict__ hist, int32_t *__restrict__ cand,
                                                           float *__restrict__ cand_v, int cap)
{
    __shared__ uint32_t h[kHistBins];
    __shared__ uint32_t s_cnt;
    __shared__ float s_thr;
    const int q = blockIdx.x;
    for (int b = threadIdx.x; b < kHistBins; b += blockDim.x) h[b] = 0;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    const float *__restrict__ seed = seed_t + (int64_t)q * seed_ld;  // this query's values, one per seed row
    const float qn = qnorm[q], iqn = inv_qnorm[q];
    if (qn == 0.f) {
        // zero query: every cosine is 0.0 -> the answer is rows 0..k-1 (ties by id); nothing else may pass
        const int m = n_seed < k ? n_seed : k;
        for (int r = threadIdx.x; r < m && r < cap; r += blockDim.x) {
            cand[(int64_t)q * cap + r] = r;
            cand_v[(int64_t)q * cap + r] = 0.f;
        }
        for (int b = threadIdx.x; b < kHistBins; b += blockDim.x) hist[(int64_t)q * kHistBins + b] = 0;
        if (threadIdx.x == 0) {
            cnt[q] = m;
            thr_key[q] = float_to_ordered(INFINITY);
        }
        return;
    }
    // Every thread keeps its 32 consecutive seed values in registers (eight 16-byte loads in flight at once; three
    // strided passes over global memory cost 36 us, this 22): all later passes run on registers.
    float v[kSeedPerThread];
    const int r0 = threadIdx.x * kSeedPerThread;
#pragma unroll
    for (int j = 0; j < kSeedPerThread; j += 4) {
        const float4 x = *reinterpret_cast<const float4 *>(seed + r0 + j);
        v[j] = x.x; v[j + 1] = x.y; v[j + 2] = x.z; v[j + 3] = x.w;
    }
#pragma unroll
    for (int j = 0; j < kSeedPerThread; ++j)
        if (r0 + j >= n_seed) v[j] = -INFINITY;   // (rows the seed pass did not write)
    // Only the top of the distribution matters.  Each thread's largest value is a distinct row, so the k-th largest of
    // the 256 per-thread maxima (`low`) is a lower bound of the k-th best seed value: rows below it need no histogram
    // entry (the bins under `low` stay under-counted, which can only make a threshold more conservative), and the
    // shared-memory atomics no longer pile up on the few bins around cosine 0.
    __shared__ float s_max[256];
    __shared__ float s_low;
    float mine = -INFINITY;
#pragma unroll
    for (int j = 0; j < kSeedPerThread; ++j) mine = fmaxf(mine, v[j]);
    s_max[threadIdx.x] = mine;
    if (threadIdx.x == 0) s_low = -INFINITY;
    __syncthreads();
    if (k <= (int)blockDim.x) {
        int above = 0;  // maxima ranked before mine under (value desc, thread asc)
        for (int t = 0; t < (int)blockDim.x; ++t) {
            const float o = s_max[t];
            above += (o > mine || (o == mine && t < (int)threadIdx.x)) ? 1 : 0;
        }
        if (above == k - 1) s_low = mine;  // exactly one thread; -inf when fewer than k threads saw a row
    }
    __syncthreads();
    const float low = s_low;
#pragma unroll
    for (int j = 0; j < kSeedPerThread; ++j)
        if (v[j] >= low && r0 + j < n_seed) atomicAdd(&h[cos_bin(v[j] * iqn)], 1u);
    __syncthreads();
    if (threadIdx.x < 32) {
        // largest bin b with count(bins >= b) >= k: lane i sums the 16 bins [512 - 16 (i + 1), 512 - 16 i), a warp scan
        // finds the lane whose chunk crosses k, that lane walks its chunk
        const int lane = threadIdx.x;
        const int hi = kHistBins - 16 * lane;   // one past my top bin
        uint32_t sum = 0;
        for (int b = hi - 1; b >= hi - 16; --b) sum += h[b];
        uint32_t incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += up;
        }
        const unsigned crossed = __ballot_sync(0xffffffffu, incl >= (uint32_t)k);
        if (crossed == 0) {
            if (lane == 0) s_thr = -INFINITY;   // fewer than k rows seen: keep everything
        } else if (lane == __ffs(crossed) - 1) {
            uint32_t acc = incl - sum;
            int b = hi - 1;
            for (; b >= hi - 16; --b) {
                acc += h[b];
                if (acc >= (uint32_t)k) break;
            }
            s_thr = (b >= 1) ? bin_floor(b) : -INFINITY;
        }
    }
    __syncthreads();
    const float thr = s_thr;
    const float tv = (thr - margin) * qn;
    for (int b = threadIdx.x; b < kHistBins; b += blockDim.x) hist[(int64_t)q * kHistBins + b] = h[b];
#pragma unroll
    for (int j = 0; j < kSeedPerThread; ++j) {
        if (v[j] >= tv && r0 + j < n_seed) {
            const uint32_t slot = atomicAdd(&s_cnt, 1u);
            if (slot < (uint32_t)cap) {
                cand[(int64_t)q * cap + slot] = r0 + j;
                cand_v[(int64_t)q * cap + slot] = v[j];
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        cnt[q] = s_cnt;
        thr_key[q] = float_to_ordered(thr);
    }
}

__global__ void query_norms_kernel(const double *__restrict__ sq, int n, float *__restrict__ qnorm,
                                   float *__restrict__ inv_qnorm)
{
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    double m = sqrt(sq[q]);
    qnorm[q] = (float)m;
    inv_qnorm[q] = m > 0.0 ? (float)(1.0 / m) : 0.f;
}

// ---- host side ---------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)sym;
    }
    return fn;
}

static int make_map(CUtensorMap *map, const void *base, bool bf16, bool f16, int64_t rows, int dim, int box_rows)
{
    EncodeTiledFn enc = get_encode();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled entry point not available");
        return ORAG_ECUDA;
    }
    const int eb = bf16 ? 2 : 4;
    cuuint64_t gdim[2] = {(cuuint64_t)dim, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)dim * eb};
    cuuint32_t box[2] = {(cuuint32_t)(128 / eb), (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, bf16 ? (f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16)
                               : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                     const_cast<void *>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld dim=%d box_rows=%d)", (int)r,
                  (long long)rows, dim, box_rows);
        return ORAG_ECUDA;
    }
    return ORAG_OK;
}

static uint32_t make_idesc(bool bf16, bool f16, int umma_n)
{
    const uint32_t fmt = bf16 ? (f16 ? 0u : 1u) : 2u;  // F16F32Format: F16 = 0, BF16 = 1, TF32 = 2
    return (1u << 4) /* D = f32 */ | (fmt << 7) | (fmt << 10) | ((uint32_t)(umma_n >> 3) << 17) |
           ((uint32_t)(kTileM >> 4) << 24);
}

// Launch one scan over rows [row_begin, row_end) of the operand behind `a_base`.
int launch_scan(bool bf16, const void *a_base, int64_t a_rows, const void *q_base, int dim, ScanParams p,
                cudaStream_t st)
{
    if (p.row_end <= p.row_begin) return ORAG_OK;
    CUtensorMap map_a, map_b;
    p.umma_n = (p.n_queries + 15) / 16 * 16;
    if (p.umma_n < 16) p.umma_n = 16;
    int rc = make_map(&map_a, a_base, bf16, p.f16 != 0, a_rows, dim, kTileM);
    if (rc) return rc;
    p.num_tiles = (int)((p.row_end - p.row_begin + kTileM - 1) / kTileM);
    if (p.pair_mode) {
        // triangular enumeration: query block qb owns the 2 (qb + 1) row tiles below its end -- an even count starting
        // at an even index, so the two CTAs of a pair always share their query block
        const int nb = (p.n_queries + kMaxN - 1) / kMaxN;
        p.num_tiles = nb * (nb + 1);
        p.umma_n = kMaxN;
    }
    // clusters of two CTAs sharing the query slabs: main scans only, enough tiles to keep every pair busy
    // (ORAG_SCAN_CLUSTER: 0 = never, 1 = default rule, 2 = whenever there are two tiles -- tests)
    static const int cluster_mode = getenv("ORAG_SCAN_CLUSTER") ? atoi(getenv("ORAG_SCAN_CLUSTER")) : 1;
    const int min_tiles = cluster_mode >= 2 ? 2 : 4 * sm_count();
    p.cluster2 = (cluster_mode && !p.dense && p.umma_n >= 32 && p.num_tiles >= min_tiles) ? 1 : 0;
    // the pairs issue one tcgen05.mma.cta_group::2 per K step unless ORAG_SCAN_2SM=0 (then: multicast pairs)
    static const int mma2_mode = getenv("ORAG_SCAN_2SM") ? atoi(getenv("ORAG_SCAN_2SM")) : 1;
    p.mma2 = (mma2_mode && p.cluster2) ? 1 : 0;
    rc = make_map(&map_b, q_base, bf16, p.f16 != 0, p.n_queries, dim, p.cluster2 ? p.umma_n / 2 : p.umma_n);
    if (rc) return rc;
    p.chunk_elems = bf16 ? 64 : 32;
    p.k_chunks = dim / p.chunk_elems;
    p.idesc = make_idesc(bf16, p.f16 != 0, p.umma_n);
    if (p.mma2) p.idesc = (p.idesc & ~(0x1Fu << 24)) | ((uint32_t)(2 * kTileM >> 4) << 24);  // M = 256 over the pair
    int grid = p.num_tiles < sm_count() ? p.num_tiles : sm_count();
    if (p.cluster2) grid &= ~1;
    if (!p.dense) profile_mark(0, 0, st);
    auto launch = [&](auto kernel, size_t smem_bytes) -> int {
        ORAG_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)grid);
        cfg.blockDim = dim3(kThreads);
        cfg.dynamicSmemBytes = smem_bytes;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = p.cluster2 ? 2 : 1;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        ORAG_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kernel, map_a, map_b, p));
        return ORAG_OK;
    };
    if (p.mma2)
        rc = bf16 ? launch(cosine_scan_kernel<true, true, kStages2>, scan_smem_bytes(true, kStages2))
                  : launch(cosine_scan_kernel<false, true, kStages2>, scan_smem_bytes(true, kStages2));
    else
        rc = bf16 ? launch(cosine_scan_kernel<true, false, kStages>, scan_smem_bytes(false, kStages))
                  : launch(cosine_scan_kernel<false, false, kStages>, scan_smem_bytes(false, kStages));
    if (rc) return rc;
    if (!p.dense) profile_mark(0, 1, st);
    ORAG_LAUNCH_CHECK();
    return ORAG_OK;
}

int launch_seed_finalize(const float *seed, int seed_ld, int n_seed, int n_queries, int k, float margin, const float *qnorm,
                         const float *inv_qnorm, uint32_t *thr_key, uint32_t *cnt, uint32_t *hist, int32_t *cand,
                         float *cand_v, int cap, cudaStream_t st)
{
    ORAG_REQUIRE(seed_ld == 256 * kSeedPerThread && n_seed <= seed_ld && (reinterpret_cast<uintptr_t>(seed) & 15) == 0,
                 "seed_finalize: seed buffer of 256 x 32 values per query");
    seed_finalize_kernel<<<n_queries, 256, 0, st>>>(seed, seed_ld, n_seed, n_queries, k, margin, qnorm, inv_qnorm, thr_key, cnt,
                                                    hist, cand, cand_v, cap);
    ORAG_LAUNCH_CHECK();
    return ORAG_OK;
}

int launch_query_norms(const double *sq, int n, float *qnorm, float *inv_qnorm, cudaStream_t st)
{
    query_norms_kernel<<<(n + 255) / 256, 256, 0, st>>>(sq, n, qnorm, inv_qnorm);
    ORAG_LAUNCH_CHECK();
    return ORAG_OK;
}

}  // namespace tc
}  // namespace orag
