// bm25_ms.cu -- BM25 top-k, fp32 MaxScore first pass + exact float64 re-score.
//
// Same result as bm25.cu (bit-exact float64 scores of rank_bm25.BM25Okapi.get_scores + the glue at
// rag/retrieval.py:324-347), an order of magnitude fewer instructions per (query, tile) pair.
//
// First pass.  The index carries a second view of the postings, (doc_in_tile << 16) | fp16(r) with
// r = tf*(k1+1)/(tf + t4[dl]) rounded to nearest, so a posting's contribution is one fp32 multiply
// w*r (w = fp32 idf, duplicates of a query term merged into one weight) and needs no document-length
// or table lookup.  Every approximate score s~ satisfies |s~ - s| <= eps*s with
// eps = 2^-11 (fp16 r) + (n_terms + 2) * 2^-24 (fp32 weight, products, sums) < 5e-4: all terms are positive.
//
// MaxScore.  prepare_queries_kernel sorts a query's terms by ascending upper bound
// ub_t = w_t * max_r(t) (max over the shard's postings of t) and stores the inclusive prefix sums.
// With the query's running threshold thr (a lower bound of the k-th best s~ seen so far, from the same
// log-scale histogram bm25.cu uses) the terms whose prefix sum stays below thr' = thr * (1 - 2^-9) are
// "non-essential": a document that contains only those cannot reach thr', so their posting runs are never
// scattered.  Only the essential runs are added into the warp's fp32 accumulators (shared-memory atomics,
// any order); each touched document is then claimed once (atomicExch resets the accumulator), and the
// non-essential terms are looked up for it by binary search in the staged runs, most valuable term first,
// stopping as soon as partial + remaining upper bound < thr'.  Documents that end at s~ >= thr' are emitted.
//
// Superset argument (as for the cosine first pass): let S_k be the true k-th best score and S~_k the k-th
// best approximate score.  At most k-1 docs have s > S_k, so S~_k <= S_k (1 + eps); thr <= S~_k always.  A
// true top-k doc (or a tie with the k-th) has s~ >= S_k (1 - eps) >= thr (1 - eps) / (1 + eps) > thr', so it
// is never pruned and always emitted; 2^-9 also covers the rounding of the bounds themselves (prefix sums
// are rounded up, thr' down).  ms_finalize_kernel keeps the candidates with s~ >= final thr', re-scores them
// with score_doc() (float64, query order, duplicates twice) and selects by (score desc, id asc).
//
// Work decomposition and pipeline follow bm25.cu: one warp per (tile, query slice), four stages in flight
// (descriptor of query i+3, run offsets of query i+2, 16-byte cp.async of the runs of query i+1 into a
// per-warp two-buffer ring, query i consumed), no CTA barrier in the loop.
#include <cuda_fp16.h>

#include "bm25_shared.cuh"

namespace orag {
namespace bm25 {

constexpr int kMsWarps = 8;
constexpr int kMsThreads = kMsWarps * 32;
constexpr int kMsTerms = 32;      // scoring terms per query on this path (one per lane)
constexpr int kMsStage = 736;     // postings per staging buffer (two per warp); multiple of 4
constexpr int kMsSurvCap = 2048;  // candidates re-scored per query
constexpr float kMsGuard = 1.0f - 1.0f / 512.0f;
constexpr float kMsMinR = 6.103515625e-05f;  // smallest normal fp16

struct MsParams {
    orag_bm25_index_t ix;
    const int32_t *q_terms;  // [n_queries, max_terms] original query tokens (exact re-score)
    const int32_t *q_lens;
    int n_queries;
    int max_terms;
    int k;
    // prepared queries: terms sorted by ascending upper bound
    int32_t *qd;       // [n_queries, kMsTerms, 4] = {term, fp32 weight (idf * multiplicity), inclusive prefix sum
                       //  of the upper bounds (rounded up), number of terms}
    // threshold state (same scheme as bm25.cu)
    unsigned long long *thr_bits;
    uint32_t *cnt;
    uint32_t *hist;
    uint32_t *topbin;
    int32_t *cand_doc;  // [n_queries, cap]
    float *cand_val;    // [n_queries, cap] approximate scores
    int cap;
    int32_t *surv_doc;  // [n_queries, kMsSurvCap]
    double *surv_score; // [n_queries, kMsSurvCap]
    int q_split;
};

// One thread per query: drop OOV / zero-idf tokens, merge duplicates, sort by upper bound.
__global__ void prepare_queries_kernel(const __grid_constant__ MsParams p)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= p.n_queries) return;
    int32_t term[kMsTerms];
    int mult[kMsTerms];
    float w[kMsTerms], ub[kMsTerms];
    int n = 0;
    const int nt = min(p.q_lens[q], min(p.max_terms, kMsTerms));
    for (int i = 0; i < nt; ++i) {
        const int t = p.q_terms[(int64_t)q * p.max_terms + i];
        if (t < 0 || t >= p.ix.vocab) continue;
        if (p.ix.d_idf[t] == 0.0) continue;
        int j = 0;
        while (j < n && term[j] != t) ++j;
        if (j < n) ++mult[j];
        else { term[n] = t; mult[n] = 1; ++n; }
    }
    for (int j = 0; j < n; ++j) {
        w[j] = __double2float_rn(__dmul_rn(p.ix.d_idf[term[j]], (double)mult[j]));
        ub[j] = __fmul_ru(w[j], p.ix.d_term_max_r[term[j]]);
    }
    // insertion sort by ascending upper bound (terms that occur nowhere in the shard, ub = 0, come first)
    for (int a = 1; a < n; ++a) {
        const int32_t t = term[a];
        const float ww = w[a], uu = ub[a];
        int b = a - 1;
        while (b >= 0 && ub[b] > uu) {
            term[b + 1] = term[b]; w[b + 1] = w[b]; ub[b + 1] = ub[b];
            --b;
        }
        term[b + 1] = t; w[b + 1] = ww; ub[b + 1] = uu;
    }
    float pre = 0.f;
    for (int j = 0; j < kMsTerms; ++j) {
        if (j < n) pre = __fadd_ru(pre, ub[j]);
        int4 v;
        v.x = j < n ? term[j] : -1;
        v.y = __float_as_int(j < n ? w[j] : 0.f);
        v.z = __float_as_int(pre);
        v.w = n;
        reinterpret_cast<int4 *>(p.qd)[(int64_t)q * kMsTerms + j] = v;
    }
}

__device__ __noinline__ void ms_emit(const MsParams &p, int q, int32_t doc, float v)
{
    const uint32_t slot = atomicAdd(p.cnt + q, 1u);
    if (slot < (uint32_t)p.cap) {
        p.cand_doc[(int64_t)q * p.cap + slot] = doc;
        p.cand_val[(int64_t)q * p.cap + slot] = v;
    }
    uint32_t *h = p.hist + (int64_t)q * kHistBins;
    const int bin = score_bin((double)v);
    atomicAdd(h + bin, 1u);
    const int old_top = (int)atomicMax(p.topbin + q, (uint32_t)bin);
    if ((slot & 7u) != 7u) return;
    tighten_threshold(h, old_top > bin ? old_top : bin, p.k, p.thr_bits + q);
}

__device__ __forceinline__ float thr_to_float(unsigned long long bits)
{
    // largest fp32 not above the float64 threshold, times the guard, rounded down
    return __fmul_rd(__double2float_rd(__longlong_as_double((long long)bits)), kMsGuard);
}

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
                 "l"(gmem_src)
                 : "memory");
}
__device__ __forceinline__ void cp_async16_s(uint32_t smem_dst, const void *gmem_src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void ms_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void ms_cp_wait_1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

__device__ __forceinline__ float post_r(uint32_t post)
{
    return __half2float(__ushort_as_half((unsigned short)(post & 0xFFFFu)));
}

struct MsDesc {  // lane i = i-th term of the prepared query
    int term;
    float w, pre;
    int n, q;
};
struct MsRun {
    int rel;     // first posting of the run, relative to the tile's 16-byte aligned base pointer
    int len;     // holds the END offset until the run is staged
    float w, pre;
    int n, q;
};
struct MsStaged {
    MsRun r;
    int soff;    // position of the run's first posting in the staging buffer
    int slen;    // postings available in the buffer (<= len); the rest is read from global memory
};

__global__ void __launch_bounds__(kMsThreads, 2) bm25_ms_kernel(const __grid_constant__ MsParams p)
{
    extern __shared__ uint4 ms_smem[];
    const int T = p.ix.tile_docs;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const size_t per_warp = (size_t)T * 4 + 2 * kMsStage * 4;
    uint8_t *mine = reinterpret_cast<uint8_t *>(ms_smem) + wib * per_warp;
    float *acc = reinterpret_cast<float *>(mine);                          // [T], all zero between queries
    uint32_t *stage = reinterpret_cast<uint32_t *>(mine + (size_t)T * 4);  // [2][kMsStage]
    const uint32_t stage_s = (uint32_t)__cvta_generic_to_shared(stage);
    for (int i = lane; i < T; i += 32) acc[i] = 0.f;
    __syncwarp();

    const int V1 = p.ix.vocab + 1;
    const int nq = p.n_queries;
    const int warps_total = gridDim.x * kMsWarps;
    const unsigned FULL = 0xffffffffu;
    const int S = p.q_split;
    const int4 *qd = reinterpret_cast<const int4 *>(p.qd);

    for (int item = blockIdx.x * kMsWarps + wib; item < p.ix.n_tiles * S; item += warps_total) {
        const int tile = item / S;
        const int part = item - tile * S;
        const int qi0 = (int)(((int64_t)nq * part) / S);
        const int qi1 = (int)(((int64_t)nq * (part + 1)) / S);
        const int64_t base_doc = (int64_t)tile * T;
        const int64_t tile_g0 = p.ix.d_tile_base[tile];
        const int tb = (int)(tile_g0 & 3);
        const uint32_t *tp_al = p.ix.d_postings_r16 + (tile_g0 - tb);  // 16-byte aligned
        const int32_t *toff = p.ix.d_tile_term_off + (int64_t)tile * V1;
        const int q_shift = (int)(((int64_t)tile * 7919) % nq);

        // ---- stage A: the prepared query (one 16-byte load per lane)
        auto stage_desc = [&](int qi) -> MsDesc {
            MsDesc d;
            d.term = -1; d.w = 0.f; d.pre = 0.f; d.n = 0; d.q = 0;
            if (qi < qi1) {
                int q = qi + q_shift;
                if (q >= nq) q -= nq;
                d.q = q;
                const int4 v = __ldg(qd + q * kMsTerms + lane);
                d.term = v.x;
                d.w = __int_as_float(v.y);
                d.pre = __int_as_float(v.z);
                d.n = v.w;
            }
            return d;
        };
        // ---- stage B: run offsets of every term inside this tile
        auto stage_run = [&](const MsDesc &d) -> MsRun {
            MsRun r;
            r.rel = 0; r.len = 0; r.w = d.w; r.pre = d.pre; r.n = d.n; r.q = d.q;
            if (lane < d.n) {
                r.rel = __ldg(toff + d.term);
                r.len = __ldg(toff + d.term + 1);
            }
            return r;
        };
        // ---- stage C: lay the runs out in the staging buffer (16-byte aligned source windows) and copy
        auto stage_posts = [&](MsRun run, int qi) -> MsStaged {
            MsStaged s;
            run.len -= run.rel;
            run.rel += tb;
            s.r = run;
            const int shift = run.rel & 3;
            const int alen = run.len > 0 ? ((shift + run.len + 3) & ~3) : 0;
            int incl = alen;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int v = __shfl_up_sync(FULL, incl, d);
                if (lane >= d) incl += v;
            }
            const int excl = incl - alen;
            const int avail = max(0, min(alen, kMsStage - excl));  // multiple of 4
            s.soff = excl + shift;
            s.slen = max(0, min(run.len, avail - shift));
            const int n16 = s.slen > 0 ? (avail >> 2) : 0;
            const uint32_t dst_s = stage_s + (uint32_t)((qi & 1) * kMsStage + excl) * 4u;
            const uint32_t *src = tp_al + (run.rel & ~3);
            // every lane copies the first two 16-byte chunks of its own run (most runs are that short) ...
            if (n16 > 0) cp_async16_s(dst_s, src);
            if (n16 > 1) cp_async16_s(dst_s + 16, src + 4);
            // ... and the long runs are copied by the whole warp, 512 bytes per instruction
            unsigned active = __ballot_sync(FULL, n16 > 2);
            while (active) {
                const int i = __ffs(active) - 1;
                active &= active - 1;
                const uint32_t d0 = __shfl_sync(FULL, dst_s, i);
                const unsigned long long s0 = __shfl_sync(FULL, (unsigned long long)(uintptr_t)src, i);
                const int cnt = __shfl_sync(FULL, n16, i);
                const uint32_t *sp = reinterpret_cast<const uint32_t *>((uintptr_t)s0);
#pragma unroll 1
                for (int c = 2 + lane; c < cnt; c += 32) cp_async16_s(d0 + 16u * c, sp + 4 * c);
            }
            ms_cp_commit();
            return s;
        };

        // prologue: fill the pipeline
        MsStaged cur = stage_posts(stage_run(stage_desc(qi0)), qi0);
        MsRun runB = stage_run(stage_desc(qi0 + 1));
        MsDesc descA = stage_desc(qi0 + 2);

        for (int qi = qi0; qi < qi1; ++qi) {
            // the current query's threshold: issued first, consumed after the stage-C work below
            unsigned long long thr_bits = 0;
            if (cur.r.n > 0) thr_bits = __ldcg(p.thr_bits + cur.r.q);
            const MsStaged nxt = stage_posts(runB, qi + 1);
            runB = stage_run(descA);
            descA = stage_desc(qi + 3);

            const uint32_t *buf = stage + (qi & 1) * kMsStage;
            ms_cp_wait_1();
            __syncwarp();

            const int n = cur.r.n;  // warp-uniform
            const float thr = thr_to_float(thr_bits);
            const int q = cur.r.q;
            const bool mine_ok = lane < n && cur.r.len > 0;
            const int n_ne = __popc(__ballot_sync(FULL, lane < n && cur.r.pre < thr));  // a prefix of the lanes
            const unsigned ess = __ballot_sync(FULL, mine_ok && lane >= n_ne);
            if (ess) {
                const unsigned non = __ballot_sync(FULL, mine_ok && lane < n_ne);
                // per-run views are broadcast with shuffles; postings beyond the staged part come from global memory
#define MS_RUN_VIEW(i)                                                         \
    const int len_ = __shfl_sync(FULL, cur.r.len, i);                          \
    const int so_ = __shfl_sync(FULL, cur.soff, i);                            \
    const int sl_ = __shfl_sync(FULL, cur.slen, i);                            \
    const int rel_ = __shfl_sync(FULL, cur.r.rel, i)
#define MS_POST(j) ((j) < sl_ ? buf[so_ + (j)] : __ldg(tp_al + rel_ + (j)))
                // ---- S: scatter the essential runs (docs are distinct inside a run: plain read-modify-write)
                for (unsigned a = ess; a; a &= a - 1) {
                    const int i = __ffs(a) - 1;
                    MS_RUN_VIEW(i);
                    const float w = __shfl_sync(FULL, cur.r.w, i);
#pragma unroll 1
                    for (int j = lane; j < len_; j += 32) {
                        const uint32_t post = MS_POST(j);
                        float *slot = acc + (post >> 16);
                        *slot = __fadd_rn(*slot, __fmul_rn(w, post_r(post)));
                    }
                    __syncwarp();
                }
                // ---- N: the non-essential runs only complete documents an essential term has touched
                for (unsigned a = non; a; a &= a - 1) {
                    const int i = __ffs(a) - 1;
                    MS_RUN_VIEW(i);
                    const float w = __shfl_sync(FULL, cur.r.w, i);
#pragma unroll 1
                    for (int j = lane; j < len_; j += 32) {
                        const uint32_t post = MS_POST(j);
                        float *slot = acc + (post >> 16);
                        const float v = *slot;
                        if (v != 0.f) *slot = __fadd_rn(v, __fmul_rn(w, post_r(post)));
                    }
                    __syncwarp();
                }
                // ---- X: claim every touched doc once (first essential run that holds it), reset, emit
                for (unsigned a = ess; a; a &= a - 1) {
                    const int i = __ffs(a) - 1;
                    MS_RUN_VIEW(i);
#pragma unroll 1
                    for (int j = lane; j < len_; j += 32) {
                        const uint32_t d = MS_POST(j) >> 16;
                        const float v = acc[d];
                        if (v != 0.f) {
                            acc[d] = 0.f;
                            if (v >= thr) ms_emit(p, q, (int32_t)(base_doc + d), v);
                        }
                    }
                    __syncwarp();
                }
#undef MS_RUN_VIEW
#undef MS_POST
            }
            cur = nxt;
        }
        // drain the (empty) copy groups still outstanding before the buffers are reused by the next item
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
    }
}

// One CTA per query: prune the candidates by the final threshold, re-score the survivors exactly, select.
__global__ void __launch_bounds__(1024) ms_finalize_kernel(const __grid_constant__ MsParams p, int64_t doc_id_base,
                                                         int normalize, int64_t *__restrict__ out_ids,
                                                         double *__restrict__ out_scores, double *__restrict__ out_max,
                                                         int32_t *__restrict__ status)
{
    __shared__ Pick scratch[32];
    __shared__ double dscratch[32];
    __shared__ uint32_t s_n;
    const int q = blockIdx.x;
    const int k = p.k;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    uint32_t n = p.cnt[q];
    bool overflow = false;
    if (n > (uint32_t)p.cap) {
        overflow = true;
        n = p.cap;
    }
    const float thr = thr_to_float(p.thr_bits[q]);
    const int32_t *cd = p.cand_doc + (int64_t)q * p.cap;
    const float *cv = p.cand_val + (int64_t)q * p.cap;
    int32_t *sd = p.surv_doc + (int64_t)q * kMsSurvCap;
    double *ss = p.surv_score + (int64_t)q * kMsSurvCap;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        if (cv[i] >= thr) {
            const uint32_t slot = atomicAdd(&s_n, 1u);
            if (slot < (uint32_t)kMsSurvCap) sd[slot] = cd[i];
        }
    }
    __syncthreads();
    uint32_t ns = s_n;
    if (ns > (uint32_t)kMsSurvCap) {
        overflow = true;
        ns = kMsSurvCap;
    }
    if (overflow && status && threadIdx.x == 0) status[q] |= ORAG_STATUS_OVERFLOW;
    const int32_t *terms = p.q_terms + (int64_t)q * p.max_terms;
    const int nt = min(p.q_lens[q], p.max_terms);
    for (uint32_t i = threadIdx.x; i < ns; i += blockDim.x) ss[i] = score_doc(p.ix, terms, nt, sd[i]);
    __syncthreads();
    select_from_list(p.ix, terms, nt, k, sd, ss, ns, doc_id_base, normalize, out_ids + (int64_t)q * k,
                     out_scores + (int64_t)q * k, out_max ? out_max + q : nullptr, scratch, dscratch);
}

__global__ void ms_init_state_kernel(unsigned long long *thr_bits, uint32_t *cnt, uint32_t *hist, uint32_t *topbin,
                                     int n_queries)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t total = (int64_t)n_queries * kHistBins;
    if (i < total) hist[i] = 0;
    if (i < n_queries) {
        thr_bits[i] = 1ull;  // smallest positive double: "score > 0"
        cnt[i] = 0;
        topbin[i] = 0;
    }
}

static int ms_cap(int n_queries)
{
    // ~128 MiB of candidate storage shared by the batch, at least 8192 slots per query
    int64_t cap = ((int64_t)128 << 20) / 8 / (n_queries > 0 ? n_queries : 1);
    if (cap < 8192) cap = 8192;
    if (cap > (1 << 22)) cap = 1 << 22;
    return (int)cap;
}

bool ms_eligible(const orag_bm25_index_t *ix, int max_terms, int flags)
{
    return ix->d_postings_r16 != nullptr && ix->d_term_max_r != nullptr && !ix->has_negative_idf &&
           max_terms <= kMsTerms && !(flags & (ORAG_BM25_EXACT_TILES | ORAG_BM25_FORCE_DENSE)) && ix->tile_docs <= 4096;
}

struct MsCarve {
    MsParams p;
    size_t bytes;
};

static MsCarve ms_carve(void *base, int n_queries)
{
    MsCarve c{};
    uint8_t *w = (uint8_t *)base;
    auto take = [&](size_t n) {
        uint8_t *r = w;
        w += align_up(n, 256);
        return r;
    };
    const size_t nq = (size_t)n_queries;
    const int cap = ms_cap(n_queries);
    c.p.cap = cap;
    c.p.thr_bits = (unsigned long long *)take(nq * 8);
    c.p.cnt = (uint32_t *)take(nq * 4);
    c.p.topbin = (uint32_t *)take(nq * 4);
    c.p.hist = (uint32_t *)take(nq * kHistBins * 4);
    c.p.qd = (int32_t *)take(nq * kMsTerms * 16);
    c.p.cand_doc = (int32_t *)take(nq * cap * 4);
    c.p.cand_val = (float *)take(nq * cap * 4);
    c.p.surv_doc = (int32_t *)take(nq * kMsSurvCap * 4);
    c.p.surv_score = (double *)take(nq * kMsSurvCap * 8);
    c.bytes = (size_t)(w - (uint8_t *)base);
    return c;
}

size_t ms_workspace_bytes(const orag_bm25_index_t *ix, int n_queries)
{
    (void)ix;
    return ms_carve(nullptr, n_queries).bytes;
}

int ms_topk(const orag_bm25_index_t *ix, int64_t doc_id_base, const int32_t *d_query_terms,
            const int32_t *d_query_lens, int n_queries, int max_terms, int k, int normalize, int64_t *d_out_ids,
            double *d_out_scores, double *d_out_max, int32_t *d_out_status, void *d_workspace, cudaStream_t st)
{
    MsParams p = ms_carve(d_workspace, n_queries).p;
    p.ix = *ix;
    p.q_terms = d_query_terms;
    p.q_lens = d_query_lens;
    p.n_queries = n_queries;
    p.max_terms = max_terms;
    p.k = k;
    {
        int64_t total = (int64_t)n_queries * kHistBins;
        ms_init_state_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(p.thr_bits, p.cnt, p.hist, p.topbin,
                                                                              n_queries);
        ORAG_LAUNCH_CHECK();
        prepare_queries_kernel<<<(n_queries + 127) / 128, 128, 0, st>>>(p);
        ORAG_LAUNCH_CHECK();
    }
    if (ix->n_tiles > 0) {
        const size_t smem = (size_t)kMsWarps * ((size_t)ix->tile_docs * 4 + 2 * kMsStage * 4);
        int per_sm = (int)((226 * 1024) / (smem + 1024));
        if (per_sm < 1) per_sm = 1;
        if (per_sm > 2) per_sm = 2;
        const int lim = sm_count() * per_sm;
        int64_t want = (int64_t)4 * lim * kMsWarps;
        int split = (int)((want + ix->n_tiles - 1) / ix->n_tiles);
        if (split > 16) split = 16;
        if (split > n_queries) split = n_queries;
        if (split < 1) split = 1;
        p.q_split = split;
        const int64_t items = (int64_t)ix->n_tiles * split;
        int grid = (int)((items + kMsWarps - 1) / kMsWarps);
        if (grid > lim) grid = lim;
        ORAG_CUDA_CHECK(cudaFuncSetAttribute(bm25_ms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        profile_mark(1, 0, st);
        bm25_ms_kernel<<<grid, kMsThreads, smem, st>>>(p);
        profile_mark(1, 1, st);
        ORAG_LAUNCH_CHECK();
    }
    ms_finalize_kernel<<<n_queries, 1024, 0, st>>>(p, doc_id_base, normalize, d_out_ids, d_out_scores, d_out_max,
                                                   d_out_status);
    ORAG_LAUNCH_CHECK();
    return ORAG_OK;
}

}  // namespace bm25
}  // namespace orag
