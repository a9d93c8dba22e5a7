// bm25_ms.cu -- BM25 top-k, fp32 MaxScore first pass over big tiles + exact float64 re-score.
//
// Same result as bm25.cu (bit-exact float64 scores of rank_bm25.BM25Okapi.get_scores + the glue at
// rag/retrieval.py:324-347) at a fraction of the instructions per posting.
//
// First-pass view (built by bm25_build.cu).  A second tiling of the postings (8192-doc tiles by default, up to
// 16384): (doc_in_tile << 16) | fp16(r), r = tf*(k1+1)/(tf + t4[dl]) rounded to nearest, so a posting's
// contribution is one fp32 multiply w*r (w = fp32 idf, duplicates of a query term merged into one weight) and needs
// no document-length or table lookup.  Every (tile, term) run starts on a 16-byte boundary and is padded to a
// multiple of four postings with copies of its last doc carrying impact 0, so the long runs are only ever moved as
// uint4.  Every approximate score s~ satisfies |s~ - s| <= eps*s with eps = 2^-11 (fp16 r) + (n_terms + 2) * 2^-24
// (fp32 weight, products, sums) < 5e-4: all terms are positive.
//
// MaxScore.  prepare_queries_kernel sorts a query's terms by ascending upper bound ub_t = w_t * max_r(t) (max over
// the shard's postings of t) and stores the inclusive prefix sums.  With the query's running threshold thr (a lower
// bound of the k-th best s~ seen so far, from the same log-scale histogram bm25.cu uses) the terms whose prefix sum
// stays below thr' = thr * (1 - 2^-9) are "non-essential": a document that contains only those cannot reach thr'.
//
// One warp works on one (query, tile) pair at a time with three small shared-memory structures -- a bitmap of the
// tile's docs, per-word prefix popcounts, and a compact fp32 accumulator indexed by a doc's RANK among the marked docs
// (sized by the docs a query touches, not by the tile) -- and reads the postings straight from L2, where a prefetch
// issued one pair ahead has put them (the whole run of a term that is essential by the threshold of that moment, the
// first 512 bytes of the others):
//   E1  mark the docs of the essential runs in the bitmap (shared-memory atomicOr)
//   R   prefix popcounts -> rank(d)
//   E2  add the essential contributions w * r into acc[rank(d)], run by run (docs are distinct inside a run, so plain
//       read-modify-writes suffice), tracking the best partial score
//   N   the non-essential runs are streamed with 16-byte loads (two in flight per lane), most valuable term first; a
//       posting only matters when its doc is marked (one shared-memory word test), in which case its contribution
//       completes acc[rank(d)].  Before every run: if the best partial score (kept up to date with one warp-wide
//       REDUX per run) plus the upper bounds of the remaining terms stays below thr', nothing of this pair can be
//       emitted and the remaining (longest) runs are not read at all
//   X   only if the best score reaches thr': lanes walk the RANKS, and a score that clears thr' gets its doc id back
//       from the prefix counts (binary search + find-n-th-set-bit) and is emitted.  Otherwise acc is just zeroed.
// A pair whose essential postings outnumber the accumulator slots (cold start: thr = 0 makes every term essential;
// or a very dense query) takes the same steps in doc sub-ranges: runs are doc-sorted, so a sub-range is a contiguous
// slice of every run (one binary search per run and boundary); the threshold is re-read between sub-ranges, and
// before a cold sub-range emits anything it derives a LOCAL threshold -- the k-th largest of the 32 lane maxima of
// its own scores, a valid lower bound of the global k-th best -- and publishes it.
// (An earlier build staged the essential runs in shared memory with cp.async; dropping that buffer doubled the
// accumulator, which is what lets the tiles grow to 8192 docs: 1.20 -> 0.93 ms at 10M docs x 256 queries.)
//
// Superset argument (as for the cosine first pass): let S_k be the true k-th best score and S~_k the k-th best
// approximate score.  At most k-1 docs have s > S_k, so S~_k <= S_k (1 + eps); thr <= S~_k always.  A true top-k doc
// (or a tie with the k-th) has s~ >= S_k (1 - eps) >= thr (1 - eps) / (1 + eps) > thr', so it is never pruned and
// always emitted; 2^-9 also covers the rounding of the bounds themselves (prefix sums are rounded up, thr' down).
// ms_finalize_kernel keeps the candidates with s~ >= final thr', re-scores them with score_doc() (float64, query
// order, duplicates twice) and selects by (score desc, id asc).
//
// Work items are (tile, slice of the query batch), handed out through an atomic counter; the prepared query of pair
// i+2 and the run offsets (+ an L2 prefetch of the run heads) of pair i+1 are in flight while pair i is consumed.
#include <cuda_fp16.h>

#include "bm25_shared.cuh"

namespace orag {
namespace bm25 {

constexpr int kMsWarps = 16;      // warps per CTA of the stand-alone configuration (2 CTAs per SM)
constexpr int kMsThreads = kMsWarps * 32;
constexpr int kMsBgWarps = 8;     // ... and of the background configuration (see ms_topk)
constexpr int kMsTerms = 32;      // scoring terms per query on this path (one per lane)
constexpr int kMsWarpBytes = 7168;    // shared memory per warp, stand-alone: 2 CTAs x (16 x 7168 + 1 KiB) fit one SM
constexpr int kMsBgWarpBytes = 3840;  // background: 8 warps < 31 KB next to a resident cosine scan CTA
constexpr int kMsMaxTile = 16384;
constexpr int kMsSurvCap = 2048;  // candidates re-scored per query
constexpr int kMsContrib = 4096;  // (survivor, token) contributions staged in shared memory by ms_finalize_kernel
constexpr float kMsGuard = 1.0f - 1.0f / 512.0f;

// per-warp shared memory: compact accumulator [cap] | bitmap [words] | per-word ranks [words].  cap = docs one
// (query, tile sub-range) may mark; a multiple of 4.
__host__ __device__ inline int ms_cap(int words, int warp_bytes)
{
    const int left = warp_bytes - words * 6;
    return left < 256 ? 0 : (left / 4) & ~3;
}

struct MsParams {
    orag_bm25_index_t ix;
    const int32_t *q_terms;  // [n_queries, max_terms] original query tokens (exact re-score)
    const int32_t *q_lens;
    int n_queries;
    int max_terms;
    int k;
    int32_t *qd;       // [n_queries, kMsTerms, 4] = {term, fp32 weight (idf * multiplicity), inclusive prefix sum
                       //  of the upper bounds (rounded up), number of terms}; terms by ascending upper bound
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
    uint32_t *work;     // next work item
    int32_t *status;    // [n_queries] or null
    int q_split;
    int n_items;
    int warp_bytes;     // kMsWarpBytes or kMsBgWarpBytes
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
    // Warm start (index->d_term_kth_r): K >= k docs contain term t with an impact of at least r_K(t), and each of them
    // scores at least fl(w_t * r_K(t)) in this pass's own arithmetic (every contribution is >= 0 and rounding is
    // monotone), so the best of these is a lower bound of the k-th best approximate score -- a valid threshold.
    if (p.ix.d_term_kth_r && p.k <= 128) {
        const int level = p.k <= 10 ? 0 : p.k <= 16 ? 1 : p.k <= 32 ? 2 : p.k <= 64 ? 3 : 4;
        float t0 = 0.f;
        for (int j = 0; j < n; ++j)
            t0 = fmaxf(t0, __fmul_rn(w[j], p.ix.d_term_kth_r[(int64_t)level * p.ix.vocab + term[j]]));
        if (t0 > 0.f) p.thr_bits[q] = (unsigned long long)__double_as_longlong((double)t0);
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

__device__ __forceinline__ float post_r(uint32_t post)
{
    return __half2float(__ushort_as_half((unsigned short)(post & 0xFFFFu)));
}

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// warp-wide maximum of non-negative floats: their bit patterns order like unsigned integers -> one REDUX
__device__ __forceinline__ float warp_max_nonneg(float v)
{
    return __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(v)));
}

struct MsDesc {  // lane i = i-th term of the prepared query
    int term;
    float w, pre;
    int n, q;
};
struct MsRun {   // lane i = i-th term's run in the current tile, in units of four postings (uint4)
    int rel4, len4;
    float w, pre;
    int n, q;
};

// kT: docs per first-pass tile when known at compile time (8192 / 4096 in the stand-alone configuration: every
// shared-memory offset becomes an immediate and the bitmap loops unroll), 0 = read tile size and layout from the call.
template <int kT>
__global__ void __launch_bounds__(kMsThreads, 2) bm25_ms_kernel(const __grid_constant__ MsParams p)
{
    extern __shared__ uint4 ms_smem[];
    const int T = kT ? kT : p.ix.fp_tile_docs;
    const int words = (T + 31) >> 5;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int warp_bytes = kT ? kMsWarpBytes : p.warp_bytes;
    const int cap = ms_cap(words, warp_bytes);
    uint8_t *mine = reinterpret_cast<uint8_t *>(ms_smem) + (size_t)wib * warp_bytes;
    float *acc = reinterpret_cast<float *>(mine);                              // [cap], zero between pairs
    uint32_t *bm = reinterpret_cast<uint32_t *>(mine + (size_t)cap * 4);       // [words] marked docs, zero between pairs
    uint16_t *pre = reinterpret_cast<uint16_t *>(bm + words);                  // [words] rank of a word's bit 0
    for (int i = lane; i < words; i += 32) bm[i] = 0u;
    for (int i = lane; i < cap; i += 32) acc[i] = 0.f;
    __syncwarp();

    const int V1 = p.ix.vocab + 1;
    const int nq = p.n_queries;
    const unsigned FULL = 0xffffffffu;
    const int S = p.q_split;
    const int4 *qd = reinterpret_cast<const int4 *>(p.qd);
    const int wpl = (words + 31) >> 5;  // bitmap words per lane in the rank pass
    const bool wpl4 = words == 128;     // 4096-doc tiles (bm and pre are 16-byte aligned: cap is a multiple of 4)
    const bool wpl8 = words == 256;     // 8192-doc tiles
    const int k = p.k;

    auto rank_of = [&](uint32_t d) -> int {
        const uint32_t wd = bm[d >> 5];
        return (int)pre[d >> 5] + __popc(wd & ((1u << (d & 31)) - 1u));
    };

    for (;;) {
        int item = 0;
        if (lane == 0) item = (int)atomicAdd(p.work, 1u);
        item = __shfl_sync(FULL, item, 0);
        if (item >= p.n_items) break;
        const int tile = item / S;
        const int part = item - tile * S;
        const int qi0 = (int)(((int64_t)nq * part) / S);
        const int qi1 = (int)(((int64_t)nq * (part + 1)) / S);
        const int64_t base_doc = (int64_t)tile * T;
        const uint32_t *tp = p.ix.d_postings_r16 + p.ix.d_fp_tile_base[tile];  // 16-byte aligned (padded layout)
        const uint4 *tp4 = reinterpret_cast<const uint4 *>(tp);
        const int32_t *toff = p.ix.d_fp_tile_term_off + (int64_t)tile * V1;
        // stagger the query order across tiles so a query's threshold is established by few warps
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
        // ---- stage B: run offsets inside this tile; L2 prefetch of the runs one pair ahead: the head (512 bytes) of
        // every run, and all of a run (up to 4 KiB) whose term is essential by the threshold as of now
        auto stage_run = [&](const MsDesc &d) -> MsRun {
            MsRun r;
            r.rel4 = 0; r.len4 = 0; r.w = d.w; r.pre = d.pre; r.n = d.n; r.q = d.q;
            if (lane < d.n) {
                const int o0 = __ldg(toff + d.term), o1 = __ldg(toff + d.term + 1);
                r.rel4 = o0 >> 2;
                r.len4 = (o1 - o0) >> 2;
                if (r.len4 > 0) {
                    const float thr_now = thr_to_float(__ldcg(p.thr_bits + d.q));
                    const int lines = !(d.pre < thr_now) ? min((r.len4 + 7) >> 3, 32) : min((r.len4 + 7) >> 3, 4);
                    for (int l = 0; l < lines; ++l) prefetch_l2(tp4 + r.rel4 + 8 * l);
                }
            }
            return r;
        };

        MsRun cur = stage_run(stage_desc(qi0));
        MsRun nxt = stage_run(stage_desc(qi0 + 1));
        MsDesc descA = stage_desc(qi0 + 2);

        for (int qi = qi0; qi < qi1; ++qi) {
            const int n = cur.n;  // warp-uniform
            const int q = cur.q;
            const bool mine_ok = lane < n && cur.len4 > 0;
            float thr = 0.f;
            if (n > 0) thr = thr_to_float(__ldcg(p.thr_bits + q));
            int n_ne = __popc(__ballot_sync(FULL, lane < n && cur.pre < thr));  // a prefix of the lanes
            unsigned ess = __ballot_sync(FULL, mine_ok && lane >= n_ne);
            if (ess) {
                // essential postings of the pair; more than the accumulator holds -> doc sub-ranges
                const int etot = __reduce_add_sync(FULL, (mine_ok && lane >= n_ne) ? cur.len4 * 4 : 0);
                // (the doc range that is split is the tile's OWN, rounded up to a power of two: the last tile of a shard may
                // hold a fraction of T docs, and equal slices of T would put all of its postings into the first few
                // sub-ranges; the rounding costs at most a factor of two, which the cap / 2 target absorbs)
                const int64_t left = (int64_t)p.ix.n_docs - (int64_t)tile * T;
                int nsub = 1, sub_docs = left >= T ? T : 1 << (32 - __clz((int)left - 1 | 31));
                while (nsub * (etot > cap ? (cap >> 1) : cap) < etot && sub_docs > 32) {
                    nsub <<= 1;
                    sub_docs >>= 1;
                }
                int cpos = 0;  // this lane's run: first posting of the current sub-range
                for (int sub = 0; sub < nsub; ++sub) {
                    int v_rel, v_len;   // this lane's run in the (sub-)range: posting slice relative to tp
                    int n4_lo, n4_cnt;  // ... and the 16-byte aligned superset of it the N phase streams
                    if (nsub == 1) {
                        v_rel = cur.rel4 * 4;
                        v_len = mine_ok ? cur.len4 * 4 : 0;
                        n4_lo = cur.rel4;
                        n4_cnt = cur.len4;
                    } else {
                        if (sub > 0) {
                            // a pair that needs several sub-ranges is a cold one: its own emissions have raised the
                            // threshold since, so re-read it and re-partition (fewer essential runs, fewer emissions)
                            thr = thr_to_float(__ldcg(p.thr_bits + q));
                            n_ne = __popc(__ballot_sync(FULL, lane < n && cur.pre < thr));
                            ess = __ballot_sync(FULL, mine_ok && lane >= n_ne);
                            if (!ess) break;  // nothing left in this tile can reach the threshold
                        }
                        int s_end = cur.len4 * 4;
                        if (sub + 1 < nsub && mine_ok) {
                            // first posting of this lane's run with doc >= the sub-range's upper boundary
                            const uint32_t bound = (uint32_t)((sub + 1) * sub_docs);
                            int lo = cpos, hi = s_end;
                            const uint32_t *run = tp + cur.rel4 * 4;
                            while (lo < hi) {
                                const int mid = (lo + hi) >> 1;
                                if ((__ldg(run + mid) >> 16) < bound) lo = mid + 1; else hi = mid;
                            }
                            s_end = lo;
                        }
                        v_rel = cur.rel4 * 4 + cpos;
                        v_len = mine_ok ? s_end - cpos : 0;
                        n4_lo = cur.rel4 + (cpos >> 2);
                        n4_cnt = v_len > 0 ? ((s_end + 3) >> 2) - (cpos >> 2) : 0;
                        cpos = s_end;
                    }
#define MS_VIEW(i)                                  \
    const int len_ = __shfl_sync(FULL, v_len, i);   \
    const uint32_t *gp_ = tp + __shfl_sync(FULL, v_rel, i)
                    // ---- E1: mark the docs of the essential runs (L2-resident: prefetched one pair ahead)
                    for (unsigned a = ess; a; a &= a - 1) {
                        const int i = __ffs(a) - 1;
                        MS_VIEW(i);
                        // four loads per lane in flight before the first is consumed: one L2 round trip per 128 postings
#pragma unroll 1
                        for (int j0 = lane; j0 < len_; j0 += 128) {
                            uint32_t pz[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) pz[u] = j0 + 32 * u < len_ ? __ldg(gp_ + j0 + 32 * u) : 0u;
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                if (j0 + 32 * u < len_) {
                                    const uint32_t d = pz[u] >> 16;
                                    atomicOr(bm + (d >> 5), 1u << (d & 31));
                                }
                            }
                        }
                    }
                    __syncwarp();
                    // ---- R: prefix popcounts (lane l owns words [l*wpl, (l+1)*wpl))
                    int mycnt = 0;
                    uint4 w4 = make_uint4(0u, 0u, 0u, 0u), x4 = w4;
                    if (wpl4) {
                        // 4096-doc tiles: one 16-byte load, one 8-byte store per lane
                        w4 = reinterpret_cast<const uint4 *>(bm)[lane];
                        mycnt = __popc(w4.x) + __popc(w4.y) + __popc(w4.z) + __popc(w4.w);
                    } else if (wpl8) {
                        w4 = reinterpret_cast<const uint4 *>(bm)[2 * lane];
                        x4 = reinterpret_cast<const uint4 *>(bm)[2 * lane + 1];
                        mycnt = __popc(w4.x) + __popc(w4.y) + __popc(w4.z) + __popc(w4.w) + __popc(x4.x) + __popc(x4.y) +
                                __popc(x4.z) + __popc(x4.w);
                    } else {
                        for (int t = 0; t < wpl; ++t) {
                            const int wi = lane * wpl + t;
                            if (wi < words) mycnt += __popc(bm[wi]);
                        }
                    }
                    int incl = mycnt;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int v = __shfl_up_sync(FULL, incl, o);
                        if (lane >= o) incl += v;
                    }
                    const int marked = __shfl_sync(FULL, incl, 31);
                    const int my_base = incl - mycnt;
                    if (wpl4 || wpl8) {
                        const uint32_t p0 = (uint32_t)my_base, p1 = p0 + __popc(w4.x), p2 = p1 + __popc(w4.y),
                                       p3 = p2 + __popc(w4.z);
                        if (wpl4) {
                            reinterpret_cast<uint2 *>(pre)[lane] = make_uint2(p0 | (p1 << 16), p2 | (p3 << 16));
                        } else {
                            const uint32_t p4 = p3 + __popc(w4.w), p5 = p4 + __popc(x4.x), p6 = p5 + __popc(x4.y),
                                           p7 = p6 + __popc(x4.z);
                            reinterpret_cast<uint4 *>(pre)[lane] =
                                make_uint4(p0 | (p1 << 16), p2 | (p3 << 16), p4 | (p5 << 16), p6 | (p7 << 16));
                        }
                    } else {
                        int run_sum = my_base;
                        for (int t = 0; t < wpl; ++t) {
                            const int wi = lane * wpl + t;
                            if (wi < words) {
                                pre[wi] = (uint16_t)run_sum;
                                run_sum += __popc(bm[wi]);
                            }
                        }
                    }
                    __syncwarp();
                    const bool fits = marked <= cap;
                    float mx = 0.f;
                    bool reachable = fits;
                    if (fits) {
                        // ---- E2: essential contributions into the compact accumulator.  Docs are distinct inside a run
                        // (a padding posting repeats the last doc with impact 0: skipped) and runs are applied one at a
                        // time, so plain read-modify-writes suffice.
                        float lmax = 0.f;
                        for (unsigned a = ess; a; a &= a - 1) {
                            const int i = __ffs(a) - 1;
                            MS_VIEW(i);
                            const float w = __shfl_sync(FULL, cur.w, i);
#pragma unroll 1
                            for (int j0 = lane; j0 < len_; j0 += 128) {
                                uint32_t pz[4];
#pragma unroll
                                for (int u = 0; u < 4; ++u) pz[u] = j0 + 32 * u < len_ ? __ldg(gp_ + j0 + 32 * u) : 0u;
#pragma unroll
                                for (int u = 0; u < 4; ++u) {
                                    const uint32_t post = pz[u];
                                    if (post & 0xFFFFu) {  // (a padding posting, or past the end of the run: impact 0)
                                        float *slot = acc + rank_of(post >> 16);
                                        const float nv = __fadd_rn(*slot, __fmul_rn(w, post_r(post)));
                                        *slot = nv;
                                        lmax = fmaxf(lmax, nv);
                                    }
                                }
                            }
                            __syncwarp();
                        }
                        mx = warp_max_nonneg(lmax);  // best partial score so far
                        // ---- N: non-essential runs complete the marked docs only, most valuable term first.  Before
                        // every run: if even the best partial score plus everything the remaining terms could add stays
                        // below thr', nothing of this (sub-)range can be emitted and the remaining (longest) runs are
                        // not read.
                        for (unsigned a = __ballot_sync(FULL, mine_ok && lane < n_ne); a;) {
                            const int i = 31 - __clz(a);
                            a &= ~(1u << i);
                            const float rest = __shfl_sync(FULL, cur.pre, i);  // upper bounds of terms 0..i
                            if (__fadd_ru(mx, rest) < thr) {
                                reachable = false;
                                break;
                            }
                            const int body4 = __shfl_sync(FULL, n4_cnt, i);
                            const uint4 *g4 = tp4 + __shfl_sync(FULL, n4_lo, i);
                            const float w = __shfl_sync(FULL, cur.w, i);
                            float hmax = 0.f;
                            uint4 v0 = make_uint4(0, 0, 0, 0), v1 = v0;
                            if (lane < body4) v0 = __ldg(g4 + lane);
                            if (lane + 32 < body4) v1 = __ldg(g4 + lane + 32);
#pragma unroll 1
                            for (int c = lane; c < body4; c += 64) {
                                const uint4 u0 = v0, u1 = v1;
                                if (c + 64 < body4) v0 = __ldg(g4 + c + 64);
                                if (c + 96 < body4) v1 = __ldg(g4 + c + 96);
                                // test all eight postings first (branch-free), then visit the rare hits in one divergent
                                // loop: ~3 % of the postings hit, but some lane of the warp does in most groups of 32, so
                                // a branch per posting would run the update path almost every time
                                auto bit = [&](uint32_t post) -> uint32_t {
                                    return (bm[post >> 21] >> ((post >> 16) & 31u)) & 1u;
                                };
                                uint32_t hits = bit(u0.x) | (bit(u0.y) << 1) | (bit(u0.z) << 2) | (bit(u0.w) << 3);
                                if (c + 32 < body4)
                                    hits |= (bit(u1.x) << 4) | (bit(u1.y) << 5) | (bit(u1.z) << 6) | (bit(u1.w) << 7);
                                while (hits) {
                                    const int b = __ffs(hits) - 1;
                                    hits &= hits - 1;
                                    const uint4 u = (b & 4) ? u1 : u0;
                                    const uint32_t lo2 = (b & 1) ? u.y : u.x, hi2 = (b & 1) ? u.w : u.z;
                                    const uint32_t post = (b & 2) ? hi2 : lo2;
                                    float *slot = acc + rank_of(post >> 16);
                                    const float nv = __fadd_rn(*slot, __fmul_rn(w, post_r(post)));
                                    *slot = nv;
                                    hmax = fmaxf(hmax, nv);
                                }
                            }
                            __syncwarp();
                            mx = fmaxf(mx, warp_max_nonneg(hmax));
                        }
                        // ---- cold pair: before emitting, derive a threshold from this sub-range alone.  Every lane
                        // takes the maximum of a strided subset of the marked docs' scores; the k-th largest of the 32
                        // lane maxima is the score of k distinct docs' worth of evidence, i.e. a lower bound of the k-th
                        // best approximate score overall -- publish it and emit only what clears it.
                        if (reachable && (nsub > 1 || thr == 0.f) && k <= 32 && marked >= 2 * k) {
                            float lm = 0.f;
                            for (int r = lane; r < marked; r += 32) lm = fmaxf(lm, acc[r]);
                            float kth = 0.f;
                            for (int r = 0; r < k; ++r) {
                                const float m = warp_max_nonneg(fmaxf(lm, 0.f));
                                kth = m;
                                const unsigned who = __ballot_sync(FULL, lm == m);
                                if (lane == __ffs(who) - 1) lm = -1.f;
                            }
                            if (kth > 0.f) {
                                const unsigned long long kb = (unsigned long long)__double_as_longlong((double)kth);
                                if (lane == 0) atomicMax(p.thr_bits + q, kb);
                                thr = fmaxf(thr, thr_to_float(kb));
                            }
                        }
                    } else if (p.status && lane == 0) {
                        // a skewed sub-range marked more docs than the accumulator holds: the caller re-runs the query
                        atomicOr(p.status + q, ORAG_STATUS_OVERFLOW | ORAG_STATUS_WHERE_TILE);
                    }
#undef MS_VIEW
                    // ---- X: only when something can clear the threshold, claim by RANK (lane-balanced, conflict-free):
                    // read and reset acc[r]; the rare score that clears the threshold needs its doc id -- the word whose
                    // rank range holds r (binary search in the prefix counts), then the (r - pre[word])-th set bit
                    if (reachable && mx >= thr) {
                        for (int r = lane; r < marked; r += 32) {
                            const float v = acc[r];
                            acc[r] = 0.f;
                            if (v >= thr) {
                                int lo = 0, hi = words - 1;  // last word with pre[word] <= r
                                while (lo < hi) {
                                    const int mid = (lo + hi + 1) >> 1;
                                    if ((int)pre[mid] <= r) lo = mid; else hi = mid - 1;
                                }
                                const int b = (int)__fns(bm[lo], 0u, r - (int)pre[lo] + 1);
                                ms_emit(p, q, (int32_t)(base_doc + lo * 32 + b), v);
                            }
                        }
                    } else if (fits) {
                        for (int r = lane; r < marked; r += 32) acc[r] = 0.f;
                    }
                    __syncwarp();
                    if (wpl4) reinterpret_cast<uint4 *>(bm)[lane] = make_uint4(0u, 0u, 0u, 0u);
                    else
                        for (int i = lane; i < words; i += 32) bm[i] = 0u;
                    __syncwarp();
                }
            }
            cur = nxt;
            nxt = stage_run(descA);
            descA = stage_desc(qi + 3);
        }
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
    int where = 0;
    if (n > (uint32_t)p.cap) {
        overflow = true;
        where |= ORAG_STATUS_WHERE_CANDIDATES;
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
        where |= ORAG_STATUS_WHERE_SURVIVORS;
        ns = kMsSurvCap;
    }
    if (overflow && status && threadIdx.x == 0) status[q] |= ORAG_STATUS_OVERFLOW | where;
    const int32_t *terms = p.q_terms + (int64_t)q * p.max_terms;
    const int nt = min(p.q_lens[q], p.max_terms);
    // exact re-score: one thread per (survivor, query token) runs the binary search, then one thread per survivor
    // adds the contributions in query order (the reference's order)
    __shared__ double contrib[kMsContrib];
    if (nt > 0 && (uint64_t)ns * (uint64_t)nt <= (uint64_t)kMsContrib) {
        for (uint32_t t = threadIdx.x; t < ns * (uint32_t)nt; t += blockDim.x)
            contrib[t] = term_contribution(p.ix, terms[t % nt], sd[t / nt]);
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < ns; i += blockDim.x) {
            double acc = 0.0;
            for (int j = 0; j < nt; ++j) acc = __dadd_rn(acc, contrib[i * nt + j]);
            ss[i] = acc;
        }
    } else {
        for (uint32_t i = threadIdx.x; i < ns; i += blockDim.x) ss[i] = score_doc(p.ix, terms, nt, sd[i]);
    }
    __syncthreads();
    // (an overflowed query is re-run by the caller: skip the serial zero-score fill for it)
    select_from_list(p.ix, terms, nt, k, sd, ss, ns, doc_id_base, normalize, out_ids + (int64_t)q * k,
                     out_scores + (int64_t)q * k, out_max ? out_max + q : nullptr, scratch, dscratch, !overflow);
}

__global__ void ms_init_state_kernel(unsigned long long *thr_bits, uint32_t *cnt, uint32_t *hist, uint32_t *topbin,
                                     uint32_t *work, int n_queries)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t total = (int64_t)n_queries * kHistBins;
    if (i < total) hist[i] = 0;
    if (i < n_queries) {
        thr_bits[i] = 1ull;  // smallest positive double: "score > 0"
        cnt[i] = 0;
        topbin[i] = 0;
    }
    if (i == 0) work[0] = 0;
}

static int ms_cand_cap(int n_queries)
{
    // ~128 MiB of candidate storage shared by the batch, at least 8192 slots per query
    int64_t cap = ((int64_t)128 << 20) / 8 / (n_queries > 0 ? n_queries : 1);
    if (cap < 8192) cap = 8192;
    if (cap > (1 << 22)) cap = 1 << 22;
    return (int)cap;
}

bool ms_eligible(const orag_bm25_index_t *ix, int max_terms, int flags)
{
    // (reserved & 1): the first-pass runs are 16-byte aligned and padded to four postings (bm25_build.cu layout)
    return ix->d_postings_r16 != nullptr && ix->d_term_max_r != nullptr && ix->d_fp_tile_base != nullptr &&
           ix->d_fp_tile_term_off != nullptr && (ix->reserved & 1) && !ix->has_negative_idf && max_terms <= kMsTerms &&
           !(flags & (ORAG_BM25_EXACT_TILES | ORAG_BM25_FORCE_DENSE)) && ix->fp_tile_docs >= 32 &&
           ix->fp_tile_docs <= kMsMaxTile && (ix->fp_tile_docs & (ix->fp_tile_docs - 1)) == 0 &&
           (reinterpret_cast<uintptr_t>(ix->d_postings_r16) & 15) == 0;
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
    const int cap = ms_cand_cap(n_queries);
    c.p.cap = cap;
    c.p.thr_bits = (unsigned long long *)take(nq * 8);
    c.p.cnt = (uint32_t *)take(nq * 4);
    c.p.topbin = (uint32_t *)take(nq * 4);
    c.p.hist = (uint32_t *)take(nq * kHistBins * 4);
    c.p.qd = (int32_t *)take(nq * kMsTerms * 16);
    c.p.work = (uint32_t *)take(256);
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
            const int32_t *d_query_lens, int n_queries, int max_terms, int k, int normalize, bool background,
            int64_t *d_out_ids, double *d_out_scores, double *d_out_max, int32_t *d_out_status, void *d_workspace,
            cudaStream_t st)
{
    MsParams p = ms_carve(d_workspace, n_queries).p;
    p.ix = *ix;
    p.q_terms = d_query_terms;
    p.q_lens = d_query_lens;
    p.n_queries = n_queries;
    p.max_terms = max_terms;
    p.k = k;
    p.status = d_out_status;
    {
        TimelineScope tl(TL_BM25_PREPARE, st);
        int64_t total = (int64_t)n_queries * kHistBins;
        ms_init_state_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(p.thr_bits, p.cnt, p.hist, p.topbin,
                                                                              p.work, n_queries);
        ORAG_LAUNCH_CHECK();
        prepare_queries_kernel<<<(n_queries + 127) / 128, 128, 0, st>>>(p);
        ORAG_LAUNCH_CHECK();
    }
    if (ix->fp_n_tiles > 0) {
        // Stand-alone: 2 CTAs x 16 warps per SM (64 registers per thread, 7 KiB of shared memory per warp).  Background
        // (ORAG_BM25_BACKGROUND): CTAs of 8 warps and < 31 KB of shared memory, so that ONE of them fits next to a
        // resident CTA of the cosine scan (199.9 KB, 384 threads, 96 registers) and rides along on the SM resources the
        // tensor-bound scan leaves idle.
        const int warps = background ? kMsBgWarps : kMsWarps;
        p.warp_bytes = background ? kMsBgWarpBytes : kMsWarpBytes;
        const size_t smem = (size_t)warps * p.warp_bytes;
        const int lim = sm_count() * 2;
        const int fixed_t = background ? 0 : ix->fp_tile_docs;
        auto kernel = fixed_t == 8192 ? bm25_ms_kernel<8192> : fixed_t == 4096 ? bm25_ms_kernel<4096> : bm25_ms_kernel<0>;
        ORAG_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (background) {
            // the query preparation above may run whenever the SMs have room; the first pass itself is ordered after the
            // point right before the latest cosine main scan, so that the scan's CTAs take the SMs first
            int rc = wait_prescan(st);
            if (rc) return rc;
        }
        profile_mark(1, 0, st);
        {
            TimelineScope tl(TL_BM25_FIRST_PASS, st);
            const int tiles = ix->fp_n_tiles;
            // ~8 work items per resident warp so that the atomic hand-out can balance uneven pairs
            int64_t want = (int64_t)8 * lim * warps;
            int64_t split = (want + tiles - 1) / tiles;
            if (split > n_queries) split = n_queries;
            if (split < 1) split = 1;
            p.q_split = (int)split;
            p.n_items = (int)((int64_t)tiles * split);
            int grid = (p.n_items + warps - 1) / warps;
            if (grid > lim) grid = lim;
            kernel<<<grid, warps * 32, smem, st>>>(p);
            ORAG_LAUNCH_CHECK();
        }
        profile_mark(1, 1, st);
    }
    TimelineScope tl_fin(TL_BM25_FINALIZE, st);
    ms_finalize_kernel<<<n_queries, 1024, 0, st>>>(p, doc_id_base, normalize, d_out_ids, d_out_scores, d_out_max,
                                                   d_out_status);
    ORAG_LAUNCH_CHECK();
    return ORAG_OK;
}

}  // namespace bm25
}  // namespace orag
