// bm25.cu -- bit-exact BM25 (rank_bm25.BM25Okapi.get_scores + rag/retrieval.py:324-347 glue)
// over a GPU-resident, doc-range-tiled inverted index.
//
// Index layout (built by optimized_rag_b200/bm25_index.py): documents are cut into tiles of
// `tile_docs` consecutive docs; each tile owns a contiguous run of 4-byte postings
// ((doc_in_tile << 16) | tf) grouped by term, ascending doc inside a term, plus a row of
// vocab+1 term offsets.  A tile is the unit of work: one WARP keeps a float64 accumulator and the
// per-document length (u16; t4 = k1*(1 - b + b*dl/avgdl) comes from a small dl-indexed table) in shared memory and
// streams, for every query of the batch, the posting runs of the query's terms with coalesced
// 4-byte reads.
//
// Exactness: the reference adds, per query token IN QUERY ORDER,
//     idf * (tf*(k1+1) / (tf + t4[d]))          (numpy float64, this operation order)
// to score[d].  Here one term is processed at a time (each doc occurs at most once per term, so
// there are no intra-term conflicts) with a CTA barrier between terms, using __dmul_rn /
// __ddiv_rn / __dadd_rn only -- no FMA contraction, no atomics on scores -- so every score is
// bit-identical to the oracle.  Selection: touched docs are drained with an atomic exchange
// (each doc is reported exactly once with its final score, and the accumulator is reset);
// docs whose score clears the query's running threshold go to its candidate list; the
// threshold is tightened from a log-scale histogram of emitted scores (a lower bound of the
// k-th best score so far, so no true top-k doc is ever dropped).
#include "common.cuh"
#include "select.cuh"
#include "exact.cuh"
#include "bm25_shared.cuh"

namespace orag {
namespace bm25 {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int kMaxTerms = 64;
struct Params {
    orag_bm25_index_t ix;
    const int32_t *q_terms;  // [n_queries, max_terms]
    const int32_t *q_lens;
    int n_queries;
    int max_terms;
    int k;
    // sparse path
    unsigned long long *thr_bits;  // [n_queries] bits of the (positive) threshold score
    uint32_t *cnt;
    uint32_t *hist;                // [n_queries, kHistBins]
    uint32_t *topbin;              // [n_queries] highest occupied histogram bin
    int32_t *cand_doc;             // [n_queries, cap] local doc id
    double *cand_score;            // [n_queries, cap]
    int cap;
    // dense path
    double *dense_out;             // [n_queries, n_docs] pre-zeroed
    int q_begin;                   // dense chunking: queries [q_begin, q_begin + n_queries)
    int q_split;                   // work items per tile (slices of the query batch), >= 1
};

// Candidate emission + threshold tightening.  Every 8th emission of a query re-derives its
// threshold from the log-scale histogram, scanning down from the highest occupied bin in batches
// of 16 bins fetched with four independent 16-byte loads.
__device__ __noinline__ void emit(const Params &p, int q, int32_t doc, double v)
{
    const uint32_t slot = atomicAdd(p.cnt + q, 1u);
    if (slot < (uint32_t)p.cap) {
        p.cand_doc[(int64_t)q * p.cap + slot] = doc;
        p.cand_score[(int64_t)q * p.cap + slot] = v;
    }
    uint32_t *h = p.hist + (int64_t)q * kHistBins;
    const int bin = score_bin(v);
    atomicAdd(h + bin, 1u);
    const int old_top = (int)atomicMax(p.topbin + q, (uint32_t)bin);
    if ((slot & 7u) != 7u) return;
    tighten_threshold(h, old_top > bin ? old_top : bin, p.k, p.thr_bits + q);
}

constexpr int kRTf = 4;      // r = tf*(k1+1)/(tf + t4[dl]) is tabulated for tf = 1..kRTf
constexpr int kRCacheDl = 384; // document lengths whose r-table rows are cached in shared memory

// contribution of one posting, in the reference's operation order (see file header)
__device__ __forceinline__ double contribution(const Params &p, const double *r_s, int rc, uint32_t post, uint32_t dl,
                                               double idf)
{
    const uint32_t tf = post & 0xFFFFu;
    double r;
    if (tf <= (uint32_t)kRTf) {
        // same IEEE ops as the slow branch, evaluated once at index build (bm25_index.py r_table)
        r = (int)dl < rc ? r_s[dl * kRTf + (tf - 1)] : __ldg(p.ix.d_r_table + dl * kRTf + (tf - 1));
    } else {
        const double f = (double)tf;
        r = __ddiv_rn(__dmul_rn(f, 2.5), __dadd_rn(f, __ldg(p.ix.d_t4_table + dl)));
    }
    return __dmul_rn(idf, r);
}

struct LaneRun {  // lane i holds the run of query term i inside the warp's tile
    int start;
    int len;
    double idf;
};

constexpr int kStage = 256;  // postings of one query staged per buffer (two buffers per warp)
constexpr uint32_t kNoPost = 0xFFFFFFFFu;  // doc_in_tile <= 2047, so no real posting has this value

__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gmem_src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
                 "l"(gmem_src)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

struct Staged {   // one query's runs (lane i = term i) and where their staged parts sit in the buffer
    LaneRun run;
    int soff;     // offset of the run's staged part in the staging buffer
    int slen;     // staged length (<= run.len); the rest is read on demand
    int q;        // query index
    int nt_all;   // query length
};

// Every warp owns one tile (a doc range of tile_docs docs, accumulators in its slice of shared
// memory) and walks the whole query batch on its own: no CTA barrier anywhere in the loop.
//   lanes <-> query terms   while fetching run descriptors (term id -> idf, run offsets)
//   lanes <-> postings      while staging / accumulating / draining a run (coalesced 4-byte reads)
// Four-stage software pipeline per warp:
//   iteration i issues   term ids of query i+3 (registers), run descriptors of query i+2
//                        (registers), cp.async of the posting runs of query i+1 (shared memory)
//   and consumes         the staged postings of query i
// so every load issued in an iteration is independent of everything the iteration consumes.
// Terms are applied in query order (ascending lane), run by run with __syncwarp in between, so
// per-document sums follow the reference's order.  The staged postings double as the drain list.
template <bool kDense>
__global__ void __launch_bounds__(kThreads, 2) bm25_tile_kernel(const __grid_constant__ Params p)
{
    extern __shared__ double smem_d[];
    const int T = p.ix.tile_docs;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const size_t per_warp = (size_t)T * 10 + 2 * kStage * 4;  // bytes (multiple of 8)
    // CTA-shared cache of the first rows of the r table (read-only after this point)
    double *r_s = smem_d;
    const int rc = min(p.ix.max_doc_len + 1, kRCacheDl);
    for (int i = threadIdx.x; i < rc * kRTf; i += kThreads) r_s[i] = p.ix.d_r_table[i];
    __syncthreads();
    uint8_t *mine = reinterpret_cast<uint8_t *>(smem_d + kRCacheDl * kRTf) + wib * per_warp;
    double *acc = reinterpret_cast<double *>(mine);                         // [T]
    uint32_t *stage = reinterpret_cast<uint32_t *>(mine + (size_t)T * 8);   // [2][kStage]
    uint16_t *dls = reinterpret_cast<uint16_t *>(stage + 2 * kStage);       // [T] document lengths

    const int V1 = p.ix.vocab + 1;
    const int nq = p.n_queries;
    const int mt = p.max_terms;
    const int warps_total = gridDim.x * kWarps;
    const unsigned FULL = 0xffffffffu;

    // work item = (tile, slice of the query batch): q_split > 1 keeps all warps busy when a shard has
    // fewer tiles than the GPU has resident warps (multi-GPU sharding, small corpora)
    const int S = p.q_split;
    for (int item = blockIdx.x * kWarps + wib; item < p.ix.n_tiles * S; item += warps_total) {
        const int tile = item / S;
        const int part = item - tile * S;
        const int qi0 = (int)(((int64_t)nq * part) / S);
        const int qi1 = (int)(((int64_t)nq * (part + 1)) / S);
        const int64_t base_doc = (int64_t)tile * T;
        const int nd = (int)min((int64_t)T, p.ix.n_docs - base_doc);
        for (int i = lane; i < T; i += 32) {
            acc[i] = 0.0;
            dls[i] = i < nd ? (uint16_t)p.ix.d_doc_len[base_doc + i] : (uint16_t)0;
        }
        const uint32_t *tile_post = p.ix.d_postings + p.ix.d_tile_base[tile];
        const int32_t *toff = p.ix.d_tile_term_off + (int64_t)tile * V1;
        // stagger the query order across tiles so a query's threshold is established by few warps
        const int q_shift = (int)(((int64_t)tile * 7919) % nq);
        __syncwarp();

        // ---- stage A: raw term id + query length (two independent loads)
        auto stage_terms = [&](int qi, int &t, int &nt) {
            t = -1; nt = 0;
            if (qi < qi1) {
                int q = qi + q_shift;  // q_shift < nq and qi < nq: one conditional subtract, no modulo
                if (q >= nq) q -= nq;
                nt = __ldg(p.q_lens + q);
                if (lane < mt) t = __ldg(p.q_terms + (int64_t)q * mt + lane);
            }
        };
        // ---- stage B: run descriptor of (validated) term t; len holds the END offset until staged
        auto stage_run = [&](int t, int nt) -> LaneRun {
            LaneRun r;
            r.start = 0; r.len = 0; r.idf = 0.0;
            if (lane < min(nt, mt) && t >= 0 && t < p.ix.vocab) {
                r.idf = __ldg(p.ix.d_idf + t);
                r.start = __ldg(toff + t);
                r.len = __ldg(toff + t + 1);
            }
            return r;
        };
        // ---- stage C: lay the runs out in the staging buffer and issue their copies
        auto stage_posts = [&](LaneRun run, int qi, int nt) -> Staged {
            Staged s;
            run.len = (run.idf != 0.0) ? run.len - run.start : 0;  // `idf.get(q) or 0`: zero idf adds nothing
            s.run = run;
            s.q = (qi < qi1) ? (qi + q_shift >= nq ? qi + q_shift - nq : qi + q_shift) : 0;
            s.nt_all = min(nt, mt);
            int incl = run.len;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int n = __shfl_up_sync(FULL, incl, d);
                if (lane >= d) incl += n;
            }
            s.soff = min(incl - run.len, kStage);
            s.slen = min(run.len, kStage - s.soff);
            uint32_t *buf = stage + (qi & 1) * kStage;
            unsigned active = __ballot_sync(FULL, s.slen > 0);
            while (active) {
                const int i = __ffs(active) - 1;
                active &= active - 1;
                const int st = __shfl_sync(FULL, run.start, i);
                const int so = __shfl_sync(FULL, s.soff, i);
                const int sl = __shfl_sync(FULL, s.slen, i);
                for (int j = lane; j < sl; j += 32) cp_async4(buf + so + j, tile_post + st + j);
            }
            cp_async_commit();
            return s;
        };

        // prologue: fill the pipeline
        int tA, ntA;
        stage_terms(qi0, tA, ntA);
        Staged cur = stage_posts(stage_run(tA, ntA), qi0, ntA);  // first query: postings in flight
        stage_terms(qi0 + 1, tA, ntA);
        LaneRun runB = stage_run(tA, ntA);                       // second query: descriptors in flight
        int ntRunB = ntA;
        stage_terms(qi0 + 2, tA, ntA);                           // third query: term ids in flight

        for (int qi = qi0; qi < qi1; ++qi) {
            // ---- issue: postings of query qi+1, descriptors of query qi+2, terms of query qi+3
            const Staged nxt = stage_posts(runB, qi + 1, ntRunB);
            runB = stage_run(tA, ntA);
            ntRunB = ntA;
            stage_terms(qi + 3, tA, ntA);

            const int q = cur.q;
            double thr = 0.0;
            if (!kDense) thr = __longlong_as_double((long long)__ldcg(p.thr_bits + q));
            const uint32_t *buf = stage + (qi & 1) * kStage;
            cp_async_wait_1();  // everything but the copies just issued has landed (this lane's part)
            __syncwarp();       // ... and every lane's part

            // ---- accumulate query qi, run by run in query order
            auto apply = [&](uint32_t post, double idf) {
                const uint32_t d = post >> 16;
                acc[d] = __dadd_rn(acc[d], contribution(p, r_s, rc, post, dls[d], idf));
            };
            auto drain_doc = [&](uint32_t d) {
                const double v = acc[d];
                if (v == 0.0) return;  // already drained through another run of this query
                if (kDense) {
                    acc[d] = 0.0;      // idempotent: racing lanes store the same values
                    p.dense_out[(int64_t)q * p.ix.n_docs + base_doc + d] = v;
                } else if (v < thr) {
                    acc[d] = 0.0;
                } else {
                    const unsigned long long bits = atomicExch((unsigned long long *)(acc + d), 0ull);
                    const double w = __longlong_as_double((long long)bits);
                    if (w != 0.0) emit(p, q, (int32_t)(base_doc + d), w);
                }
            };
            // walks the runs of the first 32 terms; mode 0 = accumulate, 1 = drain the non-staged parts
            auto walk = [&](const LaneRun &run, int soff, int slen, bool staged_ok, int mode) {
                unsigned active = __ballot_sync(FULL, run.len > 0);
                while (active) {
                    const int i = __ffs(active) - 1;
                    active &= active - 1;
                    const int st = __shfl_sync(FULL, run.start, i);
                    const int ln = __shfl_sync(FULL, run.len, i);
                    const int so = __shfl_sync(FULL, soff, i);
                    const int sl = staged_ok ? __shfl_sync(FULL, slen, i) : 0;
                    if (mode == 0) {
                        const double idf = __shfl_sync(FULL, run.idf, i);
                        for (int j = lane; j < sl; j += 32) apply(buf[so + j], idf);
                        // non-staged tail of a long run: four independent loads in flight per lane
                        for (int j = sl + lane; j < ln; j += 128) {
                            uint32_t pv[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u)
                                pv[u] = (j + 32 * u < ln) ? __ldg(tile_post + st + j + 32 * u) : kNoPost;
#pragma unroll
                            for (int u = 0; u < 4; ++u)
                                if (pv[u] != kNoPost) apply(pv[u], idf);
                        }
                        __syncwarp();  // run i fully applied before run i+1
                    } else {
                        for (int j = sl + lane; j < ln; j += 32) drain_doc(__ldg(tile_post + st + j) >> 16);
                    }
                }
            };
            auto long_query_runs = [&](int c0) -> LaneRun {  // terms 32.. of very long queries (rare)
                LaneRun run;
                run.start = 0; run.len = 0; run.idf = 0.0;
                if (c0 + lane < cur.nt_all) {
                    const int t = __ldg(p.q_terms + (int64_t)q * mt + c0 + lane);
                    if (t >= 0 && t < p.ix.vocab) {
                        run.idf = __ldg(p.ix.d_idf + t);
                        run.start = __ldg(toff + t);
                        run.len = (run.idf != 0.0) ? __ldg(toff + t + 1) - run.start : 0;
                    }
                }
                return run;
            };
            const int staged_total = __shfl_sync(FULL, cur.soff + cur.slen, 31);
            const bool any = __any_sync(FULL, cur.run.len > 0) || cur.nt_all > 32;
            if (any) {
                walk(cur.run, cur.soff, cur.slen, true, 0);
                for (int c0 = 32; c0 < cur.nt_all; c0 += 32) walk(long_query_runs(c0), 0, 0, false, 0);

                // ---- drain: every touched doc is reported once with its final score; accumulator reset
                const bool long_runs = __any_sync(FULL, cur.run.len > cur.slen) || cur.nt_all > 32;
                if (long_runs) {
                    // some postings were not staged: sweep the whole tile instead of re-reading them
                    for (int d = lane; d < T; d += 32) drain_doc((uint32_t)d);
                } else {
#pragma unroll 4
                    for (int j = lane; j < staged_total; j += 32) drain_doc(buf[j] >> 16);
                }
                __syncwarp();
            }
            cur = nxt;
        }
    }
}

// One CTA per query: exact top-k over the candidate list by (score/max desc, id asc), then
// zero-score fill (docs untouched by the query rank after all positive ones, in id order).
__global__ void __launch_bounds__(1024) finalize_kernel(const __grid_constant__ Params p, int64_t doc_id_base, int normalize,
                                                      int64_t *__restrict__ out_ids, double *__restrict__ out_scores,
                                                      double *__restrict__ out_max, int32_t *__restrict__ status)
{
    __shared__ Pick scratch[32];
    __shared__ double dscratch[32];
    const int q = blockIdx.x;
    const int k = p.k;
    uint32_t n = p.cnt[q];
    const bool overflow = n > (uint32_t)p.cap;
    if (overflow) {
        if (status && threadIdx.x == 0) status[q] |= ORAG_STATUS_OVERFLOW;
        n = p.cap;
    }
    select_from_list(p.ix, p.q_terms + (int64_t)q * p.max_terms, min(p.q_lens[q], p.max_terms), k,
                     p.cand_doc + (int64_t)q * p.cap, p.cand_score + (int64_t)q * p.cap, n, doc_id_base, normalize,
                     out_ids + (int64_t)q * k, out_scores + (int64_t)q * k, out_max ? out_max + q : nullptr, scratch,
                     dscratch, !overflow);
}

__global__ void init_state_kernel(unsigned long long *thr_bits, uint32_t *cnt, uint32_t *hist, uint32_t *topbin,
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

}  // namespace bm25

}  // namespace orag

using orag::bm25::Params;

static int validate_index(const orag_bm25_index_t *ix)
{
    ORAG_REQUIRE(ix != nullptr, "index");
    ORAG_REQUIRE(ix->n_docs >= 0 && ix->vocab > 0, "index sizes");
    ORAG_REQUIRE(ix->tile_docs >= 32 && ix->tile_docs <= 65536 && (ix->tile_docs & (ix->tile_docs - 1)) == 0,
                 "tile_docs must be a power of two in [32, 65536]");
    ORAG_REQUIRE((int64_t)ix->n_tiles == (ix->n_docs + ix->tile_docs - 1) / ix->tile_docs, "n_tiles");
    ORAG_REQUIRE(ix->tile_docs <= 2048, "tile_docs must be <= 2048 (per-warp accumulators in shared memory)");
    ORAG_REQUIRE(ix->n_docs == 0 || ix->d_r_table, "r table");
    if (ix->n_docs > 0)
        ORAG_REQUIRE(ix->d_tile_base && ix->d_tile_term_off && ix->d_postings && ix->d_doc_len && ix->d_t4_table &&
                         ix->d_idf,
                     "index pointers");
    ORAG_REQUIRE(ix->max_doc_len >= 0 && ix->max_doc_len <= 65535, "document length above 65535 tokens");
    return ORAG_OK;
}

static size_t dense_chunk_queries(const orag_bm25_index_t *ix, int n_queries)
{
    const size_t budget = (size_t)256 << 20;
    size_t per_q = (size_t)(ix->n_docs > 0 ? ix->n_docs : 1) * sizeof(double);
    size_t c = budget / per_q;
    if (c < 1) c = 1;
    if (c > (size_t)n_queries) c = (size_t)n_queries;
    return c;
}

static bool use_dense(const orag_bm25_index_t *ix, int n_queries, int flags)
{
    if (flags & ORAG_BM25_FORCE_DENSE) return true;
    if (ix->has_negative_idf) return true;
    if (flags & ORAG_BM25_FORCE_SPARSE) return false;
    return (size_t)n_queries * (size_t)ix->n_docs * sizeof(double) <= ((size_t)64 << 20);
}

static int sparse_cap(int n_queries)
{
    // ~256 MiB of candidate storage shared by the batch, at least 8192 slots per query
    int64_t cap = ((int64_t)256 << 20) / 12 / (n_queries > 0 ? n_queries : 1);
    if (cap < 8192) cap = 8192;
    if (cap > (1 << 22)) cap = 1 << 22;
    return (int)cap;
}

extern "C" size_t orag_bm25_workspace_bytes(const orag_bm25_index_t *ix, int n_queries, int k, int flags)
{
    if (!ix || n_queries <= 0) return 0;
    (void)k;
    if (use_dense(ix, n_queries, flags))
        return orag::align_up(dense_chunk_queries(ix, n_queries) * (size_t)(ix->n_docs > 0 ? ix->n_docs : 1) * 8, 256);
    const size_t cap = sparse_cap(n_queries);
    const size_t ms = (ix->d_postings_r16 && !(flags & ORAG_BM25_EXACT_TILES))
                          ? orag::bm25::ms_workspace_bytes(ix, n_queries) : 0;
    size_t b = 0;
    b += orag::align_up((size_t)n_queries * 8, 256);                        // thr_bits
    b += orag::align_up((size_t)n_queries * 4, 256);                        // cnt
    b += orag::align_up((size_t)n_queries * 4, 256);                        // topbin
    b += orag::align_up((size_t)n_queries * orag::bm25::kHistBins * 4, 256);  // hist
    b += orag::align_up((size_t)n_queries * cap * 4, 256);                  // cand_doc
    b += orag::align_up((size_t)n_queries * cap * 8, 256);                  // cand_score
    return b > ms ? b : ms;
}

// Queries longer than kMaxTerms tokens (the tile kernels keep one term per lane slot): one thread per (query, doc)
// walks the query's tokens in order and looks each up in the doc's tile (score_doc) -- the reference's arithmetic and
// summation order for any query length, at O(tokens * log(run)) per document.
__global__ void __launch_bounds__(256) dense_by_doc_kernel(orag_bm25_index_t ix, const int32_t *__restrict__ q_terms,
                                                          const int32_t *__restrict__ q_lens, int n_queries,
                                                          int max_terms, double *__restrict__ out)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= (int64_t)n_queries * ix.n_docs) return;
    const int q = (int)(i / ix.n_docs);
    const int64_t d = i - (int64_t)q * ix.n_docs;
    const int nt = min(q_lens[q], max_terms);
    out[i] = orag::bm25::score_doc(ix, q_terms + (int64_t)q * max_terms, nt, d);
}

static int launch_by_doc(const orag_bm25_index_t *ix, const int32_t *q_terms, const int32_t *q_lens, int n_queries,
                         int max_terms, double *out, cudaStream_t st)
{
    const int64_t total = (int64_t)n_queries * ix->n_docs;
    if (total == 0) return ORAG_OK;
    dense_by_doc_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(*ix, q_terms, q_lens, n_queries, max_terms, out);
    ORAG_LAUNCH_CHECK();
    return ORAG_OK;
}

static int launch_tiles(const Params &p, bool dense, cudaStream_t st)
{
    const size_t smem = (size_t)orag::bm25::kWarps * ((size_t)p.ix.tile_docs * 10 + 2 * orag::bm25::kStage * 4) +
                        (size_t)orag::bm25::kRCacheDl * orag::bm25::kRTf * 8;
    int per_sm = (int)((224 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
    const int lim = orag::sm_count() * per_sm;
    // split the query batch so that there are ~4 work items per resident warp even for small shards
    Params pp = p;
    int64_t want = (int64_t)4 * lim * orag::bm25::kWarps;
    int split = (int)((want + p.ix.n_tiles - 1) / (p.ix.n_tiles > 0 ? p.ix.n_tiles : 1));
    if (split > 16) split = 16;
    if (split > p.n_queries) split = p.n_queries;
    if (split < 1) split = 1;
    pp.q_split = split;
    int64_t items = (int64_t)p.ix.n_tiles * split;
    int grid = (int)((items + orag::bm25::kWarps - 1) / orag::bm25::kWarps);
    if (grid > lim) grid = lim;
    if (grid < 1) return ORAG_OK;
    orag::profile_mark(1, 0, st);
    if (dense) {
        ORAG_CUDA_CHECK(cudaFuncSetAttribute(orag::bm25::bm25_tile_kernel<true>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        orag::bm25::bm25_tile_kernel<true><<<grid, orag::bm25::kThreads, smem, st>>>(pp);
    } else {
        ORAG_CUDA_CHECK(cudaFuncSetAttribute(orag::bm25::bm25_tile_kernel<false>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        orag::bm25::bm25_tile_kernel<false><<<grid, orag::bm25::kThreads, smem, st>>>(pp);
    }
    orag::profile_mark(1, 1, st);
    ORAG_LAUNCH_CHECK();
    return ORAG_OK;
}

extern "C" int orag_bm25_dense(const orag_bm25_index_t *ix, const int32_t *d_query_terms, const int32_t *d_query_lens,
                               int n_queries, int max_terms, double *d_out, void *stream)
{
    int rc = validate_index(ix);
    if (rc) return rc;
    ORAG_REQUIRE(d_query_terms && d_query_lens && d_out && n_queries > 0 && max_terms > 0, "bm25_dense");
    cudaStream_t st = (cudaStream_t)stream;
    if (ix->n_docs == 0) return ORAG_OK;
    if (max_terms > orag::bm25::kMaxTerms)  // long queries: per-document evaluation, any length
        return launch_by_doc(ix, d_query_terms, d_query_lens, n_queries, max_terms, d_out, st);
    ORAG_CUDA_CHECK(cudaMemsetAsync(d_out, 0, (size_t)n_queries * ix->n_docs * sizeof(double), st));
    Params p{};
    p.ix = *ix;
    p.q_terms = d_query_terms;
    p.q_lens = d_query_lens;
    p.n_queries = n_queries;
    p.max_terms = max_terms;
    p.dense_out = d_out;
    p.q_begin = 0;
    return launch_tiles(p, true, st);
}

extern "C" int orag_bm25_topk(const orag_bm25_index_t *ix, int64_t doc_id_base, const int32_t *d_query_terms,
                              const int32_t *d_query_lens, int n_queries, int max_terms, int k, int flags,
                              int64_t *d_out_ids, double *d_out_scores, double *d_out_max, int32_t *d_out_status,
                              void *d_workspace, size_t workspace_bytes, void *stream)
{
    int rc = validate_index(ix);
    if (rc) return rc;
    ORAG_REQUIRE(d_query_terms && d_query_lens && d_out_ids && d_out_scores && n_queries > 0 && k > 0 && max_terms > 0,
                 "bm25_topk");
    ORAG_REQUIRE(max_terms <= orag::bm25::kMaxTerms || (flags & ORAG_BM25_FORCE_DENSE),
                 "queries longer than 64 tokens take the dense path: pass ORAG_BM25_FORCE_DENSE");
    cudaStream_t st = (cudaStream_t)stream;
    const int normalize = (flags & ORAG_BM25_NORMALIZE) ? 1 : 0;
    if (workspace_bytes < orag_bm25_workspace_bytes(ix, n_queries, k, flags) || (!d_workspace && ix->n_docs > 0)) {
        orag::set_error("bm25_topk: workspace too small (%zu < %zu)", workspace_bytes,
                        orag_bm25_workspace_bytes(ix, n_queries, k, flags));
        return ORAG_EWORKSPACE;
    }
    if (d_out_status) ORAG_CUDA_CHECK(cudaMemsetAsync(d_out_status, 0, (size_t)n_queries * sizeof(int32_t), st));

    if (use_dense(ix, n_queries, flags)) {
        const int chunk = (int)dense_chunk_queries(ix, n_queries);
        double *dense = (double *)d_workspace;
        for (int q0 = 0; q0 < n_queries; q0 += chunk) {
            const int nq = (n_queries - q0) < chunk ? (n_queries - q0) : chunk;
            if (ix->n_docs > 0 && max_terms > orag::bm25::kMaxTerms) {
                rc = launch_by_doc(ix, d_query_terms + (int64_t)q0 * max_terms, d_query_lens + q0, nq, max_terms, dense, st);
                if (rc) return rc;
            } else if (ix->n_docs > 0) {
                ORAG_CUDA_CHECK(cudaMemsetAsync(dense, 0, (size_t)nq * ix->n_docs * sizeof(double), st));
                Params p{};
                p.ix = *ix;
                p.q_terms = d_query_terms + (int64_t)q0 * max_terms;
                p.q_lens = d_query_lens + q0;
                p.n_queries = nq;
                p.max_terms = max_terms;
                p.dense_out = dense;
                p.q_begin = 0;
                rc = launch_tiles(p, true, st);
                if (rc) return rc;
            }
            rc = orag::launch_select_topk(dense, nullptr, nullptr, ix->n_docs, ix->n_docs, nq, k, doc_id_base,
                                          normalize ? 1 : 2, nullptr, 0, d_out_ids + (int64_t)q0 * k,
                                          d_out_scores + (int64_t)q0 * k, d_out_max ? d_out_max + q0 : nullptr, nullptr,
                                          st);
            if (rc) return rc;
        }
        return ORAG_OK;
    }

    // ---- sparse (candidate) paths ----
    if (orag::bm25::ms_eligible(ix, max_terms, flags))
        return orag::bm25::ms_topk(ix, doc_id_base, d_query_terms, d_query_lens, n_queries, max_terms, k, normalize,
                                   (flags & ORAG_BM25_BACKGROUND) != 0, d_out_ids, d_out_scores, d_out_max,
                                   d_out_status, d_workspace, st);
    const int cap = sparse_cap(n_queries);
    uint8_t *w = (uint8_t *)d_workspace;
    Params p{};
    p.ix = *ix;
    p.q_terms = d_query_terms;
    p.q_lens = d_query_lens;
    p.n_queries = n_queries;
    p.max_terms = max_terms;
    p.k = k;
    p.cap = cap;
    p.thr_bits = (unsigned long long *)w; w += orag::align_up((size_t)n_queries * 8, 256);
    p.cnt = (uint32_t *)w;                w += orag::align_up((size_t)n_queries * 4, 256);
    p.topbin = (uint32_t *)w;             w += orag::align_up((size_t)n_queries * 4, 256);
    p.hist = (uint32_t *)w;               w += orag::align_up((size_t)n_queries * orag::bm25::kHistBins * 4, 256);
    p.cand_doc = (int32_t *)w;            w += orag::align_up((size_t)n_queries * cap * 4, 256);
    p.cand_score = (double *)w;
    {
        int64_t total = (int64_t)n_queries * orag::bm25::kHistBins;
        orag::bm25::init_state_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(p.thr_bits, p.cnt, p.hist,
                                                                                      p.topbin, n_queries);
        ORAG_LAUNCH_CHECK();
    }
    rc = launch_tiles(p, false, st);
    if (rc) return rc;
    orag::bm25::finalize_kernel<<<n_queries, 1024, 0, st>>>(p, doc_id_base, normalize, d_out_ids, d_out_scores, d_out_max,
                                                           d_out_status);
    ORAG_LAUNCH_CHECK();
    return ORAG_OK;
}
