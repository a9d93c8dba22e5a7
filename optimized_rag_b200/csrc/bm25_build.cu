// bm25_build.cu -- ingest side of the BM25 path: token corpus -> the tiled inverted index orag_bm25_topk consumes.
//
// Replaces what `BM25Okapi(tokenized_corpus)` derives on every call (rag/retrieval.py:334-338; rank_bm25 0.2.2
// BM25._initialize: per-document term frequencies, document lengths, document frequencies in first-seen order) and lays
// the postings out for the two kernels of the query side (include/orag.h `orag_bm25_index_t`):
//   exact view       uint32 (doc_in_tile << 16) | tf, grouped by (tile, term), ascending doc      (bm25.cu, re-score)
//   first-pass view  uint32 (doc_in_tile << 16) | fp16(tf*(k1+1)/(tf + t4[dl])), tiles of fp_tile_docs docs; every
//                    (tile, term) run starts on a 16-byte boundary and is padded to a multiple of four postings with
//                    copies of its last doc carrying impact 0 (bm25_ms.cu streams runs with 16-byte loads only)
//
// Two phases, because the sizes of the posting arrays are an output and the library never allocates:
//   orag_bm25_index_plan   counting pass: per-(tile, term) posting counts of both tilings, doc lengths, df, first-seen
//                          positions; exclusive scans -> run offsets and tile bases (final).  The caller reads the two
//                          totals, reduces the statistics over all shards (global idf / avgdl), allocates.
//   orag_bm25_index_fill   scatter pass: the same walk over the documents writes every posting to its slot.
//
// One warp owns one "super tile" (max(tile_docs, fp_tile_docs) consecutive docs) and walks its documents IN ORDER, so
// the postings of a run come out in ascending doc order without any sort: the warp de-duplicates a document's tokens in
// a shared-memory hash table (term -> tf; documents longer than 768 tokens are split by term hash into several passes),
// then every distinct (term, tf) takes the next slot of its run through a per-(tile, term) cursor.  Different warps
// touch different tiles, so cursors and counts never contend; 10M docs = 2442 super tiles = 2442 warps in flight.
#include <cuda_fp16.h>

#include "common.cuh"

namespace orag {
namespace build {

constexpr int kSlots = 1024;          // hash slots per warp
constexpr int kPassTokens = 768;      // tokens per de-duplication pass (worst case 768 distinct terms in 1024 slots)
constexpr int kWarps = 4;             // warps per CTA (10 KB of shared memory each)
constexpr uint32_t kEmpty = 0xFFFFFFFFu;

// d_info words
constexpr int kInfoMaxTf = 0, kInfoMaxDl = 1, kInfoError = 2;
constexpr int kErrToken = 1;      // a token id outside [0, vocab)
constexpr int kErrDocLen = 2;     // a document longer than 65535 tokens
constexpr int kErrHash = 4;       // de-duplication table overflow (adversarial term hashing)
constexpr int kErrTilePostings = 8;   // a tile with >= 2^31 postings
constexpr int kErrImpact = 16;    // an impact that is not a normal fp16 number: the first-pass view is unusable

struct WarpHash {
    uint32_t key[kSlots];
    uint32_t cnt[kSlots];
    uint16_t owned[kSlots];
};

__device__ __forceinline__ uint32_t hash_term(uint32_t t) { return (t * 0x9E3779B1u) >> 22; }          // 10 bits
__device__ __forceinline__ uint32_t part_term(uint32_t t) { return (t ^ (t >> 15)) * 0x85EBCA6Bu >> 8; }

// Calls emit(term, tf) once per distinct term of the document tok[0..len), from arbitrary lanes.
template <class Emit>
__device__ __forceinline__ void doc_terms(const int32_t *__restrict__ tok, int len, int vocab, WarpHash &h, int lane,
                                          int32_t *info, Emit emit)
{
    const unsigned FULL = 0xffffffffu;
    const int parts = (len + kPassTokens - 1) / kPassTokens;
    for (int part = 0; part < parts; ++part) {
        int n_owned = 0;
        for (int j0 = 0; j0 < len; j0 += 32) {
            const int j = j0 + lane;
            int32_t t = j < len ? __ldg(tok + j) : -1;
            bool mine = j < len;
            if (mine && (t < 0 || t >= vocab)) {
                atomicOr(info + kInfoError, kErrToken);
                mine = false;
            }
            if (mine && parts > 1) mine = (int)(part_term((uint32_t)t) % (uint32_t)parts) == part;
            bool fresh = false;
            uint32_t s = 0;
            if (mine) {
                s = hash_term((uint32_t)t);
                int probe = 0;
                for (; probe < kSlots; ++probe) {
                    const uint32_t old = atomicCAS(&h.key[s], kEmpty, (uint32_t)t);
                    if (old == kEmpty) { fresh = true; break; }
                    if (old == (uint32_t)t) break;
                    s = (s + 1) & (kSlots - 1);
                }
                if (probe == kSlots) atomicOr(info + kInfoError, kErrHash);
                else atomicAdd(&h.cnt[s], 1u);
            }
            const unsigned m = __ballot_sync(FULL, fresh);
            if (fresh) h.owned[n_owned + __popc(m & ((1u << lane) - 1u))] = (uint16_t)s;
            n_owned += __popc(m);
        }
        __syncwarp();
        for (int i = lane; i < n_owned; i += 32) {
            const int s = h.owned[i];
            const uint32_t term = h.key[s], tf = h.cnt[s];
            h.key[s] = kEmpty;
            h.cnt[s] = 0u;
            emit((int)term, (int)tf);
        }
        __syncwarp();
    }
}

__device__ __forceinline__ void init_hash(WarpHash &h, int lane)
{
    for (int i = lane; i < kSlots; i += 32) {
        h.key[i] = kEmpty;
        h.cnt[i] = 0u;
    }
    __syncwarp();
}

struct Geometry {
    int64_t n_docs;
    int vocab;
    int tile_docs, fp_tile_docs, super_docs;   // super_docs = max of the two (all powers of two)
    int n_tiles, fp_n_tiles, n_super;
};

// ---- phase 1: counts -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kWarps * 32) count_kernel(const int64_t *__restrict__ doc_off,
                                                           const int32_t *__restrict__ tokens, Geometry g,
                                                           int32_t *__restrict__ doc_len, int32_t *__restrict__ cnt_e,
                                                           int32_t *__restrict__ cnt_f, int32_t *__restrict__ info)
{
    __shared__ WarpHash hs[kWarps];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    WarpHash &h = hs[wib];
    init_hash(h, lane);
    const int V = g.vocab;
    int max_tf = 0, max_dl = 0;
    for (int st = blockIdx.x * kWarps + wib; st < g.n_super; st += gridDim.x * kWarps) {
        const int64_t d0 = (int64_t)st * g.super_docs;
        const int64_t d1 = min(g.n_docs, d0 + g.super_docs);
        for (int64_t d = d0; d < d1; ++d) {
            const int64_t lo = doc_off[d], hi = doc_off[d + 1];
            int len = (int)(hi - lo);
            if (hi - lo > 65535) {
                if (lane == 0) atomicOr(info + kInfoError, kErrDocLen);
                len = 65535;
            }
            if (lane == 0) doc_len[d] = len;
            max_dl = max(max_dl, len);
            int32_t *ce = cnt_e + (int64_t)(d / g.tile_docs) * V;
            int32_t *cf = cnt_f + (int64_t)(d / g.fp_tile_docs) * V;
            doc_terms(tokens + lo, len, V, h, lane, info, [&](int term, int tf) {
                atomicAdd(ce + term, 1);
                atomicAdd(cf + term, 1);
                max_tf = max(max_tf, tf);
            });
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) max_tf = max(max_tf, __shfl_xor_sync(0xffffffffu, max_tf, o));
    if (lane == 0) {
        atomicMax(info + kInfoMaxTf, max_tf);
        atomicMax(info + kInfoMaxDl, max_dl);
    }
}

// first global token position of every term (dict insertion order of rank_bm25's `nd`: ascending first position)
__global__ void first_pos_kernel(const int32_t *__restrict__ tokens, int64_t total, int vocab, int64_t pos_base,
                                 long long *__restrict__ first_pos)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += stride) {
        const int t = __ldg(tokens + i);
        if (t < 0 || t >= vocab) continue;
        const long long gp = pos_base + i;
        if (gp < __ldcg(first_pos + t)) atomicMin(first_pos + t, gp);
    }
}

// df[t] += number of this shard's docs that contain t = column sums of the exact-view counts
__global__ void df_kernel(const int32_t *__restrict__ cnt_e, int n_tiles, int vocab, long long *__restrict__ df)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= vocab) return;
    long long s = 0;
    for (int tile = 0; tile < n_tiles; ++tile) s += cnt_e[(int64_t)tile * vocab + t];
    df[t] += s;
}

// One CTA per tile: off[tile][0..V] = exclusive scan of the (optionally padded) counts; the counts are zeroed on the way
// (they serve as the fill pass's cursors); totals[tile] = postings of the tile (padded).
__global__ void __launch_bounds__(1024) scan_rows_kernel(int32_t *__restrict__ cnt, int vocab, int pad4,
                                                        int32_t *__restrict__ off, long long *__restrict__ totals,
                                                        int32_t *__restrict__ info)
{
    __shared__ long long warp_sums[32];
    __shared__ long long s_carry;
    const int tile = blockIdx.x;
    int32_t *c = cnt + (int64_t)tile * vocab;
    int32_t *o = off + (int64_t)tile * (vocab + 1);
    const int per = (vocab + blockDim.x - 1) / blockDim.x;
    const int lo = min(vocab, (int)threadIdx.x * per), hi = min(vocab, lo + per);
    long long mine = 0;
    for (int t = lo; t < hi; ++t) {
        const int v = c[t];
        mine += pad4 ? ((v + 3) & ~3) : v;
    }
    // block-wide exclusive scan of `mine`
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    long long incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const long long v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
    }
    if (lane == 31) warp_sums[w] = incl;
    __syncthreads();
    if (w == 0) {
        long long x = lane < (int)(blockDim.x >> 5) ? warp_sums[lane] : 0;
        long long xi = x;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const long long v = __shfl_up_sync(0xffffffffu, xi, d);
            if (lane >= d) xi += v;
        }
        warp_sums[lane] = xi - x;
        if (lane == 31) s_carry = xi;
    }
    __syncthreads();
    long long run = warp_sums[w] + incl - mine;
    for (int t = lo; t < hi; ++t) {
        const int v = c[t];
        o[t] = (int32_t)run;
        run += pad4 ? ((v + 3) & ~3) : v;
        c[t] = 0;
    }
    if (threadIdx.x == 0) {
        const long long total = s_carry;
        o[vocab] = (int32_t)total;
        totals[tile] = total;
        if (total >= (1ll << 31)) atomicOr(info + kInfoError, kErrTilePostings);
    }
}

// base[0..n] = exclusive scan of totals (one warp; n is a few thousand)
__global__ void scan_bases_kernel(const long long *__restrict__ totals, int n, long long *__restrict__ base)
{
    const int lane = threadIdx.x;
    long long carry = 0;
    for (int i0 = 0; i0 < n; i0 += 32) {
        const long long v = i0 + lane < n ? totals[i0 + lane] : 0;
        long long incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const long long u = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += u;
        }
        if (i0 + lane < n) base[i0 + lane] = carry + incl - v;
        carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) base[n] = carry;
}

// ---- phase 2: scatter ------------------------------------------------------------------------------------------------
struct FillArgs {
    const int64_t *doc_off;
    const int32_t *tokens;
    const double *t4_table;   // [max_doc_len + 1]
    const long long *tile_base, *fp_tile_base;
    const int32_t *tile_term_off, *fp_tile_term_off;
    uint32_t *postings, *postings_r16;   // postings_r16 may be null (no first-pass view)
    uint32_t *term_max_bits;             // fp32 bits of the per-term maximum impact
    int32_t *cur_e, *cur_f;
    int32_t *info;
    int max_doc_len;
};

__global__ void __launch_bounds__(kWarps * 32) fill_kernel(Geometry g, FillArgs a)
{
    __shared__ WarpHash hs[kWarps];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    WarpHash &h = hs[wib];
    init_hash(h, lane);
    const int V = g.vocab;
    for (int st = blockIdx.x * kWarps + wib; st < g.n_super; st += gridDim.x * kWarps) {
        const int64_t d0 = (int64_t)st * g.super_docs;
        const int64_t d1 = min(g.n_docs, d0 + g.super_docs);
        for (int64_t d = d0; d < d1; ++d) {
            const int64_t lo = a.doc_off[d];
            const int len = (int)min((int64_t)65535, a.doc_off[d + 1] - lo);
            const int te = (int)(d / g.tile_docs), tf_ = (int)(d / g.fp_tile_docs);
            const uint32_t de = (uint32_t)(d - (int64_t)te * g.tile_docs) << 16;
            const uint32_t df_ = (uint32_t)(d - (int64_t)tf_ * g.fp_tile_docs) << 16;
            uint32_t *pe = a.postings + a.tile_base[te];
            const int32_t *oe = a.tile_term_off + (int64_t)te * (V + 1);
            int32_t *ce = a.cur_e + (int64_t)te * V;
            uint32_t *pf = a.postings_r16 ? a.postings_r16 + a.fp_tile_base[tf_] : nullptr;
            const int32_t *of = a.fp_tile_term_off + (int64_t)tf_ * (V + 1);
            int32_t *cf = a.cur_f + (int64_t)tf_ * V;
            const double t4 = a.t4_table[min(len, a.max_doc_len)];
            doc_terms(a.tokens + lo, len, V, h, lane, a.info, [&](int term, int tf) {
                const int tfc = min(tf, 0xFFFF);
                pe[oe[term] + atomicAdd(ce + term, 1)] = de | (uint32_t)tfc;
                if (pf) {
                    // impact in rank_bm25's operation order, then float64 -> fp32 -> fp16, both round-to-nearest
                    const double tfd = (double)tf;
                    const double r = __ddiv_rn(__dmul_rn(tfd, 2.5), __dadd_rn(tfd, t4));
                    const float r32 = __double2float_rn(r);
                    const __half r16 = __float2half_rn(r32);
                    const unsigned short bits = __half_as_ushort(r16);
                    if ((bits & 0x7C00u) == 0u || (bits & 0x7C00u) == 0x7C00u || (bits & 0x8000u))
                        atomicOr(a.info + kInfoError, kErrImpact);   // zero / subnormal / inf / nan / negative
                    pf[of[term] + atomicAdd(cf + term, 1)] = df_ | (uint32_t)bits;
                    const uint32_t fb = __float_as_uint(__half2float(r16));
                    if (fb > __ldcg(a.term_max_bits + term)) atomicMax(a.term_max_bits + term, fb);
                }
            });
        }
    }
}

// pads every first-pass run to a multiple of four postings: copies of the run's last doc with impact +0.0
__global__ void pad_runs_kernel(Geometry g, const long long *__restrict__ fp_tile_base,
                                const int32_t *__restrict__ fp_tile_term_off, const int32_t *__restrict__ cur_f,
                                uint32_t *__restrict__ postings_r16)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= (int64_t)g.fp_n_tiles * g.vocab) return;
    const int tile = (int)(i / g.vocab), t = (int)(i - (int64_t)tile * g.vocab);
    const int len = cur_f[i];
    if (len == 0 || (len & 3) == 0) return;
    uint32_t *run = postings_r16 + fp_tile_base[tile] + fp_tile_term_off[(int64_t)tile * (g.vocab + 1) + t];
    const uint32_t filler = run[len - 1] & 0xFFFF0000u;
    for (int j = len; j < ((len + 3) & ~3); ++j) run[j] = filler;
}

static bool pow2(int x) { return x > 0 && (x & (x - 1)) == 0; }

static int geometry(int64_t n_docs, int vocab, int tile_docs, int fp_tile_docs, Geometry *g)
{
    ORAG_REQUIRE(n_docs >= 0 && n_docs < ((int64_t)1 << 31) && vocab > 0, "bm25 build sizes");
    ORAG_REQUIRE(pow2(tile_docs) && tile_docs >= 32 && tile_docs <= 65536, "tile_docs: power of two in [32, 65536]");
    ORAG_REQUIRE(pow2(fp_tile_docs) && fp_tile_docs >= 32 && fp_tile_docs <= 16384, "fp_tile_docs: power of two in [32, 16384]");
    g->n_docs = n_docs;
    g->vocab = vocab;
    g->tile_docs = tile_docs;
    g->fp_tile_docs = fp_tile_docs;
    g->super_docs = tile_docs > fp_tile_docs ? tile_docs : fp_tile_docs;
    g->n_tiles = (int)((n_docs + tile_docs - 1) / tile_docs);
    g->fp_n_tiles = (int)((n_docs + fp_tile_docs - 1) / fp_tile_docs);
    g->n_super = (int)((n_docs + g->super_docs - 1) / g->super_docs);
    return ORAG_OK;
}

struct BuildWs {
    int32_t *cur_e, *cur_f;        // [n_tiles, V], [fp_n_tiles, V]: counts in phase 1, cursors in phase 2
    long long *totals_e, *totals_f;
    size_t bytes;
};

static BuildWs carve(void *base, const Geometry &g)
{
    BuildWs w{};
    uint8_t *p = (uint8_t *)base;
    auto take = [&](size_t n) {
        uint8_t *r = p;
        p += align_up(n, 256);
        return r;
    };
    w.cur_e = (int32_t *)take((size_t)(g.n_tiles > 0 ? g.n_tiles : 1) * g.vocab * 4);
    w.cur_f = (int32_t *)take((size_t)(g.fp_n_tiles > 0 ? g.fp_n_tiles : 1) * g.vocab * 4);
    w.totals_e = (long long *)take((size_t)(g.n_tiles + 1) * 8);
    w.totals_f = (long long *)take((size_t)(g.fp_n_tiles + 1) * 8);
    w.bytes = (size_t)(p - (uint8_t *)base);
    return w;
}

static int grid_for(int n_super)
{
    int grid = (n_super + kWarps - 1) / kWarps;
    const int lim = sm_count() * 8;
    return grid < 1 ? 1 : (grid > lim ? lim : grid);
}


// ---- k-th largest impact of every term (threshold warm start of the first pass) ---------------------------------
// One CTA per term: two passes over the term's first-pass runs -- a histogram of the high byte of the fp16 impact
// bits (positive fp16 numbers order like their bit patterns), then, for each of the ORAG_BM25_KTH_LEVELS ranks, a
// histogram of the low byte inside the bin that holds it.  Padding postings (impact +0.0) are not counted.
// The K-th largest impact of ANY subset of the postings is a valid (if weaker) bound, so a term's scan stops after the
// first block of 256 tiles that brings it to kKthEnough postings: a term that occurs in every document is done after
// one block instead of streaming 10M postings through one CTA (the table cost 0.24 s of a 0.70 s build before, ~10 ms
// now); rare terms -- the ones whose high idf makes their bound the query's threshold -- are always scanned in full.
// Inside a block every lane looks up the run of one tile (32 tiles per warp and round trip), then the warp walks the
// non-empty runs together with four loads per lane in flight.
__constant__ int c_kth_k[ORAG_BM25_KTH_LEVELS] = {10, 16, 32, 64, 128};
constexpr uint32_t kKthEnough = 16384;

template <typename F>
__device__ __forceinline__ void kth_walk_block(int block, int t, int vocab, int fp_n_tiles,
                                               const long long *__restrict__ fp_tile_base,
                                               const int32_t *__restrict__ fp_tile_term_off,
                                               const uint32_t *__restrict__ postings_r16, F &&visit)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = (block * 8 + warp) * 32 + lane;
    int lo = 0, hi = 0;
    long long base = 0;
    if (tile < fp_n_tiles) {
        const int32_t *off = fp_tile_term_off + (int64_t)tile * ((int64_t)vocab + 1) + t;
        lo = __ldg(off);
        hi = __ldg(off + 1);
        base = fp_tile_base[tile];
    }
    for (unsigned todo = __ballot_sync(0xffffffffu, hi > lo); todo; todo &= todo - 1) {
        const int j = __ffs(todo) - 1;
        const int rlo = __shfl_sync(0xffffffffu, lo, j), rhi = __shfl_sync(0xffffffffu, hi, j);
        const uint32_t *run = postings_r16 + __shfl_sync(0xffffffffu, base, j);
        for (int i = rlo + lane; i < rhi; i += 128) {
            uint32_t key[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) key[u] = i + 32 * u < rhi ? __ldg(run + i + 32 * u) & 0xFFFFu : 0u;
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (key[u]) visit(key[u]);
        }
    }
}

__global__ void __launch_bounds__(256) term_kth_kernel(int vocab, int fp_n_tiles, const long long *__restrict__ fp_tile_base,
                                                      const int32_t *__restrict__ fp_tile_term_off,
                                                      const uint32_t *__restrict__ postings_r16, float *__restrict__ kth)
{
    __shared__ uint32_t h1[256];
    __shared__ uint32_t h2[ORAG_BM25_KTH_LEVELS][256];
    __shared__ int s_bin[ORAG_BM25_KTH_LEVELS];
    __shared__ uint32_t s_rank[ORAG_BM25_KTH_LEVELS];
    __shared__ uint32_t s_seen;
    const int t = blockIdx.x;
    h1[threadIdx.x] = 0;
    for (int l = 0; l < ORAG_BM25_KTH_LEVELS; ++l) h2[l][threadIdx.x] = 0;
    if (threadIdx.x == 0) s_seen = 0;
    __syncthreads();
    const int n_blocks = (fp_n_tiles + 255) / 256;
    int used = 0;   // blocks of 256 tiles that make up the subset both passes look at
    for (; used < n_blocks;) {
        uint32_t mine = 0;
        kth_walk_block(used, t, vocab, fp_n_tiles, fp_tile_base, fp_tile_term_off, postings_r16, [&](uint32_t key) {
            atomicAdd(&h1[key >> 8], 1u);
            ++mine;
        });
        if (mine) atomicAdd(&s_seen, mine);
        ++used;
        __syncthreads();
        if (s_seen >= kKthEnough) break;   // (uniform: read after the barrier, written before it)
        __syncthreads();
    }
    if (threadIdx.x < ORAG_BM25_KTH_LEVELS) {
        const uint32_t want = (uint32_t)c_kth_k[threadIdx.x];
        uint32_t acc = 0;
        int b = 255;
        for (; b >= 0; --b) {
            if (acc + h1[b] >= want) break;
            acc += h1[b];
        }
        s_bin[threadIdx.x] = b;            // -1: the subset holds fewer postings than this rank
        s_rank[threadIdx.x] = want - acc;  // rank inside the bin, 1-based
    }
    __syncthreads();
    bool any = false;
    for (int l = 0; l < ORAG_BM25_KTH_LEVELS; ++l) any |= s_bin[l] >= 0;
    if (any) {
        for (int blk = 0; blk < used; ++blk)
            kth_walk_block(blk, t, vocab, fp_n_tiles, fp_tile_base, fp_tile_term_off, postings_r16, [&](uint32_t key) {
                const int hb = (int)(key >> 8);
#pragma unroll
                for (int l = 0; l < ORAG_BM25_KTH_LEVELS; ++l)
                    if (hb == s_bin[l]) atomicAdd(&h2[l][key & 255u], 1u);
            });
    }
    __syncthreads();
    if (threadIdx.x < ORAG_BM25_KTH_LEVELS) {
        const int l = threadIdx.x;
        float v = 0.f;
        if (s_bin[l] >= 0) {
            uint32_t acc = 0;
            int b = 255;
            for (; b > 0; --b) {
                acc += h2[l][b];
                if (acc >= s_rank[l]) break;
            }
            v = __half2float(__ushort_as_half((unsigned short)((s_bin[l] << 8) | b)));
        }
        kth[(int64_t)l * vocab + t] = v;
    }
}

}  // namespace build
}  // namespace orag

using namespace orag;
using namespace orag::build;

extern "C" size_t orag_bm25_build_workspace_bytes(int64_t n_docs, int vocab, int tile_docs, int fp_tile_docs)
{
    Geometry g;
    if (geometry(n_docs, vocab, tile_docs, fp_tile_docs, &g) != ORAG_OK) return 0;
    return carve(nullptr, g).bytes;
}

extern "C" int orag_bm25_index_plan(const int64_t *d_doc_off, const int32_t *d_tokens, int64_t n_docs, int vocab,
                                    int tile_docs, int fp_tile_docs, int64_t token_pos_base, int32_t *d_doc_len,
                                    int64_t *d_df, int64_t *d_first_pos, int64_t *d_tile_base, int32_t *d_tile_term_off,
                                    int64_t *d_fp_tile_base, int32_t *d_fp_tile_term_off, int32_t *d_info,
                                    void *d_workspace, size_t workspace_bytes, int64_t *h_totals, void *stream)
{
    Geometry g;
    int rc = geometry(n_docs, vocab, tile_docs, fp_tile_docs, &g);
    if (rc) return rc;
    ORAG_REQUIRE(d_doc_off && d_df && d_first_pos && d_tile_base && d_fp_tile_base && d_info && h_totals,
                 "bm25_index_plan pointers");
    ORAG_REQUIRE(n_docs == 0 || (d_tokens && d_doc_len && d_tile_term_off && d_fp_tile_term_off), "per-document arrays");
    if (!d_workspace || workspace_bytes < carve(nullptr, g).bytes) {
        set_error("bm25_index_plan: workspace too small");
        return ORAG_EWORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    BuildWs w = carve(d_workspace, g);
    ORAG_CUDA_CHECK(cudaMemsetAsync(d_info, 0, 4 * sizeof(int32_t), st));
    ORAG_CUDA_CHECK(cudaMemsetAsync(w.cur_e, 0, (size_t)(g.n_tiles > 0 ? g.n_tiles : 1) * vocab * 4, st));
    ORAG_CUDA_CHECK(cudaMemsetAsync(w.cur_f, 0, (size_t)(g.fp_n_tiles > 0 ? g.fp_n_tiles : 1) * vocab * 4, st));
    if (n_docs > 0) {
        count_kernel<<<grid_for(g.n_super), kWarps * 32, 0, st>>>(d_doc_off, d_tokens, g, d_doc_len, w.cur_e, w.cur_f,
                                                                  d_info);
        ORAG_LAUNCH_CHECK();
        int64_t total_tokens = 0;
        ORAG_CUDA_CHECK(cudaMemcpyAsync(&total_tokens, d_doc_off + n_docs, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        ORAG_CUDA_CHECK(cudaStreamSynchronize(st));
        if (total_tokens > 0) {
            int64_t blocks = (total_tokens + 255) / 256;
            const int64_t lim = (int64_t)sm_count() * 16;
            first_pos_kernel<<<(unsigned)(blocks > lim ? lim : blocks), 256, 0, st>>>(
                d_tokens, total_tokens, vocab, token_pos_base, (long long *)d_first_pos);
            ORAG_LAUNCH_CHECK();
        }
        df_kernel<<<(vocab + 255) / 256, 256, 0, st>>>(w.cur_e, g.n_tiles, vocab, (long long *)d_df);
        ORAG_LAUNCH_CHECK();
        scan_rows_kernel<<<g.n_tiles, 1024, 0, st>>>(w.cur_e, vocab, 0, d_tile_term_off, w.totals_e, d_info);
        ORAG_LAUNCH_CHECK();
        scan_rows_kernel<<<g.fp_n_tiles, 1024, 0, st>>>(w.cur_f, vocab, 1, d_fp_tile_term_off, w.totals_f, d_info);
        ORAG_LAUNCH_CHECK();
    }
    scan_bases_kernel<<<1, 32, 0, st>>>(w.totals_e, g.n_tiles, (long long *)d_tile_base);
    ORAG_LAUNCH_CHECK();
    scan_bases_kernel<<<1, 32, 0, st>>>(w.totals_f, g.fp_n_tiles, (long long *)d_fp_tile_base);
    ORAG_LAUNCH_CHECK();
    ORAG_CUDA_CHECK(cudaMemcpyAsync(h_totals, d_tile_base + g.n_tiles, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    ORAG_CUDA_CHECK(cudaMemcpyAsync(h_totals + 1, d_fp_tile_base + g.fp_n_tiles, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    ORAG_CUDA_CHECK(cudaStreamSynchronize(st));
    return ORAG_OK;
}

extern "C" int orag_bm25_index_fill(const int64_t *d_doc_off, const int32_t *d_tokens, int64_t n_docs, int vocab,
                                    int tile_docs, int fp_tile_docs, const double *d_t4_table, int max_doc_len,
                                    const int64_t *d_tile_base, const int32_t *d_tile_term_off, uint32_t *d_postings,
                                    const int64_t *d_fp_tile_base, const int32_t *d_fp_tile_term_off,
                                    uint32_t *d_postings_r16, float *d_term_max_r, int32_t *d_info, void *d_workspace,
                                    size_t workspace_bytes, void *stream)
{
    Geometry g;
    int rc = geometry(n_docs, vocab, tile_docs, fp_tile_docs, &g);
    if (rc) return rc;
    ORAG_REQUIRE(d_doc_off && d_t4_table && d_tile_base && d_postings && d_fp_tile_base && d_info && max_doc_len >= 0,
                 "bm25_index_fill pointers");
    ORAG_REQUIRE(n_docs == 0 || (d_tokens && d_tile_term_off && d_fp_tile_term_off), "per-document arrays");
    ORAG_REQUIRE(!d_postings_r16 || d_term_max_r, "term_max_r goes with the first-pass view");
    ORAG_REQUIRE(!d_postings_r16 || (reinterpret_cast<uintptr_t>(d_postings_r16) & 15) == 0, "postings_r16 16-byte aligned");
    if (!d_workspace || workspace_bytes < carve(nullptr, g).bytes) {
        set_error("bm25_index_fill: workspace too small");
        return ORAG_EWORKSPACE;
    }
    if (n_docs == 0) return ORAG_OK;
    cudaStream_t st = (cudaStream_t)stream;
    BuildWs w = carve(d_workspace, g);   // the cursors were left at zero by orag_bm25_index_plan's scans
    if (d_postings_r16) ORAG_CUDA_CHECK(cudaMemsetAsync(d_term_max_r, 0, (size_t)vocab * sizeof(float), st));
    FillArgs a{};
    a.doc_off = d_doc_off;
    a.tokens = d_tokens;
    a.t4_table = d_t4_table;
    a.tile_base = (const long long *)d_tile_base;
    a.fp_tile_base = (const long long *)d_fp_tile_base;
    a.tile_term_off = d_tile_term_off;
    a.fp_tile_term_off = d_fp_tile_term_off;
    a.postings = d_postings;
    a.postings_r16 = d_postings_r16;
    a.term_max_bits = (uint32_t *)d_term_max_r;
    a.cur_e = w.cur_e;
    a.cur_f = w.cur_f;
    a.info = d_info;
    a.max_doc_len = max_doc_len;
    fill_kernel<<<grid_for(g.n_super), kWarps * 32, 0, st>>>(g, a);
    ORAG_LAUNCH_CHECK();
    if (d_postings_r16) {
        const int64_t n = (int64_t)g.fp_n_tiles * vocab;
        pad_runs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(g, (const long long *)d_fp_tile_base,
                                                                    d_fp_tile_term_off, w.cur_f, d_postings_r16);
        ORAG_LAUNCH_CHECK();
    }
    return ORAG_OK;
}

extern "C" int orag_bm25_term_kth(const orag_bm25_index_t *index, float *d_term_kth_r, void *stream)
{
    ORAG_REQUIRE(index && d_term_kth_r && index->vocab > 0, "bm25_term_kth");
    ORAG_REQUIRE(index->d_postings_r16 && index->d_fp_tile_base && index->d_fp_tile_term_off, "index without first-pass view");
    cudaStream_t st = (cudaStream_t)stream;
    if (index->n_docs == 0 || index->fp_n_tiles == 0) {
        ORAG_CUDA_CHECK(cudaMemsetAsync(d_term_kth_r, 0, (size_t)ORAG_BM25_KTH_LEVELS * index->vocab * sizeof(float), st));
        return ORAG_OK;
    }
    term_kth_kernel<<<index->vocab, 256, 0, st>>>(index->vocab, index->fp_n_tiles, (const long long *)index->d_fp_tile_base,
                                                  index->d_fp_tile_term_off, index->d_postings_r16, d_term_kth_r);
    ORAG_LAUNCH_CHECK();
    return ORAG_OK;
}
