// api.cu -- C-ABI entry points that orchestrate several kernels (cosine top-k pipeline), plus
// error plumbing.  See include/orag.h for the contract of every symbol.
#include <stdarg.h>
#include <string.h>

#include <atomic>

#include "common.cuh"
#include "cosine_tc.cuh"
#include "exact.cuh"

namespace orag {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ---- profiling hooks ------------------------------------------------------------------------------
static std::atomic<unsigned long long> g_launches{0};
static int g_profile = 0;
// ring of event pairs per slot: every bracket recorded since orag_profile_enable(1) (up to kProfRing) can be read back,
// so that a roofline figure rests on all timed steps rather than on the last one
constexpr int kProfRing = 256;
static cudaEvent_t g_ev[2][kProfRing][2];
static bool g_ev_made[2][kProfRing];
static int g_ev_n[2] = {0, 0};      // completed brackets per slot (may exceed kProfRing: the ring then holds the last ones)
static int g_ev_open[2] = {-1, -1}; // ring index of the bracket whose start has been recorded

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

void profile_mark(int slot, int end, cudaStream_t st)
{
    if (!g_profile || slot < 0 || slot > 1) return;
    if (!end) {
        const int i = g_ev_n[slot] % kProfRing;
        if (!g_ev_made[slot][i]) {
            cudaEventCreate(&g_ev[slot][i][0]);
            cudaEventCreate(&g_ev[slot][i][1]);
            g_ev_made[slot][i] = true;
        }
        cudaEventRecord(g_ev[slot][i][0], st);
        g_ev_open[slot] = i;
    } else if (g_ev_open[slot] >= 0) {
        cudaEventRecord(g_ev[slot][g_ev_open[slot]][1], st);
        g_ev_open[slot] = -1;
        ++g_ev_n[slot];
    }
}

// timeline: up to kTlCap tagged (start, end) event pairs since orag_timeline_enable(1)
constexpr int kTlCap = 4096;
static int g_timeline = 0;
static cudaEvent_t g_tl_epoch = nullptr;
static cudaEvent_t g_tl_ev[kTlCap][2];
static int g_tl_tag[kTlCap];
static int g_tl_made = 0;
static std::atomic<int> g_tl_n{0};

int timeline_mark(int tag, int handle, cudaStream_t st)
{
    if (!g_timeline) return -1;
    if (handle < 0) {
        const int i = g_tl_n.fetch_add(1, std::memory_order_relaxed);
        if (i >= g_tl_made) return -1;
        g_tl_tag[i] = tag;
        cudaEventRecord(g_tl_ev[i][0], st);
        return i;
    }
    cudaEventRecord(g_tl_ev[handle][1], st);
    return handle;
}

// inv_qnorm_s = 1 / (|q| * scale) from the fp16 conversion of the query block -> the other three views of the norm
__global__ void scale_norms_kernel(const float *inv_qnorm_s, const float *scale, int n, float *qnorm_s, float *qnorm,
                                   float *inv_qnorm)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const float inv = inv_qnorm_s[i];
        const float ns = inv > 0.f ? 1.f / inv : 0.f;
        qnorm_s[i] = ns;
        qnorm[i] = ns / scale[i];        // exact: scale is a power of two
        inv_qnorm[i] = inv * scale[i];
    }
}

// auxiliary stream + fork/join events, one set per device (a stream belongs to the device it was created on)
constexpr int kMaxDevices = 64;
struct AuxStream {
    cudaStream_t stream = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
    cudaEvent_t prescan = nullptr;  // co-scheduling hook (orag_cosine_mark_prescan)
    bool prescan_recorded = false;
};
static AuxStream g_aux[kMaxDevices];

static int aux_for_current_device(AuxStream **out)
{
    int dev = 0;
    ORAG_CUDA_CHECK(cudaGetDevice(&dev));
    ORAG_REQUIRE(dev >= 0 && dev < kMaxDevices, "device ordinal");
    AuxStream &a = g_aux[dev];
    if (!a.stream) {
        ORAG_CUDA_CHECK(cudaStreamCreateWithFlags(&a.stream, cudaStreamNonBlocking));
        ORAG_CUDA_CHECK(cudaEventCreateWithFlags(&a.fork, cudaEventDisableTiming));
        ORAG_CUDA_CHECK(cudaEventCreateWithFlags(&a.join, cudaEventDisableTiming));
        ORAG_CUDA_CHECK(cudaEventCreateWithFlags(&a.prescan, cudaEventDisableTiming));
    }
    *out = &a;
    return ORAG_OK;
}

// co-scheduling (orag_cosine_mark_prescan): make `st` wait for the point right before the latest main scan
int wait_prescan(cudaStream_t st)
{
    AuxStream *aux = nullptr;
    int rc = aux_for_current_device(&aux);
    if (rc) return rc;
    if (!aux->prescan_recorded) return ORAG_OK;  // nothing recorded yet on this device: nothing to wait for
    ORAG_CUDA_CHECK(cudaStreamWaitEvent(st, aux->prescan, 0));
    return ORAG_OK;
}

__global__ void unscale_columns_kernel(float *out, int64_t n, const float *scale, int n_queries)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int q = (int)(i & 255);
    if (i < n && q < n_queries) out[i] = out[i] / scale[q];
}

// ---- co-scheduling hook ---------------------------------------------------------------------------
static int g_mark_prescan = 0;

// First-pass error bounds in cosine units (DESIGN.md "exactness of the first pass"):
//   tf32: operands truncated to 10 mantissa bits -> |rel err per product| < 2^-9 + 2^-20
//   bf16: both operands rounded to nearest, 8 significant bits (unit roundoff 2^-8 each)
//                                                   -> |rel err per product| < 2^-7 + 2^-16
//   fp16: both operands rounded to nearest, 11 significant bits, rows and queries pre-scaled by powers of two
//         (cosine_exact.cu f32_to_f16_rows_kernel)  -> |rel err per product| < 2^-10 + 2^-22
// sum |a_i b_i| <= |a||b| turns that into an absolute bound on the cosine.  On top of the operand term comes the
// accumulation term, which grows with the vector length: a chain of `dim` fp32 additions whose every partial sum is
// bounded by sum |a_i b_i| contributes at most dim * 2^-23 (round-toward-zero worst case of the tensor core's
// accumulator), plus 64 * 2^-23 for the fp32 inverse norms, the epilogue multiply and the threshold arithmetic:
// first_pass_eps(dim) -- 2.5e-4 up to dim = 2033 (the constant the 1536-d captures were taken with), linear beyond.
constexpr float kOperandTf32 = 1.954e-3f;
constexpr float kOperandBf16 = 7.83e-3f;
constexpr float kOperandF16 = 9.77e-4f;

}  // namespace orag
float orag::tc::first_pass_eps(int mode, int dim)
{
    const float operand = mode == ORAG_COS_F16 ? kOperandF16 : mode == ORAG_COS_BF16 ? kOperandBf16 : kOperandTf32;
    const float accumulate = (float)(dim + 64) * 1.1920929e-7f;  // 2^-23
    return operand + (accumulate > 2.5e-4f ? accumulate : 2.5e-4f);
}
namespace orag {

constexpr int kGroup = 256;     // queries per tensor-core pass (UMMA N)
// Rows of the dense seed pass that initialises the thresholds (64 tiles: the pass needs whole SMs, and with fewer CTAs
// than SMs it can start while the last CTAs of the previous batch's scan are still draining).  The first wave of the main
// scan (two tiles per CTA in flight, ~38k rows) still runs on the seed's threshold, which passes k / kSeedRows of the
// rows: with 2048 seed rows that wave alone left ~190 candidates per query (k = 10) for the fp32 re-score, with 8192 ~45.
constexpr int kSeedRows = 64 * 128;   // (= 256 threads x 32 values of seed_finalize_kernel)
constexpr int kCandCap = 4096;  // first-pass candidate slots per query
constexpr int kSurvCap = 512;   // candidates that survive the fp32 re-score (k <= 128 leaves ample room)

struct CosineWs {
    double *sq_q;       // [G]
    float *qnorm;       // [G]
    float *inv_qnorm;   // [G]
    uint32_t *thr_key;  // [G]
    uint32_t *cnt;      // [G]
    uint32_t *hist;     // [G, 1024]
    int32_t *cand;      // [G, cap]   first-pass survivors (local row ids)
    float *cos32;       // [G, cap]   their fp32 re-scored cosines
    int32_t *surv;      // [G, cap2]  rows within 2*eps32 of the k-th best fp32 cosine
    uint32_t *surv_cnt; // [G]
    double *cand_score; // [G, cap2]  float64 cosines of the survivors
    int64_t *cand_id;   // [G, cap2]
    float *seed;        // [256, kSeedRows] (transposed: one row of first-pass values per query)
    void *q_bf16;       // [G, dim] bf16 / fp16 copy of the query block
    float *q_scale;     // [G] power-of-two scale of each fp16 query row
    float *qnorm_s;     // [G] |q| * scale     (the units the fp16 scan's accumulators are in)
    float *inv_qnorm_s; // [G] 1 / (|q| * scale)
    size_t bytes;
};

static CosineWs carve_tc(void *base, int dim)
{
    CosineWs w{};
    uint8_t *p = (uint8_t *)base;
    auto take = [&](size_t n) {
        uint8_t *r = p;
        p += align_up(n, 256);
        return r;
    };
    w.sq_q = (double *)take(kGroup * 8);
    w.qnorm = (float *)take(kGroup * 4);
    w.inv_qnorm = (float *)take(kGroup * 4);
    w.thr_key = (uint32_t *)take(kGroup * 4);
    w.cnt = (uint32_t *)take(kGroup * 4);
    w.hist = (uint32_t *)take((size_t)kGroup * tc::kHistBins * 4);
    w.cand = (int32_t *)take((size_t)kGroup * kCandCap * 4);
    w.cos32 = (float *)take((size_t)kGroup * kCandCap * 4);
    w.surv = (int32_t *)take((size_t)kGroup * kSurvCap * 4);
    w.surv_cnt = (uint32_t *)take(kGroup * 4);
    w.cand_score = (double *)take((size_t)kGroup * kSurvCap * 8);
    w.cand_id = (int64_t *)take((size_t)kGroup * kSurvCap * 8);
    w.seed = (float *)take((size_t)kSeedRows * 256 * 4);
    w.q_bf16 = take((size_t)kGroup * dim * 2);
    w.q_scale = (float *)take(kGroup * 4);
    w.qnorm_s = (float *)take(kGroup * 4);
    w.inv_qnorm_s = (float *)take(kGroup * 4);
    w.bytes = (size_t)(p - (uint8_t *)base);
    return w;
}

}  // namespace orag

using namespace orag;

extern "C" int orag_version(void) { return 1; }
extern "C" const char *orag_last_error(void) { return orag::g_err; }

extern "C" unsigned long long orag_launch_count(void) { return orag::g_launches.load(std::memory_order_relaxed); }

extern "C" int orag_profile_enable(int on)
{
    orag::g_profile = on ? 1 : 0;
    orag::g_ev_n[0] = orag::g_ev_n[1] = 0;
    orag::g_ev_open[0] = orag::g_ev_open[1] = -1;
    return ORAG_OK;
}

extern "C" int orag_timeline_enable(int on)
{
    orag::g_timeline = 0;
    if (on) {
        ORAG_CUDA_CHECK(cudaDeviceSynchronize());
        if (!orag::g_tl_epoch) ORAG_CUDA_CHECK(cudaEventCreate(&orag::g_tl_epoch));
        for (; orag::g_tl_made < orag::kTlCap; ++orag::g_tl_made) {
            ORAG_CUDA_CHECK(cudaEventCreate(&orag::g_tl_ev[orag::g_tl_made][0]));
            ORAG_CUDA_CHECK(cudaEventCreate(&orag::g_tl_ev[orag::g_tl_made][1]));
        }
        orag::g_tl_n.store(0);
        ORAG_CUDA_CHECK(cudaEventRecord(orag::g_tl_epoch, 0));
        ORAG_CUDA_CHECK(cudaDeviceSynchronize());
        orag::g_timeline = 1;
    }
    return ORAG_OK;
}

extern "C" int orag_timeline_read(int *tags, float *begin_ms, float *end_ms, int cap)
{
    ORAG_REQUIRE((tags && begin_ms && end_ms) || cap == 0, "timeline_read");
    ORAG_CUDA_CHECK(cudaDeviceSynchronize());
    int n = orag::g_tl_n.load();
    if (n > orag::g_tl_made) n = orag::g_tl_made;
    if (n > cap) n = cap;
    for (int i = 0; i < n; ++i) {
        tags[i] = orag::g_tl_tag[i];
        ORAG_CUDA_CHECK(cudaEventElapsedTime(begin_ms + i, orag::g_tl_epoch, orag::g_tl_ev[i][0]));
        ORAG_CUDA_CHECK(cudaEventElapsedTime(end_ms + i, orag::g_tl_epoch, orag::g_tl_ev[i][1]));
    }
    return n;
}

extern "C" int orag_cosine_last_counts(const void *d_workspace, int dim, int n_queries, uint32_t *h_candidates,
                                       uint32_t *h_survivors, void *stream)
{
    ORAG_REQUIRE(d_workspace && n_queries >= 1 && n_queries <= kGroup && h_candidates && h_survivors, "last_counts");
    CosineWs w = carve_tc(const_cast<void *>(d_workspace), dim);
    cudaStream_t st = (cudaStream_t)stream;
    ORAG_CUDA_CHECK(cudaMemcpyAsync(h_candidates, w.cnt, (size_t)n_queries * 4, cudaMemcpyDeviceToHost, st));
    ORAG_CUDA_CHECK(cudaMemcpyAsync(h_survivors, w.surv_cnt, (size_t)n_queries * 4, cudaMemcpyDeviceToHost, st));
    ORAG_CUDA_CHECK(cudaStreamSynchronize(st));
    return ORAG_OK;
}

extern "C" int orag_profile_read_all(int slot, float *ms, int cap)
{
    ORAG_REQUIRE(slot >= 0 && slot <= 1 && (ms || cap == 0) && cap >= 0, "profile_read_all");
    const int total = orag::g_ev_n[slot];
    const int have = total < orag::kProfRing ? total : orag::kProfRing;
    const int n = have < cap ? have : cap;
    // oldest first among the brackets still in the ring
    for (int j = 0; j < n; ++j) {
        const int i = (total - have + j) % orag::kProfRing;
        ORAG_CUDA_CHECK(cudaEventSynchronize(orag::g_ev[slot][i][1]));
        ORAG_CUDA_CHECK(cudaEventElapsedTime(ms + j, orag::g_ev[slot][i][0], orag::g_ev[slot][i][1]));
    }
    return n;
}

extern "C" int orag_profile_read(float *scan_ms, float *bm25_ms)
{
    float *out[2] = {scan_ms, bm25_ms};
    for (int s = 0; s < 2; ++s) {
        if (!out[s]) continue;
        *out[s] = -1.f;
        const int total = orag::g_ev_n[s];
        if (total <= 0) continue;
        const int i = (total - 1) % orag::kProfRing;
        ORAG_CUDA_CHECK(cudaEventSynchronize(orag::g_ev[s][i][1]));
        ORAG_CUDA_CHECK(cudaEventElapsedTime(out[s], orag::g_ev[s][i][0], orag::g_ev[s][i][1]));
    }
    return ORAG_OK;
}

extern "C" int orag_device_info(int *sm, int *major, int *minor)
{
    int dev = 0;
    ORAG_CUDA_CHECK(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    ORAG_CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
    if (sm) *sm = prop.multiProcessorCount;
    if (major) *major = prop.major;
    if (minor) *minor = prop.minor;
    return ORAG_OK;
}

extern "C" size_t orag_cosine_workspace_bytes(int64_t n_rows, int dim, int n_queries, int k, int mode)
{
    (void)k;
    if (n_queries <= 0 || dim <= 0) return 0;
    if (mode == ORAG_COS_EXACT) {
        // queries are processed in chunks whose dense score block stays under ~512 MiB
        size_t per_q = (size_t)(n_rows > 0 ? n_rows : 1) * 8;
        size_t chunk = ((size_t)512 << 20) / per_q;
        if (chunk < 1) chunk = 1;
        if (chunk > (size_t)n_queries) chunk = (size_t)n_queries;
        return align_up((size_t)n_queries * 8, 256) + align_up(chunk * per_q, 256);
    }
    return carve_tc(nullptr, dim).bytes;
}

static int cosine_exact_path(const float *corpus, int64_t n_rows, int dim, int64_t id_base, const float *queries,
                             int n_queries, int k, int64_t *out_ids, double *out_scores, void *ws, cudaStream_t st)
{
    double *sq = (double *)ws;
    double *dense = (double *)((uint8_t *)ws + align_up((size_t)n_queries * 8, 256));
    int rc = launch_query_sq(queries, n_queries, dim, sq, st);
    if (rc) return rc;
    size_t per_q = (size_t)(n_rows > 0 ? n_rows : 1) * 8;
    size_t chunk = ((size_t)512 << 20) / per_q;
    if (chunk < 1) chunk = 1;
    if (chunk > (size_t)n_queries) chunk = (size_t)n_queries;
    for (int q0 = 0; q0 < n_queries; q0 += (int)chunk) {
        int nq = n_queries - q0 < (int)chunk ? n_queries - q0 : (int)chunk;
        if (n_rows > 0) {
            rc = launch_cosine_dense(corpus, n_rows, dim, queries + (int64_t)q0 * dim, nq, sq + q0, dense, st);
            if (rc) return rc;
        }
        rc = launch_select_topk(dense, nullptr, nullptr, n_rows, n_rows, nq, k, id_base, 0, nullptr, 0,
                                out_ids + (int64_t)q0 * k, out_scores + (int64_t)q0 * k, nullptr, nullptr, st);
        if (rc) return rc;
    }
    return ORAG_OK;
}

extern "C" int orag_f32_to_bf16(const float *d_src, void *d_dst, int64_t count, void *stream);
extern "C" int orag_f32_to_f16_rows(const float *d_src, int64_t n_rows, int dim, void *d_dst_f16,
                                    float *d_inv_norm_scaled, float *d_scale, void *stream);

extern "C" int orag_cosine_topk_phase(const float *d_corpus, const float *d_inv_norm, const void *d_shadow,
                                      const double *d_row_sq, int64_t n_rows,
                                      int dim, int64_t row_id_base, const float *d_queries, int n_queries, int k, int mode,
                                      int64_t *d_out_ids, double *d_out_scores, int32_t *d_out_status, void *d_workspace,
                                      size_t workspace_bytes, int phases, void *stream)
{
    ORAG_REQUIRE(d_queries && d_out_ids && d_out_scores && n_queries > 0 && k > 0 && dim > 0 && n_rows >= 0,
                 "cosine_topk");
    ORAG_REQUIRE(phases >= 1 && phases <= ORAG_PHASE_ALL, "phases");
    ORAG_REQUIRE(phases == ORAG_PHASE_ALL || (n_queries <= kGroup && mode != ORAG_COS_EXACT),
                 "split phases: at most 256 queries, tensor-core modes only");
    const bool do_prep = (phases & ORAG_PHASE_PREP) != 0, do_scan = (phases & ORAG_PHASE_SCAN) != 0,
               do_finish = (phases & ORAG_PHASE_FINISH) != 0;
    ORAG_REQUIRE(n_rows == 0 || d_corpus, "corpus");
    ORAG_REQUIRE(n_rows < ((int64_t)1 << 31), "n_rows per shard must fit int32");
    cudaStream_t st = (cudaStream_t)stream;
    size_t need = orag_cosine_workspace_bytes(n_rows, dim, n_queries, k, mode);
    if (workspace_bytes < need || !d_workspace) {
        set_error("cosine_topk: workspace too small (%zu < %zu)", workspace_bytes, need);
        return ORAG_EWORKSPACE;
    }
    if (d_out_status && do_prep)
        ORAG_CUDA_CHECK(cudaMemsetAsync(d_out_status, 0, (size_t)n_queries * sizeof(int32_t), st));
    if (mode == ORAG_COS_EXACT)
        return cosine_exact_path(d_corpus, n_rows, dim, row_id_base, d_queries, n_queries, k, d_out_ids, d_out_scores,
                                 d_workspace, st);

    ORAG_REQUIRE(mode == ORAG_COS_TF32 || mode == ORAG_COS_BF16 || mode == ORAG_COS_F16, "mode");
    const bool f16 = mode == ORAG_COS_F16;
    const bool bf16 = mode == ORAG_COS_BF16 || f16;  // 16-bit operands (kind::f16)
    ORAG_REQUIRE(dim % (bf16 ? 64 : 32) == 0, "dim must be a multiple of 32 (tf32) / 64 (bf16, fp16)");
    ORAG_REQUIRE(d_inv_norm != nullptr || n_rows == 0, "inv_norm required for tensor-core modes");
    ORAG_REQUIRE(!bf16 || d_shadow != nullptr || n_rows == 0, "16-bit shadow required");
    ORAG_REQUIRE((reinterpret_cast<uintptr_t>(d_corpus) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_queries) & 15) == 0,
                 "16-byte alignment");
    ORAG_REQUIRE(k <= 128, "k <= 128 for tensor-core modes");
    CosineWs w = carve_tc(d_workspace, dim);
    const float margin = 2.f * tc::first_pass_eps(mode, dim);
    const int n_seed = (int)(n_rows < kSeedRows ? n_rows : kSeedRows);

    for (int q0 = 0; q0 < n_queries; q0 += kGroup) {
        const int nq = n_queries - q0 < kGroup ? n_queries - q0 : kGroup;
        const float *q = d_queries + (int64_t)q0 * dim;
        int rc;
        AuxStream *aux = nullptr;
        if (f16 || g_mark_prescan) {
            rc = aux_for_current_device(&aux);
            if (rc) return rc;
        }
        const void *qop = bf16 ? (const void *)w.q_bf16 : (const void *)q;
        const float *qnorm_scan = f16 ? w.qnorm_s : w.qnorm, *inv_qnorm_scan = f16 ? w.inv_qnorm_s : w.inv_qnorm;
        if (do_prep) {
            if (f16) {
                // The exact float64 sum(q*q) (one sequential compensated chain per query, ~65 us) is only needed by
                // the final re-score: it runs on an auxiliary stream next to the scan.  The scan's thresholds use fp32
                // norms that fall out of the fp16 conversion of the query block (error far inside the 2.5e-4 slack).
                ORAG_CUDA_CHECK(cudaEventRecord(aux->fork, st));
                ORAG_CUDA_CHECK(cudaStreamWaitEvent(aux->stream, aux->fork, 0));
                {
                    TimelineScope tl(TL_QUERY_SQ, aux->stream);
                    rc = launch_query_sq(q, nq, dim, w.sq_q, aux->stream);
                }
                if (rc) return rc;
                ORAG_CUDA_CHECK(cudaEventRecord(aux->join, aux->stream));
                TimelineScope tl(TL_QUERY_PREP, st);
                rc = orag_f32_to_f16_rows(q, nq, dim, w.q_bf16, w.inv_qnorm_s, w.q_scale, st);
                if (rc) return rc;
                scale_norms_kernel<<<(nq + 255) / 256, 256, 0, st>>>(w.inv_qnorm_s, w.q_scale, nq, w.qnorm_s, w.qnorm,
                                                                   w.inv_qnorm);
                ORAG_LAUNCH_CHECK();
            } else {
                rc = launch_query_sq(q, nq, dim, w.sq_q, st);
                if (rc) return rc;
                rc = tc::launch_query_norms(w.sq_q, nq, w.qnorm, w.inv_qnorm, st);
                if (rc) return rc;
                if (bf16) {
                    rc = orag_f32_to_bf16(q, w.q_bf16, (int64_t)nq * dim, st);
                    if (rc) return rc;
                }
            }
        }
        if (do_scan) {
            const void *aop = bf16 ? d_shadow : (const void *)d_corpus;
            tc::ScanParams p{};
            p.f16 = f16 ? 1 : 0;
            p.n_queries = nq;
            p.inv_norm = d_inv_norm;
            p.thr_key = w.thr_key;
            p.cnt = w.cnt;
            p.hist = w.hist;
            p.cand = w.cand;
            p.cand_v = w.cos32;  // first-pass values; the fp32 re-score overwrites them in place
            p.cap = kCandCap;
            p.qnorm = qnorm_scan;
            p.inv_qnorm = inv_qnorm_scan;
            p.margin = margin;
            p.k = k;
            // seed: dense first pass over the first rows -> histogram, threshold, first candidates
            p.row_begin = 0;
            p.row_end = n_seed;
            p.dense = 2;
            p.dense_out = w.seed;
            p.dense_ld = kSeedRows;
            {
                TimelineScope tl(TL_SEED_SCAN, st);
                rc = tc::launch_scan(bf16, aop, n_rows, qop, dim, p, st);
            }
            if (rc) return rc;
            {
                TimelineScope tl(TL_SEED_FINALIZE, st);
                rc = tc::launch_seed_finalize(w.seed, kSeedRows, n_seed, nq, k, margin, qnorm_scan, inv_qnorm_scan, w.thr_key, w.cnt,
                                              w.hist, w.cand, w.cos32, kCandCap, st);
            }
            if (rc) return rc;
            // main scan over the rest of the shard (co-scheduling hook: see orag_cosine_mark_prescan)
            if (g_mark_prescan) {
                ORAG_CUDA_CHECK(cudaEventRecord(aux->prescan, st));
                aux->prescan_recorded = true;
            }
            p.row_begin = n_seed;
            p.row_end = n_rows;
            p.dense = 0;
            p.dense_out = nullptr;
            {
                TimelineScope tl(TL_MAIN_SCAN, st);
                rc = tc::launch_scan(bf16, aop, n_rows, qop, dim, p, st);
            }
            if (rc) return rc;
        }
        if (!do_finish) continue;
        // fp32 re-score of the candidate set -> the few rows that can still reach the top-k ...
        {
            TimelineScope tl(TL_PREFILTER, st);
            rc = launch_prefilter(d_corpus, dim, q, w.inv_qnorm, w.cand, w.cnt, kCandCap, nq, k, w.cos32, kSurvCap, w.surv,
                                  w.surv_cnt, d_out_status ? d_out_status + q0 : nullptr, st, w.thr_key, qnorm_scan,
                                  margin);
        }
        if (rc) return rc;
        // ... exact float64 re-score of those (the reference's arithmetic), then exact selection.  (A split call's
        // finish phase waits for the sum(q*q) its own scan phase -- the latest one on this device -- put on the
        // auxiliary stream: callers issue scan and finish of one batch back to back.)
        if (f16) ORAG_CUDA_CHECK(cudaStreamWaitEvent(st, aux->join, 0));
        {
            TimelineScope tl(TL_RESCORE, st);
            rc = launch_rescore(d_corpus, dim, row_id_base, q, w.sq_q, w.surv, w.surv_cnt, kSurvCap, nq, d_row_sq,
                                w.cand_score, w.cand_id, st);
        }
        if (rc) return rc;
        TimelineScope tl_sel(TL_SELECT, st);
        rc = launch_select_topk(w.cand_score, w.cand_id, w.surv_cnt, kSurvCap, kSurvCap, nq, k, 0, 0, nullptr, 0,
                                d_out_ids + (int64_t)q0 * k, d_out_scores + (int64_t)q0 * k, nullptr,
                                d_out_status ? d_out_status + q0 : nullptr, st);
        if (rc) return rc;
    }
    return ORAG_OK;
}

extern "C" int orag_cosine_topk(const float *d_corpus, const float *d_inv_norm, const void *d_shadow,
                                const double *d_row_sq, int64_t n_rows,
                                int dim, int64_t row_id_base, const float *d_queries, int n_queries, int k, int mode,
                                int64_t *d_out_ids, double *d_out_scores, int32_t *d_out_status, void *d_workspace,
                                size_t workspace_bytes, void *stream)
{
    return orag_cosine_topk_phase(d_corpus, d_inv_norm, d_shadow, d_row_sq, n_rows, dim, row_id_base, d_queries, n_queries,
                                  k, mode, d_out_ids, d_out_scores, d_out_status, d_workspace, workspace_bytes,
                                  ORAG_PHASE_ALL, stream);
}

extern "C" int orag_cosine_firstpass_dense(const float *d_corpus, const float *d_inv_norm, const void *d_shadow,
                                           int64_t n_rows, int dim, const float *d_queries, int n_queries, int mode,
                                           float *d_out, void *d_workspace, size_t workspace_bytes, void *stream)
{
    ORAG_REQUIRE(d_corpus && d_inv_norm && d_queries && d_out && n_rows > 0 && n_queries > 0 && n_queries <= kGroup,
                 "firstpass_dense");
    ORAG_REQUIRE(mode == ORAG_COS_TF32 || mode == ORAG_COS_BF16 || mode == ORAG_COS_F16, "mode");
    const bool f16 = mode == ORAG_COS_F16;
    const bool bf16 = mode == ORAG_COS_BF16 || f16;
    ORAG_REQUIRE(dim % (bf16 ? 64 : 32) == 0, "dim multiple of 32/64");
    cudaStream_t st = (cudaStream_t)stream;
    const void *qop = d_queries;
    float *q_scale = nullptr;
    if (bf16) {
        ORAG_REQUIRE(d_shadow && d_workspace && workspace_bytes >= (size_t)n_queries * dim * 2 + 1024, "16-bit workspace");
        int rc;
        if (f16) {
            q_scale = (float *)((uint8_t *)d_workspace + align_up((size_t)n_queries * dim * 2, 256));
            ORAG_REQUIRE(workspace_bytes >= align_up((size_t)n_queries * dim * 2, 256) + (size_t)n_queries * 4,
                         "fp16 workspace");
            rc = orag_f32_to_f16_rows(d_queries, n_queries, dim, d_workspace, nullptr, q_scale, st);
        } else {
            rc = orag_f32_to_bf16(d_queries, d_workspace, (int64_t)n_queries * dim, st);
        }
        if (rc) return rc;
        qop = d_workspace;
    }
    tc::ScanParams p{};
    p.f16 = f16 ? 1 : 0;
    p.n_queries = n_queries;
    p.inv_norm = d_inv_norm;
    p.row_begin = 0;
    p.row_end = n_rows;
    p.dense = 1;
    p.dense_out = d_out;
    int rc = tc::launch_scan(bf16, bf16 ? d_shadow : (const void *)d_corpus, n_rows, qop, dim, p, st);
    if (rc || !f16) return rc;
    // fp16: the accumulators carry the queries' power-of-two scales; divide them out (exact)
    const int64_t total = (n_rows + 127) / 128 * 128 * 256;
    unscale_columns_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d_out, total, q_scale, n_queries);
    ORAG_LAUNCH_CHECK();
    return ORAG_OK;
}

extern "C" size_t orag_pairwise_workspace_bytes(int64_t m, int dim)
{
    (void)dim;
    return align_up((size_t)(m > 0 ? m : 1) * 8, 256);
}

extern "C" int orag_cosine_mark_prescan(int enable)
{
    orag::g_mark_prescan = enable ? 1 : 0;
    return ORAG_OK;
}

extern "C" int orag_stream_wait_prescan(void *stream) { return orag::wait_prescan((cudaStream_t)stream); }
