// common.cuh -- shared helpers for the sm_100a kernels (error plumbing, hashing,
// the reference's float64 cosine arithmetic as device functions).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/orag.h"

namespace orag {

void set_error(const char *fmt, ...);
// profiling hooks (api.cu): kernel-launch counter and optional event brackets around the two
// dominant kernels (slot 0 = cosine main scan, slot 1 = BM25 tile kernel)
void count_launch();
void profile_mark(int slot, int end, cudaStream_t st);
// timeline (api.cu, orag_timeline_enable): start / end events of every tagged launch of the hybrid step, readable as
// times since one epoch event -- the per-stream picture of how the batches in flight overlap.  No-ops when disabled.
enum TimelineTag {
    TL_QUERY_PREP = 1, TL_SEED_SCAN, TL_SEED_FINALIZE, TL_MAIN_SCAN, TL_PREFILTER, TL_RESCORE, TL_SELECT,
    TL_BM25_PREPARE, TL_BM25_FIRST_PASS, TL_BM25_FINALIZE, TL_RRF, TL_PUSH, TL_MERGE, TL_QUERY_SQ, TL_WAIT
};
int timeline_mark(int tag, int handle, cudaStream_t st);
int wait_prescan(cudaStream_t st);  // api.cu: order `st` after the point right before the latest cosine main scan
struct TimelineScope {
    int h;
    cudaStream_t st;
    TimelineScope(int tag, cudaStream_t s) : h(timeline_mark(tag, -1, s)), st(s) {}
    ~TimelineScope() { if (h >= 0) timeline_mark(0, h, st); }
};

#define ORAG_CUDA_CHECK(expr)                                                              \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            orag::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return ORAG_ECUDA;                                                             \
        }                                                                                  \
    } while (0)

#define ORAG_REQUIRE(cond, msg)                                                            \
    do {                                                                                   \
        if (!(cond)) {                                                                     \
            orag::set_error("invalid argument: %s (%s) (%s:%d)", msg, #cond, __FILE__, __LINE__); \
            return ORAG_EINVAL;                                                            \
        }                                                                                  \
    } while (0)

#define ORAG_LAUNCH_CHECK()                                                                \
    do {                                                                                   \
        orag::count_launch();                                                              \
        cudaError_t _e = cudaGetLastError();                                               \
        if (_e != cudaSuccess) {                                                           \
            orag::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
            return ORAG_ECUDA;                                                             \
        }                                                                                  \
    } while (0)

inline int sm_count()
{
    static int cached = 0;
    if (!cached) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev);
        if (cached <= 0) cached = 148;
    }
    return cached;
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- counter-based hashing (mirrors optimized_rag_b200/synthetic.py) --------------------------
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
#define ORAG_K_ROW 0xD1B54A32D192ED03ull
#define ORAG_K_DOC 0x9FB21C651E98DF25ull
#define ORAG_K_DUP 0xA24BAED4963EE407ull

__host__ __device__ __forceinline__ uint64_t row_key(uint64_t seed, uint64_t row)
{
    return mix64(seed ^ (row * ORAG_K_ROW));
}

// ---- the reference's float64 cosine (rag/retrieval.py:362-371 under CPython >= 3.12) ---------
// Neumaier-compensated running sum, exactly as builtin_sum() performs it for floats.
struct NeuSum {
    double s, c;
    // builtin_sum() starts from int 0: the first float item x0 enters as 0 + x0 == 0.0 + x0, after which
    // the compensated loop runs; starting the loop itself from s = 0.0, c = 0.0 is bit-identical (the
    // first step adds an exact zero to c) and needs no special case.
    __device__ __forceinline__ void init() { s = 0.0; c = 0.0; }
    __device__ __forceinline__ void add(double x)
    {
        const double t = __dadd_rn(s, x);
        const double hi = fabs(s) >= fabs(x) ? s : x;
        const double lo = fabs(s) >= fabs(x) ? x : s;
        c = __dadd_rn(c, __dadd_rn(__dadd_rn(hi, -t), lo));
        s = t;
    }
    __device__ __forceinline__ double result() const
    {
        if (c != 0.0 && isfinite(c)) return __dadd_rn(s, c);
        return s;
    }
};

// cosine from the three sums
__device__ __forceinline__ double cosine_from_sums(double dot, double sq_q, double sq_r)
{
    double m1 = sqrt(sq_q);
    double m2 = sqrt(sq_r);
    if (m1 == 0.0 || m2 == 0.0) return 0.0;
    return __ddiv_rn(dot, __dmul_rn(m1, m2));
}

// ---- ordering helpers ---------------------------------------------------------------------------
// true when (s1,id1) ranks strictly before (s2,id2): score desc, id asc
__device__ __forceinline__ bool ranks_before(double s1, int64_t id1, double s2, int64_t id2)
{
    return (s1 > s2) || (s1 == s2 && id1 < id2);
}

// monotone map float -> uint32 so that unsigned compare == float compare (for atomicMax)
__device__ __forceinline__ uint32_t float_to_ordered(float f)
{
    uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(uint32_t u)
{
    uint32_t b = (u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u;
    return __uint_as_float(b);
}

}  // namespace orag
