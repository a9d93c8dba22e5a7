// select.cuh -- block-level exact top-k selection under (score desc, id asc).
//
// One CTA per query; k rounds of a block-wide arg-best over the entries that rank strictly
// after the previous pick.  O(k * n / blockDim) per thread: used on candidate lists (hundreds
// of entries), on gathered per-shard lists (G * k entries) and -- as the small-N / fallback
// path -- on dense score rows.
#pragma once
#include "common.cuh"

namespace orag {

struct Pick {
    double s;
    int64_t id;
    int valid;
};

__device__ __forceinline__ Pick better(Pick a, Pick b)
{
    if (!a.valid) return b;
    if (!b.valid) return a;
    return ranks_before(a.s, a.id, b.s, b.id) ? a : b;
}

__device__ __forceinline__ Pick warp_best(Pick p)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        Pick q;
        q.s = __shfl_xor_sync(0xffffffffu, p.s, o);
        q.id = __shfl_xor_sync(0xffffffffu, p.id, o);
        q.valid = __shfl_xor_sync(0xffffffffu, p.valid, o);
        p = better(p, q);
    }
    return p;
}

// Block-wide reduction of Pick; result valid in every thread.  `scratch` holds >= 32 Picks.
__device__ __forceinline__ Pick block_best(Pick p, Pick *scratch)
{
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    p = warp_best(p);
    __syncthreads();  // scratch reuse across rounds
    if (lane == 0) scratch[warp] = p;
    __syncthreads();
    Pick r;
    r.valid = 0; r.s = 0.0; r.id = 0;
    if (lane < nw) r = scratch[lane];
    r = warp_best(r);
    return r;
}

__device__ __forceinline__ double block_max(double v, double *scratch)
{
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    double r = (lane < nw) ? scratch[lane] : -INFINITY;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r = fmax(r, __shfl_xor_sync(0xffffffffu, r, o));
    return r;
}

}  // namespace orag
