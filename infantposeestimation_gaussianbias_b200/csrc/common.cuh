// common.cuh — helpers shared by the gbcodec kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "../../include/gbcodec.h"

namespace gbc {

constexpr int kWarp = 32;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kEps = 1e-8f;

// ---- status plumbing (api.cu) ------------------------------------------------
int fail(int status, const char* fmt, ...);
int check_launch(const char* what);
// every kernel launch of the library goes through this counter (gbcodec_launch_count): bench.py reports launches it
// counted, not launches it assumes
void note_launch();

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- 128-bit global access with cache hints -----------------------------------
// Streaming read of data this SM will not touch again: skip L1 allocation.
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
// Read that is repeated by this SM within the kernel: let it live in L1.
__device__ __forceinline__ float4 ldg_keep(const float4* p) { return __ldg(p); }
// Streaming store (written once, never read back by this kernel).
__device__ __forceinline__ void stg_stream(float4* p, const float4& v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// ---- fast transcendental building blocks ---------------------------------------
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// sigmoid(h) = 1/(1+2^(-h*log2e)); relative error ~2^-21, monotone in h.
__device__ __forceinline__ float sigmoid_fast(float h) { return rcp(1.0f + ex2(-kLog2e * h)); }
// exact-ish sigmoid for the two learnable scalars (once per tile)
__device__ __forceinline__ float sigmoid_acc(float h) { return 1.0f / (1.0f + expf(-h)); }

// ---- warp reductions ---------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// Programmatic dependent launch (sm_90+): a kernel launched with the programmatic-stream-serialization attribute may
// start while its predecessor in the stream still runs — once every CTA of the predecessor has executed
// pdl_launch_dependents() (or exited) — and must execute pdl_wait() before it touches anything the predecessor
// writes; pdl_wait() returns when the predecessor grid has completed and its writes are visible.  Both are no-ops in a
// kernel launched the ordinary way.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// NV running sums per lane -> lane l holds the warp total of value l >> (5 - log2 NV) (halving butterfly:
// each step exchanges half of the remaining values, so 4 values cost 2+1+3 shuffles instead of 20).
template <int NV>
__device__ __forceinline__ void warp_scatter_sum(float (&v)[NV], int lane) {
#pragma unroll
    for (int n = NV, o = 16; n > 1; n >>= 1, o >>= 1) {
        const bool up = (lane & o) != 0;
#pragma unroll
        for (int j = 0; j < n / 2; ++j) {
            const float keep = up ? v[j + n / 2] : v[j];
            const float send = up ? v[j] : v[j + n / 2];
            v[j] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
    }
#pragma unroll
    for (int o = 16 / NV; o > 0; o >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], o);
}
template <int NV>
__device__ __forceinline__ void warp_scatter_sum(float (&v)[NV]) { warp_scatter_sum<NV>(v, threadIdx.x & 31); }   // 1-D blocks

// Block-wide sums of NV (power of two, <= 16) values with ONE barrier and a run-time warp count: halving
// butterfly inside the warp, then every warp adds the per-warp partials itself in a fixed order (deterministic,
// the same bits in every thread).  `red` holds nw * NV floats and must not be in use by a previous reduction
// that some warp may still be reading (alternate two buffers).  Results are written back into v[0..NV).
template <int NV>
__device__ __forceinline__ void block_sum_1bar(float (&v)[NV], float* red, int nw, int lane, int warp) {
    constexpr int SUB = 32 / NV;                       // lanes that share one value after the warp stage
    int sh = 0;
#pragma unroll
    for (int s = SUB; s > 1; s >>= 1) ++sh;            // log2(SUB)
    warp_scatter_sum<NV>(v, lane);
    const int idx = lane >> sh, q = lane & (SUB - 1);
    if (q == 0) red[warp * NV + idx] = v[0];
    __syncthreads();
    float acc = 0.f;
    for (int ww = q; ww < nw; ww += SUB) acc += red[ww * NV + idx];
#pragma unroll
    for (int o = SUB / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = __shfl_sync(0xffffffffu, acc, k << sh);
}

// (value, index) maximum; on equal values the smaller index wins — torch.max's
// first-occurrence rule (pose_estimator.py:352).
__device__ __forceinline__ void argmax_merge(float& v, int& i, float ov, int oi) {
    if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
}
__device__ __forceinline__ void warp_argmax(float& v, int& i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, v, o);
        int oi = __shfl_xor_sync(0xffffffffu, i, o);
        argmax_merge(v, i, ov, oi);
    }
}

// Block-wide sum of NV values per thread.  `scratch` holds 32*NV + NV floats.
// Two barriers; the cross-warp stage is done by the first NV threads in a fixed
// order, so the result is deterministic and identical in every thread.
template <int NV>
__device__ __forceinline__ void block_sum(float (&v)[NV], float* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = warp_sum(v[k]);
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) scratch[k * 32 + warp] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        float s = 0.f;
        for (int w = 0; w < nwarp; ++w) s += scratch[threadIdx.x * 32 + w];
        scratch[NV * 32 + threadIdx.x] = s;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = scratch[NV * 32 + k];
    __syncthreads();   // scratch may be reused right away
}

__device__ __forceinline__ float block_max(float v, float* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
    v = warp_max(v);
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    float r = scratch[0];
    for (int w = 1; w < nwarp; ++w) r = fmaxf(r, scratch[w]);
    __syncthreads();
    return r;
}

__device__ __forceinline__ void block_argmax(float& v, int& i, float* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
    warp_argmax(v, i);
    int* iscratch = reinterpret_cast<int*>(scratch + 32);
    if (lane == 0) { scratch[warp] = v; iscratch[warp] = i; }
    __syncthreads();
    v = scratch[0]; i = iscratch[0];
    for (int w = 1; w < nwarp; ++w) argmax_merge(v, i, scratch[w], iscratch[w]);
    __syncthreads();
}

// ---- encoder geometry (datasets/coco_dataset.py:208-248) ---------------------------
struct PatchGeom {
    int ulx, uly;        // patch origin = trunc(mu - 3 sigma), may be negative
    int x_from, x_to;    // pasted column range, clipped to the map (empty if x_to <= x_from)
    int y_from, y_to;
    float weight;        // weight after the visibility / off-map rule
    int active;          // 1 if any pixel is pasted
};
struct EncodeConst {
    double radius;       // 3*sigma
    float centre;        // floor((6 sigma + 1)/2)
    float two_sigma_sq;  // (float)(2 sigma^2)
    int ntap;            // ceil(6 sigma + 1)
    int lut_size;        // 2*centre_max^2 + 1 entries of exp(-d2 / 2sigma^2)
};
static inline EncodeConst make_encode_const(double sigma) {
    EncodeConst c;
    c.radius = sigma * 3.0;
    const double extent = 2.0 * c.radius + 1.0;
    c.ntap = (int)ceil(extent);
    c.centre = (float)floor(extent / 2.0);
    c.two_sigma_sq = (float)(2.0 * sigma * sigma);
    const int far = (int)fmax((double)c.centre, (double)(c.ntap - 1) - (double)c.centre);
    c.lut_size = 2 * far * far + 1;
    return c;
}
__device__ __forceinline__ PatchGeom patch_geometry(float kx, float ky, float vis, int H, int W,
                                                    float in_w, float in_h, const EncodeConst& ec) {
    PatchGeom g;
    g.weight = vis;
    g.active = 0;
    g.ulx = g.uly = g.x_from = g.x_to = g.y_from = g.y_to = 0;
    if (vis < 0.5f) return g;
    const double mux = (double)kx / ((double)in_w / (double)W);
    const double muy = (double)ky / ((double)in_h / (double)H);
    // int() of a Python float truncates toward zero; clamp first so the cast is defined
    const double lim = 1.0e9;
    const int ulx = (int)fmin(fmax(mux - ec.radius, -lim), lim);
    const int uly = (int)fmin(fmax(muy - ec.radius, -lim), lim);
    const int brx = (int)fmin(fmax(mux + ec.radius + 1.0, -lim), lim);
    const int bry = (int)fmin(fmax(muy + ec.radius + 1.0, -lim), lim);
    if (ulx >= W || uly >= H || brx < 0 || bry < 0) { g.weight = 0.f; return g; }
    g.ulx = ulx; g.uly = uly;
    g.x_from = max(0, ulx); g.x_to = min(min(brx, W), ulx + ec.ntap);
    g.y_from = max(0, uly); g.y_to = min(min(bry, H), uly + ec.ntap);
    g.active = (g.x_to > g.x_from) && (g.y_to > g.y_from);
    return g;
}
// exp(-d2/(2 sigma^2)) for every squared tap distance that can occur, in smem.
__device__ __forceinline__ void fill_patch_lut(float* lut, const EncodeConst& ec) {
    for (int i = threadIdx.x; i < ec.lut_size; i += blockDim.x)
        lut[i] = expf(-(float)i / ec.two_sigma_sq);
}
__device__ __forceinline__ float patch_value(const float* lut, const PatchGeom& g, const EncodeConst& ec, int x, int y) {
    const int dx = x - g.ulx - (int)ec.centre, dy = y - g.uly - (int)ec.centre;
    return lut[dx * dx + dy * dy];
}

}  // namespace gbc
