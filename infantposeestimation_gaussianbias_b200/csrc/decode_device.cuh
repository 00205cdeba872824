// decode_device.cuh — device helpers shared by decode.cu and the fused step in loss.cu.
#pragma once
#include "common.cuh"
#include <cuda_fp16.h>

namespace gbc {

__device__ __forceinline__ float4 rev4(const float4& v) { return make_float4(v.w, v.z, v.y, v.x); }

// one element as float: the maps are float32, or float16 under autocast (the values are up-cast, never the sums)
__device__ __forceinline__ float ld1(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ld1(const __half* p) { return __half2float(__ldg(p)); }

// value of the (possibly flip-averaged) tile at one pixel; same arithmetic as the vector load path
template <typename T>
__device__ __forceinline__ float tile_at(const T* hm_tile, const T* hmf_tile, int W, int x, int y) {
    float v = ld1(hm_tile + y * W + x);
    if (hmf_tile) v = (v + ld1(hmf_tile + y * W + (W - 1 - x))) * 0.5f;
    return v;
}

// Bilinear read of a 2-channel offset tile at (cx,cy): grid_sample(bilinear, border,
// align_corners=True) of fusion_head.py:353-359 — coordinates clamped to the map,
// taps outside the map contribute nothing.
struct Bilinear {
    int x0, y0, x1, y1;
    float w00, w01, w10, w11;   // first index y, second x
    float okx, oky;
    float fx, fy;
};
__device__ __forceinline__ Bilinear bilinear_setup(float cx, float cy, int H, int W) {
    Bilinear b;
    const float ccx = fminf(fmaxf(cx, 0.f), (float)(W - 1));
    const float ccy = fminf(fmaxf(cy, 0.f), (float)(H - 1));
    const float fx0 = floorf(ccx), fy0 = floorf(ccy);
    b.x0 = (int)fx0; b.y0 = (int)fy0;
    b.fx = ccx - fx0; b.fy = ccy - fy0;
    b.okx = (b.x0 + 1 < W) ? 1.f : 0.f;
    b.oky = (b.y0 + 1 < H) ? 1.f : 0.f;
    b.x1 = min(b.x0 + 1, W - 1); b.y1 = min(b.y0 + 1, H - 1);
    b.w00 = (1.f - b.fx) * (1.f - b.fy); b.w01 = b.fx * (1.f - b.fy);
    b.w10 = (1.f - b.fx) * b.fy;         b.w11 = b.fx * b.fy;
    return b;
}
template <typename T>
__device__ __forceinline__ float bilinear_read(const T* ch, const Bilinear& b, int W) {
    const float v00 = ld1(ch + b.y0 * W + b.x0);
    const float v01 = ld1(ch + b.y0 * W + b.x1) * b.okx;
    const float v10 = ld1(ch + b.y1 * W + b.x0) * b.oky;
    const float v11 = ld1(ch + b.y1 * W + b.x1) * (b.okx * b.oky);
    return b.w00 * v00 + b.w01 * v01 + b.w10 * v10 + b.w11 * v11;
}

// Steps 3-6 of the decode for one tile, executed by warp 0 once the global
// soft-argmax (cx, cy) is known.  Result in lane 0.
template <typename T>
__device__ __forceinline__ void refine_and_correct(const T* hm_tile, const T* hmf_tile, const T* off_tile,
                                                   const float* alpha_param, const float* fusion_weight,
                                                   int H, int W, int radius, unsigned flags,
                                                   float& cx, float& cy, int& px_out, int& py_out) {
    const int lane = threadIdx.x & 31;
    // torch.round is round-half-to-even == rintf in the default rounding mode
    const int px = (int)fminf(fmaxf(rintf(cx), 0.f), (float)(W - 1));
    const int py = (int)fminf(fmaxf(rintf(cy), 0.f), (float)(H - 1));
    px_out = px; py_out = py;
    if (flags & GBCODEC_DECODE_REFINE) {
        const int S = 2 * radius + 1;
        float vmax = -INFINITY;
        for (int c = lane; c < S * S; c += 32) {
            const int x = px - radius + c % S, y = py - radius + c / S;
            if (x >= 0 && x < W && y >= 0 && y < H) vmax = fmaxf(vmax, tile_at(hm_tile, hmf_tile, W, x, y));
        }
        vmax = warp_max(vmax);
        float se = 0.f, sx = 0.f, sy = 0.f;
        for (int c = lane; c < S * S; c += 32) {
            const int x = px - radius + c % S, y = py - radius + c / S;
            if (x >= 0 && x < W && y >= 0 && y < H) {
                const float e = expf(tile_at(hm_tile, hmf_tile, W, x, y) - vmax);
                se += e; sx += e * (float)x; sy += e * (float)y;
            }
        }
        se = warp_sum(se); sx = warp_sum(sx); sy = warp_sum(sy);
        const float a = sigmoid_acc(__ldg(alpha_param));
        cx = a * cx + (1.f - a) * (sx / se);
        cy = a * cy + (1.f - a) * (sy / se);
    }
    if (flags & GBCODEC_DECODE_APPLY_OFFSET) {
        float fw = __ldg(fusion_weight);
        if (flags & GBCODEC_DECODE_FUSION_WEIGHT_RAW) fw = sigmoid_acc(fw);
        const Bilinear bl = bilinear_setup(cx, cy, H, W);
        const float ox = bilinear_read(off_tile, bl, W);
        const float oy = bilinear_read(off_tile + H * W, bl, W);
        cx += fw * ox;
        cy += fw * oy;
    }
}

// Sub-pixel step of the arg-max family for the peak at flat index `at` of tile `t`.
__device__ __forceinline__ void subpixel_step(const float* __restrict__ t, int at, int H, int W, int mode, float& fx, float& fy) {
    const int x = at % W, y = at / W;
    fx = (float)x; fy = (float)y;
    if (mode == GBCODEC_ARGMAX_QUARTER) {
        if (x > 0 && x < W - 1 && y > 0 && y < H - 1) {
            const float dx = t[y * W + x + 1] - t[y * W + x - 1];
            const float dy = t[(y + 1) * W + x] - t[(y - 1) * W + x];
            fx += (dx > 0.f ? 0.25f : (dx < 0.f ? -0.25f : 0.f));
            fy += (dy > 0.f ? 0.25f : (dy < 0.f ? -0.25f : 0.f));
        }
    } else if (mode == GBCODEC_ARGMAX_TAYLOR) {
        // utils/postprocess.py:57-73: strict '1 <', fp32 differences, the division in double
        if (x > 1 && x < W - 1 && y > 1 && y < H - 1) {
            const float c = t[y * W + x];
            const float xl = t[y * W + x - 1], xr = t[y * W + x + 1];
            const float yu = t[(y - 1) * W + x], yd = t[(y + 1) * W + x];
            const float dx = xr - xl, dy = yd - yu;
            const float dxx = __fadd_rn(__fsub_rn(xr, __fmul_rn(2.f, c)), xl);
            const float dyy = __fadd_rn(__fsub_rn(yd, __fmul_rn(2.f, c)), yu);
            if (dxx < 0.f) {
                double o = (double)dx / (2.0 * fabs((double)dxx));
                o = fmin(fmax(o, -0.5), 0.5);
                fx = __fadd_rn(fx, (float)o);
            }
            if (dyy < 0.f) {
                double o = (double)dy / (2.0 * fabs((double)dyy));
                o = fmin(fmax(o, -0.5), 0.5);
                fy = __fadd_rn(fy, (float)o);
            }
        }
    }
}

}  // namespace gbc
