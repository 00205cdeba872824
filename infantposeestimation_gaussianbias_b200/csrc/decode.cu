// decode.cu — keypoint decoding kernels.
//
//   decode_kernel        HeatmapRegressionHead.decode (models/fusion_head.py:309-365)
//                        with the flip-test average of PoseEstimator.inference
//                        (models/pose_estimator.py:303-319) folded into the load
//   argmax_kernel        PoseEstimator.decode_heatmaps (models/pose_estimator.py:331-373),
//                        get_max_preds / get_max_preds_with_subpixel (utils/postprocess.py:10-75)
//   centroid_kernel      coordinate_refinement (utils/postprocess.py:138-184)
//
// One CTA per (image, keypoint) tile.  The tile is read once from HBM with
// 128-bit loads and stays in registers for both softmax passes; the 5x5 window
// and the 8 offset taps are re-read through L2.  Roofline: HBM read, 4*H*W
// bytes per tile (8*H*W with the flip average).
#include "common.cuh"
#include "decode_device.cuh"
#include <stdlib.h>
#include <string.h>

namespace gbc {

// Schedule of one tile (NITER > 0: the tile fits the CTA's registers):
//   1. the whole tile with 128-bit streaming loads; the two learnable scalars at once
//   2. max            — one barrier (every warp finishes the cross-warp step itself)
//   3. exp moments    — one barrier; every thread then knows the soft-argmax (cx, cy)
//   4. the threads that hold pixels of the (2r+1)^2 window around round(cx, cy) drop them into
//      shared memory (no second trip to L2), while warp 0 already has the 4x4x2 block of offset
//      taps around floor(cx, cy) in flight — the one dependent DRAM access of the tail is issued
//      before the coordinate it depends on is final, and covers |refined - global| <= 1 px
//      (anything further takes the direct path)
//   5. one barrier; warp 0: window softmax from shared memory, blend, taps picked by shuffle.
constexpr int kMaxWin = 17 * 17;           // local_radius <= 8

template <int NITER, bool FLIP, int MAXT = 1024, int MINB = 1>
__global__ void __launch_bounds__(MAXT, MINB)
decode_kernel(const float* __restrict__ hm, const float* __restrict__ hmf, const int32_t* __restrict__ perm,
              const float* __restrict__ off, const float* __restrict__ alpha_param,
              const float* __restrict__ fusion_weight, int K, int H, int W, int radius, unsigned flags,
              float* __restrict__ coords, float* __restrict__ scores, int32_t* __restrict__ centre) {
    __shared__ float red0[32], red1[32 * 4];
    __shared__ float win[kMaxWin];
    const int tile = blockIdx.x;
    const int b = tile / K, k = tile - b * K;
    const int n = H * W, n4 = n >> 2, w4 = W >> 2;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    const float* hm_tile = hm + (size_t)tile * n;
    const float* hmf_tile = nullptr;
    if (FLIP) {
        const int kk = perm ? __ldg(perm + k) : k;
        hmf_tile = hmf + ((size_t)b * K + kk) * n;
    }
    const float4* src = reinterpret_cast<const float4*>(hm_tile);
    const float4* srcf = reinterpret_cast<const float4*>(hmf_tile);

    constexpr int R = NITER > 0 ? NITER : 1;
    float4 v[R];
    auto load = [&](int i) {
        float4 t = ldg_stream(src + i);
        if (FLIP) {
            const int y = i / w4, xq = i - y * w4;
            const float4 f = rev4(ldg_stream(srcf + y * w4 + (w4 - 1 - xq)));
            t.x = (t.x + f.x) * 0.5f; t.y = (t.y + f.y) * 0.5f; t.z = (t.z + f.z) * 0.5f; t.w = (t.w + f.w) * 0.5f;
        }
        return t;
    };
    float m = -INFINITY;
    if (NITER > 0) {
#pragma unroll
        for (int it = 0; it < R; ++it) v[it] = load(it * blockDim.x + tid);
    }
    // scalars of the tail: issued now, consumed three barriers later
    float a_raw = 0.f, fw = 0.f;
    if (warp == 0) {
        if (flags & GBCODEC_DECODE_REFINE) a_raw = __ldg(alpha_param);
        if (flags & GBCODEC_DECODE_APPLY_OFFSET) fw = __ldg(fusion_weight);
    }
    const int S = 2 * radius + 1;
    if (NITER > 0 && (flags & GBCODEC_DECODE_REFINE)) {
        for (int c = tid; c < S * S; c += blockDim.x) win[c] = -INFINITY;       // off-map window cells stay -inf
    }
    if (NITER > 0) {
#pragma unroll
        for (int it = 0; it < R; ++it) m = fmaxf(m, fmaxf(fmaxf(v[it].x, v[it].y), fmaxf(v[it].z, v[it].w)));
    } else {
        for (int i = tid; i < n4; i += blockDim.x) {
            const float4 t = load(i);
            m = fmaxf(m, fmaxf(fmaxf(t.x, t.y), fmaxf(t.z, t.w)));
        }
    }
    m = warp_max(m);
    if (lane == 0) red0[warp] = m;
    __syncthreads();
    m = warp_max(lane < nw ? red0[lane] : -INFINITY);

    float acc[3] = {0.f, 0.f, 0.f};     // sum e, sum e*x, sum e*y
    const float ml = m * kLog2e;
    auto accumulate = [&](const float4& t, int i) {
        const int y = i / w4, x = (i - y * w4) << 2;
        const float e0 = ex2(fmaf(t.x, kLog2e, -ml)), e1 = ex2(fmaf(t.y, kLog2e, -ml));
        const float e2 = ex2(fmaf(t.z, kLog2e, -ml)), e3 = ex2(fmaf(t.w, kLog2e, -ml));
        const float se = (e0 + e1) + (e2 + e3);
        acc[0] += se;
        acc[1] += fmaf((float)x, se, fmaf(3.f, e3, fmaf(2.f, e2, e1)));
        acc[2] = fmaf((float)y, se, acc[2]);
    };
    if (NITER > 0) {
#pragma unroll
        for (int it = 0; it < R; ++it) accumulate(v[it], it * blockDim.x + tid);
    } else {
        for (int i = tid; i < n4; i += blockDim.x) accumulate(load(i), i);
    }
#pragma unroll
    for (int q = 0; q < 3; ++q) acc[q] = warp_sum(acc[q]);
    if (lane == 0) { red1[warp * 4] = acc[0]; red1[warp * 4 + 1] = acc[1]; red1[warp * 4 + 2] = acc[2]; }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 3; ++q) acc[q] = warp_sum(lane < nw ? red1[lane * 4 + q] : 0.f);   // fixed order: same bits in every warp
    float cx = acc[1] / acc[0], cy = acc[2] / acc[0];

    if (NITER == 0) {
        // generic shapes: the tail re-reads its few pixels through L2
        if (tid < 32) {
            int px, py;
            refine_and_correct(hm_tile, hmf_tile, off ? off + (size_t)tile * 2 * n : nullptr, alpha_param, fusion_weight,
                               H, W, radius, flags, cx, cy, px, py);
            if (tid == 0) {
                coords[2 * tile] = cx; coords[2 * tile + 1] = cy;
                scores[tile] = m;
                if (centre) { centre[2 * tile] = px; centre[2 * tile + 1] = py; }
            }
        }
        return;
    }

    // torch.round is round-half-to-even == rintf in the default rounding mode
    const int px = (int)fminf(fmaxf(rintf(cx), 0.f), (float)(W - 1));
    const int py = (int)fminf(fmaxf(rintf(cy), 0.f), (float)(H - 1));
    // offset taps: a 4x4 block per channel around floor(clamp(cx, cy)) - 1, one tap per lane of warp 0
    const float* off_tile = off ? off + (size_t)tile * 2 * n : nullptr;
    const bool want_off = (flags & GBCODEC_DECODE_APPLY_OFFSET) != 0;
    const int bx = (int)floorf(fminf(fmaxf(cx, 0.f), (float)(W - 1))) - 1;
    const int by = (int)floorf(fminf(fmaxf(cy, 0.f), (float)(H - 1))) - 1;
    float pre = 0.f;
    if (want_off && warp == 0) {
        const int t = lane & 15;
        const int tx = min(max(bx + (t & 3), 0), W - 1), ty = min(max(by + (t >> 2), 0), H - 1);
        pre = __ldg(off_tile + (lane >> 4) * n + ty * W + tx);
    }
    if (flags & GBCODEC_DECODE_REFINE) {
#pragma unroll
        for (int it = 0; it < R; ++it) {
            const int i = it * blockDim.x + tid;
            const int y = i / w4, x0 = (i - y * w4) << 2;
            const int wy = y - py + radius;
            if (wy >= 0 && wy < S && x0 + 3 >= px - radius && x0 <= px + radius) {
                const float e[4] = {v[it].x, v[it].y, v[it].z, v[it].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int wx = x0 + j - px + radius;
                    if (wx >= 0 && wx < S) win[wy * S + wx] = e[j];
                }
            }
        }
        __syncthreads();
    }
    if (warp != 0) return;

    if (flags & GBCODEC_DECODE_REFINE) {
        float vmax = -INFINITY;
        for (int c = lane; c < S * S; c += 32) vmax = fmaxf(vmax, win[c]);
        vmax = warp_max(vmax);
        float se = 0.f, sx = 0.f, sy = 0.f;
        for (int c = lane; c < S * S; c += 32) {
            const float wv = win[c];
            if (wv != -INFINITY) {
                const int x = px - radius + c % S, y = py - radius + c / S;
                const float e = expf(wv - vmax);
                se += e; sx += e * (float)x; sy += e * (float)y;
            }
        }
        se = warp_sum(se); sx = warp_sum(sx); sy = warp_sum(sy);
        const float a = sigmoid_acc(__shfl_sync(0xffffffffu, a_raw, 0));
        cx = a * cx + (1.f - a) * (sx / se);
        cy = a * cy + (1.f - a) * (sy / se);
    }
    if (want_off) {
        if (flags & GBCODEC_DECODE_FUSION_WEIGHT_RAW) fw = sigmoid_acc(fw);
        const Bilinear bl = bilinear_setup(cx, cy, H, W);
        float ox, oy;
        const bool covered = bl.x0 >= bx && bl.x1 <= bx + 3 && bl.y0 >= by && bl.y1 <= by + 3;   // warp-uniform
        if (covered) {
            const int i00 = (bl.y0 - by) * 4 + (bl.x0 - bx), i01 = (bl.y0 - by) * 4 + (bl.x1 - bx);
            const int i10 = (bl.y1 - by) * 4 + (bl.x0 - bx), i11 = (bl.y1 - by) * 4 + (bl.x1 - bx);
            float t[2][4];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                t[c][0] = __shfl_sync(0xffffffffu, pre, c * 16 + i00);
                t[c][1] = __shfl_sync(0xffffffffu, pre, c * 16 + i01) * bl.okx;
                t[c][2] = __shfl_sync(0xffffffffu, pre, c * 16 + i10) * bl.oky;
                t[c][3] = __shfl_sync(0xffffffffu, pre, c * 16 + i11) * (bl.okx * bl.oky);
            }
            ox = bl.w00 * t[0][0] + bl.w01 * t[0][1] + bl.w10 * t[0][2] + bl.w11 * t[0][3];
            oy = bl.w00 * t[1][0] + bl.w01 * t[1][1] + bl.w10 * t[1][2] + bl.w11 * t[1][3];
        } else {
            ox = bilinear_read(off_tile, bl, W);
            oy = bilinear_read(off_tile + n, bl, W);
        }
        const float f = __shfl_sync(0xffffffffu, fw, 0);
        cx += f * ox;
        cy += f * oy;
    }
    if (lane == 0) {
        coords[2 * tile] = cx; coords[2 * tile + 1] = cy;
        scores[tile] = m;
        if (centre) { centre[2 * tile] = px; centre[2 * tile + 1] = py; }
    }
}

// ---------------------------------------------------------------------------------
// Column-owner variant for the three tile shapes of the reference's configs (64x48, 96x72,
// 128x128): TPB = (W/4) * ROWS threads, a thread owns the same four columns in every row it
// visits, so there is no index arithmetic in the pixel loops (the generic kernel spends half
// of its instructions on i / w4) and the x-moment factors out of the row loop:
//   sum e x = x0 * Z_t + (E1 + 2 E2 + 3 E3),  E_j = column sums of this thread.
// Same schedule otherwise (one barrier per reduction, window scatter, prefetched offset taps).
// ---------------------------------------------------------------------------------
// Reductions: ONE barrier for max and moments together.  A warp takes the softmax of its own pixels
// relative to its own maximum; the per-warp {Z, sum e x, sum e y, max} meet in shared memory and
// every warp rescales them to the tile maximum (exp2((m_w - m) log2 e), one per warp) while adding —
// the same sum, 23 shuffles instead of 40 and one barrier less.
__host__ __device__ constexpr int ceil_log2(int n) { return n <= 1 ? 0 : 1 + ceil_log2((n + 1) / 2); }

template <int W4, int ROWS, int NIT, bool FLIP, int MINB>
__global__ void __launch_bounds__(W4* ROWS, MINB)
decode_tile_kernel(const float* __restrict__ hm, const float* __restrict__ hmf, const int32_t* __restrict__ perm,
                   const float* __restrict__ off, const float* __restrict__ alpha_param,
                   const float* __restrict__ fusion_weight, int K, int radius, unsigned flags,
                   float* __restrict__ coords, float* __restrict__ scores, int32_t* __restrict__ centre) {
    constexpr int TPB = W4 * ROWS, NW = TPB / 32, N4 = TPB * NIT, N = 4 * N4, W = 4 * W4, H = ROWS * NIT;
    constexpr int NWP = 1 << ceil_log2(NW);                    // lanes that take part in the cross-warp step
    static_assert(TPB % 32 == 0 && NW <= 32, "CTA must be whole warps");
    __shared__ float4 red[32];
    __shared__ float win[kMaxWin];
    const int tile = blockIdx.x;
    const int tx = threadIdx.x, ty = threadIdx.y;              // block = (W4, ROWS): the column group and first row of this thread
    const int tid = ty * W4 + tx, lane = tid & 31, warp = tid >> 5;
    const bool refine = (flags & GBCODEC_DECODE_REFINE) != 0, want_off = (flags & GBCODEC_DECODE_APPLY_OFFSET) != 0;
    const float4* src = reinterpret_cast<const float4*>(hm + (size_t)tile * N) + tid;
    float4 v[NIT];
#pragma unroll
    for (int it = 0; it < NIT; ++it) v[it] = ldg_stream(src + it * TPB);
    float4 f[FLIP ? NIT : 1];
    if (FLIP) {
        const int b = tile / K, k = tile - b * K;
        const int kk = perm ? __ldg(perm + k) : k;
        // the mirrored float4 of the same row: column group W4-1-tx, lanes reversed
        const float4* srcf = reinterpret_cast<const float4*>(hmf + ((size_t)b * K + kk) * N) + ty * W4 + (W4 - 1 - tx);
#pragma unroll
        for (int it = 0; it < NIT; ++it) f[it] = ldg_stream(srcf + it * TPB);
    }
    // the two learnable scalars and their sigmoids: in the shadow of the tile loads
    float a = 0.f, fw = 0.f;
    if (warp == 0) {
        if (refine) a = sigmoid_acc(__ldg(alpha_param));
        if (want_off) {
            fw = __ldg(fusion_weight);
            if (flags & GBCODEC_DECODE_FUSION_WEIGHT_RAW) fw = sigmoid_acc(fw);
        }
    }
    if (FLIP) {
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            const float4 g = rev4(f[it]);
            v[it].x = (v[it].x + g.x) * 0.5f; v[it].y = (v[it].y + g.y) * 0.5f;
            v[it].z = (v[it].z + g.z) * 0.5f; v[it].w = (v[it].w + g.w) * 0.5f;
        }
    }
    float mw = -INFINITY;
#pragma unroll
    for (int it = 0; it < NIT; ++it) mw = fmaxf(mw, fmaxf(fmaxf(v[it].x, v[it].y), fmaxf(v[it].z, v[it].w)));
    mw = warp_max(mw);

    const float mlw = mw * kLog2e;
    float Ej[4] = {0.f, 0.f, 0.f, 0.f};
    float Yw = 0.f;
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
        const float e0 = ex2(fmaf(v[it].x, kLog2e, -mlw)), e1 = ex2(fmaf(v[it].y, kLog2e, -mlw));
        const float e2 = ex2(fmaf(v[it].z, kLog2e, -mlw)), e3 = ex2(fmaf(v[it].w, kLog2e, -mlw));
        Ej[0] += e0; Ej[1] += e1; Ej[2] += e2; Ej[3] += e3;
        Yw = fmaf((float)(it * ROWS), (e0 + e1) + (e2 + e3), Yw);
    }
    const float Zt = (Ej[0] + Ej[1]) + (Ej[2] + Ej[3]);
    float acc[4];
    acc[0] = Zt;
    acc[1] = fmaf((float)(tx << 2), Zt, fmaf(3.f, Ej[3], fmaf(2.f, Ej[2], Ej[1])));
    acc[2] = fmaf((float)ty, Zt, Yw);
    acc[3] = 0.f;
    warp_scatter_sum<4>(acc, lane);                                  // lane l: warp total of value l >> 3
    if ((lane & 7) == 0) reinterpret_cast<float*>(red + warp)[lane >> 3] = lane == 24 ? mw : acc[0];
    __syncthreads();
    // every warp: rescale the per-warp moments to the tile maximum and add them, in lane order (same bits everywhere)
    const float4 r = lane < NW ? red[lane] : make_float4(0.f, 0.f, 0.f, -INFINITY);
    float m = r.w;
#pragma unroll
    for (int o = NWP / 2; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    m = __shfl_sync(0xffffffffu, m, 0);
    const float sc = lane < NW ? ex2((r.w - m) * kLog2e) : 0.f;
    float Z = r.x * sc, X = r.y * sc, Y = r.z * sc;
#pragma unroll
    for (int o = NWP / 2; o > 0; o >>= 1) {
        Z += __shfl_xor_sync(0xffffffffu, Z, o);
        X += __shfl_xor_sync(0xffffffffu, X, o);
        Y += __shfl_xor_sync(0xffffffffu, Y, o);
    }
    // lanes 0..NWP-1 hold the totals; the window scatter below needs them in every lane
    const float iZ = 1.0f / __shfl_sync(0xffffffffu, Z, 0);
    float cx = __shfl_sync(0xffffffffu, X, 0) * iZ, cy = __shfl_sync(0xffffffffu, Y, 0) * iZ;

    // torch.round is round-half-to-even == rintf in the default rounding mode
    const int px = (int)fminf(fmaxf(rintf(cx), 0.f), (float)(W - 1));
    const int py = (int)fminf(fmaxf(rintf(cy), 0.f), (float)(H - 1));
    const int S = 2 * radius + 1;
    int bx = 0, by = 0;
    float pre = 0.f;
    const float* off_tile = nullptr;
    if (want_off && warp == 0) {
        off_tile = off + (size_t)tile * 2 * N;
        bx = (int)floorf(fminf(fmaxf(cx, 0.f), (float)(W - 1))) - 1;
        by = (int)floorf(fminf(fmaxf(cy, 0.f), (float)(H - 1))) - 1;
        const int t = lane & 15;
        const int qx = min(max(bx + (t & 3), 0), W - 1), qy = min(max(by + (t >> 2), 0), H - 1);
        pre = __ldg(off_tile + (lane >> 4) * N + qy * W + qx);
    }
    if (refine) {
        const int x0 = tx << 2;
        if (x0 + 3 >= px - radius && x0 <= px + radius) {
#pragma unroll
            for (int it = 0; it < NIT; ++it) {
                const int wy = it * ROWS + ty - py + radius;
                if (wy >= 0 && wy < S) {
                    const float e[4] = {v[it].x, v[it].y, v[it].z, v[it].w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int wx = x0 + j - px + radius;
                        if (wx >= 0 && wx < S) win[wy * S + wx] = e[j];
                    }
                }
            }
        }
        __syncthreads();
    }
    if (warp != 0) return;

    if (refine) {
        // window cells outside the map were never written: validity comes from the coordinates
        float vmax = -INFINITY;
        for (int c = lane; c < S * S; c += 32) {
            const int x = px - radius + c % S, y = py - radius + c / S;
            if (x >= 0 && x < W && y >= 0 && y < H) vmax = fmaxf(vmax, win[c]);
        }
        vmax = warp_max(vmax);
        float sw[4] = {0.f, 0.f, 0.f, 0.f};
        for (int c = lane; c < S * S; c += 32) {
            const int x = px - radius + c % S, y = py - radius + c / S;
            if (x >= 0 && x < W && y >= 0 && y < H) {
                const float e = expf(win[c] - vmax);
                sw[0] += e; sw[1] += e * (float)x; sw[2] += e * (float)y;
            }
        }
        warp_scatter_sum<4>(sw, lane);
        const float se = __shfl_sync(0xffffffffu, sw[0], 0), sx = __shfl_sync(0xffffffffu, sw[0], 8), sy = __shfl_sync(0xffffffffu, sw[0], 16);
        const float al = __shfl_sync(0xffffffffu, a, 0);
        cx = al * cx + (1.f - al) * (sx / se);
        cy = al * cy + (1.f - al) * (sy / se);
    }
    if (want_off) {
        const Bilinear bl = bilinear_setup(cx, cy, H, W);
        float ox, oy;
        const bool covered = bl.x0 >= bx && bl.x1 <= bx + 3 && bl.y0 >= by && bl.y1 <= by + 3;   // warp-uniform
        if (covered) {
            const int i00 = (bl.y0 - by) * 4 + (bl.x0 - bx), i01 = (bl.y0 - by) * 4 + (bl.x1 - bx);
            const int i10 = (bl.y1 - by) * 4 + (bl.x0 - bx), i11 = (bl.y1 - by) * 4 + (bl.x1 - bx);
            float t[2][4];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                t[c][0] = __shfl_sync(0xffffffffu, pre, c * 16 + i00);
                t[c][1] = __shfl_sync(0xffffffffu, pre, c * 16 + i01) * bl.okx;
                t[c][2] = __shfl_sync(0xffffffffu, pre, c * 16 + i10) * bl.oky;
                t[c][3] = __shfl_sync(0xffffffffu, pre, c * 16 + i11) * (bl.okx * bl.oky);
            }
            ox = bl.w00 * t[0][0] + bl.w01 * t[0][1] + bl.w10 * t[0][2] + bl.w11 * t[0][3];
            oy = bl.w00 * t[1][0] + bl.w01 * t[1][1] + bl.w10 * t[1][2] + bl.w11 * t[1][3];
        } else {
            ox = bilinear_read(off_tile, bl, W);
            oy = bilinear_read(off_tile + N, bl, W);
        }
        const float fq = __shfl_sync(0xffffffffu, fw, 0);
        cx += fq * ox;
        cy += fq * oy;
    }
    if (lane == 0) {
        coords[2 * tile] = cx; coords[2 * tile + 1] = cy;
        scores[tile] = m;
        if (centre) { centre[2 * tile] = px; centre[2 * tile + 1] = py; }
    }
}

template <int W4, int ROWS, int NIT, bool FLIP, int MINB>
static int launch_decode_tile(const float* hm, const float* hmf, const int32_t* perm, const float* off, const float* ap,
                              const float* fw, int tiles, int K, int radius, unsigned flags, float* coords, float* scores,
                              int32_t* centre, cudaStream_t s) {
    note_launch(), decode_tile_kernel<W4, ROWS, NIT, FLIP, MINB><<<tiles, dim3(W4, ROWS), 0, s>>>(hm, hmf, perm, off, ap, fw, K, radius, flags, coords, scores, centre);
    return check_launch("decode_tile_kernel");
}

// ---------------------------------------------------------------------------------
// One WARP per tile (tiles of at most 16 KB: 64x48, 64x64; no flip average).  The CTA-per-tile kernel above spends ~2 400
// warp instructions on a 12 KB tile, most of them on what six warps of 16 pixels per thread owe each other (two block
// barriers, the cross-warp sums, 16 window tests per thread for the scatter) and runs at 72 % issue utilisation: it is
// instruction-bound at 5.2 TB/s where a read-only stream reaches 7 TB/s.  Here a warp owns the whole tile: lane l holds the
// float4 32 j + l (j < N4/32: 24 x 128-bit loads in flight per lane, 512 coalesced bytes per instruction), takes the maximum
// and the softmax moments from registers, and five shuffles per sum finish the tile — no shared memory, no barrier, ~700
// instructions.  A warp's tail (window + offset taps: two dependent L2 / HBM round trips) stalls nobody else: the other warps
// of the SM are independent tiles in other phases.  The window pixels are re-read through L2 (the tile has just come
// through it); their sums run in decode_tile_kernel's order.
// Walking 32 float4 ahead moves a lane A = 32 / W4 rows down and Bc = 32 % W4 column groups right (wrapping into the next
// row), so the pixel coordinates are two running floats per lane.
// ---------------------------------------------------------------------------------
constexpr int kDecodeWarpThreads = 128;
template <int W4, int HH, int MINB>
__global__ void __launch_bounds__(kDecodeWarpThreads, MINB)
decode_warp_kernel(const float* __restrict__ hm, const float* __restrict__ off, const float* __restrict__ alpha_param,
                   const float* __restrict__ fusion_weight, int tiles, int radius, unsigned flags,
                   float* __restrict__ coords, float* __restrict__ scores, int32_t* __restrict__ centre) {
    constexpr int W = 4 * W4, H = HH, N4 = W4 * HH, N = 4 * N4, NJ = N4 / 32;
    constexpr int A = 32 / W4, Bc = 32 % W4;
    static_assert(N4 % 32 == 0 && NJ <= 32, "a lane holds at most 32 float4 of the tile");
    const int lane = threadIdx.x & 31;
    const int tile = blockIdx.x * (kDecodeWarpThreads / 32) + (threadIdx.x >> 5);
    if (tile >= tiles) return;
    const bool refine = (flags & GBCODEC_DECODE_REFINE) != 0, want_off = (flags & GBCODEC_DECODE_APPLY_OFFSET) != 0;
    const float* hm_tile = hm + (size_t)tile * N;
    const float4* src = reinterpret_cast<const float4*>(hm_tile) + lane;
    float4 v[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) v[j] = ldg_stream(src + 32 * j);
    // the two learnable scalars and their sigmoids: in the shadow of the tile loads
    float a = 0.f, fw = 0.f;
    if (refine) a = sigmoid_acc(__ldg(alpha_param));
    if (want_off) {
        fw = __ldg(fusion_weight);
        if (flags & GBCODEC_DECODE_FUSION_WEIGHT_RAW) fw = sigmoid_acc(fw);
    }
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < NJ; ++j) m = fmaxf(m, fmaxf(fmaxf(v[j].x, v[j].y), fmaxf(v[j].z, v[j].w)));
    m = warp_max(m);
    const float ml = m * kLog2e;
    float fc = (float)((lane % W4) << 2), fr = (float)(lane / W4);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};                          // Z, sum e x, sum e y
    float E1 = 0.f, E2 = 0.f, E3 = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        const float e0 = ex2(fmaf(v[j].x, kLog2e, -ml)), e1 = ex2(fmaf(v[j].y, kLog2e, -ml));
        const float e2 = ex2(fmaf(v[j].z, kLog2e, -ml)), e3 = ex2(fmaf(v[j].w, kLog2e, -ml));
        const float sr = (e0 + e1) + (e2 + e3);
        acc[0] += sr;
        acc[1] = fmaf(fc, sr, acc[1]);
        acc[2] = fmaf(fr, sr, acc[2]);
        E1 += e1; E2 += e2; E3 += e3;
        fc += (float)(4 * Bc); fr += (float)A;
        if (Bc != 0 && fc >= (float)W) { fc -= (float)W; fr += 1.f; }
    }
    acc[1] += fmaf(3.f, E3, fmaf(2.f, E2, E1));
    warp_scatter_sum<4>(acc, lane);                               // lane l: total of value l >> 3
    const float iZ = 1.0f / __shfl_sync(0xffffffffu, acc[0], 0);
    float cx = __shfl_sync(0xffffffffu, acc[0], 8) * iZ, cy = __shfl_sync(0xffffffffu, acc[0], 16) * iZ;

    // torch.round is round-half-to-even == rintf in the default rounding mode
    const int px = (int)fminf(fmaxf(rintf(cx), 0.f), (float)(W - 1));
    const int py = (int)fminf(fmaxf(rintf(cy), 0.f), (float)(H - 1));
    const int S = 2 * radius + 1;
    int bx = 0, by = 0;
    float pre = 0.f;
    const float* off_tile = off + (size_t)tile * 2 * N;
    if (want_off) {
        // the 4x4x2 block of offset taps around floor(cx, cy): in flight before the refined coordinate is final
        bx = (int)floorf(fminf(fmaxf(cx, 0.f), (float)(W - 1))) - 1;
        by = (int)floorf(fminf(fmaxf(cy, 0.f), (float)(H - 1))) - 1;
        const int t = lane & 15;
        const int qx = min(max(bx + (t & 3), 0), W - 1), qy = min(max(by + (t >> 2), 0), H - 1);
        pre = __ldg(off_tile + (lane >> 4) * N + qy * W + qx);
    }
    if (refine) {
        // window cells outside the map do not exist: validity comes from the coordinates
        float first = -INFINITY;                                  // this lane's first window pixel stays in a register
        {
            const int x = px - radius + lane % S, y = py - radius + lane / S;
            if (lane < S * S && x >= 0 && x < W && y >= 0 && y < H) first = __ldg(hm_tile + y * W + x);
        }
        float vmax = first;
        for (int c = lane + 32; c < S * S; c += 32) {
            const int x = px - radius + c % S, y = py - radius + c / S;
            if (x >= 0 && x < W && y >= 0 && y < H) vmax = fmaxf(vmax, __ldg(hm_tile + y * W + x));
        }
        vmax = warp_max(vmax);
        float sw[4] = {0.f, 0.f, 0.f, 0.f};
        for (int c = lane; c < S * S; c += 32) {
            const int x = px - radius + c % S, y = py - radius + c / S;
            if (x >= 0 && x < W && y >= 0 && y < H) {
                const float e = expf((c == lane ? first : __ldg(hm_tile + y * W + x)) - vmax);
                sw[0] += e; sw[1] += e * (float)x; sw[2] += e * (float)y;
            }
        }
        warp_scatter_sum<4>(sw, lane);
        const float se = __shfl_sync(0xffffffffu, sw[0], 0), sx = __shfl_sync(0xffffffffu, sw[0], 8), sy = __shfl_sync(0xffffffffu, sw[0], 16);
        cx = a * cx + (1.f - a) * (sx / se);
        cy = a * cy + (1.f - a) * (sy / se);
    }
    if (want_off) {
        const Bilinear bl = bilinear_setup(cx, cy, H, W);
        float ox, oy;
        const bool covered = bl.x0 >= bx && bl.x1 <= bx + 3 && bl.y0 >= by && bl.y1 <= by + 3;   // warp-uniform
        if (covered) {
            const int i00 = (bl.y0 - by) * 4 + (bl.x0 - bx), i01 = (bl.y0 - by) * 4 + (bl.x1 - bx);
            const int i10 = (bl.y1 - by) * 4 + (bl.x0 - bx), i11 = (bl.y1 - by) * 4 + (bl.x1 - bx);
            float t[2][4];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                t[c][0] = __shfl_sync(0xffffffffu, pre, c * 16 + i00);
                t[c][1] = __shfl_sync(0xffffffffu, pre, c * 16 + i01) * bl.okx;
                t[c][2] = __shfl_sync(0xffffffffu, pre, c * 16 + i10) * bl.oky;
                t[c][3] = __shfl_sync(0xffffffffu, pre, c * 16 + i11) * (bl.okx * bl.oky);
            }
            ox = bl.w00 * t[0][0] + bl.w01 * t[0][1] + bl.w10 * t[0][2] + bl.w11 * t[0][3];
            oy = bl.w00 * t[1][0] + bl.w01 * t[1][1] + bl.w10 * t[1][2] + bl.w11 * t[1][3];
        } else {
            ox = bilinear_read(off_tile, bl, W);
            oy = bilinear_read(off_tile + N, bl, W);
        }
        cx += fw * ox;
        cy += fw * oy;
    }
    if (lane == 0) {
        coords[2 * tile] = cx; coords[2 * tile + 1] = cy;
        scores[tile] = m;
        if (centre) { centre[2 * tile] = px; centre[2 * tile + 1] = py; }
    }
}

template <int W4, int HH, int MINB>
static int launch_decode_warp(const float* hm, const float* off, const float* ap, const float* fw, int tiles, int radius,
                              unsigned flags, float* coords, float* scores, int32_t* centre, cudaStream_t s) {
    const int per_cta = kDecodeWarpThreads / 32;
    note_launch(), decode_warp_kernel<W4, HH, MINB><<<(tiles + per_cta - 1) / per_cta, kDecodeWarpThreads, 0, s>>>(hm, off, ap, fw, tiles, radius, flags, coords, scores, centre);
    return check_launch("decode_warp_kernel");
}

// Thread count T (multiple of 32) with n4 == T*NITER, NITER <= 8; 0 if none.
int pick_threads(int n4, int* niter) {
    for (int pass = 0; pass < 2; ++pass) {
        const int lo = pass == 0 ? 256 : 64, hi = pass == 0 ? 1024 : 224;
        for (int t = lo; t <= hi; t += 32)
            if (n4 % t == 0 && n4 / t <= 8) { *niter = n4 / t; return t; }
    }
    *niter = 0;
    return 0;
}

template <bool FLIP>
static void launch_decode_t(int niter, int grid, int threads, cudaStream_t s,
                            const float* hm, const float* hmf, const int32_t* perm, const float* off,
                            const float* ap, const float* fw, int K, int H, int W, int radius, unsigned flags,
                            float* coords, float* scores, int32_t* centre) {
    // the 64x48 tile (256 threads x 3 float4): hold the register count down so that 6 CTAs (no flip) fit an SM
    if (niter == 3 && threads == 256) {
        note_launch(), decode_kernel<3, FLIP, 256, FLIP ? 5 : 6><<<grid, threads, 0, s>>>(hm, hmf, perm, off, ap, fw, K, H, W, radius, flags, coords, scores, centre);
        return;
    }
#define GBC_CASE(NI) case NI: note_launch(), decode_kernel<NI, FLIP><<<grid, threads, 0, s>>>(hm, hmf, perm, off, ap, fw, K, H, W, radius, flags, coords, scores, centre); break;
    switch (niter) {
        GBC_CASE(1) GBC_CASE(2) GBC_CASE(3) GBC_CASE(4) GBC_CASE(5) GBC_CASE(6) GBC_CASE(7) GBC_CASE(8)
        default: note_launch(), decode_kernel<0, FLIP><<<grid, threads, 0, s>>>(hm, hmf, perm, off, ap, fw, K, H, W, radius, flags, coords, scores, centre); break;
    }
#undef GBC_CASE
}

int launch_decode(const float* hm, const float* hmf, const int32_t* perm, const float* off,
                  const float* alpha_param, const float* fusion_weight, int B, int K, int H, int W,
                  int radius, unsigned flags, float* coords, float* scores, int32_t* centre, cudaStream_t stream) {
    // GBCODEC_DECODE_KERNEL=generic: the shape-agnostic kernel for every shape (A/B measurements)
    static const bool generic = [] { const char* e = getenv("GBCODEC_DECODE_KERNEL"); return e && !strcmp(e, "generic"); }();
    const int grid_t = B * K;
#define GBC_TILE(W4, ROWS, NIT, MB_PLAIN, MB_FLIP)                                                                          \
    do {                                                                                                                  \
        if (hmf) return launch_decode_tile<W4, ROWS, NIT, true, MB_FLIP>(hm, hmf, perm, off, alpha_param, fusion_weight, grid_t, K, radius, flags, coords, scores, centre, stream); \
        return launch_decode_tile<W4, ROWS, NIT, false, MB_PLAIN>(hm, hmf, perm, off, alpha_param, fusion_weight, grid_t, K, radius, flags, coords, scores, centre, stream); \
    } while (0)
    // GBCODEC_DECODE_KERNEL=tile: the CTA-per-tile kernel where the warp-per-tile one is the default (A/B measurements)
    const char* const sel = getenv("GBCODEC_DECODE_KERNEL");                 // read on every call: tests compare the two kernels in one process
    const bool cta_tile = sel && !strcmp(sel, "tile");
    if (!generic && !cta_tile && !hmf) {
        if (H == 64 && W == 48) return launch_decode_warp<12, 64, 4>(hm, off, alpha_param, fusion_weight, grid_t, radius, flags, coords, scores, centre, stream);
        if (H == 64 && W == 64) return launch_decode_warp<16, 64, 3>(hm, off, alpha_param, fusion_weight, grid_t, radius, flags, coords, scores, centre, stream);
    }
    if (!generic) {
        if (H == 64 && W == 48) GBC_TILE(12, 16, 4, 10, 8);     // 192 threads, 16 px each
        if (H == 64 && W == 64) GBC_TILE(16, 16, 4, 8, 6);      // 256 threads, 16 px each
        if (H == 96 && W == 72) GBC_TILE(18, 16, 6, 5, 4);      // 288 threads, 24 px each
        if (H == 128 && W == 128) GBC_TILE(32, 16, 8, 3, 2);    // 512 threads, 32 px each
    }
#undef GBC_TILE
    int niter = 0;
    int threads = pick_threads((H * W) >> 2, &niter);
    if (!threads) threads = 256;
    if (hmf) launch_decode_t<true>(niter, B * K, threads, stream, hm, hmf, perm, off, alpha_param, fusion_weight, K, H, W, radius, flags, coords, scores, centre);
    else     launch_decode_t<false>(niter, B * K, threads, stream, hm, hmf, perm, off, alpha_param, fusion_weight, K, H, W, radius, flags, coords, scores, centre);
    return check_launch("decode_kernel");
}

// coordinate_refinement (utils/postprocess.py:138-184) for one tile, by one warp: linear-weight
// centroid of the window around trunc(ix, iy); an empty window keeps the input (so does a window that
// lies wholly left of / above the map, where the reference's slice wraps around and its torch.arange
// raises).  Result in every lane.
__device__ __forceinline__ void centroid_window(const float* __restrict__ t, float ix, float iy, int H, int W, int window,
                                                float& ox, float& oy) {
    const int lane = threadIdx.x & 31;
    const int half = window / 2;
    // int() truncates toward zero; clamp first so that the cast is defined for any finite input
    const int x = (int)fminf(fmaxf(ix, -1.0e9f), 1.0e9f), y = (int)fminf(fmaxf(iy, -1.0e9f), 1.0e9f);
    const int x_min = max(0, x - half), x_max = min(W, x + half + 1);
    const int y_min = max(0, y - half), y_max = min(H, y + half + 1);
    const int ww = x_max - x_min, hh = y_max - y_min;
    ox = ix; oy = iy;
    if (ww > 0 && hh > 0) {
        float s = 0.f, sx = 0.f, sy = 0.f;
        for (int c = lane; c < ww * hh; c += 32) {
            const int xx = x_min + c % ww, yy = y_min + c / ww;
            const float v = __ldg(t + yy * W + xx);
            s += v; sx += v * (float)xx; sy += v * (float)yy;
        }
        s = warp_sum(s); sx = warp_sum(sx); sy = warp_sum(sy);
        const float d = s + kEps;
        ox = sx / d; oy = sy / d;
    }
}

// ---------------------------------------------------------------------------------
// arg-max family
// ---------------------------------------------------------------------------------
template <int NITER>
__global__ void __launch_bounds__(1024)
argmax_kernel(const float* __restrict__ hm, int H, int W, int mode,
              float* __restrict__ coords, float* __restrict__ maxvals, int32_t* __restrict__ index) {
    __shared__ float scratch[64];
    const int tile = blockIdx.x;
    const int n = H * W, n4 = n >> 2;
    const float* t = hm + (size_t)tile * n;
    const float4* src = reinterpret_cast<const float4*>(t);
    float best = -INFINITY;
    int at = 0x7fffffff;
    auto visit = [&](const float4& q, int i) {
        // ascending index inside the thread and strict '>' keep the first maximum
        const int base = i << 2;
        if (q.x > best) { best = q.x; at = base; }
        if (q.y > best) { best = q.y; at = base + 1; }
        if (q.z > best) { best = q.z; at = base + 2; }
        if (q.w > best) { best = q.w; at = base + 3; }
    };
    if (NITER > 0) {
        float4 v[NITER > 0 ? NITER : 1];
#pragma unroll
        for (int it = 0; it < NITER; ++it) v[it] = ldg_stream(src + it * blockDim.x + threadIdx.x);
#pragma unroll
        for (int it = 0; it < NITER; ++it) visit(v[it], it * blockDim.x + threadIdx.x);
    } else {
        for (int i = threadIdx.x; i < n4; i += blockDim.x) visit(ldg_stream(src + i), i);
    }
    // NaN-free tiles always find a finite maximum; an all -inf/NaN tile reports index 0 like torch.max on -inf
    if (at == 0x7fffffff) { at = 0x7ffffffe; }
    block_argmax(best, at, scratch);
    if (threadIdx.x == 0) {
        if (at >= n) { at = 0; best = t[0]; }
        float fx, fy;
        subpixel_step(t, at, H, W, mode, fx, fy);
        coords[2 * tile] = fx; coords[2 * tile + 1] = fy;
        maxvals[tile] = best;
        if (index) index[tile] = at;
    }
}

// One WARP per tile for tiles of whole 512-byte rows of lanes (n4 % 32 == 0: every shape of the reference's configs): lane l
// visits the float4 32 j + l in ascending j, eight 128-bit loads in flight at a time; strict '>' inside the lane and the
// smaller index on equal values across lanes keep torch.max's first maximum.  No shared memory, no barrier: ~350
// instructions per 12 KB tile against ~900 for the CTA-per-tile kernel above, and a tile's tail (the sub-pixel taps, through
// L2) stalls one warp, not a CTA.
constexpr int kArgmaxWarpThreads = 256;
__global__ void __launch_bounds__(kArgmaxWarpThreads)
argmax_warp_kernel(const float* __restrict__ hm, int tiles, int H, int W, int mode,
                   float* __restrict__ coords, float* __restrict__ maxvals, int32_t* __restrict__ index) {
    const int lane = threadIdx.x & 31;
    const int tile = blockIdx.x * (kArgmaxWarpThreads / 32) + (threadIdx.x >> 5);
    if (tile >= tiles) return;
    const int n = H * W, nj = n >> 7;
    const float* t = hm + (size_t)tile * n;
    const float4* src = reinterpret_cast<const float4*>(t) + lane;
    float best = -INFINITY;
    int at = 0x7ffffffe;
    auto visit = [&](const float4& q, int i) {
        const int base = i << 2;
        if (q.x > best) { best = q.x; at = base; }
        if (q.y > best) { best = q.y; at = base + 1; }
        if (q.z > best) { best = q.z; at = base + 2; }
        if (q.w > best) { best = q.w; at = base + 3; }
    };
    int j = 0;
    for (; j + 8 <= nj; j += 8) {
        float4 v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = ldg_stream(src + 32 * (j + q));
#pragma unroll
        for (int q = 0; q < 8; ++q) visit(v[q], 32 * (j + q) + lane);
    }
    for (; j < nj; ++j) visit(ldg_stream(src + 32 * j), 32 * j + lane);
    // NaN-free tiles always find a finite maximum; an all -inf/NaN tile reports index 0 like torch.max on -inf
    warp_argmax(best, at);
    if (lane == 0) {
        if (at >= n) { at = 0; best = t[0]; }
        float fx, fy;
        subpixel_step(t, at, H, W, mode, fx, fy);
        coords[2 * tile] = fx; coords[2 * tile + 1] = fy;
        maxvals[tile] = best;
        if (index) index[tile] = at;
    }
}

int launch_argmax(const float* hm, int B, int K, int H, int W, int mode,
                  float* coords, float* maxvals, int32_t* index, cudaStream_t s) {
    // GBCODEC_ARGMAX_KERNEL=tile: the CTA-per-tile kernel for every shape (A/B measurements)
    const char* const sel = getenv("GBCODEC_ARGMAX_KERNEL");                 // read on every call
    const bool cta_tile = sel && !strcmp(sel, "tile");
    if (!cta_tile && (H * W) % 128 == 0 && (reinterpret_cast<uintptr_t>(hm) & 15u) == 0) {
        const int per_cta = kArgmaxWarpThreads / 32, tiles = B * K;
        note_launch(), argmax_warp_kernel<<<(tiles + per_cta - 1) / per_cta, kArgmaxWarpThreads, 0, s>>>(hm, tiles, H, W, mode, coords, maxvals, index);
        return check_launch("argmax_warp_kernel");
    }
    int niter = 0;
    int threads = pick_threads((H * W) >> 2, &niter);
    if (!threads) threads = 256;
    const int grid = B * K;
#define GBC_CASE(NI) case NI: note_launch(), argmax_kernel<NI><<<grid, threads, 0, s>>>(hm, H, W, mode, coords, maxvals, index); break;
    switch (niter) {
        GBC_CASE(1) GBC_CASE(2) GBC_CASE(3) GBC_CASE(4) GBC_CASE(5) GBC_CASE(6) GBC_CASE(7) GBC_CASE(8)
        default: note_launch(), argmax_kernel<0><<<grid, threads, 0, s>>>(hm, H, W, mode, coords, maxvals, index); break;
    }
#undef GBC_CASE
    return check_launch("argmax_kernel");
}

// ---------------------------------------------------------------------------------
// coordinate_refinement (utils/postprocess.py:138-184): one warp per tile
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
centroid_kernel(const float* __restrict__ hm, const float* __restrict__ cin, int tiles, int H, int W, int window,
                float* __restrict__ cout) {
    const int lane = threadIdx.x & 31;
    const int tile = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (tile >= tiles) return;
    const float* t = hm + (size_t)tile * H * W;
    float ox, oy;
    centroid_window(t, cin[2 * tile], cin[2 * tile + 1], H, W, window, ox, oy);
    if (lane == 0) { cout[2 * tile] = ox; cout[2 * tile + 1] = oy; }
}

int launch_centroid(const float* hm, const float* cin, int B, int K, int H, int W, int window, float* cout, cudaStream_t s) {
    const int tiles = B * K;
    note_launch(), centroid_kernel<<<(tiles + 3) / 4, 128, 0, s>>>(hm, cin, tiles, H, W, window, cout);
    return check_launch("centroid_kernel");
}

// ---------------------------------------------------------------------------------
// postprocess_predictions (utils/postprocess.py:296-340) in one pass: fused_decode ->
// coordinate_refinement -> filter_low_confidence -> transform_preds.  One CTA per tile;
// the arg-max reads the tile once from HBM (4*H*W bytes), everything after it is a few
// dependent L1/L2 reads by warp 0.  Scalar arithmetic follows ATen's order of float32
// operations (no contraction), so coordinates in image space do not pick up FMA noise.
// ---------------------------------------------------------------------------------
struct PostParams {
    int K, H, W, mode, scale_to_image, window, filter, transform;
    float image_size, threshold, input_w, input_h;
};

// any element of the regression branch > 1.0 (or NaN): the reference then leaves it un-scaled (:119)
__global__ void __launch_bounds__(256)
regression_range_kernel(const float* __restrict__ reg, size_t n, unsigned* __restrict__ flag) {
    bool big = false;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) big |= !(reg[i] <= 1.0f);
    if (__syncthreads_or(big) && threadIdx.x == 0) atomicOr(flag, 1u);
}

template <int NITER>
__global__ void __launch_bounds__(1024)
postprocess_kernel(const float* __restrict__ hm, const __grid_constant__ PostParams P, const float* __restrict__ reg,
                   const unsigned* __restrict__ reg_flag, const float* __restrict__ center, const float* __restrict__ scale,
                   float* __restrict__ preds, float* __restrict__ maxvals, float* __restrict__ mask) {
    __shared__ float scratch[64];
    const int tile = blockIdx.x;
    const int H = P.H, W = P.W, n = H * W, n4 = n >> 2;
    const float* t = hm + (size_t)tile * n;
    const float4* src = reinterpret_cast<const float4*>(t);
    float best = -INFINITY;
    int at = 0x7fffffff;
    auto visit = [&](const float4& q, int i) {
        const int base = i << 2;
        if (q.x > best) { best = q.x; at = base; }
        if (q.y > best) { best = q.y; at = base + 1; }
        if (q.z > best) { best = q.z; at = base + 2; }
        if (q.w > best) { best = q.w; at = base + 3; }
    };
    if (NITER > 0) {
        float4 v[NITER > 0 ? NITER : 1];
#pragma unroll
        for (int it = 0; it < NITER; ++it) v[it] = ldg_stream(src + it * blockDim.x + threadIdx.x);
#pragma unroll
        for (int it = 0; it < NITER; ++it) visit(v[it], it * blockDim.x + threadIdx.x);
    } else {
        for (int i = threadIdx.x; i < n4; i += blockDim.x) visit(ldg_stream(src + i), i);
    }
    if (at == 0x7fffffff) { at = 0x7ffffffe; }
    block_argmax(best, at, scratch);
    if (threadIdx.x >= 32) return;
    if (at >= n) { at = 0; best = t[0]; }
    const int b = tile / P.K;
    float fx, fy;
    subpixel_step(t, at, H, W, P.mode, fx, fy);
    // fused_decode (:105-114): heatmap pixels -> the 256-px image
    if (P.scale_to_image) {
        fx = __fmul_rn(fx, (float)((double)P.image_size / (double)W));
        fy = __fmul_rn(fy, (float)((double)P.image_size / (double)H));
    }
    // (:116-131) confidence-adaptive blend with the regression branch
    if (reg) {
        float rx = reg[2 * tile], ry = reg[2 * tile + 1];
        if (*reg_flag == 0u) { rx = __fmul_rn(rx, P.image_size); ry = __fmul_rn(ry, P.image_size); }
        const float a = __fdiv_rn(best, __fadd_rn(best, 0.1f)), na = __fsub_rn(1.f, a);
        fx = __fadd_rn(__fmul_rn(a, fx), __fmul_rn(na, rx));
        fy = __fadd_rn(__fmul_rn(a, fy), __fmul_rn(na, ry));
    }
    if (P.window > 0) {
        float ox, oy;
        centroid_window(t, fx, fy, H, W, P.window, ox, oy);
        fx = ox; fy = oy;
    }
    float m = 1.f;
    if (P.filter) {
        m = best > P.threshold ? 1.f : 0.f;
        fx = __fmul_rn(fx, m); fy = __fmul_rn(fy, m);
    }
    if (P.transform) {
        const float cx = center[2 * b], cy = center[2 * b + 1], sx = scale[2 * b], sy = scale[2 * b + 1];
        fx = __fsub_rn(__fadd_rn(__fmul_rn(fx, __fdiv_rn(sx, P.input_w)), cx), __fmul_rn(sx, 0.5f));
        fy = __fsub_rn(__fadd_rn(__fmul_rn(fy, __fdiv_rn(sy, P.input_h)), cy), __fmul_rn(sy, 0.5f));
    }
    if (threadIdx.x == 0) {
        preds[2 * tile] = fx; preds[2 * tile + 1] = fy;
        maxvals[tile] = best;
        if (mask) mask[tile] = m;
    }
}

int launch_postprocess(const gbcodec_postprocess_desc* d, const float* hm, const float* reg, const float* center,
                       const float* scale, float* preds, float* maxvals, float* mask, void* ws, cudaStream_t s) {
    PostParams P;
    P.K = d->K; P.H = d->H; P.W = d->W; P.mode = d->argmax_mode; P.scale_to_image = d->scale_to_image;
    P.window = d->refine_window; P.filter = d->filter; P.transform = d->transform;
    P.image_size = d->image_size; P.threshold = d->threshold; P.input_w = d->input_w; P.input_h = d->input_h;
    unsigned* flag = reinterpret_cast<unsigned*>(ws);
    if (reg) {
        cudaError_t e = cudaMemsetAsync(flag, 0, 4, s);
        if (e != cudaSuccess) return fail(GBCODEC_ERR_CUDA, "cudaMemsetAsync: %s", cudaGetErrorString(e));
        const size_t n = (size_t)d->B * d->K * 2;
        const int grid = (int)((n + 255) / 256 < 148 * 4 ? (n + 255) / 256 : 148 * 4);
        note_launch(), regression_range_kernel<<<grid, 256, 0, s>>>(reg, n, flag);
    }
    int niter = 0;
    int threads = pick_threads((d->H * d->W) >> 2, &niter);
    if (!threads) threads = 256;
    const int grid = d->B * d->K;
#define GBC_CASE(NI) case NI: note_launch(), postprocess_kernel<NI><<<grid, threads, 0, s>>>(hm, P, reg, flag, center, scale, preds, maxvals, mask); break;
    switch (niter) {
        GBC_CASE(1) GBC_CASE(2) GBC_CASE(3) GBC_CASE(4) GBC_CASE(5) GBC_CASE(6) GBC_CASE(7) GBC_CASE(8)
        default: note_launch(), postprocess_kernel<0><<<grid, threads, 0, s>>>(hm, P, reg, flag, center, scale, preds, maxvals, mask); break;
    }
#undef GBC_CASE
    return check_launch("postprocess_kernel");
}

// ---------------------------------------------------------------------------------
// heatmap pixels -> input pixels -> original image (validate.py:31-36,102-119;
// inference.py:143-175), the reference's order of float32 operations:
//   c *= float32(in / hm);  c = c / in * scale + center - scale / 2
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
coords_to_image_kernel(const float* __restrict__ cin, const float* __restrict__ center, const float* __restrict__ scale,
                       int tiles, int K, float kx, float ky, float in_w, float in_h, float* __restrict__ cout) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= tiles) return;
    const int b = t / K;
    const float cx = center[2 * b], cy = center[2 * b + 1], sx = scale[2 * b], sy = scale[2 * b + 1];
    const float x = __fmul_rn(cin[2 * t], kx), y = __fmul_rn(cin[2 * t + 1], ky);
    cout[2 * t] = __fsub_rn(__fadd_rn(__fmul_rn(__fdiv_rn(x, in_w), sx), cx), __fmul_rn(sx, 0.5f));
    cout[2 * t + 1] = __fsub_rn(__fadd_rn(__fmul_rn(__fdiv_rn(y, in_h), sy), cy), __fmul_rn(sy, 0.5f));
}

int launch_coords_to_image(const float* cin, const float* center, const float* scale, int B, int K, int H, int W,
                           float in_w, float in_h, float* cout, cudaStream_t s) {
    const int tiles = B * K;
    const float kx = (float)((double)in_w / (double)W), ky = (float)((double)in_h / (double)H);
    note_launch(), coords_to_image_kernel<<<(tiles + 255) / 256, 256, 0, s>>>(cin, center, scale, tiles, K, kx, ky, in_w, in_h, cout);
    return check_launch("coords_to_image_kernel");
}

}  // namespace gbc
