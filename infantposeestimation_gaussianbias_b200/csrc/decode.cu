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

namespace gbc {

template <int NITER, bool FLIP>
__global__ void __launch_bounds__(1024)
decode_kernel(const float* __restrict__ hm, const float* __restrict__ hmf, const int32_t* __restrict__ perm,
              const float* __restrict__ off, const float* __restrict__ alpha_param,
              const float* __restrict__ fusion_weight, int K, int H, int W, int radius, unsigned flags,
              float* __restrict__ coords, float* __restrict__ scores, int32_t* __restrict__ centre) {
    __shared__ float scratch[4 * 32 + 8];
    const int tile = blockIdx.x;
    const int b = tile / K, k = tile - b * K;
    const int n = H * W, n4 = n >> 2, w4 = W >> 2;
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
        for (int it = 0; it < R; ++it) v[it] = load(it * blockDim.x + threadIdx.x);
#pragma unroll
        for (int it = 0; it < R; ++it) m = fmaxf(m, fmaxf(fmaxf(v[it].x, v[it].y), fmaxf(v[it].z, v[it].w)));
    } else {
        for (int i = threadIdx.x; i < n4; i += blockDim.x) {
            const float4 t = load(i);
            m = fmaxf(m, fmaxf(fmaxf(t.x, t.y), fmaxf(t.z, t.w)));
        }
    }
    m = block_max(m, scratch);

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
        for (int it = 0; it < R; ++it) accumulate(v[it], it * blockDim.x + threadIdx.x);
    } else {
        for (int i = threadIdx.x; i < n4; i += blockDim.x) accumulate(load(i), i);
    }
    block_sum<3>(acc, scratch);

    if (threadIdx.x < 32) {
        float cx = acc[1] / acc[0], cy = acc[2] / acc[0];
        int px, py;
        refine_and_correct(hm_tile, hmf_tile, off ? off + (size_t)tile * 2 * n : nullptr, alpha_param, fusion_weight,
                           H, W, radius, flags, cx, cy, px, py);
        if (threadIdx.x == 0) {
            coords[2 * tile] = cx; coords[2 * tile + 1] = cy;
            scores[tile] = m;
            if (centre) { centre[2 * tile] = px; centre[2 * tile + 1] = py; }
        }
    }
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
#define GBC_CASE(NI) case NI: decode_kernel<NI, FLIP><<<grid, threads, 0, s>>>(hm, hmf, perm, off, ap, fw, K, H, W, radius, flags, coords, scores, centre); break;
    switch (niter) {
        GBC_CASE(1) GBC_CASE(2) GBC_CASE(3) GBC_CASE(4) GBC_CASE(5) GBC_CASE(6) GBC_CASE(7) GBC_CASE(8)
        default: decode_kernel<0, FLIP><<<grid, threads, 0, s>>>(hm, hmf, perm, off, ap, fw, K, H, W, radius, flags, coords, scores, centre); break;
    }
#undef GBC_CASE
}

int launch_decode(const float* hm, const float* hmf, const int32_t* perm, const float* off,
                  const float* alpha_param, const float* fusion_weight, int B, int K, int H, int W,
                  int radius, unsigned flags, float* coords, float* scores, int32_t* centre, cudaStream_t stream) {
    int niter = 0;
    int threads = pick_threads((H * W) >> 2, &niter);
    if (!threads) threads = 256;
    if (hmf) launch_decode_t<true>(niter, B * K, threads, stream, hm, hmf, perm, off, alpha_param, fusion_weight, K, H, W, radius, flags, coords, scores, centre);
    else     launch_decode_t<false>(niter, B * K, threads, stream, hm, hmf, perm, off, alpha_param, fusion_weight, K, H, W, radius, flags, coords, scores, centre);
    return check_launch("decode_kernel");
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

int launch_argmax(const float* hm, int B, int K, int H, int W, int mode,
                  float* coords, float* maxvals, int32_t* index, cudaStream_t s) {
    int niter = 0;
    int threads = pick_threads((H * W) >> 2, &niter);
    if (!threads) threads = 256;
    const int grid = B * K;
#define GBC_CASE(NI) case NI: argmax_kernel<NI><<<grid, threads, 0, s>>>(hm, H, W, mode, coords, maxvals, index); break;
    switch (niter) {
        GBC_CASE(1) GBC_CASE(2) GBC_CASE(3) GBC_CASE(4) GBC_CASE(5) GBC_CASE(6) GBC_CASE(7) GBC_CASE(8)
        default: argmax_kernel<0><<<grid, threads, 0, s>>>(hm, H, W, mode, coords, maxvals, index); break;
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
    centroid_kernel<<<(tiles + 3) / 4, 128, 0, s>>>(hm, cin, tiles, H, W, window, cout);
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
        regression_range_kernel<<<grid, 256, 0, s>>>(reg, n, flag);
    }
    int niter = 0;
    int threads = pick_threads((d->H * d->W) >> 2, &niter);
    if (!threads) threads = 256;
    const int grid = d->B * d->K;
#define GBC_CASE(NI) case NI: postprocess_kernel<NI><<<grid, threads, 0, s>>>(hm, P, reg, flag, center, scale, preds, maxvals, mask); break;
    switch (niter) {
        GBC_CASE(1) GBC_CASE(2) GBC_CASE(3) GBC_CASE(4) GBC_CASE(5) GBC_CASE(6) GBC_CASE(7) GBC_CASE(8)
        default: postprocess_kernel<0><<<grid, threads, 0, s>>>(hm, P, reg, flag, center, scale, preds, maxvals, mask); break;
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
    coords_to_image_kernel<<<(tiles + 255) / 256, 256, 0, s>>>(cin, center, scale, tiles, K, kx, ky, in_w, in_h, cout);
    return check_launch("coords_to_image_kernel");
}

}  // namespace gbc
