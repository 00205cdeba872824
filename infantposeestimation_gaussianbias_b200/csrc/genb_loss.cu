// genb_loss.cu — CombinedLoss of the reference's second codec generation (models/losses.py:205-290)
// and its parts, forward and backward in one pass per tile:
//
//   heatmap     FusedPoseLoss            models/losses.py:10-47      mean(crit(p, t) * w)
//               KeypointMSELoss          models/pose_estimator.py:102-143   mean((p w - t w)^2)
//               JointsMSELoss            models/losses.py:174-202    0.5 * the same
//   morph       MorphologyShapeLoss      models/losses.py:50-135     spatial mean / variance of p and t
//   regression  OffsetRegressionLoss     models/losses.py:138-171    (B,K,2) coordinates
//
// One CTA per (image, keypoint) tile; both tiles are read from HBM once with 128-bit streaming
// loads and stay in registers for the three passes (raw moments, central moments, gradient).
// Roofline: HBM, 8N bytes read + 4N written per tile.  Closed-form backward:
//   with S = sum(p) + 1e-8, m_c = sum(p c)/S, v_c = sum(p (c - m_c)^2)/S  (c = x, y)
//   d m_c / d p_i = (c_i - m_c)/S         d v_c / d p_i = ((c_i - m_c)^2 - v_c)/S
// (the cross term 2 (d m_c/d p_i) sum q (c - m_c) = 2 (d m_c/d p_i) m_c eps/S is below fp32 resolution).
//
// float16 predictions (autocast: the head's heatmaps arrive in half, the data loader's targets stay float32): the
// HALF instantiations read pred as 8-byte vectors of four halves, up-cast where they land in registers, and round the
// gradient to half once where it leaves — what torch's autocast does with mse_loss (inputs cast to float32, the
// gradient cast back), without the two conversion passes: 2N + 4N bytes read, 2N written per tile.
#include "common.cuh"
#include "decode_device.cuh"
#include <string.h>

namespace gbc {

struct GenbParams {
    int B, K, H, W;
    unsigned terms;
    int heat_crit, coord_crit, use_target_weight;
    float heat_scale, lam_var, lam_mean;
    float w[4];              // multipliers of heatmap, morph, regression, refined in `total`
    float inv_heat;          // 1 / (Bn K N)
    float inv_pair;          // 1 / (Bn K 2)
};

struct GenbArgs {
    const float* pred; const float* target; const float* weight;
    const float* coords; const float* refined; const float* target_coords;
    const float* grad_scale;
    float* grad_pred; float* grad_coords; float* grad_refined;
    float* partial;          // [B*K][4] un-normalised per-tile numerators
    const float* eff;        // backward: device [4] effective upstream gradient per term
    const int* plan;         // backward: run only if *plan != 0
    int half_io;             // pred and grad_pred are float16 (the pointers are then __half*); everything else float32
};

// four pixels of pred / grad_pred: 16 bytes of float, 8 bytes of half
template <bool HALF> struct PredIO;
template <> struct PredIO<false> {
    using Vec = float4;
    static __device__ __forceinline__ float4 load_stream(const Vec* p) { return ldg_stream(p); }
    static __device__ __forceinline__ float4 load_keep(const Vec* p) { return ldg_keep(p); }
    static __device__ __forceinline__ void store_stream(Vec* p, const float4& v) { stg_stream(p, v); }
};
template <> struct PredIO<true> {
    using Vec = uint2;
    static __device__ __forceinline__ float4 up(const uint2& r) {
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&r.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
        return make_float4(a.x, a.y, b.x, b.y);
    }
    static __device__ __forceinline__ float4 load_stream(const Vec* p) {
        uint2 r;
        asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
        return up(r);
    }
    static __device__ __forceinline__ float4 load_keep(const Vec* p) { return up(__ldg(p)); }
    static __device__ __forceinline__ void store_stream(Vec* p, const float4& v) {
        const __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
        asm volatile("st.global.cs.v2.u32 [%0], {%1,%2};" :: "l"(p), "r"(*reinterpret_cast<const unsigned*>(&a)), "r"(*reinterpret_cast<const unsigned*>(&b)) : "memory");
    }
};

constexpr int kGenbFinBlocks = 32;
constexpr int kGenbHeaderFloats = 512;     // ticket, plan, eff[4], second-stage partials (32 x 4 doubles)
struct GenbWs { unsigned* ticket; int* plan; float* eff; double* bpart; float* partial; };
static inline size_t genb_ws_bytes(int B, int K) { return (size_t)(kGenbHeaderFloats + (size_t)B * K * 4) * sizeof(float); }
static inline GenbWs genb_carve(void* ws) {
    float* f = reinterpret_cast<float*>(ws);
    GenbWs l;
    l.ticket = reinterpret_cast<unsigned*>(f);
    l.plan = reinterpret_cast<int*>(f + 1);
    l.eff = f + 4;
    l.bpart = reinterpret_cast<double*>(f + 64);
    l.partial = f + kGenbHeaderFloats;
    return l;
}

__device__ __forceinline__ float crit_value(int crit, float d) {
    if (crit == GBCODEC_CRIT_SMOOTHL1) { const float a = fabsf(d); return a < 1.f ? 0.5f * d * d : a - 0.5f; }
    if (crit == GBCODEC_CRIT_L1) return fabsf(d);
    return d * d;
}
__device__ __forceinline__ float crit_slope(int crit, float d) {
    if (crit == GBCODEC_CRIT_SMOOTHL1) return fabsf(d) < 1.f ? d : (d > 0.f ? 1.f : -1.f);
    if (crit == GBCODEC_CRIT_L1) return d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
    return 2.f * d;
}

// ---- per-tile heatmap + morphology terms ------------------------------------------------------
template <int NITER, int MAXT, int MINB = 1, bool HALF = false>
__global__ void __launch_bounds__(MAXT, MINB)
genb_tile_kernel(const __grid_constant__ GenbParams P, const __grid_constant__ GenbArgs A) {
    using IO = PredIO<HALF>;
    using PVec = typename IO::Vec;
    __shared__ float red_a[8 * 32], red_b[4 * 32];
    if (A.plan && *A.plan == 0) return;               // stored gradients already right
    const int tile = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int n4 = (P.H * P.W) >> 2, w4 = P.W >> 2;
    const PVec* p4 = reinterpret_cast<const PVec*>(A.pred) + (size_t)tile * n4;
    const float4* t4 = reinterpret_cast<const float4*>(A.target) + (size_t)tile * n4;
    constexpr int R = NITER > 0 ? NITER : 1;
    float4 pv[R], tv[R];
    if (NITER > 0) {
#pragma unroll
        for (int it = 0; it < R; ++it) {
            pv[it] = IO::load_stream(p4 + it * blockDim.x + threadIdx.x);
            tv[it] = ldg_stream(t4 + it * blockDim.x + threadIdx.x);
        }
    }
    // pixel coordinates of the register-resident float4s, worked out once (y << 16 | x): one division per float4
    // instead of one per float4 and pass
    int yx[R];
    if (NITER > 0) {
#pragma unroll
        for (int it = 0; it < R; ++it) {
            const int i = it * blockDim.x + threadIdx.x, y = i / w4;
            yx[it] = (y << 16) | ((i - y * w4) << 2);
        }
    }
    const float wraw = A.weight ? __ldg(A.weight + tile) : 1.f;
    const bool utw = P.use_target_weight && A.weight;
    const float wh = P.heat_crit == GBCODEC_CRIT_MSE_WEIGHTED ? (utw ? wraw * wraw : 1.f) : (utw ? wraw : 1.f);
    const float wm = wraw;
    const int crit = P.heat_crit == GBCODEC_CRIT_MSE_WEIGHTED ? GBCODEC_CRIT_MSE : P.heat_crit;
    const float gs = A.grad_scale ? __ldg(A.grad_scale) : 1.f;
    const float eh = A.eff ? __ldg(A.eff) : P.w[0] * gs;
    const float em = A.eff ? __ldg(A.eff + 1) : P.w[1] * gs;

    // ---- pass 1: raw moments of both tiles and the pixel criterion -----------------------------
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    auto pass1 = [&](const float4& p, const float4& t, int c) {
        const float fx = (float)(c & 0xffff), fy = (float)(c >> 16);
        const float sp = (p.x + p.y) + (p.z + p.w), st = (t.x + t.y) + (t.z + t.w);
        acc[0] += sp;
        acc[1] += fmaf(fx, sp, fmaf(3.f, p.w, fmaf(2.f, p.z, p.y)));
        acc[2] = fmaf(fy, sp, acc[2]);
        acc[3] += st;
        acc[4] += fmaf(fx, st, fmaf(3.f, t.w, fmaf(2.f, t.z, t.y)));
        acc[5] = fmaf(fy, st, acc[5]);
        acc[6] += (crit_value(crit, p.x - t.x) + crit_value(crit, p.y - t.y)) + (crit_value(crit, p.z - t.z) + crit_value(crit, p.w - t.w));
    };
    if (NITER > 0) {
#pragma unroll
        for (int it = 0; it < R; ++it) pass1(pv[it], tv[it], yx[it]);
    } else {
        for (int i = threadIdx.x; i < n4; i += blockDim.x) { const int y = i / w4; pass1(IO::load_stream(p4 + i), ldg_stream(t4 + i), (y << 16) | ((i - y * w4) << 2)); }
    }
    block_sum_1bar<8>(acc, red_a, nw, lane, warp);
    const float Sp = acc[0] + kEps, St = acc[3] + kEps;
    const float iSp = 1.f / Sp, iSt = 1.f / St;
    const float pmx = acc[1] * iSp, pmy = acc[2] * iSp, tmx = acc[4] * iSt, tmy = acc[5] * iSt;
    const float crit_sum = acc[6];

    // ---- pass 2: central second moments about each tile's own mean ------------------------------
    float c2[4] = {0.f, 0.f, 0.f, 0.f};
    auto pass2 = [&](const float4& p, const float4& t, int c) {
        const float fx = (float)(c & 0xffff), fy = (float)(c >> 16);
        const float pe[4] = {p.x, p.y, p.z, p.w}, te[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float dp = (fx + (float)j) - pmx, dt = (fx + (float)j) - tmx;
            c2[0] = fmaf(pe[j], dp * dp, c2[0]);
            c2[2] = fmaf(te[j], dt * dt, c2[2]);
        }
        const float dyp = fy - pmy, dyt = fy - tmy;
        c2[1] = fmaf((p.x + p.y) + (p.z + p.w), dyp * dyp, c2[1]);
        c2[3] = fmaf((t.x + t.y) + (t.z + t.w), dyt * dyt, c2[3]);
    };
    if (NITER > 0) {
#pragma unroll
        for (int it = 0; it < R; ++it) pass2(pv[it], tv[it], yx[it]);
    } else {
        for (int i = threadIdx.x; i < n4; i += blockDim.x) { const int y = i / w4; pass2(IO::load_keep(p4 + i), ldg_keep(t4 + i), (y << 16) | ((i - y * w4) << 2)); }
    }
    block_sum_1bar<4>(c2, red_b, nw, lane, warp);
    const float pvx = c2[0] * iSp, pvy = c2[1] * iSp, tvx = c2[2] * iSt, tvy = c2[3] * iSt;
    const float dvx = pvx - tvx, dvy = pvy - tvy, dmx = pmx - tmx, dmy = pmy - tmy;

    if (threadIdx.x == 0 && !A.eff) {
        float* out = A.partial + (size_t)tile * 4;
        out[0] = (P.terms & GBCODEC_TERM_HEATMAP) ? wh * crit_sum : 0.f;
        out[1] = (P.terms & GBCODEC_TERM_MORPH) ? wm * (P.lam_var * (dvx * dvx + dvy * dvy) + P.lam_mean * (dmx * dmx + dmy * dmy)) : 0.f;
    }
    if (!A.grad_pred) return;

    // ---- pass 3: d(total)/d(pred) ----------------------------------------------------------------
    const float ch = (P.terms & GBCODEC_TERM_HEATMAP) ? eh * P.heat_scale * wh * P.inv_heat : 0.f;
    const float cm = (P.terms & GBCODEC_TERM_MORPH) ? em * wm * P.inv_pair * iSp : 0.f;
    const float Ax = cm * 2.f * P.lam_var * dvx, Ay = cm * 2.f * P.lam_var * dvy;
    const float Bx = cm * 2.f * P.lam_mean * dmx, By = cm * 2.f * P.lam_mean * dmy;
    const float C0 = -(Ax * pvx + Ay * pvy);
    PVec* g4 = reinterpret_cast<PVec*>(A.grad_pred) + (size_t)tile * n4;
    auto pass3 = [&](const float4& p, const float4& t, int c, int i) {
        const int x = c & 0xffff;
        const float dy = (float)(c >> 16) - pmy;
        const float rowc = fmaf(dy, fmaf(Ay, dy, By), C0);
        const float pe[4] = {p.x, p.y, p.z, p.w}, te[4] = {t.x, t.y, t.z, t.w};
        float g[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float dx = ((float)x + (float)j) - pmx;
            g[j] = fmaf(ch, crit_slope(crit, pe[j] - te[j]), fmaf(dx, fmaf(Ax, dx, Bx), rowc));
        }
        IO::store_stream(g4 + i, make_float4(g[0], g[1], g[2], g[3]));
    };
    if (NITER > 0) {
#pragma unroll
        for (int it = 0; it < R; ++it) pass3(pv[it], tv[it], yx[it], it * blockDim.x + threadIdx.x);
    } else {
        for (int i = threadIdx.x; i < n4; i += blockDim.x) { const int y = i / w4; pass3(IO::load_keep(p4 + i), ldg_keep(t4 + i), (y << 16) | ((i - y * w4) << 2), i); }
    }
}

// ---- coordinate terms: one thread per (image, keypoint) -----------------------------------------
__global__ void __launch_bounds__(256)
genb_coords_kernel(const __grid_constant__ GenbParams P, const __grid_constant__ GenbArgs A, int zero_tile_terms) {
    if (A.plan && *A.plan == 0) return;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= P.B * P.K) return;
    const float w = A.weight ? A.weight[t] : 1.f;
    const float gs = A.grad_scale ? *A.grad_scale : 1.f;
    float num[2] = {0.f, 0.f};
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const unsigned bit = q == 0 ? GBCODEC_TERM_REGRESSION : GBCODEC_TERM_REFINED;
        const float* src = q == 0 ? A.coords : A.refined;
        float* dst = q == 0 ? A.grad_coords : A.grad_refined;
        if (!(P.terms & bit)) continue;
        const float e = A.eff ? A.eff[2 + q] : P.w[2 + q] * gs;
        const float dx = src[2 * t] - A.target_coords[2 * t], dy = src[2 * t + 1] - A.target_coords[2 * t + 1];
        num[q] = w * (crit_value(P.coord_crit, dx) + crit_value(P.coord_crit, dy));
        if (dst) {
            const float c = e * w * P.inv_pair;
            dst[2 * t] = c * crit_slope(P.coord_crit, dx);
            dst[2 * t + 1] = c * crit_slope(P.coord_crit, dy);
        }
    }
    if (!A.eff) {
        float* out = A.partial + (size_t)t * 4;
        out[2] = num[0]; out[3] = num[1];
        if (zero_tile_terms) { out[0] = 0.f; out[1] = 0.f; }
    }
}

// ---- second stage: fixed-order sums, the five scalars ----------------------------------------------
__global__ void __launch_bounds__(256)
genb_finalize_kernel(const __grid_constant__ GenbParams P, const float* __restrict__ partial, double* __restrict__ bpart,
                     unsigned* __restrict__ ticket, float* __restrict__ losses5, float* __restrict__ single_out) {
    __shared__ double red[4][8];
    __shared__ bool last;
    const int tiles = P.B * P.K;
    const int per = (tiles + gridDim.x - 1) / gridDim.x;
    const int lo = blockIdx.x * per, hi = min(tiles, lo + per);
    double acc[4] = {0, 0, 0, 0};
    for (int t = lo + threadIdx.x; t < hi; t += blockDim.x) {
        const float4 a = __ldcg(reinterpret_cast<const float4*>(partial) + t);
        acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        for (int o = 16; o > 0; o >>= 1) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], o);
        if (lane == 0) red[q][warp] = acc[q];
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double s = 0.0;
        for (int wp = 0; wp < 8; ++wp) s += red[threadIdx.x][wp];
        bpart[blockIdx.x * 4 + threadIdx.x] = s;
        __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!last) return;
    __threadfence();
    __shared__ float term[4];
    if (threadIdx.x < 4) {
        const int q = threadIdx.x;
        double s = 0.0;
        for (int g = 0; g < (int)gridDim.x; ++g) s += __ldcg(bpart + g * 4 + q);
        const float v = q == 0 ? (float)s * P.inv_heat * P.heat_scale : (float)s * P.inv_pair;
        term[q] = v;
        if (losses5) losses5[q] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float total = 0.f;
        for (int q = 0; q < 4; ++q) total += P.w[q] * term[q];
        if (losses5) losses5[4] = total;
        if (single_out) *single_out = total;
        *ticket = 0u;
    }
}

// backward plan: eff[q] = g5[4] * w[q] + g5[q]; plan = 0 if that is what the forward assumed
__global__ void genb_plan_kernel(const __grid_constant__ GenbParams P, const float* __restrict__ g5,
                                 const float* __restrict__ assumed, int* __restrict__ plan, float* __restrict__ eff) {
    if (threadIdx.x != 0) return;
    const float a = assumed ? *assumed : 1.f;
    bool same = true;
    for (int q = 0; q < 4; ++q) {
        eff[q] = fmaf(g5[4], P.w[q], g5[q]);
        if (eff[q] != P.w[q] * a) same = false;
    }
    *plan = same ? 0 : 1;
}

// ---- host side ------------------------------------------------------------------------------------
static int make_genb_params(const gbcodec_combined_desc* d, GenbParams* P) {
    if (!d) return fail(GBCODEC_ERR_NULL_POINTER, "combined_loss: desc is NULL");
    if (d->B <= 0 || d->K <= 0) return fail(GBCODEC_ERR_BAD_SHAPE, "combined_loss: B,K must be positive");
    const bool tiles = d->terms & (GBCODEC_TERM_HEATMAP | GBCODEC_TERM_MORPH);
    if (tiles) {
        if (d->H <= 0 || d->W <= 0 || d->W % 4) return fail(GBCODEC_ERR_BAD_SHAPE, "combined_loss: H,W must be positive and W a multiple of 4 (got %d,%d)", d->H, d->W);
        if ((long long)d->H * d->W > GBCODEC_MAX_TILE) return fail(GBCODEC_ERR_BAD_SHAPE, "combined_loss: tile %dx%d too large", d->H, d->W);
    }
    if (!d->terms || (d->terms & ~15u)) return fail(GBCODEC_ERR_BAD_ARGUMENT, "combined_loss: terms=0x%x", d->terms);
    if (d->heatmap_criterion != GBCODEC_CRIT_MSE && d->heatmap_criterion != GBCODEC_CRIT_SMOOTHL1 && d->heatmap_criterion != GBCODEC_CRIT_MSE_WEIGHTED)
        return fail(GBCODEC_ERR_BAD_ARGUMENT, "combined_loss: heatmap_criterion=%d", d->heatmap_criterion);
    if (d->coord_criterion != GBCODEC_CRIT_MSE && d->coord_criterion != GBCODEC_CRIT_SMOOTHL1 && d->coord_criterion != GBCODEC_CRIT_L1)
        return fail(GBCODEC_ERR_BAD_ARGUMENT, "combined_loss: coord_criterion=%d", d->coord_criterion);
    if (d->norm_batch < 0) return fail(GBCODEC_ERR_BAD_ARGUMENT, "combined_loss: norm_batch=%d", d->norm_batch);
    memset(P, 0, sizeof(*P));
    P->B = d->B; P->K = d->K; P->H = d->H; P->W = d->W;
    P->terms = d->terms; P->heat_crit = d->heatmap_criterion; P->coord_crit = d->coord_criterion;
    P->use_target_weight = d->use_target_weight;
    P->heat_scale = d->heatmap_scale; P->lam_var = d->lambda_variance; P->lam_mean = d->lambda_mean;
    P->w[0] = d->w_heatmap; P->w[1] = d->w_morph; P->w[2] = d->w_reg; P->w[3] = d->w_reg;
    const double bn = (double)(d->norm_batch ? d->norm_batch : d->B) * d->K;
    P->inv_heat = tiles ? (float)(1.0 / (bn * d->H * d->W)) : 0.f;
    P->inv_pair = (float)(1.0 / (bn * 2.0));
    return GBCODEC_OK;
}

static int check_genb(const GenbParams& P, const GenbArgs& A, const void* ws, size_t ws_size, bool backward) {
    const bool tiles = P.terms & (GBCODEC_TERM_HEATMAP | GBCODEC_TERM_MORPH);
    const bool coords = P.terms & (GBCODEC_TERM_REGRESSION | GBCODEC_TERM_REFINED);
    if (tiles && (!A.pred || !A.target)) return fail(GBCODEC_ERR_NULL_POINTER, "combined_loss: d_pred / d_target is NULL");
    if ((P.terms & GBCODEC_TERM_REGRESSION) && !A.coords) return fail(GBCODEC_ERR_NULL_POINTER, "combined_loss: d_coords is NULL");
    if ((P.terms & GBCODEC_TERM_REFINED) && !A.refined) return fail(GBCODEC_ERR_NULL_POINTER, "combined_loss: d_refined is NULL");
    if (coords && !A.target_coords) return fail(GBCODEC_ERR_NULL_POINTER, "combined_loss: d_target_coords is NULL");
    if (!ws || ws_size < genb_ws_bytes(P.B, P.K)) return fail(GBCODEC_ERR_WORKSPACE, "combined_loss: workspace of %zu bytes needed", genb_ws_bytes(P.B, P.K));
    // four pixels per access: 16 bytes of float32, 8 bytes of float16 (a slice of whole images of a float16 batch need
    // not start on a 16-byte boundary)
    const auto vec_ok = [&](const void* p) { return A.half_io ? (reinterpret_cast<uintptr_t>(p) & 7u) == 0 : aligned16(p); };
    if (!aligned16(ws) || (tiles && (!vec_ok(A.pred) || !aligned16(A.target))) || (A.grad_pred && !vec_ok(A.grad_pred)))
        return fail(GBCODEC_ERR_UNALIGNED, "combined_loss: tensors must be 16-byte aligned (float16 predictions and their gradient: 8-byte)");
    if (A.half_io && !tiles) return fail(GBCODEC_ERR_BAD_ARGUMENT, "combined_loss (float16): the call has no heatmap term; use the float32 entry point");
    if (backward && tiles && !A.grad_pred) return fail(GBCODEC_ERR_NULL_POINTER, "combined_loss backward: d_grad_pred is NULL");
    return GBCODEC_OK;
}

static int launch_genb(const GenbParams& P, const GenbArgs& A, cudaStream_t s) {
    const bool tiles = P.terms & (GBCODEC_TERM_HEATMAP | GBCODEC_TERM_MORPH);
    const int nt = P.B * P.K;
    note_launch(), genb_coords_kernel<<<(nt + 255) / 256, 256, 0, s>>>(P, A, tiles ? 0 : 1);
    int st = check_launch("genb_coords_kernel");
    if (st || !tiles) return st;
    // both tiles stay in registers: up to 3 float4 each per thread with CTAs of up to 1024 threads (64 registers:
    // 64x48 -> 256 x 3, 96x72 -> 576 x 3), up to 8 with CTAs of up to 512 threads (128 registers: 128x128 -> 512 x 8);
    // other shapes re-read through L2
    const int n4 = (P.H * P.W) >> 2;
    int niter = 0, threads = 512;
    for (int t = 256; t <= 1024 && !niter; t += 32)
        if (n4 % t == 0 && n4 / t <= 3) { threads = t; niter = n4 / t; }
    for (int t = 128; t <= 512 && !niter; t += 32)
        if (n4 % t == 0 && n4 / t <= 8) { threads = t; niter = n4 / t; }
#define GBC_CASE(NI, MT) case NI: note_launch(); if (A.half_io) genb_tile_kernel<NI, MT, 1, true><<<nt, threads, 0, s>>>(P, A); \
                                                  else genb_tile_kernel<NI, MT><<<nt, threads, 0, s>>>(P, A); break;
    switch (niter) {
        GBC_CASE(1, 1024) GBC_CASE(2, 1024) GBC_CASE(3, 1024) GBC_CASE(4, 512) GBC_CASE(5, 512) GBC_CASE(6, 512) GBC_CASE(7, 512) GBC_CASE(8, 512)
        default: GBC_CASE(0, 1024)
    }
#undef GBC_CASE
    return check_launch("genb_tile_kernel");
}

size_t combined_workspace_bytes(int B, int K) { return genb_ws_bytes(B, K); }

int combined_loss(const gbcodec_combined_desc* d, const float* pred, const float* target, const float* weight,
                  const float* coords, const float* refined, const float* target_coords, const float* grad_scale,
                  float* losses5, float* gpred, float* gcoords, float* grefined, void* ws, size_t ws_size, cudaStream_t s,
                  int half_io) {
    GenbParams P;
    int st = make_genb_params(d, &P);
    if (st) return st;
    if (!losses5) return fail(GBCODEC_ERR_NULL_POINTER, "combined_loss: d_losses5 is NULL");
    GenbArgs A;
    memset(&A, 0, sizeof(A));
    A.pred = pred; A.target = target; A.weight = weight; A.coords = coords; A.refined = refined; A.target_coords = target_coords;
    A.grad_scale = grad_scale; A.grad_pred = gpred; A.grad_coords = gcoords; A.grad_refined = grefined;
    A.half_io = half_io;
    st = check_genb(P, A, ws, ws_size, false);
    if (st) return st;
    const GenbWs L = genb_carve(ws);
    A.partial = L.partial;
    cudaError_t e = cudaMemsetAsync(L.ticket, 0, 8, s);
    if (e != cudaSuccess) return fail(GBCODEC_ERR_CUDA, "cudaMemsetAsync: %s", cudaGetErrorString(e));
    st = launch_genb(P, A, s);
    if (st) return st;
    const int nt = P.B * P.K;
    const int fin = (nt + 255) / 256 < kGenbFinBlocks ? (nt + 255) / 256 : kGenbFinBlocks;
    note_launch(), genb_finalize_kernel<<<fin, 256, 0, s>>>(P, L.partial, L.bpart, L.ticket, losses5, nullptr);
    return check_launch("genb_finalize_kernel");
}

int combined_loss_backward(const gbcodec_combined_desc* d, const float* pred, const float* target, const float* weight,
                           const float* coords, const float* refined, const float* target_coords, const float* grad_scale,
                           const float* g5, float* gpred, float* gcoords, float* grefined, void* ws, size_t ws_size, cudaStream_t s,
                           int half_io) {
    GenbParams P;
    int st = make_genb_params(d, &P);
    if (st) return st;
    if (!g5) return fail(GBCODEC_ERR_NULL_POINTER, "combined_loss backward: d_grad_losses5 is NULL");
    GenbArgs A;
    memset(&A, 0, sizeof(A));
    A.pred = pred; A.target = target; A.weight = weight; A.coords = coords; A.refined = refined; A.target_coords = target_coords;
    A.grad_pred = gpred; A.grad_coords = gcoords; A.grad_refined = grefined;
    A.half_io = half_io;
    st = check_genb(P, A, ws, ws_size, true);
    if (st) return st;
    const GenbWs L = genb_carve(ws);
    A.partial = L.partial; A.eff = L.eff; A.plan = L.plan;
    note_launch(), genb_plan_kernel<<<1, 32, 0, s>>>(P, g5, grad_scale, L.plan, L.eff);
    st = check_launch("genb_plan_kernel");
    if (st) return st;
    return launch_genb(P, A, s);
}

// ---- plain heatmap head: KeypointMSELoss fwd + bwd, on-the-fly targets, arg-max decode — one pass ---------
// models/pose_estimator.py head_type='heatmap': loss = mean((p w - t w)^2) (:102-143), keypoints from
// decode_heatmaps (:331-373); target tiles as COCOPoseDataset._generate_target builds them
// (datasets/coco_dataset.py:185-250).  One CTA per tile, the tile in registers: read P (4N), write dP (4N).
struct HeatArgs {
    const float* hm; const float* target; const float* weight; const float* kps; const float* grad_scale;
    float* grad_hm; float* coords; float* maxvals; int32_t* index; float* partial;
    int H, W, mode, use_target_weight;
    float in_w, in_h, inv_bkn;
    EncodeConst ec;
};

template <int NITER, int MAXT, int MINB = 1>
__global__ void __launch_bounds__(MAXT, MINB)
heatmap_step_kernel(const __grid_constant__ HeatArgs A) {
    extern __shared__ float lut[];
    __shared__ PatchGeom geom_s;
    __shared__ float red_v[32], red_s[32];
    __shared__ int red_i[32];
    const int tile = blockIdx.x;
    const int H = A.H, W = A.W, n = H * W, n4 = n >> 2, w4 = W >> 2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const float4* p4 = reinterpret_cast<const float4*>(A.hm) + (size_t)tile * n4;
    constexpr int R = NITER > 0 ? NITER : 1;
    float4 pv[R];
    if (NITER > 0) {
#pragma unroll
        for (int it = 0; it < R; ++it) pv[it] = ldg_stream(p4 + it * blockDim.x + threadIdx.x);
    }
    if (!A.target) {
        fill_patch_lut(lut, A.ec);
        if (threadIdx.x == 0) geom_s = patch_geometry(A.kps[2 * tile], A.kps[2 * tile + 1], A.weight[tile], H, W, A.in_w, A.in_h, A.ec);
        __syncthreads();
    }
    const PatchGeom g = A.target ? PatchGeom{} : geom_s;
    const float wraw = A.target ? (A.weight ? __ldg(A.weight + tile) : 1.f) : g.weight;
    const float w2 = (A.use_target_weight && A.weight) ? wraw * wraw : 1.f;      // (p w - t w)^2 = w^2 (p - t)^2
    const float4* t4 = A.target ? reinterpret_cast<const float4*>(A.target) + (size_t)tile * n4 : nullptr;
    auto target_at = [&](int i) -> float4 {
        if (t4) return ldg_keep(t4 + i);
        const int y = i / w4, x = (i - y * w4) << 2;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (g.active && y >= g.y_from && y < g.y_to && x + 3 >= g.x_from && x < g.x_to) {
            float e[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) e[j] = (x + j >= g.x_from && x + j < g.x_to) ? patch_value(lut, g, A.ec, x + j, y) : 0.f;
            v = make_float4(e[0], e[1], e[2], e[3]);
        }
        return v;
    };
    // ---- one pass: squared error and the first maximum (ascending index per thread + strict '>') -------------
    float sq = 0.f, best = -INFINITY;
    int at = 0x7fffffff;
    float4 tv[R];
    auto visit = [&](const float4& p, const float4& t, int i) {
        const float d0 = p.x - t.x, d1 = p.y - t.y, d2 = p.z - t.z, d3 = p.w - t.w;
        sq += fmaf(d0, d0, d1 * d1) + fmaf(d2, d2, d3 * d3);
        const int base = i << 2;
        if (p.x > best) { best = p.x; at = base; }
        if (p.y > best) { best = p.y; at = base + 1; }
        if (p.z > best) { best = p.z; at = base + 2; }
        if (p.w > best) { best = p.w; at = base + 3; }
    };
    if (NITER > 0) {
#pragma unroll
        for (int it = 0; it < R; ++it) { tv[it] = target_at(it * blockDim.x + threadIdx.x); visit(pv[it], tv[it], it * blockDim.x + threadIdx.x); }
    } else {
        for (int i = threadIdx.x; i < n4; i += blockDim.x) visit(ldg_stream(p4 + i), target_at(i), i);
    }
    if (at == 0x7fffffff) at = 0x7ffffffe;
    sq = warp_sum(sq);
    warp_argmax(best, at);
    if (lane == 0) { red_s[warp] = sq; red_v[warp] = best; red_i[warp] = at; }
    // ---- gradient: does not wait for the reductions --------------------------------------------------------------
    if (A.grad_hm) {
        const float gs = A.grad_scale ? __ldg(A.grad_scale) : 1.f;
        const float c = gs * 2.f * w2 * A.inv_bkn;
        float4* g4 = reinterpret_cast<float4*>(A.grad_hm) + (size_t)tile * n4;
        auto grad = [&](const float4& p, const float4& t, int i) {
            stg_stream(g4 + i, make_float4(c * (p.x - t.x), c * (p.y - t.y), c * (p.z - t.z), c * (p.w - t.w)));
        };
        if (NITER > 0) {
#pragma unroll
            for (int it = 0; it < R; ++it) grad(pv[it], tv[it], it * blockDim.x + threadIdx.x);
        } else {
            for (int i = threadIdx.x; i < n4; i += blockDim.x) grad(ldg_keep(p4 + i), target_at(i), i);
        }
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    float s = 0.f, bv = red_v[0];
    int bi = red_i[0];
    for (int q = 0; q < nw; ++q) { s += red_s[q]; if (q) argmax_merge(bv, bi, red_v[q], red_i[q]); }
    float4* out = reinterpret_cast<float4*>(A.partial) + tile;
    *out = make_float4(w2 * s, 0.f, 0.f, 0.f);
    if (A.coords) {
        const float* t = A.hm + (size_t)tile * n;
        if (bi >= n) { bi = 0; bv = t[0]; }
        float fx, fy;
        subpixel_step(t, bi, H, W, A.mode, fx, fy);
        A.coords[2 * tile] = fx; A.coords[2 * tile + 1] = fy;
        A.maxvals[tile] = bv;
        if (A.index) A.index[tile] = bi;
    }
}

int heatmap_step(const float* hm, const float* target, const float* weight, const float* kps, int B, int K, int H, int W,
                 float in_w, float in_h, double sigma, int use_target_weight, int norm_batch, const float* grad_scale,
                 float* loss, float* grad_hm, int argmax_mode, float* coords, float* maxvals, int32_t* index,
                 void* ws, size_t ws_size, cudaStream_t s) {
    if (!hm || !loss) return fail(GBCODEC_ERR_NULL_POINTER, "heatmap_step: d_hm / d_loss is NULL");
    if (!target && (!weight || !kps)) return fail(GBCODEC_ERR_NULL_POINTER, "heatmap_step: on-the-fly targets need d_weight (visibility) and d_gt_kps");
    if (coords && !maxvals) return fail(GBCODEC_ERR_NULL_POINTER, "heatmap_step: d_maxvals is NULL");
    if (!ws || ws_size < genb_ws_bytes(B, K)) return fail(GBCODEC_ERR_WORKSPACE, "heatmap_step: workspace of %zu bytes needed", genb_ws_bytes(B, K));
    if (!aligned16(hm) || !aligned16(ws) || (target && !aligned16(target)) || (grad_hm && !aligned16(grad_hm)))
        return fail(GBCODEC_ERR_UNALIGNED, "heatmap_step: tensors must be 16-byte aligned");
    if (norm_batch < 0 || !(sigma > 0.0)) return fail(GBCODEC_ERR_BAD_ARGUMENT, "heatmap_step: norm_batch / sigma");
    HeatArgs A;
    memset(&A, 0, sizeof(A));
    A.hm = hm; A.target = target; A.weight = weight; A.kps = kps; A.grad_scale = grad_scale; A.grad_hm = grad_hm;
    A.coords = coords; A.maxvals = maxvals; A.index = index;
    A.H = H; A.W = W; A.mode = argmax_mode; A.use_target_weight = use_target_weight; A.in_w = in_w; A.in_h = in_h;
    A.ec = make_encode_const(sigma);
    const double bkn = (double)(norm_batch ? norm_batch : B) * K * H * W;
    A.inv_bkn = (float)(1.0 / bkn);
    const GenbWs L = genb_carve(ws);
    A.partial = L.partial;
    cudaError_t e = cudaMemsetAsync(L.ticket, 0, 8, s);
    if (e != cudaSuccess) return fail(GBCODEC_ERR_CUDA, "cudaMemsetAsync: %s", cudaGetErrorString(e));
    const size_t smem = target ? 0 : (size_t)A.ec.lut_size * sizeof(float);
    if (smem > 40 * 1024) return fail(GBCODEC_ERR_BAD_ARGUMENT, "heatmap_step: sigma %g needs a %zu-byte patch table", sigma, smem);
    const int nt = B * K, n4 = (H * W) >> 2;
    int niter = 0, threads = 512;
    for (int t = 256; t <= 1024 && !niter; t += 32)
        if (n4 % t == 0 && n4 / t <= 3) { threads = t; niter = n4 / t; }
    for (int t = 128; t <= 512 && !niter; t += 32)
        if (n4 % t == 0 && n4 / t <= 8) { threads = t; niter = n4 / t; }
    if (niter == 3 && threads == 256) {                     // 64x48: cap the registers for 8 resident CTAs
        note_launch(), heatmap_step_kernel<3, 256, 8><<<nt, threads, smem, s>>>(A);
        niter = -1;
    }
#define GBC_CASE(NI, MT) case NI: note_launch(), heatmap_step_kernel<NI, MT><<<nt, threads, smem, s>>>(A); break;
    switch (niter) {
        case -1: break;
        GBC_CASE(1, 1024) GBC_CASE(2, 1024) GBC_CASE(3, 1024) GBC_CASE(4, 512) GBC_CASE(5, 512) GBC_CASE(6, 512) GBC_CASE(7, 512) GBC_CASE(8, 512)
        default: note_launch(), heatmap_step_kernel<0, 1024><<<nt, threads, smem, s>>>(A); break;
    }
#undef GBC_CASE
    int st = check_launch("heatmap_step_kernel");
    if (st) return st;
    GenbParams P;
    memset(&P, 0, sizeof(P));
    P.B = B; P.K = K; P.heat_scale = 1.f; P.inv_heat = A.inv_bkn; P.w[0] = 1.f;
    const int fin = (nt + 255) / 256 < kGenbFinBlocks ? (nt + 255) / 256 : kGenbFinBlocks;
    note_launch(), genb_finalize_kernel<<<fin, 256, 0, s>>>(P, L.partial, L.bpart, L.ticket, nullptr, loss);
    return check_launch("genb_finalize_kernel");
}

}  // namespace gbc
