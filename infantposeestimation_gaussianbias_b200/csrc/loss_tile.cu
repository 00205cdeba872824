// loss_tile.cu — register-resident tile kernel of the six-term fusion loss (forward + backward
// + optional keypoint decode in one pass).  Same arithmetic as the generic kernel in loss.cu
// (FusionPoseLoss.forward, models/fusion_head.py:745-806, terms :637-743 and :405-559, and the
// autograd backward of train.py:182 in closed form); different schedule:
//
//   * one CTA per (image, keypoint) tile, TPB = (W/4) * ROWS threads; a thread owns the same four
//     columns in every row it visits, so every x-dependent factor is a per-thread constant and
//     the column moments factor out of the row loop;
//   * the tile lives in REGISTERS (NIT float4 per thread, loaded once with 128-bit streaming
//     loads); the per-pixel intermediates a later pass needs (softmax weight, sigmoid, entropy
//     derivative) are parked in thread-private shared-memory slots — no cross-thread traffic,
//     conflict-free 128-bit accesses;
//   * four block reductions per tile, ONE barrier each (halving butterfly inside the warp,
//     then every warp finishes the cross-warp sum redundantly); the per-tile scalars are
//     computed by every thread, so nothing waits for "thread 0";
//   * all limb partners of the tile are visited once (their tiles are some other CTA's own
//     tile: L2 hits); the per-pixel tie pattern the gradient needs is kept as one bit per pixel
//     and partner in a register;
//   * stores are spread over the kernel's lifetime: the (almost all zero) offset-gradient tile
//     leaves right after the loads are issued, the uniform variance-gradient tile once the
//     first reduction is known, the heatmap gradient at the end;
//   * the decode tail (window softmax, bilinear offset read) is software-pipelined through the
//     passes in warp 0 so its dependent L2 round trips hide behind the other warps' work.
//
// Algorithmic HBM bytes per tile: read hm, var (8N); write d_hm, d_var, d_off (16N).
#include "loss_common.cuh"

namespace gbc {

// ---- one-barrier block reductions -------------------------------------------------------
// Warp stage: NV running sums per lane -> lane l holds the warp total of value l >> (5 - log2 NV)
// (halving butterfly: each step exchanges half of the remaining values, so 8 values cost
// 4+2+1+2 shuffles instead of 40).
template <int NV>
__device__ __forceinline__ void warp_scatter_sum(float (&v)[NV]) {
    const int lane = threadIdx.x & 31;
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

template <int NV> struct Log2 { static constexpr int value = 1 + Log2<NV / 2>::value; };
template <> struct Log2<1> { static constexpr int value = 0; };

// Block-wide sums of NV values; `red` holds NW*NV floats and must not be the buffer of the
// previous reduction (the callers alternate two buffers).  Fixed order: deterministic, and
// every thread ends with the same bits.
template <int NV, int NW>
__device__ __forceinline__ void block_sum1(float (&v)[NV], float* red) {
    constexpr int SH = 5 - Log2<NV>::value, SUB = 32 / NV;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    warp_scatter_sum<NV>(v);
    const int idx = lane >> SH, q = lane & (SUB - 1);
    if (q == 0) red[warp * NV + idx] = v[0];
    __syncthreads();
    float acc = 0.f;
#pragma unroll
    for (int t = 0; t < (NW + SUB - 1) / SUB; ++t) {
        const int ww = q + t * SUB;
        if (ww < NW) acc += red[ww * NV + idx];
    }
#pragma unroll
    for (int o = SUB / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = __shfl_sync(0xffffffffu, acc, k << SH);
}

template <int NW>
__device__ __forceinline__ float block_max1(float m, float* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    m = warp_max(m);
    if (lane == 0) red[warp] = m;
    __syncthreads();
    return warp_max(lane < NW ? red[lane] : -INFINITY);
}

__device__ __forceinline__ float fdiv(float a, float b) { return __fdividef(a, b); }

// ---- the kernel ---------------------------------------------------------------------------------
template <int W4, int ROWS, int NIT, bool CE, bool CS, bool CA, int MINB>
__global__ void __launch_bounds__(W4* ROWS, MINB)
loss_tile_kernel(const __grid_constant__ LossParams P, const __grid_constant__ LossArgs A) {
    constexpr int TPB = W4 * ROWS, NW = TPB / 32, N4 = TPB * NIT, N = 4 * N4, W = 4 * W4, H = ROWS * NIT;
    static_assert(TPB % 32 == 0 && TPB <= 1024, "CTA must be whole warps");
    static_assert(4 * NIT <= 32, "one tie bit per owned pixel must fit a register");
    if (A.plan && *A.plan != 2) return;      // backward recompute not needed

    extern __shared__ __align__(16) float smem[];
    float4* Es = reinterpret_cast<float4*>(smem);                 // exp(h - max), later softmax weight p
    float4* Ss = Es + (CE ? N4 : 0);                              // sigmoid(h)
    float4* As = Ss + (CS ? N4 : 0);                              // a = -log(p + eps) - p / (p + eps)
    float* lut = reinterpret_cast<float*>(As + (CA ? N4 : 0));
    float* red = lut + ((P.ec.lut_size + 3) & ~3);                // 2 buffers of NW * 8 floats
    float* red0 = red, *red1 = red + NW * 8;

    const int tid = threadIdx.x, lane = tid & 31;
    const int tx = tid % W4, ty = tid / W4;
    const int x0 = tx << 2;
    const float fx0 = (float)x0, fty = (float)ty;
    const int tile = blockIdx.x;
    const int b = tile / P.K, k = tile - b * P.K;

    const bool has_target = A.target != nullptr;
    const bool grads = A.grad_hm != nullptr;
    const bool backward_only = A.lam_eff != nullptr;
    const bool decode = A.coords != nullptr;

    const float4* hm4 = reinterpret_cast<const float4*>(A.hm) + (size_t)tile * N4 + tid;
    const float4* tgt4 = has_target ? reinterpret_cast<const float4*>(A.target) + (size_t)tile * N4 + tid : nullptr;
    float4* gh4 = grads ? reinterpret_cast<float4*>(A.grad_hm) + (size_t)tile * N4 + tid : nullptr;
    float4* gv4 = (grads && A.grad_var) ? reinterpret_cast<float4*>(A.grad_var) + (size_t)tile * N4 + tid : nullptr;
    float4* go4 = grads ? reinterpret_cast<float4*>(A.grad_off) + (size_t)tile * 2 * N4 + tid : nullptr;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);

    // ---- loads first: the tile into registers -----------------------------------------------
    float4 h[NIT];
#pragma unroll
    for (int it = 0; it < NIT; ++it) h[it] = ldg_stream(hm4 + it * TPB);
    const float w = __ldg(A.weff + tile);
    const float wa = P.use_target_weight ? w : 1.f;
    const bool heavy = (w != 0.f) || !P.use_target_weight;
    float vsum = 0.f;
    if (A.var && heavy) {
        const float4* var4 = reinterpret_cast<const float4*>(A.var) + (size_t)tile * N4 + tid;
        float4 v[NIT];
#pragma unroll
        for (int it = 0; it < NIT; ++it) v[it] = ldg_stream(var4 + it * TPB);
#pragma unroll
        for (int it = 0; it < NIT; ++it) vsum += (v[it].x + v[it].y) + (v[it].z + v[it].w);
    }
    // the offset gradient is zero except on (up to) four taps per channel, patched at the end
    if (grads) {
#pragma unroll
        for (int it = 0; it < 2 * NIT; ++it) stg_stream(go4 + it * TPB, z4);
    }
    const float D = (float)__ldg(A.sums) + kEps;
    const float D5 = (float)__ldg(A.sums + 1) + kEps;
    const float gscale = A.grad_scale ? __ldg(A.grad_scale) : 1.f;
    float lam[6];
#pragma unroll
    for (int q = 0; q < 6; ++q) lam[q] = backward_only ? __ldg(A.lam_eff + q) : P.lam[q] * gscale;

    // on-the-fly target: patch geometry (from the weights pre-kernel) and the exp table
    PatchGeom geom = PatchGeom{};
    bool cols_hit = false;
    int pcx = 0, pcy = 0;
    if (!has_target && heavy) {
        geom = unpack_geom(__ldg(A.geom + tile), w);
        cols_hit = geom.active && x0 + 3 >= geom.x_from && x0 < geom.x_to;
        pcx = geom.ulx + (int)P.ec.centre; pcy = geom.uly + (int)P.ec.centre;
        if (geom.active) fill_patch_lut(lut, P.ec);
    }
    auto target4 = [&](int it) -> float4 {
        if (has_target) return ldg_keep(tgt4 + it * TPB);
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        const int y = it * ROWS + ty;
        if (cols_hit && y >= geom.y_from && y < geom.y_to) {
            const int dy2 = (y - pcy) * (y - pcy);
            float e[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int xx = x0 + j, dx = xx - pcx;
                e[j] = (xx >= geom.x_from && xx < geom.x_to) ? lut[dx * dx + dy2] : 0.f;
            }
            t = make_float4(e[0], e[1], e[2], e[3]);
        }
        return t;
    };

    // ---- reduction 1: tile maximum -----------------------------------------------------------
    float m = -INFINITY;
#pragma unroll
    for (int it = 0; it < NIT; ++it) m = fmaxf(m, fmaxf(fmaxf(h[it].x, h[it].y), fmaxf(h[it].z, h[it].w)));
    m = block_max1<NW>(m, red0);             // also publishes the exp table
    const float ml = m * kLog2e;

    // ---- pass B: softmax moments, sigmoid mass, squared error ------------------------------------
    float r8[8];
    {
        float Ej[4] = {0.f, 0.f, 0.f, 0.f};
        float Yw = 0.f, Ssum_t = 0.f, mse = 0.f;
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            const float hv[4] = {h[it].x, h[it].y, h[it].z, h[it].w};
            float e[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) { e[j] = ex2(fmaf(hv[j], kLog2e, -ml)); Ej[j] += e[j]; }
            if (CE) Es[it * TPB + tid] = make_float4(e[0], e[1], e[2], e[3]);
            Yw = fmaf((float)(it * ROWS), (e[0] + e[1]) + (e[2] + e[3]), Yw);
            if (heavy) {
                float s[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) s[j] = sigmoid_fast(hv[j]);
                if (CS) Ss[it * TPB + tid] = make_float4(s[0], s[1], s[2], s[3]);
                Ssum_t += (s[0] + s[1]) + (s[2] + s[3]);
                const float4 t = target4(it);
                const float d0 = hv[0] - t.x, d1 = hv[1] - t.y, d2 = hv[2] - t.z, d3 = hv[3] - t.w;
                mse += fmaf(d0, d0, d1 * d1) + fmaf(d2, d2, d3 * d3);
            }
        }
        const float Zt = (Ej[0] + Ej[1]) + (Ej[2] + Ej[3]);
        r8[0] = Zt;
        r8[1] = fmaf(fx0, Zt, fmaf(3.f, Ej[3], fmaf(2.f, Ej[2], Ej[1])));
        r8[2] = fmaf(fty, Zt, Yw);
        r8[3] = Ssum_t; r8[4] = mse; r8[5] = vsum; r8[6] = 0.f; r8[7] = 0.f;
    }
    block_sum1<8, NW>(r8, red1);
    const float iZ = 1.f / r8[0];
    const float cx = r8[1] * iZ, cy = r8[2] * iZ;
    const float Ssum = r8[3], mse_sum = r8[4], mV = A.var ? r8[5] / (float)N : P.sigma;

    float* gh = grads ? A.grad_hm + (size_t)tile * N : nullptr;
    const float* hm_tile = A.hm + (size_t)tile * N;
    const float* off_tile = A.off + (size_t)tile * 2 * N;

    // ---- weight 0: every term carries a factor w -> zero loss and gradient; decode only ----------
    if (!heavy) {
        if (decode && tid < 32) {
            float dx_ = cx, dy_ = cy; int px, py;
            refine_and_correct(hm_tile, nullptr, off_tile, A.alpha_param, A.fusion_weight, H, W, A.radius, A.dflags, dx_, dy_, px, py);
            if (tid == 0) { A.coords[2 * tile] = dx_; A.coords[2 * tile + 1] = dy_; A.scores[tile] = m; }
        }
        if (tid == 0 && !backward_only) {
            float4* p = reinterpret_cast<float4*>(A.partial + (size_t)tile * 8);
            p[0] = z4; p[1] = z4;
        }
        if (grads) {
#pragma unroll
            for (int it = 0; it < NIT; ++it) {
                stg_stream(gh4 + it * TPB, z4);
                if (gv4) stg_stream(gv4 + it * TPB, z4);
            }
        }
        return;
    }

    // ---- things that only need the soft-argmax: start their loads now ----------------------------
    const float ka = wa / (P.use_target_weight ? D : (float)(P.B * P.K)), kb = w / D;
    // (1) the 8 offset taps of the offset term (same addresses in every thread: one L1 line per warp)
    const Taps tp = taps_setup(cx, cy, H, W);
    float ov[2][4];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        ov[c][0] = __ldg(off_tile + c * N + tp.i00); ov[c][1] = __ldg(off_tile + c * N + tp.i01) * tp.okx;
        ov[c][2] = __ldg(off_tile + c * N + tp.i10) * tp.oky; ov[c][3] = __ldg(off_tile + c * N + tp.i11) * (tp.okx * tp.oky);
    }
    // (2) decode stage 1 (warp 0): window taps around the rounded soft-argmax
    const bool staged_decode = decode && (A.dflags & GBCODEC_DECODE_REFINE) && A.radius <= 2;
    float win = -INFINITY, winx = 0.f, winy = 0.f;
    bool win_ok = false;
    if (staged_decode && tid < 32) {
        const int px = (int)fminf(fmaxf(rintf(cx), 0.f), (float)(W - 1));
        const int py = (int)fminf(fmaxf(rintf(cy), 0.f), (float)(H - 1));
        const int S = 2 * A.radius + 1;
        const int x = px - A.radius + lane % S, y = py - A.radius + lane / S;
        win_ok = lane < S * S && x >= 0 && x < W && y >= 0 && y < H;
        winx = (float)x; winy = (float)y;
        if (win_ok) win = __ldg(hm_tile + y * W + x);
    }
    // (3) the variance-map gradient is uniform over the tile
    if (gv4) {
        const float g = lam[3] * kb * 2.f * (mV - P.sigma) / (float)N;
        const float4 g4 = make_float4(g, g, g, g);
#pragma unroll
        for (int it = 0; it < NIT; ++it) stg_stream(gv4 + it * TPB, g4);
    }

    // ---- pass C: entropy sums and relu moments about (cx, cy) ----------------------------------------
    float dxj[4], dx2j[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { dxj[j] = (fx0 + (float)j) - cx; dx2j[j] = dxj[j] * dxj[j]; }
    const float dy0 = fty - cy;
    {
        float A1 = 0.f, A2 = 0.f, Ry = 0.f, Ry2 = 0.f;
        float Rj[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            const float hv[4] = {h[it].x, h[it].y, h[it].z, h[it].w};
            float e[4];
            if (CE) { const float4 q = Es[it * TPB + tid]; e[0] = q.x; e[1] = q.y; e[2] = q.z; e[3] = q.w; }
            else {
#pragma unroll
                for (int j = 0; j < 4; ++j) e[j] = ex2(fmaf(hv[j], kLog2e, -ml));
            }
            float p[4], a[4], r[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                p[j] = e[j] * iZ;
                const float u = p[j] + kEps;
                const float l = lg2(u), rc = rcp(u);
                A1 = fmaf(p[j], l, A1);
                const float prc = p[j] * rc;
                A2 = fmaf(p[j], prc, A2);
                a[j] = fmaf(-kLn2, l, -prc);
                r[j] = fmaxf(hv[j], 0.f);
                Rj[j] += r[j];
            }
            if (CE) Es[it * TPB + tid] = make_float4(p[0], p[1], p[2], p[3]);
            if (CA) As[it * TPB + tid] = make_float4(a[0], a[1], a[2], a[3]);
            const float rs = (r[0] + r[1]) + (r[2] + r[3]);
            const float dy = dy0 + (float)(it * ROWS);
            Ry = fmaf(dy, rs, Ry);
            Ry2 = fmaf(dy * dy, rs, Ry2);
        }
        r8[0] = A1; r8[1] = A2;
        r8[2] = fmaf(dx2j[0], Rj[0], fmaf(dx2j[1], Rj[1], fmaf(dx2j[2], Rj[2], fmaf(dx2j[3], Rj[3], Ry2))));
        r8[3] = fmaf(dxj[0], Rj[0], fmaf(dxj[1], Rj[1], fmaf(dxj[2], Rj[2], dxj[3] * Rj[3])));
        r8[4] = Ry;
        r8[5] = (Rj[0] + Rj[1]) + (Rj[2] + Rj[3]);
        r8[6] = 0.f; r8[7] = 0.f;
    }
    block_sum1<8, NW>(r8, red0);

    // ---- decode stage 2 (warp 0): window softmax, blend, start the bilinear offset read ---------------
    float dcx = cx, dcy = cy, dtap = 0.f;
    Bilinear dbl = Bilinear{};
    if (decode && tid < 32) {
        if (staged_decode) {
            const float vmax = warp_max(win);
            const float e = win_ok ? expf(win - vmax) : 0.f;
            const float se = warp_sum(e), sx = warp_sum(e * winx), sy = warp_sum(e * winy);
            const float a = sigmoid_acc(__ldg(A.alpha_param));
            dcx = a * cx + (1.f - a) * (sx / se);
            dcy = a * cy + (1.f - a) * (sy / se);
        } else if (A.dflags & GBCODEC_DECODE_REFINE) {
            int px, py;
            refine_and_correct(hm_tile, nullptr, nullptr, A.alpha_param, nullptr, H, W, A.radius, GBCODEC_DECODE_REFINE, dcx, dcy, px, py);
        }
        if (A.dflags & GBCODEC_DECODE_APPLY_OFFSET) {
            dbl = bilinear_setup(dcx, dcy, H, W);
            // lane t < 8 fetches tap (t & 3) of channel (t >> 2)
            const int tap = lane & 3;
            const int yy = (tap & 2) ? dbl.y1 : dbl.y0, xx = (tap & 1) ? dbl.x1 : dbl.x0;
            if (lane < 8) dtap = __ldg(off_tile + (lane >> 2) * N + yy * W + xx);
        }
    }

    // ---- per-tile scalars (every thread: nobody waits) ---------------------------------------------------
    const float gx = __ldg(A.gt + 2 * tile) * ((float)W / P.in_w);
    const float gy = __ldg(A.gt + 2 * tile + 1) * ((float)H / P.in_h);
    const float Rp = r8[5] + kEps;
    const float iRp = 1.f / Rp;
    float c1, c4, k4, c6, fxx, fyy, go0, go1, pa_;
    {
        float sl1 = 0.f, sl1p[2], dsdx[2], dsdy[2];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const float samp = tp.w00 * ov[c][0] + tp.w01 * ov[c][1] + tp.w10 * ov[c][2] + tp.w11 * ov[c][3];
            dsdx[c] = ((1.f - tp.fy) * (ov[c][1] - ov[c][0]) + tp.fy * (ov[c][3] - ov[c][2])) * tp.inx;
            dsdy[c] = ((1.f - tp.fx) * (ov[c][2] - ov[c][0]) + tp.fx * (ov[c][3] - ov[c][1])) * tp.iny;
            const float d = samp - ((c == 0 ? gx : gy) - (c == 0 ? cx : cy));
            const float ad = fabsf(d);
            sl1 += ad < 1.f ? 0.5f * d * d : ad - 0.5f;
            sl1p[c] = ad < 1.f ? d : (d > 0.f ? 1.f : -1.f);
        }
        const float off_t = 0.5f * sl1;
        const float peak_t = (cx - gx) * (cx - gx) + (cy - gy) * (cy - gy);
        const float v = r8[2] * iRp;
        const float s = sqrtf(v + kEps);
        const float var_t = (s - P.sigma) * (s - P.sigma) + (A.var ? (mV - P.sigma) * (mV - P.sigma) : 0.f);
        const float E = -kLn2 * r8[0];
        const float pa = E - r8[1];
        const float shape_t = (E - P.e_star) * (E - P.e_star);
        if (tid == 0 && !backward_only) {
            float* p = A.partial + (size_t)tile * 8;
            p[0] = wa * (mse_sum / (float)N); p[1] = wa * off_t; p[2] = wa * peak_t;
            p[3] = w * var_t; p[5] = w * shape_t;      // p[4] (limb overlap) follows the partner pass
        }
        c1 = lam[0] * ka * 2.f / (float)N;
        const float a4 = lam[3] * kb * (s - P.sigma) / s;
        c4 = a4 * iRp;
        k4 = -c4 * v;
        c6 = lam[5] * kb * 2.f * (E - P.e_star);
        pa_ = pa;
        const float dv_dcx = -2.f * r8[3] * iRp, dv_dcy = -2.f * r8[4] * iRp;
        fxx = lam[2] * ka * 2.f * (cx - gx) + lam[1] * ka * 0.5f * (sl1p[0] * (dsdx[0] + 1.f) + sl1p[1] * dsdx[1]) + a4 * dv_dcx;
        fyy = lam[2] * ka * 2.f * (cy - gy) + lam[1] * ka * 0.5f * (sl1p[0] * dsdy[0] + sl1p[1] * (dsdy[1] + 1.f)) + a4 * dv_dcy;
        go0 = lam[1] * ka * 0.5f * sl1p[0];
        go1 = lam[1] * ka * 0.5f * sl1p[1];
    }

    // ---- limb partners: one visit each; sums for the overlap ratio, one tie bit per pixel -------------
    const int np = P.n_partner[k];
    unsigned bits[GBCODEC_MAX_PARTNERS];
    float wj_[GBCODEC_MAX_PARTNERS];
    unsigned eqflags = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) r8[q] = 0.f;
#pragma unroll
    for (int pi = 0; pi < GBCODEC_MAX_PARTNERS; ++pi) {
        bits[pi] = 0u; wj_[pi] = 0.f;
        if (pi < np) {                                              // CTA-uniform
            const int j = P.partner[k][pi];
            const float wj = __ldg(A.weff + b * P.K + j);
            wj_[pi] = wj;
            if (w != 0.f && wj != 0.f) {
                const float4* hj4 = reinterpret_cast<const float4*>(A.hm) + ((size_t)b * P.K + j) * N4 + tid;
                float4 qv[NIT];
#pragma unroll
                for (int it = 0; it < NIT; ++it) qv[it] = ldg_stream(hj4 + it * TPB);
                float Sj = 0.f, M = 0.f;
                bool anyeq = false;
#pragma unroll
                for (int it = 0; it < NIT; ++it) {
                    const float hv[4] = {h[it].x, h[it].y, h[it].z, h[it].w};
                    const float qq[4] = {qv[it].x, qv[it].y, qv[it].z, qv[it].w};
                    float sk[4];
                    if (CS) { const float4 s4 = Ss[it * TPB + tid]; sk[0] = s4.x; sk[1] = s4.y; sk[2] = s4.z; sk[3] = s4.w; }
                    else {
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) sk[jj] = sigmoid_fast(hv[jj]);
                    }
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const float sq = sigmoid_fast(qq[jj]);
                        Sj += sq;
                        // min(sigma(a), sigma(b)) = sigma(min(a, b)): decide on the logits; equal logits give equal sigmoids
                        const bool own_smaller = hv[jj] < qq[jj];
                        M += own_smaller ? sk[jj] : sq;
                        if (own_smaller) bits[pi] |= 1u << (it * 4 + jj);
                        anyeq |= hv[jj] == qq[jj];
                    }
                }
                r8[2 * pi] = Sj; r8[2 * pi + 1] = M;
                if (anyeq) eqflags |= 1u << pi;
            }
        }
    }
    float cj[GBCODEC_MAX_PARTNERS] = {0.f, 0.f, 0.f, 0.f};
    float cst = 0.f;
    bool g_live = false;
    if (np > 0) {
        block_sum1<8, NW>(r8, red1);
        float pair_loss = 0.f;
#pragma unroll
        for (int pi = 0; pi < GBCODEC_MAX_PARTNERS; ++pi) {
            if (pi < np && w != 0.f && wj_[pi] != 0.f) {
                const float Sj = r8[2 * pi], M = r8[2 * pi + 1];
                const float mm = fminf(Ssum, Sj) + kEps;
                const float rho = M / mm;
                if ((P.owner[k] >> pi) & 1) pair_loss += w * wj_[pi] * fmaxf(rho - 0.5f, 0.f);
                if (grads && rho > 0.5f) {
                    cj[pi] = lam[4] * w * wj_[pi] / D5 / mm;
                    cst += cj[pi] * rho * tie_rule(Ssum, Sj);
                    g_live = true;
                }
            }
        }
        if (tid == 0 && !backward_only) A.partial[(size_t)tile * 8 + 4] = pair_loss;
    } else if (tid == 0 && !backward_only) {
        A.partial[(size_t)tile * 8 + 4] = 0.f;
    }

    // ---- decode stage 3 (warp 0): finish the bilinear read, publish -----------------------------------
    if (decode && tid < 32) {
        if (A.dflags & GBCODEC_DECODE_APPLY_OFFSET) {
            float fw = __ldg(A.fusion_weight);
            if (A.dflags & GBCODEC_DECODE_FUSION_WEIGHT_RAW) fw = sigmoid_acc(fw);
            float t[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) t[q] = __shfl_sync(0xffffffffu, dtap, q);
            const float ox = dbl.w00 * t[0] + dbl.w01 * (t[1] * dbl.okx) + dbl.w10 * (t[2] * dbl.oky) + dbl.w11 * (t[3] * (dbl.okx * dbl.oky));
            const float oy = dbl.w00 * t[4] + dbl.w01 * (t[5] * dbl.okx) + dbl.w10 * (t[6] * dbl.oky) + dbl.w11 * (t[7] * (dbl.okx * dbl.oky));
            dcx += fw * ox;
            dcy += fw * oy;
        }
        if (tid == 0) { A.coords[2 * tile] = dcx; A.coords[2 * tile + 1] = dcy; A.scores[tile] = m; }
    }
    if (!grads) return;

    // ---- pass D: the heatmap gradient ----------------------------------------------------------------------
    float basej[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) basej[j] = dxj[j] * fxx;
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
        const float hv[4] = {h[it].x, h[it].y, h[it].z, h[it].w};
        float p[4], a[4];
        if (CE) { const float4 q = Es[it * TPB + tid]; p[0] = q.x; p[1] = q.y; p[2] = q.z; p[3] = q.w; }
        else {
#pragma unroll
            for (int j = 0; j < 4; ++j) p[j] = ex2(fmaf(hv[j], kLog2e, -ml)) * iZ;
        }
        if (CA) { const float4 q = As[it * TPB + tid]; a[0] = q.x; a[1] = q.y; a[2] = q.z; a[3] = q.w; }
        else {
#pragma unroll
            for (int j = 0; j < 4; ++j) { const float u = p[j] + kEps; a[j] = fmaf(-kLn2, lg2(u), -p[j] * rcp(u)); }
        }
        const float4 t = target4(it);
        const float tv[4] = {t.x, t.y, t.z, t.w};
        const float dy = dy0 + (float)(it * ROWS);
        const float dy2 = dy * dy, fyd = dy * fyy;
        float out[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float g = c1 * (hv[j] - tv[j]);
            g = fmaf(p[j], fmaf(c6, a[j] - pa_, basej[j] + fyd), g);
            const float rterm = fmaf(c4, dx2j[j] + dy2, k4);
            if (hv[j] > 0.f) g += rterm;
            out[j] = g;
        }
        if (g_live) {
            float G[4] = {-cst, -cst, -cst, -cst};
#pragma unroll
            for (int pi = 0; pi < GBCODEC_MAX_PARTNERS; ++pi) {
                if (cj[pi] != 0.f) {                                  // CTA-uniform
#pragma unroll
                    for (int j = 0; j < 4; ++j) if ((bits[pi] >> (it * 4 + j)) & 1u) G[j] += cj[pi];
                }
            }
            if (eqflags) {
                // rare: some logit of this thread equals its partner's; ATen's minimum splits that gradient evenly
#pragma unroll
                for (int pi = 0; pi < GBCODEC_MAX_PARTNERS; ++pi) {
                    if (((eqflags >> pi) & 1u) && cj[pi] != 0.f) {
                        const int jp = P.partner[k][pi];
                        const float4 q = ldg_keep(reinterpret_cast<const float4*>(A.hm) + ((size_t)b * P.K + jp) * N4 + tid + it * TPB);
                        const float qq[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                        for (int j = 0; j < 4; ++j) if (qq[j] == hv[j]) G[j] += 0.5f * cj[pi];
                    }
                }
            }
            float sv[4];
            if (CS) { const float4 s4 = Ss[it * TPB + tid]; sv[0] = s4.x; sv[1] = s4.y; sv[2] = s4.z; sv[3] = s4.w; }
            else {
#pragma unroll
                for (int j = 0; j < 4; ++j) sv[j] = sigmoid_fast(hv[j]);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) out[j] = fmaf(G[j] * sv[j], 1.f - sv[j], out[j]);
        }
        stg_stream(gh4 + it * TPB, make_float4(out[0], out[1], out[2], out[3]));
    }

    // the (up to) four non-zero taps per channel of the offset gradient; the zero fill of these
    // addresses was issued before the first barrier, so it is ordered before these stores
    if (tid == 0) {
        float* go = A.grad_off + (size_t)tile * 2 * N;
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
            float* o = go + ch * N;
            const float gc = ch == 0 ? go0 : go1;
            o[tp.i00] = gc * tp.w00;
            if (tp.okx != 0.f) o[tp.i01] = gc * tp.w01;
            if (tp.oky != 0.f) o[tp.i10] = gc * tp.w10;
            if (tp.okx != 0.f && tp.oky != 0.f) o[tp.i11] = gc * tp.w11;
        }
    }
}

// ---- launcher ----------------------------------------------------------------------------------------
template <int W4, int ROWS, int NIT, bool CE, bool CS, bool CA, int MINB>
static int launch_tile_t(const LossParams& P, const LossArgs& A, cudaStream_t s, cudaEvent_t e0, cudaEvent_t e1) {
    constexpr int TPB = W4 * ROWS, N4 = TPB * NIT, NW = TPB / 32;
    const size_t smem = (size_t)N4 * 16 * ((CE ? 1 : 0) + (CS ? 1 : 0) + (CA ? 1 : 0))
                      + (size_t)((P.ec.lut_size + 3) & ~3) * 4 + (size_t)2 * NW * 8 * 4;
    auto kern = loss_tile_kernel<W4, ROWS, NIT, CE, CS, CA, MINB>;
    if (smem > 227 * 1024) return 1;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(GBCODEC_ERR_CUDA, "cudaFuncSetAttribute(loss_tile_kernel): %s", cudaGetErrorString(e));
    e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return fail(GBCODEC_ERR_CUDA, "cudaFuncSetAttribute(carveout): %s", cudaGetErrorString(e));
    if (e0 && !A.plan) cudaEventRecord(e0, s);
    kern<<<P.B * P.K, TPB, smem, s>>>(P, A);
    if (e1 && !A.plan) cudaEventRecord(e1, s);
    return check_launch("loss_tile_kernel");
}

int launch_loss_tile(const LossParams& P, const LossArgs& A, cudaStream_t s, cudaEvent_t e0, cudaEvent_t e1) {
    if (P.H == 64 && P.W == 48) return launch_tile_t<12, 16, 4, true, true, true, 4>(P, A, s, e0, e1);      // 192 threads, 16 px each
    if (P.H == 96 && P.W == 72) return launch_tile_t<18, 16, 6, true, true, true, 2>(P, A, s, e0, e1);      // 288 threads, 24 px each
    if (P.H == 128 && P.W == 128) return launch_tile_t<32, 16, 8, false, true, true, 1>(P, A, s, e0, e1);   // 512 threads, 32 px each
    return 1;
}

}  // namespace gbc
