// step_tile.cu — persistent kernel of the fused codec step: target tiles generated on the fly + six-term fusion loss
// forward and backward + keypoint decode, one pass over every heatmap.
//
// Same arithmetic as loss_tile.cu (FusionPoseLoss.forward, models/fusion_head.py:745-806, terms :637-743 and :405-559,
// the autograd backward of train.py:182 in closed form; decode: HeatmapRegressionHead.decode, fusion_head.py:309-365);
// a different machine mapping, built around what bounded the one-CTA-per-tile kernel (issue slots, the MUFU pipe and
// three block barriers per tile — not HBM):
//
//   * persistent CTAs (SMs x resident CTAs) draw tiles from an atomic counter; everything that does not depend on the
//     tile (exp table of the target patch, normalisers, loss weights, barriers) is set up once per CTA;
//   * the tile, its variance tile, its 64-byte descriptor (weight, target geometry, ground truth, active limb partners —
//     written by denoms_kernel) and the limb partners' tiles arrive by bulk copies (cp.async.bulk + mbarrier
//     complete_tx, SASS UBLKCP / SYNCS) issued by ONE thread; the next tile's copies are in flight while the current
//     tile is computed (two tile buffers, two partner buffers).  The landing layout is the tile's linear layout, which
//     is also the thread-private slot layout `it * TPB + tid`: no per-thread copy instructions, no address arithmetic;
//   * ONE block reduction (one barrier) per tile: the entropy term is accumulated as sum e*t with t = (h - m) log2 e —
//     log(p + eps) = log p + eps/p - ... and sum_i p_i (eps / p_i) = N eps exactly, so the entropy and its gradient need
//     no per-pixel log when eps Z / e_min is small — and the relu moments of the variance term are taken about the tile
//     centre and shifted to the soft-argmax afterwards (the soft-argmax of a beta = 1 softmax sits within a fraction of
//     a pixel of the centre, so the shift does not cancel).  Both shortcuts are guarded per tile by the conditions
//     under which they hold to 1e-7; a tile that fails them takes the general second pass (one more reduction);
//   * sigmoid(h) = e / (e + exp(-m)) from the softmax numerator e = exp(h - m): one MUFU operation instead of two;
//   * after the reduction every warp derives the per-tile scalars itself (no role hand-off through a third barrier);
//     the eight offset taps of the loss term come from a 4x4x2 window around the tile centre that one warp requested
//     when the tile started (falls back to dependent loads when the soft-argmax leaves the window).
//
// Algorithmic HBM bytes per tile: read hm, var (8N); write d_hm, d_var, d_off (16N).
#include "loss_common.cuh"
#include "f32x2.cuh"
#include <stdlib.h>
#include <type_traits>
#include <stdio.h>

namespace gbc {

// ---- bulk copies and mbarriers (PTX) -----------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    asm volatile("{\n"
                 " .reg .pred p;\n"
                 "WAIT_%=:\n"
                 " mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 " @p bra DONE_%=;\n"
                 " bra WAIT_%=;\n"
                 "DONE_%=:\n"
                 "}" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
// the producer warp's waits: back off between polls so that its spinning does not take issue slots from the compute warps
__device__ __forceinline__ bool mbar_test(uint64_t* bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// the producer warp's waits: the hardware suspends the thread for up to the hinted time per attempt, so a long wait
// costs a handful of issue slots (a test + nanosleep loop took 11 % of the kernel's instructions)
#ifndef STEP_WAITHINT
#define STEP_WAITHINT 1
#endif
#ifndef STEP_ROLL
#define STEP_ROLL 1
#endif
#ifndef STEP_SMEMSCAL
#define STEP_SMEMSCAL 1
#endif
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, unsigned parity) {
#if !STEP_WAITHINT
    while (!mbar_test(bar, parity)) __nanosleep(128);
    return;
#endif
    asm volatile("{\n"
                 " .reg .pred p;\n"
                 "WAITB_%=:\n"
                 " mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
                 " @p bra DONEB_%=;\n"
                 " bra WAITB_%=;\n"
                 "DONEB_%=:\n"
                 "}" :: "r"(smem_u32(bar)), "r"(parity), "r"(4000u) : "memory");
}
// global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// shared -> global, tracked by the issuing thread's bulk async-groups
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ float max3f(float a, float b, float c) { float d; asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float min3f(float a, float b, float c) { float d; asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float sqrt_approx(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// ---- shared-memory plan -------------------------------------------------------------------------
constexpr float kFlatRmax = 1e-3f;              // largest eps / p for which the entropy shortcut holds to 1e-7
constexpr float kShiftCond = 4.f;               // largest sum |terms| / |result| accepted for the shifted relu moments

template <int W4, int ROWS, int NIT>
struct StepPlan {
    static constexpr int TPB = W4 * ROWS, NW = TPB / 32, N4 = TPB * NIT, N = 4 * N4, W = 4 * W4, H = ROWS * NIT;
    static constexpr int kTile = N4 * 16;
    static constexpr int oH = 0;                                 // two tile buffers
    static constexpr int oV = 2 * kTile;                         // variance tile
    static constexpr int oS = 3 * kTile;                         // sigmoid of the own tile
    static constexpr int oQ = 4 * kTile;                         // two partner buffers
    static constexpr int oDesc = 6 * kTile;                      // two descriptors
    static constexpr int oRed = oDesc + 2 * 64;                  // 2 x NW x 16 floats
    static constexpr int oRed2 = oRed + 2 * NW * 16 * 4;         // 2 x NW x 4 floats (partners 3 and 4)
    static constexpr int oRedC = oRed2 + 2 * NW * 4 * 4;         // NW x 8 floats (general second pass)
    static constexpr int oRedM = oRedC + NW * 8 * 4;             // 2 x NW x 2 floats (per-warp max, min)
    static constexpr int oTap = oRedM + 2 * NW * 2 * 4;          // 2 x 32 floats (offset-tap window)
    static constexpr int oTab = oTap + 2 * 32 * 4;               // NW x 16 floats (overlap coefficient per tie pattern)
    static constexpr int oBar = oTab + NW * 16 * 4;              // 10 mbarriers
    static constexpr int oTid = oBar + 10 * 8;                   // 2 tile indices (+ pad)
    static constexpr int oDec = oTid + 16;                       // 2 x float4: soft-argmax and maximum handed to the producer warp
    static constexpr int oCta = oDec + 32;                       // float4: 1/(sum w + eps), 1/(sum w_i w_j + eps), gradient scale
    static constexpr int oLut = oCta + 16;                       // exp table of the target patch
    static_assert(TPB % 32 == 0 && NW >= 2, "whole warps; the tap window needs a second warp");
    static_assert((oBar % 8) == 0 && (oDesc % 16) == 0 && (oLut % 16) == 0, "alignment");
};

// ---- small pieces ---------------------------------------------------------------------------------

// NV running sums per lane -> block totals, lane-distributed (value k in lanes k*SUB .. k*SUB+SUB-1 after the final
// shuffle: fetch with lane k * SUB).  First NSC values carry softmax scaling relative to the warp's maximum.
template <int NV> struct Lg2 { static constexpr int v = 1 + Lg2<NV / 2>::v; };
template <> struct Lg2<1> { static constexpr int v = 0; };

// ---- the kernel ------------------------------------------------------------------------------------
// GRADS = false: forward only (validate.py:90 evaluates the loss under no_grad every batch): no sigmoid tile, no tie
// words, no gradient pass, no stores besides the per-tile loss numerators and the decode.
template <int W4, int ROWS, int NIT, int MINB, bool GRADS>
__global__ void __launch_bounds__(W4* ROWS + 32, MINB)
step_tile_kernel(const __grid_constant__ LossParams P, const __grid_constant__ LossArgs A) {
    using L = StepPlan<W4, ROWS, NIT>;
    constexpr int TPB = L::TPB, NW = L::NW, N4 = L::N4, N = L::N, W = L::W, H = L::H;
    constexpr unsigned kTile = L::kTile;
    extern __shared__ __align__(128) unsigned char smraw[];
    float4* const Hb = reinterpret_cast<float4*>(smraw + L::oH);
    float4* const Vb = reinterpret_cast<float4*>(smraw + L::oV);
    float4* const Sb = reinterpret_cast<float4*>(smraw + L::oS);
    float4* const Qb = reinterpret_cast<float4*>(smraw + L::oQ);
    TileDesc* const Db = reinterpret_cast<TileDesc*>(smraw + L::oDesc);
    float* const red = reinterpret_cast<float*>(smraw + L::oRed);
    float* const red2 = reinterpret_cast<float*>(smraw + L::oRed2);
    float* const redC = reinterpret_cast<float*>(smraw + L::oRedC);
    float* const redM = reinterpret_cast<float*>(smraw + L::oRedM);
    float* const tapw = reinterpret_cast<float*>(smraw + L::oTap);
    float* const tabs = reinterpret_cast<float*>(smraw + L::oTab);
    uint64_t* const bars = reinterpret_cast<uint64_t*>(smraw + L::oBar);
    int* const tids = reinterpret_cast<int*>(smraw + L::oTid);
    float4* const decs = reinterpret_cast<float4*>(smraw + L::oDec);
    float* const lut = reinterpret_cast<float*>(smraw + L::oLut);
    float4* const ctas = reinterpret_cast<float4*>(smraw + L::oCta);
    uint64_t* const hfull = bars;          // [2] tile + descriptor landed
    uint64_t* const vfull = bars + 2;      // variance tile landed
    uint64_t* const qfull = bars + 3;      // [2] partner tile landed
    uint64_t* const hempty = bars + 5;     // [2] every compute warp is done with the tile buffer
    uint64_t* const qempty = bars + 7;     // [2] ... with the partner buffer
    uint64_t* const vempty = bars + 9;     // ... with the variance buffer

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tiles = P.B * P.K;
    const bool has_var = A.var != nullptr;
    const bool decode = A.coords != nullptr;

    // ---- once per CTA -------------------------------------------------------------------------------
    if (tid == 0) {
        mbar_init(hfull, 1); mbar_init(hfull + 1, 1); mbar_init(vfull, 1);
        mbar_init(qfull, 1); mbar_init(qfull + 1, 1);
        mbar_init(hempty, NW); mbar_init(hempty + 1, NW);
        mbar_init(qempty, NW); mbar_init(qempty + 1, NW); mbar_init(vempty, NW);
        fence_mbar_init();
    }
    for (int q = tid; q < P.ec.lut_size; q += TPB + 32) lut[q] = expf(-(float)q / P.ec.two_sigma_sq);
    __syncthreads();

    const float* const hm = A.hm;
    const bool producer = warp == NW;
    if (producer && lane == 0) {
        const size_t t0 = blockIdx.x;
        tids[0] = (int)blockIdx.x;                     // published by the arrive below (release)
        mbar_arrive_expect_tx(hfull, kTile + 64);
        bulk_g2s(Hb, hm + t0 * N, kTile, hfull);
        if (has_var) { mbar_arrive_expect_tx(vfull, kTile); bulk_g2s(Vb, A.var + t0 * N, kTile, vfull); }
    }
    // everything below depends on the kernel that prepared the weights (programmatic dependent launch)
    pdl_wait();
    pdl_launch_dependents();

    // ---- producer warp -------------------------------------------------------------------------------------------------
    // Per tile j (buffer s = j & 1): wait until the compute warps have left tile j - 2 and finish that tile's keypoint
    // decode (local refinement + offset correction from its soft-argmax, handed over through shared memory: dependent
    // loads that would otherwise sit on the compute warps' path); then lane 0 starts the tile + descriptor copy, reads the
    // descriptor when it has landed, starts the partner copies as partner buffers come free, then the variance tile.
    // The compute warps never wait for one another except at their one block barrier.
    if (producer) {
        if (lane == 0) bulk_g2s(Db, A.desc + blockIdx.x, 64, hfull);
        unsigned cur = blockIdx.x, nxt = 0, nxt2 = 0;
        if (lane == 0) {
            nxt = atomicAdd(A.tile_counter, 1u) + gridDim.x;
            nxt2 = atomicAdd(A.tile_counter, 1u) + gridDim.x;     // two ahead: the atomic's latency is off the path
        }
        unsigned pq = 0;                                          // partner copies started so far (lane 0)
        int told[2] = {-1, -1};                                   // tiles whose decode is still to be finished, per buffer
        auto finish_decode = [&](unsigned sb) {
            const int t = told[sb];
            if (t < 0 || !decode) return;
            const float4 d = decs[sb];
            float dx_ = d.x, dy_ = d.y;
            int px, py;
            refine_and_correct<float>(hm + (size_t)t * N, nullptr, A.off + (size_t)t * 2 * N, A.alpha_param, A.fusion_weight,
                                      H, W, A.radius, A.dflags, dx_, dy_, px, py);
            if (lane == 0) { A.coords[2 * t] = dx_; A.coords[2 * t + 1] = dy_; A.scores[t] = d.z; }
        };
        for (unsigned j = 0;; ++j) {
            const unsigned s = j & 1u;
            if (j >= 2) {
                mbar_wait_backoff(hempty + s, ((j - 2) >> 1) & 1u);
                finish_decode(s);
            }
            cur = __shfl_sync(0xffffffffu, cur, 0);
            if (cur >= (unsigned)tiles) {
                if (lane == 0) { tids[s] = -1; mbar_arrive(hfull + s); }
                // the tile in the other buffer is the last one this CTA has taken
                if (j >= 1) {
                    mbar_wait_backoff(hempty + (s ^ 1u), ((j - 1) >> 1) & 1u);
                    finish_decode(s ^ 1u);
                }
                break;
            }
            told[s] = (int)cur;
            if (lane == 0) {
                if (j > 0) {
                    tids[s] = (int)cur;
                    mbar_arrive_expect_tx(hfull + s, kTile + 64);
                    bulk_g2s(Hb + s * N4, hm + (size_t)cur * N, kTile, hfull + s);
                    bulk_g2s(Db + s, A.desc + cur, 64, hfull + s);
                }
                mbar_wait_backoff(hfull + s, (j >> 1) & 1u);
                const TileDesc* nd = Db + s;
                const int nn = (int)(nd->pk & 7u);
                const unsigned pj = nd->pj;
                if (nn > 0) {
                    const size_t b = cur / (unsigned)P.K;
                    for (int n = 0; n < nn; ++n) {
                        const unsigned q = pq & 1u;
                        if (pq >= 2) mbar_wait_backoff(qempty + q, ((pq - 2) >> 1) & 1u);
                        mbar_arrive_expect_tx(qfull + q, kTile);
                        bulk_g2s(Qb + q * N4, hm + (b * P.K + ((pj >> (8 * n)) & 0xFFu)) * N, kTile, qfull + q);
                        ++pq;
                    }
                }
                if (has_var && j > 0) {
                    mbar_wait_backoff(vempty, (j - 1) & 1u);       // every compute warp has waited for the copy and is done with it
                    mbar_arrive_expect_tx(vfull, kTile);
                    bulk_g2s(Vb, A.var + (size_t)cur * N, kTile, vfull);
                }
                cur = nxt; nxt = nxt2;
                if (nxt2 < (unsigned)tiles) nxt2 = atomicAdd(A.tile_counter, 1u) + gridDim.x;
            }
            __syncwarp();
        }
        return;
    }
    auto compute_barrier = [] { asm volatile("bar.sync 1, %0;" :: "n"(TPB) : "memory"); };

    // per-CTA scalars (the loss weights stay in the constant bank: lam[q] = P.lam[q] * gscale where they are used).
    // They go through shared memory: kept as __ldg values the compiler re-issued the global loads inside the tile loop
    // (a dependent L2 round trip in every tile's scalar chain: 6 % of the stall samples).
    if (tid == 0)
        ctas[0] = make_float4(rcp((float)__ldg(A.sums) + kEps), rcp((float)__ldg(A.sums + 1) + kEps),
                              A.grad_scale ? __ldg(A.grad_scale) : 1.f, 0.f);
    compute_barrier();

    // thread geometry: four columns x0 .. x0+3 of rows ty, ty + ROWS, ...
    const int tx = tid % W4, ty = tid / W4;
    const int x0 = tx << 2;
    const float fx0 = (float)x0, fty = (float)ty;
    constexpr float ax = 0.5f * (float)(W - 1), ay = 0.5f * (float)(H - 1);     // anchor of the relu moments
    constexpr int wx0 = (W - 1) / 2 - 1, wy0 = (H - 1) / 2 - 1;                 // origin of the 4x4 tap window
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const f2 kL2E = splat2(kLog2e), kNL2E = splat2(-kLog2e), kOne = splat2(1.f);

    unsigned qn = 0;                        // partner items consumed so far (buffer = qn & 1, parity = (qn >> 1) & 1)
    for (unsigned i = 0;; ++i) {
        const unsigned s = i & 1u;
        float4* const Hs = Hb + s * N4;
        // ---- this tile ----------------------------------------------------------------------------------
        mbar_wait(hfull + s, (i >> 1) & 1u);
        const int tile = tids[s];
        if (tile < 0) break;
        const TileDesc* dsc = Db + s;
        float4 cs4;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(cs4.x), "=f"(cs4.y), "=f"(cs4.z), "=f"(cs4.w) : "r"(smem_u32(ctas)));
#if STEP_SMEMSCAL
        const float iD = cs4.x, iD5 = cs4.y, gscale = cs4.z;
#else
        const float iD = rcp((float)__ldg(A.sums) + kEps), iD5 = rcp((float)__ldg(A.sums + 1) + kEps);
        const float gscale = A.grad_scale ? __ldg(A.grad_scale) : 1.f;
#endif
        const int4 gq = dsc->geom;
        const float4 d1 = *reinterpret_cast<const float4*>(&dsc->w);          // w, gx, gy, pk
        const float w = d1.x, gx = d1.y, gy = d1.z;
        const unsigned pk = __float_as_uint(d1.w);
        const int nact = (int)(pk & 7u);
        const float wa = P.use_target_weight ? w : 1.f;
        const bool heavy = (w != 0.f) || !P.use_target_weight;
        const size_t tile_base = (size_t)tile * N4 + tid;
        const float* off_tile = A.off + (size_t)tile * 2 * N;

        // the offset gradient is zero except on (up to) four taps per channel, patched after the block barrier
        if (GRADS) {
            float4* go4 = reinterpret_cast<float4*>(A.grad_off) + (size_t)tile * 2 * N4 + tid;
#pragma unroll
            for (int it = 0; it < 2 * NIT; ++it) stg_stream(go4 + it * TPB, z4);
        }
        // warp 1: the 4x4x2 window of offset taps around the tile centre (consumed after the reduction)
        float tapv = 0.f;
        if (heavy && warp == 1) tapv = __ldg(off_tile + (lane >> 4) * N + (wy0 + ((lane >> 2) & 3)) * W + wx0 + (lane & 3));

        // ---- maximum (and minimum) of the tile: per warp for now --------------------------------------------
        float mw = -INFINITY, mnw = INFINITY;
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            const float4 o = Hs[it * TPB + tid];
            mw = max3f(mw, o.x, o.y); mw = max3f(mw, o.z, o.w);
            if (heavy) { mnw = min3f(mnw, o.x, o.y); mnw = min3f(mnw, o.z, o.w); }
        }
        mw = warp_max(mw);
        if (heavy) mnw = warp_min(mnw);
        const float nml_w = -mw * kLog2e;

        // ---- on-the-fly target: the one row (if any) in which this thread meets the patch ---------------------
        int hit_it = -1;
        float4 thit = z4;
        if (heavy) {
            const PatchGeom geom = unpack_geom(gq, w);
            if (geom.active && x0 + 3 >= geom.x_from && x0 < geom.x_to) {
                const int it0 = max(0, (geom.y_from - ty + ROWS - 1) / ROWS);
                const int y = it0 * ROWS + ty;
                if (it0 < NIT && y < geom.y_to) {
                    hit_it = it0;
                    const int pcx = geom.ulx + (int)P.ec.centre, pcy = geom.uly + (int)P.ec.centre;
                    const int dy2 = (y - pcy) * (y - pcy);
                    float e[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int xx = x0 + j, dx = xx - pcx;
                        e[j] = (xx >= geom.x_from && xx < geom.x_to) ? lut[dx * dx + dy2] : 0.f;
                    }
                    thit = make_float4(e[0], e[1], e[2], e[3]);
                }
            }
        }

        // ---- pass B: softmax moments (relative to the warp's maximum), entropy sum, sigmoid, relu moments about the
        //      tile centre, sum h^2 -----------------------------------------------------------------------------
        float r16[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) r16[q] = 0.f;
        const bool sig = heavy && nact > 0;
        {
            const f2 kNML = splat2(nml_w);
            f2 E01 = splat2(0.f), E23 = splat2(0.f), T01 = splat2(0.f), T23 = splat2(0.f);
            f2 S2 = splat2(0.f), R01 = splat2(0.f), R23 = splat2(0.f), Hq = splat2(0.f);
            float Yw = 0.f, Ry = 0.f, Ry2 = 0.f;
            const float dyA0 = fty - ay;
            // sigmoid from the softmax numerator needs exp(-m_w) and exp(h - m_w) representable for every h that matters
            const bool fastsig = mw <= 30.f && mw >= -80.f;
            const f2 kCw = splat2(ex2(nml_w));                // exp(-m_w)
            auto body = [&](auto mode_c) {
                constexpr int MODE = decltype(mode_c)::value;     // 0 light, 1 heavy, 2 heavy + sigmoid from e, 3 heavy + plain sigmoid
                constexpr bool kRoll = STEP_ROLL != 0;
                constexpr int kUnroll = kRoll ? (MODE == 2 ? NIT : 1) : (MODE == 3 ? 1 : NIT);
#pragma unroll kUnroll
                for (int it = 0; it < NIT; ++it) {
                    const float4 o = Hs[it * TPB + tid];
                    const f4 hv = as_f4(o);
                    const f2 t01 = fma2(hv.a, kL2E, kNML), t23 = fma2(hv.b, kL2E, kNML);
                    const f2 e01 = pack2(ex2(lo2(t01)), ex2(hi2(t01))), e23 = pack2(ex2(lo2(t23)), ex2(hi2(t23)));
                    E01 = add2(E01, e01); E23 = add2(E23, e23);
                    Yw = fmaf((float)(it * ROWS), hsum2(add2(e01, e23)), Yw);
                    if (MODE >= 1) {
                        T01 = fma2(e01, t01, T01); T23 = fma2(e23, t23, T23);
                        const f2 r01 = pack2(fmaxf(o.x, 0.f), fmaxf(o.y, 0.f)), r23 = pack2(fmaxf(o.z, 0.f), fmaxf(o.w, 0.f));
                        R01 = add2(R01, r01); R23 = add2(R23, r23);
                        const float rs = hsum2(add2(r01, r23));
                        const float dy = dyA0 + (float)(it * ROWS);
                        Ry = fmaf(dy, rs, Ry);
                        Ry2 = fmaf(dy * dy, rs, Ry2);
                        Hq = fma2(hv.a, hv.a, Hq); Hq = fma2(hv.b, hv.b, Hq);
                    }
                    if (MODE == 2) {
                        const f2 g01 = add2(e01, kCw), g23 = add2(e23, kCw);
                        const f2 s01 = mul2(e01, pack2(rcp(lo2(g01)), rcp(hi2(g01)))), s23 = mul2(e23, pack2(rcp(lo2(g23)), rcp(hi2(g23))));
                        Sb[it * TPB + tid] = as_float4(f4{s01, s23});
                        S2 = add2(S2, add2(s01, s23));
                    }
                    if (MODE == 3) {
                        const f2 u01 = mul2(hv.a, kNL2E), u23 = mul2(hv.b, kNL2E);
                        const f2 g01 = add2(pack2(ex2(lo2(u01)), ex2(hi2(u01))), kOne), g23 = add2(pack2(ex2(lo2(u23)), ex2(hi2(u23))), kOne);
                        const f2 s01 = pack2(rcp(lo2(g01)), rcp(hi2(g01))), s23 = pack2(rcp(lo2(g23)), rcp(hi2(g23)));
                        Sb[it * TPB + tid] = as_float4(f4{s01, s23});
                        S2 = add2(S2, add2(s01, s23));
                    }
                }
            };
            if (!heavy) body(std::integral_constant<int, 0>{});
            else if (!sig) body(std::integral_constant<int, 1>{});
            else if (fastsig) body(std::integral_constant<int, 2>{});
            else body(std::integral_constant<int, 3>{});
            float Ej[4], Rj[4];
            unpack2(E01, Ej[0], Ej[1]); unpack2(E23, Ej[2], Ej[3]);
            unpack2(R01, Rj[0], Rj[1]); unpack2(R23, Rj[2], Rj[3]);
            const float Zt = (Ej[0] + Ej[1]) + (Ej[2] + Ej[3]);
            r16[0] = Zt;
            r16[1] = fmaf(fx0, Zt, fmaf(3.f, Ej[3], fmaf(2.f, Ej[2], Ej[1])));
            r16[2] = fmaf(fty, Zt, Yw);
            if (heavy) {
                r16[3] = hsum2(add2(T01, T23));
                r16[4] = hsum2(S2);
                const float xa0 = fx0 - ax, xa1 = xa0 + 1.f, xa2 = xa0 + 2.f, xa3 = xa0 + 3.f;
                r16[6] = (Rj[0] + Rj[1]) + (Rj[2] + Rj[3]);
                r16[7] = fmaf(xa0, Rj[0], fmaf(xa1, Rj[1], fmaf(xa2, Rj[2], xa3 * Rj[3])));
                r16[8] = Ry;
                r16[9] = fmaf(xa0 * xa0, Rj[0], fmaf(xa1 * xa1, Rj[1], fmaf(xa2 * xa2, Rj[2], fmaf(xa3 * xa3, Rj[3], Ry2))));
                // squared error: sum h^2 everywhere, corrected in the one row that meets the target patch
                float h2 = hsum2(Hq);
                if (hit_it >= 0) {
                    const float4 o = Hs[hit_it * TPB + tid];
                    const float d0 = o.x - thit.x, d1_ = o.y - thit.y, d2 = o.z - thit.z, d3 = o.w - thit.w;
                    h2 += (fmaf(d0, d0, -o.x * o.x) + fmaf(d1_, d1_, -o.y * o.y)) + (fmaf(d2, d2, -o.z * o.z) + fmaf(d3, d3, -o.w * o.w));
                }
                r16[10] = h2;
            }
        }

        // ---- limb partners: one visit each; sums for the overlap ratio, one tie bit per pixel and partner ------
        unsigned words[NIT];                  // one byte per pixel of the float4: (tie pattern over the partners) << 2
#pragma unroll
        for (int q = 0; q < NIT; ++q) words[q] = 0u;
        float mind = INFINITY;                // smallest |own - partner| logit difference seen (0 = a tie)
        float r4[4] = {0.f, 0.f, 0.f, 0.f};
        for (int n = 0; n < nact; ++n) {
            const unsigned q = qn & 1u;
            mbar_wait(qfull + q, (qn >> 1) & 1u);
            const float4* Qs = Qb + q * N4;
            f2 Sj2 = splat2(0.f), M2 = splat2(0.f);
#pragma unroll
            for (int it = 0; it < NIT; ++it) {
                const float4 q4 = Qs[it * TPB + tid];
                const float4 o = Hs[it * TPB + tid];
                const float4 s4 = Sb[it * TPB + tid];
                const f4 hv = as_f4(o), qv = as_f4(q4);
                const f2 u01 = mul2(qv.a, kNL2E), u23 = mul2(qv.b, kNL2E);
                const f2 g01 = add2(pack2(ex2(lo2(u01)), ex2(hi2(u01))), kOne), g23 = add2(pack2(ex2(lo2(u23)), ex2(hi2(u23))), kOne);
                const float sq[4] = {rcp(lo2(g01)), rcp(hi2(g01)), rcp(lo2(g23)), rcp(hi2(g23))};
                const float sk[4] = {s4.x, s4.y, s4.z, s4.w};
                // min(sigma(a), sigma(b)) = sigma(min(a, b)): decide on the logits; equal logits give equal sigmoids
                const f2 d01 = sub2(hv.a, qv.a), d23 = sub2(hv.b, qv.b);
                const float d[4] = {lo2(d01), hi2(d01), lo2(d23), hi2(d23)};
                unsigned tw = 0u;
                float sel[4];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const bool own_smaller = d[jj] < 0.f;
                    sel[jj] = own_smaller ? sk[jj] : sq[jj];
                    if (GRADS && own_smaller) tw |= 4u << (8 * jj);
                }
                if (GRADS) {
                    mind = min3f(mind, fabsf(d[0]), fabsf(d[1]));
                    mind = min3f(mind, fabsf(d[2]), fabsf(d[3]));
                    words[it] |= tw << n;
                }
                Sj2 = add2(Sj2, add2(pack2(sq[0], sq[1]), pack2(sq[2], sq[3])));
                M2 = add2(M2, add2(pack2(sel[0], sel[1]), pack2(sel[2], sel[3])));
            }
            const float Sj = hsum2(Sj2), M = hsum2(M2);
            if (n == 0) { r16[11] = Sj; r16[12] = M; }
            if (n == 1) { r16[13] = Sj; r16[14] = M; }
            if (n == 2) { r4[0] = Sj; r4[1] = M; }
            if (n == 3) { r4[2] = Sj; r4[3] = M; }
            ++qn;
            __syncwarp();
            if (lane == 0) mbar_arrive(qempty + q);          // this warp has consumed the buffer
        }

        // ---- variance tile: its sum only -------------------------------------------------------------------------
        if (has_var) mbar_wait(vfull, i & 1u);        // also when the sum is not needed: the buffer's full / empty phases alternate strictly
        if (has_var && heavy) {
            f2 V2 = splat2(0.f);
#pragma unroll
            for (int it = 0; it < NIT; ++it) {
                const f4 v = as_f4(Vb[it * TPB + tid]);
                V2 = add2(V2, add2(v.a, v.b));
            }
            r16[5] = hsum2(V2);
        }
        if (has_var) {
            __syncwarp();
            if (lane == 0) mbar_arrive(vempty);
        }

        // ---- the one block reduction ------------------------------------------------------------------------------
        float* const redb = red + s * (NW * 16);
        float* const red2b = red2 + s * (NW * 4);
        float* const redMb = redM + s * (NW * 2);
        warp_scatter_sum<16>(r16, lane);
        if ((lane & 1) == 0) redb[warp * 16 + (lane >> 1)] = r16[0];
        if (nact > 2) {
            warp_scatter_sum<4>(r4, lane);
            if ((lane & 7) == 0) red2b[warp * 4 + (lane >> 3)] = r4[0];
        }
        if (lane == 0) { redMb[warp * 2] = mw; redMb[warp * 2 + 1] = mnw; }
        if (heavy && warp == 1) tapw[s * 32 + lane] = tapv;
        compute_barrier();
        float m = redMb[0], hmin = redMb[1];
#pragma unroll
        for (int ww = 1; ww < NW; ++ww) { m = fmaxf(m, redMb[2 * ww]); hmin = fminf(hmin, redMb[2 * ww + 1]); }
        float acc = 0.f;
        {
            const int idx = lane >> 1, q = lane & 1;
#pragma unroll
            for (int t = 0; t < (NW + 1) / 2; ++t) {
                const int ww = q + 2 * t;
                if (ww < NW) {
                    float x = redb[ww * 16 + idx];
                    const float dl = (redMb[2 * ww] - m) * kLog2e;            // (m_w - m) log2 e <= 0
                    if (idx == 3) x = fmaf(dl, redb[ww * 16], x);             // sum e t: t is relative to the warp's maximum too
                    if (idx < 4) x *= ex2(dl);
                    acc += x;
                }
            }
            acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        }
        auto val = [&](int k) -> float { return __shfl_sync(0xffffffffu, acc, k << 1); };
        const float Zs = val(0);
        const float iZ = rcp(Zs);
        const float cx = val(1) * iZ, cy = val(2) * iZ;
        const float ml = m * kLog2e;

        float* const gh = A.grad_hm;
        float4* gh4 = GRADS ? reinterpret_cast<float4*>(gh) + tile_base : nullptr;
        float4* gv4 = (GRADS && has_var) ? reinterpret_cast<float4*>(A.grad_var) + tile_base : nullptr;

        // ---- weight 0: every term carries a factor w -> zero loss and gradient; decode only ------------------------
        if (!heavy) {
            if (tid == 0) {
                float4* p = reinterpret_cast<float4*>(A.partial + (size_t)tile * 8);
                p[0] = z4; p[1] = z4;
                decs[s] = make_float4(cx, cy, m, 0.f);         // the producer warp finishes the decode
            }
            if (GRADS) {
#pragma unroll
                for (int it = 0; it < NIT; ++it) {
                    stg_stream(gh4 + it * TPB, z4);
                    if (gv4) stg_stream(gv4 + it * TPB, z4);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(hempty + s);
            continue;
        }

        // ---- per-tile scalars, derived by every warp from the same sums in the same order --------------------------
        const float ET = val(3), Ssum = val(4), Vsum = val(5), Rm = val(6), Rxa = val(7), Rya = val(8), M2a = val(9), mse_sum = val(10);
        const float mV = has_var ? Vsum * P.inv_n : P.sigma;
        const float ka = P.use_target_weight ? wa * iD : 1.f / (float)(P.B * P.K), kb = w * iD;
        // relu moments shifted from the tile centre to (cx, cy)
        const float dcx = cx - ax, dcy = cy - ay;
        const float sh1 = dcx * Rxa, sh2 = dcy * Rya, dd = dcx * dcx + dcy * dcy;
        float M2c = M2a - 2.f * (sh1 + sh2) + dd * Rm;
        float sXc = Rxa - dcx * Rm, sYc = Rya - dcy * Rm;
        const float shift_abs = M2a + 2.f * (fabsf(sh1) + fabsf(sh2)) + dd * Rm;
        // entropy: E = ln Z - ln2 * sum p t - N eps ; sum p a = E - 1 + N eps
        const float tbar = ET * iZ;
        float Ent = kLn2 * (lg2(Zs) - tbar) - (float)N * kEps;
        const float rmax = kEps * Zs * ex2((m - hmin) * kLog2e);
        const bool flat = (rmax <= kFlatRmax) && (shift_abs <= kShiftCond * M2c || shift_abs == 0.f);
        float pa = Ent - 1.f + (float)N * kEps;
        if (!flat) {
            // ---- general second pass: entropy sums with the per-pixel log, relu moments about (cx, cy) -----------------
            float r8[8];
            {
                const f2 kIZ = splat2(iZ), kEps2 = splat2(kEps), kNML = splat2(-ml);
                f2 A1 = splat2(0.f), A2 = splat2(0.f), R01 = splat2(0.f), R23 = splat2(0.f);
                float Ry = 0.f, Ry2 = 0.f;
                const float dy0 = fty - cy;
#pragma unroll 1
                for (int it = 0; it < NIT; ++it) {
                    const float4 o = Hs[it * TPB + tid];
                    const f4 hv = as_f4(o);
                    const f2 t01 = fma2(hv.a, kL2E, kNML), t23 = fma2(hv.b, kL2E, kNML);
                    const f2 p01 = mul2(pack2(ex2(lo2(t01)), ex2(hi2(t01))), kIZ), p23 = mul2(pack2(ex2(lo2(t23)), ex2(hi2(t23))), kIZ);
                    const f2 u01 = add2(p01, kEps2), u23 = add2(p23, kEps2);
                    const f2 l01 = pack2(lg2(lo2(u01)), lg2(hi2(u01))), l23 = pack2(lg2(lo2(u23)), lg2(hi2(u23)));
                    const f2 c01 = pack2(rcp(lo2(u01)), rcp(hi2(u01))), c23 = pack2(rcp(lo2(u23)), rcp(hi2(u23)));
                    A1 = fma2(p01, l01, A1); A1 = fma2(p23, l23, A1);
                    A2 = fma2(p01, mul2(p01, c01), A2); A2 = fma2(p23, mul2(p23, c23), A2);
                    const f2 r01 = pack2(fmaxf(o.x, 0.f), fmaxf(o.y, 0.f)), r23 = pack2(fmaxf(o.z, 0.f), fmaxf(o.w, 0.f));
                    R01 = add2(R01, r01); R23 = add2(R23, r23);
                    const float rs = hsum2(add2(r01, r23));
                    const float dy = dy0 + (float)(it * ROWS);
                    Ry = fmaf(dy, rs, Ry);
                    Ry2 = fmaf(dy * dy, rs, Ry2);
                }
                float Rj[4];
                unpack2(R01, Rj[0], Rj[1]); unpack2(R23, Rj[2], Rj[3]);
                float dxj[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) dxj[j] = (fx0 + (float)j) - cx;
                r8[0] = hsum2(A1); r8[1] = hsum2(A2);
                r8[2] = fmaf(dxj[0] * dxj[0], Rj[0], fmaf(dxj[1] * dxj[1], Rj[1], fmaf(dxj[2] * dxj[2], Rj[2], fmaf(dxj[3] * dxj[3], Rj[3], Ry2))));
                r8[3] = fmaf(dxj[0], Rj[0], fmaf(dxj[1], Rj[1], fmaf(dxj[2], Rj[2], dxj[3] * Rj[3])));
                r8[4] = Ry; r8[5] = 0.f; r8[6] = 0.f; r8[7] = 0.f;
            }
            warp_scatter_sum<8>(r8, lane);
            if ((lane & 3) == 0) redC[warp * 8 + (lane >> 2)] = r8[0];
            compute_barrier();
            float a8 = 0.f;
            {
                const int idx = lane >> 2, q = lane & 3;
#pragma unroll
                for (int t = 0; t < (NW + 3) / 4; ++t) {
                    const int ww = q + 4 * t;
                    if (ww < NW) a8 += redC[ww * 8 + idx];
                }
                a8 += __shfl_xor_sync(0xffffffffu, a8, 1);
                a8 += __shfl_xor_sync(0xffffffffu, a8, 2);
            }
            const float A1s = __shfl_sync(0xffffffffu, a8, 0), A2s = __shfl_sync(0xffffffffu, a8, 4);
            M2c = __shfl_sync(0xffffffffu, a8, 8);
            sXc = __shfl_sync(0xffffffffu, a8, 12);
            sYc = __shfl_sync(0xffffffffu, a8, 16);
            Ent = -kLn2 * A1s;
            pa = Ent - A2s;
        }
        const float iRp = rcp(Rm + kEps);
        const float v = M2c * iRp;
        const float sd = sqrt_approx(v + kEps);
        const float a4 = (P.lam[3] * gscale) * kb * (sd - P.sigma) * rcp(sd);
        const float c4 = a4 * iRp, k4 = -c4 * v;
        const float c1 = (P.lam[0] * gscale) * ka * 2.f * P.inv_n;
        const float c6 = (P.lam[5] * gscale) * kb * 2.f * (Ent - P.e_star);
        // offset term: the eight taps around the soft-argmax
        const Taps tp = taps_setup(cx, cy, H, W);
        float ov[2][4];
        {
            const int bx = (tp.i00 % W) - wx0, by = (tp.i00 / W) - wy0;
            if (bx >= 0 && bx <= 2 && by >= 0 && by <= 2) {
                const float* tw_ = tapw + s * 32 + by * 4 + bx;
#pragma unroll
                for (int c = 0; c < 2; ++c) { ov[c][0] = tw_[c * 16]; ov[c][1] = tw_[c * 16 + 1]; ov[c][2] = tw_[c * 16 + 4]; ov[c][3] = tw_[c * 16 + 5]; }
            } else {
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    ov[c][0] = __ldg(off_tile + c * N + tp.i00); ov[c][1] = __ldg(off_tile + c * N + tp.i01);
                    ov[c][2] = __ldg(off_tile + c * N + tp.i10); ov[c][3] = __ldg(off_tile + c * N + tp.i11);
                }
            }
        }
        float sl1 = 0.f, sl1p[2], dsdx[2], dsdy[2];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            ov[c][1] *= tp.okx; ov[c][2] *= tp.oky; ov[c][3] *= tp.okx * tp.oky;
            const float samp = tp.w00 * ov[c][0] + tp.w01 * ov[c][1] + tp.w10 * ov[c][2] + tp.w11 * ov[c][3];
            dsdx[c] = ((1.f - tp.fy) * (ov[c][1] - ov[c][0]) + tp.fy * (ov[c][3] - ov[c][2])) * tp.inx;
            dsdy[c] = ((1.f - tp.fx) * (ov[c][2] - ov[c][0]) + tp.fx * (ov[c][3] - ov[c][1])) * tp.iny;
            const float d = samp - ((c == 0 ? gx : gy) - (c == 0 ? cx : cy));
            const float ad = fabsf(d);
            sl1 += ad < 1.f ? 0.5f * d * d : ad - 0.5f;
            sl1p[c] = ad < 1.f ? d : (d > 0.f ? 1.f : -1.f);
        }
        const float h2c = (P.lam[1] * gscale) * ka * 0.5f;
        const float fxx = (P.lam[2] * gscale) * ka * 2.f * (cx - gx) + a4 * (-2.f * sXc * iRp) + h2c * (sl1p[0] * (dsdx[0] + 1.f) + sl1p[1] * dsdx[1]);
        const float fyy = (P.lam[2] * gscale) * ka * 2.f * (cy - gy) + a4 * (-2.f * sYc * iRp) + h2c * (sl1p[0] * dsdy[0] + sl1p[1] * (dsdy[1] + 1.f));
        // limb overlap: ratios -> loss numerator and the per-partner gradient scale
        float cj[4] = {0.f, 0.f, 0.f, 0.f};
        float cst = 0.f, pair_loss = 0.f;
        bool g_live = false;
        if (nact > 0) {
            float SjM[4][2];
            SjM[0][0] = val(11); SjM[0][1] = val(12); SjM[1][0] = val(13); SjM[1][1] = val(14);
            SjM[2][0] = SjM[2][1] = SjM[3][0] = SjM[3][1] = 0.f;
            if (nact > 2) {
                float a4s = 0.f;
                const int idx = lane >> 3, q = lane & 7;
#pragma unroll
                for (int t = 0; t < (NW + 7) / 8; ++t) {
                    const int ww = q + 8 * t;
                    if (ww < NW) a4s += red2b[ww * 4 + idx];
                }
                a4s += __shfl_xor_sync(0xffffffffu, a4s, 1);
                a4s += __shfl_xor_sync(0xffffffffu, a4s, 2);
                a4s += __shfl_xor_sync(0xffffffffu, a4s, 4);
                SjM[2][0] = __shfl_sync(0xffffffffu, a4s, 0); SjM[2][1] = __shfl_sync(0xffffffffu, a4s, 8);
                SjM[3][0] = __shfl_sync(0xffffffffu, a4s, 16); SjM[3][1] = __shfl_sync(0xffffffffu, a4s, 24);
            }
            const float4 wj4 = *reinterpret_cast<const float4*>(dsc->wj);
            const float wjv[4] = {wj4.x, wj4.y, wj4.z, wj4.w};
#pragma unroll
            for (int n = 0; n < 4; ++n) {
                if (n < nact) {
                    const float Sj = SjM[n][0], M = SjM[n][1];
                    const float imm = rcp(fminf(Ssum, Sj) + kEps);
                    const float rho = M * imm;
                    if ((pk >> (4 + n)) & 1u) pair_loss += w * wjv[n] * fmaxf(rho - 0.5f, 0.f);
                    if (GRADS && rho > 0.5f) {
                        cj[n] = (P.lam[4] * gscale) * w * wjv[n] * iD5 * imm;
                        cst += cj[n] * rho * tie_rule(Ssum, Sj);
                        g_live = true;
                    }
                }
            }
        }
        // loss numerators of the tile (second stage: finalize_kernel)
        if (tid == 0) {
            const float peak_t = (cx - gx) * (cx - gx) + (cy - gy) * (cy - gy);
            const float var_t = (sd - P.sigma) * (sd - P.sigma) + (has_var ? (mV - P.sigma) * (mV - P.sigma) : 0.f);
            const float shape_t = (Ent - P.e_star) * (Ent - P.e_star);
            float4* p = reinterpret_cast<float4*>(A.partial + (size_t)tile * 8);
            p[0] = make_float4(wa * (mse_sum * P.inv_n), wa * (0.5f * sl1), wa * peak_t, w * var_t);
            p[1] = make_float4(pair_loss, w * shape_t, 0.f, 0.f);
        }

        // the producer warp finishes the decode (local refinement + offset correction) from the soft-argmax
        if (tid == 0) decs[s] = make_float4(cx, cy, m, 0.f);

        if (GRADS) {
            // overlap coefficient of a pixel as a function of its tie pattern: per-warp table
            float* const tab = tabs + warp * 16;
            if (g_live) {
                if (lane < 16) {
                    float g = -cst;
#pragma unroll
                    for (int n = 0; n < 4; ++n) if ((lane >> n) & 1) g += cj[n];
                    tab[lane] = g;
                }
                __syncwarp();
            }
            // ---- pass D: the heatmap gradient -----------------------------------------------------------------------
            // g = c1 (h - t) + p (c6 (a - pa) + (x - cx) Fx + (y - cy) Fy) + [h > 0] (c4 ((x-cx)^2 + (y-cy)^2) + k4) + overlap
            // flat tiles: c6 (a - pa) = c6 ln2 (tbar - t_i)
            float dxj[4], dx2j[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) { dxj[j] = (fx0 + (float)j) - cx; dx2j[j] = dxj[j] * dxj[j]; }
            const float dy0 = fty - cy;
            const f2 kIZ = splat2(iZ), kNML = splat2(-ml), kC1 = splat2(c1), kC4 = splat2(c4), kK4 = splat2(k4);
            const f2 dx2_01 = pack2(dx2j[0], dx2j[1]), dx2_23 = pack2(dx2j[2], dx2j[3]);
            const float gvar = (P.lam[3] * gscale) * kb * 2.f * (mV - P.sigma) * P.inv_n;
            const float4 gv = make_float4(gvar, gvar, gvar, gvar);
            auto pass_d = [&](auto flat_c) {
                constexpr bool FLAT = decltype(flat_c)::value;
                const float kc = c6 * kLn2;
                const float addc = FLAT ? kc * tbar : -c6 * pa;
                const f2 kNKC = splat2(-kc), kNC6 = splat2(-c6), kEps2 = splat2(kEps), kLn2v = splat2(kLn2);
                const f2 base01 = pack2(fmaf(dxj[0], fxx, addc), fmaf(dxj[1], fxx, addc)), base23 = pack2(fmaf(dxj[2], fxx, addc), fmaf(dxj[3], fxx, addc));
#pragma unroll
                for (int it = 0; it < NIT; ++it) {
                    const float4 o = Hs[it * TPB + tid];
                    const f4 hv = as_f4(o);
                    const f2 t01 = fma2(hv.a, kL2E, kNML), t23 = fma2(hv.b, kL2E, kNML);
                    const f2 p01 = mul2(pack2(ex2(lo2(t01)), ex2(hi2(t01))), kIZ), p23 = mul2(pack2(ex2(lo2(t23)), ex2(hi2(t23))), kIZ);
                    const float dy = dy0 + (float)(it * ROWS);
                    const f2 fyd = splat2(dy * fyy), dy2 = splat2(dy * dy);
                    f2 in01 = add2(base01, fyd), in23 = add2(base23, fyd);
                    if (FLAT) {
                        in01 = fma2(t01, kNKC, in01); in23 = fma2(t23, kNKC, in23);
                    } else {
                        // an = ln2 lg2(p + eps) + p / (p + eps) = -a
                        const f2 u01 = add2(p01, kEps2), u23 = add2(p23, kEps2);
                        const f2 l01 = pack2(lg2(lo2(u01)), lg2(hi2(u01))), l23 = pack2(lg2(lo2(u23)), lg2(hi2(u23)));
                        const f2 c01 = pack2(rcp(lo2(u01)), rcp(hi2(u01))), c23 = pack2(rcp(lo2(u23)), rcp(hi2(u23)));
                        const f2 an01 = fma2(kLn2v, l01, mul2(p01, c01)), an23 = fma2(kLn2v, l23, mul2(p23, c23));
                        in01 = fma2(kNC6, an01, in01); in23 = fma2(kNC6, an23, in23);
                    }
                    f2 g01 = mul2(kC1, hv.a), g23 = mul2(kC1, hv.b);
                    if (it == hit_it) {
                        const f4 tv = as_f4(thit);
                        g01 = mul2(kC1, sub2(hv.a, tv.a)); g23 = mul2(kC1, sub2(hv.b, tv.b));
                    }
                    g01 = fma2(p01, in01, g01); g23 = fma2(p23, in23, g23);
                    // relu branch of the variance term: rterm * [h > 0] (the mask is exactly 0 or 1)
                    const f2 rt01 = fma2(kC4, add2(dx2_01, dy2), kK4), rt23 = fma2(kC4, add2(dx2_23, dy2), kK4);
                    const f2 m01 = pack2(o.x > 0.f ? 1.f : 0.f, o.y > 0.f ? 1.f : 0.f), m23 = pack2(o.z > 0.f ? 1.f : 0.f, o.w > 0.f ? 1.f : 0.f);
                    g01 = fma2(rt01, m01, g01); g23 = fma2(rt23, m23, g23);
                    if (g_live) {
                        const f4 sv = as_f4(Sb[it * TPB + tid]);
                        const unsigned wd = words[it];
                        float G[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            G[j] = *reinterpret_cast<const float*>(reinterpret_cast<const char*>(tab) + ((wd >> (8 * j)) & 0xFFu));
                        const f2 kNeg = splat2(-1.f);
                        g01 = fma2(pack2(G[0], G[1]), fma2(mul2(sv.a, kNeg), sv.a, sv.a), g01);
                        g23 = fma2(pack2(G[2], G[3]), fma2(mul2(sv.b, kNeg), sv.b, sv.b), g23);
                    }
                    stg_stream(gh4 + it * TPB, as_float4(f4{g01, g23}));
                    if (gv4) stg_stream(gv4 + it * TPB, gv);
                }
            };
            if (flat) pass_d(std::true_type{}); else pass_d(std::false_type{});

            // rare: a logit of this thread equals its partner's — ATen's minimum splits that gradient evenly.
            // Patch the pixels this thread has just written (same thread, program order).
            if (g_live && mind == 0.f) {
                const int b = tile / P.K;
                const unsigned pj = dsc->pj;
                for (int n = 0; n < nact; ++n) {
                    const float cjp = cj[n == 0 ? 0 : (n == 1 ? 1 : (n == 2 ? 2 : 3))];
                    if (cjp == 0.f) continue;
                    const float4* src = reinterpret_cast<const float4*>(hm) + ((size_t)b * P.K + ((pj >> (8 * n)) & 0xFFu)) * N4 + tid;
                    for (int it = 0; it < NIT; ++it) {
                        const float4 q = ldg_keep(src + it * TPB), o = Hs[it * TPB + tid];
                        if (q.x == o.x || q.y == o.y || q.z == o.z || q.w == o.w) {
                            float4 g = gh4[it * TPB];
                            float sg;
                            if (q.x == o.x) { sg = sigmoid_fast(o.x); g.x = fmaf(0.5f * cjp * sg, 1.f - sg, g.x); }
                            if (q.y == o.y) { sg = sigmoid_fast(o.y); g.y = fmaf(0.5f * cjp * sg, 1.f - sg, g.y); }
                            if (q.z == o.z) { sg = sigmoid_fast(o.z); g.z = fmaf(0.5f * cjp * sg, 1.f - sg, g.z); }
                            if (q.w == o.w) { sg = sigmoid_fast(o.w); g.w = fmaf(0.5f * cjp * sg, 1.f - sg, g.w); }
                            gh4[it * TPB] = g;
                        }
                    }
                }
            }
            // the (up to) four non-zero taps per channel of the offset gradient; the zero fill of these addresses was
            // issued before the block barrier, so it is ordered before these stores
            if (tid == 0) {
                float* go = A.grad_off + (size_t)tile * 2 * N;
#pragma unroll
                for (int ch = 0; ch < 2; ++ch) {
                    float* o = go + ch * N;
                    const float gc = h2c * sl1p[ch];
                    o[tp.i00] = gc * tp.w00;
                    if (tp.okx != 0.f) o[tp.i01] = gc * tp.w01;
                    if (tp.oky != 0.f) o[tp.i10] = gc * tp.w10;
                    if (tp.okx != 0.f && tp.oky != 0.f) o[tp.i11] = gc * tp.w11;
                }
            }
        }

        __syncwarp();
        if (lane == 0) mbar_arrive(hempty + s);
    }
}

// ---- launcher ------------------------------------------------------------------------------------------
template <int W4, int ROWS, int NIT, int MINB, bool GRADS>
static int launch_step_t(const LossParams& P, const LossArgs& A, cudaStream_t s, cudaEvent_t e0, cudaEvent_t e1) {
    using L = StepPlan<W4, ROWS, NIT>;
    const size_t smem = (size_t)L::oLut + (size_t)((P.ec.lut_size + 3) & ~3) * 4;
    auto kern = step_tile_kernel<W4, ROWS, NIT, MINB, GRADS>;
    int dev = 0, sms = 0, max_optin = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (max_optin > 0 && smem > (size_t)max_optin) return 1;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return fail(GBCODEC_ERR_CUDA, "cudaFuncSetAttribute(step_tile_kernel): %s", cudaGetErrorString(e));
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, L::TPB + 32, smem);
    if (e != cudaSuccess || per_sm <= 0) return fail(GBCODEC_ERR_CUDA, "step_tile_kernel does not fit an SM (%zu bytes of shared memory)", smem);
    const int tiles = P.B * P.K;
    const int grid = tiles < sms * per_sm ? tiles : sms * per_sm;
    if (e0) cudaEventRecord(e0, s);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(L::TPB + 32);                  // compute warps + the producer warp
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    note_launch();
    e = cudaLaunchKernelEx(&cfg, kern, P, A);
    if (e != cudaSuccess) return fail(GBCODEC_ERR_CUDA, "cudaLaunchKernelEx(step_tile_kernel): %s", cudaGetErrorString(e));
    if (e1) cudaEventRecord(e1, s);
    return check_launch("step_tile_kernel");
}

// GBCODEC_STEP_KERNEL=tile selects the one-CTA-per-tile kernel of loss_tile.cu (A/B measurements, and tests that compare
// the float32 entry point bit for bit with the float16 / per-tile-mean entry points, which only that kernel serves).
// Read on every call: a test sets it around single calls.
static bool step_disabled() {
    const char* e = getenv("GBCODEC_STEP_KERNEL");
    return e && !strcmp(e, "tile");
}

int launch_step_tile(const LossParams& P, const LossArgs& A, cudaStream_t s, cudaEvent_t e0, cudaEvent_t e1) {
    // float32 maps, target generated on the fly with a patch no taller than the CTA's rows, the ordinary forward (+ backward) call
    if (step_disabled() || A.half_io || A.target || A.lam_eff || A.plan || A.var_mean || A.grad_var_mean || !A.desc || !A.tile_counter) return 1;
    if (A.coords && A.radius > 8) return 1;
    const bool grads = A.grad_hm != nullptr;
    if (P.H == 64 && P.W == 48 && P.ec.ntap <= 16) {
        return grads ? launch_step_t<12, 16, 4, 3, true>(P, A, s, e0, e1) : launch_step_t<12, 16, 4, 3, false>(P, A, s, e0, e1);
    }
    return 1;
}

}  // namespace gbc
