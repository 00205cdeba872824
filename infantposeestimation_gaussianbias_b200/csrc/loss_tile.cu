// loss_tile.cu — tile kernel of the six-term fusion loss (forward + backward
// + optional keypoint decode in one pass).  Same arithmetic as the generic kernel in loss.cu
// (FusionPoseLoss.forward, models/fusion_head.py:745-806, terms :637-743 and :405-559, and the
// autograd backward of train.py:182 in closed form); different schedule:
//
//   * one CTA per (image, keypoint) tile, TPB = (W/4) * ROWS threads; a thread owns the same four
//     columns in every row it visits, so every x-dependent factor is a per-thread constant and
//     the column moments factor out of the row loop;
//   * the tile is copied once (cp.async, 16 B per thread and row) into thread-private shared-memory
//     slots and every pass reads it from there with rolled row loops (the shipped instantiations;
//     ROLL = false keeps it in registers with unrolled loops); the per-pixel intermediates a later
//     pass needs (sigmoid, entropy derivative) are parked in slots of the same kind — no
//     cross-thread traffic, conflict-free 128-bit accesses; the variance tile, needed for its sum
//     only, takes the same road through slots that are still unused at that time;
//   * two block reductions per tile, ONE barrier each (halving butterfly inside the warp,
//     then every warp finishes the cross-warp sum redundantly); the tile maximum has no
//     reduction of its own: the softmax moments are accumulated relative to each warp's
//     maximum and rescaled where the warps' partial sums meet;
//   * the limb partners' tiles (some other CTA's own tile: L2 hits) stream through thread-private
//     slots with cp.async — a whole tile where shared memory has room for it (the next partner's
//     copy in flight while the current one is consumed), a two-row ring where it has not; each is
//     visited once, the per-pixel tie pattern the gradient needs is kept as one bit per pixel and
//     partner, and the gradient pass turns the 4-bit pattern of a pixel into its overlap
//     coefficient with one table look-up;
//   * stores are spread over the kernel's lifetime: the (almost all zero) offset-gradient tile
//     leaves right after the loads are issued, the uniform variance-gradient tile once the
//     first reduction is known, the heatmap gradient at the end;
//   * the decode tail (window softmax, bilinear offset read) is software-pipelined through the
//     passes in warp 0 so its dependent L2 round trips hide behind the other warps' work.
//
// Algorithmic HBM bytes per tile: read hm, var (8N); write d_hm, d_var, d_off (16N).
#include "loss_common.cuh"
#include "f32x2.cuh"
#include <cuda_fp16.h>
#include <stdlib.h>

namespace gbc {

// ---- one-barrier block reductions -------------------------------------------------------
// Warp stage: warp_scatter_sum (common.cuh).
template <int NV> struct Log2 { static constexpr int value = 1 + Log2<NV / 2>::value; };
template <> struct Log2<1> { static constexpr int value = 0; };

// Block-wide sums of NV values; `red` holds NW*NV floats and must not be the buffer of the
// previous reduction (the callers alternate two buffers).  Fixed order: deterministic, and
// every thread ends with the same bits.  Returns the lane-distributed totals: value k sits in
// lane k << (5 - log2 NV) of every warp (fetch it with __shfl_sync).
template <int NV, int NW>
__device__ __forceinline__ float block_sum1(float (&v)[NV], float* red) {
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
    return acc;
}
// The same sums when the first NSC values were accumulated relative to each WARP's own maximum (softmax numerators
// exp(h - m_warp)): the per-warp maxima travel with the partials, every thread takes their maximum after the one
// barrier and rescales the first NSC values of each warp by exp(m_warp - m).  `m` enters as the warp's maximum and
// leaves as the tile's (the raw value: it is also the decode score).
template <int NV, int NW, int NSC>
__device__ __forceinline__ float block_sum1_rescaled(float (&v)[NV], float& m, float* red, float* redm) {
    constexpr int SH = 5 - Log2<NV>::value, SUB = 32 / NV;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    warp_scatter_sum<NV>(v);
    const int idx = lane >> SH, q = lane & (SUB - 1);
    if (q == 0) red[warp * NV + idx] = v[0];
    if (lane == 0) redm[warp] = m;
    __syncthreads();
    float mm = redm[0];
#pragma unroll
    for (int ww = 1; ww < NW; ++ww) mm = fmaxf(mm, redm[ww]);
    const float mml = mm * kLog2e;
    float acc = 0.f;
#pragma unroll
    for (int t = 0; t < (NW + SUB - 1) / SUB; ++t) {
        const int ww = q + t * SUB;
        if (ww < NW) {
            const float x = red[ww * NV + idx];
            acc += idx < NSC ? x * ex2(fmaf(redm[ww], kLog2e, -mml)) : x;
        }
    }
#pragma unroll
    for (int o = SUB / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    m = mm;
    return acc;
}
template <int NV>
__device__ __forceinline__ float lane_value(float acc, int k) { return __shfl_sync(0xffffffffu, acc, k << (5 - Log2<NV>::value)); }

template <int NW>
__device__ __forceinline__ float block_max1(float m, float* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    m = warp_max(m);
    if (lane == 0) red[warp] = m;
    __syncthreads();
    return warp_max(lane < NW ? red[lane] : -INFINITY);
}

// ---- small helpers -------------------------------------------------------------------------
__device__ __forceinline__ float fsqrt_fast(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_but_one() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
// bit i of the low byte -> bit 4*i
__device__ __forceinline__ unsigned spread8(unsigned t) {
    unsigned x = t & 0xFFu;
    x = (x | (x << 12)) & 0x000F000Fu;
    x = (x | (x << 6)) & 0x03030303u;
    x = (x | (x << 3)) & 0x11111111u;
    return x;
}
template <typename T>
__device__ __forceinline__ T pick4(int i, T a, T b, T c, T d) { return i == 0 ? a : (i == 1 ? b : (i == 2 ? c : d)); }

// Out-of-line copy of the un-staged decode tail for the rare paths (weight-0 tiles, radius > 2),
// so that the hot path's code stays small.
template <typename T>
__device__ __noinline__ void refine_and_correct_cold(const T* hm_tile, const T* off_tile, const float* alpha_param,
                                                     const float* fusion_weight, int H, int W, int radius, unsigned flags,
                                                     float* cx, float* cy) {
    int px, py;
    refine_and_correct<T>(hm_tile, nullptr, off_tile, alpha_param, fusion_weight, H, W, radius, flags, *cx, *cy, px, py);
}

// ---- element type of the maps -------------------------------------------------------------------
// float32, or float16 under autocast (train.py:171).  Half maps are up-cast value by value where they enter
// (registers / the shared-memory slots hold float32 either way), gradients are rounded once where they leave;
// everything in between is the same code.  A "vector" is four pixels: 16 bytes of float, 8 bytes of half.
__device__ __forceinline__ float4 half4_to_float4(const uint2& r) {
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&r.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
    return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ uint2 float4_to_half4(const float4& v) {
    const __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
    uint2 r;
    r.x = *reinterpret_cast<const unsigned*>(&a); r.y = *reinterpret_cast<const unsigned*>(&b);
    return r;
}
template <bool HALF> struct TileIO;
template <> struct TileIO<false> {
    using Vec = float4;
    using Elem = float;
    static __device__ __forceinline__ float4 load_stream(const Vec* p) { return ldg_stream(p); }
    static __device__ __forceinline__ float4 load_keep(const Vec* p) { return ldg_keep(p); }
    static __device__ __forceinline__ float4 load_plain(const Vec* p) { return *p; }
    static __device__ __forceinline__ void store_stream(Vec* p, const float4& v) { stg_stream(p, v); }
    static __device__ __forceinline__ void store_plain(Vec* p, const float4& v) { *p = v; }
    static __device__ __forceinline__ void copy_async(float4* slot, const Vec* g) { cp_async16(slot, g); }
    static __device__ __forceinline__ float4 from_slot(const float4* slot) { return *slot; }
    static __device__ __forceinline__ void st1(Elem* p, float v) { *p = v; }
};
template <> struct TileIO<true> {
    using Vec = uint2;
    using Elem = __half;
    static __device__ __forceinline__ float4 load_stream(const Vec* p) {
        uint2 r;
        asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
        return half4_to_float4(r);
    }
    static __device__ __forceinline__ float4 load_keep(const Vec* p) { return half4_to_float4(__ldg(p)); }
    static __device__ __forceinline__ float4 load_plain(const Vec* p) { return half4_to_float4(*p); }
    static __device__ __forceinline__ void store_stream(Vec* p, const float4& v) {
        const uint2 r = float4_to_half4(v);
        asm volatile("st.global.cs.v2.u32 [%0], {%1,%2};" :: "l"(p), "r"(r.x), "r"(r.y) : "memory");
    }
    static __device__ __forceinline__ void store_plain(Vec* p, const float4& v) { *p = float4_to_half4(v); }
    // the raw halves land in the first 8 bytes of the thread's 16-byte slot; from_slot up-casts them
    static __device__ __forceinline__ void copy_async(float4* slot, const Vec* g) {
        const unsigned sa = (unsigned)__cvta_generic_to_shared(slot);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(sa), "l"(g) : "memory");
    }
    static __device__ __forceinline__ float4 from_slot(const float4* slot) { return half4_to_float4(*reinterpret_cast<const uint2*>(slot)); }
    static __device__ __forceinline__ void st1(Elem* p, float v) { *p = __float2half_rn(v); }
};

// Target modes: where the target tile comes from
constexpr int kTargetOneHit = 0;   // generated on the fly; the patch is no taller than ROWS, so a thread meets it in at most one row
constexpr int kTargetGlobal = 1;   // read from HBM (stand-alone loss with d_target)
constexpr int kTargetLut = 2;      // generated on the fly, any patch height

// ---- the kernel ---------------------------------------------------------------------------------
// ROLL = false: the tile stays in registers and the row loops are fully unrolled (most ILP, ~80 registers, large code).
// ROLL = true : the tile is parked in a thread-private shared-memory slot as well and the row loops stay rolled
//               (4x smaller loop code, <= 64 registers -> one more CTA per SM).
// The tile's work is a device function: the forward kernel runs it once per CTA (grid = (K, B)), the backward kernel
// (loss_tile_backward_kernel below) walks a few tiles per CTA so that the common case — nothing to recompute — costs a
// small grid of CTAs that read one word and leave, not one CTA per tile.
// MG = true : the tile maximum is not reduced on its own: pass B runs relative to each warp's maximum and the moments
//              are rescaled where the warps' sums meet (one barrier and one sweep over the tile less); the squared
//              error then waits for pass C, since the target's exp table is published by that same barrier.
template <int W4, int ROWS, int NIT, bool CE, bool CS, bool CA, bool CQ, bool ROLL, int TM, bool HALF, bool MG = false, bool VSLOT = false>
__device__ __forceinline__ void loss_tile_body(const LossParams& P, const LossArgs& A, const int k, const int b, const int kdim) {
    using IO = TileIO<HALF>;
    using Vec = typename IO::Vec;
    using Elem = typename IO::Elem;
    constexpr int TPB = W4 * ROWS, NW = TPB / 32, N4 = TPB * NIT, N = 4 * N4, W = 4 * W4, H = ROWS * NIT;
    static_assert(TPB % 32 == 0 && TPB <= 1024, "CTA must be whole warps");      // NW <= 32: redm holds 32 floats
    static_assert(GBCODEC_MAX_PARTNERS == 4, "tie patterns are nibbles");

    extern __shared__ __align__(16) float smem[];
    constexpr int UNR = ROLL ? 1 : NIT;
    constexpr bool AQ = CQ && !CA;                  // the partner slot is free once the partners are done: park `a` there
    // QRING: no room for a whole partner tile -> its rows stream through a two-row ring of thread-private slots, each
    // row requested (cp.async) two rows before it is consumed, the next partner's first rows during the last two
    constexpr bool QRING = !CQ && ROLL;
    static_assert(!QRING || NIT % 2 == 0, "the ring's slot parity must carry over from one partner to the next");
    float4* Qs = reinterpret_cast<float4*>(smem);                 // partner tile, thread-private slots
    float4* Hs = Qs + (CQ ? N4 : (QRING ? 2 * TPB : 0));          // own tile (ROLL only)
    float4* Es = Hs + (ROLL ? N4 : 0);                            // exp(h - max), later softmax weight p
    float4* Ss = Es + (CE ? N4 : 0);                              // sigmoid(h)
    float4* As = Ss + (CS ? N4 : 0);                              // a = -log(p + eps) - p / (p + eps)
    float* red0 = reinterpret_cast<float*>(As + (CA ? N4 : 0));   // two reduction buffers of NW * 16 floats
    float* red1 = red0 + NW * 16;
    float* lutG = red1 + NW * 16;                                 // 12 tile coefficients + 16 overlap coefficients + 4 partner scales
    float* redm = lutG + 32;                                      // per-warp maxima (MG), one float per warp of the largest CTA
    unsigned* Ws = reinterpret_cast<unsigned*>(redm + 32);        // tie-pattern words (ROLL only), thread-private
    float* lut = reinterpret_cast<float*>(Ws + (ROLL ? N4 : 0));  // exp table of the target patch

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = b * kdim + k;

    // ---- bulk loads first: the tile into registers, the variance map into its sum ----------------
    // (unconditional: whether the tile carries weight is only known one L2 round trip later)
    const size_t toff = (size_t)tile * N4 + tid;
    const Vec* hmb = reinterpret_cast<const Vec*>(A.hm) + tid;
    const Vec* hm4 = reinterpret_cast<const Vec*>(A.hm) + toff;
    // VS: the variance tile is needed for its sum only; it rides through the (still unused) sigmoid slots as an
    // asynchronous copy and is added up by pass B just before a slot receives its sigmoid — no registers held
    // across the wait for HBM, no second wait
    constexpr bool VS = ROLL && CS && VSLOT;
    // no sigmoid slots, but a partner ring: the variance rows go through the ring ahead of the first partner's
    constexpr bool VRING = !CQ && ROLL && !VS;
    const bool vring = VRING && A.var != nullptr;
    const Vec* var4r = reinterpret_cast<const Vec*>(A.var) + toff;
    float4 h[NIT];
    if (ROLL) {
#pragma unroll
        for (int it = 0; it < NIT; ++it) IO::copy_async(Hs + it * TPB + tid, hm4 + it * TPB);
        if (VS && A.var) {
            const Vec* var4 = reinterpret_cast<const Vec*>(A.var) + toff;
#pragma unroll
            for (int it = 0; it < NIT; ++it) IO::copy_async(Ss + it * TPB + tid, var4 + it * TPB);
        }
        cp_async_commit();
        if (vring) {
            IO::copy_async(Qs + tid, var4r);
            cp_async_commit();
            IO::copy_async(Qs + TPB + tid, var4r + TPB);
            cp_async_commit();
        }
    } else {
#pragma unroll
        for (int it = 0; it < NIT; ++it) h[it] = IO::load_stream(hm4 + it * TPB);
    }
    auto own4 = [&](int it) -> float4 { return ROLL ? Hs[it * TPB + tid] : h[it]; };
    float4 vv[NIT];
    if (A.var && !VS && !vring) {
        const Vec* var4 = reinterpret_cast<const Vec*>(A.var) + toff;
#pragma unroll
        for (int it = 0; it < NIT; ++it) vv[it] = IO::load_stream(var4 + it * TPB);
    }
    // then every scalar the tile will need.  The forward kernel is launched programmatically dependent on the kernel that
    // writes them (the bulk copies above are already on their way); the finalize kernel may be scheduled from here on.
    pdl_wait();
    pdl_launch_dependents();
    const float w = __ldg(A.weff + tile);
    const int np = P.n_partner[k];
    const int4 pj4 = make_int4(P.partner[k][0], P.partner[k][1], P.partner[k][2], P.partner[k][3]);
    float wjv[4];
#pragma unroll
    for (int pi = 0; pi < 4; ++pi) wjv[pi] = pi < np ? __ldg(A.weff + b * P.K + pick4(pi, pj4.x, pj4.y, pj4.z, pj4.w)) : 0.f;
    int4 gq = make_int4(0, 0, 0, 0);
    if (TM != kTargetGlobal) gq = __ldg(A.geom + tile);

    const bool grads = A.grad_hm != nullptr;
    const bool backward_only = A.lam_eff != nullptr;
    const bool decode = A.coords != nullptr;
    Vec* gh4 = reinterpret_cast<Vec*>(A.grad_hm) + toff;
    Vec* gv4 = (grads && A.grad_var) ? reinterpret_cast<Vec*>(A.grad_var) + toff : nullptr;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    // the offset gradient is zero except on (up to) four taps per channel, patched at the end
    if (grads) {
        Vec* go4 = reinterpret_cast<Vec*>(A.grad_off) + (size_t)tile * 2 * N4 + tid;
#pragma unroll
        for (int it = 0; it < 2 * NIT; ++it) IO::store_stream(go4 + it * TPB, z4);
    }

    const int tx = tid % W4, ty = tid / W4;
    const int x0 = tx << 2;
    const float fx0 = (float)x0, fty = (float)ty;
    const float wa = P.use_target_weight ? w : 1.f;
    const bool heavy = (w != 0.f) || !P.use_target_weight;
    float vsum = 0.f;
    if (A.var && !VS && !vring) {
#pragma unroll
        for (int it = 0; it < NIT; ++it) vsum += (vv[it].x + vv[it].y) + (vv[it].z + vv[it].w);
    }
    // active partners (both weights non-zero), as a 4-bit mask; CTA-uniform
    unsigned act = 0;
#pragma unroll
    for (int pi = 0; pi < 4; ++pi) if (w != 0.f && wjv[pi] != 0.f) act |= 1u << pi;

    // on-the-fly target: patch geometry (from the weights pre-kernel) and the exp table
    PatchGeom geom = PatchGeom{};
    bool cols_hit = false;
    int pcx = 0, pcy = 0;
    if (TM != kTargetGlobal && heavy) {
        geom = unpack_geom(gq, w);
        cols_hit = geom.active && x0 + 3 >= geom.x_from && x0 < geom.x_to;
        pcx = geom.ulx + (int)P.ec.centre; pcy = geom.uly + (int)P.ec.centre;
        if (geom.active) fill_patch_lut(lut, P.ec);
    }
    auto patch_row = [&](int y) -> float4 {
        const int dy2 = (y - pcy) * (y - pcy);
        float e[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int xx = x0 + j, dx = xx - pcx;
            e[j] = (xx >= geom.x_from && xx < geom.x_to) ? lut[dx * dx + dy2] : 0.f;
        }
        return make_float4(e[0], e[1], e[2], e[3]);
    };

    // ---- reduction 1: tile maximum -----------------------------------------------------------
    float m = -INFINITY;
    if (ROLL) {
        cp_async_wait_all();                  // own slots only: no barrier needed
        if (HALF) {
#pragma unroll
            for (int it = 0; it < NIT; ++it) Hs[it * TPB + tid] = IO::from_slot(Hs + it * TPB + tid);
        }
    }
#pragma unroll UNR
    for (int it = 0; it < NIT; ++it) { const float4 o = own4(it); m = fmaxf(m, fmaxf(fmaxf(o.x, o.y), fmaxf(o.z, o.w))); }
    static_assert(!(MG && CE), "the parked numerators of pass B would be relative to the warp's maximum");
    if (MG) m = warp_max(m);                 // this warp's maximum for now
    else m = block_max1<NW>(m, red0);        // the barrier also publishes the exp table
    float ml = m * kLog2e;

    // first active partner's tile -> thread-private smem slots; it has all of pass B to arrive
    int cur = act ? __ffs(act) - 1 : -1;
    if (CQ && cur >= 0) {
        const Vec* src = hmb + ((size_t)b * P.K + pick4(cur, pj4.x, pj4.y, pj4.z, pj4.w)) * N4;
#pragma unroll
        for (int it = 0; it < NIT; ++it) IO::copy_async(Qs + it * TPB + tid, src + it * TPB);
        cp_async_commit();
    }
    const Vec* src_first = hmb + ((size_t)b * P.K + pick4(cur < 0 ? 0 : cur, pj4.x, pj4.y, pj4.z, pj4.w)) * N4;
    if (QRING && cur >= 0 && !vring) {
        const Vec* src = hmb + ((size_t)b * P.K + pick4(cur, pj4.x, pj4.y, pj4.z, pj4.w)) * N4;
        IO::copy_async(Qs + tid, src);
        cp_async_commit();
        IO::copy_async(Qs + TPB + tid, src + TPB);
        cp_async_commit();
    }

    // kTargetOneHit: the one row (if any) in which this thread meets the patch
    int hit_it = -1;
    float4 thit = z4;
    auto find_hit = [&]() {
        if (TM == kTargetOneHit && cols_hit) {
            const int it0 = max(0, (geom.y_from - ty + ROWS - 1) / ROWS);
            const int y = it0 * ROWS + ty;
            if (it0 < NIT && y < geom.y_to) { hit_it = it0; thit = patch_row(y); }
        }
    };
    if (!MG) find_hit();
    auto target4 = [&](int it) -> float4 {
        if (TM == kTargetGlobal) return ldg_keep(reinterpret_cast<const float4*>(A.target) + (size_t)tile * N4 + tid + it * TPB);
        if (TM == kTargetOneHit) {
            const bool on = it == hit_it;
            return make_float4(on ? thit.x : 0.f, on ? thit.y : 0.f, on ? thit.z : 0.f, on ? thit.w : 0.f);
        }
        const int y = it * ROWS + ty;
        return (cols_hit && y >= geom.y_from && y < geom.y_to) ? patch_row(y) : z4;
    };

    // ---- pass B: softmax moments, sigmoid mass, squared error ------------------------------------
    // Pixel arithmetic runs on packed pairs (f32x2.cuh): (x, y) and (z, w) of the float4.
    const f2 kL2E = splat2(kLog2e), kNL2E = splat2(-kLog2e), kOne = splat2(1.f);
    f2 kNML = splat2(-ml);
    float r8[8];
    {
        f2 E01 = splat2(0.f), E23 = splat2(0.f), S2 = splat2(0.f), mse2 = splat2(0.f);
        float Yw = 0.f;
#pragma unroll UNR
        for (int it = 0; it < NIT; ++it) {
            const f4 hv = as_f4(own4(it));
            const f2 t01 = fma2(hv.a, kL2E, kNML), t23 = fma2(hv.b, kL2E, kNML);
            const f2 e01 = pack2(ex2(lo2(t01)), ex2(hi2(t01))), e23 = pack2(ex2(lo2(t23)), ex2(hi2(t23)));
            E01 = add2(E01, e01); E23 = add2(E23, e23);
            if (CE) Es[it * TPB + tid] = as_float4(f4{e01, e23});
            Yw = fmaf((float)(it * ROWS), hsum2(add2(e01, e23)), Yw);
            if (VS && A.var) { const float4 v4 = IO::from_slot(Ss + it * TPB + tid); vsum += (v4.x + v4.y) + (v4.z + v4.w); }
            if (vring) {
                // the ring's row `it` is a variance row; its slot then takes variance row it + 2 or, in the last two
                // rounds, the first partner's rows 0 and 1
                cp_async_wait_but_one();
                const float4 v4 = IO::from_slot(Qs + (it & 1) * TPB + tid);
                vsum += (v4.x + v4.y) + (v4.z + v4.w);
                if (it + 2 < NIT) IO::copy_async(Qs + (it & 1) * TPB + tid, var4r + (it + 2) * TPB);
                else if (cur >= 0) IO::copy_async(Qs + (it & 1) * TPB + tid, src_first + (it + 2 - NIT) * TPB);
                cp_async_commit();
            }
            if (heavy) {
                const f2 u01 = mul2(hv.a, kNL2E), u23 = mul2(hv.b, kNL2E);
                const f2 g01 = add2(pack2(ex2(lo2(u01)), ex2(hi2(u01))), kOne), g23 = add2(pack2(ex2(lo2(u23)), ex2(hi2(u23))), kOne);
                const f2 s01 = pack2(rcp(lo2(g01)), rcp(hi2(g01))), s23 = pack2(rcp(lo2(g23)), rcp(hi2(g23)));
                if (CS) Ss[it * TPB + tid] = as_float4(f4{s01, s23});
                S2 = add2(S2, add2(s01, s23));
                if (!MG) {
                    const f4 tv = as_f4(target4(it));
                    const f2 d01 = sub2(hv.a, tv.a), d23 = sub2(hv.b, tv.b);
                    mse2 = fma2(d01, d01, mse2);
                    mse2 = fma2(d23, d23, mse2);
                }
            }
        }
        float Ej[4];
        unpack2(E01, Ej[0], Ej[1]); unpack2(E23, Ej[2], Ej[3]);
        const float Zt = (Ej[0] + Ej[1]) + (Ej[2] + Ej[3]);
        r8[0] = Zt;
        r8[1] = fmaf(fx0, Zt, fmaf(3.f, Ej[3], fmaf(2.f, Ej[2], Ej[1])));
        r8[2] = fmaf(fty, Zt, Yw);
        r8[3] = hsum2(S2); r8[4] = hsum2(mse2); r8[5] = vsum; r8[6] = 0.f; r8[7] = 0.f;
    }
    float acc8;
    if (MG) {
        acc8 = block_sum1_rescaled<8, NW, 3>(r8, m, red1, redm);     // m: the tile's maximum from here on
        ml = m * kLog2e;
        kNML = splat2(-ml);
        find_hit();
    } else {
        acc8 = block_sum1<8, NW>(r8, red1);
    }
    const float iZ = rcp(lane_value<8>(acc8, 0));
    const float cx = lane_value<8>(acc8, 1) * iZ, cy = lane_value<8>(acc8, 2) * iZ;

    const Elem* hm_tile = reinterpret_cast<const Elem*>(A.hm) + (size_t)tile * N;
    const Elem* off_tile = reinterpret_cast<const Elem*>(A.off) + (size_t)tile * 2 * N;

    // ---- weight 0: every term carries a factor w -> zero loss and gradient; decode only ----------
    if (!heavy) {
        if (decode && tid < 32) {
            float dx_ = cx, dy_ = cy;
            refine_and_correct_cold<Elem>(hm_tile, off_tile, A.alpha_param, A.fusion_weight, H, W, A.radius, A.dflags, &dx_, &dy_);
            if (tid == 0) { A.coords[2 * tile] = dx_; A.coords[2 * tile + 1] = dy_; A.scores[tile] = m; }
        }
        if (tid == 0 && !backward_only) {
            float4* p = reinterpret_cast<float4*>(A.partial + (size_t)tile * 8);
            p[0] = z4; p[1] = z4;
        }
        if (grads) {
#pragma unroll
            for (int it = 0; it < NIT; ++it) {
                IO::store_stream(gh4 + it * TPB, z4);
                if (gv4) IO::store_stream(gv4 + it * TPB, z4);
            }
            if (A.grad_var_mean && tid == 0) A.grad_var_mean[tile] = 0.f;
        }
        return;
    }
    const float Ssum = lane_value<8>(acc8, 3);
    const bool has_var = A.var != nullptr || A.var_mean != nullptr;
    const float mV = A.var ? lane_value<8>(acc8, 5) * P.inv_n : (A.var_mean ? __ldg(A.var_mean + tile) : P.sigma);

    // ---- warp roles for the per-tile scalars (a warp each, the others do not repeat the work) ------------
    //   warp RD: decode tail                       warp RO: offset term (8 taps, SmoothL1, its share of dL/dc)
    //   warp RV: peak / variance / entropy terms    warp RP: limb overlap ratios and the tie-pattern table
    constexpr int RD = 0, RO = 1 % NW, RV = 2 % NW, RP = 3 % NW;
    // loads whose values are first needed after the partner visits: issue now, consume then
    const float gtx = __ldg(A.gt + 2 * tile), gty = __ldg(A.gt + 2 * tile + 1);
    const double sum_w = __ldg(A.sums), sum_p = __ldg(A.sums + 1);
    const float gscale = A.grad_scale ? __ldg(A.grad_scale) : 1.f;
    float lam_eff[6];
    if (backward_only) {
#pragma unroll
        for (int q = 0; q < 6; ++q) lam_eff[q] = __ldg(A.lam_eff + q);
    }
    // offset term: the 8 taps around the soft-argmax (warp RO only)
    float ov[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    if (warp == RO) {
        const Taps t0 = taps_setup(cx, cy, H, W);   // recomputed when the values are consumed: only the 8 loads stay live
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            ov[c][0] = ld1(off_tile + c * N + t0.i00); ov[c][1] = ld1(off_tile + c * N + t0.i01);
            ov[c][2] = ld1(off_tile + c * N + t0.i10); ov[c][3] = ld1(off_tile + c * N + t0.i11);
        }
    }
    // decode stage 1 (warp RD): window taps around the rounded soft-argmax
    const bool staged_decode = decode && (A.dflags & GBCODEC_DECODE_REFINE) && A.radius <= 2;
    float win = -INFINITY, winx = 0.f, winy = 0.f;
    bool win_ok = false;
    if (staged_decode && warp == RD) {
        const int px = (int)fminf(fmaxf(rintf(cx), 0.f), (float)(W - 1));
        const int py = (int)fminf(fmaxf(rintf(cy), 0.f), (float)(H - 1));
        const int S = 2 * A.radius + 1;
        const int x = px - A.radius + lane % S, y = py - A.radius + lane / S;
        win_ok = lane < S * S && x >= 0 && x < W && y >= 0 && y < H;
        winx = (float)x; winy = (float)y;
        if (win_ok) win = ld1(hm_tile + y * W + x);
    }

    float r16[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) r16[q] = 0.f;
    // ---- limb partners: one visit each; sums for the overlap ratio, one tie bit per pixel -------------
    // words[it]: one byte per pixel of the float4, holding (4-bit partner pattern) << 2 — a byte offset into
    // the per-warp coefficient table of pass D
    unsigned words[NIT];
#pragma unroll
    for (int q = 0; q < NIT; ++q) { words[q] = 0u; if (ROLL) Ws[q * TPB + tid] = 0u; }
    float mind = INFINITY;                            // smallest |own - partner| logit difference seen (0 = a tie)
    while (cur >= 0) {
        const unsigned rest = act & ~((2u << cur) - 1u);
        const int nxt = rest ? __ffs(rest) - 1 : -1;
        const Vec* src = hmb + ((size_t)b * P.K + pick4(nxt < 0 ? 0 : nxt, pj4.x, pj4.y, pj4.z, pj4.w)) * N4;
        asm volatile("" : "+l"(src));                 // keep the pointer in registers instead of re-deriving it per row
        const Vec* srcc = hmb + ((size_t)b * P.K + pick4(cur, pj4.x, pj4.y, pj4.z, pj4.w)) * N4;
        if (CQ) cp_async_wait_all();                  // this thread's slots hold partner `cur`
        f2 Sj2 = splat2(0.f), M2 = splat2(0.f);
#pragma unroll UNR
        for (int it = 0; it < NIT; ++it) {
            float4 q4;
            if (CQ) q4 = IO::from_slot(Qs + it * TPB + tid);
            else if (QRING) { cp_async_wait_but_one(); q4 = IO::from_slot(Qs + (it & 1) * TPB + tid); }   // one group per row: all but the newest have landed
            else q4 = IO::load_stream(srcc + it * TPB);                  // register-resident tile without slots: straight from L2
            const float4 o = own4(it);
            const f4 hv = as_f4(o), qv = as_f4(q4);
            float sk[4];
            if (CS) { const float4 s4 = Ss[it * TPB + tid]; sk[0] = s4.x; sk[1] = s4.y; sk[2] = s4.z; sk[3] = s4.w; }
            else {
                const float hh[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) sk[jj] = sigmoid_fast(hh[jj]);
            }
            const f2 u01 = mul2(qv.a, kNL2E), u23 = mul2(qv.b, kNL2E);
            const f2 g01 = add2(pack2(ex2(lo2(u01)), ex2(hi2(u01))), kOne), g23 = add2(pack2(ex2(lo2(u23)), ex2(hi2(u23))), kOne);
            float sq[4] = {rcp(lo2(g01)), rcp(hi2(g01)), rcp(lo2(g23)), rcp(hi2(g23))};
            // the slot has been consumed (its value went through the sigmoid): refill it with the next partner
            if (CQ && nxt >= 0) IO::copy_async(Qs + it * TPB + tid, src + it * TPB);
            if (QRING) {
                if (it + 2 < NIT) IO::copy_async(Qs + (it & 1) * TPB + tid, srcc + (it + 2) * TPB);
                else if (nxt >= 0) IO::copy_async(Qs + (it & 1) * TPB + tid, src + (it + 2 - NIT) * TPB);
                cp_async_commit();
            }
            // min(sigma(a), sigma(b)) = sigma(min(a, b)): decide on the logits; equal logits give equal sigmoids
            const f2 d01 = sub2(hv.a, qv.a), d23 = sub2(hv.b, qv.b);
            const float d[4] = {lo2(d01), hi2(d01), lo2(d23), hi2(d23)};
            unsigned tw = 0u;
            float sel[4];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const bool own_smaller = d[jj] < 0.f;
                sel[jj] = own_smaller ? sk[jj] : sq[jj];
                if (own_smaller) tw |= 4u << (8 * jj);
                mind = fminf(mind, fabsf(d[jj]));
            }
            Sj2 = add2(Sj2, add2(pack2(sq[0], sq[1]), pack2(sq[2], sq[3])));
            M2 = add2(M2, add2(pack2(sel[0], sel[1]), pack2(sel[2], sel[3])));
            if (ROLL) Ws[it * TPB + tid] |= tw << cur; else words[it] |= tw << cur;
        }
        const float Sj = hsum2(Sj2), M = hsum2(M2);
        if (CQ) cp_async_commit();
#pragma unroll
        for (int s = 0; s < 4; ++s) if (cur == s) { r16[6 + 2 * s] = Sj; r16[7 + 2 * s] = M; }
        cur = nxt;
    }
    const bool anyeq = mind == 0.f;

    // the normalisers have arrived by now
    const float D = (float)sum_w + kEps, D5 = (float)sum_p + kEps;
    const float gx = gtx * P.sx, gy = gty * P.sy;
    float lam[6];
#pragma unroll
    for (int q = 0; q < 6; ++q) lam[q] = backward_only ? lam_eff[q] : P.lam[q] * gscale;
    const float iD = rcp(D);
    const float ka = P.use_target_weight ? wa * iD : 1.f / (float)(P.B * P.K), kb = w * iD;
    // the variance-map gradient is uniform over the tile
    if (gv4) {
        const float g = lam[3] * kb * 2.f * (mV - P.sigma) * P.inv_n;
        const float4 g4 = make_float4(g, g, g, g);
#pragma unroll
        for (int it = 0; it < NIT; ++it) IO::store_stream(gv4 + it * TPB, g4);
    }
    if (grads && A.grad_var_mean && tid == 0) A.grad_var_mean[tile] = lam[3] * kb * 2.f * (mV - P.sigma);

    // ---- pass C: entropy sums and relu moments about (cx, cy) ----------------------------------------
    float dxj[4], dx2j[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { dxj[j] = (fx0 + (float)j) - cx; dx2j[j] = dxj[j] * dxj[j]; }
    const float dy0 = fty - cy;
    // the entropy derivative is parked negated: an = ln2 * lg2(p + eps) + p / (p + eps) = -a
    {
        const f2 kIZ = splat2(iZ), kEps2 = splat2(kEps), kLn2v = splat2(kLn2);
        f2 A1 = splat2(0.f), A2 = splat2(0.f), R01 = splat2(0.f), R23 = splat2(0.f), mse2 = splat2(0.f);
        float Ry = 0.f, Ry2 = 0.f;
#pragma unroll UNR
        for (int it = 0; it < NIT; ++it) {
            const float4 o = own4(it);
            const f4 hv = as_f4(o);
            f2 e01, e23;
            if (CE) { const f4 q = as_f4(Es[it * TPB + tid]); e01 = q.a; e23 = q.b; }
            else {
                const f2 t01 = fma2(hv.a, kL2E, kNML), t23 = fma2(hv.b, kL2E, kNML);
                e01 = pack2(ex2(lo2(t01)), ex2(hi2(t01))); e23 = pack2(ex2(lo2(t23)), ex2(hi2(t23)));
            }
            const f2 p01 = mul2(e01, kIZ), p23 = mul2(e23, kIZ);
            const f2 u01 = add2(p01, kEps2), u23 = add2(p23, kEps2);
            const f2 l01 = pack2(lg2(lo2(u01)), lg2(hi2(u01))), l23 = pack2(lg2(lo2(u23)), lg2(hi2(u23)));
            const f2 c01 = pack2(rcp(lo2(u01)), rcp(hi2(u01))), c23 = pack2(rcp(lo2(u23)), rcp(hi2(u23)));
            A1 = fma2(p01, l01, A1); A1 = fma2(p23, l23, A1);
            const f2 prc01 = mul2(p01, c01), prc23 = mul2(p23, c23);
            A2 = fma2(p01, prc01, A2); A2 = fma2(p23, prc23, A2);
            const f2 an01 = fma2(kLn2v, l01, prc01), an23 = fma2(kLn2v, l23, prc23);
            const f2 r01 = pack2(fmaxf(o.x, 0.f), fmaxf(o.y, 0.f)), r23 = pack2(fmaxf(o.z, 0.f), fmaxf(o.w, 0.f));
            R01 = add2(R01, r01); R23 = add2(R23, r23);
            if (MG) {
                const f4 tv = as_f4(target4(it));
                const f2 d01 = sub2(hv.a, tv.a), d23 = sub2(hv.b, tv.b);
                mse2 = fma2(d01, d01, mse2);
                mse2 = fma2(d23, d23, mse2);
            }
            if (CE) Es[it * TPB + tid] = as_float4(f4{p01, p23});
            if (CA) As[it * TPB + tid] = as_float4(f4{an01, an23});
            if (AQ) Qs[it * TPB + tid] = as_float4(f4{an01, an23});
            const float rs = hsum2(add2(r01, r23));
            const float dy = dy0 + (float)(it * ROWS);
            Ry = fmaf(dy, rs, Ry);
            Ry2 = fmaf(dy * dy, rs, Ry2);
        }
        float Rj[4];
        unpack2(R01, Rj[0], Rj[1]); unpack2(R23, Rj[2], Rj[3]);
        r16[0] = hsum2(A1); r16[1] = hsum2(A2);
        r16[2] = fmaf(dx2j[0], Rj[0], fmaf(dx2j[1], Rj[1], fmaf(dx2j[2], Rj[2], fmaf(dx2j[3], Rj[3], Ry2))));
        r16[3] = fmaf(dxj[0], Rj[0], fmaf(dxj[1], Rj[1], fmaf(dxj[2], Rj[2], dxj[3] * Rj[3])));
        r16[4] = Ry;
        r16[5] = (Rj[0] + Rj[1]) + (Rj[2] + Rj[3]);
        if (MG) r16[14] = hsum2(mse2);
    }

    // ---- reduction 3: entropy / variance sums and the partner sums in one go ------------------------------
    const float acc16 = block_sum1<16, NW>(r16, red0);
    const float mse_sum = MG ? lane_value<16>(acc16, 14) : lane_value<8>(acc8, 4);

    // ---- per-tile scalars, one warp per group; results meet in shared memory ---------------------------------
    // coef[0..3] = c1, c4, k4, c6   coef[4] = pa   coef[5..6] = dL/dc from warp RV   coef[7..8] = dL/dc from warp RO
    // coef[9] = 1 if any limb gradient is live;  tab[0..15] overlap coefficient per tie pattern, tab[16..19] partner scales
    float* coef = lutG;                   // 12 floats
    float* tab = lutG + 12;               // 20 floats
    float dcx = cx, dcy = cy, dtap = 0.f;
    if (warp == RD && decode) {
        // decode stage 2: window softmax, blend, start the bilinear offset read
        if (staged_decode) {
            const float vmax = warp_max(win);
            const float e = win_ok ? expf(win - vmax) : 0.f;
            const float se = warp_sum(e), sx = warp_sum(e * winx), sy = warp_sum(e * winy);
            const float a = sigmoid_acc(__ldg(A.alpha_param));
            dcx = a * cx + (1.f - a) * (sx / se);
            dcy = a * cy + (1.f - a) * (sy / se);
        } else if (A.dflags & GBCODEC_DECODE_REFINE) {
            refine_and_correct_cold<Elem>(hm_tile, nullptr, A.alpha_param, nullptr, H, W, A.radius, GBCODEC_DECODE_REFINE, &dcx, &dcy);
        }
        if (A.dflags & GBCODEC_DECODE_APPLY_OFFSET) {
            const Bilinear dbl = bilinear_setup(dcx, dcy, H, W);
            // lane t < 8 fetches tap (t & 3) of channel (t >> 2)
            const int tap = lane & 3;
            const int yy = (tap & 2) ? dbl.y1 : dbl.y0, xx = (tap & 1) ? dbl.x1 : dbl.x0;
            if (lane < 8) dtap = ld1(off_tile + (lane >> 2) * N + yy * W + xx);
        }
    }
    if (warp == RO) {
        const Taps tp = taps_setup(cx, cy, H, W);
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
        const float h2 = lam[1] * ka * 0.5f;
        if (lane == 0) {
            if (!backward_only) A.partial[(size_t)tile * 8 + 1] = wa * (0.5f * sl1);
            coef[7] = h2 * (sl1p[0] * (dsdx[0] + 1.f) + sl1p[1] * dsdx[1]);
            coef[8] = h2 * (sl1p[0] * dsdy[0] + sl1p[1] * (dsdy[1] + 1.f));
            if (grads) {
                // the (up to) four non-zero taps per channel of the offset gradient; the zero fill of these
                // addresses was issued before the first barrier, so it is ordered before these stores
                Elem* go = reinterpret_cast<Elem*>(A.grad_off) + (size_t)tile * 2 * N;
#pragma unroll
                for (int ch = 0; ch < 2; ++ch) {
                    Elem* o = go + ch * N;
                    const float gc = h2 * sl1p[ch];
                    IO::st1(o + tp.i00, gc * tp.w00);
                    if (tp.okx != 0.f) IO::st1(o + tp.i01, gc * tp.w01);
                    if (tp.oky != 0.f) IO::st1(o + tp.i10, gc * tp.w10);
                    if (tp.okx != 0.f && tp.oky != 0.f) IO::st1(o + tp.i11, gc * tp.w11);
                }
            }
        }
    }
    if (warp == RV) {
        const float iRp = rcp(lane_value<16>(acc16, 5) + kEps);
        const float peak_t = (cx - gx) * (cx - gx) + (cy - gy) * (cy - gy);
        const float v = lane_value<16>(acc16, 2) * iRp;
        const float sd = fsqrt_fast(v + kEps);
        const float var_t = (sd - P.sigma) * (sd - P.sigma) + (has_var ? (mV - P.sigma) * (mV - P.sigma) : 0.f);
        const float E = -kLn2 * lane_value<16>(acc16, 0);
        const float pa = E - lane_value<16>(acc16, 1);
        const float shape_t = (E - P.e_star) * (E - P.e_star);
        const float a4 = lam[3] * kb * (sd - P.sigma) * rcp(sd);
        const float c4 = a4 * iRp;
        const float dv_dcx = -2.f * lane_value<16>(acc16, 3) * iRp, dv_dcy = -2.f * lane_value<16>(acc16, 4) * iRp;
        if (lane == 0) {
            if (!backward_only) {
                float* p = A.partial + (size_t)tile * 8;
                p[0] = wa * (mse_sum * P.inv_n); p[2] = wa * peak_t; p[3] = w * var_t; p[5] = w * shape_t;
            }
            coef[0] = lam[0] * ka * 2.f * P.inv_n;
            coef[1] = c4;
            coef[2] = -c4 * v;
            coef[3] = lam[5] * kb * 2.f * (E - P.e_star);
            coef[4] = pa;
            coef[5] = lam[2] * ka * 2.f * (cx - gx) + a4 * dv_dcx;
            coef[6] = lam[2] * ka * 2.f * (cy - gy) + a4 * dv_dcy;
        }
    }
    if (warp == RP) {
        // overlap ratios -> loss numerator and the per-partner gradient scale
        float cj[4] = {0.f, 0.f, 0.f, 0.f};
        float cst = 0.f, pair_loss = 0.f;
        bool live = false;
        const float iD5 = rcp(D5);
#pragma unroll
        for (int pi = 0; pi < 4; ++pi) {
            if ((act >> pi) & 1u) {                     // warp-uniform
                const float Sj = lane_value<16>(acc16, 6 + 2 * pi), M = lane_value<16>(acc16, 7 + 2 * pi);
                const float imm = rcp(fminf(Ssum, Sj) + kEps);
                const float rho = M * imm;
                if ((P.owner[k] >> pi) & 1) pair_loss += w * wjv[pi] * fmaxf(rho - 0.5f, 0.f);
                if (grads && rho > 0.5f) {
                    cj[pi] = lam[4] * w * wjv[pi] * iD5 * imm;
                    cst += cj[pi] * rho * tie_rule(Ssum, Sj);
                    live = true;
                }
            }
        }
        if (lane == 0) {
            if (!backward_only) A.partial[(size_t)tile * 8 + 4] = pair_loss;
            coef[9] = live ? 1.f : 0.f;
        }
        // the overlap coefficient of a pixel as a function of its 4-bit tie pattern
        if (lane < 16) {
            float g = -cst;
#pragma unroll
            for (int pi = 0; pi < 4; ++pi) if ((lane >> pi) & 1) g += cj[pi];
            tab[lane] = g;
        } else if (lane < 20) {
            tab[lane] = pick4(lane - 16, cj[0], cj[1], cj[2], cj[3]);
        }
    }
    __syncthreads();
    const float4 cf0 = *reinterpret_cast<const float4*>(coef), cf1 = *reinterpret_cast<const float4*>(coef + 4);
    const float c1 = cf0.x, c4 = cf0.y, k4 = cf0.z, c6 = cf0.w, pa_ = cf1.x;
    const float fxx = cf1.y + coef[7], fyy = cf1.z + coef[8];
    const bool g_live = coef[9] != 0.f;
    const float* lutw = tab;
    if (!grads) {
        if (decode && warp == RD && (A.dflags & GBCODEC_DECODE_APPLY_OFFSET)) {
            float fw = __ldg(A.fusion_weight);
            if (A.dflags & GBCODEC_DECODE_FUSION_WEIGHT_RAW) fw = sigmoid_acc(fw);
            const Bilinear dbl = bilinear_setup(dcx, dcy, H, W);
            float t[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) t[q] = __shfl_sync(0xffffffffu, dtap, q);
            const float ox = dbl.w00 * t[0] + dbl.w01 * (t[1] * dbl.okx) + dbl.w10 * (t[2] * dbl.oky) + dbl.w11 * (t[3] * (dbl.okx * dbl.oky));
            const float oy = dbl.w00 * t[4] + dbl.w01 * (t[5] * dbl.okx) + dbl.w10 * (t[6] * dbl.oky) + dbl.w11 * (t[7] * (dbl.okx * dbl.oky));
            dcx += fw * ox;
            dcy += fw * oy;
        }
        if (decode && tid == RD * 32) { A.coords[2 * tile] = dcx; A.coords[2 * tile + 1] = dcy; A.scores[tile] = m; }
        return;
    }

    // ---- pass D: the heatmap gradient ----------------------------------------------------------------------
    // g = c1 (h - t) + p (c6 (a - pa) + (x - cx) Fx + (y - cy) Fy) + [h > 0] (c4 ((x-cx)^2 + (y-cy)^2) + k4) + overlap
    // with an = -a parked by pass C: c6 (a - pa) = -c6 (an + pa)
    {
        const f2 kIZ = splat2(iZ), kEps2 = splat2(kEps), kLn2v = splat2(kLn2);
        const f2 kC1 = splat2(c1), kNC6 = splat2(-c6), kPa = splat2(pa_), kC4 = splat2(c4), kK4 = splat2(k4);
        const f2 base01 = pack2(dxj[0] * fxx, dxj[1] * fxx), base23 = pack2(dxj[2] * fxx, dxj[3] * fxx);
        const f2 dx2_01 = pack2(dx2j[0], dx2j[1]), dx2_23 = pack2(dx2j[2], dx2j[3]);
#pragma unroll UNR
        for (int it = 0; it < NIT; ++it) {
            const float4 o = own4(it);
            const f4 hv = as_f4(o);
            f2 p01, p23, an01, an23;
            if (CE) { const f4 q = as_f4(Es[it * TPB + tid]); p01 = q.a; p23 = q.b; }
            else {
                const f2 t01 = fma2(hv.a, kL2E, kNML), t23 = fma2(hv.b, kL2E, kNML);
                p01 = mul2(pack2(ex2(lo2(t01)), ex2(hi2(t01))), kIZ); p23 = mul2(pack2(ex2(lo2(t23)), ex2(hi2(t23))), kIZ);
            }
            if (CA || AQ) { const f4 q = as_f4(CA ? As[it * TPB + tid] : Qs[it * TPB + tid]); an01 = q.a; an23 = q.b; }
            else {
                const f2 u01 = add2(p01, kEps2), u23 = add2(p23, kEps2);
                const f2 l01 = pack2(lg2(lo2(u01)), lg2(hi2(u01))), l23 = pack2(lg2(lo2(u23)), lg2(hi2(u23)));
                const f2 c01 = pack2(rcp(lo2(u01)), rcp(hi2(u01))), c23 = pack2(rcp(lo2(u23)), rcp(hi2(u23)));
                an01 = fma2(kLn2v, l01, mul2(p01, c01)); an23 = fma2(kLn2v, l23, mul2(p23, c23));
            }
            const f4 tv = as_f4(target4(it));
            const float dy = dy0 + (float)(it * ROWS);
            const f2 fyd = splat2(dy * fyy), dy2 = splat2(dy * dy);
            f2 g01 = mul2(kC1, sub2(hv.a, tv.a)), g23 = mul2(kC1, sub2(hv.b, tv.b));
            g01 = fma2(p01, fma2(kNC6, add2(an01, kPa), add2(base01, fyd)), g01);
            g23 = fma2(p23, fma2(kNC6, add2(an23, kPa), add2(base23, fyd)), g23);
            // relu branch of the variance term: rterm * [h > 0] (the mask is exactly 0 or 1)
            const f2 rt01 = fma2(kC4, add2(dx2_01, dy2), kK4), rt23 = fma2(kC4, add2(dx2_23, dy2), kK4);
            const f2 m01 = pack2(o.x > 0.f ? 1.f : 0.f, o.y > 0.f ? 1.f : 0.f), m23 = pack2(o.z > 0.f ? 1.f : 0.f, o.w > 0.f ? 1.f : 0.f);
            g01 = fma2(rt01, m01, g01); g23 = fma2(rt23, m23, g23);
            if (g_live) {
                f2 s01, s23;
                if (CS) { const f4 q = as_f4(Ss[it * TPB + tid]); s01 = q.a; s23 = q.b; }
                else { s01 = pack2(sigmoid_fast(o.x), sigmoid_fast(o.y)); s23 = pack2(sigmoid_fast(o.z), sigmoid_fast(o.w)); }
                const unsigned wd = ROLL ? Ws[it * TPB + tid] : words[it];
                float G[4];
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    G[j] = *reinterpret_cast<const float*>(reinterpret_cast<const char*>(lutw) + ((wd >> (8 * j)) & 0xFFu));
                const f2 kNeg = splat2(-1.f);
                g01 = fma2(pack2(G[0], G[1]), fma2(mul2(s01, kNeg), s01, s01), g01);
                g23 = fma2(pack2(G[2], G[3]), fma2(mul2(s23, kNeg), s23, s23), g23);
            }
            IO::store_stream(gh4 + it * TPB, as_float4(f4{g01, g23}));
        }
    }

    // rare: a logit of this thread equals its partner's — ATen's minimum splits that gradient evenly.
    // Patch the pixels this thread has just written (same thread, program order).
    if (g_live && anyeq) {
        for (int pi = 0; pi < 4; ++pi) {
            const float cjp = lutw[16 + pi];
            if (!((act >> pi) & 1u) || cjp == 0.f) continue;
            const Vec* src = hmb + ((size_t)b * P.K + pick4(pi, pj4.x, pj4.y, pj4.z, pj4.w)) * N4;
            for (int it = 0; it < NIT; ++it) {
                const float4 q = IO::load_keep(src + it * TPB), o = IO::load_keep(hm4 + it * TPB);
                if (q.x == o.x || q.y == o.y || q.z == o.z || q.w == o.w) {
                    float4 g = IO::load_plain(gh4 + it * TPB);
                    float s;
                    if (q.x == o.x) { s = sigmoid_fast(o.x); g.x = fmaf(0.5f * cjp * s, 1.f - s, g.x); }
                    if (q.y == o.y) { s = sigmoid_fast(o.y); g.y = fmaf(0.5f * cjp * s, 1.f - s, g.y); }
                    if (q.z == o.z) { s = sigmoid_fast(o.z); g.z = fmaf(0.5f * cjp * s, 1.f - s, g.z); }
                    if (q.w == o.w) { s = sigmoid_fast(o.w); g.w = fmaf(0.5f * cjp * s, 1.f - s, g.w); }
                    IO::store_plain(gh4 + it * TPB, g);
                }
            }
        }
    }

    // ---- decode stage 3 (warp RD): finish the bilinear read, publish -----------------------------------
    if (decode && warp == RD) {
        if (A.dflags & GBCODEC_DECODE_APPLY_OFFSET) {
            float fw = __ldg(A.fusion_weight);
            if (A.dflags & GBCODEC_DECODE_FUSION_WEIGHT_RAW) fw = sigmoid_acc(fw);
            const Bilinear dbl = bilinear_setup(dcx, dcy, H, W);
            float t[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) t[q] = __shfl_sync(0xffffffffu, dtap, q);
            const float ox = dbl.w00 * t[0] + dbl.w01 * (t[1] * dbl.okx) + dbl.w10 * (t[2] * dbl.oky) + dbl.w11 * (t[3] * (dbl.okx * dbl.oky));
            const float oy = dbl.w00 * t[4] + dbl.w01 * (t[5] * dbl.okx) + dbl.w10 * (t[6] * dbl.oky) + dbl.w11 * (t[7] * (dbl.okx * dbl.oky));
            dcx += fw * ox;
            dcy += fw * oy;
        }
        if (lane == 0) { A.coords[2 * tile] = dcx; A.coords[2 * tile + 1] = dcy; A.scores[tile] = m; }
    }
}

// forward (and fused step): one CTA per tile, grid = (K, B) — no division
template <int W4, int ROWS, int NIT, bool CE, bool CS, bool CA, bool CQ, bool ROLL, int TM, int MINB, bool HALF = false, bool MG = false, bool VSLOT = false>
__global__ void __launch_bounds__(W4* ROWS, MINB)
loss_tile_kernel(const __grid_constant__ LossParams P, const __grid_constant__ LossArgs A) {
    loss_tile_body<W4, ROWS, NIT, CE, CS, CA, CQ, ROLL, TM, HALF, MG, VSLOT>(P, A, blockIdx.x, blockIdx.y, gridDim.x);
}

// backward call: nothing to do if the stored gradients are already right (plan 0); float32 gradients that are off by one
// common factor are rescaled in place by rescale_kernel (plan 1), float16 ones are computed again (a stored half cannot be
// rescaled without a second rounding).  The grid is one wave of resident CTAs; each walks its share of the tiles.
template <int W4, int ROWS, int NIT, bool CE, bool CS, bool CA, bool CQ, bool ROLL, int TM, int MINB, bool HALF = false, bool MG = false, bool VSLOT = false>
__global__ void __launch_bounds__(W4* ROWS, MINB)
loss_tile_backward_kernel(const __grid_constant__ LossParams P, const __grid_constant__ LossArgs A, const int tiles_per_cta) {
    if (A.plan && (HALF ? *A.plan == 0 : *A.plan != 2)) return;
    const int tiles = P.B * P.K;
    const int t0 = blockIdx.x * tiles_per_cta, t1 = min(tiles, t0 + tiles_per_cta);
    for (int t = t0; t < t1; ++t) {
        const int b = t / P.K;
        loss_tile_body<W4, ROWS, NIT, CE, CS, CA, CQ, ROLL, TM, HALF, MG, VSLOT>(P, A, t - b * P.K, b, P.K);
        __syncthreads();                         // the next tile re-uses every shared-memory buffer
    }
}

// ---- launcher ----------------------------------------------------------------------------------------
template <int W4, int ROWS, int NIT, bool CE, bool CS, bool CA, bool CQ, bool ROLL, int TM, int MINB, bool HALF = false, bool MG = false, bool VSLOT = false>
static int launch_tile_t(const LossParams& P, const LossArgs& A, cudaStream_t s, cudaEvent_t e0, cudaEvent_t e1) {
    constexpr int TPB = W4 * ROWS, N4 = TPB * NIT, NW = TPB / 32;
    const size_t smem = (size_t)N4 * 16 * ((CQ ? 1 : 0) + (ROLL ? 1 : 0) + (CE ? 1 : 0) + (CS ? 1 : 0) + (CA ? 1 : 0))
                      + (size_t)((!CQ && ROLL) ? 2 * TPB * 16 : 0)
                      + (size_t)(2 * NW * 16 + 32 + 32 + (ROLL ? N4 : 0)) * 4 + (size_t)((P.ec.lut_size + 3) & ~3) * 4;
    if (smem > 227 * 1024 || P.B > 65535) return 1;
    if (A.lam_eff != nullptr) {                        // fusion_loss_backward: per-term upstream weights on the device
        auto bk = loss_tile_backward_kernel<W4, ROWS, NIT, CE, CS, CA, CQ, ROLL, TM, MINB, HALF, MG, VSLOT>;
        cudaError_t e = cudaFuncSetAttribute(bk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(bk, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return fail(GBCODEC_ERR_CUDA, "cudaFuncSetAttribute(loss_tile_backward_kernel): %s", cudaGetErrorString(e));
        int sms = 0, dev = 0;                           // per call: the process may drive devices of different sizes
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
        const int tiles = P.B * P.K, resident = sms * MINB;
        const int per = (tiles + resident - 1) / resident;            // one wave: balanced when there is work, ~sms*MINB CTAs to dismiss when there is none
        note_launch(), bk<<<(tiles + per - 1) / per, TPB, smem, s>>>(P, A, per);
        return check_launch("loss_tile_backward_kernel");
    }
    auto kern = loss_tile_kernel<W4, ROWS, NIT, CE, CS, CA, CQ, ROLL, TM, MINB, HALF, MG, VSLOT>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(GBCODEC_ERR_CUDA, "cudaFuncSetAttribute(loss_tile_kernel): %s", cudaGetErrorString(e));
    e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return fail(GBCODEC_ERR_CUDA, "cudaFuncSetAttribute(carveout): %s", cudaGetErrorString(e));
    if (e0) cudaEventRecord(e0, s);
    {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(P.K, P.B); cfg.blockDim = dim3(TPB); cfg.dynamicSmemBytes = smem; cfg.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        note_launch();
        e = cudaLaunchKernelEx(&cfg, kern, P, A);
        if (e != cudaSuccess) return fail(GBCODEC_ERR_CUDA, "cudaLaunchKernelEx(loss_tile_kernel): %s", cudaGetErrorString(e));
    }
    if (e1) cudaEventRecord(e1, s);
    return check_launch("loss_tile_kernel");
}

template <int W4, int ROWS, int NIT, bool CE, bool CS, bool CA, int MINB, bool CQ = true, bool ROLL = false, bool HALF = false, bool MG = false, bool VSLOT = false>
static int launch_tile_tm(const LossParams& P, const LossArgs& A, cudaStream_t s, cudaEvent_t e0, cudaEvent_t e1) {
    if (A.target) return launch_tile_t<W4, ROWS, NIT, CE, CS, CA, CQ, ROLL, kTargetGlobal, MINB, HALF, MG, VSLOT>(P, A, s, e0, e1);
    if (P.ec.ntap <= ROWS) return launch_tile_t<W4, ROWS, NIT, CE, CS, CA, CQ, ROLL, kTargetOneHit, MINB, HALF, MG, VSLOT>(P, A, s, e0, e1);
    return launch_tile_t<W4, ROWS, NIT, CE, CS, CA, CQ, ROLL, kTargetLut, MINB, HALF, MG, VSLOT>(P, A, s, e0, e1);
}

int launch_loss_tile(const LossParams& P, const LossArgs& A, cudaStream_t s, cudaEvent_t e0, cudaEvent_t e1) {
    // The instantiations with sigmoid slots (CS) take the variance tile through them as an asynchronous copy (VSLOT): no
    // registers held across the first wait for HBM (64x48 0.2755 -> 0.2670 ms, 96x72 0.659 -> 0.616, 64x64 0.360 -> 0.345).
    // Every default instantiation merges the maximum into the moment reduction (MG): 1-2 % faster on all four shapes than
    // a reduction of its own (64x48 0.2790 -> 0.2756 ms, 96x72 0.665 -> 0.659, 128x128 K=13 1.147 -> 1.121, 64x64 0.364 -> 0.360).
    if (A.half_io) {
        // float16 maps (autocast): the default instantiation of each shape; other shapes have no half path
        if (P.H == 64 && P.W == 48) return launch_tile_tm<12, 16, 4, false, true, false, 5, true, true, true, true, true>(P, A, s, e0, e1);
        if (P.H == 96 && P.W == 72) return launch_tile_tm<18, 16, 6, false, true, false, 3, false, true, true, true, true>(P, A, s, e0, e1);
        if (P.H == 128 && P.W == 128) return launch_tile_tm<32, 16, 8, false, false, false, 2, false, true, true, true>(P, A, s, e0, e1);
        if (P.H == 64 && P.W == 64) return launch_tile_tm<16, 16, 4, false, true, false, 4, true, true, true, true, true>(P, A, s, e0, e1);
        return 1;
    }
    // (DESIGN.md §4 lists the CTA shapes that were tried and measured slower — 96- and 384-thread CTAs, register-resident
    // unrolled loops, 4-8 CTAs per SM, a maximum reduction of its own, the variance tile through registers; their
    // instantiations are in the history of this file)
    if (P.H == 64 && P.W == 48) {
        return launch_tile_tm<12, 16, 4, false, true, false, 5, true, true, false, true, true>(P, A, s, e0, e1);   // 192 threads, 16 px each; H,S,Q in smem, 5 CTAs
    }
    if (P.H == 96 && P.W == 72) {
        return launch_tile_tm<18, 16, 6, false, true, false, 3, false, true, false, true, true>(P, A, s, e0, e1);  // rolled, 288 threads, 24 px each; H,S in smem, partners through a two-row ring: 3 CTAs (0.581 ms)
    }
    if (P.H == 64 && P.W == 64)          // the reference's other default map (data/pose_transforms.py:391): 256 threads, H,S,Q in smem, 4 CTAs
        return launch_tile_tm<16, 16, 4, false, true, false, 4, true, true, false, true, true>(P, A, s, e0, e1);
    if (P.H == 128 && P.W == 128) {
        return launch_tile_tm<32, 16, 8, false, false, false, 2, false, true, false, true>(P, A, s, e0, e1); // rolled, 512 threads, 32 px each; own tile + a two-row ring for the variance and partner rows in smem: 2 CTAs (1.067 ms)
    }
    return 1;
}

}  // namespace gbc
