// step_pipe.cu — the fused codec step as a software pipeline of specialised warps: target tiles generated on the fly +
// six-term fusion loss forward and backward + keypoint decode, one pass over every heatmap.
//
// Same arithmetic as step_tile.cu / loss_tile.cu (FusionPoseLoss.forward, models/fusion_head.py:745-806, terms :637-743
// and :405-559, the autograd backward of train.py:182 in closed form; decode: HeatmapRegressionHead.decode,
// fusion_head.py:309-365).  What changes is who does what, and when.  The profile of step_tile.cu (profiles/
// r02_step_tile_*): its pixel loops are 36 % of the instructions and a quarter of the warps' time; the rest is the serial
// per-tile chain every compute warp walks after the block reduction (cross-warp sums, ~230 dependent instructions of
// per-tile scalars, the offset taps), the barrier in front of it and the waits for data behind it.  Here:
//
//   * NW compute warps run ONLY pixel loops.  Per tile they do a FRONT half (maximum, softmax / relu moments, sigmoid,
//     limb-partner visits, variance sum -> per-warp partial sums published in shared memory) and a BACK half (the
//     gradient pass).  They execute front(i) and then back(i - 1): by the time a warp starts back(i - 1) the constants
//     it needs were finished long ago by
//   * the scalar warp, which owns the per-tile chain: it adds up the warps' partial sums of tile i while the compute
//     warps are in front(i + 1), derives every per-tile scalar, writes the ~36 constants of the gradient pass into shared
//     memory, then — off everybody's critical path — stores the loss numerators and the (up to) 8 non-zero offset
//     gradient taps and finishes the keypoint decode (local refinement + offset correction: three dependent L2 round
//     trips that no compute warp ever sees);
//   * the producer warp (one lane) draws tiles from the atomic counter and moves all bytes with bulk copies
//     (cp.async.bulk + mbarrier complete_tx; SASS UBLKCP / SYNCS): the tile + its 64-byte descriptor into a THREE-deep
//     ring (front(i), back(i - 1) and the prefetch of i + 1 are in flight together), limb partners and the variance tile
//     through a two-deep ring;
//   * no block barrier anywhere: every hand-over is an mbarrier (full / empty pairs), and each of them is normally
//     complete before it is waited for.
//
// Algorithmic HBM bytes per tile: read hm, var (8N); write d_hm, d_var, d_off (16N).
#include "loss_common.cuh"
#include "f32x2.cuh"
#include "decode_device.cuh"
#include <stdlib.h>
#include <type_traits>
#include <stdio.h>

namespace gbc {
namespace {

// ---- bulk copies and mbarriers (PTX) -----------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
// an mbarrier by its 32-bit shared-memory address: the generic-to-shared conversion happens once per kernel, not at every
// one of the ~40 arrive / wait sites (it was 135 instructions of the once-per-tile path)
struct Bar {
    unsigned a;
    __device__ __forceinline__ Bar operator+(unsigned k) const { return Bar{a + 8u * k}; }
    __device__ __forceinline__ Bar operator+(int k) const { return Bar{a + 8u * (unsigned)k}; }
};
__device__ __forceinline__ unsigned smem_u32(Bar b) { return b.a; }
__device__ __forceinline__ void mbar_init(Bar bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(Bar bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(Bar bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
// try_wait suspends the thread in hardware until the phase completes or a time limit passes: a wait costs a few issue
// slots however long it lasts
#define PIPE_INLINE __forceinline__
__device__ PIPE_INLINE void mbar_wait(Bar bar, unsigned parity) {
    asm volatile("{\n"
                 " .reg .pred p;\n"
                 "WAIT_%=:\n"
                 " mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 " @p bra DONE_%=;\n"
                 " bra WAIT_%=;\n"
                 "DONE_%=:\n"
                 "}" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
// waits that are long by design (the scalar warp between tiles, the producer for a free tile buffer): give the issue
// slots to the compute warps — polling took 15 % of the kernel's instructions
#ifndef PIPE_SLEEP
#define PIPE_SLEEP 1
#endif
#ifndef PIPE_SLEEP_NS
#define PIPE_SLEEP_NS 400
#endif
__device__ PIPE_INLINE void mbar_wait_idle(Bar bar, unsigned parity) {
#if PIPE_SLEEP == 1
    for (;;) {
        unsigned ok;
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (ok) return;
        __nanosleep(PIPE_SLEEP_NS);
    }
#elif PIPE_SLEEP == 2
    asm volatile("{\n"
                 " .reg .pred p;\n"
                 "WAITI_%=:\n"
                 " mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
                 " @p bra DONEI_%=;\n"
                 " bra WAITI_%=;\n"
                 "DONEI_%=:\n"
                 "}" :: "r"(smem_u32(bar)), "r"(parity), "r"(20000u) : "memory");
#else
    mbar_wait(bar, parity);
#endif
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, Bar bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ float max3f(float a, float b, float c) { float d; asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float min3f(float a, float b, float c) { float d; asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float sqrt_approx(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float4 lds4(const void* p) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_u32(p)));
    return v;
}

// Unroll factor of the hot pixel loops.  The L0 instruction cache holds ~6 KB, L1.5 32 KB: fully unrolled (x4) the three
// roles' hot paths are ~40 KB of SASS and a quarter of all stall samples are "no instruction"; rolled, an iteration's
// body runs from L0.  The tie words are rotated through four registers so that a rolled loop needs no indexed array.
#ifndef PIPE_UNROLL
#define PIPE_UNROLL 1
#endif
// per-loop overrides (A/B measurements): the small loops (zero fill, maximum, variance sum) cost a few instructions per
// iteration and can be unrolled without growing the hot path much; the three big bodies are what the caches feel
#ifndef PIPE_UNROLL_SMALL
#define PIPE_UNROLL_SMALL 4
#endif
#ifndef PIPE_UNROLL_B
#define PIPE_UNROLL_B PIPE_UNROLL
#endif
#ifndef PIPE_UNROLL_P
#define PIPE_UNROLL_P 4
#endif
#ifndef PIPE_UNROLL_D
#define PIPE_UNROLL_D PIPE_UNROLL
#endif
#ifndef PIPE_SRED_UNROLL
#define PIPE_SRED_UNROLL 1
#endif
constexpr int kSredU = PIPE_SRED_UNROLL;
// the producer's loop over the ring items of a tile, rolled: the compiler's unrolled version was 175 instructions longer,
// all of them on the once-per-tile path
#ifndef PIPE_PROD_UNROLL
#define PIPE_PROD_UNROLL 1
#endif
constexpr int kProdU = PIPE_PROD_UNROLL;
#define PIPE_TROW_INLINE PIPE_INLINE
// PIPE_TAIL_OUTSIDE: when the call decodes, the scalar warp leaves the soft-argmax in d_coords and the two offset-gradient
// factors in the spare words of the tile's numerator row; finalize_kernel's tail CTAs (loss.cu: one warp per tile, same
// operations in the same order) place the offset-gradient taps and finish the decode.  ~270 instructions less on the
// once-per-tile path of this kernel, whose time follows that path's footprint.
#ifndef PIPE_TAIL_OUTSIDE
#define PIPE_TAIL_OUTSIDE 1
#endif
// Ring order: the variance tile is the FIRST ring item of a tile, the limb partners follow.  It is the one ring item that
// always comes from HBM (the partners are other CTAs' tiles of the same image: L2 hits); as the last item of the two-deep
// ring it could only be requested once the first partner had been consumed, and the compute warps sat out most of a DRAM
// round trip per tile in front of it (10 % of the kernel's stall samples).  As the first item it is requested while the
// previous tile's partners are still being consumed and has landed before pass B is over: 0.2300 -> 0.2220 ms.
// Measured and NOT kept (profiles/r02_step_variants_s4.jsonl; the switches are in the history of this
// file): the variance sum at the very start of the front half or right in front of pass B (+2 %, +1.5 %: the copy has not
// always landed), the scalar warp reading the variance tile itself with plain loads so that the ring carries partners only
// (+35 %: three DRAM round trips on its chain make it the pipeline's period), the next tile requested before the ring
// items (+4 %), the target row / squared-error correction / zero fill between the variance sum and the partner visits (+9 %).
//
// PIPE_MERGE_B1: heavy tiles without an active limb partner (6-8 % of the tiles) run pass B with the sigmoid too (its
// result unused) instead of a pass-B instantiation of their own: ~1 KB less of warm code (0.2140 -> 0.2118 ms).
#ifndef PIPE_MERGE_B1
#define PIPE_MERGE_B1 1
#endif
// ring depth of the float16 instantiation (its tiles are 6 KB: four ring buffers and the rest are 55 KB per CTA; a
// two-deep ring measures the same)
#ifndef PIPE_HALF_RING
#define PIPE_HALF_RING 4
#endif
// PIPE_CARRY_TARGET: the row in which a thread meets the target patch (target_row: ~100 instructions) is worked out in the
// front half and carried in five registers to the tile's back half one iteration later, instead of being worked out
// again there: one inlined copy of target_row less on the once-per-tile path (0.2220 -> 0.2140 ms).
#ifndef PIPE_CARRY_TARGET
#define PIPE_CARRY_TARGET 1
#endif
constexpr int kPUs = PIPE_UNROLL_SMALL, kPUb = PIPE_UNROLL_B, kPUp = PIPE_UNROLL_P, kPUd = PIPE_UNROLL_D;
// Lane sums: the compute warps do not reduce their 16 running sums across lanes (a 31-shuffle butterfly, ~125
// instructions per warp and tile); every lane stores its sums into its own four slots of the sigmoid tile — free once the
// partner visits are over — and the scalar warp adds up the 6 x 32 lanes (off the critical path, a rolled loop).
// Also measured and not kept in earlier sessions (profiles/r02_step_variants.jsonl; switches in this file's history):
// helpers out of line, decode loads requested early in the scalar warp (+3 %: ~100 once-per-tile instructions more),
// roles laid out per CTA rank for balanced schedulers, a code-size probe (4 KB of hot code more = +5 %).

constexpr float kFlatRmax = 1e-3f;              // largest eps / p for which the entropy shortcut holds to 1e-7
constexpr float kShiftCond = 4.f;               // largest sum |terms| / |result| accepted for the shifted relu moments

// flags in the constants block
constexpr unsigned kFHeavy = 1u, kFFlat = 2u, kFLive = 4u, kFFastSig = 8u;

// constants of the gradient pass, per tile (floats): written by the scalar warp, read by the compute warps
//  0 iZ   1 ml   2 cx   3 cy | 4 c1   5 c4   6 k4   7 fxx | 8 fyy   9 ke (flat: c6 ln2, else c6)  10 addc  11 gvar |
// 12 flags  13 eM (= exp(m), for the sigmoid derivative)  14,15 - | 16..31 overlap coefficient per tie pattern | 32..35 cj
constexpr int kConsFloats = 48;

// ---- element type of the maps: float32, or float16 under autocast (train.py:171) ---------------------------------------
// A "vector" is four pixels: 16 bytes of float, 8 bytes of half.  Half maps travel as they are (bulk copies of raw halves:
// a 64x48 tile is 6 KB) and are up-cast value by value where a warp reads them from shared memory; gradients are rounded
// once where they leave.  Everything in between is the same code, so losses and decode equal the float32 kernel's on the
// up-cast maps bit for bit.
template <bool HALF> struct PX;
template <> struct PX<false> {
    using Vec = float4;
    using Elem = float;
    static __device__ __forceinline__ float4 lds(const Vec* p) { return *p; }
    static __device__ __forceinline__ float4 ldg(const Vec* p) { return ldg_keep(p); }
    static __device__ __forceinline__ float4 ld_plain(const Vec* p) { return *p; }
    static __device__ __forceinline__ void stg(Vec* p, const float4& v) { stg_stream(p, v); }
    static __device__ __forceinline__ void st_plain(Vec* p, const float4& v) { *p = v; }
    static __device__ __forceinline__ void st1(Elem* p, float v) { *p = v; }
};
template <> struct PX<true> {
    using Vec = uint2;
    using Elem = __half;
    static __device__ __forceinline__ float4 up(const uint2& r) {
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&r.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
        return make_float4(a.x, a.y, b.x, b.y);
    }
    static __device__ __forceinline__ uint2 down(const float4& v) {
        const __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
        uint2 r;
        r.x = *reinterpret_cast<const unsigned*>(&a); r.y = *reinterpret_cast<const unsigned*>(&b);
        return r;
    }
    static __device__ __forceinline__ float4 lds(const Vec* p) { return up(*p); }
    static __device__ __forceinline__ float4 ldg(const Vec* p) { return up(__ldg(p)); }
    static __device__ __forceinline__ float4 ld_plain(const Vec* p) { return up(*p); }
    static __device__ __forceinline__ void stg(Vec* p, const float4& v) {
        const uint2 r = down(v);
        asm volatile("st.global.cs.v2.u32 [%0], {%1,%2};" :: "l"(p), "r"(r.x), "r"(r.y) : "memory");
    }
    static __device__ __forceinline__ void st_plain(Vec* p, const float4& v) { *p = down(v); }
    static __device__ __forceinline__ void st1(Elem* p, float v) { *p = __float2half_rn(v); }
};

// RD: depth of the ring that carries the variance tile and the limb partners (a power of two).  Two for float32 tiles
// (six 12 KB buffers are what fits an SM three times); four for float16 tiles, whose buffers are half the size.
template <int W4, int ROWS, int NIT, bool HALF = false, int RD = 2>
struct PipePlan {
    static constexpr int TPB = W4 * ROWS, NW = TPB / 32, N4 = TPB * NIT, N = 4 * N4, W = 4 * W4, H = ROWS * NIT;
    static constexpr int kTile = N4 * (HALF ? 8 : 16);           // bytes of one tile as it travels
    static constexpr int oH = 0;                                 // three tile buffers
    static constexpr int oS = 3 * kTile;                         // sigmoid of the tile in its front half (float32 either way)
    static constexpr int oR = oS + N4 * 16;                      // RD ring buffers: variance tile, limb partners
    static constexpr int oDesc = oR + RD * kTile;                // three descriptors
    static constexpr int oRed = oDesc + 3 * 64;                  // 2 x NW x 16 floats
    static constexpr int oRed2 = oRed + 2 * NW * 16 * 4;         // 2 x NW x 4 floats (partners 3 and 4)
    static constexpr int oRedM = oRed2 + 2 * NW * 4 * 4;         // 2 x NW x 2 floats (per-warp max, min)
    static constexpr int oRedC = oRedM + 2 * NW * 2 * 4;         // NW x 8 floats (general second pass)
    static constexpr int oCons = oRedC + NW * 8 * 4;             // 2 x kConsFloats
    static constexpr int oBar = oCons + 2 * kConsFloats * 4;     // 20 + 2 RD mbarriers
    static constexpr int oTid = oBar + (20 + 2 * RD) * 8;        // 3 tile indices (+ pad)
    static constexpr int oCta = oTid + 16;                       // float4: 1/(sum w + eps), 1/(sum w_i w_j + eps), gradient scale
    static constexpr int oLut = oCta + 16;                       // exp table of the target patch
    static_assert(TPB % 32 == 0 && NW >= 2 && NW <= 16, "whole warps");
    static_assert(RD >= 2 && (RD & (RD - 1)) == 0, "ring depth: a power of two");
    static_assert((oS % 16) == 0 && (oR % 16) == 0 && (kTile % 16) == 0, "bulk copies land on 16-byte boundaries");
    static_assert((oBar % 8) == 0 && (oDesc % 16) == 0 && (oLut % 16) == 0 && (oCons % 16) == 0 && (oRed % 16) == 0, "alignment");
};

// the one row (if any) in which this thread meets the on-the-fly target patch
template <int ROWS, int NIT>
__device__ PIPE_TROW_INLINE int target_row(const int4& gq, float w, int x0, int ty, const EncodeConst& ec, const float* lut, float4& thit) {
    thit = make_float4(0.f, 0.f, 0.f, 0.f);
    const PatchGeom geom = unpack_geom(gq, w);
    if (!(geom.active && x0 + 3 >= geom.x_from && x0 < geom.x_to)) return -1;
    const int it0 = max(0, (geom.y_from - ty + ROWS - 1) / ROWS);
    const int y = it0 * ROWS + ty;
    if (!(it0 < NIT && y < geom.y_to)) return -1;
    const int pcx = geom.ulx + (int)ec.centre, pcy = geom.uly + (int)ec.centre;
    const int dy2 = (y - pcy) * (y - pcy);
    float e[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int xx = x0 + j, dx = xx - pcx;
        e[j] = (xx >= geom.x_from && xx < geom.x_to) ? lut[dx * dx + dy2] : 0.f;
    }
    thit = make_float4(e[0], e[1], e[2], e[3]);
    return it0;
}

// ---- the kernel ------------------------------------------------------------------------------------
// GRADS = false: forward only (validate.py:90 evaluates the loss under no_grad every batch): no tie words, no gradient
// pass, no stores besides the per-tile loss numerators and the decode.
template <int W4, int ROWS, int NIT, int MINB, bool GRADS, bool HALF = false, int RD = 2>
__global__ void __launch_bounds__(W4* ROWS + 64, MINB)
step_pipe_kernel(const __grid_constant__ LossParams P, const __grid_constant__ LossArgs A) {
    using L = PipePlan<W4, ROWS, NIT, HALF, RD>;
    using IO = PX<HALF>;
    using Vec = typename IO::Vec;
    using Elem = typename IO::Elem;
    static_assert(!HALF || PIPE_TAIL_OUTSIDE, "the float16 instantiation leaves the decode tail to finalize_kernel");
    constexpr int TPB = L::TPB, NW = L::NW, N4 = L::N4, N = L::N, W = L::W, H = L::H;
    constexpr unsigned kTile = L::kTile;
    extern __shared__ __align__(128) unsigned char smraw[];
    Vec* const Hb = reinterpret_cast<Vec*>(smraw + L::oH);
    float4* const Sb = reinterpret_cast<float4*>(smraw + L::oS);
    Vec* const Rb = reinterpret_cast<Vec*>(smraw + L::oR);
    TileDesc* const Db = reinterpret_cast<TileDesc*>(smraw + L::oDesc);
    float* const red = reinterpret_cast<float*>(smraw + L::oRed);
    float* const red2 = reinterpret_cast<float*>(smraw + L::oRed2);
    float* const redM = reinterpret_cast<float*>(smraw + L::oRedM);
    float* const redC = reinterpret_cast<float*>(smraw + L::oRedC);
    float* const cons = reinterpret_cast<float*>(smraw + L::oCons);
    uint64_t* const bars = reinterpret_cast<uint64_t*>(smraw + L::oBar);
    int* const tids = reinterpret_cast<int*>(smraw + L::oTid);
    float4* const ctas = reinterpret_cast<float4*>(smraw + L::oCta);
    float* const lut = reinterpret_cast<float*>(smraw + L::oLut);
    const Bar bar0{smem_u32(static_cast<const void*>(bars))};
    const Bar hfull = bar0;                  // [3] tile + descriptor landed                       (producer -> all)
    const Bar hempty = bar0 + 3;             // [3] every compute warp is done with the tile       (compute -> producer)
    const Bar rfull = bar0 + 20;             // [RD] ring item landed                              (producer -> compute)
    const Bar rempty = bar0 + (20 + RD);     // [RD] every compute warp has consumed it            (compute -> producer)
    const Bar sfull = bar0 + 10;             // [2] every compute warp has published its sums      (compute -> scalar)
    const Bar sempty = bar0 + 12;            // [2] the scalar warp has read them                  (scalar -> compute)
    const Bar cfull = bar0 + 14;             // [2] constants of the gradient pass written         (scalar -> compute)
    const Bar cempty = bar0 + 16;            // [2] every compute warp is done with them           (compute -> scalar)
    const Bar p2full = bar0 + 18;            // general second pass: sums published                (compute -> scalar)
    const Bar p2done = bar0 + 19;            // ... final constants written                        (scalar -> compute)

    const int lane = threadIdx.x & 31;
    const int tiles = P.B * P.K;
    const bool has_var = A.var != nullptr;
    const bool has_vmean = A.var_mean != nullptr;            // the variance branch reduced to mean_N(V) by the head: no variance tile
    const bool decode = A.coords != nullptr;
    const Elem* const hm = reinterpret_cast<const Elem*>(A.hm);
    const Elem* const var_maps = reinterpret_cast<const Elem*>(A.var);

    // ---- once per CTA -------------------------------------------------------------------------------
    if (threadIdx.x == 0) {
        tids[3] = 0;
#pragma unroll
        for (int q = 0; q < 3; ++q) { mbar_init(hfull + q, 1); mbar_init(hempty + q, NW); }
#pragma unroll
        for (int q = 0; q < RD; ++q) { mbar_init(rfull + q, 1); mbar_init(rempty + q, NW); }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            mbar_init(sfull + q, NW); mbar_init(sempty + q, 1);
            mbar_init(cfull + q, 1); mbar_init(cempty + q, NW);
        }
        mbar_init(p2full, NW); mbar_init(p2done, 1);
        fence_mbar_init();
    }
    for (int q = threadIdx.x; q < P.ec.lut_size; q += TPB + 64) lut[q] = expf(-(float)q / P.ec.two_sigma_sq);
    __syncthreads();

    // ---- roles --------------------------------------------------------------------------------------
    // A warp's scheduler (SM sub-partition) is its slot in the SM modulo 4, and the CTA's eight warps take consecutive
    // slots.  With the producer and the scalar warp as warps 6 and 7 of every CTA, the three CTAs of an SM put 6 compute
    // warps on each of the sub-partitions 0 and 1 and 3 on each of 2 and 3 — and the compute warps of a tile move in lock
    // step (ring buffers and sums are handed over when ALL of them are done), so the kernel ran at the pace of the two
    // crowded schedulers.  Laid out per CTA rank, the six light warps of an SM sit 2 / 2 / 1 / 1 on the sub-partitions
    // and the compute warps 4 / 4 / 5 / 5 (scalar warps, the heavier of the light ones, where only 4 compute warps are):
    //   rank 0: scalar = warp 4 (sub-partition 0), producer = warp 5 (1)
    //   rank 1: scalar = warp 4 (0),               producer = warp 6 (2)
    //   rank 2: scalar = warp 5 (1),               producer = warp 7 (3)
    // The logical warp index used below: compute warps 0 .. NW-1 in slot order, producer = NW, scalar = NW + 1.
    const int warp = threadIdx.x >> 5;
    const int tid = warp * 32 + lane;

    // ======================================================================================================
    // producer warp: one lane
    // ======================================================================================================
    if (warp == NW) {
        if (lane != 0) return;
        unsigned cur = blockIdx.x;
        tids[0] = (int)cur;                                  // published by the arrive below (release)
        mbar_arrive_expect_tx(hfull, kTile + 64);
        bulk_g2s(Hb, hm + (size_t)cur * N, kTile, hfull);    // the maps do not depend on the predecessor kernel
        pdl_wait();                                          // the descriptors do
        pdl_launch_dependents();
        bulk_g2s(Db, A.desc + cur, 64, hfull);
        unsigned nxt = atomicAdd(A.tile_counter, 1u) + gridDim.x;
        unsigned nxt2 = atomicAdd(A.tile_counter, 1u) + gridDim.x;      // two ahead: the atomic's latency is off the path
        // What the producer needs of a descriptor (weight, active partners) it reads itself, one tile ahead, with plain
        // loads: the ring items of a tile can then be requested before the tile itself has landed.
        auto desc_w = [&](unsigned t) { return t < (unsigned)tiles ? __ldg(&A.desc[t].w) : 0.f; };
        auto desc_pk = [&](unsigned t) { return t < (unsigned)tiles ? __ldg(&A.desc[t].pk) : 0u; };
        auto desc_pj = [&](unsigned t) { return t < (unsigned)tiles ? __ldg(&A.desc[t].pj) : 0u; };
        float w_c = desc_w(cur), w_n = desc_w(nxt);
        unsigned pk_c = desc_pk(cur), pk_n = desc_pk(nxt), pj_c = desc_pj(cur), pj_n = desc_pj(nxt);
        unsigned rq = 0;                                     // ring items started so far
        unsigned ph_empty = 0;                               // phase bits per tile buffer
        for (unsigned j = 0;; ++j) {
            const unsigned s1 = (j + 1) % 3u;
            // ring items of tile j: its variance tile, then its limb partners
            const bool heavy = (w_c != 0.f) || !P.use_target_weight;
            if (heavy) {
                const int nn = (int)(pk_c & 7u);
                const size_t b = cur / (unsigned)P.K;
#pragma unroll kProdU
                for (int n = 0; n < nn + (has_var ? 1 : 0); ++n) {
                    const unsigned q = rq & (unsigned)(RD - 1);
                    if (rq >= (unsigned)RD) mbar_wait(rempty + q, ((rq - RD) / RD) & 1u);
                    const int pn = n - (has_var ? 1 : 0);       // -1: the variance tile
                    const Elem* src = pn >= 0 ? hm + (b * P.K + ((pj_c >> (8 * pn)) & 0xFFu)) * N : var_maps + (size_t)cur * N;
                    mbar_arrive_expect_tx(rfull + q, kTile);
                    bulk_g2s(Rb + q * N4, src, kTile, rfull + q);
                    ++rq;
                }
            }
            // tile j + 1 into the buffer tile j - 2 has left
            if (j >= 2) { mbar_wait_idle(hempty + s1, (ph_empty >> s1) & 1u); ph_empty ^= 1u << s1; }
            cur = nxt; nxt = nxt2;
            w_c = w_n; pk_c = pk_n; pj_c = pj_n;
            if (cur >= (unsigned)tiles) {
                tids[s1] = -1;
                mbar_arrive(hfull + s1);                     // the sentinel: everybody leaves at tile j + 1
                break;
            }
            w_n = desc_w(nxt); pk_n = desc_pk(nxt); pj_n = desc_pj(nxt);
            if (nxt2 < (unsigned)tiles) nxt2 = atomicAdd(A.tile_counter, 1u) + gridDim.x;
            tids[s1] = (int)cur;
            mbar_arrive_expect_tx(hfull + s1, kTile + 64);
            bulk_g2s(Hb + s1 * N4, hm + (size_t)cur * N, kTile, hfull + s1);
            bulk_g2s(Db + s1, A.desc + cur, 64, hfull + s1);
        }
        return;
    }

    // everything below depends on the kernel that prepared the weights (programmatic dependent launch)
    pdl_wait();
    pdl_launch_dependents();

    // ======================================================================================================
    // scalar warp: cross-warp sums -> per-tile scalars -> constants of the gradient pass; loss numerators, offset
    // gradient taps, keypoint decode
    // ======================================================================================================
    if (warp == NW + 1) {
        const float iD = rcp((float)__ldg(A.sums) + kEps), iD5 = rcp((float)__ldg(A.sums + 1) + kEps);
        const float gscale = A.grad_scale ? __ldg(A.grad_scale) : 1.f;
        constexpr float ax = 0.5f * (float)(W - 1), ay = 0.5f * (float)(H - 1);     // anchor of the relu moments
        constexpr int wx0 = (W - 1) / 2 - 1, wy0 = (H - 1) / 2 - 1;                 // origin of the 4x4 tap window
        // decode constants of the call
        const int wside = 2 * A.radius + 1;
        const bool win_small = decode && (A.dflags & GBCODEC_DECODE_REFINE) && wside * wside <= 32;
        const bool in_win = lane < wside * wside;
        const int wdx = lane % wside - A.radius, wdy = lane / wside - A.radius;
        (void)in_win; (void)wdx; (void)wdy;             // used by the in-kernel decode only (PIPE_TAIL_OUTSIDE = 0)
        float a_blend = 1.f, fw_dec = 0.f;
        if (win_small) {
            a_blend = sigmoid_acc(__ldg(A.alpha_param));
            if (A.dflags & GBCODEC_DECODE_APPLY_OFFSET) {
                fw_dec = __ldg(A.fusion_weight);
                if (A.dflags & GBCODEC_DECODE_FUSION_WEIGHT_RAW) fw_dec = sigmoid_acc(fw_dec);
            }
        }
        unsigned ph_full = 0, np2 = 0;
        for (unsigned i = 0;; ++i) {
            const unsigned s = i % 3u, b = i & 1u;
            mbar_wait_idle(hfull + s, (ph_full >> s) & 1u); ph_full ^= 1u << s;
            const int tile = tids[s];
            if (tile < 0) break;
            const TileDesc* dsc = Db + s;
            const float4 d1 = *reinterpret_cast<const float4*>(&dsc->w);          // w, gx, gy, pk
            const float w = d1.x, gx = d1.y, gy = d1.z;
            const unsigned pk = __float_as_uint(d1.w);
            const float4 wj4 = *reinterpret_cast<const float4*>(dsc->wj);
            const int nact = (int)(pk & 7u);
            const float wa = P.use_target_weight ? w : 1.f;
            const bool heavy = (w != 0.f) || !P.use_target_weight;
            const Elem* off_tile = reinterpret_cast<const Elem*>(A.off) + (size_t)tile * 2 * N;
            // the 4x4x2 window of offset taps around the tile centre (consumed after the sums)
            float tapv = 0.f;
            if (heavy) tapv = ld1(off_tile + (lane >> 4) * N + (wy0 + ((lane >> 2) & 3)) * W + wx0 + (lane & 3));
            float vmean_in = 0.f;
            if (has_vmean && heavy) vmean_in = __ldg(A.var_mean + tile);

            // ---- the warps' partial sums of tile i --------------------------------------------------------------
            mbar_wait_idle(sfull + b, (i >> 1) & 1u);
            const float* red2b = red2 + b * (NW * 4);
            const float* redMb = redM + b * (NW * 2);
            float m = redMb[0], hmin = redMb[1];
#pragma unroll
            for (int ww = 1; ww < NW; ++ww) { m = fmaxf(m, redMb[2 * ww]); hmin = fminf(hmin, redMb[2 * ww + 1]); }
            // value k = 4 * row + c of warp ww, lane l sits in component c of Sb[row * TPB + ww * 32 + l].  This lane takes
            // row = lane >> 3 and the source lanes g, g + 8, g + 16, g + 24 (g = lane & 7: conflict-free 128-bit reads),
            // then the eight lanes of a row add up.  Row 0 carries the softmax sums, relative to each warp's own maximum.
            float acc4[4] = {0.f, 0.f, 0.f, 0.f};
            {
                const int row = lane >> 3, g = lane & 7;
#pragma unroll kSredU
                for (int ww = 0; ww < NW; ++ww) {
                    const float4* src = Sb + row * TPB + ww * 32 + g;
                    const float4 a = src[0], b4 = src[8], c = src[16], d = src[24];
                    float x0 = (a.x + b4.x) + (c.x + d.x), x1 = (a.y + b4.y) + (c.y + d.y);
                    float x2 = (a.z + b4.z) + (c.z + d.z), x3 = (a.w + b4.w) + (c.w + d.w);
                    if (row == 0) {
                        const float dl = (redMb[2 * ww] - m) * kLog2e;            // (m_w - m) log2 e <= 0
                        const float sc = ex2(dl);
                        x3 = fmaf(dl, x0, x3);                                    // sum e t: t is relative to the warp's maximum too
                        x0 *= sc; x1 *= sc; x2 *= sc; x3 *= sc;
                    }
                    acc4[0] += x0; acc4[1] += x1; acc4[2] += x2; acc4[3] += x3;
                }
#pragma unroll
                for (int o = 1; o < 8; o <<= 1) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) acc4[c] += __shfl_xor_sync(0xffffffffu, acc4[c], o);
                }
            }
            float a4s = 0.f;
            if (nact > 2) {
                const int idx = lane >> 3, q = lane & 7;
#pragma unroll
                for (int t = 0; t < (NW + 7) / 8; ++t) {
                    const int ww = q + 8 * t;
                    if (ww < NW) a4s += red2b[ww * 4 + idx];
                }
                a4s += __shfl_xor_sync(0xffffffffu, a4s, 1);
                a4s += __shfl_xor_sync(0xffffffffu, a4s, 2);
                a4s += __shfl_xor_sync(0xffffffffu, a4s, 4);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(sempty + b);          // the sums are in registers
            auto val = [&](int k) -> float { return __shfl_sync(0xffffffffu, acc4[k & 3], (k >> 2) << 3); };
            const float Zs = val(0);
            const float iZ = rcp(Zs);
            const float cx = val(1) * iZ, cy = val(2) * iZ;
            const float ml = m * kLog2e;
            float* const cb = cons + b * kConsFloats;

            if (!heavy) {
                // weight 0: every term carries a factor w -> zero loss and gradient; decode only
                if (i >= 2) mbar_wait(cempty + b, ((i - 2) >> 1) & 1u);
                if (lane == 0) {
                    *reinterpret_cast<float4*>(cb) = make_float4(iZ, ml, cx, cy);
                    *reinterpret_cast<float4*>(cb + 12) = make_float4(__uint_as_float(0u), 0.f, 0.f, 0.f);
                    float4* p = reinterpret_cast<float4*>(A.partial + (size_t)tile * 8);
                    p[0] = make_float4(0.f, 0.f, 0.f, 0.f); p[1] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (GRADS && has_vmean) A.grad_var_mean[tile] = 0.f;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(cfull + b);
            } else {
                // ---- per-tile scalars ------------------------------------------------------------------------------
                const float Vsum = val(5);
                const float ET = val(3), Ssum = val(4), Rm = val(6), Rxa = val(7), Rya = val(8), M2a = val(9), mse_sum = val(10);
                const float mV = has_var ? Vsum * P.inv_n : (has_vmean ? vmean_in : P.sigma);
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
                const bool fastsig = m <= 30.f && m >= -80.f;
                if (i >= 2) mbar_wait(cempty + b, ((i - 2) >> 1) & 1u);
                if (!flat) {
                    // ---- general second pass: the compute warps run it in their back half with these ------------------
                    if (lane == 0) {
                        *reinterpret_cast<float4*>(cb) = make_float4(iZ, ml, cx, cy);
                        *reinterpret_cast<float4*>(cb + 12) = make_float4(__uint_as_float(kFHeavy), 0.f, 0.f, 0.f);
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(cfull + b);
                    mbar_wait(p2full, np2 & 1u);
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
                        const int l0 = by * 4 + bx;
#pragma unroll
                        for (int c = 0; c < 2; ++c) {
                            ov[c][0] = __shfl_sync(0xffffffffu, tapv, c * 16 + l0); ov[c][1] = __shfl_sync(0xffffffffu, tapv, c * 16 + l0 + 1);
                            ov[c][2] = __shfl_sync(0xffffffffu, tapv, c * 16 + l0 + 4); ov[c][3] = __shfl_sync(0xffffffffu, tapv, c * 16 + l0 + 5);
                        }
                    } else {
#pragma unroll
                        for (int c = 0; c < 2; ++c) {
                            ov[c][0] = ld1(off_tile + c * N + tp.i00); ov[c][1] = ld1(off_tile + c * N + tp.i01);
                            ov[c][2] = ld1(off_tile + c * N + tp.i10); ov[c][3] = ld1(off_tile + c * N + tp.i11);
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
                    SjM[2][0] = __shfl_sync(0xffffffffu, a4s, 0); SjM[2][1] = __shfl_sync(0xffffffffu, a4s, 8);
                    SjM[3][0] = __shfl_sync(0xffffffffu, a4s, 16); SjM[3][1] = __shfl_sync(0xffffffffu, a4s, 24);
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
                // ---- constants of the gradient pass ------------------------------------------------------------------
                {
                    const float kc = c6 * kLn2;
                    const float gvar = (P.lam[3] * gscale) * kb * 2.f * (mV - P.sigma) * P.inv_n;
                    const unsigned flags = kFHeavy | (flat ? kFFlat : 0u) | (g_live ? kFLive : 0u) | (fastsig ? kFFastSig : 0u);
                    if (lane == 0) {
                        float4* c4p = reinterpret_cast<float4*>(cb);
                        c4p[0] = make_float4(iZ, ml, cx, cy);
                        c4p[1] = make_float4(c1, c4, k4, fxx);
                        c4p[2] = make_float4(fyy, flat ? kc : c6, flat ? kc * tbar : -c6 * pa, gvar);
                        c4p[3] = make_float4(__uint_as_float(flags), ex2(ml), 0.f, 0.f);
                        c4p[8] = make_float4(cj[0], cj[1], cj[2], cj[3]);
                    }
                    if (g_live && lane < 16) {
                        float g = -cst;
#pragma unroll
                        for (int n = 0; n < 4; ++n) if ((lane >> n) & 1) g += cj[n];
                        cb[16 + lane] = g;
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(flat ? cfull + b : p2done);
                    if (!flat) ++np2;
                }
                // ---- off the critical path: loss numerators, offset gradient taps ----------------------------------------
                if (lane == 0) {
                    const float peak_t = (cx - gx) * (cx - gx) + (cy - gy) * (cy - gy);
                    const float var_t = (sd - P.sigma) * (sd - P.sigma) + ((has_var || has_vmean) ? (mV - P.sigma) * (mV - P.sigma) : 0.f);
                    if (GRADS && has_vmean) A.grad_var_mean[tile] = (P.lam[3] * gscale) * kb * 2.f * (mV - P.sigma);      // d(total) / d(mean_N(V))
                    const float shape_t = (Ent - P.e_star) * (Ent - P.e_star);
                    float4* p = reinterpret_cast<float4*>(A.partial + (size_t)tile * 8);
                    p[0] = make_float4(wa * (mse_sum * P.inv_n), wa * (0.5f * sl1), wa * peak_t, w * var_t);
#if PIPE_TAIL_OUTSIDE
                    const bool taps_here = !decode;
                    p[1] = make_float4(pair_loss, w * shape_t, (GRADS && decode) ? h2c * sl1p[0] : 0.f, (GRADS && decode) ? h2c * sl1p[1] : 0.f);
#else
                    const bool taps_here = true;
                    p[1] = make_float4(pair_loss, w * shape_t, 0.f, 0.f);
#endif
                    if (GRADS && taps_here) {
                        // the zero fill of these addresses was issued by the compute warps before they published their
                        // sums (release) and this warp has waited for that (acquire): these stores come after
                        Elem* go = reinterpret_cast<Elem*>(A.grad_off) + (size_t)tile * 2 * N;
#pragma unroll
                        for (int ch = 0; ch < 2; ++ch) {
                            Elem* o = go + ch * N;
                            const float gc = h2c * sl1p[ch];
                            IO::st1(o + tp.i00, gc * tp.w00);
                            if (tp.okx != 0.f) IO::st1(o + tp.i01, gc * tp.w01);
                            if (tp.oky != 0.f) IO::st1(o + tp.i10, gc * tp.w10);
                            if (tp.okx != 0.f && tp.oky != 0.f) IO::st1(o + tp.i11, gc * tp.w11);
                        }
                    }
                }
            }
            // ---- keypoint decode: local refinement + offset correction from the soft-argmax -------------------------
            // (steps 3-6 of decode_device.cuh's refine_and_correct with the per-call constants hoisted and, for windows of
            // at most 32 pixels, one window pixel per lane: the same operations in the same order, a sixth of the code —
            // this warp's instruction footprint competes with the compute warps' loops for the 32 KB L1.5 I-cache)
#if PIPE_TAIL_OUTSIDE
            if (decode) {
                if (lane == 0) { A.coords[2 * tile] = cx; A.coords[2 * tile + 1] = cy; A.scores[tile] = m; }
            }
#else
            if (decode) {
                float dx_ = cx, dy_ = cy;
                if (win_small) {
                    const int px = (int)fminf(fmaxf(rintf(cx), 0.f), (float)(W - 1));
                    const int py = (int)fminf(fmaxf(rintf(cy), 0.f), (float)(H - 1));
                    const int x = px + wdx, y = py + wdy;
                    const bool ok = in_win && x >= 0 && x < W && y >= 0 && y < H;
                    const float vpx = ok ? ld1(hm + (size_t)tile * N + y * W + x) : -INFINITY;
                    const float vmax = warp_max(vpx);
                    const float e = ok ? expf(vpx - vmax) : 0.f;
                    float sw4[4] = {e, e * (float)x, e * (float)y, 0.f};
                    warp_scatter_sum<4>(sw4, lane);
                    const float se = __shfl_sync(0xffffffffu, sw4[0], 0), sx = __shfl_sync(0xffffffffu, sw4[0], 8), sy = __shfl_sync(0xffffffffu, sw4[0], 16);
                    dx_ = a_blend * cx + (1.f - a_blend) * (sx / se);
                    dy_ = a_blend * cy + (1.f - a_blend) * (sy / se);
                    if (A.dflags & GBCODEC_DECODE_APPLY_OFFSET) {
                        const Bilinear bl = bilinear_setup(dx_, dy_, H, W);
                        float ox, oy;
                        {
                            ox = bilinear_read(off_tile, bl, W);
                            oy = bilinear_read(off_tile + N, bl, W);
                        }
                        dx_ += fw_dec * ox;
                        dy_ += fw_dec * oy;
                    }
                } else {
                    int px, py;
                    refine_and_correct<Elem>(hm + (size_t)tile * N, nullptr, off_tile, A.alpha_param, A.fusion_weight,
                                              H, W, A.radius, A.dflags, dx_, dy_, px, py);
                }
                if (lane == 0) { A.coords[2 * tile] = dx_; A.coords[2 * tile + 1] = dy_; A.scores[tile] = m; }
            }
#endif
        }
        return;
    }

    // ======================================================================================================
    // compute warps: front(i), then back(i - 1)
    // ======================================================================================================
    // thread geometry: four columns x0 .. x0+3 of rows ty, ty + ROWS, ...
    const int tx = tid % W4, ty = tid / W4;
    const int x0 = tx << 2;
    const float fx0 = (float)x0, fty = (float)ty;
    constexpr float ax = 0.5f * (float)(W - 1), ay = 0.5f * (float)(H - 1);
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const f2 kL2E = splat2(kLog2e), kNL2E = splat2(-kLog2e), kOne = splat2(1.f);

    unsigned rq = 0;                        // ring items consumed so far (buffer = rq % RD, parity = (rq / RD) & 1)
    unsigned ph_full = 0, np2 = 0;
    unsigned wprev[NIT];                    // tie words of the tile whose back half is due
#pragma unroll
    for (int q = 0; q < NIT; ++q) wprev[q] = 0u;
    bool tie_prev = false;
#if PIPE_CARRY_TARGET
    int hit_prev = -1;
    float4 thit_prev = z4;
#endif

    for (unsigned i = 0;; ++i) {
        const unsigned s = i % 3u, b = i & 1u;
        unsigned words[NIT];                  // one byte per pixel of the float4: (tie pattern over the partners) << 2
#pragma unroll
        for (int q = 0; q < NIT; ++q) words[q] = 0u;
        bool tie_now = false;
#if PIPE_CARRY_TARGET
        int hit_now = -1;
        float4 thit_now = z4;
#endif
        // ================================================== front(i) ==================================================
        mbar_wait(hfull + s, (ph_full >> s) & 1u); ph_full ^= 1u << s;
        const int tile = tids[s];
        if (tile >= 0) {
            const Vec* const Hs = Hb + s * N4;
            const TileDesc* dsc = Db + s;
            const int4 gq = dsc->geom;
            const float w = dsc->w;
            const int nact = (int)(dsc->pk & 7u);
            const bool heavy = (w != 0.f) || !P.use_target_weight;

            // the offset gradient is zero except on (up to) four taps per channel, patched by the scalar warp
            if (GRADS) {
                constexpr int kZ = HALF ? NIT : 2 * NIT;                // 16-byte stores per thread: 2N elements per tile
                float4* go4 = reinterpret_cast<float4*>(A.grad_off) + (size_t)tile * (kZ * TPB) + tid;
#pragma unroll (2 * kPUs)
                for (int it = 0; it < kZ; ++it) stg_stream(go4 + it * TPB, z4);
            }
            // ---- maximum (and minimum) of the tile: per warp ------------------------------------------------------
            float mw = -INFINITY, mnw = INFINITY;
#pragma unroll kPUs
            for (int it = 0; it < NIT; ++it) {
                const float4 o = IO::lds(Hs + it * TPB + tid);
                mw = max3f(mw, o.x, o.y); mw = max3f(mw, o.z, o.w);
                mnw = min3f(mnw, o.x, o.y); mnw = min3f(mnw, o.z, o.w);
            }
            mw = warp_max(mw);
            if (heavy) mnw = warp_min(mnw);
            const float nml_w = -mw * kLog2e;

            // ---- on-the-fly target: the one row (if any) in which this thread meets the patch ---------------------
            int hit_it = -1;
            float4 thit = z4;
            if (heavy) hit_it = target_row<ROWS, NIT>(gq, w, x0, ty, P.ec, lut, thit);
#if PIPE_CARRY_TARGET
            hit_now = hit_it; thit_now = thit;
#endif

            // ---- pass B: softmax moments (relative to the warp's maximum), entropy sum, sigmoid, relu moments about the
            //      tile centre, sum h^2 -----------------------------------------------------------------------------
            float r16[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) r16[q] = 0.f;
            const bool sig = heavy && (nact > 0 || PIPE_MERGE_B1);
            (void)sig;
            // the sigmoid slots still hold the lanes' sums of tile i - 1 until the scalar warp has taken them
            if (i >= 1) mbar_wait(sempty + ((i - 1) & 1u), ((i - 1) >> 1) & 1u);
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
                    constexpr int kUnroll = MODE == 2 ? kPUb : 1;
#pragma unroll kUnroll
                    for (int it = 0; it < NIT; ++it) {
                        const float4 o = IO::lds(Hs + it * TPB + tid);
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
#if !PIPE_MERGE_B1
                else if (!sig) body(std::integral_constant<int, 1>{});
#endif
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
                        const float4 o = IO::lds(Hs + hit_it * TPB + tid);
                        const float d0 = o.x - thit.x, d1_ = o.y - thit.y, d2 = o.z - thit.z, d3 = o.w - thit.w;
                        h2 += (fmaf(d0, d0, -o.x * o.x) + fmaf(d1_, d1_, -o.y * o.y)) + (fmaf(d2, d2, -o.z * o.z) + fmaf(d3, d3, -o.w * o.w));
                    }
                    r16[10] = h2;
                }
            }

            // ---- limb partners: one visit each; sums for the overlap ratio, one tie bit per pixel and partner ------
            float mind = INFINITY;                // smallest |own - partner| logit difference seen (0 = a tie)
            float r4[4] = {0.f, 0.f, 0.f, 0.f};
            if (heavy) {
                // ---- variance tile: its sum only ---------------------------------------------------------------------
                if (has_var) {
                    const unsigned q = rq & (unsigned)(RD - 1);
                    mbar_wait(rfull + q, (rq / RD) & 1u);
                    const Vec* Vs = Rb + q * N4;
                    f2 V2 = splat2(0.f);
#pragma unroll kPUs
                    for (int it = 0; it < NIT; ++it) {
                        const f4 v = as_f4(IO::lds(Vs + it * TPB + tid));
                        V2 = add2(V2, add2(v.a, v.b));
                    }
                    r16[5] = hsum2(V2);
                    ++rq;
                    __syncwarp();
                    if (lane == 0) mbar_arrive(rempty + q);
                }
                for (int n = 0; n < nact; ++n) {
                    const unsigned q = rq & (unsigned)(RD - 1);
                    mbar_wait(rfull + q, (rq / RD) & 1u);
                    const Vec* Qs = Rb + q * N4;
                    f2 Sj2 = splat2(0.f), M2 = splat2(0.f);
#pragma unroll kPUp
                    for (int it = 0; it < NIT; ++it) {
                        const float4 q4 = IO::lds(Qs + it * TPB + tid);
                        const float4 o = IO::lds(Hs + it * TPB + tid);
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
                            // the word of this iteration sits in words[0]; rotate (NIT rotations restore the order)
                            const unsigned w0 = words[0] | (tw << n);
#pragma unroll
                            for (int q = 0; q + 1 < NIT; ++q) words[q] = words[q + 1];
                            words[NIT - 1] = w0;
                        }
                        Sj2 = add2(Sj2, add2(pack2(sq[0], sq[1]), pack2(sq[2], sq[3])));
                        M2 = add2(M2, add2(pack2(sel[0], sel[1]), pack2(sel[2], sel[3])));
                    }
                    const float Sj = hsum2(Sj2), M = hsum2(M2);
                    if (n == 0) { r16[11] = Sj; r16[12] = M; }
                    if (n == 1) { r16[13] = Sj; r16[14] = M; }
                    if (n == 2) { r4[0] = Sj; r4[1] = M; }
                    if (n == 3) { r4[2] = Sj; r4[3] = M; }
                    ++rq;
                    __syncwarp();
                    if (lane == 0) mbar_arrive(rempty + q);          // this warp has consumed the buffer
                }
            }
            tie_now = GRADS && mind == 0.f;

            // ---- publish this warp's sums --------------------------------------------------------------------------
            float* const red2b = red2 + b * (NW * 4);
            float* const redMb = redM + b * (NW * 2);
#pragma unroll
            for (int q = 0; q < 4; ++q) Sb[q * TPB + tid] = make_float4(r16[4 * q], r16[4 * q + 1], r16[4 * q + 2], r16[4 * q + 3]);
            if (nact > 2) {
                warp_scatter_sum<4>(r4, lane);
                if ((lane & 7) == 0) red2b[warp * 4 + (lane >> 3)] = r4[0];
            }
            if (lane == 0) { redMb[warp * 2] = mw; redMb[warp * 2 + 1] = mnw; }
            __syncwarp();
            if (lane == 0) mbar_arrive(sfull + b);
        }

        // ================================================== back(i - 1) ===============================================
        if (i >= 1) {
            const unsigned sp = (i - 1) % 3u, bp = (i - 1) & 1u;
            const Vec* const Hs = Hb + sp * N4;
            const int tile_p = tids[sp];
            const float* const cb = cons + bp * kConsFloats;
            mbar_wait(cfull + bp, ((i - 1) >> 1) & 1u);
            float4 k0 = lds4(cb), k3 = lds4(cb + 12);
            unsigned flags = __float_as_uint(k3.x);
            if ((flags & kFHeavy) && !(flags & kFFlat)) {
                // ---- general second pass: entropy sums with the per-pixel log, relu moments about (cx, cy) --------------
                const float iZ = k0.x, ml = k0.y, cx = k0.z, cy = k0.w;
                float r8[8];
                {
                    const f2 kIZ = splat2(iZ), kEps2 = splat2(kEps), kNML = splat2(-ml);
                    f2 A1 = splat2(0.f), A2 = splat2(0.f), R01 = splat2(0.f), R23 = splat2(0.f);
                    float Ry = 0.f, Ry2 = 0.f;
                    const float dy0 = fty - cy;
#pragma unroll 1
                    for (int it = 0; it < NIT; ++it) {
                        const float4 o = IO::lds(Hs + it * TPB + tid);
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
                __syncwarp();
                if (lane == 0) mbar_arrive(p2full);
                mbar_wait(p2done, np2 & 1u);
                ++np2;
                k0 = lds4(cb); k3 = lds4(cb + 12);
                flags = __float_as_uint(k3.x);
            }
            if (GRADS) {
                Vec* gh4 = reinterpret_cast<Vec*>(A.grad_hm) + (size_t)tile_p * N4 + tid;
                Vec* gv4 = has_var ? reinterpret_cast<Vec*>(A.grad_var) + (size_t)tile_p * N4 + tid : nullptr;
                if (!(flags & kFHeavy)) {
#pragma unroll
                    for (int it = 0; it < NIT; ++it) {
                        IO::stg(gh4 + it * TPB, z4);
                        if (gv4) IO::stg(gv4 + it * TPB, z4);
                    }
                } else {
                    const float4 k1 = lds4(cb + 4), k2 = lds4(cb + 8);
                    const float iZ = k0.x, ml = k0.y, cx = k0.z, cy = k0.w;
                    const float c1 = k1.x, c4 = k1.y, k4 = k1.z, fxx = k1.w;
                    const float fyy = k2.x, ke = k2.y, addc = k2.z, gvar = k2.w;
                    const bool flat = (flags & kFFlat) != 0u, g_live = (flags & kFLive) != 0u;
                    const TileDesc* dsc = Db + sp;
#if PIPE_CARRY_TARGET
                    const float4 thit = thit_prev;
                    const int hit_it = hit_prev;
#else
                    float4 thit;
                    const int hit_it = target_row<ROWS, NIT>(dsc->geom, dsc->w, x0, ty, P.ec, lut, thit);
#endif
                    // ---- pass D: the heatmap gradient ---------------------------------------------------------------------
                    // g = c1 (h - t) + p (c6 (a - pa) + (x - cx) Fx + (y - cy) Fy) + [h > 0] (c4 ((x-cx)^2 + (y-cy)^2) + k4) + overlap
                    // flat tiles: c6 (a - pa) = c6 ln2 (tbar - t_i)
                    float dxj[4], dx2j[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) { dxj[j] = (fx0 + (float)j) - cx; dx2j[j] = dxj[j] * dxj[j]; }
                    const float dy0 = fty - cy;
                    const f2 kIZ = splat2(iZ), kNML = splat2(-ml), kC1 = splat2(c1), kC4 = splat2(c4), kK4 = splat2(k4);
                    const f2 dx2_01 = pack2(dx2j[0], dx2j[1]), dx2_23 = pack2(dx2j[2], dx2j[3]);
                    const float4 gv = make_float4(gvar, gvar, gvar, gvar);
                    const float* const tab = cb + 16;
                    // sigmoid'(h) = x / (1 + x)^2 with x = exp(h) = e * exp(m); the plain sigmoid where exp(m) is not representable
                    const f2 kEM = splat2(k3.y);
                    const bool fastsig = (flags & kFFastSig) != 0u;
                    auto pass_d = [&](auto flat_c) {
                        constexpr bool FLAT = decltype(flat_c)::value;
                        const f2 kNKE = splat2(-ke), kEps2 = splat2(kEps), kLn2v = splat2(kLn2);
                        const f2 base01 = pack2(fmaf(dxj[0], fxx, addc), fmaf(dxj[1], fxx, addc)), base23 = pack2(fmaf(dxj[2], fxx, addc), fmaf(dxj[3], fxx, addc));
                        constexpr int kUnrollD = FLAT ? kPUd : 1;
#pragma unroll kUnrollD
                        for (int it = 0; it < NIT; ++it) {
                            const float4 o = IO::lds(Hs + it * TPB + tid);
                            const f4 hv = as_f4(o);
                            const f2 t01 = fma2(hv.a, kL2E, kNML), t23 = fma2(hv.b, kL2E, kNML);
                            const f2 e01 = pack2(ex2(lo2(t01)), ex2(hi2(t01))), e23 = pack2(ex2(lo2(t23)), ex2(hi2(t23)));
                            const f2 p01 = mul2(e01, kIZ), p23 = mul2(e23, kIZ);
                            const float dy = dy0 + (float)(it * ROWS);
                            const f2 fyd = splat2(dy * fyy), dy2 = splat2(dy * dy);
                            f2 in01 = add2(base01, fyd), in23 = add2(base23, fyd);
                            if (FLAT) {
                                in01 = fma2(t01, kNKE, in01); in23 = fma2(t23, kNKE, in23);
                            } else {
                                // an = ln2 lg2(p + eps) + p / (p + eps) = -a
                                const f2 u01 = add2(p01, kEps2), u23 = add2(p23, kEps2);
                                const f2 l01 = pack2(lg2(lo2(u01)), lg2(hi2(u01))), l23 = pack2(lg2(lo2(u23)), lg2(hi2(u23)));
                                const f2 c01 = pack2(rcp(lo2(u01)), rcp(hi2(u01))), c23 = pack2(rcp(lo2(u23)), rcp(hi2(u23)));
                                const f2 an01 = fma2(kLn2v, l01, mul2(p01, c01)), an23 = fma2(kLn2v, l23, mul2(p23, c23));
                                in01 = fma2(kNKE, an01, in01); in23 = fma2(kNKE, an23, in23);
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
                                f2 sd01, sd23;                     // sigmoid (1 - sigmoid)
                                if (fastsig) {
                                    const f2 x01 = mul2(e01, kEM), x23 = mul2(e23, kEM);
                                    const f2 y01 = add2(x01, kOne), y23 = add2(x23, kOne);
                                    const f2 r01 = pack2(rcp(lo2(y01)), rcp(hi2(y01))), r23 = pack2(rcp(lo2(y23)), rcp(hi2(y23)));
                                    sd01 = mul2(mul2(x01, r01), r01); sd23 = mul2(mul2(x23, r23), r23);
                                } else {
                                    const f2 s01 = pack2(sigmoid_fast(o.x), sigmoid_fast(o.y)), s23 = pack2(sigmoid_fast(o.z), sigmoid_fast(o.w));
                                    const f2 kNeg = splat2(-1.f);
                                    sd01 = fma2(mul2(s01, kNeg), s01, s01); sd23 = fma2(mul2(s23, kNeg), s23, s23);
                                }
                                const unsigned wd = wprev[0];
#pragma unroll
                                for (int q = 0; q + 1 < NIT; ++q) wprev[q] = wprev[q + 1];
                                wprev[NIT - 1] = wd;
                                float G[4];
#pragma unroll
                                for (int j = 0; j < 4; ++j)
                                    G[j] = *reinterpret_cast<const float*>(reinterpret_cast<const char*>(tab) + ((wd >> (8 * j)) & 0xFCu));
                                g01 = fma2(pack2(G[0], G[1]), sd01, g01);
                                g23 = fma2(pack2(G[2], G[3]), sd23, g23);
                            }
                            IO::stg(gh4 + it * TPB, as_float4(f4{g01, g23}));
                            if (gv4) IO::stg(gv4 + it * TPB, gv);
                        }
                    };
                    if (flat) pass_d(std::true_type{}); else pass_d(std::false_type{});

                    // rare: a logit of this thread equals its partner's — ATen's minimum splits that gradient evenly.
                    // Patch the pixels this thread has just written (same thread, program order).
                    if (g_live && tie_prev) {
                        const float4 cj4 = lds4(cb + 32);
                        const float cjv[4] = {cj4.x, cj4.y, cj4.z, cj4.w};
                        const int bimg = tile_p / P.K;
                        const unsigned pj = dsc->pj;
                        const int nact_p = (int)(dsc->pk & 7u);
                        for (int n = 0; n < nact_p; ++n) {
                            const float cjp = cjv[n == 0 ? 0 : (n == 1 ? 1 : (n == 2 ? 2 : 3))];
                            if (cjp == 0.f) continue;
                            const Vec* src = reinterpret_cast<const Vec*>(hm) + ((size_t)bimg * P.K + ((pj >> (8 * n)) & 0xFFu)) * N4 + tid;
                            for (int it = 0; it < NIT; ++it) {
                                const float4 q = IO::ldg(src + it * TPB), o = IO::lds(Hs + it * TPB + tid);
                                if (q.x == o.x || q.y == o.y || q.z == o.z || q.w == o.w) {
                                    float4 g = IO::ld_plain(gh4 + it * TPB);
                                    float sg;
                                    if (q.x == o.x) { sg = sigmoid_fast(o.x); g.x = fmaf(0.5f * cjp * sg, 1.f - sg, g.x); }
                                    if (q.y == o.y) { sg = sigmoid_fast(o.y); g.y = fmaf(0.5f * cjp * sg, 1.f - sg, g.y); }
                                    if (q.z == o.z) { sg = sigmoid_fast(o.z); g.z = fmaf(0.5f * cjp * sg, 1.f - sg, g.z); }
                                    if (q.w == o.w) { sg = sigmoid_fast(o.w); g.w = fmaf(0.5f * cjp * sg, 1.f - sg, g.w); }
                                    IO::st_plain(gh4 + it * TPB, g);
                                }
                            }
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) { mbar_arrive(cempty + bp); mbar_arrive(hempty + sp); }
        }
        if (tile < 0) break;
#pragma unroll
        for (int q = 0; q < NIT; ++q) wprev[q] = words[q];
        tie_prev = tie_now;
#if PIPE_CARRY_TARGET
        hit_prev = hit_now; thit_prev = thit_now;
#endif
    }
}

// ---- launcher ------------------------------------------------------------------------------------------
template <int W4, int ROWS, int NIT, int MINB, bool GRADS, bool HALF = false, int RD = 2>
int launch_pipe_t(const LossParams& P, const LossArgs& A, cudaStream_t s, cudaEvent_t e0, cudaEvent_t e1) {
    using L = PipePlan<W4, ROWS, NIT, HALF, RD>;
    const size_t smem = (size_t)L::oLut + (size_t)((P.ec.lut_size + 3) & ~3) * 4;
    auto kern = step_pipe_kernel<W4, ROWS, NIT, MINB, GRADS, HALF, RD>;
    int dev = 0, sms = 0, max_optin = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (max_optin > 0 && smem > (size_t)max_optin) return 1;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return fail(GBCODEC_ERR_CUDA, "cudaFuncSetAttribute(step_pipe_kernel): %s", cudaGetErrorString(e));
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, L::TPB + 64, smem);
    if (e != cudaSuccess || per_sm <= 0) return fail(GBCODEC_ERR_CUDA, "step_pipe_kernel does not fit an SM (%zu bytes of shared memory)", smem);
    const int tiles = P.B * P.K;
    // GBCODEC_PIPE_CTAS=<n>: fewer resident CTAs per SM than fit (measurement only: how the kernel scales with the
    // number of tile pipelines in flight)
    if (const char* lim = getenv("GBCODEC_PIPE_CTAS")) { const int n = atoi(lim); if (n >= 1 && n < per_sm) per_sm = n; }
    const int grid = tiles < sms * per_sm ? tiles : sms * per_sm;
    if (e0) cudaEventRecord(e0, s);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(L::TPB + 64);                  // compute warps + the producer warp + the scalar warp
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    note_launch();
    e = cudaLaunchKernelEx(&cfg, kern, P, A);
    if (e != cudaSuccess) return fail(GBCODEC_ERR_CUDA, "cudaLaunchKernelEx(step_pipe_kernel): %s", cudaGetErrorString(e));
    if (e1) cudaEventRecord(e1, s);
    return check_launch("step_pipe_kernel");
}

}  // namespace

// GBCODEC_STEP_KERNEL=tile | persist selects the one-CTA-per-tile kernel of loss_tile.cu / the persistent kernel of
// step_tile.cu instead (A/B measurements, and tests that compare entry points bit for bit).  Read on every call.
bool step_pipe_tail_outside() { return PIPE_TAIL_OUTSIDE != 0; }

int launch_step_pipe(const LossParams& P, const LossArgs& A, cudaStream_t s, cudaEvent_t e0, cudaEvent_t e1) {
    const char* env = getenv("GBCODEC_STEP_KERNEL");
    if (env && (!strcmp(env, "tile") || !strcmp(env, "persist"))) return 1;
    // target generated on the fly with a patch no taller than the CTA's rows, the ordinary forward (+ backward) call;
    // float32 maps, or float16 maps (autocast) with a four-deep ring
    if (A.target || A.lam_eff || A.plan || !A.desc || !A.tile_counter) return 1;
    if ((A.var_mean || A.grad_var_mean) && (A.half_io || A.var)) return 1;          // per-tile variance means: float32 maps only
    if (A.coords && A.radius > 8) return 1;
    const bool grads = A.grad_hm != nullptr;
    if (P.H == 64 && P.W == 48 && P.ec.ntap <= 16) {
        if (A.half_io) {
            // GBCODEC_STEP_F16=tile keeps float16 maps on loss_tile_kernel<..., HALF> (A/B measurements)
            const char* h = getenv("GBCODEC_STEP_F16");
            if (h && !strcmp(h, "tile")) return 1;
            return grads ? launch_pipe_t<12, 16, 4, 3, true, true, PIPE_HALF_RING>(P, A, s, e0, e1)
                         : launch_pipe_t<12, 16, 4, 3, false, true, PIPE_HALF_RING>(P, A, s, e0, e1);
        }
        return grads ? launch_pipe_t<12, 16, 4, 3, true>(P, A, s, e0, e1) : launch_pipe_t<12, 16, 4, 3, false>(P, A, s, e0, e1);
    }
    return 1;
}

}  // namespace gbc
