// loss_common.cuh — descriptors shared by the loss kernels (loss.cu: host side, pre/post
// kernels and the generic tile kernel; loss_tile.cu: the register-resident tile kernel).
#pragma once
#include "common.cuh"
#include "decode_device.cuh"
#include <string.h>

namespace gbc {

// ---- kernel-side description -------------------------------------------------------
struct LossParams {
    int B, K, H, W;
    float in_w, in_h;
    float lam[6];
    float sigma;            // target sigma
    float e_star;           // log(2 pi e sigma^2)
    float inv_n;            // 1 / (H*W)
    float sx, sy;           // W / in_w, H / in_h (fusion_head.py:679-684)
    int use_target_weight;
    int n_pairs;
    EncodeConst ec;
    int8_t n_partner[GBCODEC_MAX_K];
    int8_t partner[GBCODEC_MAX_K][GBCODEC_MAX_PARTNERS];
    uint8_t owner[GBCODEC_MAX_K];       // bit p: this channel is the first index of the pair with partner p
    int16_t pair_i[GBCODEC_MAX_PAIRS], pair_j[GBCODEC_MAX_PAIRS];
};

struct LossArgs {
    const float* hm; const float* off; const float* var; const float* target;
    const float* weight; const float* gt;
    const float* grad_scale;            // device scalar or null
    float* grad_hm; float* grad_off; float* grad_var;
    // fused decode (null coords = off)
    const float* alpha_param; const float* fusion_weight; float* coords; float* scores;
    int radius; unsigned dflags;
    // workspace
    const double* sums;                 // [2] raw sums of w and w_i*w_j
    const float* weff;                  // [B*K] weights after the encoder's rule
    const int4* geom;                   // [B*K] packed patch geometry of the on-the-fly target (pack_geom)
    float* partial;                     // [B*K][8] un-normalised per-tile loss numerators
    const float* lam_eff;               // backward recompute: device [6] per-term multipliers
    const int* plan;                    // backward recompute: run only if *plan == 2
    int half_io;                        // hm, off, var and the three gradients are float16 (the pointers are then __half*)
    // the variance branch reduced to its per-tile mean by the caller (a head that fuses the mean into its last
    // convolution's epilogue): replaces var / grad_var, 16N instead of 24N bytes per tile
    const float* var_mean;              // (B,K) or null
    float* grad_var_mean;               // (B,K): d(total)/d(mean_N(V))
    // persistent step kernel (step_tile.cu): per-tile descriptors written by denoms_kernel, dynamic tile counter
    const struct TileDesc* desc;
    unsigned* tile_counter;
    unsigned* sm_slots;                 // [kSmSlots] running count of step CTAs started per SM (never reset: only its value mod 3 is used)
};

// Everything the persistent step kernel needs to know about one tile besides its pixels, in one 64-byte record
// (one bulk copy per tile instead of a chain of dependent scalar loads).  Written by denoms_kernel.
struct __align__(16) TileDesc {
    int4 geom;            // packed patch geometry of the on-the-fly target (pack_geom)
    float w;              // weight after the encoder's rule
    float gx, gy;         // ground-truth keypoint in heatmap pixels (fusion_head.py:679-684)
    unsigned pk;          // bits 0-2: number of ACTIVE limb partners (both weights non-zero), bits 4-7: owner bits (compacted)
    float wj[4];          // weights of the active partners, in partner order
    unsigned pj;          // channels of the active partners, one byte each
    unsigned pad[3];
};
static_assert(sizeof(TileDesc) == 64, "one 64-byte bulk copy per tile");

constexpr int kFinBlocks = 32;          // CTAs of the second-stage reduction
constexpr int kSmSlots = 512;           // per-SM counters of the step kernel (step_pipe.cu: which CTA of its SM a CTA is)
constexpr int kWsHeaderFloats = 1024;   // sums (2 doubles), plan, ticket, lam_eff, second-stage partials (2 KB); per-SM counters (2 KB)
struct WsLayout {
    double* sums; int* plan; unsigned* ticket; unsigned* tile_counter; float* lam_eff; double* bpart; float* weff; int4* geom; float* partial;
    TileDesc* desc;
    unsigned* sm_slots;
};
static inline size_t ws_bytes(int B, int K) {
    return (size_t)(kWsHeaderFloats + (size_t)B * K * (13 + 16) + 8 + 16) * sizeof(float);
}
static inline WsLayout ws_carve(void* ws, int B, int K) {
    float* f = reinterpret_cast<float*>(ws);
    WsLayout l;
    l.sums = reinterpret_cast<double*>(f);          // f[0..3]
    l.plan = reinterpret_cast<int*>(f + 4);         // f[4]
    l.ticket = reinterpret_cast<unsigned*>(f + 5);  // f[5]
    l.tile_counter = reinterpret_cast<unsigned*>(f + 6);   // f[6]: inside the 32 bytes every call clears
    l.lam_eff = f + 8;                              // f[8..15]
    l.bpart = reinterpret_cast<double*>(f + 64);    // kFinBlocks * 6 doubles
    l.sm_slots = reinterpret_cast<unsigned*>(f + 512);   // kSmSlots words
    // geom rows are 16 bytes and partial rows 32 bytes: keep both aligned
    const size_t tiles = (size_t)B * K, tiles8 = (tiles + 7) & ~(size_t)7;
    l.geom = reinterpret_cast<int4*>(f + kWsHeaderFloats);
    l.partial = f + kWsHeaderFloats + 4 * tiles8;
    l.weff = l.partial + 8 * tiles;
    // 64-byte records, 16-byte aligned (the workspace is)
    l.desc = reinterpret_cast<TileDesc*>(f + kWsHeaderFloats + 4 * tiles8 + 8 * tiles + ((tiles + 3) & ~(size_t)3));
    return l;
}

// patch geometry in 16 bytes: origin, then [from, to) ranges packed as from | to << 16
__device__ __forceinline__ int4 pack_geom(const PatchGeom& g) {
    return make_int4(g.ulx, g.uly, g.active ? (g.x_from | (g.x_to << 16)) : 0, g.active ? (g.y_from | (g.y_to << 16)) : 0);
}
__device__ __forceinline__ PatchGeom unpack_geom(const int4& v, float weight) {
    PatchGeom g;
    g.ulx = v.x; g.uly = v.y;
    g.x_from = v.z & 0xffff; g.x_to = v.z >> 16;
    g.y_from = v.w & 0xffff; g.y_to = v.w >> 16;
    g.weight = weight;
    g.active = (g.x_to > g.x_from) && (g.y_to > g.y_from);
    return g;
}

// ---- peer-memory exchange of a batch-sharded job (one process per GPU, NVLink / NVSwitch) -----------
// Every rank owns one PeerMail in its own HBM; peers write into it with plain stores over NVLink
// (cudaIpc-mapped pointers) and publish with a system-scope release store of the call's sequence
// number; the owner polls its LOCAL copy with acquire loads.  Two parity slots: a peer can be at
// most one call ahead (its call s+1 needs this rank's data of call s+1).
struct PeerMail {
    double den[2][GBCODEC_MAX_PEERS][2];                 // [parity][source rank]{sum w, sum w_i w_j}
    unsigned long long den_seq[2][GBCODEC_MAX_PEERS];
    float loss[4][GBCODEC_MAX_PEERS][8];                 // [seq & 3][source rank] six loss terms; four slots: a deferred reader
    unsigned long long loss_seq[4][GBCODEC_MAX_PEERS];   // (gbcodec_peer_collect_losses_f32) may be two steps behind the writers
    unsigned int timeouts;                               // bounded spins that gave up (a peer died)
};
struct PeerView {
    PeerMail* mail[GBCODEC_MAX_PEERS];                   // mail[rank] is the local one
    int rank, world;
    unsigned long long seq;                              // sequence number of the step (loss exchange)
    unsigned long long den_seq;                          // ... of the normaliser exchange (it may run one step ahead of the steps)
    int defer;                                           // the step publishes its loss terms and does not wait for the peers'
    unsigned long long timeout_ns;                       // how long a kernel waits for a peer before it gives up
    unsigned int* failed;                                // sticky flag in mapped host memory (device alias)
};
struct PeerCtx {
    PeerView view;
    cudaIpcMemHandle_t handle;
    int connected;
    int device;
    unsigned int* h_failed;                              // the same flag, host side: read at the entry of every sharded call
};
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// Wait until a peer has published `seq`.  Bounded by the context's time-out (a dead peer must not hang the GPU); on
// expiry the sticky flag is raised and the caller poisons what it could not complete with NaN — a mailbox slot that
// was not published for this call holds the values of two calls ago and must never be used.
__device__ __forceinline__ bool wait_seq(const unsigned long long* flag, unsigned long long seq, const PeerView& peer, unsigned int* timeouts) {
    const unsigned long long t0 = global_ns();
    for (unsigned it = 0;; ++it) {
        if (ld_acquire_sys(flag) == seq) return true;
        __nanosleep(200);
        if ((it & 63u) == 63u && global_ns() - t0 > peer.timeout_ns) break;
    }
    atomicAdd(timeouts, 1u);
    if (peer.failed) { *reinterpret_cast<volatile unsigned int*>(peer.failed) = 1u; __threadfence_system(); }
    return false;
}

// ---- bilinear helpers (same convention as decode.cu) ---------------------------------
struct Taps {
    int i00, i01, i10, i11;          // flat indices inside a channel
    float w00, w01, w10, w11;        // weights, already zero for taps outside the map
    float fx, fy, inx, iny, okx, oky;
};
__device__ __forceinline__ Taps taps_setup(float cx, float cy, int H, int W) {
    Taps t;
    const float ccx = fminf(fmaxf(cx, 0.f), (float)(W - 1));
    const float ccy = fminf(fmaxf(cy, 0.f), (float)(H - 1));
    t.inx = (cx >= 0.f && cx <= (float)(W - 1)) ? 1.f : 0.f;
    t.iny = (cy >= 0.f && cy <= (float)(H - 1)) ? 1.f : 0.f;
    const float fx0 = floorf(ccx), fy0 = floorf(ccy);
    const int x0 = (int)fx0, y0 = (int)fy0;
    t.fx = ccx - fx0; t.fy = ccy - fy0;
    t.okx = (x0 + 1 < W) ? 1.f : 0.f;
    t.oky = (y0 + 1 < H) ? 1.f : 0.f;
    const int x1 = min(x0 + 1, W - 1), y1 = min(y0 + 1, H - 1);
    t.i00 = y0 * W + x0; t.i01 = y0 * W + x1; t.i10 = y1 * W + x0; t.i11 = y1 * W + x1;
    t.w00 = (1.f - t.fx) * (1.f - t.fy);
    t.w01 = t.fx * (1.f - t.fy) * t.okx;
    t.w10 = (1.f - t.fx) * t.fy * t.oky;
    t.w11 = t.fx * t.fy * t.okx * t.oky;
    return t;
}

__device__ __forceinline__ float tie_rule(float a, float b) { return a < b ? 1.f : (a == b ? 0.5f : 0.f); }

// per-tile scalars broadcast from thread 0 to the CTA
struct TileCoef {
    float c1;        // lambda1 * wa/(Da N) * 2
    float c4;        // lambda4 * w/D * (s - sigma)/s / R'
    float v;         // variance of the tile
    float c6;        // lambda6 * w/D * 2 (E - E*)
    float pa;        // sum p a
    float fx, fy;    // d(loss)/d(cx, cy)
    float gv;        // uniform gradient of the variance map
    float go[2];     // lambda2 * wa/(2 Da) * sl1'(d_ch)
};


// loss_tile.cu: register-resident kernel for tiles whose rows split evenly over the CTA;
// returns 1 if it has no instantiation for this shape (the caller then uses the generic kernel).
int launch_loss_tile(const LossParams& P, const LossArgs& A, cudaStream_t s, cudaEvent_t ev_start, cudaEvent_t ev_stop);
// step_tile.cu: persistent kernel (bulk-copy pipeline, one block reduction per tile) for float32 maps with the target
// generated on the fly; returns 1 if it does not cover this call (the caller then uses launch_loss_tile).
int launch_step_tile(const LossParams& P, const LossArgs& A, cudaStream_t s, cudaEvent_t ev_start, cudaEvent_t ev_stop);
// step_pipe.cu: the same step as a software pipeline of specialised warps (compute warps run only pixel loops, a scalar
// warp owns the per-tile chain, a producer lane moves the bytes); same coverage and return convention as launch_step_tile.
int launch_step_pipe(const LossParams& P, const LossArgs& A, cudaStream_t s, cudaEvent_t ev_start, cudaEvent_t ev_stop);
// true if launch_step_pipe's kernel leaves the offset-gradient taps and the decode's refinement of a decoding call to
// finalize_kernel's tail CTAs (d_coords then holds the soft-argmax, words 6 and 7 of a tile's numerator row the two factors)
bool step_pipe_tail_outside();

}  // namespace gbc
