// loss.cu — six-term fusion loss, forward and backward in one pass per tile.
//
// Restates FusionPoseLoss.forward (models/fusion_head.py:745-806) with its terms
// (:637-743) and GaussianDistributionConstraint (:405-559), plus the autograd
// backward train.py:182 triggers, in closed form (DESIGN.md "Loss kernel math";
// the algebra is checked against autograd in tests/test_closed_form.py).
//
// One CTA per (image, keypoint) tile.  The tile is read from HBM once with
// 128-bit loads into shared memory and every pass (max, softmax moments,
// entropy/variance, limb overlap, gradient) runs out of shared memory; the limb
// partners' tiles are re-read through L2 (they are some other CTA's own tile, so
// HBM sees each heatmap once).  Gradients leave with 128-bit streaming stores.
// Algorithmic HBM bytes per tile: read hm, var (8N); write d_hm, d_var, d_off (16N);
// +4N when the target tiles come from HBM instead of being generated on the fly.
#include "loss_common.cuh"
#include <stdlib.h>

namespace gbc {

// ---- weights after the encoder rule + the two batch sums ----------------------------
// One thread per tile; a CTA owns 256 / K whole images so that the limb products of an image
// stay inside one CTA.  The sums are accumulated in double: with the reference's weights
// (visibility flags 0/1/2, coco_dataset.py:214) every partial sum is an exact integer, so the
// order of the two atomics per CTA does not change the result.
// `sums` may be null (the caller supplies the normalisers): the kernel then only prepares the weights, the patch
// geometry and the tile descriptors.
__global__ void __launch_bounds__(256)
denoms_kernel(const __grid_constant__ LossParams P, const float* __restrict__ weight,
              const float* __restrict__ gt, int target_given, float* __restrict__ weff, int4* __restrict__ geom,
              TileDesc* __restrict__ desc,
              double* __restrict__ sums, unsigned* __restrict__ ticket, const __grid_constant__ PeerView peer,
              float* __restrict__ global_out, const float* __restrict__ denoms_in = nullptr, double* __restrict__ sums_dst = nullptr) {
    __shared__ float wsm[256];
    __shared__ double red[2][8];
    pdl_launch_dependents();                           // the tile kernel may start its bulk loads; it waits before it reads weff / geom / sums
    const int ipb = 256 / P.K;                         // images per CTA
    const int img0 = blockIdx.x * ipb;
    const int nimg = min(ipb, P.B - img0);
    const int local = threadIdx.x;
    double sw = 0.0, sp = 0.0;
    int4 gpk = make_int4(0, 0, 0, 0);
    if (local < nimg * P.K) {
        const int t = img0 * P.K + local;
        float wk = weight[t];
        if (!target_given) {
            const PatchGeom g = patch_geometry(gt[2 * t], gt[2 * t + 1], wk, P.H, P.W, P.in_w, P.in_h, P.ec);
            wk = g.weight;
            gpk = pack_geom(g);
            if (geom) geom[t] = gpk;
        }
        wsm[local] = wk;
        if (weff) weff[t] = wk;
        sw = (double)wk;
    }
    __syncthreads();
    // the tile's descriptor for the persistent step kernel: its weight, target geometry, ground truth in heatmap pixels
    // and the limb partners that carry weight (compacted, in partner order)
    if (desc && local < nimg * P.K) {
        const int t = img0 * P.K + local;
        const int im = local / P.K, k = local - im * P.K;
        TileDesc d;
        d.geom = gpk;
        d.w = wsm[local];
        d.gx = gt ? gt[2 * t] * P.sx : 0.f; d.gy = gt ? gt[2 * t + 1] * P.sy : 0.f;
        unsigned nact = 0, own = 0, pj = 0;
        d.wj[0] = d.wj[1] = d.wj[2] = d.wj[3] = 0.f;
        for (int pi = 0; pi < P.n_partner[k]; ++pi) {
            const int j = P.partner[k][pi];
            const float wj = wsm[im * P.K + j];
            if (d.w != 0.f && wj != 0.f) {
                d.wj[nact] = wj;
                pj |= (unsigned)j << (8 * nact);
                if ((P.owner[k] >> pi) & 1) own |= 1u << nact;
                ++nact;
            }
        }
        d.pk = nact | (own << 4);
        d.pj = pj;
        d.pad[0] = d.pad[1] = d.pad[2] = 0u;
        desc[t] = d;
    }
    if (denoms_in && blockIdx.x == 0) {
        // normalisers given by the caller (a shard of a global batch): into the workspace as the tile kernel reads them,
        // and the finalize ticket / tile counter (adjacent words) cleared — what a memset + a one-warp kernel used to do
        if (threadIdx.x < 2) sums_dst[threadIdx.x] = (double)denoms_in[threadIdx.x];
        if (threadIdx.x == 2) { ticket[0] = 0u; ticket[1] = 0u; }
    }
    if (!sums) return;
    for (int q = local; q < nimg * P.n_pairs; q += 256) {
        const int im = q / P.n_pairs, p = q - im * P.n_pairs;
        sp += (double)(wsm[im * P.K + P.pair_i[p]] * wsm[im * P.K + P.pair_j[p]]);
    }
    for (int o = 16; o > 0; o >>= 1) {
        sw += __shfl_xor_sync(0xffffffffu, sw, o);
        sp += __shfl_xor_sync(0xffffffffu, sp, o);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { red[0][warp] = sw; red[1][warp] = sp; }
    __syncthreads();
    if (threadIdx.x < 2) {
        double a = 0.0;
        for (int i = 0; i < 8; ++i) a += red[threadIdx.x][i];
        atomicAdd(sums + threadIdx.x, a);
    }
    if (peer.world <= 1) return;
    // ---- batch-sharded job: the CTA that draws the last ticket holds this rank's sums; it writes them into
    // every peer's mailbox over NVLink, waits for the peers' sums in its own mailbox and leaves the GLOBAL
    // sums (added in rank order: every rank gets the same bits) in the workspace.  No NCCL on this path.
    __shared__ bool last;
    __shared__ double gath[GBCODEC_MAX_PEERS][2];
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!last) return;
    __threadfence();
    const int q = (int)(peer.den_seq & 1ull), r = threadIdx.x;
    if (r < peer.world) {
        const double sw = __ldcg(sums), sp = __ldcg(sums + 1);
        PeerMail* dst = peer.mail[r];
        dst->den[q][peer.rank][0] = sw;
        dst->den[q][peer.rank][1] = sp;
        __threadfence_system();
        st_release_sys(&dst->den_seq[q][peer.rank], peer.den_seq);
        PeerMail* mine = peer.mail[peer.rank];
        const bool ok = wait_seq(&mine->den_seq[q][r], peer.den_seq, peer, &mine->timeouts);
        gath[r][0] = ok ? mine->den[q][r][0] : (double)NAN;
        gath[r][1] = ok ? mine->den[q][r][1] : (double)NAN;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double gw = 0.0, gp = 0.0;
        for (int i = 0; i < peer.world; ++i) { gw += gath[i][0]; gp += gath[i][1]; }
        sums[0] = gw; sums[1] = gp;
        if (global_out) { global_out[0] = (float)gw; global_out[1] = (float)gp; }    // what the backward takes as d_denoms
        *ticket = 0u;
    }
}

__global__ void sums_to_float_kernel(const double* __restrict__ sums, float* __restrict__ out2) {
    if (threadIdx.x < 2) out2[threadIdx.x] = (float)sums[threadIdx.x];
}
template <int TPB, int NITER, int CACHE>
__global__ void __launch_bounds__(TPB)
loss_kernel(const __grid_constant__ LossParams P, const __grid_constant__ LossArgs A) {
    if (A.plan && *A.plan != 2) return;      // backward recompute not needed
    extern __shared__ __align__(16) float smem[];
    const int H = P.H, W = P.W, n = H * W, n4 = n >> 2, w4 = W >> 2;
    const int niter = NITER > 0 ? NITER : (n4 + TPB - 1) / TPB;
    const int tile = blockIdx.x, tid = threadIdx.x;
    const int b = tile / P.K, k = tile - b * P.K;

    float4* Hs = reinterpret_cast<float4*>(smem);
    float4* Gs = Hs + n4;
    float4* Es = Gs + n4;                                   // CACHE >= 1
    float4* As = Es + (CACHE >= 1 ? n4 : 0);                // CACHE == 2
    float4* Ss = As + (CACHE == 2 ? n4 : 0);                // CACHE == 2
    float* lut = reinterpret_cast<float*>(Ss + (CACHE == 2 ? n4 : 0));
    float* scratch = lut + ((P.ec.lut_size + 3) & ~3);      // 8*32 + 8 floats
    __shared__ PatchGeom geom_s;
    __shared__ TileCoef coef_s;

    const bool has_target = A.target != nullptr;
    const bool grads = A.grad_hm != nullptr;
    const bool backward_only = A.lam_eff != nullptr;
    const float w = __ldg(A.weff + tile);
    const float wa = P.use_target_weight ? w : 1.f;
    const bool heavy = (w != 0.f) || !P.use_target_weight;
    const bool decode = A.coords != nullptr;

    const float4* hm4 = reinterpret_cast<const float4*>(A.hm) + (size_t)tile * n4;
    const float4* var4 = A.var ? reinterpret_cast<const float4*>(A.var) + (size_t)tile * n4 : nullptr;
    const float4* tgt4 = has_target ? reinterpret_cast<const float4*>(A.target) + (size_t)tile * n4 : nullptr;

    if (!has_target) {
        fill_patch_lut(lut, P.ec);
        if (tid == 0) geom_s = patch_geometry(__ldg(A.gt + 2 * tile), __ldg(A.gt + 2 * tile + 1), w, H, W, P.in_w, P.in_h, P.ec);
    }

    // ---- load: own tile -> smem, running max; variance map -> running sum -----------
    float m = -INFINITY, vsum = 0.f;
    for (int it = 0; it < niter; ++it) {
        const int i = it * TPB + tid;
        if (NITER > 0 || i < n4) {
            const float4 v = ldg_stream(hm4 + i);
            Hs[i] = v;
            m = fmaxf(m, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
        }
    }
    if (var4 && heavy) {
        for (int it = 0; it < niter; ++it) {
            const int i = it * TPB + tid;
            if (NITER > 0 || i < n4) { const float4 v = ldg_stream(var4 + i); vsum += (v.x + v.y) + (v.z + v.w); }
        }
    }
    m = block_max(m, scratch);           // barriers: Hs, lut and geom_s are visible from here on
    const PatchGeom geom = has_target ? PatchGeom{} : geom_s;

    float gscale = A.grad_scale ? __ldg(A.grad_scale) : 1.f;
    float lam[6];
#pragma unroll
    for (int q = 0; q < 6; ++q) lam[q] = backward_only ? __ldg(A.lam_eff + q) : P.lam[q] * gscale;

    // target value for 4 consecutive pixels of row y starting at column x
    auto target4 = [&](int i, int x, int y) -> float4 {
        if (has_target) return ldg_keep(tgt4 + i);
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (geom.active && y >= geom.y_from && y < geom.y_to && x + 3 >= geom.x_from && x < geom.x_to) {
            float e[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int xx = x + j;
                e[j] = (xx >= geom.x_from && xx < geom.x_to) ? patch_value(lut, geom, P.ec, xx, y) : 0.f;
            }
            t = make_float4(e[0], e[1], e[2], e[3]);
        }
        return t;
    };

    // ---- pass B: softmax moments, sigmoid mass, relu mass, squared error ---------------
    const float ml = m * kLog2e;
    float r7[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, vsum};   // Z, X, Y, S, R, MSE, Vsum
    for (int it = 0; it < niter; ++it) {
        const int i = it * TPB + tid;
        if (NITER > 0 || i < n4) {
            const float4 h = Hs[i];
            const int y = i / w4, x = (i - y * w4) << 2;
            float4 e;
            e.x = ex2(fmaf(h.x, kLog2e, -ml)); e.y = ex2(fmaf(h.y, kLog2e, -ml));
            e.z = ex2(fmaf(h.z, kLog2e, -ml)); e.w = ex2(fmaf(h.w, kLog2e, -ml));
            if (CACHE >= 1) Es[i] = e;
            const float se = (e.x + e.y) + (e.z + e.w);
            r7[0] += se;
            r7[1] += fmaf((float)x, se, fmaf(3.f, e.w, fmaf(2.f, e.z, e.y)));
            r7[2] = fmaf((float)y, se, r7[2]);
            if (heavy) {
                float4 s;
                s.x = sigmoid_fast(h.x); s.y = sigmoid_fast(h.y); s.z = sigmoid_fast(h.z); s.w = sigmoid_fast(h.w);
                if (CACHE == 2) Ss[i] = s;
                r7[3] += (s.x + s.y) + (s.z + s.w);
                r7[4] += (fmaxf(h.x, 0.f) + fmaxf(h.y, 0.f)) + (fmaxf(h.z, 0.f) + fmaxf(h.w, 0.f));
                const float4 t = target4(i, x, y);
                const float d0 = h.x - t.x, d1 = h.y - t.y, d2 = h.z - t.z, d3 = h.w - t.w;
                r7[5] += fmaf(d0, d0, d1 * d1) + fmaf(d2, d2, d3 * d3);
            }
        }
    }
    block_sum<7>(r7, scratch);
    const float Z = r7[0], iZ = 1.f / Z;
    const float cx = r7[1] * iZ, cy = r7[2] * iZ;
    const float Ssum = r7[3], Rp = r7[4] + kEps;

    // ---- fused decode (warp 0) ------------------------------------------------------------
    if (decode && tid < 32) {
        float dx_ = cx, dy_ = cy; int px, py;
        refine_and_correct<float>(A.hm + (size_t)tile * n, nullptr, A.off ? A.off + (size_t)tile * 2 * n : nullptr,
                           A.alpha_param, A.fusion_weight, H, W, A.radius, A.dflags, dx_, dy_, px, py);
        if (tid == 0) { A.coords[2 * tile] = dx_; A.coords[2 * tile + 1] = dy_; A.scores[tile] = m; }
    }

    float* gh = grads ? A.grad_hm + (size_t)tile * n : nullptr;
    float* gv = (grads && A.grad_var) ? A.grad_var + (size_t)tile * n : nullptr;
    float* go = grads ? A.grad_off + (size_t)tile * 2 * n : nullptr;

    if (!heavy) {
        // weight 0: every term of this tile carries a factor w -> zero loss, zero gradient
        if (tid == 0 && !backward_only) {
            float4* p = reinterpret_cast<float4*>(A.partial + (size_t)tile * 8);
            p[0] = make_float4(0.f, 0.f, 0.f, 0.f); p[1] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (grads) {
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int it = 0; it < niter; ++it) {
                const int i = it * TPB + tid;
                if (NITER > 0 || i < n4) {
                    stg_stream(reinterpret_cast<float4*>(gh) + i, z);
                    if (gv) stg_stream(reinterpret_cast<float4*>(gv) + i, z);
                    stg_stream(reinterpret_cast<float4*>(go) + i, z);
                    stg_stream(reinterpret_cast<float4*>(go) + n4 + i, z);
                }
            }
        }
        return;
    }

    // thread 0 starts the 8 offset taps now; they are consumed after pass C
    Taps tp;
    float ov[2][4];
    if (tid == 0) {
        tp = taps_setup(cx, cy, H, W);
        const float* o = A.off + (size_t)tile * 2 * n;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            ov[c][0] = __ldg(o + c * n + tp.i00); ov[c][1] = __ldg(o + c * n + tp.i01) * tp.okx;
            ov[c][2] = __ldg(o + c * n + tp.i10) * tp.oky; ov[c][3] = __ldg(o + c * n + tp.i11) * (tp.okx * tp.oky);
        }
    }

    // ---- pass C: entropy sums and spatial variance around (cx, cy) --------------------------
    float r5[5] = {0.f, 0.f, 0.f, 0.f, 0.f};       // A1 = sum p lg2(u), A2 = sum p^2/u, sum r e, sum r dx, sum r dy
    for (int it = 0; it < niter; ++it) {
        const int i = it * TPB + tid;
        if (NITER > 0 || i < n4) {
            const float4 h = Hs[i];
            const int y = i / w4, x = (i - y * w4) << 2;
            float ev[4];
            if (CACHE >= 1) { const float4 e = Es[i]; ev[0] = e.x; ev[1] = e.y; ev[2] = e.z; ev[3] = e.w; }
            else { ev[0] = ex2(fmaf(h.x, kLog2e, -ml)); ev[1] = ex2(fmaf(h.y, kLog2e, -ml));
                   ev[2] = ex2(fmaf(h.z, kLog2e, -ml)); ev[3] = ex2(fmaf(h.w, kLog2e, -ml)); }
            const float hv[4] = {h.x, h.y, h.z, h.w};
            const float dy = (float)y - cy, dy2 = dy * dy;
            float av[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float p = ev[j] * iZ, u = p + kEps;
                const float l = lg2(u), rc = rcp(u);
                r5[0] = fmaf(p, l, r5[0]);
                const float prc = p * rc;
                r5[1] = fmaf(p, prc, r5[1]);
                av[j] = fmaf(-kLn2, l, -prc);
                const float dx = (float)(x + j) - cx;
                const float r = fmaxf(hv[j], 0.f);
                r5[2] = fmaf(r, fmaf(dx, dx, dy2), r5[2]);
                r5[3] = fmaf(r, dx, r5[3]);
                r5[4] = fmaf(r, dy, r5[4]);
            }
            if (CACHE == 2) As[i] = make_float4(av[0], av[1], av[2], av[3]);
        }
    }
    block_sum<5>(r5, scratch);

    // ---- per-tile scalars (thread 0) -------------------------------------------------------
    const float D = (float)A.sums[0] + kEps;
    const float D5 = (float)A.sums[1] + kEps;
    if (tid == 0) {
        const float Da = P.use_target_weight ? D : (float)(P.B * P.K);
        const float ka = wa / Da, kb = w / D;
        const float gx = __ldg(A.gt + 2 * tile) * ((float)W / P.in_w);
        const float gy = __ldg(A.gt + 2 * tile + 1) * ((float)H / P.in_h);
        // offset term
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
        // variance term
        const float v = r5[2] / Rp;
        const float s = sqrtf(v + kEps);
        const float mV = A.var ? r7[6] / (float)n : P.sigma;
        const float var_t = (s - P.sigma) * (s - P.sigma) + (A.var ? (mV - P.sigma) * (mV - P.sigma) : 0.f);
        // shape term
        const float E = -kLn2 * r5[0];
        const float pa = E - r5[1];
        const float shape_t = (E - P.e_star) * (E - P.e_star);

        if (!backward_only) {
            float* p = A.partial + (size_t)tile * 8;
            p[0] = wa * (r7[5] / (float)n); p[1] = wa * off_t; p[2] = wa * peak_t;
            p[3] = w * var_t; p[5] = w * shape_t;      // p[4] (limb overlap) is written after the partner loop
        }
        TileCoef c;
        c.c1 = lam[0] * ka * 2.f / (float)n;
        const float a4 = lam[3] * kb * (s - P.sigma) / s;
        c.c4 = a4 / Rp;
        c.v = v;
        c.c6 = lam[5] * kb * 2.f * (E - P.e_star);
        c.pa = pa;
        const float dv_dcx = -2.f * r5[3] / Rp, dv_dcy = -2.f * r5[4] / Rp;
        c.fx = lam[2] * ka * 2.f * (cx - gx) + lam[1] * ka * 0.5f * (sl1p[0] * (dsdx[0] + 1.f) + sl1p[1] * dsdx[1]) + a4 * dv_dcx;
        c.fy = lam[2] * ka * 2.f * (cy - gy) + lam[1] * ka * 0.5f * (sl1p[0] * dsdy[0] + sl1p[1] * (dsdy[1] + 1.f)) + a4 * dv_dcy;
        c.gv = lam[3] * kb * 2.f * (mV - P.sigma) / (float)n;
        c.go[0] = lam[1] * ka * 0.5f * sl1p[0];
        c.go[1] = lam[1] * ka * 0.5f * sl1p[1];
        coef_s = c;
    }

    // ---- limb partners: overlap mass, then the per-pixel tie pattern into Gs -----------------
    float pair_loss = 0.f, cst = 0.f;
    bool g_live = false;
    const int np = P.n_partner[k];
    for (int pi = 0; pi < np; ++pi) {
        const int j = P.partner[k][pi];
        const float wj = __ldg(A.weff + b * P.K + j);
        if (w == 0.f || wj == 0.f) continue;                // CTA-uniform
        const float4* hj4 = reinterpret_cast<const float4*>(A.hm) + ((size_t)b * P.K + j) * n4;
        constexpr int R = NITER > 0 ? NITER : 1;
        float4 hj[R];
        float r2[2] = {0.f, 0.f};                            // S_j, M_kj
        for (int it = 0; it < niter; ++it) {
            const int i = it * TPB + tid;
            if (NITER > 0 || i < n4) {
                const float4 q = ldg_keep(hj4 + i);
                if (NITER > 0) hj[it] = q;
                const float4 h = Hs[i];
                float4 sk;
                if (CACHE == 2) sk = Ss[i];
                else { sk.x = sigmoid_fast(h.x); sk.y = sigmoid_fast(h.y); sk.z = sigmoid_fast(h.z); sk.w = sigmoid_fast(h.w); }
                const float s0 = sigmoid_fast(q.x), s1 = sigmoid_fast(q.y), s2 = sigmoid_fast(q.z), s3 = sigmoid_fast(q.w);
                r2[0] += (s0 + s1) + (s2 + s3);
                // min(sigma(a), sigma(b)) = sigma(min(a, b)): pick by comparing the logits
                r2[1] += ((q.x < h.x ? s0 : sk.x) + (q.y < h.y ? s1 : sk.y)) + ((q.z < h.z ? s2 : sk.z) + (q.w < h.w ? s3 : sk.w));
            }
        }
        block_sum<2>(r2, scratch);
        const float Sj = r2[0], M = r2[1];
        const float mm = fminf(Ssum, Sj) + kEps;
        const float rho = M / mm;
        if ((P.owner[k] >> pi) & 1) pair_loss += w * wj * fmaxf(rho - 0.5f, 0.f);
        if (grads && rho > 0.5f) {
            const float cj = lam[4] * w * wj / D5 / mm;
            cst += cj * rho * tie_rule(Ssum, Sj);
            for (int it = 0; it < niter; ++it) {
                const int i = it * TPB + tid;
                if (NITER > 0 || i < n4) {
                    const float4 q = NITER > 0 ? hj[it] : ldg_keep(hj4 + i);
                    const float4 h = Hs[i];
                    float4 g = g_live ? Gs[i] : make_float4(0.f, 0.f, 0.f, 0.f);
                    g.x = fmaf(cj, tie_rule(h.x, q.x), g.x); g.y = fmaf(cj, tie_rule(h.y, q.y), g.y);
                    g.z = fmaf(cj, tie_rule(h.z, q.z), g.z); g.w = fmaf(cj, tie_rule(h.w, q.w), g.w);
                    Gs[i] = g;
                }
            }
            g_live = true;
        }
    }
    if (tid == 0 && !backward_only) A.partial[(size_t)tile * 8 + 4] = pair_loss;
    __syncthreads();            // coef_s visible (Gs is only read back by its own writer)
    if (!grads) return;

    // ---- pass D: gradient ----------------------------------------------------------------------
    const TileCoef c = coef_s;
    const float4 gv4 = make_float4(c.gv, c.gv, c.gv, c.gv);
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int it = 0; it < niter; ++it) {
        const int i = it * TPB + tid;
        if (NITER > 0 || i < n4) {
            const float4 h = Hs[i];
            const int y = i / w4, x = (i - y * w4) << 2;
            const float hv[4] = {h.x, h.y, h.z, h.w};
            float ev[4], av[4], sv[4], gl[4] = {0.f, 0.f, 0.f, 0.f};
            if (CACHE >= 1) { const float4 e = Es[i]; ev[0] = e.x; ev[1] = e.y; ev[2] = e.z; ev[3] = e.w; }
            else { for (int j = 0; j < 4; ++j) ev[j] = ex2(fmaf(hv[j], kLog2e, -ml)); }
            if (CACHE == 2) { const float4 a = As[i]; av[0] = a.x; av[1] = a.y; av[2] = a.z; av[3] = a.w; }
            if (g_live) {
                const float4 g = Gs[i]; gl[0] = g.x; gl[1] = g.y; gl[2] = g.z; gl[3] = g.w;
                if (CACHE == 2) { const float4 s = Ss[i]; sv[0] = s.x; sv[1] = s.y; sv[2] = s.z; sv[3] = s.w; }
                else { for (int j = 0; j < 4; ++j) sv[j] = sigmoid_fast(hv[j]); }
            }
            const float4 t = target4(i, x, y);
            const float tv[4] = {t.x, t.y, t.z, t.w};
            const float dy = (float)y - cy, dy2 = dy * dy, fyd = dy * c.fy;
            float out[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float p = ev[j] * iZ;
                float a;
                if (CACHE == 2) a = av[j];
                else { const float u = p + kEps; a = fmaf(-kLn2, lg2(u), -p * rcp(u)); }
                const float dx = (float)(x + j) - cx;
                float g = c.c1 * (hv[j] - tv[j]);
                g = fmaf(p, fmaf(c.c6, a - c.pa, fmaf(dx, c.fx, fyd)), g);
                if (hv[j] > 0.f) g = fmaf(c.c4, fmaf(dx, dx, dy2) - c.v, g);
                if (g_live) g = fmaf((gl[j] - cst) * sv[j], 1.f - sv[j], g);
                out[j] = g;
            }
            stg_stream(reinterpret_cast<float4*>(gh) + i, make_float4(out[0], out[1], out[2], out[3]));
            if (gv) stg_stream(reinterpret_cast<float4*>(gv) + i, gv4);
            stg_stream(reinterpret_cast<float4*>(go) + i, z4);
            stg_stream(reinterpret_cast<float4*>(go) + n4 + i, z4);
        }
    }
    // the offset gradient is zero except on the (up to) four taps of each channel
    __syncthreads();
    if (tid == 0) {
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
            float* o = go + ch * n;
            o[tp.i00] = c.go[ch] * tp.w00;
            if (tp.okx != 0.f) o[tp.i01] = c.go[ch] * tp.w01;
            if (tp.oky != 0.f) o[tp.i10] = c.go[ch] * tp.w10;
            if (tp.okx != 0.f && tp.oky != 0.f) o[tp.i11] = c.go[ch] * tp.w11;
        }
    }
}

// ---- second stage: fixed-order sum of the per-tile numerators -----------------------------
// Up to kFinBlocks CTAs each sum a contiguous slice of the tiles in double and publish six
// partials; the CTA that draws the last ticket adds the partials in slice order and writes the
// seven losses.  The order of every addition is fixed by the launch shape: bit-reproducible.
// The tail of a decoding step whose step kernel (step_pipe.cu) left it here: CTAs beyond the `fin_blocks` summing ones,
// one warp per tile.  The soft-argmax (cx, cy) is in d_coords, the two offset-gradient factors lambda2 ka/2 * SmoothL1'
// in words 6 and 7 of the tile's numerator row.  Same operations in the same order as the in-kernel versions
// (step_tile.cu, loss_tile.cu): the taps of the bilinear sample around (cx, cy) into the zero-filled d_grad_off
// (fusion_head.py:353-359, 686-700), then local refinement + offset correction (fusion_head.py:309-365).
__device__ __forceinline__ void st1(float* p, float v) { *p = v; }
__device__ __forceinline__ void st1(__half* p, float v) { *p = __float2half_rn(v); }
struct TailArgs {
    const float* hm; const float* off; float* grad_off; float* coords;
    const float* alpha_param; const float* fusion_weight;
    int radius; unsigned dflags; int fin_blocks; int enabled;
    int half_io;          // hm, off and grad_off are float16 (the pointers are then __half*)
};

// tiles per warp (every load of a warp's tiles is requested before the first is consumed).  Measured on the default
// workload: 1 tile per warp (6 CTAs per SM, 2.4 waves) 0.2541 ms per step, 2 tiles 0.2568, 3 tiles 0.2571, 4 tiles (one
// wave) 0.2586 — the tail is not bound by its waves; one tile per warp it stays
#ifndef FIN_TPW
#define FIN_TPW 1
#endif
constexpr int kTailTPW = FIN_TPW;
template <typename E>
__device__ __forceinline__ void step_tail(const LossParams& P, const TailArgs& T, const float* __restrict__ partial) {
    const E* const t_hm = reinterpret_cast<const E*>(T.hm);
    const E* const t_off = reinterpret_cast<const E*>(T.off);
    E* const t_goff = reinterpret_cast<E*>(T.grad_off);
    const int H = P.H, W = P.W, N = H * W, tiles = P.B * P.K;
    const int lane = threadIdx.x & 31;
    const int tile0 = ((blockIdx.x - T.fin_blocks) * 8 + (threadIdx.x >> 5)) * kTailTPW;
    if (tile0 >= tiles) return;
    const int wside = 2 * T.radius + 1;
    const bool win_small = (T.dflags & GBCODEC_DECODE_REFINE) && wside * wside <= 32;
    const bool want_off = (T.dflags & GBCODEC_DECODE_APPLY_OFFSET) != 0;
    const int wdx = lane % wside - T.radius, wdy = lane / wside - T.radius;
    const bool in_win = lane < wside * wside;
    float cx[kTailTPW], cy[kTailTPW], g0[kTailTPW], g1[kTailTPW], vpx[kTailTPW], pre[kTailTPW];
#pragma unroll
    for (int u = 0; u < kTailTPW; ++u) {
        const int tile = min(tile0 + u, tiles - 1);
        cx[u] = __ldcg(T.coords + 2 * tile); cy[u] = __ldcg(T.coords + 2 * tile + 1);
        g0[u] = 0.f; g1[u] = 0.f;
        if (T.grad_off && lane == 0) { g0[u] = __ldcg(partial + (size_t)tile * 8 + 6); g1[u] = __ldcg(partial + (size_t)tile * 8 + 7); }
    }
    // the decode's loads (one window pixel per lane; the 4x4x2 block of offset taps around floor(cx, cy), which covers the
    // refined coordinate unless it moves by more than a pixel): one round trip for all of them
#pragma unroll
    for (int u = 0; u < kTailTPW; ++u) {
        const int tile = min(tile0 + u, tiles - 1);
        const int px = (int)fminf(fmaxf(rintf(cx[u]), 0.f), (float)(W - 1));
        const int py = (int)fminf(fmaxf(rintf(cy[u]), 0.f), (float)(H - 1));
        const int wx = px + wdx, wy = py + wdy;
        const bool okw = win_small && in_win && wx >= 0 && wx < W && wy >= 0 && wy < H;
        vpx[u] = okw ? ld1(t_hm + (size_t)tile * N + wy * W + wx) : -INFINITY;
        pre[u] = 0.f;
        if (win_small && want_off) {
            const int pbx = (int)floorf(fminf(fmaxf(cx[u], 0.f), (float)(W - 1))) - 1;
            const int pby = (int)floorf(fminf(fmaxf(cy[u], 0.f), (float)(H - 1))) - 1;
            const int tq = lane & 15;
            const int qx = min(max(pbx + (tq & 3), 0), W - 1), qy = min(max(pby + (tq >> 2), 0), H - 1);
            pre[u] = ld1(t_off + (size_t)tile * 2 * N + (lane >> 4) * N + qy * W + qx);
        }
    }
    float a_blend = 1.f, fw_dec = 0.f;
    if (win_small) {
        a_blend = sigmoid_acc(__ldg(T.alpha_param));
        if (want_off) {
            fw_dec = __ldg(T.fusion_weight);
            if (T.dflags & GBCODEC_DECODE_FUSION_WEIGHT_RAW) fw_dec = sigmoid_acc(fw_dec);
        }
    }
#pragma unroll
    for (int u = 0; u < kTailTPW; ++u) {
        const int tile = tile0 + u;
        if (tile >= tiles) break;
        const E* off_tile = t_off + (size_t)tile * 2 * N;
        if (lane == 0 && (g0[u] != 0.f || g1[u] != 0.f)) {
            const Taps tp = taps_setup(cx[u], cy[u], H, W);
            E* go = t_goff + (size_t)tile * 2 * N;
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                E* o = go + ch * N;
                const float gc = ch == 0 ? g0[u] : g1[u];
                st1(o + tp.i00, gc * tp.w00);
                if (tp.okx != 0.f) st1(o + tp.i01, gc * tp.w01);
                if (tp.oky != 0.f) st1(o + tp.i10, gc * tp.w10);
                if (tp.okx != 0.f && tp.oky != 0.f) st1(o + tp.i11, gc * tp.w11);
            }
        }
        float dx_ = cx[u], dy_ = cy[u];
        if (win_small) {
            const int px = (int)fminf(fmaxf(rintf(cx[u]), 0.f), (float)(W - 1));
            const int py = (int)fminf(fmaxf(rintf(cy[u]), 0.f), (float)(H - 1));
            const int x = px + wdx, y = py + wdy;
            const bool ok = in_win && x >= 0 && x < W && y >= 0 && y < H;
            const float vmax = warp_max(vpx[u]);
            const float e = ok ? expf(vpx[u] - vmax) : 0.f;
            float sw4[4] = {e, e * (float)x, e * (float)y, 0.f};
            warp_scatter_sum<4>(sw4, lane);
            const float se = __shfl_sync(0xffffffffu, sw4[0], 0), sx = __shfl_sync(0xffffffffu, sw4[0], 8), sy = __shfl_sync(0xffffffffu, sw4[0], 16);
            dx_ = a_blend * cx[u] + (1.f - a_blend) * (sx / se);
            dy_ = a_blend * cy[u] + (1.f - a_blend) * (sy / se);
            if (want_off) {
                const int pbx = (int)floorf(fminf(fmaxf(cx[u], 0.f), (float)(W - 1))) - 1;
                const int pby = (int)floorf(fminf(fmaxf(cy[u], 0.f), (float)(H - 1))) - 1;
                const Bilinear bl = bilinear_setup(dx_, dy_, H, W);
                float ox, oy;
                if (bl.x0 >= pbx && bl.x1 <= pbx + 3 && bl.y0 >= pby && bl.y1 <= pby + 3) {          // warp-uniform
                    // the same four values times the same factors in bilinear_read's order
                    const int i00 = (bl.y0 - pby) * 4 + (bl.x0 - pbx), i01 = (bl.y0 - pby) * 4 + (bl.x1 - pbx);
                    const int i10 = (bl.y1 - pby) * 4 + (bl.x0 - pbx), i11 = (bl.y1 - pby) * 4 + (bl.x1 - pbx);
                    float tq[2][4];
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        tq[c][0] = __shfl_sync(0xffffffffu, pre[u], c * 16 + i00);
                        tq[c][1] = __shfl_sync(0xffffffffu, pre[u], c * 16 + i01) * bl.okx;
                        tq[c][2] = __shfl_sync(0xffffffffu, pre[u], c * 16 + i10) * bl.oky;
                        tq[c][3] = __shfl_sync(0xffffffffu, pre[u], c * 16 + i11) * (bl.okx * bl.oky);
                    }
                    ox = bl.w00 * tq[0][0] + bl.w01 * tq[0][1] + bl.w10 * tq[0][2] + bl.w11 * tq[0][3];
                    oy = bl.w00 * tq[1][0] + bl.w01 * tq[1][1] + bl.w10 * tq[1][2] + bl.w11 * tq[1][3];
                } else {
                    ox = bilinear_read(off_tile, bl, W);
                    oy = bilinear_read(off_tile + N, bl, W);
                }
                dx_ += fw_dec * ox;
                dy_ += fw_dec * oy;
            }
        } else {
            int qx_, qy_;
            refine_and_correct<E>(t_hm + (size_t)tile * N, nullptr, off_tile, T.alpha_param, T.fusion_weight, H, W, T.radius, T.dflags, dx_, dy_, qx_, qy_);
        }
        if (lane == 0) { T.coords[2 * tile] = dx_; T.coords[2 * tile + 1] = dy_; }
    }
}

#ifndef FIN_MINB
#define FIN_MINB 6
#endif
__global__ void __launch_bounds__(256, FIN_MINB)
finalize_kernel(const __grid_constant__ LossParams P, const float* __restrict__ partial, const double* __restrict__ sums,
                double* __restrict__ bpart, unsigned* __restrict__ ticket, float* __restrict__ losses7,
                const __grid_constant__ PeerView peer, const __grid_constant__ TailArgs tail) {
    __shared__ double red[6][8];
    __shared__ bool last;
    pdl_wait();                                        // launched while the tile kernel's last wave is still running
    const int nblk = tail.enabled ? tail.fin_blocks : (int)gridDim.x;      // the summing CTAs
    if ((int)blockIdx.x >= nblk) {
        if (tail.half_io) step_tail<__half>(P, tail, partial); else step_tail<float>(P, tail, partial);
        return;
    }
    const int tiles = P.B * P.K;
    const int per = (tiles + nblk - 1) / nblk;
    const int lo = blockIdx.x * per, hi = min(tiles, lo + per);
    double acc[6] = {0, 0, 0, 0, 0, 0};
    for (int t = lo + threadIdx.x; t < hi; t += blockDim.x) {
        const float4 a = __ldcg(reinterpret_cast<const float4*>(partial + (size_t)t * 8));
        const float4 c = __ldcg(reinterpret_cast<const float4*>(partial + (size_t)t * 8 + 4));
        acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w; acc[4] += c.x; acc[5] += c.y;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < 6; ++q) {
        for (int o = 16; o > 0; o >>= 1) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], o);
        if (lane == 0) red[q][warp] = acc[q];
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        double s = 0.0;
        for (int wp = 0; wp < 8; ++wp) s += red[threadIdx.x][wp];
        bpart[blockIdx.x * 6 + threadIdx.x] = s;
        __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(ticket, 1u) == (unsigned)nblk - 1u;
    __syncthreads();
    if (!last) return;
    __threadfence();
    __shared__ float term[6];
    if (threadIdx.x < 6) {
        const int q = threadIdx.x;
        double s = 0.0;
        for (int g = 0; g < nblk; ++g) s += __ldcg(bpart + g * 6 + q);
        const float D = (float)sums[0] + kEps, D5 = (float)sums[1] + kEps;
        const float Da = P.use_target_weight ? D : (float)tiles;
        const float den = q < 3 ? Da : (q == 4 ? D5 : D);
        term[q] = P.lam[q] * ((float)s / den);
        losses7[q] = term[q];
    }
    __syncthreads();
    if (peer.world <= 1) {
        if (threadIdx.x == 0) {
            float total = 0.f;
            for (int q = 0; q < 6; ++q) total += term[q];
            losses7[6] = total;
            *ticket = 0u;
        }
        return;
    }
    // ---- batch-sharded job: every rank's terms are already divided by the GLOBAL normalisers, so the global
    // losses are their sums over the ranks.  Same mailbox protocol as denoms_kernel; fixed rank order.
    __shared__ float gl[GBCODEC_MAX_PEERS][8];
    const int pq = (int)(peer.seq & 3ull), r = threadIdx.x;
    if (r < peer.world) {
        PeerMail* dst = peer.mail[r];
        for (int q = 0; q < 6; ++q) dst->loss[pq][peer.rank][q] = term[q];
        __threadfence_system();
        st_release_sys(&dst->loss_seq[pq][peer.rank], peer.seq);
    }
    if (peer.defer) {
        // deferred mode: the terms are on their way to every peer; nobody waits inside the step.  d_losses7 holds THIS
        // rank's share (already divided by the global normalisers); gbcodec_peer_collect_losses_f32 adds the shares up
        // whenever the caller wants the numbers (they are logging-only: fusion_head.py:795-806 feeds nothing back).
        if (threadIdx.x == 0) {
            float total = 0.f;
            for (int q = 0; q < 6; ++q) total += term[q];
            losses7[6] = total;
            *ticket = 0u;
        }
        return;
    }
    if (r < peer.world) {
        PeerMail* mine = peer.mail[peer.rank];
        const bool ok = wait_seq(&mine->loss_seq[pq][r], peer.seq, peer, &mine->timeouts);
        for (int q = 0; q < 6; ++q) gl[r][q] = ok ? mine->loss[pq][r][q] : NAN;
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        float v = 0.f;
        for (int i = 0; i < peer.world; ++i) v += gl[i][threadIdx.x];
        term[threadIdx.x] = v;
        losses7[threadIdx.x] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float total = 0.f;
        for (int q = 0; q < 6; ++q) total += term[q];
        losses7[6] = total;
        *ticket = 0u;
    }
}

// Sum of every rank's loss terms of step `seq`, in rank order (every rank gets the same bits).  One warp.
__global__ void collect_losses_kernel(const __grid_constant__ PeerView peer, unsigned long long seq, float* __restrict__ losses7) {
    __shared__ float gl[GBCODEC_MAX_PEERS][8];
    const int pq = (int)(seq & 3ull), r = threadIdx.x;
    if (r < peer.world) {
        PeerMail* mine = peer.mail[peer.rank];
        const bool ok = wait_seq(&mine->loss_seq[pq][r], seq, peer, &mine->timeouts);
        for (int q = 0; q < 6; ++q) gl[r][q] = ok ? mine->loss[pq][r][q] : NAN;
    }
    __syncthreads();
    float v = 0.f;
    if (threadIdx.x < 6) {
        for (int i = 0; i < peer.world; ++i) v += gl[i][threadIdx.x];
        losses7[threadIdx.x] = v;
    }
    float total = 0.f;
    for (int q = 0; q < 6; ++q) total += __shfl_sync(0xffffffffu, v, q);
    if (threadIdx.x == 0) losses7[6] = total;
}

// ---- backward for an arbitrary upstream gradient --------------------------------------------
// plan: 0 = stored gradients already right, 1 = scale them by ratio, 2 = recompute with lam_eff.
// `held` (6 floats, caller-owned for the lifetime of one forward's stored gradients) records the upstream factor each
// term's share of the stored gradients carries NOW: a backward may rescale or recompute them in place, so the next
// backward through the same graph (retain_graph, per-term losses followed by total_loss) compares with what is there,
// not with what the forward assumed.  held_valid == 0: first backward, every term carries *assumed (1 if NULL).
__global__ void plan_kernel(const __grid_constant__ LossParams P, const float* __restrict__ g7,
                            const float* __restrict__ assumed, float* __restrict__ held, int held_valid,
                            int* __restrict__ plan, float* __restrict__ lam_eff) {
    if (threadIdx.x != 0) return;
    const float a = assumed ? *assumed : 1.f;
    float r[6], h[6];
    bool r_uniform = true, h_uniform = true, same = true;
    for (int q = 0; q < 6; ++q) {
        r[q] = g7[6] + g7[q];
        h[q] = (held && held_valid) ? held[q] : a;
        lam_eff[q] = P.lam[q] * r[q];
        if (r[q] != r[0]) r_uniform = false;
        if (h[q] != h[0]) h_uniform = false;
        if (r[q] != h[q]) same = false;
    }
    const float ratio = r[0] / h[0];
    if (same) plan[0] = 0;
    else if (r_uniform && h_uniform && h[0] != 0.f && isfinite(ratio)) { plan[0] = 1; lam_eff[6] = ratio; }
    else plan[0] = 2;
    if (held) for (int q = 0; q < 6; ++q) held[q] = r[q];
}
// plan 1: the three gradient tensors times one factor, one launch (n4a + n4b + n4c float4s)
__global__ void __launch_bounds__(256)
rescale_kernel(const int* __restrict__ plan, const float* __restrict__ lam_eff, float4* __restrict__ ga, size_t n4a,
               float4* __restrict__ gb, size_t n4b, float4* __restrict__ gc, size_t n4c) {
    if (*plan != 1) return;
    const float r = lam_eff[6];
    const size_t n = n4a + n4b + n4c;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float4* g = i < n4a ? ga + i : (i < n4a + n4b ? gb + (i - n4a) : gc + (i - n4a - n4b));
        float4 v = *g;
        v.x *= r; v.y *= r; v.z *= r; v.w *= r;
        *g = v;
    }
}

// ---- host side ---------------------------------------------------------------------------------
static int make_params(const gbcodec_loss_desc* d, LossParams* P) {
    if (d->B <= 0 || d->K <= 0 || d->H <= 0 || d->W <= 0) return fail(GBCODEC_ERR_BAD_SHAPE, "loss: B,K,H,W must be positive");
    if (d->K > GBCODEC_MAX_K) return fail(GBCODEC_ERR_BAD_SHAPE, "loss: K=%d exceeds %d", d->K, GBCODEC_MAX_K);
    if (d->W % 4) return fail(GBCODEC_ERR_BAD_SHAPE, "loss: W=%d is not a multiple of 4", d->W);
    if ((long long)d->H * d->W > GBCODEC_MAX_TILE) return fail(GBCODEC_ERR_BAD_SHAPE, "loss: tile %dx%d too large", d->H, d->W);
    if (d->n_pairs < 0 || d->n_pairs > GBCODEC_MAX_PAIRS) return fail(GBCODEC_ERR_BAD_ARGUMENT, "loss: n_pairs=%d", d->n_pairs);
    if (!(d->target_sigma > 0.0) || !(d->encode_sigma > 0.0)) return fail(GBCODEC_ERR_BAD_ARGUMENT, "loss: sigma must be positive");
    memset(P, 0, sizeof(*P));
    P->B = d->B; P->K = d->K; P->H = d->H; P->W = d->W; P->in_w = d->in_w; P->in_h = d->in_h;
    for (int q = 0; q < 6; ++q) P->lam[q] = d->lambdas[q];
    P->sigma = (float)d->target_sigma;
    P->e_star = (float)log(2.0 * M_PI * M_E * d->target_sigma * d->target_sigma);
    P->inv_n = 1.0f / (float)(d->H * d->W);
    P->sx = (float)d->W / d->in_w; P->sy = (float)d->H / d->in_h;
    P->use_target_weight = d->use_target_weight;
    P->ec = make_encode_const(d->encode_sigma);
    int np = 0;
    for (int p = 0; p < d->n_pairs; ++p) {
        const int i = d->pairs[p][0], j = d->pairs[p][1];
        if (i < 0 || j < 0) return fail(GBCODEC_ERR_BAD_ARGUMENT, "loss: negative channel in pair %d", p);
        if (i >= d->K || j >= d->K) continue;                   // fusion_head.py:504-505
        if (i == j) return fail(GBCODEC_ERR_BAD_ARGUMENT, "loss: pair %d joins channel %d with itself", p, i);
        if (P->n_partner[i] >= GBCODEC_MAX_PARTNERS || P->n_partner[j] >= GBCODEC_MAX_PARTNERS)
            return fail(GBCODEC_ERR_BAD_ARGUMENT, "loss: a channel takes part in more than %d pairs", GBCODEC_MAX_PARTNERS);
        P->owner[i] |= (uint8_t)(1u << P->n_partner[i]);
        P->partner[i][P->n_partner[i]++] = (int8_t)j;
        P->partner[j][P->n_partner[j]++] = (int8_t)i;
        P->pair_i[np] = (int16_t)i; P->pair_j[np] = (int16_t)j; ++np;
    }
    P->n_pairs = np;
    return GBCODEC_OK;
}

static thread_local cudaEvent_t g_prof_start = nullptr, g_prof_stop = nullptr;
void set_profile_events(cudaEvent_t a, cudaEvent_t b) { g_prof_start = a; g_prof_stop = b; }

template <int TPB, int NITER, int CACHE>
static int launch_loss_t(const LossParams& P, const LossArgs& A, cudaStream_t s) {
    const int n = P.H * P.W;
    const size_t smem = (size_t)n * 4 * (2 + (CACHE >= 1 ? 1 : 0) + (CACHE == 2 ? 2 : 0))
                      + (size_t)((P.ec.lut_size + 3) & ~3) * 4 + (8 * 32 + 8) * 4;
    auto kern = loss_kernel<TPB, NITER, CACHE>;
    if (smem > 227 * 1024) return fail(GBCODEC_ERR_BAD_SHAPE, "loss: tile %dx%d needs %zu bytes of shared memory", P.H, P.W, smem);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(GBCODEC_ERR_CUDA, "cudaFuncSetAttribute(loss_kernel): %s", cudaGetErrorString(e));
    if (g_prof_start && !A.plan) cudaEventRecord(g_prof_start, s);
    note_launch(), kern<<<P.B * P.K, TPB, smem, s>>>(P, A);
    if (g_prof_stop && !A.plan) cudaEventRecord(g_prof_stop, s);
    return check_launch("loss_kernel");
}

// GBCODEC_LOSS_KERNEL=generic forces the shared-memory kernel below for every shape (A/B measurements).
static bool force_generic() {
    static const int v = [] { const char* e = getenv("GBCODEC_LOSS_KERNEL"); return (e && !strcmp(e, "generic")) ? 1 : 0; }();
    return v != 0;
}

static int launch_loss_kernel(const LossParams& P, const LossArgs& A, cudaStream_t s, bool* pipe_used = nullptr) {
    if (!force_generic()) {
        // the persistent step kernels cover float32 maps with the target generated on the fly
        int st = launch_step_pipe(P, A, s, g_prof_start, g_prof_stop);
        if (st != 1) { if (pipe_used) *pipe_used = (st == 0); return st; }
        st = launch_step_tile(P, A, s, g_prof_start, g_prof_stop);
        if (st != 1) return st;
    }
    if (!force_generic() || A.half_io) {
        const int st = launch_loss_tile(P, A, s, g_prof_start, g_prof_stop);
        if (st != 1) return st;
        if (A.half_io) return fail(GBCODEC_ERR_BAD_SHAPE, "loss: float16 maps are supported for 64x48, 64x64, 96x72 and 128x128 tiles (got %dx%d)", P.H, P.W);
    }
    const int n4 = (P.H * P.W) >> 2;
    if (n4 == 256 * 3) return launch_loss_t<256, 3, 2>(P, A, s);           // 64x48
    if (n4 == 576 * 3) return launch_loss_t<576, 3, 1>(P, A, s);           // 96x72
    if (n4 == 1024 * 4) return launch_loss_t<1024, 4, 1>(P, A, s);         // 128x128
    if ((size_t)n4 * 16 * 5 <= 160 * 1024) return launch_loss_t<256, 0, 2>(P, A, s);
    if ((size_t)n4 * 16 * 3 <= 200 * 1024) return launch_loss_t<512, 0, 1>(P, A, s);
    return launch_loss_t<1024, 0, 0>(P, A, s);
}

size_t loss_workspace_bytes(int B, int K) { return ws_bytes(B, K) + 64; }

static int check_common(const gbcodec_loss_desc* d, const float* hm, const float* off, const float* weight,
                        const float* gt, void* ws, size_t ws_size) {
    if (!d || !hm || !off || !weight || !gt) return fail(GBCODEC_ERR_NULL_POINTER, "loss: a required pointer is NULL");
    if (!ws || ws_size < loss_workspace_bytes(d->B, d->K))
        return fail(GBCODEC_ERR_WORKSPACE, "loss: workspace of %zu bytes needed, %zu given", loss_workspace_bytes(d->B, d->K), ws ? ws_size : (size_t)0);
    if (!aligned16(hm) || !aligned16(off) || !aligned16(ws)) return fail(GBCODEC_ERR_UNALIGNED, "loss: tensors must be 16-byte aligned");
    return GBCODEC_OK;
}

static const PeerView kNoPeers = {};

static int prepare_weights(const LossParams& P, const WsLayout& L, const float* weight, const float* gt,
                           int target_given, const float* denoms, cudaStream_t s, const PeerView& peer = kNoPeers,
                           float* global_out = nullptr) {
    const int ipb = 256 / P.K;
    const int grid = (P.B + ipb - 1) / ipb;
    if (denoms) {
        note_launch(), denoms_kernel<<<grid, 256, 0, s>>>(P, weight, gt, target_given, L.weff, L.geom, L.desc, nullptr, L.ticket, kNoPeers, nullptr, denoms, L.sums);
    } else {
        // sums, plan, the finalize ticket and the step kernel's tile counter share the first 32 bytes of the workspace
        cudaError_t e = cudaMemsetAsync(L.sums, 0, 32, s);
        if (e != cudaSuccess) return fail(GBCODEC_ERR_CUDA, "cudaMemsetAsync: %s", cudaGetErrorString(e));
        note_launch(), denoms_kernel<<<grid, 256, 0, s>>>(P, weight, gt, target_given, L.weff, L.geom, L.desc, L.sums, L.ticket, peer, peer.world > 1 ? global_out : nullptr);
    }
    return check_launch("denoms_kernel");
}

int loss_denominators(const gbcodec_loss_desc* d, const float* weight, const float* gt, int target_given,
                      float* out2, void* ws, size_t ws_size, cudaStream_t s) {
    LossParams P;
    int st = make_params(d, &P);
    if (st) return st;
    if (!weight || !out2 || (!target_given && !gt)) return fail(GBCODEC_ERR_NULL_POINTER, "denominators: NULL pointer");
    if (!ws || ws_size < loss_workspace_bytes(d->B, d->K) || !aligned16(ws))
        return fail(GBCODEC_ERR_WORKSPACE, "denominators: workspace of %zu bytes needed", loss_workspace_bytes(d->B, d->K));
    const WsLayout L = ws_carve(ws, P.B, P.K);
    st = prepare_weights(P, L, weight, gt, target_given, nullptr, s);
    if (st) return st;
    note_launch(), sums_to_float_kernel<<<1, 32, 0, s>>>(L.sums, out2);
    return check_launch("sums_to_float_kernel");
}

int fusion_loss(const gbcodec_loss_desc* d, const float* hm, const float* off, const float* var, const float* target,
                const float* weight, const float* gt, const float* denoms, const float* grad_scale,
                float* losses7, float* ghm, float* goff, float* gvar,
                const float* alpha_param, const float* fusion_weight, int radius, unsigned dflags, float* coords, float* scores,
                void* ws, size_t ws_size, cudaStream_t s, void* peer_ctx, float* denoms_out, int half_io,
                const float* var_mean, float* grad_var_mean, int defer_losses) {
    int st = check_common(d, hm, off, weight, gt, ws, ws_size);
    if (var_mean && var) return fail(GBCODEC_ERR_BAD_ARGUMENT, "loss: give the variance maps or their per-tile means, not both");
    if (var_mean && (ghm != nullptr) != (grad_var_mean != nullptr)) return fail(GBCODEC_ERR_NULL_POINTER, "loss: d_grad_var_mean goes with the other gradients");
    if (st) return st;
    PeerView peer = kNoPeers;
    if (peer_ctx) {
        PeerCtx* pc = reinterpret_cast<PeerCtx*>(peer_ctx);
        if (!pc->connected) return fail(GBCODEC_ERR_BAD_ARGUMENT, "sharded step: the peer context is not connected");
        if (defer_losses && !denoms) return fail(GBCODEC_ERR_BAD_ARGUMENT, "sharded step: deferred losses go with prefetched normalisers (d_denoms_global)");
        if (pc->h_failed && *reinterpret_cast<volatile unsigned int*>(pc->h_failed))
            return fail(GBCODEC_ERR_PEER_TIMEOUT, "sharded step: an earlier call of rank %d gave up waiting for a peer (its normalisers / losses were NaN); "
                                                  "the exchange is out of step — tear the job down", pc->view.rank);
    }
    if (!losses7) return fail(GBCODEC_ERR_NULL_POINTER, "loss: d_losses7 is NULL");
    const bool grads = ghm || goff || gvar;
    if (grads && (!ghm || !goff || (!gvar) != (!var))) return fail(GBCODEC_ERR_NULL_POINTER, "loss: give all gradient pointers or none (d_grad_var iff d_var)");
    if ((var && !aligned16(var)) || (target && !aligned16(target)) || (ghm && !aligned16(ghm)) || (goff && !aligned16(goff)) || (gvar && !aligned16(gvar)))
        return fail(GBCODEC_ERR_UNALIGNED, "loss: tensors must be 16-byte aligned");
    if (coords) {
        if (!scores) return fail(GBCODEC_ERR_NULL_POINTER, "step: d_scores is NULL");
        if ((dflags & GBCODEC_DECODE_REFINE) && !alpha_param) return fail(GBCODEC_ERR_NULL_POINTER, "step: d_alpha_param is NULL");
        if ((dflags & GBCODEC_DECODE_APPLY_OFFSET) && !fusion_weight) return fail(GBCODEC_ERR_NULL_POINTER, "step: d_fusion_weight is NULL");
        if (radius < 0 || radius > 8) return fail(GBCODEC_ERR_BAD_ARGUMENT, "step: local_radius=%d", radius);
    }
    LossParams P;
    st = make_params(d, &P);
    if (st) return st;
    if (peer_ctx) {
        // every check has passed: only now does this call take its place in the exchange's sequence (a call that
        // fails validation on one rank must not put that rank one step ahead of the others for the rest of the job)
        PeerCtx* pc = reinterpret_cast<PeerCtx*>(peer_ctx);
        pc->view.seq += 1;
        if (!denoms) pc->view.den_seq += 1;              // the normalisers are exchanged inside this call
        peer = pc->view;
        peer.defer = defer_losses;
    }
    const WsLayout L = ws_carve(ws, P.B, P.K);
    // normalisers given (prefetched with gbcodec_peer_denominators_f32, or all-reduced by the caller): no exchange in front
    st = prepare_weights(P, L, weight, gt, target != nullptr, denoms, s, denoms ? kNoPeers : peer, denoms_out);
    if (st) return st;
    if (denoms_out && peer.world <= 1) {                 // with peers the exchanging CTA has written them already
        note_launch(), sums_to_float_kernel<<<1, 32, 0, s>>>(L.sums, denoms_out);
        st = check_launch("sums_to_float_kernel");
        if (st) return st;
    }
    bool pipe_used = false;
    LossArgs A;
    memset(&A, 0, sizeof(A));
    A.hm = hm; A.off = off; A.var = var; A.target = target; A.weight = weight; A.gt = gt; A.grad_scale = grad_scale;
    A.grad_hm = ghm; A.grad_off = goff; A.grad_var = gvar;
    A.alpha_param = alpha_param; A.fusion_weight = fusion_weight; A.coords = coords; A.scores = scores;
    A.radius = radius; A.dflags = dflags;
    A.sums = L.sums; A.weff = L.weff; A.geom = L.geom; A.partial = L.partial;
    A.desc = L.desc; A.tile_counter = L.tile_counter; A.sm_slots = L.sm_slots;
    A.half_io = half_io;
    A.var_mean = var_mean; A.grad_var_mean = grad_var_mean;
    if (var_mean) {
        // the persistent step kernel (64x48, targets built in the kernel) and the tile kernels take the means
        st = force_generic() ? 1 : launch_step_pipe(P, A, s, g_prof_start, g_prof_stop);
        if (st == 0) pipe_used = true;
        if (st == 1) st = launch_loss_tile(P, A, s, g_prof_start, g_prof_stop);
        if (st == 1) return fail(GBCODEC_ERR_BAD_SHAPE, "loss: per-tile variance means are supported for 64x48, 64x64, 96x72 and 128x128 tiles (got %dx%d)", P.H, P.W);
    } else
    st = launch_loss_kernel(P, A, s, &pipe_used);
    if (st) return st;
    const int tiles = P.B * P.K;
    const int fin_blocks = (tiles + 255) / 256 < kFinBlocks ? (tiles + 255) / 256 : kFinBlocks;
    {
        // a decoding step through step_pipe_kernel leaves its tail (offset-gradient taps, refinement of the decode) to
        // extra CTAs of this launch, one warp per tile
        TailArgs tail;
        memset(&tail, 0, sizeof(tail));
        tail.fin_blocks = fin_blocks;
        if (pipe_used && coords && step_pipe_tail_outside()) {
            tail.enabled = 1;
            tail.hm = hm; tail.off = off; tail.grad_off = goff; tail.coords = coords;
            tail.alpha_param = alpha_param; tail.fusion_weight = fusion_weight; tail.radius = radius; tail.dflags = dflags;
            tail.half_io = half_io;
        }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(fin_blocks + (tail.enabled ? (tiles + 8 * kTailTPW - 1) / (8 * kTailTPW) : 0)); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        const float* partial_c = L.partial; const double* sums_c = L.sums;
        note_launch();
        cudaError_t e = cudaLaunchKernelEx(&cfg, finalize_kernel, P, partial_c, sums_c, L.bpart, L.ticket, losses7, peer, tail);
        if (e != cudaSuccess) return fail(GBCODEC_ERR_CUDA, "cudaLaunchKernelEx(finalize_kernel): %s", cudaGetErrorString(e));
    }
    return check_launch("finalize_kernel");
}

// The normaliser exchange on its own: this rank's raw sums into every peer's mailbox, the global sums out.  A caller
// that knows the next batch's visibility flags and keypoints (train.py: as soon as the batch is loaded) issues this on a
// side stream during the current step; the step itself then starts with the global sums in hand.
int peer_denominators(const gbcodec_loss_desc* d, const float* weight, const float* gt, int target_given,
                      float* out2, void* ws, size_t ws_size, void* peer_ctx, cudaStream_t s) {
    LossParams P;
    int st = make_params(d, &P);
    if (st) return st;
    if (!weight || !out2 || !peer_ctx || (!target_given && !gt)) return fail(GBCODEC_ERR_NULL_POINTER, "peer_denominators: NULL pointer");
    if (!ws || ws_size < loss_workspace_bytes(d->B, d->K) || !aligned16(ws))
        return fail(GBCODEC_ERR_WORKSPACE, "peer_denominators: workspace of %zu bytes needed", loss_workspace_bytes(d->B, d->K));
    PeerCtx* pc = reinterpret_cast<PeerCtx*>(peer_ctx);
    if (!pc->connected) return fail(GBCODEC_ERR_BAD_ARGUMENT, "peer_denominators: the peer context is not connected");
    if (pc->h_failed && *reinterpret_cast<volatile unsigned int*>(pc->h_failed))
        return fail(GBCODEC_ERR_PEER_TIMEOUT, "peer_denominators: an earlier call of rank %d gave up waiting for a peer", pc->view.rank);
    pc->view.den_seq += 1;
    const WsLayout L = ws_carve(ws, P.B, P.K);
    st = prepare_weights(P, L, weight, gt, target_given, nullptr, s, pc->view, out2);
    if (st) return st;
    if (pc->view.world <= 1) {
        note_launch(), sums_to_float_kernel<<<1, 32, 0, s>>>(L.sums, out2);
        return check_launch("sums_to_float_kernel");
    }
    return GBCODEC_OK;
}

int peer_collect_losses(void* peer_ctx, int steps_back, float* losses7, cudaStream_t s) {
    if (!peer_ctx || !losses7) return fail(GBCODEC_ERR_NULL_POINTER, "peer_collect_losses: NULL pointer");
    PeerCtx* pc = reinterpret_cast<PeerCtx*>(peer_ctx);
    if (steps_back < 0 || steps_back > 2 || (unsigned long long)steps_back >= pc->view.seq)
        return fail(GBCODEC_ERR_BAD_ARGUMENT, "peer_collect_losses: steps_back=%d (0..2, and a step that has been made)", steps_back);
    note_launch(), collect_losses_kernel<<<1, 32, 0, s>>>(pc->view, pc->view.seq - (unsigned long long)steps_back, losses7);
    return check_launch("collect_losses_kernel");
}

int fusion_loss_backward(const gbcodec_loss_desc* d, const float* hm, const float* off, const float* var, const float* target,
                         const float* weight, const float* gt, const float* denoms, const float* grad_scale, const float* g7,
                         float* ghm, float* goff, float* gvar, void* ws, size_t ws_size, cudaStream_t s, int half_io,
                         int have_stash, float* held, int held_valid, int ws_from_forward) {
    int st = check_common(d, hm, off, weight, gt, ws, ws_size);
    if (st) return st;
    if (!g7 || !ghm || !goff || (!gvar) != (!var)) return fail(GBCODEC_ERR_NULL_POINTER, "backward: NULL pointer");
    if (held_valid && !held) return fail(GBCODEC_ERR_NULL_POINTER, "backward: held_valid without d_held6");
    LossParams P;
    st = make_params(d, &P);
    if (st) return st;
    const WsLayout L = ws_carve(ws, P.B, P.K);
    // the workspace still holds the weights, patch geometry and sums of the forward if the caller kept it
    // (ws_from_forward); otherwise they are computed again, which is cheap
    if (!ws_from_forward) {
        st = prepare_weights(P, L, weight, gt, target != nullptr, denoms, s);
        if (st) return st;
    }
    note_launch(), plan_kernel<<<1, 32, 0, s>>>(P, g7, grad_scale, held, held_valid, L.plan, L.lam_eff);
    if (!half_io && have_stash) {
        const size_t n4 = (size_t)P.B * P.K * P.H * P.W / 4;
        note_launch(), rescale_kernel<<<148 * 8, 256, 0, s>>>(L.plan, L.lam_eff, reinterpret_cast<float4*>(ghm), n4,
                                                reinterpret_cast<float4*>(goff), 2 * n4, reinterpret_cast<float4*>(gvar), gvar ? n4 : 0);
    }
    LossArgs A;
    memset(&A, 0, sizeof(A));
    A.hm = hm; A.off = off; A.var = var; A.target = target; A.weight = weight; A.gt = gt;
    A.grad_hm = ghm; A.grad_off = goff; A.grad_var = gvar;
    A.sums = L.sums; A.weff = L.weff; A.geom = L.geom; A.partial = L.partial; A.lam_eff = L.lam_eff;
    // float16 maps: a gradient must meet its upstream factor before it is rounded to half.  Either the forward stored
    // gradients for an ASSUMED upstream factor (d_grad_scale) and this call only re-computes them if the actual one
    // differs (plan != 0), or it stored nothing and this call always computes them (per-term weights of plan_kernel)
    A.plan = have_stash ? L.plan : nullptr;
    A.half_io = half_io;
    return launch_loss_kernel(P, A, s);
}

}  // namespace gbc
