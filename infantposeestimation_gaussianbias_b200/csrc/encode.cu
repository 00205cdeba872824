// encode.cu — Gaussian target tiles written straight into HBM.
// Restates COCOPoseDataset._generate_target (datasets/coco_dataset.py:185-250) and the two
// second-generation encoders (data/coco_dataset.py:222-287, data/pose_transforms.py:385-457)
// for a whole batch: one CTA per (image, keypoint) tile, 128-bit streaming stores.
// Roofline: HBM write, 4*H*W bytes per tile (nothing but 12 bytes is read).
#include "common.cuh"

namespace gbc {

__global__ void __launch_bounds__(256)
encode_kernel(const float* __restrict__ kps, const float* __restrict__ vis,
              float* __restrict__ target, float* __restrict__ weight,
              int tiles, int H, int W, float in_w, float in_h, EncodeConst ec) {
    extern __shared__ float lut[];
    __shared__ PatchGeom geom;
    fill_patch_lut(lut, ec);
    const int n4 = (H * W) >> 2;
    const int w4 = W >> 2;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        __syncthreads();   // lut ready / previous geom consumed
        if (threadIdx.x == 0) {
            geom = patch_geometry(kps[2 * tile], kps[2 * tile + 1], vis[tile], H, W, in_w, in_h, ec);
            weight[tile] = geom.weight;
        }
        __syncthreads();
        const PatchGeom g = geom;
        float4* out = reinterpret_cast<float4*>(target) + (size_t)tile * n4;
        for (int i = threadIdx.x; i < n4; i += blockDim.x) {
            const int y = i / w4, x = (i - y * w4) << 2;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (g.active && y >= g.y_from && y < g.y_to && x + 3 >= g.x_from && x < g.x_to) {
                float e[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int xx = x + j;
                    e[j] = (xx >= g.x_from && xx < g.x_to) ? patch_value(lut, g, ec, xx, y) : 0.f;
                }
                v = make_float4(e[0], e[1], e[2], e[3]);
            }
            stg_stream(out + i, v);
        }
    }
}

// Warp-per-tile variant for small tiles (64x48: 12 KB): no block barrier per tile — every lane works the patch geometry
// out for itself (the same instructions a single lane would issue) and the warp streams its tile out 512 bytes per
// store instruction.  The CTA shares only the exp table, filled once.
__global__ void __launch_bounds__(256)
encode_warp_kernel(const float* __restrict__ kps, const float* __restrict__ vis,
                   float* __restrict__ target, float* __restrict__ weight,
                   int tiles, int H, int W, float in_w, float in_h, EncodeConst ec) {
    extern __shared__ float lut[];
    fill_patch_lut(lut, ec);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int n4 = (H * W) >> 2, w4 = W >> 2;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    const int sx = 32 % w4, sy = 32 / w4;                      // how (column group, row) move when the float4 index grows by 32
    for (int tile = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; tile < tiles; tile += warps) {
        const PatchGeom g = patch_geometry(__ldg(kps + 2 * tile), __ldg(kps + 2 * tile + 1), __ldg(vis + tile), H, W, in_w, in_h, ec);
        if (lane == 0) weight[tile] = g.weight;
        float4* out = reinterpret_cast<float4*>(target) + (size_t)tile * n4;
        int y = lane / w4, x4 = lane - y * w4;
        for (int i = lane; i < n4; i += 32) {
            const int x = x4 << 2;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (g.active && y >= g.y_from && y < g.y_to && x + 3 >= g.x_from && x < g.x_to) {
                float e[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int xx = x + j;
                    e[j] = (xx >= g.x_from && xx < g.x_to) ? patch_value(lut, g, ec, xx, y) : 0.f;
                }
                v = make_float4(e[0], e[1], e[2], e[3]);
            }
            stg_stream(out + i, v);
            x4 += sx; y += sy;
            if (x4 >= w4) { x4 -= w4; ++y; }
        }
    }
}

// ---- the reference's second-generation encoders ------------------------------------------------
// data/coco_dataset.py:222-287 (_generate_heatmaps): mu in float32 (float32 joint * weak Python
// float), weight 1 if visible and mu inside the map, patch origin clamped to 0 before the patch
// slice is derived, so the pasted block is g[0 : br - ul_c] at [ul_c, br).
__device__ __forceinline__ PatchGeom patch_geometry_clipped(float kx, float ky, float vis, int H, int W,
                                                            float sx, float sy, float radius_f, const EncodeConst& ec) {
    PatchGeom g;
    g.weight = 0.f;
    g.active = 0;
    g.ulx = g.uly = g.x_from = g.x_to = g.y_from = g.y_to = 0;
    if (!(vis > 0.f)) return g;
    const float mux = __fmul_rn(kx, sx), muy = __fmul_rn(ky, sy);
    if (mux < 0.f || muy < 0.f || mux >= (float)W || muy >= (float)H) return g;   // data/coco_dataset.py:250
    g.weight = 1.f;
    const int ulx = max(0, (int)__fsub_rn(mux, radius_f)), uly = max(0, (int)__fsub_rn(muy, radius_f));
    const int brx = min(W, (int)__fadd_rn(__fadd_rn(mux, radius_f), 1.f)), bry = min(H, (int)__fadd_rn(__fadd_rn(muy, radius_f), 1.f));
    g.ulx = ulx; g.uly = uly;
    g.x_from = ulx; g.x_to = min(brx, ulx + ec.ntap);
    g.y_from = uly; g.y_to = min(bry, uly + ec.ntap);
    g.active = (g.x_to > g.x_from) && (g.y_to > g.y_from);
    return g;
}

template <int MODE>
__global__ void __launch_bounds__(256)
encode_genb_kernel(const float* __restrict__ kps, const float* __restrict__ vis,
                   float* __restrict__ target, float* __restrict__ weight,
                   int tiles, int H, int W, float sx, float sy, float radius_f, EncodeConst ec) {
    extern __shared__ float lut[];
    __shared__ PatchGeom geom;
    __shared__ float centre[2];
    if (MODE == GBCODEC_ENCODE_PATCH_CLIPPED) fill_patch_lut(lut, ec);
    const int n4 = (H * W) >> 2;
    const int w4 = W >> 2;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        __syncthreads();
        if (threadIdx.x == 0) {
            if (MODE == GBCODEC_ENCODE_PATCH_CLIPPED) {
                geom = patch_geometry_clipped(kps[2 * tile], kps[2 * tile + 1], vis[tile], H, W, sx, sy, radius_f, ec);
                weight[tile] = geom.weight;
            } else {
                // data/pose_transforms.py:427-451: centre = kp * (hm / in) in float32, inside test on the centre
                const float cx = __fmul_rn(kps[2 * tile], sx), cy = __fmul_rn(kps[2 * tile + 1], sy);
                const bool on = vis[tile] > 0.f && cx >= 0.f && cx < (float)W && cy >= 0.f && cy < (float)H;
                centre[0] = cx; centre[1] = cy;
                geom.active = on ? 1 : 0;
                weight[tile] = on ? 1.f : 0.f;
            }
        }
        __syncthreads();
        const PatchGeom g = geom;
        float4* out = reinterpret_cast<float4*>(target) + (size_t)tile * n4;
        if (MODE == GBCODEC_ENCODE_PATCH_CLIPPED) {
            for (int i = threadIdx.x; i < n4; i += blockDim.x) {
                const int y = i / w4, x = (i - y * w4) << 2;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (g.active && y >= g.y_from && y < g.y_to && x + 3 >= g.x_from && x < g.x_to) {
                    float e[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int xx = x + j;
                        e[j] = (xx >= g.x_from && xx < g.x_to) ? patch_value(lut, g, ec, xx, y) : 0.f;
                    }
                    v = make_float4(e[0], e[1], e[2], e[3]);
                }
                stg_stream(out + i, v);
            }
        } else {
            const float cx = centre[0], cy = centre[1];
            for (int i = threadIdx.x; i < n4; i += blockDim.x) {
                const int y = i / w4, x = (i - y * w4) << 2;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (g.active) {
                    // exp(-((x - x0)^2 + (y - y0)^2) / (2 sigma^2)), every step rounded to float32 as numpy does
                    const float dy = __fsub_rn((float)y, cy), dy2 = __fmul_rn(dy, dy);
                    float e[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float dx = __fsub_rn((float)(x + j), cx);
                        e[j] = expf(__fdiv_rn(-__fadd_rn(__fmul_rn(dx, dx), dy2), ec.two_sigma_sq));
                    }
                    v = make_float4(e[0], e[1], e[2], e[3]);
                }
                stg_stream(out + i, v);
            }
        }
    }
}

int launch_encode(const float* kps, const float* vis, float* target, float* weight,
                  int B, int K, int H, int W, float in_w, float in_h, double sigma, int mode, cudaStream_t stream) {
    const EncodeConst ec = make_encode_const(sigma);
    const int tiles = B * K;
    const int grid = tiles < 148 * 16 ? tiles : 148 * 16;
    const size_t smem = (size_t)ec.lut_size * sizeof(float);
    if (smem > 40 * 1024) return fail(GBCODEC_ERR_BAD_ARGUMENT, "encode: sigma %g needs a %zu-byte patch table", sigma, smem);
    // Gen-B scales: float32(W / in_w) — a weak Python float meeting a float32 array (NEP 50)
    const float sx = (float)((double)W / (double)in_w), sy = (float)((double)H / (double)in_h);
    switch (mode) {
        case GBCODEC_ENCODE_PATCH:
            if (H * W <= 8192 && tiles >= 148 * 8) {               // small tiles, enough of them: one warp per tile
                const int wgrid = (tiles + 7) / 8 < 148 * 8 ? (tiles + 7) / 8 : 148 * 8;
                note_launch(), encode_warp_kernel<<<wgrid, 256, smem, stream>>>(kps, vis, target, weight, tiles, H, W, in_w, in_h, ec);
                return check_launch("encode_warp_kernel");
            }
            note_launch(), encode_kernel<<<grid, 256, smem, stream>>>(kps, vis, target, weight, tiles, H, W, in_w, in_h, ec);
            return check_launch("encode_kernel");
        case GBCODEC_ENCODE_PATCH_CLIPPED:
            note_launch(), encode_genb_kernel<GBCODEC_ENCODE_PATCH_CLIPPED><<<grid, 256, smem, stream>>>(kps, vis, target, weight, tiles, H, W, sx, sy, (float)(sigma * 3.0), ec);
            return check_launch("encode_genb_kernel<clipped>");
        case GBCODEC_ENCODE_DENSE:
            note_launch(), encode_genb_kernel<GBCODEC_ENCODE_DENSE><<<grid, 256, 0, stream>>>(kps, vis, target, weight, tiles, H, W, sx, sy, 0.f, ec);
            return check_launch("encode_genb_kernel<dense>");
        default:
            return fail(GBCODEC_ERR_BAD_ARGUMENT, "encode: mode=%d", mode);
    }
}

}  // namespace gbc
