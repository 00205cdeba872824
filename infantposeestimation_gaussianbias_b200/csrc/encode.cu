// encode.cu — Gaussian target tiles written straight into HBM.
// Restates COCOPoseDataset._generate_target (datasets/coco_dataset.py:185-250)
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

int launch_encode(const float* kps, const float* vis, float* target, float* weight,
                  int B, int K, int H, int W, float in_w, float in_h, double sigma, cudaStream_t stream) {
    const EncodeConst ec = make_encode_const(sigma);
    const int tiles = B * K;
    const int grid = tiles < 148 * 16 ? tiles : 148 * 16;
    const size_t smem = (size_t)ec.lut_size * sizeof(float);
    if (smem > 40 * 1024) return fail(GBCODEC_ERR_BAD_ARGUMENT, "encode: sigma %g needs a %zu-byte patch table", sigma, smem);
    encode_kernel<<<grid, 256, smem, stream>>>(kps, vis, target, weight, tiles, H, W, in_w, in_h, ec);
    return check_launch("encode_kernel");
}

}  // namespace gbc
