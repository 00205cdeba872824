// head_epilogue.cu — the variance branch's tail reduced to what the loss reads.
//
// The head ends its variance branch in Softplus (models/fusion_head.py:245-251) and the loss uses that map only through
// its per-tile mean (compute_heatmap_variance / variance_alignment_loss, :467-478: `pred_var.mean(dim=(2, 3))`).  Stock, the
// (B,K,H,W) map makes five trips through HBM per step (written by Softplus, read by the loss, its gradient written by the
// loss and read by Softplus' backward, which also re-reads the map) — 28N bytes per tile with the conv's own write and
// gradient read.  Here the last 1x1 convolution's RAW output (the caller drops the Softplus module) is reduced straight to
// mean_N(softplus(raw)): (B,K) floats feed gbcodec_fusion_step_vmean_f32, whose d(total)/d(mean) comes back through the
// second kernel as the gradient of the raw map:  d raw_i = g_tile / N * sigmoid(raw_i).
// Algorithmic bytes per tile: 4N (forward) + 8N (backward) = 12N; the fused step itself drops from 24N to 16N.
//
// torch.nn.Softplus(beta = 1, threshold = 20): x > 20 -> x, else log1p(exp(x)); its derivative sigmoid(x), 1 above the
// threshold.  One warp per tile, 128-bit loads, fixed summation order (bit-reproducible).
#include "common.cuh"

namespace gbc {

__device__ __forceinline__ float softplus_t20(float x) { return x > 20.f ? x : log1pf(expf(x)); }
__device__ __forceinline__ float softplus_grad_t20(float x) { return x > 20.f ? 1.f : 1.f / (1.f + expf(-x)); }

__global__ void __launch_bounds__(256)
softplus_mean_kernel(const float4* __restrict__ raw, float* __restrict__ mean, int tiles, int n4, float inv_n) {
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    for (int tile = blockIdx.x * wpb + (threadIdx.x >> 5); tile < tiles; tile += gridDim.x * wpb) {
        const float4* src = raw + (size_t)tile * n4;
        float acc = 0.f;
        for (int i = lane; i < n4; i += 32) {
            const float4 v = ldg_stream(src + i);
            acc += (softplus_t20(v.x) + softplus_t20(v.y)) + (softplus_t20(v.z) + softplus_t20(v.w));
        }
        acc = warp_sum(acc);
        if (lane == 0) mean[tile] = acc * inv_n;
    }
}

__global__ void __launch_bounds__(256)
softplus_mean_backward_kernel(const float4* __restrict__ raw, const float* __restrict__ grad_mean, float4* __restrict__ grad_raw,
                              int tiles, int n4, float inv_n) {
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    for (int tile = blockIdx.x * wpb + (threadIdx.x >> 5); tile < tiles; tile += gridDim.x * wpb) {
        const float g = __ldg(grad_mean + tile) * inv_n;
        const float4* src = raw + (size_t)tile * n4;
        float4* dst = grad_raw + (size_t)tile * n4;
        for (int i = lane; i < n4; i += 32) {
            const float4 v = ldg_stream(src + i);
            stg_stream(dst + i, make_float4(g * softplus_grad_t20(v.x), g * softplus_grad_t20(v.y), g * softplus_grad_t20(v.z), g * softplus_grad_t20(v.w)));
        }
    }
}

int launch_softplus_mean(const float* raw, float* mean, int B, int K, int H, int W, cudaStream_t s) {
    const int tiles = B * K, n4 = (H * W) >> 2;
    const int grid = (tiles + 7) / 8 < 148 * 8 ? (tiles + 7) / 8 : 148 * 8;
    note_launch(), softplus_mean_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<const float4*>(raw), mean, tiles, n4, 1.0f / (float)(H * W));
    return check_launch("softplus_mean_kernel");
}

int launch_softplus_mean_backward(const float* raw, const float* grad_mean, float* grad_raw, int B, int K, int H, int W, cudaStream_t s) {
    const int tiles = B * K, n4 = (H * W) >> 2;
    const int grid = (tiles + 7) / 8 < 148 * 8 ? (tiles + 7) / 8 : 148 * 8;
    note_launch(), softplus_mean_backward_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<const float4*>(raw), grad_mean,
                                                                      reinterpret_cast<float4*>(grad_raw), tiles, n4, 1.0f / (float)(H * W));
    return check_launch("softplus_mean_backward_kernel");
}

}  // namespace gbc
