// bench_loss.cu — stand-alone timing harness for the fused codec step through the C ABI
// (no Python, no torch): synthetic batch on the device, W warm-up + K timed calls of
// gbcodec_fusion_step_f32, CUDA-event time of the whole call sequence and of the tile kernel.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/bench_loss.cu -o tools/bench_loss \
//        -Iinclude -Linfantposeestimation_gaussianbias_b200 -lgbcodec -Xlinker -rpath=$PWD/infantposeestimation_gaussianbias_b200
//   tools/bench_loss [B=1024] [K=17] [H=64] [W=48] [steps=20] [warmup=5] [sigma=2.0]
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "gbcodec.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)
#define GB(x) do { int s_ = (x); if (s_ != 0) { fprintf(stderr, "%s:%d gbcodec %d: %s\n", __FILE__, __LINE__, s_, gbcodec_last_error()); exit(1); } } while (0)

__device__ __forceinline__ unsigned hash32(unsigned x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }
__device__ __forceinline__ float urand(unsigned long long i, unsigned seed) { return (hash32((unsigned)i ^ hash32((unsigned)(i >> 32) + seed)) >> 8) * (1.0f / 16777216.0f); }
__device__ __forceinline__ float nrand(unsigned long long i, unsigned seed) {
    const float u1 = fmaxf(urand(2 * i, seed), 1e-7f), u2 = urand(2 * i + 1, seed);
    return sqrtf(-2.f * logf(u1)) * cospif(2.f * u2);
}
__global__ void fill_normal(float* p, size_t n, float scale, unsigned seed, int softplus) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float v = nrand(i, seed);
        p[i] = softplus ? log1pf(expf(v)) : scale * v;
    }
}
__global__ void make_kps(float* kps, float* jit, float* vis, float* vis2, int n, float in_w, float in_h, float stride) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float u = urand(i, 11);
    vis[i] = u < 0.15f ? 0.f : (u < 0.40f ? 1.f : 2.f);
    vis2[i] = 2.f;
    kps[2 * i] = (urand(i, 12) * 1.2f - 0.1f) * in_w;
    kps[2 * i + 1] = (urand(i, 13) * 1.2f - 0.1f) * in_h;
    jit[2 * i] = kps[2 * i] + nrand(i, 14) * 1.5f * stride;
    jit[2 * i + 1] = kps[2 * i + 1] + nrand(i, 15) * 1.5f * stride;
}
__global__ void shape_hm(float* hm, size_t n, int tile, unsigned seed) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float amp = 0.3f + 0.9f * urand(i / tile, seed);
        hm[i] = amp * hm[i] + 0.05f * nrand(i, seed + 1);
    }
}

int main(int argc, char** argv) {
    const int B = argc > 1 ? atoi(argv[1]) : 1024, K = argc > 2 ? atoi(argv[2]) : 17;
    const int H = argc > 3 ? atoi(argv[3]) : 64, W = argc > 4 ? atoi(argv[4]) : 48;
    const int steps = argc > 5 ? atoi(argv[5]) : 20, warm = argc > 6 ? atoi(argv[6]) : 5;
    const double sigma = argc > 7 ? atof(argv[7]) : 2.0;
    const float in_w = 4.f * W, in_h = 4.f * H;
    const size_t tiles = (size_t)B * K, n = tiles * H * W;
    float *hm, *off, *var, *ghm, *goff, *gvar, *kps, *jit, *vis, *vis2, *wdummy, *losses, *coords, *scores, *alpha;
    CK(cudaMalloc(&hm, n * 4)); CK(cudaMalloc(&off, 2 * n * 4)); CK(cudaMalloc(&var, n * 4));
    CK(cudaMalloc(&ghm, n * 4)); CK(cudaMalloc(&goff, 2 * n * 4)); CK(cudaMalloc(&gvar, n * 4));
    CK(cudaMalloc(&kps, tiles * 8)); CK(cudaMalloc(&jit, tiles * 8)); CK(cudaMalloc(&vis, tiles * 4)); CK(cudaMalloc(&vis2, tiles * 4));
    CK(cudaMalloc(&wdummy, tiles * 4)); CK(cudaMalloc(&losses, 64)); CK(cudaMalloc(&coords, tiles * 8)); CK(cudaMalloc(&scores, tiles * 4));
    CK(cudaMalloc(&alpha, 8));
    const float ab[2] = {0.5f, 0.6224593312018546f};
    CK(cudaMemcpy(alpha, ab, 8, cudaMemcpyHostToDevice));
    make_kps<<<(int)((tiles + 255) / 256), 256>>>(kps, jit, vis, vis2, (int)tiles, in_w, in_h, 4.f);
    GB(gbcodec_encode_f32(jit, vis2, hm, wdummy, B, K, H, W, in_w, in_h, sigma, nullptr));
    shape_hm<<<148 * 8, 256>>>(hm, n, H * W, 21);
    fill_normal<<<148 * 8, 256>>>(off, 2 * n, 0.3f, 31, 0);
    fill_normal<<<148 * 8, 256>>>(var, n, 1.f, 41, 1);
    CK(cudaDeviceSynchronize());

    gbcodec_loss_desc d = {};
    d.B = B; d.K = K; d.H = H; d.W = W; d.in_w = in_w; d.in_h = in_h;
    const float lam[6] = {1.f, 1.f, .5f, .1f, .05f, .05f};
    for (int q = 0; q < 6; ++q) d.lambdas[q] = lam[q];
    d.target_sigma = sigma; d.encode_sigma = sigma; d.use_target_weight = 1;
    const int sk[16][2] = {{0,1},{0,2},{1,3},{2,4},{5,6},{5,7},{7,9},{6,8},{8,10},{5,11},{6,12},{11,12},{11,13},{13,15},{12,14},{14,16}};
    d.n_pairs = 16;
    for (int p = 0; p < 16; ++p) { d.pairs[p][0] = sk[p][0]; d.pairs[p][1] = sk[p][1]; }
    const size_t wsb = gbcodec_loss_workspace_bytes(B, K, H, W);
    void* ws; CK(cudaMalloc(&ws, wsb));
    cudaStream_t s; CK(cudaStreamCreate(&s));
    std::vector<cudaEvent_t> e0(steps), e1(steps);
    for (int i = 0; i < steps; ++i) { CK(cudaEventCreate(&e0[i])); CK(cudaEventCreate(&e1[i])); }
    cudaEvent_t ta, tb; CK(cudaEventCreate(&ta)); CK(cudaEventCreate(&tb));
    // BENCH_NODECODE=1: the step without its decode (measurement: what the decode costs inside the step kernel)
    const bool nodecode = getenv("BENCH_NODECODE") != nullptr;
    auto step = [&]() {
        if (nodecode) GB(gbcodec_fusion_loss_f32(&d, hm, off, var, nullptr, vis, kps, nullptr, nullptr, losses, ghm, goff, gvar, ws, wsb, s));
        else GB(gbcodec_fusion_step_f32(&d, hm, off, var, nullptr, vis, kps, nullptr, nullptr, losses, ghm, goff, gvar,
                                        alpha, alpha + 1, 2, GBCODEC_DECODE_REFINE | GBCODEC_DECODE_APPLY_OFFSET, coords, scores, ws, wsb, s));
    };
    for (int i = 0; i < warm; ++i) step();
    CK(cudaStreamSynchronize(s));
    CK(cudaEventRecord(ta, s));
    for (int i = 0; i < steps; ++i) { GB(gbcodec_profile_loss_kernel(e0[i], e1[i])); step(); }
    CK(cudaEventRecord(tb, s));
    CK(cudaStreamSynchronize(s));
    GB(gbcodec_profile_loss_kernel(nullptr, nullptr));
    float total = 0.f, kern = 0.f, kmin = 1e9f;
    CK(cudaEventElapsedTime(&total, ta, tb));
    for (int i = 0; i < steps; ++i) { float t; CK(cudaEventElapsedTime(&t, e0[i], e1[i])); kern += t; kmin = t < kmin ? t : kmin; }
    float hl[7]; CK(cudaMemcpy(hl, losses, 28, cudaMemcpyDeviceToHost));
    const double bytes = 24.0 * n;
    printf("{\"B\": %d, \"K\": %d, \"H\": %d, \"W\": %d, \"ms_per_step\": %.4f, \"kernel_ms_mean\": %.4f, \"kernel_ms_min\": %.4f, "
           "\"hm_per_s\": %.4g, \"kernel_GBps\": %.1f, \"total_loss\": %.6f, \"variant\": \"%s\"}\n",
           B, K, H, W, total / steps, kern / steps, kmin, tiles * steps / (total * 1e-3), bytes / (kern / steps * 1e-3) / 1e9, hl[6],
           getenv("GBCODEC_LOSS_KERNEL") ? getenv("GBCODEC_LOSS_KERNEL") : "default");
    return 0;
}
