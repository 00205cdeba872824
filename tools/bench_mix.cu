// bench_mix.cu — HBM bandwidth of streaming kernels with the read:write mix of the codec kernels
// (no arithmetic): what a kernel with the fused step's traffic — read 8N, write 16N bytes per tile —
// can reach at best on this part, next to pure read, pure write and 1:1 copy.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/bench_mix.cu -o tools/bench_mix && tools/bench_mix
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void stg_stream(float4* p, const float4& v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// one CTA per tile of N4 float4: reads NR input tiles, writes NW output tiles
template <int NR, int NW, int NIT>
__global__ void __launch_bounds__(192) mix_kernel(const float4* __restrict__ in, float4* __restrict__ out, int n4) {
    const size_t tile = blockIdx.x;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 v[NR > 0 ? NR * NIT : 1];
#pragma unroll
    for (int r = 0; r < NR; ++r)
#pragma unroll
        for (int it = 0; it < NIT; ++it) v[r * NIT + it] = ldg_stream(in + ((size_t)r * gridDim.x + tile) * n4 + it * blockDim.x + threadIdx.x);
#pragma unroll
    for (int i = 0; i < NR * NIT; ++i) { acc.x += v[i].x; acc.y += v[i].y; acc.z += v[i].z; acc.w += v[i].w; }
    if (NW == 0) { if (acc.x == 12345.678f) out[tile] = acc; return; }
#pragma unroll
    for (int w = 0; w < NW; ++w)
#pragma unroll
        for (int it = 0; it < NIT; ++it) stg_stream(out + ((size_t)w * gridDim.x + tile) * n4 + it * blockDim.x + threadIdx.x, acc);
}

template <int NR, int NW>
static void run(const char* name, const float4* in, float4* out, int tiles, int n4) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    for (int i = 0; i < 3; ++i) mix_kernel<NR, NW, 4><<<tiles, 192>>>(in, out, n4);
    CK(cudaEventRecord(a));
    const int reps = 20;
    for (int i = 0; i < reps; ++i) mix_kernel<NR, NW, 4><<<tiles, 192>>>(in, out, n4);
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b)); ms /= reps;
    const double bytes = (double)(NR + NW) * tiles * n4 * 16.0;
    printf("{\"kernel\": \"%s\", \"read_tiles\": %d, \"write_tiles\": %d, \"tiles\": %d, \"ms\": %.4f, \"GBps\": %.1f}\n", name, NR, NW, tiles, ms, bytes / ms / 1e6);
}

int main(int argc, char** argv) {
    const int tiles = argc > 1 ? atoi(argv[1]) : 17408, n4 = 768;
    float4 *in, *out;
    CK(cudaMalloc(&in, (size_t)4 * tiles * n4 * 16)); CK(cudaMalloc(&out, (size_t)4 * tiles * n4 * 16));
    CK(cudaMemset(in, 0, (size_t)4 * tiles * n4 * 16));
    run<1, 0>("read 1", in, out, tiles, n4);
    run<2, 0>("read 2", in, out, tiles, n4);
    run<4, 0>("read 4", in, out, tiles, n4);
    run<0, 1>("write 1", in, out, tiles, n4);
    run<0, 4>("write 4", in, out, tiles, n4);
    run<1, 1>("copy 1:1", in, out, tiles, n4);
    run<2, 2>("copy 2:2", in, out, tiles, n4);
    run<2, 4>("fused-step mix 2:4", in, out, tiles, n4);
    run<3, 4>("loss mix 3:4", in, out, tiles, n4);
    run<2, 1>("genb mix 2:1", in, out, tiles, n4);
    return 0;
}
