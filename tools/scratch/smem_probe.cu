#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
__global__ void __launch_bounds__(224, 3) probe(int off, int mode, float* out) {
    extern __shared__ __align__(128) unsigned char sm[];
    if (mode == 0) { if (threadIdx.x == 0) *reinterpret_cast<volatile float*>(sm + off) = 1.f; }
    else if (threadIdx.x == 0) {
        unsigned a = (unsigned)__cvta_generic_to_shared(sm + off);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(a), "r"(1) : "memory");
    }
    __syncthreads();
    if (threadIdx.x == 1) out[0] = *reinterpret_cast<volatile float*>(sm + off);
}
int main() {
    float* out; cudaMalloc(&out, 4);
    for (int threads : {192, 224, 256}) for (int smem : {40000, 76144, 76640, 100000}) for (int mode : {0, 1}) {
        cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(probe, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        probe<<<1, threads, smem>>>(smem - 64, mode, out);
        cudaError_t e = cudaDeviceSynchronize();
        printf("threads %d smem %d mode %d: %s\n", threads, smem, mode, cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
    }
    return 0;
}
