#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
template <int T>
__global__ void __launch_bounds__(T, 3) probe(int off, const float* src, float* out) {
    extern __shared__ __align__(128) unsigned char sm[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + off);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == T - 32) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(4096) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(smem_u32(sm)), "l"(src), "r"(4096), "r"(smem_u32(bar)) : "memory");
    }
    asm volatile("{\n .reg .pred p;\nW_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra D_%=;\n bra W_%=;\nD_%=:\n}" :: "r"(smem_u32(bar)), "r"(0) : "memory");
    if (threadIdx.x == 1) out[0] = reinterpret_cast<float*>(sm)[5];
}
template <int T> int run(const float* src, float* out) {
    for (int smem : {40000, 76144}) {
        cudaFuncSetAttribute(probe<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(probe<T>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        probe<T><<<4, T, smem>>>(smem - 64, src, out);
        cudaError_t e = cudaDeviceSynchronize();
        printf("threads %d smem %d: %s\n", T, smem, cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
    }
    return 0;
}
int main() {
    float *out, *src; cudaMalloc(&out, 4); cudaMalloc(&src, 1 << 20);
    if (run<192>(src, out)) return 1;
    if (run<224>(src, out)) return 1;
    return 0;
}
