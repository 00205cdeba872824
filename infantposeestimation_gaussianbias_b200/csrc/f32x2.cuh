// f32x2.cuh — packed single-precision pairs (sm_100 FFMA2 / FADD2 / FMUL2).
//
// Blackwell issues one instruction for two independent fp32 operations on an aligned register
// pair (PTX fma.rn.f32x2 / add.rn.f32x2 / mul.rn.f32x2).  Each half is the same IEEE operation as
// its scalar form, so results do not change; what changes is the number of issue slots — the tile
// kernels are bound by instruction issue, not by the FMA pipe.  A float4 read from shared memory
// or HBM already sits in two aligned pairs, packing and unpacking are register renames.
#pragma once
#include <cuda_runtime.h>

namespace gbc {

struct f2 { unsigned long long r; };

__device__ __forceinline__ f2 pack2(float lo, float hi) { f2 v; asm("mov.b64 %0, {%1,%2};" : "=l"(v.r) : "f"(lo), "f"(hi)); return v; }
__device__ __forceinline__ f2 splat2(float a) { return pack2(a, a); }
__device__ __forceinline__ void unpack2(f2 v, float& lo, float& hi) { asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v.r)); }
__device__ __forceinline__ float lo2(f2 v) { float a, b; unpack2(v, a, b); return a; }
__device__ __forceinline__ float hi2(f2 v) { float a, b; unpack2(v, a, b); return b; }
__device__ __forceinline__ float hsum2(f2 v) { float a, b; unpack2(v, a, b); return a + b; }

__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 v; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(v.r) : "l"(a.r), "l"(b.r), "l"(c.r)); return v; }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { f2 v; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(v.r) : "l"(a.r), "l"(b.r)); return v; }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { f2 v; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(v.r) : "l"(a.r), "l"(b.r)); return v; }
// a - b, exactly: the product (-1) * b is exact, so the fused form rounds once like the subtraction
__device__ __forceinline__ f2 sub2(f2 a, f2 b) { return fma2(b, splat2(-1.f), a); }

// a float4 as two pairs and back
struct f4 { f2 a, b; };     // a = (x, y), b = (z, w)
__device__ __forceinline__ f4 as_f4(const float4& v) { f4 r; r.a = pack2(v.x, v.y); r.b = pack2(v.z, v.w); return r; }
__device__ __forceinline__ float4 as_float4(const f4& v) {
    float4 r;
    unpack2(v.a, r.x, r.y); unpack2(v.b, r.z, r.w);
    return r;
}

}  // namespace gbc
