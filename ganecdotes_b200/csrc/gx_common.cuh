// Shared helpers for the ganecdotes_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include "../../include/ganecdotes_b200.h"

#define GX_SM_COUNT_FALLBACK 148

#define GX_CHECK_ARG(cond)            \
  do {                                \
    if (!(cond)) return GX_ERR_ARG;   \
  } while (0)

#define GX_CHECK_CUDA(expr)                          \
  do {                                               \
    cudaError_t _e = (expr);                         \
    if (_e != cudaSuccess) {                         \
      gx_set_last_cuda_error((int)_e);               \
      return GX_ERR_CUDA;                            \
    }                                                \
  } while (0)

#define GX_LAUNCH_CHECK() GX_CHECK_CUDA(cudaGetLastError())

void gx_set_last_cuda_error(int e);
int gx_sm_count();
int gx_umma_cta_budget();
int gx_stream_cta_budget();

static inline int gx_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

#ifdef __CUDACC__

// ---------------------------------------------------------------------------
// split-bf16: x = hi + lo with hi = bf16(x), lo = bf16(x - hi).  Three bf16 MMAs
// (hi*hi + hi*lo + lo*hi) then reproduce an fp32 product to ~2^-16 relative.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void gx_split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

__device__ __forceinline__ uint32_t gx_pack_bf16x2(__nv_bfloat16 a, __nv_bfloat16 b) {
  return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}

// two floats -> one packed bf16x2 word (a in the low half), round to nearest even: ONE F2FP.BF16.F32.PACK_AB instead of
// two single conversions and a PRMT
__device__ __forceinline__ uint32_t gx_cvt_bf16x2(float a, float b) {
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a));
  return d;
}
// split two floats into the packed hi word and the packed lo word (same values as gx_split_bf16 per element)
__device__ __forceinline__ void gx_split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  hi = gx_cvt_bf16x2(a, b);
  lo = gx_cvt_bf16x2(a - __uint_as_float(hi << 16), b - __uint_as_float(hi & 0xffff0000u));
}

// split 4 floats into packed hi (2 words) and lo (2 words)
__device__ __forceinline__ void gx_split4(const float4 v, uint2& hi, uint2& lo) {
  gx_split2(v.x, v.y, hi.x, lo.x);
  gx_split2(v.z, v.w, hi.y, lo.y);
}

__device__ __forceinline__ float gx_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float gx_warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// streaming 128-bit accesses that do not pollute L1
__device__ __forceinline__ float4 gx_ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void gx_stg_stream(float4* p, const float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}

// packed fp32x2 arithmetic (Blackwell FFMA2 / FADD2 / FMUL2): halves the FMA-pipe instruction count
// of the streaming kernels, which are bound by instruction issue rather than by arithmetic
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;"
      : "=l"(reinterpret_cast<unsigned long long&>(d))
      : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)),
        "l"(reinterpret_cast<unsigned long long&>(c)));
  return d;
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  float2 d;
  asm("add.rn.f32x2 %0, %1, %2;"
      : "=l"(reinterpret_cast<unsigned long long&>(d))
      : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)));
  return d;
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  float2 d;
  asm("mul.rn.f32x2 %0, %1, %2;"
      : "=l"(reinterpret_cast<unsigned long long&>(d))
      : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)));
  return d;
}

#endif  // __CUDACC__
