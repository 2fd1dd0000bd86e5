// common.cuh -- shared device/host helpers for libdrtk_b200 (sm_100a only).
//
// Nothing here is copied from the reference; where a helper has to reproduce the
// reference's *semantics* the reference file:line is cited.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/drtk_b200.h"

namespace drtk {

// SM count of the CURRENT device (148 on B200: 2 dies x 74 SMs), queried once per device and cached
// (capi.cu); sizes every persistent / one-wave grid.
int num_sms();
// images per launch when the batch index rides on gridDim.y / gridDim.z (limit 65535): larger batches are
// processed in consecutive slices by the entry points (batch items are independent in every kernel)
constexpr int64_t kMaxBatchPerLaunch = 32768;

struct Strides3 { int64_t s0, s1, s2; };
struct Strides4 { int64_t s0, s1, s2, s3; };

static inline Strides3 make3(const int64_t* s) { return Strides3{s[0], s[1], s[2]}; }
static inline Strides4 make4(const int64_t* s) { return Strides4{s[0], s[1], s[2], s[3]}; }

#define DRTK_CHECK_LAUNCH()                                  \
  do {                                                       \
    cudaError_t e__ = cudaGetLastError();                    \
    if (e__ != cudaSuccess) return static_cast<int>(e__);    \
  } while (0)

#define DRTK_CUDA(call)                                      \
  do {                                                       \
    cudaError_t e__ = (call);                                \
    if (e__ != cudaSuccess) return static_cast<int>(e__);    \
  } while (0)

// ---- exactly-rounded building blocks (never contracted by nvcc) -------------------------
// The library is compiled with -ftz=true, so these map to FMUL.FTZ / FADD.FTZ / FFMA.FTZ,
// the same SASS the reference build (--use_fast_math) uses.
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fma_rn(float a, float b, float c) { return __fmaf_rn(a, b, c); }

// MUFU.RCP, the reciprocal the reference's fast-math division / `1.0f / x` compile to.
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float rsqrt_approx(float x) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// epsclamp: keep |v| >= 1e-8 with the sign of v; non-negative and NaN take the +eps branch
// (src/include/cuda_math_helper.h:1036-1041, eps :63-65).
__device__ __forceinline__ float epsclamp(float v) {
  return (v < 0.f) ? fminf(v, -1e-8f) : fmaxf(v, 1e-8f);
}

// x*y - z*w the way nvcc contracts it in the reference build: second product rounded,
// first fused (observed in the sm_100 SASS of oracle/_ref/*.so).
__device__ __forceinline__ float diff_of_products(float x, float y, float z, float w) {
  return fma_rn(x, y, -mul_rn(z, w));
}

// ---- streaming (touch-once) global accesses ---------------------------------------------
__device__ __forceinline__ float4 ldg_stream_f4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ int4 ldg_stream_i4(const int32_t* p) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float ldg_stream_f(const float* p) {
  float r;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream_f4(float* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void stg_stream_i4(int32_t* p, int4 v) {
  asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// 256-bit global load (sm_100+, SASS LDG.E.ENL2.256): address must be 32-B aligned
struct float8 { float4 lo, hi; };
__device__ __forceinline__ float8 ldg_f8(const float* p) {
  float8 r;
  asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r.lo.x), "=f"(r.lo.y), "=f"(r.lo.z), "=f"(r.lo.w), "=f"(r.hi.x), "=f"(r.hi.y), "=f"(r.hi.z), "=f"(r.hi.w)
               : "l"(p));
  return r;
}

__device__ __forceinline__ void prefetch_l1(const void* p) {
  asm volatile("prefetch.global.L1 [%0];" :: "l"(p));
}

// fire-and-forget float reduction (REDG.ADD.F32)
__device__ __forceinline__ void red_add(float* p, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" :: "l"(p), "f"(v) : "memory");
}
// 128-bit vector reduction (sm_90+): one request for four consecutive floats
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// Segmented inclusive "suffix" sum over consecutive lanes sharing a key: after the call the
// first lane (head) of every run of equal keys holds the sum over the run.  `tail` marks the
// last lane of a run.  5 shuffle steps regardless of the run structure.
template <int NV>
__device__ __forceinline__ void seg_reduce_to_head(float (&val)[NV], unsigned tail_mask, int lane) {
  // distance from this lane to the end of its run
  const unsigned above = tail_mask >> lane;          // bit0 = own tail flag
  const int dist_to_tail = __ffs(above) - 1;         // >= 0 because lane 31 is always a tail
  // runs are short on fine meshes (~5 px): when no run of this warp is longer than 8 lanes, three
  // shuffle steps suffice (warp-uniform choice)
  const bool short_runs = __all_sync(0xffffffffu, dist_to_tail < 8);
#pragma unroll
  for (int off = 1; off < 8; off <<= 1) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float o = __shfl_down_sync(0xffffffffu, val[i], off);
      if (off <= dist_to_tail) val[i] += o;
    }
  }
  if (!short_runs) {
#pragma unroll
    for (int off = 8; off < 32; off <<= 1) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const float o = __shfl_down_sync(0xffffffffu, val[i], off);
        if (off <= dist_to_tail) val[i] += o;
      }
    }
  }
}

struct VecOk {
  // true when a [.., H, W] image with the given innermost strides can be accessed with
  // aligned 128-bit vectors along W
  static inline bool image(const void* p, int64_t W, int64_t sW, int64_t sH, int64_t sOuter0,
                           int64_t sOuter1 = 0) {
    return sW == 1 && (W % 4) == 0 && (sH % 4) == 0 && (sOuter0 % 4) == 0 && (sOuter1 % 4) == 0 &&
           (reinterpret_cast<uintptr_t>(p) % 16) == 0;
  }
};

}  // namespace drtk
