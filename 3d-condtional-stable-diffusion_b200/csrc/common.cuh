// Shared helpers for libb200dm (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/b200dm.h"

void b200dm_set_error(const char* fmt, ...);

// ---- 16-bit activation / weight storage type ------------------------------------------------------------
// The library is built twice from the same sources: libb200dm.so stores activations and packed weights as bf16
// (8 significand bits, fp32 range), libb200dm_f16.so (-DB200DM_ACT_FP16) as IEEE fp16 (11 significand bits: 8x smaller
// storage rounding, range 65504).  tcgen05.mma kind::f16 runs both at the same rate and always accumulates in fp32.
#ifdef B200DM_ACT_FP16
typedef __half act_t;
typedef __half2 act2_t;
__device__ __forceinline__ float2 act2_to_float2(const act2_t v) { return __half22float2(v); }
__device__ __forceinline__ act2_t floats_to_act2(float a, float b) { return __floats2half2_rn(a, b); }
__device__ __forceinline__ float act_to_float(const act_t v) { return __half2float(v); }
__device__ __forceinline__ act_t float_to_act(float v) { return __float2half_rn(v); }
#define B200DM_ACT_NAME "fp16"
#else
typedef __nv_bfloat16 act_t;
typedef __nv_bfloat162 act2_t;
__device__ __forceinline__ float2 act2_to_float2(const act2_t v) { return __bfloat1622float2(v); }
__device__ __forceinline__ act2_t floats_to_act2(float a, float b) { return __floats2bfloat162_rn(a, b); }
__device__ __forceinline__ float act_to_float(const act_t v) { return __bfloat162float(v); }
__device__ __forceinline__ act_t float_to_act(float v) { return __float2bfloat16_rn(v); }
#define B200DM_ACT_NAME "bf16"
#endif

#define B2_CHECK_ARG(cond, ...)                 \
  do {                                          \
    if (!(cond)) {                              \
      b200dm_set_error(__VA_ARGS__);            \
      return B200DM_ERR_INVALID;                \
    }                                           \
  } while (0)

#define B2_CHECK_CUDA(expr)                                                                  \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      b200dm_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return B200DM_ERR_CUDA;                                                                \
    }                                                                                        \
  } while (0)

#define B2_CHECK_LAUNCH() B2_CHECK_CUDA(cudaGetLastError())

static inline int b2_num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  return sms;
}

// ---- programmatic dependent launch (PDL) ----------------------------------------------------------------
// Every per-step kernel calls pdl_launch_dependents() first (the next kernel of the stream / graph may start its
// prologue as soon as all CTAs of this one are resident) and pdl_wait() before its first access to global memory that
// a predecessor may have written (or may still read): barrier init, TMEM allocation, tensor-map prefetch and launch
// latency of kernel k+1 overlap the tail of kernel k.  B200DM_PDL=0 launches without the attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

bool b200dm_pdl_enabled();

template <typename... KArgs, typename... Args>
static inline cudaError_t b2_launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = b200dm_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// ---- small device helpers ------------------------------------------------------------------
struct __align__(16) bf16x8 {
  act2_t v[4];
};

__device__ __forceinline__ bf16x8 ldg_bf16x8(const bf16x8* p) {   // 16-byte read-only load
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  bf16x8 r;
  memcpy(&r, &u, 16);
  return r;
}
__device__ __forceinline__ void unpack8(const bf16x8& p, float (&f)[8]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = act2_to_float2(p.v[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ bf16x8 pack8(const float (&f)[8]) {
  bf16x8 p;
#pragma unroll
  for (int i = 0; i < 4; ++i) p.v[i] = floats_to_act2(f[2 * i], f[2 * i + 1]);
  return p;
}
// x * sigmoid(x).  ex2.approx.ftz == __expf outside the denormal range, and there 1 + e == 1 either way: same values as
// __fdividef(x, 1 + __expf(-x)) without the per-element range fix-up (FSETP + two predicated FMULs) of the non-ftz form.
__device__ __forceinline__ float silu_f(float x) {
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
  return __fdividef(x, 1.0f + e);
}
__device__ __forceinline__ float apply_act(float x, int act) {
  if (act == B200DM_ACT_SILU) return silu_f(x);
  if (act == B200DM_ACT_RELU) return fmaxf(x, 0.0f);
  return x;
}
// Runtime activation of a register vector with the kind test OUTSIDE the element loop.  With the test inside, the compiler
// emitted a branch per element and the N dependent MUFU.EX2 -> FADD -> MUFU.RCP -> FMUL chains ran one after the other
// (~80 cycles each): the SiLU epilogue of a 64 -> 64 conv tile took 9.3 k cycles against 3.6 k without activation and made
// the ResidualBlock conv1 launches epilogue-bound (clock64 timeline, tools/conv_trace.py EPI=1).
template <int N>
__device__ __forceinline__ void apply_act_vec(float (&v)[N], int act) {
  if (act == B200DM_ACT_SILU) {
#pragma unroll
    for (int j = 0; j < N; ++j) v[j] = silu_f(v[j]);
  } else if (act == B200DM_ACT_RELU) {
#pragma unroll
    for (int j = 0; j < N; ++j) v[j] = fmaxf(v[j], 0.0f);
  }
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
