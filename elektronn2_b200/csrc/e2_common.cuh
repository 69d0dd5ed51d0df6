// Shared internals of libe2b200.so (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>

#include "e2b200.h"

struct e2_handle {
  int device;
  int sm_count;
  int64_t launches;
  char err[512];
  void* tmap_cache;  // tcgen05 path: host-side tensor-map cache (opaque)
};

// launch counter (e2_launch_count): atomic, a handle may be driven from several host threads / streams
static inline void e2_count_launch(e2_handle* h) { __atomic_fetch_add(&h->launches, (int64_t)1, __ATOMIC_RELAXED); }

static inline int e2_fail(e2_handle* h, int code, const char* fmt, ...) {
  if (h) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(h->err, sizeof(h->err), fmt, ap);
    va_end(ap);
  }
  return code;
}

#define E2_REQUIRE(h, cond, ...)                                   \
  do {                                                             \
    if (!(cond)) return e2_fail((h), E2_ERR_INVALID, __VA_ARGS__); \
  } while (0)

#define E2_CUDA_CHECK(h, what)                                                                       \
  do {                                                                                               \
    cudaError_t e__ = cudaGetLastError();                                                            \
    if (e__ != cudaSuccess) return e2_fail((h), E2_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e__)); \
  } while (0)

static inline bool e2_tensor_ok(const e2_tensor* t) {
  return t && t->n > 0 && t->z > 0 && t->x > 0 && t->y > 0 && t->c > 0 && t->c_pitch >= t->c;
}
static inline int64_t e2_positions(const e2_tensor* t) { return (int64_t)t->n * t->z * t->x * t->y; }

static inline int e2_grid_1d(int64_t work, int threads, int sm_count, int max_waves = 32) {
  int64_t g = (work + threads - 1) / threads;
  int64_t cap = (int64_t)sm_count * max_waves;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// Division of a small index (< 2^20) by a launch-constant divisor (< 4096) as one multiply-high; m == 0 selects
// the plain division (divisor 1 or out of the exact range).
struct E2FastDiv {
  uint32_t m, d;
  __device__ __forceinline__ uint32_t div(uint32_t t) const { return m ? __umulhi(t, m) : t / d; }
};
static inline E2FastDiv e2_fastdiv(uint32_t d, uint64_t max_t) {
  E2FastDiv f;
  f.d = d;
  f.m = (d > 1 && d < 4096 && max_t < (1u << 20)) ? (uint32_t)(((1ull << 32) + d - 1) / d) : 0u;
  return f;
}

// lin / relu (every BASELINE config) stay inline; the transcendental activations live behind a call so that
// unrolled epilogues do not carry a copy of tanhf / expf per element (instruction-cache footprint).
static __device__ __noinline__ float e2_apply_act_slow(float v, int act) {
  switch (act) {
    case E2_ACT_TANH: return tanhf(v);
    case E2_ACT_SIGMOID: return 1.f / (1.f + __expf(-v));
    case E2_ACT_ABS: return fabsf(v);
    case E2_ACT_SOFTPLUS: return v > 20.f ? v : log1pf(expf(v));                      // T.nnet.softplus
    case E2_ACT_ELU: return v > 0.f ? v : expm1f(v);                                   // T.nnet.elu(x, alpha=1)
    case E2_ACT_SELU: return 1.0507009873554805f * (v > 0.f ? v : 1.6732632423543772f * expm1f(v));   // computations.py:89-98
    default: return v;
  }
}
__device__ __forceinline__ float e2_apply_act(float v, int act) {
  if (act == E2_ACT_LIN) return v;
  if (act == E2_ACT_RELU) return v > 0.f ? v : 0.f;
  return e2_apply_act_slow(v, act);
}

// round-to-nearest fp32 -> tf32 (kept in an fp32 container)
__device__ __forceinline__ float e2_round_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}

// average / sum pooling (e2_epilogue.cu), reached through e2_maxpool3d_fwd / _bwd with d->mode != E2_POOL_MAX
int e2_launch_avgpool_fwd(e2_handle* h, const e2_pool_desc* d, const float* x, float* y, cudaStream_t s);
int e2_launch_avgpool_bwd(e2_handle* h, const e2_pool_desc* d, const float* dy, float* dx, cudaStream_t s);

// internal launchers shared between translation units
int e2_conv_tc_supported(const e2_handle* h, int c_in, int c_out, int c_in_pitch, int c_out_pitch);
