// Shared helpers for libmmtpsm_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "mmtpsm_b200.h"

namespace mmt {

// thread-local last-error text + process-wide launch counter (defined in status.cu)
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

// Optional event bracketing of a kernel family (status.cu); a no-op unless mmt_prof_enable(1).
bool prof_enabled();
void prof_begin(cudaStream_t stream);
void prof_end(int family, double flops, double bytes, cudaStream_t stream);
struct ProfScope {
  int family;
  double flops, bytes;
  cudaStream_t stream;
  bool on;
  ProfScope(int family_, double flops_, double bytes_, cudaStream_t s)
      : family(family_), flops(flops_), bytes(bytes_), stream(s), on(prof_enabled()) {
    if (on) prof_begin(stream);
  }
  ~ProfScope() {
    if (on) prof_end(family, flops, bytes, stream);
  }
};

inline int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  set_error("%s", buf);
  return code;
}

// Check the launch that was just issued (no synchronisation: asynchronous errors surface on the
// caller's next sync, exactly like any CUDA library).
inline int check_launch(const char* what) {
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(MMT_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return MMT_OK;
}

#define MMT_REQUIRE(cond, ...)                                     \
  do {                                                             \
    if (!(cond)) return ::mmt::fail(MMT_ERR_INVALID_ARGUMENT, __VA_ARGS__); \
  } while (0)

#define MMT_CUDA(call)                                                                   \
  do {                                                                                   \
    cudaError_t e__ = (call);                                                            \
    if (e__ != cudaSuccess)                                                              \
      return ::mmt::fail(MMT_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e__));        \
  } while (0)

constexpr int kNumSMs = 148;  // B200

template <typename T>
__host__ __device__ constexpr T ceil_div(T a, T b) {
  return (a + b - 1) / b;
}
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum_int(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum for blockDim.x <= 1024 (multiple of 32); result valid in every thread.
__device__ __forceinline__ float block_sum(float v, float* smem32) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) smem32[wid] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  float r = (lane < nw) ? smem32[lane] : 0.f;
  r = warp_sum(r);
  return r;
}

// 16-byte vector reduction to global memory (sm_90+): one transaction for four fp32 adds.
__device__ __forceinline__ void red_add_v4(float* addr, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}

}  // namespace mmt
