// dq_common.cuh -- shared device/host helpers for the DivQuant B200 path (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

// The reference's error convention for this path is "message on stderr + abort()"
// (check_mem, DivQuantMapColors.cpp:43-51; size mismatch, DivQuantCluster.cpp:1021-1026).
// A CUDA failure has no better recovery inside void entry points, so it follows the same rule:
// there is no CPU fallback.
#define DQ_CUDA_CHECK(expr)                                                                      \
  do {                                                                                           \
    cudaError_t dq_err__ = (expr);                                                               \
    if (dq_err__ != cudaSuccess) {                                                               \
      fprintf(stderr, "divquant_b200: CUDA error %s at %s:%d: %s\n", cudaGetErrorName(dq_err__), \
              __FILE__, __LINE__, cudaGetErrorString(dq_err__));                                 \
      abort();                                                                                   \
    }                                                                                            \
  } while (0)

#include <atomic>

namespace dq {

// Function attributes (the dynamic shared-memory limit) belong to a device's context: "already raised" is remembered
// per device ordinal.  Several host threads (pipeline lanes) may ask at once: raising it twice is harmless.
struct PerDeviceLimit {
  std::atomic<size_t> raised[64];
  PerDeviceLimit() {
    for (auto &r : raised) r = 0;
  }
  // true when the caller has to raise the limit to `bytes` on the current device (and should call done afterwards)
  bool needs(size_t bytes, int *device_out) {
    int dev = 0;
    cudaGetDevice(&dev);
    *device_out = dev;
    return dev < 0 || dev >= 64 || bytes > raised[dev];
  }
  void done(size_t bytes, int dev) {
    if (dev >= 0 && dev < 64) raised[dev] = bytes;
  }
};
#define DQ_RAISE_SMEM(kernel, bytes)                                                                              \
  do {                                                                                                            \
    static ::dq::PerDeviceLimit dq_lim__;                                                                         \
    int dq_dev__ = 0;                                                                                             \
    if ((bytes) > 48 * 1024 && dq_lim__.needs((bytes), &dq_dev__)) {                                              \
      DQ_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)));     \
      dq_lim__.done((bytes), dq_dev__);                                                                           \
    }                                                                                                             \
  } while (0)

// Number of 24-bit colours = bins of the direct-addressed table.
constexpr uint32_t kColourBins = 1u << 24;

// ---- IEEE-754 double arithmetic that can never be contracted into FMAs -------------------------
// The reference build (x86-64 SSE2) rounds after every multiply and add; the decisions of the
// divisive phase (cut test, hyperplane test, TSE arg-max, .5 rounding) are taken on those values.
__device__ __forceinline__ double fmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double fadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double fsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double fdiv(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ double fsq(double a) { return __dmul_rn(a, a); }

// ---- L2-coherent accesses for data exchanged between CTAs inside one persistent kernel ----------
__device__ __forceinline__ uint64_t ld_cg_u64(const uint64_t *p) {
  return __ldcg(reinterpret_cast<const unsigned long long *>(p));
}
__device__ __forceinline__ uint32_t ld_cg_u32(const uint32_t *p) { return __ldcg(p); }
__device__ __forceinline__ double ld_cg_f64(const double *p) { return __ldcg(p); }
__device__ __forceinline__ uint2 ld_cg_u2(const uint2 *p) { return __ldcg(p); }

// ---- grid-wide barrier for the persistent (cooperatively launched) split kernel ------------------
// One monotonically increasing arrival counter; every CTA's thread 0 keeps the value it waits for.
__device__ __forceinline__ void grid_barrier(unsigned int *counter, unsigned int &target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += gridDim.x;
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(counter), "r"(1u) : "memory");
    unsigned int seen;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
    } while (seen < target);
  }
  __syncthreads();
}

__device__ __forceinline__ uint64_t warp_sum_u64(uint64_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace dq
