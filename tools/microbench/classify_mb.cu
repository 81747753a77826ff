// How long does one classification pass over 2 (or PPT) points per thread take for a 512-thread CTA?
// Variants isolate the FP64 test, the integer accumulation and the REDUX stage-1 reduction.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../clusteringsegmentation-1_b200/csrc/dq_split_math.cuh"
using namespace dq;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint64_t warp_sum_redux(uint64_t v) {
  if (!__any_sync(0xffffffffu, (v >> 27) != 0)) return __reduce_add_sync(0xffffffffu, (unsigned)v);
  const unsigned lo = __reduce_add_sync(0xffffffffu, (unsigned)(v & 0xFFFFFFu));
  const unsigned hi = __reduce_add_sync(0xffffffffu, (unsigned)(v >> 24));
  return (uint64_t)lo + ((uint64_t)hi << 24);
}

__device__ __forceinline__ uint64_t warp_sum_2limb(uint64_t v) {
  const unsigned lo = __reduce_add_sync(0xffffffffu, (unsigned)(v & 0xFFFFFFu));
  const unsigned hi = __reduce_add_sync(0xffffffffu, (unsigned)(v >> 24));
  return (uint64_t)lo + ((uint64_t)hi << 24);
}
template <int PPT, int MODE, int RED>
__global__ void __launch_bounds__(512, 1) k(const uint2 *pts, int iters, uint64_t *sink, long long *cycles) {
  __shared__ PassParams spp;
  __shared__ uint64_t red[16][8];
  if (threadIdx.x == 0) { spp.a = 1234.5; spp.r[0] = 3.25; spp.r[1] = -1.5; spp.r[2] = 7.125; spp.axis = 1; }
  __syncthreads();
  uint2 p[PPT];
#pragma unroll
  for (int k2 = 0; k2 < PPT; ++k2) p[k2] = pts[threadIdx.x + k2 * 512];
  uint64_t total = 0;
  long long t0 = clock64(), tc = 0, tr = 0;
  for (int it = 0; it < iters; ++it) {
    long long a0 = clock64();
    PassParams pp = spp;
    pp.a += it;  // keep the loop from being hoisted
    AccD acc = acc_zero();
#pragma unroll
    for (int k2 = 0; k2 < PPT; ++k2) {
      if (MODE == 0) { if (goes_new_t<false>(pp, to_point(p[k2]))) acc_add(acc, p[k2], false); }
      if (MODE == 1) { if (goes_new_t<false>(pp, to_point(p[k2]))) acc.n += 1; }             // test only
      if (MODE == 2) { if ((p[k2].x + it) & 1) acc_add(acc, p[k2], false); }                  // accumulate only
      if (MODE == 3) {  // branch-free: multiplicity forced to zero for points that stay
        const bool t = goes_new_t<false>(pp, to_point(p[k2]));
        const uint32_t c = t ? p[k2].y : 0u;
        const uint32_t R = (p[k2].x >> 16) & 0xFFu, G = (p[k2].x >> 8) & 0xFFu, B = p[k2].x & 0xFFu;
        acc.cnt += c; acc.r += (uint64_t)c * R; acc.g += (uint64_t)c * G; acc.b += (uint64_t)c * B; acc.n += t ? 1u : 0u;
      }
    }
    uint64_t v[8];
    acc_words(acc, v);
    long long a1 = clock64();
    // stage 1
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (RED == 0) {
#pragma unroll
      for (int w = 0; w < 5; ++w) {
        const uint64_t s = (w == 4) ? (uint64_t)__reduce_add_sync(0xffffffffu, (unsigned)v[w]) : warp_sum_redux(v[w]);
        if (lane == 0) red[warp][w] = s;
      }
    } else if (RED == 1) {
      uint64_t s[5];
#pragma unroll
      for (int w = 0; w < 4; ++w) s[w] = warp_sum_2limb(v[w]);
      s[4] = __reduce_add_sync(0xffffffffu, (unsigned)v[4]);
      if (lane < 5) red[warp][lane] = s[lane == 0 ? 0 : lane == 1 ? 1 : lane == 2 ? 2 : lane == 3 ? 3 : 4];
    } else {
      // one vote for all words, then 1 or 2 REDUX per word without further branches
      const bool big = __any_sync(0xffffffffu, ((v[0] | v[1] | v[2] | v[3]) >> 27) != 0);
      uint64_t s[5];
      if (!big) {
#pragma unroll
        for (int w = 0; w < 5; ++w) s[w] = __reduce_add_sync(0xffffffffu, (unsigned)v[w]);
      } else {
#pragma unroll
        for (int w = 0; w < 4; ++w) s[w] = warp_sum_2limb(v[w]);
        s[4] = __reduce_add_sync(0xffffffffu, (unsigned)v[4]);
      }
      if (lane == 0) {
#pragma unroll
        for (int w = 0; w < 5; ++w) red[warp][w] = s[w];
      }
    }
    __syncthreads();
    long long a2 = clock64();
    total += red[(it + warp) & 15][it & 3];
    tc += a1 - a0; tr += a2 - a1;
    __syncthreads();
  }
  long long t1 = clock64();
  sink[blockIdx.x * blockDim.x + threadIdx.x] = total;
  if (threadIdx.x == 0 && blockIdx.x == 0) { cycles[0] = (t1 - t0) / iters; cycles[1] = tc / iters; cycles[2] = tr / iters; }
}

template <int PPT, int MODE, int RED>
int run(const uint2 *pts, uint64_t *sink, long long *cyc, const char *name) {
  k<PPT, MODE, RED><<<148, 512>>>(pts, 2000, sink, cyc);
  CK(cudaDeviceSynchronize());
  long long h[3]; CK(cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
  printf("%-34s ppt=%d: %5lld cycles/iter (classify %lld, stage1+sync %lld)\n", name, PPT, h[0], h[1], h[2]);
  return 0;
}
int main() {
  uint2 *pts; uint64_t *sink; long long *cyc;
  CK(cudaMalloc(&pts, 512 * 16 * 8)); CK(cudaMalloc(&sink, 148 * 512 * 8)); CK(cudaMalloc(&cyc, 64));
  uint2 h[512 * 16];
  for (int i = 0; i < 512 * 16; ++i) h[i] = make_uint2(((i * 2654435761u) >> 8) & 0xFFFFFF, 1 + (i % 97));
  CK(cudaMemcpy(pts, h, sizeof(h), cudaMemcpyHostToDevice));
  run<2, 0, 0>(pts, sink, cyc, "adaptive per word");
  run<2, 0, 1>(pts, sink, cyc, "unconditional 2-limb");
  run<2, 0, 2>(pts, sink, cyc, "one vote, then 1 or 2 limbs");
  run<2, 3, 2>(pts, sink, cyc, "branch-free accumulate, one vote");
  run<8, 3, 2>(pts, sink, cyc, "branch-free accumulate, one vote");
  run<8, 0, 0>(pts, sink, cyc, "adaptive per word");
  run<8, 0, 2>(pts, sink, cyc, "one vote, then 1 or 2 limbs");
  return 0;
}
