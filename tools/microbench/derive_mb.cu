// Warp-0 part of a pass: stage-2 reduction over the per-warp partials + derivation of the next hyperplane.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../clusteringsegmentation-1_b200/csrc/dq_split_math.cuh"
namespace dq {
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
struct JobConst { double tw, tm[3], cut; int32_t axis, buf; uint32_t begin, size; };
struct Sh { uint64_t red[32][8]; uint64_t tot[8]; PassParams pp; };

__device__ __forceinline__ void stage2_v0(Sh &S, int rows) {
  const int lane = threadIdx.x & 31;
  uint64_t x[5];
#pragma unroll
  for (int w = 0; w < 5; ++w) x[w] = lane < rows ? S.red[lane][w] : 0ull;
#pragma unroll
  for (int w = 0; w < 5; ++w) {
    const unsigned a = __reduce_add_sync(0xffffffffu, (unsigned)(x[w] & 0x7FFFFFFu));
    const unsigned c = __reduce_add_sync(0xffffffffu, (unsigned)(x[w] >> 27));
    if (lane == 0) S.tot[w] = (uint64_t)a + ((uint64_t)c << 27);
  }
  __syncwarp();
}
__device__ __forceinline__ void derive_v0(Sh &S, const JobConst &jc, double norm) {
  const int c = min((int)threadIdx.x, 2);
  const double nw = fmul(u52_to_double(S.tot[kAccCnt]), norm);
  const double nm = fdiv(fmul(u52_to_double(S.tot[kAccR + c]), norm), nw);
  const double ow = fsub(jc.tw, nw);
  const double om = fdiv(fsub(fmul(jc.tw, jc.tm[c]), fmul(nw, nm)), ow);
  const double a = fsq(om), b = fsq(nm);
  const double a1 = __shfl_sync(0xffffffffu, a, 1), b1 = __shfl_sync(0xffffffffu, b, 1);
  const double a2 = __shfl_sync(0xffffffffu, a, 2), b2 = __shfl_sync(0xffffffffu, b, 2);
  if (threadIdx.x < 3) S.pp.r[c] = fsub(om, nm);
  if (threadIdx.x == 0) {
    double l = fsub(a, b); l = fadd(l, a1); l = fsub(l, b1); l = fadd(l, a2); l = fsub(l, b2);
    S.pp.a = fmul(0.5, l);
  }
}
// variant: totals stay in registers (REDUX broadcasts), one vote decides 1 or 2 limbs
__device__ __forceinline__ void stage2_derive_v1(Sh &S, int rows, const JobConst &jc, double norm) {
  const int lane = threadIdx.x & 31;
  uint64_t x[5], t[5];
#pragma unroll
  for (int w = 0; w < 5; ++w) x[w] = lane < rows ? S.red[lane][w] : 0ull;
  const bool big = __any_sync(0xffffffffu, ((x[0] | x[1] | x[2] | x[3]) >> 27) != 0);
  if (!big) {
#pragma unroll
    for (int w = 0; w < 5; ++w) t[w] = __reduce_add_sync(0xffffffffu, (unsigned)x[w]);
  } else {
#pragma unroll
    for (int w = 0; w < 5; ++w) {
      const unsigned a = __reduce_add_sync(0xffffffffu, (unsigned)(x[w] & 0x7FFFFFFu));
      const unsigned c = __reduce_add_sync(0xffffffffu, (unsigned)(x[w] >> 27));
      t[w] = (uint64_t)a + ((uint64_t)c << 27);
    }
  }
  if (lane < 5) S.tot[lane] = t[lane == 0 ? 0 : lane == 1 ? 1 : lane == 2 ? 2 : lane == 3 ? 3 : 4];
  const int c = min(lane, 2);
  const uint64_t sc = c == 0 ? t[1] : (c == 1 ? t[2] : t[3]);
  const double nw = fmul(u52_to_double(t[0]), norm);
  const double nm = fdiv(fmul(u52_to_double(sc), norm), nw);
  const double ow = fsub(jc.tw, nw);
  const double om = fdiv(fsub(fmul(jc.tw, jc.tm[c]), fmul(nw, nm)), ow);
  const double a = fsq(om), b = fsq(nm);
  const double a1 = __shfl_sync(0xffffffffu, a, 1), b1 = __shfl_sync(0xffffffffu, b, 1);
  const double a2 = __shfl_sync(0xffffffffu, a, 2), b2 = __shfl_sync(0xffffffffu, b, 2);
  if (lane < 3) S.pp.r[c] = fsub(om, nm);
  if (lane == 0) {
    double l = fsub(a, b); l = fadd(l, a1); l = fsub(l, b1); l = fadd(l, a2); l = fsub(l, b2);
    S.pp.a = fmul(0.5, l);
  }
}
template <int V>
__global__ void __launch_bounds__(512, 1) k(int iters, double *sink, long long *cyc) {
  __shared__ Sh S;
  JobConst jc; jc.tw = 0.37; jc.tm[0] = 101.5; jc.tm[1] = 77.25; jc.tm[2] = 140.0;
  const double norm = 1.0 / 8294400.0;
  for (int i = threadIdx.x; i < 32 * 8; i += blockDim.x) S.red[i / 8][i % 8] = 1000 + 37 * i;
  __syncthreads();
  long long t0 = clock64(), acc = 0;
  for (int it = 0; it < iters; ++it) {
    if (threadIdx.x < 32) {
      long long a0 = clock64();
      if (V == 0) { stage2_v0(S, 16); derive_v0(S, jc, norm); }
      if (V == 1) stage2_derive_v1(S, 16, jc, norm);
      acc += clock64() - a0;
      if (threadIdx.x == 0) S.red[it & 15][1] += 3;
    }
    __syncthreads();
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) { cyc[0] = (t1 - t0) / iters; cyc[1] = acc / iters; sink[0] = S.pp.a + S.pp.r[0]; }
}
}  // namespace dq
using dq::k;
int main() {
  double *sink; long long *cyc; CK(cudaMalloc(&sink, 64)); CK(cudaMalloc(&cyc, 64));
  long long h[2];
  k<0><<<148, 512>>>(2000, sink, cyc); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost));
  printf("stage2 + derive, current      : %lld cycles per pass incl. sync (warp 0 alone: %lld)\n", h[0], h[1]);
  k<1><<<148, 512>>>(2000, sink, cyc); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost));
  printf("stage2 + derive, regs + 1 vote: %lld cycles per pass incl. sync (warp 0 alone: %lld)\n", h[0], h[1]);
  return 0;
}
