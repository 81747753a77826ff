// Micro-benchmarks that size the latency budget of the split kernel on B200 (sm_100a).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mb mb.cu ; run: ./mb
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__global__ void k_dfma(double *out, int iters) {
  double a = threadIdx.x * 1e-3, b = 1.0000001, c = 1e-9, d0 = a, d1 = a + 1, d2 = a + 2, d3 = a + 3;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) { d0 = __fma_rn(d0, b, c); d1 = __fma_rn(d1, b, c); d2 = __fma_rn(d2, b, c); d3 = __fma_rn(d3, b, c); }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = d0 + d1 + d2 + d3;
  if (threadIdx.x == 0 && blockIdx.x == 0) printf("DFMA   : %d warps/SM, %.2f cycles per warp-instr per SM (4 chains)\n", blockDim.x / 32, (double)(t1 - t0) / (4.0 * iters * (blockDim.x / 32)));
}
__global__ void k_dsetp(int *out, int iters) {
  double a = threadIdx.x * 1e-3; int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) { double x = (double)i; c0 += (a < x); c1 += (a + 1 < x); c2 += (a + 2 > x); c3 += (a == x); }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = c0 + c1 + c2 + c3;
  if (threadIdx.x == 0 && blockIdx.x == 0) printf("DSETP+ : %d warps/SM, %.2f cycles per loop iteration per warp per SM (4 compares + cvt)\n", blockDim.x / 32, (double)(t1 - t0) / (1.0 * iters * (blockDim.x / 32)));
}
__global__ void k_ffma(float *out, int iters) {
  float a = threadIdx.x * 1e-3f, b = 1.0000001f, c = 1e-9f, d0 = a, d1 = a + 1, d2 = a + 2, d3 = a + 3;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) { d0 = fmaf(d0, b, c); d1 = fmaf(d1, b, c); d2 = fmaf(d2, b, c); d3 = fmaf(d3, b, c); }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = d0 + d1 + d2 + d3;
  if (threadIdx.x == 0 && blockIdx.x == 0) printf("FFMA   : %d warps/SM, %.3f cycles per warp-instr per SM\n", blockDim.x / 32, (double)(t1 - t0) / (4.0 * iters * (blockDim.x / 32)));
}
__global__ void k_imad(int *out, int iters) {
  int a = threadIdx.x, b = 3, d0 = a, d1 = a + 1, d2 = a + 2, d3 = a + 3;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) { d0 = d0 * b + i; d1 = d1 * b + i; d2 = d2 * b + i; d3 = d3 * b + i; }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = d0 + d1 + d2 + d3;
  if (threadIdx.x == 0 && blockIdx.x == 0) printf("IMAD   : %d warps/SM, %.3f cycles per warp-instr per SM\n", blockDim.x / 32, (double)(t1 - t0) / (4.0 * iters * (blockDim.x / 32)));
}
__global__ void k_ddiv(double *out, int iters) {
  double a = 1.0 + threadIdx.x * 1e-3, b = 1.0000001;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) a = __ddiv_rn(a, b);
  long long t1 = clock64();
  out[threadIdx.x] = a;
  if (threadIdx.x == 0) printf("DDIV   : dependent chain latency %.1f cycles (1 warp)\n", (double)(t1 - t0) / iters);
}
__global__ void k_dchain(double *out, int iters) {
  double a = 1.0 + threadIdx.x * 1e-3, b = 1.0000001;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) a = __dmul_rn(a, b);
  long long t1 = clock64();
  out[threadIdx.x] = a;
  if (threadIdx.x == 0) printf("DMUL   : dependent chain latency %.1f cycles (1 warp)\n", (double)(t1 - t0) / iters);
}
__device__ __forceinline__ unsigned ldr32(const unsigned *p) { unsigned v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned long long ldr64(const unsigned long long *p) { unsigned long long v; asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void str64(unsigned long long *p, unsigned long long v) { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory"); }
__global__ void k_barrier(unsigned *counter, int iters, int fenced) {
  unsigned target = 0;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    __syncthreads();
    if (threadIdx.x == 0) {
      target += gridDim.x;
      if (fenced) __threadfence();
      atomicAdd(counter, 1u);
      while (ldr32(counter) < target) {}
      if (fenced) __threadfence();
    }
    __syncthreads();
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) printf("BARRIER: %s grid barrier over %d CTAs: %.0f cycles\n", fenced ? "fenced" : "unfenced", gridDim.x, (double)(t1 - t0) / iters);
}
// all-gather of 8 tagged words per CTA: publish + everybody reads everybody
__global__ void k_exchange(unsigned long long *slots, int iters, unsigned long long *sink) {
  const int G = gridDim.x, tid = threadIdx.x;
  unsigned long long acc = 0;
  long long t0 = clock64();
  for (int it = 1; it <= iters; ++it) {
    unsigned long long *base = slots + (size_t)(it & 1) * G * 8;
    if (tid < 8) str64(base + blockIdx.x * 8 + tid, ((unsigned long long)(it & 0xFFFF) << 48) | (unsigned long long)(blockIdx.x + tid));
    const int w = tid & 7;
    unsigned long long sum = 0;
    for (int i = tid >> 3; i < G; i += blockDim.x / 8) {
      unsigned long long v;
      do { v = ldr64(base + i * 8 + w); } while ((unsigned)(v >> 48) != (unsigned)(it & 0xFFFF));
      sum += v & 0xFFFFFFFFFFFFull;
    }
    acc += sum;
    __syncthreads();
  }
  long long t1 = clock64();
  sink[blockIdx.x * blockDim.x + tid] = acc;
  if (tid == 0 && blockIdx.x == 0) printf("EXCHANG: tagged-slot all-gather over %d CTAs (+1 syncthreads): %.0f cycles per round\n", G, (double)(t1 - t0) / iters);
}
__global__ void k_sync(int iters) {
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) printf("SYNCTHR: %d threads: %.0f cycles\n", blockDim.x, (double)(t1 - t0) / iters);
}
__global__ void k_shfl64(unsigned long long *out, int iters) {
  unsigned long long v = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = v;
  if (threadIdx.x == 0 && blockIdx.x == 0) printf("WSUM64 : warp_sum_u64, %d warps/SM: %.0f cycles per reduction (all warps concurrently)\n", blockDim.x / 32, (double)(t1 - t0) / iters);
}
__global__ void k_redux(unsigned *out, int iters) {
  unsigned v = threadIdx.x, acc = 0;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    acc += __reduce_add_sync(0xffffffffu, v + i);
    acc += __reduce_add_sync(0xffffffffu, v ^ i);
    acc += __reduce_add_sync(0xffffffffu, v * 3 + i);
    acc += __reduce_add_sync(0xffffffffu, v + 2 * i);
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) printf("REDUX  : %d warps/SM: %.2f cycles per warp-REDUX per SM\n", blockDim.x / 32, (double)(t1 - t0) / (4.0 * iters * (blockDim.x / 32)));
}
__global__ void k_redux_lat(unsigned *out, int iters) {
  unsigned v = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) v = __reduce_add_sync(0xffffffffu, v) + threadIdx.x;
  long long t1 = clock64();
  out[threadIdx.x] = v;
  if (threadIdx.x == 0) printf("REDUX  : dependent latency %.1f cycles\n", (double)(t1 - t0) / iters);
}
__global__ void k_lds_bcast(double *out, int iters) {
  __shared__ double tab[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) tab[i] = i * 0.37;
  __syncthreads();
  double t = threadIdx.x * 0.5; int above = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it)
    for (int m = 0; m < 1024; ++m) above += (tab[m] > t);
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = above;
  if (threadIdx.x == 0 && blockIdx.x == 0) printf("RANKCNT: %d warps: %.2f cycles per (LDS.64 bcast + DSETP + add) per warp\n", blockDim.x / 32, (double)(t1 - t0) / (1024.0 * iters * (blockDim.x / 32)));
}
int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int G = prop.multiProcessorCount;
  printf("%s, %d SMs, clock %d kHz\n", prop.name, G, prop.clockRate);
  void *buf; CK(cudaMalloc(&buf, 64 << 20)); CK(cudaMemset(buf, 0, 64 << 20));
  for (int w : {128, 512, 1024}) { k_dfma<<<G, w>>>((double *)buf, 20000); CK(cudaDeviceSynchronize()); }
  for (int w : {128, 1024}) { k_dsetp<<<G, w>>>((int *)buf, 20000); CK(cudaDeviceSynchronize()); }
  for (int w : {256, 1024}) { k_ffma<<<G, w>>>((float *)buf, 20000); CK(cudaDeviceSynchronize()); }
  for (int w : {256, 1024}) { k_imad<<<G, w>>>((int *)buf, 20000); CK(cudaDeviceSynchronize()); }
  for (int w : {256, 1024}) { k_redux<<<G, w>>>((unsigned *)buf, 20000); CK(cudaDeviceSynchronize()); }
  k_redux_lat<<<1, 32>>>((unsigned *)buf, 2000); CK(cudaDeviceSynchronize());
  for (int w : {512, 1024}) { k_lds_bcast<<<1, w>>>((double *)buf, 4); CK(cudaDeviceSynchronize()); }
  k_ddiv<<<1, 32>>>((double *)buf, 2000); CK(cudaDeviceSynchronize());
  k_dchain<<<1, 32>>>((double *)buf, 2000); CK(cudaDeviceSynchronize());
  k_sync<<<1, 1024>>>(2000); CK(cudaDeviceSynchronize());
  k_shfl64<<<G, 1024>>>((unsigned long long *)buf, 2000); CK(cudaDeviceSynchronize());
  for (int fenced = 0; fenced < 2; ++fenced) {
    CK(cudaMemset(buf, 0, 4096));
    unsigned *counter = (unsigned *)buf; int iters = 2000;
    void *args[] = {&counter, &iters, &fenced};
    CK(cudaLaunchCooperativeKernel((void *)k_barrier, dim3(G), dim3(1024), args, 0, 0)); CK(cudaDeviceSynchronize());
  }
  {
    CK(cudaMemset(buf, 0, 1 << 20));
    unsigned long long *slots = (unsigned long long *)buf, *sink = slots + (1 << 17); int iters = 2000;
    void *args[] = {&slots, &iters, &sink};
    CK(cudaLaunchCooperativeKernel((void *)k_exchange, dim3(G), dim3(1024), args, 0, 0)); CK(cudaDeviceSynchronize());
  }
  printf("done\n");
  return 0;
}
