// dq_map.cu -- nearest-palette remap.
//
// Reference: map_colors_mps, DivQuant/DivQuantMapColors.cpp:243-539.  The reference sorts the palette
// by r+g+b, starts at lut_init[r+g+b] and walks up/down alternately, pruning with (delta sum)^2/3.
// Its result is exactly  argmin_k (dist(k), rank(k))  over the WHOLE sorted palette, with
// rank(s)=0, rank(s+d)=2d-1, rank(s-d)=2d (re-verified against the compiled reference
// in tests/test_oracle_vs_ref.py), because the prune bound never exceeds the true distance and a
// candidate only replaces the incumbent when strictly closer.
//
// The sorted palette and lut_init come from the host shim, which calls the same std::sort as the
// reference (the order of equal sums is observable).
//
// Two device formulations, both bit-exact:
//   * brute force over pixels (map_pixels_*): N*K distance evaluations.
//   * de-duplicated (map_unique + map_gather): the K evaluations are done once per UNIQUE colour and
//     the answer is parked in the 2^24-entry direct table; pixels then gather through it.  Used
//     whenever a histogram of the same pixels is at hand (quant_recurse always has one).
#include "dq_kernels.cuh"

namespace dq {
namespace {

constexpr int kMapThreads = 256;
constexpr uint32_t kLutEntries = 3 * 255 + 1;

// Winner in the sorted palette for one colour.  `pal` holds {r, g, b, packed 0x00RRGGBB}.
__device__ __forceinline__ uint32_t nearest_entry(uint32_t colour, const int4 *pal, int num_colors, const int *lut) {
  const int r = (colour >> 16) & 0xFF, g = (colour >> 8) & 0xFF, b = colour & 0xFF;
  const int s = lut[r + g + b];
  // lexicographic (dist, rank): dist < 2^18, rank < 2K
  uint64_t best = ~uint64_t(0);
  int win = s;
#pragma unroll 4
  for (int k = 0; k < num_colors; ++k) {
    const int4 e = pal[k];
    const int dr = r - e.x, dg = g - e.y, db = b - e.z;
    const uint32_t dist = (uint32_t)(dr * dr + dg * dg + db * db);
    const uint32_t rank = (k > s) ? (uint32_t)(2 * (k - s) - 1) : (uint32_t)(2 * (s - k));
    const uint64_t key = ((uint64_t)dist << 32) | rank;
    if (key < best) {
      best = key;
      win = k;
    }
  }
  return (uint32_t)pal[win].w;
}

__device__ __forceinline__ void stage_palette(const uint32_t *sorted, int num_colors, const int *lut_init, int4 *s_pal,
                                              int *s_lut) {
  for (int k = threadIdx.x; k < num_colors; k += blockDim.x) {
    const uint32_t c = sorted[k] & 0x00FFFFFFu;
    s_pal[k] = make_int4((c >> 16) & 0xFF, (c >> 8) & 0xFF, c & 0xFF, (int)c);
  }
  for (int i = threadIdx.x; i < (int)kLutEntries; i += blockDim.x) s_lut[i] = lut_init[i];
  __syncthreads();
}

// Brute force over pixels, palette staged in shared memory.
__global__ void __launch_bounds__(kMapThreads) map_pixels_kernel(const uint32_t *__restrict__ in, uint32_t n,
                                                                uint32_t *__restrict__ out, const uint32_t *sorted,
                                                                int num_colors, const int *lut_init) {
  extern __shared__ __align__(16) unsigned char smem[];
  int *s_lut = reinterpret_cast<int *>(smem);
  int4 *s_pal = reinterpret_cast<int4 *>(smem + ((kLutEntries * 4 + 15) & ~15u));
  stage_palette(sorted, num_colors, lut_init, s_pal, s_lut);
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    out[i] = nearest_entry(in[i], s_pal, num_colors, s_lut);
}

// One evaluation per unique colour; the answer goes to table[colour] with bit 31 set (so that a
// mapped colour of 0x000000 is distinguishable from an untouched entry while debugging).
__global__ void __launch_bounds__(kMapThreads) map_unique_kernel(const uint32_t *__restrict__ uniq, const uint32_t *ucount,
                                                                uint32_t *table, const uint32_t *sorted, int num_colors,
                                                                const int *lut_init) {
  extern __shared__ __align__(16) unsigned char smem[];
  int *s_lut = reinterpret_cast<int *>(smem);
  int4 *s_pal = reinterpret_cast<int4 *>(smem + ((kLutEntries * 4 + 15) & ~15u));
  stage_palette(sorted, num_colors, lut_init, s_pal, s_lut);
  const uint32_t u = *ucount;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < u; i += gridDim.x * blockDim.x) {
    const uint32_t c = uniq[i];
    table[c] = 0x80000000u | nearest_entry(c, s_pal, num_colors, s_lut);
  }
}

// Same, palette read from global memory (palettes too large for shared memory).
__global__ void __launch_bounds__(kMapThreads) map_pixels_big_kernel(const uint32_t *__restrict__ in, uint32_t n,
                                                                    uint32_t *__restrict__ out, const int4 *pal,
                                                                    int num_colors, const int *lut_init) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    out[i] = nearest_entry(in[i], pal, num_colors, lut_init);
}

__global__ void __launch_bounds__(kMapThreads) expand_palette_kernel(const uint32_t *sorted, int num_colors, int4 *pal) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < num_colors; k += gridDim.x * blockDim.x) {
    const uint32_t c = sorted[k] & 0x00FFFFFFu;
    pal[k] = make_int4((c >> 16) & 0xFF, (c >> 8) & 0xFF, c & 0xFF, (int)c);
  }
}

__global__ void __launch_bounds__(kMapThreads) map_gather_kernel(const uint32_t *__restrict__ in, uint32_t n,
                                                                uint32_t *__restrict__ out, const uint32_t *table,
                                                                uint32_t word_mask, uint32_t shift) {
  const uint32_t nvec = n >> 2;
  const uint4 *in4 = reinterpret_cast<const uint4 *>(in);
  uint4 *out4 = reinterpret_cast<uint4 *>(out);
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += gridDim.x * blockDim.x) {
    const uint4 p = __ldcs(in4 + i);
    uint4 q;
    q.x = table[(p.x & word_mask) >> shift] & 0x00FFFFFFu;
    q.y = table[(p.y & word_mask) >> shift] & 0x00FFFFFFu;
    q.z = table[(p.z & word_mask) >> shift] & 0x00FFFFFFu;
    q.w = table[(p.w & word_mask) >> shift] & 0x00FFFFFFu;
    __stcs(out4 + i, q);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3u)) {
    const uint32_t i = (nvec << 2) + threadIdx.x;
    out[i] = table[(in[i] & word_mask) >> shift] & 0x00FFFFFFu;
  }
}

__global__ void __launch_bounds__(kMapThreads) map_gather_scalar_kernel(const uint32_t *__restrict__ in, uint32_t n,
                                                                       uint32_t *__restrict__ out, const uint32_t *table) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    out[i] = table[in[i] & 0x00FFFFFFu] & 0x00FFFFFFu;
}

inline int blocks_for(uint64_t items, int threads, int sm_count, int per_sm) {
  uint64_t want = (items + threads - 1) / threads;
  uint64_t cap = (uint64_t)sm_count * per_sm;
  if (want < 1) want = 1;
  return (int)(want < cap ? want : cap);
}

inline size_t map_smem_bytes(int num_colors) { return ((kLutEntries * 4 + 15) & ~15u) + (size_t)num_colors * sizeof(int4); }

}  // namespace

int map_smem_palette_limit() { return 8192; }  // 128 KB of int4 entries + the LUT fit the 227 KB carve-out

void map_pixels(const uint32_t *d_in, uint32_t n, uint32_t *d_out, const uint32_t *d_sorted, int num_colors,
                const int *d_lut, int4 *d_pal_scratch, int sm_count, cudaStream_t st) {
  if (n == 0) return;
  if (num_colors <= map_smem_palette_limit()) {
    const size_t smem = map_smem_bytes(num_colors);
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
      DQ_CUDA_CHECK(cudaFuncSetAttribute(map_pixels_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      configured = smem;
    }
    map_pixels_kernel<<<blocks_for(n, kMapThreads, sm_count, 8), kMapThreads, smem, st>>>(d_in, n, d_out, d_sorted,
                                                                                         num_colors, d_lut);
  } else {
    expand_palette_kernel<<<blocks_for(num_colors, kMapThreads, sm_count, 1), kMapThreads, 0, st>>>(d_sorted, num_colors,
                                                                                                   d_pal_scratch);
    map_pixels_big_kernel<<<blocks_for(n, kMapThreads, sm_count, 8), kMapThreads, 0, st>>>(d_in, n, d_out, d_pal_scratch,
                                                                                          num_colors, d_lut);
  }
  DQ_CUDA_CHECK(cudaGetLastError());
}

void map_unique(const uint32_t *d_uniq, const uint32_t *d_ucount, uint32_t u_hint, uint32_t *d_table,
                const uint32_t *d_sorted, int num_colors, const int *d_lut, int sm_count, cudaStream_t st) {
  const size_t smem = map_smem_bytes(num_colors);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    DQ_CUDA_CHECK(cudaFuncSetAttribute(map_unique_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  map_unique_kernel<<<blocks_for(u_hint, kMapThreads, sm_count, 8), kMapThreads, smem, st>>>(d_uniq, d_ucount, d_table,
                                                                                            d_sorted, num_colors, d_lut);
  DQ_CUDA_CHECK(cudaGetLastError());
}

void map_gather(const uint32_t *d_in, uint32_t n, uint32_t *d_out, const uint32_t *d_table, int sm_count, cudaStream_t st) {
  if (n == 0) return;
  const bool aligned = ((reinterpret_cast<uintptr_t>(d_in) | reinterpret_cast<uintptr_t>(d_out)) & 15u) == 0;
  if (aligned)
    map_gather_kernel<<<blocks_for((n >> 2) + 1, kMapThreads, sm_count, 8), kMapThreads, 0, st>>>(d_in, n, d_out, d_table,
                                                                                                 0x00FFFFFFu, 0u);
  else
    map_gather_scalar_kernel<<<blocks_for(n, kMapThreads, sm_count, 8), kMapThreads, 0, st>>>(d_in, n, d_out, d_table);
  DQ_CUDA_CHECK(cudaGetLastError());
}

}  // namespace dq
