// dq_split_exact.cu -- stand-alone launch of the small-input split (dq_split_exact.cuh) for the callers whose split
// kernel does not carry it (K > 512: the generic kernel of dq_split.cu).  The latency-optimised kernel (dq_split2.cu)
// runs the same body itself, without extra launches.
#include "dq_split_exact.cuh"
#include <atomic>
#include <mutex>

#include <algorithm>

namespace dq {
namespace {

constexpr int kExactThreads = 256;

__global__ void __launch_bounds__(256) exact_first_seen_kernel(const ExactSampling q, const uint32_t *ucount, uint32_t max_points,
                                                              uint32_t *first_seen) {
  if (*ucount > max_points) return;
  exact_first_seen(q, first_seen, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}

__global__ void __launch_bounds__(kExactThreads) split_exact_kernel(const SplitArgs A, unsigned char *scratch, const uint32_t *uniq,
                                                                    uint32_t *table, const uint32_t *first_seen,
                                                                    double *g_weight, double *g_tse, double *g_mean,
                                                                    double *g_var, int32_t *g_size) {
  extern __shared__ __align__(16) unsigned char exact_smem[];
  const uint32_t U = *A.num_points_dev;
  if (U > A.exact_small_max || U == 0) return;
  exact::split_exact_body<kExactThreads>(A, (int)U, exact_smem, scratch, uniq, table, first_seen, g_weight, g_tse, g_mean, g_var,
                                         g_size);
}

}  // namespace

size_t split_exact_smem_bytes() { return sizeof(exact::Shared); }
size_t split_exact_scratch_bytes() { return exact::kScratchBytes; }

ExactSampling exact_sampling(const uint32_t *d_in, uint32_t num_rows, uint32_t num_cols, uint32_t dec, int bits) {
  ExactSampling q;
  const uint32_t nr = (num_rows + dec - 1) / dec, nc = (num_cols + dec - 1) / dec;
  const uint32_t shift = 8u - (uint32_t)bits;
  const uint32_t byte_mask = (0xFFu >> shift) << shift;
  q.in = d_in;
  q.samples_per_row = nc;
  q.num_samples = nr * nc;
  q.num_rows = num_rows;
  q.dec = dec;
  q.word_mask = (byte_mask << 16) | (byte_mask << 8) | byte_mask;
  q.shift = shift;
  return q;
}

void first_seen_launch(const ExactSampling &q, uint32_t *d_first_seen, cudaStream_t st) {
  static const uint32_t kNoLimit = 0xFFFFFFFFu;
  static uint32_t *d_zero = nullptr;  // a device word holding 0: "U <= limit" is always true
  static std::atomic<bool> ready{false};
  if (!ready) {
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    if (!d_zero) {
      DQ_CUDA_CHECK(cudaMalloc(&d_zero, sizeof(uint32_t)));
      DQ_CUDA_CHECK(cudaMemset(d_zero, 0, sizeof(uint32_t)));
    }
    ready = true;
  }
  const unsigned blocks = (unsigned)std::min<uint64_t>(((uint64_t)q.num_samples + 255) / 256, 1184u);
  exact_first_seen_kernel<<<std::max(blocks, 1u), 256, 0, st>>>(q, d_zero, kNoLimit, d_first_seen);
  DQ_CUDA_CHECK(cudaGetLastError());
}

void split_exact_launch(const SplitArgs &args, const ExactSampling &q, unsigned char *d_scratch, const uint32_t *d_uniq,
                        uint32_t *d_table, uint32_t *d_first_seen, double *g_f64, int32_t *g_i32, cudaStream_t st) {
  const unsigned blocks = (unsigned)std::min<uint64_t>(((uint64_t)q.num_samples + 255) / 256, 64u);
  exact_first_seen_kernel<<<std::max(blocks, 1u), 256, 0, st>>>(q, args.num_points_dev, args.exact_small_max, d_first_seen);
  DQ_RAISE_SMEM(split_exact_kernel, sizeof(exact::Shared));
  const size_t K = args.num_colors;
  split_exact_kernel<<<1, kExactThreads, sizeof(exact::Shared), st>>>(args, d_scratch, d_uniq, d_table, d_first_seen, g_f64, g_f64 + K,
                                                                           g_f64 + 2 * K, g_f64 + 5 * K, g_i32);
  DQ_CUDA_CHECK(cudaGetLastError());
}

}  // namespace dq
