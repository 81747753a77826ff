// dq_split_exact.cu -- the divisive phase for SMALL inputs, in the reference's own summation order.
//
// Reference: DivQuantCluster<UW=false,...>, DivQuant/DivQuantCluster.cpp:133-1097 (weighted path) and
// DivQuantClusterInitMeanAndVar, :49-123.
//
// Why this exists.  In the weighted path the reference accumulates  weight_i * channel_i  as doubles, one point
// after the other, in the order calc_color_table emitted the unique colours (hash bucket ascending, most
// recently first-seen colour first inside a bucket, MapColors.cpp:157-198).  The large-input kernels
// (dq_split2.cu) use exact integer sums instead, which differ from those doubles by ~1e-16 relative.  That
// never matters unless a decision sits exactly on a tie -- a mean on an integer, two equal variances or TSEs --
// which is what small or synthetic inputs (few colours, symmetric counts) produce.  There the reference decides
// by its own rounding noise, so to return its palette bit for bit the noise has to be reproduced: for
// U <= kExactMaxPoints this kernel walks the reference's loop with the same sequential double accumulations
// (one lane per accumulator chain, non-contracted IEEE operations in the reference's order).
//
// One CTA.  Launched for every weighted call; returns at once when U is larger (the split kernels return at
// once when it is not), so no host round trip is needed to choose.
#include "dq_split_math.cuh"

#include <algorithm>
#include <cfloat>

namespace dq {
namespace {

constexpr int kExactThreads = 256;
constexpr int kExactSortCap = 4096;  // power of two >= kExactMaxPoints
constexpr int kExactIndexBits = 12;
static_assert(kExactSortCap >= (int)kExactMaxPoints, "sort capacity");

struct ExactShared {
  unsigned long long keys[kExactSortCap];
  double w[kExactMaxPoints];        // weights[] of calc_color_table (:172,185)
  uint32_t colour[kExactMaxPoints];
  uint32_t member[kExactMaxPoints];
  uint16_t cur[kExactMaxPoints];    // points of the cluster being split, ascending original order (:929-1019)
  uint16_t sel[kExactMaxPoints];    // this pass: the points of cur on the new side, same order
  double chain[8];
  int32_t warp_tmp[kExactThreads / 32];
  double red_val[kExactThreads / 32];
  int32_t red_idx[kExactThreads / 32];
  int32_t cur_n, old_index, new_size, scan_carry;
  // scalars of the current split (thread 0 writes, everybody reads after a barrier)
  double tw, tm[3], tv[3], nw, ow, nm[3], om[3], nv[3];
  double lhs, rr[3], cut;
  int32_t axis;
};

__device__ __forceinline__ double chan(uint32_t p, int c) { return byte_to_double((p >> (16 - 8 * c)) & 0xFFu); }
__device__ __forceinline__ double chan_sq(uint32_t p, int c) {
  const uint32_t v = (p >> (16 - 8 * c)) & 0xFFu;
  return u52_to_double((uint64_t)(v * v));  // the reference squares in int, then converts (:98-100, :741-743)
}

// Sequential accumulation chains over list[0..n): lane l < nchains owns chain l.
//   chains 0..2: sum w*c      3: sum w      4..6: sum w*(c*c)
// Every chain adds in ascending list order, exactly like the reference's single loop over the cluster's points
// (which skips the points of the other side).  list == nullptr: the identity.
__device__ __forceinline__ void run_chains(ExactShared &S, const uint16_t *list, int n, int nchains) {
  const int lane = threadIdx.x;
  if (lane < nchains) {
    double acc = 0.0;
#pragma unroll 4
    for (int j = 0; j < n; ++j) {
      const int idx = list ? (int)list[j] : j;
      const double wt = S.w[idx];
      const uint32_t p = S.colour[idx];
      double term;
      if (lane < 3) term = fmul(wt, chan(p, lane));
      else if (lane == 3) term = wt;
      else term = fmul(wt, chan_sq(p, lane - 4));
      acc = fadd(acc, term);
    }
    S.chain[lane] = acc;
  }
}

// Classifies the points of cur with `pred` (true = new side) and leaves the new side, in order, in S.sel;
// S.new_size = how many.  Thread t owns the contiguous piece [t*per, (t+1)*per) so that the order survives.
template <typename Pred, typename Each>
__device__ __forceinline__ void select_new(ExactShared &S, int cur_n, Pred pred, Each each) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int per = (cur_n + kExactThreads - 1) / kExactThreads;  // <= 16
  const int lo = min(tid * per, cur_n), hi = min(lo + per, cur_n);
  unsigned mask = 0;
  for (int j = lo; j < hi; ++j) {
    const int idx = S.cur[j];
    const bool is_new = pred(S.colour[idx]);
    each(idx, is_new);
    mask |= (unsigned)is_new << (j - lo);
  }
  const int mine = __popc(mask);
  int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) S.warp_tmp[warp] = incl;
  __syncthreads();
  int pos = incl - mine;
  for (int q = 0; q < warp; ++q) pos += S.warp_tmp[q];
  for (int j = lo; j < hi; ++j)
    if ((mask >> (j - lo)) & 1u) S.sel[pos++] = S.cur[j];
  if (tid == kExactThreads - 1) S.new_size = pos;
  __syncthreads();
}

}  // namespace

// first_seen[c] = 0xFFFFFFFF for the unique colours of a small input (no-op for large ones)
__global__ void __launch_bounds__(256) exact_prepare_kernel(const uint32_t *__restrict__ uniq, const uint32_t *ucount,
                                                           uint32_t *first_seen) {
  const uint32_t u = *ucount;
  if (u > kExactMaxPoints) return;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < u; i += gridDim.x * blockDim.x) first_seen[uniq[i]] = 0xFFFFFFFFu;
}

// first_seen[c] = smallest sample index holding colour c (sampling and bit cutting as in hist_insert)
__global__ void __launch_bounds__(256) exact_first_seen_kernel(const uint32_t *__restrict__ in, const uint32_t *ucount,
                                                              uint32_t samples_per_row, uint32_t num_samples,
                                                              uint32_t num_rows, uint32_t dec, uint32_t word_mask,
                                                              uint32_t shift, uint32_t *first_seen) {
  if (*ucount > kExactMaxPoints) return;
  for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < num_samples; s += gridDim.x * blockDim.x) {
    const uint32_t ir = (s / samples_per_row) * dec, ic = (s % samples_per_row) * dec;
    // the reference's addressing, bug included (MapColors.cpp:120-122); bits cut as in hist_insert
    const uint32_t c = (in[ic + ir * num_rows] & word_mask) >> shift;
    if (s < first_seen[c]) atomicMin(first_seen + c, s);  // the plain read only skips atomics that cannot win
  }
}

__global__ void __launch_bounds__(kExactThreads) split_exact_kernel(const SplitArgs A, const uint32_t *uniq, uint32_t *table,
                                                                    const uint32_t *first_seen, double *g_weight,
                                                                    double *g_tse, double *g_mean, double *g_var,
                                                                    int32_t *g_size) {
  extern __shared__ __align__(16) unsigned char exact_smem[];
  ExactShared &S = *reinterpret_cast<ExactShared *>(exact_smem);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int U = (int)*A.num_points_dev;
  if (U > (int)kExactMaxPoints || U == 0) return;
  const int K = (int)A.num_colors;
  const int last_it = A.max_iters - 1;

  // ---- points in calc_color_table's emission order: (bucket asc, first seen desc) ----
  for (int i = tid; i < kExactSortCap; i += kExactThreads) {
    unsigned long long key = ~0ull;
    if (i < U) {
      const uint32_t c = uniq[i];
      const long R = (c >> 16) & 0xFF, G = (c >> 8) & 0xFF, B = c & 0xFF;
      const unsigned long long bucket = (unsigned long long)(((R * 33023 + G * 30013 + B * 27011) & 0x7fffffff) % 20023);
      key = (bucket << (32 + kExactIndexBits)) | ((unsigned long long)(0xFFFFFFFFu - first_seen[c]) << kExactIndexBits) | (unsigned long long)i;
    }
    S.keys[i] = key;
  }
  __syncthreads();
  for (int k = 2; k <= kExactSortCap; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < kExactSortCap; i += kExactThreads) {
        const int partner = i ^ j;
        if (partner > i) {
          const unsigned long long a = S.keys[i], b = S.keys[partner];
          const bool up = (i & k) == 0;
          if ((a > b) == up) {
            S.keys[i] = b;
            S.keys[partner] = a;
          }
        }
      }
      __syncthreads();
    }
  }
  for (int i = tid; i < U; i += kExactThreads) {
    const uint32_t c = uniq[(int)(S.keys[i] & ((1ull << kExactIndexBits) - 1ull))];
    const uint32_t count = table[c];
    table[c] = 0u;  // the count table is all-zero again when the call ends
    S.colour[i] = c;
    S.w[i] = fmul(A.norm, (double)(int)count);  // weights[i] = weight * count (:185)
    S.member[i] = 0u;
    S.cur[i] = (uint16_t)i;
    A.pts[0][i] = make_uint2(c, count);
  }
  for (int i = tid; i < K; i += kExactThreads) {  // `new T[n]()` of the reference (:296-324)
    g_weight[i] = 0.0;
    g_tse[i] = 0.0;
    g_size[i] = 0;
    for (int c = 0; c < 3; ++c) g_mean[3 * i + c] = 0.0, g_var[3 * i + c] = 0.0;
  }
  __syncthreads();

  // ---- DivQuantClusterInitMeanAndVar (:60-104): chains 0..2 mean, 4..6 second moments ----
  run_chains(S, nullptr, U, 7);
  __syncthreads();
  if (tid == 0) {
    for (int c = 0; c < 3; ++c) {
      S.tm[c] = S.chain[c];
      S.tv[c] = fsub(S.chain[4 + c], fsq(S.chain[c]));
    }
    g_weight[0] = 1.0;
    g_size[0] = U;
    S.old_index = 0;
    S.cur_n = U;
  }
  __syncthreads();

  for (int new_index = 1; new_index < K; ++new_index) {
    const int old_index = S.old_index, cur_n = S.cur_n;
    if (tid == 0) {
      S.tw = g_weight[old_index];
      if (new_index > 1) {
        for (int c = 0; c < 3; ++c) S.tm[c] = g_mean[3 * old_index + c], S.tv[c] = g_var[3 * old_index + c];
      }
      int axis;
      double cut;
      choose_cut(S.tv, S.tm, axis, cut);
      S.axis = axis;
      S.cut = cut;
    }
    __syncthreads();
    // ---- split pass (:438-559) ----
    {
      const int axis = S.axis;
      const double cut = S.cut;
      select_new(S, cur_n, [&](uint32_t p) { return cut < chan(p, axis); }, [](int, bool) {});
    }
    run_chains(S, S.sel, S.new_size, 4);
    __syncthreads();
    if (tid == 0) {
      S.nw = S.chain[3];
      S.ow = fsub(S.tw, S.nw);
      for (int c = 0; c < 3; ++c) {
        S.nm[c] = fdiv(S.chain[c], S.nw);
        S.om[c] = fdiv(fsub(fmul(S.tw, S.tm[c]), fmul(S.nw, S.nm[c])), S.ow);  // combined mean (:579-581)
      }
    }
    __syncthreads();
    // ---- local 2-means (:613-811) ----
    for (int it = 0; it <= last_it; ++it) {
      if (tid == 0) {
        double l = fsub(fsq(S.om[0]), fsq(S.nm[0]));
        l = fadd(l, fsq(S.om[1]));
        l = fsub(l, fsq(S.nm[1]));
        l = fadd(l, fsq(S.om[2]));
        l = fsub(l, fsq(S.nm[2]));
        S.lhs = fmul(0.5, l);
        for (int c = 0; c < 3; ++c) S.rr[c] = fsub(S.om[c], S.nm[c]);
      }
      __syncthreads();
      {
        const double lhs = S.lhs, r0 = S.rr[0], r1 = S.rr[1], r2 = S.rr[2];
        const bool last = (it == last_it);
        select_new(
            S, cur_n,
            [&](uint32_t p) {
              const double dot = fadd(fadd(fmul(r0, chan(p, 0)), fmul(r1, chan(p, 1))), fmul(r2, chan(p, 2)));
              return !(lhs < dot);  // (:683)
            },
            [&](int idx, bool is_new) {
              if (last) S.member[idx] = is_new ? (uint32_t)new_index : (uint32_t)old_index;
            });
      }
      run_chains(S, S.sel, S.new_size, it == last_it ? 7 : 4);
      __syncthreads();
      if (tid == 0) {
        S.nw = S.chain[3];
        for (int c = 0; c < 3; ++c) {
          S.nm[c] = fdiv(S.chain[c], S.nw);
          if (it == last_it) S.nv[c] = S.chain[4 + c];
        }
        S.ow = fsub(S.tw, S.nw);
        for (int c = 0; c < 3; ++c) S.om[c] = fdiv(fsub(fmul(S.tw, S.tm[c]), fmul(S.nw, S.nm[c])), S.ow);
      }
      __syncthreads();
    }

    if (tid == 0) {
      g_size[old_index] = cur_n - S.new_size;
      g_size[new_index] = S.new_size;
      for (int c = 0; c < 3; ++c) g_mean[3 * new_index + c] = S.nm[c], g_mean[3 * old_index + c] = S.om[c];
      SplitRecord r;
      if (A.records != nullptr) {
        r.new_index = new_index;
        r.old_index = old_index;
        r.cut_axis = S.axis;
        r.num_points = cur_n;
        r.new_size = S.new_size;
        r.is_last = (new_index == K - 1);
        r.cut_pos = S.cut;
        r.total_weight = S.tw;
        r.new_weight = S.nw;
        r.old_weight = S.ow;
        for (int c = 0; c < 3; ++c) {
          r.new_mean[c] = S.nm[c], r.old_mean[c] = S.om[c];
          r.new_var[c] = r.old_var[c] = 0.0;
        }
        r.new_tse = r.old_tse = 0.0;
      }
      if (new_index < K - 1) {  // the last split leaves without touching var / weight / tse (:823-832)
        double nv[3], ov[3];
        for (int c = 0; c < 3; ++c) {
          nv[c] = fsub(fdiv(S.nv[c], S.nw), fsq(S.nm[c]));  // (:836-838)
          ov[c] = fsub(fdiv(fsub(fmul(S.tw, S.tv[c]), fmul(S.nw, fadd(nv[c], fsq(fsub(S.nm[c], S.tm[c]))))), S.ow),
                       fsq(fsub(S.om[c], S.tm[c])));          // combined variance (:844-855)
          g_var[3 * new_index + c] = nv[c];
          g_var[3 * old_index + c] = ov[c];
        }
        g_weight[old_index] = S.ow;
        g_weight[new_index] = S.nw;
        g_tse[old_index] = fmul(S.ow, fadd(fadd(ov[0], ov[1]), ov[2]));  // (:871)
        g_tse[new_index] = fmul(S.nw, fadd(fadd(nv[0], nv[1]), nv[2]));
        if (A.records != nullptr) {
          for (int c = 0; c < 3; ++c) r.new_var[c] = nv[c], r.old_var[c] = ov[c];
          r.new_tse = g_tse[new_index];
          r.old_tse = g_tse[old_index];
        }
      }
      if (A.records != nullptr) A.records[new_index - 1] = r;
    }
    __syncthreads();
    if (new_index == K - 1) break;

    // ---- next cluster: strictly-greater scan seeded with DBL_MIN; stale old_index otherwise (:876-887) ----
    {
      double best = DBL_MIN;
      int best_i = -1;
      for (int ic = tid; ic <= new_index; ic += kExactThreads) {
        const double t = g_tse[ic];
        if (best < t) best = t, best_i = ic;
      }
      for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
        if (oi >= 0 && (best_i < 0 || best < ob || (ob == best && oi < best_i))) best = ob, best_i = oi;
      }
      if (lane == 0) S.red_val[warp] = best, S.red_idx[warp] = best_i;
      __syncthreads();
      if (tid == 0) {
        for (int q = 1; q < kExactThreads / 32; ++q) {
          const double ob = S.red_val[q];
          const int oi = S.red_idx[q];
          if (oi >= 0 && (best_i < 0 || best < ob || (ob == best && oi < best_i))) best = ob, best_i = oi;
        }
        if (best_i >= 0) S.old_index = best_i;
        S.scan_carry = 0;
      }
      __syncthreads();
    }
    // ---- gather its points in ascending original order (:929-1019) ----
    {
      const uint32_t want = (uint32_t)S.old_index;
      for (int base = 0; base < U; base += kExactThreads) {
        const int i = base + tid;
        const bool hit = i < U && S.member[i] == want;
        const unsigned ballot = __ballot_sync(0xffffffffu, hit);
        if (lane == 0) S.warp_tmp[warp] = __popc(ballot);
        __syncthreads();
        int before = S.scan_carry;
        for (int q = 0; q < warp; ++q) before += S.warp_tmp[q];
        if (hit) S.cur[before + __popc(ballot & ((1u << lane) - 1u))] = (uint16_t)i;
        __syncthreads();
        if (tid == 0) {
          int tot = 0;
          for (int q = 0; q < kExactThreads / 32; ++q) tot += S.warp_tmp[q];
          S.scan_carry += tot;
        }
        __syncthreads();
      }
      if (tid == 0) {
        S.cur_n = S.scan_carry;
        if (S.cur_n != g_size[S.old_index]) A.ctl[kCtlError] = 7;  // "Cluster to be split is expected to be of size ..." (:1013)
      }
      __syncthreads();
    }
  }

  // ---- palette = rounded means of the non-empty clusters in index order (:1030-1065) ----
  if (tid == 0) S.scan_carry = 0;
  __syncthreads();
  for (int base = 0; base < K; base += kExactThreads) {
    const int ic = base + tid;
    uint32_t colour = 0;
    int size = 0;
    if (ic < K) {
      size = (K > 1) ? g_size[ic] : U;
      double mean[3] = {0.0, 0.0, 0.0};  // K == 1 never assigns mean[0]
      if (K > 1) mean[0] = g_mean[3 * ic], mean[1] = g_mean[3 * ic + 1], mean[2] = g_mean[3 * ic + 2];
      if (size > 0) {
        const uint32_t Rr = (__double2uint_rz(fadd(mean[0], 0.5)) & 0xFFu) << A.shift;
        const uint32_t Gg = (__double2uint_rz(fadd(mean[1], 0.5)) & 0xFFu) << A.shift;
        const uint32_t Bb = (__double2uint_rz(fadd(mean[2], 0.5)) & 0xFFu) << A.shift;
        colour = (Rr << 16) | (Gg << 8) | Bb;
      }
      A.cluster_size[ic] = (uint32_t)size;
      for (int c = 0; c < 3; ++c) A.cluster_mean[3 * ic + c] = mean[c];
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, size > 0);
    if (lane == 0) S.warp_tmp[warp] = __popc(ballot);
    __syncthreads();
    int before = S.scan_carry;
    for (int q = 0; q < warp; ++q) before += S.warp_tmp[q];
    if (size > 0) A.palette[before + __popc(ballot & ((1u << lane) - 1u))] = colour;
    __syncthreads();
    if (tid == 0) {
      int tot = 0;
      for (int q = 0; q < kExactThreads / 32; ++q) tot += S.warp_tmp[q];
      S.scan_carry += tot;
    }
    __syncthreads();
  }
  if (tid == 0) {
    A.result[0] = (uint32_t)S.scan_carry;
    A.result[1] = (uint32_t)(K - S.scan_carry);
    A.ctl[kCtlDone] = 1;
    A.ctl[kCtlRounds] = (uint32_t)(K - 1);
    A.ctl[kCtlSplits] = (uint32_t)(K - 1);
  }
}

void split_exact_launch(const SplitArgs &args, const uint32_t *d_in, uint32_t num_rows, uint32_t num_cols, uint32_t dec,
                        int bits, const uint32_t *d_uniq, uint32_t *d_table, uint32_t *d_first_seen, double *g_f64,
                        int32_t *g_i32, int sm_count, cudaStream_t st) {
  const uint32_t nr = (num_rows + dec - 1) / dec, nc = (num_cols + dec - 1) / dec;
  const uint32_t samples = nr * nc;
  const uint32_t shift = 8u - (uint32_t)bits;
  const uint32_t byte_mask = (0xFFu >> shift) << shift;
  const uint32_t word_mask = (byte_mask << 16) | (byte_mask << 8) | byte_mask;
  exact_prepare_kernel<<<(kExactMaxPoints + 255) / 256, 256, 0, st>>>(d_uniq, args.num_points_dev, d_first_seen);
  const unsigned blocks = (unsigned)std::min<uint64_t>(((uint64_t)samples + 255) / 256, (uint64_t)sm_count * 8);
  exact_first_seen_kernel<<<std::max(blocks, 1u), 256, 0, st>>>(d_in, args.num_points_dev, nc, samples, num_rows, dec, word_mask,
                                                                shift, d_first_seen);
  const size_t K = args.num_colors;
  static bool configured = false;
  if (!configured) {
    DQ_CUDA_CHECK(cudaFuncSetAttribute(split_exact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ExactShared)));
    configured = true;
  }
  split_exact_kernel<<<1, kExactThreads, sizeof(ExactShared), st>>>(args, d_uniq, d_table, d_first_seen, g_f64, g_f64 + K, g_f64 + 2 * K,
                                                  g_f64 + 5 * K, g_i32);
  DQ_CUDA_CHECK(cudaGetLastError());
}

}  // namespace dq
