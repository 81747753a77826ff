// dq_split_exact.cuh -- the divisive phase for SMALL inputs, in the reference's own summation order.
//
// Reference: DivQuantCluster<UW=false,...>, DivQuant/DivQuantCluster.cpp:133-1097 (weighted path) and
// DivQuantClusterInitMeanAndVar, :49-123.
//
// Why this exists.  In the weighted path the reference accumulates  weight_i * channel_i  as doubles, one point
// after the other, in the order calc_color_table emitted the unique colours (hash bucket ascending, most
// recently first-seen colour first inside a bucket, MapColors.cpp:157-198).  The large-input kernels
// (dq_split2.cu) use exact integer sums instead, which differ from those doubles by ~1e-16 relative.  That
// never matters unless a decision sits exactly on a tie -- a mean on an integer, two equal variances or TSEs --
// which is what small or synthetic inputs (few colours, equal counts, tiny clusters) produce.  There the
// reference decides by its own rounding noise, so to return its palette bit for bit the noise has to be
// reproduced: for U <= kExactMaxPoints this kernel walks the reference's loop with the same sequential double
// accumulations (one lane per accumulator chain, non-contracted IEEE operations in the reference's order).
//
// One CTA.  Launched for every weighted call; returns at once when U is larger (the split kernels return at
// once when it is not), so no host round trip is needed to choose.
//
// Structure of one pass over the cluster being split (cur[0..cur_n), ascending original order):
//   select   every thread classifies a contiguous piece, an exclusive scan gives each new-side point its rank
//   terms    w*c (and w, w*c*c) of the new-side points are written to shared memory in rank order
//   chains   lane l < 7 of warp 0 adds terms[l][0..n) one after the other -- the reference's sum, add for add
//   scalars  lanes 0..2 derive the centres per channel, lane 0 the hyperplane
// Clusters of up to 512 points are handled by warp 0 alone (no block barriers inside the 11 passes).
#pragma once

#include "dq_split_math.cuh"

#include <cfloat>

namespace dq {
namespace exact {

constexpr int kExactSortCap = 4096;  // power of two >= kExactMaxPoints
constexpr int kExactIndexBits = 12;
constexpr int kExactTile = 1024;     // new-side points whose terms are staged at a time
constexpr int kExactSolo = 512;      // clusters up to this size are split by warp 0 alone
constexpr int kExactSmemColors = 1024;
static_assert(kExactSortCap >= (int)kExactMaxPoints, "sort capacity");
static_assert(kExactSolo <= 16 * 32 && kExactSolo <= kExactTile, "a solo pass fits one tile and 16 points per lane");
static_assert(kExactMaxPoints <= 16 * 256, "a block pass has at most 16 points per thread");

struct ExactShared {
  union {
    unsigned long long keys[kExactSortCap];  // only while the points are put in order
    double terms[7][kExactTile];
  };
  double w[kExactMaxPoints];         // weights[] of calc_color_table (:172,185)
  uint32_t colour[kExactMaxPoints];
  uint16_t member[kExactMaxPoints];  // K <= kExactMaxColors <= 65535
  uint16_t cur[kExactMaxPoints];     // points of the cluster being split, ascending original order (:929-1019)
  // per-cluster arrays of the reference (:296-324) when K fits, else the global scratch is used
  double k_weight[kExactSmemColors], k_tse[kExactSmemColors], k_mean[3 * kExactSmemColors], k_var[3 * kExactSmemColors];
  int32_t k_size[kExactSmemColors];
  double chain[8];
  int32_t warp_tmp[32];
  double red_val[32];
  int32_t red_idx[32];
  int32_t cur_n, old_index, new_size, emitted;
  // scalars of the current split
  double tw, tm[3], tv[3], nw, ow, nm[3], om[3], nv[3];
  double lhs, rr[3], cut;
  int32_t axis;
};

__device__ __forceinline__ double chan(uint32_t p, int c) { return byte_to_double((p >> (16 - 8 * c)) & 0xFFu); }
__device__ __forceinline__ double chan_sq(uint32_t p, int c) {
  const uint32_t v = (p >> (16 - 8 * c)) & 0xFFu;
  return u52_to_double((uint64_t)(v * v));  // the reference squares in int, then converts (:98-100, :741-743)
}

template <bool SOLO>
__device__ __forceinline__ void group_sync() {
  if (SOLO) __syncwarp();
  else __syncthreads();
}

__device__ __forceinline__ void store_terms(ExactShared &S, int slot, int idx, int nchains) {
  const double wt = S.w[idx];
  const uint32_t p = S.colour[idx];
  S.terms[0][slot] = fmul(wt, chan(p, 0));
  S.terms[1][slot] = fmul(wt, chan(p, 1));
  S.terms[2][slot] = fmul(wt, chan(p, 2));
  S.terms[3][slot] = wt;
  if (nchains > 4) {
    S.terms[4][slot] = fmul(wt, chan_sq(p, 0));
    S.terms[5][slot] = fmul(wt, chan_sq(p, 1));
    S.terms[6][slot] = fmul(wt, chan_sq(p, 2));
  }
}

// lane l < nchains of warp 0 continues chain l over the n staged terms, in order
__device__ __forceinline__ void add_terms(ExactShared &S, int n, int nchains, double &acc) {
  const int t = threadIdx.x;
  if (t < nchains) {
    const double *src = S.terms[t];
    int j = 0;
    for (; j + 8 <= n; j += 8) {
      const double a0 = src[j], a1 = src[j + 1], a2 = src[j + 2], a3 = src[j + 3];
      const double a4 = src[j + 4], a5 = src[j + 5], a6 = src[j + 6], a7 = src[j + 7];
      acc = fadd(acc, a0);
      acc = fadd(acc, a1);
      acc = fadd(acc, a2);
      acc = fadd(acc, a3);
      acc = fadd(acc, a4);
      acc = fadd(acc, a5);
      acc = fadd(acc, a6);
      acc = fadd(acc, a7);
    }
    for (; j < n; ++j) acc = fadd(acc, src[j]);
  }
}

// One pass: classify cur[0..cur_n) with pred (true = new side; `each` sees every point), then sum the new side's
// terms in order.  Leaves chain[0..nchains) and new_size in S.  SOLO: executed by warp 0 only.
// Returns true when the new side is the same set of points as in the previous pass (prev_mask, updated).
template <int THREADS, bool SOLO, typename Pred, typename Each>
__device__ __forceinline__ bool pass_sums(ExactShared &S, int cur_n, int nchains, unsigned &prev_mask, Pred pred, Each each) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gsize = SOLO ? 32 : THREADS;
  const int per = (cur_n + gsize - 1) / gsize;  // <= 16
  const int lo = min(tid * per, cur_n), hi = min(lo + per, cur_n);
  unsigned mask = 0;
  for (int j = lo; j < hi; ++j) {
    const int idx = S.cur[j];
    const bool is_new = pred(S.colour[idx]);
    each(idx, is_new);
    mask |= (unsigned)is_new << (j - lo);
  }
  const int mine = __popc(mask);
  const bool changed = (mask != prev_mask);
  prev_mask = mask;
  int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  int first = incl - mine, total;
  bool any_changed;
  if (SOLO) {
    total = __shfl_sync(0xffffffffu, incl, 31);
    any_changed = __any_sync(0xffffffffu, changed);
  } else {
    if (lane == 31) S.warp_tmp[warp] = incl;
    any_changed = __syncthreads_or(changed) != 0;
    total = 0;
    for (int q = 0; q < THREADS / 32; ++q) {
      if (q < warp) first += S.warp_tmp[q];
      total += S.warp_tmp[q];
    }
  }
  double acc = 0.0;
  for (int tile = 0; tile < total; tile += kExactTile) {
    // my new-side points have ranks [first, first + mine): stage those that fall into this tile
    int rank = first;
    for (int j = lo; j < hi; ++j) {
      if ((mask >> (j - lo)) & 1u) {
        if (rank >= tile && rank < tile + kExactTile) store_terms(S, rank - tile, S.cur[j], nchains);
        ++rank;
      }
    }
    group_sync<SOLO>();
    add_terms(S, min(total - tile, kExactTile), nchains, acc);
    group_sync<SOLO>();
  }
  if (tid < nchains) S.chain[tid] = acc;
  if (tid == 0) S.new_size = total;
  group_sync<SOLO>();
  return !any_changed;
}

// new centre from the chains, old centre by the 'combined mean' (:561-581, :780-810); lanes 0..2 = channels
template <bool SOLO>
__device__ __forceinline__ void derive_centres(ExactShared &S, bool with_squares) {
  const int t = threadIdx.x;
  if (t < 3) {
    const double nw = S.chain[3], ow = fsub(S.tw, nw);
    const double nm = fdiv(S.chain[t], nw);
    S.nm[t] = nm;
    S.om[t] = fdiv(fsub(fmul(S.tw, S.tm[t]), fmul(nw, nm)), ow);
    if (with_squares) S.nv[t] = S.chain[4 + t];
    if (t == 0) S.nw = nw, S.ow = ow;
  }
  group_sync<SOLO>();
}

// lhs / rhs of the hyperplane test (:616-623)
template <bool SOLO>
__device__ __forceinline__ void derive_hyperplane(ExactShared &S) {
  if (threadIdx.x == 0) {
    double l = fsub(fsq(S.om[0]), fsq(S.nm[0]));
    l = fadd(l, fsq(S.om[1]));
    l = fsub(l, fsq(S.nm[1]));
    l = fadd(l, fsq(S.om[2]));
    l = fsub(l, fsq(S.nm[2]));
    S.lhs = fmul(0.5, l);
    for (int c = 0; c < 3; ++c) S.rr[c] = fsub(S.om[c], S.nm[c]);
  }
  group_sync<SOLO>();
}

// split pass + max_iters LKM passes of one split (:438-811)
template <int THREADS, bool SOLO>
__device__ __forceinline__ void split_passes(ExactShared &S, int cur_n, int max_iters, int new_index, int old_index) {
  unsigned prev_mask = 0xFFFFFFFFu;  // a piece has at most 16 points: never a real mask
  {
    const int axis = S.axis;
    const double cut = S.cut;
    pass_sums<THREADS, SOLO>(S, cur_n, 4, prev_mask, [&](uint32_t p) { return cut < chan(p, axis); }, [](int, bool) {});
  }
  derive_centres<SOLO>(S, false);
  for (int it = 0; it < max_iters; ++it) {
    derive_hyperplane<SOLO>(S);
    const double lhs = S.lhs, r0 = S.rr[0], r1 = S.rr[1], r2 = S.rr[2];
    const bool last = (it == max_iters - 1);
    const bool fixed_point = pass_sums<THREADS, SOLO>(
        S, cur_n, last ? 7 : 4, prev_mask,
        [&](uint32_t p) {
          const double dot = fadd(fadd(fmul(r0, chan(p, 0)), fmul(r1, chan(p, 1))), fmul(r2, chan(p, 2)));
          return !(lhs < dot);  // (:683)
        },
        [&](int idx, bool is_new) {
          if (last) S.member[idx] = (uint16_t)(is_new ? new_index : old_index);
        });
    derive_centres<SOLO>(S, last);
    // Same new side as in the previous pass: the same sums, hence the same centres and the same side again, until
    // the last iteration -- which is the only one left to run (it also sums w*c*c and writes member[]).
    if (fixed_point && it < max_iters - 2) it = max_iters - 2;
  }
}


// The whole divisive phase of one small input by one CTA of THREADS threads.  smem: sizeof(ExactShared) bytes.
// uniq / table: the histogram (unique colours in arrival order, counts; the counts are zeroed on the way);
// first_seen[c]: smallest sample index of colour c; g_*: per-cluster scratch in global memory, used when K > 1024.
template <int THREADS>
__device__ void split_exact_body(const SplitArgs &A, int U, unsigned char *smem, const uint32_t *uniq, uint32_t *table,
                                 const uint32_t *first_seen, double *g_weight, double *g_tse, double *g_mean, double *g_var,
                                 int32_t *g_size) {
  ExactShared &S = *reinterpret_cast<ExactShared *>(smem);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int K = (int)A.num_colors;
  const bool in_smem = K <= kExactSmemColors;
  double *const weight = in_smem ? S.k_weight : g_weight;
  double *const tse = in_smem ? S.k_tse : g_tse;
  double *const mean = in_smem ? S.k_mean : g_mean;
  double *const var = in_smem ? S.k_var : g_var;
  int32_t *const size = in_smem ? S.k_size : g_size;

  // ---- points in calc_color_table's emission order: (bucket asc, first seen desc) ----
  int sort_n = 32;
  while (sort_n < U) sort_n <<= 1;  // padding keys are all-ones: sorting the first power of two >= U is enough
  for (int i = tid; i < sort_n; i += THREADS) {
    unsigned long long key = ~0ull;
    if (i < U) {
      const uint32_t c = uniq[i];
      const long R = (c >> 16) & 0xFF, G = (c >> 8) & 0xFF, B = c & 0xFF;
      const unsigned long long bucket = (unsigned long long)(((R * 33023 + G * 30013 + B * 27011) & 0x7fffffff) % 20023);
      key = (bucket << (32 + kExactIndexBits)) | ((unsigned long long)(0xFFFFFFFFu - ld_cg_u32(first_seen + c)) << kExactIndexBits) |
            (unsigned long long)i;
    }
    S.keys[i] = key;
  }
  __syncthreads();
  for (int k = 2; k <= sort_n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < sort_n; i += THREADS) {
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
  for (int i = tid; i < U; i += THREADS) {
    const uint32_t c = uniq[(int)(S.keys[i] & ((1ull << kExactIndexBits) - 1ull))];
    const uint32_t count = table[c];
    table[c] = 0u;  // the count table is all-zero again when the call ends
    S.colour[i] = c;
    S.w[i] = fmul(A.norm, (double)(int)count);  // weights[i] = weight * count (:185)
    S.member[i] = 0;
    S.cur[i] = (uint16_t)i;
    A.pts[0][i] = make_uint2(c, count);
  }
  for (int i = tid; i < K; i += THREADS) {  // `new T[n]()` of the reference (:296-324)
    weight[i] = 0.0;
    tse[i] = 0.0;
    size[i] = 0;
    for (int c = 0; c < 3; ++c) mean[3 * i + c] = 0.0, var[3 * i + c] = 0.0;
  }
  __syncthreads();  // keys are dead from here on: the union now holds terms

  // ---- DivQuantClusterInitMeanAndVar (:60-104): chains 0..2 mean, 4..6 second moments ----
  unsigned no_prev = 0xFFFFFFFFu;
  pass_sums<THREADS, false>(S, U, 7, no_prev, [](uint32_t) { return true; }, [](int, bool) {});
  if (tid == 0) {
    for (int c = 0; c < 3; ++c) {
      S.tm[c] = S.chain[c];
      S.tv[c] = fsub(S.chain[4 + c], fsq(S.chain[c]));
    }
    weight[0] = 1.0;
    size[0] = U;
    S.old_index = 0;
    S.cur_n = U;
  }
  __syncthreads();

  for (int new_index = 1; new_index < K; ++new_index) {
    const int old_index = S.old_index, cur_n = S.cur_n;
    if (tid == 0) {
      S.tw = weight[old_index];
      if (new_index > 1) {
        for (int c = 0; c < 3; ++c) S.tm[c] = mean[3 * old_index + c], S.tv[c] = var[3 * old_index + c];
      }
      int axis;
      double cut;
      choose_cut(S.tv, S.tm, axis, cut);
      S.axis = axis;
      S.cut = cut;
    }
    __syncthreads();
    if (cur_n <= kExactSolo) {
      if (warp == 0) split_passes<THREADS, true>(S, cur_n, A.max_iters, new_index, old_index);
    } else {
      split_passes<THREADS, false>(S, cur_n, A.max_iters, new_index, old_index);
    }
    __syncthreads();

    if (tid == 0) {
      size[old_index] = cur_n - S.new_size;
      size[new_index] = S.new_size;
      for (int c = 0; c < 3; ++c) mean[3 * new_index + c] = S.nm[c], mean[3 * old_index + c] = S.om[c];
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
          var[3 * new_index + c] = nv[c];
          var[3 * old_index + c] = ov[c];
        }
        weight[old_index] = S.ow;
        weight[new_index] = S.nw;
        tse[old_index] = fmul(S.ow, fadd(fadd(ov[0], ov[1]), ov[2]));  // (:871)
        tse[new_index] = fmul(S.nw, fadd(fadd(nv[0], nv[1]), nv[2]));
        if (A.records != nullptr) {
          for (int c = 0; c < 3; ++c) r.new_var[c] = nv[c], r.old_var[c] = ov[c];
          r.new_tse = tse[new_index];
          r.old_tse = tse[old_index];
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
      for (int ic = tid; ic <= new_index; ic += THREADS) {
        const double t = tse[ic];
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
        for (int q = 1; q < THREADS / 32; ++q) {
          const double ob = S.red_val[q];
          const int oi = S.red_idx[q];
          if (oi >= 0 && (best_i < 0 || best < ob || (ob == best && oi < best_i))) best = ob, best_i = oi;
        }
        if (best_i >= 0) S.old_index = best_i;
      }
      __syncthreads();
    }
    // ---- gather its points in ascending original order (:929-1019): contiguous pieces + exclusive scan ----
    {
      const uint16_t want = (uint16_t)S.old_index;
      const int per = (U + THREADS - 1) / THREADS;  // <= 16
      const int lo = min(tid * per, U), hi = min(lo + per, U);
      unsigned mask = 0;
      for (int i = lo; i < hi; ++i) mask |= (unsigned)(S.member[i] == want) << (i - lo);
      const int mine = __popc(mask);
      int incl = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      if (lane == 31) S.warp_tmp[warp] = incl;
      __syncthreads();
      int pos = incl - mine, total = 0;
      for (int q = 0; q < THREADS / 32; ++q) {
        if (q < warp) pos += S.warp_tmp[q];
        total += S.warp_tmp[q];
      }
      for (int i = lo; i < hi; ++i)
        if ((mask >> (i - lo)) & 1u) S.cur[pos++] = (uint16_t)i;
      if (tid == 0) {
        S.cur_n = total;
        if (total != size[S.old_index]) A.ctl[kCtlError] = 7;  // "Cluster to be split is expected to be of size ..." (:1013)
      }
      __syncthreads();
    }
  }

  // ---- palette = rounded means of the non-empty clusters in index order (:1030-1065) ----
  if (tid == 0) S.emitted = 0;
  __syncthreads();
  for (int base = 0; base < K; base += THREADS) {
    const int ic = base + tid;
    uint32_t colour = 0;
    int sz = 0;
    if (ic < K) {
      sz = (K > 1) ? size[ic] : U;
      double m[3] = {0.0, 0.0, 0.0};  // K == 1 never assigns mean[0]
      if (K > 1) m[0] = mean[3 * ic], m[1] = mean[3 * ic + 1], m[2] = mean[3 * ic + 2];
      if (sz > 0) {
        const uint32_t Rr = (__double2uint_rz(fadd(m[0], 0.5)) & 0xFFu) << A.shift;
        const uint32_t Gg = (__double2uint_rz(fadd(m[1], 0.5)) & 0xFFu) << A.shift;
        const uint32_t Bb = (__double2uint_rz(fadd(m[2], 0.5)) & 0xFFu) << A.shift;
        colour = (Rr << 16) | (Gg << 8) | Bb;
      }
      A.cluster_size[ic] = (uint32_t)sz;
      for (int c = 0; c < 3; ++c) A.cluster_mean[3 * ic + c] = m[c];
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, sz > 0);
    if (lane == 0) S.warp_tmp[warp] = __popc(ballot);
    __syncthreads();
    int before = S.emitted;
    for (int q = 0; q < warp; ++q) before += S.warp_tmp[q];
    if (sz > 0) A.palette[before + __popc(ballot & ((1u << lane) - 1u))] = colour;
    __syncthreads();
    if (tid == 0) {
      int tot = 0;
      for (int q = 0; q < THREADS / 32; ++q) tot += S.warp_tmp[q];
      S.emitted += tot;
    }
    __syncthreads();
  }
  if (tid == 0) {
    A.result[0] = (uint32_t)S.emitted;
    A.result[1] = (uint32_t)(K - S.emitted);
    A.ctl[kCtlDone] = 1;
    A.ctl[kCtlRounds] = (uint32_t)(K - 1);
    A.ctl[kCtlSplits] = (uint32_t)(K - 1);
  }
}

}  // namespace exact

// first_seen pass shared by the fused and the stand-alone form: smallest sample index of every colour (sampling and bit
// cutting as in hist_insert, whose first toucher of every colour has set the entry to 0xFFFFFFFF).
__device__ __forceinline__ void exact_first_seen(const ExactSampling &q, uint32_t *first_seen, uint32_t thread, uint32_t threads) {
  for (uint32_t s = thread; s < q.num_samples; s += threads) {
    const uint32_t ir = (s / q.samples_per_row) * q.dec, ic = (s % q.samples_per_row) * q.dec;
    // the reference's addressing, bug included (MapColors.cpp:120-122)
    const uint32_t c = (q.in[ic + ir * q.num_rows] & q.word_mask) >> q.shift;
    if (s < first_seen[c]) atomicMin(first_seen + c, s);  // the plain read only skips atomics that cannot win
  }
}

}  // namespace dq
