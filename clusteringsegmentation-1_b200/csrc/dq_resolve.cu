// dq_resolve.cu -- a final cluster's centre in the reference's own arithmetic, for the palette entries whose rounding the
// tie audit flagged (kTieRound, dq_tie.cuh): (uint8)(mean + 0.5) with the exact mean on x.5 (DivQuantCluster.cpp:1050-1052).
//
// On such a tie the reference decides by the rounding noise of its sequential double sums, so the sums have to be redone
// in its order -- but only the few that feed this one mean.  mean[ic] of a final cluster is (:561-581, :780-810)
//   * for a "new" side: sum(w c) / sum(w) over its own points, added one after the other in calc_color_table's
//     emission order (hash bucket ascending, most recently first-seen colour first, MapColors.cpp:157-198);
//   * for an "old" side: (tw tm - nw nm) / (tw - nw) from its parent's (tw, tm) and its sibling's sums -- recursively
//     up the chain of "old" sides to the nearest "new" side (or to the root, whose sums run over all points).
// Every set involved is a node of the exact-integer split tree (the memberships are not in doubt: only the rounding is),
// and all of them lie inside the top node of that chain.  So one CTA per flagged cluster: collect the top node's points
// from the leaves below it, sort them into emission order (first-seen index from the first-seen pass), add the chains
// sequentially -- one thread per (set, accumulator) -- and replay the scalar formulas down the chain.
// Chains through large nodes (more than kResolveMax points, e.g. the ever-shrinking cluster 0 of a 4K frame) get status 2
// from that kernel and go through the large form below: the points under the top nodes are put into emission order by a
// global sort (any number of them), and one warp per 8 sets streams the sorted list, lane = (set, accumulator).
#include <algorithm>
#include <cfloat>

#include "dq_kernels.cuh"
#include "dq_split_math.cuh"

namespace dq {
namespace {

constexpr int kResolveThreads = 512;
constexpr int kResolveMax = 8192;   // points of the chain's top node
constexpr int kResolveChain = 24;   // "old" steps between the flagged cluster and the top node
constexpr int kResolveLeaves = 1024;
constexpr int kResolveNodes = 8 * 512 + 16;  // node table of the split kernel for K <= kSplit2MaxColors

struct ResolveShared {
  union {
    unsigned long long keys[kResolveMax];  // sort keys; afterwards ...
    double w_sorted[kResolveMax];          // ... the weights in emission order
  };
  union {
    uint2 pts[kResolveMax];         // (colour, count) by position inside the top node's range; afterwards ...
    struct {
      uint32_t colour_sorted[kResolveMax];  // ... colours and positions in emission order
      uint16_t pos_sorted[kResolveMax];
    } so;
  };
  // the split tree's links and segments, staged once (walking them in global memory costs a round trip per step)
  uint32_t nd_begin[kResolveNodes], nd_size[kResolveNodes];
  int16_t nd_parent[kResolveNodes], nd_child[kResolveNodes];  // node ids fit: < 4112
  uint8_t nd_buf[kResolveNodes];
  double sums[kResolveChain + 1][4];
  int32_t leaf[kResolveLeaves];
  int32_t sib[kResolveChain];
  uint32_t sib_begin[kResolveChain + 1], sib_size[kResolveChain + 1];
  int32_t n_leaves, n_sib, top, fail, top_is_root;
};

// The chain of "old" sides above node_x: S.sib[] = the "new" siblings whose sums the chain was derived from, S.top = the node
// whose statistics are its own sums (a "new" side, or the root), S.sib_begin/size[0] = its range, [k + 1] = sibling k's.
// Thread 0 only.
template <class SH>
__device__ __forceinline__ void walk_chain(SH &S, int node_x) {
  S.fail = 0;
  S.n_sib = 0;
  S.top_is_root = 0;
  int cur = node_x;
  for (;;) {
    const int p = S.nd_parent[cur];
    if (p < 0) {  // the root: its statistics are sums over every point (:60-104)
      S.top_is_root = 1;
      break;
    }
    const int child0 = S.nd_child[p];
    if (cur == child0 + 1) break;  // a "new" side: its statistics are its own sums
    if (S.n_sib >= kResolveChain) {
      S.fail = 1;
      break;
    }
    S.sib[S.n_sib++] = child0 + 1;  // the sibling whose sums this "old" side was derived from
    cur = p;
  }
  S.top = cur;
  if (S.nd_size[cur] == 0u) S.fail = 1;
  S.sib_begin[0] = S.nd_begin[cur];
  S.sib_size[0] = S.nd_size[cur];
  for (int k = 0; k < S.n_sib && !S.fail; ++k) {
    S.sib_begin[k + 1] = S.nd_begin[S.sib[k]];
    S.sib_size[k + 1] = S.nd_size[S.sib[k]];
  }
}

template <class SH>
__device__ __forceinline__ void stage_nodes(SH &S, const SplitNode *nodes, uint32_t num_nodes, int tid, int threads) {
  for (uint32_t i = tid; i < num_nodes; i += threads) {
    const SplitNode *nd = nodes + i;
    S.nd_begin[i] = __ldcg(&nd->begin);
    S.nd_size[i] = __ldcg(&nd->size);
    S.nd_parent[i] = (int16_t)__ldcg(&nd->parent);
    S.nd_child[i] = (int16_t)__ldcg(&nd->child);
  }
}

// The scalar formulas down the chain from S.sums (set 0 = the top node): the node's mean as the reference holds it.
template <class SH>
__device__ __forceinline__ void replay_chain(const SH &S, double *tm) {
  double tw;
  if (S.top_is_root) {
    tw = 1.0;  // weight[0] = 1.0 (:343); the root's means are the plain sums (:107-112)
    for (int c = 0; c < 3; ++c) tm[c] = S.sums[0][c];
  } else {
    tw = S.sums[0][3];
    for (int c = 0; c < 3; ++c) tm[c] = fdiv(S.sums[0][c], tw);
  }
  for (int k = S.n_sib - 1; k >= 0; --k) {  // from the top node down to the flagged node
    const double nw = S.sums[k + 1][3];
    const double ow = fsub(tw, nw);
    for (int c = 0; c < 3; ++c) {
      const double nm = fdiv(S.sums[k + 1][c], nw);
      tm[c] = fdiv(fsub(fmul(tw, tm[c]), fmul(nw, nm)), ow);  // 'combined mean' (:579-581, :805-810)
    }
    tw = ow;
  }
}

// What the resolved mean settles.  A rounding entry (item < n_round): the palette word (:1050-1052), status 1.
// A cut entry: the split of this node sent a point to the new side iff  cut_pos < value  (:473) with cut_pos = the mean on
// the cut axis; the values are integers, so the reference's cut and this kernel's separate the same points iff they
// have the same floor: status 1 (nothing to change) or 3 (the reference splits this node differently).
template <class SH>
__device__ __forceinline__ void settle(const SH &S, const SplitNode *nodes, int node_x, bool is_cut, int slot, int shift,
                                       uint32_t *palette, uint32_t *status_word, CutOverride *ref_cut) {
  double tm[3];
  replay_chain(S, tm);
  if (is_cut) {
    const int axis = __ldcg(&nodes[node_x].axis);
    const double cut_here = __ldcg(&nodes[node_x].cut);
    // what the host needs to have this node cut where the reference cuts it (SplitArgs::cut_overrides)
    ref_cut->begin = __ldcg(&nodes[node_x].begin);
    ref_cut->size = __ldcg(&nodes[node_x].size);
    ref_cut->mean_here = cut_here;
    ref_cut->mean_ref = tm[axis];
    *status_word = (floor(tm[axis]) == floor(cut_here)) ? 1u : 3u;
    return;
  }
  const uint32_t Rr = (__double2uint_rz(fadd(tm[0], 0.5)) & 0xFFu) << shift;
  const uint32_t Gg = (__double2uint_rz(fadd(tm[1], 0.5)) & 0xFFu) << shift;
  const uint32_t Bb = (__double2uint_rz(fadd(tm[2], 0.5)) & 0xFFu) << shift;
  palette[slot] = (Rr << 16) | (Gg << 8) | Bb;
  *status_word = 1u;
}
// item -> (node, palette slot, kind) from the two lists
__device__ __forceinline__ void item_of(const uint32_t *list, uint32_t n_round, int item, int &node_x, int &slot, bool &is_cut) {
  is_cut = (uint32_t)item >= n_round;
  node_x = is_cut ? (int)list[kTieCutList + (uint32_t)item - n_round] : (int)list[4 * item + 1];
  slot = is_cut ? 0 : (int)list[4 * item + 2];
}

__global__ void __launch_bounds__(kResolveThreads) tie_resolve_kernel(const SplitNode *nodes, uint2 *pts0, uint2 *pts1,
                                                                      const uint32_t *first_seen, double norm, int shift,
                                                                      const uint32_t *list, uint32_t n_round, uint32_t num_nodes,
                                                                      uint32_t *palette, uint32_t *status, CutOverride *ref_cuts) {
  extern __shared__ __align__(16) unsigned char resolve_smem[];
  ResolveShared &S = *reinterpret_cast<ResolveShared *>(resolve_smem);
  const int tid = threadIdx.x, item = blockIdx.x;
  int node_x, slot;
  bool is_cut;
  item_of(list, n_round, item, node_x, slot, is_cut);

  if (num_nodes > (uint32_t)kResolveNodes) {
    if (tid == 0) status[item] = 2u;
    return;
  }
  for (uint32_t i = tid; i < num_nodes; i += kResolveThreads) {
    const SplitNode *nd = nodes + i;
    S.nd_begin[i] = __ldcg(&nd->begin);
    S.nd_size[i] = __ldcg(&nd->size);
    S.nd_parent[i] = (int16_t)__ldcg(&nd->parent);
    S.nd_child[i] = (int16_t)__ldcg(&nd->child);
    S.nd_buf[i] = (uint8_t)__ldcg(&nd->buf);
  }
  __syncthreads();
  // ---- the chain of "old" sides above the flagged cluster, and the leaves below its top node ----
  if (tid == 0) {
    walk_chain(S, node_x);
    const int cur = S.top;
    if (S.nd_size[cur] > (uint32_t)kResolveMax) S.fail = 1;
    // leaves below the top node (depth-first, explicit stack)
    S.n_leaves = 0;
    if (!S.fail) {
      int stack[64], sp = 0;
      stack[sp++] = cur;
      while (sp > 0) {
        const int nd = stack[--sp];
        const int ch = S.nd_child[nd];
        if (ch < 0) {
          if (S.n_leaves >= kResolveLeaves) {
            S.fail = 1;
            break;
          }
          S.leaf[S.n_leaves++] = nd;
        } else {
          if (sp + 2 > 64) {
            S.fail = 1;
            break;
          }
          stack[sp++] = ch;
          stack[sp++] = ch + 1;
        }
      }
    }
  }
  __syncthreads();
  if (S.fail) {
    if (tid == 0) status[item] = 2u;
    return;
  }
  const uint32_t base = S.sib_begin[0], n = S.sib_size[0];
  // ---- the top node's points: every leaf keeps its segment in its own buffer ----
  for (int l = 0; l < S.n_leaves; ++l) {
    const int lf = S.leaf[l];
    const uint32_t lb = S.nd_begin[lf], ls = S.nd_size[lf];
    const uint2 *src = (S.nd_buf[lf] ? pts1 : pts0) + lb;
    for (uint32_t i = tid; i < ls; i += kResolveThreads) S.pts[lb - base + i] = __ldcg(src + i);
  }
  __syncthreads();
  // ---- emission order: (bucket asc, first seen desc), MapColors.cpp:59-62, 157-198 ----
  int sort_n = 32;
  while (sort_n < (int)n) sort_n <<= 1;
  for (int i = tid; i < sort_n; i += kResolveThreads) {
    unsigned long long key = ~0ull;
    if (i < (int)n) {
      const uint32_t c = S.pts[i].x;
      const long R = (c >> 16) & 0xFF, G = (c >> 8) & 0xFF, B = c & 0xFF;
      const unsigned long long bucket = (unsigned long long)(((R * 33023 + G * 30013 + B * 27011) & 0x7fffffff) % 20023);
      key = (bucket << 44) | ((unsigned long long)(0x7FFFFFFFu - __ldcg(first_seen + c)) << 13) | (unsigned long long)i;
    }
    S.keys[i] = key;
  }
  __syncthreads();
  for (int k = 2; k <= sort_n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < sort_n; i += kResolveThreads) {
        const int partner = i ^ j;
        if (partner > i) {
          const unsigned long long a = S.keys[i], b = S.keys[partner];
          if ((a > b) == ((i & k) == 0)) {
            S.keys[i] = b;
            S.keys[partner] = a;
          }
        }
      }
      __syncthreads();
    }
  }
  // ---- weights, colours and positions in emission order (through registers: the arrays alias keys / pts) ----
  {
    constexpr int kPer = kResolveMax / kResolveThreads;
    double w_r[kPer];
    uint32_t c_r[kPer];
    uint16_t i_r[kPer];
#pragma unroll
    for (int q = 0; q < kPer; ++q) {
      const int r = tid + q * kResolveThreads;
      w_r[q] = 0.0, c_r[q] = 0u, i_r[q] = 0;
      if (r < (int)n) {
        const uint32_t i = (uint32_t)(S.keys[r] & 0x1FFFull);
        const uint2 p = S.pts[i];
        w_r[q] = fmul(norm, (double)(int)p.y);  // weights[i] = norm * count (MapColors.cpp:185)
        c_r[q] = p.x;
        i_r[q] = (uint16_t)i;
      }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < kPer; ++q) {
      const int r = tid + q * kResolveThreads;
      if (r < (int)n) {
        S.w_sorted[r] = w_r[q];
        S.so.colour_sorted[r] = c_r[q];
        S.so.pos_sorted[r] = i_r[q];
      }
    }
    __syncthreads();
  }
  // ---- the reference's sums: thread (set, chain) adds its terms one after the other; a point outside the set adds +0.0,
  //      which leaves a non-negative sum as it is, so every thread walks the same list ----
  const int n_sets = S.n_sib + 1;
  if (tid < 4 * n_sets) {
    const int set = tid >> 2, chain = tid & 3;
    const uint32_t lo = S.sib_begin[set] - base, hi = lo + S.sib_size[set];
    const int sh = 16 - 8 * chain;
    double acc = 0.0;
    uint32_t r = 0;
    for (; r + 4 <= n; r += 4) {
      double t[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t i = S.so.pos_sorted[r + q];
        const double w = S.w_sorted[r + q];
        const double term = (chain == 3) ? w : fmul(w, byte_to_double((S.so.colour_sorted[r + q] >> sh) & 0xFFu));
        t[q] = (i >= lo && i < hi) ? term : 0.0;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) acc = fadd(acc, t[q]);
    }
    for (; r < n; ++r) {
      const uint32_t i = S.so.pos_sorted[r];
      const double w = S.w_sorted[r];
      const double term = (chain == 3) ? w : fmul(w, byte_to_double((S.so.colour_sorted[r] >> sh) & 0xFFu));
      acc = fadd(acc, (i >= lo && i < hi) ? term : 0.0);
    }
    S.sums[set][chain] = acc;
  }
  __syncthreads();
  // ---- the scalar formulas down the chain ----
  if (tid == 0) settle(S, nodes, node_x, is_cut, slot, shift, palette, status + item, ref_cuts + (is_cut ? (uint32_t)item - n_round : 0u));
}


// ---- large form -------------------------------------------------------------------------------------------------------
// Three steps on the stream, any number of points:
//   flatten   every leaf copies its segment to one flat array (a node's range [begin, begin + size) is the same in either
//             buffer, so position p names a point and a set is a range of positions); points under the top node of some
//             flagged cluster get their emission-order key (bucket asc, first seen desc), all others the padding key
//   sort      keys with the position as payload (order_sort: global bitonic network, dq_hist.cu)
//   stream    one CTA per flagged cluster; lane = (set, accumulator) as above, 32 of them per warp; a warp reads the sorted
//             positions 32 at a time, gathers their points (one ahead), and hands the ones inside the top node's range
//             around by shuffle, each lane adding its term or +0.0 -- the reference's sum, add for add.
constexpr int kBigThreads = 32 * ((4 * (kResolveChain + 1) + 31) / 32);  // every (set, accumulator) has a lane

struct BigShared {
  uint32_t nd_begin[kResolveNodes], nd_size[kResolveNodes];
  int16_t nd_parent[kResolveNodes], nd_child[kResolveNodes];
  double sums[kResolveChain + 1][4];
  int32_t sib[kResolveChain];
  uint32_t sib_begin[kResolveChain + 1], sib_size[kResolveChain + 1];
  int32_t n_sib, top, fail, top_is_root;
  uint32_t range_lo[2 * kTieListCap], range_hi[2 * kTieListCap];
  int32_t n_ranges;
};

__global__ void __launch_bounds__(256) resolve_flatten_kernel(const SplitNode *nodes, uint32_t num_nodes, const uint2 *pts0,
                                                             const uint2 *pts1, const uint32_t *first_seen, const uint32_t *list,
                                                             uint32_t n_round, uint32_t count, unsigned long long *keys,
                                                             uint2 *flat, uint32_t *in_range_total) {
  extern __shared__ __align__(16) unsigned char resolve_smem[];
  BigShared &S = *reinterpret_cast<BigShared *>(resolve_smem);
  const int tid = threadIdx.x;
  if (num_nodes > (uint32_t)kResolveNodes) return;  // (the stream kernel reports it)
  stage_nodes(S, nodes, num_nodes, tid, 256);
  __syncthreads();
  if (tid == 0) {
    S.n_ranges = 0;
    for (uint32_t it = 0; it < count; ++it) {
      int node_x, slot;
      bool is_cut;
      item_of(list, n_round, (int)it, node_x, slot, is_cut);
      walk_chain(S, node_x);
      if (S.fail) continue;
      S.range_lo[S.n_ranges] = S.sib_begin[0];
      S.range_hi[S.n_ranges] = S.sib_begin[0] + S.sib_size[0];
      S.n_ranges++;
    }
  }
  __syncthreads();
  const int n_ranges = S.n_ranges;
  uint32_t mine = 0;
  for (uint32_t nd = blockIdx.x; nd < num_nodes; nd += gridDim.x) {
    if (S.nd_child[nd] >= 0) continue;  // leaves only
    const uint32_t lb = S.nd_begin[nd], ls = S.nd_size[nd];
    const uint2 *src = (__ldcg(&nodes[nd].buf) ? pts1 : pts0) + lb;
    for (uint32_t i = tid; i < ls; i += 256) {
      const uint2 p = __ldcg(src + i);
      const uint32_t pos = lb + i;
      flat[pos] = p;
      bool wanted = false;
      for (int r = 0; r < n_ranges; ++r) wanted = wanted || (pos >= S.range_lo[r] && pos < S.range_hi[r]);
      unsigned long long key = ~0ull;
      if (wanted) {
        const uint32_t c = p.x;
        const long R = (c >> 16) & 0xFF, G = (c >> 8) & 0xFF, B = c & 0xFF;
        const unsigned long long bucket = (unsigned long long)(((R * 33023 + G * 30013 + B * 27011) & 0x7fffffff) % 20023);
        key = (bucket << 31) | (unsigned long long)(0x7FFFFFFFu - __ldcg(first_seen + c));
        ++mine;
      }
      keys[pos] = key;
    }
  }
  // how many keys are real: the stream stops there
  for (int off = 16; off > 0; off >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, off);
  if ((tid & 31) == 0 && mine) atomicAdd(in_range_total, mine);
}

__global__ void __launch_bounds__(kBigThreads) tie_resolve_big_kernel(const SplitNode *nodes, uint32_t num_nodes,
                                                                      const uint32_t *__restrict__ sorted_pos,
                                                                      const uint2 *__restrict__ flat, const uint32_t *in_range_total,
                                                                      double norm, int shift, const uint32_t *list,
                                                                      uint32_t n_round, uint32_t *palette, uint32_t *status,
                                                                      CutOverride *ref_cuts) {
  extern __shared__ __align__(16) unsigned char resolve_smem[];
  BigShared &S = *reinterpret_cast<BigShared *>(resolve_smem);
  const int tid = threadIdx.x, lane = tid & 31, item = blockIdx.x;
  int node_x, slot;
  bool is_cut;
  item_of(list, n_round, item, node_x, slot, is_cut);
  if (num_nodes > (uint32_t)kResolveNodes) {
    if (tid == 0) status[item] = 2u;
    return;
  }
  stage_nodes(S, nodes, num_nodes, tid, kBigThreads);
  __syncthreads();
  if (tid == 0) walk_chain(S, node_x);
  __syncthreads();
  if (S.fail) {
    if (tid == 0) status[item] = 2u;
    return;
  }
  const uint32_t total = __ldcg(in_range_total);
  const uint32_t t_lo = S.sib_begin[0], t_hi = t_lo + S.sib_size[0];
  const int n_sets = S.n_sib + 1;
  const int set = tid >> 2, chain = tid & 3;
  const bool live = set < n_sets;
  const uint32_t lo = live ? S.sib_begin[set] : 0u, hi = live ? lo + S.sib_size[set] : 0u;
  const int sh = 16 - 8 * chain;
  if ((tid & ~31) < 4 * n_sets) {  // warps with at least one live lane
    double acc = 0.0;
    // one block of 32 sorted positions ahead: its gather is in flight while the current block is added
    uint32_t pos_n = 0xFFFFFFFFu;
    uint2 pt_n = make_uint2(0u, 0u);
    if ((uint32_t)lane < total) {
      pos_n = __ldg(sorted_pos + lane);
      if (pos_n >= t_lo && pos_n < t_hi) pt_n = __ldg(flat + pos_n);
    }
    for (uint32_t r0 = 0; r0 < total; r0 += 32) {
      const uint32_t pos = pos_n;
      const uint2 pt = pt_n;
      pos_n = 0xFFFFFFFFu;
      if (r0 + 32 + lane < total) {
        pos_n = __ldg(sorted_pos + r0 + 32 + lane);
        if (pos_n >= t_lo && pos_n < t_hi) pt_n = __ldg(flat + pos_n);
      }
      const double w = fmul(norm, (double)(int)pt.y);  // weights[i] = norm * count (MapColors.cpp:185)
      unsigned inside = __ballot_sync(0xffffffffu, pos >= t_lo && pos < t_hi);
      while (inside) {
        const int q = __ffs(inside) - 1;
        inside &= inside - 1u;
        const uint32_t pq = __shfl_sync(0xffffffffu, pos, q);
        const uint32_t cq = __shfl_sync(0xffffffffu, pt.x, q);
        const double wq = __shfl_sync(0xffffffffu, w, q);
        const double term = (chain == 3) ? wq : fmul(wq, byte_to_double((cq >> sh) & 0xFFu));
        acc = fadd(acc, (pq >= lo && pq < hi) ? term : 0.0);
      }
    }
    if (live) S.sums[set][chain] = acc;
  }
  __syncthreads();
  if (tid == 0) settle(S, nodes, node_x, is_cut, slot, shift, palette, status + item, ref_cuts + (is_cut ? (uint32_t)item - n_round : 0u));
}

}  // namespace

// the cut entries' records inside the tie_list buffer (layout: dq_split.cuh)
static inline CutOverride *ref_cuts_of(const uint32_t *d_list) {
  return reinterpret_cast<CutOverride *>(const_cast<uint32_t *>(d_list) + kTieRefCut);
}

void tie_resolve_launch(const SplitNode *d_nodes, uint32_t num_nodes, uint2 *const *pts, const uint32_t *d_first_seen, double norm,
                        int shift, const uint32_t *d_list, uint32_t n_round, uint32_t n_cut, uint32_t *d_palette, uint32_t *d_status,
                        cudaStream_t st) {
  DQ_RAISE_SMEM(tie_resolve_kernel, sizeof(ResolveShared));
  tie_resolve_kernel<<<n_round + n_cut, kResolveThreads, sizeof(ResolveShared), st>>>(d_nodes, pts[0], pts[1], d_first_seen, norm,
                                                                                      shift, d_list, n_round, num_nodes, d_palette,
                                                                                      d_status, ref_cuts_of(d_list));
  DQ_CUDA_CHECK(cudaGetLastError());
}

// The large form: d_keys [n_pow2], d_vals [n_pow2], d_flat [U], d_counter [1] are scratch; n_pow2 = the power of two >= U.
void tie_resolve_big_launch(const SplitNode *d_nodes, uint32_t num_nodes, uint2 *const *pts, const uint32_t *d_first_seen, uint32_t u,
                            double norm, int shift, const uint32_t *d_list, uint32_t n_round, uint32_t n_cut, uint64_t *d_keys,
                            uint32_t *d_vals, uint2 *d_flat, uint32_t *d_counter, uint32_t *d_palette, uint32_t *d_status,
                            int sm_count, cudaStream_t st) {
  const uint32_t count = n_round + n_cut;
  DQ_CUDA_CHECK(cudaMemsetAsync(d_counter, 0, sizeof(uint32_t), st));
  DQ_RAISE_SMEM(resolve_flatten_kernel, sizeof(BigShared));
  DQ_RAISE_SMEM(tie_resolve_big_kernel, sizeof(BigShared));
  resolve_flatten_kernel<<<std::min<uint32_t>(num_nodes, (uint32_t)(4 * sm_count)), 256, sizeof(BigShared), st>>>(
      d_nodes, num_nodes, pts[0], pts[1], d_first_seen, d_list, n_round, count, reinterpret_cast<unsigned long long *>(d_keys), d_flat,
      d_counter);
  order_sort(d_keys, d_vals, u, sm_count, st);
  tie_resolve_big_kernel<<<count, kBigThreads, sizeof(BigShared), st>>>(d_nodes, num_nodes, d_vals, d_flat, d_counter, norm, shift, d_list, n_round,
                                                                        d_palette, d_status, ref_cuts_of(d_list));
  DQ_CUDA_CHECK(cudaGetLastError());
}

}  // namespace dq
