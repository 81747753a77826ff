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
// Chains through large nodes (more than kResolveMax points, e.g. the ever-shrinking cluster 0 of a 4K frame) are not
// handled here: status 2, and the host re-runs the frame on the ordered path.
#include <cfloat>

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

__device__ __forceinline__ SplitNode load_node_g(const SplitNode *nodes, int id) {
  SplitNode nd;
  const double *src = reinterpret_cast<const double *>(nodes + id);
  double *dst = reinterpret_cast<double *>(&nd);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(SplitNode) / 8); ++i) dst[i] = __ldcg(src + i);
  return nd;
}

__global__ void __launch_bounds__(kResolveThreads) tie_resolve_kernel(const SplitNode *nodes, uint2 *pts0, uint2 *pts1,
                                                                      const uint32_t *first_seen, double norm, int shift,
                                                                      const uint32_t *list, uint32_t num_nodes, uint32_t *palette,
                                                                      uint32_t *status) {
  extern __shared__ __align__(16) unsigned char resolve_smem[];
  ResolveShared &S = *reinterpret_cast<ResolveShared *>(resolve_smem);
  const int tid = threadIdx.x, item = blockIdx.x;
  const int node_x = (int)list[4 * item + 1], slot = (int)list[4 * item + 2];

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
    if (S.nd_size[cur] > (uint32_t)kResolveMax || S.nd_size[cur] == 0u) S.fail = 1;
    S.sib_begin[0] = S.nd_begin[cur];
    S.sib_size[0] = S.nd_size[cur];
    for (int k = 0; k < S.n_sib && !S.fail; ++k) {
      S.sib_begin[k + 1] = S.nd_begin[S.sib[k]];
      S.sib_size[k + 1] = S.nd_size[S.sib[k]];
    }
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
  if (tid == 0) {
    double tw, tm[3];
    if (S.top_is_root) {
      tw = 1.0;  // weight[0] = 1.0 (:343); the root's means are the plain sums (:107-112)
      for (int c = 0; c < 3; ++c) tm[c] = S.sums[0][c];
    } else {
      tw = S.sums[0][3];
      for (int c = 0; c < 3; ++c) tm[c] = fdiv(S.sums[0][c], tw);
    }
    for (int k = S.n_sib - 1; k >= 0; --k) {  // from the top node down to the flagged cluster
      const double nw = S.sums[k + 1][3];
      const double ow = fsub(tw, nw);
      for (int c = 0; c < 3; ++c) {
        const double nm = fdiv(S.sums[k + 1][c], nw);
        tm[c] = fdiv(fsub(fmul(tw, tm[c]), fmul(nw, nm)), ow);  // 'combined mean' (:579-581, :805-810)
      }
      tw = ow;
    }
    const uint32_t Rr = (__double2uint_rz(fadd(tm[0], 0.5)) & 0xFFu) << shift;
    const uint32_t Gg = (__double2uint_rz(fadd(tm[1], 0.5)) & 0xFFu) << shift;
    const uint32_t Bb = (__double2uint_rz(fadd(tm[2], 0.5)) & 0xFFu) << shift;
    palette[slot] = (Rr << 16) | (Gg << 8) | Bb;
    status[item] = 1u;
  }
}

}  // namespace

void tie_resolve_launch(const SplitNode *d_nodes, uint32_t num_nodes, uint2 *const *pts, const uint32_t *d_first_seen, double norm,
                        int shift, const uint32_t *d_list, uint32_t count, uint32_t *d_palette, uint32_t *d_status, cudaStream_t st) {
  DQ_RAISE_SMEM(tie_resolve_kernel, sizeof(ResolveShared));
  tie_resolve_kernel<<<count, kResolveThreads, sizeof(ResolveShared), st>>>(d_nodes, pts[0], pts[1], d_first_seen, norm, shift,
                                                                            d_list, num_nodes, d_palette, d_status);
  DQ_CUDA_CHECK(cudaGetLastError());
}

}  // namespace dq
