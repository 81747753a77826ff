// dq_split_ordered.cuh -- the ordered (reference-summation-order) split on ALL CTAs of the split kernel.
// A fragment of dq_split2.cu's translation unit: included there after spin_expired / grid_barrier2.
//
// dq_split_exact.cuh walks the reference's K-1 splits one after the other on one CTA.  What a split produces depends
// only on the cluster's own points (in calc_color_table order) and the statistics it inherited; only WHICH cluster is
// split next depends on the global arg-max of the TSEs (DivQuantCluster.cpp:876-887).  So here
//   * CTA 0 is the controller: it replays the reference's sequence over a tree of splits (cluster index -> node),
//     consuming a node's split when it exists, computing it itself when nobody has started it;
//   * every other CTA is a worker: it claims the not-yet-split leaf of largest TSE that can still be among the
//     K-1 splits the reference makes (fewer than K-1 known nodes have a larger TSE) and splits it with the same
//     ordered passes (exact::split_passes), ahead of the controller.
// A split's points are a segment of an index list in calc_color_table order; a split partitions its segment
// stably into [old | new] in the other buffer, so every child sees its points in that order again.
// Node states: 3 not valid yet, 0 leaf, 1 being split, 2 split (children valid).  All waits are bounded.
#pragma once

namespace ordered {

using exact::Points;
using exact::Shared;
using exact::kMaxChunks;
using exact::kPiece;
using exact::kSolo;

constexpr uint32_t kInvalid = 3u, kLeaf = 0u, kBusy = 1u, kSplitDone = 2u;

// global scratch of one call (all offsets in bytes from Split2Extra::exact_scratch)
struct Scratch {
  unsigned long long *keys;  // [kExactMaxPoints]
  double *w;                 // [kExactMaxPoints]
  uint32_t *colour;          // [kExactMaxPoints]
  uint32_t *idx[2];          // [kExactMaxPoints] each: index lists, double-buffered per node
  uint32_t *state;           // [node_cap]
  uint32_t *counters;        // [0] nodes allocated  [1] ready  [2] done
};
constexpr size_t kScratchFixed = exact::kScratchBytes;

__device__ __forceinline__ Scratch carve(unsigned char *base, uint32_t node_cap) {
  Scratch s;
  s.keys = reinterpret_cast<unsigned long long *>(base);
  s.w = reinterpret_cast<double *>(base + (size_t)kExactMaxPoints * 8);
  s.colour = reinterpret_cast<uint32_t *>(base + (size_t)kExactMaxPoints * 16);
  s.idx[0] = reinterpret_cast<uint32_t *>(base + (size_t)kExactMaxPoints * 20);
  s.idx[1] = reinterpret_cast<uint32_t *>(base + (size_t)kExactMaxPoints * 24);
  s.state = reinterpret_cast<uint32_t *>(base + kScratchFixed);
  s.counters = s.state + node_cap;
  return s;
}

__device__ __forceinline__ void st_u32(uint32_t *p, uint32_t v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ SplitNode load_node_cg(const SplitNode *nodes, int id) {
  SplitNode nd;
  const double *src = reinterpret_cast<const double *>(nodes + id);
  double *dst = reinterpret_cast<double *>(&nd);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(SplitNode) / 8); ++i) dst[i] = __ldcg(src + i);
  return nd;
}
__device__ __forceinline__ void store_node(SplitNode *nodes, int id, const SplitNode &nd) {
  double *dst = reinterpret_cast<double *>(nodes + id);
  const double *src = reinterpret_cast<const double *>(&nd);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(SplitNode) / 8); ++i) __stcg(dst + i, src[i]);
}

// Splits node `id` (claimed by this CTA): ordered passes, children statistics, stable partition of its index segment.
// Every thread of the CTA calls.  Returns false when the node table is full.
__device__ bool split_node(const SplitArgs &A, Shared &S, const Scratch &G, int id) {
  const int tid = threadIdx.x, warp = tid >> 5;
  __shared__ SplitNode s_nd;
  __shared__ int s_child;
  if (tid == 0) {
    s_nd = load_node_cg(A.nodes, id);
    S.tw = s_nd.tw;
    for (int c = 0; c < 3; ++c) S.tm[c] = s_nd.tm[c], S.tv[c] = s_nd.tv[c];
    int axis;
    double cut;
    choose_cut(S.tv, S.tm, axis, cut);
    S.axis = axis;
    S.cut = cut;
    const uint32_t c0 = atomicAdd(G.counters, 2u);
    s_child = (c0 + 2u <= A.node_cap) ? (int)c0 : -1;
  }
  __syncthreads();
  if (s_child < 0) return false;
  const int cur_n = (int)s_nd.size;
  Points P;
  P.keys = G.keys, P.w = G.w, P.colour = G.colour, P.member = nullptr;
  P.cur = G.idx[s_nd.buf] + s_nd.begin;
  P.cur_shared_across_ctas = true;
  P.in_global = true;
  unsigned masks[kMaxChunks];
  const bool solo = cur_n <= kSolo;
  if (solo) {
    if (warp == 0) exact::split_passes<T, true>(S, P, cur_n, A.max_iters, 0, 0, masks, false);
  } else {
    exact::split_passes<T, false>(S, P, cur_n, A.max_iters, 0, 0, masks, false);
  }
  __syncthreads();
  const int new_size = S.new_size, old_size = cur_n - new_size;
  // ---- children (:800-871; the reference skips var / tse after its last split, nobody reads them then) ----
  if (tid == 0) {
    SplitNode o, n;
    double nv[3], ov[3];
    for (int c = 0; c < 3; ++c) {
      nv[c] = fsub(fdiv(S.nv[c], S.nw), fsq(S.nm[c]));  // (:836-838)
      ov[c] = fsub(fdiv(fsub(fmul(S.tw, S.tv[c]), fmul(S.nw, fadd(nv[c], fsq(fsub(S.nm[c], S.tm[c]))))), S.ow),
                   fsq(fsub(S.om[c], S.tm[c])));          // combined variance (:844-855)
      n.tm[c] = S.nm[c], o.tm[c] = S.om[c];
      n.tv[c] = nv[c], o.tv[c] = ov[c];
    }
    o.tw = S.ow, n.tw = S.nw;
    o.tse = fmul(S.ow, fadd(fadd(ov[0], ov[1]), ov[2]));  // (:871)
    n.tse = fmul(S.nw, fadd(fadd(nv[0], nv[1]), nv[2]));
    o.cut = n.cut = 0.0;
    o.axis = n.axis = 0;
    o.parent = n.parent = id;
    o.child = n.child = -1;
    o.buf = n.buf = s_nd.buf ^ 1;
    o.begin = s_nd.begin, o.size = (uint32_t)old_size;
    n.begin = s_nd.begin + (uint32_t)old_size, n.size = (uint32_t)new_size;
    store_node(A.nodes, s_child, o);
    store_node(A.nodes, s_child + 1, n);
    SplitNode me = s_nd;
    me.cut = S.cut, me.axis = S.axis, me.child = s_child;
    store_node(A.nodes, id, me);
  }
  // ---- stable partition of the index segment into [old | new] of the other buffer ----
  {
    uint32_t *dst = G.idx[s_nd.buf ^ 1] + s_nd.begin;
    const int gsize = solo ? 32 : T;
    const int chunk = gsize * kPiece;
    int done_new = 0, done_old = 0, ci = 0;
    if (!solo || warp == 0) {
      for (int base = 0; base < cur_n; base += chunk, ++ci) {
        const int n_here = min(chunk, cur_n - base);
        const int per = (n_here + gsize - 1) / gsize;
        const int lo = base + min(tid * per, n_here), hi = base + min(tid * per + per, n_here);
        const unsigned mask = masks[ci];
        const int mine_new = __popc(mask), mine_old = (hi - lo) - mine_new;
        int first_new, total_new, first_old, total_old;
        bool unused;
        if (solo) {
          exact::group_ranks<T, true>(S, mine_new, false, first_new, total_new, unused);
          exact::group_ranks<T, true>(S, mine_old, false, first_old, total_old, unused);
        } else {
          exact::group_ranks<T, false>(S, mine_new, false, first_new, total_new, unused);
          exact::group_ranks<T, false>(S, mine_old, false, first_old, total_old, unused);
        }
        int pn = old_size + done_new + first_new, po = done_old + first_old;
        for (int j = lo; j < hi; ++j) {
          const uint32_t v = (uint32_t)exact::load_cur(P, j);
          if ((mask >> (j - lo)) & 1u) dst[pn++] = v;
          else dst[po++] = v;
        }
        done_new += total_new;
        done_old += total_old;
      }
    }
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    st_u32(G.state + s_child, kLeaf);
    st_u32(G.state + s_child + 1, kLeaf);
    __threadfence();
    st_u32(G.state + id, kSplitDone);
  }
  __syncthreads();
  return true;
}

__device__ __forceinline__ int block_sum_int(int v) {
  __shared__ int s_part[T / 32];
  __shared__ int s_total;
  v = __reduce_add_sync(0xffffffffu, v);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int q = 0; q < T / 32; ++q) t += s_part[q];
    s_total = t;
  }
  __syncthreads();
  const int out = s_total;
  __syncthreads();
  return out;
}

// Worker: the leaf of largest TSE that is still worth splitting, or -1.  Every thread calls; CTA-uniform result.
__device__ int pick_leaf(const SplitArgs &A, const Scratch &G, int K) {
  __shared__ double s_best[T / 32];
  __shared__ int s_best_i[T / 32];
  __shared__ int s_pick;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = (int)min(ld_relaxed_u32(G.counters), A.node_cap);
  double best = DBL_MIN;
  int best_i = -1;
  for (int i = tid; i < n; i += T) {
    if (ld_relaxed_u32(G.state + i) != kLeaf) continue;
    const double t = (i == 0) ? __longlong_as_double(0x7ff0000000000000ll) : __ldcg(&A.nodes[i].tse);
    if (best < t) best = t, best_i = i;
  }
  for (int o = 16; o > 0; o >>= 1) {
    const double ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
    if (oi >= 0 && (best_i < 0 || best < ob || (ob == best && oi < best_i))) best = ob, best_i = oi;
  }
  if (lane == 0) s_best[warp] = best, s_best_i[warp] = best_i;
  __syncthreads();
  if (tid == 0) {
    for (int q = 1; q < T / 32; ++q)
      if (s_best_i[q] >= 0 && (best_i < 0 || best < s_best[q] || (s_best[q] == best && s_best_i[q] < best_i))) best = s_best[q], best_i = s_best_i[q];
    s_pick = best_i;
    s_best[0] = best;
  }
  __syncthreads();
  const int cand = s_pick;
  if (cand <= 0) return cand;  // nothing, or the root (always worth it)
  // worth it only while fewer than K-1 known nodes have a larger TSE (a child's TSE is below its parent's); the root's
  // key is +inf, so it always counts
  const double t = s_best[0];
  int above = 0;
  for (int i = tid; i < n; i += T) {
    if (ld_relaxed_u32(G.state + i) == kInvalid) continue;
    above += (i == 0) ? 1 : (int)(__ldcg(&A.nodes[i].tse) > t);
  }
  const int total = block_sum_int(above);
  return (total < K - 1) ? cand : -1;
}

// The whole divisive phase of one input on all CTAs.  Called by every thread of every CTA after the first-seen pass and
// the grid barrier that follows it; state[] = kInvalid and counters[] = 0 were set before that barrier.
__device__ __noinline__ void run(const SplitArgs &A, const Split2Extra &X, int U, unsigned char *smem, int b, int nctas, unsigned int &bar_target) {
  Shared &S = *reinterpret_cast<Shared *>(smem);
  const Scratch G = carve(X.exact_scratch, A.node_cap);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int K = (int)A.num_colors;
  __shared__ uint32_t s_flag;

  // ---- points in calc_color_table's emission order: (bucket asc, first seen desc) -- bitonic sort of the keys by the
  //      whole grid (one barrier per step; every thread loads its pairs before it stores any of them) ----
  {
    int sort_n = 32;
    while (sort_n < U) sort_n <<= 1;
    const int gthreads = nctas * T, gtid = b * T + tid;
    for (int i = gtid; i < sort_n; i += gthreads) {
      unsigned long long key = ~0ull;
      if (i < U) {
        const uint32_t c = X.collect_uniq[i];
        const long R = (c >> 16) & 0xFF, Gc = (c >> 8) & 0xFF, B = c & 0xFF;
        const unsigned long long bucket = (unsigned long long)(((R * 33023 + Gc * 30013 + B * 27011) & 0x7fffffff) % 20023);
        key = (bucket << (31 + exact::kIndexBits)) |
              ((unsigned long long)(0x7FFFFFFFu - ld_cg_u32(X.exact_first_seen + c)) << exact::kIndexBits) | (unsigned long long)i;
      }
      G.keys[i] = key;
    }
    grid_barrier2(A, bar_target, X.progress);
    const int half = sort_n >> 1;
    for (int k = 2; k <= sort_n; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int p0 = gtid; p0 < half; p0 += 2 * gthreads) {  // two compare-exchanges per trip
          int i[2];
          unsigned long long x[2], y[2];
          bool live[2];
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int pr = p0 + q * gthreads;
            live[q] = pr < half;
            i[q] = ((pr & ~(j - 1)) << 1) | (pr & (j - 1));  // the pr-th index whose bit j is clear
            x[q] = live[q] ? __ldcg(G.keys + i[q]) : 0ull;
            y[q] = live[q] ? __ldcg(G.keys + (i[q] | j)) : 0ull;
          }
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const bool up = (i[q] & k) == 0;
            if (live[q] && (x[q] > y[q]) == up) {
              __stcg(G.keys + i[q], y[q]);
              __stcg(G.keys + (i[q] | j), x[q]);
            }
          }
        }
        grid_barrier2(A, bar_target, X.progress);
      }
    }
    // the points in that order: colour, weight, identity index list; the count table is zeroed on the way
    for (int i = gtid; i < U; i += gthreads) {
      const uint32_t c = X.collect_uniq[(int)(__ldcg(G.keys + i) & ((1ull << exact::kIndexBits) - 1ull))];
      const uint32_t count = X.collect_table[c];
      X.collect_table[c] = 0u;  // the count table is all-zero again when the call ends
      G.colour[i] = c;
      G.w[i] = fmul(A.norm, (double)(int)count);  // weights[i] = weight * count (:185)
      G.idx[0][i] = (uint32_t)i;
      A.pts[0][i] = make_uint2(c, count);
    }
    grid_barrier2(A, bar_target, X.progress);
  }

  if (b != 0) {
    // ================= worker =================
    long long t0 = clock64();
    unsigned polls = 0;
    for (;;) {  // wait for the root
      if (tid == 0) s_flag = ld_relaxed_u32(G.counters + 1) | (ld_relaxed_u32(G.counters + 2) << 1) | (ld_relaxed_u32(A.ctl + kCtlError) << 2);
      __syncthreads();
      const uint32_t f = s_flag;
      __syncthreads();
      if (f & 6u) return;  // done or error
      if (f & 1u) break;
      if (spin_expired(A, t0, polls, 5, b)) return;
      __nanosleep(200);
    }
    __threadfence();
    for (;;) {
      if (tid == 0) {
        // leave the last 2K (+ a racing claim per CTA) nodes to the controller: it must never find the table full
        const bool room = (uint64_t)ld_relaxed_u32(G.counters) + 2u + 2u * (uint32_t)K + 2u * gridDim.x <= A.node_cap;
        const uint32_t stop = ld_relaxed_u32(G.counters + 2) | ld_relaxed_u32(A.ctl + kCtlError);
        s_flag = (stop ? 1u : 0u) | (room ? 2u : 0u);
      }
      __syncthreads();
      const uint32_t f = s_flag;
      __syncthreads();
      if (f & 1u) return;
      const int cand = (f & 2u) ? pick_leaf(A, G, K) : -1;
      if (cand < 0) {
        __nanosleep(500);
        continue;
      }
      if (tid == 0) s_flag = (atomicCAS(G.state + cand, kLeaf, kBusy) == kLeaf) ? 1u : 0u;
      __syncthreads();
      const uint32_t mine = s_flag;
      __syncthreads();
      if (!mine) continue;
      __threadfence();
      if (!split_node(A, S, G, cand)) {
        // node table full: give the leaf back, the controller splits what it needs itself into the reserved tail
        if (tid == 0) st_u32(G.state + cand, kLeaf);
        return;
      }
    }
  }

  // ================= controller (CTA 0) =================
  Points P;
  P.keys = G.keys, P.w = G.w, P.colour = G.colour, P.member = nullptr, P.cur = G.idx[0];
  P.cur_shared_across_ctas = true;
  P.in_global = true;
  // cluster arrays of the reference (:296-324): cluster -> node and tse[]
  double *ctse = (K <= exact::kSmemColors) ? S.k_tse : X.exact_f64;
  int32_t *cnode = (K <= exact::kSmemColors) ? S.k_size : X.exact_i32;
  for (int i = tid; i < K; i += T) ctse[i] = 0.0, cnode[i] = 0;
  __syncthreads();
  // ---- DivQuantClusterInitMeanAndVar (:60-104) ----
  {
    unsigned no_prev[kMaxChunks];
#pragma unroll
    for (int i = 0; i < kMaxChunks; ++i) no_prev[i] = 0xFFFFFFFFu;
    exact::pass_sums<T, false>(S, P, U, 7, no_prev, [](uint32_t) { return true; }, [](int, bool) {});
  }
  if (tid == 0) {
    SplitNode root;
    root.tw = 1.0;  // weight[0] = 1.0 (:343)
    for (int c = 0; c < 3; ++c) {
      root.tm[c] = S.chain[c];
      root.tv[c] = fsub(S.chain[4 + c], fsq(S.chain[c]));
    }
    root.tse = 0.0, root.cut = 0.0;
    root.begin = 0, root.size = (uint32_t)U, root.buf = 0, root.child = -1, root.axis = 0, root.parent = -1;
    store_node(A.nodes, 0, root);
    __threadfence();
    st_u32(G.counters, 1u);      // nodes allocated
    st_u32(G.state, kLeaf);      // the root may be claimed
    __threadfence();
    st_u32(G.counters + 1, 1u);  // ready
  }
  __syncthreads();

  // ---- the reference's sequence (:333-1019) over the tree ----
  int old_index = 0;
  bool failed = false;
  for (int new_index = 1; new_index < K && !failed; ++new_index) {
    const int node = cnode[old_index];
    long long t0 = clock64();
    unsigned polls = 0;
    for (;;) {  // make sure `node` is split
      if (tid == 0) {
        uint32_t st = ld_relaxed_u32(G.state + node);
        if (st == kLeaf && atomicCAS(G.state + node, kLeaf, kBusy) == kLeaf) st = 100u;  // ours
        s_flag = st;
      }
      __syncthreads();
      const uint32_t st = s_flag;
      __syncthreads();
      if (st == kSplitDone) break;
      if (st == 100u) {
        __threadfence();
        if (!split_node(A, S, G, node)) {
          if (tid == 0) atomicCAS(A.ctl + kCtlError, 0u, 8u);  // node table full
          failed = true;
        }
        break;
      }
      if (spin_expired(A, t0, polls, 6, node)) {
        failed = true;
        break;
      }
    }
    if (failed) break;
    __threadfence();
    __shared__ int s_child2;
    __shared__ int s_steady;
    if (tid == 0) {
      const SplitNode nd = load_node_cg(A.nodes, node);
      const int child = nd.child;
      s_child2 = child;
      cnode[old_index] = child;
      cnode[new_index] = child + 1;
      const SplitNode o = load_node_cg(A.nodes, child), n = load_node_cg(A.nodes, child + 1);
      {
        // steady state (see dq_split_exact.cuh): nothing went to the new side and the old side's statistics are the
        // parent's bit for bit -- if the same cluster is up again, every split that is left repeats this one
        bool same = n.size == 0u && exact::same_input(o.tw, nd.tw);
        for (int c = 0; c < 3; ++c) same = same && exact::same_input(o.tm[c], nd.tm[c]) && exact::same_input(o.tv[c], nd.tv[c]);
        s_steady = same ? 1 : 0;
      }
      if (new_index < K - 1) {  // the last split leaves tse[] alone (:823-832)
        ctse[old_index] = o.tse;
        ctse[new_index] = n.tse;
      }
      if (A.records != nullptr) {
        SplitRecord r;
        r.new_index = new_index, r.old_index = old_index, r.cut_axis = nd.axis, r.num_points = (int32_t)nd.size;
        r.new_size = (int32_t)n.size, r.is_last = (new_index == K - 1);
        r.cut_pos = nd.cut, r.total_weight = nd.tw, r.new_weight = n.tw, r.old_weight = o.tw;
        for (int c = 0; c < 3; ++c) {
          r.new_mean[c] = n.tm[c], r.old_mean[c] = o.tm[c];
          r.new_var[c] = r.is_last ? 0.0 : n.tv[c], r.old_var[c] = r.is_last ? 0.0 : o.tv[c];
        }
        r.new_tse = r.is_last ? 0.0 : n.tse, r.old_tse = r.is_last ? 0.0 : o.tse;
        A.records[new_index - 1] = r;
      }
    }
    __syncthreads();
    if (new_index == K - 1) break;
    // next cluster: strictly-greater scan seeded with DBL_MIN; stale old_index otherwise (:876-887)
    {
      double best = DBL_MIN;
      int best_i = -1;
      for (int ic = tid; ic <= new_index; ic += T) {
        const double t = ctse[ic];
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
        for (int q = 1; q < T / 32; ++q) {
          const double ob = S.red_val[q];
          const int oi = S.red_idx[q];
          if (oi >= 0 && (best_i < 0 || best < ob || (ob == best && oi < best_i))) best = ob, best_i = oi;
        }
        S.old_index = (best_i >= 0) ? best_i : old_index;
      }
      __syncthreads();
      const bool repeats = s_steady != 0 && S.old_index == old_index;
      old_index = S.old_index;
      __syncthreads();
      if (repeats) {
        // the remaining clusters are all this split's empty new side; the old cluster keeps the old child (same statistics)
        for (int ni = new_index + 1 + tid; ni < K; ni += T) {
          cnode[ni] = cnode[new_index];
          if (ni < K - 1) ctse[ni] = ctse[new_index];
          if (A.records != nullptr) {
            SplitRecord r = A.records[new_index - 1];
            r.new_index = ni;
            r.is_last = (ni == K - 1);
            if (r.is_last) {
              for (int c = 0; c < 3; ++c) r.new_var[c] = r.old_var[c] = 0.0;
              r.new_tse = r.old_tse = 0.0;
            }
            A.records[ni - 1] = r;
          }
        }
        __syncthreads();
        break;
      }
    }
  }
  if (tid == 0) {
    __threadfence();
    st_u32(G.counters + 2, 1u);  // done: workers leave after their current split
  }
  if (failed) return;

  // ---- palette = rounded means of the non-empty clusters in index order (:1030-1065) ----
  if (tid == 0) S.emitted = 0;
  __syncthreads();
  for (int base = 0; base < K; base += T) {
    const int ic = base + tid;
    uint32_t colour = 0;
    int sz = 0;
    if (ic < K) {
      double m[3] = {0.0, 0.0, 0.0};  // K == 1 never assigns mean[0]
      sz = U;
      if (K > 1) {
        const SplitNode nd = load_node_cg(A.nodes, cnode[ic]);
        sz = (int)nd.size;
        m[0] = nd.tm[0], m[1] = nd.tm[1], m[2] = nd.tm[2];
      }
      if (sz > 0) {
        const uint32_t Rr = (__double2uint_rz(fadd(m[0], 0.5)) & 0xFFu) << A.shift;
        const uint32_t Gg = (__double2uint_rz(fadd(m[1], 0.5)) & 0xFFu) << A.shift;
        const uint32_t Bb = (__double2uint_rz(fadd(m[2], 0.5)) & 0xFFu) << A.shift;
        colour = (Rr << 16) | (Gg << 8) | Bb;
      }
      A.cluster_size[ic] = (uint32_t)sz;
      for (int c = 0; c < 3; ++c) A.cluster_mean[3 * ic + c] = m[c];
    }
    int first, total;
    bool unused;
    exact::group_ranks<T, false>(S, sz > 0 ? 1 : 0, false, first, total, unused);
    if (sz > 0) A.palette[S.emitted + first] = colour;
    __syncthreads();
    if (tid == 0) S.emitted += total;
    __syncthreads();
  }
  if (tid == 0) {
    A.result[0] = (uint32_t)S.emitted;
    A.result[1] = (uint32_t)(K - S.emitted);
    A.ctl[kCtlDone] = 1;
    A.ctl[kCtlRounds] = (uint32_t)(K - 1);
    A.ctl[kCtlSplits] = ld_relaxed_u32(G.counters) / 2u;
  }
}

}  // namespace ordered
