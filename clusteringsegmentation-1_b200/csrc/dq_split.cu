// dq_split.cu -- the divisive phase as ONE persistent, cooperatively launched sm_100a kernel.
//
// Reference: DivQuantClusterInitMeanAndVar / DivQuantCluster, DivQuant/DivQuantCluster.cpp:49-1097.
// See dq_split.cuh for the formulation (speculative level-synchronous tree + scalar controller).
//
// Pass structure of one round (P = max_iters):
//   pass 0      split test   cut_pos < c[axis]                       (:438-559)
//   pass 1..P   LKM test     !(lhs < rhs . c)  -> "new" side         (:613-811)
//               pass P also accumulates count*c*c                    (:735-745)
//   partition   scatter every split cluster's points into [old | new] segments of the other buffer
// Every pass reduces, per split cluster, exact u64 sums {count, count*R, count*G, count*B, #points
// [, count*R^2, count*G^2, count*B^2]} with warp shuffles + one global atomic per CTA and word; the
// parameters of the next pass are re-derived from those sums by every CTA that needs them, so a
// pass costs exactly one grid barrier.
#include "dq_split_math.cuh"
#include <atomic>

#include <cfloat>

namespace dq {
namespace {

struct CtaShared {
  uint64_t red[32][kAccWords];
  uint32_t num_points;
  PassParams pp;
  uint32_t npts_new;  // partition pass: size of the "new" child
  int32_t cached_job;
  // controller-only (CTA 0)
  int32_t new_index, old_index;
  int32_t node_count;
  int32_t prev_jobs;
  int32_t njobs, ncand, budget, scan_carry;
  int32_t warp_tmp[32];
};

// Block-wide sum of `words` u64 values per thread, result added atomically to dst[0..words).
template <int WORDS>
__device__ __forceinline__ void block_accumulate(CtaShared &S, const uint64_t (&v)[kAccWords], uint64_t *dst) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int w = 0; w < WORDS; ++w) {
    uint64_t s = warp_sum_u64(v[w]);
    if (lane == 0) S.red[warp][w] = s;
  }
  __syncthreads();
  if (threadIdx.x < WORDS * 32) {
    // warp w reduces word w across the 32 per-warp partials
    const int w = threadIdx.x >> 5;
    uint64_t s = warp_sum_u64(S.red[lane][w]);
    if (lane == 0 && s != 0) atomicAdd(reinterpret_cast<unsigned long long *>(dst + w), (unsigned long long)s);
  }
  __syncthreads();
}

// In-place exclusive scan of a[0..n) held in shared memory; returns the total in S.scan_carry.
__device__ void block_exclusive_scan(CtaShared &S, uint32_t *a, int n) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int per = (n + kSplitThreads - 1) / kSplitThreads;
  const int lo = min(tid * per, n), hi = min(lo + per, n);
  uint32_t sum = 0;
  for (int i = lo; i < hi; ++i) sum += a[i];
  uint32_t incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) S.warp_tmp[warp] = (int32_t)incl;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = (uint32_t)S.warp_tmp[lane], wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    S.warp_tmp[lane] = (int32_t)(wi - w);
    if (lane == 31) S.scan_carry = (int32_t)wi;
  }
  __syncthreads();
  uint32_t run = (uint32_t)S.warp_tmp[warp] + (incl - sum);
  for (int i = lo; i < hi; ++i) {
    uint32_t t = a[i];
    a[i] = run;
    run += t;
  }
  __syncthreads();
}

// Controller arrays (shared memory of CTA 0, or global memory for very large K).
struct CtlArrays {
  int32_t *cluster_node;  // [K]    node currently standing for cluster ic
  double *cluster_tse;    // [K]    tse[] of the reference (:310)
  int32_t *node_child;    // [node_cap]
  double *node_tse;       // [node_cap]
  uint32_t *cand;         // [K]    scratch
  uint32_t *tiles;        // [K+1]  scratch: tiles per job -> tile0
};

__device__ __forceinline__ void trace(const SplitArgs &A, int tag, int arg) {
  if (A.timeline != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
    const unsigned long long n = A.timeline[0];
    if (2 * n + 3 < A.timeline_cap) {
      A.timeline[1 + 2 * n] = ((unsigned long long)tag << 32) | (unsigned)arg;
      A.timeline[2 + 2 * n] = clock64();
      A.timeline[0] = n + 1;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Controller, executed by CTA 0 between rounds.
// ---------------------------------------------------------------------------------------------
__device__ void controller(const SplitArgs &A, CtaShared &S, const CtlArrays &C, int round) {
  const int tid = threadIdx.x, lane = tid & 31;
  const int K = (int)A.num_colors, P = A.max_iters;
  const int prev_par = (round - 1) & 1, par = round & 1;
  const size_t acc_job_stride = (size_t)(P + 1) * kAccWords;

  // ---- phase A: turn last round's sums into child nodes (:800-871) ----
  if (round == 0) {
    if (tid == 0) {
      SplitNode root;
      root.tw = 1.0;  // weight[0] = 1.0 (:343)
      double mean[3], var[3];
      for (int c = 0; c < 3; ++c) {
        mean[c] = fmul(__ull2double_rn(ld_cg_u64(A.root_acc + kAccR + c)), A.norm);  // (:107-112)
        var[c] = fsub(fmul(__ull2double_rn(ld_cg_u64(A.root_acc + kAccRR + c)), A.norm), fsq(mean[c]));
        root.tm[c] = mean[c];
        root.tv[c] = var[c];
      }
      root.tse = 0.0;
      root.cut = 0.0;
      root.begin = 0;
      root.size = S.num_points;
      root.buf = 0;
      root.child = -1;
      root.axis = 0;
      root.parent = -1;
      A.nodes[0] = root;
      C.node_child[0] = -1;
      C.node_tse[0] = 0.0;
      C.cluster_node[0] = 0;
      C.cluster_tse[0] = 0.0;
      S.new_index = 1;
      S.old_index = 0;
      S.node_count = 1;
      S.prev_jobs = 0;
    }
  } else {
    const SplitJob *jobs = A.jobs + (size_t)prev_par * K;
    const uint64_t *acc = A.acc + (size_t)prev_par * K * acc_job_stride;
    for (int j = tid; j < S.prev_jobs; j += kSplitThreads) {
      const SplitJob job = jobs[j];
      const uint64_t *a = acc + (size_t)j * acc_job_stride + (size_t)P * kAccWords;
      const uint64_t cnt = ld_cg_u64(a + kAccCnt), npts = ld_cg_u64(a + kAccPts);
      Means m;
      derive_means(job.tw, job.tm, A.norm, cnt, ld_cg_u64(a + kAccR), ld_cg_u64(a + kAccG), ld_cg_u64(a + kAccB), m);
      const SplitNode parent = A.nodes[job.node];
      SplitNode o, n;
      double tse_o = 0.0, tse_n = 0.0;
      for (int c = 0; c < 3; ++c) {
        // new side: sum(w x^2)/sum(w) - mean^2 (:836-838); old side: combined variance (:844-855)
        const double sq_sum = fmul(__ull2double_rn(ld_cg_u64(a + kAccRR + c)), A.norm);
        n.tv[c] = fsub(fdiv(sq_sum, m.nw), fsq(m.nm[c]));
        o.tv[c] = fsub(fdiv(fsub(fmul(job.tw, parent.tv[c]),
                                 fmul(m.nw, fadd(n.tv[c], fsq(fsub(m.nm[c], job.tm[c]))))),
                            m.ow),
                       fsq(fsub(m.om[c], job.tm[c])));
        n.tm[c] = m.nm[c];
        o.tm[c] = m.om[c];
      }
      tse_o = fmul(m.ow, fadd(fadd(o.tv[0], o.tv[1]), o.tv[2]));  // (:871)
      tse_n = fmul(m.nw, fadd(fadd(n.tv[0], n.tv[1]), n.tv[2]));
      o.tw = m.ow;
      n.tw = m.nw;
      o.tse = tse_o;
      n.tse = tse_n;
      o.cut = n.cut = 0.0;
      o.axis = n.axis = 0;
      o.parent = n.parent = job.node;
      o.child = n.child = -1;
      o.buf = n.buf = job.buf ^ 1;
      const uint32_t size_new = (uint32_t)npts;
      o.begin = job.begin;
      o.size = job.size - size_new;  // size[old] = tmp_num_points - new_size (:819)
      n.begin = job.begin + o.size;
      n.size = size_new;
      A.nodes[job.child0] = o;
      A.nodes[job.child0 + 1] = n;
      A.nodes[job.node].child = job.child0;
      C.node_child[job.node] = job.child0;
      C.node_child[job.child0] = -1;
      C.node_child[job.child0 + 1] = -1;
      C.node_tse[job.child0] = tse_o;
      C.node_tse[job.child0 + 1] = tse_n;
    }
  }
  __syncthreads();
  trace(A, kTracePhaseA, S.prev_jobs);

  // ---- phase B: replay the reference's sequential selection over the cached splits ----
  if (tid < 32) {
    int new_index = S.new_index, old_index = S.old_index;
    while (new_index < K) {
      const int node = C.cluster_node[old_index];
      const int child = C.node_child[node];
      if (child < 0) break;  // this split has not been computed yet
      __syncwarp();
      if (lane == 0) {
        C.cluster_node[old_index] = child;
        C.cluster_node[new_index] = child + 1;
        if (A.records != nullptr) {
          const SplitNode p = A.nodes[node], o = A.nodes[child], n = A.nodes[child + 1];
          SplitRecord r;
          r.new_index = new_index;
          r.old_index = old_index;
          r.cut_axis = p.axis;
          r.num_points = (int32_t)p.size;
          r.new_size = (int32_t)n.size;
          r.is_last = (new_index == K - 1);
          r.cut_pos = p.cut;
          r.total_weight = p.tw;
          r.new_weight = n.tw;
          r.old_weight = o.tw;
          for (int c = 0; c < 3; ++c) {
            r.new_mean[c] = n.tm[c];
            r.old_mean[c] = o.tm[c];
            r.new_var[c] = n.tv[c];
            r.old_var[c] = o.tv[c];
          }
          r.new_tse = n.tse;
          r.old_tse = o.tse;
          A.records[new_index - 1] = r;
        }
      }
      if (new_index == K - 1) {  // last split: no TSE bookkeeping (:823-832)
        new_index = K;
        break;
      }
      if (lane == 0) {
        C.cluster_tse[old_index] = C.node_tse[child];
        C.cluster_tse[new_index] = C.node_tse[child + 1];
      }
      __syncwarp();
      // arg-max with DBL_MIN seed and strict '<': lowest index among equal maxima, and
      // old_index is left untouched when nothing exceeds DBL_MIN (:876-887)
      double best = DBL_MIN;
      int best_i = -1;
      for (int ic = lane; ic <= new_index; ic += 32) {
        const double t = C.cluster_tse[ic];
        if (best < t) {
          best = t;
          best_i = ic;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
        const bool take = (oi >= 0) && (best_i < 0 || best < ob || (ob == best && oi < best_i));
        if (take) {
          best = ob;
          best_i = oi;
        }
      }
      if (best_i >= 0) old_index = best_i;
      ++new_index;
    }
    if (lane == 0) {
      S.new_index = new_index;
      S.old_index = old_index;
    }
  }
  __syncthreads();

  const int new_index = S.new_index, old_index = S.old_index;
  trace(A, kTracePhaseB, new_index);
  if (new_index >= K) {
    // ---- finished: palette = rounded means of non-empty clusters in index order (:1030-1065) ----
    if (tid == 0) S.scan_carry = 0;
    __syncthreads();
    for (int base = 0; base < K; base += kSplitThreads) {
      const int ic = base + tid;
      uint32_t colour = 0, size = 0;
      if (ic < K) {
        double mean[3] = {0.0, 0.0, 0.0};  // K == 1 never assigns mean[0]
        if (K > 1) {
          const SplitNode nd = A.nodes[C.cluster_node[ic]];
          size = nd.size;
          mean[0] = nd.tm[0], mean[1] = nd.tm[1], mean[2] = nd.tm[2];
        } else {
          size = S.num_points;
        }
        if (size > 0) {
          const uint32_t R = (__double2uint_rz(fadd(mean[0], 0.5)) & 0xFFu) << A.shift;
          const uint32_t G = (__double2uint_rz(fadd(mean[1], 0.5)) & 0xFFu) << A.shift;
          const uint32_t B = (__double2uint_rz(fadd(mean[2], 0.5)) & 0xFFu) << A.shift;
          colour = (R << 16) | (G << 8) | B;
        }
        A.cluster_size[ic] = size;
        A.cluster_mean[3 * ic + 0] = mean[0];
        A.cluster_mean[3 * ic + 1] = mean[1];
        A.cluster_mean[3 * ic + 2] = mean[2];
      }
      const unsigned ballot = __ballot_sync(0xffffffffu, size > 0);
      if (lane == 0) S.warp_tmp[tid >> 5] = __popc(ballot);
      __syncthreads();
      int before = S.scan_carry;
      for (int w = 0; w < (tid >> 5); ++w) before += S.warp_tmp[w];
      if (size > 0) A.palette[before + __popc(ballot & ((1u << lane) - 1u))] = colour;
      __syncthreads();
      if (tid == 0) {
        int tot = 0;
        for (int w = 0; w < 32; ++w) tot += S.warp_tmp[w];
        S.scan_carry += tot;
      }
      __syncthreads();
    }
    if (tid == 0) {
      A.result[0] = (uint32_t)S.scan_carry;
      A.result[1] = (uint32_t)(K - S.scan_carry);
      A.ctl[kCtlDone] = 1;
      A.ctl[kCtlJobs] = 0;
      A.ctl[kCtlTiles] = 0;
      A.ctl[kCtlNodes] = (uint32_t)S.node_count;
      A.ctl[kCtlRounds] = (uint32_t)round;
    }
    return;
  }

  // ---- phase C: request the stalled split plus the leaves inside the remaining budget ----
  // The sequential process always takes the current max-TSE leaf, so whatever it will still split
  // in its remaining R = K - new_index steps is among the top-R current leaves (and descendants).
  if (tid == 0) {
    S.ncand = 0;
    S.njobs = 0;
    S.budget = (K - new_index) - 1;  // besides the mandatory one
  }
  __syncthreads();
  for (int ic = tid; ic < new_index; ic += kSplitThreads) {
    if (ic == old_index) continue;
    const int node = C.cluster_node[ic];
    if (C.node_child[node] >= 0) continue;
    if (!(C.cluster_tse[ic] > DBL_MIN)) continue;
    C.cand[atomicAdd(&S.ncand, 1)] = (uint32_t)ic;
  }
  __syncthreads();
  const int ncand = S.ncand, budget = S.budget;
  SplitJob *jobs = A.jobs + (size_t)par * K;
  auto make_job = [&](int ic) {
    const int node = C.cluster_node[ic];
    SplitNode nd = A.nodes[node];
    // axis of greatest variance, cut at its mean (:388-403)
    double max_val = nd.tv[0];
    int axis = 0;
    double cut = nd.tm[0];
    if (max_val < nd.tv[1]) {
      max_val = nd.tv[1];
      axis = 1;
      cut = nd.tm[1];
    }
    if (max_val < nd.tv[2]) {
      axis = 2;
      cut = nd.tm[2];
    }
    const int j = atomicAdd(&S.njobs, 1);
    SplitJob job;
    job.cut = cut;
    job.tw = nd.tw;
    job.tm[0] = nd.tm[0], job.tm[1] = nd.tm[1], job.tm[2] = nd.tm[2];
    job.node = node;
    job.axis = axis;
    job.begin = nd.begin;
    job.size = nd.size;
    job.buf = nd.buf;
    job.tile0 = 0;
    job.cur_old = 0;
    job.cur_new = 0;
    job.child0 = S.node_count + 2 * j;
    job.pad = 0;
    jobs[j] = job;
    A.nodes[node].axis = axis;
    A.nodes[node].cut = cut;
    C.tiles[j] = (nd.size + kSplitTile - 1) / kSplitTile;
  };
  if (tid == 0) make_job(old_index);
  for (int i = tid; i < ncand; i += kSplitThreads) {
    const uint32_t ic = C.cand[i];
    bool take = true;
    if (ncand > budget) {
      const double mine = C.cluster_tse[ic];
      int rank = 0;
      for (int q = 0; q < ncand; ++q) {
        const uint32_t oc = C.cand[q];
        const double t = C.cluster_tse[oc];
        rank += (t > mine) || (t == mine && oc < ic);
      }
      take = rank < budget;
    }
    if (take) make_job((int)ic);
  }
  __syncthreads();
  const int njobs = S.njobs;
  block_exclusive_scan(S, C.tiles, njobs);
  for (int j = tid; j < njobs; j += kSplitThreads) jobs[j].tile0 = C.tiles[j];
  if (tid == 0) {
    C.tiles[njobs] = (uint32_t)S.scan_carry;
    S.prev_jobs = njobs;
    S.node_count += 2 * njobs;
    A.ctl[kCtlJobs] = (uint32_t)njobs;
    A.ctl[kCtlTiles] = (uint32_t)S.scan_carry;
    A.ctl[kCtlDone] = 0;
    A.ctl[kCtlSplits] += (uint32_t)njobs;
    if ((uint32_t)S.node_count > A.node_cap) A.ctl[kCtlError] = 1;
  }
  __syncthreads();
  trace(A, kTracePhaseC, njobs);
}

// ---------------------------------------------------------------------------------------------
// The persistent kernel.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSplitThreads, 1) split_kernel(const SplitArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  CtaShared &S = *reinterpret_cast<CtaShared *>(smem_raw);
  const int K = (int)A.num_colors, P = A.max_iters;
  // shared-memory carve-up after CtaShared: tile0[K+1] for every CTA; controller arrays for CTA 0
  unsigned char *cursor = smem_raw + ((sizeof(CtaShared) + 15) & ~size_t(15));
  uint32_t *s_tile0 = reinterpret_cast<uint32_t *>(cursor);
  cursor += ((size_t)(K + 1) * 4 + 15) & ~size_t(15);
  CtlArrays C;
  C.tiles = s_tile0;
  if (A.use_smem_ctl) {
    C.cluster_tse = reinterpret_cast<double *>(cursor);
    cursor += (size_t)K * 8;
    C.node_tse = reinterpret_cast<double *>(cursor);
    cursor += (size_t)A.node_cap * 8;
    C.cluster_node = reinterpret_cast<int32_t *>(cursor);
    cursor += (size_t)K * 4;
    C.node_child = reinterpret_cast<int32_t *>(cursor);
    cursor += (size_t)A.node_cap * 4;
    C.cand = reinterpret_cast<uint32_t *>(cursor);
  } else {
    C.cluster_tse = A.g_cluster_tse;
    C.node_tse = A.g_cluster_tse + K;
    C.cluster_node = A.g_cluster_node;
    C.node_child = A.g_cluster_node + K;
    C.cand = reinterpret_cast<uint32_t *>(A.g_cluster_node + K + A.node_cap);
  }

  const int tid = threadIdx.x;
  const uint32_t U = A.num_points_dev ? ld_cg_u32(A.num_points_dev) : A.num_points;
  if (A.exact_small_max != 0u && U <= A.exact_small_max) return;  // split_exact_kernel has done this input
  if (tid == 0) S.num_points = U;
  unsigned int bar_target = 0;
  const size_t acc_job_stride = (size_t)(P + 1) * kAccWords;
  const size_t acc_par_words = (size_t)K * acc_job_stride;

  // ---- global statistics of all points (DivQuantClusterInitMeanAndVar, :60-104) ----
  {
    uint64_t v[kAccWords] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (size_t i = (size_t)blockIdx.x * kSplitThreads + tid; i < U; i += (size_t)gridDim.x * kSplitThreads) {
      const uint2 p = ld_cg_u2(A.pts[0] + i);
      const uint64_t cnt = p.y, R = (p.x >> 16) & 0xFFu, G = (p.x >> 8) & 0xFFu, B = p.x & 0xFFu;
      v[kAccCnt] += cnt;
      v[kAccR] += cnt * R;
      v[kAccG] += cnt * G;
      v[kAccB] += cnt * B;
      v[kAccPts] += 1;
      v[kAccRR] += cnt * (R * R);
      v[kAccGG] += cnt * (G * G);
      v[kAccBB] += cnt * (B * B);
    }
    block_accumulate<kAccWords>(S, v, A.root_acc);
    // both parities of the per-job accumulators start at zero
    for (size_t i = (size_t)blockIdx.x * kSplitThreads + tid; i < 2 * acc_par_words; i += (size_t)gridDim.x * kSplitThreads)
      A.acc[i] = 0;
  }
  grid_barrier(A.barrier, bar_target);
  trace(A, kTraceRoot, 0);

  for (int round = 0;; ++round) {
    trace(A, kTraceRoundBegin, round);
    if (blockIdx.x == 0) controller(A, S, C, round);
    grid_barrier(A.barrier, bar_target);
    trace(A, kTraceCtlBarrier, round);
    if (ld_cg_u32(A.ctl + kCtlDone) != 0 || ld_cg_u32(A.ctl + kCtlError) != 0) break;
    const int par = round & 1;
    const int njobs = (int)ld_cg_u32(A.ctl + kCtlJobs);
    const uint32_t ntiles = ld_cg_u32(A.ctl + kCtlTiles);
    const SplitJob *jobs = A.jobs + (size_t)par * K;
    uint64_t *acc = A.acc + (size_t)par * acc_par_words;
    if (blockIdx.x != 0) {  // CTA 0 already holds the tile prefix in shared memory
      for (int j = tid; j < njobs; j += kSplitThreads) s_tile0[j] = ld_cg_u32(&jobs[j].tile0);
      if (tid == 0) s_tile0[njobs] = ntiles;
    }
    if (tid == 0) S.cached_job = -1;
    // the other parity is used by the next round: clear it now
    {
      uint64_t *other = A.acc + (size_t)(par ^ 1) * acc_par_words;
      for (size_t i = (size_t)blockIdx.x * kSplitThreads + tid; i < acc_par_words; i += (size_t)gridDim.x * kSplitThreads)
        other[i] = 0;
    }
    __syncthreads();

    for (int pass = 0; pass <= P + 1; ++pass) {
      const bool partition = (pass == P + 1);
      // which previous pass defines the classification: the partition repeats the last LKM test
      const int src_pass = partition ? P - 1 : pass - 1;
      for (uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        // job of tile t: last j with tile0[j] <= t
        int lo = 0, hi = njobs;  // invariant: tile0[lo] <= t < tile0[hi]
        while (hi - lo > 1) {
          const int mid = (lo + hi) >> 1;
          if (s_tile0[mid] <= t) lo = mid; else hi = mid;
        }
        const int j = lo;
        if (tid == 0) {
          const SplitJob job = jobs[j];
          PassParams pp;
          pp.axis = job.axis;
          pp.buf = job.buf;
          pp.begin = job.begin;
          pp.size = job.size;
          if (src_pass < 0) {
            pp.a = job.cut;
            pp.r[0] = pp.r[1] = pp.r[2] = 0.0;
          } else {
            const uint64_t *a = acc + (size_t)j * acc_job_stride + (size_t)src_pass * kAccWords;
            Means m;
            derive_means(job.tw, job.tm, A.norm, ld_cg_u64(a + kAccCnt), ld_cg_u64(a + kAccR),
                         ld_cg_u64(a + kAccG), ld_cg_u64(a + kAccB), m);
            // lhs = 0.5*(|old|^2 - |new|^2) summed channel by channel, left to right (:616-619)
            double l = fsub(fsq(m.om[0]), fsq(m.nm[0]));
            l = fadd(l, fsq(m.om[1]));
            l = fsub(l, fsq(m.nm[1]));
            l = fadd(l, fsq(m.om[2]));
            l = fsub(l, fsq(m.nm[2]));
            pp.a = fmul(0.5, l);
            pp.r[0] = fsub(m.om[0], m.nm[0]);
            pp.r[1] = fsub(m.om[1], m.nm[1]);
            pp.r[2] = fsub(m.om[2], m.nm[2]);
          }
          S.pp = pp;
          if (partition) S.npts_new = (uint32_t)ld_cg_u64(acc + (size_t)j * acc_job_stride + (size_t)P * kAccWords + kAccPts);
        }
        __syncthreads();
        const PassParams pp = S.pp;
        const uint32_t off = (t - s_tile0[j]) * kSplitTile + tid;
        const bool valid = off < pp.size;
        uint2 p = make_uint2(0, 0);
        if (valid) p = ld_cg_u2(A.pts[pp.buf] + pp.begin + off);
        const bool to_new = valid && goes_new(pp, pass == 0, p.x);
        if (!partition) {
          uint64_t v[kAccWords] = {0, 0, 0, 0, 0, 0, 0, 0};
          if (to_new) {
            const uint64_t cnt = p.y, R = (p.x >> 16) & 0xFFu, G = (p.x >> 8) & 0xFFu, B = p.x & 0xFFu;
            v[kAccCnt] = cnt;
            v[kAccR] = cnt * R;
            v[kAccG] = cnt * G;
            v[kAccB] = cnt * B;
            v[kAccPts] = 1;
            if (pass == P) {
              v[kAccRR] = cnt * (R * R);
              v[kAccGG] = cnt * (G * G);
              v[kAccBB] = cnt * (B * B);
            }
          }
          uint64_t *dst = acc + (size_t)j * acc_job_stride + (size_t)pass * kAccWords;
          if (pass == P) block_accumulate<kAccWords>(S, v, dst);
          else block_accumulate<5>(S, v, dst);
        } else {
          // scatter into [old | new] segments of the other buffer; order inside a segment is free
          // because every later sum is an exact integer sum
          const uint32_t size_old = pp.size - S.npts_new;
          const unsigned m_new = __ballot_sync(0xffffffffu, to_new);
          const unsigned m_old = __ballot_sync(0xffffffffu, valid && !to_new);
          const int lane = tid & 31;
          uint32_t base_new = 0, base_old = 0;
          if (lane == 0) {
            SplitJob *jw = A.jobs + (size_t)par * K + j;
            if (m_new) base_new = atomicAdd(&jw->cur_new, (uint32_t)__popc(m_new));
            if (m_old) base_old = atomicAdd(&jw->cur_old, (uint32_t)__popc(m_old));
          }
          base_new = __shfl_sync(0xffffffffu, base_new, 0);
          base_old = __shfl_sync(0xffffffffu, base_old, 0);
          if (valid) {
            const unsigned below = (1u << lane) - 1u;
            const uint32_t dst = to_new ? pp.begin + size_old + base_new + __popc(m_new & below)
                                        : pp.begin + base_old + __popc(m_old & below);
            A.pts[pp.buf ^ 1][dst] = p;
          }
          __syncthreads();
        }
      }
      if (!partition) grid_barrier(A.barrier, bar_target);
      trace(A, partition ? kTracePartition : kTracePass, pass);
    }
  }
}

}  // namespace

size_t split_acc_words(uint32_t num_colors, int max_iters) {
  return (size_t)2 * num_colors * (size_t)(max_iters + 1) * kAccWords;
}

static size_t split_ctl_smem(uint32_t K, uint32_t node_cap) {
  return (size_t)K * 8 + (size_t)node_cap * 8 + (size_t)K * 4 + (size_t)node_cap * 4 + (size_t)K * 4 + 64;
}

SplitLaunch split_plan(int sm_count, uint32_t num_colors) {
  SplitLaunch plan;
  plan.grid = sm_count;  // one persistent CTA per SM
  const uint32_t node_cap = 4 * num_colors + 8;
  size_t base = ((sizeof(CtaShared) + 15) & ~size_t(15)) + (((size_t)(num_colors + 1) * 4 + 15) & ~size_t(15));
  size_t with_ctl = base + split_ctl_smem(num_colors, node_cap);
  plan.smem_bytes = (with_ctl <= 200 * 1024) ? with_ctl : base;
  return plan;
}

void split_launch(const SplitArgs &args_in, const SplitLaunch &plan, cudaStream_t stream) {
  SplitArgs args = args_in;
  const uint32_t K = args.num_colors;
  size_t base = ((sizeof(CtaShared) + 15) & ~size_t(15)) + (((size_t)(K + 1) * 4 + 15) & ~size_t(15));
  args.use_smem_ctl = plan.smem_bytes > base ? 1 : 0;
  DQ_RAISE_SMEM(split_kernel, plan.smem_bytes);
  void *kargs[] = {(void *)&args};
  DQ_CUDA_CHECK(cudaLaunchCooperativeKernel((const void *)split_kernel, dim3(plan.grid), dim3(kSplitThreads), kargs,
                                            plan.smem_bytes, stream));
}

}  // namespace dq
