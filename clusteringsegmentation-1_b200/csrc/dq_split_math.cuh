// dq_split_math.cuh -- scalar formulas of the divisive phase shared by the split kernels.
// Every expression follows the reference's operation order (DivQuantCluster.cpp, lines cited) and
// uses the never-contracted IEEE helpers of dq_common.cuh.
#pragma once

#include "dq_split.cuh"

namespace dq {

struct Means {
  double nw, ow;
  double nm[3], om[3];
};

// new/old weights and centres from the integer sums of a pass
// (:561-581 after the split pass, :780-810 after an LKM iteration; uniform-weight form).
__device__ __forceinline__ void derive_means(double tw, const double *tm, double norm, uint64_t cnt, uint64_t sr,
                                             uint64_t sg, uint64_t sb, Means &m) {
  m.nw = fmul(__ull2double_rn(cnt), norm);
  m.nm[0] = fdiv(fmul(__ull2double_rn(sr), norm), m.nw);
  m.nm[1] = fdiv(fmul(__ull2double_rn(sg), norm), m.nw);
  m.nm[2] = fdiv(fmul(__ull2double_rn(sb), norm), m.nw);
  m.ow = fsub(tw, m.nw);
#pragma unroll
  for (int c = 0; c < 3; ++c) m.om[c] = fdiv(fsub(fmul(tw, tm[c]), fmul(m.nw, m.nm[c])), m.ow);
}

// What every thread needs to classify a point in the current pass.
struct PassParams {
  double a;     // pass 0: cut position; later: lhs (:616-619)
  double r[3];  // rhs = old_mean - new_mean (:621-623)
  int32_t axis;
  int32_t buf;
  uint32_t begin;
  uint32_t size;
};

// lhs / rhs of the hyperplane test from the two centres (:616-623)
__device__ __forceinline__ void hyperplane(const Means &m, PassParams &pp) {
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

__device__ __forceinline__ bool goes_new(const PassParams &pp, bool split_pass, uint32_t colour) {
  const uint32_t R = (colour >> 16) & 0xFFu, G = (colour >> 8) & 0xFFu, B = colour & 0xFFu;
  if (split_pass) {
    const uint32_t ch = (pp.axis == 0) ? R : ((pp.axis == 1) ? G : B);
    return pp.a < (double)ch;  // (:473)
  }
  const double dot = fadd(fadd(fmul(pp.r[0], (double)R), fmul(pp.r[1], (double)G)), fmul(pp.r[2], (double)B));
  return !(pp.a < dot);  // (:683) -- false on NaN, exactly like the reference's else branch
}

// axis of greatest variance, cut at its mean (:388-403)
__device__ __forceinline__ void choose_cut(const double *tv, const double *tm, int &axis, double &cut) {
  double max_val = tv[0];
  axis = 0;
  cut = tm[0];
  if (max_val < tv[1]) {
    max_val = tv[1];
    axis = 1;
    cut = tm[1];
  }
  if (max_val < tv[2]) {
    axis = 2;
    cut = tm[2];
  }
}

// The two children of a split from the final sums of its last pass (:800-871).
// sums = {cnt, R, G, B, npts, RR, GG, BB}.
__device__ __forceinline__ void make_children(const SplitNode &parent, int parent_id, int child0, double norm,
                                              const uint64_t *sums, SplitNode &o, SplitNode &n) {
  Means m;
  derive_means(parent.tw, parent.tm, norm, sums[kAccCnt], sums[kAccR], sums[kAccG], sums[kAccB], m);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    // new side: sum(w x^2)/sum(w) - mean^2 (:836-838); old side: combined variance (:844-855)
    const double sq_sum = fmul(__ull2double_rn(sums[kAccRR + c]), norm);
    n.tv[c] = fsub(fdiv(sq_sum, m.nw), fsq(m.nm[c]));
    o.tv[c] = fsub(fdiv(fsub(fmul(parent.tw, parent.tv[c]), fmul(m.nw, fadd(n.tv[c], fsq(fsub(m.nm[c], parent.tm[c]))))),
                        m.ow),
                   fsq(fsub(m.om[c], parent.tm[c])));
    n.tm[c] = m.nm[c];
    o.tm[c] = m.om[c];
  }
  o.tse = fmul(m.ow, fadd(fadd(o.tv[0], o.tv[1]), o.tv[2]));  // (:871)
  n.tse = fmul(m.nw, fadd(fadd(n.tv[0], n.tv[1]), n.tv[2]));
  o.tw = m.ow;
  n.tw = m.nw;
  o.cut = n.cut = 0.0;
  o.axis = n.axis = 0;
  o.parent = n.parent = parent_id;
  o.child = n.child = -1;
  o.buf = n.buf = parent.buf ^ 1;
  const uint32_t size_new = (uint32_t)sums[kAccPts];
  o.begin = parent.begin;
  o.size = parent.size - size_new;  // size[old] = tmp_num_points - new_size (:819)
  n.begin = parent.begin + o.size;
  n.size = size_new;
  (void)child0;
}

}  // namespace dq
