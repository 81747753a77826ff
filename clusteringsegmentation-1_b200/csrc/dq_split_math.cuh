// dq_split_math.cuh -- scalar formulas of the divisive phase shared by the split kernels.
// Every expression follows the reference's operation order (DivQuantCluster.cpp, lines cited) and
// uses the never-contracted IEEE helpers of dq_common.cuh.
#pragma once

#include "dq_split.cuh"
#include "dq_tie.cuh"

namespace dq {

// Exact u8 -> double without the (slow, XU-pipe) I2F.F64: 2^52 + x has x in its low mantissa bits.
__device__ __forceinline__ double byte_to_double(uint32_t x) {
  return __dsub_rn(__hiloint2double(0x43300000, (int)x), 4503599627370496.0);
}

// Exact integer -> double for 0 <= x < 2^52 (sums of a pass are below 2^48) and back, again without
// the conversion pipe: one FP64 add each way.
__device__ __forceinline__ double u52_to_double(uint64_t x) {
  return __dsub_rn(__longlong_as_double((long long)(0x4330000000000000ull | x)), 4503599627370496.0);
}
__device__ __forceinline__ uint64_t double_to_u52(double x) {  // x a non-negative integer below 2^52
  return (uint64_t)__double_as_longlong(__dadd_rn(x, 4503599627370496.0)) & 0x000FFFFFFFFFFFFFull;
}

struct Means {
  double nw, ow;
  double nm[3], om[3];
};

// new/old weights and centres from the integer sums of a pass
// (:561-581 after the split pass, :780-810 after an LKM iteration; uniform-weight form).
__device__ __forceinline__ void derive_means(double tw, const double *tm, double norm, uint64_t cnt, uint64_t sr,
                                             uint64_t sg, uint64_t sb, Means &m) {
  m.nw = fmul(u52_to_double(cnt), norm);
  m.nm[0] = fdiv(fmul(u52_to_double(sr), norm), m.nw);
  m.nm[1] = fdiv(fmul(u52_to_double(sg), norm), m.nw);
  m.nm[2] = fdiv(fmul(u52_to_double(sb), norm), m.nw);
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
  // tie audit: high word (+1) of the tolerance tie::PassExt::tol -- a point whose test value is within tol of `a` is
  // inside the reference's rounding noise; negative when the audit is off
  int32_t tol_hi;
  int32_t pad;
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

// A point's channels as exact doubles for the hyperplane test (3 conversions + 6 FP64 operations per point).
struct PointD {
  double r, g, b;
};
__device__ __forceinline__ PointD to_point(uint2 p) {
  PointD d;
  d.r = byte_to_double((p.x >> 16) & 0xFFu);
  d.g = byte_to_double((p.x >> 8) & 0xFFu);
  d.b = byte_to_double(p.x & 0xFFu);
  return d;
}
__device__ __forceinline__ bool goes_new(const PassParams &pp, bool split_pass, const PointD &d) {
  if (split_pass) {
    const double ch = (pp.axis == 0) ? d.r : ((pp.axis == 1) ? d.g : d.b);
    return pp.a < ch;  // (:473)
  }
  const double dot = fadd(fadd(fmul(pp.r[0], d.r), fmul(pp.r[1], d.g)), fmul(pp.r[2], d.b));
  return !(pp.a < dot);  // (:683)
}
template <bool SPLIT>
__device__ __forceinline__ bool goes_new_t(const PassParams &pp, const PointD &d) {
  if (SPLIT) {
    const double ch = (pp.axis == 0) ? d.r : ((pp.axis == 1) ? d.g : d.b);
    return pp.a < ch;  // (:473)
  }
  const double dot = fadd(fadd(fmul(pp.r[0], d.r), fmul(pp.r[1], d.g)), fmul(pp.r[2], d.b));
  return !(pp.a < dot);  // (:683)
}
// Same decision, plus the tie audit: `near` is raised when the test value lies within pp.tol of the threshold.
// x - a has the sign of the comparison (a difference of doubles is zero only for equal operands), so one subtraction
// serves both; a NaN on either side compares false everywhere, like the reference's else branch.
template <bool SPLIT>
__device__ __forceinline__ bool goes_new_audit(const PassParams &pp, const tie::PassExt *ext, const PointD &d, bool &near) {
  double x;
  if (SPLIT) {
    x = (pp.axis == 0) ? d.r : ((pp.axis == 1) ? d.g : d.b);
  } else {
    x = fadd(fadd(fmul(pp.r[0], d.r), fmul(pp.r[1], d.g)), fmul(pp.r[2], d.b));
  }
  const double m = fsub(x, pp.a);
  // The sign and the magnitude class of m are read off its high word on the integer pipe (which idles next to the FP64
  // pipe here): m > 0 <=> hi > 0 (a non-zero difference of these operands is never below 2^-1022), and
  // |m| <= tol  =>  (hi & 0x7fffffff) <= hi(tol), a filter that passes a handful of points per frame; the exact test
  // and the bound for the very point follow only for those.  tol < 0 (audit off) has a negative high word: never passes.
  const int hi = __double2hiint(m);
  near = near || ((hi & 0x7fffffff) <= pp.tol_hi);  // the filter only: recheck_near settles it, off the hot loop
  (void)ext;
  return SPLIT ? (hi > 0) : !(hi > 0);
}
// Cold path of the audit: a point that passed the integer filter, tested for real.
template <bool SPLIT>
__device__ __forceinline__ bool recheck_near(const PassParams &pp, const tie::PassExt *ext, const PointD &d) {
  double x;
  if (SPLIT) {
    x = (pp.axis == 0) ? d.r : ((pp.axis == 1) ? d.g : d.b);
  } else {
    x = fadd(fadd(fmul(pp.r[0], d.r), fmul(pp.r[1], d.g)), fmul(pp.r[2], d.b));
  }
  const double m = fabs(fsub(x, pp.a));
  if (!(m <= ext->tol)) return false;
  return SPLIT ? true : (m <= tie::point_tol(*ext, d.r, d.g, d.b));
}
// Accumulation stays on the integer pipe (IMAD.WIDE), which runs next to the FP64 pipe that classifies:
// exact u64 sums of count*c and count*c*c.
struct AccD {
  uint64_t cnt, r, g, b, rr, gg, bb;
  uint32_t n;
};
__device__ __forceinline__ AccD acc_zero() {
  AccD a;
  a.cnt = a.r = a.g = a.b = a.rr = a.gg = a.bb = 0ull;
  a.n = 0;
  return a;
}
__device__ __forceinline__ void acc_add(AccD &a, uint2 p, bool with_squares) {
  const uint32_t R = (p.x >> 16) & 0xFFu, G = (p.x >> 8) & 0xFFu, B = p.x & 0xFFu, c = p.y;
  a.cnt += c;
  a.r += (uint64_t)c * R;
  a.g += (uint64_t)c * G;
  a.b += (uint64_t)c * B;
  a.n += 1;
  if (with_squares) {
    a.rr += (uint64_t)c * (R * R);
    a.gg += (uint64_t)c * (G * G);
    a.bb += (uint64_t)c * (B * B);
  }
}
__device__ __forceinline__ void acc_words(const AccD &a, uint64_t (&v)[kAccWords]) {
  v[kAccCnt] = a.cnt;
  v[kAccR] = a.r;
  v[kAccG] = a.g;
  v[kAccB] = a.b;
  v[kAccPts] = a.n;
  v[kAccRR] = a.rr;
  v[kAccGG] = a.gg;
  v[kAccBB] = a.bb;
}

__device__ __forceinline__ bool goes_new(const PassParams &pp, bool split_pass, uint32_t colour) {
  const uint32_t R = (colour >> 16) & 0xFFu, G = (colour >> 8) & 0xFFu, B = colour & 0xFFu;
  if (split_pass) {
    const uint32_t ch = (pp.axis == 0) ? R : ((pp.axis == 1) ? G : B);
    return pp.a < byte_to_double(ch);  // (:473)
  }
  const double dot = fadd(fadd(fmul(pp.r[0], byte_to_double(R)), fmul(pp.r[1], byte_to_double(G))),
                          fmul(pp.r[2], byte_to_double(B)));
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
                                              const uint64_t *sums, SplitNode &o, SplitNode &n, bool audit = false) {
  Means m;
  derive_means(parent.tw, parent.tm, norm, sums[kAccCnt], sums[kAccR], sums[kAccG], sums[kAccB], m);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    // new side: sum(w x^2)/sum(w) - mean^2 (:836-838); old side: combined variance (:844-855)
    const double sq_sum = fmul(u52_to_double(sums[kAccRR + c]), norm);
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
  o.eW = o.eS = o.eQ = o.eM = o.eV = o.eT = n.eW = n.eS = n.eQ = n.eM = n.eV = n.eT = 0.0;
  o.tie = n.tie = 0u;
  o.pad = n.pad = 0u;
  if (audit) {
    const tie::PassErr fe = tie::pass_err_fast(parent.eW, parent.eS, parent.tw, tie::max3abs(parent.tm), m.nw, tie::max3abs(m.nm), m.ow,
                                               tie::max3abs(m.om), (double)size_new);
    tie::Bounds bp, bn, bo;
    bp.eW = parent.eW, bp.eS = parent.eS, bp.eQ = parent.eQ;
    tie::child_bounds_fast(bp, fe, parent.tw, m.nw, m.nm, n.tv, n.tse, m.ow, m.om, o.tv, o.tse, (double)size_new, bn, bo);
    n.eW = bn.eW, n.eS = bn.eS, n.eQ = bn.eQ, n.eM = bn.eM, n.eV = bn.eV, n.eT = bn.eT;
    o.eW = bo.eW, o.eS = bo.eS, o.eQ = bo.eQ, o.eM = bo.eM, o.eV = bo.eV, o.eT = bo.eT;
  }
  (void)child0;
}

}  // namespace dq
