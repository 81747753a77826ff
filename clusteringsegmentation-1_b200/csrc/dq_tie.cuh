// dq_tie.cuh -- tie audit of the exact-integer divisive phase.
//
// The large-input split kernels (dq_split2.cu) sum count*c exactly and evaluate the reference's scalar formulas once
// per pass; the reference (DivQuantCluster.cpp:78-86, 493-497, 721-723) adds the doubles w_i*c one after the other.
// The two values of a statistic differ by the rounding noise of the reference's sums, so a decision of the reference
// (D1 axis :388-403, D2 cut :473, D3 hyperplane :683, D4 arg-max TSE :876-887, D5 rounding :1050-1052) is the same here
// unless its two sides are closer than that noise.  This header carries first-order bounds on that noise through the
// split tree; a decision inside its bound raises a TieBit, and the host re-runs a flagged frame in the reference's own
// summation order (the ordered path, dq_split_ordered.cuh).  A frame without a flag equals the reference bit for bit.
//
// What is bounded, with u = 2^-53: for every cluster the absolute deviation between this kernel's and the reference's
//   W = weight,  S_c = weight * mean_c,  Q_c = weight * (var_c + mean_c^2)          (eW, eS, eQ; max over channels)
// i.e. of the three sums the statistics stand for.  In these variables a split is linear (old = parent - new: :579-581,
// :844-855 are exact identities), so bounds simply add along the chain of "old" sides, plus a few u per formula step:
//   * a fresh sequential sum of n rounded products deviates by gamma(n) = min(n + 4, lambda sqrt(n + 4)) u relative.
//     n u is the worst case; rounding errors are not aligned, and lambda sqrt(n) u holds with probability
//     1 - 2 exp(-lambda^2 / 2) (Higham & Mary, "A new approach to probabilistic rounding error analysis", SISC 2019);
//     lambda = 8: 3e-14 per sum.  The rms of such a sum's error is ~0.3 sqrt(n) u, so the bound sits ~25 sigma out.
//   * mean, variance and TSE bounds follow by first-order propagation: mean = S / W, var = Q / W - mean^2,
//     TSE = sum_c Q_c - S_c^2 / W.
// oracle/divquant_oracle.cpp (TieAudit) is the CPU model of exactly these formulas; tests compare the two.
#pragma once

#include "dq_split.cuh"

namespace dq {
namespace tie {

constexpr double kU = 1.1102230246251565e-16;  // 2^-53
constexpr double kLambda = 8.0;

__host__ __device__ __forceinline__ double gamma_n(double n) {
  const double a = n + 4.0, b = kLambda * sqrt(n + 4.0);
  return (a < b ? a : b) * kU;
}
__host__ __device__ __forceinline__ double max3abs(const double *v) { return fmax(fabs(v[0]), fmax(fabs(v[1]), fabs(v[2]))); }
// largest second moment var_c + mean_c^2 over the channels
__host__ __device__ __forceinline__ double max3moment(const double *var, const double *mean) {
  return fmax(fabs(var[0]) + mean[0] * mean[0], fmax(fabs(var[1]) + mean[1] * mean[1], fabs(var[2]) + mean[2] * mean[2]));
}

// Bounds of one cluster: the three sums, and what follows for mean / variance / TSE.
struct Bounds {
  double eW, eS, eQ;
  double eM, eV, eT;
};
__host__ __device__ __forceinline__ void derive(Bounds &b, double tw, const double *tm, const double *tv, double tse) {
  const double m1 = max3abs(tm), m2 = max3moment(tv, tm), w = fabs(tw);
  b.eM = (b.eS + m1 * b.eW) / w + 2.0 * kU * m1;
  b.eV = (b.eQ + m2 * b.eW) / w + 2.0 * m1 * b.eM + 4.0 * kU * m2;
  b.eT = 3.0 * (b.eQ + 2.0 * m1 * b.eS + m1 * m1 * b.eW) + 8.0 * kU * fabs(tse);
}

// Root statistics (DivQuantClusterInitMeanAndVar, :60-104): U sequential adds per sum; weight[0] = 1.0 in both.
__host__ __device__ __forceinline__ Bounds root_bounds(double U, const double *tm, const double *tv) {
  Bounds b;
  b.eW = 0.0;
  b.eS = (gamma_n(U) + 2.0 * kU) * max3abs(tm);
  b.eQ = (gamma_n(U) + 4.0 * kU) * max3moment(tv, tm);
  derive(b, 1.0, tm, tv, 0.0);
  return b;
}

// Bounds of the two centres after a pass whose new side has n_new points (:561-581, :780-810).
// m1_* = largest |mean_c| of the parent, the new and the old centre.
struct PassErr {
  double e_nw, eS_n, e_nm, e_ow, eS_o, e_om;
};
__host__ __device__ __forceinline__ PassErr pass_err(double p_eW, double p_eS, double tw, double m1_t, double nw, double m1_n, double ow,
                                                     double m1_o, double n_new) {
  PassErr r;
  const double g = gamma_n(n_new);
  r.e_nw = g * nw;
  r.eS_n = (g + 2.0 * kU) * nw * m1_n;
  r.e_nm = (2.0 * g + 4.0 * kU) * m1_n;  // = (eS_n + m1_n e_nw) / nw + 2 u m1_n
  r.e_ow = p_eW + r.e_nw + kU * fabs(ow);
  r.eS_o = p_eS + r.eS_n + 4.0 * kU * (tw * m1_t + nw * m1_n);
  r.e_om = (r.eS_o + m1_o * r.e_ow) / fabs(ow) + 2.0 * kU * m1_o;
  return r;
}
#ifdef __CUDACC__
// The same bounds on the device, where they sit on the split's critical path: single-precision square root and reciprocal
// rounded UP (the results are bounds: a few ulp of float more is as good, and never less than the CPU model's), no
// double-precision division or square root.
__device__ __forceinline__ double gamma_fast(double n) {
  const float nf = __double2float_ru(n + 4.0);
  return (double)fminf(nf, __fmul_ru((float)kLambda, __fsqrt_ru(nf))) * kU;
}
__device__ __forceinline__ double rcp_up(double w) { return (double)__frcp_ru(__double2float_rd(fabs(w))); }
__device__ __forceinline__ void derive_fast(Bounds &b, double tw, const double *tm, const double *tv, double tse) {
  const double m1 = max3abs(tm), m2 = max3moment(tv, tm), iw = rcp_up(tw);
  b.eM = (b.eS + m1 * b.eW) * iw + 2.0 * kU * m1;
  b.eV = (b.eQ + m2 * b.eW) * iw + 2.0 * m1 * b.eM + 4.0 * kU * m2;
  b.eT = 3.0 * (b.eQ + 2.0 * m1 * b.eS + m1 * m1 * b.eW) + 8.0 * kU * fabs(tse);
}
__device__ __forceinline__ PassErr pass_err_fast(double p_eW, double p_eS, double tw, double m1_t, double nw, double m1_n, double ow,
                                                 double m1_o, double n_new) {
  PassErr r;
  const double g = gamma_fast(n_new);
  const double inv_ow = rcp_up(ow);
  r.e_nw = g * nw;
  r.eS_n = (g + 2.0 * kU) * nw * m1_n;
  r.e_nm = (2.0 * g + 4.0 * kU) * m1_n;
  r.e_ow = p_eW + r.e_nw + kU * fabs(ow);
  r.eS_o = p_eS + r.eS_n + 4.0 * kU * (tw * m1_t + nw * m1_n);
  r.e_om = (r.eS_o + m1_o * r.e_ow) * inv_ow + 2.0 * kU * m1_o;
  return r;
}
#endif

// Tolerance of the hyperplane test  lhs < rhs . x  (:616-623, :683) for centres with these bounds.
// dot - lhs = (|x - nm|^2 - |x - om|^2) / 2, so a perturbation of the centres moves it by
// sum_c |om_c - x_c| e_om + |nm_c - x_c| e_nm to first order; 2^23 u covers the rounding of the evaluation itself.
// hyperplane_tol: for any point (|om_c - x_c| <= 256), a filter; point_tol: for the point at hand.
__host__ __device__ __forceinline__ double hyperplane_tol(const PassErr &q) { return 1536.0 * (q.e_om + q.e_nm) + 8388608.0 * kU; }
struct PassExt {  // centres and their bounds of the pass in flight (shared memory; read only for points inside the filter)
  double om[3], nm[3], e_om, e_nm;
  double tol;  // the filter's threshold itself (its high word travels with the classification parameters)
};
__host__ __device__ __forceinline__ double point_tol(const PassExt &x, double r, double g, double b) {
  const double d_o = fabs(x.om[0] - r) + fabs(x.om[1] - g) + fabs(x.om[2] - b);
  const double d_n = fabs(x.nm[0] - r) + fabs(x.nm[1] - g) + fabs(x.nm[2] - b);
  return d_o * x.e_om + d_n * x.e_nm + 8388608.0 * kU;
}

#ifdef __CUDACC__
__device__ __forceinline__ void child_bounds_fast(const Bounds &p, const PassErr &fe, double tw, double nw, const double *nm,
                                                  const double *nv, double tse_n, double ow, const double *om, const double *ov,
                                                  double tse_o, double n_new, Bounds &bn, Bounds &bo) {
  bn.eW = fe.e_nw;
  bn.eS = fe.eS_n;
  bn.eQ = (gamma_fast(n_new) + 4.0 * kU) * nw * max3moment(nv, nm);
  bo.eW = fe.e_ow;
  bo.eS = fe.eS_o;
  bo.eQ = p.eQ + bn.eQ + 16.0 * kU * 65536.0 * (tw + nw);
  derive_fast(bn, nw, nm, nv, tse_n);
  derive_fast(bo, ow, om, ov, tse_o);
}
#endif
// Bounds of the two children of a finished split (:800-871) from the parent's and the last pass's.
__host__ __device__ __forceinline__ void child_bounds(const Bounds &p, const PassErr &fe, double tw, const double *tm, const double *tv,
                                                      double nw, const double *nm, const double *nv, double tse_n, double ow,
                                                      const double *om, const double *ov, double tse_o, double n_new, Bounds &bn,
                                                      Bounds &bo) {
  bn.eW = fe.e_nw;
  bn.eS = fe.eS_n;
  bn.eQ = (gamma_n(n_new) + 4.0 * kU) * nw * max3moment(nv, nm);
  bo.eW = fe.e_ow;
  bo.eS = fe.eS_o;
  bo.eQ = p.eQ + bn.eQ + 16.0 * kU * 65536.0 * (tw + nw);  // a dozen roundings of the combined-variance formula (:844-855)
  derive(bn, nw, nm, nv, tse_n);
  derive(bo, ow, om, ov, tse_o);
  (void)tm;
  (void)tv;
}

// D1: the comparisons choose_cut makes (:388-403).
__host__ __device__ __forceinline__ bool axis_tie(const double *tv, double eV) {
  double best = tv[0];
  bool tie = fabs(best - tv[1]) <= 2.0 * eV;
  if (best < tv[1]) best = tv[1];
  tie = tie || (fabs(best - tv[2]) <= 2.0 * eV);
  return tie;
}

// D5: (uint8)(mean + 0.5) (:1050-1052).
__host__ __device__ __forceinline__ bool round_tie(double mean, double eM) {
  const double v = mean + 0.5;
  return fabs(v - rint(v)) <= eM + 512.0 * kU;
}

}  // namespace tie
}  // namespace dq
