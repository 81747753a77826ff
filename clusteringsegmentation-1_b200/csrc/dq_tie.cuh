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
// Bounds are on |value here - value in the reference| with u = 2^-53:
//   * a sequential sum of n rounded products: gamma(n) = min(n + 4, lambda sqrt(n + 4)) u relative.  n u is the worst
//     case; rounding errors are not aligned, and lambda sqrt(n) u holds with probability 1 - 2 exp(-lambda^2 / 2)
//     (Higham & Mary, "A new approach to probabilistic rounding error analysis", 2019); lambda = 8: 3e-14 per sum.
//     The measured rms of such a sum's error is ~0.3 sqrt(n) u, so the bound sits ~25 sigma out.
//   * every scalar formula propagates its operands' bounds to first order, with |channel| <= 256, |channel^2| <= 65536.
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

// Root statistics (DivQuantClusterInitMeanAndVar, :60-104): U sequential adds per sum.
__host__ __device__ __forceinline__ void root_bounds(double U, double &eW, double &eM, double &eV) {
  eW = 0.0;  // weight[0] = 1.0 in both
  eM = gamma_n(U) * 256.0;
  eV = (3.0 * gamma_n(U) + 8.0 * kU) * 65536.0;
}

// Bounds of the centres after a pass whose new side has n_new points (:561-581, :780-810).
struct PassErr {
  double e_nw, e_nm, e_ow, e_om;
};
__host__ __device__ __forceinline__ PassErr pass_err(double p_eW, double p_eM, double tw, double nw, double ow, double n_new) {
  PassErr r;
  const double g = gamma_n(n_new);
  r.e_nw = g * nw;
  r.e_nm = (2.0 * g + 4.0 * kU) * 256.0;
  r.e_ow = p_eW + r.e_nw + kU * fabs(ow);
  const double e_num = p_eW * 256.0 + tw * p_eM + r.e_nw * 256.0 + nw * r.e_nm + 4.0 * kU * (tw + nw) * 256.0;
  r.e_om = (e_num + 256.0 * r.e_ow) / fabs(ow) + 2.0 * kU * 256.0;
  return r;
}

// Tolerance of the hyperplane test  lhs < rhs . x  (:616-623, :683) for centres with these bounds.
// dot - lhs = (|x - nm|^2 - |x - om|^2) / 2, so a perturbation of the centres moves it by
// sum_c |om_c - x_c| e_om + |nm_c - x_c| e_nm to first order; 2^23 u covers the rounding of the evaluation itself.
// hyperplane_tol: for any point (|om_c - x_c| <= 256), a filter; point_tol: for the point at hand.
__host__ __device__ __forceinline__ double hyperplane_tol(const PassErr &q) { return 1536.0 * (q.e_om + q.e_nm) + 8388608.0 * kU; }
struct PassExt {  // centres and their bounds of the pass in flight (shared memory; read only for points inside the filter)
  double om[3], nm[3], e_om, e_nm;
};
__host__ __device__ __forceinline__ double point_tol(const PassExt &x, double r, double g, double b) {
  const double d_o = fabs(x.om[0] - r) + fabs(x.om[1] - g) + fabs(x.om[2] - b);
  const double d_n = fabs(x.nm[0] - r) + fabs(x.nm[1] - g) + fabs(x.nm[2] - b);
  return d_o * x.e_om + d_n * x.e_nm + 8388608.0 * kU;
}

// Bounds of the two children's variance and TSE (:836-871).  nv/ov = the children's variances as computed.
__host__ __device__ __forceinline__ void child_var_bounds(const PassErr &fe, double p_eW, double p_eM, double p_eV, double tw,
                                                          const double *tm, const double *tv, double nw, double ow, const double *nm,
                                                          const double *om, const double *nv, const double *ov, double n_new,
                                                          double tse_new, double tse_old, double &eVn, double &eVo, double &eTn,
                                                          double &eTo) {
  eVn = (2.0 * gamma_n(n_new) + 8.0 * kU) * 65536.0 + 512.0 * fe.e_nm;
  eVo = 0.0;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const double dn = fabs(nm[c] - tm[c]), dmo = fabs(om[c] - tm[c]);
    const double inner = fabs(nv[c]) + dn * dn;
    const double e_inner = eVn + 2.0 * dn * (fe.e_nm + p_eM) + 3.0 * kU * inner;
    const double e_num = p_eW * fabs(tv[c]) + tw * p_eV + fe.e_nw * inner + nw * e_inner + 3.0 * kU * (tw * fabs(tv[c]) + nw * inner);
    const double q = (tw * fabs(tv[c]) + nw * inner) / fabs(ow);
    const double e_q = (e_num + q * fe.e_ow) / fabs(ow) + kU * q;
    const double e = e_q + 2.0 * dmo * (fe.e_om + p_eM) + 3.0 * kU * (q + dmo * dmo);
    if (!(e <= eVo)) eVo = e;  // NaN propagates
  }
  eTn = fe.e_nw * fabs(nv[0] + nv[1] + nv[2]) + nw * 3.0 * eVn + 4.0 * kU * fabs(tse_new);
  eTo = fe.e_ow * fabs(ov[0] + ov[1] + ov[2]) + fabs(ow) * 3.0 * eVo + 4.0 * kU * fabs(tse_old);
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
