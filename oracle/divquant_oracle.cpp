// oracle/divquant_oracle.cpp -- TEST INFRASTRUCTURE ONLY.  NOT PART OF THE PRODUCT.
//
// An independent CPU restatement of the reference's DivQuant hot path (caomw/ClusteringSegmentation-1,
// DivQuant/*).  It exists to check the CUDA path in csrc/ and is itself pinned two ways:
//   (1) against the seven known-answer palettes of the reference's Test/DivQuantTest.m
//       (tests/golden/divquant_kat.json), and
//   (2) against the UNMODIFIED reference sources compiled from /root/reference into
//       oracle/_ref/libdivquant_ref.so (tests/test_oracle_golden.py; fixtures in tests/golden/).
// Parity status: PINNED (both of the above pass bit-for-bit, including the floating-point
// summation order of the weighted path, which this file follows operation by operation).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this library.  Every function cites the reference file:line it restates.  The arithmetic
// is written so that g++ evaluates the same IEEE-754 double operations in the same order as the
// reference build (compile with -ffp-contract=off; see oracle/Makefile).
#include "divquant_oracle.h"

#include <algorithm>
#include <cassert>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unordered_map>
#include <vector>

namespace {

inline uint32_t chan_r(uint32_t p) { return (p >> 16) & 0xFFu; }
inline uint32_t chan_g(uint32_t p) { return (p >> 8) & 0xFFu; }
inline uint32_t chan_b(uint32_t p) { return p & 0xFFu; }
inline double sq(double x) { return x * x; }

struct Vec3 {
  double r, g, b;
};

// ---------------------------------------------------------------------------------------------
// 24-bit colour histogram.  Reference: calc_color_table, DivQuantMapColors.cpp:82-203.
// The reference keeps a 20 023-bucket chained hash (HASH macro :55-62) whose chains are pushed at
// the head, then walks buckets in ascending order.  The emission order matters downstream because
// the weighted statistics are summed sequentially in doubles in exactly that order.
// ---------------------------------------------------------------------------------------------
const int kBuckets = 20023;

inline int bucket_of(uint32_t r, uint32_t g, uint32_t b) {
  long h = (long)r * 33023 + (long)g * 30013 + (long)b * 27011;
  return (int)((h & 0x7fffffff) % kBuckets);
}

struct ChainNode {
  uint32_t colour;  // 0x00RRGGBB
  uint32_t count;
  int next;  // index of the next node in the chain, -1 = end
};

}  // namespace

extern "C" int oracle_calc_color_table(const uint32_t *in, uint32_t /*num_pixels*/, uint32_t num_rows,
                                       uint32_t num_cols, int dec_factor, uint32_t *unique_out,
                                       double *weights_out, uint32_t *counts_out) {
  if (dec_factor <= 0) {
    fprintf(stderr, "Decimation factor ( %d ) should be positive !\n", dec_factor);
    return -1;
  }
  std::vector<int> head(kBuckets, -1);
  std::vector<ChainNode> nodes;
  // Sampling grid, including the reference's `ic + ir * numRows` addressing (:124) -- harmless in
  // practice because the only caller passes numRows == 1 (quant_util.cpp:60).
  for (int ir = 0; ir < (int)num_rows; ir += dec_factor) {
    for (int ic = 0; ic < (int)num_cols; ic += dec_factor) {
      uint32_t colour = in[ic + ir * (int)num_rows] & 0x00FFFFFFu;
      int h = bucket_of(chan_r(colour), chan_g(colour), chan_b(colour));
      int at = head[h];
      while (at >= 0 && nodes[at].colour != colour) at = nodes[at].next;
      if (at >= 0) {
        nodes[at].count++;
      } else {
        ChainNode fresh = {colour, 1u, head[h]};  // push at the chain head (:157-158)
        head[h] = (int)nodes.size();
        nodes.push_back(fresh);
      }
    }
  }
  // Frequencies -> probabilities (:172).
  double norm = 1.0 / (std::ceil(num_rows / (double)dec_factor) * std::ceil(num_cols / (double)dec_factor));
  int emitted = 0;
  for (int h = 0; h < kBuckets; ++h) {
    for (int at = head[h]; at >= 0; at = nodes[at].next) {
      unique_out[emitted] = nodes[at].colour;
      if (weights_out) weights_out[emitted] = norm * (int)nodes[at].count;  // (:185) bucket->value is int
      if (counts_out) counts_out[emitted] = nodes[at].count;
      ++emitted;
    }
  }
  return emitted;
}

// Reference: cut_bits, DivQuantUni.cpp:28-100 (validate_num_bits, DivQuantMisc.cpp:36-46).
extern "C" int oracle_cut_bits(const uint32_t *in, uint32_t num_pixels, uint32_t *out, int rbits, int gbits,
                               int bbits) {
  const int bits[3] = {rbits, gbits, bbits};
  for (int i = 0; i < 3; ++i) {
    if (!(0 < bits[i] && bits[i] <= 8)) {
      fprintf(stderr, "Number of bits per channel ( %d ) must be in [1,8] !\n", bits[i]);
      return 0;
    }
  }
  const uint32_t sr = 8 - rbits, sg = 8 - gbits, sb = 8 - bbits;
  if (sr == sg && sr == sb) {
    // Whole-word variant (:60-74): the alpha byte is masked away, then everything shifts together.
    const uint32_t byte_mask = (0xFFu >> sr) << sr;
    const uint32_t word_mask = (byte_mask << 16) | (byte_mask << 8) | byte_mask;
    for (uint32_t i = 0; i < num_pixels; ++i) out[i] = (in[i] & word_mask) >> sr;
  } else {
    for (uint32_t i = 0; i < num_pixels; ++i) {
      uint32_t p = in[i];
      out[i] = ((chan_r(p) >> sr) << 16) | ((chan_g(p) >> sg) << 8) | (chan_b(p) >> sb);
    }
  }
  return 1;
}

namespace {

// ---------------------------------------------------------------------------------------------
// Divisive phase.  Reference: DivQuantCluster<UW,MT,KM=true>, DivQuantCluster.cpp:133-1097, and
// DivQuantClusterInitMeanAndVar, :49-123.
//
// `uniform` mirrors the UW template flag: every point carries weight `w_uniform` and the reference
// accumulates plain integers (u32 chunks of <= 0xFFFF points, :440-455), which is exact, so one
// exact u64 accumulation per pass is the same number.  Otherwise `w[i]` are the doubles produced
// by calc_color_table and every sum is a sequential double accumulation in point order.
// ---------------------------------------------------------------------------------------------
struct DivisiveState {
  int num_points;
  const uint32_t *data;
  const double *w;  // nullptr when uniform
  double w_uniform;
  bool uniform;
  // Device-arithmetic model only (oracle_quant_varpart_fast_exact): integer multiplicity of each
  // point.  nullptr = 1, which is the reference's own uniform path.
  const uint32_t *counts;
};

// Global weighted mean / variance of all points (:60-104).
void initial_mean_and_var(const DivisiveState &s, Vec3 *mean, Vec3 *var) {
  double mr = 0.0, mg = 0.0, mb = 0.0, vr = 0.0, vg = 0.0, vb = 0.0;
  for (int ip = 0; ip < s.num_points; ++ip) {
    uint32_t p = s.data[ip];
    uint32_t R = chan_r(p), G = chan_g(p), B = chan_b(p);
    if (s.uniform) {
      double c = s.counts ? (double)s.counts[ip] : 1.0;  // exact integers either way
      mr += c * R;
      mg += c * G;
      mb += c * B;
      vr += c * (R * R);
      vg += c * (G * G);
      vb += c * (B * B);
    } else {
      double wt = s.w[ip];
      mr += wt * R;
      mg += wt * G;
      mb += wt * B;
      vr += wt * (R * R);
      vg += wt * (G * G);
      vb += wt * (B * B);
    }
  }
  if (s.uniform) {
    mr *= s.w_uniform;
    mg *= s.w_uniform;
    mb *= s.w_uniform;
    vr *= s.w_uniform;
    vg *= s.w_uniform;
    vb *= s.w_uniform;
  }
  vr -= sq(mr);
  vg -= sq(mg);
  vb -= sq(mb);
  mean->r = mr;
  mean->g = mg;
  mean->b = mb;
  var->r = vr;
  var->g = vg;
  var->b = vb;
}

}  // namespace

// Diagnostics for DESIGN.md: in how many LKM iterations does a split reach its fixed point?
static int g_converged_at[4096];
static int g_converged_n = 0;
extern "C" int oracle_debug_convergence(int *out, int cap) {
  int n = g_converged_n < cap ? g_converged_n : cap;
  for (int i = 0; i < n; ++i) out[i] = g_converged_at[i];
  return n;
}

// ---------------------------------------------------------------------------------------------
// Tie audit of the device-arithmetic model (exact_counts): a CPU model of csrc/dq_tie.cuh.
//
// The exact-integer path computes every sum over points exactly and the reference sums doubles
// sequentially, so the two agree on a decision unless the compared values are closer than the
// rounding noise of the reference's sums.  For every cluster the audit carries first-order bounds on
// |model - reference| of (weight, mean, variance, TSE) and flags a decision whose margin is inside the
// bound: D1 axis choice, D2 cut test, D3 hyperplane test, D4 TSE arg-max, D5 final .5 rounding.
// A frame with no flag is guaranteed (to first order, with a safety factor) to equal the reference.
// ---------------------------------------------------------------------------------------------
namespace {
const double kU = 1.1102230246251565e-16;  // 2^-53
const double kLambda = 8.0;
// Rounding errors of a sequential sum of n terms: n*u is the worst case, but the errors are not aligned; the
// probabilistic bound lambda*sqrt(n)*u (Higham & Mary 2019) fails with probability ~2 exp(-lambda^2/2) (3e-14 at 8).
inline double gamma_n(double n) {
  const double a = n + 4.0, b = kLambda * std::sqrt(n + 4.0);
  return (a < b ? a : b) * kU;
}
inline double max3abs(const Vec3 &v) { return std::fmax(std::fabs(v.r), std::fmax(std::fabs(v.g), std::fabs(v.b))); }
inline double max3moment(const Vec3 &var, const Vec3 &mean) {
  return std::fmax(std::fabs(var.r) + mean.r * mean.r, std::fmax(std::fabs(var.g) + mean.g * mean.g, std::fabs(var.b) + mean.b * mean.b));
}
// bounds on |model - reference| of W, S = W*mean, Q = W*(var + mean^2) and what follows for mean / variance / TSE
struct ClusterErr {
  double eW, eS, eQ, eM, eV, eT;
};
inline void derive_err(ClusterErr &b, double tw, const Vec3 &tm, const Vec3 &tv, double tse) {
  const double m1 = max3abs(tm), m2 = max3moment(tv, tm), w = std::fabs(tw);
  b.eM = (b.eS + m1 * b.eW) / w + 2.0 * kU * m1;
  b.eV = (b.eQ + m2 * b.eW) / w + 2.0 * m1 * b.eM + 4.0 * kU * m2;
  b.eT = 3.0 * (b.eQ + 2.0 * m1 * b.eS + m1 * m1 * b.eW) + 8.0 * kU * std::fabs(tse);
}
struct TieAudit {
  uint32_t flags[6];  // [1..5] = D1..D5 counts, [0] = total
  // CPU model of the device's tie resolver (csrc/dq_resolve.cu + the forced cuts of csrc/dq_context.cu): next to the exact
  // integer sums, the reference's own sequential double sums are carried for the SAME memberships ("shadow" statistics:
  // what the resolver recomputes from the split tree).  A flagged cut is taken from the shadow mean when the two means
  // have different floors, a flagged rounding rounds the shadow mean.  [0] roundings taken from the shadow, [1] cuts
  // confirmed (same floor), [2] cuts forced.
  bool resolve;
  uint32_t resolved[3];
  void hit(int d) {
    flags[d]++;
    flags[0]++;
  }
};
// bounds after a pass: the new side has n_new points
struct PassErr {
  double e_nw, eS_n, e_nm, e_ow, eS_o, e_om;
};
inline PassErr pass_err(const ClusterErr &pe, double tw, const Vec3 &tm, double nw, const Vec3 &nm, double ow, const Vec3 &om, int n_new) {
  PassErr r;
  const double g = gamma_n(n_new), m1_t = max3abs(tm), m1_n = max3abs(nm), m1_o = max3abs(om);
  r.e_nw = g * nw;
  r.eS_n = (g + 2.0 * kU) * nw * m1_n;
  r.e_nm = (r.eS_n + m1_n * r.e_nw) / nw + 2.0 * kU * m1_n;
  r.e_ow = pe.eW + r.e_nw + kU * std::fabs(ow);
  r.eS_o = pe.eS + r.eS_n + 4.0 * kU * (tw * m1_t + nw * m1_n);
  r.e_om = (r.eS_o + m1_o * r.e_ow) / std::fabs(ow) + 2.0 * kU * m1_o;
  return r;
}
}  // namespace

static int varpart_impl(uint32_t num_pixels, const uint32_t *in, uint32_t num_rows, uint32_t num_cols,
                        uint32_t *num_clusters, uint32_t *colortable, int num_bits, int dec_factor,
                        int max_iters, int all_pixels_unique, oracle_split_record *records, int *num_records,
                        bool exact_counts, TieAudit *audit = nullptr) {
  assert(0 < num_bits && num_bits <= 8);  // (:1115-1118)
  assert(max_iters >= 1);                 // KM is hard-wired true; 0 iterations is degenerate
  if (num_records) *num_records = 0;

  // ---- path selection, quant_varpart_fast :1130-1147 ----
  std::vector<uint32_t> points, counts;
  std::vector<double> weights;
  DivisiveState s;
  if (all_pixels_unique && num_bits == 8 && dec_factor == 1) {
    s.uniform = true;
    s.w_uniform = 1.0 / (std::ceil(1 / (double)1) * std::ceil((int)num_pixels / (double)1));  // (:215)
    s.w = nullptr;
    s.counts = nullptr;
    s.data = in;
    s.num_points = (int)num_pixels;
  } else {
    std::vector<uint32_t> sampled;
    const uint32_t *src = in;
    if (!(!all_pixels_unique && num_bits == 8)) {
      sampled.resize(num_pixels);
      oracle_cut_bits(in, num_pixels, sampled.data(), num_bits, num_bits, num_bits);
      src = sampled.data();
    }
    points.resize(num_pixels);
    weights.resize(num_pixels);
    counts.resize(num_pixels);
    int u = oracle_calc_color_table(src, num_pixels, num_rows, num_cols, dec_factor, points.data(),
                                    weights.data(), counts.data());
    assert(u > 0);
    s.uniform = false;
    s.w_uniform = 0.0;
    s.w = weights.data();
    s.counts = nullptr;
    s.data = points.data();
    s.num_points = u;
    if (exact_counts) {
      // Device-arithmetic model: the same scalar formulas, but every sum over points is an exact
      // integer sum of count*c scaled once by norm = 1/#samples (what csrc/ computes on the GPU).
      s.uniform = true;
      s.counts = counts.data();
      s.w_uniform = 1.0 / (std::ceil(num_rows / (double)dec_factor) * std::ceil(num_cols / (double)dec_factor));
    }
  }

  const int K = (int)*num_clusters;
  assert(K > 0);
  const int U = s.num_points;

  // Per-cluster bookkeeping (all zero-initialised like the reference's `new T[n]()`, :296-324).
  std::vector<double> weight(K, 0.0), tse(K, 0.0);
  std::vector<int> size(K, 0);
  std::vector<Vec3> mean(K, Vec3{0, 0, 0}), var(K, Vec3{0, 0, 0});
  std::vector<uint32_t> member(U, 0u);
  std::vector<ClusterErr> cerr(K, ClusterErr{0, 0, 0, 0, 0, 0});
  std::vector<char> eq_rg, eq_gb, eq_rb;  // (D1 experiment below)
  const bool resolve = audit != nullptr && audit->resolve && exact_counts && s.counts != nullptr;
  if (audit) {
    memset(audit, 0, sizeof(*audit));
    audit->resolve = resolve;
  }
  // shadow statistics of the resolver model: the reference's weight and mean of every cluster for the model's memberships
  std::vector<double> s_weight(resolve ? K : 0, 0.0);
  std::vector<Vec3> s_mean(resolve ? K : 0, Vec3{0, 0, 0});
  if (resolve) s_weight[0] = 1.0;

  // The cluster being split: `cur` lists original point indices in ascending order (the reference
  // keeps tmp_data/point_index, :929-1019; before the first gather it is the identity).
  std::vector<int> cur(U);
  for (int i = 0; i < U; ++i) cur[i] = i;

  int old_index = 0;
  weight[0] = 1.0;
  size[0] = U;
  int cur_n = U;
  const int last_it = max_iters - 1;

  for (int new_index = 1; new_index < K; ++new_index) {
    const double tw = weight[old_index];
    Vec3 tm, tv;
    if (new_index == 1) {
      initial_mean_and_var(s, &tm, &tv);
      // reference: U sequential adds of rounded products per sum
      cerr[0].eW = 0.0;
      cerr[0].eS = (gamma_n(U) + 2.0 * kU) * max3abs(tm);
      cerr[0].eQ = (gamma_n(U) + 4.0 * kU) * max3moment(tv, tm);
      derive_err(cerr[0], 1.0, tm, tv, 0.0);
    } else {
      tm = mean[old_index];
      tv = var[old_index];
    }

    // Cutting axis = channel of largest variance, cut at the mean (:388-403). The blue branch does
    // not refresh max_val in the reference; nothing reads it afterwards.
    double best = tv.r;
    int axis = 0;
    double cut = tm.r;
    const ClusterErr pe = cerr[old_index];
    // (experiment, ORACLE_D1_EQUAL_CHANNELS=1, not on the device yet: two channels that are equal in EVERY point the cluster's
    // statistics were summed over -- its own points for a "new" side, inherited along "old" sides -- go through the same
    // operations on the same values in the reference and here, so their variances are bit-identical in both and comparing
    // them is no tie to worry about)
    static const bool d1_equal = getenv("ORACLE_D1_EQUAL_CHANNELS") != nullptr;
    if (audit && d1_equal && new_index == 1) {
      eq_rg.assign(K, 0), eq_gb.assign(K, 0), eq_rb.assign(K, 0);
      bool a = true, b2 = true, c2 = true;
      for (int j = 0; j < U; ++j) {
        const uint32_t p = s.data[j];
        a = a && chan_r(p) == chan_g(p), b2 = b2 && chan_g(p) == chan_b(p), c2 = c2 && chan_r(p) == chan_b(p);
      }
      eq_rg[0] = a, eq_gb[0] = b2, eq_rb[0] = c2;
    }
    const bool rg_same = audit && d1_equal && eq_rg[old_index], gb_same = audit && d1_equal && eq_gb[old_index],
               rb_same = audit && d1_equal && eq_rb[old_index];
    if (audit && std::fabs(best - tv.g) <= 2.0 * pe.eV && !(rg_same && tv.r == tv.g)) audit->hit(1);
    if (best < tv.g) {
      best = tv.g;
      axis = 1;
      cut = tm.g;
    }
    if (audit && std::fabs(best - tv.b) <= 2.0 * pe.eV && !((axis == 0 ? rb_same : gb_same) && best == tv.b)) audit->hit(1);
    if (best < tv.b) {
      axis = 2;
      cut = tm.b;
    }

    // resolver model: the reference's own mean of this cluster (root: the sequential weighted sums of :60-104)
    double s_tw = 0.0;
    Vec3 s_tm{0, 0, 0};
    if (resolve) {
      s_tw = s_weight[old_index];
      if (new_index == 1) {
        DivisiveState ref_state = s;
        ref_state.uniform = false;
        ref_state.counts = nullptr;
        ref_state.w = weights.data();
        Vec3 unused_var;
        initial_mean_and_var(ref_state, &s_tm, &unused_var);
      } else {
        s_tm = s_mean[old_index];
      }
      // a flagged cut (some point within the bound of the cut): the points are integers, so the two cuts separate the same
      // points iff they have the same floor; otherwise the reference's cut is taken (the device's forced cut)
      const double s_cut = (axis == 0) ? s_tm.r : ((axis == 1) ? s_tm.g : s_tm.b);
      bool flagged = false;
      for (int j = 0; j < cur_n && !flagged; ++j) {
        const uint32_t p = s.data[cur[j]];
        const double proj = (axis == 0) ? chan_r(p) : ((axis == 1) ? chan_g(p) : chan_b(p));
        flagged = std::fabs(proj - cut) <= pe.eM;
      }
      if (flagged) {
        if (std::floor(s_cut) == std::floor(cut)) {
          audit->resolved[1]++;
        } else {
          audit->resolved[2]++;
          cut = s_cut;
        }
      }
    }

    Vec3 &nm = mean[new_index];
    Vec3 &nv = var[new_index];
    Vec3 &om = mean[old_index];
    nm = Vec3{0, 0, 0};

    // ---- split pass (:438-559): points strictly above the cut seed the new cluster ----
    double nw = 0.0;
    int cut_new = 0;
    {
      uint64_t ir = 0, ig = 0, ib = 0, cnt = 0;
      for (int j = 0; j < cur_n; ++j) {
        int idx = cur[j];
        uint32_t p = s.data[idx];
        uint32_t R = chan_r(p), G = chan_g(p), B = chan_b(p);
        double proj = (axis == 0) ? R : ((axis == 1) ? G : B);
        if (audit && !resolve && std::fabs(proj - cut) <= pe.eM) audit->hit(2);
        if (cut < proj) {
          ++cut_new;
          if (s.uniform) {
            uint64_t c = s.counts ? s.counts[idx] : 1u;
            ir += c * R;
            ig += c * G;
            ib += c * B;
            cnt += c;
          } else {
            double wt = s.w[idx];
            nm.r += wt * R;
            nm.g += wt * G;
            nm.b += wt * B;
            nw += wt;
          }
        }
      }
      if (s.uniform) {
        nm.r += (double)ir;
        nm.g += (double)ig;
        nm.b += (double)ib;
        nm.r *= s.w_uniform;
        nm.g *= s.w_uniform;
        nm.b *= s.w_uniform;
        nw = (double)cnt * s.w_uniform;
      }
    }
    double ow = tw - nw;
    nm.r /= nw;
    nm.g /= nw;
    nm.b /= nw;
    // 'combined mean' (:579-581)
    om.r = (tw * tm.r - nw * nm.r) / ow;
    om.g = (tw * tm.g - nw * nm.g) / ow;
    om.b = (tw * tm.b - nw * nm.b) / ow;

    // ---- local 2-means refinement (:613-811) ----
    int new_size = cut_new;
    int prev_size = -1, converged_at = max_iters;
    uint64_t prev_sig = 0;
    if (new_index == 1) g_converged_n = 0;
    for (int it = 0; it < max_iters; ++it) {
      const double lhs =
          0.5 * (sq(om.r) - sq(nm.r) + sq(om.g) - sq(nm.g) + sq(om.b) - sq(nm.b));  // (:616-619)
      const double rr = om.r - nm.r, rg = om.g - nm.g, rb = om.b - nm.b;
      double eH = 0.0;
      PassErr q = {0, 0, 0, 0, 0, 0};
      const Vec3 om_pass = om, nm_pass = nm;  // the centres this pass classifies with
      if (audit) {
        q = pass_err(pe, tw, tm, nw, nm, ow, om, new_size);
        eH = 1536.0 * (q.e_om + q.e_nm) + 8388608.0 * kU;  // filter: holds for any point
      }
      nw = 0.0;
      new_size = 0;
      nm = Vec3{0, 0, 0};
      nv = Vec3{0, 0, 0};
      uint64_t ir = 0, ig = 0, ib = 0, irr = 0, igg = 0, ibb = 0, cnt = 0;
      double s_nw = 0.0;  // resolver model: the reference's sums of the new side, term by term in emission order
      Vec3 s_nm{0, 0, 0};
      for (int j = 0; j < cur_n; ++j) {
        int idx = cur[j];
        uint32_t p = s.data[idx];
        uint32_t R = chan_r(p), G = chan_g(p), B = chan_b(p);
        double red = R, green = G, blue = B;
        if (audit && std::fabs(((rr * red) + (rg * green) + (rb * blue)) - lhs) <= eH) {
          // inside the filter: the bound for this very point, sum_c |om_c - x_c| e_om + |nm_c - x_c| e_nm (dq_tie.cuh)
          const double d_o = std::fabs(om_pass.r - red) + std::fabs(om_pass.g - green) + std::fabs(om_pass.b - blue);
          const double d_n = std::fabs(nm_pass.r - red) + std::fabs(nm_pass.g - green) + std::fabs(nm_pass.b - blue);
          if (std::fabs(((rr * red) + (rg * green) + (rb * blue)) - lhs) <= d_o * q.e_om + d_n * q.e_nm + 8388608.0 * kU)
            audit->hit(3);
        }
        if (lhs < ((rr * red) + (rg * green) + (rb * blue))) {
          // closer to the old centre (:683)
          if (it == last_it) member[idx] = (uint32_t)old_index;
        } else {
          if (s.uniform) {
            uint64_t c = s.counts ? s.counts[idx] : 1u;
            ir += c * R;
            ig += c * G;
            ib += c * B;
            cnt += c;
            if (it == last_it) {
              irr += c * (R * R);
              igg += c * (G * G);
              ibb += c * (B * B);
              if (resolve) {
                const double wt = weights[idx];
                s_nm.r += wt * red;
                s_nm.g += wt * green;
                s_nm.b += wt * blue;
                s_nw += wt;
              }
            }
          } else {
            double wt = s.w[idx];
            nm.r += wt * red;
            nm.g += wt * green;
            nm.b += wt * blue;
            if (it == last_it) {
              nv.r += wt * (R * R);
              nv.g += wt * (G * G);
              nv.b += wt * (B * B);
            }
            nw += wt;
          }
          if (it == last_it) member[idx] = (uint32_t)new_index;
          new_size++;
        }
      }
      {
        // signature of the membership sums of this iteration (exact in the integer paths)
        uint64_t sig = s.uniform ? (ir * 1000003ull + ig * 10007ull + ib * 101ull + cnt) : (uint64_t)new_size;
        if (converged_at == max_iters && prev_size == new_size && sig == prev_sig) converged_at = it;
        prev_size = new_size;
        prev_sig = sig;
      }
      if (s.uniform) {
        nm.r += (double)ir;
        nm.g += (double)ig;
        nm.b += (double)ib;
        nv.r += (double)irr;
        nv.g += (double)igg;
        nv.b += (double)ibb;
        nm.r *= s.w_uniform;
        nm.g *= s.w_uniform;
        nm.b *= s.w_uniform;
        nw = (double)cnt * s.w_uniform;  // == new_size * data_weight when every count is 1 (:788)
        nv.r *= s.w_uniform;
        nv.g *= s.w_uniform;
        nv.b *= s.w_uniform;
      }
      nm.r /= nw;
      nm.g /= nw;
      nm.b /= nw;
      ow = tw - nw;
      om.r = (tw * tm.r - nw * nm.r) / ow;
      om.g = (tw * tm.g - nw * nm.g) / ow;
      om.b = (tw * tm.b - nw * nm.b) / ow;
      if (resolve && it == last_it) {
        // the reference's centres after its last pass (:780-810), from its own sums over the model's memberships
        s_nm.r /= s_nw;
        s_nm.g /= s_nw;
        s_nm.b /= s_nw;
        const double s_ow = s_tw - s_nw;
        s_mean[new_index] = s_nm;
        s_mean[old_index] = Vec3{(s_tw * s_tm.r - s_nw * s_nm.r) / s_ow, (s_tw * s_tm.g - s_nw * s_nm.g) / s_ow,
                                 (s_tw * s_tm.b - s_nw * s_nm.b) / s_ow};
        if (new_index < K - 1) {  // (the last split leaves the weights alone, :823-832)
          s_weight[old_index] = s_ow;
          s_weight[new_index] = s_nw;
        }
      }
    }

    size[old_index] = cur_n - new_size;
    size[new_index] = new_size;
    if (!eq_rg.empty()) {  // (D1 experiment) the new side's sums run over its own points; the old side inherits
      bool a = true, b2 = true, c2 = true;
      for (int j = 0; j < cur_n; ++j) {
        if (member[cur[j]] != (uint32_t)new_index) continue;
        const uint32_t p = s.data[cur[j]];
        a = a && chan_r(p) == chan_g(p), b2 = b2 && chan_g(p) == chan_b(p), c2 = c2 && chan_r(p) == chan_b(p);
      }
      eq_rg[new_index] = a, eq_gb[new_index] = b2, eq_rb[new_index] = c2;
    }
    if (g_converged_n < 4096) g_converged_at[g_converged_n++] = converged_at;
    PassErr fe = {0, 0, 0, 0, 0, 0};
    if (audit) {
      fe = pass_err(pe, tw, tm, nw, nm, ow, om, new_size);
      ClusterErr bn = {fe.e_nw, fe.eS_n, 0, 0, 0, 0}, bo = {fe.e_ow, fe.eS_o, pe.eQ, 0, 0, 0};
      // the last split computes no variances (:823-832): the means' bounds are all that is read afterwards
      derive_err(bn, nw, nm, Vec3{0, 0, 0}, 0.0);
      derive_err(bo, ow, om, Vec3{0, 0, 0}, 0.0);
      cerr[new_index] = bn;
      cerr[old_index] = bo;
    }

    oracle_split_record rec;
    memset(&rec, 0, sizeof(rec));
    rec.new_index = new_index;
    rec.old_index = old_index;
    rec.cut_axis = axis;
    rec.num_points = cur_n;
    rec.new_size = new_size;
    rec.cut_pos = cut;
    rec.total_weight = tw;
    rec.new_weight = nw;
    rec.old_weight = ow;
    rec.new_mean[0] = nm.r, rec.new_mean[1] = nm.g, rec.new_mean[2] = nm.b;
    rec.old_mean[0] = om.r, rec.old_mean[1] = om.g, rec.old_mean[2] = om.b;

    if (new_index == K - 1) {
      // Last split: the reference leaves without touching var/weight/tse (:823-832).
      rec.is_last = 1;
      if (records) records[new_index - 1] = rec;
      if (num_records) *num_records = new_index;
      break;
    }

    // Variances: new side from its own sums, old side by the 'combined variance' formula (:836-855).
    nv.r = nv.r / nw - sq(nm.r);
    nv.g = nv.g / nw - sq(nm.g);
    nv.b = nv.b / nw - sq(nm.b);
    Vec3 &ov = var[old_index];
    ov.r = ((tw * tv.r - nw * (nv.r + sq(nm.r - tm.r))) / ow) - sq(om.r - tm.r);
    ov.g = ((tw * tv.g - nw * (nv.g + sq(nm.g - tm.g))) / ow) - sq(om.g - tm.g);
    ov.b = ((tw * tv.b - nw * (nv.b + sq(nm.b - tm.b))) / ow) - sq(om.b - tm.b);
    weight[old_index] = ow;
    weight[new_index] = nw;
    tse[old_index] = ow * (ov.r + ov.g + ov.b);
    tse[new_index] = nw * (nv.r + nv.g + nv.b);

    if (audit) {
      // Q = W (var + mean^2): fresh sum on the new side, parent - new on the old side (:836-855)
      ClusterErr bn = {fe.e_nw, fe.eS_n, (gamma_n(new_size) + 4.0 * kU) * nw * max3moment(nv, nm), 0, 0, 0};
      ClusterErr bo = {fe.e_ow, fe.eS_o, pe.eQ + bn.eQ + 16.0 * kU * 65536.0 * (tw + nw), 0, 0, 0};
      derive_err(bn, nw, nm, nv, tse[new_index]);
      derive_err(bo, ow, om, ov, tse[old_index]);
      cerr[new_index] = bn;
      cerr[old_index] = bo;
    }
    rec.new_var[0] = nv.r, rec.new_var[1] = nv.g, rec.new_var[2] = nv.b;
    rec.old_var[0] = ov.r, rec.old_var[1] = ov.g, rec.old_var[2] = ov.b;
    rec.new_tse = tse[new_index];
    rec.old_tse = tse[old_index];
    if (records) records[new_index - 1] = rec;
    if (num_records) *num_records = new_index;

    // Next victim: strictly-greater scan seeded with DBL_MIN, so old_index goes stale when no
    // cluster has a TSE above it (:876-887).
    double top = DBL_MIN;
    for (int ic = 0; ic <= new_index; ++ic) {
      if (top < tse[ic]) {
        top = tse[ic];
        old_index = ic;
      }
    }
    if (audit) {
      // D4: the winner must beat every other cluster (and DBL_MIN) by more than the bounds
      bool tie = !(top - cerr[old_index].eT > DBL_MIN);
      for (int ic = 0; ic <= new_index && !tie; ++ic)
        if (ic != old_index && !(top - tse[ic] > cerr[old_index].eT + cerr[ic].eT)) {
          tie = true;
          if (getenv("ORACLE_AUDIT_VERBOSE"))
            fprintf(stderr, "D4 step %d: top %.17g (cluster %d, eT %.3e size %d) vs %.17g (cluster %d, eT %.3e size %d) diff %.3e\n", new_index, top,
                    old_index, cerr[old_index].eT, size[old_index], tse[ic], ic, cerr[ic].eT, size[ic], top - tse[ic]);
        }
      if (tie) audit->hit(4);
    }

    // Gather its points in ascending original order (:929-1019).
    cur_n = 0;
    for (int ip = 0; ip < U; ++ip) {
      if (member[ip] == (uint32_t)old_index) cur[cur_n++] = ip;
    }
    if (cur_n != size[old_index]) {
      fprintf(stderr, "Cluster to be split is expected to be of size %d not %d !\n", size[old_index], cur_n);
      abort();
    }
  }

  // Palette = rounded means of the non-empty clusters in index order (:1030-1065).
  const int shift = 8 - num_bits;
  int empty = 0, emitted = 0;
  for (int ic = 0; ic < K; ++ic) {
    if (size[ic] > 0) {
      Vec3 m = mean[ic];
      if (audit) {
        const double mm[3] = {mean[ic].r, mean[ic].g, mean[ic].b};
        bool flagged = false;
        for (int c = 0; c < 3; ++c) {
          const double v = mm[c] + 0.5;
          if (std::fabs(v - std::rint(v)) <= cerr[ic].eM + 512.0 * kU) {
            flagged = true;
            if (!resolve) audit->hit(5);
          }
        }
        if (flagged && resolve && K > 1) {  // the resolver's answer: the reference's own mean decides the rounding
          m = s_mean[ic];
          audit->resolved[0]++;
        }
      }
      uint32_t R = ((uint8_t)(m.r + 0.5)) << shift;
      uint32_t G = ((uint8_t)(m.g + 0.5)) << shift;
      uint32_t B = ((uint8_t)(m.b + 0.5)) << shift;
      colortable[emitted++] = (R << 16) | (G << 8) | B;
    } else {
      ++empty;
    }
  }
  *num_clusters = (uint32_t)(K - empty);
  return empty;
}

extern "C" int oracle_quant_varpart_fast(uint32_t num_pixels, const uint32_t *in, uint32_t num_rows,
                                         uint32_t num_cols, uint32_t *num_clusters, uint32_t *colortable,
                                         int num_bits, int dec_factor, int max_iters, int all_pixels_unique,
                                         oracle_split_record *records, int *num_records) {
  return varpart_impl(num_pixels, in, num_rows, num_cols, num_clusters, colortable, num_bits, dec_factor,
                      max_iters, all_pixels_unique, records, num_records, false);
}

// The device-arithmetic model with its tie audit: flags_out[0] = total, [1..5] = D1..D5 (see TieAudit).
extern "C" int oracle_quant_varpart_fast_exact_audit(uint32_t num_pixels, const uint32_t *in, uint32_t num_rows,
                                                     uint32_t num_cols, uint32_t *num_clusters, uint32_t *colortable,
                                                     int num_bits, int dec_factor, int max_iters, int all_pixels_unique,
                                                     uint32_t *flags_out) {
  TieAudit audit;
  audit.resolve = false;
  int rc = varpart_impl(num_pixels, in, num_rows, num_cols, num_clusters, colortable, num_bits, dec_factor, max_iters,
                        all_pixels_unique, nullptr, nullptr, true, &audit);
  for (int i = 0; i < 6; ++i) flags_out[i] = audit.flags[i];
  return rc;
}

// The device-arithmetic model WITH the resolver model (see TieAudit): flags_out[0..5] as above for what is left flagged
// (axis / hyperplane / TSE: the memberships may differ from the reference's, nothing is resolved then), [6] roundings taken
// from the reference's mean, [7] cuts confirmed, [8] cuts forced.  With flags_out[0] == 0 the palette is the reference's.
extern "C" int oracle_quant_varpart_fast_exact_resolved(uint32_t num_pixels, const uint32_t *in, uint32_t num_rows,
                                                        uint32_t num_cols, uint32_t *num_clusters, uint32_t *colortable,
                                                        int num_bits, int dec_factor, int max_iters, int all_pixels_unique,
                                                        uint32_t *flags_out) {
  TieAudit audit;
  audit.resolve = true;
  int rc = varpart_impl(num_pixels, in, num_rows, num_cols, num_clusters, colortable, num_bits, dec_factor, max_iters,
                        all_pixels_unique, nullptr, nullptr, true, &audit);
  for (int i = 0; i < 6; ++i) flags_out[i] = audit.flags[i];
  for (int i = 0; i < 3; ++i) flags_out[6 + i] = audit.resolved[i];
  return rc;
}

extern "C" int oracle_quant_varpart_fast_exact(uint32_t num_pixels, const uint32_t *in, uint32_t num_rows,
                                               uint32_t num_cols, uint32_t *num_clusters,
                                               uint32_t *colortable, int num_bits, int dec_factor,
                                               int max_iters, int all_pixels_unique,
                                               oracle_split_record *records, int *num_records) {
  return varpart_impl(num_pixels, in, num_rows, num_cols, num_clusters, colortable, num_bits, dec_factor,
                      max_iters, all_pixels_unique, records, num_records, true);
}

namespace {

// Reference Pixel_Int (DivQuantHeader.h:40-44): the element type std::sort permutes.  Keeping the
// same layout and comparator keeps libstdc++'s (unstable) introsort permutation identical.
struct PaletteEntry {
  int red, green, blue;
  int weight;  // r+g+b
};

inline bool by_sum(const PaletteEntry &a, const PaletteEntry &b) { return a.weight < b.weight; }

struct SearchTables {
  std::vector<PaletteEntry> sorted;
  int lut_init[3 * 255 + 1];
};

// Reference: map_colors_mps set-up, DivQuantMapColors.cpp:267-383.
void build_tables(const uint32_t *colortable, int n, SearchTables *t) {
  std::vector<PaletteEntry> v((size_t)n);
  for (int i = 0; i < n; ++i) {
    uint32_t p = colortable[i];
    v[i].red = (int)chan_r(p);
    v[i].green = (int)chan_g(p);
    v[i].blue = (int)chan_b(p);
    v[i].weight = v[i].red + v[i].green + v[i].blue;
  }
  std::sort(v.begin(), v.end(), by_sum);
  t->sorted = v;

  const int lut_n = 3 * 255 + 1;
  int low, high;
  // Entry 0 owns sums below the rounded midpoint of the first two sums; the last entry owns sums
  // from the rounded midpoint of the last two (:331-358). Single-colour palettes use 1 for both.
  low = (n >= 2) ? (int)(0.5 * (v[0].weight + v[1].weight) + 0.5) : 1;
  for (int k = 0; k < low; ++k) t->lut_init[k] = 0;
  high = (n >= 2) ? (int)(0.5 * (v[n - 2].weight + v[n - 1].weight) + 0.5) : 1;
  for (int k = high; k < lut_n; ++k) t->lut_init[k] = n - 1;
  for (int ic = 1; ic < n - 1; ++ic) {
    low = (int)(0.5 * (v[ic - 1].weight + v[ic].weight) + 0.5);
    high = (int)(0.5 * (v[ic].weight + v[ic + 1].weight) + 0.5);
    for (int k = low; k < high; ++k) t->lut_init[k] = ic;
  }
}

inline int dist2(int r, int g, int b, const PaletteEntry &e) {
  int d = r - e.red, acc = d * d;
  d = g - e.green;
  acc += d * d;
  d = b - e.blue;
  acc += d * d;
  return acc;
}

inline uint32_t pack(const PaletteEntry &e) {
  return ((uint32_t)(uint8_t)e.red << 16) | ((uint32_t)(uint8_t)e.green << 8) | (uint32_t)(uint8_t)e.blue;
}

}  // namespace

extern "C" void oracle_build_search_tables(const uint32_t *colortable, int num_colors, uint32_t *sorted_out,
                                           int32_t *lut_init_out) {
  SearchTables t;
  build_tables(colortable, num_colors, &t);
  for (int i = 0; i < num_colors; ++i) sorted_out[i] = pack(t.sorted[i]);
  for (int i = 0; i < 766; ++i) lut_init_out[i] = t.lut_init[i];
}

// Reference: map_colors_mps search loop, DivQuantMapColors.cpp:385-532.
extern "C" void oracle_map_colors_mps(const uint32_t *in, uint32_t num_pixels, uint32_t *out,
                                      const uint32_t *colortable, int num_colors) {
  assert(num_colors > 0);
  SearchTables t;
  build_tables(colortable, num_colors, &t);
  const PaletteEntry *pal = t.sorted.data();
  // Lower bound on the distance from the gap in channel sums: (int)(d*d/3.0)  (:285-311).
  std::vector<int> bound(2 * 765 + 1);
  for (int d = 0; d <= 765; ++d) bound[765 + d] = bound[765 - d] = (int)((d * d) / 3.0);
  const int *lb = bound.data() + 765;

  for (uint32_t i = 0; i < num_pixels; ++i) {
    uint32_t p = in[i];
    int r = (int)chan_r(p), g = (int)chan_g(p), b = (int)chan_b(p);
    int sum = r + g + b;
    int win = t.lut_init[sum];
    int best = dist2(r, g, b, pal[win]);
    int hi = win, lo = win;
    bool go_up = true, go_down = true;
    while (go_up || go_down) {
      if (go_up) {
        ++hi;
        if (hi > num_colors - 1 || lb[sum - pal[hi].weight] >= best) {
          go_up = false;
        } else {
          int d = dist2(r, g, b, pal[hi]);
          if (d < best) {
            best = d;
            win = hi;
          }
        }
      }
      if (go_down) {
        --lo;
        if (lo < 0 || lb[sum - pal[lo].weight] >= best) {
          go_down = false;
        } else {
          int d = dist2(r, g, b, pal[lo]);
          if (d < best) {
            best = d;
            win = lo;
          }
        }
      }
    }
    out[i] = pack(pal[win]);  // alpha cleared (:523-527)
  }
}

// Closed form of the same search: argmin over the whole sorted palette of
// (distance, visiting rank) with rank(s)=0, rank(s+d)=2d-1, rank(s-d)=2d.
extern "C" void oracle_map_colors_bruteforce(const uint32_t *in, uint32_t num_pixels, uint32_t *out,
                                             const uint32_t *colortable, int num_colors) {
  assert(num_colors > 0);
  SearchTables t;
  build_tables(colortable, num_colors, &t);
  const PaletteEntry *pal = t.sorted.data();
  for (uint32_t i = 0; i < num_pixels; ++i) {
    uint32_t p = in[i];
    int r = (int)chan_r(p), g = (int)chan_g(p), b = (int)chan_b(p);
    int s = t.lut_init[r + g + b];
    int64_t best_key = INT64_MAX;
    int win = s;
    for (int k = 0; k < num_colors; ++k) {
      int rank = (k == s) ? 0 : (k > s ? 2 * (k - s) - 1 : 2 * (s - k));
      int64_t key = ((int64_t)dist2(r, g, b, pal[k]) << 32) | (int64_t)rank;
      if (key < best_key) {
        best_key = key;
        win = k;
      }
    }
    out[i] = pack(pal[win]);
  }
}

// Reference: quant_recurse, quant_util.cpp:20-158.
extern "C" void oracle_quant_recurse(uint32_t num_pixels, const uint32_t *in, uint32_t *out,
                                     uint32_t *num_clusters, uint32_t *colortable, int all_pixels_unique) {
  oracle_quant_varpart_fast(num_pixels, in, 1, num_pixels, num_clusters, colortable, 8, 1, 10,
                            all_pixels_unique, nullptr, nullptr);
  // First occurrence of each palette word survives, order kept (:93-118).
  int n = (int)*num_clusters, kept = 0;
  for (int i = 0; i < n; ++i) {
    bool seen = false;
    for (int j = 0; j < kept && !seen; ++j) seen = (colortable[j] == colortable[i]);
    if (!seen) colortable[kept++] = colortable[i];
  }
  *num_clusters = (uint32_t)kept;
  oracle_map_colors_mps(in, num_pixels, out, colortable, kept);
}

// Reference: mapQuantPixelsToColortableIndexes, superpixels/OpenCVUtil.cpp:787-849 (pixel -> index
// map filled in palette order, so the LAST duplicate wins; alpha is forced opaque on both sides).
extern "C" int oracle_colortable_indexes(const uint32_t *quant_pixels, uint32_t num_pixels,
                                         const uint32_t *colortable, int num_colors, uint32_t *labels_out) {
  for (uint32_t i = 0; i < num_pixels; ++i) {
    uint32_t want = quant_pixels[i] & 0x00FFFFFFu;
    int found = -1;
    for (int k = 0; k < num_colors; ++k)
      if ((colortable[k] & 0x00FFFFFFu) == want) found = k;
    if (found < 0) return -1;
    labels_out[i] = (uint32_t)found;
  }
  return 0;
}

// Reference: genHistogramsForBlocks, ClusteringSegmentation/ClusteringSegmentation.cpp:417-563 (block size
// = ceil(dim/superpixelDim) as in ClusteringSegmentationMain.cpp:139-149).  Every superpixelDim x superpixelDim
// block (clipped at the image border) is represented by its most frequent quantized pixel; ties go to the
// first maximum in the iteration order of the std::unordered_map the counts live in -- the same libstdc++
// container here, filled in the same (row-major) order, hence the same order.
extern "C" void oracle_block_vote(const uint32_t *quant_pixels, uint32_t width, uint32_t height, uint32_t dim,
                                  uint32_t *block_out) {
  const uint32_t bw = (width + dim - 1) / dim, bh = (height + dim - 1) / dim;
  for (uint32_t by = 0; by < bh; ++by) {
    for (uint32_t bx = 0; bx < bw; ++bx) {
      std::vector<uint32_t> px;
      bool all_same = true;
      for (uint32_t y = by * dim; y < by * dim + dim; ++y)
        for (uint32_t x = bx * dim; x < bx * dim + dim; ++x) {
          if (x > width - 1 || y > height - 1) continue;
          const uint32_t q = quant_pixels[(size_t)y * width + x];
          if (!px.empty() && q != px[0]) all_same = false;
          px.push_back(q);
        }
      uint32_t best = 0;
      if (all_same) {
        best = px[0];
      } else {
        std::unordered_map<uint32_t, uint32_t> counts;
        for (uint32_t q : px) counts[q] += 1;
        int max_count = 0;
        for (auto it = counts.begin(); it != counts.end(); ++it) {
          if ((int)it->second > max_count) {
            max_count = (int)it->second;
            best = it->first;
          }
        }
      }
      block_out[(size_t)by * bw + bx] = best;
    }
  }
}

// Reference: SRM/srm.c:103-121 (diff = largest per-channel absolute difference, channels B,G,R at offset
// row*widthStep + channels*col), :135-177 (edge list: for every pixel of the (h-1) x (w-1) interior its right and
// lower neighbour, then the last column's lower neighbours, then the last row's right neighbours) and :226-246
// (stable 256-bin counting sort by diff).  pairs_out: 3 words per pair (r1, r2, diff).  Returns n_pairs.
extern "C" uint32_t oracle_srm_sorted_edges(const uint8_t *in, uint32_t width, uint32_t height, uint32_t channels,
                                            uint32_t width_step, uint32_t *pairs_out) {
  const uint32_t n = 2 * (width - 1) * (height - 1) + (height - 1) + (width - 1);
  if (!pairs_out) return n;
  std::vector<uint32_t> pairs((size_t)n * 3);
  auto diff = [&](uint32_t a, uint32_t b) {
    const uint8_t *pa = in + (size_t)(a / width) * width_step + (size_t)channels * (a % width);
    const uint8_t *pb = in + (size_t)(b / width) * width_step + (size_t)channels * (b % width);
    uint32_t d = 0;
    for (int c = 0; c < 3; ++c) {
      const uint32_t x = pa[c], y = pb[c];
      const uint32_t dc = x > y ? x - y : y - x;
      if (dc > d) d = dc;
    }
    return d;
  };
  size_t p = 0;
  auto emit = [&](uint32_t a, uint32_t b) {
    pairs[3 * p] = a;
    pairs[3 * p + 1] = b;
    pairs[3 * p + 2] = diff(a, b);
    ++p;
  };
  for (uint32_t i = 0; i + 1 < height; ++i)
    for (uint32_t j = 0; j + 1 < width; ++j) {
      emit(i * width + j, i * width + j + 1);
      emit(i * width + j, i * width + j + width);
    }
  for (uint32_t i = 0; i + 1 < height; ++i) emit(i * width + width - 1, i * width + width - 1 + width);
  for (uint32_t j = 0; j + 1 < width; ++j) emit((height - 1) * width + j, (height - 1) * width + j + 1);
  uint32_t start[257] = {0};
  for (size_t q = 0; q < n; ++q) start[pairs[3 * q + 2] + 1]++;
  for (int b = 0; b < 256; ++b) start[b + 1] += start[b];
  for (size_t q = 0; q < n; ++q) {
    const uint32_t at = start[pairs[3 * q + 2]]++;
    pairs_out[3 * at] = pairs[3 * q];
    pairs_out[3 * at + 1] = pairs[3 * q + 1];
    pairs_out[3 * at + 2] = pairs[3 * q + 2];
  }
  return n;
}

extern "C" uint64_t oracle_hash_words(const uint32_t *words, uint64_t n) {
  uint64_t h = 0xcbf29ce484222325ull;
  for (uint64_t i = 0; i < n; ++i) {
    h ^= words[i];
    h *= 0x100000001b3ull;
  }
  return h;
}

extern "C" void oracle_generate(int kind, uint32_t width, uint32_t height, uint64_t seed, uint32_t *out) {
  uint64_t s = seed;
  for (uint32_t y = 0; y < height; ++y) {
    for (uint32_t x = 0; x < width; ++x) {
      s += 0x9E3779B97F4A7C15ull;
      uint64_t z = s;
      z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
      z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
      z ^= z >> 31;
      uint32_t px;
      if (kind == 1) {
        int r = (int)(x * 255u / width), g = (int)(y * 255u / height);
        int b = (int)((x + y) * 255u / (width + height));
        int n = (int)(z % 9) - 4;
        r = std::min(255, std::max(0, r + n));
        g = std::min(255, std::max(0, g + n));
        b = std::min(255, std::max(0, b + n));
        px = 0xFF000000u + ((uint32_t)r << 16) + ((uint32_t)g << 8) + (uint32_t)b;
      } else {
        px = 0xFF000000u + (uint32_t)(z & 0xFFFFFFu);
      }
      out[(size_t)y * width + x] = px;
    }
  }
}
