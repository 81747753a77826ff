// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE ONLY.
//
// extern "C" trampolines onto the reference's C++-linkage DivQuant entry points
// (DivQuant/DivQuantHeader.h:52-96), so that tests/ and bench.py's cpu_baseline leg can drive the
// UNMODIFIED reference build (oracle/_ref/libdivquant_ref.so) through ctypes.  This file is compiled
// together with the reference sources where they lie under /root/reference; it contains no
// reference code, only calls.
#include "DivQuantHeader.h"
#include "quant_util.h"

#include <cstring>

extern "C" {

void ref_quant_recurse(uint32_t n, const uint32_t *in, uint32_t *out, uint32_t *num_clusters,
                       uint32_t *colortable, int all_unique) {
  quant_recurse(n, in, out, num_clusters, colortable, all_unique);
}

void ref_quant_varpart_fast(uint32_t n, const uint32_t *in, uint32_t *tmp, uint32_t rows, uint32_t cols,
                            uint32_t *num_clusters, uint32_t *colortable, int num_bits, int dec_factor,
                            int max_iters, int all_unique) {
  quant_varpart_fast(n, in, tmp, rows, cols, num_clusters, colortable, num_bits, dec_factor, max_iters,
                     all_unique);
}

void ref_map_colors_mps(const uint32_t *in, uint32_t n, uint32_t *out, uint32_t *colortable, int k) {
  map_colors_mps(in, n, out, colortable, k);
}

// Returns the number of unique colours; weights_out (capacity >= n) receives the reference's doubles.
int ref_calc_color_table(const uint32_t *in, uint32_t n, uint32_t *unique_out, uint32_t rows, uint32_t cols,
                         int dec_factor, double *weights_out) {
  int num = 0;
  double *w = calc_color_table(in, n, unique_out, rows, cols, dec_factor, &num);
  if (w == nullptr) return -1;
  if (weights_out) std::memcpy(weights_out, w, sizeof(double) * (size_t)num);
  delete[] w;
  return num;
}

void ref_cut_bits(const uint32_t *in, uint32_t n, uint32_t *out, int rbits, int gbits, int bbits) {
  cut_bits(in, n, out, (uchar)rbits, (uchar)gbits, (uchar)bbits);
}

double ref_get_double_scale(const uint32_t *in, uint32_t n) { return get_double_scale(in, n); }

}  // extern "C"
