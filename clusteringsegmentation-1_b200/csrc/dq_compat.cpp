// dq_compat.cpp -- the reference's own entry points, exported with the reference's own (C++-mangled)
// symbol names, forwarding to the C ABI of include/divquant_b200.h.
//
// A build of the reference that links libdivquant_b200.so instead of compiling DivQuant/*.cpp gets
// _Z18quant_varpart_fastjPKjPjjjS1_S1_iiii, _Z14map_colors_mpsPKjjPjS1_i, _Z16calc_color_tablePKjjPjjjiPi,
// _Z16get_double_scalePKjj, _Z8cut_bitsPKjjPjhhh, _Z17validate_num_bitsh, _Z9check_memi, _Z11start_timerv,
// _Z10stop_timerl, _Z8timediffll and extern "C" quant_recurse from here.
#include <stdio.h>
#include <stdlib.h>

#include <new>

#include "../../include/DivQuantHeader.h"
#include "../../include/divquant_b200.h"
#include "../../include/quant_util.h"

clock_t start_timer(void) { return clock(); }

double stop_timer(const clock_t start) { return ((double)(clock() - start)) / CLOCKS_PER_SEC; }

long timediff(clock_t t1, clock_t t2) { return (long)(((double)t2 - t1) / CLOCKS_PER_SEC * 1000); }

int validate_num_bits(const uchar num_bits) { return dq_validate_num_bits(num_bits); }

void check_mem(const int failed) {
  if (failed != 0) {
    fprintf(stderr, "Insufficient memory !\n");
    abort();
  }
}

double get_double_scale(const uint32_t *inPixels, const uint32_t numPixels) {
  return dq_get_double_scale(inPixels, numPixels);
}

void map_colors_mps(const uint32_t *inPixelsPtr, uint32_t numPixels, uint32_t *outPixelsPtr, uint32_t *outColortablePtr,
                    int colormapSize) {
  dq_map_colors_mps(inPixelsPtr, numPixels, outPixelsPtr, outColortablePtr, colormapSize);
}

double *calc_color_table(const uint32_t *inPixels, const uint32_t numPixels, uint32_t *outPixels, const uint32_t numRows,
                         const uint32_t numCols, const int dec_factor, int *num_colors) {
  if (dec_factor <= 0) {
    fprintf(stderr, "Decimation factor ( %d ) should be positive !\n", dec_factor);
    return NULL;
  }
  const uint32_t dec = (uint32_t)dec_factor;
  const size_t samples = (size_t)((numRows + dec - 1) / dec) * ((numCols + dec - 1) / dec);
  double *scratch = new double[samples ? samples : 1];
  dq_calc_color_table(inPixels, numPixels, outPixels, numRows, numCols, dec_factor, num_colors, scratch);
  // hand back an array of exactly U doubles, allocated with new[] as the reference does (:169)
  double *weights = new double[*num_colors > 0 ? *num_colors : 1];
  for (int i = 0; i < *num_colors; ++i) weights[i] = scratch[i];
  delete[] scratch;
  return weights;
}

void cut_bits(const uint32_t *inPixels, const uint32_t numPixels, uint32_t *outPixels, const uchar num_bits_red,
              const uchar num_bits_green, const uchar num_bits_blue) {
  dq_cut_bits(inPixels, numPixels, outPixels, num_bits_red, num_bits_green, num_bits_blue);
}

void quant_varpart_fast(const uint32_t numPixels, const uint32_t *inPixels, uint32_t *tmpPixels, const uint32_t numRows,
                        const uint32_t numCols, uint32_t *numClustersPtr, uint32_t *colortablePtr, const int num_bits,
                        const int dec_factor, const int max_iters, const int allPixelsUnique) {
  dq_quant_varpart_fast(numPixels, inPixels, tmpPixels, numRows, numCols, numClustersPtr, colortablePtr, num_bits, dec_factor,
                        max_iters, allPixelsUnique);
}

extern "C" void quant_recurse(uint32_t numPixels, const uint32_t *inPixelsPtr, uint32_t *outPixelsPtr,
                              uint32_t *numClustersPtr, uint32_t *outColortablePtr, int allPixelsUnique) {
  dq_quant_recurse(numPixels, inPixelsPtr, outPixelsPtr, numClustersPtr, outColortablePtr, allPixelsUnique);
}
