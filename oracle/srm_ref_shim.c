/* oracle/srm_ref_shim.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Drives the UNMODIFIED reference SRM (SRM/srm.c + SRM/unionfind.c, compiled where they lie under /root/reference
 * by oracle/Makefile into oracle/_ref/libsrm_ref.so) and hands back the sorted edge list it builds inside
 * segmentation() (srm.c:135-177 edge generation, :226-246 bucket sort) plus the segmented image.  Contains no
 * reference code, only calls and reads of the public struct of SRM/srm.h. */
#include <stdint.h>
#include <string.h>

#include "srm.h"

/* Returns n_pairs; pairs_out (3 words per pair: r1, r2, diff) may be NULL to only query the count. */
unsigned int ref_srm_sorted_edges(double Q, unsigned int width, unsigned int height, unsigned int channels,
                                  unsigned int width_step, uint8_t *in, uint8_t *out, uint32_t *pairs_out) {
  struct srm *s = srm_new(Q, width, height, channels, 0);
  unsigned int n = s->n_pairs;
  if (pairs_out) {
    srm_run(s, width_step, in, width_step, out);
    memcpy(pairs_out, s->ordered_pairs, (size_t)n * sizeof(struct my_pair));
  }
  srm_delete(s);
  return n;
}
