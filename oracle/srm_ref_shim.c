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

/* BASELINE config 5 harness ("GPU front half feeding the reference's host SRM"): the reference's SRM with the sorted edge
 * list SUPPLIED by the caller (the GPU's, dq_srm_sorted_edges) instead of built by segmentation() (srm.c:135-177, 226-246).
 * Everything else is the reference's own code, called in srm_run()'s order (srm.c:78-88): initialize, the union-find merge
 * loop over the pairs (restated below from srm.c:179-190 -- six lines, all calls into the reference), merge_small_regions,
 * finalize. */
void initialize(struct srm *srm);
void merge_small_regions(struct srm *srm);
void finalize(struct srm *srm);
unsigned int merge_predicate(struct srm *srm, unsigned int reg1, unsigned int reg2);
void merge_regions(struct srm *srm, unsigned int r1, unsigned int r2);
#include "unionfind.h"

void ref_srm_run_with_pairs(double Q, unsigned int width, unsigned int height, unsigned int channels, unsigned int width_step,
                            uint8_t *in, uint8_t *out, const uint32_t *pairs) {
  struct srm *s = srm_new(Q, width, height, channels, 0);
  s->in = in;
  s->widthStep_in = width_step;
  s->out = out;
  s->widthStep_out = width_step;
  initialize(s);
  for (unsigned int i = 0; i < s->n_pairs; i++) { /* srm.c:179-190 */
    unsigned int reg1 = unionfind_find(s->uf, pairs[3 * i + 0]);
    unsigned int reg2 = unionfind_find(s->uf, pairs[3 * i + 1]);
    if ((reg1 != reg2) && (merge_predicate(s, reg1, reg2))) merge_regions(s, reg1, reg2);
  }
  merge_small_regions(s);
  finalize(s);
  srm_delete(s);
}
