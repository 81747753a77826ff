/* oracle/divquant_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * C-ABI of the CPU restatement of the reference's DivQuant path (see divquant_oracle.cpp).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library, and only as the checker.  The product (csrc/) never links or calls it.
 */
#ifndef DIVQUANT_ORACLE_H
#define DIVQUANT_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* One record per executed split of the divisive phase (reference loop DivQuantCluster.cpp:346-1027). */
typedef struct {
  int32_t new_index;     /* cluster created by this split                         */
  int32_t old_index;     /* cluster that was split                                */
  int32_t cut_axis;      /* 0=R 1=G 2=B                                           */
  int32_t num_points;    /* unique-colour points in the cluster before the split  */
  int32_t new_size;      /* points that ended in new_index after the last LKM it. */
  int32_t is_last;       /* 1 when new_index == K-1 (variance/TSE not computed)   */
  double cut_pos;
  double total_weight, new_weight, old_weight;
  double new_mean[3], old_mean[3];
  double new_var[3], old_var[3];   /* undefined when is_last */
  double new_tse, old_tse;         /* undefined when is_last */
} oracle_split_record;

/* calc_color_table (DivQuantMapColors.cpp:82-203).  Returns U, or -1 on dec_factor<=0.
 * unique_out / weights_out need capacity for every sampled pixel. counts_out may be NULL. */
int oracle_calc_color_table(const uint32_t *in, uint32_t num_pixels, uint32_t num_rows, uint32_t num_cols,
                            int dec_factor, uint32_t *unique_out, double *weights_out, uint32_t *counts_out);

/* cut_bits (DivQuantUni.cpp:28-100). Returns 0 when a bit count is rejected (output untouched). */
int oracle_cut_bits(const uint32_t *in, uint32_t num_pixels, uint32_t *out, int rbits, int gbits, int bbits);

/* quant_varpart_fast (DivQuantCluster.cpp:1099-1179).  *num_clusters in: requested K, out: actual.
 * records (capacity >= K-1) may be NULL; *num_records receives the number of splits executed.
 * Returns the number of empty clusters that were dropped. */
int oracle_quant_varpart_fast(uint32_t num_pixels, const uint32_t *in, uint32_t num_rows, uint32_t num_cols,
                              uint32_t *num_clusters, uint32_t *colortable, int num_bits, int dec_factor,
                              int max_iters, int all_pixels_unique, oracle_split_record *records,
                              int *num_records);

/* NOT the reference's arithmetic: a CPU model of the device path's arithmetic, kept here so that CPU
 * tests can measure how far it is from the reference.  Identical control flow and scalar formulas,
 * but in the weighted path every sum over points is the exact integer sum of count*c (and count*c*c),
 * scaled once by 1/#samples -- i.e. the reference's own uniform-weight arithmetic applied to
 * (colour, count) points.  Summation order therefore cannot matter. */
int oracle_quant_varpart_fast_exact(uint32_t num_pixels, const uint32_t *in, uint32_t num_rows,
                                    uint32_t num_cols, uint32_t *num_clusters, uint32_t *colortable,
                                    int num_bits, int dec_factor, int max_iters, int all_pixels_unique,
                                    oracle_split_record *records, int *num_records);
/* The same model with the device's tie audit (csrc/dq_tie.cuh): flags_out[0] = decisions inside the reference's rounding
 * noise in all, [1..5] = by kind (axis :388-403, cut :473, hyperplane :683, TSE arg-max :876-887, rounding :1050-1052). */
int oracle_quant_varpart_fast_exact_audit(uint32_t num_pixels, const uint32_t *in, uint32_t num_rows, uint32_t num_cols,
                                          uint32_t *num_clusters, uint32_t *colortable, int num_bits, int dec_factor,
                                          int max_iters, int all_pixels_unique, uint32_t *flags_out);
/* ... and with the model of the device's resolver (csrc/dq_resolve.cu, forced cuts of csrc/dq_context.cu): the reference's
 * own sequential sums are carried for the model's memberships; a flagged rounding rounds the reference's mean, a flagged
 * cut is the reference's when the two means have different floors.  flags_out[0..5] = what is LEFT flagged (kinds 1, 3, 4),
 * [6] roundings resolved, [7] cuts confirmed, [8] cuts forced.  flags_out[0] == 0 => the palette is the reference's. */
int oracle_quant_varpart_fast_exact_resolved(uint32_t num_pixels, const uint32_t *in, uint32_t num_rows, uint32_t num_cols,
                                             uint32_t *num_clusters, uint32_t *colortable, int num_bits, int dec_factor,
                                             int max_iters, int all_pixels_unique, uint32_t *flags_out);

/* map_colors_mps (DivQuantMapColors.cpp:243-539), restated as the reference's pruned two-way search. */
void oracle_map_colors_mps(const uint32_t *in, uint32_t num_pixels, uint32_t *out, const uint32_t *colortable,
                           int num_colors);

/* Same result through the closed form argmin_k (dist, rank) of SURVEY.md section 8a (no pruning). */
void oracle_map_colors_bruteforce(const uint32_t *in, uint32_t num_pixels, uint32_t *out,
                                  const uint32_t *colortable, int num_colors);

/* Sorted palette and start-index table exactly as map_colors_mps builds them
 * (DivQuantMapColors.cpp:314-383).  sorted_out[num_colors], lut_init_out[766]. */
void oracle_build_search_tables(const uint32_t *colortable, int num_colors, uint32_t *sorted_out,
                                int32_t *lut_init_out);

/* quant_recurse (quant_util.cpp:20-158): quantize (max_iters 10, 8 bits, no decimation),
 * first-occurrence palette dedup, remap.  Silent (no timing lines). */
void oracle_quant_recurse(uint32_t num_pixels, const uint32_t *in, uint32_t *out, uint32_t *num_clusters,
                          uint32_t *colortable, int all_pixels_unique);

/* Label image: index of each mapped pixel in the caller's palette, last duplicate wins
 * (superpixels/OpenCVUtil.cpp:787-849). Returns 0, or -1 if a pixel is not in the palette. */
int oracle_colortable_indexes(const uint32_t *quant_pixels, uint32_t num_pixels, const uint32_t *colortable,
                              int num_colors, uint32_t *labels_out);

/* Block majority vote of genHistogramsForBlocks (ClusteringSegmentation.cpp:417-563): block_out has
 * ceil(width/dim) * ceil(height/dim) entries, row-major. */
void oracle_block_vote(const uint32_t *quant_pixels, uint32_t width, uint32_t height, uint32_t dim, uint32_t *block_out);

/* Sorted edge list of SRM's segmentation() (SRM/srm.c:135-177 + :226-246): 3 words per pair (r1, r2, diff); returns
 * n_pairs = 2(w-1)(h-1) + (h-1) + (w-1); pairs_out may be NULL to query the count. */
uint32_t oracle_srm_sorted_edges(const uint8_t *in, uint32_t width, uint32_t height, uint32_t channels, uint32_t width_step,
                                 uint32_t *pairs_out);

/* FNV-1a style fingerprint over u32 words used by SURVEY.md section 8c. */
uint64_t oracle_hash_words(const uint32_t *words, uint64_t n);

/* Synthetic generators of SURVEY.md section 8d (splitmix64). kind 1 = G1 natural-like, 2 = G2 uniform. */
void oracle_generate(int kind, uint32_t width, uint32_t height, uint64_t seed, uint32_t *out);

#ifdef __cplusplus
}
#endif
#endif
