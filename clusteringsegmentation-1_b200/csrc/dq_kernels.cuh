// dq_kernels.cuh -- host-callable launchers of the histogram and remap kernels.
#pragma once

#include "dq_common.cuh"

namespace dq {

// ---- dq_hist.cu ----
void hist_insert(const uint32_t *d_in, uint32_t n, uint32_t num_rows, uint32_t num_cols, uint32_t dec, int num_bits,
                 uint32_t *d_table, uint32_t *d_uniq, uint32_t *d_ucount, int sm_count, cudaStream_t st,
                 uint32_t *d_seen_init = nullptr);
void hist_collect(const uint32_t *d_uniq, const uint32_t *d_ucount, uint32_t u_hint, uint32_t *d_table, uint2 *d_pts,
                  bool clear_table, int sm_count, cudaStream_t st);
void table_clear(const uint32_t *d_uniq, const uint32_t *d_ucount, uint32_t u_hint, uint32_t *d_table, int sm_count,
                 cudaStream_t st);
void points_from_pixels(const uint32_t *d_in, uint32_t n, uint2 *d_pts, int sm_count, cudaStream_t st);
void cut_bits_device(const uint32_t *d_in, uint32_t n, uint32_t *d_out, int rbits, int gbits, int bbits, int sm_count,
                     cudaStream_t st);
void order_keys(const uint32_t *d_in, uint32_t num_rows, uint32_t num_cols, uint32_t dec, int num_bits, const uint32_t *d_uniq,
                const uint2 *d_pts, uint32_t u, uint32_t *d_table, uint64_t *d_keys, int sm_count, cudaStream_t st);

// Sorts the u (unique) keys of order_keys with their point index (d_keys / d_vals need room for the next power of two),
// then writes the unique colours and their weights norm * count in that order: calc_color_table's output, on the device.
void order_sort(uint64_t *d_keys, uint32_t *d_vals, uint32_t u, int sm_count, cudaStream_t st);
void order_sort_emit(const uint2 *d_pts, uint32_t u, uint64_t *d_keys, uint32_t *d_vals, double norm, uint32_t *d_colours,
                     double *d_weights, int sm_count, cudaStream_t st);
void hist_export(const uint2 *d_pts, const uint32_t *d_ucount, uint32_t u_hint, uint32_t *d_colours, uint32_t *d_counts,
                 int sm_count, cudaStream_t st);
void hist_export_padded(const uint2 *d_pts, const uint32_t *d_ucount, uint32_t cap, uint32_t *d_colours, uint32_t *d_counts,
                        uint32_t *d_overflow, int sm_count, cudaStream_t st);
void hist_merge(const uint32_t *d_colours, const uint32_t *d_counts, uint32_t num_entries, uint32_t *d_table, uint32_t *d_uniq,
                uint32_t *d_ucount, int sm_count, cudaStream_t st);

// ---- dq_map.cu ----
int map_smem_palette_limit();
void map_pixels(const uint32_t *d_in, uint32_t n, uint32_t *d_out, const uint32_t *d_sorted, int num_colors,
                const int *d_lut, int4 *d_pal_scratch, int sm_count, cudaStream_t st);
void map_unique(const uint32_t *d_uniq, const uint32_t *d_ucount, uint32_t u_hint, uint32_t *d_table,
                const uint32_t *d_sorted, int num_colors, const int *d_lut, int sm_count, cudaStream_t st);
// Sorted palette and start-index table small enough (K <= 256) to travel inside the kernel's parameter block: no
// upload, no extra dependency between the host's table building and the launch.
struct MapTablesParam {
  uint32_t sorted[256];
  uint16_t lut[768];  // lut_init[0..765], values < K
};
void map_unique_params(const MapTablesParam &tables, const uint32_t *d_uniq, const uint32_t *d_ucount, uint32_t u_hint,
                       uint32_t *d_table, int num_colors, int sm_count, cudaStream_t st);
// ---- palette handling on the device (frame pipeline): quant_util.cpp:93-118 + DivQuantMapColors.cpp:267-383 ----
// What a frame of the pipeline hands back to the host in one copy.
constexpr int kFrameMaxColors = 512;
struct FrameResult {
  uint32_t num_colors;      // palette entries after the duplicates are dropped
  uint32_t num_points;      // U
  uint32_t result[4];       // SplitArgs::result
  uint32_t ctl[16];         // SplitArgs::ctl (kCtlWords <= 16)
  uint32_t palette[kFrameMaxColors];  // first occurrence of every word, order kept (quant_util.cpp:93-118)
};
// One CTA: drops duplicate palette words, orders the rest by r+g+b exactly as the reference's std::sort does
// (dq_stdsort.cuh), builds lut_init.  d_sorted: [max_colors palette words | 766 lut entries] as upload_search_tables lays
// them out; d_frame receives the FrameResult.
void palette_post(const uint32_t *d_palette, const uint32_t *d_result, const uint32_t *d_ctl, const uint32_t *d_ucount,
                  int max_colors, uint32_t *d_sorted, FrameResult *d_frame, cudaStream_t st);
// map_unique with the number of colours read from the device (FrameResult::num_colors).
void map_unique_dev(const uint32_t *d_uniq, const uint32_t *d_ucount, uint32_t u_hint, uint32_t *d_table, const uint32_t *d_sorted,
                    int max_colors, const FrameResult *d_frame, int sm_count, cudaStream_t st);
void block_vote(const uint32_t *d_quant, uint32_t width, uint32_t height, uint32_t dim, uint32_t *d_blocks, int sm_count,
                cudaStream_t st);
void map_labels(const uint32_t *d_in, uint32_t n, uint32_t *d_out, const uint2 *d_pairs, int num_pairs, int greyscale,
                uint32_t *d_error, int sm_count, cudaStream_t st);
void map_gather(const uint32_t *d_in, uint32_t n, uint32_t *d_out, const uint32_t *d_table, int sm_count, cudaStream_t st);

// ---- dq_srm.cu ----
uint32_t srm_num_pairs(uint32_t width, uint32_t height);
size_t srm_scratch_words(uint32_t width, uint32_t height);
// Returns the number of kernels launched.  d_pairs: 3 words per edge (r1, r2, diff) in merge order.
int srm_sorted_edges(const uint8_t *d_in, uint32_t width, uint32_t height, uint32_t channels, uint32_t width_step,
                     uint32_t *d_pairs, uint32_t *d_scratch, cudaStream_t st);

}  // namespace dq
