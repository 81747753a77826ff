/* DivQuantHeader.h -- drop-in replacement for the reference's DivQuant/DivQuantHeader.h.
 *
 * Written from scratch for libdivquant_b200.so: it declares the same C++-linkage functions, with the
 * same signatures, that reference callers (ClusteringSegmentation/ClusteringSegmentation.cpp, the
 * XCTest files) include today, so that they compile and link unchanged against the B200 library.
 * Reference declarations: DivQuant/DivQuantHeader.h:33-96.  The functions are implemented in
 * clusteringsegmentation-1_b200/csrc/dq_compat.cpp on top of the C ABI in divquant_b200.h.
 */
#ifndef DivQuantHeader_h
#define DivQuantHeader_h

#include <stdint.h>
#include <time.h>

#include <vector>

#define MAX_RGB (255)
#define MAX_RGB_SQR (65025)
#define MAX_COLORS (256)

typedef unsigned char uchar;
typedef unsigned short ushort;
typedef unsigned int uint;
typedef unsigned long ulong;

typedef struct {
  int red, green, blue;
  int weight;
} Pixel_Int;

typedef struct {
  double red, green, blue;
  double weight;
} Pixel_Double;

#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

/* DivQuantMisc.cpp:18-46 */
clock_t start_timer(void);
double stop_timer(const clock_t start);
long timediff(clock_t t1, clock_t t2);
int validate_num_bits(const uchar num_bits);

/* DivQuantMapColors.cpp:43-51, 205-220 */
void check_mem(const int failed);
double get_double_scale(const uint32_t *inPixels, const uint32_t numPixels);

/* DivQuantMapColors.cpp:243-539 */
void map_colors_mps(const uint32_t *inPixelsPtr, uint32_t numPixels, uint32_t *outPixelsPtr,
                    uint32_t *outColortablePtr, int colormapSize);

/* DivQuantMapColors.cpp:82-203.  Returns new double[*num_colors]; the caller delete[]s it. */
double *calc_color_table(const uint32_t *inPixels, const uint32_t numPixels, uint32_t *outPixels,
                         const uint32_t numRows, const uint32_t numCols, const int dec_factor, int *num_colors);

/* DivQuantUni.cpp:28-100 */
void cut_bits(const uint32_t *inPixels, const uint32_t numPixels, uint32_t *outPixels, const uchar num_bits_red,
              const uchar num_bits_green, const uchar num_bits_blue);

/* DivQuantCluster.cpp:1099-1179.
 * Deviations from the reference, both fatal here instead of silently degenerate there:
 *  - max_iters must be in 1..32.  The reference accepts 0, but its KM template flag is hard-wired true (:1154-1166), so
 *    with no local k-means pass member[] is never written: every split leaves all points in the old cluster and the call
 *    returns one colour plus "# empty clusters: K-1" (SURVEY.md 7).  Nothing calls it that way (quant_util.cpp:31 ships 10).
 *  - dec_factor <= 0: message + abort (the reference dereferences the NULL calc_color_table returns, :1136).
 * The entry points of this header run on one lazily created context and are serialised by a lock: safe to call from
 * several threads, but one call runs at a time (the reference's functions are re-entrant); concurrent streams of frames go
 * through divquant_b200.h (dq_context_create / dq_pipeline_*). */
void quant_varpart_fast(const uint32_t numPixels, const uint32_t *inPixels, uint32_t *tmpPixels,
                        const uint32_t numRows, const uint32_t numCols, uint32_t *numClustersPtr,
                        uint32_t *colortablePtr, const int num_bits, const int dec_factor, const int max_iters,
                        const int allPixelsUnique);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif

#endif /* DivQuantHeader_h */
