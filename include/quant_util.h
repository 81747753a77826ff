/* quant_util.h -- drop-in replacement for the reference's DivQuant/quant_util.h (C or C++).
 * Reference declaration: DivQuant/quant_util.h:10; implementation replaced: quant_util.cpp:20-158. */
#ifndef quant_util_h
#define quant_util_h

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif
void quant_recurse(uint32_t numPixels, const uint32_t *inPixelsPtr, uint32_t *outColorTableOffsetPtr,
                   uint32_t *numClustersPtr, uint32_t *outColortablePtr, int allPixelsUnique);
#if defined(__GNUC__)
#pragma GCC visibility pop
#endif

#ifdef __cplusplus
}
#endif

#endif /* quant_util_h */
