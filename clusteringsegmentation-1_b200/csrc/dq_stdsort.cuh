// dq_stdsort.cuh -- libstdc++'s std::sort, replayed step for step, for the <= K palette entries.
//
// map_colors_mps sorts the palette by r+g+b with std::sort (DivQuantMapColors.cpp:314-323, sort_color :227-238).  std::sort
// is not stable and equal sums are everywhere (the 125-colour grid has 13 distinct sums), and WHICH of two equidistant
// palette colours a pixel gets depends on their order, so the remap is bit-exact only with the very
// permutation the reference's build produces.  The oracle and the compiled reference here are libstdc++ (GCC 13); on the
// host the shim simply calls the same std::sort.  This is the same algorithm written out, so that the frame pipeline can
// keep the palette on the device (no host round trip between the split and the remap):
//   introsort: median-of-three pivot to the front, unguarded Hoare partition, recursion on the right part, depth limit
//   2 floor(log2 n) with heap sort beyond it, ranges of <= 16 left to one final insertion sort (bits/stl_algo.h:
//   __introsort_loop, __unguarded_partition_pivot, __move_median_to_first, __final_insertion_sort; bits/stl_heap.h).
// Elements are words  key << 16 | payload ; only the key takes part in comparisons, exactly like the reference's comparator,
// which looks at Pixel_Int::weight alone.  tests/test_stdsort.py checks it against std::sort itself on the CPU
// (random, few distinct keys, sorted / reversed / organ-pipe inputs and median-of-three killers that reach the heap sort).
#pragma once

#include <stdint.h>

#ifndef __CUDACC__
#define __host__
#define __device__
#endif

namespace dq {
namespace stdsort {

__host__ __device__ inline bool less(uint32_t a, uint32_t b) { return (a >> 16) < (b >> 16); }
__host__ __device__ inline void swap_at(uint32_t *v, int i, int j) {
  const uint32_t t = v[i];
  v[i] = v[j];
  v[j] = t;
}

// bits/stl_heap.h
__host__ __device__ inline void push_heap(uint32_t *first, int hole, int top, uint32_t value) {
  int parent = (hole - 1) / 2;
  while (hole > top && less(first[parent], value)) {
    first[hole] = first[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  first[hole] = value;
}
__host__ __device__ inline void adjust_heap(uint32_t *first, int hole, int len, uint32_t value) {
  const int top = hole;
  int second = hole;
  while (second < (len - 1) / 2) {
    second = 2 * (second + 1);
    if (less(first[second], first[second - 1])) second--;
    first[hole] = first[second];
    hole = second;
  }
  if ((len & 1) == 0 && second == (len - 2) / 2) {
    second = 2 * (second + 1);
    first[hole] = first[second - 1];
    hole = second - 1;
  }
  push_heap(first, hole, top, value);
}
// __partial_sort(first, last, last): __heap_select degenerates to __make_heap, then __sort_heap
__host__ __device__ inline void heap_sort(uint32_t *first, int len) {
  if (len >= 2) {
    int parent = (len - 2) / 2;
    for (;;) {
      const uint32_t value = first[parent];
      adjust_heap(first, parent, len, value);
      if (parent == 0) break;
      parent--;
    }
  }
  int last = len;
  while (last > 1) {
    --last;
    const uint32_t value = first[last];
    first[last] = first[0];
    adjust_heap(first, 0, last, value);
  }
}

// bits/stl_algo.h
__host__ __device__ inline void move_median_to_first(uint32_t *v, int result, int a, int b, int c) {
  if (less(v[a], v[b])) {
    if (less(v[b], v[c])) swap_at(v, result, b);
    else if (less(v[a], v[c])) swap_at(v, result, c);
    else swap_at(v, result, a);
  } else if (less(v[a], v[c])) {
    swap_at(v, result, a);
  } else if (less(v[b], v[c])) {
    swap_at(v, result, c);
  } else {
    swap_at(v, result, b);
  }
}
__host__ __device__ inline int unguarded_partition(uint32_t *v, int first, int last, int pivot) {
  for (;;) {
    while (less(v[first], v[pivot])) ++first;
    --last;
    while (less(v[pivot], v[last])) --last;
    if (!(first < last)) return first;
    swap_at(v, first, last);
    ++first;
  }
}
__host__ __device__ inline void unguarded_linear_insert(uint32_t *v, int last) {
  const uint32_t val = v[last];
  int next = last - 1;
  while (less(val, v[next])) {
    v[last] = v[next];
    last = next;
    --next;
  }
  v[last] = val;
}
__host__ __device__ inline void insertion_sort(uint32_t *v, int first, int last) {
  if (first == last) return;
  for (int i = first + 1; i != last; ++i) {
    if (less(v[i], v[first])) {
      const uint32_t val = v[i];
      for (int j = i; j > first; --j) v[j] = v[j - 1];  // move_backward(first, i, i + 1)
      v[first] = val;
    } else {
      unguarded_linear_insert(v, i);
    }
  }
}

constexpr int kThreshold = 16;  // _S_threshold

// std::sort(v, v + n, by key).  n <= 65536 (the explicit stack holds one frame per level of the depth limit).
__host__ __device__ inline void sort(uint32_t *v, int n) {
  if (n <= 0) return;
  int lg = 0;
  while ((n >> (lg + 1)) != 0) ++lg;  // std::__lg
  // __introsort_loop(first, last, 2 lg): the recursion on [cut, last) becomes a stack of pending ranges; the ranges are
  // disjoint, so the order in which they are worked off does not change what any of them ends up as
  int stack_first[40], stack_last[40], stack_depth[40];
  int sp = 0;
  stack_first[sp] = 0, stack_last[sp] = n, stack_depth[sp] = 2 * lg;
  ++sp;
  while (sp > 0) {
    --sp;
    const int first = stack_first[sp];
    int last = stack_last[sp], depth = stack_depth[sp];
    while (last - first > kThreshold) {
      if (depth == 0) {
        heap_sort(v + first, last - first);
        break;
      }
      --depth;
      const int mid = first + (last - first) / 2;
      move_median_to_first(v, first, first + 1, mid, last - 1);
      const int cut = unguarded_partition(v, first + 1, last, first);
      stack_first[sp] = cut, stack_last[sp] = last, stack_depth[sp] = depth;  // __introsort_loop(cut, last, depth)
      ++sp;
      last = cut;
    }
  }
  // __final_insertion_sort
  if (n > kThreshold) {
    insertion_sort(v, 0, kThreshold);
    for (int i = kThreshold; i != n; ++i) unguarded_linear_insert(v, i);
  } else {
    insertion_sort(v, 0, n);
  }
}

}  // namespace stdsort
}  // namespace dq
