// mstep_math.hpp -- the K-length pieces of the M step, written once for host and device.
//
// The device M step (mstep.cu) must order the sticks of StickBreak / GDirichlet exactly as the reference's
// `std::sort(..., greater-count-first)` does (src/distributions.cpp:146), ties included: an exact tie (two empty
// clusters) decides which cluster gets which E[log pi].  std::sort is not stable, so this header restates the
// libstdc++ algorithm (introsort with a median-of-three pivot, threshold 16, heap-sort fallback, final insertion
// sort) on an index array; tests/test_abi.py checks it against std::sort on the host.
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define LCB_HD __host__ __device__ __forceinline__
#else
#define LCB_HD inline
#endif

namespace lcb {
namespace mm {

// psi(x), x > 0: upward recurrence to x >= 8, then the Stirling series (same code as host_model.cpp)
LCB_HD double digamma(double x) {
  double acc = 0.0;
  for (; x < 8.0; x += 1.0) acc += 1.0 / x;
  const double i2 = 1.0 / (x * x);
  const double series =
      i2 * (1.0 / 12 - i2 * (1.0 / 120 - i2 * (1.0 / 252 - i2 * (1.0 / 240 - i2 * (1.0 / 132 -
      i2 * (691.0 / 32760 - i2 * (1.0 / 12 - i2 * (3617.0 / 8160))))))));
  return log(x) - 0.5 / x - series - acc;
}

// ceil(log2(x)) for finite x > 0, exactly
LCB_HD int ceil_log2(double x) {
  int e;
  const double f = frexp(x, &e);  // x = f 2^e, f in [0.5, 1)
  return f == 0.5 ? e - 1 : e;
}
// floor(log2(x)) for finite x > 0, exactly
LCB_HD int floor_log2(double x) {
  int e;
  (void)frexp(x, &e);
  return e - 1;
}

// ---- std::sort of libstdc++ on (index, count) pairs, "greater count first" ---------------------------------------
struct DescSorter {
  int* id;          // permutation being sorted
  const double* v;  // counts, addressed through id

  LCB_HD bool lt(int a, int b) const { return v[a] > v[b]; }  // comp(a, b)
  LCB_HD void swp(int i, int j) {
    const int t = id[i];
    id[i] = id[j];
    id[j] = t;
  }
  LCB_HD void median_to_first(int result, int a, int b, int c) {
    if (lt(id[a], id[b])) {
      if (lt(id[b], id[c])) swp(result, b);
      else if (lt(id[a], id[c])) swp(result, c);
      else swp(result, a);
    } else if (lt(id[a], id[c])) swp(result, a);
    else if (lt(id[b], id[c])) swp(result, c);
    else swp(result, b);
  }
  LCB_HD int partition(int first, int last, int pivot) {
    for (;;) {
      while (lt(id[first], id[pivot])) ++first;
      --last;
      while (lt(id[pivot], id[last])) --last;
      if (!(first < last)) return first;
      swp(first, last);
      ++first;
    }
  }
  LCB_HD void push_heap(int first, int hole, int top, int value) {
    int parent = (hole - 1) / 2;
    while (hole > top && lt(id[first + parent], value)) {
      id[first + hole] = id[first + parent];
      hole = parent;
      parent = (hole - 1) / 2;
    }
    id[first + hole] = value;
  }
  LCB_HD void adjust_heap(int first, int hole, int len, int value) {
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
      child = 2 * (child + 1);
      if (lt(id[first + child], id[first + child - 1])) --child;
      id[first + hole] = id[first + child];
      hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
      child = 2 * (child + 1);
      id[first + hole] = id[first + child - 1];
      hole = child - 1;
    }
    push_heap(first, hole, top, value);
  }
  LCB_HD void heap_sort(int first, int last) {  // std::__partial_sort(first, last, last)
    const int len = last - first;
    if (len >= 2) {
      for (int parent = (len - 2) / 2;; --parent) {
        adjust_heap(first, parent, len, id[first + parent]);
        if (parent == 0) break;
      }
    }
    while (last - first > 1) {
      --last;
      const int value = id[last];
      id[last] = id[first];
      adjust_heap(first, 0, last - first, value);
    }
  }
  LCB_HD void unguarded_linear_insert(int last) {
    const int val = id[last];
    int next = last - 1;
    while (lt(val, id[next])) {
      id[last] = id[next];
      last = next;
      --next;
    }
    id[last] = val;
  }
  LCB_HD void insertion_sort(int first, int last) {
    if (first == last) return;
    for (int i = first + 1; i != last; ++i) {
      if (lt(id[i], id[first])) {
        const int val = id[i];
        for (int j = i; j > first; --j) id[j] = id[j - 1];
        id[first] = val;
      } else {
        unguarded_linear_insert(i);
      }
    }
  }
  LCB_HD void sort(int n) {
    if (n <= 0) return;
    // __introsort_loop with an explicit stack of (first, last, depth) for the right-hand parts
    int stk_first[64], stk_last[64], stk_depth[64];
    int sp = 0;
    int lg = 0;
    for (int t = n; t > 1; t >>= 1) ++lg;
    stk_first[0] = 0;
    stk_last[0] = n;
    stk_depth[0] = 2 * lg;
    sp = 1;
    while (sp > 0) {
      --sp;
      const int first = stk_first[sp];
      int last = stk_last[sp], depth = stk_depth[sp];
      // the reference recursion handles [cut, last) before it continues with [first, cut): ranges are disjoint, so
      // the order in which they are processed does not change the result
      while (last - first > 16) {
        if (depth == 0) {
          heap_sort(first, last);
          break;
        }
        --depth;
        const int mid = first + (last - first) / 2;
        median_to_first(first, first + 1, mid, last - 1);
        const int cut = partition(first + 1, last, first);
        if (sp < 64) {
          stk_first[sp] = cut;
          stk_last[sp] = last;
          stk_depth[sp] = depth;
          ++sp;
        }
        last = cut;
      }
    }
    if (n > 16) {
      insertion_sort(0, 16);
      for (int i = 16; i != n; ++i) unguarded_linear_insert(i);
    } else {
      insertion_sort(0, n);
    }
  }
};

}  // namespace mm
}  // namespace lcb
