// TEST INFRASTRUCTURE ONLY — not product code.
//
// Stand-in for the three ParlayLib scheduling primitives the reference uses
// (ParlayLib itself is an un-vendored, un-pinned `git clone` in the reference's
// CMakeLists.txt:89-98 and cannot be fetched offline).  No arithmetic lives in
// ParlayLib; the reference calls only
//   parlay::par_do        (src/Suffix_Array.cpp:126,426)
//   parlay::parallel_for  (src/Suffix_Array.cpp:154,179,245,316,364,384,404,443)
//   parlay::blocked_for   (src/main.cpp:63)
// This header maps them onto OpenMP tasks so that the unmodified reference
// sources compile from where they lie under /root/reference (see oracle/Makefile).
// Thread count: PARLAY_NUM_THREADS (default: all cores), like ParlayLib.
#pragma once

#include <atomic>
#include <algorithm>
#include <cstddef>
#include <cstdlib>
#include <utility>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace parlay {

namespace shim_detail {

inline int thread_budget() {
  static const int t = [] {
    const char* e = std::getenv("PARLAY_NUM_THREADS");
    int v = e ? std::atoi(e) : 0;
#ifdef _OPENMP
    if (v <= 0) v = omp_get_num_procs();
#else
    v = 1;
#endif
    return v;
  }();
  return t;
}

template <class Body>
inline void run_forked(Body&& body) {
#ifdef _OPENMP
  if (omp_in_parallel()) {
    body();
  } else {
#pragma omp parallel num_threads(thread_budget())
#pragma omp single nowait
    body();
  }
#else
  body();
#endif
}

}  // namespace shim_detail

inline size_t num_workers() { return static_cast<size_t>(shim_detail::thread_budget()); }

template <class L, class R>
inline void par_do(L&& left, R&& right, bool /*conservative*/ = false) {
#ifdef _OPENMP
  shim_detail::run_forked([&] {
#pragma omp task default(shared)
    left();
    right();
#pragma omp taskwait
  });
#else
  left();
  right();
#endif
}

namespace shim_detail {

// Binary fork-join over [lo, hi) with explicit tasks (libgomp's taskloop ran sequentially
// under `single nowait` on this toolchain, so it is not used).
template <class F>
inline void split_range(size_t lo, size_t hi, size_t grain, F& f) {
  if (hi - lo <= grain) {
    for (size_t i = lo; i < hi; ++i) f(i);
    return;
  }
  const size_t mid = lo + (hi - lo) / 2;
#ifdef _OPENMP
#pragma omp task default(shared) firstprivate(lo, mid, grain)
  split_range(lo, mid, grain, f);
  split_range(mid, hi, grain, f);
#pragma omp taskwait
#else
  split_range(lo, mid, grain, f);
  split_range(mid, hi, grain, f);
#endif
}

}  // namespace shim_detail

template <class F>
inline void parallel_for(size_t start, size_t end, F&& f, long granularity = 0,
                         bool /*conservative*/ = false) {
  if (end <= start) return;
  const size_t count = end - start;
  const size_t grain = granularity > 0 ? static_cast<size_t>(granularity)
                                       : std::max<size_t>(1, count / (32 * num_workers()));
  shim_detail::run_forked([&] { shim_detail::split_range(start, end, grain, f); });
}

// f(block_index, block_start, block_end) with block_end clamped to `end`.
template <class F>
inline void blocked_for(size_t start, size_t end, size_t block_size, F&& f,
                        bool /*conservative*/ = false) {
  if (end <= start) return;
  const size_t blocks = (end - start + block_size - 1) / block_size;
  parallel_for(0, blocks, [&](size_t b) {
    const size_t s = start + b * block_size;
    f(b, s, std::min(end, s + block_size));
  });
}

}  // namespace parlay
