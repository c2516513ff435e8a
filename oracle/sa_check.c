/* TEST INFRASTRUCTURE ONLY — not product code.  See oracle_port.h.
 *
 * Algorithm-independent validator for (SA, LCP).  At the default (unbounded) context the
 * reference's output is the unique suffix array under signed-char order with the shorter
 * suffix first (src/Suffix_Array.cpp:71,76-77) and LCP[k] = lcp(SA[k-1], SA[k]), LCP[0] = 0
 * — the predicate of the reference's own (never called) is_sorted(), :512-536.  The order
 * is checked in O(n) through the inverse permutation instead of by character scans, so it
 * is usable on highly repetitive texts; LCP is recomputed with Kasai's algorithm.
 */
#include "oracle_port.h"

#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* Launchers such as torchrun export OMP_NUM_THREADS=1 to every rank; the multi-threaded checkers
 * below would then crawl on one core (3.1 G suffixes: ten minutes instead of 25 s).  The caller
 * says how many threads the checker may use. */
void caps_oracle_set_threads(int threads) {
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#else
  (void)threads;
#endif
}

static inline uint64_t get_idx(const void* arr, int w, uint64_t k) {
  return w == 4 ? (uint64_t)((const uint32_t*)arr)[k] : ((const uint64_t*)arr)[k];
}

int caps_check_sa_lcp(const char* text, uint64_t n, const void* sa, const void* lcp,
                      int idx_bytes, uint64_t* bad_pos) {
  if ((idx_bytes != 4 && idx_bytes != 8) || !text || !sa || !lcp) return 4;
  if (n == 0) return 0;
  const signed char* t = (const signed char*)text;
  uint64_t* rank = malloc((n + 1) * sizeof(uint64_t)); /* rank+1, 0 = unseen / empty suffix */
  if (!rank) return 4;
  memset(rank, 0, (n + 1) * sizeof(uint64_t));
  int rc = 0;
  uint64_t bad = 0;

  for (uint64_t k = 0; k < n; ++k) { /* permutation of [0, n) */
    const uint64_t s = get_idx(sa, idx_bytes, k);
    if (s >= n || rank[s] != 0) {
      rc = 1, bad = k;
      goto done;
    }
    rank[s] = k + 1;
  }
  rank[n] = 0; /* the empty suffix precedes everything */

  for (uint64_t k = 1; k < n; ++k) { /* suffix a must precede suffix b */
    const uint64_t a = get_idx(sa, idx_bytes, k - 1), b = get_idx(sa, idx_bytes, k);
    if (t[a] > t[b] || (t[a] == t[b] && rank[a + 1] >= rank[b + 1])) {
      rc = 2, bad = k;
      goto done;
    }
  }

  if (get_idx(lcp, idx_bytes, 0) != 0) {
    rc = 3, bad = 0;
    goto done;
  }
  { /* Kasai: walk the text, h drops by at most one per step */
    uint64_t h = 0;
    for (uint64_t i = 0; i < n; ++i) {
      const uint64_t k = rank[i] - 1;
      if (k == 0) {
        h = 0;
        continue;
      }
      const uint64_t j = get_idx(sa, idx_bytes, k - 1);
      while (i + h < n && j + h < n && t[i + h] == t[j + h]) ++h;
      if (get_idx(lcp, idx_bytes, k) != h) {
        rc = 3;
        if (bad == 0 || k < bad) bad = k;
      }
      if (rc) break;
      if (h) --h;
    }
  }

done:
  if (rc && bad_pos) *bad_pos = bad;
  free(rank);
  return rc;
}

/* Validator for the 1 Gbp periodic text of BASELINE config 4(i), where the sequential Kasai
 * walk above is too slow to run on a GPU box's host: text = its first `period` bytes repeated.
 * Same predicates — permutation, suffix order through the inverse permutation — but OpenMP
 * loops, and the LCP from the closed form for periodic texts: with L = n - max(a, b),
 *   lcp(a, b) = L                           if a = b (mod period)
 *             = min(L, mis[a % P][b % P])   otherwise, mis = first offset at which the two
 *                                           rotations of the unit differ (computed by direct
 *                                           comparison of the rotations).
 * Returns the codes of caps_check_sa_lcp, 5 if the text is not period-periodic. */
int caps_check_sa_lcp_periodic(const char* text, uint64_t n, uint64_t period, const void* sa, const void* lcp,
                               int idx_bytes, uint64_t* bad_pos) {
  if ((idx_bytes != 4 && idx_bytes != 8) || !text || !sa || !lcp || period == 0 || period > 4096) return 4;
  if (n == 0) return 0;
  const signed char* t = (const signed char*)text;
  const uint64_t P = period;
  int rc = 0;
  uint64_t bad = UINT64_MAX;
#define CAPS_FAIL(code, where)                        \
  do {                                                \
    _Pragma("omp critical(caps_check_fail)") {        \
      if ((where) < bad) rc = (code), bad = (where);  \
    }                                                 \
  } while (0)

  int periodic = 1;
#pragma omp parallel for schedule(static) reduction(&& : periodic)
  for (uint64_t i = P; i < n; ++i) periodic = periodic && t[i] == t[i - P];
  if (!periodic) return 5;

  uint32_t* mis = malloc(P * P * sizeof(uint32_t));
  uint64_t* rank = malloc((n + 1) * sizeof(uint64_t));
  unsigned char* seen = calloc(n, 1);
  if (!mis || !rank || !seen) {
    free(mis), free(rank), free(seen);
    return 4;
  }
#pragma omp parallel for schedule(dynamic, 8)
  for (uint64_t r = 0; r < P; ++r)
    for (uint64_t q = 0; q < P; ++q) {
      uint64_t m = 0;
      while (m < P && t[(r + m) % P] == t[(q + m) % P]) ++m;
      mis[r * P + q] = m < P ? (uint32_t)m : UINT32_MAX; /* equal rotations: unit not primitive */
    }

#pragma omp parallel for schedule(static)
  for (uint64_t k = 0; k < n; ++k) { /* permutation of [0, n) */
    const uint64_t s = get_idx(sa, idx_bytes, k);
    if (s >= n || __atomic_exchange_n(&seen[s], 1, __ATOMIC_RELAXED)) {
      CAPS_FAIL(1, k);
    } else {
      rank[s] = k + 1;
    }
  }
  rank[n] = 0; /* the empty suffix precedes everything */
  if (rc) goto done;

#pragma omp parallel for schedule(static)
  for (uint64_t k = 1; k < n; ++k) {
    const uint64_t a = get_idx(sa, idx_bytes, k - 1), b = get_idx(sa, idx_bytes, k);
    if (t[a] > t[b] || (t[a] == t[b] && rank[a + 1] >= rank[b + 1])) CAPS_FAIL(2, k);
    const uint64_t L = n - (a > b ? a : b);
    const uint64_t m = a % P == b % P ? UINT64_MAX : mis[(a % P) * P + b % P];
    const uint64_t want = m == UINT32_MAX || m > L ? L : m;
    if (get_idx(lcp, idx_bytes, k) != want) CAPS_FAIL(3, k);
  }
  if (get_idx(lcp, idx_bytes, 0) != 0) CAPS_FAIL(3, 0);

done:
#undef CAPS_FAIL
  if (rc && bad_pos) *bad_pos = bad;
  free(mis), free(rank), free(seen);
  return rc;
}

/* Multi-threaded form of caps_check_sa_lcp for the sizes bench.py builds (3.1 G suffixes): the
 * same three predicates — permutation, suffix order through the inverse permutation, LCP by
 * Kasai's walk — as OpenMP loops.  Kasai's walk is sequential in the text position (h drops by
 * at most one per step); here the text is cut into `pieces` ranges of positions and every range
 * starts its own walk at h = 0, which costs one from-scratch comparison per range and changes
 * nothing else.  The inverse permutation is stored at the index width of the input (4 bytes for
 * a 32-bit suffix array), so 3.1 G suffixes need 12.4 GB on top of the arrays being checked.
 * Same return codes; bad_pos = the smallest offending SA position of the failing predicate. */
int caps_check_sa_lcp_mt_pieces(const char* text, uint64_t n, const void* sa, const void* lcp, int idx_bytes,
                                uint64_t max_pieces, uint64_t* bad_pos);

int caps_check_sa_lcp_mt(const char* text, uint64_t n, const void* sa, const void* lcp, int idx_bytes,
                         uint64_t* bad_pos) {
  return caps_check_sa_lcp_mt_pieces(text, n, sa, lcp, idx_bytes, 0, bad_pos);
}

/* max_pieces = the number of ranges Kasai's walk is cut into (0: up to 4096).  Every range pays one
 * comparison from scratch, i.e. up to the longest LCP: on highly repetitive texts (Fibonacci words)
 * use about one range per thread. */
int caps_check_sa_lcp_mt_pieces(const char* text, uint64_t n, const void* sa, const void* lcp, int idx_bytes,
                                uint64_t max_pieces, uint64_t* bad_pos) {
  if ((idx_bytes != 4 && idx_bytes != 8) || !text || !sa || !lcp) return 4;
  if (n == 0) return 0;
  const signed char* t = (const signed char*)text;
  const int w = idx_bytes;
  void* rank = malloc(n * (size_t)w); /* rank[s] = SA position of suffix s */
  unsigned char* seen = calloc((n + 7) / 8, 1);
  if (!rank || !seen) {
    free(rank), free(seen);
    return 4;
  }
  int rc = 0;
  uint64_t bad = UINT64_MAX;
#define CAPS_FAIL(code, where)                       \
  do {                                               \
    _Pragma("omp critical(caps_check_mt_fail)") {    \
      if ((where) < bad) rc = (code), bad = (where); \
    }                                                \
  } while (0)
#define RANK_OF(s) (w == 4 ? (uint64_t)((const uint32_t*)rank)[s] : ((const uint64_t*)rank)[s])

#pragma omp parallel for schedule(static)
  for (uint64_t k = 0; k < n; ++k) { /* permutation of [0, n) */
    const uint64_t s = get_idx(sa, w, k);
    if (s >= n) {
      CAPS_FAIL(1, k);
      continue;
    }
    const unsigned char bit = (unsigned char)(1u << (s & 7u));
    if (__atomic_fetch_or(&seen[s >> 3], bit, __ATOMIC_RELAXED) & bit) {
      CAPS_FAIL(1, k);
      continue;
    }
    if (w == 4)
      ((uint32_t*)rank)[s] = (uint32_t)k;
    else
      ((uint64_t*)rank)[s] = k;
  }
  if (rc) goto done;

#pragma omp parallel for schedule(static)
  for (uint64_t k = 1; k < n; ++k) { /* suffix a must precede suffix b; the empty suffix precedes everything */
    const uint64_t a = get_idx(sa, w, k - 1), b = get_idx(sa, w, k);
    int ok;
    if (t[a] != t[b]) {
      ok = t[a] < t[b];
    } else if (a + 1 == n) {
      ok = 1;
    } else if (b + 1 == n) {
      ok = 0;
    } else {
      ok = RANK_OF(a + 1) < RANK_OF(b + 1);
    }
    if (!ok) CAPS_FAIL(2, k);
  }
  if (rc) goto done;

  if (get_idx(lcp, w, 0) != 0) CAPS_FAIL(3, 0);
  {
    uint64_t pieces = n / 65536 + 1;
    if (pieces > 4096) pieces = 4096;
    if (max_pieces && pieces > max_pieces) pieces = max_pieces;
#pragma omp parallel for schedule(dynamic, 1)
    for (uint64_t p = 0; p < pieces; ++p) {
      const uint64_t lo = n / pieces * p, hi = p + 1 == pieces ? n : n / pieces * (p + 1);
      uint64_t h = 0;
      for (uint64_t i = lo; i < hi; ++i) {
        const uint64_t k = RANK_OF(i);
        if (k == 0) {
          h = 0;
          continue;
        }
        const uint64_t j = get_idx(sa, w, k - 1);
        while (i + h < n && j + h < n && t[i + h] == t[j + h]) ++h;
        if (get_idx(lcp, w, k) != h) CAPS_FAIL(3, k);
        if (h) --h;
      }
    }
  }

done:
#undef CAPS_FAIL
#undef RANK_OF
  if (rc && bad_pos) *bad_pos = bad;
  free(rank), free(seen);
  return rc;
}

typedef struct {
  const signed char* t;
  uint64_t n;
} naive_env;
static naive_env g_env; /* qsort has no context argument; tiny inputs, single-threaded use */

static int naive_cmp(const void* pa, const void* pb) {
  const uint64_t a = *(const uint64_t*)pa, b = *(const uint64_t*)pb;
  const uint64_t la = g_env.n - a, lb = g_env.n - b, m = la < lb ? la : lb;
  for (uint64_t k = 0; k < m; ++k)
    if (g_env.t[a + k] != g_env.t[b + k]) return g_env.t[a + k] < g_env.t[b + k] ? -1 : 1;
  return la < lb ? -1 : (la > lb ? 1 : 0);
}

int caps_naive_sa_lcp(const char* text, uint64_t n, uint64_t* sa_out, uint64_t* lcp_out) {
  g_env.t = (const signed char*)text;
  g_env.n = n;
  for (uint64_t i = 0; i < n; ++i) sa_out[i] = i;
  qsort(sa_out, n, sizeof(uint64_t), naive_cmp);
  for (uint64_t k = 0; k < n; ++k) {
    uint64_t h = 0;
    if (k) {
      const uint64_t a = sa_out[k - 1], b = sa_out[k];
      while (a + h < n && b + h < n && g_env.t[a + h] == g_env.t[b + h]) ++h;
    }
    lcp_out[k] = h;
  }
  return 0;
}
