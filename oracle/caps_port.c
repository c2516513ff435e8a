/* TEST INFRASTRUCTURE ONLY — not product code.  See oracle_port.h for the pinning status.
 *
 * Plain-C restatement of the reference's CPU algorithm (samplesort over suffixes with an
 * LCP-carrying two-way merge).  It is written from the behaviour of the reference, one
 * function per reference routine, sequential except for OpenMP over the independent
 * subproblems (results do not depend on scheduling).  Indices are 64-bit throughout; the
 * reference instantiates uint32_t/uint64_t (src/Suffix_Array.cpp:543-544) with identical
 * results.
 */
#include "oracle_port.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef uint64_t ix;

typedef struct {
  const signed char* text; /* compared as signed char: src/Suffix_Array.cpp:77,289 */
  ix n;
  ix parts;   /* p_   : src/Suffix_Array.cpp:24 */
  ix context; /* max_context : src/Suffix_Array.cpp:25 */
  ix samples_per_part; /* pivot_per_part_ : src/Suffix_Array.cpp:27 */
} job_t;

/* Common-prefix length of a and b, at most `limit` symbols.  Scalar statement of
 * LCP<N>/LCP_unrolled<N> (include/Suffix_Array.hpp:195-241); the AVX2 blocking there
 * changes speed only. */
static ix common_prefix(const signed char* a, const signed char* b, ix limit) {
  ix k = 0;
  while (k + 8 <= limit) {
    uint64_t wa, wb;
    memcpy(&wa, a + k, 8);
    memcpy(&wb, b + k, 8);
    if (wa != wb) return k + (ix)(__builtin_ctzll(wa ^ wb) >> 3);
    k += 8;
  }
  while (k < limit && a[k] == b[k]) ++k;
  return k;
}

/* A sorted run of suffixes with its LCP array (lcp[t] = LCP(run[t-1], run[t])). */
typedef struct {
  const ix* suf;
  const ix* lcp;
  ix len, at;
} run_t;

/* LCP-aware two-way merge: reference merge(), src/Suffix_Array.cpp:48-109.
 * `carry` is the LCP between the head of the passive run and the last emitted suffix
 * (the reference's `m`).  The active run is the one that emitted last; the reference
 * swaps X and Y to keep that true (:85-92) — here two cursors and an index do the same. */
static void merge_runs(const job_t* J, const ix* xs, ix nx, const ix* xl, const ix* ys, ix ny,
                       const ix* yl, ix* out, ix* out_lcp) {
  run_t r[2] = {{xs, xl, nx, 0}, {ys, yl, ny, 0}};
  int act = 0; /* active run */
  ix carry = 0, k = 0;

  while (r[act].at < r[act].len && r[act ^ 1].at < r[act ^ 1].len) {
    run_t* A = &r[act];
    run_t* B = &r[act ^ 1];
    const ix sa = A->suf[A->at], sb = B->suf[B->at];
    const ix la = A->lcp[A->at];
    int take_active;

    if (la > carry) { /* :61-64 */
      take_active = 1;
      out_lcp[k] = la;
    } else if (la < carry) { /* :65-68 */
      take_active = 0;
      out_lcp[k] = carry;
      carry = la;
    } else { /* :69-80 — compare characters from offset `carry` on */
      const ix deeper = sa > sb ? sa : sb;
      const ix shorter_len = J->n - deeper;                                /* :71 */
      const ix ctx = J->context < shorter_len ? J->context : shorter_len; /* :72 */
      const ix full = carry + common_prefix(J->text + sa + carry, J->text + sb + carry,
                                            ctx - carry);                  /* :73 */
      ix winner;
      if (full == shorter_len)
        winner = deeper; /* the shorter suffix is a prefix of the other: it goes first (:76) */
      else
        winner = J->text[sa + full] < J->text[sb + full] ? sa : sb; /* :77 (ties -> passive) */
      take_active = (winner == sa);
      out_lcp[k] = take_active ? la : carry; /* :78 */
      carry = full;                          /* :79 */
    }

    if (take_active) {
      out[k] = sa;
      A->at++;
    } else {
      out[k] = sb;
      B->at++;
      act ^= 1; /* :85-92 */
    }
    ++k;
  }

  /* Tails (:98-108): exactly one run still has elements; its first copied LCP is `carry`. */
  for (int s = 0; s < 2; ++s) {
    const run_t* R = &r[s];
    if (R->at < R->len) {
      memcpy(out + k, R->suf + R->at, (R->len - R->at) * sizeof(ix));
      memcpy(out_lcp + k, R->lcp + R->at, (R->len - R->at) * sizeof(ix));
    }
  }
  if (k < nx + ny) out_lcp[k] = carry;
}

/* Top-down merge sort with ping-pong buffers: reference merge_sort(), :112-129.
 * Precondition (as there): dst == src element-wise.  Result in dst/dst_lcp. */
static void sort_run(const job_t* J, ix* src, ix* dst, ix count, ix* dst_lcp, ix* tmp_lcp) {
  if (count == 1) {
    dst_lcp[0] = 0;
    return;
  }
  const ix half = count / 2;
  sort_run(J, dst, src, half, tmp_lcp, dst_lcp);
  sort_run(J, dst + half, src + half, count - half, tmp_lcp + half, dst_lcp + half);
  merge_runs(J, src, half, tmp_lcp, src + half, count - half, tmp_lcp + half, dst, dst_lcp);
}

/* Regular sampling: reference sample_pivots(), :187-194. */
static void regular_sample(const ix* run, ix len, ix want, ix* out) {
  const ix gap = len / want;
  for (ix t = 0; t < want; ++t) out[t] = run[(t + 1) * gap - 1];
}

/* First position in a sorted run whose suffix is greater than the pattern (itself a suffix
 * of the text): reference upper_bound(), :252-297, including its 65536-symbol cap. */
static ix run_upper_bound(const job_t* J, const ix* run, ix len, ix pat_pos) {
  const signed char* pat = J->text + pat_pos;
  const ix pat_len = J->n - pat_pos;
  const ix cap = 65536; /* :261 */
  int64_t lo = -1, hi = (int64_t)len;
  ix answer = len, lcp_lo = 0, lcp_hi = 0;

  while (hi - lo > 1) {
    const ix mid = (ix)((lo + hi) / 2);
    const signed char* suf = J->text + run[mid];
    const ix suf_len = J->n - run[mid];
    ix known = lcp_lo < lcp_hi ? lcp_lo : lcp_hi; /* :269 */
    if (known > cap) known = cap;
    ix bound = suf_len < pat_len ? suf_len : pat_len; /* :271 */
    if (bound > J->context) bound = J->context;
    if (bound > cap) bound = cap;
    known += common_prefix(suf + known, pat + known, bound - known);

    if (known == bound) { /* :275-287 */
      if (known == pat_len) {
        if (pat_len == suf_len) return mid + 1; /* the pattern is this very suffix */
        hi = (int64_t)mid, lcp_hi = known, answer = mid;
      } else {
        lo = (int64_t)mid, lcp_lo = known;
      }
    } else if (suf[known] < pat[known]) { /* :289-292 */
      lo = (int64_t)mid, lcp_lo = known;
    } else {
      hi = (int64_t)mid, lcp_hi = known, answer = mid;
    }
  }
  return answer;
}

/* Balanced binary merge tree over `runs` sorted runs laid flat: reference sort_partition(),
 * :412-428.  ruler[0..runs] are the run offsets.  Precondition: dst == src (and LCPs). */
static void merge_tree(const job_t* J, ix* src, ix* dst, ix runs, const ix* ruler, ix* src_lcp,
                       ix* dst_lcp) {
  if (runs == 1) return;
  const ix half = runs / 2;
  const ix left = ruler[half] - ruler[0];
  const ix right = ruler[runs] - ruler[half];
  merge_tree(J, dst, src, half, ruler, dst_lcp, src_lcp);
  merge_tree(J, dst + left, src + left, runs - half, ruler + half, dst_lcp + left, src_lcp + left);
  merge_runs(J, src, left, src_lcp, src + left, right, src_lcp + left, dst, dst_lcp);
}

int caps_port_construct(const char* text, uint64_t n, uint64_t subproblems, uint64_t max_context,
                        uint64_t* sa, uint64_t* lcp) {
  if (n < 16) return -1;
  job_t J;
  J.text = (const signed char*)text;
  J.n = n;
  /* ctor, src/Suffix_Array.cpp:17-38 */
  J.parts = subproblems > 0 ? subproblems : 8192;
  if (J.parts > n / 16) J.parts = n / 16;
  J.context = max_context ? max_context : n;
  {
    const ix by_log = (ix)ceil(32.0 * log((double)n));
    const ix by_size = n / J.parts - 1;
    J.samples_per_part = by_log < by_size ? by_log : by_size;
  }
  const ix p = J.parts, slice = n / p, spp = J.samples_per_part;

  ix* sa_w = malloc(n * sizeof(ix));
  ix* lcp_w = malloc(n * sizeof(ix));
  ix* pivots = malloc(p * spp * sizeof(ix));
  ix* where = malloc(p * (p + 1) * sizeof(ix));  /* P        : :481 */
  ix* ruler = malloc(p * (p + 1) * sizeof(ix));  /* part_ruler_ */
  ix* part_at = malloc((p + 1) * sizeof(ix));    /* part_size_scan_ */
  if (!sa_w || !lcp_w || !pivots || !where || !ruler || !part_at) return -1;

  /* permute, :148-158 */
  for (ix i = 0; i < n; ++i) sa[i] = sa_w[i] = i;

  /* sort_subarrays, :161-184 — slice i covers text positions [i*slice, ...), last one
   * also takes the n % p remainder (:172). */
#pragma omp parallel for schedule(dynamic, 1)
  for (ix i = 0; i < p; ++i) {
    const ix len = slice + (i + 1 < p ? 0 : n % p);
    sort_run(&J, sa_w + i * slice, sa + i * slice, len, lcp + i * slice, lcp_w + i * slice);
  }

  /* select_pivots, :197-222 */
  {
    const ix total = p * spp;
    for (ix i = 0; i < p; ++i)
      regular_sample(sa + i * slice, slice + (i + 1 < p ? 0 : n % p), spp, pivots + i * spp);
    ix* sorted = malloc(total * sizeof(ix));
    ix* t1 = malloc(total * sizeof(ix));
    ix* t2 = malloc(total * sizeof(ix));
    if (!sorted || !t1 || !t2) return -1;
    memcpy(sorted, pivots, total * sizeof(ix));
    sort_run(&J, pivots, sorted, total, t1, t2);
    regular_sample(sorted, total, p - 1, pivots);
    free(sorted), free(t1), free(t2);
  }

  /* locate_pivots, :225-249 */
#pragma omp parallel for schedule(dynamic, 1)
  for (ix i = 0; i < p; ++i) {
    const ix len = slice + (i + 1 < p ? 0 : n % p);
    ix* row = where + i * (p + 1);
    row[0] = 0, row[p] = len;
    for (ix j = 0; j + 1 < p; ++j)
      row[j + 1] = run_upper_bound(&J, sa + i * slice, len, pivots[j]);
  }

  /* partition_sub_subarrays, :300-368 */
  {
    ix acc = 0;
    for (ix j = 0; j < p; ++j) {
      ix size = 0;
      for (ix i = 0; i < p; ++i) size += where[i * (p + 1) + j + 1] - where[i * (p + 1) + j];
      part_at[j] = acc;
      acc += size;
    }
    part_at[p] = acc;
    if (acc != n) return -1;
  }
#pragma omp parallel for schedule(dynamic, 1)
  for (ix j = 0; j < p; ++j) {
    ix* dst = sa_w + part_at[j];
    ix* dst_lcp = lcp_w + part_at[j];
    ix* marks = ruler + j * (p + 1);
    ix fill = 0;
    for (ix i = 0; i < p; ++i) {
      const ix from = where[i * (p + 1) + j], cnt = where[i * (p + 1) + j + 1] - from;
      marks[i] = fill;
      if (!cnt) continue;
      memcpy(dst + fill, sa + i * slice + from, cnt * sizeof(ix));
      memcpy(dst_lcp + fill, lcp + i * slice + from, cnt * sizeof(ix));
      dst_lcp[fill] = 0; /* :356 */
      fill += cnt;
    }
    marks[p] = fill;
  }

  /* merge_sub_subarrays, :371-409 */
  memcpy(sa, sa_w, n * sizeof(ix));
  memcpy(lcp, lcp_w, n * sizeof(ix));
#pragma omp parallel for schedule(dynamic, 1)
  for (ix j = 0; j < p; ++j)
    merge_tree(&J, sa_w + part_at[j], sa + part_at[j], p, ruler + j * (p + 1), lcp_w + part_at[j],
               lcp + part_at[j]);

  /* compute_partition_boundary_lcp, :431-447 (unbounded compare).  The reference writes
   * LCP_[part_at[j]] even when that is n (trailing empty partitions) — out of bounds
   * there; guarded here. */
  for (ix j = 1; j < p; ++j) {
    const ix at = part_at[j];
    if (at == 0 || at >= n) continue;
    const ix a = sa[at - 1], b = sa[at];
    lcp[at] = common_prefix(J.text + a, J.text + b, n - (a > b ? a : b));
  }

  free(sa_w), free(lcp_w), free(pivots), free(where), free(ruler), free(part_at);
  return 0;
}

int caps_port_construct_u32(const char* text, uint64_t n, uint64_t subproblems,
                            uint64_t max_context, uint32_t* sa_out, uint32_t* lcp_out) {
  if (n > 0xFFFFFFFFull) return -1;
  uint64_t* s = malloc(n * sizeof(uint64_t));
  uint64_t* l = malloc(n * sizeof(uint64_t));
  if (!s || !l) return -1;
  const int rc = caps_port_construct(text, n, subproblems, max_context, s, l);
  if (rc == 0)
    for (uint64_t i = 0; i < n; ++i) sa_out[i] = (uint32_t)s[i], lcp_out[i] = (uint32_t)l[i];
  free(s), free(l);
  return rc;
}

/* CLI byte mapping, reference src/main.cpp:61-70: every byte (headers and newlines too)
 * becomes "ACTG"[(toupper(c) & 6) >> 1]. */
void caps_port_map_acgt(char* text, uint64_t n) {
  static const char table[4] = {'A', 'C', 'T', 'G'};
  for (uint64_t i = 0; i < n; ++i) {
    int c = (unsigned char)text[i];
    if (c >= 'a' && c <= 'z') c -= 32; /* toupper in the C locale */
    text[i] = table[(c & 6) >> 1];
  }
}
