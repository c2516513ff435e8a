// Building blocks shared by the single-GPU construction (sa_build.cu) and the sharded one
// (sharded_build.cu):
//   sort_suffixes_by_key / sort_suffix_slice
//                       key sort of suffixes on their packed-prefix key: the packed-record MSD sort
//                       of msd_sort.cuh (msd_partition + msd_local_sort) for 32-bit indices, stable
//                       LSD passes otherwise (replaces permute + sort_subarrays/merge_sort,
//                       reference src/Suffix_Array.cpp:112-184);
//   refine_shallow / refine_deep / fix_group_edges (refine_tied_groups = the three in sequence)
//                       refinement restricted to the suffixes that are still tied — pair chains,
//                       text rounds, then prefix doubling on ranks (replaces the character
//                       compares inside merge, :69-80);
//                       where the ranks live is a policy: one array (LocalRanks) or sharded by
//                       text position with an exchange per round (ShardedRanks);
//   key_lcp_count_kernel, tied_collect_kernel
//                       first pass over the sorted keys: LCP of neighbours with different keys
//                       (clz(key_a ^ key_b)), bitmaps of the tied positions and of the starts of
//                       their key groups; the lists of the tied positions from the bitmaps;
//   plcp_for_pairs      LCP of tied neighbours by the permuted-LCP recurrence
//                       PLCP[i] = PLCP[i-1] - 1 on reducible positions and a packed-word
//                       comparison on the irreducible ones (replaces the LCPs carried
//                       through merge, :61-68,:78).
#pragma once

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <type_traits>

#include "engine.cuh"

namespace capsb {

template <class IdxT>
struct IdxTraits;
template <>
struct IdxTraits<uint32_t> {
  using Comp = uint64_t;  // (group head + in-range bit) << 32 | second rank
  static constexpr unsigned kField = 32;
};
template <>
struct IdxTraits<uint64_t> {
  using Comp = unsigned __int128;
  static constexpr unsigned kField = 64;
};

template <class IdxT>
struct IdxPair {
  IdxT a, b;
};

template <class IdxT>
void chain_pairs_local(Engine& eng, const PackedText& pt, IdxT* hi, const IdxT* lo, uint64_t m, uint64_t known,
                       IdxPair<IdxT>* answer);
template <class IdxT, class PosJ, class Out>
void plcp_for_pairs(Engine& eng, const PackedText& pt, const IdxT* pos_i, PosJ pos_j, uint64_t m, uint64_t known,
                    Out out);

inline unsigned bit_length(uint64_t v) {
  unsigned b = 0;
  while (v) ++b, v >>= 1;
  return b ? b : 1;
}
inline unsigned round_up8(unsigned b) { return (b + 7u) & ~7u; }

// Width of the sort key.  Every 8 bits cost one radix pass over all n suffixes (24 B/suffix of
// HBM traffic with 32-bit indices), while a suffix left tied costs a few hundred bytes in the
// refinement; log2(n) + 8 bits leave about n/512 accidental ties on random text, so the key
// stops there instead of always spending all 64 bits.  A multiple of 8 is a whole number of
// symbols for every code width (1, 2, 4, 8 bits).
inline unsigned choose_key_bits(uint64_t n) {
  unsigned bits = round_up8(bit_length(n) + 8);
  if (const char* env = std::getenv("CAPSB_KEY_BITS")) bits = round_up8(static_cast<unsigned>(std::atoi(env)));
  return bits < 16 ? 16 : (bits > 64 ? 64 : bits);
}
inline uint64_t key_mask_of(unsigned key_bits) { return key_bits >= 64 ? ~0ull : ~0ull << (64 - key_bits); }

// First radix pass reads keys from the packed text: key = window at suffix base + i (leading
// key bits only), value = base + i.
template <class IdxT>
struct TextSource {
  PackedText pt;
  uint64_t mask;
  uint64_t base;
  __device__ __forceinline__ uint64_t key(uint64_t i) const { return pt.window(base + i) & mask; }
  __device__ __forceinline__ IdxT val(uint64_t i) const { return static_cast<IdxT>(base + i); }
  // the text window is re-read from L1/L2; compulsory traffic is the packed text itself (< 1 B)
  static constexpr uint64_t bytes_read_per_item() { return 1; }
  // suffix i + 1 follows suffix i in the text: the MSD sort's level A loads the packed words once per thread
  static constexpr bool kSequentialText = true;
};

// CAPSB_TRACE=1: per-stage / per-round log on stderr (host clock, pool occupancy)
inline bool trace_enabled() {
  static const bool on = std::getenv("CAPSB_TRACE") != nullptr;
  return on;
}
inline void trace_point(Engine& eng, const char* what) {
  if (!trace_enabled()) return;
  static const auto t0 = std::chrono::steady_clock::now();
  cudaStreamSynchronize(eng.stream);
  const uint64_t reserved = eng.arena.reserved(), used = eng.arena.used();
  const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  static thread_local double last_ms = 0;
  // (no cudaMemGetInfo here: it takes milliseconds and lands in whichever interval follows)
  std::fprintf(stderr, "[capsb dev%d] %10.3f ms (+%8.3f)  %-40s arena reserved %.2f GB used %.2f GB\n", eng.dev.device, ms,
               ms - last_ms, what, reserved / 1e9, used / 1e9);
  last_ms = ms;
}

// Stage timer: records an event now; elapsed times are read at the end.
struct StageClock {
  Engine& eng;
  std::vector<cudaEvent_t> marks;
  explicit StageClock(Engine& e) : eng(e) {}
  void mark(const char* what = "stage") {
    trace_point(eng, what);
    cudaEvent_t ev;
    if (!eng.events.empty()) {
      ev = eng.events.back();
      eng.events.pop_back();
    } else {
      CAPSB_CUDA(cudaEventCreate(&ev));
    }
    CAPSB_CUDA(cudaEventRecord(ev, eng.stream));
    marks.push_back(ev);
  }
  float between(size_t a, size_t b) {
    float ms = 0;
    CAPSB_CUDA(cudaEventElapsedTime(&ms, marks[a], marks[b]));
    return ms;
  }
  ~StageClock() {
    for (cudaEvent_t e : marks) eng.events.push_back(e);
  }
};

// Scan in two steps so the total (a count) is known before the outputs are allocated.
template <class T, class Op, class In>
T scan_total(Engine& eng, uint64_t n, In in) {
  ScanScratch<T>& sc = eng.scan_scratch<T>();
  if (n == 0) return Op::template identity<T>();
  const Chunking ck = make_chunking(n, kScanTile, sc.max_blocks);
  CAPSB_LAUNCH((scan_reduce_kernel<T, Op, In>), ck.blocks, kScanThreads, 0, eng.stream, n, ck.chunk, in,
               sc.partial.get());
  CAPSB_LAUNCH((scan_spine_kernel<T, Op>), 1, kScanThreads, 0, eng.stream, ck.blocks, sc.partial.get(),
               sc.total.get());
  T total;
  read_back(eng.stream, &total, sc.total.get(), sizeof(T));
  return total;
}
// Must directly follow scan_total / another scan of the same n (reuses the partials).
template <class T, class Op, bool Inclusive, class In, class Out>
void scan_finish(Engine& eng, uint64_t n, In in, Out out) {
  ScanScratch<T>& sc = eng.scan_scratch<T>();
  if (n == 0) return;
  const Chunking ck = make_chunking(n, kScanTile, sc.max_blocks);
  CAPSB_LAUNCH((scan_apply_kernel<T, Op, Inclusive, In, Out>), ck.blocks, kScanThreads, 0, eng.stream, n,
               ck.chunk, in, out, sc.partial.get());
}
// Must directly follow scan_total<T, OpSum> over 0/1 flags: out(i, slot) for the flagged i only.
// other(i) is called for the rest.
template <class T, class Flag, class Out, class Other = NoOther>
void select_finish(Engine& eng, uint64_t n, Flag flag, Out out, Other other = Other()) {
  ScanScratch<T>& sc = eng.scan_scratch<T>();
  if (n == 0) return;
  const Chunking ck = make_chunking(n, kScanTile, sc.max_blocks);
  CAPSB_LAUNCH((select_apply_kernel<T, Flag, Out, Other>), ck.blocks, kScanThreads, 0, eng.stream, n, ck.chunk, flag,
               out, other, sc.partial.get());
}
template <class T, class Op, bool Inclusive, class In, class Out>
void scan_full(Engine& eng, uint64_t n, In in, Out out) {
  device_scan<T, Op, Inclusive>(eng.dev, eng.stream, eng.scan_scratch<T>(), n, in, out);
}

// ---------------------------------------------------------------------------------------
// Key sort of the suffixes [base, base + count): on return keys_out / sa_out hold them in key
// order (keys masked to key_bits).  keys_out and sa_out are caller-owned, `count` entries.
// ---------------------------------------------------------------------------------------
// CAPSB_SORT=lsd keeps the key sort on the LSD passes of radix_sort.cuh (read per construction:
// the tests flip it); default is the packed-record MSD sort where it applies.
inline bool msd_sort_enabled() {
  const char* env = std::getenv("CAPSB_SORT");
  return !(env && std::string(env) == "lsd");
}

// The packed-record MSD sort (msd_sort.cuh) of `count` suffixes given by `first` (key(i) = masked
// text window of the i-th suffix, val(i) = the suffix), in two parts: msd_partition runs levels A
// and B over everything (keys_out doubles as the level-B record buffer); msd_local_sort finishes
// the buckets [q_begin, q_end) — sorted keys in place of the records, suffixes to sa_out — so a
// caller can finish the suffix array a range of positions at a time.
struct MsdSorted {
  unsigned a = 0, b = 0, key_bits = 0;
  uint64_t count = 0;
  uint32_t nbuckets = 0;
  DevBuf<uint32_t> start_b;        // nbuckets + 1 bucket starts
  std::vector<uint32_t> a_starts;  // host copy of the 2^a + 1 level-A bucket starts (on request)
  DevBuf<uint32_t> large_list, large_count;  // two lists (oversized for the first / the second local kernel), two counts
  uint64_t large_capacity = 0;
};

template <class FirstSrc>
void msd_partition(Engine& eng, FirstSrc first, uint64_t count, unsigned key_bits, uint64_t* keys_out, MsdSorted& ms,
                   bool want_a_starts) {
  cudaStream_t st = eng.stream;
  const DeviceInfo& dev = eng.dev;
  const MsdPlan plan = msd_plan(count, key_bits);
  const unsigned a = plan.a, b = plan.b;
  const unsigned rem_a = key_bits - a;  // <= 32: the record's key field after level A
  ms.a = a, ms.b = b, ms.key_bits = key_bits, ms.count = count;
  ms.nbuckets = 1u << (a + b);
  eng.stats.msd_a_bits = a;
  eng.stats.msd_b_bits = b;
  DevBuf<uint32_t> root(2, st), start_a((1u << a) + 1, st);
  ms.start_b.alloc(static_cast<uint64_t>(ms.nbuckets) + 1, st);
  {
    uint32_t* r = root.get();
    const uint32_t c = static_cast<uint32_t>(count);
    launch_map(dev, st, 1, [=] __device__(uint64_t) { r[0] = 0, r[1] = c; });
  }
  {
    DevBuf<uint64_t> rec_a(count, st);
    // level A: digit = top a bits of the key, records = (remaining bits << 32) | suffix
    // CAPSB_MSD_SHORT: which level-A passes may shift all of a thread's windows out of the first one
    // (bit 0 the counting pass, bit 1 the scatter pass; default both)
    unsigned short_windows = 3;
    if (const char* env = std::getenv("CAPSB_MSD_SHORT")) short_windows = static_cast<unsigned>(std::atoi(env)) & 3u;
    msd_partition_level<MsdFirstSource<FirstSrc>, false>(
        dev, st, eng.msd_timers, eng.msd_timers.scatter_a,
        MsdFirstSource<FirstSrc>{first, 64u - key_bits, rem_a, short_windows},
        count, root.get(), 1, a, 2, 4, FirstSrc::bytes_read_per_item(), start_a.get(), rec_a.get());
    // level B: every level-A bucket by the next b bits
    msd_partition_level<MsdRecordSource, true>(dev, st, eng.msd_timers, eng.msd_timers.scatter_b,
                                               MsdRecordSource{rec_a.get(), 32u + rem_a - b, (1u << b) - 1u}, count,
                                               start_a.get(), 1u << a, b, 8, 1, sizeof(uint64_t), ms.start_b.get(),
                                               keys_out);
  }
  if (want_a_starts) {
    ms.a_starts.resize((1u << a) + 1);
    read_back(st, ms.a_starts.data(), start_a.get(), ms.a_starts.size() * sizeof(uint32_t));
  }
  ms.large_capacity = count / kMsdLocalCap + 1;
  ms.large_list.alloc(2 * ms.large_capacity, st);
  ms.large_count.alloc(2, st);
}

inline void msd_local_sort(Engine& eng, MsdSorted& ms, uint64_t* keys_out, uint32_t* sa_out, uint32_t q_begin,
                           uint32_t q_end, uint64_t records) {
  cudaStream_t st = eng.stream;
  const DeviceInfo& dev = eng.dev;
  if (q_begin >= q_end) return;
  uint32_t* counts = ms.large_count.get();  // [0] buckets too large for the first kernel, [1] for the second
  CAPSB_CUDA(cudaMemsetAsync(counts, 0, 2 * sizeof(uint32_t), st));
  uint32_t* huge_list = ms.large_list.get() + ms.large_capacity;
  // CAPSB_MSD_LOCAL_EXTRA=0: one counter per two records in the first local counting pass (default 1: one per record)
  unsigned extra_bits = 1;
  if (const char* env = std::getenv("CAPSB_MSD_LOCAL_EXTRA")) extra_bits = static_cast<unsigned>(std::atoi(env)) & 3u;
  // Groups of records that share the first pass's digit are ordered by comparison up to this size, by a
  // second counting pass above it.  A pass costs a bucket a fixed few microseconds (barriers, the scan
  // of the counters) whatever the group's size, the comparison loop size^2 shared-memory reads spread
  // over the CTA: the pass only pays for groups of hundreds.  3.1 Gbp, local sort per construction:
  // 40.2 ms with 32 / 32, 36.9 with 32 / 256, 36.0 with 128 / 256, 36.9 with 256 / 512 (profiles/r02 s17).
  // (CAPSB_MSD_SMALL_GROUP[_BIG]: tuning.)
  auto group_limit = [](const char* name, unsigned dflt) {
    const char* env = std::getenv(name);
    const long v = env ? std::atol(env) : static_cast<long>(dflt);
    return static_cast<unsigned>(v < static_cast<long>(kMsdSmallGroup) ? kMsdSmallGroup : v > 2048 ? 2048 : v);
  };
  // CAPSB_MSD_L2_PREFETCH=0: no prefetch hint for the bucket a CTA takes next
  const bool l2_prefetch = !(std::getenv("CAPSB_MSD_L2_PREFETCH") && std::getenv("CAPSB_MSD_L2_PREFETCH")[0] == '0');
  const unsigned small_group = group_limit("CAPSB_MSD_SMALL_GROUP", 128);
  const unsigned small_group_big = group_limit("CAPSB_MSD_SMALL_GROUP_BIG", 256);
  {
    using Small = MsdLocalSmem<kMsdLocalCap>;
    msd_allow_smem(msd_local_kernel<kMsdThreads, CAPSB_MSD_MIN_CTAS>, sizeof(Small));
    const uint32_t grid = std::min<uint32_t>(q_end - q_begin, static_cast<uint32_t>(dev.sm_count) * 2u);
    MsdTimed timed(eng.msd_timers.local, st, records * (2 * sizeof(uint64_t) + sizeof(uint32_t)));
    CAPSB_LAUNCH((msd_local_kernel<kMsdThreads, CAPSB_MSD_MIN_CTAS>), grid, kMsdThreads, sizeof(Small), st, keys_out,
                 ms.start_b.get(), static_cast<const uint32_t*>(nullptr), q_begin, q_end, ms.key_bits, ms.a + ms.b,
                 extra_bits, small_group, l2_prefetch, sa_out, ms.large_list.get(), counts);
  }
  uint32_t nlarge = 0;
  read_back(st, &nlarge, counts, sizeof(uint32_t));
  if (nlarge == 0) return;
  {  // the listed buckets with twice the threads and staging area, one CTA per SM
    using Big = MsdLocalSmem<kMsdBigThreads * kMsdItems>;
    msd_allow_smem(msd_local_kernel<kMsdBigThreads, 1>, sizeof(Big));
    const uint32_t grid = std::min<uint32_t>(nlarge, static_cast<uint32_t>(dev.sm_count));
    MsdTimed timed(eng.msd_timers.local, st, 0);  // (its records are counted with the first launch)
    CAPSB_LAUNCH((msd_local_kernel<kMsdBigThreads, 1>), grid, kMsdBigThreads, sizeof(Big), st, keys_out,
                 ms.start_b.get(), static_cast<const uint32_t*>(ms.large_list.get()), 0u, nlarge, ms.key_bits,
                 ms.a + ms.b, extra_bits, small_group_big, l2_prefetch, sa_out, huge_list, counts + 1);
  }
  uint32_t nhuge = 0;
  read_back(st, &nhuge, counts + 1, sizeof(uint32_t));
  eng.stats.msd_large_buckets += nhuge;
  if (nhuge > 0) {
    uint64_t large_records = 0;
    msd_sort_large_buckets(dev, st, eng.radix, keys_out, sa_out, ms.start_b.get(), huge_list, nhuge, ms.key_bits,
                           ms.a + ms.b, eng.scan32, &large_records);
    eng.stats.msd_large_records += large_records;
  }
}

template <class FirstSrc>
void msd_sort_suffixes(Engine& eng, FirstSrc first, uint64_t count, unsigned key_bits, uint64_t* keys_out,
                       uint32_t* sa_out) {
  MsdSorted ms;
  msd_partition(eng, first, count, key_bits, keys_out, ms, false);
  msd_local_sort(eng, ms, keys_out, sa_out, 0, ms.nbuckets, count);
}

// The LSD sort: stable 8-bit passes over (u64 key, suffix) pairs, the first one reading the keys
// from `first` (64-bit indices, keys too wide for the packed records, CAPSB_SORT=lsd).
template <class IdxT, class FirstSrc>
void lsd_sort_suffixes(Engine& eng, FirstSrc first, uint64_t count, unsigned key_bits, uint64_t* keys_out, IdxT* sa_out) {
  cudaStream_t st = eng.stream;
  const unsigned passes = key_bits / 8;
  DevBuf<uint64_t> key_tmp(count, st);
  DevBuf<IdxT> val_tmp(count, st);
  // pass q writes buffer pair q & 1; the last pass must leave its output in the caller's arrays
  const unsigned last = (passes - 1) & 1u;
  uint64_t* key_buf[2];
  IdxT* val_buf[2];
  key_buf[last] = keys_out, val_buf[last] = sa_out;
  key_buf[last ^ 1u] = key_tmp.get(), val_buf[last ^ 1u] = val_tmp.get();
  radix_pass<uint64_t, IdxT>(st, eng.radix, first, count, 64 - key_bits, key_buf[0], val_buf[0]);
  for (unsigned q = 1; q < passes; ++q) {
    const unsigned in = (q - 1) & 1u, out = q & 1u;
    radix_pass<uint64_t, IdxT>(st, eng.radix, ArraySource<uint64_t, IdxT>{key_buf[in], val_buf[in]}, count,
                               64 - key_bits + 8 * q, key_buf[out], val_buf[out]);
  }
}

template <class IdxT, class FirstSrc>
void sort_suffixes_by_key(Engine& eng, FirstSrc first, uint64_t count, unsigned key_bits, uint64_t* keys_out,
                          IdxT* sa_out) {
  if (count == 0) return;
  if constexpr (sizeof(IdxT) == 4) {
    if (msd_sort_enabled() && msd_applicable(key_bits) && count <= 0xFFFFFFFFull) {
      msd_sort_suffixes(eng, first, count, key_bits, keys_out, sa_out);
      return;
    }
  }
  lsd_sort_suffixes<IdxT>(eng, first, count, key_bits, keys_out, sa_out);
}

// The suffixes [base, base + count) of the text.
template <class IdxT>
void sort_suffix_slice(Engine& eng, const PackedText& pt, uint64_t base, uint64_t count, unsigned key_bits,
                       uint64_t* keys_out, IdxT* sa_out) {
  sort_suffixes_by_key<IdxT>(eng, TextSource<IdxT>{pt, key_mask_of(key_bits), base}, count, key_bits, keys_out, sa_out);
}

// First radix pass over a list of suffixes: key = window at suffix idx[i].  The lists this is
// used on are runs of increasing text positions, so the window reads stream through the
// packed text.
template <class IdxT>
struct SuffixListSource {
  PackedText pt;
  uint64_t mask;
  const IdxT* idx;
  __device__ __forceinline__ uint64_t key(uint64_t i) const { return pt.window(idx[i]) & mask; }
  __device__ __forceinline__ IdxT val(uint64_t i) const { return idx[i]; }
  static constexpr uint64_t bytes_read_per_item() { return 2 * sizeof(IdxT) + 1; }
};

// Rank (= final position in the sorted bucket keys[0, count) / sa[0, count)) of a suffix j that holds
// no published rank.  Only the suffixes that are still tied when the prefix doubling starts publish
// ranks; every other suffix is final, and its position can be found:
//   * its key is unique: the position of the key (binary search);
//   * the text rounds separated it from every other suffix within its first `depth` symbols: inside
//     its key group — sorted by now, except for the still-tied subgroups, whose members all compare
//     the same way against j — a binary search that compares text from the end of the key on;
//   * the pair-chain step ordered it against the one suffix it was tied with: the two are
//     neighbours, so the search ends next to it.
// (Publishing those ranks instead is a random 4-byte scatter over the whole text: 15 of the 100 ms the
// refinement took at 3.1 Gbp, and an all-to-all in the sharded path.)
template <class IdxT>
__device__ __forceinline__ uint64_t rank_of_unpublished(const PackedText& pt, const uint64_t* __restrict__ keys,
                                                        const IdxT* __restrict__ sa, uint64_t count, uint64_t want,
                                                        uint64_t j, uint64_t key_symbols, uint64_t depth) {
  uint64_t lo = 0, hi = count;
  while (lo < hi) {  // first position whose key is >= want
    const uint64_t mid = (lo + hi) >> 1;
    if (keys[mid] < want)
      lo = mid + 1;
    else
      hi = mid;
  }
  const uint64_t g0 = lo;
  if (g0 + 1 >= count || keys[g0 + 1] != want) return g0;  // the key is unique
  uint64_t g1 = g0 + 2;
  {  // end of the key group: gallop, then bisect
    uint64_t step = 2;
    while (g1 < count && keys[g1] == want) {
      g1 += step;
      step <<= 1;
    }
    uint64_t a = g1 - (step >> 1), b = g1 < count ? g1 : count;  // keys[a] == want (or a == g0 + 1), keys[b] != want
    if (a < g0 + 1) a = g0 + 1;
    while (b - a > 1) {
      const uint64_t mid = (a + b) >> 1;
      if (keys[mid] == want)
        a = mid;
      else
        b = mid;
    }
    g1 = b;
  }
  const unsigned spw = pt.syms_per_word();
  const unsigned limit = static_cast<unsigned>((depth > key_symbols ? depth - key_symbols : 0) / spw) + 2u;
  uint64_t a = g0, b = g1;
  while (a < b) {
    const uint64_t mid = (a + b) >> 1;
    const uint64_t s = sa[mid];
    if (s == j) return mid;
    uint64_t l = 0;
    const bool decided = pt.common_prefix(j, s, key_symbols, limit, &l);
    if (!decided) {  // equal beyond `depth` symbols: j's pair partner (or a group j is next to): look around
      for (uint64_t r = 1; r < g1 - g0; ++r) {
        if (mid >= g0 + r && sa[mid - r] == j) return mid - r;
        if (mid + r < g1 && sa[mid + r] == j) return mid + r;
      }
      return g0;  // not reached: j is in its key group
    }
    const uint64_t shorter = pt.n - (j > s ? j : s);
    const bool j_first = l >= shorter ? j > s : pt.symbol(j + l) < pt.symbol(s + l);
    if (j_first)
      b = mid;
    else
      a = mid + 1;
  }
  for (uint64_t k = g0; k < g1; ++k)  // not reached (kept so that a wrong assumption costs time, not correctness)
    if (sa[k] == j) return k;
  return g0;
}

// ---------------------------------------------------------------------------------------
// Rank storage of the single-GPU path: isa[text position] = first SA position of the
// suffix's current group — but only for suffixes that were ever tied.  A suffix whose key is
// unique never enters the refinement; its rank is its final SA position, which is the position
// of its key in the sorted key array.  Writing those n ranks up front is a random scatter into
// an array far larger than L2 (a 32 B read-modify-write in HBM per 4-byte rank: 130 ms at
// 3.1 G suffixes, and partitioning the scatter into L2-sized windows first was no faster —
// profiles/r01/README.md), so the array is filled with a sentinel instead (one streaming
// memset) and a lookup that hits the sentinel binary-searches the suffix's key.
// ---------------------------------------------------------------------------------------
template <class IdxT>
struct LocalRanks {
  using Comp = typename IdxTraits<IdxT>::Comp;
  static constexpr IdxT kUnset = ~IdxT(0);  // never a group head: heads of tied groups are <= n - 2
  Engine& eng;
  uint64_t n;
  PackedText pt;
  const uint64_t* keys;  // sorted (masked) keys of all n suffixes
  uint64_t key_mask;
  const IdxT* sa;             // the suffix array under construction (all n positions)
  uint64_t key_symbols;       // symbols the key covers
  uint64_t resolved_depth = 0;  // symbols within which every final suffix differs from all others (set by refine_deep)
  static constexpr bool kPublishesAllTied = false;  // ranks of final suffixes are found, not published (rank_of_unpublished)
  DevBuf<IdxT> isa;  // allocated when the refinement first needs ranks (reset)
  LocalRanks(Engine& e, uint64_t n_, const PackedText& pt_, const uint64_t* keys_, uint64_t key_mask_, const IdxT* sa_,
             uint64_t key_symbols_)
      : eng(e), n(n_), pt(pt_), keys(keys_), key_mask(key_mask_), sa(sa_), key_symbols(key_symbols_) {}

  uint64_t global_sum(uint64_t v) { return v; }  // over the ranks of the construction

  // every suffix starts with its SA position as rank: implicit (see above)
  void reset() {
    if (!isa) isa.alloc(n, eng.stream);
    CAPSB_CUDA(cudaMemsetAsync(isa.get(), 0xFF, n * sizeof(IdxT), eng.stream));
  }

  // isa[idx[t]] = head[t] for t in [0, m)
  void publish(const IdxT* idx, const IdxT* head, uint64_t m) {
    IdxT* d_isa = isa.get();
    launch_map(eng.dev, eng.stream, m, [=] __device__(uint64_t t) { d_isa[idx[t]] = head[t]; });
  }

  // answer[j] = {lcp, hi sorts first} for the pairs (lo[j], hi[j]); see chain_pairs_local
  void chain_pairs(IdxT* hi, const IdxT* lo, uint64_t m, uint64_t known, IdxPair<IdxT>* answer) {
    chain_pairs_local<IdxT>(eng, pt, hi, lo, m, known, answer);
  }

  // second[t] = rank of suffix idx[t] + h, or n - 1 - idx[t] when that is beyond the end of the
  // text (shorter suffix first = larger position first; the caller ranks those below every
  // suffix that reaches depth h)
  void second_ranks(const IdxT* idx, uint64_t m, uint64_t h, IdxT* second_out) {
    const IdxT* d_isa = isa.get();
    const uint64_t n_ = n;
    const PackedText text = pt;
    const uint64_t* sorted_keys = keys;
    const uint64_t mask = key_mask;
    const IdxT* d_sa = sa;
    const uint64_t ksym = key_symbols, depth = resolved_depth;
    launch_map(eng.dev, eng.stream, m, [=] __device__(uint64_t t) {
      const uint64_t i = idx[t];
      const uint64_t ih = i + h;
      uint64_t second = n_ - 1 - i;
      if (ih < n_) {
        const IdxT r = d_isa[ih];
        second = r != kUnset ? static_cast<uint64_t>(r)  // still tied when the doubling started: its group head
                             : rank_of_unpublished<IdxT>(text, sorted_keys, d_sa, n_, text.window(ih) & mask, ih, ksym, depth);
      }
      second_out[t] = static_cast<IdxT>(second);
    });
  }
};

// LCP of two suffixes a, b (text positions) whose keys differ: clz(key_a ^ key_b) / bits, bounded
// by the shorter suffix (zero padding can agree with real code-0 symbols past the end).
template <class IdxT>
__device__ __forceinline__ IdxT key_lcp_value(uint64_t key_a, uint64_t key_b, uint64_t a, uint64_t b, uint64_t n,
                                              unsigned log2_bits) {
  const uint64_t shorter = n - (a > b ? a : b);
  const uint64_t l = static_cast<uint64_t>(__clzll(static_cast<long long>(key_a ^ key_b))) >> log2_bits;
  return static_cast<IdxT>(l < shorter ? l : shorter);
}

// The local SA positions that were tied after the key sort (members of key groups of two or
// more), in increasing order: what the stages after the refinement iterate over instead of
// all n positions.
template <class IdxT>
struct TiedSet {
  uint64_t m = 0;
  DevBuf<IdxT> pos;
};

constexpr unsigned kSmallGroup = 32;  // groups up to this size are ordered by counting, not sorting

// d_lcp value of a position whose LCP is still to be computed (an LCP is at most n - 1)
template <class IdxT>
constexpr IdxT kLcpUnset = ~IdxT(0);

// ---------------------------------------------------------------------------------------
// First pass over the key-sorted suffixes (one streaming sweep, 128-bit accesses, four
// consecutive positions per thread): writes the LCP of every position whose key differs from
// its predecessor's (clz of the key XOR), marks the others unset, records in one bitmap which
// positions sit in a key group of two or more (= still tied) and in another which of them start
// their group, and leaves per chunk the count of the tied and (1 +) the last group start —
// what tied_collect_kernel needs to turn the bitmaps into lists without reading the keys again.
// Algorithmic traffic: 8 + w read, w written per suffix.
// Entries next to a tied group are provisional (their bound depends on which member ends up at
// the group's edge); fix_group_edges rewrites them at the end.
// ---------------------------------------------------------------------------------------
constexpr int kKeyLcpThreads = 256;

__device__ __forceinline__ void load4(const uint32_t* p, uint32_t (&v)[4]) {
  const uint4 q = *reinterpret_cast<const uint4*>(p);
  v[0] = q.x, v[1] = q.y, v[2] = q.z, v[3] = q.w;
}
__device__ __forceinline__ void load4(const uint64_t* p, uint64_t (&v)[4]) {
  const ulonglong2 a = *reinterpret_cast<const ulonglong2*>(p), b = *reinterpret_cast<const ulonglong2*>(p + 2);
  v[0] = a.x, v[1] = a.y, v[2] = b.x, v[3] = b.y;
}
__device__ __forceinline__ void store4(uint32_t* p, const uint32_t (&v)[4]) {
  *reinterpret_cast<uint4*>(p) = make_uint4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void store4(uint64_t* p, const uint64_t (&v)[4]) {
  *reinterpret_cast<ulonglong2*>(p) = make_ulonglong2(v[0], v[1]);
  *reinterpret_cast<ulonglong2*>(p + 2) = make_ulonglong2(v[2], v[3]);
}

template <class IdxT>
__global__ void __launch_bounds__(kKeyLcpThreads) key_lcp_count_kernel(const uint64_t* __restrict__ keys,
                                                                       const IdxT* __restrict__ sa,
                                                                       IdxT* __restrict__ lcp,
                                                                       uint32_t* __restrict__ tied_bits,
                                                                       uint32_t* __restrict__ head_bits, uint64_t count,
                                                                       uint64_t skip, uint64_t chunk, uint64_t n,
                                                                       unsigned log2_bits, IdxT* __restrict__ partial,
                                                                       uint64_t* __restrict__ head_partial) {
  __shared__ unsigned warp_sums[kKeyLcpThreads / 32];
  __shared__ uint64_t warp_heads[kKeyLcpThreads / 32];
  uint64_t last_head = 0;  // 1 + the last position of this chunk that starts a key group (0: none)
  const uint64_t begin = static_cast<uint64_t>(blockIdx.x) * chunk;  // a multiple of 128
  const uint64_t end = begin + chunk < count ? begin + chunk : count;
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  unsigned tied_here = 0;
  for (uint64_t wbase = begin + warp * 128ull; wbase < end; wbase += (kKeyLcpThreads / 32) * 128ull) {
    const uint64_t k0 = wbase + 4ull * lane;
    uint64_t key[4];
    IdxT s[4], l[4];
    uint64_t prev_key, next_key;
    IdxT prev_sa;
    // (the first `skip` positions pad the arrays to 16-byte alignment: they belong to the range
    // of positions before this one and are neither counted nor written)
    const bool whole_row = wbase + 128 <= end && wbase >= skip;
    if (wbase + 128 <= end) {  // warp-uniform: the whole row of 128 positions exists
      load4(keys + k0, key);
      load4(sa + k0, s);
      prev_key = __shfl_up_sync(0xffffffffu, key[3], 1);
      prev_sa = __shfl_up_sync(0xffffffffu, s[3], 1);
      next_key = __shfl_down_sync(0xffffffffu, key[0], 1);
      if (lane == 0) {
        prev_key = k0 > 0 ? keys[k0 - 1] : ~key[0];
        prev_sa = k0 > 0 ? sa[k0 - 1] : IdxT(0);
      }
      if (lane == 31) next_key = k0 + 4 < count ? keys[k0 + 4] : ~key[3];
    } else {  // the last row of the array
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint64_t k = k0 + j < count ? k0 + j : count - 1;
        key[j] = keys[k];
        s[j] = sa[k];
      }
      const uint64_t kp = k0 > 0 ? (k0 - 1 < count ? k0 - 1 : count - 1) : 0;
      prev_key = k0 > 0 ? keys[kp] : ~key[0];
      prev_sa = k0 > 0 ? sa[kp] : IdxT(0);
      next_key = k0 + 4 < count ? keys[k0 + 4] : ~key[3];
    }
    unsigned nib = 0, head_nib = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint64_t pk = j ? key[j - 1] : prev_key;
      const uint64_t ps = j ? s[j - 1] : prev_sa;
      const uint64_t nk = j < 3 ? key[j + 1] : next_key;
      const bool exists = (k0 + j < end) & (k0 + j >= skip);
      const bool eq_prev = pk == key[j];
      // past the end of the array the clamped loads repeat the last element: not a neighbour
      const bool eq_next = (nk == key[j]) & (k0 + j + 1 < count);
      l[j] = eq_prev ? kLcpUnset<IdxT> : key_lcp_value<IdxT>(pk, key[j], ps, s[j], n, log2_bits);
      nib |= (exists & (eq_prev | eq_next)) ? 1u << j : 0u;
      head_nib |= (exists & !eq_prev & eq_next) ? 1u << j : 0u;
    }
    if (whole_row) {
      store4(lcp + k0, l);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (k0 + j < end && k0 + j >= skip) lcp[k0 + j] = l[j];
    }
    tied_here += __popc(nib);
    // 8 lanes x 4 flags = one 32-bit word of the bitmap
    unsigned word = nib << (4u * (lane & 7u));
    word |= __shfl_xor_sync(0xffffffffu, word, 1);
    word |= __shfl_xor_sync(0xffffffffu, word, 2);
    word |= __shfl_xor_sync(0xffffffffu, word, 4);
    if ((lane & 7u) == 0 && k0 < end) tied_bits[k0 >> 5] = word;
    // the same for the positions that start a key group of two or more
    if (head_nib) last_head = k0 + (32u - static_cast<unsigned>(__clz(head_nib)));  // (rows come in increasing order)
    unsigned head_word = head_nib << (4u * (lane & 7u));
    head_word |= __shfl_xor_sync(0xffffffffu, head_word, 1);
    head_word |= __shfl_xor_sync(0xffffffffu, head_word, 2);
    head_word |= __shfl_xor_sync(0xffffffffu, head_word, 4);
    if ((lane & 7u) == 0 && k0 < end) head_bits[k0 >> 5] = head_word;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    tied_here += __shfl_xor_sync(0xffffffffu, tied_here, d);
    const uint64_t other = __shfl_xor_sync(0xffffffffu, last_head, d);
    last_head = other > last_head ? other : last_head;
  }
  if (lane == 0) warp_sums[warp] = tied_here, warp_heads[warp] = last_head;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint64_t total = 0, head = 0;
#pragma unroll
    for (int w = 0; w < kKeyLcpThreads / 32; ++w) {
      total += warp_sums[w];
      head = warp_heads[w] > head ? warp_heads[w] : head;
    }
    partial[blockIdx.x] = static_cast<IdxT>(total);
    head_partial[blockIdx.x] = head;
  }
}

// Second half of the first pass: the tied positions (bits of tied_bits) leave for the active list
// — position, suffix, and the head of their key group (the last bit of head_bits at or before
// them) — one bitmap word (32 positions) per lane and step (a compaction over all n positions, one
// flag each, and a running-maximum scan over the tied ones took 10 ms at 3.1 Gbp).  slot_base /
// head_base: per chunk, the exclusive scans (sum / max) of what key_lcp_count_kernel left in
// partial / head_partial.  Positions count from the aligned start (see `skip` above).
template <class IdxT>
__global__ void __launch_bounds__(kScanThreads) tied_collect_kernel(const uint32_t* __restrict__ tied_bits,
                                                                    const uint32_t* __restrict__ head_bits,
                                                                    const IdxT* __restrict__ sa, uint64_t count,
                                                                    uint64_t skip, uint64_t chunk, uint64_t pos_base,
                                                                    const IdxT* __restrict__ slot_base,
                                                                    const uint64_t* __restrict__ head_base,
                                                                    IdxT* __restrict__ pos_a, IdxT* __restrict__ pos_b,
                                                                    IdxT* __restrict__ idx, IdxT* __restrict__ group) {
  constexpr int kWarps = kScanThreads / 32;
  __shared__ unsigned warp_cnt[kWarps];
  __shared__ uint64_t warp_head[kWarps];
  __shared__ uint16_t step_rel[kScanThreads * 32];  // tied positions of one step (one word per thread), relative to its first
  __shared__ uint32_t step_heads[kScanThreads];     // the step's words of head_bits
  __shared__ uint64_t step_before[kScanThreads];    // 1 + the last group start before each word
  const uint64_t begin = static_cast<uint64_t>(blockIdx.x) * chunk;  // a multiple of 128
  const uint64_t end = begin + chunk < count ? begin + chunk : count;
  const uint64_t w_begin = begin >> 5, w_end = (end + 31) >> 5;
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  uint64_t carry_slot = slot_base[blockIdx.x];
  uint64_t carry_head = head_base[blockIdx.x];  // 1 + position of the last group start before this chunk (0: none)
  for (uint64_t wt = w_begin; wt < w_end; wt += kScanThreads) {
    const uint64_t w = wt + threadIdx.x;
    const uint32_t tied = w < w_end ? tied_bits[w] : 0u;
    const uint32_t heads = w < w_end ? head_bits[w] : 0u;
    const unsigned cnt = __popc(tied);
    unsigned inc = cnt;
    uint64_t hmax = heads ? w * 32u + (32u - static_cast<unsigned>(__clz(heads))) : 0ull;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned oc = __shfl_up_sync(0xffffffffu, inc, d);
      const uint64_t oh = __shfl_up_sync(0xffffffffu, hmax, d);
      if (lane >= static_cast<unsigned>(d)) {
        inc += oc;
        hmax = oh > hmax ? oh : hmax;
      }
    }
    if (lane == 31) warp_cnt[warp] = inc, warp_head[warp] = hmax;
    __syncthreads();
    unsigned cnt_before = 0, cnt_total = 0;
    uint64_t head_before = 0, head_total = 0;
#pragma unroll
    for (int v = 0; v < kWarps; ++v) {
      const unsigned c = warp_cnt[v];
      const uint64_t h = warp_head[v];
      if (static_cast<unsigned>(v) < warp) {
        cnt_before += c;
        head_before = h > head_before ? h : head_before;
      }
      cnt_total += c;
      head_total = h > head_total ? h : head_total;
    }
    uint64_t my_head = __shfl_up_sync(0xffffffffu, hmax, 1);         // last group start before this lane's word
    if (lane == 0) my_head = 0;
    my_head = my_head > head_before ? my_head : head_before;
    my_head = my_head > carry_head ? my_head : carry_head;
    // Every lane lists the set bits of its word in shared memory, at the slots they will have in
    // this step's output; then the step's output is written slot by slot, coalesced.  (Measured at
    // 3.1 Gbp, one position in ten tied: the warp taking its words one after the other, lane b for
    // bit b, 5.5 ms; every lane storing its own bits straight to global memory, 9.6 ms.)
    {
      unsigned at = cnt_before + (inc - cnt);  // slot within this step of the word's first tied position
      uint32_t bits = tied;
      while (bits) {
        const unsigned b = static_cast<unsigned>(__ffs(static_cast<int>(bits))) - 1u;
        bits &= bits - 1u;
        step_rel[at++] = static_cast<uint16_t>(threadIdx.x * 32u + b);
      }
      step_heads[threadIdx.x] = heads;
      step_before[threadIdx.x] = my_head;
    }
    __syncthreads();
    for (unsigned j = threadIdx.x; j < cnt_total; j += kScanThreads) {
      const unsigned rel = step_rel[j], wl = rel >> 5, b = rel & 31u;
      const uint64_t word_first = (wt + wl) * 32u;
      const uint64_t k = word_first + b;
      const uint32_t upto = step_heads[wl] & (0xffffffffu >> (31u - b));
      // (a tied position always has a group start at or before it: step_before >= 1 when upto == 0)
      const uint64_t head = upto ? word_first + (31u - static_cast<unsigned>(__clz(upto))) : step_before[wl] - 1;
      const uint64_t slot = carry_slot + j;
      pos_a[slot] = pos_b[slot] = static_cast<IdxT>(k - skip);
      idx[slot] = sa[k];
      group[slot] = static_cast<IdxT>(pos_base + head - skip);
    }
    carry_slot += cnt_total;
    carry_head = head_total > carry_head ? head_total : carry_head;
    __syncthreads();  // the warp totals and the step's lists are rewritten by the next step
  }
}

// The suffixes still being ordered, in SA order; the members of a group are consecutive.
template <class IdxT>
struct ActiveList {
  uint64_t m = 0;
  DevBuf<IdxT> pos;    // local SA position
  DevBuf<IdxT> idx;    // suffix (text position)
  DevBuf<IdxT> group;  // global SA position of the group's first member
};

// ---------------------------------------------------------------------------------------
// Groups of exactly two suffixes {lo < hi} are finished in one step instead of riding through
// the doubling rounds (an exact duplicate of length L keeps ~L pairs tied for log2(L) rounds):
// the pair is compared directly.  What keeps that linear is the permuted-LCP chain rule applied
// before the order is known: if (lo - 1, hi - 1) is also a group of two, then
// lcp(lo, hi) = lcp(lo - 1, hi - 1) - 1 and the two pairs are ordered the same way, so only the
// first pair of a chain of consecutive text positions is compared (plcp_for_pairs); the rest
// inherit.  The LCP of the pair falls out as well, so these positions skip the deep-LCP stage.
//
// chain_pairs_local: answer[j] = {lcp(lo_j, hi_j), 1 if suffix hi_j sorts first else 0} for m
// pairs in any order.  hi[] is overwritten (used as sort buffer).
// ---------------------------------------------------------------------------------------
template <class IdxT>
void chain_pairs_local(Engine& eng, const PackedText& pt, IdxT* hi, const IdxT* lo, uint64_t m, uint64_t known,
                       IdxPair<IdxT>* answer) {
  if (m == 0) return;
  cudaStream_t st = eng.stream;
  DevBuf<IdxT> pos_b(m, st);
  DevBuf<IdxPair<IdxT>> tag_a(m, st), tag_b(m, st);
  {
    IdxPair<IdxT>* ta = tag_a.get();
    launch_map(eng.dev, st, m, [=] __device__(uint64_t j) { ta[j] = IdxPair<IdxT>{lo[j], static_cast<IdxT>(j)}; });
  }
  const unsigned pos_bits = round_up8(bit_length(pt.n - 1));
  const int where = radix_sort_pairs<IdxT, IdxPair<IdxT>>(st, eng.radix, hi, tag_a.get(), pos_b.get(), tag_b.get(), m,
                                                         0, pos_bits);
  const IdxT* pos_i = where ? pos_b.get() : hi;
  const IdxPair<IdxT>* tag = where ? tag_b.get() : tag_a.get();
  plcp_for_pairs<IdxT>(
      eng, pt, pos_i, [=] __device__(uint64_t t) -> uint64_t { return tag[t].a; }, m, known,
      [=] __device__(uint64_t t, IdxT lcp, uint64_t head, IdxT head_lcp) {
        // the order of the chain's first pair decides the whole chain
        const uint64_t a = tag[head].a, b = pos_i[head];  // a < b
        const uint64_t l = head_lcp;
        const bool hi_first = (b + l >= pt.n) ? true : pt.symbol(b + l) < pt.symbol(a + l);
        answer[tag[t].b] = IdxPair<IdxT>{lcp, static_cast<IdxT>(hi_first ? 1 : 0)};
      });
}

// Finishes every group of two in the active list: final order into d_sa, their LCP into d_lcp,
// ranks published (once the refinement keeps ranks), and the pairs leave the list.  Collective in the sharded path (the pairs
// travel to the rank that owns text position hi, where the chains are contiguous).
template <class IdxT, class Ranks>
void resolve_pairs(Engine& eng, Ranks& ranks, ActiveList<IdxT>& act, IdxT* d_sa, IdxT* d_lcp, uint64_t pos_base,
                   uint64_t known, bool publish) {
  cudaStream_t st = eng.stream;
  const uint64_t m = act.m;
  const IdxT* p = act.pos.get();
  const IdxT* s = act.idx.get();
  const IdxT* g = act.group.get();
  auto pair_first = [=] __device__(uint64_t t) -> IdxT {
    const IdxT grp = g[t];
    const IdxT g1 = g[t + 1 < m ? t + 1 : t], g2 = g[t + 2 < m ? t + 2 : t];  // unconditional loads
    const bool first = static_cast<uint64_t>(p[t]) + pos_base == static_cast<uint64_t>(grp);
    return (first & (t + 1 < m) & (g1 == grp) & ((t + 2 >= m) | (g2 != grp))) ? IdxT(1) : IdxT(0);
  };
  const uint64_t pairs = scan_total<IdxT, OpSum>(eng, m, pair_first);
  eng.stats.pairs_chained += pairs;
  DevBuf<IdxT> hi(pairs, st), lo(pairs, st), slot(pairs, st);
  DevBuf<IdxPair<IdxT>> answer(pairs, st);
  if (pairs > 0) {
    IdxT* h = hi.get();
    IdxT* l = lo.get();
    IdxT* sl = slot.get();
    select_finish<IdxT>(eng, m, pair_first, [=] __device__(uint64_t t, IdxT j) {
      const IdxT a = s[t], b = s[t + 1];
      h[j] = a > b ? a : b;
      l[j] = a > b ? b : a;
      sl[j] = static_cast<IdxT>(t);
    });
  }
  ranks.chain_pairs(hi.get(), lo.get(), pairs, known, answer.get());
  DevBuf<IdxT> pub_idx(2 * pairs, st), pub_head(2 * pairs, st);
  if (pairs > 0) {
    const IdxT* l = lo.get();
    const IdxT* sl = slot.get();
    const IdxPair<IdxT>* ans = answer.get();
    IdxT* pi = pub_idx.get();
    IdxT* ph = pub_head.get();
    // hi[] was consumed by the sort: the larger position is the pair's other member
    launch_map(eng.dev, st, pairs, [=] __device__(uint64_t j) {
      const uint64_t t = sl[j];
      const IdxT a = s[t], b = s[t + 1];
      const IdxT low = l[j], high = a == low ? b : a;
      const bool hi_first = ans[j].b != 0;
      const uint64_t k = p[t];
      const IdxT small = hi_first ? high : low, large = hi_first ? low : high;
      d_sa[k] = small;
      d_sa[k + 1] = large;
      d_lcp[k + 1] = ans[j].a;
      pi[2 * j] = small, ph[2 * j] = static_cast<IdxT>(pos_base + k);
      pi[2 * j + 1] = large, ph[2 * j + 1] = static_cast<IdxT>(pos_base + k + 1);
    });
  }
  if (publish) ranks.publish(pub_idx.get(), pub_head.get(), 2 * pairs);

  if (pairs == 0) return;
  auto stays = [=] __device__(uint64_t t) -> IdxT {
    const IdxT grp = g[t];
    const uint64_t first = t - (static_cast<uint64_t>(p[t]) + pos_base - static_cast<uint64_t>(grp));
    const bool pair = first + 1 < m && g[first + 1] == grp && (first + 2 >= m || g[first + 2] != grp);
    return pair ? IdxT(0) : IdxT(1);
  };
  const uint64_t m_next = scan_total<IdxT, OpSum>(eng, m, stays);  // = m - 2 * pairs
  DevBuf<IdxT> n_pos(m_next, st), n_idx(m_next, st), n_group(m_next, st);
  if (m_next > 0) {
    IdxT* np = n_pos.get();
    IdxT* ns = n_idx.get();
    IdxT* ngp = n_group.get();
    select_finish<IdxT>(eng, m, stays, [=] __device__(uint64_t t, IdxT out) {
      np[out] = p[t];
      ns[out] = s[t];
      ngp[out] = g[t];
    });
  }
  act.pos = std::move(n_pos);
  act.idx = std::move(n_idx);
  act.group = std::move(n_group);
  act.m = m_next;
}

// Groups of kSmallGroup < size <= kMidGroup suffixes (the bulk of a repeat family's ties) are
// sorted on chip, by a bitonic network on (comp, suffix) held in REGISTERS: two passes over their
// elements in HBM instead of the 7-9 radix passes the global sort would spend on them.  The
// cooperating threads (a warp, or the CTA for the larger groups) hold kE consecutive network
// positions each, so a compare-exchange step is
//   * register moves                      when its stride stays inside a thread  (stride < kE),
//   * one shuffle per 32-bit word         when it stays inside a warp            (stride < 32 kE),
//   * a trip through shared memory        otherwise (6 of the 78 steps of a 4096-element sort).
// (The first version kept the whole group in shared memory and paid four loads and up to four
// stores per exchange plus a barrier per step: 17.5 ms per construction at 3.1 Gbp, profiles/r02.)
constexpr unsigned kMidGroup = 4096;
constexpr int kGroupSortThreads = 256;
constexpr unsigned kWarpGroup = 256;  // up to here a warp sorts the group, above it a CTA does

template <class T>
__device__ __forceinline__ T shfl_xor_words(T x, unsigned lane_mask) {
  static_assert(sizeof(T) % 4 == 0, "shuffled 32 bits at a time");
  constexpr int kWords = sizeof(T) / 4;
  uint32_t w[kWords];
  memcpy(w, &x, sizeof(T));
#pragma unroll
  for (int i = 0; i < kWords; ++i) w[i] = __shfl_xor_sync(0xffffffffu, w[i], lane_mask);
  T r;
  memcpy(&r, w, sizeof(T));
  return r;
}

// (comp, suffix) pairs are distinct except for the padding, whose copies are interchangeable.
template <class CompT, class IdxT>
__device__ __forceinline__ bool pair_greater(CompT a, IdxT va, CompT b, IdxT vb) {
  return (a > b) | ((a == b) & (va > vb));
}

// Selects, not branches: which way an exchange goes differs from lane to lane.
template <class CompT, class IdxT>
__device__ __forceinline__ void compare_exchange(CompT& a, IdxT& va, CompT& b, IdxT& vb, bool up) {
  const bool greater = pair_greater(a, va, b, vb);
  const bool swap = greater == up;
  const CompT lo = swap ? b : a, hi = swap ? a : b;
  const IdxT vlo = swap ? vb : va, vhi = swap ? va : vb;
  a = lo, b = hi;
  va = vlo, vb = vhi;
}

template <class CompT, class IdxT>
__device__ __forceinline__ void keep_one(CompT& k, IdxT& v, CompT other_k, IdxT other_v, bool keep_min) {
  const bool mine_greater = pair_greater(k, v, other_k, other_v);
  const bool take = mine_greater == keep_min;
  k = take ? other_k : k;
  v = take ? other_v : v;
}

// Sorts kE * kT pairs ascending: thread t of the kT cooperating ones (a warp: kT == 32, or the
// whole CTA) passes any kE of them and leaves with the pairs of rank t * kE .. t * kE + kE - 1.
// sk / sv: kE * kT staging elements in shared memory, used only when kT > 32.
template <class CompT, class IdxT, int kE, int kT>
__device__ __forceinline__ void bitonic_sort_blocked(CompT (&k)[kE], IdxT (&v)[kE], unsigned t, CompT* sk, IdxT* sv) {
  constexpr unsigned kTotal = static_cast<unsigned>(kE) * kT;
  const unsigned base = t * kE;  // network position of k[0]
  // merges that stay inside a thread
#pragma unroll
  for (unsigned span = 2; span <= static_cast<unsigned>(kE); span <<= 1) {
#pragma unroll
    for (unsigned stride = span >> 1; stride > 0; stride >>= 1) {
#pragma unroll
      for (unsigned i = 0; i < static_cast<unsigned>(kE); ++i)
        if ((i & stride) == 0) compare_exchange(k[i], v[i], k[i | stride], v[i | stride], ((base + i) & span) == 0);
    }
  }
#pragma unroll 1
  for (unsigned span = 2 * kE; span <= kTotal; span <<= 1) {
    const bool up = (base & span) == 0;  // the same for all kE positions: span > kE
    unsigned stride = span >> 1;
    if constexpr (kT > 32) {
#pragma unroll 1
      for (; stride >= 32u * kE; stride >>= 1) {  // the partner thread sits in another warp
        const unsigned partner = t ^ (stride / kE);
        const bool keep_min = (t < partner) == up;
        __syncthreads();  // the previous trip's readers are done
#pragma unroll
        for (int i = 0; i < kE; ++i) sk[i * kT + t] = k[i], sv[i * kT + t] = v[i];
        __syncthreads();
#pragma unroll
        for (int i = 0; i < kE; ++i) keep_one(k[i], v[i], sk[i * kT + partner], sv[i * kT + partner], keep_min);
      }
    }
#pragma unroll 1
    for (; stride >= static_cast<unsigned>(kE); stride >>= 1) {  // the partner is a lane of this warp
      const unsigned lane_mask = stride / kE;
      const bool keep_min = ((t & lane_mask) == 0) == up;
#pragma unroll
      for (int i = 0; i < kE; ++i) keep_one(k[i], v[i], shfl_xor_words(k[i], lane_mask), shfl_xor_words(v[i], lane_mask), keep_min);
    }
#pragma unroll
    for (unsigned s = kE >> 1; s > 0; s >>= 1) {
#pragma unroll
      for (unsigned i = 0; i < static_cast<unsigned>(kE); ++i)
        if ((i & s) == 0) compare_exchange(k[i], v[i], k[i | s], v[i | s], up);
    }
  }
}

// One group of `size` <= kE * kT pairs starting at `first`: loaded (coalesced; which network
// position a pair starts from does not matter), padded with pairs that sort last, sorted, stored.
template <class CompT, class IdxT, int kE, int kT>
__device__ __forceinline__ void sort_one_group(uint64_t first, unsigned size, unsigned t, const CompT* __restrict__ comp_in,
                                               const IdxT* __restrict__ idx_in, CompT* __restrict__ comp_out,
                                               IdxT* __restrict__ idx_out, CompT* sk, IdxT* sv) {
  CompT k[kE];
  IdxT v[kE];
#pragma unroll
  for (int i = 0; i < kE; ++i) {
    const unsigned e = i * kT + t;
    // padding sorts last: largest comp and a suffix index no suffix has
    k[i] = e < size ? comp_in[first + e] : ~CompT(0);
    v[i] = e < size ? idx_in[first + e] : ~IdxT(0);
  }
  bitonic_sort_blocked<CompT, IdxT, kE, kT>(k, v, t, sk, sv);
#pragma unroll
  for (int i = 0; i < kE; ++i) {
    const unsigned e = t * kE + i;
    if (e < size) comp_out[first + e] = k[i], idx_out[first + e] = v[i];
  }
}

// kWarpGroup < size <= kMidGroup: one CTA per group at a time.
template <class CompT, class IdxT>
__global__ void __launch_bounds__(kGroupSortThreads) group_sort_kernel(const IdxT* __restrict__ group_first,
                                                                       const IdxT* __restrict__ group_size,
                                                                       const unsigned long long* __restrict__ group_count,
                                                                       const CompT* __restrict__ comp_in,
                                                                       const IdxT* __restrict__ idx_in,
                                                                       CompT* __restrict__ comp_out,
                                                                       IdxT* __restrict__ idx_out) {
  extern __shared__ __align__(16) unsigned char group_sort_smem[];
  CompT* sk = reinterpret_cast<CompT*>(group_sort_smem);
  IdxT* sv = reinterpret_cast<IdxT*>(group_sort_smem + sizeof(CompT) * kMidGroup);
  constexpr int kT = kGroupSortThreads;
  static_assert(kMidGroup == 16u * kT, "the largest group fills sixteen registers per thread");
  const unsigned long long groups = *group_count;
  for (unsigned long long grp = blockIdx.x; grp < groups; grp += gridDim.x) {
    const uint64_t first = group_first[grp];
    const unsigned size = static_cast<unsigned>(group_size[grp]);
    const unsigned t = threadIdx.x;
    if (size <= 2u * kT)
      sort_one_group<CompT, IdxT, 2, kT>(first, size, t, comp_in, idx_in, comp_out, idx_out, sk, sv);
    else if (size <= 4u * kT)
      sort_one_group<CompT, IdxT, 4, kT>(first, size, t, comp_in, idx_in, comp_out, idx_out, sk, sv);
    else if (size <= 8u * kT)
      sort_one_group<CompT, IdxT, 8, kT>(first, size, t, comp_in, idx_in, comp_out, idx_out, sk, sv);
    else
      sort_one_group<CompT, IdxT, 16, kT>(first, size, t, comp_in, idx_in, comp_out, idx_out, sk, sv);
  }
}

// kSmallGroup < size <= kWarpGroup — the bulk of the mid groups: a repeat family's 20-mer with one
// substitution is shared by ~100 copies — one warp per group, eight groups per CTA at a time, no
// shared memory and no barriers at all.
template <class CompT, class IdxT>
__global__ void __launch_bounds__(kGroupSortThreads) group_sort_warp_kernel(const IdxT* __restrict__ group_first,
                                                                            const IdxT* __restrict__ group_size,
                                                                            const unsigned long long* __restrict__ group_count,
                                                                            const CompT* __restrict__ comp_in,
                                                                            const IdxT* __restrict__ idx_in,
                                                                            CompT* __restrict__ comp_out,
                                                                            IdxT* __restrict__ idx_out) {
  constexpr int kWarps = kGroupSortThreads / 32;
  static_assert(kWarpGroup == 8u * 32u, "the largest warp-sorted group fills eight registers per lane");
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const unsigned long long groups = *group_count;
  for (unsigned long long grp = static_cast<unsigned long long>(blockIdx.x) * kWarps + warp; grp < groups;
       grp += static_cast<unsigned long long>(gridDim.x) * kWarps) {
    const uint64_t first = group_first[grp];
    const unsigned size = static_cast<unsigned>(group_size[grp]);
    if (size <= 64u)
      sort_one_group<CompT, IdxT, 2, 32>(first, size, lane, comp_in, idx_in, comp_out, idx_out, nullptr, nullptr);
    else if (size <= 128u)
      sort_one_group<CompT, IdxT, 4, 32>(first, size, lane, comp_in, idx_in, comp_out, idx_out, nullptr, nullptr);
    else
      sort_one_group<CompT, IdxT, 8, 32>(first, size, lane, comp_in, idx_in, comp_out, idx_out, nullptr, nullptr);
  }
}

// One refinement round on the active list.  comp_a[t] orders the members of a group: the low
// second_bits bits are the second key (text beyond the current depth, or a rank); with
// kGroupInComp the group head (plus one for suffixes that reach the current depth inside the
// text) sits above it, at bit 32 (64-bit comps) or 64 (128-bit comps), otherwise the comps
// carry no group and only the large-group sort prepends it.  Orders every group, writes the new
// order to d_sa, publishes the new group heads (when the refinement keeps ranks) and drops the
// suffixes that are now alone in their group.
//
// Text rounds (comps = flag bit . next 63 bits of packed text, see refine_tied_groups) also settle
// LCPs: where two neighbours of one group part ways in this round, their common prefix is the
// depth the group agreed on plus the common leading symbols of their comps — the same
// clz-of-XOR rule that gives the key-derived LCPs, one round deeper — so the entry is written here
// and the pair never reaches the deep-LCP stage.  The value does not depend on which members end
// up at the edges of the two new groups; only its bound by the shorter suffix does, and
// refine_tied_groups applies that bound to every tied position once the order is final.
template <class IdxT>
struct TextRoundLcp {
  IdxT* d_lcp = nullptr;  // null: not a text round
  uint64_t depth = 0;     // symbols every member of a group agrees on before this round
  uint64_t n = 0;
  unsigned log2_bits = 0;
};

template <class IdxT, class CompT, bool kGroupInComp, class Ranks>
void refine_round(Engine& eng, ActiveList<IdxT>& act, IdxT* d_sa, uint64_t pos_base, DevBuf<CompT> comp_a,
                  unsigned second_bits, unsigned rank_bits, Ranks* publish_to,
                  TextRoundLcp<IdxT> text_lcp = TextRoundLcp<IdxT>()) {
  using Wide = unsigned __int128;
  cudaStream_t st = eng.stream;
  const DeviceInfo& dev = eng.dev;
  const uint64_t m = act.m;
  DevBuf<CompT> comp_b(m, st);
  DevBuf<IdxT> idx_b(m, st);

  // Order every group by the second field.  Groups of at most kSmallGroup suffixes — after the
  // first rounds nearly all of them — are ranked by counting inside the group (one pass, no
  // sort); groups of up to kMidGroup are sorted by one CTA each in shared memory; only the rest
  // (poly-A runs, tandem arrays, periodic texts) is radix-sorted by (group, second).
  CompT* sorted_c = comp_b.get();
  IdxT* sorted_i = idx_b.get();
  trace_point(eng, "  round: comps built");
  DevBuf<uint8_t> big_flag(m, st);
  const uint64_t mid_capacity = m / (kSmallGroup + 1) + 1;
  DevBuf<IdxT> mid_first(mid_capacity, st), mid_size(mid_capacity, st);      // kWarpGroup < size <= kMidGroup
  DevBuf<IdxT> warp_first(mid_capacity, st), warp_size(mid_capacity, st);    // kSmallGroup < size <= kWarpGroup
  DevBuf<unsigned long long> mid_count(2, st);                               // [0] CTA-sorted, [1] warp-sorted
  CAPSB_CUDA(cudaMemsetAsync(mid_count.get(), 0, 2 * sizeof(unsigned long long), st));
  {
    const CompT* c = comp_a.get();
    const IdxT* s = act.idx.get();
    const IdxT* g = act.group.get();
    const IdxT* p = act.pos.get();
    uint8_t* big = big_flag.get();
    IdxT* mf = mid_first.get();
    IdxT* ms = mid_size.get();
    IdxT* wf = warp_first.get();
    IdxT* ws = warp_size.get();
    unsigned long long* mc = mid_count.get();
    static const unsigned warp_limit_env = [] {  // tuning knob: groups up to here go to the warp kernel
      const char* e = std::getenv("CAPSB_WARP_GROUP");
      const long v = e ? std::atol(e) : static_cast<long>(kWarpGroup);
      return static_cast<unsigned>(v < static_cast<long>(kSmallGroup) ? kSmallGroup : v > static_cast<long>(kWarpGroup) ? kWarpGroup : v);
    }();
    const unsigned warp_limit = warp_limit_env;
    launch_map(dev, st, m, [=] __device__(uint64_t t) {
      const IdxT grp = g[t];
      // the members of a group are consecutive in the list, in SA order
      const uint64_t first = t - (static_cast<uint64_t>(p[t]) - (static_cast<uint64_t>(grp) - pos_base));
      if (first + kMidGroup < m && g[first + kMidGroup] == grp) {
        big[t] = 1;
        return;
      }
      big[t] = 0;
      if (first + kSmallGroup < m && g[first + kSmallGroup] == grp) {
        if (t == first) {  // the group's first member announces it: last index with the same head
          uint64_t lo = first + kSmallGroup, hi = first + kMidGroup < m ? first + kMidGroup : m;  // g[lo] == grp, g[hi] != grp
          while (hi - lo > 1) {
            const uint64_t mid = (lo + hi) >> 1;
            if (g[mid] == grp)
              lo = mid;
            else
              hi = mid;
          }
          const bool by_warp = hi - first <= warp_limit;
          const unsigned long long slot = atomicAdd(mc + (by_warp ? 1 : 0), 1ull);
          (by_warp ? wf : mf)[slot] = static_cast<IdxT>(first);
          (by_warp ? ws : ms)[slot] = static_cast<IdxT>(hi - first);
        }
        return;
      }
      const CompT mine = c[t];
      unsigned rank = 0;
      for (uint64_t u = first; u < m && u < first + kSmallGroup && g[u] == grp; ++u) {
        const CompT other = c[u];
        rank += (other < mine || (other == mine && u < t)) ? 1u : 0u;
      }
      sorted_c[first + rank] = mine;
      sorted_i[first + rank] = s[t];
    });
    trace_point(eng, "  round: classified + small groups counted");
    constexpr size_t kSortSmem = (sizeof(CompT) + sizeof(IdxT)) * kMidGroup;
    // (every launch: the attribute is per context, and host threads of other ranks may be here too)
    CAPSB_CUDA(cudaFuncSetAttribute(group_sort_kernel<CompT, IdxT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(kSortSmem)));
    const uint64_t want = mid_capacity < static_cast<uint64_t>(dev.sm_count) * 8 ? mid_capacity
                                                                                   : static_cast<uint64_t>(dev.sm_count) * 8;
    CAPSB_LAUNCH((group_sort_kernel<CompT, IdxT>), static_cast<unsigned>(want), kGroupSortThreads, kSortSmem, st,
                 mid_first.get(), mid_size.get(), mid_count.get(), c, s, sorted_c, sorted_i);
    const uint64_t warp_ctas = mid_capacity / (kGroupSortThreads / 32) + 1;
    const uint64_t warp_grid = warp_ctas < static_cast<uint64_t>(dev.sm_count) * 16 ? warp_ctas
                                                                                     : static_cast<uint64_t>(dev.sm_count) * 16;
    CAPSB_LAUNCH((group_sort_warp_kernel<CompT, IdxT>), static_cast<unsigned>(warp_grid), kGroupSortThreads, 0, st,
                 warp_first.get(), warp_size.get(), mid_count.get() + 1, c, s, sorted_c, sorted_i);
  }
  trace_point(eng, "  round: mid groups sorted");
  {
    const uint8_t* big = big_flag.get();
    auto is_big = [=] __device__(uint64_t t) -> IdxT { return big[t]; };
    const uint64_t big_count = scan_total<IdxT, OpSum>(eng, m, is_big);
    eng.stats.refine_sorted += big_count;
    eng.stats.refine_counted += m - big_count;
    if (trace_enabled())
      std::fprintf(stderr, "[capsb]   round: %llu suffixes in small groups (counted), %llu in large groups (sorted)\n",
                   (unsigned long long)(m - big_count), (unsigned long long)big_count);
    if (big_count > 0) {
      // sort key of the large groups: the group head above the comp (already there when kGroupInComp)
      using BigKey = std::conditional_t<kGroupInComp, CompT, Wide>;
      constexpr unsigned kGroupShift = kGroupInComp ? sizeof(CompT) * 4 : 64;
      DevBuf<BigKey> bc_a(big_count, st), bc_b(big_count, st);
      DevBuf<IdxT> bi_a(big_count, st), bi_b(big_count, st), slot_of(big_count, st);
      {
        const CompT* c = comp_a.get();
        const IdxT* s = act.idx.get();
        const IdxT* g = act.group.get();
        BigKey* bc = bc_a.get();
        IdxT* bi = bi_a.get();
        IdxT* so = slot_of.get();
        select_finish<IdxT>(eng, m, is_big, [=] __device__(uint64_t t, IdxT j) {
          bc[j] = kGroupInComp ? static_cast<BigKey>(c[t])
                               : static_cast<BigKey>((static_cast<Wide>(static_cast<uint64_t>(g[t])) << 64) |
                                                     static_cast<Wide>(c[t]));
          bi[j] = s[t];
          so[j] = static_cast<IdxT>(t);
        });
      }
      // LSD over the second field, then the group field
      BigKey* kin = bc_a.get();
      IdxT* vin = bi_a.get();
      BigKey* kout = bc_b.get();
      IdxT* vout = bi_b.get();
      auto passes = [&](unsigned lo, unsigned hi) {
        for (unsigned shift = lo; shift < hi; shift += 8) {
          radix_pass<BigKey, IdxT>(st, eng.radix, ArraySource<BigKey, IdxT>{kin, vin}, big_count, shift, kout, vout);
          std::swap(kin, kout);
          std::swap(vin, vout);
        }
      };
      passes(0, second_bits);
      {  // the group field (with the in-range bit added into it when kGroupInComp)
        const unsigned top = kGroupShift + rank_bits + 8;
        passes(kGroupShift, top < sizeof(BigKey) * 8 ? top : static_cast<unsigned>(sizeof(BigKey) * 8));
      }
      // the sorted sub-list keeps the groups in list order, so its j-th element belongs in the
      // j-th slot that a large group occupies
      const IdxT* so = slot_of.get();
      const BigKey* sc = kin;
      const IdxT* si = vin;
      launch_map(dev, st, big_count, [=] __device__(uint64_t j) {
        sorted_c[so[j]] = static_cast<CompT>(sc[j]);
        sorted_i[so[j]] = si[j];
      });
    }
  }
  const CompT* sorted_comp = sorted_c;
  const IdxT* sorted_idx = sorted_i;
  trace_point(eng, "  round: large groups sorted");

  const IdxT* old_group = act.group.get();
  const IdxT* p = act.pos.get();
  DevBuf<IdxT> n_pos, n_idx, n_group;
  uint64_t m_next = 0;
  // The LCP a text round settles between position t - 1 and t of the sorted list (see above).
  auto settle_lcp = [=] __device__(uint64_t t, IdxT here) {
    // (captured outside the constexpr-if: an extended lambda may not first-capture inside one)
    const TextRoundLcp<IdxT> tl = text_lcp;
    const IdxT* og = old_group;
    const CompT* sc = sorted_comp;
    const IdxT* si = sorted_idx;
    const IdxT* pp = p;
    if constexpr (!kGroupInComp && sizeof(CompT) == 8) {
      if (tl.d_lcp != nullptr && t > 0 && og[t] == og[t - 1]) {
        const uint64_t ca = static_cast<uint64_t>(sc[t - 1]), cb = static_cast<uint64_t>(sc[t]);
        if (ca != cb) {
          uint64_t l;
          if ((ca & cb) >> 63) {  // both suffixes reach the depth: compare the 63 text bits
            const uint64_t x = (ca ^ cb) << 1;
            l = tl.depth + (static_cast<uint64_t>(__clzll(static_cast<long long>(x))) >> tl.log2_bits);
          } else {  // one of them ends before the depth: it is a prefix of its neighbour
            const uint64_t a = si[t - 1], b = here;
            l = tl.n - (a > b ? a : b);
          }
          tl.d_lcp[pp[t]] = static_cast<IdxT>(l);
        }
      }
    }
    (void)t, (void)here, (void)tl, (void)og, (void)sc, (void)si, (void)pp;
  };
  if constexpr (sizeof(IdxT) == 4) {
    // One scan does it all (32-bit indices: the list index of a group's first member and the count of
    // the suffixes that stay tied share a 64-bit value): the first pass counts, so that the new
    // list can be allocated; the second writes the order, the LCPs, the new group heads and the new list.
    // A new group starts where the comp changes or the old group does (comps without a group field
    // can agree across a boundary); a suffix leaves the list when it is alone in its new group.
    // (all loads unconditional, on clamped indices: a short-circuit would chain their latencies)
    auto starts = [=] __device__(uint64_t t) -> bool {
      const uint64_t tp = t > 0 ? t - 1 : 0;
      return (t == 0) | (sorted_comp[t] != sorted_comp[tp]) | (old_group[t] != old_group[tp]);
    };
    auto head_and_tied = [=] __device__(uint64_t t) -> uint64_t {
      const uint64_t tn = t + 1 < m ? t + 1 : t;
      const bool s0 = starts(t), s1 = starts(tn) | (t + 1 == m);
      return ((s0 ? t : 0ull) << 32) | ((s0 & s1) ? 0u : 1u);
    };
    m_next = static_cast<uint32_t>(scan_total<uint64_t, OpMaxHiSumLo>(eng, m, head_and_tied));
    n_pos.alloc(m_next, st), n_idx.alloc(m_next, st), n_group.alloc(m_next, st);
    DevBuf<IdxT> new_group(publish_to ? m : 0, st);
    IdxT* ng = publish_to ? new_group.get() : nullptr;
    IdxT* np = n_pos.get();
    IdxT* ns = n_idx.get();
    IdxT* ngp = n_group.get();
    scan_finish<uint64_t, OpMaxHiSumLo, true>(eng, m, head_and_tied, [=] __device__(uint64_t t, uint64_t v) {
      const uint64_t head = v >> 32;  // list index of the first member of t's new group
      const IdxT here = sorted_idx[t];
      const IdxT at = p[t];
      const IdxT group_head = static_cast<IdxT>(pos_base + p[head]);
      d_sa[at] = here;
      if (ng) ng[t] = group_head;
      settle_lcp(t, here);
      const uint64_t tn = t + 1 < m ? t + 1 : t;
      const bool alone = (head == t) & (starts(tn) | (t + 1 == m));
      if (!alone) {
        const uint32_t slot = static_cast<uint32_t>(v) - 1u;  // inclusive count of the tied, this one included
        np[slot] = at;
        ns[slot] = here;
        ngp[slot] = group_head;
      }
    });
    trace_point(eng, "  round: heads + sa + new list written");
    if (publish_to) {
      publish_to->publish(sorted_idx, new_group.get(), m);
      trace_point(eng, "  round: ranks published");
    }
  } else {
    DevBuf<IdxT> head_slot(m, st);
    IdxT* hs = head_slot.get();
    scan_full<IdxT, OpMax, true>(
        eng, m,
        [=] __device__(uint64_t t) -> IdxT {
          // a new group starts where the comp changes or the old group does (comps without a group
          // field can agree across a boundary)
          return (t > 0 && (sorted_comp[t] != sorted_comp[t - 1] || old_group[t] != old_group[t - 1])) ? static_cast<IdxT>(t)
                                                                                                    : IdxT(0);
        },
        [=] __device__(uint64_t t, IdxT head) { hs[t] = head; });

    DevBuf<IdxT> new_group(m, st);
    {
      IdxT* ng = new_group.get();
      launch_map(dev, st, m, [=] __device__(uint64_t t) {
        const IdxT here = sorted_idx[t];
        d_sa[p[t]] = here;
        ng[t] = static_cast<IdxT>(pos_base + p[hs[t]]);
        settle_lcp(t, here);
      });
    }
    trace_point(eng, "  round: heads + sa written");
    if (publish_to) {
      publish_to->publish(sorted_idx, new_group.get(), m);
      trace_point(eng, "  round: ranks published");
    }

    auto still_tied = [=] __device__(uint64_t t) -> IdxT {
      const IdxT h0 = hs[t], h1 = hs[t + 1 < m ? t + 1 : t];  // unconditional loads
      const bool single = (h0 == t) & ((t + 1 == m) | (h1 == t + 1));
      return single ? IdxT(0) : IdxT(1);
    };
    m_next = scan_total<IdxT, OpSum>(eng, m, still_tied);
    n_pos.alloc(m_next, st), n_idx.alloc(m_next, st), n_group.alloc(m_next, st);
    if (m_next > 0) {
      const IdxT* ng = new_group.get();
      IdxT* np = n_pos.get();
      IdxT* ns = n_idx.get();
      IdxT* ngp = n_group.get();
      select_finish<IdxT>(eng, m, still_tied, [=] __device__(uint64_t t, IdxT slot) {
        np[slot] = p[t];
        ns[slot] = sorted_idx[t];
        ngp[slot] = ng[t];
      });
    }
  }
  act.pos = std::move(n_pos);
  act.idx = std::move(n_idx);
  act.group = std::move(n_group);
  act.m = m_next;
}

// ---------------------------------------------------------------------------------------
// Refinement of the ties the key sort leaves.  d_sa[0..count) holds suffixes in key order (SA
// positions pos_base .. pos_base + count of the final array); keys[] are their (masked) keys,
// which order them by their first h0 symbols.  Groups of equal keys must be complete inside
// [0, count).  On return d_sa is in suffix order, d_lcp holds every entry except those of
// neighbours that stayed tied into the rank rounds (kLcpUnset: the permuted-LCP stage of the
// caller, collect_deep_pairs + plcp_for_pairs), and `tied` lists the positions that were tied.
//
//   first pass   key-derived LCPs, bitmap and list of the tied positions (key_lcp_count_kernel);
//   pair chains  groups of two are finished by one comparison per chain (resolve_pairs);
//   text rounds  groups ordered by the next 63 bits of text: no rank array, no lookups, in the
//                sharded path no exchange; the LCPs of the neighbours they separate fall out;
//   rank rounds  prefix doubling on ranks kept by the Ranks policy, which only ever stores ranks
//                of suffixes that have been tied (the rest is implicit, LocalRanks / ShardedRanks);
//   last pass    LCPs at the edges of the key groups, bound by the shorter suffix inside them.
// ---------------------------------------------------------------------------------------
// The refinement runs in three pieces so that the single-GPU path can finish the suffix array a
// range of positions at a time (sa_build.cu streams every finished range to the host while the
// next one is refined):
//   refine_shallow  first pass, pair chains and text rounds on positions [0, count) of d_sa: local
//                   to the range, no rank array.  Leaves the suffixes it could not separate (deep
//                   ties: tandem arrays, long duplications, periodic texts) in `act`, at depth h.
//   refine_deep     pair chains and rank rounds (prefix doubling) on whatever is left; needs the
//                   ranks of the whole text, i.e. every range at the end of its shallow phase.
//   fix_group_edges key-derived LCPs at the edges of the key groups and the shorter-suffix bound
//                   inside them, for the listed tied positions (idempotent once the order is final).
// refine_tied_groups runs the three in sequence over one range (the sharded path: a rank's bucket).
template <class IdxT>
struct RefineState {
  ActiveList<IdxT> act;
  uint64_t h = 0;             // symbols every group of act agrees on
  uint64_t h_resolved = 0;    // symbols within which every suffix that left the list differs from all others
                              // (pairs ordered by the pair-chain step excepted); >= h
  uint64_t total_active = 0;  // over the ranks of the construction
};

template <class IdxT, class Ranks>
void refine_shallow(Engine& eng, Ranks& ranks, const PackedText& pt, unsigned key_bits, const uint64_t* keys, IdxT* d_sa,
                    IdxT* d_lcp, uint64_t count, uint64_t pos_base, uint64_t n, TiedSet<IdxT>& tied,
                    RefineState<IdxT>& state) {
  cudaStream_t st = eng.stream;
  const unsigned log2_bits = pt.log2_bits;

  // The suffixes still to be ordered: members of key groups with at least two suffixes.  The same
  // pass writes every LCP that the keys alone decide (neighbours with different keys) and marks
  // the rest unset; entries next to a tied group are provisional (their bound depends on which
  // member ends up at the group's edge) and are rewritten by fix_group_edges.
  // (all loads unconditional, on clamped indices: a short-circuit would chain their latencies)
  auto in_group = [=] __device__(uint64_t k) -> IdxT {
    const uint64_t kp = k > 0 ? k - 1 : 0, kn = k + 1 < count ? k + 1 : k;
    const uint64_t a = keys[kp], b = keys[k], c = keys[kn];
    return ((kp != k && a == b) | (kn != k && c == b)) ? IdxT(1) : IdxT(0);
  };
  ActiveList<IdxT>& act = state.act;
  // The streaming pass wants 16-byte aligned arrays.  A range that starts in the middle of larger
  // arrays (sa_build.cu: one range of positions at a time) is reached by stepping back `skip` < 4
  // positions — entries of the range before it, which the pass reads but neither counts nor writes.
  const uint64_t skip = (reinterpret_cast<uintptr_t>(d_sa) & 15u) / sizeof(IdxT);
  const bool vector_ok =
      skip <= pos_base &&
      ((reinterpret_cast<uintptr_t>(keys - skip) | reinterpret_cast<uintptr_t>(d_sa - skip) |
        reinterpret_cast<uintptr_t>(d_lcp - skip)) & 15u) == 0;
  auto alloc_lists = [&] {
    act.pos.alloc(act.m, st);
    act.idx.alloc(act.m, st);
    act.group.alloc(act.m, st);
    tied.m = act.m;
    tied.pos.alloc(act.m, st);
  };
  if (vector_ok) {
    // one streaming sweep: key-derived LCPs, bitmaps of the tied positions and of the group starts,
    // their count / last group start per chunk (key_lcp_count_kernel); then the lists, 32 positions
    // per lane at a time (tied_collect_kernel)
    ScanScratch<IdxT>& sc = eng.scan_scratch<IdxT>();
    const Chunking ck = make_chunking(count + skip, kScanTile, sc.max_blocks);
    const uint64_t words = (count + skip) / 32 + 8;
    DevBuf<uint32_t> tied_bits(words, st), head_bits(words, st);
    DevBuf<uint64_t> head_partial(ck.blocks, st);
    CAPSB_LAUNCH((key_lcp_count_kernel<IdxT>), ck.blocks, kKeyLcpThreads, 0, st, keys - skip, d_sa - skip, d_lcp - skip,
                 tied_bits.get(), head_bits.get(), count + skip, skip, ck.chunk, n, log2_bits, sc.partial.get(),
                 head_partial.get());
    CAPSB_LAUNCH((scan_spine_kernel<IdxT, OpSum>), 1, kScanThreads, 0, st, ck.blocks, sc.partial.get(), sc.total.get());
    CAPSB_LAUNCH((scan_spine_kernel<uint64_t, OpMax>), 1, kScanThreads, 0, st, ck.blocks, head_partial.get(),
                 static_cast<uint64_t*>(nullptr));
    IdxT total;
    read_back(st, &total, sc.total.get(), sizeof(IdxT));
    act.m = total;
    alloc_lists();
    if (act.m > 0)
      CAPSB_LAUNCH((tied_collect_kernel<IdxT>), ck.blocks, kScanThreads, 0, st, tied_bits.get(), head_bits.get(),
                   d_sa - skip, count + skip, skip, ck.chunk, pos_base, sc.partial.get(), head_partial.get(),
                   act.pos.get(), tied.pos.get(), act.idx.get(), act.group.get());
  } else {  // arrays that are not 16-byte aligned: the same, element by element
    auto in_group_and_lcp = [=] __device__(uint64_t k) -> IdxT {
      const uint64_t kp = k > 0 ? k - 1 : 0, kn = k + 1 < count ? k + 1 : k;
      const uint64_t a = keys[kp], b = keys[k], c = keys[kn];
      const IdxT sa_prev = d_sa[kp], sa_here = d_sa[k];
      if (k > 0)
        d_lcp[k] = a == b ? kLcpUnset<IdxT>  // until the pair-chain step or the deep-LCP stage
                          : key_lcp_value<IdxT>(a, b, sa_prev, sa_here, n, log2_bits);
      return ((kp != k && a == b) | (kn != k && c == b)) ? IdxT(1) : IdxT(0);
    };
    act.m = scan_total<IdxT, OpSum>(eng, count, in_group_and_lcp);
    alloc_lists();
    IdxT* p = act.pos.get();
    IdxT* p0 = tied.pos.get();
    IdxT* s = act.idx.get();
    IdxT* g = act.group.get();
    select_finish<IdxT>(eng, count, in_group, [=] __device__(uint64_t k, IdxT slot) {
      p[slot] = p0[slot] = static_cast<IdxT>(k);
      s[slot] = d_sa[k];
    });
    // group head (global SA position) of every active suffix: running maximum of the heads
    scan_full<IdxT, OpMax, true>(
        eng, act.m,
        [=] __device__(uint64_t t) -> IdxT {
          const uint64_t k = p[t];
          return (k == 0 || keys[k] != keys[k - 1]) ? static_cast<IdxT>(pos_base + k) : IdxT(0);
        },
        [=] __device__(uint64_t t, IdxT head) { g[t] = head; });
  }

  uint64_t total_active = ranks.global_sum(act.m);
  const unsigned rank_bits = round_up8(bit_length(n - 1));
  uint64_t h = key_bits >> log2_bits;  // the key sort ordered the suffixes by that many symbols
  const bool trace = trace_enabled();
  std::chrono::steady_clock::time_point round_start;
  if (trace) {
    std::fprintf(stderr, "[capsb] refine: count=%llu active=%llu h0=%llu\n", (unsigned long long)count,
                 (unsigned long long)act.m, (unsigned long long)h);
    trace_point(eng, "tied set built, key LCPs written");
  }
  auto lap = [&](const char* what, uint64_t before) {
    if (!trace) return;
    CAPSB_CUDA(cudaStreamSynchronize(st));
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - round_start).count();
    std::fprintf(stderr, "[capsb] %s: h=%llu active %llu -> %llu  %.3f ms\n", what, (unsigned long long)h,
                 (unsigned long long)before, (unsigned long long)total_active, ms);
  };

  // TEXT rounds order every group by the text that follows the current depth (the next 63 bits of
  // the packed text, read directly): they need no rank array, so in the sharded path they involve
  // no exchange at all, and they clear the shallow ties — repeat families, accidental key
  // collisions — which are nearly all of them.  They stop when a round no longer thins the list:
  // what is left needs the depth to double (refine_deep).
  const uint64_t text_step = 63u >> log2_bits;  // symbols a text round advances
  bool try_pairs = true;
  // The first text round runs before any pair-chain step: it separates most pairs anyway (and settles
  // their LCPs), so chaining them first costs more passes over the full list than it saves the round
  // (3.1 Gbp: 182.3 -> 179.1 ms, profiles/r02 s14).  CAPSB_PAIRS_FIRST=1 restores the other order.
  static const bool pairs_first = std::getenv("CAPSB_PAIRS_FIRST") && std::getenv("CAPSB_PAIRS_FIRST")[0] == '1';
  for (unsigned iter = 0; total_active > 0; ++iter) {
    // groups of two are finished directly (order and LCP).  The step is a few passes over the
    // list, so it stops once it no longer thins the list out.
    if (try_pairs && (pairs_first || iter > 0)) {
      if (trace) round_start = std::chrono::steady_clock::now();
      const uint64_t before = total_active;
      resolve_pairs<IdxT>(eng, ranks, act, d_sa, d_lcp, pos_base, h, false);
      total_active = ranks.global_sum(act.m);
      try_pairs = iter < (pairs_first ? 2u : 3u) || (before - total_active) * 16 >= before;
      lap("pair chains", before);
      if (total_active == 0) break;
    }
    eng.stats.refine_rounds++;
    if (trace) round_start = std::chrono::steady_clock::now();
    const uint64_t m = act.m;
    const uint64_t before = total_active;
    const IdxT* idx = act.idx.get();
    // comp = 1 . (next 63 bits of text) for suffixes that reach depth h, else 0 . (n - 1 - i):
    // the shorter suffix first, i.e. the larger position first
    const PackedText text = pt;
    const uint64_t depth = h;
    DevBuf<uint64_t> comp(m, st);
    uint64_t* c = comp.get();
    launch_map(eng.dev, st, m, [=] __device__(uint64_t t) {
      const uint64_t i = idx[t];
      const uint64_t ih = i + depth;
      c[t] = ih < n ? (1ull << 63) | (text.window(ih) >> 1) : (n - 1 - i);
    });
    refine_round<IdxT, uint64_t, false, Ranks>(eng, act, d_sa, pos_base, std::move(comp), 64, rank_bits, nullptr,
                                               TextRoundLcp<IdxT>{d_lcp, h, n, log2_bits});
    h += text_step;
    total_active = ranks.global_sum(act.m);
    lap("text round", before);
    // a text round that leaves more than 3/4 of the list tied: the rest is deep
    if (total_active > 0 && total_active * 4 > before * 3) break;
  }
  state.h = h;
  state.h_resolved = h;
  state.total_active = total_active;
}

// Prefix doubling proper on the suffixes refine_shallow left tied: second key = rank of suffix
// i + h; h doubles, so deep ties cost O(log maxLCP) rounds.  `tied` lists every position of d_sa
// that was ever tied (all ranges): those publish their SA position as rank once, the still-tied
// ones then their group head; never-tied suffixes keep implicit ranks (LocalRanks / ShardedRanks).
// state.act positions index d_sa, state.h is a depth every group agrees on.
template <class IdxT, class Ranks>
void refine_deep(Engine& eng, Ranks& ranks, IdxT* d_sa, IdxT* d_lcp, uint64_t pos_base, uint64_t n,
                 const TiedSet<IdxT>& tied, RefineState<IdxT>& state) {
  using Comp = typename IdxTraits<IdxT>::Comp;
  constexpr unsigned kField = IdxTraits<IdxT>::kField;
  cudaStream_t st = eng.stream;
  ActiveList<IdxT>& act = state.act;
  uint64_t h = state.h;
  uint64_t total_active = state.total_active;
  if (total_active == 0) return;
  const unsigned rank_bits = round_up8(bit_length(n - 1));
  const bool trace = trace_enabled();
  std::chrono::steady_clock::time_point round_start;
  auto lap = [&](const char* what, uint64_t before) {
    if (!trace) return;
    CAPSB_CUDA(cudaStreamSynchronize(st));
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - round_start).count();
    std::fprintf(stderr, "[capsb] %s: h=%llu active %llu -> %llu  %.3f ms\n", what, (unsigned long long)h,
                 (unsigned long long)before, (unsigned long long)total_active, ms);
  };
  if (trace) round_start = std::chrono::steady_clock::now();
  ranks.reset();
  ranks.resolved_depth = state.h_resolved;
  if (Ranks::kPublishesAllTied) {  // every ever-tied suffix publishes its SA position; the still-tied ones then their group head
    const IdxT* p0 = tied.pos.get();
    DevBuf<IdxT> all_idx(tied.m, st), all_pos(tied.m, st);
    IdxT* ai = all_idx.get();
    IdxT* ap = all_pos.get();
    launch_map(eng.dev, st, tied.m, [=] __device__(uint64_t t) {
      ai[t] = d_sa[p0[t]];
      ap[t] = static_cast<IdxT>(pos_base + p0[t]);
    });
    ranks.publish(ai, ap, tied.m);
  }
  ranks.publish(act.idx.get(), act.group.get(), act.m);
  lap("ranks published (switch to doubling)", total_active);

  bool try_pairs = true;
  for (unsigned iter = 0; total_active > 0; ++iter) {
    if (try_pairs) {
      if (trace) round_start = std::chrono::steady_clock::now();
      const uint64_t before = total_active;
      resolve_pairs<IdxT>(eng, ranks, act, d_sa, d_lcp, pos_base, h, true);
      total_active = ranks.global_sum(act.m);
      try_pairs = iter < 2 || (before - total_active) * 16 >= before;
      lap("pair chains", before);
      if (total_active == 0) break;
    }
    eng.stats.refine_rounds++;
    if (trace) round_start = std::chrono::steady_clock::now();
    const uint64_t m = act.m;
    const uint64_t before = total_active;
    const IdxT* idx = act.idx.get();
    DevBuf<IdxT> second(m, st);
    ranks.second_ranks(idx, m, h, second.get());
    DevBuf<Comp> comp(m, st);
    {
      const IdxT* sec = second.get();
      const IdxT* group = act.group.get();
      Comp* c = comp.get();
      launch_map(eng.dev, st, m, [=] __device__(uint64_t t) {
        const bool inside = static_cast<uint64_t>(idx[t]) + h < n;
        c[t] = (static_cast<Comp>(static_cast<uint64_t>(group[t]) + (inside ? 1u : 0u)) << kField) |
               static_cast<Comp>(sec[t]);
      });
    }
    second.release();
    refine_round<IdxT, Comp, true, Ranks>(eng, act, d_sa, pos_base, std::move(comp), rank_bits, rank_bits, &ranks);
    if (h > (~0ull >> 2)) fail("internal: refinement did not converge");
    h <<= 1;
    total_active = ranks.global_sum(act.m);
    lap("rank round", before);
  }
  state.h = h;
  state.total_active = 0;
}

// Key-derived LCPs at the two edges of every key group and the shorter-suffix bound on the entries
// the text rounds wrote inside them, for the tied positions pos[0, m) of d_sa[0, count) (valid
// once the members at those positions are final; running it again later is harmless).
template <class IdxT>
void fix_group_edges(Engine& eng, const IdxT* pos, uint64_t m, const uint64_t* keys, const IdxT* d_sa, IdxT* d_lcp,
                     uint64_t count, uint64_t n, unsigned log2_bits) {
  launch_map(eng.dev, eng.stream, m, [=] __device__(uint64_t t) {
    const uint64_t k = pos[t];
    if (k > 0) {
      if (keys[k] != keys[k - 1]) {
        d_lcp[k] = key_lcp_value<IdxT>(keys[k - 1], keys[k], d_sa[k - 1], d_sa[k], n, log2_bits);
      } else {  // inside a key group: a text round's entry is still to be bounded by the shorter suffix
        const IdxT v = d_lcp[k];
        const uint64_t a = d_sa[k - 1], b = d_sa[k];
        const uint64_t shorter = n - (a > b ? a : b);
        if (v != kLcpUnset<IdxT> && static_cast<uint64_t>(v) > shorter) d_lcp[k] = static_cast<IdxT>(shorter);
      }
    }
    if (k + 1 < count && keys[k + 1] != keys[k])
      d_lcp[k + 1] = key_lcp_value<IdxT>(keys[k], keys[k + 1], d_sa[k], d_sa[k + 1], n, log2_bits);
  });
}

// ---------------------------------------------------------------------------------------
// Refinement of the ties the key sort leaves.  d_sa[0..count) holds suffixes in key order (SA
// positions pos_base .. pos_base + count of the final array); keys[] are their (masked) keys,
// which order them by their first h0 symbols.  Groups of equal keys must be complete inside
// [0, count).  On return d_sa is in suffix order, d_lcp holds every entry except those of
// neighbours that stayed tied into the rank rounds (kLcpUnset: the permuted-LCP stage of the
// caller, collect_deep_pairs + plcp_for_pairs), and `tied` lists the positions that were tied.
// ---------------------------------------------------------------------------------------
template <class IdxT, class Ranks>
void refine_tied_groups(Engine& eng, Ranks& ranks, const PackedText& pt, unsigned key_bits, const uint64_t* keys,
                        IdxT* d_sa, IdxT* d_lcp, uint64_t count, uint64_t pos_base, uint64_t n, TiedSet<IdxT>& tied) {
  RefineState<IdxT> state;
  refine_shallow<IdxT>(eng, ranks, pt, key_bits, keys, d_sa, d_lcp, count, pos_base, n, tied, state);
  refine_deep<IdxT>(eng, ranks, d_sa, d_lcp, pos_base, n, tied, state);
  trace_point(eng, "ties resolved");
  fix_group_edges<IdxT>(eng, tied.pos.get(), tied.m, keys, d_sa, d_lcp, count, n, pt.log2_bits);
}

// LCP of local position 0: against (prev_key, prev_idx), the last suffix of the bucket before
// this one (its key differs), or 0 at the very start of the suffix array.
template <class IdxT>
void first_position_lcp(Engine& eng, const uint64_t* keys, const IdxT* d_sa, IdxT* d_lcp, uint64_t count, uint64_t n,
                        unsigned log2_bits, bool has_prev, uint64_t prev_key, uint64_t prev_idx) {
  if (count == 0) return;
  launch_map(eng.dev, eng.stream, 1, [=] __device__(uint64_t) {
    d_lcp[0] = has_prev ? key_lcp_value<IdxT>(prev_key, keys[0], prev_idx, d_sa[0], n, log2_bits) : IdxT(0);
  });
}

// The tied neighbours whose LCP is still unset after the refinement (everything the pair-chain
// step did not settle): pair_i = SA[k], pair_j = SA[k-1], pair_k = k.  Returns their number.
template <class IdxT>
uint64_t collect_deep_pairs(Engine& eng, const TiedSet<IdxT>& tied, const uint64_t* keys, const IdxT* d_sa,
                            const IdxT* d_lcp, DevBuf<IdxT>& pair_i, DevBuf<IdxT>& pair_j, DevBuf<IdxT>& pair_k) {
  const IdxT* p0 = tied.pos.get();
  auto deep = [=] __device__(uint64_t t) -> IdxT {
    const uint64_t k = p0[t];
    const uint64_t a = keys[k > 0 ? k - 1 : 0], b = keys[k];  // unconditional loads (see in_group)
    const IdxT l = d_lcp[k];
    return ((k > 0) & (a == b) & (l == kLcpUnset<IdxT>)) ? IdxT(1) : IdxT(0);
  };
  const uint64_t m = scan_total<IdxT, OpSum>(eng, tied.m, deep);
  pair_i.alloc(m, eng.stream);
  pair_j.alloc(m, eng.stream);
  pair_k.alloc(m, eng.stream);
  IdxT* pi = pair_i.get();
  IdxT* pj = pair_j.get();
  IdxT* pk = pair_k.get();
  select_finish<IdxT>(eng, tied.m, deep, [=] __device__(uint64_t t, IdxT slot) {
    const uint64_t k = p0[t];
    pi[slot] = d_sa[k];
    pj[slot] = d_sa[k - 1];
    pk[slot] = static_cast<IdxT>(k);
  });
  return m;
}

// Block-wide comparison for the few very long common prefixes (one CTA per pair).
template <class IdxT, class PosJ>
__global__ void __launch_bounds__(256) long_lcp_kernel(PackedText pt, const IdxT* __restrict__ pos_i, PosJ pos_j,
                                                       const IdxT* __restrict__ todo, uint64_t todo_count,
                                                       IdxT* __restrict__ plcp) {
  __shared__ unsigned long long best;
  constexpr int kPerThread = 4;
  const unsigned spw = pt.syms_per_word();
  for (uint64_t e = blockIdx.x; e < todo_count; e += gridDim.x) {
    const uint64_t t = todo[e];
    const uint64_t i = pos_i[t], j = pos_j(t);
    const uint64_t shorter = pt.n - (i > j ? i : j);
    uint64_t base = plcp[t];  // symbols already known equal (multiple of spw)
    while (true) {
      if (threadIdx.x == 0) best = ~0ull;
      __syncthreads();
      unsigned long long mine = ~0ull;
#pragma unroll
      for (int q = 0; q < kPerThread; ++q) {
        const uint64_t off = base + (static_cast<uint64_t>(q) * 256 + threadIdx.x) * spw;
        if (off >= shorter) {
          if (shorter < mine) mine = shorter;
        } else {
          const uint64_t x = pt.window(i + off) ^ pt.window(j + off);
          if (x != 0) {
            const uint64_t l = off + (static_cast<uint64_t>(__clzll(static_cast<long long>(x))) >> pt.log2_bits);
            if (l < mine) mine = l;
          }
        }
      }
      if (mine != ~0ull) atomicMin(&best, mine);
      __syncthreads();
      const unsigned long long got = best;
      __syncthreads();
      if (got != ~0ull) {
        if (threadIdx.x == 0) plcp[t] = static_cast<IdxT>(got < shorter ? got : shorter);
        break;
      }
      base += static_cast<uint64_t>(kPerThread) * 256 * spw;
    }
  }
}

// LCPs of m suffix pairs (i_t, j_t), given in increasing order of i_t (pos_i) with j_t =
// pos_j(t) the suffix that precedes i_t in the suffix array.  Pair t is reducible when the
// pair (i_t - 1, j_t - 1) is pair t-1 of the list and the preceding symbols agree; then
// LCP_t = LCP_{t-1} - 1 (Karkkainen-Manzini-Puglisi).  Irreducible pairs are compared
// directly: 16 packed words per thread, then one CTA per pair for the rare long ones.
// `known` symbols are known to agree for every pair (when both suffixes are that long).
// out(t, lcp, head, head_lcp) is called once per pair; head is the list index of the irreducible
// pair its chain starts at and head_lcp that pair's LCP.
template <class IdxT, class PosJ, class Out>
void plcp_for_pairs(Engine& eng, const PackedText& pt, const IdxT* pos_i, PosJ pos_j, uint64_t m, uint64_t known,
                    Out out) {
  if (m == 0) return;
  cudaStream_t st = eng.stream;
  const DeviceInfo& dev = eng.dev;
  DevBuf<IdxT> plcp(m, st), todo(m, st), chain_head(m, st);
  DevBuf<unsigned long long> counters(2, st);
  CAPSB_CUDA(cudaMemsetAsync(counters.get(), 0, 2 * sizeof(unsigned long long), st));
  {
    IdxT* pl = plcp.get();
    IdxT* td = todo.get();
    IdxT* ch = chain_head.get();
    unsigned long long* cnt = counters.get();
    launch_map(dev, st, m, [=] __device__(uint64_t t) {
      const uint64_t i = pos_i[t];
      const uint64_t j = pos_j(t);
      // reducible: the pair (i-1, j-1) precedes it in the list (then it IS the list's previous
      // pair) and the preceding symbols agree
      const bool chained =
          t > 0 && static_cast<uint64_t>(pos_i[t - 1]) + 1 == i && i > 0 && j > 0 && pt.symbol(i - 1) == pt.symbol(j - 1);
      ch[t] = chained ? IdxT(0) : static_cast<IdxT>(t);
      // statistics: one atomic per warp, not per pair (every lane of a warp adds to one address)
      const unsigned active = __activemask();
      const unsigned heads = __ballot_sync(active, !chained);
      if (heads != 0 && (threadIdx.x & 31u) == static_cast<unsigned>(__ffs(static_cast<int>(heads)) - 1))
        atomicAdd(cnt + 0, static_cast<unsigned long long>(__popc(heads)));
      if (!chained) {
        uint64_t l = 0;
        const bool done = pt.common_prefix(i, j, known, 16, &l);
        pl[t] = static_cast<IdxT>(l);
        if (!done) td[atomicAdd(cnt + 1, 1ull)] = static_cast<IdxT>(t);
      }
    });
  }
  unsigned long long h_cnt[2];
  read_back(st, h_cnt, counters.get(), sizeof(h_cnt));
  eng.stats.deep_lcp_direct += h_cnt[0];
  eng.stats.deep_lcp_long += h_cnt[1];
  if (h_cnt[1] > 0) {
    const unsigned grid =
        static_cast<unsigned>(std::min<uint64_t>(h_cnt[1], static_cast<uint64_t>(dev.sm_count) * 8));
    CAPSB_LAUNCH((long_lcp_kernel<IdxT, PosJ>), grid, 256, 0, st, pt, pos_i, pos_j, todo.get(),
                 static_cast<uint64_t>(h_cnt[1]), plcp.get());
  }
  {
    const IdxT* pl = plcp.get();
    const IdxT* ch = chain_head.get();
    scan_full<IdxT, OpMax, true>(
        eng, m, [=] __device__(uint64_t t) -> IdxT { return ch[t]; },
        [=] __device__(uint64_t t, IdxT head) {
          const uint64_t back = static_cast<uint64_t>(pos_i[t]) - static_cast<uint64_t>(pos_i[head]);
          out(t, static_cast<IdxT>(static_cast<uint64_t>(pl[head]) - back), static_cast<uint64_t>(head), pl[head]);
        });
  }
}

}  // namespace capsb
