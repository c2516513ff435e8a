// Key sort of suffixes, most significant digit first, on packed 8-byte records — the sort the
// single-GPU path and every rank's bucket sort use for 32-bit indices (replaces permute +
// sort_subarrays / merge_sort, reference src/Suffix_Array.cpp:112-184).
//
// Why not the LSD sort of radix_sort.cuh: its passes carry (u64 key, u32 suffix) pairs, five
// times through HBM for a 40-bit key, and each pass is bound by the SM's shared-memory pipe, not
// by HBM (profiles/r01).  Here a suffix is ONE 64-bit record, (remaining key bits << 32) | suffix:
// digits already consumed are implied by where the record lies, so they are dropped.
//
//   level A   the suffixes are partitioned by the top `a` bits of their key, read straight from
//             the packed text; nothing but the 8-byte records is written;
//   level B   every level-A bucket is partitioned by the next `b` bits: 2^(a+b) buckets of a few
//             thousand records each (a + b is chosen from the number of suffixes);
//   local     one CTA per bucket: the bucket is loaded into shared memory, ordered by the
//             remaining bits (one counting pass on their top bits, then each group of up to a
//             hundred-odd records is ordered by comparison, larger groups by one more counting
//             pass), and leaves as the sorted keys — written in place of the records — plus the
//             suffix array.  Two instantiations: 512 threads, two CTAs per SM, for buckets of up
//             to 6144 records; 1024 threads, one CTA per SM, for those up to 12 288.
// Three trips through HBM instead of five, 16 B per record and trip instead of 24.  A partition
// need not be stable (whatever order equal digits come in, the next level sorts them), so a
// record's rank inside its tile is ONE shared-memory atomicAdd on a per-CTA counter instead of
// the per-warp match masks, leader counters and three table look-ups of the stable LSD pass.
// Buckets too large for shared memory (texts dominated by one key: periodic, all-equal) fall
// back to the LSD sort over just those buckets.
#pragma once

#include "packed_text.cuh"
#include "radix_sort.cuh"

namespace capsb {

#ifndef CAPSB_MSD_THREADS  // tools/msd_bench.cu builds other layouts for A/B runs
#define CAPSB_MSD_THREADS 512
#endif
#ifndef CAPSB_MSD_ITEMS
#define CAPSB_MSD_ITEMS 12
#endif
#ifndef CAPSB_MSD_MIN_CTAS
#define CAPSB_MSD_MIN_CTAS 2
#endif
constexpr int kMsdThreads = CAPSB_MSD_THREADS;
constexpr int kMsdItems = CAPSB_MSD_ITEMS;
constexpr int kMsdTile = kMsdThreads * kMsdItems;  // 6144 records = 48 KB
constexpr int kMsdMaxBits = 10;                    // digit width of levels A and B
constexpr int kMsdMaxBins = 1 << kMsdMaxBits;
constexpr int kMsdLocalBits = 12;                  // digit width of the counting passes inside a bucket
constexpr int kMsdLocalBins = 1 << kMsdLocalBits;
constexpr int kMsdLocalCap = kMsdTile;             // largest bucket the two-CTAs-per-SM local sort takes
constexpr int kMsdBigThreads = 1024;               // second instantiation: one CTA per SM, twice the bucket
constexpr unsigned kMsdSmallGroup = 32;            // groups up to AT LEAST this size are ordered by comparison (the limit is a kernel argument)
static_assert(kMsdTile <= 8192 && kMsdThreads % 32 == 0, "ranks are packed in 13 / 16 bits");

// A piece = the part of one parent bucket that one CTA partitions: records [begin, end).
struct MsdPiece {
  uint32_t parent, begin, end, reserved;
};

// ---- sources -----------------------------------------------------------------------------
// load_tile(tile, valid, bins, rec, d): this thread's kMsdItems records of the tile [tile, tile +
// valid) and their digits at this level (d = bins where the thread has no record).  Which records
// a thread takes is the source's choice — a partition need not be stable.
struct MsdRecordSource {
  const uint64_t* in;
  unsigned shift, mask;
  __device__ __forceinline__ unsigned digit_of(uint64_t rec) const { return static_cast<unsigned>(rec >> shift) & mask; }
  __device__ __forceinline__ void load_tile(uint64_t tile, unsigned valid, unsigned bins, uint64_t (&rec)[kMsdItems],
                                            unsigned (&d)[kMsdItems], bool /*counting_only*/ = false) const {
#pragma unroll
    for (int t = 0; t < kMsdItems; ++t) {  // striped: consecutive threads read consecutive records
      const unsigned e = static_cast<unsigned>(t) * kMsdThreads + threadIdx.x;
      rec[t] = e < valid ? ld_stream_u64(in + tile + e) : 0ull;
      d[t] = e < valid ? digit_of(rec[t]) : bins;
    }
  }
};

// The 64-bit text windows of kMsdItems consecutive suffixes from the (at most four) packed words
// w[] they span; the first window starts o0 < 64 bits into w[0].  Window t + 1 is window t shifted
// by one symbol with the next symbol of the text shifted in, so after three runtime shifts
// everything is shifts by constants.
template <int kLog2Bits>
__device__ __forceinline__ void msd_roll_windows(const uint64_t (&w)[4], unsigned o0, uint64_t (&win)[kMsdItems]) {
  constexpr unsigned kBits = 1u << kLog2Bits;
  static_assert(kMsdItems * kBits <= 128, "the symbols shifted in come from two tail words");
  const uint64_t cur0 = o0 ? (w[0] << o0) | (w[1] >> (64u - o0)) : w[0];
  const uint64_t tail0 = o0 ? (w[1] << o0) | (w[2] >> (64u - o0)) : w[1];  // the 64 bits after cur0
  const uint64_t tail1 = o0 ? (w[2] << o0) | (w[3] >> (64u - o0)) : w[2];  // and the 64 after those
  uint64_t cur = cur0;
#pragma unroll
  for (int t = 0; t < kMsdItems; ++t) {
    win[t] = cur;
    constexpr uint64_t kSymMask = (1ull << kBits) - 1ull;
    const unsigned at = static_cast<unsigned>(t) * kBits;  // bit offset of the next symbol in the tail
    const uint64_t tail = at < 64u ? tail0 : tail1;
    const uint64_t sym = (tail >> (64u - kBits - (at & 63u))) & kSymMask;
    cur = (cur << kBits) | sym;
  }
}

// The same when the key bits of all kMsdItems windows lie inside the first one (key_bits +
// (kMsdItems - 1) symbols <= 64: DNA with its 40-bit key): window t is window 0 shifted left by t
// symbols — garbage-free in its key bits, two packed words instead of four, one shift per window.
template <int kLog2Bits>
__device__ __forceinline__ void msd_shift_windows(uint64_t cur0, uint64_t (&win)[kMsdItems]) {
#pragma unroll
  for (int t = 0; t < kMsdItems; ++t) {
    constexpr unsigned kShiftCap = 63u;
    const unsigned sh = static_cast<unsigned>(t) << kLog2Bits;
    win[t] = cur0 << (sh < kShiftCap ? sh : kShiftCap);  // (sh < 64 whenever the caller's condition holds)
  }
}

template <class First, class = void>
struct MsdIsSequentialText : std::false_type {};
template <class First>
struct MsdIsSequentialText<First, std::enable_if_t<First::kSequentialText>> : std::true_type {};

// Level A: the key comes from the first-pass source of the LSD sort (TextSource: window of the
// packed text at suffix base + pos; SuffixListSource: window at suffix idx[pos]); the digit is
// its top bits and does not travel with the record.
template <class First>
struct MsdFirstSource {
  First first;
  unsigned key_shift;  // 64 - key_bits
  unsigned rem_bits;   // key_bits - a
  unsigned short_windows = 3;  // bit 0: the counting pass, bit 1: the scatter pass may take msd_shift_windows
  __device__ __forceinline__ unsigned digit_of(uint64_t) const { return 0; }  // not recoverable: kDigitInRec = false
  __device__ __forceinline__ void split(uint64_t key, uint64_t suffix, uint64_t& rec, unsigned& d) const {
    const uint64_t k = key >> key_shift;
    d = static_cast<unsigned>(k >> rem_bits);
    rec = ((k & ((1ull << rem_bits) - 1ull)) << 32) | suffix;
  }
  __device__ __forceinline__ void load_tile(uint64_t tile, unsigned valid, unsigned bins, uint64_t (&rec)[kMsdItems],
                                            unsigned (&d)[kMsdItems], bool counting_only = false) const {
    if constexpr (MsdIsSequentialText<First>::value) {
      // Consecutive suffixes of the text: the thread takes kMsdItems consecutive ones, loads the
      // (at most four) packed words they span ONCE and shifts every window out of them — one
      // wait on memory per tile instead of one per suffix.
      const PackedText& pt = first.pt;
      const unsigned e0 = threadIdx.x * kMsdItems;
      const uint64_t p0 = first.base + tile + e0;
      const unsigned bits = pt.bits();
      const uint64_t bit0 = p0 << pt.log2_bits;
      const uint64_t w0 = bit0 >> 6;
      const uint64_t nwords = (((pt.n << pt.log2_bits) + 63) >> 6) + 2;
      const unsigned o0 = static_cast<unsigned>(bit0 & 63u);
      uint64_t win[kMsdItems];
      const bool key_bits_in_first_window =
          ((static_cast<unsigned>(kMsdItems) - 1u) << pt.log2_bits) <= key_shift && ((short_windows >> (counting_only ? 0 : 1)) & 1u);
      if (key_bits_in_first_window) {  // (uniform over the grid)
        uint64_t w[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) w[q] = (e0 < valid && w0 + q < nwords) ? __ldg(pt.words + w0 + q) : 0ull;
        const uint64_t cur0 = o0 ? (w[0] << o0) | (w[1] >> (64u - o0)) : w[0];
        switch (pt.log2_bits) {  // constant shifts per symbol width
          case 0: msd_shift_windows<0>(cur0, win); break;
          case 1: msd_shift_windows<1>(cur0, win); break;
          case 2: msd_shift_windows<2>(cur0, win); break;
          default: msd_shift_windows<3>(cur0, win); break;
        }
      } else {
        uint64_t w[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) w[q] = (e0 < valid && w0 + q < nwords) ? __ldg(pt.words + w0 + q) : 0ull;
        switch (pt.log2_bits) {  // constant shifts per symbol width
          case 0: msd_roll_windows<0>(w, o0, win); break;
          case 1: msd_roll_windows<1>(w, o0, win); break;
          case 2: msd_roll_windows<2>(w, o0, win); break;
          default: msd_roll_windows<3>(w, o0, win); break;
        }
      }
#pragma unroll
      for (int t = 0; t < kMsdItems; ++t) {
        rec[t] = 0;
        d[t] = bins;
        if (e0 + t < valid) split(win[t] & first.mask, p0 + t, rec[t], d[t]);
      }
    } else {
#pragma unroll
      for (int t = 0; t < kMsdItems; ++t) {
        const unsigned e = static_cast<unsigned>(t) * kMsdThreads + threadIdx.x;
        rec[t] = 0;
        d[t] = bins;
        if (e < valid) split(first.key(tile + e), static_cast<uint64_t>(first.val(tile + e)), rec[t], d[t]);
      }
    }
  }
};

// ---- block-wide exclusive sum ------------------------------------------------------------
// One value per thread; warp_tot[32] is scratch (the caller separates consecutive uses by a
// barrier).  Contains one __syncthreads.
template <int kThreads>
__device__ __forceinline__ unsigned msd_block_excl_scan(unsigned v, unsigned* warp_tot, unsigned* total = nullptr) {
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  unsigned inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned o = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= static_cast<unsigned>(d)) inc += o;
  }
  if (lane == 31u) warp_tot[warp] = inc;
  __syncthreads();
  unsigned prefix = 0, all = 0;
#pragma unroll
  for (int w = 0; w < kThreads / 32; ++w) {
    const unsigned t = warp_tot[w];
    if (static_cast<unsigned>(w) < warp) prefix += t;
    all += t;
  }
  if (total) *total = all;
  return prefix + inc - v;
}

// Rank of a record among the records of its digit counted so far in this tile: normally one
// shared-memory atomicAdd.  When a probe row finds the warp's digits heavily repeated (texts
// dominated by one symbol), the lanes that share a digit elect a leader that adds their number
// once — same-address atomics of one warp instruction would serialise.  d == bins marks a lane
// without a record (its count goes to the spare counter).  Warp-uniform `aggregate`.
__device__ __forceinline__ unsigned msd_count(unsigned* cnt, unsigned d, bool has, bool aggregate, unsigned lane,
                                              unsigned lt) {
  if (!aggregate) return has ? atomicAdd(&cnt[d], 1u) : 0u;
  const unsigned peers = __match_any_sync(0xffffffffu, d);
  const int leader = __ffs(static_cast<int>(peers)) - 1;
  unsigned base = 0;
  if (static_cast<int>(lane) == leader) base = atomicAdd(&cnt[d], static_cast<unsigned>(__popc(peers)));
  base = __shfl_sync(0xffffffffu, base, leader);
  return base + static_cast<unsigned>(__popc(peers & lt));
}
__device__ __forceinline__ bool msd_probe(unsigned d, bool has) {
  const unsigned peers = __match_any_sync(0xffffffffu, has ? d : 0xffffffffu);
  return __any_sync(0xffffffffu, has && __popc(peers) >= 8);
}

// ---- planning: parents -> pieces ---------------------------------------------------------
// One CTA.  Parent x (records [parent_start[x], parent_start[x+1])) is cut into
// ceil(size / target) pieces of equal length (a multiple of the tile); piece_first[x] is its
// first piece, piece_first[nparents] = *piece_count the number of pieces.
static __global__ void __launch_bounds__(1024) msd_plan_kernel(const uint32_t* __restrict__ parent_start,
                                                               unsigned nparents, uint32_t target,
                                                               MsdPiece* __restrict__ pieces,
                                                               uint32_t* __restrict__ piece_first,
                                                               uint32_t* __restrict__ piece_count) {
  __shared__ unsigned warp_tot[32];
  const unsigned x = threadIdx.x;
  uint32_t begin = 0, size = 0;
  if (x < nparents) {
    begin = parent_start[x];
    size = parent_start[x + 1] - begin;
  }
  const uint32_t np = size ? (size - 1) / target + 1 : 0;
  unsigned total;
  const unsigned first = msd_block_excl_scan<1024>(np, warp_tot, &total);
  if (x < nparents) piece_first[x] = first;
  if (x == 0) {
    piece_first[nparents] = total;
    *piece_count = total;
  }
  if (np) {
    uint64_t per = (static_cast<uint64_t>(size) + np - 1) / np;
    per = (per + kMsdTile - 1) / kMsdTile * kMsdTile;
    const uint64_t stop = static_cast<uint64_t>(begin) + size;
    for (uint32_t j = 0; j < np; ++j) {
      uint64_t b = begin + j * per, e = b + per;
      if (b > stop) b = stop;
      if (e > stop) e = stop;
      pieces[first + j] = MsdPiece{x, static_cast<uint32_t>(b), static_cast<uint32_t>(e), 0u};
    }
  }
}

// ---- histogram of one piece --------------------------------------------------------------
// gridDim.y CTAs share a piece (its tiles dealt round-robin) and add their counts into hist, which
// the caller has zeroed.
template <class Src>
__global__ void __launch_bounds__(kMsdThreads) msd_hist_kernel(Src src, const MsdPiece* __restrict__ pieces,
                                                               const uint32_t* __restrict__ piece_count, unsigned bins,
                                                               uint32_t* __restrict__ hist) {
  __shared__ unsigned cnt[kMsdMaxBins + 1];
  if (blockIdx.x >= *piece_count) return;
  const MsdPiece pc = pieces[blockIdx.x];
  const unsigned tid = threadIdx.x, lane = tid & 31u;
  for (unsigned b = tid; b <= bins; b += kMsdThreads) cnt[b] = 0;
  __syncthreads();
  for (uint64_t tile = pc.begin + static_cast<uint64_t>(blockIdx.y) * kMsdTile; tile < pc.end;
       tile += static_cast<uint64_t>(gridDim.y) * kMsdTile) {
    const unsigned valid = static_cast<unsigned>(pc.end - tile < kMsdTile ? pc.end - tile : kMsdTile);
    uint64_t rec[kMsdItems];
    unsigned d[kMsdItems];
    src.load_tile(tile, valid, bins, rec, d, true);
    const bool aggregate = msd_probe(d[0], d[0] != bins);
#pragma unroll
    for (int t = 0; t < kMsdItems; ++t) {
      if (!aggregate) {
        if (d[t] != bins) atomicAdd(&cnt[d[t]], 1u);
      } else {
        const unsigned peers = __match_any_sync(0xffffffffu, d[t]);
        if (static_cast<int>(lane) == __ffs(static_cast<int>(peers)) - 1)
          atomicAdd(&cnt[d[t]], static_cast<unsigned>(__popc(peers)));
      }
    }
  }
  __syncthreads();
  for (unsigned b = tid; b < bins; b += kMsdThreads)
    if (cnt[b]) atomicAdd(&hist[static_cast<uint64_t>(blockIdx.x) * bins + b], cnt[b]);
}

// ---- offsets -----------------------------------------------------------------------------
// One CTA per parent, one thread per bin.  In: hist[piece][bin] = counts.  Out, in place:
// hist[piece][bin] = where the piece's records of that bin go; child_start[parent * bins + bin]
// = start of the child bucket (the children of a parent tile it in bin order), and one closing
// entry child_start[nparents * bins] = end of the last parent.
static __global__ void __launch_bounds__(kMsdMaxBins) msd_offsets_kernel(const uint32_t* __restrict__ parent_start,
                                                                         unsigned nparents,
                                                                         const uint32_t* __restrict__ piece_first,
                                                                         unsigned bins, uint32_t* __restrict__ hist,
                                                                         uint32_t* __restrict__ child_start) {
  __shared__ unsigned warp_tot[32];
  const unsigned x = blockIdx.x, b = threadIdx.x;
  const uint32_t p0 = piece_first[x], p1 = piece_first[x + 1];
  const bool live = b < bins;
  uint32_t tot = 0;
  if (live) {
    uint32_t p = p0;
    for (; p + 8 <= p1; p += 8) {  // eight independent loads in flight
      uint32_t v[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = hist[static_cast<uint64_t>(p + q) * bins + b];
#pragma unroll
      for (int q = 0; q < 8; ++q) tot += v[q];
    }
    for (; p < p1; ++p) tot += hist[static_cast<uint64_t>(p) * bins + b];
  }
  const unsigned excl = msd_block_excl_scan<kMsdMaxBins>(tot, warp_tot);
  if (!live) return;
  const uint32_t base = parent_start[x] + excl;
  child_start[static_cast<uint64_t>(x) * bins + b] = base;
  if (x + 1 == nparents && b + 1 == bins) child_start[static_cast<uint64_t>(nparents) * bins] = parent_start[nparents];
  uint32_t run = base;
  uint32_t p = p0;
  for (; p + 8 <= p1; p += 8) {
    uint32_t v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = hist[static_cast<uint64_t>(p + q) * bins + b];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      hist[static_cast<uint64_t>(p + q) * bins + b] = run;
      run += v[q];
    }
  }
  for (; p < p1; ++p) {
    const uint32_t v = hist[static_cast<uint64_t>(p) * bins + b];
    hist[static_cast<uint64_t>(p) * bins + b] = run;
    run += v;
  }
}

// ---- partition of one piece --------------------------------------------------------------
// Per tile: every thread takes kMsdItems records, ranks each with one shared-memory atomicAdd, the
// threads scan the <= 1024 counters, the records go to their place in the staging area (sorted
// by digit), and consecutive threads write consecutive staged records to the output.
//
// Only whole, aligned 32-byte sectors (four records) leave the SM.  With 1024 digits a tile holds
// about six records per digit; writing such runs as they come leaves a partly written sector at
// both ends of every run, and with a million runs open across the chip L2 evicts them before the
// next tile completes them: 40 % more DRAM writes plus as many fill reads (ncu, profiles/r02).  So
// a digit's records that do not fill a sector yet (at most three) are carried into the next tile,
// where they head the digit's run; a tile takes fewer new records to make room for them.  The
// piece's last tile writes everything.
template <bool kDigitInRec>
struct MsdScatterSmem {
  uint64_t stage[kMsdTile];
  uint64_t carry[kMsdMaxBins * 3];  // the records carried over, three slots per digit
  uint2 shlim[kMsdMaxBins];         // x: output slot of staged slot s is x + s (mod 2^32); y: staged slots below y leave now
  unsigned cnt[kMsdMaxBins + 1];    // [bins] collects the lanes without a record; starts a tile at the digit's carried count
  unsigned start[kMsdMaxBins];      // first staged slot of each digit in this tile
  unsigned warp_tot[32];
  unsigned carried_next;            // records carried into the next tile (accumulated by the digit owners)
  uint16_t sdig[kDigitInRec ? 2 : kMsdTile];  // digit of each staged record when the record does not hold it
};

template <class Src, bool kDigitInRec>
__global__ void __launch_bounds__(kMsdThreads, CAPSB_MSD_MIN_CTAS) msd_scatter_kernel(Src src, const MsdPiece* __restrict__ pieces,
                                                                     const uint32_t* __restrict__ piece_count,
                                                                     unsigned bins, const uint32_t* __restrict__ off,
                                                                     uint64_t* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char msd_smem_raw[];
  MsdScatterSmem<kDigitInRec>& sm = *reinterpret_cast<MsdScatterSmem<kDigitInRec>*>(msd_smem_raw);
  if (blockIdx.x >= *piece_count) return;
  const MsdPiece pc = pieces[blockIdx.x];
  const unsigned tid = threadIdx.x, lane = tid & 31u;
  const unsigned lt = lanemask_lt();
  // this thread owns the digits tid * kPer .. + kPer - 1: their next output slot and carried count
  constexpr int kPer = (kMsdMaxBins + kMsdThreads - 1) / kMsdThreads;
  unsigned gout[kPer], kprev[kPer];
#pragma unroll
  for (int q = 0; q < kPer; ++q) {
    const unsigned b = tid * kPer + q;
    gout[q] = b < bins ? off[static_cast<uint64_t>(blockIdx.x) * bins + b] : 0u;
    kprev[q] = 0;
  }
  for (unsigned b = tid; b <= bins; b += kMsdThreads) sm.cnt[b] = 0;
  if (tid == 0) sm.carried_next = 0;
  __syncthreads();
  unsigned carried = 0;
  for (uint64_t pos = pc.begin; pos < pc.end;) {
    const uint64_t left = pc.end - pos;
    const unsigned room = kMsdTile - carried;
    const unsigned valid = static_cast<unsigned>(left < room ? left : room);
    const bool last = left <= room;
    uint64_t rec[kMsdItems];
    unsigned dr[kMsdItems];  // digit, then digit | rank << 11 (rank < 8192)
    src.load_tile(pos, valid, bins, rec, dr);
    const bool aggregate = msd_probe(dr[0], dr[0] != bins);
#pragma unroll
    for (int t = 0; t < kMsdItems; ++t) {
      const unsigned d = dr[t];
      const unsigned r = msd_count(sm.cnt, d, d != bins, aggregate, lane, lt);
      dr[t] = d | (r << 11);
    }
    __syncthreads();
    {  // digit owners: starts inside the tile, what leaves now, what is carried, output shifts
      unsigned c[kPer], sum = 0;
#pragma unroll
      for (int q = 0; q < kPer; ++q) {
        const unsigned b = tid * kPer + q;
        c[q] = b < bins ? sm.cnt[b] : 0u;
        sum += c[q];
      }
      unsigned run = msd_block_excl_scan<kMsdThreads>(sum, sm.warp_tot);
      unsigned knew[kPer], ksum = 0;
#pragma unroll
      for (int q = 0; q < kPer; ++q) {
        const unsigned b = tid * kPer + q;
        knew[q] = 0;
        if (b < bins) {
          const unsigned g = gout[q], e = g + c[q];
          const unsigned aligned = e & ~3u;
          // records beyond the last sector boundary wait (all of them if the run has not reached one)
          const unsigned k = last ? 0u : e - (aligned > g ? aligned : g);
          const unsigned f = c[q] - k;
          sm.start[b] = run;
          sm.shlim[b] = make_uint2(g - run, run + f);
          gout[q] = g + f;
          sm.cnt[b] = k;
          knew[q] = k;
          ksum += k;
        }
        run += c[q];
      }
      if (tid == 0) sm.cnt[bins] = 0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ksum += __shfl_xor_sync(0xffffffffu, ksum, o);
      if (lane == 0 && ksum) atomicAdd(&sm.carried_next, ksum);
      __syncthreads();
      // the records carried in head their digit's run
#pragma unroll
      for (int q = 0; q < kPer; ++q) {
        const unsigned b = tid * kPer + q;
        for (unsigned j = 0; j < kprev[q]; ++j) {
          const unsigned slot = sm.start[b] + j;
          sm.stage[slot] = sm.carry[b * 3 + j];
          if (!kDigitInRec) sm.sdig[slot] = static_cast<uint16_t>(b);
        }
        kprev[q] = knew[q];
      }
    }
#pragma unroll
    for (int t = 0; t < kMsdItems; ++t) {
      const unsigned d = dr[t] & 0x7FFu;
      if (d != bins) {
        const unsigned slot = sm.start[d] + (dr[t] >> 11);
        sm.stage[slot] = rec[t];
        if (!kDigitInRec) sm.sdig[slot] = static_cast<uint16_t>(d);
      }
    }
    const unsigned total = valid + carried;
    carried = sm.carried_next;
    __syncthreads();
    if (tid == 0) sm.carried_next = 0;
#pragma unroll
    for (int t = 0; t < kMsdItems; ++t) {
      const unsigned s = static_cast<unsigned>(t) * kMsdThreads + tid;
      if (s < total) {
        const uint64_t r = sm.stage[s];
        const unsigned d = kDigitInRec ? src.digit_of(r) : static_cast<unsigned>(sm.sdig[s]);
        const uint2 sl = sm.shlim[d];
        if (s < sl.y)
          out[static_cast<uint32_t>(sl.x + s)] = r;
        else
          sm.carry[d * 3 + (s - sl.y)] = r;
      }
    }
    pos += valid;
    // no barrier here: the next tile touches the staging area, the tables and the carry slots only
    // after its own first barrier, which every thread reaches after it has finished this loop
  }
}

// ---- local sort: one CTA per bucket ------------------------------------------------------
// Bucket q = records [child_start[q], child_start[q+1]) of `recs`, all with the same leading
// prefix_bits = a + b key bits (= q); the remaining rem_bits sit at bits [32, 32 + rem_bits) of
// the record.  On return the same slots hold the sorted keys (key_bits bits, left-aligned, the
// form the rest of the pipeline uses) and sa_out the suffixes.  Buckets larger than the shared
// memory staging area are left alone and appended to large_list.
// Counter b lives at word b + (b >> 5): the scan's threads own eight consecutive counters each, and
// the extra word per 32 spreads those stride-8 accesses over all banks (plain indexing: 8-way
// conflicts, a third of the kernel's shared-memory traffic — ncu, profiles/r02).
__device__ __forceinline__ unsigned msd_pad(unsigned b) { return b + (b >> 5); }
constexpr int kMsdLocalPadded = kMsdLocalBins + kMsdLocalBins / 32 + 2;

template <int kCap>
struct MsdLocalSmem {
  uint64_t stage[kCap];
  unsigned cnt[kMsdLocalPadded];
  unsigned start[kMsdLocalPadded];
  unsigned warp_tot[32];
  unsigned big[kCap / (kMsdSmallGroup + 1) + 1];  // (start, size - 1) of the groups too large for comparison ordering
  unsigned nbig;
};

// Counting pass over records held in registers (rec[t] for the elements e = t * threads + tid <
// count of the range [base, base + count) of the staging area, count <= kN * threads): ranks, scan
// of `bins` counters (bins <= 4096), records written back sorted by the digit at `shift`.
// start[] holds the digit starts (relative to base) afterwards, start[bins] = count.  All threads
// of the CTA call it.
template <int kN, int kT, class Smem>
__device__ __forceinline__ void msd_local_pass(Smem& sm, uint64_t (&rec)[kN], unsigned base, unsigned count,
                                               unsigned shift, unsigned bits) {
  const unsigned tid = threadIdx.x, lane = tid & 31u;
  const unsigned lt = lanemask_lt();
  const unsigned bins = 1u << bits, mask = bins - 1u;
  unsigned rk[(kN + 1) / 2];  // two 16-bit ranks per register; the digit is recomputed from the record
  const bool has0 = tid < count;
  const bool aggregate = msd_probe(has0 ? (static_cast<unsigned>(rec[0] >> shift) & mask) : bins, has0);
  if (!aggregate) {
#pragma unroll
    for (int t = 0; t < kN; ++t) {
      const unsigned e = static_cast<unsigned>(t) * kT + tid;
      unsigned r = 0;
      if (e < count) r = atomicAdd(&sm.cnt[msd_pad(static_cast<unsigned>(rec[t] >> shift) & mask)], 1u);
      if (t & 1)
        rk[t >> 1] |= r << 16;
      else
        rk[t >> 1] = r;
    }
  } else {
#pragma unroll
    for (int t = 0; t < kN; ++t) {
      const unsigned e = static_cast<unsigned>(t) * kT + tid;
      const unsigned d = e < count ? (static_cast<unsigned>(rec[t] >> shift) & mask) : bins;
      const unsigned r = msd_count(sm.cnt, msd_pad(d), d != bins, true, lane, lt);
      if (t & 1)
        rk[t >> 1] |= r << 16;
      else
        rk[t >> 1] = r;
    }
  }
  __syncthreads();
  {  // eight consecutive counters per thread
    constexpr int kPer = (kMsdLocalBins + kT - 1) / kT;
    unsigned c[kPer], sum = 0;
#pragma unroll
    for (int q = 0; q < kPer; ++q) {
      const unsigned b = tid * kPer + q;
      c[q] = b < bins ? sm.cnt[msd_pad(b)] : 0u;
      sum += c[q];
    }
    unsigned run = msd_block_excl_scan<kT>(sum, sm.warp_tot);
#pragma unroll
    for (int q = 0; q < kPer; ++q) {
      const unsigned b = tid * kPer + q;
      if (b < bins) {
        sm.start[msd_pad(b)] = run;
        sm.cnt[msd_pad(b)] = 0;
      }
      run += c[q];
    }
    if (tid == 0) {
      sm.start[msd_pad(bins)] = count;
      sm.cnt[msd_pad(bins)] = 0;
    }
  }
  __syncthreads();
#pragma unroll
  for (int t = 0; t < kN; ++t) {
    const unsigned e = static_cast<unsigned>(t) * kT + tid;
    if (e < count) {
      const unsigned d = static_cast<unsigned>(rec[t] >> shift) & mask;
      sm.stage[base + sm.start[msd_pad(d)] + ((rk[t >> 1] >> ((t & 1) * 16)) & 0xFFFFu)] = rec[t];
    }
  }
  __syncthreads();
}

// One bucket of at most kN * threads records (the kernel picks kN by the bucket's size: most
// buckets of a multi-Gbp text hold about 3000 records, half of what the staging area takes, and
// the unrolled per-record code of the full size would be half idle on them).
template <int kN, int kT, class Smem>
__device__ __forceinline__ void msd_local_bucket(Smem& sm, uint64_t* __restrict__ recs,
                                                 uint32_t* __restrict__ sa_out, uint32_t q, uint32_t beg, uint32_t count,
                                                 unsigned rem_bits, unsigned key_shift, unsigned extra_bits,
                                                 unsigned small_group, uint32_t next_beg, uint32_t next_count) {
  const unsigned tid = threadIdx.x;
  // The first counting pass takes the top hb of the remaining bits: about one counter per record
  // (extra_bits = 1; per two records with 0 — the scan of the counters is per-bucket overhead, the
  // comparison loop of the groups grows with the records per counter), at least rem_bits - 12 so
  // that one more pass can finish a large group, at most 12.
  unsigned hb = (count > 4 ? 31u - static_cast<unsigned>(__clz(count - 1u)) : 1u) + extra_bits;  // 2^(hb+1-extra) >= count
  if (hb > static_cast<unsigned>(kMsdLocalBits)) hb = kMsdLocalBits;
  if (hb + kMsdLocalBits < rem_bits) hb = rem_bits - kMsdLocalBits;
  if (hb > rem_bits) hb = rem_bits;
  const unsigned lb = rem_bits - hb;
  // the key bits every record of this bucket shares, in place above the remaining ones
  const uint64_t prefix = static_cast<uint64_t>(q) << rem_bits;
  const uint64_t rem_mask = (1ull << rem_bits) - 1ull;
  uint64_t rec[kN];
#pragma unroll
  for (int t = 0; t < kN; ++t) {
    const unsigned e = static_cast<unsigned>(t) * kT + tid;
    rec[t] = e < count ? ld_stream_u64(recs + beg + e) : 0ull;
  }
  // While this bucket is ordered, the records of the CTA's next one travel to L2 (one 128-byte line
  // per thread: no registers held, unlike loading them ahead): the loads above are what the kernel
  // waits on longest after its barriers (long-scoreboard stalls, profiles/r02 s24).
  if (next_count > 0) {
    const uintptr_t first_line = reinterpret_cast<uintptr_t>(recs + next_beg) & ~static_cast<uintptr_t>(127);
    const uintptr_t end = reinterpret_cast<uintptr_t>(recs + next_beg + next_count);
    const uintptr_t line = first_line + static_cast<uintptr_t>(tid) * 128u;
    if (line < end) asm volatile("prefetch.global.L2 [%0];" ::"l"(line));
  }
  if (hb > 0 && count > 1) {
    msd_local_pass<kN, kT>(sm, rec, 0, count, 32u + lb, hb);
    if (lb == 0) {  // the pass consumed every remaining bit
      for (unsigned s = tid; s < count; s += kT) {
        const uint64_t r = sm.stage[s];
        recs[beg + s] = (prefix | ((r >> 32) & rem_mask)) << key_shift;
        sa_out[beg + s] = static_cast<uint32_t>(r);
      }
    } else {
      // groups = runs of equal top digits, now contiguous.  A record of a small group counts the
      // records of its group that precede it (by the remaining bits, then by slot) and leaves
      // for its final place at once: key in place of the records, suffix to the suffix array.
      // Large groups are listed and get a counting pass of their own.
      const unsigned hmask = (1u << hb) - 1u;
#pragma unroll 2
      for (unsigned s = tid; s < count; s += kT) {
        const uint64_t r = sm.stage[s];
        const unsigned d = static_cast<unsigned>(r >> (32u + lb)) & hmask;
        const unsigned g0 = sm.start[msd_pad(d)], g1 = sm.start[msd_pad(d + 1)];
        if (g1 - g0 <= small_group) {
          const unsigned mine = static_cast<unsigned>(r >> 32);
          unsigned rank = 0;
          if (g1 - g0 == 2) {  // (most records sit in groups of one or two: no loop for them)
            const unsigned u = s == g0 ? g0 + 1 : g0;
            const unsigned other = static_cast<unsigned>(sm.stage[u] >> 32);
            rank = (other < mine || (other == mine && u < s)) ? 1u : 0u;
          } else if (g1 - g0 > 2) {
            for (unsigned u = g0; u < g1; ++u) {
              const unsigned other = static_cast<unsigned>(sm.stage[u] >> 32);
              rank += (other < mine || (other == mine && u < s)) ? 1u : 0u;
            }
          }
          recs[beg + g0 + rank] = (prefix | (static_cast<uint64_t>(mine) & rem_mask)) << key_shift;
          sa_out[beg + g0 + rank] = static_cast<uint32_t>(r);
        } else if (s == g0) {
          sm.big[atomicAdd(&sm.nbig, 1u)] = g0 | ((g1 - g0 - 1u) << 16);  // start, size - 1 < 65536
        }
      }
      __syncthreads();
      const unsigned nbig = sm.nbig;
      for (unsigned j = 0; j < nbig; ++j) {
        const unsigned g0 = sm.big[j] & 0xFFFFu, gsize = (sm.big[j] >> 16) + 1u;
#pragma unroll
        for (int t = 0; t < kN; ++t) {
          const unsigned e = static_cast<unsigned>(t) * kT + tid;
          rec[t] = e < gsize ? sm.stage[g0 + e] : 0ull;
        }
        __syncthreads();  // every record of the group is in registers before any is written back
        msd_local_pass<kN, kT>(sm, rec, g0, gsize, 32u, lb);
        for (unsigned e = tid; e < gsize; e += kT) {
          const uint64_t r = sm.stage[g0 + e];
          recs[beg + g0 + e] = (prefix | ((r >> 32) & rem_mask)) << key_shift;
          sa_out[beg + g0 + e] = static_cast<uint32_t>(r);
        }
      }
      if (nbig > 0) {
        __syncthreads();
        if (tid == 0) sm.nbig = 0;
      }
    }
    __syncthreads();  // the staging area is refilled by the next bucket
  } else {
#pragma unroll
    for (int t = 0; t < kN; ++t) {
      const unsigned e = static_cast<unsigned>(t) * kT + tid;
      if (e < count) {
        recs[beg + e] = (prefix | ((rec[t] >> 32) & rem_mask)) << key_shift;
        sa_out[beg + e] = static_cast<uint32_t>(rec[t]);
      }
    }
  }
}

// kT threads take buckets of up to kT * kMsdItems records.  Two instantiations: 512 threads, two CTAs
// per SM, over a range of bucket numbers; and 1024 threads, one CTA per SM, over the list of buckets the
// first one found too large (repeat families: 10-mers shared by thousands of copies).  Buckets too
// large for this instantiation are appended to large_list.
template <int kT, int kMinCtas>
__global__ void __launch_bounds__(kT, kMinCtas) msd_local_kernel(uint64_t* __restrict__ recs,
                                                                 const uint32_t* __restrict__ child_start,
                                                                 const uint32_t* __restrict__ bucket_list,
                                                                 uint32_t q_begin, uint32_t q_end, unsigned key_bits,
                                                                 unsigned prefix_bits, unsigned extra_bits,
                                                                 unsigned small_group, bool l2_prefetch,
                                                                 uint32_t* __restrict__ sa_out,
                                                                 uint32_t* __restrict__ large_list,
                                                                 uint32_t* __restrict__ large_count) {
  extern __shared__ __align__(16) unsigned char msd_smem_raw[];
  using Smem = MsdLocalSmem<kT * kMsdItems>;
  Smem& sm = *reinterpret_cast<Smem*>(msd_smem_raw);
  const unsigned tid = threadIdx.x;
  const unsigned rem_bits = key_bits - prefix_bits;  // <= 24
  const unsigned key_shift = 64u - key_bits;
  constexpr int kHalf = (kMsdItems + 1) / 2;
  for (unsigned b = tid; b < static_cast<unsigned>(kMsdLocalPadded); b += kT) sm.cnt[b] = 0;
  if (tid == 0) sm.nbig = 0;
  __syncthreads();
  for (uint32_t at = q_begin + blockIdx.x; at < q_end; at += gridDim.x) {
    const uint32_t q = bucket_list ? bucket_list[at] : at;
    const uint32_t beg = child_start[q];
    const uint32_t count = child_start[q + 1] - beg;
    uint32_t next_beg = 0, next_count = 0;  // the bucket this CTA takes after this one (for the L2 prefetch)
    if (l2_prefetch && at + gridDim.x < q_end) {
      const uint32_t nq = bucket_list ? bucket_list[at + gridDim.x] : at + gridDim.x;
      next_beg = child_start[nq];
      next_count = child_start[nq + 1] - next_beg;
      if (next_count > static_cast<uint32_t>(kT * kMsdItems)) next_count = 0;  // (not this kernel's: it only lists it)
    }
    if (count == 0) continue;
    if (count > static_cast<uint32_t>(kT * kMsdItems)) {
      if (tid == 0) large_list[atomicAdd(large_count, 1u)] = q;
      continue;
    }
    if (count <= static_cast<uint32_t>(kHalf * kT))
      msd_local_bucket<kHalf, kT>(sm, recs, sa_out, q, beg, count, rem_bits, key_shift, extra_bits, small_group, next_beg,
                                  next_count);
    else
      msd_local_bucket<kMsdItems, kT>(sm, recs, sa_out, q, beg, count, rem_bits, key_shift, extra_bits, small_group,
                                      next_beg, next_count);
  }
}

// ---- host side ---------------------------------------------------------------------------
struct MsdPlan {
  unsigned a = 0, b = 0;  // digit widths of levels A and B
};

inline unsigned msd_ceil_log2(uint64_t v) {
  unsigned b = 0;
  while ((1ull << b) < v) ++b;
  return b;
}
inline unsigned round_up8_bits(unsigned b) { return b == 0 ? 8u : (b + 7u) & ~7u; }

// Can the packed-record sort take this job?  The key must leave room for a 32-bit suffix next
// to its remaining bits after level A (a <= 10), and the buckets' remaining bits must fit the two
// counting passes of the local sort.
inline bool msd_applicable(unsigned key_bits) { return key_bits >= 16 && key_bits <= 32u + kMsdMaxBits; }

inline MsdPlan msd_plan(uint64_t count, unsigned key_bits) {
  // about 2048 records per bucket; at least key_bits - 24 leading bits so that the local sort's
  // two passes (12 + 12) cover the rest
  unsigned p = msd_ceil_log2((count + 2047) / 2048);
  if (p < 2) p = 2;
  if (key_bits > 24 && p < key_bits - 24) p = key_bits - 24;
  if (p > 2 * kMsdMaxBits) p = 2 * kMsdMaxBits;
  MsdPlan plan;
  plan.a = (p + 1) / 2;
  if (key_bits > 32 && plan.a < key_bits - 32) plan.a = key_bits - 32;
  plan.b = p > plan.a ? p - plan.a : 1;
  if (plan.b > static_cast<unsigned>(kMsdMaxBits)) plan.b = kMsdMaxBits;
  return plan;
}

// Per-kernel-class CUDA-event timing (bench.py's roofline figures), same scheme as KernelTimer.
struct MsdTimers {
  bool enabled = false;
  KernelTimer scatter_a, scatter_b, local, hist;
  void reset() { scatter_a.reset(), scatter_b.reset(), local.reset(), hist.reset(); }
  void set_enabled(bool on) { enabled = scatter_a.enabled = scatter_b.enabled = local.enabled = hist.enabled = on; }
};

struct MsdTimed {
  KernelTimer& timer;
  cudaStream_t stream;
  cudaEvent_t t0 = nullptr, t1 = nullptr;
  MsdTimed(KernelTimer& t, cudaStream_t st, uint64_t bytes) : timer(t), stream(st) {
    if (!timer.enabled) return;
    t0 = timer.get();
    t1 = timer.get();
    timer.bytes += bytes;
    cudaEventRecord(t0, stream);
  }
  ~MsdTimed() {
    if (!timer.enabled) return;
    cudaEventRecord(t1, stream);
    timer.pending.emplace_back(t0, t1);
  }
};

template <class Kernel>
inline void msd_allow_smem(Kernel kernel, size_t bytes) {
  // per call: the attribute belongs to the current context (one per device), and the call costs
  // a few hundred nanoseconds
  CAPSB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)));
}

// One partition level: the records of `parents` (nparents + 1 starts on the device) are
// partitioned by the digit `src` yields into `bins` children each; child_start gets
// nparents * bins + 1 entries.
template <class Src, bool kDigitInRec>
inline void msd_partition_level(const DeviceInfo& dev, cudaStream_t st, MsdTimers& timers, KernelTimer& scatter_timer,
                                Src src, uint64_t count, const uint32_t* parent_start, unsigned nparents,
                                unsigned bits, unsigned pieces_per_sm, unsigned hist_split,
                                uint64_t in_bytes_per_record, uint32_t* child_start, uint64_t* out) {
  const unsigned bins = 1u << bits;
  const uint64_t want_pieces = static_cast<uint64_t>(dev.sm_count) * pieces_per_sm;
  uint64_t target = (count + want_pieces - 1) / want_pieces;
  target = (target + kMsdTile - 1) / kMsdTile * kMsdTile;
  const unsigned max_pieces = static_cast<unsigned>(count / target + nparents + 1);
  DevBuf<MsdPiece> pieces(max_pieces, st);
  DevBuf<uint32_t> piece_first(nparents + 1, st), piece_count(1, st);
  DevBuf<uint32_t> hist(static_cast<uint64_t>(max_pieces) * bins, st);
  CAPSB_LAUNCH(msd_plan_kernel, 1, 1024, 0, st, parent_start, nparents, static_cast<uint32_t>(target), pieces.get(),
               piece_first.get(), piece_count.get());
  CAPSB_CUDA(cudaMemsetAsync(hist.get(), 0, static_cast<uint64_t>(max_pieces) * bins * sizeof(uint32_t), st));
  {
    MsdTimed timed(timers.hist, st, count * in_bytes_per_record);
    CAPSB_LAUNCH((msd_hist_kernel<Src>), dim3(max_pieces, hist_split), kMsdThreads, 0, st, src, pieces.get(),
                 piece_count.get(), bins, hist.get());
  }
  CAPSB_LAUNCH(msd_offsets_kernel, nparents, kMsdMaxBins, 0, st, parent_start, nparents, piece_first.get(), bins,
               hist.get(), child_start);
  {
    constexpr size_t kSmem = sizeof(MsdScatterSmem<kDigitInRec>);
    msd_allow_smem(msd_scatter_kernel<Src, kDigitInRec>, kSmem);
    MsdTimed timed(scatter_timer, st, count * (in_bytes_per_record + sizeof(uint64_t)));
    CAPSB_LAUNCH((msd_scatter_kernel<Src, kDigitInRec>), max_pieces, kMsdThreads, kSmem, st, src, pieces.get(),
                 piece_count.get(), bins, hist.get(), out);
  }
}

// The buckets the local sort could not take (more than kMsdLocalCap records): their records are
// gathered, sorted by (bucket, remaining key bits) with the LSD sort, and written back as keys
// and suffixes.
inline void msd_sort_large_buckets(const DeviceInfo& dev, cudaStream_t st, RadixScratch& radix, uint64_t* recs,
                                   uint32_t* sa_out, const uint32_t* child_start, const uint32_t* large_list,
                                   uint32_t nlarge, unsigned key_bits, unsigned prefix_bits,
                                   ScanScratch<uint32_t>& scan, uint64_t* records_out) {
  const unsigned rem_bits = key_bits - prefix_bits;
  DevBuf<uint32_t> loff(nlarge + 1, st);
  uint32_t total = 0;
  {
    uint32_t* lo = loff.get();
    device_scan<uint32_t, OpSum, false>(
        dev, st, scan, nlarge,
        [=] __device__(uint64_t j) -> uint32_t { return child_start[large_list[j] + 1] - child_start[large_list[j]]; },
        [=] __device__(uint64_t j, uint32_t v) { lo[j] = v; }, &total);
    launch_map(dev, st, 1, [=] __device__(uint64_t) { lo[nlarge] = total; });
  }
  if (records_out) *records_out = total;
  DevBuf<uint64_t> key_a(total, st), key_b(total, st);
  DevBuf<uint32_t> val_a(total, st), val_b(total, st);
  const uint32_t* lo = loff.get();
  auto bucket_of = [=] __device__(uint64_t e) -> uint32_t {  // last j with loff[j] <= e
    uint32_t a = 0, b = nlarge;
    while (b - a > 1) {
      const uint32_t mid = (a + b) >> 1;
      if (lo[mid] <= e)
        a = mid;
      else
        b = mid;
    }
    return a;
  };
  {
    uint64_t* ka = key_a.get();
    uint32_t* va = val_a.get();
    const uint64_t rem_mask = (1ull << rem_bits) - 1ull;
    launch_map(dev, st, total, [=] __device__(uint64_t e) {
      const uint32_t j = bucket_of(e);
      const uint64_t r = recs[child_start[large_list[j]] + (e - lo[j])];
      ka[e] = (static_cast<uint64_t>(j) << rem_bits) | ((r >> 32) & rem_mask);
      va[e] = static_cast<uint32_t>(r);
    });
  }
  const unsigned sort_bits = round_up8_bits(rem_bits + msd_ceil_log2(nlarge));
  const int where = radix_sort_pairs<uint64_t, uint32_t>(st, radix, key_a.get(), val_a.get(), key_b.get(), val_b.get(),
                                                        total, 0, sort_bits);
  {
    const uint64_t* ks = where ? key_b.get() : key_a.get();
    const uint32_t* vs = where ? val_b.get() : val_a.get();
    const uint64_t rem_mask = (1ull << rem_bits) - 1ull;
    const unsigned key_shift = 64u - key_bits;
    launch_map(dev, st, total, [=] __device__(uint64_t e) {
      const uint64_t k = ks[e];
      const uint32_t j = static_cast<uint32_t>(k >> rem_bits);
      const uint32_t q = large_list[j];
      const uint64_t pos = child_start[q] + (e - lo[j]);
      recs[pos] = ((static_cast<uint64_t>(q) << rem_bits) | (k & rem_mask)) << key_shift;
      sa_out[pos] = vs[e];
    });
  }
}

}  // namespace capsb
