// LSD radix sort, 8 bits per pass, stable, hand-written for sm_100a.
//
// One pass = three launches over a fixed grid of `blocks` CTAs, each owning a contiguous
// chunk of the input (so a pass needs no inter-CTA waiting and cannot hang):
//   radix_hist_kernel     per-CTA digit histogram                 -> hist[digit][cta]
//   radix_offsets_kernel  (256 CTAs) exclusive scan of each digit row + digit totals
//   radix_scatter_kernel  re-reads the chunk tile by tile, ranks the tile's keys inside each
//                         warp (stable), and scatters keys + values
// HBM traffic per pass: keys twice + values once in, keys + values once out.
//
// Keys come from a `Src` functor, so the first pass of the suffix sort reads the packed
// text directly (key = 64-bit window at suffix i, value = i) and never materialises an
// unsorted key array.
#pragma once

#include <cstdlib>
#include <type_traits>
#include <vector>

#include "common.cuh"

namespace capsb {

constexpr int kRadixBits = 8;
constexpr int kRadixSize = 1 << kRadixBits;
constexpr int kRsThreads = 256;         // histogram kernel
constexpr int kRsWarps = kRsThreads / 32;
constexpr int kRsTile = 8192;           // chunks are multiples of this; every scatter layout's tile (threads x items) divides it

template <class KeyT>
__device__ __forceinline__ unsigned radix_digit(KeyT key, unsigned shift) {
  return static_cast<unsigned>(key >> shift) & (kRadixSize - 1);
}

// Source functors -----------------------------------------------------------------------
template <class KeyT, class ValT>
struct ArraySource {
  const KeyT* keys;
  const ValT* vals;
  __device__ __forceinline__ KeyT key(uint64_t i) const { return keys[i]; }
  __device__ __forceinline__ ValT val(uint64_t i) const { return vals[i]; }
  static constexpr uint64_t bytes_read_per_item() { return sizeof(KeyT) + sizeof(ValT); }
  // keys are plain memory: the scatter kernel may fetch the next tile's keys asynchronously
  static constexpr bool kDirectKeys = true;
  __device__ __forceinline__ const KeyT* key_ptr(uint64_t i) const { return keys + i; }
};

template <class Src, class = void>
struct HasDirectKeys : std::false_type {};
template <class Src>
struct HasDirectKeys<Src, std::enable_if_t<Src::kDirectKeys>> : std::true_type {};

// Kernels -------------------------------------------------------------------------------
template <class KeyT, class Src>
__global__ void __launch_bounds__(kRsThreads) radix_hist_kernel(Src src, uint64_t n, uint64_t chunk,
                                                                unsigned shift, uint64_t* hist) {
  __shared__ unsigned counts[kRsWarps][kRadixSize];
  for (int i = threadIdx.x; i < kRsWarps * kRadixSize; i += kRsThreads) (&counts[0][0])[i] = 0;
  __syncthreads();
  const uint64_t begin = static_cast<uint64_t>(blockIdx.x) * chunk;
  const uint64_t end = begin + chunk < n ? begin + chunk : n;
  unsigned* mine = counts[threadIdx.x >> 5];
  for (uint64_t i = begin + threadIdx.x; i < end; i += kRsThreads)
    atomicAdd(&mine[radix_digit<KeyT>(src.key(i), shift)], 1u);
  __syncthreads();
  unsigned total = 0;
#pragma unroll
  for (int w = 0; w < kRsWarps; ++w) total += counts[w][threadIdx.x];
  hist[static_cast<uint64_t>(threadIdx.x) * gridDim.x + blockIdx.x] = total;
}

// CTA d scans row d of hist (one entry per sorting CTA) in place (exclusive) and writes the
// row total to digit_total[d].
static __global__ void __launch_bounds__(kScanThreads) radix_offsets_kernel(uint64_t* hist, unsigned blocks,
                                                                     uint64_t* digit_total) {
  __shared__ uint64_t smem[kScanThreads / 32];
  uint64_t* row = hist + static_cast<uint64_t>(blockIdx.x) * blocks;
  uint64_t carry = 0;
  for (unsigned base = 0; base < blocks; base += kScanThreads) {
    const unsigned i = base + threadIdx.x;
    const uint64_t v = i < blocks ? row[i] : 0;
    uint64_t inc, total;
    const uint64_t excl = block_scan<uint64_t, OpSum>(v, &inc, &total, smem);
    if (i < blocks) row[i] = carry + excl;
    carry += total;
  }
  if (threadIdx.x == 0) digit_total[blockIdx.x] = carry;
}

// ---------------------------------------------------------------------------------------
// Scatter kernel.  A CTA of kThreads threads walks its chunk in tiles of kThreads * kItems
// elements.  Per tile:
//   1. coalesced key loads (warp-striped: (warp, item, lane) lexicographic == input order),
//      value loads issued right behind them so their latency hides under the ranking;
//   2. stable in-warp ranking: every lane ORs its lane bit into the warp's per-digit mask word
//      in shared memory and reads the word back (= the lanes of this row that hold the same
//      digit); the lowest such lane adds the row's count to the warp's digit counter.
//      (~15 instructions per row; MATCH.ANY saturates the XU pipe and a ballot per digit bit
//      costs ~50 — profiles/r01/README.md.)  kRsMatchSlots rows are in flight together;
//   3. digit d: exclusive scan over the warps, digit starts inside the tile, global shifts;
//   4. keys and values are placed in digit-sorted order in shared memory (separate staging
//      areas when they fit, else the values reuse the key bytes);
//   5. consecutive threads write consecutive staged elements: only whole digit runs leave the
//      SM, so HBM sees ~1x the algorithmic write traffic.
// Occupancy is bounded by the register file (64 K registers per SM): kItems = 8 with 512
// threads keeps the 4096-element tile (16-element digit runs on average) at 64 registers per
// thread, i.e. 32 resident warps per SM instead of the 16 of a 256 x 16 layout (ncu of that
// layout: warps active 25 %, stalls spread over barrier / long / short scoreboard, DRAM 41 %).
// ---------------------------------------------------------------------------------------
constexpr int kRsMatchSlots = 2;  // rows of a warp whose peer masks are in flight together

template <class KeyT, class ValT, int kThreads, int kItems>
struct ScatterCfg {
  static constexpr int kWarps = kThreads / 32;
  static constexpr int kTile = kThreads * kItems;
  // both staging areas at once if that keeps two CTAs of this layout resident on an SM
  static constexpr bool kSeparate = (sizeof(KeyT) + sizeof(ValT)) * kTile <= 56 * 1024;
};

template <class KeyT, class ValT, int kThreads, int kItems, bool kPrefetch = false>
struct ScatterSmem {
  using Cfg = ScatterCfg<KeyT, ValT, kThreads, kItems>;
  // next tile's keys, fetched with cp.async while this tile is written out: [item][thread]
  alignas(16) KeyT prefetch[kPrefetch ? Cfg::kTile : 1];
  static constexpr size_t kStageBytes =
      Cfg::kSeparate ? (sizeof(KeyT) + sizeof(ValT)) * Cfg::kTile
                     : (sizeof(KeyT) > sizeof(ValT) ? sizeof(KeyT) : sizeof(ValT)) * Cfg::kTile;
  alignas(16) unsigned char stage[kStageBytes];
  uint64_t run_base[kRadixSize];     // next free global slot per digit for this CTA
  uint64_t out_shift[kRadixSize];    // global slot of tile-sorted position s is out_shift[digit] + s
  unsigned digit_start[kRadixSize];  // first tile-sorted position of each digit
  unsigned warp_cnt[Cfg::kWarps][kRadixSize + 1];  // [..][256] collects out-of-range lanes
  unsigned match[Cfg::kWarps][kRsMatchSlots][kRadixSize + 1];  // lanes of the warp holding each digit (one row)
  uint64_t scan_tmp[8];
  __device__ __forceinline__ KeyT* keys() { return reinterpret_cast<KeyT*>(stage); }
  __device__ __forceinline__ ValT* vals() {
    return reinterpret_cast<ValT*>(stage + (Cfg::kSeparate ? sizeof(KeyT) * Cfg::kTile : 0));
  }
};

// Exclusive sum over the 256 digits (threads 0..255 contribute `v`; called by every thread).
template <int kThreads>
__device__ __forceinline__ uint64_t digit_scan(uint64_t v, uint64_t* smem /*[8]*/) {
  const unsigned tid = threadIdx.x;
  const unsigned warp = tid >> 5;
  uint64_t inc = 0;
  if (tid < kRadixSize) {
    inc = warp_scan_inclusive<uint64_t, OpSum>(v);
    if ((tid & 31u) == 31u) smem[warp] = inc;
  }
  __syncthreads();
  uint64_t prefix = 0;
#pragma unroll
  for (int w = 0; w < kRadixSize / 32; ++w)
    if (static_cast<unsigned>(w) < warp) prefix += smem[w];
  return prefix + inc - v;
}

// One tile of the scatter pass.  kFull = the tile has all kTile elements (no bounds checks).
// kPrefetched: this tile's keys are already on their way to sm.prefetch (cp.async);
// next_tile/next_valid (kPrefetch only): the tile whose keys to fetch while this one is written out.
template <bool kFull, bool kPrefetch, class KeyT, class ValT, class Src, int kThreads, int kItems>
__device__ __forceinline__ void scatter_tile(ScatterSmem<KeyT, ValT, kThreads, kItems, kPrefetch>& sm, const Src& src,
                                             uint64_t tile, unsigned valid, unsigned shift,
                                             KeyT* __restrict__ keys_out, ValT* __restrict__ vals_out,
                                             bool prefetched, uint64_t next_tile, unsigned next_valid) {
  using Cfg = ScatterCfg<KeyT, ValT, kThreads, kItems>;
  constexpr int kWarps = Cfg::kWarps;
  constexpr bool kEarlyVals = Cfg::kSeparate && sizeof(ValT) <= 8;
  const unsigned tid = threadIdx.x;
  const unsigned lane = tid & 31u;
  const unsigned warp = tid >> 5;
  const unsigned lt = lanemask_lt();
  const unsigned lane_bit = 1u << lane;

  for (int i = tid; i < kWarps * (kRadixSize + 1); i += kThreads) (&sm.warp_cnt[0][0])[i] = 0;
  __syncthreads();  // also orders the previous tile's shared-memory reads before new writes

  const unsigned warp_first = warp * (32 * kItems) + lane;
  KeyT key[kItems];
  ValT val[kItems];
  unsigned slot[kItems];  // rank among equal digits of the warp -> tile-sorted position
  if (kPrefetch && prefetched) {
    asm volatile("cp.async.wait_all;" ::: "memory");  // own copies only: each thread reads what it fetched
#pragma unroll
    for (int t = 0; t < kItems; ++t) key[t] = sm.prefetch[t * kThreads + tid];
  } else {
#pragma unroll
    for (int t = 0; t < kItems; ++t) {
      const unsigned off = warp_first + t * 32;
      key[t] = (kFull || off < valid) ? src.key(tile + off) : KeyT(0);
    }
  }
  if (kEarlyVals) {
#pragma unroll
    for (int t = 0; t < kItems; ++t) {
      const unsigned off = warp_first + t * 32;
      val[t] = (kFull || off < valid) ? src.val(tile + off) : ValT{};
    }
  }
#pragma unroll
  for (int t0 = 0; t0 < kItems; t0 += kRsMatchSlots) {
    unsigned d[kRsMatchSlots];
#pragma unroll
    for (int u = 0; u < kRsMatchSlots; ++u) {
      const int t = t0 + u;
      const bool ok = kFull || warp_first + t * 32 < valid;
      d[u] = ok ? radix_digit<KeyT>(key[t], shift) : static_cast<unsigned>(kRadixSize);
      atomicOr(&sm.match[warp][u][d[u]], lane_bit);
    }
    __syncwarp();
    unsigned peers[kRsMatchSlots];
#pragma unroll
    for (int u = 0; u < kRsMatchSlots; ++u) peers[u] = sm.match[warp][u][d[u]];
    __syncwarp();  // every lane has its mask before the words are cleared
#pragma unroll
    for (int u = 0; u < kRsMatchSlots; ++u) {
      const int leader = __ffs(static_cast<int>(peers[u])) - 1;
      unsigned before = 0;
      if (static_cast<int>(lane) == leader) {
        sm.match[warp][u][d[u]] = 0;
        before = atomicAdd(&sm.warp_cnt[warp][d[u]], __popc(peers[u]));
      }
      before = __shfl_sync(0xffffffffu, before, leader);
      slot[t0 + u] = before + __popc(peers[u] & lt);
    }
    __syncwarp();  // the cleared mask words are visible before the next rows use them
  }
  __syncthreads();

  {  // digit `tid`: exclusive scan over warps, digit starts inside the tile, global shifts
    unsigned run = 0;
    if (tid < kRadixSize) {
#pragma unroll
      for (int w = 0; w < kWarps; ++w) {
        const unsigned c = sm.warp_cnt[w][tid];
        sm.warp_cnt[w][tid] = run;
        run += c;
      }
    }
    const uint64_t start = digit_scan<kThreads>(static_cast<uint64_t>(run), sm.scan_tmp);
    if (tid < kRadixSize) {
      const uint64_t base = sm.run_base[tid];
      sm.digit_start[tid] = static_cast<unsigned>(start);
      sm.out_shift[tid] = base - start;  // modular arithmetic: out_shift + s is exact
      sm.run_base[tid] = base + run;
    }
  }
  __syncthreads();

  KeyT* stage_keys = sm.keys();
  ValT* stage_vals = sm.vals();
#pragma unroll
  for (int t = 0; t < kItems; ++t) {
    if (kFull || warp_first + t * 32 < valid) {
      const unsigned d = radix_digit<KeyT>(key[t], shift);
      slot[t] += sm.digit_start[d] + sm.warp_cnt[warp][d];
      stage_keys[slot[t]] = key[t];
      if (kEarlyVals) stage_vals[slot[t]] = val[t];
    }
  }
  if constexpr (kPrefetch) {
    if (next_valid) {  // next tile's keys travel while this tile is written out
#pragma unroll
      for (int t = 0; t < kItems; ++t) {
        const unsigned off = warp_first + t * 32;
        if (off < next_valid) {
          const unsigned dst = static_cast<unsigned>(__cvta_generic_to_shared(&sm.prefetch[t * kThreads + tid]));
          asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(dst), "l"(src.key_ptr(next_tile + off)),
                       "n"(sizeof(KeyT))
                       : "memory");
        } else {
          sm.prefetch[t * kThreads + tid] = KeyT(0);
        }
      }
    }
  }
  if (!kEarlyVals) {  // the values are needed only after the keys have left; issue their loads now
#pragma unroll
    for (int t = 0; t < kItems; ++t) {
      const unsigned off = warp_first + t * 32;
      val[t] = (kFull || off < valid) ? src.val(tile + off) : ValT{};
    }
  }
  __syncthreads();

  unsigned digits[(kItems + 3) / 4];  // digit of tile-sorted position t*kThreads+tid, 8 bits each
#pragma unroll
  for (int t = 0; t < kItems; ++t) {
    const unsigned s = static_cast<unsigned>(t) * kThreads + tid;
    if ((t & 3) == 0) digits[t >> 2] = 0;
    if (kFull || s < valid) {
      const KeyT k = stage_keys[s];
      const unsigned d = radix_digit<KeyT>(k, shift);
      digits[t >> 2] |= d << (8 * (t & 3));
#ifdef CAPSB_RADIX_EXPERIMENT_SEQ_WRITE  // tools/radix_bench.cu only: timing without the scatter pattern
      keys_out[tile + s] = k;
      if (kEarlyVals) vals_out[tile + s] = stage_vals[s];
#else
      const uint64_t g = sm.out_shift[d] + s;
      keys_out[g] = k;
      if (kEarlyVals) vals_out[g] = stage_vals[s];
#endif
    }
  }
  if (!kEarlyVals) {
    if (!Cfg::kSeparate) __syncthreads();  // every key has been read; the staging bytes now take the values
#pragma unroll
    for (int t = 0; t < kItems; ++t)
      if (kFull || warp_first + t * 32 < valid) stage_vals[slot[t]] = val[t];
    __syncthreads();
#pragma unroll
    for (int t = 0; t < kItems; ++t) {
      const unsigned s = static_cast<unsigned>(t) * kThreads + tid;
      if (kFull || s < valid) {
        const unsigned d = (digits[t >> 2] >> (8 * (t & 3))) & 0xFFu;
#ifdef CAPSB_RADIX_EXPERIMENT_SEQ_WRITE
        vals_out[tile + s] = stage_vals[s];
#else
        vals_out[sm.out_shift[d] + s] = stage_vals[s];
#endif
      }
    }
  }
}

template <class KeyT, class ValT, class Src, int kThreads, int kItems, int kMinBlocks, bool kPrefetch>
__global__ void __launch_bounds__(kThreads, kMinBlocks) radix_scatter_kernel(Src src, uint64_t n, uint64_t chunk,
                                                                    unsigned shift,
                                                                    const uint64_t* __restrict__ hist,
                                                                    const uint64_t* __restrict__ digit_total,
                                                                    KeyT* __restrict__ keys_out,
                                                                    ValT* __restrict__ vals_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using Smem = ScatterSmem<KeyT, ValT, kThreads, kItems, kPrefetch>;
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  constexpr int kTile = kThreads * kItems;
  const unsigned tid = threadIdx.x;

  {  // global base of digit d = sum of totals of smaller digits; plus this CTA's row offset
    const uint64_t t = tid < kRadixSize ? digit_total[tid] : 0;
    const uint64_t excl = digit_scan<kThreads>(t, sm.scan_tmp);
    if (tid < kRadixSize) sm.run_base[tid] = excl + hist[static_cast<uint64_t>(tid) * gridDim.x + blockIdx.x];
  }
  for (int i = tid; i < (kThreads / 32) * kRsMatchSlots * (kRadixSize + 1); i += kThreads)
    (&sm.match[0][0][0])[i] = 0;

#ifdef CAPSB_RADIX_EXPERIMENT_TILE_ORDER
  // tools/radix_bench.cu only (results are garbage): tiles dealt round-robin to the CTAs and written
  // where a tile-ordered (onesweep-style) pass would put them, assuming 16 elements per digit and tile
  __shared__ uint64_t digit_base[kRadixSize];
  __syncthreads();
  if (tid < kRadixSize) digit_base[tid] = sm.run_base[tid] - hist[static_cast<uint64_t>(tid) * gridDim.x + blockIdx.x];
  const uint64_t full_tiles = n / kTile;
  for (uint64_t ti = blockIdx.x; ti < full_tiles; ti += gridDim.x) {
    __syncthreads();
    if (tid < kRadixSize) sm.run_base[tid] = digit_base[tid] + ti * (kTile / kRadixSize);
    scatter_tile<true, false, KeyT, ValT, Src, kThreads, kItems>(
        *reinterpret_cast<ScatterSmem<KeyT, ValT, kThreads, kItems, false>*>(smem_raw), src, ti * kTile, kTile, shift,
        keys_out, vals_out, false, 0, 0);
  }
  (void)chunk;
#else
  const uint64_t begin = static_cast<uint64_t>(blockIdx.x) * chunk;
  const uint64_t end = begin + chunk < n ? begin + chunk : n;
  uint64_t tile = begin;
  bool prefetched = false;
  for (; tile + kTile <= end; tile += kTile) {
    const uint64_t next = tile + kTile;
    const unsigned next_valid = next < end ? static_cast<unsigned>(end - next < kTile ? end - next : kTile) : 0u;
    scatter_tile<true, kPrefetch, KeyT, ValT, Src, kThreads, kItems>(sm, src, tile, kTile, shift, keys_out, vals_out,
                                                                    prefetched, next, next_valid);
    prefetched = kPrefetch && next_valid != 0;
  }
  if (tile < end)
    scatter_tile<false, kPrefetch, KeyT, ValT, Src, kThreads, kItems>(sm, src, tile, static_cast<unsigned>(end - tile),
                                                                     shift, keys_out, vals_out, prefetched, 0, 0);
#endif
}

// Host driver ---------------------------------------------------------------------------
// Optional per-launch timing of the scatter kernel (the dominant kernel of the pipeline).
struct KernelTimer {
  bool enabled = false;
  std::vector<cudaEvent_t> pool;                             // recycled events
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;  // recorded, not yet read
  uint64_t bytes = 0;
  cudaEvent_t get() {
    cudaEvent_t e;
    if (!pool.empty()) {
      e = pool.back();
      pool.pop_back();
    } else {
      CAPSB_CUDA(cudaEventCreate(&e));
    }
    return e;
  }
  // Sums and recycles the pending intervals (call after the stream has been synchronised).
  float drain(uint32_t* launches) {
    float total = 0;
    for (auto& pr : pending) {
      float ms = 0;
      CAPSB_CUDA(cudaEventElapsedTime(&ms, pr.first, pr.second));
      total += ms;
      pool.push_back(pr.first);
      pool.push_back(pr.second);
    }
    *launches = static_cast<uint32_t>(pending.size());
    pending.clear();
    return total;
  }
  void reset() {
    uint32_t ignored;
    if (!pending.empty()) drain(&ignored);
    bytes = 0;
  }
  ~KernelTimer() {
    for (auto& pr : pending) cudaEventDestroy(pr.first), cudaEventDestroy(pr.second);
    for (cudaEvent_t e : pool) cudaEventDestroy(e);
  }
};

struct RadixScratch {
  DevBuf<uint64_t> hist;         // [256][blocks]
  DevBuf<uint64_t> digit_total;  // [256]
  unsigned max_blocks = 0;
  int variant = 0;  // tools/radix_bench.cu only: 0 = the product choice (launch_scatter)
  int device = 0;
  KernelTimer timer;
  void init(const DeviceInfo& dev, cudaStream_t stream) {
    device = dev.device;
    if (const char* env = std::getenv("CAPSB_SCATTER_VARIANT")) variant = std::atoi(env);
    unsigned per_sm = 6;  // CTAs per SM in the grid (three waves of the two resident CTAs)
    if (const char* env = std::getenv("CAPSB_SCATTER_CTAS_PER_SM")) per_sm = static_cast<unsigned>(std::atoi(env));
    max_blocks = static_cast<unsigned>(dev.sm_count) * (per_sm ? per_sm : 6);
    hist.alloc(static_cast<uint64_t>(kRadixSize) * max_blocks, stream);
    digit_total.alloc(kRadixSize, stream);
  }
};

template <class KeyT, class ValT, class Src, int kThreads, int kItems, int kMinBlocks, bool kPrefetch>
inline void launch_scatter_variant(cudaStream_t stream, RadixScratch& rs, const Chunking& ck, Src src, uint64_t n,
                                   unsigned shift, KeyT* keys_out, ValT* vals_out) {
  constexpr size_t kSmem = sizeof(ScatterSmem<KeyT, ValT, kThreads, kItems, kPrefetch>);
  // every launch: the attribute belongs to the current context (one per device), several host threads
  // may drive several devices at once, and the call costs a few hundred nanoseconds
  CAPSB_CUDA(cudaFuncSetAttribute(radix_scatter_kernel<KeyT, ValT, Src, kThreads, kItems, kMinBlocks, kPrefetch>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmem)));
  CAPSB_LAUNCH((radix_scatter_kernel<KeyT, ValT, Src, kThreads, kItems, kMinBlocks, kPrefetch>), ck.blocks, kThreads,
               kSmem, stream, src, n, ck.chunk, shift, rs.hist.get(), rs.digit_total.get(), keys_out, vals_out);
}

template <class KeyT, class ValT, class Src>
inline void launch_scatter(cudaStream_t stream, RadixScratch& rs, const Chunking& ck, Src src, uint64_t n,
                           unsigned shift, KeyT* keys_out, ValT* vals_out) {
  // wide elements: one CTA per SM is all the shared memory allows
  constexpr int kMin = (sizeof(KeyT) + sizeof(ValT)) * 4096 <= 96 * 1024 ? 2 : 1;
  // asynchronous key prefetch where the keys are plain memory and the extra staging area fits
  constexpr bool kPre = HasDirectKeys<Src>::value && (2 * sizeof(KeyT) + sizeof(ValT)) * 4096 <= 80 * 1024;
#ifdef CAPSB_RADIX_ALL_VARIANTS  // tools/radix_bench.cu
  switch (rs.variant) {
    case 1: launch_scatter_variant<KeyT, ValT, Src, 512, 8, kMin, false>(stream, rs, ck, src, n, shift, keys_out, vals_out); return;
    case 2: launch_scatter_variant<KeyT, ValT, Src, 256, 16, kMin, false>(stream, rs, ck, src, n, shift, keys_out, vals_out); return;
    case 3: launch_scatter_variant<KeyT, ValT, Src, 256, 16, kMin, kPre>(stream, rs, ck, src, n, shift, keys_out, vals_out); return;
    case 4: launch_scatter_variant<KeyT, ValT, Src, 1024, 8, 1, false>(stream, rs, ck, src, n, shift, keys_out, vals_out); return;
    case 5: launch_scatter_variant<KeyT, ValT, Src, 512, 16, 1, false>(stream, rs, ck, src, n, shift, keys_out, vals_out); return;
    default: break;
  }
#endif
  // Measured (profiles/r01, 400 M (u64, u32) pairs): with the asynchronous key prefetch the
  // 512 x 8 layout is the fastest (3.02 TB/s); without it 256 x 16 is (2.96 vs 2.70 TB/s).
  if constexpr (kPre)
    launch_scatter_variant<KeyT, ValT, Src, 512, 8, kMin, true>(stream, rs, ck, src, n, shift, keys_out, vals_out);
  else
    launch_scatter_variant<KeyT, ValT, Src, 256, 16, kMin, false>(stream, rs, ck, src, n, shift, keys_out, vals_out);
}

// One stable counting pass on the 8-bit digit at `shift`.
template <class KeyT, class ValT, class Src>
inline void radix_pass(cudaStream_t stream, RadixScratch& rs, Src src, uint64_t n, unsigned shift,
                       KeyT* keys_out, ValT* vals_out) {
  if (n == 0) return;
  const Chunking ck = make_chunking(n, kRsTile, rs.max_blocks);
  CAPSB_LAUNCH((radix_hist_kernel<KeyT, Src>), ck.blocks, kRsThreads, 0, stream, src, n, ck.chunk, shift,
               rs.hist.get());
  CAPSB_LAUNCH(radix_offsets_kernel, kRadixSize, kScanThreads, 0, stream, rs.hist.get(), ck.blocks,
               rs.digit_total.get());
  cudaEvent_t t0 = nullptr, t1 = nullptr;
  if (rs.timer.enabled) {
    t0 = rs.timer.get();
    t1 = rs.timer.get();
    CAPSB_CUDA(cudaEventRecord(t0, stream));
  }
  launch_scatter<KeyT, ValT, Src>(stream, rs, ck, src, n, shift, keys_out, vals_out);
  if (rs.timer.enabled) {
    CAPSB_CUDA(cudaEventRecord(t1, stream));
    rs.timer.pending.emplace_back(t0, t1);
    rs.timer.bytes += n * (sizeof(KeyT) + sizeof(ValT)) + n * Src::bytes_read_per_item();
  }
}

// Sorts (keys, vals) by the key bits [begin_bit, end_bit) using ping-pong buffers.
// Returns 0 if the result is in (keys_a, vals_a), 1 if it is in (keys_b, vals_b).
// Input is read from (keys_a, vals_a).
template <class KeyT, class ValT>
inline int radix_sort_pairs(cudaStream_t stream, RadixScratch& rs, KeyT* keys_a, ValT* vals_a, KeyT* keys_b,
                            ValT* vals_b, uint64_t n, unsigned begin_bit, unsigned end_bit) {
  int cur = 0;
  for (unsigned shift = begin_bit; shift < end_bit; shift += kRadixBits) {
    KeyT* kin = cur ? keys_b : keys_a;
    ValT* vin = cur ? vals_b : vals_a;
    KeyT* kout = cur ? keys_a : keys_b;
    ValT* vout = cur ? vals_a : vals_b;
    radix_pass<KeyT, ValT>(stream, rs, ArraySource<KeyT, ValT>{kin, vin}, n, shift, kout, vout);
    cur ^= 1;
  }
  return cur;
}

}  // namespace capsb
