// LSD radix sort, 8 bits per pass, stable, hand-written for sm_100a.
//
// One pass = three launches over a fixed grid of `blocks` CTAs, each owning a contiguous
// chunk of the input (so a pass needs no inter-CTA waiting and cannot hang):
//   radix_hist_kernel     per-CTA digit histogram                 -> hist[digit][cta]
//   radix_offsets_kernel  (256 CTAs) exclusive scan of each digit row + digit totals
//   radix_scatter_kernel  re-reads the chunk tile by tile, ranks the tile's keys with
//                         warp match_any (stable), and scatters keys + values
// HBM traffic per pass: keys twice + values once in, keys + values once out.
//
// Keys come from a `Src` functor, so the first pass of the suffix sort reads the packed
// text directly (key = 64-bit window at suffix i, value = i) and never materialises an
// unsorted key array.
#pragma once

#include <cstdlib>
#include <vector>

#include "common.cuh"

namespace capsb {

constexpr int kRadixBits = 8;
constexpr int kRadixSize = 1 << kRadixBits;
constexpr int kRsThreads = 256;
constexpr int kRsWarps = kRsThreads / 32;
constexpr int kRsItems = 16;
constexpr int kRsTile = kRsThreads * kRsItems;

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
};

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

// Shared-memory layout of the scatter kernel.  The staging area holds the tile in
// digit-sorted order: first the keys, then (reusing the same bytes) the values.
constexpr int kRsMatchSlots = 2;  // rows of a warp whose peer masks are in flight together
template <class KeyT, class ValT>
struct ScatterSmem {
  union Stage {
    KeyT keys[kRsTile];
    ValT vals[kRsTile];
  } stage;
  uint64_t run_base[kRadixSize];    // next free global slot per digit for this CTA
  uint64_t out_shift[kRadixSize];   // global slot of tile-sorted position s is out_shift[digit] + s
  unsigned digit_start[kRadixSize]; // first tile-sorted position of each digit
  unsigned warp_cnt[kRsWarps][kRadixSize + 1];  // [..][256] collects out-of-range lanes
  unsigned match[kRsWarps][kRsMatchSlots][kRadixSize + 1];  // lanes of the warp holding each digit (one row)
  uint64_t scan_tmp[kScanThreads / 32];
};

// One tile of the scatter pass.  kFull = the tile has all kRsTile elements (no bounds checks).
//
// Ranking: every lane ORs its lane bit into the warp's per-digit mask word in shared memory,
// reads the word back (= the lanes of this row that hold the same digit), and the lowest such
// lane adds the row's count to the warp's digit counter.  That is ~15 instructions per row
// where a ballot per digit bit costs ~50 and MATCH.ANY saturates the XU pipe
// (profiles/r01/README.md).  Two rows are in flight at a time (kRsMatchSlots).
template <bool kFull, class KeyT, class ValT, class Src>
__device__ __forceinline__ void scatter_tile(ScatterSmem<KeyT, ValT>& sm, const Src& src, uint64_t tile, unsigned valid,
                                             unsigned shift, KeyT* __restrict__ keys_out,
                                             ValT* __restrict__ vals_out) {
  const unsigned tid = threadIdx.x;
  const unsigned lane = tid & 31u;
  const unsigned warp = tid >> 5;
  const unsigned lt = lanemask_lt();
  const unsigned lane_bit = 1u << lane;

  for (int i = tid; i < kRsWarps * (kRadixSize + 1); i += kRsThreads) (&sm.warp_cnt[0][0])[i] = 0;
  __syncthreads();  // also orders the previous tile's shared-memory reads before new writes

  // warp-striped layout keeps global loads coalesced and defines the stable order:
  // (warp, item, lane) lexicographic == increasing input index.
  const unsigned warp_first = warp * (32 * kRsItems) + lane;
  KeyT key[kRsItems];
  unsigned slot[kRsItems];  // rank among equal digits of the warp -> tile-sorted position
#pragma unroll
  for (int t = 0; t < kRsItems; ++t) {
    const unsigned off = warp_first + t * 32;
    key[t] = (kFull || off < valid) ? src.key(tile + off) : KeyT(0);
  }
#pragma unroll
  for (int t0 = 0; t0 < kRsItems; t0 += kRsMatchSlots) {
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
#pragma unroll
    for (int w = 0; w < kRsWarps; ++w) {
      const unsigned c = sm.warp_cnt[w][tid];
      sm.warp_cnt[w][tid] = run;
      run += c;
    }
    uint64_t inc, total;
    const uint64_t start = block_scan<uint64_t, OpSum>(static_cast<uint64_t>(run), &inc, &total, sm.scan_tmp);
    const uint64_t base = sm.run_base[tid];
    sm.digit_start[tid] = static_cast<unsigned>(start);
    sm.out_shift[tid] = base - start;  // modular arithmetic: out_shift + s is exact
    sm.run_base[tid] = base + run;
  }
  __syncthreads();

#pragma unroll
  for (int t = 0; t < kRsItems; ++t) {
    if (kFull || warp_first + t * 32 < valid) {
      const unsigned d = radix_digit<KeyT>(key[t], shift);
      slot[t] += sm.digit_start[d] + sm.warp_cnt[warp][d];
      sm.stage.keys[slot[t]] = key[t];
    }
  }
  // the values are needed only after the keys have left; issue their loads now
  ValT val[kRsItems];
#pragma unroll
  for (int t = 0; t < kRsItems; ++t) {
    const unsigned off = warp_first + t * 32;
    val[t] = (kFull || off < valid) ? src.val(tile + off) : ValT{};
  }
  __syncthreads();

  unsigned digits[kRsItems / 4];  // digit of tile-sorted position t*256+tid, 8 bits each
#pragma unroll
  for (int t = 0; t < kRsItems; ++t) {
    const unsigned s = static_cast<unsigned>(t) * kRsThreads + tid;
    if ((t & 3) == 0) digits[t >> 2] = 0;
    if (kFull || s < valid) {
      const KeyT k = sm.stage.keys[s];
      const unsigned d = radix_digit<KeyT>(k, shift);
      digits[t >> 2] |= d << (8 * (t & 3));
#ifdef CAPSB_RADIX_EXPERIMENT_SEQ_WRITE  // tools/radix_bench.cu only: timing without the scatter pattern
      keys_out[tile + s] = k;
#else
      keys_out[sm.out_shift[d] + s] = k;
#endif
    }
  }
  __syncthreads();  // every key has been read; the staging bytes now take the values

#pragma unroll
  for (int t = 0; t < kRsItems; ++t)
    if (kFull || warp_first + t * 32 < valid) sm.stage.vals[slot[t]] = val[t];
  __syncthreads();
#pragma unroll
  for (int t = 0; t < kRsItems; ++t) {
    const unsigned s = static_cast<unsigned>(t) * kRsThreads + tid;
    if (kFull || s < valid) {
      const unsigned d = (digits[t >> 2] >> (8 * (t & 3))) & 0xFFu;
#ifdef CAPSB_RADIX_EXPERIMENT_SEQ_WRITE
      vals_out[tile + s] = sm.stage.vals[s] + d;
#else
      vals_out[sm.out_shift[d] + s] = sm.stage.vals[s];
#endif
    }
  }
}

// Tile pipeline: coalesced key loads -> stable in-tile ranking -> keys reordered by digit in
// shared memory -> coalesced runs written to each digit's global range -> the same for the
// values through the same staging bytes.  Only full sectors leave the SM except at run
// boundaries, so HBM sees ~1x the algorithmic write traffic.
template <class KeyT, class ValT, class Src, int kMinBlocks>
__global__ void __launch_bounds__(kRsThreads, kMinBlocks) radix_scatter_kernel(Src src, uint64_t n, uint64_t chunk,
                                                                      unsigned shift,
                                                                      const uint64_t* __restrict__ hist,
                                                                      const uint64_t* __restrict__ digit_total,
                                                                      KeyT* __restrict__ keys_out,
                                                                      ValT* __restrict__ vals_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ScatterSmem<KeyT, ValT>& sm = *reinterpret_cast<ScatterSmem<KeyT, ValT>*>(smem_raw);
  const unsigned tid = threadIdx.x;

  {  // global base of digit d = sum of totals of smaller digits; plus this CTA's row offset
    const uint64_t t = digit_total[tid];
    uint64_t inc, total;
    const uint64_t excl = block_scan<uint64_t, OpSum>(t, &inc, &total, sm.scan_tmp);
    sm.run_base[tid] = excl + hist[static_cast<uint64_t>(tid) * gridDim.x + blockIdx.x];
  }
  for (int i = tid; i < kRsWarps * kRsMatchSlots * (kRadixSize + 1); i += kRsThreads) (&sm.match[0][0][0])[i] = 0;

  const uint64_t begin = static_cast<uint64_t>(blockIdx.x) * chunk;
  const uint64_t end = begin + chunk < n ? begin + chunk : n;
  uint64_t tile = begin;
  for (; tile + kRsTile <= end; tile += kRsTile)
    scatter_tile<true, KeyT, ValT, Src>(sm, src, tile, kRsTile, shift, keys_out, vals_out);
  if (tile < end)
    scatter_tile<false, KeyT, ValT, Src>(sm, src, tile, static_cast<unsigned>(end - tile), shift, keys_out, vals_out);
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
  int min_blocks = 2;  // resident scatter CTAs per SM the kernel is compiled for (measured best: profiles/r01)
  int device = 0;
  KernelTimer timer;
  void init(const DeviceInfo& dev, cudaStream_t stream) {
    device = dev.device;
    if (const char* env = std::getenv("CAPSB_SCATTER_MIN_BLOCKS")) min_blocks = std::atoi(env);
    unsigned per_sm = 6;  // CTAs per SM in the grid: a multiple of every min_blocks variant's residency
    if (const char* env = std::getenv("CAPSB_SCATTER_CTAS_PER_SM")) per_sm = static_cast<unsigned>(std::atoi(env));
    max_blocks = static_cast<unsigned>(dev.sm_count) * (per_sm ? per_sm : 6);
    hist.alloc(static_cast<uint64_t>(kRadixSize) * max_blocks, stream);
    digit_total.alloc(kRadixSize, stream);
  }
};

// Launches the scatter kernel variant selected by rs.min_blocks (register budget: 2 -> 128,
// 3 -> 80, 4 -> 64 registers per thread; tuning knob CAPSB_SCATTER_MIN_BLOCKS).
template <class KeyT, class ValT, class Src, int kMinBlocks>
inline void launch_scatter_variant(cudaStream_t stream, RadixScratch& rs, const Chunking& ck, Src src, uint64_t n,
                                   unsigned shift, KeyT* keys_out, ValT* vals_out) {
  constexpr size_t kSmem = sizeof(ScatterSmem<KeyT, ValT>);
  static bool configured[64] = {};  // per template instantiation and device (the attribute is per context)
  const int slot = rs.device & 63;
  if (!configured[slot] || rs.device >= 64) {
    CAPSB_CUDA(cudaFuncSetAttribute(radix_scatter_kernel<KeyT, ValT, Src, kMinBlocks>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmem)));
    configured[slot] = true;
  }
  CAPSB_LAUNCH((radix_scatter_kernel<KeyT, ValT, Src, kMinBlocks>), ck.blocks, kRsThreads, kSmem, stream, src, n,
               ck.chunk, shift, rs.hist.get(), rs.digit_total.get(), keys_out, vals_out);
}

template <class KeyT, class ValT, class Src>
inline void launch_scatter(cudaStream_t stream, RadixScratch& rs, const Chunking& ck, Src src, uint64_t n,
                           unsigned shift, KeyT* keys_out, ValT* vals_out) {
  switch (rs.min_blocks) {
    case 2: launch_scatter_variant<KeyT, ValT, Src, 2>(stream, rs, ck, src, n, shift, keys_out, vals_out); break;
    case 4: launch_scatter_variant<KeyT, ValT, Src, 4>(stream, rs, ck, src, n, shift, keys_out, vals_out); break;
    default: launch_scatter_variant<KeyT, ValT, Src, 3>(stream, rs, ck, src, n, shift, keys_out, vals_out); break;
  }
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
