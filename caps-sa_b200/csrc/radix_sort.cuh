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

template <class KeyT, class ValT, class Src>
__global__ void __launch_bounds__(kRsThreads) radix_scatter_kernel(Src src, uint64_t n, uint64_t chunk,
                                                                   unsigned shift,
                                                                   const uint64_t* __restrict__ hist,
                                                                   const uint64_t* __restrict__ digit_total,
                                                                   KeyT* __restrict__ keys_out,
                                                                   ValT* __restrict__ vals_out) {
  __shared__ uint64_t run_base[kRadixSize];          // next free global slot per digit for this CTA
  __shared__ uint64_t tile_base[kRadixSize];         // run_base at the start of the current tile
  __shared__ unsigned warp_cnt[kRsWarps][kRadixSize + 1];  // [..][256] collects out-of-range lanes
  __shared__ uint64_t scan_smem[kScanThreads / 32];

  const unsigned lane = lane_id();
  const unsigned warp = threadIdx.x >> 5;
  const unsigned lt = lanemask_lt();

  {  // global base of digit d = sum of totals of smaller digits; plus this CTA's row offset
    const uint64_t t = digit_total[threadIdx.x];
    uint64_t inc, total;
    const uint64_t excl = block_scan<uint64_t, OpSum>(t, &inc, &total, scan_smem);
    run_base[threadIdx.x] = excl + hist[static_cast<uint64_t>(threadIdx.x) * gridDim.x + blockIdx.x];
  }
  __syncthreads();

  const uint64_t begin = static_cast<uint64_t>(blockIdx.x) * chunk;
  const uint64_t end = begin + chunk < n ? begin + chunk : n;

  for (uint64_t tile = begin; tile < end; tile += kRsTile) {
    for (int i = threadIdx.x; i < kRsWarps * (kRadixSize + 1); i += kRsThreads) (&warp_cnt[0][0])[i] = 0;
    __syncthreads();

    // warp-striped layout keeps global loads coalesced and defines the stable order:
    // (warp, item, lane) lexicographic == increasing input index.
    const uint64_t warp_first = tile + static_cast<uint64_t>(warp) * (32 * kRsItems);
    KeyT key[kRsItems];
    unsigned dig[kRsItems];
    unsigned rank[kRsItems];
#pragma unroll
    for (int t = 0; t < kRsItems; ++t) {
      const uint64_t i = warp_first + static_cast<uint64_t>(t) * 32 + lane;
      const bool ok = i < end;
      key[t] = ok ? src.key(i) : KeyT(0);
      dig[t] = ok ? radix_digit<KeyT>(key[t], shift) : static_cast<unsigned>(kRadixSize);
    }
#pragma unroll
    for (int t = 0; t < kRsItems; ++t) {
      const unsigned peers = __match_any_sync(0xffffffffu, dig[t]);
      const int leader = __ffs(static_cast<int>(peers)) - 1;
      unsigned before = 0;
      if (static_cast<int>(lane) == leader) {
        before = warp_cnt[warp][dig[t]];
        warp_cnt[warp][dig[t]] = before + __popc(peers);
      }
      before = __shfl_sync(0xffffffffu, before, leader);
      rank[t] = before + __popc(peers & lt);
      __syncwarp();
    }
    __syncthreads();

    {  // digit threadIdx.x: exclusive scan over warps, then advance the CTA's running base
      unsigned run = 0;
#pragma unroll
      for (int w = 0; w < kRsWarps; ++w) {
        const unsigned c = warp_cnt[w][threadIdx.x];
        warp_cnt[w][threadIdx.x] = run;
        run += c;
      }
      const uint64_t base = run_base[threadIdx.x];
      tile_base[threadIdx.x] = base;
      run_base[threadIdx.x] = base + run;
    }
    __syncthreads();

#pragma unroll
    for (int t = 0; t < kRsItems; ++t) {
      if (dig[t] < static_cast<unsigned>(kRadixSize)) {
        const uint64_t i = warp_first + static_cast<uint64_t>(t) * 32 + lane;
        const uint64_t pos = tile_base[dig[t]] + warp_cnt[warp][dig[t]] + rank[t];
        keys_out[pos] = key[t];
        vals_out[pos] = src.val(i);
      }
    }
    __syncthreads();
  }
}

// Host driver ---------------------------------------------------------------------------
struct RadixScratch {
  DevBuf<uint64_t> hist;         // [256][blocks]
  DevBuf<uint64_t> digit_total;  // [256]
  unsigned max_blocks = 0;
  void init(const DeviceInfo& dev, cudaStream_t stream) {
    max_blocks = static_cast<unsigned>(dev.sm_count) * 4;
    hist.alloc(static_cast<uint64_t>(kRadixSize) * max_blocks, stream);
    digit_total.alloc(kRadixSize, stream);
  }
};

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
  CAPSB_LAUNCH((radix_scatter_kernel<KeyT, ValT, Src>), ck.blocks, kRsThreads, 0, stream, src, n, ck.chunk,
               shift, rs.hist.get(), rs.digit_total.get(), keys_out, vals_out);
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
