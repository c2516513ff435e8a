// Bucket merge (replaces merge_sub_subarrays / sort_partition / merge, reference
// src/Suffix_Array.cpp:371-428 and :48-109): after the exchange a rank holds one key-sorted
// run per source rank; a balanced tree of two-way merge-path merges turns them into one
// sorted bucket.  Every CTA produces a fixed tile of the output: it finds its stretch of the
// two runs with one diagonal binary search per side, stages the stretch in shared memory,
// every thread then finds its own diagonal in the staged stretch and merges kMpItems
// elements serially.  Ties keep the left run first, so the tree is a stable merge in source
// rank order.  LCPs are not carried through the merge: they are recovered afterwards from the
// keys of neighbours (pipeline.cuh, key_lcp), which is one streaming pass instead of a
// second payload in every merge level.
#pragma once

#include <vector>

#include "common.cuh"

namespace capsb {

constexpr int kMpThreads = 256;
constexpr int kMpItems = 8;
constexpr int kMpTile = kMpThreads * kMpItems;

// Number of elements of a[0..na) among the first `diag` outputs of the stable merge of a and b.
template <class KeyT>
__device__ __forceinline__ uint64_t merge_path_split(const KeyT* a, uint64_t na, const KeyT* b, uint64_t nb,
                                                     uint64_t diag) {
  uint64_t lo = diag > nb ? diag - nb : 0;
  uint64_t hi = diag < na ? diag : na;
  while (lo < hi) {
    const uint64_t mid = (lo + hi) >> 1;
    // taking `mid` from a leaves b[diag-1-mid] as the last element taken from b; it may only
    // precede a[mid] if it is strictly smaller (ties go to a)
    if (b[diag - 1 - mid] < a[mid])
      hi = mid;
    else
      lo = mid + 1;
  }
  return lo;
}

template <class KeyT, class ValT>
__global__ void __launch_bounds__(kMpThreads) merge_path_kernel(const KeyT* __restrict__ ka, const ValT* __restrict__ va,
                                                                uint64_t na, const KeyT* __restrict__ kb,
                                                                const ValT* __restrict__ vb, uint64_t nb,
                                                                KeyT* __restrict__ kout, ValT* __restrict__ vout) {
  __shared__ KeyT skeys[kMpTile];
  __shared__ KeyT sout[kMpTile];
  __shared__ unsigned ssrc[kMpTile];  // position in the staged stretch each output came from
  __shared__ uint64_t split[2];
  const uint64_t total = na + nb;
  const uint64_t tiles = (total + kMpTile - 1) / kMpTile;
  for (uint64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const uint64_t d0 = tile * kMpTile;
    const uint64_t d1 = d0 + kMpTile < total ? d0 + kMpTile : total;
    __syncthreads();  // previous iteration's shared-memory reads are done
    if (threadIdx.x < 2) split[threadIdx.x] = merge_path_split<KeyT>(ka, na, kb, nb, threadIdx.x ? d1 : d0);
    __syncthreads();
    const uint64_t a0 = split[0], a1 = split[1];
    const uint64_t b0 = d0 - a0, b1 = d1 - a1;
    const unsigned ca = static_cast<unsigned>(a1 - a0), cb = static_cast<unsigned>(b1 - b0);
    const unsigned count = ca + cb;
    for (unsigned s = threadIdx.x; s < count; s += kMpThreads) skeys[s] = s < ca ? ka[a0 + s] : kb[b0 + (s - ca)];
    __syncthreads();

    const KeyT* sa_ = skeys;
    const KeyT* sb_ = skeys + ca;
    const unsigned diag = threadIdx.x * kMpItems < count ? threadIdx.x * kMpItems : count;
    unsigned i = static_cast<unsigned>(merge_path_split<KeyT>(sa_, ca, sb_, cb, diag));
    unsigned j = diag - i;
#pragma unroll
    for (int q = 0; q < kMpItems; ++q) {
      const unsigned o = diag + q;
      if (o < count) {
        const bool take_a = j >= cb || (i < ca && !(sb_[j] < sa_[i]));
        const unsigned from = take_a ? i : ca + j;
        sout[o] = skeys[from];
        ssrc[o] = from;
        if (take_a)
          ++i;
        else
          ++j;
      }
    }
    __syncthreads();
    for (unsigned s = threadIdx.x; s < count; s += kMpThreads) {
      const unsigned from = ssrc[s];
      kout[d0 + s] = sout[s];
      vout[d0 + s] = from < ca ? va[a0 + from] : vb[b0 + (from - ca)];
    }
  }
}

// Merges the `runs` sorted runs stored back to back in (keys, vals) — run r occupies
// [offsets[r], offsets[r+1]) — using (keys_tmp, vals_tmp) as the ping-pong partner.
// Returns 0 if the merged bucket ends in (keys, vals), 1 if it ends in the partner.
template <class KeyT, class ValT>
inline int merge_sorted_runs(const DeviceInfo& dev, cudaStream_t st, KeyT* keys, ValT* vals, KeyT* keys_tmp,
                             ValT* vals_tmp, std::vector<uint64_t> offsets) {
  int cur = 0;
  while (offsets.size() > 2) {
    KeyT* kin = cur ? keys_tmp : keys;
    ValT* vin = cur ? vals_tmp : vals;
    KeyT* kout = cur ? keys : keys_tmp;
    ValT* vout = cur ? vals : vals_tmp;
    std::vector<uint64_t> next;
    next.push_back(offsets[0]);
    const size_t runs = offsets.size() - 1;
    for (size_t r = 0; r < runs; r += 2) {
      const uint64_t lo = offsets[r];
      if (r + 1 < runs) {
        const uint64_t mid = offsets[r + 1], hi = offsets[r + 2];
        const uint64_t total = hi - lo;
        if (total) {
          const uint64_t tiles = ceil_div(total, kMpTile);
          const uint64_t cap = static_cast<uint64_t>(dev.sm_count) * 8;
          const unsigned grid = static_cast<unsigned>(tiles < cap ? tiles : cap);
          CAPSB_LAUNCH((merge_path_kernel<KeyT, ValT>), grid, kMpThreads, 0, st, kin + lo, vin + lo, mid - lo,
                       kin + mid, vin + mid, hi - mid, kout + lo, vout + lo);
        }
        next.push_back(hi);
      } else {  // odd run out: carried to the next level unchanged
        const uint64_t hi = offsets[r + 1];
        if (hi > lo) {
          CAPSB_CUDA(cudaMemcpyAsync(kout + lo, kin + lo, (hi - lo) * sizeof(KeyT), cudaMemcpyDeviceToDevice, st));
          CAPSB_CUDA(cudaMemcpyAsync(vout + lo, vin + lo, (hi - lo) * sizeof(ValT), cudaMemcpyDeviceToDevice, st));
        }
        next.push_back(hi);
      }
    }
    offsets.swap(next);
    cur ^= 1;
  }
  return cur;
}

}  // namespace capsb
