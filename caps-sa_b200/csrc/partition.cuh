// Dedicated G-way (G <= 8) stable partition: the sharded construction's partition pass.
//
// Every suffix of a rank's text slice goes to the rank that owns its key bucket (reference
// locate_pivots + partition_sub_subarrays, src/Suffix_Array.cpp:225-368).  With at most eight
// buckets the ranking needs no shared-memory tables — one ballot per bucket and row, counts in
// registers — and a tile leaves the SM in runs of ~tile/G elements, long enough to be written
// straight into the owners' buffers: the destinations are a table of pointers, which are
// peer-mapped buffers of the other ranks when the caller fuses the pass with the exchange
// (sharded_build.cu, CAPSB_SHARD_P2P=1) and local buffers in tools/partition_bench.cu, where the
// kernels are timed and checked against the host on one GPU.
//   partition_count_kernel    per-CTA bucket counts              -> hist[bucket][cta]
//   partition_offsets_kernel  exclusive prefix over the CTAs, bucket totals
//   partition_scatter_kernel  stable scatter of the chunk through the pointer table
// `Src` supplies key(i) for i in [0, n); the value written for element i is base + i.
#pragma once

#include "common.cuh"

namespace capsb {

constexpr int kPartMaxBuckets = 8;
constexpr int kPartThreads = 256;
constexpr int kPartItems = 16;
constexpr int kPartWarps = kPartThreads / 32;
constexpr int kPartTile = kPartThreads * kPartItems;

struct PartPivots {
  uint64_t p[kPartMaxBuckets - 1];
  unsigned count;  // buckets - 1
  // bucket = number of pivots below the key (keys <= pivot j go to buckets <= j), as
  // BucketSource::key in sharded_build.cu
  __device__ __forceinline__ unsigned bucket(uint64_t key) const {
    unsigned b = 0;
#pragma unroll
    for (int j = 0; j < kPartMaxBuckets - 1; ++j) b += (static_cast<unsigned>(j) < count && p[j] < key) ? 1u : 0u;
    return b;
  }
};

template <class IdxT>
struct PartDestinations {
  IdxT* ptr[kPartMaxBuckets];  // where bucket q's elements of this rank start (a peer's memory in the product)
};

// Per-CTA bucket counts of the CTA's chunk -> hist[bucket * gridDim.x + cta].
template <class Src>
__global__ void __launch_bounds__(kPartThreads) partition_count_kernel(Src src, uint64_t n, uint64_t chunk, PartPivots piv,
                                                                       uint64_t* hist) {
  __shared__ unsigned totals[kPartMaxBuckets];
  if (threadIdx.x < kPartMaxBuckets) totals[threadIdx.x] = 0;
  __syncthreads();
  const uint64_t begin = static_cast<uint64_t>(blockIdx.x) * chunk;
  const uint64_t end = begin + chunk < n ? begin + chunk : n;
  unsigned mine[kPartMaxBuckets] = {};
  for (uint64_t i = begin + threadIdx.x; i < end; i += kPartThreads) {
    const unsigned b = piv.bucket(src.key(i));
#pragma unroll
    for (int q = 0; q < kPartMaxBuckets; ++q) mine[q] += (b == static_cast<unsigned>(q)) ? 1u : 0u;
  }
#pragma unroll
  for (int q = 0; q < kPartMaxBuckets; ++q) {
    unsigned v = mine[q];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    if ((threadIdx.x & 31u) == 0 && v) atomicAdd(&totals[q], v);
  }
  __syncthreads();
  if (threadIdx.x < kPartMaxBuckets) hist[static_cast<uint64_t>(threadIdx.x) * gridDim.x + blockIdx.x] = totals[threadIdx.x];
}

// hist[bucket][cta] -> exclusive prefix over the CTAs (in place); bucket totals -> total[bucket].
static __global__ void partition_offsets_kernel(uint64_t* hist, unsigned blocks, uint64_t* total) {
  const unsigned q = blockIdx.x;
  if (threadIdx.x != 0) return;  // a few hundred entries per bucket: serial is fine here
  uint64_t run = 0;
  for (unsigned c = 0; c < blocks; ++c) {
    const uint64_t v = hist[static_cast<uint64_t>(q) * blocks + c];
    hist[static_cast<uint64_t>(q) * blocks + c] = run;
    run += v;
  }
  total[q] = run;
}

// Stable partition of the CTA's chunk: element i (value base + i) goes to
// dst.ptr[bucket] + cta_base[bucket][cta] + (its rank among the CTA's elements of that bucket).
template <class IdxT, class Src>
__global__ void __launch_bounds__(kPartThreads) partition_scatter_kernel(Src src, uint64_t n, uint64_t chunk, uint64_t base,
                                                                         PartPivots piv, const uint64_t* __restrict__ cta_base,
                                                                         PartDestinations<IdxT> dst) {
  __shared__ IdxT stage[kPartTile];
  __shared__ unsigned warp_tot[kPartWarps][kPartMaxBuckets];   // per warp and bucket: elements in this tile
  __shared__ unsigned warp_base[kPartWarps][kPartMaxBuckets];  // tile-sorted position of the warp's first one
  __shared__ unsigned start[kPartMaxBuckets + 1];              // tile-sorted position of each bucket
  __shared__ uint64_t run[kPartMaxBuckets];                    // next free slot of this CTA in each destination
  __shared__ IdxT* out_ptr[kPartMaxBuckets];                   // (a dynamically indexed kernel parameter would live in local memory)
  const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const unsigned lt = lanemask_lt();
  if (tid < kPartMaxBuckets) {
    run[tid] = cta_base[static_cast<uint64_t>(tid) * gridDim.x + blockIdx.x];
#pragma unroll
    for (int q = 0; q < kPartMaxBuckets; ++q)
      if (tid == static_cast<unsigned>(q)) out_ptr[q] = dst.ptr[q];
  }
  const uint64_t begin = static_cast<uint64_t>(blockIdx.x) * chunk;
  const uint64_t end = begin + chunk < n ? begin + chunk : n;
  for (uint64_t tile = begin; tile < end; tile += kPartTile) {
    const unsigned valid = end - tile < kPartTile ? static_cast<unsigned>(end - tile) : kPartTile;
    const unsigned warp_first = warp * (32 * kPartItems) + lane;
    // 1. bucket of every element; stable rank inside the warp: ballots, counts stay in registers
    unsigned bucket_of[kPartItems], pos[kPartItems];
    unsigned cnt[kPartMaxBuckets] = {};
    uint64_t key[kPartItems];
#pragma unroll
    for (int t = 0; t < kPartItems; ++t) {
      const unsigned off = warp_first + t * 32;
      key[t] = off < valid ? src.key(tile + off) : 0;
    }
#pragma unroll
    for (int t = 0; t < kPartItems; ++t) {
      const unsigned off = warp_first + t * 32;
      const unsigned b = off < valid ? piv.bucket(key[t]) : static_cast<unsigned>(kPartMaxBuckets);
      bucket_of[t] = b;
      pos[t] = 0;
#pragma unroll
      for (int q = 0; q < kPartMaxBuckets; ++q) {
        const unsigned mask = __ballot_sync(0xffffffffu, b == static_cast<unsigned>(q));
        if (b == static_cast<unsigned>(q)) pos[t] = cnt[q] + __popc(mask & lt);
        cnt[q] += __popc(mask);
      }
    }
    // 2. positions of the warps' runs inside the tile (bucket-major, then warp, then input order)
    if (lane < kPartMaxBuckets) {
      unsigned v = 0;
#pragma unroll
      for (int q = 0; q < kPartMaxBuckets; ++q) v = lane == static_cast<unsigned>(q) ? cnt[q] : v;
      warp_tot[warp][lane] = v;
    }
    __syncthreads();
    if (tid < kPartMaxBuckets) {  // one thread per bucket: totals; thread 0 then scans the eight totals
      unsigned total = 0;
#pragma unroll
      for (int w = 0; w < kPartWarps; ++w) total += warp_tot[w][tid];
      start[tid + 1] = total;
    }
    __syncthreads();
    if (tid == 0) {
      unsigned at = 0;
      start[0] = 0;
#pragma unroll
      for (int q = 0; q < kPartMaxBuckets; ++q) {
        const unsigned c = start[q + 1];
        start[q] = at;
        at += c;
      }
      start[kPartMaxBuckets] = at;
    }
    __syncthreads();
    if (tid < kPartWarps * kPartMaxBuckets) {
      const unsigned w = tid / kPartMaxBuckets, q = tid % kPartMaxBuckets;
      unsigned at = start[q];
      for (unsigned v = 0; v < w; ++v) at += warp_tot[v][q];
      warp_base[w][q] = at;
    }
    __syncthreads();
    // 3. stage the values in tile-sorted order
#pragma unroll
    for (int t = 0; t < kPartItems; ++t) {
      const unsigned off = warp_first + t * 32;
      if (off < valid) stage[warp_base[warp][bucket_of[t]] + pos[t]] = static_cast<IdxT>(base + tile + off);
    }
    __syncthreads();
    // 4. consecutive threads write consecutive staged elements: runs of ~tile/G leave the SM
#pragma unroll
    for (int j = 0; j < kPartItems; ++j) {
      const unsigned s = static_cast<unsigned>(j) * kPartThreads + tid;
      if (s < valid) {
        unsigned q = 0;
#pragma unroll
        for (int r = 1; r < kPartMaxBuckets; ++r) q += s >= start[r] ? 1u : 0u;
        // empty buckets share a start: the count above lands on the last bucket that starts at or before s,
        // which is the non-empty one holding s
        out_ptr[q][run[q] + (s - start[q])] = stage[s];
      }
    }
    __syncthreads();
    if (tid < kPartMaxBuckets) run[tid] += start[tid + 1] - start[tid];
    // (the next tile's first barrier orders this update and the reuse of the staging area)
  }
}

}  // namespace capsb
