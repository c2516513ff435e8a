// Micro-benchmark + self-check of a dedicated G-way partition kernel (not product code yet).
//
// The sharded construction's partition pass (sharded_build.cu, "partition first") sends every
// suffix of a rank's text slice to the rank that owns its key bucket.  Today that is one pass of
// the generic 256-digit radix machinery with the bucket as digit (15 ms for 1.55 G suffixes at
// G = 2) followed by an ncclSend/Recv all-to-all (17 ms).  With at most eight buckets the ranking
// needs no shared-memory tables at all — one ballot per bucket and row, counts in registers — and
// the tile leaves the SM in runs of ~tile/G elements, long enough to be written straight into the
// owners' buffers over NVLink.  This file is that kernel with the destinations given as a table
// of pointers (local buffers here; peer-mapped buffers in the product), so that it can be timed
// and validated on one GPU before it replaces the two steps.
//
//   partition_bench [n] [buckets] [reps]      (defaults 4e8, 8, 10)
//
// Self-check (always on): every bucket's output must be the increasing list of exactly the
// indices whose key falls between its pivots (stable partition), verified on the host.
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../caps-sa_b200/csrc/common.cuh"

namespace capsb {
std::atomic<uint64_t> g_kernel_launches{0};
}
using namespace capsb;

constexpr int kPartMaxBuckets = 8;
constexpr int kPartThreads = 256;
constexpr int kPartItems = 16;
constexpr int kPartWarps = kPartThreads / 32;
constexpr int kPartTile = kPartThreads * kPartItems;

struct Pivots {
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
struct Destinations {
  IdxT* ptr[kPartMaxBuckets];  // where bucket q's elements of this rank start (a peer's memory in the product)
};

struct ArrayKeys {
  const uint64_t* keys;
  __device__ __forceinline__ uint64_t key(uint64_t i) const { return keys[i]; }
};

// Per-CTA bucket counts of the CTA's chunk -> hist[bucket * gridDim.x + cta].
template <class Src>
__global__ void __launch_bounds__(kPartThreads) partition_count_kernel(Src src, uint64_t n, uint64_t chunk, Pivots piv,
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
                                                                         Pivots piv, const uint64_t* __restrict__ cta_base,
                                                                         Destinations<IdxT> dst) {
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

__global__ void fill_random(uint64_t* keys, uint64_t n) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    uint64_t z = i * 0x9E3779B97F4A7C15ull + 0x1234567ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    keys[i] = z ^ (z >> 31);
  }
}

int main(int argc, char** argv) {
  const uint64_t n = argc > 1 ? (uint64_t)atof(argv[1]) : 400000000ull;
  const unsigned buckets = argc > 2 ? (unsigned)atoi(argv[2]) : 8;
  const int reps = argc > 3 ? atoi(argv[3]) : 10;
  if (buckets < 1 || buckets > kPartMaxBuckets || n == 0 || n > 0xFFFFFFFFull) {
    fprintf(stderr, "usage: partition_bench [n <= 2^32-1] [buckets 1..8] [reps]\n");
    return 2;
  }
  try {
    cudaDeviceProp prop;
    CAPSB_CUDA(cudaGetDeviceProperties(&prop, 0));
    cudaStream_t st;
    CAPSB_CUDA(cudaStreamCreate(&st));
    Arena arena;
    ArenaScope scope(&arena);
    const unsigned max_blocks = prop.multiProcessorCount * 8;
    const Chunking ck = make_chunking(n, kPartTile, max_blocks);
    DevBuf<uint64_t> keys(n, st), hist((uint64_t)kPartMaxBuckets * ck.blocks, st), total(kPartMaxBuckets, st);
    DevBuf<uint32_t> out(n, st);
    fill_random<<<prop.multiProcessorCount * 8, 256, 0, st>>>(keys.get(), n);
    Pivots piv{};
    piv.count = buckets - 1;
    for (unsigned j = 0; j + 1 < buckets; ++j) piv.p[j] = (~0ull / buckets) * (j + 1);  // skewed a little on purpose below
    if (buckets > 2) piv.p[0] /= 3;  // an uneven split: bucket 0 small, bucket 1 large
    cudaEvent_t e0, e1, e2;
    cudaEventCreate(&e0), cudaEventCreate(&e1), cudaEventCreate(&e2);
    float ms_count = 0, ms_scatter = 0;
    std::vector<uint64_t> h_total(kPartMaxBuckets);
    for (int r = -2; r < reps; ++r) {  // two warm-up rounds
      cudaEventRecord(e0, st);
      CAPSB_LAUNCH((partition_count_kernel<ArrayKeys>), ck.blocks, kPartThreads, 0, st, ArrayKeys{keys.get()}, n, ck.chunk,
                   piv, hist.get());
      CAPSB_LAUNCH(partition_offsets_kernel, kPartMaxBuckets, 32, 0, st, hist.get(), ck.blocks, total.get());
      cudaEventRecord(e1, st);
      // destinations: the buckets packed one after the other in `out` (needs the totals: tiny read-back)
      CAPSB_CUDA(cudaMemcpyAsync(h_total.data(), total.get(), kPartMaxBuckets * 8, cudaMemcpyDeviceToHost, st));
      CAPSB_CUDA(cudaStreamSynchronize(st));
      Destinations<uint32_t> dst{};
      uint64_t at = 0;
      for (int q = 0; q < kPartMaxBuckets; ++q) dst.ptr[q] = out.get() + at, at += h_total[q];
      cudaEventRecord(e1, st);
      CAPSB_LAUNCH((partition_scatter_kernel<uint32_t, ArrayKeys>), ck.blocks, kPartThreads, 0, st, ArrayKeys{keys.get()}, n,
                   ck.chunk, (uint64_t)0, piv, hist.get(), dst);
      cudaEventRecord(e2, st);
      CAPSB_CUDA(cudaStreamSynchronize(st));
      if (r >= 0) {
        float a = 0, b = 0;
        cudaEventElapsedTime(&a, e0, e1);  // includes the host round trip of the totals
        cudaEventElapsedTime(&b, e1, e2);
        ms_count += a, ms_scatter += b;
      }
    }
    printf("n=%llu buckets=%u: count+offsets %.3f ms, scatter %.3f ms (%.1f GB/s for 8 B read + 4 B written per element)\n",
           (unsigned long long)n, buckets, ms_count / reps, ms_scatter / reps, 12.0 * n / (ms_scatter / reps) / 1e6);
    // ---- self-check on the host ---------------------------------------------------------------
    std::vector<uint64_t> hk(n);
    std::vector<uint32_t> ho(n);
    CAPSB_CUDA(cudaMemcpy(hk.data(), keys.get(), n * 8, cudaMemcpyDeviceToHost));
    CAPSB_CUDA(cudaMemcpy(ho.data(), out.get(), n * 4, cudaMemcpyDeviceToHost));
    auto bucket_of = [&](uint64_t k) {
      unsigned b = 0;
      for (unsigned j = 0; j < piv.count; ++j) b += piv.p[j] < k ? 1u : 0u;
      return b;
    };
    std::vector<uint64_t> want(kPartMaxBuckets, 0);
    for (uint64_t i = 0; i < n; ++i) want[bucket_of(hk[i])]++;
    bool ok = true;
    uint64_t at = 0;
    for (int q = 0; q < kPartMaxBuckets && ok; ++q) {
      if (want[q] != h_total[q]) {
        printf("bucket %d: %llu elements, expected %llu\n", q, (unsigned long long)h_total[q], (unsigned long long)want[q]);
        ok = false;
        break;
      }
      for (uint64_t j = 0; j < want[q]; ++j) {
        const uint32_t i = ho[at + j];
        if (i >= n || bucket_of(hk[i]) != (unsigned)q || (j > 0 && ho[at + j - 1] >= i)) {
          printf("bucket %d, slot %llu: element %u is out of place\n", q, (unsigned long long)j, i);
          ok = false;
          break;
        }
      }
      at += want[q];
    }
    printf("stable partition verified on the host: %s\n", ok ? "yes" : "NO");
    return ok ? 0 : 1;
  } catch (const std::exception& e) {
    fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
}
