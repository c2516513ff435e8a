// Micro-benchmark + self-check of the dedicated G-way partition kernels (caps-sa_b200/csrc/partition.cuh).
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

#include "../caps-sa_b200/csrc/partition.cuh"

namespace capsb {
std::atomic<uint64_t> g_kernel_launches{0};
}
using namespace capsb;

struct ArrayKeys {
  const uint64_t* keys;
  __device__ __forceinline__ uint64_t key(uint64_t i) const { return keys[i]; }
};

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
    PartPivots piv{};
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
      PartDestinations<uint32_t> dst{};
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
