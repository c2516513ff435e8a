// Micro-benchmark of one radix pass (not product code): times radix_hist / radix_scatter on
// uniformly random 64-bit keys + 32-bit values.  usage: radix_bench [n] [reps] [shift]
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CAPSB_RADIX_ALL_VARIANTS 1
#include "../caps-sa_b200/csrc/radix_sort.cuh"

namespace capsb {
std::atomic<uint64_t> g_kernel_launches{0};
}
using namespace capsb;

__global__ void fill_random(uint64_t* keys, uint32_t* vals, uint64_t n) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    uint64_t z = i * 0x9E3779B97F4A7C15ull + 0x1234567ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    keys[i] = z ^ (z >> 31);
    vals[i] = (uint32_t)i;
  }
}

int main(int argc, char** argv) {
  const uint64_t n = argc > 1 ? (uint64_t)atof(argv[1]) : 100000000ull;
  const int reps = argc > 2 ? atoi(argv[2]) : 10;
  const unsigned shift = argc > 3 ? atoi(argv[3]) : 0;
  try {
    DeviceInfo dev;
    cudaDeviceProp prop;
    CAPSB_CUDA(cudaGetDeviceProperties(&prop, 0));
    dev.sm_count = prop.multiProcessorCount;
    cudaStream_t st;
    CAPSB_CUDA(cudaStreamCreate(&st));
    RadixScratch rs;
    rs.init(dev, st);
    DevBuf<uint64_t> ka(n, st), kb(n + (8u << 20), st);  // slack: the TILE_ORDER experiment overshoots
    DevBuf<uint32_t> va(n, st), vb(n + (8u << 20), st);
    fill_random<<<dev.sm_count * 8, 256, 0, st>>>(ka.get(), va.get(), n);
    CAPSB_CUDA(cudaStreamSynchronize(st));
    rs.timer.enabled = true;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    const char* names[] = {"512x8 prefetch (product)", "512x8", "256x16", "256x16 prefetch", "1024x8 tile 8192", "512x16 tile 8192"};
    for (int variant = 0; variant < 6; ++variant) {
      if (argc > 4 && atoi(argv[4]) != variant) continue;
      rs.variant = variant;
      for (int w = 0; w < 3; ++w)
        radix_pass<uint64_t, uint32_t>(st, rs, ArraySource<uint64_t, uint32_t>{ka.get(), va.get()}, n, shift, kb.get(),
                                       vb.get());
      CAPSB_CUDA(cudaStreamSynchronize(st));
      rs.timer.reset();
      cudaEventRecord(e0, st);
      for (int r = 0; r < reps; ++r)
        radix_pass<uint64_t, uint32_t>(st, rs, ArraySource<uint64_t, uint32_t>{ka.get(), va.get()}, n, shift, kb.get(),
                                       vb.get());
      cudaEventRecord(e1, st);
      CAPSB_CUDA(cudaStreamSynchronize(st));
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      uint32_t launches = 0;
      const float scatter_ms = rs.timer.drain(&launches);
      const double bytes = 24.0 * n;
      printf("n=%llu variant %d (%s) grid_per_sm=%u: pass %.4f ms (%.1f GB/s incl. hist), scatter %.4f ms (%.1f GB/s, %.3f of 6546)\n",
             (unsigned long long)n, variant, names[variant], rs.max_blocks / dev.sm_count, ms / reps,
             (bytes + 8.0 * n) / (ms / reps) / 1e6, scatter_ms / launches, bytes / (scatter_ms / launches) / 1e6,
             bytes / (scatter_ms / launches) / 1e6 / 6546.0);
    }
    // (u32 key, u32 value) records through the product layout: what a pass would cost if the four
    // upper passes of the suffix sort carried a 32-bit high key
    if (argc <= 4) {
      DevBuf<uint32_t> k32a(n, st), k32b(n + (8u << 20), st);
      {
        uint32_t* d = k32a.get();
        const uint64_t* src = ka.get();
        launch_map(dev, st, n, [=] __device__(uint64_t i) { d[i] = static_cast<uint32_t>(src[i] >> 32); });
      }
      rs.variant = 0;
      for (int w = 0; w < 3; ++w)
        radix_pass<uint32_t, uint32_t>(st, rs, ArraySource<uint32_t, uint32_t>{k32a.get(), va.get()}, n, shift, k32b.get(),
                                       vb.get());
      CAPSB_CUDA(cudaStreamSynchronize(st));
      rs.timer.reset();
      for (int r = 0; r < reps; ++r)
        radix_pass<uint32_t, uint32_t>(st, rs, ArraySource<uint32_t, uint32_t>{k32a.get(), va.get()}, n, shift, k32b.get(),
                                       vb.get());
      CAPSB_CUDA(cudaStreamSynchronize(st));
      uint32_t launches = 0;
      const float scatter_ms = rs.timer.drain(&launches);
      printf("n=%llu (u32,u32) records, product layout: scatter %.4f ms (%.1f GB/s of 16 B/element)\n",
             (unsigned long long)n, scatter_ms / launches, 16.0 * n / (scatter_ms / launches) / 1e6);
      // leave the (u64,u32) output of the last variant in kb/vb for the check below
      rs.variant = 5;
      radix_pass<uint64_t, uint32_t>(st, rs, ArraySource<uint64_t, uint32_t>{ka.get(), va.get()}, n, shift, kb.get(), vb.get());
      CAPSB_CUDA(cudaStreamSynchronize(st));
    }
    // correctness of the last variant run: output sorted by the digit, stable
    {
      std::vector<uint64_t> hk(1 << 20);
      std::vector<uint32_t> hv(1 << 20);
      const uint64_t take = n < hk.size() ? n : hk.size();
      cudaMemcpy(hk.data(), kb.get(), take * 8, cudaMemcpyDeviceToHost);
      cudaMemcpy(hv.data(), vb.get(), take * 4, cudaMemcpyDeviceToHost);
      bool ok = true;
      for (uint64_t i = 1; i < take; ++i) {
        const unsigned a = (hk[i - 1] >> shift) & 255, b = (hk[i] >> shift) & 255;
        if (a > b || (a == b && hv[i - 1] >= hv[i])) ok = false;
      }
      printf("first %llu outputs sorted and stable: %s\n", (unsigned long long)take, ok ? "yes" : "NO");
    }
  } catch (const std::exception& e) {
    fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  return 0;
}
