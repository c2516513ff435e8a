// Micro-benchmark (not product code): cost of the warp-ranking primitives on one SM under load.
// Each CTA of 256 threads runs ITER rounds of one primitive on pseudo-random 8-bit digits;
// reports SM cycles per warp-row (32 elements) with `ctas_per_sm` CTAs resident.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ITER = 4096;

template <int MODE>
__global__ void __launch_bounds__(256) bench(unsigned* out, unsigned long long* cycles) {
  __shared__ unsigned tab[8][257];
  __shared__ unsigned cnt[8][257];
  const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < 8 * 257; i += 256) (&tab[0][0])[i] = 0, (&cnt[0][0])[i] = 0;
  __syncthreads();
  unsigned x = tid * 2654435761u + blockIdx.x * 40503u + 12345u;
  unsigned acc = 0;
  const unsigned lt = (1u << lane) - 1u;
  const long long t0 = clock64();
  for (int it = 0; it < ITER; ++it) {
    x = x * 1664525u + 1013904223u;
    const unsigned d = x >> 24;
    if (MODE == 0) {  // atomicOr (no return) + read back
      atomicOr(&tab[warp][d], 1u << lane);
      __syncwarp();
      acc += tab[warp][d];
      __syncwarp();
    } else if (MODE == 1) {  // plain store + read back
      tab[warp][d] = lane;
      __syncwarp();
      acc += tab[warp][d];
      __syncwarp();
    } else if (MODE == 2) {  // atomicAdd with return
      acc += atomicAdd(&cnt[warp][d], 1u);
    } else if (MODE == 3) {  // match_any
      const unsigned peers = __match_any_sync(0xffffffffu, d);
      acc += __popc(peers & lt);
    } else if (MODE == 4) {  // 8 ballots
      unsigned peers = 0xffffffffu;
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        const unsigned v = __ballot_sync(0xffffffffu, (d >> b) & 1u);
        peers &= ((d >> b) & 1u) ? v : ~v;
      }
      acc += __popc(peers & lt);
    } else if (MODE == 5) {  // the full current ranking step (or, read, leader add, shfl)
      atomicOr(&tab[warp][d], 1u << lane);
      __syncwarp();
      const unsigned peers = tab[warp][d];
      __syncwarp();
      const int leader = __ffs((int)peers) - 1;
      unsigned before = 0;
      if ((int)lane == leader) {
        tab[warp][d] = 0;
        before = atomicAdd(&cnt[warp][d], __popc(peers));
      }
      before = __shfl_sync(0xffffffffu, before, leader);
      acc += before + __popc(peers & lt);
      __syncwarp();
    } else if (MODE == 6) {  // same with a plain read-modify-write by the leader
      atomicOr(&tab[warp][d], 1u << lane);
      __syncwarp();
      const unsigned peers = tab[warp][d];
      __syncwarp();
      const int leader = __ffs((int)peers) - 1;
      unsigned before = 0;
      if ((int)lane == leader) {
        tab[warp][d] = 0;
        before = cnt[warp][d];
        cnt[warp][d] = before + __popc(peers);
      }
      before = __shfl_sync(0xffffffffu, before, leader);
      acc += before + __popc(peers & lt);
      __syncwarp();
    } else if (MODE == 7) {  // byte counters private to the lane: [digit][lane]
      unsigned char* c = reinterpret_cast<unsigned char*>(&tab[0][0]) ;  // 8 KB per warp needs 64 KB: use modulo for timing
      unsigned char* p = c + ((warp * 8192u + d * 32u + lane) & 8191u);
      const unsigned before = *p;
      *p = (unsigned char)(before + 1);
      acc += before;
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * 256 + tid] = acc;
  if (tid == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
}

template <int MODE>
void run(const char* name, int ctas_per_sm, int sms) {
  unsigned* out;
  unsigned long long* cyc;
  const int grid = sms * ctas_per_sm;
  cudaMalloc(&out, grid * 256 * 4);
  cudaMalloc(&cyc, grid * 8);
  bench<MODE><<<grid, 256>>>(out, cyc);
  bench<MODE><<<grid, 256>>>(out, cyc);
  cudaDeviceSynchronize();
  unsigned long long* h = new unsigned long long[grid];
  cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < grid; ++i) avg += h[i];
  avg /= grid;
  // per SM: ctas_per_sm * 8 warps * ITER rows in `avg` cycles
  printf("%-44s ctas/sm %d: %.1f cycles per warp-row per warp, %.2f SM-cycles per row (%.3f per element)\n", name,
         ctas_per_sm, avg / ITER, avg / ITER / (ctas_per_sm * 8), avg / ITER / (ctas_per_sm * 8) / 32);
  cudaFree(out), cudaFree(cyc);
  delete[] h;
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  for (int c : {1, 2, 4}) {
    run<0>("atomicOr + read back", c, sms);
    run<1>("plain store + read back", c, sms);
    run<2>("atomicAdd with return", c, sms);
    run<3>("match_any", c, sms);
    run<4>("8 ballots", c, sms);
    run<5>("full step: or, read, leader atomicAdd, shfl", c, sms);
    run<6>("full step with plain leader update", c, sms);
    run<7>("lane-private byte counter ++", c, sms);
  }
  return 0;
}
