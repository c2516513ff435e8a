// Text staging and key packing (north-star subsystem 1; replaces the CLI's byte-mapping
// loop, reference src/main.cpp:61-70, and prepares the operand of every later kernel).
//
//   alphabet_scan_kernel : which of the 256 byte values occur (256-bit presence mask)
//   pack_kernel<LOG2BITS>: byte -> dense order-preserving code -> MSB-first 64-bit words,
//                          one output word per thread, 128-bit input loads
//   map_acgt_kernel      : the CLI's in-place mapping "ACTG"[(c & 6) >> 1] on the device
#include "engine.cuh"

namespace capsb {

namespace {

__global__ void __launch_bounds__(256) alphabet_scan_kernel(const uint8_t* __restrict__ text,
                                                            uint64_t n, uint32_t* present /*[8]*/) {
  // one flag byte per byte value in shared memory: a plain store per input byte (lanes that
  // hit the same word merge), no read-modify-write and no per-thread bitmap arithmetic
  __shared__ uint8_t seen[256];
  seen[threadIdx.x] = 0;
  __syncthreads();
  const uint64_t gtid = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const uint64_t gsize = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  const bool aligned = (reinterpret_cast<uintptr_t>(text) & 15u) == 0;
  uint64_t vec_end = 0;
  if (aligned) {
    const uint64_t nvec = n / 16;
    vec_end = nvec * 16;
    const uint4* v = reinterpret_cast<const uint4*>(text);
    for (uint64_t i = gtid; i < nvec; i += 2 * gsize) {
      const uint64_t j = i + gsize;
      const uint4 q0 = __ldg(v + i);
      const uint4 q1 = __ldg(v + (j < nvec ? j : i));
      const uint32_t w[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
      for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) seen[(w[a] >> (8 * b)) & 0xFFu] = 1;
    }
  }
  for (uint64_t i = vec_end + gtid; i < n; i += gsize) seen[text[i]] = 1;
  __syncthreads();
  const unsigned mine = seen[threadIdx.x] ? 1u : 0u;
  const unsigned word = __ballot_sync(0xffffffffu, mine);  // warp w covers byte values 32w .. 32w+31
  if (lane_id() == 0 && word) atomicOr(present + (threadIdx.x >> 5), word);
}

// One thread builds one 64-bit output word from S = 64 >> LOG2BITS input bytes.
template <int LOG2BITS>
__global__ void __launch_bounds__(256) pack_kernel(const uint8_t* __restrict__ text, uint64_t n,
                                                   const uint8_t* __restrict__ lut_global,
                                                   uint64_t* __restrict__ words, uint64_t nwords_total) {
  constexpr unsigned BITS = 1u << LOG2BITS;
  constexpr unsigned S = 64u >> LOG2BITS;  // input bytes per output word: 64, 32, 16, 8
  __shared__ uint8_t lut[256];
  lut[threadIdx.x] = lut_global[threadIdx.x];
  __syncthreads();
  const bool aligned = (reinterpret_cast<uintptr_t>(text) & 15u) == 0;
  const uint64_t gsize = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  for (uint64_t w = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; w < nwords_total;
       w += gsize) {
    const uint64_t first = w * S;
    uint64_t acc = 0;
    if (first + S <= n && aligned && S >= 16) {
      const uint4* v = reinterpret_cast<const uint4*>(text + first);
#pragma unroll
      for (unsigned q = 0; q < S / 16; ++q) {
        const uint4 x = __ldg(v + q);
        const uint32_t part[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b)
            acc = (acc << BITS) | lut[(part[a] >> (8 * b)) & 0xFFu];
      }
    } else if (first + S <= n && aligned && S == 8) {
      const uint2 x = __ldg(reinterpret_cast<const uint2*>(text + first));
      const uint32_t part[2] = {x.x, x.y};
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc = (acc << BITS) | lut[(part[a] >> (8 * b)) & 0xFFu];
    } else {
      // tail word, padding words, or unaligned input: byte loads, zero padding
      for (unsigned k = 0; k < S; ++k) {
        const uint64_t i = first + k;
        const uint64_t code = i < n ? lut[text[i]] : 0u;
        acc = (acc << BITS) | code;
      }
    }
    words[w] = acc;
  }
}

__global__ void __launch_bounds__(256) map_acgt_kernel(uint8_t* text, uint64_t n) {
  // "ACTG"[(toupper(c) & 6) >> 1] — toupper never touches bits 1-2 (reference src/main.cpp:68)
  const uint32_t table = ('A') | ('C' << 8) | ('T' << 16) | ('G' << 24);
  const uint64_t gtid = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const uint64_t gsize = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  const bool aligned = (reinterpret_cast<uintptr_t>(text) & 15u) == 0;
  uint64_t vec_end = 0;
  if (aligned) {
    const uint64_t nvec = n / 16;
    vec_end = nvec * 16;
    uint4* v = reinterpret_cast<uint4*>(text);
    for (uint64_t i = gtid; i < nvec; i += gsize) {
      uint4 q = v[i];
      uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        uint32_t r = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const uint32_t c = (w[a] >> (8 * b)) & 0xFFu;
          r |= ((table >> (8 * ((c >> 1) & 3u))) & 0xFFu) << (8 * b);
        }
        w[a] = r;
      }
      v[i] = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
  for (uint64_t i = vec_end + gtid; i < n; i += gsize) {
    const uint32_t c = text[i];
    text[i] = static_cast<uint8_t>((table >> (8 * ((c >> 1) & 3u))) & 0xFFu);
  }
}

}  // namespace

void map_acgt_device(Engine& eng, uint8_t* d_text, uint64_t n) {
  if (n == 0) return;
  const unsigned grid = static_cast<unsigned>(std::min<uint64_t>(ceil_div(ceil_div(n, 16), 256),
                                                                 static_cast<uint64_t>(eng.dev.sm_count) * 8));
  CAPSB_LAUNCH(map_acgt_kernel, grid ? grid : 1, 256, 0, eng.stream, d_text, n);
}

// Builds the order-preserving code table from the presence mask: codes are assigned in
// increasing `signed char` order (0x80..0xFF first, then 0x00..0x7F).
static unsigned build_code_table(const uint32_t present[8], uint8_t lut[256], unsigned* sigma_out) {
  unsigned next = 0;
  for (int k = 0; k < 256; ++k) {
    const unsigned byte = static_cast<unsigned>((k + 128) & 255);  // signed order
    const bool on = (present[byte >> 5] >> (byte & 31)) & 1u;
    lut[byte] = on ? static_cast<uint8_t>(next++) : 0;
  }
  *sigma_out = next;
  unsigned log2_bits = 0;  // bits per symbol in {1,2,4,8}
  while ((1u << (1u << log2_bits)) < next) ++log2_bits;
  return log2_bits;
}

PackedTextBuf pack_text(Engine& eng, const uint8_t* d_text, uint64_t n) {
  PackedTextBuf out;
  cudaStream_t st = eng.stream;
  DevBuf<uint32_t> present(8, st);
  CAPSB_CUDA(cudaMemsetAsync(present.get(), 0, 8 * sizeof(uint32_t), st));
  {
    const unsigned grid = static_cast<unsigned>(std::min<uint64_t>(
        std::max<uint64_t>(1, ceil_div(ceil_div(n, 16), 256)), static_cast<uint64_t>(eng.dev.sm_count) * 8));
    CAPSB_LAUNCH(alphabet_scan_kernel, grid, 256, 0, st, d_text, n, present.get());
  }
  uint32_t h_present[8];
  CAPSB_CUDA(cudaMemcpyAsync(h_present, present.get(), sizeof(h_present), cudaMemcpyDeviceToHost, st));
  CAPSB_CUDA(cudaStreamSynchronize(st));

  uint8_t lut[256];
  out.log2_bits = build_code_table(h_present, lut, &out.sigma);
  DevBuf<uint8_t> d_lut(256, st);
  CAPSB_CUDA(cudaMemcpyAsync(d_lut.get(), lut, 256, cudaMemcpyHostToDevice, st));

  const unsigned bits = 1u << out.log2_bits;
  out.nwords = ceil_div(n * bits, 64) + 2;
  out.words.alloc(out.nwords, st);
  const unsigned grid = static_cast<unsigned>(std::min<uint64_t>(ceil_div(out.nwords, 256),
                                                                 static_cast<uint64_t>(eng.dev.sm_count) * 16));
  switch (out.log2_bits) {
    case 0: CAPSB_LAUNCH(pack_kernel<0>, grid, 256, 0, st, d_text, n, d_lut.get(), out.words.get(), out.nwords); break;
    case 1: CAPSB_LAUNCH(pack_kernel<1>, grid, 256, 0, st, d_text, n, d_lut.get(), out.words.get(), out.nwords); break;
    case 2: CAPSB_LAUNCH(pack_kernel<2>, grid, 256, 0, st, d_text, n, d_lut.get(), out.words.get(), out.nwords); break;
    default: CAPSB_LAUNCH(pack_kernel<3>, grid, 256, 0, st, d_text, n, d_lut.get(), out.words.get(), out.nwords); break;
  }
  // d_lut is freed stream-ordered after the kernel; lut[] was copied synchronously enough:
  CAPSB_CUDA(cudaStreamSynchronize(st));  // lut is a stack array used by the async H2D copy
  return out;
}

}  // namespace capsb
