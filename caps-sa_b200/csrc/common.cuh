// Shared definitions for the CaPS-SA B200 construction engine (sm_100a only).
//
// Conventions
//   * every kernel launch goes through CAPSB_LAUNCH so launches are counted
//     (bench.py reports the count as "gpu_launches");
//   * every CUDA call goes through CAPSB_CUDA, which throws capsb::Error; the C-ABI
//     layer (capi.cu) turns that into a non-zero return code + caps_sa_gpu_last_error();
//   * device scratch comes from the stream-ordered pool (cudaMallocAsync) through DevBuf;
//   * streaming kernels use a fixed grid (a multiple of the SM count) and walk
//     contiguous chunks, so per-block partials stay small and scans are 3 short kernels.
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <utility>

namespace capsb {

struct Error : std::runtime_error {
  using std::runtime_error::runtime_error;
};

[[noreturn]] inline void fail(const std::string& what) { throw Error(what); }

#define CAPSB_CUDA(expr)                                                                   \
  do {                                                                                     \
    cudaError_t err__ = (expr);                                                            \
    if (err__ != cudaSuccess) {                                                            \
      ::capsb::fail(std::string("CUDA error: ") + cudaGetErrorString(err__) + " at " +     \
                    __FILE__ + ":" + std::to_string(__LINE__) + " (" #expr ")");           \
    }                                                                                      \
  } while (0)

extern std::atomic<uint64_t> g_kernel_launches;

#define CAPSB_LAUNCH(kernel, grid, block, smem, stream, ...)                               \
  do {                                                                                     \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                            \
    ::capsb::g_kernel_launches.fetch_add(1, std::memory_order_relaxed);                    \
    CAPSB_CUDA(cudaGetLastError());                                                        \
  } while (0)

// ---------------------------------------------------------------------------------------
// Device properties / grid sizing
// ---------------------------------------------------------------------------------------
struct DeviceInfo {
  int device = 0;
  int sm_count = 148;
};

inline uint64_t ceil_div(uint64_t a, uint64_t b) { return (a + b - 1) / b; }

// Fixed-grid chunking: `blocks` CTAs, each owning `chunk` consecutive elements
// (chunk is a multiple of `tile`).
struct Chunking {
  unsigned blocks;
  uint64_t chunk;
};

inline Chunking make_chunking(uint64_t n, unsigned tile, unsigned max_blocks) {
  uint64_t tiles = ceil_div(n ? n : 1, tile);
  unsigned blocks = static_cast<unsigned>(tiles < max_blocks ? tiles : max_blocks);
  uint64_t tiles_per_block = ceil_div(tiles, blocks);
  blocks = static_cast<unsigned>(ceil_div(tiles, tiles_per_block));
  return {blocks, tiles_per_block * tile};
}

// ---------------------------------------------------------------------------------------
// Stream-ordered device buffer
// ---------------------------------------------------------------------------------------
template <class T>
class DevBuf {
 public:
  DevBuf() = default;
  DevBuf(uint64_t count, cudaStream_t stream) { alloc(count, stream); }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept { *this = std::move(o); }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) {
      release();
      ptr_ = o.ptr_, count_ = o.count_, stream_ = o.stream_;
      o.ptr_ = nullptr, o.count_ = 0;
    }
    return *this;
  }
  ~DevBuf() { release(); }

  void alloc(uint64_t count, cudaStream_t stream) {
    release();
    stream_ = stream;
    count_ = count;
    if (count == 0) return;
    void* p = nullptr;
    cudaError_t err = cudaMallocAsync(&p, count * sizeof(T), stream);
    if (err != cudaSuccess) {
      cudaGetLastError();
      fail("out of device memory: cudaMallocAsync(" + std::to_string(count * sizeof(T)) +
           " bytes) failed: " + cudaGetErrorString(err));
    }
    ptr_ = static_cast<T*>(p);
  }
  void release() {
    if (ptr_) cudaFreeAsync(ptr_, stream_);
    ptr_ = nullptr;
    count_ = 0;
  }
  T* get() const { return ptr_; }
  uint64_t size() const { return count_; }
  explicit operator bool() const { return ptr_ != nullptr; }

 private:
  T* ptr_ = nullptr;
  uint64_t count_ = 0;
  cudaStream_t stream_ = nullptr;
};

// ---------------------------------------------------------------------------------------
// Small device helpers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

// Streaming (read-once) 64-bit load that does not allocate in L1.
__device__ __forceinline__ uint64_t ld_stream_u64(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}

// ---------------------------------------------------------------------------------------
// Generic element-wise kernel: f(i) for i in [0, n), grid-stride.
// ---------------------------------------------------------------------------------------
template <class F>
__global__ void __launch_bounds__(256) map_kernel(uint64_t n, F f) {
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    f(i);
}

template <class F>
inline void launch_map(const DeviceInfo& dev, cudaStream_t stream, uint64_t n, F f) {
  if (n == 0) return;
  const uint64_t want = ceil_div(n, 256);
  const uint64_t cap = static_cast<uint64_t>(dev.sm_count) * 16;
  const unsigned grid = static_cast<unsigned>(want < cap ? want : cap);
  CAPSB_LAUNCH((map_kernel<F>), grid, 256, 0, stream, n, f);
}

// ---------------------------------------------------------------------------------------
// Device-wide scans (3 short kernels over a fixed grid).
//   out[i] = op-scan of in(i); `in` and `out` are functors so producers/consumers fuse.
// ---------------------------------------------------------------------------------------
struct OpSum {
  template <class T>
  __host__ __device__ static T identity() { return T(0); }
  template <class T>
  __host__ __device__ static T apply(T a, T b) { return a + b; }
};
struct OpMax {
  template <class T>
  __host__ __device__ static T identity() { return T(0); }
  template <class T>
  __host__ __device__ static T apply(T a, T b) { return a > b ? a : b; }
};

constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

template <class T, class Op>
__device__ __forceinline__ T warp_scan_inclusive(T v) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    T o = __shfl_up_sync(0xffffffffu, v, d);
    if (lane_id() >= static_cast<unsigned>(d)) v = Op::template apply<T>(o, v);
  }
  return v;
}

// Scan of one value per thread across a 256-thread block.
// Returns the exclusive prefix; *inclusive and *block_total are filled in.
template <class T, class Op>
__device__ __forceinline__ T block_scan(T v, T* inclusive, T* block_total, T* smem /*[8]*/) {
  const unsigned warp = threadIdx.x >> 5;
  const T inc = warp_scan_inclusive<T, Op>(v);
  T lane_excl = __shfl_up_sync(0xffffffffu, inc, 1);
  if (lane_id() == 0) lane_excl = Op::template identity<T>();
  __syncthreads();  // guards reuse of smem across calls
  if (lane_id() == 31) smem[warp] = inc;
  __syncthreads();
  T prefix = Op::template identity<T>();
  T total = Op::template identity<T>();
#pragma unroll
  for (int w = 0; w < kScanThreads / 32; ++w) {
    const T s = smem[w];
    if (static_cast<unsigned>(w) < warp) prefix = Op::template apply<T>(prefix, s);
    total = Op::template apply<T>(total, s);
  }
  *block_total = total;
  *inclusive = Op::template apply<T>(prefix, inc);
  return Op::template apply<T>(prefix, lane_excl);
}

template <class T, class Op, class In>
__global__ void __launch_bounds__(kScanThreads) scan_reduce_kernel(uint64_t n, uint64_t chunk, In in,
                                                                   T* partial) {
  __shared__ T smem[kScanThreads / 32];
  const uint64_t begin = static_cast<uint64_t>(blockIdx.x) * chunk;
  const uint64_t end = begin + chunk < n ? begin + chunk : n;
  T acc = Op::template identity<T>();
  for (uint64_t i = begin + threadIdx.x; i < end; i += kScanThreads)
    acc = Op::template apply<T>(acc, in(i));
  T inc, total;
  block_scan<T, Op>(acc, &inc, &total, smem);
  if (threadIdx.x == 0) partial[blockIdx.x] = total;
}

// Single block: exclusive scan of partial[0..count) in place; total -> *total_out (optional).
template <class T, class Op>
__global__ void __launch_bounds__(kScanThreads) scan_spine_kernel(unsigned count, T* partial,
                                                                  T* total_out) {
  __shared__ T smem[kScanThreads / 32];
  T carry = Op::template identity<T>();
  for (unsigned base = 0; base < count; base += kScanThreads) {
    const unsigned i = base + threadIdx.x;
    const T v = i < count ? partial[i] : Op::template identity<T>();
    T inc, total;
    const T excl = block_scan<T, Op>(v, &inc, &total, smem);
    if (i < count) partial[i] = Op::template apply<T>(carry, excl);
    carry = Op::template apply<T>(carry, total);
  }
  if (threadIdx.x == 0 && total_out) *total_out = carry;
}

template <class T, class Op, bool Inclusive, class In, class Out>
__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(uint64_t n, uint64_t chunk, In in,
                                                                  Out out, const T* partial) {
  __shared__ T smem[kScanThreads / 32];
  const uint64_t begin = static_cast<uint64_t>(blockIdx.x) * chunk;
  const uint64_t end = begin + chunk < n ? begin + chunk : n;
  T carry = partial[blockIdx.x];
  for (uint64_t tile = begin; tile < end; tile += kScanTile) {
    // blocked arrangement: thread t owns items [t*kScanItems, (t+1)*kScanItems)
    const uint64_t first = tile + static_cast<uint64_t>(threadIdx.x) * kScanItems;
    T vals[kScanItems];
    T local = Op::template identity<T>();
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
      const uint64_t i = first + k;
      vals[k] = i < end ? in(i) : Op::template identity<T>();
      local = Op::template apply<T>(local, vals[k]);
    }
    T inc, total;
    const T excl = block_scan<T, Op>(local, &inc, &total, smem);
    T run = Op::template apply<T>(carry, excl);
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
      const uint64_t i = first + k;
      const T next = Op::template apply<T>(run, vals[k]);
      if (i < end) out(i, Inclusive ? next : run);
      run = next;
    }
    carry = Op::template apply<T>(carry, total);
  }
}

// Scratch for scans (per-block partials); owned by the pipeline context.
template <class T>
struct ScanScratch {
  DevBuf<T> partial;
  DevBuf<T> total;
  unsigned max_blocks = 0;
  void init(const DeviceInfo& dev, cudaStream_t stream) {
    max_blocks = static_cast<unsigned>(dev.sm_count) * 8;
    partial.alloc(max_blocks, stream);
    total.alloc(1, stream);
  }
};

// out(i, scan) for i in [0, n).  If total_host != nullptr the grand total is copied back
// (synchronises the stream).
template <class T, class Op, bool Inclusive, class In, class Out>
inline void device_scan(const DeviceInfo& dev, cudaStream_t stream, ScanScratch<T>& scratch,
                        uint64_t n, In in, Out out, T* total_host = nullptr) {
  if (n == 0) {
    if (total_host) *total_host = Op::template identity<T>();
    return;
  }
  (void)dev;
  const Chunking ck = make_chunking(n, kScanTile, scratch.max_blocks);
  CAPSB_LAUNCH((scan_reduce_kernel<T, Op, In>), ck.blocks, kScanThreads, 0, stream, n, ck.chunk, in,
               scratch.partial.get());
  CAPSB_LAUNCH((scan_spine_kernel<T, Op>), 1, kScanThreads, 0, stream, ck.blocks,
               scratch.partial.get(), scratch.total.get());
  CAPSB_LAUNCH((scan_apply_kernel<T, Op, Inclusive, In, Out>), ck.blocks, kScanThreads, 0, stream, n,
               ck.chunk, in, out, scratch.partial.get());
  if (total_host) {
    CAPSB_CUDA(cudaMemcpyAsync(total_host, scratch.total.get(), sizeof(T), cudaMemcpyDeviceToHost,
                               stream));
    CAPSB_CUDA(cudaStreamSynchronize(stream));
  }
}

}  // namespace capsb
