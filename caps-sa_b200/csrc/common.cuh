// Shared definitions for the CaPS-SA B200 construction engine (sm_100a only).
//
// Conventions
//   * every kernel launch goes through CAPSB_LAUNCH so launches are counted
//     (bench.py reports the count as "gpu_launches");
//   * every CUDA call goes through CAPSB_CUDA, which throws capsb::Error; the C-ABI
//     layer (capi.cu) turns that into a non-zero return code + caps_sa_gpu_last_error();
//   * device scratch comes from the engine's slab arena (Arena, below) through DevBuf;
//   * streaming kernels use a fixed grid (a multiple of the SM count) and walk
//     contiguous chunks, so per-block partials stay small and scans are 3 short kernels.
#pragma once

#include <cstring>
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <iterator>
#include <map>
#include <unordered_map>
#include <vector>
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
// Device memory arena
//
// All scratch of one engine comes from a few large cudaMalloc slabs that the engine keeps
// between constructions, sub-allocated first-fit with coalescing on the host.  Everything an
// engine does is ordered on one stream, so a block can be handed out again the moment it is
// freed.  (The stream-ordered pool of the driver, cudaMallocAsync, was measured to remap
// physical pages when multi-GB blocks of changing sizes are recycled: stages of the sharded
// path took 1-2 s instead of 50 ms — profiles/r01/README.md.)
// ---------------------------------------------------------------------------------------
class Arena {
 public:
  Arena() = default;
  Arena(const Arena&) = delete;
  Arena& operator=(const Arena&) = delete;
  ~Arena() { release_all(); }

  void* alloc(size_t bytes) {
    const size_t need = round_up(bytes ? bytes : 1, kAlign);
    if (void* p = take(need)) return p;
    add_slab(need);  // no hole fits
    void* p = take(need);
    if (!p) fail("internal: arena allocation failed after adding a slab");
    return p;
  }

  void free(void* p) {
    auto it = live_.find(p);
    if (it == live_.end()) return;
    const Live l = it->second;
    live_.erase(it);
    used_ -= l.size;
    Slab& slab = slabs_[l.slab];
    size_t off = static_cast<size_t>(static_cast<char*>(p) - slab.base), len = l.size;
    auto next = slab.free_blocks.lower_bound(off);
    if (next != slab.free_blocks.end() && off + len == next->first) {
      len += next->second;
      next = slab.free_blocks.erase(next);
    }
    if (next != slab.free_blocks.begin()) {
      auto prev = std::prev(next);
      if (prev->first + prev->second == off) {
        off = prev->first;
        len += prev->second;
        slab.free_blocks.erase(prev);
      }
    }
    slab.free_blocks.emplace(off, len);
  }

  // Returns every slab to the driver (all blocks must have been freed or are abandoned).
  void release_all() {
    for (Slab& slab : slabs_)
      if (slab.base) cudaFree(slab.base);
    slabs_.clear();
    live_.clear();
    reserved_ = used_ = 0;
  }

  size_t reserved() const { return reserved_; }
  size_t used() const { return used_; }

  // The arena DevBuf allocates from on this host thread (nullptr: the driver's stream-ordered pool).
  static Arena*& current() {
    static thread_local Arena* arena = nullptr;
    return arena;
  }

 private:
  static constexpr size_t kAlign = 512;
  static constexpr size_t kMinSlab = 32ull << 20;
  static constexpr size_t kSlabAlign = 2ull << 20;
  static size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

  struct Slab {
    char* base = nullptr;
    size_t size = 0;
    std::map<size_t, size_t> free_blocks;  // offset -> length
  };
  struct Live {
    size_t slab;
    size_t size;
  };

  // First fit over the slabs' free blocks.
  void* take(size_t need) {
    for (size_t s = 0; s < slabs_.size(); ++s) {
      Slab& slab = slabs_[s];
      for (auto it = slab.free_blocks.begin(); it != slab.free_blocks.end(); ++it) {
        if (it->second < need) continue;
        const size_t off = it->first, len = it->second;
        slab.free_blocks.erase(it);
        if (len > need) slab.free_blocks.emplace(off + need, len - need);
        void* p = slab.base + off;
        live_.emplace(p, Live{s, need});
        used_ += need;
        return p;
      }
    }
    return nullptr;
  }

  // A new slab for a request of `need` bytes (small requests share slabs of kMinSlab).
  void add_slab(size_t need) {
    const size_t slab_bytes = round_up(need > kMinSlab ? need : kMinSlab, kSlabAlign);
    void* base = nullptr;
    cudaError_t err = cudaMalloc(&base, slab_bytes);
    if (err != cudaSuccess && drop_empty_slabs()) {  // give unused slabs back and retry once
      cudaGetLastError();
      err = cudaMalloc(&base, slab_bytes);
    }
    if (err != cudaSuccess) {
      cudaGetLastError();
      fail("out of device memory: cudaMalloc(" + std::to_string(slab_bytes) + " bytes) failed: " +
           cudaGetErrorString(err) + " (arena holds " + std::to_string(reserved_) + " bytes, " +
           std::to_string(used_) + " in use)");
    }
    Slab slab;
    slab.base = static_cast<char*>(base);
    slab.size = slab_bytes;
    slab.free_blocks.emplace(0, slab_bytes);
    slabs_.push_back(std::move(slab));
    reserved_ += slab_bytes;
  }

  // Frees slabs that hold no live block (the vector keeps their slots: indices stay valid).
  bool drop_empty_slabs() {
    bool any = false;
    for (Slab& slab : slabs_) {
      if (slab.base && slab.free_blocks.size() == 1 && slab.free_blocks.begin()->second == slab.size) {
        cudaFree(slab.base);
        reserved_ -= slab.size;
        slab.base = nullptr;
        slab.size = 0;
        slab.free_blocks.clear();
        any = true;
      }
    }
    return any;
  }

  std::vector<Slab> slabs_;
  std::unordered_map<void*, Live> live_;
  size_t reserved_ = 0, used_ = 0;
};

struct ArenaScope {
  Arena* saved;
  explicit ArenaScope(Arena* a) : saved(Arena::current()) { Arena::current() = a; }
  ~ArenaScope() { Arena::current() = saved; }
  ArenaScope(const ArenaScope&) = delete;
  ArenaScope& operator=(const ArenaScope&) = delete;
};

// ---------------------------------------------------------------------------------------
// Device buffer: from the calling thread's arena if one is active, else stream-ordered
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
      ptr_ = o.ptr_, count_ = o.count_, stream_ = o.stream_, arena_ = o.arena_;
      o.ptr_ = nullptr, o.count_ = 0, o.arena_ = nullptr;
    }
    return *this;
  }
  ~DevBuf() { release(); }

  void alloc(uint64_t count, cudaStream_t stream) {
    release();
    stream_ = stream;
    count_ = count;
    if (count == 0) return;
    arena_ = Arena::current();
    if (arena_) {
      ptr_ = static_cast<T*>(arena_->alloc(count * sizeof(T)));
      return;
    }
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
    if (ptr_) {
      if (arena_)
        arena_->free(ptr_);
      else
        cudaFreeAsync(ptr_, stream_);
    }
    ptr_ = nullptr;
    count_ = 0;
    arena_ = nullptr;
  }
  T* get() const { return ptr_; }
  uint64_t size() const { return count_; }
  explicit operator bool() const { return ptr_ != nullptr; }

 private:
  T* ptr_ = nullptr;
  uint64_t count_ = 0;
  cudaStream_t stream_ = nullptr;
  Arena* arena_ = nullptr;
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
// Small device -> host reads (counts, totals) that do not use the copy engine: a one-warp
// kernel stores the words into mapped pinned host memory.  A cudaMemcpyAsync of a few bytes
// queues behind whatever the device-to-host engine is busy with — the multi-GB copy of the
// finished suffix array that overlaps the LCP stage (Engine::sa_is_final) stalled that stage
// for its whole duration this way.  Synchronises the stream, like the copies it replaces.
// ---------------------------------------------------------------------------------------
static __global__ void mailbox_kernel(const uint32_t* __restrict__ src, volatile uint32_t* dst, unsigned words) {
  for (unsigned i = threadIdx.x; i < words; i += blockDim.x) dst[i] = src[i];
}

inline void read_back(cudaStream_t stream, void* host_dst, const void* dev_src, size_t bytes) {
  constexpr size_t kMailboxBytes = 4096;
  if (bytes == 0) return;
  if (bytes > kMailboxBytes || (bytes & 3u) || (reinterpret_cast<uintptr_t>(dev_src) & 3u)) {
    CAPSB_CUDA(cudaMemcpyAsync(host_dst, dev_src, bytes, cudaMemcpyDeviceToHost, stream));
    CAPSB_CUDA(cudaStreamSynchronize(stream));
    return;
  }
  struct Box {  // one per host thread (a thread drives one engine at a time)
    void* p = nullptr;
    ~Box() {
      if (p) cudaFreeHost(p);
    }
  };
  thread_local Box box;
  if (!box.p) CAPSB_CUDA(cudaHostAlloc(&box.p, kMailboxBytes, cudaHostAllocPortable | cudaHostAllocMapped));
  CAPSB_LAUNCH(mailbox_kernel, 1, 32, 0, stream, static_cast<const uint32_t*>(dev_src),
               static_cast<volatile uint32_t*>(box.p), static_cast<unsigned>(bytes / 4));
  CAPSB_CUDA(cudaStreamSynchronize(stream));
  std::memcpy(host_dst, box.p, bytes);
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

// Two scans in one over 64-bit values: running maximum of the high halves, sum of the low halves
// (a sum that stays below 2^32).
struct OpMaxHiSumLo {
  template <class T>
  __host__ __device__ static T identity() { return T(0); }
  template <class T>
  __host__ __device__ static T apply(T a, T b) {
    static_assert(sizeof(T) == 8, "two 32-bit fields");
    const T ha = a >> 32, hb = b >> 32;
    return ((ha > hb ? ha : hb) << 32) | static_cast<uint32_t>(a + b);
  }
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
  uint64_t i = begin + threadIdx.x;
  for (; i + 3 * kScanThreads < end; i += 4 * kScanThreads) {  // four independent elements in flight
    const T a = in(i), b = in(i + kScanThreads), c = in(i + 2 * kScanThreads), d = in(i + 3 * kScanThreads);
    acc = Op::template apply<T>(Op::template apply<T>(acc, Op::template apply<T>(a, b)), Op::template apply<T>(c, d));
  }
  for (; i < end; i += kScanThreads) acc = Op::template apply<T>(acc, in(i));
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

// Shared-memory slot of tile element e: one pad word per 8 elements makes both access patterns
// conflict-free — striped (e = k*256 + t, the coalesced global order) and blocked (e = 8t + k,
// the order the scan needs).
__device__ __forceinline__ unsigned scan_slot(unsigned e) { return e + (e >> 3); }

// `in` and `out` are called in striped order (consecutive threads = consecutive elements) so the
// loads/stores inside the functors coalesce; the values cross to the blocked arrangement of the
// scan through shared memory.
template <class T, class Op, bool Inclusive, class In, class Out>
__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(uint64_t n, uint64_t chunk, In in,
                                                                  Out out, const T* partial) {
  __shared__ T smem[kScanThreads / 32];
  __shared__ T stage[kScanTile + kScanTile / 8];
  const uint64_t begin = static_cast<uint64_t>(blockIdx.x) * chunk;
  const uint64_t end = begin + chunk < n ? begin + chunk : n;
  T carry = partial[blockIdx.x];
  for (uint64_t tile = begin; tile < end; tile += kScanTile) {
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
      const unsigned e = static_cast<unsigned>(k) * kScanThreads + threadIdx.x;
      const uint64_t i = tile + e;
      stage[scan_slot(e)] = i < end ? in(i) : Op::template identity<T>();
    }
    __syncthreads();
    // blocked arrangement: thread t owns elements [t*kScanItems, (t+1)*kScanItems)
    T vals[kScanItems];
    T local = Op::template identity<T>();
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
      vals[k] = stage[scan_slot(threadIdx.x * kScanItems + k)];
      local = Op::template apply<T>(local, vals[k]);
    }
    T inc, total;
    const T excl = block_scan<T, Op>(local, &inc, &total, smem);
    T run = Op::template apply<T>(carry, excl);
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
      const T next = Op::template apply<T>(run, vals[k]);
      stage[scan_slot(threadIdx.x * kScanItems + k)] = Inclusive ? next : run;
      run = next;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
      const unsigned e = static_cast<unsigned>(k) * kScanThreads + threadIdx.x;
      const uint64_t i = tile + e;
      if (i < end) out(i, stage[scan_slot(e)]);
    }
    carry = Op::template apply<T>(carry, total);
    __syncthreads();  // the staging area is rewritten by the next tile
  }
}

// Stream compaction: out(i, slot) for every i in the CTA's chunk with flag(i) != 0, slot = number
// of flagged elements before i (partial[] holds the chunks' exclusive bases, as left by
// scan_reduce + scan_spine over the same flags).  Flags are ballots, not values moved through
// shared memory: per tile one ballot per row, an exclusive scan of the 64 (row, warp) counts by
// one warp, and a popc — the general scan_apply_kernel costs ~3x as much on sparse selections.
// other(i) is called for the elements that are not flagged.
struct NoOther {
  __device__ __forceinline__ void operator()(uint64_t) const {}
};
template <class T, class Flag, class Out, class Other>
__global__ void __launch_bounds__(kScanThreads) select_apply_kernel(uint64_t n, uint64_t chunk, Flag flag, Out out,
                                                                    Other other, const T* partial) {
  static_assert(kScanItems * (kScanThreads / 32) == 64, "the (row, warp) counts are scanned by one warp, two each");
  __shared__ unsigned counts[kScanItems * (kScanThreads / 32)];
  __shared__ unsigned tile_total;
  constexpr int kWarps = kScanThreads / 32;
  const uint64_t begin = static_cast<uint64_t>(blockIdx.x) * chunk;
  const uint64_t end = begin + chunk < n ? begin + chunk : n;
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const unsigned lt = lanemask_lt();
  T carry = partial[blockIdx.x];
  for (uint64_t tile = begin; tile < end; tile += kScanTile) {
    unsigned votes[kScanItems];
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
      const uint64_t i = tile + static_cast<unsigned>(k) * kScanThreads + threadIdx.x;
      // flag() of a clamped index rather than a branch around it: the rows' loads stay independent
      const bool f = (flag(i < end ? i : end - 1) != 0) & (i < end);
      votes[k] = __ballot_sync(0xffffffffu, f);
      if (lane == 0) counts[k * kWarps + warp] = __popc(votes[k]);
    }
    __syncthreads();
    if (warp == 0) {  // exclusive scan of the 64 counts in (row, warp) order = element order
      const unsigned a = counts[2 * lane], b = counts[2 * lane + 1];
      unsigned inc = a + b;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const unsigned o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= static_cast<unsigned>(d)) inc += o;
      }
      counts[2 * lane] = inc - a - b;
      counts[2 * lane + 1] = inc - b;
      if (lane == 31) tile_total = inc;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
      const uint64_t i = tile + static_cast<unsigned>(k) * kScanThreads + threadIdx.x;
      if (votes[k] >> lane & 1u)
        out(i, static_cast<T>(carry + counts[k * kWarps + warp] + __popc(votes[k] & lt)));
      else if (i < end)
        other(i);
    }
    carry += tile_total;
    __syncthreads();  // counts are rewritten by the next tile
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
    read_back(stream, total_host, scratch.total.get(), sizeof(T));
  }
}

}  // namespace capsb
