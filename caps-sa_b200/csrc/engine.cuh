// Engine: per-device context (stream, scratch, statistics) and the stage entry points.
#pragma once

#include <algorithm>
#include <memory>
#include <vector>

#include "common.cuh"
#include "packed_text.cuh"
#include "radix_sort.cuh"
#include "msd_sort.cuh"

namespace capsb {

// Per-construction statistics, mirrored by caps_sa_gpu_stats in include/caps_sa_gpu.h.
struct Stats {
  uint64_t n = 0;
  uint32_t idx_bytes = 0;
  uint32_t bits_per_symbol = 0;
  uint32_t alphabet_size = 0;
  uint32_t refine_rounds = 0;
  uint64_t tied_after_key_sort = 0;  // suffixes in key groups of two or more
  uint64_t refine_sorted = 0;        // suffix-rounds ordered by the radix sort (large groups)
  uint64_t refine_counted = 0;       // suffix-rounds ordered by counting inside small groups
  uint64_t deep_lcp_direct = 0;      // irreducible deep LCPs computed by comparison
  uint64_t deep_lcp_long = 0;        // ... of which needed the block-wide compare
  uint64_t kernel_launches = 0;
  float ms_pack = 0, ms_sort = 0, ms_heads = 0, ms_refine = 0, ms_deep_lcp = 0, ms_total = 0;
  float ms_h2d = 0, ms_d2h = 0;
  uint32_t scatter_launches = 0;
  float ms_scatter = 0;
  uint64_t scatter_bytes = 0;
  uint32_t key_bits = 0;  // leading bits of the packed prefix the key sort used
  // packed-record MSD sort (msd_sort.cuh): digit widths, per-kernel-class times and algorithmic bytes
  uint32_t msd_a_bits = 0, msd_b_bits = 0;
  uint32_t msd_large_buckets = 0;  // buckets that fell back to the LSD sort
  uint64_t msd_large_records = 0;
  float ms_msd_scatter_a = 0, ms_msd_scatter_b = 0, ms_msd_local = 0, ms_msd_hist = 0;
  uint64_t msd_scatter_a_bytes = 0, msd_scatter_b_bytes = 0, msd_local_bytes = 0, msd_hist_bytes = 0;
  uint32_t msd_scatter_a_launches = 0, msd_scatter_b_launches = 0, msd_local_launches = 0, msd_hist_launches = 0;
  // sharded construction only
  float ms_partition = 0;     // pivots, pivot location and the (key, suffix) all-to-all
  float ms_merge = 0;         // merge-path tree over the received runs
  uint64_t comm_bytes = 0;    // bytes this rank moved to other ranks
  uint64_t shard_offset = 0;  // this rank owns SA/LCP positions [shard_offset, shard_offset + shard_count)
  uint64_t shard_count = 0;
  uint64_t pairs_chained = 0;  // groups of two finished by the pair-chain step
};

struct PackedTextBuf {
  DevBuf<uint64_t> words;
  uint64_t nwords = 0;
  unsigned log2_bits = 3;
  unsigned sigma = 0;
  PackedText view(uint64_t n) const { return PackedText{words.get(), n, log2_bits}; }
};

struct Comm;  // comm.cuh

// Result of a sharded construction on one rank: entries [offset, offset + count) of SA and LCP.
template <class IdxT>
struct ShardResult {
  DevBuf<IdxT> sa, lcp;
  uint64_t offset = 0, count = 0;
};

struct Engine {
  Arena arena;  // first member: outlives every buffer below
  DeviceInfo dev;
  cudaStream_t stream = nullptr;      // stream all work is issued on
  cudaStream_t own_stream = nullptr;  // created with the engine
  RadixScratch radix;
  MsdTimers msd_timers;
  ScanScratch<uint32_t> scan32;
  ScanScratch<uint64_t> scan64;
  Stats stats;
  std::vector<cudaEvent_t> events;
  // Host-buffer entry points: the suffix array is final once the ties are resolved, a stage
  // before the LCP array, so its device-to-host copy starts there, on copy_stream, and overlaps
  // the deep-LCP stage.  sa_sink = host array (indexed like the full SA), or null.
  cudaStream_t copy_stream = nullptr;
  void* sa_sink = nullptr;
  cudaEvent_t sa_sink_started = nullptr;  // recorded on copy_stream right before the copy (caller-owned)
  bool sa_sunk = false;
  // d_sa[0, count) = SA[first, first + count) is final
  void sa_is_final(const void* d_sa, uint64_t first, uint64_t count, size_t idx_bytes);
  // Single-GPU host-buffer construction: the path finishes SA and LCP a range of positions at a
  // time and copies every finished range to the caller's arrays while the next one is computed
  // (PCIe is the longest leg of the call).  The few entries that only the last stage settles — the
  // deep ties — are written afterwards, straight into the host arrays by a kernel, which needs
  // them mapped into the device's address space (pinned memory); lcp_sink_dev == nullptr turns the
  // streaming off and everything is copied once at the end.
  void* lcp_sink = nullptr;
  void* sa_sink_dev = nullptr;
  void* lcp_sink_dev = nullptr;
  bool results_streamed = false;  // every entry of SA and LCP has been sent to the host arrays
  bool can_stream() const { return sa_sink && lcp_sink && sa_sink_dev && lcp_sink_dev; }
  // entries [first, first + count) of both device arrays (indexed like the full arrays) are final
  void range_is_final(const void* d_sa, const void* d_lcp, uint64_t first, uint64_t count, size_t idx_bytes);
  // sharded construction (multi-process): the transport this engine joined and its last shard
  std::unique_ptr<Comm> comm;
  ShardResult<uint32_t> shard32;
  ShardResult<uint64_t> shard64;

  explicit Engine(int device);
  ~Engine();
  Engine(const Engine&) = delete;
  Engine& operator=(const Engine&) = delete;

  template <class T>
  ScanScratch<T>& scan_scratch();
};

template <>
inline ScanScratch<uint32_t>& Engine::scan_scratch<uint32_t>() { return scan32; }
template <>
inline ScanScratch<uint64_t>& Engine::scan_scratch<uint64_t>() { return scan64; }

// Temporarily run an engine on a caller-provided stream.
struct StreamScope {
  Engine& eng;
  cudaStream_t saved;
  StreamScope(Engine& e, cudaStream_t s) : eng(e), saved(e.stream) { eng.stream = s; }
  ~StreamScope() { eng.stream = saved; }
};

// text_pack.cu
PackedTextBuf pack_text(Engine& eng, const uint8_t* d_text, uint64_t n);
void map_acgt_device(Engine& eng, uint8_t* d_text, uint64_t n);

// Moves the per-kernel-class timings of the last construction into eng.stats (after the stream has
// been synchronised).
inline void collect_msd_timings(Engine& eng) {
  MsdTimers& t = eng.msd_timers;
  if (!t.enabled) return;
  eng.stats.msd_scatter_a_bytes = t.scatter_a.bytes;
  eng.stats.ms_msd_scatter_a = t.scatter_a.drain(&eng.stats.msd_scatter_a_launches);
  eng.stats.msd_scatter_b_bytes = t.scatter_b.bytes;
  eng.stats.ms_msd_scatter_b = t.scatter_b.drain(&eng.stats.msd_scatter_b_launches);
  eng.stats.msd_local_bytes = t.local.bytes;
  eng.stats.ms_msd_local = t.local.drain(&eng.stats.msd_local_launches);
  eng.stats.msd_hist_bytes = t.hist.bytes;
  eng.stats.ms_msd_hist = t.hist.drain(&eng.stats.msd_hist_launches);
}

// sa_build.cu — the construction path (reference construct(), src/Suffix_Array.cpp:466-494)
template <class IdxT>
void build_sa_lcp(Engine& eng, const uint8_t* d_text, uint64_t n, IdxT* d_sa, IdxT* d_lcp);

// sharded_build.cu — the same path with one rank per GPU; collective over the ranks of `comm`
template <class IdxT>
void build_sa_lcp_sharded(Engine& eng, Comm& comm, const uint8_t* d_text, uint64_t n, ShardResult<IdxT>& out);

// Test hook: the key sort of all suffixes (packed-record MSD sort, or the LSD passes when use_lsd);
// returns the key width.  Fills the sort-related statistics.
int stage_key_sort_u32(Engine& eng, const uint8_t* d_text, uint64_t n, bool use_lsd, uint64_t* d_keys, uint32_t* d_sa);

// Test hook: the two scan flavours the pipeline uses (inclusive max, exclusive sum).
void stage_scan_u32(Engine& eng, const uint32_t* d_in, uint32_t* d_out, uint64_t n, bool inclusive_max);

}  // namespace capsb
