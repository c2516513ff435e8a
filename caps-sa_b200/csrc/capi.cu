// C-ABI layer (include/caps_sa_gpu.h): argument checks, host<->device staging, and the
// translation of capsb::Error into return codes.  No CPU fallback anywhere: without a CUDA
// device every entry point fails loudly.
#include <cstdlib>
#include <cstring>
#include <new>
#include <thread>

#include "../../include/caps_sa_gpu.h"
#include "comm.cuh"
#include "engine.cuh"

using capsb::Engine;

struct caps_sa_gpu_engine {
  Engine impl;
  explicit caps_sa_gpu_engine(int device) : impl(device) {}
};

namespace {

thread_local std::string g_last_error;
std::atomic<int> g_cli_byte_mapping{0};  // caps_sa_gpu_set_cli_byte_mapping

template <class F>
int guarded(F&& body) {
  try {
    g_last_error.clear();
    return body();
  } catch (const capsb::Error& e) {
    g_last_error = e.what();
    // leave the device usable for the next call if the error was recoverable
    cudaGetLastError();
    return CAPS_SA_GPU_ERR_CUDA;
  } catch (const std::bad_alloc&) {
    g_last_error = "host allocation failed";
    return CAPS_SA_GPU_ERR_CUDA;
  } catch (const std::exception& e) {
    g_last_error = e.what();
    return CAPS_SA_GPU_ERR_CUDA;
  }
}

int bad_args(const char* why) {
  g_last_error = why;
  return CAPS_SA_GPU_ERR_ARGS;
}

struct EventPair {
  cudaEvent_t a = nullptr, b = nullptr;
  EventPair() {
    CAPSB_CUDA(cudaEventCreate(&a));
    CAPSB_CUDA(cudaEventCreate(&b));
  }
  ~EventPair() {
    if (a) cudaEventDestroy(a);
    if (b) cudaEventDestroy(b);
  }
  float ms() {
    float v = 0;
    CAPSB_CUDA(cudaEventElapsedTime(&v, a, b));
    return v;
  }
};

// Bounded context (reference ctor argument 4, src/Suffix_Array.cpp:25).  The
// reference then compares suffixes on their first max_context + 1 symbols only (:72-77): its SA
// is sorted on that prefix with an order inside ties that depends on the subproblem count, and
// its LCP entries are min(lcp, max_context) (except at its p - 1 partition boundaries, which it
// patches with an unbounded compare, :440).  Here the construction is always exact — the exact
// order is one of the orders the bounded comparison allows — and the LCP array is clamped, so
// the documented properties hold for every subproblem count ("parity modulo ties").
template <class IdxT>
void clamp_lcp(Engine& eng, IdxT* d_lcp, uint64_t count, uint64_t max_context, uint64_t n) {
  if (max_context == 0 || max_context >= n || count == 0) return;
  const IdxT cap = static_cast<IdxT>(max_context);
  capsb::launch_map(eng.dev, eng.stream, count, [=] __device__(uint64_t k) {
    const IdxT v = d_lcp[k];
    if (v > cap) d_lcp[k] = cap;
  });
}

// Host-buffer entry points: while in scope, the engine copies the suffix array to `sa_out` as
// soon as it is final (Engine::sa_is_final); finish() copies the rest once the construction is
// done.  All copies run on the engine's copy stream.
template <class IdxT>
struct ResultCopy {
  Engine& eng;
  EventPair d2h;
  ResultCopy(Engine& e, IdxT* sa_out, IdxT* lcp_out = nullptr) : eng(e) {
    // CAPSB_EARLY_SA=0: measurement switch, both copies after the construction
    const char* env = std::getenv("CAPSB_EARLY_SA");
    const bool early = !(env && env[0] == '0');
    eng.sa_sink = early ? sa_out : nullptr;
    eng.lcp_sink = early ? lcp_out : nullptr;
    eng.sa_sink_dev = eng.lcp_sink_dev = nullptr;
    if (early && sa_out && lcp_out) {
      // streaming needs the caller's arrays in the device's address space (pinned host memory)
      void *sa_dev = nullptr, *lcp_dev = nullptr;
      if (cudaHostGetDevicePointer(&sa_dev, sa_out, 0) == cudaSuccess &&
          cudaHostGetDevicePointer(&lcp_dev, lcp_out, 0) == cudaSuccess) {
        eng.sa_sink_dev = sa_dev;
        eng.lcp_sink_dev = lcp_dev;
      } else {
        cudaGetLastError();
      }
    }
    eng.sa_sink_started = d2h.a;
    eng.sa_sunk = false;
    eng.results_streamed = false;
  }
  ~ResultCopy() {
    eng.sa_sink = eng.lcp_sink = nullptr;
    eng.sa_sink_dev = eng.lcp_sink_dev = nullptr;
    eng.sa_sink_started = nullptr;
    cudaStreamSynchronize(eng.copy_stream);  // no copy into the caller's arrays outlives the call
  }
  // d_sa / d_lcp [0, count) = entries [first, first + count) of the full arrays
  void finish(const IdxT* d_sa, const IdxT* d_lcp, IdxT* sa_out, IdxT* lcp_out, uint64_t first, uint64_t count,
              bool lcp_changed_after_build = false) {
    cudaStream_t cs = eng.copy_stream;
    eng.sa_sink = nullptr;
    CAPSB_CUDA(cudaEventRecord(d2h.b, eng.stream));
    CAPSB_CUDA(cudaStreamWaitEvent(cs, d2h.b, 0));
    const bool streamed = eng.results_streamed && !lcp_changed_after_build;
    if (!eng.sa_sunk) {
      CAPSB_CUDA(cudaEventRecord(d2h.a, cs));
      if (count)
        CAPSB_CUDA(cudaMemcpyAsync(sa_out + first, d_sa, count * sizeof(IdxT), cudaMemcpyDeviceToHost, cs));
    }
    if (count && !streamed)
      CAPSB_CUDA(cudaMemcpyAsync(lcp_out + first, d_lcp, count * sizeof(IdxT), cudaMemcpyDeviceToHost, cs));
    CAPSB_CUDA(cudaEventRecord(d2h.b, cs));
    CAPSB_CUDA(cudaStreamSynchronize(cs));
    CAPSB_CUDA(cudaStreamSynchronize(eng.stream));
    eng.stats.ms_d2h = d2h.ms();
  }
};

template <class IdxT>
int construct_host(caps_sa_gpu_engine* engine, const char* text, uint64_t n, IdxT* sa_out, IdxT* lcp_out,
                   uint64_t max_context) {
  if (!engine) return bad_args("engine is NULL");
  if (n > 0 && (!text || !sa_out || !lcp_out)) return bad_args("NULL buffer");
  if (sizeof(IdxT) == 4 && n > 0xFFFFFFFFull) return bad_args("n does not fit 32-bit indices");
  return guarded([&]() -> int {
    Engine& eng = engine->impl;
    capsb::ArenaScope arena_scope(&eng.arena);
    CAPSB_CUDA(cudaSetDevice(eng.dev.device));
    if (n == 0) return CAPS_SA_GPU_OK;
    cudaStream_t st = eng.stream;
    capsb::DevBuf<uint8_t> d_text(n, st);
    capsb::DevBuf<IdxT> d_sa(n, st), d_lcp(n, st);
    EventPair h2d;
    // (a bounded context clamps the LCP array after the construction: no streaming then)
    const bool clamps = max_context != 0 && max_context < n;
    ResultCopy<IdxT> results(eng, sa_out, clamps ? nullptr : lcp_out);
    CAPSB_CUDA(cudaEventRecord(h2d.a, st));
    CAPSB_CUDA(cudaMemcpyAsync(d_text.get(), text, n, cudaMemcpyHostToDevice, st));
    CAPSB_CUDA(cudaEventRecord(h2d.b, st));
    if (g_cli_byte_mapping.load()) capsb::map_acgt_device(eng, d_text.get(), n);
    capsb::build_sa_lcp<IdxT>(eng, d_text.get(), n, d_sa.get(), d_lcp.get());
    clamp_lcp<IdxT>(eng, d_lcp.get(), n, max_context, n);
    results.finish(d_sa.get(), d_lcp.get(), sa_out, lcp_out, 0, n);
    eng.stats.ms_h2d = h2d.ms();
    return CAPS_SA_GPU_OK;
  });
}

template <class IdxT>
int construct_device(caps_sa_gpu_engine* engine, const void* d_text, uint64_t n, IdxT* d_sa, IdxT* d_lcp,
                     void* stream) {
  if (!engine) return bad_args("engine is NULL");
  if (n > 0 && (!d_text || !d_sa || !d_lcp)) return bad_args("NULL buffer");
  if (sizeof(IdxT) == 4 && n > 0xFFFFFFFFull) return bad_args("n does not fit 32-bit indices");
  return guarded([&]() -> int {
    Engine& eng = engine->impl;
    capsb::ArenaScope arena_scope(&eng.arena);
    capsb::StreamScope scope(eng, stream ? static_cast<cudaStream_t>(stream) : eng.stream);
    capsb::build_sa_lcp<IdxT>(eng, static_cast<const uint8_t*>(d_text), n, d_sa, d_lcp);
    return CAPS_SA_GPU_OK;
  });
}

void fill_stats(const capsb::Stats& s, caps_sa_gpu_stats* out) {
  out->n = s.n;
  out->idx_bytes = s.idx_bytes;
  out->bits_per_symbol = s.bits_per_symbol;
  out->alphabet_size = s.alphabet_size;
  out->refine_rounds = s.refine_rounds;
  out->tied_after_key_sort = s.tied_after_key_sort;
  out->deep_lcp_direct = s.deep_lcp_direct;
  out->deep_lcp_long = s.deep_lcp_long;
  out->kernel_launches = s.kernel_launches;
  out->ms_pack = s.ms_pack, out->ms_sort = s.ms_sort, out->ms_heads = s.ms_heads;
  out->ms_refine = s.ms_refine, out->ms_deep_lcp = s.ms_deep_lcp, out->ms_total = s.ms_total;
  out->ms_h2d = s.ms_h2d, out->ms_d2h = s.ms_d2h;
  out->scatter_launches = s.scatter_launches;
  out->ms_scatter = s.ms_scatter;
  out->scatter_bytes = s.scatter_bytes;
  out->key_bits = s.key_bits;
  out->reserved = 0;
  out->ms_partition = s.ms_partition, out->ms_merge = s.ms_merge;
  out->comm_bytes = s.comm_bytes;
  out->shard_offset = s.shard_offset, out->shard_count = s.shard_count;
  out->pairs_chained = s.pairs_chained;
  out->msd_a_bits = s.msd_a_bits, out->msd_b_bits = s.msd_b_bits;
  out->msd_large_buckets = s.msd_large_buckets;
  out->msd_reserved = 0;
  out->msd_large_records = s.msd_large_records;
  out->ms_msd_scatter_a = s.ms_msd_scatter_a, out->ms_msd_scatter_b = s.ms_msd_scatter_b;
  out->ms_msd_local = s.ms_msd_local, out->ms_msd_hist = s.ms_msd_hist;
  out->msd_scatter_a_bytes = s.msd_scatter_a_bytes, out->msd_scatter_b_bytes = s.msd_scatter_b_bytes;
  out->msd_local_bytes = s.msd_local_bytes, out->msd_hist_bytes = s.msd_hist_bytes;
  out->msd_scatter_a_launches = s.msd_scatter_a_launches, out->msd_scatter_b_launches = s.msd_scatter_b_launches;
  out->msd_local_launches = s.msd_local_launches, out->msd_hist_launches = s.msd_hist_launches;
}

template <class IdxT>
capsb::ShardResult<IdxT>& shard_of(Engine& eng);
template <>
capsb::ShardResult<uint32_t>& shard_of<uint32_t>(Engine& eng) { return eng.shard32; }
template <>
capsb::ShardResult<uint64_t>& shard_of<uint64_t>(Engine& eng) { return eng.shard64; }


// Text staging of the sharded construction from a host buffer: every rank uploads only its own
// 1/world of the text over its own PCIe link and the ranks all-gather the pieces over NVLink
// (in place), instead of every rank pulling the whole text through PCIe.
struct StagedText {
  capsb::DevBuf<uint8_t> buf;  // piece * world bytes >= n
  float ms_h2d = 0;
};
StagedText stage_text_sharded(Engine& eng, capsb::Comm& comm, const char* text, uint64_t n) {
  cudaStream_t st = eng.stream;
  const uint64_t world = static_cast<uint64_t>(comm.world), rank = static_cast<uint64_t>(comm.rank);
  const uint64_t piece = (capsb::ceil_div(n, world) + 15) / 16 * 16;
  StagedText out;
  out.buf.alloc(piece * world, st);
  const uint64_t lo = rank * piece < n ? rank * piece : n;
  const uint64_t hi = (rank + 1) * piece < n ? (rank + 1) * piece : n;
  EventPair h2d;
  CAPSB_CUDA(cudaEventRecord(h2d.a, st));
  if (hi > lo) CAPSB_CUDA(cudaMemcpyAsync(out.buf.get() + lo, text + lo, hi - lo, cudaMemcpyHostToDevice, st));
  CAPSB_CUDA(cudaEventRecord(h2d.b, st));
  if (world > 1) comm.all_gather_device(out.buf.get() + rank * piece, out.buf.get(), piece, st);
  if (g_cli_byte_mapping.load()) capsb::map_acgt_device(eng, out.buf.get(), n);
  CAPSB_CUDA(cudaStreamSynchronize(st));
  out.ms_h2d = h2d.ms();
  return out;
}

// One process, one host thread per rank, peer copies between the ranks (ThreadComm).
template <class IdxT>
int construct_multi(const int* devices, int num_ranks, const char* text, uint64_t n, IdxT* sa_out, IdxT* lcp_out,
                    uint64_t max_context, caps_sa_gpu_stats* stats_out) {
  if (!devices || num_ranks < 1 || num_ranks > 64) return bad_args("bad device list");
  if (n > 0 && (!text || !sa_out || !lcp_out)) return bad_args("NULL buffer");
  if (sizeof(IdxT) == 4 && n > 0xFFFFFFFFull) return bad_args("n does not fit 32-bit indices");
  return guarded([&]() -> int {
    int visible = 0;
    CAPSB_CUDA(cudaGetDeviceCount(&visible));
    for (int r = 0; r < num_ranks; ++r)
      if (devices[r] < 0 || devices[r] >= visible)
        capsb::fail("no such CUDA device: " + std::to_string(devices[r]) + " (visible: " + std::to_string(visible) + ")");
    auto group = std::make_shared<capsb::ThreadGroup>(num_ranks);
    std::vector<capsb::Stats> stats(num_ranks);
    std::vector<std::thread> threads;
    for (int r = 0; r < num_ranks; ++r) {
      threads.emplace_back([&, r] {
        try {
          Engine eng(devices[r]);
          capsb::ArenaScope arena_scope(&eng.arena);
          capsb::ThreadComm comm(group, r, devices[r]);
          if (n) {
            cudaStream_t st = eng.stream;
            capsb::ShardResult<IdxT> shard;
            ResultCopy<IdxT> results(eng, sa_out);
            {
              StagedText staged = stage_text_sharded(eng, comm, text, n);
              capsb::build_sa_lcp_sharded<IdxT>(eng, comm, staged.buf.get(), n, shard);
              eng.stats.ms_h2d = staged.ms_h2d;
            }
            clamp_lcp<IdxT>(eng, shard.lcp.get(), shard.count, max_context, n);
            results.finish(shard.sa.get(), shard.lcp.get(), sa_out, lcp_out, shard.offset, shard.count);
          }
          stats[r] = eng.stats;
        } catch (const std::exception& e) {
          group->fail("rank " + std::to_string(r) + ": " + e.what());
        }
      });
    }
    for (std::thread& t : threads) t.join();
    if (group->failed()) capsb::fail(group->failure());
    if (stats_out)
      for (int r = 0; r < num_ranks; ++r) fill_stats(stats[r], stats_out + r);
    return CAPS_SA_GPU_OK;
  });
}

template <class IdxT>
int construct_sharded_device(caps_sa_gpu_engine* engine, const void* d_text, uint64_t n, void* stream) {
  if (!engine) return bad_args("engine is NULL");
  if (n > 0 && !d_text) return bad_args("NULL buffer");
  if (sizeof(IdxT) == 4 && n > 0xFFFFFFFFull) return bad_args("n does not fit 32-bit indices");
  if (!engine->impl.comm) return bad_args("the engine has not joined a communicator (caps_sa_gpu_engine_comm_init)");
  return guarded([&]() -> int {
    Engine& eng = engine->impl;
    capsb::ArenaScope arena_scope(&eng.arena);
    capsb::StreamScope scope(eng, stream ? static_cast<cudaStream_t>(stream) : eng.stream);
    capsb::build_sa_lcp_sharded<IdxT>(eng, *eng.comm, static_cast<const uint8_t*>(d_text), n, shard_of<IdxT>(eng));
    return CAPS_SA_GPU_OK;
  });
}

template <class IdxT>
int construct_sharded_host(caps_sa_gpu_engine* engine, const char* text, uint64_t n, IdxT* sa_out, IdxT* lcp_out) {
  if (!engine) return bad_args("engine is NULL");
  if (n > 0 && (!text || !sa_out || !lcp_out)) return bad_args("NULL buffer");
  if (sizeof(IdxT) == 4 && n > 0xFFFFFFFFull) return bad_args("n does not fit 32-bit indices");
  if (!engine->impl.comm) return bad_args("the engine has not joined a communicator (caps_sa_gpu_engine_comm_init)");
  return guarded([&]() -> int {
    Engine& eng = engine->impl;
    capsb::ArenaScope arena_scope(&eng.arena);
    CAPSB_CUDA(cudaSetDevice(eng.dev.device));
    if (n == 0) return CAPS_SA_GPU_OK;
    cudaStream_t st = eng.stream;
    capsb::ShardResult<IdxT>& shard = shard_of<IdxT>(eng);
    float ms_h2d = 0;
    ResultCopy<IdxT> results(eng, sa_out);
    {
      StagedText staged = stage_text_sharded(eng, *eng.comm, text, n);
      capsb::build_sa_lcp_sharded<IdxT>(eng, *eng.comm, staged.buf.get(), n, shard);
      ms_h2d = staged.ms_h2d;
    }
    results.finish(shard.sa.get(), shard.lcp.get(), sa_out, lcp_out, shard.offset, shard.count);
    eng.stats.ms_h2d = ms_h2d;
    return CAPS_SA_GPU_OK;
  });
}

}  // namespace

extern "C" {

int caps_sa_gpu_device_count(void) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return count;
}

const char* caps_sa_gpu_last_error(void) { return g_last_error.c_str(); }

caps_sa_gpu_engine* caps_sa_gpu_engine_create(int device) {
  caps_sa_gpu_engine* out = nullptr;
  const int rc = guarded([&]() -> int {
    int count = 0;
    CAPSB_CUDA(cudaGetDeviceCount(&count));
    if (device < 0 || device >= count)
      capsb::fail("no such CUDA device: " + std::to_string(device) + " (visible: " + std::to_string(count) + ")");
    out = new caps_sa_gpu_engine(device);
    return CAPS_SA_GPU_OK;
  });
  return rc == CAPS_SA_GPU_OK ? out : nullptr;
}

void caps_sa_gpu_engine_destroy(caps_sa_gpu_engine* engine) { delete engine; }

int caps_sa_gpu_engine_stats(const caps_sa_gpu_engine* engine, caps_sa_gpu_stats* out) {
  if (!engine || !out) return bad_args("NULL argument");
  fill_stats(engine->impl.stats, out);
  return CAPS_SA_GPU_OK;
}

int caps_sa_gpu_engine_set_stream(caps_sa_gpu_engine* engine, void* stream) {
  if (!engine) return bad_args("engine is NULL");
  engine->impl.stream = stream ? static_cast<cudaStream_t>(stream) : engine->impl.own_stream;
  return CAPS_SA_GPU_OK;
}

int caps_sa_gpu_engine_set_kernel_timing(caps_sa_gpu_engine* engine, int enabled) {
  if (!engine) return bad_args("engine is NULL");
  engine->impl.radix.timer.enabled = enabled != 0;
  engine->impl.msd_timers.set_enabled(enabled != 0);
  return CAPS_SA_GPU_OK;
}

int caps_sa_gpu_construct_u32(caps_sa_gpu_engine* engine, const char* text, uint64_t n, uint32_t* sa_out,
                              uint32_t* lcp_out, uint64_t /*subproblem_count*/, uint64_t max_context) {
  return construct_host<uint32_t>(engine, text, n, sa_out, lcp_out, max_context);
}
int caps_sa_gpu_construct_u64(caps_sa_gpu_engine* engine, const char* text, uint64_t n, uint64_t* sa_out,
                              uint64_t* lcp_out, uint64_t /*subproblem_count*/, uint64_t max_context) {
  return construct_host<uint64_t>(engine, text, n, sa_out, lcp_out, max_context);
}
int caps_sa_gpu_construct_device_u32(caps_sa_gpu_engine* engine, const void* d_text, uint64_t n, uint32_t* d_sa,
                                     uint32_t* d_lcp, void* stream) {
  return construct_device<uint32_t>(engine, d_text, n, d_sa, d_lcp, stream);
}
int caps_sa_gpu_construct_device_u64(caps_sa_gpu_engine* engine, const void* d_text, uint64_t n, uint64_t* d_sa,
                                     uint64_t* d_lcp, void* stream) {
  return construct_device<uint64_t>(engine, d_text, n, d_sa, d_lcp, stream);
}

int caps_sa_gpu_construct_multi_u32(const int* devices, int num_ranks, const char* text, uint64_t n, uint32_t* sa_out,
                                    uint32_t* lcp_out, uint64_t /*subproblem_count*/, uint64_t max_context,
                                    caps_sa_gpu_stats* stats_out) {
  return construct_multi<uint32_t>(devices, num_ranks, text, n, sa_out, lcp_out, max_context, stats_out);
}
int caps_sa_gpu_construct_multi_u64(const int* devices, int num_ranks, const char* text, uint64_t n, uint64_t* sa_out,
                                    uint64_t* lcp_out, uint64_t /*subproblem_count*/, uint64_t max_context,
                                    caps_sa_gpu_stats* stats_out) {
  return construct_multi<uint64_t>(devices, num_ranks, text, n, sa_out, lcp_out, max_context, stats_out);
}

int caps_sa_gpu_comm_unique_id(void* id_out) {
  if (!id_out) return bad_args("NULL argument");
  return guarded([&]() -> int {
    capsb::nccl_unique_id(id_out);
    return CAPS_SA_GPU_OK;
  });
}

int caps_sa_gpu_engine_comm_init(caps_sa_gpu_engine* engine, const void* id, int rank, int world) {
  if (!engine || !id) return bad_args("NULL argument");
  if (world < 1 || rank < 0 || rank >= world) return bad_args("bad rank / world size");
  return guarded([&]() -> int {
    Engine& eng = engine->impl;
    capsb::ArenaScope arena_scope(&eng.arena);
    eng.comm.reset();
    if (world == 1)
      eng.comm.reset(new capsb::SelfComm());
    else
      eng.comm.reset(new capsb::NcclComm(id, rank, world, eng.dev.device));
    return CAPS_SA_GPU_OK;
  });
}

int caps_sa_gpu_construct_sharded_device_u32(caps_sa_gpu_engine* engine, const void* d_text, uint64_t n, void* stream) {
  return construct_sharded_device<uint32_t>(engine, d_text, n, stream);
}
int caps_sa_gpu_construct_sharded_device_u64(caps_sa_gpu_engine* engine, const void* d_text, uint64_t n, void* stream) {
  return construct_sharded_device<uint64_t>(engine, d_text, n, stream);
}
int caps_sa_gpu_construct_sharded_u32(caps_sa_gpu_engine* engine, const char* text, uint64_t n, uint32_t* sa_out,
                                      uint32_t* lcp_out) {
  return construct_sharded_host<uint32_t>(engine, text, n, sa_out, lcp_out);
}
int caps_sa_gpu_construct_sharded_u64(caps_sa_gpu_engine* engine, const char* text, uint64_t n, uint64_t* sa_out,
                                      uint64_t* lcp_out) {
  return construct_sharded_host<uint64_t>(engine, text, n, sa_out, lcp_out);
}

int caps_sa_gpu_shard_copy(caps_sa_gpu_engine* engine, void* sa_dst, void* lcp_dst, int to_host) {
  if (!engine) return bad_args("engine is NULL");
  return guarded([&]() -> int {
    Engine& eng = engine->impl;
    capsb::ArenaScope arena_scope(&eng.arena);
    CAPSB_CUDA(cudaSetDevice(eng.dev.device));
    const cudaMemcpyKind kind = to_host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    const bool wide = eng.stats.idx_bytes == 8;
    const uint64_t count = wide ? eng.shard64.count : eng.shard32.count;
    const size_t bytes = count * (wide ? 8 : 4);
    const void* sa = wide ? static_cast<const void*>(eng.shard64.sa.get()) : eng.shard32.sa.get();
    const void* lcp = wide ? static_cast<const void*>(eng.shard64.lcp.get()) : eng.shard32.lcp.get();
    if (bytes && (!sa_dst || !lcp_dst)) capsb::fail("NULL destination");
    if (bytes) {
      CAPSB_CUDA(cudaMemcpyAsync(sa_dst, sa, bytes, kind, eng.stream));
      CAPSB_CUDA(cudaMemcpyAsync(lcp_dst, lcp, bytes, kind, eng.stream));
    }
    CAPSB_CUDA(cudaStreamSynchronize(eng.stream));
    return CAPS_SA_GPU_OK;
  });
}

int caps_sa_gpu_map_acgt(caps_sa_gpu_engine* engine, char* text, uint64_t n) {
  if (!engine) return bad_args("engine is NULL");
  if (n > 0 && !text) return bad_args("NULL buffer");
  return guarded([&]() -> int {
    Engine& eng = engine->impl;
    capsb::ArenaScope arena_scope(&eng.arena);
    CAPSB_CUDA(cudaSetDevice(eng.dev.device));
    if (n == 0) return CAPS_SA_GPU_OK;
    capsb::DevBuf<uint8_t> d(n, eng.stream);
    CAPSB_CUDA(cudaMemcpyAsync(d.get(), text, n, cudaMemcpyHostToDevice, eng.stream));
    capsb::map_acgt_device(eng, d.get(), n);
    CAPSB_CUDA(cudaMemcpyAsync(text, d.get(), n, cudaMemcpyDeviceToHost, eng.stream));
    CAPSB_CUDA(cudaStreamSynchronize(eng.stream));
    return CAPS_SA_GPU_OK;
  });
}

int caps_sa_gpu_set_cli_byte_mapping(int enabled) { return g_cli_byte_mapping.exchange(enabled ? 1 : 0); }

void* caps_sa_gpu_host_alloc(size_t bytes) {
  void* p = nullptr;
  const cudaError_t err = cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable);
  if (err != cudaSuccess) {
    cudaGetLastError();
    g_last_error = std::string("CUDA error: cudaHostAlloc failed: ") + cudaGetErrorString(err);
    return nullptr;
  }
  return p;
}
void caps_sa_gpu_host_free(void* ptr) {
  if (ptr) cudaFreeHost(ptr);
}

int caps_sa_gpu_stage_pack(caps_sa_gpu_engine* engine, const char* text, uint64_t n, uint64_t* words_out,
                           uint64_t* nwords_out, uint32_t* alphabet_size_out) {
  if (!engine || !text || !words_out || !nwords_out) return -bad_args("NULL argument");
  int bits = 0;
  const int rc = guarded([&]() -> int {
    Engine& eng = engine->impl;
    capsb::ArenaScope arena_scope(&eng.arena);
    CAPSB_CUDA(cudaSetDevice(eng.dev.device));
    capsb::DevBuf<uint8_t> d(n ? n : 1, eng.stream);
    CAPSB_CUDA(cudaMemcpyAsync(d.get(), text, n, cudaMemcpyHostToDevice, eng.stream));
    capsb::PackedTextBuf p = capsb::pack_text(eng, d.get(), n);
    CAPSB_CUDA(cudaMemcpyAsync(words_out, p.words.get(), p.nwords * sizeof(uint64_t), cudaMemcpyDeviceToHost,
                               eng.stream));
    CAPSB_CUDA(cudaStreamSynchronize(eng.stream));
    *nwords_out = p.nwords;
    if (alphabet_size_out) *alphabet_size_out = p.sigma;
    bits = 1 << p.log2_bits;
    return CAPS_SA_GPU_OK;
  });
  return rc == CAPS_SA_GPU_OK ? bits : -rc;
}

int caps_sa_gpu_stage_radix_sort_u64_u32(caps_sa_gpu_engine* engine, uint64_t* keys, uint32_t* vals, uint64_t n,
                                         unsigned begin_bit, unsigned end_bit) {
  if (!engine || (n && (!keys || !vals))) return bad_args("NULL argument");
  if (end_bit > 64 || begin_bit >= end_bit) return bad_args("bad bit range");
  return guarded([&]() -> int {
    Engine& eng = engine->impl;
    capsb::ArenaScope arena_scope(&eng.arena);
    CAPSB_CUDA(cudaSetDevice(eng.dev.device));
    if (n == 0) return CAPS_SA_GPU_OK;
    cudaStream_t st = eng.stream;
    capsb::DevBuf<uint64_t> ka(n, st), kb(n, st);
    capsb::DevBuf<uint32_t> va(n, st), vb(n, st);
    CAPSB_CUDA(cudaMemcpyAsync(ka.get(), keys, n * 8, cudaMemcpyHostToDevice, st));
    CAPSB_CUDA(cudaMemcpyAsync(va.get(), vals, n * 4, cudaMemcpyHostToDevice, st));
    const int where = capsb::radix_sort_pairs<uint64_t, uint32_t>(st, eng.radix, ka.get(), va.get(), kb.get(),
                                                                  vb.get(), n, begin_bit, end_bit);
    CAPSB_CUDA(cudaMemcpyAsync(keys, where ? kb.get() : ka.get(), n * 8, cudaMemcpyDeviceToHost, st));
    CAPSB_CUDA(cudaMemcpyAsync(vals, where ? vb.get() : va.get(), n * 4, cudaMemcpyDeviceToHost, st));
    CAPSB_CUDA(cudaStreamSynchronize(st));
    return CAPS_SA_GPU_OK;
  });
}

int caps_sa_gpu_stage_key_sort_u32(caps_sa_gpu_engine* engine, const char* text, uint64_t n, int use_lsd,
                                   uint64_t* keys_out, uint32_t* sa_out) {
  if (!engine || (n && (!text || !keys_out || !sa_out))) return -bad_args("NULL argument");
  if (n > 0xFFFFFFFFull) return -bad_args("n does not fit 32-bit indices");
  int key_bits = 0;
  const int rc = guarded([&]() -> int {
    Engine& eng = engine->impl;
    capsb::ArenaScope arena_scope(&eng.arena);
    CAPSB_CUDA(cudaSetDevice(eng.dev.device));
    if (n == 0) return CAPS_SA_GPU_OK;
    cudaStream_t st = eng.stream;
    capsb::DevBuf<uint8_t> d_text(n, st);
    capsb::DevBuf<uint64_t> d_keys(n, st);
    capsb::DevBuf<uint32_t> d_sa(n, st);
    CAPSB_CUDA(cudaMemcpyAsync(d_text.get(), text, n, cudaMemcpyHostToDevice, st));
    key_bits = capsb::stage_key_sort_u32(eng, d_text.get(), n, use_lsd != 0, d_keys.get(), d_sa.get());
    CAPSB_CUDA(cudaMemcpyAsync(keys_out, d_keys.get(), n * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    CAPSB_CUDA(cudaMemcpyAsync(sa_out, d_sa.get(), n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CAPSB_CUDA(cudaStreamSynchronize(st));
    return CAPS_SA_GPU_OK;
  });
  return rc == CAPS_SA_GPU_OK ? key_bits : -rc;
}

int caps_sa_gpu_stage_scan_u32(caps_sa_gpu_engine* engine, uint32_t* data, uint64_t n, int inclusive_max) {
  if (!engine || (n && !data)) return bad_args("NULL argument");
  return guarded([&]() -> int {
    Engine& eng = engine->impl;
    capsb::ArenaScope arena_scope(&eng.arena);
    CAPSB_CUDA(cudaSetDevice(eng.dev.device));
    if (n == 0) return CAPS_SA_GPU_OK;
    cudaStream_t st = eng.stream;
    capsb::DevBuf<uint32_t> in(n, st), out(n, st);
    CAPSB_CUDA(cudaMemcpyAsync(in.get(), data, n * 4, cudaMemcpyHostToDevice, st));
    capsb::stage_scan_u32(eng, in.get(), out.get(), n, inclusive_max != 0);
    CAPSB_CUDA(cudaMemcpyAsync(data, out.get(), n * 4, cudaMemcpyDeviceToHost, st));
    CAPSB_CUDA(cudaStreamSynchronize(st));
    return CAPS_SA_GPU_OK;
  });
}

}  // extern "C"
