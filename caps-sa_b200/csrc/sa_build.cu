// Suffix-array + LCP construction on one B200 (the hot path).
//
// Replaces the reference's construct() pipeline (src/Suffix_Array.cpp:466-494):
//   permute + sort_subarrays/merge_sort/merge (:112-184)  ->  a key sort of all suffixes on the top
//       key_bits of their packed text: most significant digit first on packed 8-byte records
//       (msd_sort.cuh: two partition levels read straight from the packed text, then one CTA per
//       bucket in shared memory); stable LSD passes (radix_sort.cuh) for 64-bit indices;
//   character-compare tie resolution inside merge (:69-80)  ->  refinement restricted to the
//       suffixes that are still tied: pair chains, text rounds, then prefix doubling on ranks
//       (pipeline.cuh: refine_shallow / refine_deep / fix_group_edges);
//   LCP carried through merges (:61-68,:78)  ->  clz(key_a ^ key_b) for neighbours with
//       different keys, the common leading symbols of the comps for neighbours a text round
//       separates; for the rest the permuted-LCP recurrence PLCP[i] = PLCP[i-1] - 1 on
//       reducible positions and a direct packed-word comparison on the irreducible ones
//       (sum of irreducible LCPs <= 2 n log n).
// With result arrays in pinned host memory the construction runs a range of SA positions at a
// time (plan_ranges) and copies every finished range out while the next one is refined.
// The output is the canonical SA/LCP under signed-char order, shorter suffix first —
// bit-identical to the reference at its default (unbounded) context.
#include <cstdlib>
#include <vector>

#include "comm.cuh"
#include "pipeline.cuh"

namespace capsb {

std::atomic<uint64_t> g_kernel_launches{0};

Engine::Engine(int device) {
  dev.device = device;
  CAPSB_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  CAPSB_CUDA(cudaGetDeviceProperties(&prop, device));
  dev.sm_count = prop.multiProcessorCount;
  ArenaScope scope(&arena);
  CAPSB_CUDA(cudaStreamCreateWithFlags(&own_stream, cudaStreamNonBlocking));
  stream = own_stream;
  CAPSB_CUDA(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
  radix.init(dev, stream);
  scan32.init(dev, stream);
  scan64.init(dev, stream);
}

Engine::~Engine() {
  cudaSetDevice(dev.device);
  if (own_stream) cudaStreamSynchronize(own_stream);
  comm.reset();
  shard32 = ShardResult<uint32_t>();
  shard64 = ShardResult<uint64_t>();
  radix = RadixScratch();
  scan32 = ScanScratch<uint32_t>();
  scan64 = ScanScratch<uint64_t>();
  for (cudaEvent_t e : events) cudaEventDestroy(e);
  if (copy_stream) cudaStreamDestroy(copy_stream);
  if (own_stream) cudaStreamDestroy(own_stream);
}

void Engine::sa_is_final(const void* d_sa, uint64_t first, uint64_t count, size_t idx_bytes) {
  if (!sa_sink || count == 0) return;
  cudaEvent_t ready;
  if (!events.empty()) {
    ready = events.back();
    events.pop_back();
  } else {
    CAPSB_CUDA(cudaEventCreate(&ready));
  }
  CAPSB_CUDA(cudaEventRecord(ready, stream));
  CAPSB_CUDA(cudaStreamWaitEvent(copy_stream, ready, 0));
  events.push_back(ready);  // the wait has captured this recording
  if (sa_sink_started) CAPSB_CUDA(cudaEventRecord(sa_sink_started, copy_stream));
  CAPSB_CUDA(cudaMemcpyAsync(static_cast<char*>(sa_sink) + first * idx_bytes, d_sa, count * idx_bytes,
                             cudaMemcpyDeviceToHost, copy_stream));
  sa_sunk = true;
}

void Engine::range_is_final(const void* d_sa, const void* d_lcp, uint64_t first, uint64_t count, size_t idx_bytes) {
  if (!can_stream() || count == 0) return;
  cudaEvent_t ready;
  if (!events.empty()) {
    ready = events.back();
    events.pop_back();
  } else {
    CAPSB_CUDA(cudaEventCreate(&ready));
  }
  CAPSB_CUDA(cudaEventRecord(ready, stream));
  CAPSB_CUDA(cudaStreamWaitEvent(copy_stream, ready, 0));
  events.push_back(ready);  // the wait has captured this recording
  if (!sa_sunk && sa_sink_started) CAPSB_CUDA(cudaEventRecord(sa_sink_started, copy_stream));
  CAPSB_CUDA(cudaMemcpyAsync(static_cast<char*>(sa_sink) + first * idx_bytes, static_cast<const char*>(d_sa) + first * idx_bytes,
                             count * idx_bytes, cudaMemcpyDeviceToHost, copy_stream));
  CAPSB_CUDA(cudaMemcpyAsync(static_cast<char*>(lcp_sink) + first * idx_bytes, static_cast<const char*>(d_lcp) + first * idx_bytes,
                             count * idx_bytes, cudaMemcpyDeviceToHost, copy_stream));
  sa_sunk = true;
}

namespace {

// Ranges of SA positions that the single-GPU path finishes one after the other.  Boundaries are
// starts of level-A buckets of the key sort, so no key group (hence no tie group) spans one.
struct RangePlan {
  std::vector<uint64_t> bound;    // bound[r] .. bound[r + 1]: SA positions of range r
  std::vector<uint32_t> a_index;  // a_index[r] .. a_index[r + 1]: its level-A buckets
};

RangePlan plan_ranges(uint64_t n, const std::vector<uint32_t>& a_starts, uint64_t want) {
  RangePlan plan;
  const uint32_t buckets = static_cast<uint32_t>(a_starts.size()) - 1;
  plan.bound.push_back(0);
  plan.a_index.push_back(0);
  if (want < 1) want = 1;
  const uint64_t target = (n + want - 1) / want;
  for (uint32_t x = 1; x < buckets; ++x) {
    const uint64_t here = a_starts[x];
    if (here - plan.bound.back() >= target && n - here > 0 && here > plan.bound.back()) {
      plan.bound.push_back(here);
      plan.a_index.push_back(x);
    }
  }
  plan.bound.push_back(n);
  plan.a_index.push_back(buckets);
  return plan;
}

uint64_t ranges_wanted(const Engine& eng, uint64_t n) {
  if (const char* env = std::getenv("CAPSB_STREAM_RANGES")) return static_cast<uint64_t>(std::atoll(env));
  if (!eng.can_stream()) return 1;
  // a range's copy (8 bytes per position over PCIe) should dwarf the fixed cost of refining it
  const uint64_t want = n / (96ull << 20);
  return want < 1 ? 1 : (want > 32 ? 32 : want);
}

}  // namespace

template <class IdxT>
void build_sa_lcp(Engine& eng, const uint8_t* d_text, uint64_t n, IdxT* d_sa, IdxT* d_lcp) {
  CAPSB_CUDA(cudaSetDevice(eng.dev.device));
  cudaStream_t st = eng.stream;
  const uint64_t launches_before = g_kernel_launches.load();
  eng.stats = Stats();
  eng.radix.timer.reset();
  eng.msd_timers.reset();
  eng.stats.n = n;
  eng.stats.idx_bytes = sizeof(IdxT);
  eng.results_streamed = false;
  if (n == 0) return;
  if (sizeof(IdxT) == 4 && n > 0xFFFFFFFFull) fail("text too long for 32-bit indices");

  StageClock clock(eng);
  clock.mark();  // 0

  // ---- 1. text staging + key packing -----------------------------------------------------
  PackedTextBuf packed = pack_text(eng, d_text, n);
  const PackedText pt = packed.view(n);
  const unsigned log2_bits = pt.log2_bits;
  eng.stats.bits_per_symbol = pt.bits();
  eng.stats.alphabet_size = packed.sigma;
  clock.mark();  // 1

  // ---- 2. key sort (leading key_bits of the packed prefix) ---------------------------------
  // Packed-record MSD sort: the two partition levels run over everything, the buckets are then
  // finished range by range, each range followed at once by its share of step 3 and by its copy to
  // the host.  LSD passes (64-bit indices): one range.
  const unsigned key_bits = choose_key_bits(n);
  eng.stats.key_bits = key_bits;
  DevBuf<uint64_t> key_buf(n, st);
  uint64_t* keys = key_buf.get();
  const TextSource<IdxT> first{pt, key_mask_of(key_bits), 0};
  bool msd = false;
  MsdSorted ms;
  RangePlan ranges;
  ranges.bound = {0, n};
  ranges.a_index = {0, 0};
  if constexpr (sizeof(IdxT) == 4) {
    msd = msd_sort_enabled() && msd_applicable(key_bits);
    if (msd) {
      const uint64_t want = ranges_wanted(eng, n);
      msd_partition(eng, first, n, key_bits, keys, ms, true);
      ranges = plan_ranges(n, ms.a_starts, want);
    }
  }
  if (!msd) lsd_sort_suffixes<IdxT>(eng, first, n, key_bits, keys, d_sa);
  const size_t range_count = ranges.bound.size() - 1;
  const bool streaming = range_count > 1 && eng.can_stream();
  clock.mark();  // 2
  float ms_local = 0, ms_shallow = 0;

  // ---- 3. ties: per range, the shallow part of the refinement (pair chains + text rounds) ------
  LocalRanks<IdxT> ranks(eng, n, pt, keys, key_mask_of(key_bits), d_sa, key_bits >> log2_bits);
  std::vector<TiedSet<IdxT>> tied_of(range_count);
  std::vector<RefineState<IdxT>> state_of(range_count);
  for (size_t r = 0; r < range_count; ++r) {
    const uint64_t lo = ranges.bound[r], count = ranges.bound[r + 1] - lo;
    clock.mark("range: begin");
    if constexpr (sizeof(IdxT) == 4) {
      if (msd)
        msd_local_sort(eng, ms, keys, d_sa, ranges.a_index[r] << ms.b, ranges.a_index[r + 1] << ms.b, count);
    }
    clock.mark("range: buckets sorted");
    refine_shallow<IdxT>(eng, ranks, pt, key_bits, keys + lo, d_sa + lo, d_lcp + lo, count, lo, n, tied_of[r],
                         state_of[r]);
    fix_group_edges<IdxT>(eng, tied_of[r].pos.get(), tied_of[r].m, keys + lo, d_sa + lo, d_lcp + lo, count, n, log2_bits);
    {  // the range's first position against the last suffix of the range before it (a different key)
      IdxT* lcp_here = d_lcp + lo;
      const IdxT* sa_here = d_sa + lo;
      const uint64_t* keys_here = keys + lo;
      const bool has_prev = lo > 0;
      launch_map(eng.dev, st, 1, [=] __device__(uint64_t) {
        lcp_here[0] = has_prev ? key_lcp_value<IdxT>(keys_here[-1], keys_here[0], sa_here[-1], sa_here[0], n, log2_bits)
                               : IdxT(0);
      });
    }
    if (streaming) eng.range_is_final(d_sa, d_lcp, lo, count, sizeof(IdxT));
    clock.mark("range: shallow ties resolved");
    eng.stats.tied_after_key_sort += tied_of[r].m;
  }

  // ---- 3b. what the text rounds could not separate: prefix doubling over all ranges at once ------
  // (the lists are concatenated in range order = SA order; positions become positions of d_sa)
  TiedSet<IdxT> tied;
  RefineState<IdxT> deep;
  deep.h = ~0ull;
  for (size_t r = 0; r < range_count; ++r) {
    tied.m += tied_of[r].m;
    deep.act.m += state_of[r].act.m;
    if (state_of[r].act.m && state_of[r].h < deep.h) deep.h = state_of[r].h;
    if (state_of[r].h_resolved > deep.h_resolved) deep.h_resolved = state_of[r].h_resolved;
  }
  deep.total_active = deep.act.m;
  DevBuf<IdxT> deep_pos(deep.act.m, st);  // the positions the deep stages may still change
  if (range_count == 1) {
    tied = std::move(tied_of[0]);
    deep.act = std::move(state_of[0].act);
  } else {
    tied.pos.alloc(tied.m, st);
    deep.act.pos.alloc(deep.act.m, st);
    deep.act.idx.alloc(deep.act.m, st);
    deep.act.group.alloc(deep.act.m, st);
    uint64_t tied_at = 0, act_at = 0;
    for (size_t r = 0; r < range_count; ++r) {
      const IdxT lo = static_cast<IdxT>(ranges.bound[r]);
      {
        const IdxT* src = tied_of[r].pos.get();
        IdxT* dst = tied.pos.get() + tied_at;
        launch_map(eng.dev, st, tied_of[r].m, [=] __device__(uint64_t t) { dst[t] = src[t] + lo; });
        tied_at += tied_of[r].m;
      }
      {
        const ActiveList<IdxT>& a = state_of[r].act;
        const IdxT* sp = a.pos.get();
        const IdxT* si = a.idx.get();
        const IdxT* sg = a.group.get();
        IdxT* dp = deep.act.pos.get() + act_at;
        IdxT* di = deep.act.idx.get() + act_at;
        IdxT* dg = deep.act.group.get() + act_at;
        launch_map(eng.dev, st, a.m, [=] __device__(uint64_t t) {
          dp[t] = sp[t] + lo;
          di[t] = si[t];
          dg[t] = sg[t];
        });
        act_at += a.m;
      }
      tied_of[r] = TiedSet<IdxT>();
      state_of[r] = RefineState<IdxT>();
    }
  }
  const uint64_t deep_count = deep.act.m;
  if (deep_count) CAPSB_CUDA(cudaMemcpyAsync(deep_pos.get(), deep.act.pos.get(), deep_count * sizeof(IdxT), cudaMemcpyDeviceToDevice, st));
  refine_deep<IdxT>(eng, ranks, d_sa, d_lcp, 0, n, tied, deep);
  if (deep_count)
    fix_group_edges<IdxT>(eng, deep_pos.get(), deep_count, keys, d_sa, d_lcp, n, n, log2_bits);
  if (!streaming) eng.sa_is_final(d_sa, 0, n, sizeof(IdxT));
  clock.mark();  // after the deep ties

  // ---- 4. LCP of the tied neighbours still open: permuted-LCP recurrence ---------------------
  // (only neighbours that stayed tied into the rank rounds: positions listed in deep_pos)
  if (deep_count > 0) {
    TiedSet<IdxT> open;
    open.m = deep_count;
    open.pos = std::move(deep_pos);
    // the pairs (i = SA[k], j = SA[k-1]) keyed by i; j and k travel as the sort's value
    DevBuf<IdxT> pos_a, pair_j, pair_k;
    const uint64_t m = collect_deep_pairs<IdxT>(eng, open, keys, d_sa, d_lcp, pos_a, pair_j, pair_k);
    DevBuf<IdxT> pos_b(m, st);
    DevBuf<IdxPair<IdxT>> tag_a(m, st), tag_b(m, st);
    {
      const IdxT* pj = pair_j.get();
      const IdxT* pk = pair_k.get();
      IdxPair<IdxT>* ta = tag_a.get();
      launch_map(eng.dev, st, m, [=] __device__(uint64_t t) { ta[t] = IdxPair<IdxT>{pj[t], pk[t]}; });
    }
    pair_j.release();
    pair_k.release();
    const unsigned pos_bits = round_up8(bit_length(n - 1));
    const int where = radix_sort_pairs<IdxT, IdxPair<IdxT>>(st, eng.radix, pos_a.get(), tag_a.get(), pos_b.get(),
                                                           tag_b.get(), m, 0, pos_bits);
    const IdxT* pos_i = where ? pos_b.get() : pos_a.get();
    const IdxPair<IdxT>* tag = where ? tag_b.get() : tag_a.get();
    plcp_for_pairs<IdxT>(
        eng, pt, pos_i, [=] __device__(uint64_t t) -> uint64_t { return tag[t].a; }, m, 0,
        [=] __device__(uint64_t t, IdxT lcp, uint64_t, IdxT) { d_lcp[tag[t].b] = lcp; });
    deep_pos = std::move(open.pos);
  }
  // ---- 5. streamed results: the entries the deep stages settled go to the host arrays directly ----
  if (streaming) {
    if (deep_count * 8 > n) {
      // most of the text is deep ties (periodic texts): nothing was gained by the early copies,
      // the caller copies both arrays again
      eng.results_streamed = false;
      eng.sa_sunk = false;
    } else {
      if (deep_count > 0) {
        cudaEvent_t done;
        if (!eng.events.empty()) {
          done = eng.events.back();
          eng.events.pop_back();
        } else {
          CAPSB_CUDA(cudaEventCreate(&done));
        }
        CAPSB_CUDA(cudaEventRecord(done, st));
        CAPSB_CUDA(cudaStreamWaitEvent(eng.copy_stream, done, 0));
        eng.events.push_back(done);
        // on the copy stream, behind the ranges' copies: the later write wins
        const IdxT* pos = deep_pos.get();
        const IdxT* sa_dev = d_sa;
        const IdxT* lcp_dev = d_lcp;
        IdxT* sa_host = static_cast<IdxT*>(eng.sa_sink_dev);
        IdxT* lcp_host = static_cast<IdxT*>(eng.lcp_sink_dev);
        launch_map(eng.dev, eng.copy_stream, deep_count, [=] __device__(uint64_t t) {
          const uint64_t k = pos[t];
          sa_host[k] = sa_dev[k];
          lcp_host[k] = lcp_dev[k];
          // the entry after a group's last member depends on that member as well
          if (k + 1 < n && (t + 1 == deep_count || static_cast<uint64_t>(pos[t + 1]) != k + 1)) lcp_host[k + 1] = lcp_dev[k + 1];
        });
        CAPSB_CUDA(cudaStreamSynchronize(eng.copy_stream));  // deep_pos is released on return
      }
      eng.results_streamed = true;
    }
  }
  clock.mark();  // end

  CAPSB_CUDA(cudaStreamSynchronize(st));
  // marks: 0 begin, 1 packed, 2 partitioned, then per range (begin, sorted, shallow done), then deep done, end
  const size_t base = 3;
  for (size_t r = 0; r < range_count; ++r) {
    ms_local += clock.between(base + 3 * r, base + 3 * r + 1);
    ms_shallow += clock.between(base + 3 * r + 1, base + 3 * r + 2);
  }
  const size_t after = base + 3 * range_count;
  eng.stats.ms_pack = clock.between(0, 1);
  eng.stats.ms_sort = clock.between(1, 2) + ms_local;
  eng.stats.ms_heads = 0;
  eng.stats.ms_refine = ms_shallow + clock.between(after - 1, after);
  eng.stats.ms_deep_lcp = clock.between(after, after + 1);
  eng.stats.ms_total = clock.between(0, after + 1);
  eng.stats.kernel_launches = g_kernel_launches.load() - launches_before;
  if (eng.radix.timer.enabled) {
    eng.stats.scatter_bytes = eng.radix.timer.bytes;
    eng.stats.ms_scatter = eng.radix.timer.drain(&eng.stats.scatter_launches);
  }
  collect_msd_timings(eng);
}

int stage_key_sort_u32(Engine& eng, const uint8_t* d_text, uint64_t n, bool use_lsd, uint64_t* d_keys, uint32_t* d_sa) {
  eng.stats = Stats();
  eng.radix.timer.reset();
  eng.msd_timers.reset();
  PackedTextBuf packed = pack_text(eng, d_text, n);
  const PackedText pt = packed.view(n);
  const unsigned key_bits = choose_key_bits(n);
  const TextSource<uint32_t> first{pt, key_mask_of(key_bits), 0};
  StageClock clock(eng);
  clock.mark();
  if (use_lsd || !msd_applicable(key_bits))
    lsd_sort_suffixes<uint32_t>(eng, first, n, key_bits, d_keys, d_sa);
  else
    msd_sort_suffixes(eng, first, n, key_bits, d_keys, d_sa);
  clock.mark();
  CAPSB_CUDA(cudaStreamSynchronize(eng.stream));
  eng.stats.n = n;
  eng.stats.key_bits = key_bits;
  eng.stats.ms_sort = clock.between(0, 1);
  if (eng.radix.timer.enabled) {
    eng.stats.scatter_bytes = eng.radix.timer.bytes;
    eng.stats.ms_scatter = eng.radix.timer.drain(&eng.stats.scatter_launches);
  }
  collect_msd_timings(eng);
  return static_cast<int>(key_bits);
}

void stage_scan_u32(Engine& eng, const uint32_t* d_in, uint32_t* d_out, uint64_t n, bool inclusive_max) {
  auto in = [=] __device__(uint64_t i) -> uint32_t { return d_in[i]; };
  auto out = [=] __device__(uint64_t i, uint32_t v) { d_out[i] = v; };
  if (inclusive_max)
    scan_full<uint32_t, OpMax, true>(eng, n, in, out);
  else
    scan_full<uint32_t, OpSum, false>(eng, n, in, out);
}

template void build_sa_lcp<uint32_t>(Engine&, const uint8_t*, uint64_t, uint32_t*, uint32_t*);
template void build_sa_lcp<uint64_t>(Engine&, const uint8_t*, uint64_t, uint64_t*, uint64_t*);

}  // namespace capsb
