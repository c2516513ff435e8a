// Suffix-array + LCP construction on one B200 (the hot path).
//
// Replaces the reference's construct() pipeline (src/Suffix_Array.cpp:466-494):
//   permute + sort_subarrays/merge_sort/merge (:112-184)  ->  one stable LSD radix sort of all
//       suffixes on their 64-bit packed-prefix key (radix_sort.cuh), keys read straight
//       from the packed text;
//   character-compare tie resolution inside merge (:69-80)  ->  prefix-doubling rank
//       refinement restricted to the suffixes that are still tied ("discarding");
//   LCP carried through merges (:61-68,:78)  ->  clz(key_a ^ key_b) for neighbours with
//       different keys; for tied neighbours the permuted-LCP recurrence
//       PLCP[i] = PLCP[i-1] - 1 on reducible positions and a direct packed-word
//       comparison on the irreducible ones (sum of irreducible LCPs <= 2 n log n).
// The output is the canonical SA/LCP under signed-char order, shorter suffix first —
// bit-identical to the reference at its default (unbounded) context.
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

  // ---- 2. key sort of all suffixes (leading key_bits of the packed prefix, stable LSD passes) ---
  const unsigned key_bits = choose_key_bits(n);
  eng.stats.key_bits = key_bits;
  DevBuf<uint64_t> key_buf(n, st);
  sort_suffix_slice<IdxT>(eng, pt, 0, n, key_bits, key_buf.get(), d_sa);
  const uint64_t* keys = key_buf.get();
  clock.mark();  // 2
  clock.mark();  // 3

  // ---- 3. ties: prefix-doubling refinement + pair chains; LCPs the keys decide ---------------
  TiedSet<IdxT> tied;
  {
    LocalRanks<IdxT> ranks(eng, n, pt, keys, key_mask_of(key_bits));
    refine_tied_groups<IdxT>(eng, ranks, pt, key_bits, keys, d_sa, d_lcp, n, 0, n, tied);
  }
  eng.sa_is_final(d_sa, 0, n, sizeof(IdxT));
  first_position_lcp<IdxT>(eng, keys, d_sa, d_lcp, n, n, log2_bits, false, 0, 0);
  eng.stats.tied_after_key_sort = tied.m;
  clock.mark();  // 4

  // ---- 4. LCP of the tied neighbours still open: permuted-LCP recurrence ---------------------
  if (tied.m > 0) {
    // the pairs (i = SA[k], j = SA[k-1]) keyed by i; j and k travel as the sort's value
    DevBuf<IdxT> pos_a, pair_j, pair_k;
    const uint64_t m = collect_deep_pairs<IdxT>(eng, tied, keys, d_sa, d_lcp, pos_a, pair_j, pair_k);
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
  }
  clock.mark();  // 5

  clock.mark();  // 6
  CAPSB_CUDA(cudaStreamSynchronize(st));
  eng.stats.ms_pack = clock.between(0, 1);
  eng.stats.ms_sort = clock.between(1, 2);
  eng.stats.ms_heads = clock.between(2, 3);
  eng.stats.ms_refine = clock.between(3, 4);
  eng.stats.ms_deep_lcp = clock.between(4, 5);
  eng.stats.ms_total = clock.between(0, 6);
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
