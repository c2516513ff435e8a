// Sharded (multi-GPU) suffix-array + LCP construction: the samplesort shape of the reference
// (src/Suffix_Array.cpp:466-494) with one rank per GPU.  Default order of stages (partition first):
//
//   text              replicated: every rank stages and packs the whole text in its HBM
//   select_pivots     (:197-222)  1024 sampled keys of the rank's slice of the text (last slice takes the
//                                 remainder, as :172), all-gathered; every rank sorts the same sample set
//                                 and picks the same G-1 key pivots
//   locate_pivots + partition (:225-368)  one counting pass over the slice with the bucket (binary search
//                                 among the pivots) as digit -> the slice's suffixes grouped by bucket, in
//                                 text order; send counts all-gathered -> the G x G matrix P
//   collate           (:335-364)  all-to-all of suffix INDICES over NVLink (keys are re-derived from the
//                                 replicated text); CAPSB_SHARD_P2P=1: the partition kernel stores straight
//                                 into the owners' buckets instead (partition.cuh; measured slower at two GPUs)
//   sort_subarrays / merge_sub_subarrays (:161-184, :371-428)  the rank key-sorts its bucket (packed-record
//                                 MSD sort, msd_sort.cuh)
//   ties                          pair chains and text rounds are local to the rank; prefix doubling on ranks
//                                 sharded by text position: per round one request/response exchange and one
//                                 update exchange; ranks of final suffixes are found by the owner of their key
//   LCP               (:61-68,:78,:431-447) neighbours with different keys: from the keys (the
//                                 predecessor of a bucket's first suffix comes from the previous rank = the
//                                 reference's boundary patch); tied neighbours travel to the rank that owns
//                                 their text position, where the permuted-LCP chains are contiguous, and back.
// CAPSB_SHARD_MODE=merge keeps the reference's own order (sort the slices, locate the pivots in the sorted
// slices, exchange sorted (key, suffix) runs, merge-path tree over the received runs).
// Pivots are keys, so suffixes with equal keys land on one rank and tie groups never span
// ranks; rank r ends up owning the contiguous range [offset, offset + count) of SA and LCP.
#include "comm.cuh"
#include "merge_path.cuh"
#include "partition.cuh"
#include "pipeline.cuh"

namespace capsb {

namespace {

// Slices of text positions: rank r owns [r * slice, (r + 1) * slice), the last rank to n.
struct SliceMap {
  uint64_t n;
  uint64_t slice;
  unsigned world;
  __host__ __device__ unsigned owner(uint64_t pos) const {
    if (slice == 0) return world - 1;
    const uint64_t o = pos / slice;
    return o < world - 1 ? static_cast<unsigned>(o) : world - 1;
  }
  __host__ __device__ uint64_t begin(unsigned r) const { return static_cast<uint64_t>(r) * slice; }
  __host__ __device__ uint64_t end(unsigned r) const {
    return r + 1 == world ? n : static_cast<uint64_t>(r + 1) * slice;
  }
};

// ---- routing items to owner ranks ------------------------------------------------------
// A route is a stable partition of m local items by destination rank (one counting pass of
// the radix machinery on the destination as an 8-bit digit) plus the exchanged counts.
template <class IdxT>
struct Route {
  uint64_t m = 0;
  uint64_t recv_total = 0;
  DevBuf<IdxT> perm;  // perm[j] = local item that travels in position j
  std::vector<uint64_t> send_counts, recv_counts;
};

template <class IdxT, class DestFn>
struct DestSource {
  DestFn dest;
  __device__ __forceinline__ uint8_t key(uint64_t t) const { return static_cast<uint8_t>(dest(t)); }
  __device__ __forceinline__ IdxT val(uint64_t t) const { return static_cast<IdxT>(t); }
  static constexpr uint64_t bytes_read_per_item() { return sizeof(IdxT); }
};

template <class IdxT, class DestFn>
Route<IdxT> plan_route(Engine& eng, Comm& comm, uint64_t m, DestFn dest) {
  cudaStream_t st = eng.stream;
  const unsigned world = static_cast<unsigned>(comm.world);
  if (world > static_cast<unsigned>(kRadixSize)) fail("more ranks than the routing digit can hold");
  Route<IdxT> r;
  r.m = m;
  r.send_counts.assign(world, 0);
  r.recv_counts.assign(world, 0);
  if (m) {
    r.perm.alloc(m, st);
    DevBuf<uint8_t> sorted_dest(m, st);
    radix_pass<uint8_t, IdxT>(st, eng.radix, DestSource<IdxT, DestFn>{dest}, m, 0, sorted_dest.get(), r.perm.get());
    read_back(st, r.send_counts.data(), eng.radix.digit_total.get(), world * sizeof(uint64_t));
  }
  std::vector<uint64_t> matrix(static_cast<size_t>(world) * world);
  comm.all_gather_host(r.send_counts.data(), world * sizeof(uint64_t), matrix.data(), st);
  for (unsigned s = 0; s < world; ++s) {
    r.recv_counts[s] = matrix[static_cast<size_t>(s) * world + comm.rank];
    r.recv_total += r.recv_counts[s];
  }
  return r;
}

// recv[j] (j over the items this rank receives, grouped by source rank) = in(item)
template <class T, class IdxT, class In>
void route_forward(Engine& eng, Comm& comm, const Route<IdxT>& r, In in, T* recv) {
  cudaStream_t st = eng.stream;
  DevBuf<T> send(r.m, st);
  {
    T* s = send.get();
    const IdxT* perm = r.perm.get();
    launch_map(eng.dev, st, r.m, [=] __device__(uint64_t j) { s[j] = in(static_cast<uint64_t>(perm[j])); });
  }
  comm.all_to_all_v(send.get(), r.send_counts.data(), recv, r.recv_counts.data(), sizeof(T), st);
}

// The reverse trip: answers[j] for every received item go back to the sender, which gets
// out(item, answer).
template <class T, class IdxT, class Out>
void route_backward(Engine& eng, Comm& comm, const Route<IdxT>& r, const T* answers, Out out) {
  cudaStream_t st = eng.stream;
  DevBuf<T> back(r.m, st);
  comm.all_to_all_v(answers, r.recv_counts.data(), back.get(), r.send_counts.data(), sizeof(T), st);
  const T* b = back.get();
  const IdxT* perm = r.perm.get();
  launch_map(eng.dev, st, r.m, [=] __device__(uint64_t j) { out(static_cast<uint64_t>(perm[j]), b[j]); });
}

// A question to the rank that owns a key: the rank (SA position) of `suffix`, whose key it is.
struct KeyAsk {
  uint64_t key, suffix;
};

// ---- ranks sharded by text position ----------------------------------------------------
// isa_local[pos - lo] = first SA position of suffix pos's group, for suffixes that have been
// tied; everything else keeps the sentinel.  A lookup that hits the sentinel asks a second
// owner: the rank whose bucket holds the suffix's (unique) key, which answers with the key's
// global position (bucket offset + binary search in its sorted bucket).
template <class IdxT>
struct ShardedRanks {
  using Comp = typename IdxTraits<IdxT>::Comp;
  static constexpr IdxT kUnset = ~IdxT(0);
  static constexpr bool kPublishesAllTied = false;  // ranks of final suffixes are found by the bucket's owner (rank_of_unpublished)
  uint64_t resolved_depth = 0;  // set by refine_deep
  Engine& eng;
  Comm& comm;
  SliceMap map;
  PackedText pt;
  uint64_t key_mask;
  const uint64_t* sorted_samples;  // pivot j = sorted_samples[(j + 1) * sample_stride - 1]
  uint64_t sample_stride;
  const uint64_t* bucket_keys;     // this rank's sorted bucket
  const IdxT* bucket_sa;           // ... and its suffixes (the rank's range of the suffix array, under construction)
  uint64_t key_symbols;            // symbols the key covers
  uint64_t bucket_count, bucket_offset;
  uint64_t lo;              // first text position of this rank's slice
  uint64_t slice_len = 0;
  DevBuf<IdxT> isa_local;

  ShardedRanks(Engine& e, Comm& c, SliceMap m, const PackedText& text, uint64_t mask, const uint64_t* samples,
               uint64_t stride, const uint64_t* keys, const IdxT* sa, uint64_t key_syms, uint64_t count, uint64_t offset)
      : eng(e), comm(c), map(m), pt(text), key_mask(mask), sorted_samples(samples), sample_stride(stride),
        bucket_keys(keys), bucket_sa(sa), key_symbols(key_syms), bucket_count(count), bucket_offset(offset),
        lo(m.begin(static_cast<unsigned>(c.rank))) {
    slice_len = m.end(static_cast<unsigned>(c.rank)) - lo;  // isa_local is allocated by reset()
  }

  void reset() {
    if (!isa_local) isa_local.alloc(slice_len ? slice_len : 1, eng.stream);
    CAPSB_CUDA(cudaMemsetAsync(isa_local.get(), 0xFF, isa_local.size() * sizeof(IdxT), eng.stream));
  }

  uint64_t global_sum(uint64_t v) {
    std::vector<uint64_t> all(static_cast<size_t>(comm.world));
    comm.all_gather_host(&v, sizeof(v), all.data(), eng.stream);
    uint64_t sum = 0;
    for (uint64_t x : all) sum += x;
    return sum;
  }

  // isa[idx[t]] = head[t] on the rank that owns text position idx[t], for t in [0, m)
  void publish(const IdxT* idx, const IdxT* head, uint64_t m) {
    const SliceMap mp = map;
    Route<IdxT> rt = plan_route<IdxT>(eng, comm, m, [=] __device__(uint64_t t) -> unsigned { return mp.owner(idx[t]); });
    DevBuf<IdxPair<IdxT>> recv(rt.recv_total, eng.stream);
    route_forward<IdxPair<IdxT>>(
        eng, comm, rt, [=] __device__(uint64_t t) -> IdxPair<IdxT> { return IdxPair<IdxT>{idx[t], head[t]}; }, recv.get());
    IdxT* isa = isa_local.get();
    const IdxPair<IdxT>* rv = recv.get();
    const uint64_t lo_ = lo;
    launch_map(eng.dev, eng.stream, rt.recv_total, [=] __device__(uint64_t j) { isa[rv[j].a - lo_] = rv[j].b; });
  }

  // The pairs travel to the rank that owns text position hi: consecutive text positions land in
  // unrelated buckets, so only there are the chains of chain_pairs_local contiguous.
  void chain_pairs(IdxT* hi, const IdxT* lo, uint64_t m, uint64_t known, IdxPair<IdxT>* answer) {
    cudaStream_t st = eng.stream;
    const SliceMap mp = map;
    Route<IdxT> rt = plan_route<IdxT>(eng, comm, m, [=] __device__(uint64_t t) -> unsigned { return mp.owner(hi[t]); });
    const uint64_t got = rt.recv_total;
    DevBuf<IdxPair<IdxT>> recv(got, st), ans(got, st);
    route_forward<IdxPair<IdxT>>(
        eng, comm, rt, [=] __device__(uint64_t t) -> IdxPair<IdxT> { return IdxPair<IdxT>{hi[t], lo[t]}; }, recv.get());
    DevBuf<IdxT> r_hi(got, st), r_lo(got, st);
    {
      const IdxPair<IdxT>* rv = recv.get();
      IdxT* h = r_hi.get();
      IdxT* l = r_lo.get();
      launch_map(eng.dev, st, got, [=] __device__(uint64_t j) { h[j] = rv[j].a, l[j] = rv[j].b; });
    }
    chain_pairs_local<IdxT>(eng, pt, r_hi.get(), r_lo.get(), got, known, ans.get());
    route_backward<IdxPair<IdxT>>(eng, comm, rt, ans.get(),
                                  [=] __device__(uint64_t t, IdxPair<IdxT> a) { answer[t] = a; });
  }

  // second_out[t] = rank of suffix idx[t] + h, or n - 1 - idx[t] beyond the end of the text
  void second_ranks(const IdxT* idx, uint64_t m, uint64_t h, IdxT* second_out) {
    cudaStream_t st = eng.stream;
    const SliceMap mp = map;
    const uint64_t n = map.n, lo_ = lo;
    const unsigned self = static_cast<unsigned>(comm.rank);
    IdxT* sec = second_out;

    // 1. ask the owner of text position i + h for its rank.  Suffixes that run past the end
    //    need none; they ride along as a request to this rank.
    {
      Route<IdxT> rt = plan_route<IdxT>(eng, comm, m, [=] __device__(uint64_t t) -> unsigned {
        const uint64_t ih = static_cast<uint64_t>(idx[t]) + h;
        return ih < n ? mp.owner(ih) : self;
      });
      DevBuf<IdxT> asked(rt.recv_total, st), answers(rt.recv_total, st);
      route_forward<IdxT>(
          eng, comm, rt,
          [=] __device__(uint64_t t) -> IdxT {
            const uint64_t ih = static_cast<uint64_t>(idx[t]) + h;
            return static_cast<IdxT>(ih < n ? ih : lo_);
          },
          asked.get());
      const IdxT* isa = isa_local.get();
      const IdxT* q = asked.get();
      IdxT* a = answers.get();
      launch_map(eng.dev, st, rt.recv_total, [=] __device__(uint64_t j) { a[j] = isa[q[j] - lo_]; });
      route_backward<IdxT>(eng, comm, rt, answers.get(), [=] __device__(uint64_t t, IdxT r) { sec[t] = r; });
    }

    // 2. never-tied suffixes: their rank is the global position of their key — ask the rank
    //    whose bucket holds that key
    {
      auto unknown = [=] __device__(uint64_t t) -> IdxT {
        return (static_cast<uint64_t>(idx[t]) + h < n && sec[t] == kUnset) ? IdxT(1) : IdxT(0);
      };
      const uint64_t pending = scan_total<IdxT, OpSum>(eng, m, unknown);
      DevBuf<IdxT> slot_of(pending, st);
      {
        IdxT* so = slot_of.get();
        scan_finish<IdxT, OpSum, false>(eng, m, unknown, [=] __device__(uint64_t t, IdxT j) {
          if (static_cast<uint64_t>(idx[t]) + h < n && sec[t] == kUnset) so[j] = static_cast<IdxT>(t);
        });
      }
      const IdxT* so = slot_of.get();
      const PackedText text = pt;
      const uint64_t mask = key_mask;
      const uint64_t* samples = sorted_samples;
      const uint64_t stride = sample_stride;
      const unsigned pivots = static_cast<unsigned>(comm.world) - 1;
      Route<IdxT> rt = plan_route<IdxT>(eng, comm, pending, [=] __device__(uint64_t j) -> unsigned {
        const uint64_t key = text.window(static_cast<uint64_t>(idx[so[j]]) + h) & mask;
        unsigned bucket = 0;  // number of pivots below the key (keys <= pivot j went to ranks <= j)
        while (bucket < pivots && samples[static_cast<uint64_t>(bucket + 1) * stride - 1] < key) ++bucket;
        return bucket;
      });
      // the question names the suffix as well as its key: a key shared by several (final) suffixes
      // is resolved inside the owner's bucket by comparing text (rank_of_unpublished)
      DevBuf<KeyAsk> asked(rt.recv_total, st);
      DevBuf<IdxT> answers(rt.recv_total, st);
      route_forward<KeyAsk>(
          eng, comm, rt,
          [=] __device__(uint64_t j) -> KeyAsk {
            const uint64_t ih = static_cast<uint64_t>(idx[so[j]]) + h;
            return KeyAsk{text.window(ih) & mask, ih};
          },
          asked.get());
      const uint64_t* keys = bucket_keys;
      const IdxT* sa = bucket_sa;
      const uint64_t count = bucket_count, offset = bucket_offset;
      const uint64_t ksym = key_symbols, depth = resolved_depth;
      const KeyAsk* q = asked.get();
      IdxT* a = answers.get();
      launch_map(eng.dev, st, rt.recv_total, [=] __device__(uint64_t j) {
        a[j] = static_cast<IdxT>(offset + rank_of_unpublished<IdxT>(text, keys, sa, count, q[j].key, q[j].suffix, ksym, depth));
      });
      route_backward<IdxT>(eng, comm, rt, answers.get(), [=] __device__(uint64_t j, IdxT r) { sec[so[j]] = r; });
    }

    launch_map(eng.dev, st, m, [=] __device__(uint64_t t) {
      const uint64_t i = idx[t];
      if (i + h >= n) sec[t] = static_cast<IdxT>(n - 1 - i);
    });
  }
};

// ---- pivot location --------------------------------------------------------------------
// One warp per pivot: 32-ary upper-bound search in the sorted keys (first position whose key
// is greater than the pivot).  bounds[0] = 0, bounds[j + 1] = upper_bound(pivot j),
// bounds[pivots + 1] = count.  Pivot j is sorted_samples[(j + 1) * stride - 1].
__global__ void __launch_bounds__(256) locate_pivots_kernel(const uint64_t* __restrict__ keys, uint64_t count,
                                                            const uint64_t* __restrict__ sorted_samples,
                                                            uint64_t stride, unsigned pivots,
                                                            uint64_t* __restrict__ bounds) {
  const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned lane = threadIdx.x & 31u;
  if (warp == 0 && lane == 0) {
    bounds[0] = 0;
    bounds[pivots + 1] = count;
  }
  if (warp >= pivots) return;
  const uint64_t pivot = sorted_samples[static_cast<uint64_t>(warp + 1) * stride - 1];
  uint64_t lo = 0, hi = count;  // answer in [lo, hi]
  while (hi - lo > 32) {
    const uint64_t width = hi - lo;
    const uint64_t probe = lo + (static_cast<uint64_t>(lane + 1) * width) / 33;  // strictly inside [lo, hi)
    const bool le = keys[probe] <= pivot;
    const unsigned vote = __ballot_sync(0xffffffffu, le);
    const int c = __popc(vote);  // probes are increasing, so `le` is a prefix of the lanes
    const uint64_t below = __shfl_sync(0xffffffffu, probe, c > 0 ? c - 1 : 0);
    const uint64_t above = __shfl_sync(0xffffffffu, probe, c < 32 ? c : 31);
    if (c > 0) lo = below + 1;
    if (c < 32) hi = above;
  }
  const bool le = lo + lane < hi && keys[lo + lane] <= pivot;
  const unsigned vote = __ballot_sync(0xffffffffu, le);
  if (lane == 0) bounds[warp + 1] = lo + static_cast<uint64_t>(__popc(vote));
}

constexpr unsigned kSamplesPerRank = 1024;

// Partition pass of the partition-first mode: "digit" = the bucket the suffix belongs to
// (number of pivots below its key: keys <= pivot j go to ranks <= j), value = the suffix.
template <class IdxT>
struct BucketSource {
  PackedText pt;
  uint64_t mask;
  uint64_t base;
  const uint64_t* pivot;  // pivots[0 .. pivots)
  unsigned pivots;
  __device__ __forceinline__ uint8_t key(uint64_t i) const {
    const uint64_t k = pt.window(base + i) & mask;
    unsigned lo = 0, hi = pivots;  // first pivot that is >= k
    while (lo < hi) {
      const unsigned mid = (lo + hi) >> 1;
      if (pivot[mid] < k)
        lo = mid + 1;
      else
        hi = mid;
    }
    return static_cast<uint8_t>(lo);
  }
  __device__ __forceinline__ IdxT val(uint64_t i) const { return static_cast<IdxT>(base + i); }
  static constexpr uint64_t bytes_read_per_item() { return 1; }
};

// Keys of the suffixes of a text slice for the dedicated partition kernels (partition.cuh).
struct SliceKeys {
  PackedText pt;
  uint64_t mask;
  uint64_t base;
  __device__ __forceinline__ uint64_t key(uint64_t i) const { return pt.window(base + i) & mask; }
};

// A device buffer straight from cudaMalloc (not from the arena): its address is the base of an
// allocation, which is what CUDA IPC can export to the other ranks' processes.
struct RawDeviceBuffer {
  void* ptr = nullptr;
  explicit RawDeviceBuffer(size_t bytes) {
    const cudaError_t err = cudaMalloc(&ptr, bytes ? bytes : 256);
    if (err != cudaSuccess) {
      cudaGetLastError();
      fail(std::string("out of device memory: cudaMalloc of the peer-visible bucket buffer failed: ") +
           cudaGetErrorString(err));
    }
  }
  ~RawDeviceBuffer() {
    if (ptr) cudaFree(ptr);
  }
  RawDeviceBuffer(const RawDeviceBuffer&) = delete;
  RawDeviceBuffer& operator=(const RawDeviceBuffer&) = delete;
};

}  // namespace

template <class IdxT>
void build_sa_lcp_sharded(Engine& eng, Comm& comm, const uint8_t* d_text, uint64_t n, ShardResult<IdxT>& out) {
  CAPSB_CUDA(cudaSetDevice(eng.dev.device));
  cudaStream_t st = eng.stream;
  const DeviceInfo& dev = eng.dev;
  const unsigned world = static_cast<unsigned>(comm.world);
  const unsigned rank = static_cast<unsigned>(comm.rank);
  const uint64_t launches_before = g_kernel_launches.load();
  const uint64_t comm_bytes_before = comm.bytes_sent;
  eng.stats = Stats();
  eng.radix.timer.reset();
  eng.msd_timers.reset();
  eng.stats.n = n;
  eng.stats.idx_bytes = sizeof(IdxT);
  out.offset = out.count = 0;
  out.sa.release();
  out.lcp.release();
  if (n == 0) return;
  if (sizeof(IdxT) == 4 && n > 0xFFFFFFFFull) fail("text too long for 32-bit indices");

  StageClock clock(eng);
  clock.mark("begin");  // 0

  // ---- text staging + key packing (whole text, every rank) -------------------------------
  PackedTextBuf packed = pack_text(eng, d_text, n);
  const PackedText pt = packed.view(n);
  const unsigned log2_bits = pt.log2_bits;
  eng.stats.bits_per_symbol = pt.bits();
  eng.stats.alphabet_size = packed.sigma;
  clock.mark("packed");  // 1

  const SliceMap map{n, n / world, world};
  const uint64_t lo = map.begin(rank);
  const uint64_t slice_count = map.end(rank) - lo;
  const unsigned key_bits = choose_key_bits(n);
  const uint64_t key_mask = key_mask_of(key_bits);
  eng.stats.key_bits = key_bits;
  // CAPSB_SHARD_MODE=merge: the reference's order of stages (sort the slices, exchange sorted
  // runs, merge).  Default: partition first (exchange suffix indices only, sort the bucket) —
  // no merge passes and a third of the NVLink bytes.
  const char* mode_env = std::getenv("CAPSB_SHARD_MODE");  // read per construction: the tests flip it
  const bool merge_mode = mode_env && std::string(mode_env) == "merge";
  if (world > static_cast<unsigned>(kRadixSize)) fail("more ranks than the bucket digit can hold");

  const uint64_t sample_total = static_cast<uint64_t>(kSamplesPerRank) * world;
  DevBuf<uint64_t> samples(kSamplesPerRank, st), all_a(sample_total, st), all_b(sample_total, st);
  DevBuf<uint32_t> dummy_a(sample_total, st), dummy_b(sample_total, st);
  const uint64_t* sorted_samples = nullptr;
  // every rank sorts the same all-gathered sample set and so picks the same G-1 pivots
  // (reference select_pivots, src/Suffix_Array.cpp:197-222): pivot j = sorted_samples[(j+1)*1024 - 1]
  auto agree_on_pivots = [&]() {
    CAPSB_CUDA(cudaMemsetAsync(dummy_a.get(), 0, sample_total * sizeof(uint32_t), st));
    comm.all_gather_device(samples.get(), all_a.get(), kSamplesPerRank * sizeof(uint64_t), st);
    const int sorted_in_b = radix_sort_pairs<uint64_t, uint32_t>(st, eng.radix, all_a.get(), dummy_a.get(), all_b.get(),
                                                                dummy_b.get(), sample_total, 64 - key_bits, 64);
    sorted_samples = sorted_in_b ? all_b.get() : all_a.get();
  };
  std::vector<uint64_t> send_counts(world), recv_counts(world), matrix(static_cast<size_t>(world) * world);
  uint64_t bucket_count = 0, bucket_offset = 0;
  // all-gather of the send counts -> the G x G count matrix (the reference's P, :481) -> what
  // this rank receives and where its bucket starts in the suffix array (part_size_scan_, :319-330)
  auto exchange_counts = [&]() {
    comm.all_gather_host(send_counts.data(), world * sizeof(uint64_t), matrix.data(), st);
    for (unsigned s = 0; s < world; ++s) {
      recv_counts[s] = matrix[static_cast<size_t>(s) * world + rank];
      bucket_count += recv_counts[s];
      for (unsigned q = 0; q < rank; ++q) bucket_offset += matrix[static_cast<size_t>(s) * world + q];
    }
    out.offset = bucket_offset;
    out.count = bucket_count;
    eng.stats.shard_offset = bucket_offset;
    eng.stats.shard_count = bucket_count;
  };
  DevBuf<uint64_t> bucket_keys;

  if (!merge_mode) {
    // ---- pivots from a sample of this rank's slice of the text --------------------------------
    {
      uint64_t* s = samples.get();
      const PackedText text = pt;
      launch_map(dev, st, kSamplesPerRank, [=] __device__(uint64_t t) {
        // scrambled positions (a regular stride could alias with a periodic text); an empty
        // slice contributes the largest key so it does not pull the pivots down
        uint64_t z = (t + 1) * 0x9E3779B97F4A7C15ull;
        z = (z ^ (z >> 29)) * 0xBF58476D1CE4E5B9ull;
        z ^= z >> 32;
        s[t] = slice_count ? (text.window(lo + z % slice_count) & key_mask) : ~0ull;
      });
    }
    agree_on_pivots();
    DevBuf<uint64_t> pivots(world, st);
    {
      uint64_t* pv = pivots.get();
      const uint64_t* ss = sorted_samples;
      launch_map(dev, st, world - 1, [=] __device__(uint64_t j) { pv[j] = ss[(j + 1) * kSamplesPerRank - 1]; });
    }
    // CAPSB_SHARD_P2P=1 (experimental, off by default): the partition pass and the exchange as ONE
    // kernel — the dedicated G-way partition (partition.cuh) stores every suffix index straight
    // into its owner's bucket buffer over NVLink instead of grouping the slice locally and
    // handing it to ncclSend/Recv.  Same bucket layout (runs by source rank, text order inside a
    // run), so everything after it is unchanged.
    const char* p2p_env = std::getenv("CAPSB_SHARD_P2P");
    const bool p2p = p2p_env && p2p_env[0] == '1' && world <= static_cast<unsigned>(kPartMaxBuckets);
    std::unique_ptr<RawDeviceBuffer> peer_bucket;  // the bucket's suffix indices in P2P mode
    DevBuf<IdxT> bucket_idx;
    const IdxT* bucket_suffixes = nullptr;
    if (p2p) {
      PartPivots piv{};
      piv.count = world - 1;
      if (world > 1) read_back(st, piv.p, pivots.get(), (world - 1) * sizeof(uint64_t));
      const SliceKeys slice_keys{pt, key_mask, lo};
      const Chunking ck = make_chunking(slice_count, kPartTile, static_cast<unsigned>(dev.sm_count) * 8);
      DevBuf<uint64_t> hist(static_cast<uint64_t>(kPartMaxBuckets) * ck.blocks, st), totals(kPartMaxBuckets, st);
      CAPSB_LAUNCH((partition_count_kernel<SliceKeys>), ck.blocks, kPartThreads, 0, st, slice_keys, slice_count, ck.chunk,
                   piv, hist.get());
      CAPSB_LAUNCH(partition_offsets_kernel, kPartMaxBuckets, 32, 0, st, hist.get(), ck.blocks, totals.get());
      uint64_t h_totals[kPartMaxBuckets];
      read_back(st, h_totals, totals.get(), sizeof(h_totals));
      for (unsigned q = 0; q < world; ++q) send_counts[q] = h_totals[q];
      exchange_counts();
      clock.mark("partitioned");  // 2
      peer_bucket = std::make_unique<RawDeviceBuffer>(bucket_count * sizeof(IdxT));
      const std::vector<void*> peers = comm.open_peer_buffers(peer_bucket->ptr, st);
      PartDestinations<IdxT> dst{};
      for (unsigned q = 0; q < static_cast<unsigned>(kPartMaxBuckets); ++q) {
        if (q >= world) {
          dst.ptr[q] = static_cast<IdxT*>(peer_bucket->ptr);  // never written: no key maps to these buckets
          continue;
        }
        uint64_t before = 0;  // what the ranks below this one send to rank q comes first in its bucket
        for (unsigned s = 0; s < rank; ++s) before += matrix[static_cast<size_t>(s) * world + q];
        dst.ptr[q] = static_cast<IdxT*>(peers[q]) + before;
        if (q != rank) comm.bytes_sent += send_counts[q] * sizeof(IdxT);
      }
      if (slice_count)
        CAPSB_LAUNCH((partition_scatter_kernel<IdxT, SliceKeys>), ck.blocks, kPartThreads, 0, st, slice_keys, slice_count,
                     ck.chunk, lo, piv, hist.get(), dst);
      comm.close_peer_buffers(peers, st);  // every rank's stores have landed in every bucket
      bucket_suffixes = static_cast<const IdxT*>(peer_bucket->ptr);
      clock.mark("exchanged");  // 3
    } else {
      // ---- partition: the slice's suffixes grouped by bucket, in text order inside a bucket ----
      DevBuf<IdxT> slice_idx(slice_count, st);
      {
        DevBuf<uint8_t> bucket_of(slice_count, st);
        CAPSB_CUDA(cudaMemsetAsync(eng.radix.digit_total.get(), 0, kRadixSize * sizeof(uint64_t), st));
        radix_pass<uint8_t, IdxT>(st, eng.radix, BucketSource<IdxT>{pt, key_mask, lo, pivots.get(), world - 1},
                                  slice_count, 0, bucket_of.get(), slice_idx.get());
        CAPSB_CUDA(cudaMemcpyAsync(send_counts.data(), eng.radix.digit_total.get(), world * sizeof(uint64_t),
                                   cudaMemcpyDeviceToHost, st));
        CAPSB_CUDA(cudaStreamSynchronize(st));
      }
      exchange_counts();
      clock.mark("partitioned");  // 2
      // ---- collate: suffix indices move to the rank that owns their bucket ----------------------
      bucket_idx.alloc(bucket_count, st);
      comm.all_to_all_v(slice_idx.get(), send_counts.data(), bucket_idx.get(), recv_counts.data(), sizeof(IdxT), st);
      slice_idx.release();
      bucket_suffixes = bucket_idx.get();
      clock.mark("exchanged");  // 3
    }
    // ---- bucket sort: keys come from the text again (each received run is in text order) ------
    bucket_keys.alloc(bucket_count, st);
    out.sa.alloc(bucket_count, st);
    sort_suffixes_by_key<IdxT>(eng, SuffixListSource<IdxT>{pt, key_mask, bucket_suffixes}, bucket_count, key_bits,
                               bucket_keys.get(), out.sa.get());
    bucket_idx.release();
    if (peer_bucket) {  // cudaFree waits for the sort that read it
      CAPSB_CUDA(cudaStreamSynchronize(st));
      peer_bucket.reset();
    }
    clock.mark("bucket sorted");  // 4
  } else {
    // ---- slice sort ------------------------------------------------------------------------
    DevBuf<uint64_t> slice_keys(slice_count, st);
    DevBuf<IdxT> slice_idx(slice_count, st);
    sort_suffix_slice<IdxT>(eng, pt, lo, slice_count, key_bits, slice_keys.get(), slice_idx.get());
    clock.mark("slice sorted");  // 2

    // ---- pivots: regular samples of the sorted slice (reference sample_pivots, :187-194) ------
    {
      uint64_t* s = samples.get();
      const uint64_t* k = slice_keys.get();
      launch_map(dev, st, kSamplesPerRank, [=] __device__(uint64_t t) {
        // an empty slice contributes the largest key so it does not pull the pivots down
        s[t] = slice_count ? k[((t + 1) * slice_count + kSamplesPerRank - 1) / kSamplesPerRank - 1] : ~0ull;
      });
    }
    agree_on_pivots();

    // ---- locate the pivots in the sorted slice -> send counts ----------------------------------
    DevBuf<uint64_t> d_bounds(world + 1, st);
    {
      const unsigned pivots = world - 1;
      const unsigned blocks = pivots ? static_cast<unsigned>(ceil_div(pivots, 8)) : 1;
      CAPSB_LAUNCH(locate_pivots_kernel, blocks, 256, 0, st, slice_keys.get(), slice_count, sorted_samples,
                   static_cast<uint64_t>(kSamplesPerRank), pivots, d_bounds.get());
    }
    std::vector<uint64_t> bounds(world + 1);
    CAPSB_CUDA(cudaMemcpyAsync(bounds.data(), d_bounds.get(), (world + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    CAPSB_CUDA(cudaStreamSynchronize(st));
    for (unsigned p = 0; p < world; ++p) send_counts[p] = bounds[p + 1] - bounds[p];
    exchange_counts();

    // ---- collate: (key, suffix) runs move to the rank that owns their bucket -----------------
    DevBuf<uint64_t> bucket_keys_tmp(bucket_count, st);
    DevBuf<IdxT> bucket_idx_tmp(bucket_count, st);
    bucket_keys.alloc(bucket_count, st);
    out.sa.alloc(bucket_count, st);
    comm.all_to_all_v(slice_keys.get(), send_counts.data(), bucket_keys.get(), recv_counts.data(), sizeof(uint64_t), st);
    comm.all_to_all_v(slice_idx.get(), send_counts.data(), out.sa.get(), recv_counts.data(), sizeof(IdxT), st);
    slice_keys.release();
    slice_idx.release();
    clock.mark("exchanged");  // 3

    // ---- bucket merge ------------------------------------------------------------------------
    std::vector<uint64_t> offsets(world + 1, 0);
    for (unsigned s = 0; s < world; ++s) offsets[s + 1] = offsets[s] + recv_counts[s];
    const int in_tmp = merge_sorted_runs<uint64_t, IdxT>(dev, st, bucket_keys.get(), out.sa.get(),
                                                        bucket_keys_tmp.get(), bucket_idx_tmp.get(), offsets);
    if (in_tmp) {
      std::swap(bucket_keys, bucket_keys_tmp);
      std::swap(out.sa, bucket_idx_tmp);
    }
    clock.mark("merged");  // 4
  }
  const uint64_t* keys = bucket_keys.get();
  IdxT* d_sa = out.sa.get();

  // ---- ties (collective: every rank runs the same number of rounds) ---------------------------
  out.lcp.alloc(bucket_count, st);
  IdxT* d_lcp = out.lcp.get();
  TiedSet<IdxT> tied;
  {
    ShardedRanks<IdxT> ranks(eng, comm, map, pt, key_mask_of(key_bits), sorted_samples, kSamplesPerRank, keys, d_sa,
                             key_bits >> log2_bits, bucket_count, bucket_offset);
    refine_tied_groups<IdxT>(eng, ranks, pt, key_bits, keys, d_sa, d_lcp, bucket_count, bucket_offset, n, tied);
  }
  eng.stats.tied_after_key_sort = tied.m;
  eng.sa_is_final(d_sa, bucket_offset, bucket_count, sizeof(IdxT));
  clock.mark("ties resolved");  // 5

  // ---- LCP ----------------------------------------------------------------------------------
  {
    // the suffix that precedes this bucket is the last one of the nearest non-empty bucket below
    // (the reference's partition-boundary patch, src/Suffix_Array.cpp:431-447)
    struct Edge {
      uint64_t count, last_key, last_idx;
    } mine{bucket_count, 0, 0};
    if (bucket_count) {
      IdxT last_idx;
      read_back(st, &mine.last_key, keys + bucket_count - 1, sizeof(uint64_t));
      read_back(st, &last_idx, d_sa + bucket_count - 1, sizeof(IdxT));
      mine.last_idx = last_idx;
    }
    std::vector<Edge> edges(world);
    comm.all_gather_host(&mine, sizeof(Edge), edges.data(), st);
    bool has_prev = false;
    uint64_t prev_key = 0, prev_idx = 0;
    for (unsigned q = 0; q < rank; ++q)
      if (edges[q].count) has_prev = true, prev_key = edges[q].last_key, prev_idx = edges[q].last_idx;
    first_position_lcp<IdxT>(eng, keys, d_sa, d_lcp, bucket_count, n, log2_bits, has_prev, prev_key, prev_idx);
  }
  {
    // tied neighbours (always inside one bucket) that the pair-chain step has not settled: the
    // pair (i = SA[k], j = SA[k-1]) goes to the rank that owns text position i
    DevBuf<IdxT> pair_i, pair_j, pair_k;
    const uint64_t deep_count = collect_deep_pairs<IdxT>(eng, tied, keys, d_sa, d_lcp, pair_i, pair_j, pair_k);
    const SliceMap mp = map;
    const IdxT* pi = pair_i.get();
    const IdxT* pj = pair_j.get();
    const IdxT* pk = pair_k.get();
    Route<IdxT> rt =
        plan_route<IdxT>(eng, comm, deep_count, [=] __device__(uint64_t t) -> unsigned { return mp.owner(pi[t]); });
    const uint64_t got = rt.recv_total;
    DevBuf<IdxPair<IdxT>> recv(got, st);
    route_forward<IdxPair<IdxT>>(
        eng, comm, rt, [=] __device__(uint64_t t) -> IdxPair<IdxT> { return IdxPair<IdxT>{pi[t], pj[t]}; }, recv.get());

    // order the received pairs by text position so the permuted-LCP chains are contiguous;
    // the predecessor j and the arrival slot travel with the position as the sort's value
    DevBuf<IdxT> pos_a(got, st), pos_b(got, st), answers(got, st);
    DevBuf<IdxPair<IdxT>> tag_a(got, st), tag_b(got, st);
    {
      const IdxPair<IdxT>* rv = recv.get();
      IdxT* pa = pos_a.get();
      IdxPair<IdxT>* ta = tag_a.get();
      launch_map(dev, st, got, [=] __device__(uint64_t j) {
        pa[j] = rv[j].a;
        ta[j] = IdxPair<IdxT>{rv[j].b, static_cast<IdxT>(j)};
      });
    }
    const unsigned pos_bits = round_up8(bit_length(n - 1));
    const int where = radix_sort_pairs<IdxT, IdxPair<IdxT>>(st, eng.radix, pos_a.get(), tag_a.get(), pos_b.get(),
                                                           tag_b.get(), got, 0, pos_bits);
    const IdxT* pos_i = where ? pos_b.get() : pos_a.get();
    const IdxPair<IdxT>* tag = where ? tag_b.get() : tag_a.get();
    {
      IdxT* ans = answers.get();
      plcp_for_pairs<IdxT>(
          eng, pt, pos_i, [=] __device__(uint64_t t) -> uint64_t { return tag[t].a; }, got, 0,
          [=] __device__(uint64_t t, IdxT lcp, uint64_t, IdxT) { ans[tag[t].b] = lcp; });
    }
    route_backward<IdxT>(eng, comm, rt, answers.get(),
                         [=] __device__(uint64_t t, IdxT lcp) { d_lcp[pk[t]] = lcp; });
  }
  clock.mark("lcp done");  // 6
  CAPSB_CUDA(cudaStreamSynchronize(st));

  eng.stats.ms_pack = clock.between(0, 1);
  if (merge_mode) {
    eng.stats.ms_sort = clock.between(1, 2);
    eng.stats.ms_partition = clock.between(2, 3);
    eng.stats.ms_merge = clock.between(3, 4);
  } else {
    eng.stats.ms_partition = clock.between(1, 3);  // pivots, partition pass, suffix all-to-all
    eng.stats.ms_sort = clock.between(3, 4);       // bucket sort
    eng.stats.ms_merge = 0;
  }
  eng.stats.ms_refine = clock.between(4, 5);
  eng.stats.ms_deep_lcp = clock.between(5, 6);
  eng.stats.ms_total = clock.between(0, 6);
  eng.stats.comm_bytes = comm.bytes_sent - comm_bytes_before;
  eng.stats.kernel_launches = g_kernel_launches.load() - launches_before;
  if (eng.radix.timer.enabled) {
    eng.stats.scatter_bytes = eng.radix.timer.bytes;
    eng.stats.ms_scatter = eng.radix.timer.drain(&eng.stats.scatter_launches);
  }
  collect_msd_timings(eng);
}

template void build_sa_lcp_sharded<uint32_t>(Engine&, Comm&, const uint8_t*, uint64_t, ShardResult<uint32_t>&);
template void build_sa_lcp_sharded<uint64_t>(Engine&, Comm&, const uint8_t*, uint64_t, ShardResult<uint64_t>&);

}  // namespace capsb
