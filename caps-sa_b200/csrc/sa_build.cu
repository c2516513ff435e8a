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
#include "engine.cuh"

namespace capsb {

std::atomic<uint64_t> g_kernel_launches{0};

Engine::Engine(int device) {
  dev.device = device;
  CAPSB_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  CAPSB_CUDA(cudaGetDeviceProperties(&prop, device));
  dev.sm_count = prop.multiProcessorCount;
  CAPSB_CUDA(cudaStreamCreateWithFlags(&own_stream, cudaStreamNonBlocking));
  stream = own_stream;
  // keep freed scratch cached in the pool between constructions
  cudaMemPool_t pool;
  CAPSB_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
  uint64_t threshold = ~0ull;
  CAPSB_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold));
  radix.init(dev, stream);
  scan32.init(dev, stream);
  scan64.init(dev, stream);
}

Engine::~Engine() {
  cudaSetDevice(dev.device);
  if (own_stream) cudaStreamSynchronize(own_stream);
  radix = RadixScratch();
  scan32 = ScanScratch<uint32_t>();
  scan64 = ScanScratch<uint64_t>();
  for (cudaEvent_t e : events) cudaEventDestroy(e);
  if (own_stream) cudaStreamDestroy(own_stream);
}

namespace {

template <class IdxT>
struct IdxTraits;
template <>
struct IdxTraits<uint32_t> {
  using Comp = uint64_t;  // (group head + in-range bit) << 32 | second rank
  static constexpr unsigned kField = 32;
};
template <>
struct IdxTraits<uint64_t> {
  using Comp = unsigned __int128;
  static constexpr unsigned kField = 64;
};

inline unsigned bit_length(uint64_t v) {
  unsigned b = 0;
  while (v) ++b, v >>= 1;
  return b ? b : 1;
}
inline unsigned round_up8(unsigned b) { return (b + 7u) & ~7u; }

// First radix pass reads keys from the packed text: key = window at suffix i, value = i.
template <class IdxT>
struct TextSource {
  PackedText pt;
  __device__ __forceinline__ uint64_t key(uint64_t i) const { return pt.window(i); }
  __device__ __forceinline__ IdxT val(uint64_t i) const { return static_cast<IdxT>(i); }
  // the text window is re-read from L1/L2; compulsory traffic is the packed text itself (< 1 B)
  static constexpr uint64_t bytes_read_per_item() { return 1; }
};

// Stage timer: records an event now; elapsed times are read at the end.
struct StageClock {
  Engine& eng;
  std::vector<cudaEvent_t> marks;
  explicit StageClock(Engine& e) : eng(e) {}
  void mark() {
    cudaEvent_t ev;
    if (!eng.events.empty()) {
      ev = eng.events.back();
      eng.events.pop_back();
    } else {
      CAPSB_CUDA(cudaEventCreate(&ev));
    }
    CAPSB_CUDA(cudaEventRecord(ev, eng.stream));
    marks.push_back(ev);
  }
  float between(size_t a, size_t b) {
    float ms = 0;
    CAPSB_CUDA(cudaEventElapsedTime(&ms, marks[a], marks[b]));
    return ms;
  }
  ~StageClock() {
    for (cudaEvent_t e : marks) eng.events.push_back(e);
  }
};

// Scan in two steps so the total (a count) is known before the outputs are allocated.
template <class T, class Op, class In>
T scan_total(Engine& eng, uint64_t n, In in) {
  ScanScratch<T>& sc = eng.scan_scratch<T>();
  if (n == 0) return Op::template identity<T>();
  const Chunking ck = make_chunking(n, kScanTile, sc.max_blocks);
  CAPSB_LAUNCH((scan_reduce_kernel<T, Op, In>), ck.blocks, kScanThreads, 0, eng.stream, n, ck.chunk, in,
               sc.partial.get());
  CAPSB_LAUNCH((scan_spine_kernel<T, Op>), 1, kScanThreads, 0, eng.stream, ck.blocks, sc.partial.get(),
               sc.total.get());
  T total;
  CAPSB_CUDA(cudaMemcpyAsync(&total, sc.total.get(), sizeof(T), cudaMemcpyDeviceToHost, eng.stream));
  CAPSB_CUDA(cudaStreamSynchronize(eng.stream));
  return total;
}
// Must directly follow scan_total / another scan of the same n (reuses the partials).
template <class T, class Op, bool Inclusive, class In, class Out>
void scan_finish(Engine& eng, uint64_t n, In in, Out out) {
  ScanScratch<T>& sc = eng.scan_scratch<T>();
  if (n == 0) return;
  const Chunking ck = make_chunking(n, kScanTile, sc.max_blocks);
  CAPSB_LAUNCH((scan_apply_kernel<T, Op, Inclusive, In, Out>), ck.blocks, kScanThreads, 0, eng.stream, n,
               ck.chunk, in, out, sc.partial.get());
}
template <class T, class Op, bool Inclusive, class In, class Out>
void scan_full(Engine& eng, uint64_t n, In in, Out out) {
  device_scan<T, Op, Inclusive>(eng.dev, eng.stream, eng.scan_scratch<T>(), n, in, out);
}

// Block-wide comparison for the few very long common prefixes (one CTA per pair).
template <class IdxT>
__global__ void __launch_bounds__(256) long_lcp_kernel(PackedText pt, const uint64_t* __restrict__ pos_i,
                                                       const IdxT* __restrict__ pos_j,
                                                       const IdxT* __restrict__ todo, uint64_t todo_count,
                                                       IdxT* __restrict__ plcp) {
  __shared__ unsigned long long best;
  constexpr int kPerThread = 4;
  const unsigned spw = pt.syms_per_word();
  for (uint64_t e = blockIdx.x; e < todo_count; e += gridDim.x) {
    const uint64_t t = todo[e];
    const uint64_t i = pos_i[t], j = pos_j[t];
    const uint64_t shorter = pt.n - (i > j ? i : j);
    uint64_t base = plcp[t];  // symbols already known equal (multiple of spw)
    while (true) {
      if (threadIdx.x == 0) best = ~0ull;
      __syncthreads();
      unsigned long long mine = ~0ull;
#pragma unroll
      for (int q = 0; q < kPerThread; ++q) {
        const uint64_t off = base + (static_cast<uint64_t>(q) * 256 + threadIdx.x) * spw;
        if (off >= shorter) {
          if (shorter < mine) mine = shorter;
        } else {
          const uint64_t x = pt.window(i + off) ^ pt.window(j + off);
          if (x != 0) {
            const uint64_t l = off + (static_cast<uint64_t>(__clzll(static_cast<long long>(x))) >> pt.log2_bits);
            if (l < mine) mine = l;
          }
        }
      }
      if (mine != ~0ull) atomicMin(&best, mine);
      __syncthreads();
      const unsigned long long got = best;
      __syncthreads();
      if (got != ~0ull) {
        if (threadIdx.x == 0) plcp[t] = static_cast<IdxT>(got < shorter ? got : shorter);
        break;
      }
      base += static_cast<uint64_t>(kPerThread) * 256 * spw;
    }
  }
}

// LCP of neighbours with different keys comes from the keys alone: clz(key_a ^ key_b) / bits,
// bounded by the shorter suffix.  It needs the FINAL predecessor (the bound depends on which
// member of the previous group ends up last), so it runs after the refinement.
template <class IdxT>
void key_lcp(Engine& eng, const uint64_t* keys, const IdxT* d_sa, IdxT* d_lcp, uint64_t n, unsigned log2_bits) {
  launch_map(eng.dev, eng.stream, n, [=] __device__(uint64_t k) {
    if (k == 0) {
      d_lcp[0] = 0;
      return;
    }
    const uint64_t x = keys[k] ^ keys[k - 1];
    if (x != 0) {
      const uint64_t a = d_sa[k - 1], b = d_sa[k];
      const uint64_t shorter = n - (a > b ? a : b);
      const uint64_t l = static_cast<uint64_t>(__clzll(static_cast<long long>(x))) >> log2_bits;
      d_lcp[k] = static_cast<IdxT>(l < shorter ? l : shorter);
    }
  });
}

}  // namespace

template <class IdxT>
void build_sa_lcp(Engine& eng, const uint8_t* d_text, uint64_t n, IdxT* d_sa, IdxT* d_lcp) {
  using Comp = typename IdxTraits<IdxT>::Comp;
  constexpr unsigned kField = IdxTraits<IdxT>::kField;
  CAPSB_CUDA(cudaSetDevice(eng.dev.device));
  cudaStream_t st = eng.stream;
  const DeviceInfo& dev = eng.dev;
  const uint64_t launches_before = g_kernel_launches.load();
  eng.stats = Stats();
  eng.radix.timer.reset();
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
  const unsigned spw = pt.syms_per_word();
  eng.stats.bits_per_symbol = pt.bits();
  eng.stats.alphabet_size = packed.sigma;
  clock.mark();  // 1

  // ---- 2. key sort of all suffixes (64-bit packed prefix, 8 stable passes) ---------------
  DevBuf<uint64_t> key_a(n, st), key_b(n, st);
  DevBuf<IdxT> val_a(n, st);
  {
    radix_pass<uint64_t, IdxT>(st, eng.radix, TextSource<IdxT>{pt}, n, 0, key_a.get(), val_a.get());
    uint64_t* kin = key_a.get();
    IdxT* vin = val_a.get();
    uint64_t* kout = key_b.get();
    IdxT* vout = d_sa;
    for (unsigned shift = 8; shift < 64; shift += 8) {
      radix_pass<uint64_t, IdxT>(st, eng.radix, ArraySource<uint64_t, IdxT>{kin, vin}, n, shift, kout, vout);
      std::swap(kin, kout);
      std::swap(vin, vout);
    }
    // 8 passes: the last one wrote (key_b, d_sa); after the final swap kin/vin point there.
    if (kin != key_b.get() || vin != d_sa) fail("internal: radix ping-pong parity");
  }
  key_a.release();
  val_a.release();
  const uint64_t* keys = key_b.get();
  clock.mark();  // 2

  // ---- 3. count the suffixes whose key ties with their predecessor --------------------------
  auto tied = [=] __device__(uint64_t k) -> uint64_t { return (k > 0 && keys[k] == keys[k - 1]) ? 1u : 0u; };
  const uint64_t ties = scan_total<uint64_t, OpSum>(eng, n, tied);
  eng.stats.tied_after_key_sort = ties;
  clock.mark();  // 3

  if (ties == 0) {
    key_lcp<IdxT>(eng, keys, d_sa, d_lcp, n, log2_bits);
    clock.mark();  // 4
    clock.mark();  // 5
  } else {
    // ---- 4. prefix-doubling refinement of the tied groups --------------------------------
    {
      DevBuf<IdxT> isa(n, st);
      DevBuf<IdxT> group_of(n, st);  // SA position -> first SA position of its group
      IdxT* d_isa = isa.get();
      IdxT* d_group = group_of.get();
      scan_full<IdxT, OpMax, true>(
          eng, n,
          [=] __device__(uint64_t k) -> IdxT { return (k > 0 && keys[k] != keys[k - 1]) ? static_cast<IdxT>(k) : IdxT(0); },
          [=] __device__(uint64_t k, IdxT head) {
            d_group[k] = head;
            d_isa[d_sa[k]] = head;
          });

      auto in_group = [=] __device__(uint64_t k) -> IdxT {
        const bool single = d_group[k] == k && (k + 1 == n || d_group[k + 1] == k + 1);
        return single ? IdxT(0) : IdxT(1);
      };
      uint64_t m = scan_total<IdxT, OpSum>(eng, n, in_group);
      DevBuf<IdxT> a_pos(m, st), a_idx(m, st), a_group(m, st);
      {
        IdxT* p = a_pos.get();
        IdxT* s = a_idx.get();
        IdxT* g = a_group.get();
        scan_finish<IdxT, OpSum, false>(eng, n, in_group, [=] __device__(uint64_t k, IdxT slot) {
          const bool single = d_group[k] == k && (k + 1 == n || d_group[k + 1] == k + 1);
          if (!single) {
            p[slot] = static_cast<IdxT>(k);
            s[slot] = d_sa[k];
            g[slot] = d_group[k];
          }
        });
      }
      group_of.release();

      const unsigned rank_bits = round_up8(bit_length(n - 1));
      uint64_t h = spw;  // the key sort ordered the suffixes by their first `spw` symbols
      while (m > 0) {
        eng.stats.refine_rounds++;
        DevBuf<Comp> comp_a(m, st), comp_b(m, st);
        DevBuf<IdxT> idx_b(m, st), head_slot(m, st);
        Comp* ca = comp_a.get();
        {
          const IdxT* s = a_idx.get();
          const IdxT* g = a_group.get();
          launch_map(dev, st, m, [=] __device__(uint64_t t) {
            const uint64_t i = s[t];
            const uint64_t ih = i + h;
            const bool inside = ih < n;
            // beyond the end: shorter suffix first, i.e. larger position first
            const uint64_t second = inside ? static_cast<uint64_t>(d_isa[ih]) : (n - 1 - i);
            ca[t] = (static_cast<Comp>(static_cast<uint64_t>(g[t]) + (inside ? 1u : 0u)) << kField) |
                    static_cast<Comp>(second);
          });
        }
        // sort by (group, second rank): LSD over the second-rank field, then the group field
        Comp* kin = comp_a.get();
        IdxT* vin = a_idx.get();
        Comp* kout = comp_b.get();
        IdxT* vout = idx_b.get();
        for (unsigned field = 0; field < 2; ++field)
          for (unsigned shift = field * kField; shift < field * kField + rank_bits; shift += 8) {
            radix_pass<Comp, IdxT>(st, eng.radix, ArraySource<Comp, IdxT>{kin, vin}, m, shift, kout, vout);
            std::swap(kin, kout);
            std::swap(vin, vout);
          }
        const Comp* sorted_comp = kin;
        const IdxT* sorted_idx = vin;

        IdxT* hs = head_slot.get();
        scan_full<IdxT, OpMax, true>(
            eng, m,
            [=] __device__(uint64_t t) -> IdxT {
              return (t > 0 && sorted_comp[t] != sorted_comp[t - 1]) ? static_cast<IdxT>(t) : IdxT(0);
            },
            [=] __device__(uint64_t t, IdxT head) { hs[t] = head; });

        DevBuf<IdxT> new_group(m, st);
        {
          const IdxT* p = a_pos.get();
          IdxT* ng = new_group.get();
          launch_map(dev, st, m, [=] __device__(uint64_t t) {
            const IdxT suffix = sorted_idx[t];
            const IdxT head_pos = p[hs[t]];
            d_sa[p[t]] = suffix;
            d_isa[suffix] = head_pos;
            ng[t] = head_pos;
          });
        }
        auto still_tied = [=] __device__(uint64_t t) -> IdxT {
          const bool single = hs[t] == t && (t + 1 == m || hs[t + 1] == t + 1);
          return single ? IdxT(0) : IdxT(1);
        };
        const uint64_t m_next = scan_total<IdxT, OpSum>(eng, m, still_tied);
        DevBuf<IdxT> n_pos(m_next, st), n_idx(m_next, st), n_group(m_next, st);
        if (m_next > 0) {
          const IdxT* p = a_pos.get();
          const IdxT* ng = new_group.get();
          IdxT* np = n_pos.get();
          IdxT* ns = n_idx.get();
          IdxT* ngp = n_group.get();
          scan_finish<IdxT, OpSum, false>(eng, m, still_tied, [=] __device__(uint64_t t, IdxT slot) {
            const bool single = hs[t] == t && (t + 1 == m || hs[t + 1] == t + 1);
            if (!single) {
              np[slot] = p[t];
              ns[slot] = sorted_idx[t];
              ngp[slot] = ng[t];
            }
          });
        }
        a_pos = std::move(n_pos);
        a_idx = std::move(n_idx);
        a_group = std::move(n_group);
        m = m_next;
        if (h > (~0ull >> 2)) fail("internal: refinement did not converge");
        h <<= 1;
      }
    }
    key_lcp<IdxT>(eng, keys, d_sa, d_lcp, n, log2_bits);
    clock.mark();  // 4

    // ---- 5. LCP of the tied neighbours: permuted-LCP recurrence on the deep positions ----
    {
      const uint64_t m = ties;
      DevBuf<uint64_t> pos_a(m, st), pos_b(m, st);  // text position i of the later suffix
      DevBuf<IdxT> rank_a(m, st), rank_b(m, st);     // its SA position k
      {
        uint64_t* pa = pos_a.get();
        IdxT* ra = rank_a.get();
        scan_full<uint64_t, OpSum, false>(eng, n, tied, [=] __device__(uint64_t k, uint64_t slot) {
          if (k > 0 && keys[k] == keys[k - 1]) {
            pa[slot] = d_sa[k];
            ra[slot] = static_cast<IdxT>(k);
          }
        });
      }
      const unsigned pos_bits = round_up8(bit_length(n - 1));
      const int where = radix_sort_pairs<uint64_t, IdxT>(st, eng.radix, pos_a.get(), rank_a.get(), pos_b.get(),
                                                        rank_b.get(), m, 0, pos_bits);
      const uint64_t* pos_i = where ? pos_b.get() : pos_a.get();
      const IdxT* sa_rank = where ? rank_b.get() : rank_a.get();

      DevBuf<IdxT> pos_j(m, st), plcp(m, st), todo(m, st), chain_head(m, st);
      DevBuf<unsigned long long> counters(2, st);
      CAPSB_CUDA(cudaMemsetAsync(counters.get(), 0, 2 * sizeof(unsigned long long), st));
      {
        IdxT* pj = pos_j.get();
        IdxT* pl = plcp.get();
        IdxT* td = todo.get();
        IdxT* ch = chain_head.get();
        unsigned long long* cnt = counters.get();
        launch_map(dev, st, m, [=] __device__(uint64_t t) {
          const uint64_t i = pos_i[t];
          const uint64_t j = d_sa[sa_rank[t] - 1];
          pj[t] = static_cast<IdxT>(j);
          // reducible: the pair (i-1, j-1) precedes it in the chain and the preceding symbols
          // agree, so PLCP[i] = PLCP[i-1] - 1 (Karkkainen-Manzini-Puglisi).
          const bool chained = t > 0 && pos_i[t - 1] + 1 == i && i > 0 && j > 0 &&
                               pt.symbol(i - 1) == pt.symbol(j - 1);
          ch[t] = chained ? IdxT(0) : static_cast<IdxT>(t);
          if (!chained) {
            uint64_t l = 0;
            const bool done = pt.common_prefix(i, j, 0, 16, &l);
            pl[t] = static_cast<IdxT>(l);
            atomicAdd(cnt + 0, 1ull);
            if (!done) td[atomicAdd(cnt + 1, 1ull)] = static_cast<IdxT>(t);
          }
        });
      }
      unsigned long long h_cnt[2];
      CAPSB_CUDA(cudaMemcpyAsync(h_cnt, counters.get(), sizeof(h_cnt), cudaMemcpyDeviceToHost, st));
      CAPSB_CUDA(cudaStreamSynchronize(st));
      eng.stats.deep_lcp_direct = h_cnt[0];
      eng.stats.deep_lcp_long = h_cnt[1];
      if (h_cnt[1] > 0) {
        const unsigned grid = static_cast<unsigned>(std::min<uint64_t>(h_cnt[1], static_cast<uint64_t>(dev.sm_count) * 8));
        CAPSB_LAUNCH((long_lcp_kernel<IdxT>), grid, 256, 0, st, pt, pos_i, pos_j.get(), todo.get(),
                     static_cast<uint64_t>(h_cnt[1]), plcp.get());
      }
      {
        const IdxT* pl = plcp.get();
        const IdxT* ch = chain_head.get();
        scan_full<IdxT, OpMax, true>(
            eng, m, [=] __device__(uint64_t t) -> IdxT { return ch[t]; },
            [=] __device__(uint64_t t, IdxT head) {
              const uint64_t back = pos_i[t] - pos_i[head];
              d_lcp[sa_rank[t]] = static_cast<IdxT>(static_cast<uint64_t>(pl[head]) - back);
            });
      }
    }
    clock.mark();  // 5
  }

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
