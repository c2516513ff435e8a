/* caps_sa_gpu.h — C-ABI of the B200-native CaPS-SA construction engine.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++ or torch
 * types.  The C++ class shell in include/Suffix_Array.hpp (same public surface as the
 * reference's CaPS_SA::Suffix_Array<idx_t>, reference include/Suffix_Array.hpp:22-181) and
 * the Python mirror in caps-sa_b200/ bind exactly these entry points.  There is no CPU
 * fallback: every function fails (non-zero return + caps_sa_gpu_last_error()) when no CUDA
 * device / kernel image is available.
 *
 * Semantics (identical to the reference at its default, unbounded context):
 *   SA  = suffix array of text[0..n) under `signed char` order (reference
 *         src/Suffix_Array.cpp:77,289), the shorter suffix first when one is a prefix of
 *         the other (:71,76);
 *   LCP[0] = 0, LCP[k] = lcp(text[SA[k-1]..], text[SA[k]..]).
 * Results do not depend on subproblem_count (a tuning hint here, as in the reference the
 * output is independent of p — SURVEY.md §0).
 */
#ifndef CAPS_SA_GPU_H
#define CAPS_SA_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CAPS_SA_GPU_OK 0
#define CAPS_SA_GPU_ERR_ARGS 1     /* bad arguments (NULL pointers, n too large for the index width) */
#define CAPS_SA_GPU_ERR_CUDA 2     /* CUDA / NCCL failure, including out of device memory */
#define CAPS_SA_GPU_ERR_UNSUPPORTED 3 /* reserved; nothing returns it at present */

/* Opaque per-device engine: stream, scratch pools, statistics. */
typedef struct caps_sa_gpu_engine caps_sa_gpu_engine;

/* Statistics of the last construction on an engine (times in ms, CUDA events on the
 * engine's stream). */
typedef struct caps_sa_gpu_stats {
  uint64_t n;
  uint32_t idx_bytes;
  uint32_t bits_per_symbol;     /* 1, 2, 4 or 8: width of the packed codes */
  uint32_t alphabet_size;
  uint32_t refine_rounds;       /* text and rank (prefix-doubling) rounds run on the tied suffixes */
  uint64_t tied_after_key_sort; /* suffixes whose sort key is shared with another suffix */
  uint64_t deep_lcp_direct;     /* irreducible deep LCPs computed by direct comparison */
  uint64_t deep_lcp_long;       /* ... that needed the block-wide comparison */
  uint64_t kernel_launches;     /* launches of this library's kernels */
  float ms_pack, ms_sort, ms_heads, ms_refine, ms_deep_lcp, ms_total;
  float ms_h2d, ms_d2h;         /* host-buffer entry points only */
  /* Dominant kernel (radix_scatter_kernel), measured with CUDA events around each launch when
   * kernel timing is enabled (caps_sa_gpu_engine_set_kernel_timing): */
  uint32_t scatter_launches;
  float ms_scatter;             /* summed duration of those launches */
  uint64_t scatter_bytes;       /* summed algorithmic bytes (keys+values read once, written once) */
  uint32_t key_bits;            /* leading bits of the packed prefix used as the sort key */
  uint32_t reserved;
  /* sharded construction only: */
  float ms_partition;           /* pivots, partition pass and suffix all-to-all (merge mode: pivot location and
                                   the (key, suffix) all-to-all) */
  float ms_merge;               /* merge-path tree over the received runs (merge mode only) */
  uint64_t comm_bytes;          /* bytes this rank moved to other ranks */
  uint64_t shard_offset;        /* this rank owns SA/LCP positions [shard_offset, shard_offset + shard_count) */
  uint64_t shard_count;
  uint64_t pairs_chained;       /* groups of two suffixes finished by the pair-chain step (order + LCP) */
  /* Key sort on packed 8-byte records, most significant digit first (32-bit indices): level A
   * partitions the suffixes by the top msd_a_bits of their key straight from the text, level B every
   * level-A bucket by the next msd_b_bits, the local sort orders each bucket in shared memory.
   * Per kernel class, with kernel timing enabled: summed launch durations, algorithmic bytes
   * (bytes each record is read and written with, once per launch) and launch counts. */
  uint32_t msd_a_bits, msd_b_bits;
  uint32_t msd_large_buckets;   /* buckets too large for shared memory, sorted by the LSD passes instead */
  uint32_t msd_reserved;
  uint64_t msd_large_records;
  float ms_msd_scatter_a, ms_msd_scatter_b, ms_msd_local, ms_msd_hist;
  uint64_t msd_scatter_a_bytes, msd_scatter_b_bytes, msd_local_bytes, msd_hist_bytes;
  uint32_t msd_scatter_a_launches, msd_scatter_b_launches, msd_local_launches, msd_hist_launches;
} caps_sa_gpu_stats;

/* Number of CUDA devices visible to the library (0 if none / driver missing). */
int caps_sa_gpu_device_count(void);

/* Message of the last failure on the calling thread ("" if none). */
const char* caps_sa_gpu_last_error(void);

/* Engine lifetime.  caps_sa_gpu_engine_create returns NULL on failure. */
caps_sa_gpu_engine* caps_sa_gpu_engine_create(int device);
void caps_sa_gpu_engine_destroy(caps_sa_gpu_engine* engine);
int caps_sa_gpu_engine_stats(const caps_sa_gpu_engine* engine, caps_sa_gpu_stats* out);

/* Run all of the engine's work on `stream` (a cudaStream_t; NULL = the engine's own stream),
 * so that a caller's CUDA events on that stream bracket the work. */
int caps_sa_gpu_engine_set_stream(caps_sa_gpu_engine* engine, void* stream);

/* Enable/disable per-launch CUDA-event timing of the dominant kernel (off by default). */
int caps_sa_gpu_engine_set_kernel_timing(caps_sa_gpu_engine* engine, int enabled);

/* ---- Host-buffer construction: what Suffix_Array<idx_t>::construct() binds --------------
 * Replaces the reference's construct() (src/Suffix_Array.cpp:466-494).  `text` is borrowed
 * host memory of n bytes; sa_out / lcp_out are caller-owned host arrays of n entries
 * (pinned memory from caps_sa_gpu_host_alloc makes the copies run at PCIe speed).
 * Blocking.  max_context: 0 or >= n = the reference's default (unbounded).  A bounded context
 * (reference ctor argument 4, src/Suffix_Array.cpp:25,72-77) is implementation-defined in the
 * reference — the order inside groups of suffixes that agree on their first max_context + 1
 * symbols depends on its subproblem count.  Here the suffix array is always the exact one
 * (one of the orders the bounded comparison allows) and the LCP entries are
 * min(lcp, max_context), the reference's values everywhere except at its p - 1 partition
 * boundaries (:440).  subproblem_count is accepted for compatibility and ignored. */
int caps_sa_gpu_construct_u32(caps_sa_gpu_engine* engine, const char* text, uint64_t n,
                              uint32_t* sa_out, uint32_t* lcp_out, uint64_t subproblem_count,
                              uint64_t max_context);
int caps_sa_gpu_construct_u64(caps_sa_gpu_engine* engine, const char* text, uint64_t n,
                              uint64_t* sa_out, uint64_t* lcp_out, uint64_t subproblem_count,
                              uint64_t max_context);

/* ---- Device-resident construction --------------------------------------------------------
 * d_text (n bytes), d_sa, d_lcp (n entries each) are device pointers on the engine's
 * device; `stream` is a cudaStream_t (NULL = the engine's own stream).  The call returns
 * after the work has completed on that stream (it contains small device-to-host reads). */
int caps_sa_gpu_construct_device_u32(caps_sa_gpu_engine* engine, const void* d_text, uint64_t n,
                                     uint32_t* d_sa, uint32_t* d_lcp, void* stream);
int caps_sa_gpu_construct_device_u64(caps_sa_gpu_engine* engine, const void* d_text, uint64_t n,
                                     uint64_t* d_sa, uint64_t* d_lcp, void* stream);

/* ---- Sharded construction: several GPUs, one rank each -------------------------------------
 * The samplesort shape of the reference (src/Suffix_Array.cpp:466-494) across ranks: the text
 * is replicated; pivots are agreed from all-gathered key samples of every rank's slice of the text
 * (select_pivots, :197-222); one counting pass groups the slice's suffixes by bucket
 * (locate_pivots + partition_sub_subarrays, :225-368); suffix indices move to the rank that owns
 * their bucket in one all-to-all (the collate step, :335-364); each rank key-sorts its bucket and
 * resolves its ties (sort_subarrays / merge_sub_subarrays, :161-184, :371-428); the LCP at bucket
 * boundaries comes from the neighbouring rank (compute_partition_boundary_lcp, :431-447).
 * CAPSB_SHARD_MODE=merge selects the reference's own order of stages instead (sort the slices, locate
 * the pivots in them, exchange sorted (key, suffix) runs, merge).  Rank r ends up owning the
 * contiguous range [shard_offset, shard_offset + shard_count) of SA and LCP.
 *
 * (1) One process, one host thread per rank (what Suffix_Array<idx_t>::construct() binds when
 *     CAPS_SA_GPUS > 1).  devices[r] is the CUDA device of rank r; listing a device more than
 *     once runs several ranks on it (used by the parity tests on a one-GPU box).  Ranks
 *     exchange data by peer copies (NVLink when the devices differ).  text, sa_out, lcp_out
 *     are host buffers as in caps_sa_gpu_construct_u32; every rank writes its shard into
 *     sa_out / lcp_out.  stats_out (may be NULL) receives num_ranks entries. */
int caps_sa_gpu_construct_multi_u32(const int* devices, int num_ranks, const char* text, uint64_t n,
                                    uint32_t* sa_out, uint32_t* lcp_out, uint64_t subproblem_count,
                                    uint64_t max_context, caps_sa_gpu_stats* stats_out);
int caps_sa_gpu_construct_multi_u64(const int* devices, int num_ranks, const char* text, uint64_t n,
                                    uint64_t* sa_out, uint64_t* lcp_out, uint64_t subproblem_count,
                                    uint64_t max_context, caps_sa_gpu_stats* stats_out);

/* (2) One process per GPU (torchrun): the ranks join an NCCL communicator and exchange over
 *     NVLink / NVSwitch.  Rank 0 creates an id (128 bytes) and the host side broadcasts it (e.g.
 *     torch.distributed); every rank then calls caps_sa_gpu_engine_comm_init, a collective. */
#define CAPS_SA_GPU_COMM_ID_BYTES 128
int caps_sa_gpu_comm_unique_id(void* id_out);
int caps_sa_gpu_engine_comm_init(caps_sa_gpu_engine* engine, const void* id, int rank, int world);

/* Collective over the engine's communicator.  d_text is this rank's device copy of the whole
 * text.  The shard stays in device memory owned by the engine until the next construction;
 * caps_sa_gpu_engine_stats reports shard_offset / shard_count. */
int caps_sa_gpu_construct_sharded_device_u32(caps_sa_gpu_engine* engine, const void* d_text, uint64_t n,
                                             void* stream);
int caps_sa_gpu_construct_sharded_device_u64(caps_sa_gpu_engine* engine, const void* d_text, uint64_t n,
                                             void* stream);
/* Same with host buffers: stages the text (H2D), constructs, and copies this rank's shard to
 * sa_out + shard_offset / lcp_out + shard_offset (the arrays have n entries; other ranges are
 * left untouched). */
int caps_sa_gpu_construct_sharded_u32(caps_sa_gpu_engine* engine, const char* text, uint64_t n,
                                      uint32_t* sa_out, uint32_t* lcp_out);
int caps_sa_gpu_construct_sharded_u64(caps_sa_gpu_engine* engine, const char* text, uint64_t n,
                                      uint64_t* sa_out, uint64_t* lcp_out);
/* Copies the last shard (shard_count entries each) to sa_dst / lcp_dst: host memory when
 * to_host != 0, else device memory on the engine's device. */
int caps_sa_gpu_shard_copy(caps_sa_gpu_engine* engine, void* sa_dst, void* lcp_dst, int to_host);

/* ---- CLI byte mapping ----------------------------------------------------------------------
 * In-place text[i] = "ACTG"[(toupper(text[i]) & 6) >> 1] for every byte, run on the device
 * (reference src/main.cpp:61-70).  `text` is host memory. */
int caps_sa_gpu_map_acgt(caps_sa_gpu_engine* engine, char* text, uint64_t n);

/* The same mapping fused into the staging of the host-buffer construction calls: when enabled
 * (process-wide), every caps_sa_gpu_construct_* / _multi_* / _sharded_* call with a host text applies
 * it to its device copy of the text right after the upload, so a CLI need not push the text
 * through PCIe three times (up and down for the mapping, up again for the construction).  The
 * caller's host text is left as it was.  Returns the previous setting. */
int caps_sa_gpu_set_cli_byte_mapping(int enabled);

/* ---- Pinned host memory ---------------------------------------------------------------------
 * The class shell allocates SA_/LCP_ with these (reference: malloc in the ctor,
 * src/Suffix_Array.cpp:20-21, include/Suffix_Array.hpp:137-142). */
void* caps_sa_gpu_host_alloc(size_t bytes);
void caps_sa_gpu_host_free(void* ptr);

/* ---- Stage-level entry points (parity tests of the individual kernels) ----------------------
 * All operate on host arrays and run the production kernels on the engine's device. */

/* Packs `text` and returns bits per symbol (1/2/4/8) or a negative error; words_out must
 * hold ceil(n*8/64)+2 entries (enough for any width); *nwords_out receives the count. */
int caps_sa_gpu_stage_pack(caps_sa_gpu_engine* engine, const char* text, uint64_t n,
                           uint64_t* words_out, uint64_t* nwords_out, uint32_t* alphabet_size_out);

/* Stable LSD radix sort of (keys, vals) on key bits [begin_bit, end_bit), in place. */
int caps_sa_gpu_stage_radix_sort_u64_u32(caps_sa_gpu_engine* engine, uint64_t* keys, uint32_t* vals,
                                         uint64_t n, unsigned begin_bit, unsigned end_bit);

/* Key sort of all suffixes of `text` (the stage that replaces permute + sort_subarrays, reference
 * src/Suffix_Array.cpp:148-184): keys_out[k] = the leading key bits of suffix sa_out[k] (left-aligned
 * in 64 bits, the rest zero), ascending; suffixes with equal keys come in no particular order.
 * use_lsd != 0 forces the LSD passes of the 64-bit-index path.  Returns the key width in bits, or a
 * negative error. */
int caps_sa_gpu_stage_key_sort_u32(caps_sa_gpu_engine* engine, const char* text, uint64_t n, int use_lsd,
                                   uint64_t* keys_out, uint32_t* sa_out);

/* Inclusive max-scan and exclusive sum-scan of a uint32 array (the two scan flavours the
 * pipeline uses), in place. */
int caps_sa_gpu_stage_scan_u32(caps_sa_gpu_engine* engine, uint32_t* data, uint64_t n, int inclusive_max);

#ifdef __cplusplus
}
#endif
#endif /* CAPS_SA_GPU_H */
