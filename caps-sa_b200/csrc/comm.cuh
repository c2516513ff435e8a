// Exchange steps of the sharded (multi-GPU) construction.
//
// The sharded pipeline (sharded_build.cu) needs three collectives: a variable all-to-all of
// device buffers (suffix indices / keys moving to the rank that owns their bucket or their
// text position — the reference's collate step, src/Suffix_Array.cpp:335-364), an all-gather
// of a few host words (send-count matrix ≙ the reference's P, :481; convergence flags) and an
// all-gather of small device buffers (pivot samples, :207-216).  Two transports implement
// them:
//   NcclComm   — one process per GPU (torchrun); NCCL over NVLink 5 / NVSwitch.  NCCL is
//                resolved with dlopen at first use so the single-GPU path has no link-time
//                dependency on it.
//   ThreadComm — one host thread per rank inside one process, peer copies between the
//                ranks' buffers (cudaMemcpyAsync; NVLink P2P when the devices differ).  Drives
//                the C++ class / CLI across several GPUs and lets several virtual ranks share
//                one device for the parity tests.
#pragma once

#include <condition_variable>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"

namespace capsb {

struct Comm {
  int rank = 0;
  int world = 1;
  virtual ~Comm() = default;

  // Rank r sends send_counts[p] elements (elem_bytes each, consecutive in `send` in rank
  // order) to every rank p and receives recv_counts[p] elements from it, consecutive in
  // `recv` in rank order.  Stream-ordered for the caller: later work on `st` sees the data.
  virtual void all_to_all_v(const void* send, const uint64_t* send_counts, void* recv, const uint64_t* recv_counts,
                            size_t elem_bytes, cudaStream_t st) = 0;

  // Every rank contributes `bytes` of host memory; `out` receives world * bytes in rank
  // order.  Blocking (synchronises `st`).
  virtual void all_gather_host(const void* in, size_t bytes, void* out, cudaStream_t st) = 0;

  // Every rank contributes `bytes` of device memory; `recv` (device) receives world * bytes.
  virtual void all_gather_device(const void* send, void* recv, size_t bytes, cudaStream_t st) = 0;

  // Peer-visible buffers, for kernels that store straight into the other ranks' memory over
  // NVLink (the fused partition pass, sharded_build.cu).  Collective: every rank passes a buffer
  // it obtained from cudaMalloc itself (the base address of the allocation — what CUDA IPC can
  // export) and gets, for every rank p, an address through which its kernels can write rank p's
  // buffer (entry `rank` is `mine`).  close_peer_buffers is collective too: it synchronises `st`,
  // waits until every rank has done so (all stores into every buffer have landed) and releases
  // the mappings.
  virtual std::vector<void*> open_peer_buffers(void* mine, cudaStream_t st) = 0;
  virtual void close_peer_buffers(const std::vector<void*>& peers, cudaStream_t st) = 0;

  // bytes moved by this rank over the transport so far (for the NVLink roofline)
  uint64_t bytes_sent = 0;
};

// ---- in-process transport ---------------------------------------------------------------
// Shared state of the ranks of one ThreadComm group.
class ThreadGroup {
 public:
  explicit ThreadGroup(int world) : world_(world), slots_(world) {}
  int world() const { return world_; }

  struct Slot {
    const void* ptr = nullptr;
    const uint64_t* counts = nullptr;
    int device = 0;
  };
  Slot& slot(int rank) { return slots_[rank]; }

  // Reusable barrier; throws when any rank has failed so that nobody waits forever.
  void arrive_and_wait();
  void fail(const std::string& why);
  bool failed() const { return failed_; }
  std::string failure() const { return failure_; }

 private:
  int world_;
  std::vector<Slot> slots_;
  std::mutex mu_;
  std::condition_variable cv_;
  int waiting_ = 0;
  uint64_t generation_ = 0;
  bool failed_ = false;
  std::string failure_;
};

class ThreadComm : public Comm {
 public:
  ThreadComm(std::shared_ptr<ThreadGroup> group, int rank_, int device);
  void all_to_all_v(const void* send, const uint64_t* send_counts, void* recv, const uint64_t* recv_counts,
                    size_t elem_bytes, cudaStream_t st) override;
  void all_gather_host(const void* in, size_t bytes, void* out, cudaStream_t st) override;
  void all_gather_device(const void* send, void* recv, size_t bytes, cudaStream_t st) override;
  std::vector<void*> open_peer_buffers(void* mine, cudaStream_t st) override;
  void close_peer_buffers(const std::vector<void*>& peers, cudaStream_t st) override;

 private:
  std::shared_ptr<ThreadGroup> group_;
  int device_;
};

// ---- NCCL transport ---------------------------------------------------------------------
constexpr size_t kCommIdBytes = 128;  // NCCL_UNIQUE_ID_BYTES

// Fills a fresh NCCL unique id (rank 0 calls this; the host side broadcasts the bytes).
void nccl_unique_id(void* out128);

class NcclComm : public Comm {
 public:
  // Collective over all ranks: joins the communicator described by `id128`.
  NcclComm(const void* id128, int rank_, int world_, int device);
  ~NcclComm() override;
  void all_to_all_v(const void* send, const uint64_t* send_counts, void* recv, const uint64_t* recv_counts,
                    size_t elem_bytes, cudaStream_t st) override;
  void all_gather_host(const void* in, size_t bytes, void* out, cudaStream_t st) override;
  void all_gather_device(const void* send, void* recv, size_t bytes, cudaStream_t st) override;
  std::vector<void*> open_peer_buffers(void* mine, cudaStream_t st) override;
  void close_peer_buffers(const std::vector<void*>& peers, cudaStream_t st) override;

 private:
  void* comm_ = nullptr;  // ncclComm_t
  int device_;
};

// Trivial transport for world == 1 (no peers): copies.
class SelfComm : public Comm {
 public:
  void all_to_all_v(const void* send, const uint64_t* send_counts, void* recv, const uint64_t* recv_counts,
                    size_t elem_bytes, cudaStream_t st) override;
  void all_gather_host(const void* in, size_t bytes, void* out, cudaStream_t st) override;
  void all_gather_device(const void* send, void* recv, size_t bytes, cudaStream_t st) override;
  std::vector<void*> open_peer_buffers(void* mine, cudaStream_t st) override;
  void close_peer_buffers(const std::vector<void*>& peers, cudaStream_t st) override;
};

}  // namespace capsb
