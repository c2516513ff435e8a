// Transports of the sharded construction (see comm.cuh).
#include "comm.cuh"

#include <dlfcn.h>

#include <cstring>

namespace capsb {

// ---- ThreadGroup ------------------------------------------------------------------------
void ThreadGroup::arrive_and_wait() {
  std::unique_lock<std::mutex> lock(mu_);
  if (failed_) throw Error("a peer rank failed: " + failure_);
  const uint64_t gen = generation_;
  if (++waiting_ == world_) {
    waiting_ = 0;
    ++generation_;
    cv_.notify_all();
    return;
  }
  cv_.wait(lock, [&] { return generation_ != gen || failed_; });
  if (generation_ == gen) throw Error("a peer rank failed: " + failure_);
}

void ThreadGroup::fail(const std::string& why) {
  std::lock_guard<std::mutex> lock(mu_);
  if (!failed_) {
    failed_ = true;
    failure_ = why;
  }
  cv_.notify_all();
}

// ---- ThreadComm -------------------------------------------------------------------------
ThreadComm::ThreadComm(std::shared_ptr<ThreadGroup> group, int rank_, int device)
    : group_(std::move(group)), device_(device) {
  rank = rank_;
  world = group_->world();
  group_->slot(rank).device = device;
  group_->arrive_and_wait();
  // direct peer copies over NVLink where the ranks sit on different devices
  for (int p = 0; p < world; ++p) {
    const int other = group_->slot(p).device;
    if (other == device_) continue;
    int can = 0;
    if (cudaDeviceCanAccessPeer(&can, device_, other) == cudaSuccess && can) {
      const cudaError_t err = cudaDeviceEnablePeerAccess(other, 0);
      if (err != cudaSuccess) cudaGetLastError();  // already enabled is fine
    } else {
      cudaGetLastError();
    }
  }
}

void ThreadComm::all_to_all_v(const void* send, const uint64_t* send_counts, void* recv,
                              const uint64_t* recv_counts, size_t elem_bytes, cudaStream_t st) {
  CAPSB_CUDA(cudaStreamSynchronize(st));  // my send buffer is complete
  ThreadGroup::Slot& mine = group_->slot(rank);
  mine.ptr = send;
  mine.counts = send_counts;
  group_->arrive_and_wait();
  uint64_t recv_off = 0;
  for (int s = 0; s < world; ++s) {
    const ThreadGroup::Slot& src = group_->slot(s);
    uint64_t src_off = 0;
    for (int p = 0; p < rank; ++p) src_off += src.counts[p];
    const uint64_t count = src.counts[rank];
    if (count != recv_counts[s]) {
      group_->fail("all_to_all_v: count mismatch between ranks");
      fail("all_to_all_v: count mismatch between ranks");
    }
    if (count) {
      const char* from = static_cast<const char*>(src.ptr) + src_off * elem_bytes;
      char* to = static_cast<char*>(recv) + recv_off * elem_bytes;
      if (src.device == device_)
        CAPSB_CUDA(cudaMemcpyAsync(to, from, count * elem_bytes, cudaMemcpyDeviceToDevice, st));
      else
        CAPSB_CUDA(cudaMemcpyPeerAsync(to, device_, from, src.device, count * elem_bytes, st));
      if (s != rank) bytes_sent += count * elem_bytes;  // symmetric on average; counted on the receiving side
    }
    recv_off += count;
  }
  CAPSB_CUDA(cudaStreamSynchronize(st));
  group_->arrive_and_wait();  // peers may now reuse their send buffers
}

void ThreadComm::all_gather_host(const void* in, size_t bytes, void* out, cudaStream_t) {
  group_->slot(rank).ptr = in;
  group_->arrive_and_wait();
  for (int s = 0; s < world; ++s)
    std::memcpy(static_cast<char*>(out) + static_cast<size_t>(s) * bytes, group_->slot(s).ptr, bytes);
  group_->arrive_and_wait();
}

void ThreadComm::all_gather_device(const void* send, void* recv, size_t bytes, cudaStream_t st) {
  CAPSB_CUDA(cudaStreamSynchronize(st));
  group_->slot(rank).ptr = send;
  group_->arrive_and_wait();
  for (int s = 0; s < world; ++s) {
    const ThreadGroup::Slot& src = group_->slot(s);
    char* to = static_cast<char*>(recv) + static_cast<size_t>(s) * bytes;
    if (!bytes || to == src.ptr) continue;  // in place: the own piece is already there
    if (src.device == device_)
      CAPSB_CUDA(cudaMemcpyAsync(to, src.ptr, bytes, cudaMemcpyDeviceToDevice, st));
    else
      CAPSB_CUDA(cudaMemcpyPeerAsync(to, device_, src.ptr, src.device, bytes, st));
  }
  CAPSB_CUDA(cudaStreamSynchronize(st));
  group_->arrive_and_wait();
}

// ---- SelfComm ---------------------------------------------------------------------------
std::vector<void*> ThreadComm::open_peer_buffers(void* mine, cudaStream_t) {
  // one process: the peers' addresses are usable as they are (peer access was enabled in the
  // constructor where the ranks sit on different devices)
  group_->slot(rank).ptr = mine;
  group_->arrive_and_wait();
  std::vector<void*> peers(static_cast<size_t>(world));
  for (int p = 0; p < world; ++p) peers[static_cast<size_t>(p)] = const_cast<void*>(group_->slot(p).ptr);
  group_->arrive_and_wait();  // everybody has read the slots before they are reused
  return peers;
}

void ThreadComm::close_peer_buffers(const std::vector<void*>&, cudaStream_t st) {
  CAPSB_CUDA(cudaStreamSynchronize(st));
  group_->arrive_and_wait();
}

void SelfComm::all_to_all_v(const void* send, const uint64_t* send_counts, void* recv, const uint64_t* recv_counts,
                            size_t elem_bytes, cudaStream_t st) {
  if (send_counts[0] != recv_counts[0]) fail("all_to_all_v: count mismatch");
  if (send_counts[0])
    CAPSB_CUDA(cudaMemcpyAsync(recv, send, send_counts[0] * elem_bytes, cudaMemcpyDeviceToDevice, st));
}
void SelfComm::all_gather_host(const void* in, size_t bytes, void* out, cudaStream_t) { std::memcpy(out, in, bytes); }
std::vector<void*> SelfComm::open_peer_buffers(void* mine, cudaStream_t) { return {mine}; }
void SelfComm::close_peer_buffers(const std::vector<void*>&, cudaStream_t st) { CAPSB_CUDA(cudaStreamSynchronize(st)); }
void SelfComm::all_gather_device(const void* send, void* recv, size_t bytes, cudaStream_t st) {
  if (bytes && send != recv) CAPSB_CUDA(cudaMemcpyAsync(recv, send, bytes, cudaMemcpyDeviceToDevice, st));
}

// ---- NCCL, resolved at run time ---------------------------------------------------------
namespace {

struct IdBlob {
  char internal[kCommIdBytes];
};
constexpr int kNcclUint8 = 1;  // ncclUint8 in nccl.h's ncclDataType_t

struct NcclApi {
  int (*GetUniqueId)(IdBlob*) = nullptr;
  int (*CommInitRank)(void**, int, IdBlob, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};

const NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  static std::string problem;
  std::call_once(once, [] {
    // prefer the copy the host process already loaded (torch bundles its own libnccl.so.2)
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
      problem = std::string("cannot load NCCL: ") + dlerror();
      return;
    }
    auto sym = [&](const char* name) -> void* {
      void* p = dlsym(h, name);
      if (!p && problem.empty()) problem = std::string("NCCL lacks symbol ") + name;
      return p;
    };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
    api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
    api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
  });
  if (!problem.empty()) fail(problem);
  return api;
}

void nccl_check(int rc, const char* what) {
  if (rc != 0) fail(std::string("NCCL error in ") + what + ": " + nccl().GetErrorString(rc));
}

}  // namespace

void nccl_unique_id(void* out128) {
  IdBlob id;
  nccl_check(nccl().GetUniqueId(&id), "ncclGetUniqueId");
  std::memcpy(out128, id.internal, kCommIdBytes);
}

NcclComm::NcclComm(const void* id128, int rank_, int world_, int device) : device_(device) {
  rank = rank_;
  world = world_;
  IdBlob id;
  std::memcpy(id.internal, id128, kCommIdBytes);
  CAPSB_CUDA(cudaSetDevice(device));
  nccl_check(nccl().CommInitRank(&comm_, world, id, rank), "ncclCommInitRank");
}

NcclComm::~NcclComm() {
  if (comm_) nccl().CommDestroy(comm_);
}

void NcclComm::all_to_all_v(const void* send, const uint64_t* send_counts, void* recv, const uint64_t* recv_counts,
                            size_t elem_bytes, cudaStream_t st) {
  const NcclApi& api = nccl();
  // the piece that stays on this rank is a plain device copy (through NCCL it would be a
  // send/recv pair sharing the channels with the NVLink traffic)
  {
    uint64_t soff = 0, roff = 0;
    for (int p = 0; p < rank; ++p) soff += send_counts[p], roff += recv_counts[p];
    if (send_counts[rank] != recv_counts[rank]) fail("all_to_all_v: count mismatch");
    if (send_counts[rank])
      CAPSB_CUDA(cudaMemcpyAsync(static_cast<char*>(recv) + roff * elem_bytes,
                                 static_cast<const char*>(send) + soff * elem_bytes, send_counts[rank] * elem_bytes,
                                 cudaMemcpyDeviceToDevice, st));
  }
  nccl_check(api.GroupStart(), "ncclGroupStart");
  uint64_t soff = 0, roff = 0;
  for (int p = 0; p < world; ++p) {
    if (p != rank) {
      if (send_counts[p]) {
        nccl_check(api.Send(static_cast<const char*>(send) + soff * elem_bytes, send_counts[p] * elem_bytes,
                            kNcclUint8, p, comm_, st),
                   "ncclSend");
        bytes_sent += send_counts[p] * elem_bytes;
      }
      if (recv_counts[p])
        nccl_check(api.Recv(static_cast<char*>(recv) + roff * elem_bytes, recv_counts[p] * elem_bytes, kNcclUint8, p,
                            comm_, st),
                   "ncclRecv");
    }
    soff += send_counts[p];
    roff += recv_counts[p];
  }
  nccl_check(api.GroupEnd(), "ncclGroupEnd");
}

void NcclComm::all_gather_device(const void* send, void* recv, size_t bytes, cudaStream_t st) {
  if (!bytes) return;
  nccl_check(nccl().AllGather(send, recv, bytes, kNcclUint8, comm_, st), "ncclAllGather");
  bytes_sent += bytes * static_cast<size_t>(world - 1);
}

// One process per rank: the buffers cross the process boundary as CUDA IPC handles (all-gathered
// as host bytes) and are mapped with peer access over NVLink.
std::vector<void*> NcclComm::open_peer_buffers(void* mine, cudaStream_t st) {
  cudaIpcMemHandle_t handle;
  CAPSB_CUDA(cudaIpcGetMemHandle(&handle, mine));
  std::vector<cudaIpcMemHandle_t> all(static_cast<size_t>(world));
  all_gather_host(&handle, sizeof(handle), all.data(), st);
  std::vector<void*> peers(static_cast<size_t>(world), nullptr);
  for (int p = 0; p < world; ++p) {
    if (p == rank) {
      peers[static_cast<size_t>(p)] = mine;
      continue;
    }
    void* mapped = nullptr;
    CAPSB_CUDA(cudaIpcOpenMemHandle(&mapped, all[static_cast<size_t>(p)], cudaIpcMemLazyEnablePeerAccess));
    peers[static_cast<size_t>(p)] = mapped;
  }
  return peers;
}

void NcclComm::close_peer_buffers(const std::vector<void*>& peers, cudaStream_t st) {
  CAPSB_CUDA(cudaStreamSynchronize(st));
  // barrier: nobody unmaps (or reads its own buffer) before every rank's stores have completed
  const uint32_t token = 1;
  std::vector<uint32_t> tokens(static_cast<size_t>(world));
  all_gather_host(&token, sizeof(token), tokens.data(), st);
  for (int p = 0; p < world; ++p)
    if (p != rank && peers[static_cast<size_t>(p)]) CAPSB_CUDA(cudaIpcCloseMemHandle(peers[static_cast<size_t>(p)]));
  // second barrier: an owner may free its buffer only after every importer has unmapped it
  all_gather_host(&token, sizeof(token), tokens.data(), st);
}

void NcclComm::all_gather_host(const void* in, size_t bytes, void* out, cudaStream_t st) {
  if (!bytes) return;
  DevBuf<unsigned char> d_in(bytes, st), d_out(bytes * static_cast<size_t>(world), st);
  CAPSB_CUDA(cudaMemcpyAsync(d_in.get(), in, bytes, cudaMemcpyHostToDevice, st));
  nccl_check(nccl().AllGather(d_in.get(), d_out.get(), bytes, kNcclUint8, comm_, st), "ncclAllGather");
  read_back(st, out, d_out.get(), bytes * static_cast<size_t>(world));
}

}  // namespace capsb
