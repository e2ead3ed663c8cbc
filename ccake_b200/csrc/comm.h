// Point-to-point transport between the contexts of a slab decomposition (one context per GPU).
//
//   NcclLink  : one process per GPU, NCCL send/recv over NVLink (libnccl.so.2 is dlopen'ed on
//                    first use, so a single-GPU run has no NCCL dependency at all);
//   LocalLink : several contexts inside ONE process (each driven by its own host thread),
//                    exchanging through device mailboxes. It exists so that the whole slab logic
//                    (ghost columns, migration, reductions) is exercised by the single-GPU test run.
//
// Both expose the same small interface; inside a group every rank issues its sends before its
// receives, and both sides of a pair issue their messages in the same order (FIFO matching).
#pragma once

#include <cuda_runtime.h>
#include <dlfcn.h>

#include <condition_variable>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

namespace ccb {

struct Link {
  virtual ~Link() {}
  virtual const char* name() const = 0;
  virtual int group_begin() = 0;
  virtual int group_end() = 0;
  virtual int send(const void* dbuf, size_t bytes, int peer, cudaStream_t s) = 0;
  virtual int recv(void* dbuf, size_t bytes, int peer, cudaStream_t s) = 0;
  virtual int allreduce_sum(double* dbuf, int n, cudaStream_t s) = 0;  // in place, device buffer
  // Peer-memory halo: hand the neighbours (rank -+ 1; -1 = none) a view of this rank's record array and flag words and
  // obtain theirs, as device pointers valid on THIS device (CUDA IPC between processes, plain pointers inside one).
  virtual int exchange_pointers(void* rec, void* flags, int left, int right, void* peer_rec[2], void* peer_flags[2], cudaStream_t s) = 0;
  std::string err;
};

// ---- NCCL, resolved at run time -------------------------------------------------------------
struct NcclApi {
  typedef struct ncclComm* comm_t;
  struct unique_id { char internal[128]; };
  void* h = nullptr;
  int (*GetUniqueId)(unique_id*) = nullptr;
  int (*CommInitRank)(comm_t*, int, unique_id, int) = nullptr;
  int (*CommDestroy)(comm_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, comm_t, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, comm_t, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, comm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool load(std::string& err) {
    if (h) return true;
    h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { err = std::string("cannot load libnccl.so.2: ") + dlerror(); return false; }
#define CCB_SYM(field, sym) field = reinterpret_cast<decltype(field)>(dlsym(h, sym)); if (!field) { err = std::string("libnccl: missing ") + sym; return false; }
    CCB_SYM(GetUniqueId, "ncclGetUniqueId") CCB_SYM(CommInitRank, "ncclCommInitRank") CCB_SYM(CommDestroy, "ncclCommDestroy")
    CCB_SYM(Send, "ncclSend") CCB_SYM(Recv, "ncclRecv") CCB_SYM(AllReduce, "ncclAllReduce") CCB_SYM(GroupStart, "ncclGroupStart")
    CCB_SYM(GroupEnd, "ncclGroupEnd") CCB_SYM(GetErrorString, "ncclGetErrorString")
#undef CCB_SYM
    return true;
  }
  static NcclApi& get() { static NcclApi api; return api; }
};

struct NcclLink : Link {
  NcclApi::comm_t comm = nullptr;
  NcclApi& api = NcclApi::get();
  const char* name() const override { return "nccl"; }
  int init(const char id[128], int rank, int nranks) {
    if (!api.load(err)) return 1;
    NcclApi::unique_id uid;
    std::memcpy(uid.internal, id, 128);
    return chk(api.CommInitRank(&comm, nranks, uid, rank));
  }
  ~NcclLink() override { if (comm) api.CommDestroy(comm); }
  int chk(int rc) { if (rc) { err = std::string("NCCL: ") + api.GetErrorString(rc); return 1; } return 0; }
  int group_begin() override { return chk(api.GroupStart()); }
  int group_end() override { return chk(api.GroupEnd()); }
  int send(const void* d, size_t bytes, int peer, cudaStream_t s) override { return chk(api.Send(d, bytes, /*ncclChar*/ 0, peer, comm, s)); }
  int recv(void* d, size_t bytes, int peer, cudaStream_t s) override { return chk(api.Recv(d, bytes, 0, peer, comm, s)); }
  int allreduce_sum(double* d, int n, cudaStream_t s) override { return chk(api.AllReduce(d, d, (size_t)n, /*ncclDouble*/ 8, /*ncclSum*/ 0, comm, s)); }
  std::vector<void*> opened;
  int exchange_pointers(void* rec, void* flags, int left, int right, void* peer_rec[2], void* peer_flags[2], cudaStream_t s) override {
    for (void* p : opened) cudaIpcCloseMemHandle(p);
    opened.clear();
    // A local failure must not leave a neighbour waiting in the handle exchange: the exchange always runs (zeroed handles on
    // failure) and the caller agrees on the outcome across ranks before using the pointers.
    int failed = 0;
    cudaIpcMemHandle_t mine[2], theirs[2][2];
    std::memset(mine, 0, sizeof(mine));
    if (cudaIpcGetMemHandle(&mine[0], rec) != cudaSuccess || cudaIpcGetMemHandle(&mine[1], flags) != cudaSuccess) {
      err = std::string("cudaIpcGetMemHandle failed: ") + cudaGetErrorString(cudaGetLastError());
      std::memset(mine, 0, sizeof(mine));
      failed = 1;
    }
    char* d = nullptr;
    const size_t hb = sizeof(mine);
    if (cudaMalloc(&d, 3 * hb) != cudaSuccess) { err = "cudaMalloc failed"; return 1; }
    cudaMemcpyAsync(d, mine, hb, cudaMemcpyHostToDevice, s);
    const int peer[2] = {left, right};
    if (group_begin()) return 1;
    for (int k = 0; k < 2; ++k) if (peer[k] >= 0 && send(d, hb, peer[k], s)) return 1;
    for (int k = 0; k < 2; ++k) if (peer[k] >= 0 && recv(d + (1 + k) * hb, hb, peer[k], s)) return 1;
    if (group_end()) return 1;
    cudaMemcpyAsync(theirs, d + hb, 2 * hb, cudaMemcpyDeviceToHost, s);
    if (cudaStreamSynchronize(s) != cudaSuccess) { err = "handle exchange failed"; return 1; }
    cudaFree(d);
    for (int k = 0; k < 2; ++k) {
      peer_rec[k] = peer_flags[k] = nullptr;
      if (peer[k] < 0) continue;
      if (failed) continue;
      if (cudaIpcOpenMemHandle(&peer_rec[k], theirs[k][0], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
          cudaIpcOpenMemHandle(&peer_flags[k], theirs[k][1], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        err = std::string("cudaIpcOpenMemHandle failed: ") + cudaGetErrorString(cudaGetLastError());
        failed = 1;
        continue;
      }
      opened.push_back(peer_rec[k]); opened.push_back(peer_flags[k]);
    }
    return failed;
  }
};

// ---- in-process transport ---------------------------------------------------------------------
struct LocalHub {
  struct Msg { void* dptr; size_t bytes; };
  std::mutex mu;
  std::condition_variable cv;
  std::map<std::pair<int, int>, std::deque<Msg>> q;  // (src, dst) -> FIFO
  std::map<int, std::pair<void*, void*>> ptrs;       // rank -> (record array, flag words): exchange_pointers
  std::map<int, int> ptr_gen;                        // rank -> number of times it has published: a second exchange (re-upload) must not read the first one's pointers
  int nranks = 0;
  static std::shared_ptr<LocalHub> get(int group, int nranks) {
    static std::mutex gmu;
    static std::map<int, std::weak_ptr<LocalHub>> hubs;
    std::lock_guard<std::mutex> l(gmu);
    auto sp = hubs[group].lock();
    if (!sp) { sp = std::make_shared<LocalHub>(); sp->nranks = nranks; hubs[group] = sp; }
    return sp;
  }
};

struct LocalLink : Link {
  std::shared_ptr<LocalHub> hub;
  int rank = 0, nranks = 1;
  const char* name() const override { return "local"; }
  LocalLink(int group, int rank_, int nranks_) : hub(LocalHub::get(group, nranks_)), rank(rank_), nranks(nranks_) {}
  int group_begin() override { return 0; }
  int group_end() override { return 0; }
  int send(const void* d, size_t bytes, int peer, cudaStream_t s) override {
    void* box = nullptr;
    if (bytes) {
      if (cudaMalloc(&box, bytes) != cudaSuccess) { err = "local transport: cudaMalloc failed"; return 1; }
      if (cudaMemcpyAsync(box, d, bytes, cudaMemcpyDeviceToDevice, s) != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess) {
        err = "local transport: copy to mailbox failed"; return 1;
      }
    }
    { std::lock_guard<std::mutex> l(hub->mu); hub->q[{rank, peer}].push_back({box, bytes}); }
    hub->cv.notify_all();
    return 0;
  }
  int recv(void* d, size_t bytes, int peer, cudaStream_t s) override {
    LocalHub::Msg m;
    {
      std::unique_lock<std::mutex> l(hub->mu);
      auto& dq = hub->q[{peer, rank}];
      hub->cv.wait(l, [&] { return !dq.empty(); });
      m = dq.front();
      dq.pop_front();
    }
    if (m.bytes != bytes) { err = "local transport: message size mismatch"; if (m.dptr) cudaFree(m.dptr); return 1; }
    if (bytes) {
      if (cudaMemcpyAsync(d, m.dptr, bytes, cudaMemcpyDeviceToDevice, s) != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess) {
        err = "local transport: copy from mailbox failed"; return 1;
      }
      cudaFree(m.dptr);
    }
    return 0;
  }
  int exchange_pointers(void* rec, void* flags, int left, int right, void* peer_rec[2], void* peer_flags[2], cudaStream_t) override {
    std::unique_lock<std::mutex> l(hub->mu);
    hub->ptrs[rank] = {rec, flags};
    const int gen = ++hub->ptr_gen[rank];
    hub->cv.notify_all();
    const int peer[2] = {left, right};
    for (int k = 0; k < 2; ++k) {
      peer_rec[k] = peer_flags[k] = nullptr;
      if (peer[k] < 0) continue;
      hub->cv.wait(l, [&] { return hub->ptr_gen.count(peer[k]) != 0 && hub->ptr_gen[peer[k]] >= gen; });
      peer_rec[k] = hub->ptrs[peer[k]].first; peer_flags[k] = hub->ptrs[peer[k]].second;
    }
    return 0;
  }
  int allreduce_sum(double* d, int n, cudaStream_t s) override {
    std::vector<double> mine(n), acc(n, 0.0), other(n);
    if (cudaMemcpyAsync(mine.data(), d, sizeof(double) * n, cudaMemcpyDeviceToHost, s) != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess) return 1;
    for (int p = 0; p < nranks; ++p) if (p != rank && send(d, sizeof(double) * n, p, s)) return 1;
    double* tmp = nullptr;
    if (cudaMalloc(&tmp, sizeof(double) * n) != cudaSuccess) return 1;
    for (int p = 0; p < nranks; ++p) {  // summed in rank order on every rank: identical totals everywhere
      if (p == rank) { for (int k = 0; k < n; ++k) acc[k] += mine[k]; continue; }
      if (recv(tmp, sizeof(double) * n, p, s)) { cudaFree(tmp); return 1; }
      cudaMemcpy(other.data(), tmp, sizeof(double) * n, cudaMemcpyDeviceToHost);
      for (int k = 0; k < n; ++k) acc[k] += other[k];
    }
    cudaFree(tmp);
    cudaMemcpyAsync(d, acc.data(), sizeof(double) * n, cudaMemcpyHostToDevice, s);
    return cudaStreamSynchronize(s) == cudaSuccess ? 0 : 1;
  }
};

}  // namespace ccb
