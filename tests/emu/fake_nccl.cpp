// fake_nccl.cpp -> libfake_nccl.so: an in-process stand-in for the NCCL entry points sharded.cu resolves with dlsym
// (TEST INFRASTRUCTURE ONLY, tests/emu).  A "rank" is a host thread; all ranks of a communicator live in one process and
// "device memory" is host memory, so a collective is: publish my pointers, barrier, copy, barrier.
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "fake_cuda/nccl.h"

namespace {
struct Group {
  int n = 0;
  std::mutex m;
  std::condition_variable cv;
  int arrived = 0;
  unsigned gen = 0;
  std::vector<const void*> send;
  void barrier() {
    std::unique_lock<std::mutex> lk(m);
    const unsigned g = gen;
    if (++arrived == n) {
      arrived = 0;
      ++gen;
      cv.notify_all();
    } else {
      cv.wait(lk, [&] { return gen != g; });
    }
  }
};
std::mutex g_mu;
std::map<std::string, std::shared_ptr<Group>> g_groups;
unsigned long long g_next_id = 1;
size_t dtype_size(ncclDataType_t t) {
  switch (t) {
    case ncclInt8: case ncclUint8: return 1;
    case ncclFloat16: return 2;
    case ncclInt32: case ncclUint32: case ncclFloat32: return 4;
    default: return 8;
  }
}
}  // namespace

struct ncclComm {
  std::shared_ptr<Group> g;
  int rank;
};

extern "C" {
ncclResult_t ncclGetUniqueId(ncclUniqueId* id) {
  std::lock_guard<std::mutex> lk(g_mu);
  memset(id, 0, sizeof(*id));
  const unsigned long long v = g_next_id++;
  memcpy(id->internal, &v, sizeof(v));
  memcpy(id->internal + 8, "fake-nccl", 9);
  return ncclSuccess;
}
ncclResult_t ncclCommInitRank(ncclComm_t* comm, int nranks, ncclUniqueId id, int rank) {
  if (nranks < 1 || rank < 0 || rank >= nranks) return ncclInvalidArgument;
  std::shared_ptr<Group> g;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto& slot = g_groups[std::string(id.internal, sizeof(id.internal))];
    if (!slot) {
      slot = std::make_shared<Group>();
      slot->n = nranks;
      slot->send.assign(nranks, nullptr);
    }
    if (slot->n != nranks) return ncclInvalidArgument;
    g = slot;
  }
  *comm = new ncclComm{g, rank};
  g->barrier();  // like the real call: returns once every rank has joined
  return ncclSuccess;
}
ncclResult_t ncclCommDestroy(ncclComm_t c) {
  delete c;
  return ncclSuccess;
}
ncclResult_t ncclAllGather(const void* send, void* recv, size_t count, ncclDataType_t t, ncclComm_t c, cudaStream_t) {
  Group& g = *c->g;
  const size_t bytes = count * dtype_size(t);
  g.send[c->rank] = send;
  g.barrier();
  for (int r = 0; r < g.n; ++r) memmove(static_cast<char*>(recv) + (size_t)r * bytes, g.send[r], bytes);  // in place allowed
  g.barrier();
  return ncclSuccess;
}
ncclResult_t ncclBroadcast(const void* send, void* recv, size_t count, ncclDataType_t t, int root, ncclComm_t c, cudaStream_t) {
  Group& g = *c->g;
  if (c->rank == root) g.send[root] = send;
  g.barrier();
  if (g.send[root] != recv) memmove(recv, g.send[root], count * dtype_size(t));
  g.barrier();
  return ncclSuccess;
}
ncclResult_t ncclGroupStart() { return ncclSuccess; }
ncclResult_t ncclGroupEnd() { return ncclSuccess; }
const char* ncclGetErrorString(ncclResult_t r) { return r == ncclSuccess ? "no error" : "fake NCCL error"; }
ncclResult_t ncclGetVersion(int* v) {
  *v = 22809;
  return ncclSuccess;
}
}
