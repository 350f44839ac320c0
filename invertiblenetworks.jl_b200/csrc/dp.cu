// dp.cu - NCCL plumbing of the data-parallel plane (dp.cuh).  The library does not link NCCL: the handful of entry
// points it needs are resolved with dlopen/dlsym on first use, preferring a libnccl.so.2 that is already mapped into
// the process (torch's, NCCL.jl's artifact) so that a communicator created by the host framework and the calls made
// here go through the same library instance.
#include "dp.cuh"

#include <dlfcn.h>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace inb {

namespace {
// the slice of nccl.h (2.10+) used here, restated so that the build needs no NCCL headers
typedef void* ncclComm_t;
struct ncclUniqueId { char internal[128]; };
enum { ncclSuccess = 0 };
enum { ncclSum = 0, ncclAvg = 4 };
enum { ncclFloat32 = 7, ncclFloat64 = 8 };

struct Nccl {
  void* h = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*CommCount)(ncclComm_t, int*) = nullptr;
  int (*CommUserRank)(ncclComm_t, int*) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  std::string err;
};

Nccl& nccl() {
  static Nccl n;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* env = getenv("INB200_NCCL_LIB");
    const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
      if (!nm || !nm[0]) continue;
      n.h = dlopen(nm, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);  // the instance the host framework already loaded
      if (n.h) break;
    }
    for (const char* nm : names) {
      if (n.h) break;
      if (!nm || !nm[0]) continue;
      n.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    }
    if (!n.h) {
      n.err = "NCCL is not available: libnccl.so.2 could not be loaded (set INB200_NCCL_LIB to its path)";
      return;
    }
    auto sym = [&](const char* s) -> void* {
      void* p = dlsym(n.h, s);
      if (!p && n.err.empty()) n.err = std::string("NCCL symbol missing: ") + s;
      return p;
    };
    n.GetUniqueId = (decltype(n.GetUniqueId))sym("ncclGetUniqueId");
    n.CommInitRank = (decltype(n.CommInitRank))sym("ncclCommInitRank");
    n.CommDestroy = (decltype(n.CommDestroy))sym("ncclCommDestroy");
    n.CommCount = (decltype(n.CommCount))sym("ncclCommCount");
    n.CommUserRank = (decltype(n.CommUserRank))sym("ncclCommUserRank");
    n.AllReduce = (decltype(n.AllReduce))sym("ncclAllReduce");
    n.Broadcast = (decltype(n.Broadcast))sym("ncclBroadcast");
    n.GroupStart = (decltype(n.GroupStart))sym("ncclGroupStart");
    n.GroupEnd = (decltype(n.GroupEnd))sym("ncclGroupEnd");
    n.GetErrorString = (decltype(n.GetErrorString))sym("ncclGetErrorString");
  });
  if (!n.err.empty()) fail(1, "%s", n.err.c_str());
  return n;
}

void nccl_check(int rc, const char* what) {
  if (rc != ncclSuccess) fail(2, "%s failed: %s", what, nccl().GetErrorString ? nccl().GetErrorString(rc) : "NCCL error");
}
}  // namespace

cudaEvent_t DpComm::event() {
  if (next_ev == (int)ev.size()) {
    cudaEvent_t e;
    INB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    ev.push_back(e);
  }
  return ev[next_ev++];
}

void dp_allreduce_sum_f64(DpComm* dp, double* buf, size_t count, cudaStream_t st) {
  if (!dp || dp->nranks <= 1 || count == 0) return;
  nccl_check(nccl().AllReduce(buf, buf, count, ncclFloat64, ncclSum, dp->comm, st), "ncclAllReduce(sum, f64)");
  ++dp->calls;
  dp->bytes += (long long)count * 8;
}

void dp_allreduce_avg_bucket(DpComm* dp, float* const* ptrs, const long long* numel, int first, int n, long long max_gap,
                             cudaStream_t st) {
  if (!dp || dp->nranks <= 1 || n <= 0) return;
  Nccl& N = nccl();
  nccl_check(N.GroupStart(), "ncclGroupStart");
  int i = first;
  const int end = first + n;
  while (i < end) {
    float* base = ptrs[i];
    long long len = numel[i];
    int j = i + 1;
    while (j < end) {
      const long long gap = ptrs[j] - (base + len);
      if (gap < 0 || gap > max_gap) break;
      len += gap + numel[j];
      ++j;
    }
    const int rc = N.AllReduce(base, base, (size_t)len, ncclFloat32, ncclAvg, dp->comm, st);
    if (rc != ncclSuccess) {
      N.GroupEnd();
      nccl_check(rc, "ncclAllReduce(avg, f32)");
    }
    ++dp->calls;
    dp->bytes += len * 4;
    i = j;
  }
  nccl_check(N.GroupEnd(), "ncclGroupEnd");
}

void dp_fork(DpComm* dp, cudaStream_t main, cudaStream_t side) {
  cudaEvent_t e = dp->event();
  INB_CUDA(cudaEventRecord(e, main));
  INB_CUDA(cudaStreamWaitEvent(dp->st, e, 0));
  if (side) {
    cudaEvent_t s = dp->event();
    INB_CUDA(cudaEventRecord(s, side));
    INB_CUDA(cudaStreamWaitEvent(dp->st, s, 0));
  }
  dp->pending = true;
}
void dp_join(DpComm* dp, cudaStream_t main) {
  if (dp->pending) {
    cudaEvent_t e = dp->event();
    INB_CUDA(cudaEventRecord(e, dp->st));
    INB_CUDA(cudaStreamWaitEvent(main, e, 0));
  }
  dp->pending = false;
  dp->next_ev = 0;
}

// ---------------------------------------------------------------- communicator life cycle (used by api.cu)
void dp_unique_id(char* id128) {
  ncclUniqueId id;
  nccl_check(nccl().GetUniqueId(&id), "ncclGetUniqueId");
  memcpy(id128, id.internal, 128);
}
DpComm* dp_create(int nranks, int rank, const char* id128) {
  INB_CHECK(nranks >= 1 && rank >= 0 && rank < nranks, "communicator: rank %d outside [0, %d)", rank, nranks);
  ncclUniqueId id;
  memcpy(id.internal, id128, 128);
  ncclComm_t comm = nullptr;
  nccl_check(nccl().CommInitRank(&comm, nranks, id, rank), "ncclCommInitRank");
  DpComm* dp = new DpComm();
  dp->comm = comm;
  dp->nranks = nranks;
  dp->rank = rank;
  dp->owned = true;
  INB_CUDA(cudaStreamCreateWithFlags(&dp->st, cudaStreamNonBlocking));
  return dp;
}
DpComm* dp_wrap(void* nccl_comm) {
  INB_CHECK(nccl_comm != nullptr, "null ncclComm_t");
  DpComm* dp = new DpComm();
  dp->comm = nccl_comm;
  int n = 1, r = 0;
  nccl_check(nccl().CommCount(nccl_comm, &n), "ncclCommCount");
  nccl_check(nccl().CommUserRank(nccl_comm, &r), "ncclCommUserRank");
  dp->nranks = n;
  dp->rank = r;
  dp->owned = false;
  INB_CUDA(cudaStreamCreateWithFlags(&dp->st, cudaStreamNonBlocking));
  return dp;
}
void dp_destroy(DpComm* dp) {
  if (!dp) return;
  for (cudaEvent_t e : dp->ev) cudaEventDestroy(e);
  if (dp->st) cudaStreamDestroy(dp->st);
  if (dp->owned && dp->comm) nccl().CommDestroy(dp->comm);
  delete dp;
}
void dp_broadcast_f32(DpComm* dp, float* const* ptrs, const long long* numel, int n, long long max_gap, int root,
                      cudaStream_t st) {
  if (!dp || dp->nranks <= 1 || n <= 0) return;
  INB_CHECK(root >= 0 && root < dp->nranks, "broadcast root %d outside [0, %d)", root, dp->nranks);
  Nccl& N = nccl();
  nccl_check(N.GroupStart(), "ncclGroupStart");
  int i = 0;
  while (i < n) {
    float* base = ptrs[i];
    long long len = numel[i];
    int j = i + 1;
    while (j < n) {
      const long long gap = ptrs[j] - (base + len);
      if (gap < 0 || gap > max_gap) break;
      len += gap + numel[j];
      ++j;
    }
    const int rc = N.Broadcast(base, base, (size_t)len, ncclFloat32, root, dp->comm, st);
    if (rc != ncclSuccess) {
      N.GroupEnd();
      nccl_check(rc, "ncclBroadcast");
    }
    i = j;
  }
  nccl_check(N.GroupEnd(), "ncclGroupEnd");
}

}  // namespace inb
