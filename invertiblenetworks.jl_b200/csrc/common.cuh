// common.cuh - shared host/device helpers of libinb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <string>
#include <vector>

namespace inb {

// ---------------------------------------------------------------- errors
struct Error {
  int code;
  std::string msg;
};
void set_last_error(const std::string& s);
[[noreturn]] void fail(int code, const char* fmt, ...);

#define INB_CUDA(expr)                                                                   \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess)                                                               \
      ::inb::fail(2, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

#define INB_CHECK(cond, ...)                       \
  do {                                             \
    if (!(cond)) ::inb::fail(1, __VA_ARGS__);      \
  } while (0)

// ---------------------------------------------------------------- geometry / tensor views
// A tensor is (B, C, [D,] H, W) float32 with W fastest; channel stride = px = W*H*D.
struct Geo {
  int nd;        // 2 or 3
  int W, H, D;   // nx, ny, nz
  long long px;  // W*H*D
};
inline Geo make_geo(int nd, int nx, int ny, int nz) {
  Geo g;
  g.nd = nd;
  g.W = nx;
  g.H = ny;
  g.D = (nd == 3) ? nz : 1;
  g.px = (long long)g.W * g.H * g.D;
  return g;
}
inline Geo half_geo(const Geo& g) {
  return make_geo(g.nd, g.W / 2, g.H / 2, g.nd == 3 ? g.D / 2 : 1);
}

// A channel-range view into such a tensor: p points at channel 0 of the view in sample 0,
// bs is the batch stride in elements (>= C*px: lets split halves / latent blocks be addressed
// in place, no tensor_split / tensor_cat copies).
struct View {
  float* p;
  long long bs;
};
inline View view(float* p, long long bs) { return View{p, bs}; }
inline View sub(const View& v, long long chan, long long px) { return View{v.p + chan * px, v.bs}; }

// ---------------------------------------------------------------- workspace arena
// Bump allocator over one device block owned by a plan (or a stream-ordered temporary block for
// the layer-level entry points).  `dry` = sizing pass: nothing is launched, sizes are summed.
struct Arena {
  char* base = nullptr;
  size_t cap = 0, off = 0, peak = 0;
  bool dry = false;
  void* alloc_bytes(size_t n) {
    size_t a = (off + 255) & ~size_t(255);
    off = a + n;
    if (off > peak) peak = off;
    if (dry) return (void*)(uintptr_t)(256 + a);  // fake but aligned, never dereferenced
    if (off > cap) fail(3, "workspace arena overflow (%zu > %zu bytes)", off, cap);
    return base + a;
  }
  float* f32(size_t n) { return (float*)alloc_bytes(n * sizeof(float)); }
  double* f64(size_t n) { return (double*)alloc_bytes(n * sizeof(double)); }
  size_t mark() const { return off; }
  void release(size_t m) { off = m; }
};

// A second stream for small kernels nothing later in the same call waits for (the closed-form Householder / ActNorm
// gradient kernels: one CTA, fp64, ~15 us each, 96 of them per cfg2 backward).  fork() makes the lane wait for the
// work enqueued so far on the main stream, join() makes the main stream wait for the lane; inside a graph capture both
// become edges of the graph.  What such kernels read lives in `pool`, which is not recycled within a call.
struct SideLane {
  cudaStream_t st = nullptr;
  cudaEvent_t* ev = nullptr;
  int nev = 0, next = 0;
  char* pool = nullptr;
  size_t pool_bytes = 0, pool_off = 0;
  bool used = false;
  // two more streams for the weight gradients of a block when each of them fills less than half of the SMs (small batch
  // shards, coarse scales): the three split-K kernels then run side by side (op_wgrad2_tc_multi)
  cudaStream_t wst[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t* wev = nullptr;
  int nwev = 0, wnext = 0;
  // deferred mode (Ctx::wg_defer): all three kernels and their reduction leave the main stream and overlap the NEXT
  // flow step; the steps alternate between two workspace regions (wparity) and the main stream waits for wdone[q]
  // before a region is written again (drive_reverse)
  int wparity = 0;
  cudaEvent_t wdone[2] = {nullptr, nullptr};
  bool wpending[2] = {false, false};
  void wait_deferred(cudaStream_t main, int q) {
    if (wpending[q]) INB_CUDA(cudaStreamWaitEvent(main, wdone[q], 0));
    wpending[q] = false;
  }
  void* take(size_t n) {
    size_t a = (pool_off + 255) & ~size_t(255);
    if (!pool || a + n > pool_bytes || next + 2 > nev) return nullptr;
    pool_off = a + n;
    return pool + a;
  }
  void fork(cudaStream_t main) {
    INB_CUDA(cudaEventRecord(ev[next], main));
    INB_CUDA(cudaStreamWaitEvent(st, ev[next], 0));
    ++next;
    used = true;
  }
  void join(cudaStream_t main) {
    if (used) {
      INB_CUDA(cudaEventRecord(ev[next], st));
      INB_CUDA(cudaStreamWaitEvent(main, ev[next], 0));
    }
    used = false;
    next = 0;
    pool_off = 0;
  }
};

struct DpComm;  // dp.cuh
struct Ctx {
  cudaStream_t st;
  Arena* ar;
  int prec;  // INB_PREC_*
  SideLane* lane = nullptr;
  DpComm* dp = nullptr;   // data-parallel communicator attached to the plan (global-batch ActNorm statistics)
  bool wg_defer = false;  // weight gradients of this flow step may leave the main stream (SideLane::wparity)
  bool dry() const { return ar->dry; }
};

inline long long cdiv(long long a, long long b) { return (a + b - 1) / b; }
// INB_PREC_*: number of products per contraction term (3-term split or single pass) and the operand format
inline int prec_terms(int prec) { return (prec == 1 || prec == 3) ? 3 : 1; }
inline bool prec_f16(int prec) { return prec == 3; }

// ---------------------------------------------------------------- launch accounting / profiling
// Every op brackets its kernel launches with a Prof scope: the launch counter is always kept
// (inb_launch_count, the "gpu_launches" of bench.py); when profiling is enabled (inb_prof_enable)
// a CUDA event pair on the launching stream times the scope and the op's ALGORITHMIC flops/bytes
// are recorded beside it (bench.py's roofline object comes from these).
enum Family {
  F_SQUEEZE = 0, F_COPY, F_AN_STATS, F_AN_HH_FWD, F_HH_AN_INV, F_HH_AN_BWD, F_GRAD_FINISH, F_COUPLING_FWD,
  F_COUPLING_INV, F_COUPLING_BWD, F_PACK, F_CONV_SIMT, F_WGRAD_SIMT, F_CHANNEL_SUM, F_NLL, F_MISC,
  F_CONV_TC, F_WGRAD_TC, F_LAYOUT_TC, F_COL2IM, F_COUNT
};
struct Prof {
  cudaStream_t st;
  int fam;
  void* slot;
  Prof(const Ctx& c, int fam, int launches, double flops, double bytes);
  ~Prof();
};

bool prof_is_enabled();
long long launch_count_now();
void launch_count_add(long long n);

// Int(round(C/2)) with ties-to-even (dimensionality_operations.jl:408)
inline int split_k(int C) {
  int h = C / 2;
  if (C % 2 == 0) return h;
  // C/2 = h + 0.5 -> ties to even
  return (h % 2 == 0) ? h : h + 1;
}

}  // namespace inb
