// api.cu - the C ABI of libinb200 (include/inb200.h): argument checking, the plan / workspace,
// and the network drivers that walk the L x K flow steps inside the library.
#include "../../include/inb200.h"
#include "glow.cuh"
#include "dp.cuh"
#include <mutex>
#include <vector>

#include <cstring>
#include <memory>
#include <algorithm>
#include <stdexcept>
#include <initializer_list>
#include <cstdlib>

namespace inb {

static thread_local std::string g_last_error;
void set_last_error(const std::string& s) { g_last_error = s; }
void fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  throw Error{code, buf};
}

template <class F>
static int guarded(F&& f) {
  try {
    f();
    return INB_OK;
  } catch (const Error& e) {
    set_last_error(e.msg);
    return e.code;
  } catch (const std::exception& e) {
    set_last_error(std::string("internal error: ") + e.what());
    return INB_ERR_INVALID;
  } catch (...) {
    set_last_error("unknown internal error");
    return INB_ERR_INVALID;
  }
}

// stream-ordered temporary arena for the layer-level entry points.  The library keeps its own memory pool per
// device with an unlimited release threshold: the default pool hands its pages back to the driver at every
// synchronisation, which made each layer-level call re-map its whole workspace (tens of ms for a cfg2 block).
static cudaMemPool_t temp_pool() {
  static std::mutex mu;
  static cudaMemPool_t pools[64] = {};
  int dev = 0;
  INB_CUDA(cudaGetDevice(&dev));
  INB_CHECK(dev >= 0 && dev < 64, "device index out of range");
  std::lock_guard<std::mutex> lk(mu);
  if (!pools[dev]) {
    cudaMemPoolProps props{};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = dev;
    INB_CUDA(cudaMemPoolCreate(&pools[dev], &props));
    unsigned long long keep = ~0ull;
    INB_CUDA(cudaMemPoolSetAttribute(pools[dev], cudaMemPoolAttrReleaseThreshold, &keep));
  }
  return pools[dev];
}
struct TempArena {
  Arena ar;
  cudaStream_t st;
  TempArena(cudaStream_t s) : st(s) {}
  void reserve(size_t bytes) {
    void* p = nullptr;
    INB_CUDA(cudaMallocFromPoolAsync(&p, bytes + 512, temp_pool(), st));
    ar.base = (char*)p;
    ar.cap = bytes + 512;
    ar.off = 0;
    ar.dry = false;
  }
  ~TempArena() {
    if (ar.base) cudaFreeAsync(ar.base, st);
  }
};
// run `body` once dry to size the workspace, then for real
template <class F>
static void with_temp_arena(cudaStream_t st, int prec, F&& body) {
  Arena dry;
  dry.dry = true;
  Ctx dc{st, &dry, prec};
  body(dc);
  TempArena t(st);
  t.reserve(dry.peak);
  Ctx c{st, &t.ar, prec};
  body(c);
}

}  // namespace inb

using namespace inb;

// ====================================================================== plan
struct ScaleInfo {
  Geo g;        // geometry of the flow steps at this scale
  int C;        // channels during the flow steps
  int Ccond;    // condition channels at this scale
  int has_split;
  int kc;       // channels that continue after the split (== C when no split)
};

// Identity of a captured call: batch, communicator and every pointer it was captured with, folded into two independent
// 64-bit hashes (a replay with stale pointers would corrupt memory silently, so a single 64-bit hash is not enough)
struct GraphKey {
  uint64_t a = 0, b = 0;
  bool operator==(const GraphKey& o) const { return a == o.a && b == o.b; }
  bool operator!=(const GraphKey& o) const { return !(*this == o); }
  void add(uint64_t v) {
    a ^= v + 0x9E3779B97F4A7C15ull + (a << 6) + (a >> 2);
    b = (b ^ (v * 0xFF51AFD7ED558CCDull)) * 0xC4CEB9FE1A85EC53ull;
    b ^= b >> 29;
  }
};
// graph-replay state shared by the network plans (run_graphed)
struct GraphCache {
  // CUDA graphs of the network-level calls (slot 0 forward, 1 inverse, 2 backward): a call with the same
  // pointers and batch as the captured one replays ~1500 launches with a single cudaGraphLaunch
  // (a caller's allocator typically cycles through a few addresses for its outputs: kGraphWays entries per
  // slot, least recently used replaced)
  static constexpr int kGraphWays = 8;
  struct GraphSlot {
    GraphKey key;
    cudaGraphExec_t exec = nullptr;
    long long launches = 0;
    unsigned long long used = 0;
  } graphs[3][kGraphWays];
  unsigned long long graph_clock = 0;
  // per call type: hit (0) / miss (1) history of the last 32 calls, and counters for inb_graph_stats.  A caller
  // whose allocator never repeats an address combination would otherwise pay a capture + instantiate (tens
  // of ms) on every call: once the ways are full and misses keep coming, such calls are launched directly.
  uint32_t graph_hist[3] = {0, 0, 0};
  long long graph_captures = 0, graph_replays = 0, graph_direct = 0;
  cudaStream_t capture_stream = nullptr;
  void destroy_graphs() {
    for (auto& slot : graphs)
      for (auto& g : slot)
        if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
    if (capture_stream) { cudaStreamDestroy(capture_stream); capture_stream = nullptr; }
  }
};

struct inb_plan : GraphCache {
  inb_glow_desc d;
  bool cond;
  std::vector<ScaleInfo> sc;
  Geo g0;  // input geometry
  Arena ar;
  size_t persist = 0;  // bytes at the start of the arena that survive calls (logdet accumulator)
  size_t need = 0;     // workspace bytes (sizing pass at plan creation)
  double* ld = nullptr;
  SideLane lane;
  DpComm* dp = nullptr;  // attached communicator (inb_glow_plan_set_comm), not owned
  DpComm* dp_warm = nullptr;  // the communicator whose collectives have run once outside a capture
  std::vector<long long> numel;     // element count of every parameter (get_params order)
  std::vector<long long> flat_off;  // canonical flat layout: element offset of every parameter, 64-element aligned
  long long flat_total = 0;
};
struct inb_comm : inb::DpComm {};
namespace inb {
void dp_unique_id(char* id128);
DpComm* dp_create(int nranks, int rank, const char* id128);
DpComm* dp_wrap(void* nccl_comm);
void dp_destroy(DpComm* dp);
void dp_broadcast_f32(DpComm* dp, float* const* ptrs, const long long* numel, int n, long long max_gap, int root,
                      cudaStream_t st);
}

static int an_index(const inb_plan* p, int i, int j, int which) { return 2 * (i * p->d.K + j) + which; }
static int cl_index(const inb_plan* p, int i, int j, int which) {
  return 2 * p->d.L * p->d.K + (p->cond ? 2 : 0) + 8 * (i * p->d.K + j) + which;
}
static int anc_index(const inb_plan* p, int which) { return 2 * p->d.L * p->d.K + which; }

// element count of parameter `index` (get_params order)
static long long plan_param_numel(const inb_plan* p, int index) {
  const int LK = p->d.L * p->d.K;
  if (index < 2 * LK) return p->sc[(index / 2) / p->d.K].C;
  index -= 2 * LK;
  if (p->cond) {
    if (index < 2) return p->d.n_cond;
    index -= 2;
  }
  const int layer = index / 8, which = index % 8;
  const ScaleInfo& s = p->sc[layer / p->d.K];
  const int C1 = split_k(s.C), nh = p->d.n_hidden;
  const int Cin = s.C - C1 + (p->cond ? s.Ccond : 0), Cout = 2 * C1;
  long long t1 = 1, t2 = 1;
  for (int a = 0; a < p->d.ndims; ++a) { t1 *= p->d.k1; t2 *= p->d.k2; }
  switch (which) {
    case 0: case 1: case 2: return s.C;
    case 3: return t1 * Cin * nh;
    case 4: return t2 * nh * nh;
    case 5: return t1 * Cout * nh;
    default: return nh;
  }
}
static void plan_layout(inb_plan* p) {
  const int n = 10 * p->d.L * p->d.K + (p->cond ? 2 : 0);
  p->numel.resize(n);
  p->flat_off.resize(n);
  long long o = 0;
  for (int i = 0; i < n; ++i) {
    p->numel[i] = plan_param_numel(p, i);
    p->flat_off[i] = o;
    o += (p->numel[i] + 63) / 64 * 64;  // every tensor starts on a 256-byte boundary
  }
  p->flat_total = o;
}
// the caller's table follows the canonical flat layout (inb_glow_flat_layout): the gaps between consecutive tensors are
// the caller's own alignment padding and a bucket may be reduced as one range across them
static long long table_max_gap(const inb_plan* p, float* const* t) {
  for (size_t i = 0; i < p->flat_off.size(); ++i)
    if (t[i] != t[0] + p->flat_off[i]) return 0;
  return 63;
}


static void plan_scales(inb_plan* p) {
  const inb_glow_desc& d = p->d;
  p->g0 = make_geo(d.ndims, d.nx, d.ny, d.nz);
  Geo g = p->g0;
  int c = d.n_in, cc = d.n_cond;
  const int f = 1 << d.ndims;
  for (int i = 0; i < d.L; ++i) {
    ScaleInfo s{};
    if (d.split_scales) {
      INB_CHECK(!(g.W % 2) && !(g.H % 2) && (g.nd == 2 || !(g.D % 2)),
                "Input dimensions must be multiple of 2");  // dimensionality_operations.jl:82-84
      c *= f;
      cc *= f;
      g = half_geo(g);
    }
    s.g = g;
    s.C = c;
    s.Ccond = cc;
    // invertible_network_glow.jl:120 (i < L || i == 1) ; conditional_glow.jl:122 (i < L)
    s.has_split = d.split_scales && (i < d.L - 1 || (!p->cond && i == 0));
    s.kc = s.has_split ? split_k(c) : c;
    INB_CHECK(c >= 2, "a coupling layer needs at least 2 channels (scale %d has %d)", i + 1, c);
    p->sc.push_back(s);
    if (d.split_scales && i < d.L - 1) c = c / 2;  // ctor :100
    if (s.has_split) INB_CHECK(s.kc == c || i == d.L - 1, "internal: split bookkeeping");
  }
}

static FlowShape flow_shape(const inb_plan* p, int i, int B) {
  const ScaleInfo& s = p->sc[i];
  FlowShape f{};
  f.g = s.g;
  f.B = B;
  f.C = s.C;
  f.ccond = p->cond ? s.Ccond : 0;
  f.nh = p->d.n_hidden;
  f.k1 = p->d.k1;
  f.k2 = p->d.k2;
  f.low = p->d.sig_low;
  f.high = p->d.sig_high;
  f.logdet = p->d.logdet;
  f.freeze = p->d.freeze_conv;
  return f;
}
static FlowParams flow_params(const inb_plan* p, int i, int j, float* const* prm) {
  FlowParams fp{};
  if (!prm) return fp;
  fp.s = prm[an_index(p, i, j, 0)];
  fp.b = prm[an_index(p, i, j, 1)];
  fp.v1 = prm[cl_index(p, i, j, 0)];
  fp.v2 = prm[cl_index(p, i, j, 1)];
  fp.v3 = prm[cl_index(p, i, j, 2)];
  fp.rb = RBParams{prm[cl_index(p, i, j, 3)], prm[cl_index(p, i, j, 4)], prm[cl_index(p, i, j, 5)],
                   prm[cl_index(p, i, j, 6)], prm[cl_index(p, i, j, 7)]};
  return fp;
}
static FlowGrads flow_grads(const inb_plan* p, int i, int j, float* const* gr) {
  FlowGrads fg{};
  if (!gr) return fg;
  fg.s = gr[an_index(p, i, j, 0)];
  fg.b = gr[an_index(p, i, j, 1)];
  fg.v1 = gr[cl_index(p, i, j, 0)];
  fg.v2 = gr[cl_index(p, i, j, 1)];
  fg.v3 = gr[cl_index(p, i, j, 2)];
  fg.rb = RBGrads{gr[cl_index(p, i, j, 3)], gr[cl_index(p, i, j, 4)], gr[cl_index(p, i, j, 5)],
                  gr[cl_index(p, i, j, 6)], gr[cl_index(p, i, j, 7)]};
  return fg;
}

// element offset of latent block `i` in the flat Z vector and of the tail (block index L)
static long long z_offset(const inb_plan* p, int B, int upto) {
  long long off = 0;
  for (int i = 0; i < upto && i < (int)p->sc.size(); ++i) {
    const ScaleInfo& s = p->sc[i];
    if (s.has_split) off += (long long)B * (s.C - s.kc) * s.g.px;
  }
  return off;
}

// ---------------------------------------------------------------- drivers
// `prm` null in the sizing pass (pointers are never dereferenced on the host).
static void drive_forward(inb_plan* p, Ctx& c, int B, const float* X, const float* Cnd, float* const* prm,
                          float* Z, float* ZC, float* logdet, int init) {
  const inb_glow_desc& d = p->d;
  const long long tot = (long long)d.n_in * p->g0.px;  // elements per sample
  size_t m = c.ar->mark();
  float* buf[2] = {c.ar->f32((size_t)B * tot), c.ar->f32((size_t)B * tot)};
  float* cbuf[2] = {nullptr, nullptr};
  const long long ctot = (long long)d.n_cond * p->g0.px;
  if (p->cond) {
    cbuf[0] = c.ar->f32((size_t)B * ctot);
    cbuf[1] = c.ar->f32((size_t)B * ctot);
  }
  if (d.logdet) op_zero(c, p->ld, sizeof(double));
  View cur = view(const_cast<float*>(X), tot);
  int which = 0;  // buffer that is free to write
  View cond = view(nullptr, 0);
  int cwhich = 0;
  if (p->cond) {
    // C = AN_C.forward(C), logdet = false                        conditional_glow.jl:111
    float* s = prm ? prm[anc_index(p, 0)] : nullptr;
    float* b = prm ? prm[anc_index(p, 1)] : nullptr;
    View cin = view(const_cast<float*>(Cnd), ctot);
    if (init) op_actnorm_init(c, p->g0.px, B, d.n_cond, cin, s, b);
    bool last = !d.split_scales;
    View cout = view(last ? ZC : cbuf[0], ctot);
    op_an_hh_fwd(c, p->g0.px, B, d.n_cond, cin, cout, s, b, nullptr, nullptr, nullptr, nullptr);
    cond = cout;
    cwhich = 1;
  }
  Geo g = p->g0;
  int chan = d.n_in, cchan = d.n_cond;
  for (int i = 0; i < d.L; ++i) {
    const ScaleInfo& s = p->sc[i];
    if (d.split_scales) {
      View out = view(buf[which], (long long)s.C * s.g.px);
      op_squeeze(c, g, B, chan, cur, out);  // :114
      cur = out;
      which ^= 1;
      if (p->cond) {
        bool last = (i == d.L - 1);
        View co = view(last ? ZC : cbuf[cwhich], (long long)s.Ccond * s.g.px);
        op_squeeze(c, g, B, cchan, cond, co);  // conditional_glow.jl:116
        cond = co;
        cwhich ^= 1;
      }
      g = s.g;
      chan = s.C;
      cchan = s.Ccond;
    }
    FlowShape f = flow_shape(p, i, B);
    // the chain operands of the scale's K blocks are packed in one launch before the first step
    size_t mpack = c.ar->mark();
    std::vector<PackedW> packs(d.K);
    bool prepacked = false;
    {  // also in the sizing pass (no parameters there): the planes must be counted
      std::vector<RBParams> rbp(d.K);
      for (int j = 0; j < d.K; ++j) rbp[j] = flow_params(p, i, j, prm).rb;
      prepacked = rb_prepack_chain(c, f.rb(), rbp.data(), d.K, 0, packs.data());
    }
    for (int j = 0; j < d.K; ++j) {
      FlowParams fp = flow_params(p, i, j, prm);
      if (prepacked) fp.rb.pre[0] = &packs[j];
      if (init) op_actnorm_init(c, s.g.px, B, s.C, cur, const_cast<float*>(fp.s), const_cast<float*>(fp.b));
      View out = view(buf[which], (long long)s.C * s.g.px);
      flow_forward(c, f, cur, out, cond, fp, p->ld);  // :116-118
      cur = out;
      which ^= 1;
    }
    c.ar->release(mpack);
    if (s.has_split) {  // :120-124: X = first part, Z = second part
      long long off = z_offset(p, B, i);
      int zc = s.C - s.kc;
      op_copy(c, s.g.px, B, zc, sub(cur, s.kc, s.g.px), view(Z + off, (long long)zc * s.g.px));
      chan = s.kc;
    }
  }
  // tail: cat_states (:126) / plain output
  {
    const ScaleInfo& s = p->sc[d.L - 1];
    long long off = z_offset(p, B, d.L);
    op_copy(c, s.g.px, B, chan, cur, view(Z + off, (long long)chan * s.g.px));
  }
  if (d.logdet && logdet) op_ld_finish(c, p->ld, logdet);
  c.ar->release(m);
}

static bool wgrad_defer_enabled() {
  static const bool on = [] { const char* e = getenv("INB_WGRAD_DEFER"); return !(e && e[0] == '0'); }();
  return on;
}
// inverse (grads == false) and backward (grads == true) share the reverse sweep
static void drive_reverse(inb_plan* p, Ctx& c, int B, bool grads, const float* dZ, const float* Z,
                          const float* ZC, float* const* prm, float* const* gr, float* dX, float* X,
                          float* dC) {
  const inb_glow_desc& d = p->d;
  const long long tot = (long long)d.n_in * p->g0.px;
  const long long ctot = (long long)d.n_cond * p->g0.px;
  size_t m = c.ar->mark();
  float* yb[2] = {c.ar->f32((size_t)B * tot), c.ar->f32((size_t)B * tot)};
  float* db[2] = {nullptr, nullptr};
  if (grads) {
    db[0] = c.ar->f32((size_t)B * tot);
    db[1] = c.ar->f32((size_t)B * tot);
  }
  float* cb[2] = {nullptr, nullptr};
  float* dcb[2] = {nullptr, nullptr};
  if (p->cond) {
    cb[0] = c.ar->f32((size_t)B * ctot);
    cb[1] = c.ar->f32((size_t)B * ctot);
    if (grads) {
      dcb[0] = c.ar->f32((size_t)B * ctot);
      dcb[1] = c.ar->f32((size_t)B * ctot);
    }
  }
  int which = 0, cwhich = 0;
  // tail block -> first kc channels of the scale-L buffer
  const ScaleInfo& sl = p->sc[d.L - 1];
  {
    long long off = z_offset(p, B, d.L);
    View dst = view(yb[which], (long long)sl.C * sl.g.px);
    op_copy(c, sl.g.px, B, sl.kc, view(const_cast<float*>(Z) + off, (long long)sl.kc * sl.g.px), dst);
    if (grads) {
      View ddst = view(db[which], (long long)sl.C * sl.g.px);
      op_copy(c, sl.g.px, B, sl.kc, view(const_cast<float*>(dZ) + off, (long long)sl.kc * sl.g.px), ddst);
    }
  }
  View cond = view(const_cast<float*>(ZC), p->cond ? (long long)sl.Ccond * sl.g.px : 0);
  View dcond = view(nullptr, 0);
  if (p->cond && grads) {
    dcond = view(dcb[cwhich], cond.bs);
    op_zero(c, dcond.p, (size_t)B * ctot * sizeof(float));  // conditional_glow.jl:160
  }
  for (int i = d.L - 1; i >= 0; --i) {
    const ScaleInfo& s = p->sc[i];
    const long long bs = (long long)s.C * s.g.px;
    View y = view(yb[which], bs);
    View dy = view(grads ? db[which] : nullptr, bs);
    if (s.has_split) {  // :168-169 tensor_cat with the saved latent
      long long off = z_offset(p, B, i);
      int zc = s.C - s.kc;
      op_copy(c, s.g.px, B, zc, view(const_cast<float*>(Z) + off, (long long)zc * s.g.px), sub(y, s.kc, s.g.px));
      if (grads)
        op_copy(c, s.g.px, B, zc, view(const_cast<float*>(dZ) + off, (long long)zc * s.g.px), sub(dy, s.kc, s.g.px));
    }
    FlowShape f = flow_shape(p, i, B);
    size_t mpack = c.ar->mark();
    std::vector<PackedW> packs_f(d.K), packs_b(d.K);
    bool pre_f = false, pre_b = false;
    {
      std::vector<RBParams> rbp(d.K);
      for (int j = 0; j < d.K; ++j) rbp[j] = flow_params(p, i, j, prm).rb;
      pre_f = rb_prepack_chain(c, f.rb(), rbp.data(), d.K, 0, packs_f.data());
      if (grads) pre_b = rb_prepack_chain(c, f.rb(), rbp.data(), d.K, 1, packs_b.data());
    }
    // Small shards / coarse scales (a weight-gradient kernel fills at most half of the SMs): the weight gradients of a
    // step leave the main stream and overlap the next step.  The steps then alternate between two workspace regions
    // (the second starts `foot` bytes above the first), and a region is reused only after its gradients are done.
    size_t foot = 0;
    if (grads && pre_b && ((long long)B * s.g.px / 512 <= 74 || wgrad_overlap_ctas() > 0) && wgrad_defer_enabled()) {
      Arena tmp;
      tmp.dry = true;
      Ctx tc{nullptr, &tmp, c.prec};
      FlowParams fp0 = flow_params(p, i, 0, nullptr);
      fp0.rb.pre[0] = &packs_f[0];
      fp0.rb.pre[1] = &packs_b[0];
      flow_backward(tc, f, dy, y, dy, y, cond, dcond, fp0, FlowGrads{});
      foot = (tmp.peak + 1023) & ~size_t(1023);
    }
    for (int j = d.K - 1; j >= 0; --j) {
      FlowParams fp = flow_params(p, i, j, prm);
      if (pre_f) fp.rb.pre[0] = &packs_f[j];
      if (pre_b) fp.rb.pre[1] = &packs_b[j];
      View xo = y, dxo = dy;
      const bool final_step = (i == 0 && j == 0 && !d.split_scales);
      if (final_step) {  // write straight into the caller's buffers
        xo = view(X, bs);
        dxo = view(dX, bs);
      }
      if (grads) {
        FlowGrads fg = flow_grads(p, i, j, gr);
        const size_t mstep = c.ar->mark();
        Ctx sc = c;
        if (foot) {
          const int q = (d.K - 1 - j) & 1;
          if (q) c.ar->alloc_bytes(foot);
          if (c.lane && !c.dry()) {
            c.lane->wait_deferred(c.st, q);
            c.lane->wparity = q;
            sc.wg_defer = true;
          }
        }
        flow_backward(sc, f, dy, y, dxo, xo, cond, dcond, fp, fg);  // :173-174
        c.ar->release(mstep);
      } else {
        flow_inverse(c, f, y, xo, cond, fp);  // :139-140
      }
    }
    if (foot && c.lane && !c.dry()) {  // the scale's workspace is recycled below
      c.lane->wait_deferred(c.st, 0);
      c.lane->wait_deferred(c.st, 1);
    }
    c.ar->release(mpack);
    if (grads && gr && c.dp && c.dp->nranks > 1 && !c.dry()) {
      // data parallel: this scale's 10*K gradients are final - average them over the ranks on the
      // communicator's stream while the remaining scales run (two ranges when the caller's table is the flat layout)
      dp_fork(c.dp, c.st, (c.lane && c.lane->used) ? c.lane->st : nullptr);
      const long long gap = table_max_gap(p, gr);
      dp_allreduce_avg_bucket(c.dp, gr, p->numel.data(), an_index(p, i, 0, 0), 2 * d.K, gap, c.dp->st);
      dp_allreduce_avg_bucket(c.dp, gr, p->numel.data(), cl_index(p, i, 0, 0), 8 * d.K, gap, c.dp->st);
    }
    if (d.split_scales) {  // :186-187 unsqueeze
      Geo gout = (i == 0) ? p->g0 : make_geo(d.ndims, s.g.W * 2, s.g.H * 2, s.g.D * 2);
      int cout = s.C >> d.ndims;
      long long obs = (i == 0) ? tot : (long long)p->sc[i - 1].C * p->sc[i - 1].g.px;
      float* xdst = (i == 0) ? X : yb[which ^ 1];
      op_unsqueeze(c, gout, B, cout, y, view(xdst, obs));
      if (grads) {
        float* ddst = (i == 0) ? dX : db[which ^ 1];
        op_unsqueeze(c, gout, B, cout, dy, view(ddst, obs));
      }
      which ^= 1;
      if (p->cond) {
        int ccout = s.Ccond >> d.ndims;
        long long cbs = (long long)ccout * gout.px;
        View cn = view(cb[cwhich], cbs);
        op_unsqueeze(c, gout, B, ccout, cond, cn);  // conditional_glow.jl:171
        if (grads) {
          View dn = view(dcb[cwhich ^ 1], cbs);
          op_unsqueeze(c, gout, B, ccout, dcond, dn);  // :172
          dcond = dn;
        }
        cond = cn;
        cwhich ^= 1;
      }
    }
  }
  if (p->cond && grads) {
    // dC, C = AN_C.backward(dC, C)                               conditional_glow.jl:179
    float* s = prm ? prm[anc_index(p, 0)] : nullptr;
    float* b = prm ? prm[anc_index(p, 1)] : nullptr;
    size_t m2 = c.ar->mark();
    double* dsdb = c.ar->f64(2 * (size_t)d.n_cond);
    op_zero(c, dsdb, 2 * d.n_cond * sizeof(double));
    // cond lives in cb[] or is the caller's ZC (no split scales): never write into ZC
    float* scratch = c.ar->f32((size_t)B * ctot);
    op_hh_an_bwd(c, p->g0.px, B, d.n_cond, dcond, cond, view(dC, ctot), view(scratch, ctot), s, b, nullptr,
                 nullptr, nullptr, nullptr, dsdb);
    op_an_grad_finish(c, d.n_cond, p->g0.px, dsdb, s, 0, gr ? gr[anc_index(p, 0)] : nullptr,
                      gr ? gr[anc_index(p, 1)] : nullptr);
    c.ar->release(m2);
    if (gr && c.dp && c.dp->nranks > 1 && !c.dry()) {
      dp_fork(c.dp, c.st, nullptr);
      dp_allreduce_avg_bucket(c.dp, gr, p->numel.data(), anc_index(p, 0), 2, table_max_gap(p, gr), c.dp->st);
    }
  }
  if (c.dp && !c.dry()) dp_join(c.dp, c.st);  // the averaged gradients belong to this call
  if (c.lane) c.lane->join(c.st);  // the gradient kernels on the side lane belong to this call
  c.ar->release(m);
}

static void check_desc(const inb_glow_desc* d) {
  INB_CHECK(d != nullptr, "null descriptor");
  INB_CHECK(d->ndims == 2 || d->ndims == 3, "ndims must be 2 or 3 (got %d)", d->ndims);
  INB_CHECK(d->nx > 0 && d->ny > 0 && (d->ndims == 2 || d->nz > 0), "spatial sizes must be positive");
  INB_CHECK(d->n_in >= 1 && d->n_hidden >= 1 && d->L >= 1 && d->K >= 1 && d->batch >= 1,
            "n_in, n_hidden, L, K, batch must be >= 1");
  INB_CHECK(d->n_cond >= 0, "n_cond must be >= 0");
  INB_CHECK((d->k1 == 1 || d->k1 == 3) && (d->k2 == 1 || d->k2 == 3),
            "supported ResidualBlock kernel sizes are 1 and 3 (got k1=%d k2=%d)", d->k1, d->k2);
  INB_CHECK(d->p1 == (d->k1 - 1) / 2 && d->p2 == (d->k2 - 1) / 2,
            "only 'same' padding is supported (p = (k-1)/2; got p1=%d p2=%d)", d->p1, d->p2);
  INB_CHECK(d->precision >= 0 && d->precision <= 3, "unknown precision mode %d", d->precision);
  INB_CHECK(d->sig_high > d->sig_low, "sigmoid high must exceed low");
}

static void check_call(inb_plan* p, int batch, bool want_cond) {
  INB_CHECK(p != nullptr, "null plan");
  INB_CHECK(p->cond == want_cond, want_cond ? "plan is not conditional (n_cond == 0)"
                                            : "plan is conditional: use the inb_cglow_* entry points");
  INB_CHECK(batch >= 1 && batch <= p->d.batch, "batch %d outside the plan's range [1, %d]", batch, p->d.batch);
}
static Ctx call_ctx(inb_plan* p, void* stream) {
  if (!p->ar.base) {
    void* base = nullptr;
    cudaError_t e = cudaMalloc(&base, p->need);
    if (e != cudaSuccess) fail(INB_ERR_NOMEM, "workspace of %zu bytes: %s", p->need, cudaGetErrorString(e));
    p->ar.base = (char*)base;
    p->ar.cap = p->need;
    p->ar.dry = false;
    p->ar.off = 0;
    p->ld = (double*)p->ar.alloc_bytes(256);
    p->persist = p->ar.off;
  }
  p->ar.off = p->persist;
  static const bool no_lane = [] { const char* e = getenv("INB_SIDE_LANE"); return e && e[0] == '0'; }();
  if (!p->lane.st && !no_lane) {  // side lane: stream, L*K + 2 events, one Gram block per flow step
    const int steps = p->d.L * p->d.K;
    int cmax = 1;
    for (const ScaleInfo& s : p->sc) cmax = std::max(cmax, s.C);
    p->lane.nev = steps + 2;
    p->lane.ev = new cudaEvent_t[p->lane.nev];
    for (int i = 0; i < p->lane.nev; ++i) INB_CUDA(cudaEventCreateWithFlags(&p->lane.ev[i], cudaEventDisableTiming));
    p->lane.pool_bytes = (size_t)steps * ((((size_t)cmax * cmax + 2 * cmax) * sizeof(double) + 255) & ~size_t(255));
    INB_CUDA(cudaMalloc(&p->lane.pool, p->lane.pool_bytes));
    INB_CUDA(cudaStreamCreateWithFlags(&p->lane.st, cudaStreamNonBlocking));
    static const bool no_wlanes = [] { const char* e = getenv("INB_WGRAD_LANES"); return e && e[0] == '0'; }();
    if (!no_wlanes) {
      p->lane.nwev = 4 * steps;
      p->lane.wev = new cudaEvent_t[p->lane.nwev];
      for (int i = 0; i < p->lane.nwev; ++i) INB_CUDA(cudaEventCreateWithFlags(&p->lane.wev[i], cudaEventDisableTiming));
      INB_CUDA(cudaStreamCreateWithFlags(&p->lane.wst[0], cudaStreamNonBlocking));
      INB_CUDA(cudaStreamCreateWithFlags(&p->lane.wst[1], cudaStreamNonBlocking));
      INB_CUDA(cudaStreamCreateWithFlags(&p->lane.wst[2], cudaStreamNonBlocking));
    }
  }
  Ctx c{(cudaStream_t)stream, &p->ar, p->d.precision};
  if (p->lane.st) {
    p->lane.next = 0;
    p->lane.pool_off = 0;
    p->lane.used = false;
    p->lane.wnext = 0;
    p->lane.wparity = 0;
    p->lane.wpending[0] = p->lane.wpending[1] = false;
    c.lane = &p->lane;
  }
  c.dp = p->dp;
  return c;
}

// ---------------------------------------------------------------- CUDA-graph replay of whole-network calls
static GraphKey key_of(const inb_plan* p, int batch, std::initializer_list<const void*> ptrs, float* const* params,
                       float* const* grads) {
  GraphKey k;
  k.add(0x1234567ull);
  k.add((uint64_t)batch);
  k.add((uint64_t)(uintptr_t)p->dp);  // a graph captured with a communicator holds its collectives
  for (const void* q : ptrs) k.add((uint64_t)(uintptr_t)q);
  const int n = 10 * p->d.L * p->d.K + (p->cond ? 2 : 0);
  for (int i = 0; i < n; ++i) k.add((uint64_t)(uintptr_t)params[i]);
  if (grads)
    for (int i = 0; i < n; ++i) k.add((uint64_t)(uintptr_t)grads[i]);
  k.a |= 1ull;
  return k;
}
static bool graphs_enabled() {
  static const bool on = [] {
    const char* e = getenv("INB_GRAPHS");
    return !(e && e[0] == '0');
  }();
  return on;
}
// `enqueue(ctx)` launches the call's kernels on ctx.st.  First call with a given key: capture on the plan's
// private stream (nothing executes during capture), instantiate, launch on the caller's stream; later calls
// with the same key: one cudaGraphLaunch.  Falls back to direct launches while profiling, inside a caller's own
// capture, or with INB_GRAPHS=0.
// With a communicator attached, the first backward of a plan is launched kernel by kernel: NCCL sets up its channels and
// buffers for the message sizes of this plan on first use, which must not happen inside a stream capture.
static bool dp_needs_warmup(inb_plan* p, int slot) {
  if (!p->dp || p->dp->nranks <= 1 || slot != 2) return false;
  if (p->dp_warm == p->dp) return false;
  p->dp_warm = p->dp;
  return true;
}
template <class Plan>
static bool dp_needs_warmup(Plan*, int) { return false; }
template <class Plan, class F>
static void run_graphed(Plan* p, int slot, GraphKey key, void* stream, F&& enqueue) {
  cudaStream_t st = (cudaStream_t)stream;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  const bool capturing = st != nullptr && cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs != cudaStreamCaptureStatusNone;
  cudaGetLastError();
  if (!graphs_enabled() || capturing || prof_is_enabled() || dp_needs_warmup(p, slot)) {
    Ctx c = call_ctx(p, stream);
    enqueue(c);
    return;
  }
  GraphCache::GraphSlot* gp = nullptr;
  bool full = true;
  for (auto& w : p->graphs[slot]) {
    if (w.exec && w.key == key) gp = &w;
    if (!w.exec) full = false;
  }
  p->graph_hist[slot] = (p->graph_hist[slot] << 1) | (gp ? 0u : 1u);
  if (!gp) {
    if (full && __builtin_popcount(p->graph_hist[slot]) > 8) {  // thrashing: no point in capturing this call
      Ctx c = call_ctx(p, stream);
      enqueue(c);
      ++p->graph_direct;
      return;
    }
    // least recently used way
    gp = &p->graphs[slot][0];
    for (auto& w : p->graphs[slot])
      if (w.used < gp->used) gp = &w;
  }
  GraphCache::GraphSlot& g = *gp;
  g.used = ++p->graph_clock;
  if (g.exec == nullptr || g.key != key) {
    if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
    if (!p->capture_stream) INB_CUDA(cudaStreamCreateWithFlags(&p->capture_stream, cudaStreamNonBlocking));
    Ctx c = call_ctx(p, p->capture_stream);
    const long long n0 = launch_count_now();
    INB_CUDA(cudaStreamBeginCapture(p->capture_stream, cudaStreamCaptureModeThreadLocal));
    cudaGraph_t graph = nullptr;
    try {
      enqueue(c);
    } catch (...) {
      cudaStreamEndCapture(p->capture_stream, &graph);
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();
      throw;
    }
    INB_CUDA(cudaStreamEndCapture(p->capture_stream, &graph));
    g.launches = launch_count_now() - n0;
    launch_count_add(-g.launches);  // counted again at every replay below
    cudaError_t e = cudaGraphInstantiate(&g.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) {
      g.exec = nullptr;
      fail(2, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
    }
    g.key = key;
    ++p->graph_captures;
  }
  INB_CUDA(cudaGraphLaunch(g.exec, st));
  launch_count_add(g.launches);
  ++p->graph_replays;
}

extern "C" {

const char* inb_last_error(void) { return g_last_error.c_str(); }
int inb_version(void) { return 100; }
int inb_device_ok(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) return 0;
  int dev = 0;
  cudaDeviceProp pr;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&pr, dev) != cudaSuccess) return 0;
  return pr.major == 10 ? 1 : 0;
}

int inb_glow_plan_create(const inb_glow_desc* desc, inb_plan** out) {
  return guarded([&] {
    INB_CHECK(out != nullptr, "null output");
    *out = nullptr;
    check_desc(desc);
    std::unique_ptr<inb_plan> p(new inb_plan());
    p->d = *desc;
    if (p->d.n_in == 1 && p->d.n_cond == 0) p->d.split_scales = 1;  // invertible_network_glow.jl:79
    if (p->d.ndims == 2) p->d.nz = 1;
    p->cond = p->d.n_cond > 0;
    if (p->cond) p->d.logdet = 1;  // the conditional network always carries logdet (:107-130)
    plan_scales(p.get());
    plan_layout(p.get());
    // sizing pass
    Arena dry;
    dry.dry = true;
    dry.alloc_bytes(256);  // persistent: logdet accumulator
    size_t persist = dry.off;
    Ctx dc{nullptr, &dry, p->d.precision};
    const int B = p->d.batch;
    drive_forward(p.get(), dc, B, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 1);
    drive_reverse(p.get(), dc, B, true, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
    // the device block is allocated on first use, so a plan can be created (and its parameter /
    // latent bookkeeping queried) on a host without a GPU
    p->need = dry.peak + 1024;
    (void)persist;
    *out = p.release();
  });
}

int inb_glow_plan_destroy(inb_plan* p) {
  return guarded([&] {
    if (!p) return;
    p->destroy_graphs();
    if (p->lane.st) cudaStreamDestroy(p->lane.st);
    for (int i = 0; i < p->lane.nev; ++i) cudaEventDestroy(p->lane.ev[i]);
    delete[] p->lane.ev;
    if (p->lane.pool) cudaFree(p->lane.pool);
    for (int i = 0; i < 3; ++i)
      if (p->lane.wst[i]) cudaStreamDestroy(p->lane.wst[i]);
    for (int i = 0; i < p->lane.nwev; ++i) cudaEventDestroy(p->lane.wev[i]);
    delete[] p->lane.wev;
    if (p->ar.base) cudaFree(p->ar.base);
    delete p;
  });
}

int inb_glow_graph_stats(const inb_plan* p, long long* captures, long long* replays, long long* direct) {
  return guarded([&] {
    INB_CHECK(p != nullptr, "null plan");
    if (captures) *captures = p->graph_captures;
    if (replays) *replays = p->graph_replays;
    if (direct) *direct = p->graph_direct;
  });
}

int inb_glow_num_params(const inb_plan* p) {
  if (!p) return -1;
  return 10 * p->d.L * p->d.K + (p->cond ? 2 : 0);
}

int inb_glow_param_numel(const inb_plan* p, int index, long long* numel) {
  return guarded([&] {
    INB_CHECK(p && numel, "null argument");
    INB_CHECK(index >= 0 && index < inb_glow_num_params(p), "parameter index %d out of range", index);
    *numel = p->numel[index];
  });
}

int inb_glow_flat_layout(const inb_plan* p, long long* offsets, long long* total) {
  return guarded([&] {
    INB_CHECK(p != nullptr, "null plan");
    if (offsets)
      for (size_t i = 0; i < p->flat_off.size(); ++i) offsets[i] = p->flat_off[i];
    if (total) *total = p->flat_total;
  });
}

long long inb_glow_workspace_bytes(const inb_plan* p) { return p ? (long long)p->need : -1; }

int inb_glow_zdims(const inb_plan* p, int batch, int scale, int* dims5) {
  if (!p || !dims5 || scale < 0 || scale >= p->d.L) return -1;
  const ScaleInfo& s = p->sc[scale];
  int zc = s.has_split ? s.C - s.kc : s.C;
  int n = 0;
  dims5[n++] = batch;
  dims5[n++] = zc;
  if (p->d.ndims == 3) dims5[n++] = s.g.D;
  dims5[n++] = s.g.H;
  dims5[n++] = s.g.W;
  return n;
}

// ---------------------------------------------------------------- data-parallel plane
int inb_comm_unique_id(char* id128) {
  return guarded([&] {
    INB_CHECK(id128 != nullptr, "null id buffer");
    dp_unique_id(id128);
  });
}
int inb_comm_create(int nranks, int rank, const char* id128, inb_comm** out) {
  return guarded([&] {
    INB_CHECK(out && id128, "null argument");
    *out = static_cast<inb_comm*>(dp_create(nranks, rank, id128));
  });
}
int inb_comm_wrap(void* nccl_comm, inb_comm** out) {
  return guarded([&] {
    INB_CHECK(out != nullptr, "null output");
    *out = static_cast<inb_comm*>(dp_wrap(nccl_comm));
  });
}
int inb_comm_destroy(inb_comm* comm) {
  return guarded([&] { dp_destroy(comm); });
}
int inb_comm_info(const inb_comm* comm, int* nranks, int* rank, long long* allreduce_calls, long long* allreduce_bytes) {
  return guarded([&] {
    INB_CHECK(comm != nullptr, "null communicator");
    if (nranks) *nranks = comm->nranks;
    if (rank) *rank = comm->rank;
    if (allreduce_calls) *allreduce_calls = comm->calls;
    if (allreduce_bytes) *allreduce_bytes = comm->bytes;
  });
}
int inb_glow_plan_set_comm(inb_plan* p, inb_comm* comm) {
  return guarded([&] {
    INB_CHECK(p != nullptr, "null plan");
    if (p->dp != comm) {
      // graphs captured with the previous communicator hold its collectives: NCCL requires them to be destroyed
      // before the communicator is (ncclCommDestroy otherwise waits for them forever)
      for (auto& g : p->graphs[2])
        if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; g.key = GraphKey(); }
    }
    p->dp = comm;
  });
}
int inb_allreduce_grads(inb_plan* p, float* const* grads, inb_comm* comm, void* stream) {
  return guarded([&] {
    INB_CHECK(p && grads && comm, "null argument");
    dp_allreduce_avg_bucket(comm, grads, p->numel.data(), 0, (int)p->numel.size(), table_max_gap(p, grads),
                            (cudaStream_t)stream);
  });
}
int inb_broadcast_params(inb_plan* p, float* const* params, inb_comm* comm, int root, void* stream) {
  return guarded([&] {
    INB_CHECK(p && params && comm, "null argument");
    dp_broadcast_f32(comm, params, p->numel.data(), (int)p->numel.size(), table_max_gap(p, params), root,
                     (cudaStream_t)stream);
  });
}

int inb_glow_forward(inb_plan* p, int batch, const float* X, float* const* params, float* Z, float* logdet,
                     int init_actnorm, void* stream) {
  return guarded([&] {
    check_call(p, batch, false);
    INB_CHECK(X && params && Z, "null tensor argument");
    if (init_actnorm) {  // data-dependent initialisation: never replayed
      Ctx c = call_ctx(p, stream);
      drive_forward(p, c, batch, X, nullptr, params, Z, nullptr, logdet, init_actnorm);
      return;
    }
    run_graphed(p, 0, key_of(p, batch, {X, Z, logdet}, params, nullptr), stream,
                [&](Ctx& c) { drive_forward(p, c, batch, X, nullptr, params, Z, nullptr, logdet, 0); });
  });
}
int inb_glow_inverse(inb_plan* p, int batch, const float* Z, float* const* params, float* X, void* stream) {
  return guarded([&] {
    check_call(p, batch, false);
    INB_CHECK(X && params && Z, "null tensor argument");
    run_graphed(p, 1, key_of(p, batch, {Z, X}, params, nullptr), stream, [&](Ctx& c) {
      drive_reverse(p, c, batch, false, nullptr, Z, nullptr, params, nullptr, nullptr, X, nullptr);
    });
  });
}
int inb_glow_backward(inb_plan* p, int batch, const float* dZ, const float* Z, float* const* params,
                      float* const* grads, float* dX, float* X, void* stream) {
  return guarded([&] {
    check_call(p, batch, false);
    INB_CHECK(dZ && Z && params && grads && dX && X, "null tensor argument");
    run_graphed(p, 2, key_of(p, batch, {dZ, Z, dX, X}, params, grads), stream, [&](Ctx& c) {
      drive_reverse(p, c, batch, true, dZ, Z, nullptr, params, grads, dX, X, nullptr);
    });
  });
}
int inb_cglow_forward(inb_plan* p, int batch, const float* X, const float* C, float* const* params,
                      float* ZX, float* ZC, float* logdet, int init_actnorm, void* stream) {
  return guarded([&] {
    check_call(p, batch, true);
    INB_CHECK(X && C && params && ZX && ZC, "null tensor argument");
    if (init_actnorm) {
      Ctx c = call_ctx(p, stream);
      drive_forward(p, c, batch, X, C, params, ZX, ZC, logdet, init_actnorm);
      return;
    }
    run_graphed(p, 0, key_of(p, batch, {X, C, ZX, ZC, logdet}, params, nullptr), stream,
                [&](Ctx& c) { drive_forward(p, c, batch, X, C, params, ZX, ZC, logdet, 0); });
  });
}
int inb_cglow_inverse(inb_plan* p, int batch, const float* ZX, const float* ZC, float* const* params,
                      float* X, void* stream) {
  return guarded([&] {
    check_call(p, batch, true);
    INB_CHECK(ZX && ZC && params && X, "null tensor argument");
    run_graphed(p, 1, key_of(p, batch, {ZX, ZC, X}, params, nullptr), stream, [&](Ctx& c) {
      drive_reverse(p, c, batch, false, nullptr, ZX, ZC, params, nullptr, nullptr, X, nullptr);
    });
  });
}
int inb_cglow_backward(inb_plan* p, int batch, const float* dZX, const float* ZX, const float* ZC,
                       float* const* params, float* const* grads, float* dX, float* X, float* dC,
                       void* stream) {
  return guarded([&] {
    check_call(p, batch, true);
    INB_CHECK(dZX && ZX && ZC && params && grads && dX && X && dC, "null tensor argument");
    run_graphed(p, 2, key_of(p, batch, {dZX, ZX, ZC, dX, X, dC}, params, grads), stream, [&](Ctx& c) {
      drive_reverse(p, c, batch, true, dZX, ZX, ZC, params, grads, dX, X, dC);
    });
  });
}

// ====================================================================== layer level
int inb_actnorm_init(int B, int C, long long sp, const float* X, float* s, float* b, void* stream) {
  return guarded([&] {
    INB_CHECK(X && s && b && B > 0 && C > 0 && sp > 0, "bad argument");
    with_temp_arena((cudaStream_t)stream, 0, [&](Ctx& c) {
      op_actnorm_init(c, sp, B, C, view(const_cast<float*>(X), C * sp), s, b);
    });
  });
}
int inb_actnorm_forward(int B, int C, long long sp, const float* X, const float* s, const float* b, float* Y,
                        float* logdet, void* stream) {
  return guarded([&] {
    INB_CHECK(X && s && b && Y && B > 0 && C > 0 && sp > 0, "bad argument");
    with_temp_arena((cudaStream_t)stream, 0, [&](Ctx& c) {
      double* ld = logdet ? c.ar->f64(1) : nullptr;
      if (ld) op_zero(c, ld, sizeof(double));
      op_an_hh_fwd(c, sp, B, C, view(const_cast<float*>(X), C * sp), view(Y, C * sp), s, b, nullptr, nullptr,
                   nullptr, ld);
      if (ld) op_ld_finish(c, ld, logdet);
    });
  });
}
int inb_actnorm_inverse(int B, int C, long long sp, const float* Y, const float* s, const float* b, float* X,
                        void* stream) {
  return guarded([&] {
    INB_CHECK(X && s && b && Y && B > 0 && C > 0 && sp > 0, "bad argument");
    with_temp_arena((cudaStream_t)stream, 0, [&](Ctx& c) {
      op_hh_an_inv(c, sp, B, C, view(const_cast<float*>(Y), C * sp), view(X, C * sp), s, b, nullptr, nullptr, nullptr);
    });
  });
}
int inb_actnorm_backward(int B, int C, long long sp, const float* dY, const float* Y, const float* s,
                         const float* b, int logdet, float* dX, float* X, float* ds, float* db, void* stream) {
  return guarded([&] {
    INB_CHECK(dY && Y && s && b && dX && X && ds && db && B > 0 && C > 0 && sp > 0, "bad argument");
    with_temp_arena((cudaStream_t)stream, 0, [&](Ctx& c) {
      double* dsdb = c.ar->f64(2 * (size_t)C);
      op_zero(c, dsdb, 2 * C * sizeof(double));
      op_hh_an_bwd(c, sp, B, C, view(const_cast<float*>(dY), C * sp), view(const_cast<float*>(Y), C * sp),
                   view(dX, C * sp), view(X, C * sp), s, b, nullptr, nullptr, nullptr, nullptr, dsdb);
      op_an_grad_finish(c, C, sp, dsdb, s, logdet, ds, db);
    });
  });
}

int inb_conv1x1_forward(int B, int C, long long sp, const float* X, const float* v1, const float* v2,
                        const float* v3, float* Y, void* stream) {
  return guarded([&] {
    INB_CHECK(X && v1 && v2 && v3 && Y && B > 0 && C > 0 && sp > 0, "bad argument");
    with_temp_arena((cudaStream_t)stream, 0, [&](Ctx& c) {
      op_an_hh_fwd(c, sp, B, C, view(const_cast<float*>(X), C * sp), view(Y, C * sp), nullptr, nullptr, v1, v2, v3, nullptr);
    });
  });
}
int inb_conv1x1_inverse(int B, int C, long long sp, const float* Y, const float* v1, const float* v2,
                        const float* v3, float* X, void* stream) {
  return guarded([&] {
    INB_CHECK(X && v1 && v2 && v3 && Y && B > 0 && C > 0 && sp > 0, "bad argument");
    with_temp_arena((cudaStream_t)stream, 0, [&](Ctx& c) {
      op_hh_an_inv(c, sp, B, C, view(const_cast<float*>(Y), C * sp), view(X, C * sp), nullptr, nullptr, v1, v2, v3);
    });
  });
}
int inb_conv1x1_backward(int B, int C, long long sp, const float* dY, const float* Y, const float* v1,
                         const float* v2, const float* v3, int freeze, float* dX, float* X, float* dv1,
                         float* dv2, float* dv3, void* stream) {
  return guarded([&] {
    INB_CHECK(dY && Y && v1 && v2 && v3 && dX && X && dv1 && dv2 && dv3 && B > 0 && C > 0 && sp > 0, "bad argument");
    with_temp_arena((cudaStream_t)stream, 0, [&](Ctx& c) {
      double* gram = c.ar->f64((size_t)C * C);
      op_zero(c, gram, (size_t)C * C * sizeof(double));
      op_hh_an_bwd(c, sp, B, C, view(const_cast<float*>(dY), C * sp), view(const_cast<float*>(Y), C * sp),
                   view(dX, C * sp), view(X, C * sp), nullptr, nullptr, v1, v2, v3, gram, nullptr);
      op_hh_grad_finish(c, C, gram, v1, v2, v3, freeze, dv1, dv2, dv3);
    });
  });
}

static RBShape rb_shape_of(int ndims, int nx, int ny, int nz, int B, int Cin, int nh, int Cout, int k1, int k2) {
  INB_CHECK(ndims == 2 || ndims == 3, "ndims must be 2 or 3");
  INB_CHECK(B > 0 && Cin > 0 && nh > 0 && Cout > 0, "bad shape");
  INB_CHECK((k1 == 1 || k1 == 3) && (k2 == 1 || k2 == 3), "supported kernel sizes are 1 and 3");
  return RBShape{make_geo(ndims, nx, ny, nz), B, Cin, 0, nh, Cout, k1, k2};
}

int inb_resblock_forward(int ndims, int nx, int ny, int nz, int B, int Cin, int nh, int Cout, int k1, int k2,
                         int precision, const float* X, const float* W1, const float* W2, const float* W3,
                         const float* b1, const float* b2, float* Y, void* stream) {
  return guarded([&] {
    INB_CHECK(X && W1 && W2 && W3 && b1 && b2 && Y, "null argument");
    RBShape s = rb_shape_of(ndims, nx, ny, nz, B, Cin, nh, Cout, k1, k2);
    with_temp_arena((cudaStream_t)stream, precision, [&](Ctx& c) {
      RBHidden h{c.ar->f32(rb_hidden_elems(s)), c.ar->f32(rb_hidden_elems(s)), nullptr};
      float* Y3 = c.ar->f32((size_t)B * Cout * s.g.px);
      rb_forward(c, s, view(const_cast<float*>(X), Cin * s.g.px), view(nullptr, 0),
                 RBParams{W1, W2, W3, b1, b2}, h, Y3);
      op_relu_copy(c, (long long)B * Cout * s.g.px, Y3, Y);  // layer_residual_block.jl:133
    });
  });
}
int inb_resblock_backward(int ndims, int nx, int ny, int nz, int B, int Cin, int nh, int Cout, int k1, int k2,
                          int precision, const float* dY, const float* X, const float* W1, const float* W2,
                          const float* W3, const float* b1, const float* b2, float* dX, float* dW1,
                          float* dW2, float* dW3, float* db1, float* db2, void* stream) {
  return guarded([&] {
    INB_CHECK(dY && X && W1 && W2 && W3 && b1 && b2 && dX && dW1 && dW2 && dW3 && db1 && db2, "null argument");
    RBShape s = rb_shape_of(ndims, nx, ny, nz, B, Cin, nh, Cout, k1, k2);
    with_temp_arena((cudaStream_t)stream, precision, [&](Ctx& c) {
      RBHidden h{c.ar->f32(rb_hidden_elems(s)), c.ar->f32(rb_hidden_elems(s)), c.ar->f32(rb_hidden_elems(s))};
      float* Y3 = c.ar->f32((size_t)B * Cout * s.g.px);
      View x = view(const_cast<float*>(X), Cin * s.g.px);
      RBParams p{W1, W2, W3, b1, b2};
      rb_forward(c, s, x, view(nullptr, 0), p, h, Y3);                       // :143
      op_relu_grad(c, (long long)B * Cout * s.g.px, dY, Y3, Y3);             // :150
      rb_backward(c, s, Y3, x, view(nullptr, 0), p, h, RBGrads{dW1, dW2, dW3, db1, db2},
                  view(dX, Cin * s.g.px), nullptr, 0, view(nullptr, 0));
    });
  });
}

static FlowShape coupling_shape(int ndims, int nx, int ny, int nz, int B, int C, int n_cond, int nh, int k1,
                                int k2, float low, float high, int logdet, int freeze) {
  INB_CHECK(ndims == 2 || ndims == 3, "ndims must be 2 or 3");
  INB_CHECK(B > 0 && C >= 2 && nh > 0 && n_cond >= 0, "bad shape");
  FlowShape f{};
  f.g = make_geo(ndims, nx, ny, nz);
  f.B = B; f.C = C; f.ccond = n_cond; f.nh = nh; f.k1 = k1; f.k2 = k2;
  f.low = low; f.high = high; f.logdet = logdet; f.freeze = freeze;
  return f;
}
static FlowParams coupling_params(float* const* cp) {
  FlowParams p{};
  p.v1 = cp[0]; p.v2 = cp[1]; p.v3 = cp[2];
  p.rb = RBParams{cp[3], cp[4], cp[5], cp[6], cp[7]};
  return p;
}

int inb_coupling_forward(int ndims, int nx, int ny, int nz, int B, int C, int n_cond, int nh, int k1, int k2,
                         float low, float high, int precision, const float* X, const float* Cond,
                         float* const* cparams, float* Y, float* logdet, void* stream) {
  return guarded([&] {
    INB_CHECK(X && Y && cparams && (n_cond == 0 || Cond), "null argument");
    FlowShape f = coupling_shape(ndims, nx, ny, nz, B, C, n_cond, nh, k1, k2, low, high, logdet != nullptr, 0);
    with_temp_arena((cudaStream_t)stream, precision, [&](Ctx& c) {
      double* ld = logdet ? c.ar->f64(1) : nullptr;
      if (ld) op_zero(c, ld, sizeof(double));
      flow_forward(c, f, view(const_cast<float*>(X), C * f.g.px), view(Y, C * f.g.px),
                   view(const_cast<float*>(Cond), n_cond * f.g.px), coupling_params(cparams), ld);
      if (ld) op_ld_finish(c, ld, logdet);
    });
  });
}
int inb_coupling_inverse(int ndims, int nx, int ny, int nz, int B, int C, int n_cond, int nh, int k1, int k2,
                         float low, float high, int precision, const float* Y, const float* Cond,
                         float* const* cparams, float* X, void* stream) {
  return guarded([&] {
    INB_CHECK(X && Y && cparams && (n_cond == 0 || Cond), "null argument");
    FlowShape f = coupling_shape(ndims, nx, ny, nz, B, C, n_cond, nh, k1, k2, low, high, 0, 0);
    with_temp_arena((cudaStream_t)stream, precision, [&](Ctx& c) {
      float* tmp = c.ar->f32((size_t)B * C * f.g.px);
      View t = view(tmp, C * f.g.px);
      op_copy(c, f.g.px, B, C, view(const_cast<float*>(Y), C * f.g.px), t);
      flow_inverse(c, f, t, view(X, C * f.g.px), view(const_cast<float*>(Cond), n_cond * f.g.px),
                   coupling_params(cparams));
    });
  });
}
int inb_coupling_backward(int ndims, int nx, int ny, int nz, int B, int C, int n_cond, int nh, int k1, int k2,
                          float low, float high, int logdet, int freeze, int precision, const float* dY,
                          const float* Y, const float* Cond, float* const* cparams, float* const* cgrads,
                          float* dX, float* X, float* dCond, void* stream) {
  return guarded([&] {
    INB_CHECK(dY && Y && dX && X && cparams && cgrads && (n_cond == 0 || (Cond && dCond)), "null argument");
    FlowShape f = coupling_shape(ndims, nx, ny, nz, B, C, n_cond, nh, k1, k2, low, high, logdet, freeze);
    with_temp_arena((cudaStream_t)stream, precision, [&](Ctx& c) {
      const long long bs = C * f.g.px;
      View x = view(X, bs), dx = view(dX, bs);
      op_copy(c, f.g.px, B, C, view(const_cast<float*>(Y), bs), x);
      op_copy(c, f.g.px, B, C, view(const_cast<float*>(dY), bs), dx);
      View dcond = view(dCond, n_cond * f.g.px);
      if (n_cond) op_zero(c, dCond, (size_t)B * n_cond * f.g.px * sizeof(float));
      FlowGrads g{};
      g.v1 = cgrads[0]; g.v2 = cgrads[1]; g.v3 = cgrads[2];
      g.rb = RBGrads{cgrads[3], cgrads[4], cgrads[5], cgrads[6], cgrads[7]};
      flow_backward(c, f, dx, x, dx, x, view(const_cast<float*>(Cond), n_cond * f.g.px), dcond,
                    coupling_params(cparams), g);
    });
  });
}

int inb_squeeze(int ndims, int nx, int ny, int nz, int B, int C, const float* X, float* Y, void* stream) {
  return guarded([&] {
    INB_CHECK(X && Y && B > 0 && C > 0, "bad argument");
    INB_CHECK(ndims == 2 || ndims == 3, "ndims must be 2 or 3");
    Geo g = make_geo(ndims, nx, ny, nz);
    Arena a;
    a.dry = false;
    Ctx c{(cudaStream_t)stream, &a, 0};
    op_squeeze(c, g, B, C, view(const_cast<float*>(X), C * g.px), view(Y, C * g.px));
  });
}
int inb_unsqueeze(int ndims, int nx, int ny, int nz, int B, int C, const float* Y, float* X, void* stream) {
  return guarded([&] {
    // (nx,ny,nz) and C describe the squeezed tensor Y (dimensionality_operations.jl:137-166)
    INB_CHECK(X && Y && B > 0 && C > 0, "bad argument");
    INB_CHECK(ndims == 2 || ndims == 3, "ndims must be 2 or 3");
    INB_CHECK(C % (1 << ndims) == 0, "number of channels must be divisible by %d", 1 << ndims);  // :141-143
    Geo gs = make_geo(ndims, nx, ny, nz);
    Geo g = make_geo(ndims, nx * 2, ny * 2, nz * 2);
    Arena a;
    Ctx c{(cudaStream_t)stream, &a, 0};
    op_unsqueeze(c, g, B, C >> ndims, view(const_cast<float*>(Y), C * gs.px), view(X, C * gs.px));
  });
}

int inb_nll_grad(long long n, int B, const float* Z, float* dZ, float* loss, void* stream) {
  return guarded([&] {
    INB_CHECK(Z && n > 0 && B > 0, "bad argument");
    with_temp_arena((cudaStream_t)stream, 0, [&](Ctx& c) {
      double* acc = loss ? c.ar->f64(1) : nullptr;
      op_nll_grad(c, n, B, Z, dZ, acc, loss);
    });
  });
}

int inb_adam_update(long long n, float* params, const float* grads, float* m, float* v, float lr, float beta1, float beta2,
                    float eps, float beta1_pow_t, float beta2_pow_t, void* stream) {
  return guarded([&] {
    INB_CHECK(params && grads && m && v && n > 0, "bad argument");
    INB_CHECK(beta1_pow_t < 1.f && beta2_pow_t < 1.f, "beta^t must be < 1 (t >= 1)");
    Arena none;
    Ctx c{(cudaStream_t)stream, &none, 0};
    op_adam(c, n, params, grads, m, v, lr, beta1, beta2, eps, beta1_pow_t, beta2_pow_t);
  });
}

// ====================================================================== HINT family
int inb_haar_squeeze(int nx, int ny, int B, int C, int type, const float* X, float* Y, void* stream) {
  return guarded([&] {
    INB_CHECK(X && Y && B > 0 && C > 0, "bad argument");
    Geo g = make_geo(2, nx, ny, 1);
    Arena a;
    Ctx c{(cudaStream_t)stream, &a, 0};
    op_haar_squeeze(c, g, B, C, type, view(const_cast<float*>(X), C * g.px), view(Y, C * g.px));
  });
}
int inb_haar_unsqueeze(int nx, int ny, int B, int C, int type, const float* Y, float* X, void* stream) {
  return guarded([&] {
    INB_CHECK(X && Y && B > 0 && C > 0, "bad argument");
    INB_CHECK(C % 4 == 0, "number of channels must be divisible by 4");
    Geo g = make_geo(2, nx * 2, ny * 2, 1);
    Arena a;
    Ctx c{(cudaStream_t)stream, &a, 0};
    op_haar_unsqueeze(c, g, B, C / 4, type, view(const_cast<float*>(Y), (C / 4) * g.px), view(X, (C / 4) * g.px));
  });
}

static HintShape hint_shape(int ndims, int nx, int ny, int nz, int B, int C, int nh, int k1, int k2, float low,
                            float high, int permute, int logdet, int shared) {
  INB_CHECK(ndims == 2 || ndims == 3, "ndims must be 2 or 3");
  INB_CHECK(B > 0 && nh > 0, "bad shape");
  INB_CHECK((k1 == 1 || k1 == 3) && (k2 == 1 || k2 == 3), "supported kernel sizes are 1 and 3");
  INB_CHECK(high > low, "sigmoid high must exceed low");
  HintShape h{};
  h.g = make_geo(ndims, nx, ny, nz);
  h.B = B; h.C = C; h.nh = nh; h.k1 = k1; h.k2 = k2; h.low = low; h.high = high;
  h.logdet = logdet; h.permute = permute; h.shared_last = shared;
  hint_check(h);
  return h;
}
// table in the layer's get_params order -> HintParams / HintGrads
static HintParams hint_params(const HintShape& h, float* const* t) {
  HintParams p;
  const int n = h.depth();
  p.cl.resize(n);
  if (!t) return p;
  for (int j = 0; j < n; ++j) p.cl[j] = RBParams{t[5 * j], t[5 * j + 1], t[5 * j + 2], t[5 * j + 3], t[5 * j + 4]};
  if (h.permute != HINT_PERMUTE_NONE) { p.v1 = t[5 * n]; p.v2 = t[5 * n + 1]; p.v3 = t[5 * n + 2]; }
  return p;
}
static HintGrads hint_grads(const HintShape& h, float* const* t) {
  HintGrads g;
  const int n = h.depth();
  g.cl.resize(n);
  if (!t) return g;
  for (int j = 0; j < n; ++j) g.cl[j] = RBGrads{t[5 * j], t[5 * j + 1], t[5 * j + 2], t[5 * j + 3], t[5 * j + 4]};
  if (h.permute != HINT_PERMUTE_NONE) { g.v1 = t[5 * n]; g.v2 = t[5 * n + 1]; g.v3 = t[5 * n + 2]; }
  return g;
}

static HintShape basic_shape(int ndims, int nx, int ny, int nz, int B, int C1, int nh, int k1, int k2, float low,
                             float high, int logdet) {
  INB_CHECK(ndims == 2 || ndims == 3, "ndims must be 2 or 3");
  INB_CHECK(B > 0 && nh > 0 && C1 > 0, "bad shape");
  INB_CHECK((k1 == 1 || k1 == 3) && (k2 == 1 || k2 == 3), "supported kernel sizes are 1 and 3");
  INB_CHECK(high > low, "sigmoid high must exceed low");
  HintShape h{};
  h.g = make_geo(ndims, nx, ny, nz);
  h.B = B; h.C = 2 * C1; h.nh = nh; h.k1 = k1; h.k2 = k2; h.low = low; h.high = high; h.logdet = logdet;
  return h;
}
int inb_basic_coupling_forward(int ndims, int nx, int ny, int nz, int B, int C1, int nh, int k1, int k2, float low,
                               float high, int precision, const float* X1, const float* X2, float* const* rbparams,
                               float* Y2, float* logdet, void* stream) {
  return guarded([&] {
    INB_CHECK(X1 && X2 && Y2 && rbparams, "null argument");
    HintShape h = basic_shape(ndims, nx, ny, nz, B, C1, nh, k1, k2, low, high, logdet != nullptr);
    RBParams p{rbparams[0], rbparams[1], rbparams[2], rbparams[3], rbparams[4]};
    with_temp_arena((cudaStream_t)stream, precision, [&](Ctx& c) {
      const long long bs = C1 * h.g.px;
      double* ld = logdet ? c.ar->f64(1) : nullptr;
      if (ld) op_zero(c, ld, sizeof(double));
      op_copy(c, h.g.px, B, C1, view(const_cast<float*>(X2), bs), view(Y2, bs));
      basic_forward(c, h, C1, view(const_cast<float*>(X1), bs), view(Y2, bs), p, ld);
      if (ld) op_ld_finish(c, ld, logdet);
    });
  });
}
int inb_basic_coupling_inverse(int ndims, int nx, int ny, int nz, int B, int C1, int nh, int k1, int k2, float low,
                               float high, int precision, const float* Y1, const float* Y2, float* const* rbparams,
                               float* X2, void* stream) {
  return guarded([&] {
    INB_CHECK(Y1 && Y2 && X2 && rbparams, "null argument");
    HintShape h = basic_shape(ndims, nx, ny, nz, B, C1, nh, k1, k2, low, high, 0);
    RBParams p{rbparams[0], rbparams[1], rbparams[2], rbparams[3], rbparams[4]};
    with_temp_arena((cudaStream_t)stream, precision, [&](Ctx& c) {
      const long long bs = C1 * h.g.px;
      op_copy(c, h.g.px, B, C1, view(const_cast<float*>(Y2), bs), view(X2, bs));
      basic_inverse(c, h, C1, view(const_cast<float*>(Y1), bs), view(X2, bs), p);
    });
  });
}
int inb_basic_coupling_backward(int ndims, int nx, int ny, int nz, int B, int C1, int nh, int k1, int k2, float low,
                                float high, int logdet, int precision, const float* dY1, const float* dY2,
                                const float* Y1, const float* Y2, float* const* rbparams, float* const* rbgrads,
                                float* dX1, float* dX2, float* X2, void* stream) {
  return guarded([&] {
    INB_CHECK(dY1 && dY2 && Y1 && Y2 && rbparams && rbgrads && dX1 && dX2 && X2, "null argument");
    HintShape h = basic_shape(ndims, nx, ny, nz, B, C1, nh, k1, k2, low, high, logdet);
    RBParams p{rbparams[0], rbparams[1], rbparams[2], rbparams[3], rbparams[4]};
    RBGrads g{rbgrads[0], rbgrads[1], rbgrads[2], rbgrads[3], rbgrads[4]};
    with_temp_arena((cudaStream_t)stream, precision, [&](Ctx& c) {
      const long long bs = C1 * h.g.px;
      op_copy(c, h.g.px, B, C1, view(const_cast<float*>(Y2), bs), view(X2, bs));
      op_copy(c, h.g.px, B, C1, view(const_cast<float*>(dY2), bs), view(dX2, bs));
      op_copy(c, h.g.px, B, C1, view(const_cast<float*>(dY1), bs), view(dX1, bs));  // + dY1 of basic.jl:137
      basic_backward(c, h, C1, view(const_cast<float*>(Y1), bs), view(dX1, bs), view(X2, bs), view(dX2, bs), p, g, false);
    });
  });
}

int inb_hint_depth(int C) {
  HintShape h{};
  h.C = C;
  return h.depth();
}
int inb_hint_coupling_forward(int ndims, int nx, int ny, int nz, int B, int C, int nh, int k1, int k2, float low,
                              float high, int permute, int precision, const float* X, float* const* hparams,
                              float* Y, float* logdet, void* stream) {
  return guarded([&] {
    INB_CHECK(X && Y && hparams, "null argument");
    HintShape h = hint_shape(ndims, nx, ny, nz, B, C, nh, k1, k2, low, high, permute, logdet != nullptr, 0);
    HintParams p = hint_params(h, hparams);
    with_temp_arena((cudaStream_t)stream, precision, [&](Ctx& c) {
      double* ld = logdet ? c.ar->f64(1) : nullptr;
      if (ld) op_zero(c, ld, sizeof(double));
      hint_forward(c, h, view(const_cast<float*>(X), C * h.g.px), view(Y, C * h.g.px), p, ld);
      if (ld) op_ld_finish(c, ld, logdet);
    });
  });
}
int inb_hint_coupling_inverse(int ndims, int nx, int ny, int nz, int B, int C, int nh, int k1, int k2, float low,
                              float high, int permute, int precision, const float* Y, float* const* hparams,
                              float* X, void* stream) {
  return guarded([&] {
    INB_CHECK(X && Y && hparams, "null argument");
    HintShape h = hint_shape(ndims, nx, ny, nz, B, C, nh, k1, k2, low, high, permute, 0, 0);
    HintParams p = hint_params(h, hparams);
    with_temp_arena((cudaStream_t)stream, precision, [&](Ctx& c) {
      View x = view(X, C * h.g.px);
      op_copy(c, h.g.px, B, C, view(const_cast<float*>(Y), C * h.g.px), x);
      hint_inverse(c, h, x, x, p);
    });
  });
}
int inb_hint_coupling_backward(int ndims, int nx, int ny, int nz, int B, int C, int nh, int k1, int k2, float low,
                               float high, int permute, int logdet, int shared_grads, int precision,
                               const float* dY, const float* Y, float* const* hparams, float* const* hgrads,
                               float* dX, float* X, void* stream) {
  return guarded([&] {
    INB_CHECK(dY && Y && dX && X && hparams && hgrads, "null argument");
    HintShape h = hint_shape(ndims, nx, ny, nz, B, C, nh, k1, k2, low, high, permute, logdet, shared_grads);
    HintParams p = hint_params(h, hparams);
    HintGrads g = hint_grads(h, hgrads);
    with_temp_arena((cudaStream_t)stream, precision, [&](Ctx& c) {
      const long long bs = C * h.g.px;
      View x = view(X, bs), dx = view(dX, bs);
      op_copy(c, h.g.px, B, C, view(const_cast<float*>(Y), bs), x);
      op_copy(c, h.g.px, B, C, view(const_cast<float*>(dY), bs), dx);
      hint_backward(c, h, dx, x, dx, x, p, g);
    });
  });
}

// ---------------------------------------------------------------- NetworkMultiScaleHINT
struct HintScale {
  Geo g;
  int C;     // channels of the flow steps
  int kc;    // channels that continue to the next scale (== C when nothing is split off)
  int depth;
  int pbase; // index of CL[i,1]'s first parameter
};
struct inb_hint_plan : GraphCache {
  inb_hint_desc d;
  Geo g0;
  std::vector<HintScale> sc;
  int nparams = 0;
  Arena ar;
  size_t need = 0, persist = 0;
  double* ld = nullptr;
};

static HintShape hplan_shape(const inb_hint_plan* p, int i, int B) {
  const HintScale& s = p->sc[i];
  HintShape h{};
  h.g = s.g; h.B = B; h.C = s.C; h.nh = p->d.n_hidden; h.k1 = p->d.k1; h.k2 = p->d.k2;
  h.low = p->d.sig_low; h.high = p->d.sig_high; h.logdet = 1; h.permute = HINT_PERMUTE_FULL;
  h.shared_last = p->d.shared_grads;
  return h;
}
static int hplan_cl_base(const inb_hint_plan* p, int i, int j) { return p->sc[i].pbase + j * (5 * p->sc[i].depth + 3); }
static HintParams hplan_params(const inb_hint_plan* p, int i, int j, float* const* t) {
  HintShape h = hplan_shape(p, i, 1);
  HintParams hp = hint_params(h, t ? t + hplan_cl_base(p, i, j) : nullptr);
  if (t) { hp.s = t[2 * (i * p->d.K + j)]; hp.b = t[2 * (i * p->d.K + j) + 1]; }
  return hp;
}
static HintGrads hplan_grads(const inb_hint_plan* p, int i, int j, float* const* t) {
  HintShape h = hplan_shape(p, i, 1);
  HintGrads hg = hint_grads(h, t ? t + hplan_cl_base(p, i, j) : nullptr);
  if (t) { hg.s = t[2 * (i * p->d.K + j)]; hg.b = t[2 * (i * p->d.K + j) + 1]; }
  return hg;
}
static long long hplan_zoff(const inb_hint_plan* p, int B, int upto) {
  long long off = 0;
  for (int i = 0; i < upto && i < (int)p->sc.size(); ++i) off += (long long)B * (p->sc[i].C - p->sc[i].kc) * p->sc[i].g.px;
  return off;
}

static void hint_drive_forward(inb_hint_plan* p, Ctx& c, int B, const float* X, float* const* prm, float* Z,
                               float* logdet, int init) {
  const inb_hint_desc& d = p->d;
  const long long tot = (long long)d.n_in * p->g0.px;
  size_t m = c.ar->mark();
  float* buf[2] = {c.ar->f32((size_t)B * tot), c.ar->f32((size_t)B * tot)};
  op_zero(c, p->ld, sizeof(double));
  View cur = view(const_cast<float*>(X), tot);
  int which = 0, chan = d.n_in;
  Geo g = p->g0;
  for (int i = 0; i < d.L; ++i) {
    const HintScale& s = p->sc[i];
    const long long bs = (long long)s.C * s.g.px;
    {
      View out = view(buf[which], bs);
      op_haar_squeeze(c, g, B, chan, d.squeeze_type, cur, out);  // hint_multiscale.jl:102
      cur = out;
      which ^= 1;
    }
    const HintShape h = hplan_shape(p, i, B);
    for (int j = 0; j < d.K; ++j) {
      HintParams hp = hplan_params(p, i, j, prm);
      if (init) op_actnorm_init(c, s.g.px, B, s.C, cur, const_cast<float*>(hp.s), const_cast<float*>(hp.b));
      View out = view(buf[which], bs);
      hint_forward(c, h, cur, out, hp, p->ld);  // :104-106
      cur = out;
      which ^= 1;
    }
    if (s.kc != s.C)  // :108-112: X = first part, Z = second part
      op_copy(c, s.g.px, B, s.C - s.kc, sub(cur, s.kc, s.g.px), view(Z + hplan_zoff(p, B, i), (long long)(s.C - s.kc) * s.g.px));
    chan = s.kc;
    g = s.g;
  }
  {
    const HintScale& s = p->sc[d.L - 1];
    op_copy(c, s.g.px, B, chan, cur, view(Z + hplan_zoff(p, B, d.L), (long long)chan * s.g.px));  // :114
  }
  if (logdet) op_ld_finish(c, p->ld, logdet);
  c.ar->release(m);
}

static void hint_drive_reverse(inb_hint_plan* p, Ctx& c, int B, bool grads, const float* dZ, const float* Z,
                               float* const* prm, float* const* gr, float* dX, float* X) {
  const inb_hint_desc& d = p->d;
  const long long tot = (long long)d.n_in * p->g0.px;
  size_t m = c.ar->mark();
  float* yb[2] = {c.ar->f32((size_t)B * tot), c.ar->f32((size_t)B * tot)};
  float* db[2] = {nullptr, nullptr};
  if (grads) { db[0] = c.ar->f32((size_t)B * tot); db[1] = c.ar->f32((size_t)B * tot); }
  int which = 0;
  {
    const HintScale& sl = p->sc[d.L - 1];
    const long long off = hplan_zoff(p, B, d.L), bs = (long long)sl.C * sl.g.px;
    op_copy(c, sl.g.px, B, sl.C, view(const_cast<float*>(Z) + off, bs), view(yb[which], bs));
    if (grads) op_copy(c, sl.g.px, B, sl.C, view(const_cast<float*>(dZ) + off, bs), view(db[which], bs));
  }
  for (int i = d.L - 1; i >= 0; --i) {
    const HintScale& s = p->sc[i];
    const long long bs = (long long)s.C * s.g.px;
    View y = view(yb[which], bs), dy = view(grads ? db[which] : nullptr, bs);
    if (s.kc != s.C) {  // :144-147 tensor_cat with the saved latent
      const long long off = hplan_zoff(p, B, i);
      const int zc = s.C - s.kc;
      op_copy(c, s.g.px, B, zc, view(const_cast<float*>(Z) + off, (long long)zc * s.g.px), sub(y, s.kc, s.g.px));
      if (grads) op_copy(c, s.g.px, B, zc, view(const_cast<float*>(dZ) + off, (long long)zc * s.g.px), sub(dy, s.kc, s.g.px));
    }
    const HintShape h = hplan_shape(p, i, B);
    for (int j = d.K - 1; j >= 0; --j) {
      HintParams hp = hplan_params(p, i, j, prm);
      if (grads) hint_backward(c, h, dy, y, dy, y, hp, hplan_grads(p, i, j, gr));  // :150-152
      else hint_inverse(c, h, y, y, hp);                                          // :127-130
    }
    // wavelet_unsqueeze (:131, :163-164) into the first channels of the coarser scale's buffer / the caller's output
    const Geo gout = (i == 0) ? p->g0 : p->sc[i - 1].g;
    const long long obs = (i == 0) ? tot : (long long)p->sc[i - 1].C * p->sc[i - 1].g.px;
    op_haar_unsqueeze(c, gout, B, s.C / 4, d.squeeze_type, y, view(i == 0 ? X : yb[which ^ 1], obs));
    if (grads) op_haar_unsqueeze(c, gout, B, s.C / 4, d.squeeze_type, dy, view(i == 0 ? dX : db[which ^ 1], obs));
    which ^= 1;
  }
  c.ar->release(m);
}

static Ctx call_ctx(inb_hint_plan* p, void* stream) {
  if (!p->ar.base) {
    void* base = nullptr;
    cudaError_t e = cudaMalloc(&base, p->need);
    if (e != cudaSuccess) fail(INB_ERR_NOMEM, "workspace of %zu bytes: %s", p->need, cudaGetErrorString(e));
    p->ar.base = (char*)base;
    p->ar.cap = p->need;
    p->ar.dry = false;
    p->ar.off = 0;
    p->ld = (double*)p->ar.alloc_bytes(256);
    p->persist = p->ar.off;
  }
  p->ar.off = p->persist;
  return Ctx{(cudaStream_t)stream, &p->ar, p->d.precision};
}
static GraphKey hint_key(const inb_hint_plan* p, int batch, std::initializer_list<const void*> ptrs,
                         float* const* params, float* const* grads) {
  GraphKey k;
  k.add(0x48494e54ull);
  k.add((uint64_t)batch);
  for (const void* q : ptrs) k.add((uint64_t)(uintptr_t)q);
  for (int i = 0; i < p->nparams; ++i) k.add((uint64_t)(uintptr_t)params[i]);
  if (grads)
    for (int i = 0; i < p->nparams; ++i) k.add((uint64_t)(uintptr_t)grads[i]);
  k.a |= 1ull;
  return k;
}
static void hint_check_call(inb_hint_plan* p, int batch) {
  INB_CHECK(p != nullptr, "null plan");
  INB_CHECK(batch >= 1 && batch <= p->d.batch, "batch %d outside the plan's range [1, %d]", batch, p->d.batch);
}

int inb_hint_plan_create(const inb_hint_desc* d, inb_hint_plan** out) {
  return guarded([&] {
    INB_CHECK(out != nullptr, "null output");
    *out = nullptr;
    INB_CHECK(d != nullptr, "null descriptor");
    INB_CHECK(d->nx > 0 && d->ny > 0 && d->n_in >= 1 && d->n_hidden >= 1 && d->L >= 1 && d->K >= 1 && d->batch >= 1,
              "nx, ny, n_in, n_hidden, L, K, batch must be >= 1");
    INB_CHECK((d->k1 == 1 || d->k1 == 3) && (d->k2 == 1 || d->k2 == 3),
              "supported ResidualBlock kernel sizes are 1 and 3 (got k1=%d k2=%d)", d->k1, d->k2);
    INB_CHECK(d->p1 == (d->k1 - 1) / 2 && d->p2 == (d->k2 - 1) / 2, "only 'same' padding is supported");
    INB_CHECK(d->precision >= 0 && d->precision <= 3, "unknown precision mode %d", d->precision);
    INB_CHECK(d->sig_high > d->sig_low, "sigmoid high must exceed low");
    INB_CHECK(d->squeeze_type == 0 || d->squeeze_type == 1, "squeeze_type must be 0 (wavelet) or 1 (Haar)");
    std::unique_ptr<inb_hint_plan> p(new inb_hint_plan());
    p->d = *d;
    p->g0 = make_geo(2, d->nx, d->ny, 1);
    Geo g = p->g0;
    int chan = d->n_in, idx = 2 * d->L * d->K;
    for (int i = 0; i < d->L; ++i) {
      INB_CHECK(!(g.W % 2) && !(g.H % 2), "Input dimensions must be multiple of 2");
      HintScale s{};
      s.g = half_geo(g);
      s.C = 4 * chan;  // ctor :87
      s.kc = (d->split_scales && i < d->L - 1) ? s.C / 2 : s.C;
      p->sc.push_back(s);
      HintShape h = hplan_shape(p.get(), i, d->batch);
      hint_check(h);
      p->sc[i].depth = h.depth();
      p->sc[i].pbase = idx;
      idx += d->K * (5 * h.depth() + 3);
      chan = s.kc;
      g = s.g;
    }
    p->nparams = idx;
    Arena dry;
    dry.dry = true;
    dry.alloc_bytes(256);
    Ctx dc{nullptr, &dry, p->d.precision};
    hint_drive_forward(p.get(), dc, d->batch, nullptr, nullptr, nullptr, nullptr, 1);
    hint_drive_reverse(p.get(), dc, d->batch, true, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
    p->need = dry.peak + 1024;
    *out = p.release();
  });
}
int inb_hint_plan_destroy(inb_hint_plan* p) {
  return guarded([&] {
    if (!p) return;
    p->destroy_graphs();
    if (p->ar.base) cudaFree(p->ar.base);
    delete p;
  });
}
int inb_hint_num_params(const inb_hint_plan* p) { return p ? p->nparams : -1; }
long long inb_hint_workspace_bytes(const inb_hint_plan* p) { return p ? (long long)p->need : -1; }
int inb_hint_param_numel(const inb_hint_plan* p, int index, long long* numel) {
  return guarded([&] {
    INB_CHECK(p && numel && index >= 0 && index < p->nparams, "bad argument");
    const inb_hint_desc& d = p->d;
    if (index < 2 * d.L * d.K) {
      *numel = p->sc[index / (2 * d.K)].C;
      return;
    }
    for (int i = d.L - 1; i >= 0; --i) {
      if (index < p->sc[i].pbase) continue;
      const HintScale& s = p->sc[i];
      const int per = 5 * s.depth + 3, w = (index - s.pbase) % per;
      if (w >= 5 * s.depth) { *numel = s.C; return; }
      const int j = w / 5 + 1, c = s.C >> j;  // CL[j] = CouplingLayerBasic(C / 2^j), hint.jl:86
      const long long t1 = (long long)d.k1 * d.k1, t2 = (long long)d.k2 * d.k2, nh = d.n_hidden;
      switch (w % 5) {
        case 0: *numel = nh * c * t1; break;
        case 1: *numel = nh * nh * t2; break;
        case 2: *numel = nh * 2 * c * t1; break;
        default: *numel = nh;
      }
      return;
    }
  });
}
int inb_hint_forward(inb_hint_plan* p, int batch, const float* X, float* const* params, float* Z, float* logdet,
                     int init_actnorm, void* stream) {
  return guarded([&] {
    hint_check_call(p, batch);
    INB_CHECK(X && params && Z, "null tensor argument");
    if (init_actnorm) {  // data-dependent initialisation: never replayed
      Ctx c = call_ctx(p, stream);
      hint_drive_forward(p, c, batch, X, params, Z, logdet, 1);
      return;
    }
    run_graphed(p, 0, hint_key(p, batch, {X, Z, logdet}, params, nullptr), stream,
                [&](Ctx& c) { hint_drive_forward(p, c, batch, X, params, Z, logdet, 0); });
  });
}
int inb_hint_inverse(inb_hint_plan* p, int batch, const float* Z, float* const* params, float* X, void* stream) {
  return guarded([&] {
    hint_check_call(p, batch);
    INB_CHECK(X && params && Z, "null tensor argument");
    run_graphed(p, 1, hint_key(p, batch, {Z, X}, params, nullptr), stream,
                [&](Ctx& c) { hint_drive_reverse(p, c, batch, false, nullptr, Z, params, nullptr, nullptr, X); });
  });
}
int inb_hint_backward(inb_hint_plan* p, int batch, const float* dZ, const float* Z, float* const* params,
                      float* const* grads, float* dX, float* X, void* stream) {
  return guarded([&] {
    hint_check_call(p, batch);
    INB_CHECK(dZ && Z && params && grads && dX && X, "null tensor argument");
    run_graphed(p, 2, hint_key(p, batch, {dZ, Z, dX, X}, params, grads), stream,
                [&](Ctx& c) { hint_drive_reverse(p, c, batch, true, dZ, Z, params, grads, dX, X); });
  });
}

}  // extern "C"
