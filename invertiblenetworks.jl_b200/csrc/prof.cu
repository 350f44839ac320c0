// prof.cu - launch counter and optional CUDA-event profiling of the op families.
#include "../../include/inb200.h"
#include "common.cuh"
#include "conv_tc.cuh"

#include <atomic>
#include <mutex>
#include <cstring>

namespace inb {

static const char* kFamilyNames[F_COUNT] = {
    "squeeze", "copy", "actnorm_stats", "actnorm_hh_fwd", "hh_actnorm_inv", "hh_actnorm_bwd", "grad_finish",
    "coupling_fwd", "coupling_inv", "coupling_bwd", "pack_weights", "conv_simt", "wgrad_simt", "channel_sum",
    "nll_grad", "misc", "conv_tc", "wgrad_tc", "layout_tc", "col2im"};

struct Pending {
  cudaEvent_t a, b;
  int fam;
};
struct FamStat {
  long long launches = 0, scopes = 0;
  double ms = 0, flops = 0, bytes = 0;
};
static std::atomic<long long> g_launches{0};
static std::atomic<int> g_on{0};
static std::mutex g_mu;
static std::vector<Pending> g_pending;
static std::vector<cudaEvent_t> g_pool;
static FamStat g_stat[F_COUNT];

static cudaEvent_t get_event() {
  if (!g_pool.empty()) {
    cudaEvent_t e = g_pool.back();
    g_pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}

Prof::Prof(const Ctx& c, int fam_, int launches, double flops, double bytes) : st(c.st), fam(fam_), slot(nullptr) {
  if (c.ar->dry) { fam = -1; return; }
  g_launches.fetch_add(launches, std::memory_order_relaxed);
  if (!g_on.load(std::memory_order_relaxed)) { fam = -1; return; }
  std::lock_guard<std::mutex> lk(g_mu);
  g_stat[fam].launches += launches;
  g_stat[fam].scopes += 1;
  g_stat[fam].flops += flops;
  g_stat[fam].bytes += bytes;
  Pending p{get_event(), get_event(), fam};
  cudaEventRecord(p.a, st);
  g_pending.push_back(p);
  slot = (void*)(uintptr_t)g_pending.size();  // index + 1
}
Prof::~Prof() {
  if (fam < 0 || !slot) return;
  std::lock_guard<std::mutex> lk(g_mu);
  size_t i = (size_t)(uintptr_t)slot - 1;
  if (i < g_pending.size()) cudaEventRecord(g_pending[i].b, st);
}

static void collect_locked() {
  for (auto& p : g_pending) {
    cudaEventSynchronize(p.b);
    float ms = 0;
    if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) g_stat[p.fam].ms += ms;
    g_pool.push_back(p.a);
    g_pool.push_back(p.b);
  }
  g_pending.clear();
}

}  // namespace inb

namespace inb {
bool prof_is_enabled() { return g_on.load(std::memory_order_relaxed) != 0; }
long long launch_count_now() { return g_launches.load(); }
void launch_count_add(long long n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace inb

using namespace inb;

extern "C" {

long long inb_launch_count(void) { return g_launches.load(); }

int inb_debug_chain_trace(void* dev_buf) {
  inb::chain_set_trace((long long*)dev_buf);
  return 0;
}
int inb_prof_enable(int on) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!on) collect_locked();
  g_on.store(on ? 1 : 0);
  return 0;
}
int inb_prof_reset(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  collect_locked();
  for (auto& s : g_stat) s = FamStat();
  return 0;
}
int inb_prof_num(void) { return F_COUNT; }
int inb_prof_get(int i, char* name, int name_len, long long* launches, long long* scopes, double* ms,
                 double* flops, double* bytes) {
  if (i < 0 || i >= F_COUNT) return 1;
  std::lock_guard<std::mutex> lk(g_mu);
  collect_locked();
  if (name && name_len > 0) {
    strncpy(name, kFamilyNames[i], name_len - 1);
    name[name_len - 1] = 0;
  }
  if (launches) *launches = g_stat[i].launches;
  if (scopes) *scopes = g_stat[i].scopes;
  if (ms) *ms = g_stat[i].ms;
  if (flops) *flops = g_stat[i].flops;
  if (bytes) *bytes = g_stat[i].bytes;
  return 0;
}
}
