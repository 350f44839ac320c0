// elementwise.cu - the HBM-bound kernels of the Glow path: squeeze index maps, ActNorm,
// the Householder 1x1 channel mix, the affine coupling and their reductions.
//
// Layout (B, C, px) fp32, pixel fastest.  The per-pixel kernels give one thread V consecutive
// pixels and all C channels of them in registers: every global access of a warp is a contiguous
// 128*V-byte run per channel (coalesced, vectorised), each element is read once and written once.
#include "ops.cuh"
#include "dp.cuh"
#include <algorithm>

namespace inb {

// ---------------------------------------------------------------- small device helpers
template <int V>
__device__ __forceinline__ void ldv(const float* __restrict__ p, float (&r)[V]) {
  if constexpr (V == 4) {
    float4 t = *reinterpret_cast<const float4*>(p);
    r[0] = t.x; r[1] = t.y; r[2] = t.z; r[3] = t.w;
  } else if constexpr (V == 2) {
    float2 t = *reinterpret_cast<const float2*>(p);
    r[0] = t.x; r[1] = t.y;
  } else {
    r[0] = *p;
  }
}
template <int V>
__device__ __forceinline__ void stv(float* __restrict__ p, const float (&r)[V]) {
  if constexpr (V == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(r[0], r[1], r[2], r[3]);
  } else if constexpr (V == 2) {
    *reinterpret_cast<float2*>(p) = make_float2(r[0], r[1]);
  } else {
    *p = r[0];
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// block-wide sum (blockDim.x multiple of 32, <= 1024); result valid in thread 0
__device__ __forceinline__ double block_sum(double v) {
  __shared__ double red[32];
  v = warp_sum(v);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) red[w] = v;
  __syncthreads();
  double r = 0;
  if (w == 0) {
    r = (l < (int)(blockDim.x >> 5)) ? red[l] : 0.0;
    r = warp_sum(r);
  }
  __syncthreads();
  return r;
}

// activation_functions.jl:160-162: high / (1 + e^-x) + low / (1 + e^x) = low + (high - low) / (1 + e^-x), one
// exponential and one division instead of two of each (the coupling kernels are ALU-bound, not HBM-bound)
__device__ __forceinline__ float sigmoid_lh(float x, float low, float high) {
  return low + (high - low) / (1.f + expf(-x));
}

// ---------------------------------------------------------------- squeeze / unsqueeze / copy
// dimensionality_operations.jl:40-47,97-102: out[b, p*C+c, z',y',x'] = in[b,c,2z'+iz,2y'+iy,2x'+ix],
// p = ix + 2*iy + 4*iz.  One thread moves an x-pair: 8-byte access on the full-resolution side.
template <bool FWD>
__global__ void k_squeeze(const float* __restrict__ src, long long sbs, float* __restrict__ dst,
                          long long dbs, int C, int W, int H, int D, int B) {
  // (W,H,D) is the full-resolution geometry; C the full-resolution channel count
  const int Wh = W >> 1, Hh = H >> 1;
  const long long pxf = (long long)W * H * D;
  const long long pxh = pxf >> ((D > 1) ? 3 : 2);
  const long long total = (long long)B * C * D * H * Wh;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long t = i;
    int xh = (int)(t % Wh); t /= Wh;
    int y = (int)(t % H); t /= H;
    int z = (int)(t % D); t /= D;
    int c = (int)(t % C); t /= C;
    int b = (int)t;
    int iy = y & 1, iz = z & 1;
    int p0 = 2 * iy + 4 * iz;  // ix = 0; ix = 1 is p0 + 1
    long long full = (long long)c * pxf + ((long long)z * H + y) * W + 2 * xh;
    long long half = ((long long)(z >> 1) * Hh + (y >> 1)) * Wh + xh;
    long long h0 = ((long long)p0 * C + c) * pxh + half;
    long long h1 = ((long long)(p0 + 1) * C + c) * pxh + half;
    if (FWD) {
      float2 v = *reinterpret_cast<const float2*>(src + b * sbs + full);
      dst[b * dbs + h0] = v.x;
      dst[b * dbs + h1] = v.y;
    } else {
      float2 v = make_float2(src[b * sbs + h0], src[b * sbs + h1]);
      *reinterpret_cast<float2*>(dst + b * dbs + full) = v;
    }
  }
}

// The same map with four full-resolution x per thread: one 16-byte access on the full-resolution side, two 8-byte
// accesses on the squeezed side, and 32-bit index arithmetic inside a (sample, channel) plane (grid.y walks the planes).
// The two-elements-per-thread kernel above is bound by its 64-bit division chain (1.4 TB/s); this one by HBM.
template <bool FWD>
__global__ void __launch_bounds__(256)
k_squeeze4(const float* __restrict__ src, long long sbs, float* __restrict__ dst, long long dbs, int C, int W, int H,
           int D, int planes) {
  const unsigned Wq = (unsigned)W >> 2, Wh = (unsigned)W >> 1, Hh = (unsigned)H >> 1;
  const unsigned nq = Wq * (unsigned)H * (unsigned)D;  // quads per plane
  const long long pxf = (long long)W * H * D;
  const long long pxh = pxf >> ((D > 1) ? 3 : 2);
  for (int pl = blockIdx.y; pl < planes; pl += gridDim.y) {
    const int b = pl / C, c = pl - b * C;
    const float* fs = (FWD ? src + b * sbs : dst + b * dbs) + (long long)c * pxf;
    const float* hsb = FWD ? dst + b * dbs : src + b * sbs;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < nq; i += gridDim.x * blockDim.x) {
      const unsigned row = i / Wq, xq = i - row * Wq;   // row = z * H + y
      const unsigned z = row / (unsigned)H, y = row - z * (unsigned)H;
      const unsigned p0 = 2 * (y & 1) + 4 * (z & 1);
      const long long half = ((long long)(z >> 1) * Hh + (y >> 1)) * Wh + 2 * xq;
      float* h0 = const_cast<float*>(hsb) + ((long long)p0 * C + c) * pxh + half;
      float* h1 = h0 + (long long)C * pxh;
      float* f = const_cast<float*>(fs) + (long long)row * W + 4 * xq;
      if (FWD) {
        const float4 v = *reinterpret_cast<const float4*>(f);
        *reinterpret_cast<float2*>(h0) = make_float2(v.x, v.z);
        *reinterpret_cast<float2*>(h1) = make_float2(v.y, v.w);
      } else {
        const float2 a = *reinterpret_cast<const float2*>(h0), bq = *reinterpret_cast<const float2*>(h1);
        *reinterpret_cast<float4*>(f) = make_float4(a.x, bq.x, a.y, bq.y);
      }
    }
  }
}
template <bool FWD>
static bool launch_squeeze4(Ctx& c, const Geo& g, int B, int C, View full, View half) {
  const long long pxh = g.px >> (g.nd == 3 ? 3 : 2);
  const bool ok = g.W % 4 == 0 && full.bs % 4 == 0 && ((uintptr_t)full.p & 15) == 0 && half.bs % 2 == 0 &&
                  ((uintptr_t)half.p & 7) == 0 && pxh % 2 == 0 && g.px / 4 < (1ll << 31) && (long long)B * C < (1ll << 31);
  if (!ok) return false;
  const unsigned nq = (unsigned)(g.px / 4);
  const int planes = B * C;
  // ~4 blocks per SM in flight over the whole launch
  unsigned gx = (nq + 255) / 256;
  const unsigned want = (unsigned)std::max<long long>(1, cdiv(148LL * 8, planes));
  if (gx > want) gx = want;
  const dim3 grid(gx, (unsigned)std::min(planes, 65535), 1);
  if (FWD) k_squeeze4<true><<<grid, 256, 0, c.st>>>(full.p, full.bs, half.p, half.bs, C, g.W, g.H, g.D, planes);
  else k_squeeze4<false><<<grid, 256, 0, c.st>>>(half.p, half.bs, full.p, full.bs, C, g.W, g.H, g.D, planes);
  return true;
}

static int grid_for(long long total, int block, int per_sm = 8) {
  long long g = cdiv(total, block);
  long long cap = 148LL * per_sm;
  return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

void op_squeeze(Ctx& c, const Geo& g, int B, int C, View in, View out) {
  INB_CHECK(!(g.W % 2) && !(g.H % 2) && (g.nd == 2 || !(g.D % 2)),
            "Input dimensions must be multiple of 2");  // dimensionality_operations.jl:82-84
  if (c.dry()) return;
  Prof pf(c, F_SQUEEZE, 1, 0, 8.0 * B * C * g.px);
  long long total = (long long)B * C * g.D * g.H * (g.W / 2);
  if (!launch_squeeze4<true>(c, g, B, C, in, out))
    k_squeeze<true><<<grid_for(total, 256), 256, 0, c.st>>>(in.p, in.bs, out.p, out.bs, C, g.W, g.H, g.D, B);
  INB_CUDA(cudaGetLastError());
}
void op_unsqueeze(Ctx& c, const Geo& g, int B, int C, View in, View out) {
  if (c.dry()) return;
  Prof pf(c, F_SQUEEZE, 1, 0, 8.0 * B * C * g.px);
  long long total = (long long)B * C * g.D * g.H * (g.W / 2);
  if (!launch_squeeze4<false>(c, g, B, C, out, in))
    k_squeeze<false><<<grid_for(total, 256), 256, 0, c.st>>>(in.p, in.bs, out.p, out.bs, C, g.W, g.H, g.D, B);
  INB_CUDA(cudaGetLastError());
}

// One-level orthonormal Haar transform of every (sample, channel) image followed by the patch squeeze, 2-D.
// wavelet_squeeze (dimensionality_operations.jl:199-216, WT.db1) and Haar_squeeze (:318-331) compute the same four
// butterflies of a 2x2 block and differ in the output channel order only:
//   a  = (p00 + p10 + p01 + p11) / 2       p[ix][iy] = in[2x'+ix, 2y'+iy]
//   dx = (p00 - p10 + p01 - p11) / 2       detail along x (HaarLift(., 1): first - second, :272-281)
//   dy = (p00 + p10 - p01 - p11) / 2       detail along y
//   dd = (p00 - p10 - p01 + p11) / 2
// TYPE 0 (wavelet): channel 4c + q, q = (a, dx, dy, dd) - the quadrants of dwt() in patch order (:30-36)
// TYPE 1 (Haar lifting): channel q*C + c, q = (a, v = dy, h = dx, d = dd) - cat(a, v, h, d) of :330
// One thread owns one half-resolution pixel of one (b, c): two 8-byte accesses on the full-resolution side, four
// 4-byte accesses (coalesced over x') on the squeezed side.
template <bool FWD>
__global__ void k_haar(const float* __restrict__ src, long long sbs, float* __restrict__ dst, long long dbs, int C,
                       int W, int H, int B, int type) {
  const int Wh = W >> 1, Hh = H >> 1;
  const long long pxf = (long long)W * H, pxh = pxf >> 2;
  const long long total = (long long)B * C * Hh * Wh;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long t = i;
    const int xh = (int)(t % Wh); t /= Wh;
    const int yh = (int)(t % Hh); t /= Hh;
    const int c = (int)(t % C);
    const int b = (int)(t / C);
    const long long full = (long long)b * (FWD ? sbs : dbs) + (long long)c * pxf + (long long)(2 * yh) * W + 2 * xh;
    const long long half = (long long)b * (FWD ? dbs : sbs) + (long long)yh * Wh + xh;
    long long ch[4];  // squeezed-side channel of (a, dx, dy, dd)
    if (type == 0) {
      ch[0] = 4LL * c; ch[1] = 4LL * c + 1; ch[2] = 4LL * c + 2; ch[3] = 4LL * c + 3;
    } else {
      ch[0] = c; ch[1] = 2LL * C + c; ch[2] = (long long)C + c; ch[3] = 3LL * C + c;
    }
    if (FWD) {
      const float2 r0 = *reinterpret_cast<const float2*>(src + full);      // p00 p10
      const float2 r1 = *reinterpret_cast<const float2*>(src + full + W);  // p01 p11
      const float s0 = r0.x + r0.y, d0 = r0.x - r0.y, s1 = r1.x + r1.y, d1 = r1.x - r1.y;
      dst[half + ch[0] * pxh] = 0.5f * (s0 + s1);
      dst[half + ch[1] * pxh] = 0.5f * (d0 + d1);
      dst[half + ch[2] * pxh] = 0.5f * (s0 - s1);
      dst[half + ch[3] * pxh] = 0.5f * (d0 - d1);
    } else {
      const float a = src[half + ch[0] * pxh], dx = src[half + ch[1] * pxh];
      const float dy = src[half + ch[2] * pxh], dd = src[half + ch[3] * pxh];
      const float s0 = a + dy, s1 = a - dy, d0 = dx + dd, d1 = dx - dd;
      *reinterpret_cast<float2*>(dst + full) = make_float2(0.5f * (s0 + d0), 0.5f * (s0 - d0));
      *reinterpret_cast<float2*>(dst + full + W) = make_float2(0.5f * (s1 + d1), 0.5f * (s1 - d1));
    }
  }
}
static void haar_check(const Geo& g, View full) {
  INB_CHECK(g.nd == 2, "the Haar / wavelet squeeze is implemented for 2-D tensors only");
  INB_CHECK(!(g.W % 2) && !(g.H % 2), "Input dimensions must be multiple of 2");
  INB_CHECK(full.bs % 2 == 0 && ((uintptr_t)full.p & 7) == 0, "Haar squeeze: full-resolution tensor must be 8-byte aligned");
}
// g = full-resolution geometry, C = full-resolution channels
void op_haar_squeeze(Ctx& c, const Geo& g, int B, int C, int type, View in, View out) {
  INB_CHECK(type == 0 || type == 1, "squeeze type must be 0 (wavelet db1) or 1 (Haar lifting)");
  if (c.dry()) { INB_CHECK(g.nd == 2 && !(g.W % 2) && !(g.H % 2), "Input dimensions must be multiple of 2 (2-D only)"); return; }
  haar_check(g, in);
  Prof pf(c, F_SQUEEZE, 1, 0, 8.0 * B * C * g.px);
  k_haar<true><<<grid_for((long long)B * C * (g.px / 4), 256), 256, 0, c.st>>>(in.p, in.bs, out.p, out.bs, C, g.W, g.H, B, type);
  INB_CUDA(cudaGetLastError());
}
void op_haar_unsqueeze(Ctx& c, const Geo& g, int B, int C, int type, View in, View out) {
  INB_CHECK(type == 0 || type == 1, "squeeze type must be 0 (wavelet db1) or 1 (Haar lifting)");
  if (c.dry()) return;
  haar_check(g, out);
  Prof pf(c, F_SQUEEZE, 1, 0, 8.0 * B * C * g.px);
  k_haar<false><<<grid_for((long long)B * C * (g.px / 4), 256), 256, 0, c.st>>>(in.p, in.bs, out.p, out.bs, C, g.W, g.H, B, type);
  INB_CUDA(cudaGetLastError());
}

// dst[i] += src[i] for up to five (src, dst, n) pairs in one launch (grid row = pair): the five ResidualBlock gradients
// of a coupling layer that the HINT recursion visits more than once
struct AccumArgs {
  const float* src[5];
  float* dst[5];
  long long n[5];
};
__global__ void k_accum(const AccumArgs a) {
  const float* __restrict__ src = a.src[blockIdx.y];
  float* __restrict__ dst = a.dst[blockIdx.y];
  const long long n = a.n[blockIdx.y];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] += src[i];
}
void op_accum(Ctx& c, int npairs, const float* const* src, float* const* dst, const long long* n) {
  INB_CHECK(npairs >= 1 && npairs <= 5, "op_accum: 1 to 5 pairs");
  if (c.dry()) return;
  AccumArgs a{};
  long long nmax = 0, tot = 0;
  for (int i = 0; i < npairs; ++i) {
    a.src[i] = src[i]; a.dst[i] = dst[i]; a.n[i] = n[i];
    nmax = std::max(nmax, n[i]);
    tot += n[i];
  }
  Prof pf(c, F_MISC, 1, 0, 12.0 * tot);
  k_accum<<<dim3(grid_for(nmax, 256, 2), npairs, 1), 256, 0, c.st>>>(a);
  INB_CUDA(cudaGetLastError());
}

template <int V>
__global__ void k_copy(const float* __restrict__ src, long long sbs, float* __restrict__ dst,
                       long long dbs, long long per_sample /* C*px */, int B) {
  const long long nv = per_sample / V;
  const long long total = nv * B;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long b = i / nv, e = (i - b * nv) * V;
    float r[V];
    ldv<V>(src + b * sbs + e, r);
    stv<V>(dst + b * dbs + e, r);
  }
}
static bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

void op_copy(Ctx& c, long long px, int B, int C, View in, View out) {
  if (c.dry()) return;
  Prof pf(c, F_COPY, 1, 0, 8.0 * B * C * px);
  long long per = (long long)C * px;
  if (per == 0 || B == 0) return;
  bool v4 = (per % 4 == 0) && (in.bs % 4 == 0) && (out.bs % 4 == 0) && aligned16(in.p) && aligned16(out.p);
  if (v4)
    k_copy<4><<<grid_for(per / 4 * B, 256), 256, 0, c.st>>>(in.p, in.bs, out.p, out.bs, per, B);
  else
    k_copy<1><<<grid_for(per * B, 256), 256, 0, c.st>>>(in.p, in.bs, out.p, out.bs, per, B);
  INB_CUDA(cudaGetLastError());
}
void op_zero(Ctx& c, void* p, size_t bytes) {
  if (c.dry() || bytes == 0) return;
  INB_CUDA(cudaMemsetAsync(p, 0, bytes, c.st));
}

// ---------------------------------------------------------------- per-channel statistics
// out[c] += sum over (b, pix) of f(x) with f = x - shift[c] (MODE 0: shift null -> plain sum) or
// (x - shift[c])^2 (MODE 1).  grid (chunks, C); fp32 per-thread partials, fp64 across threads.
template <int MODE>
__global__ void k_chan_stat(const float* __restrict__ x, long long bs, long long px, int B,
                            const double* __restrict__ shift, double scale_shift, double* __restrict__ out) {
  const int c = blockIdx.y;
  const float sh = shift ? (float)(shift[c] * scale_shift) : 0.f;
  const long long n = px * B;
  double acc = 0.0;
  float part = 0.f;
  int cnt = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    long long b = i / px, p = i - b * px;
    float v = x[b * bs + c * px + p] - sh;
    part += MODE ? v * v : v;
    if (++cnt == 64) { acc += part; part = 0.f; cnt = 0; }
  }
  acc += part;
  double r = block_sum(acc);
  if (threadIdx.x == 0) atomicAdd(out + c, r);
}

__global__ void k_actnorm_init_finish(const double* __restrict__ sum, const double* __restrict__ ss,
                                      double n, int C, float* __restrict__ s, float* __restrict__ b) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float mu = (float)(sum[c] / n);
  float var = (float)(ss[c] / (n - 1.0));  // Statistics.var: unbiased
  float sd = sqrtf(var);
  s[c] = 1.f / sd;     // actnorm.jl:70
  b[c] = -mu / sd;     // actnorm.jl:71
}

static dim3 stat_grid(long long n, int C) {
  long long chunks = cdiv(n, 256LL * 16);
  long long cap = cdiv(148LL * 8, C);
  if (chunks > cap) chunks = cap;
  if (chunks < 1) chunks = 1;
  return dim3((unsigned)chunks, (unsigned)C);
}

void op_actnorm_init(Ctx& c, long long px, int B, int C, View x, float* s, float* b) {
  size_t m = c.ar->mark();
  double* sum = c.ar->f64(2 * (size_t)C);
  double* ss = sum + C;
  if (!c.dry()) {
    Prof pf(c, F_AN_STATS, 3, 0, 8.0 * B * C * px);
    INB_CUDA(cudaMemsetAsync(sum, 0, 2 * C * sizeof(double), c.st));
    double n = (double)px * B;
    dim3 g = stat_grid(px * B, C);
    k_chan_stat<0><<<g, 256, 0, c.st>>>(x.p, x.bs, px, B, nullptr, 0.0, sum);
    if (c.dp && c.dp->nranks > 1) {
      // data parallel: the statistics of the GLOBAL batch (invertible_layer_actnorm.jl:67-72 on the concatenated
      // shards; equal shards): sum over the ranks of the per-shard sums, then of the squared deviations from the
      // global mean - the same two-pass mean / unbiased variance as on one device
      dp_allreduce_sum_f64(c.dp, sum, C, c.st);
      n *= c.dp->nranks;
    }
    k_chan_stat<1><<<g, 256, 0, c.st>>>(x.p, x.bs, px, B, sum, 1.0 / n, ss);
    if (c.dp && c.dp->nranks > 1) dp_allreduce_sum_f64(c.dp, ss, C, c.st);
    k_actnorm_init_finish<<<(C + 127) / 128, 128, 0, c.st>>>(sum, ss, n, C, s, b);
    INB_CUDA(cudaGetLastError());
  }
  c.ar->release(m);
}

__global__ void k_cast_d2f(const double* __restrict__ in, float* __restrict__ out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (float)in[i];
}
void op_channel_sum(Ctx& c, long long px, int B, int C, const float* in, float* out) {
  size_t m = c.ar->mark();
  double* acc = c.ar->f64(C);
  if (!c.dry()) {
    Prof pf(c, F_CHANNEL_SUM, 2, 0, 4.0 * B * C * px);
    INB_CUDA(cudaMemsetAsync(acc, 0, C * sizeof(double), c.st));
    k_chan_stat<0><<<stat_grid(px * B, C), 256, 0, c.st>>>(in, (long long)C * px, px, B, nullptr, 0.0, acc);
    k_cast_d2f<<<(C + 127) / 128, 128, 0, c.st>>>(acc, out, C);
    INB_CUDA(cudaGetLastError());
  }
  c.ar->release(m);
}

// ---------------------------------------------------------------- ActNorm + Householder, forward
// One reflection of compute_utils.jl:24-27: t = a.v ; t *= -2/(v.v) ; a += t*v
template <int C, int V>
__device__ __forceinline__ void reflect(float (&a)[C][V], const float* __restrict__ v, float n) {
#pragma unroll
  for (int u = 0; u < V; ++u) {
    float t = 0.f;
#pragma unroll
    for (int ch = 0; ch < C; ++ch) t = fmaf(a[ch][u], v[ch], t);
    t *= n;
#pragma unroll
    for (int ch = 0; ch < C; ++ch) a[ch][u] = fmaf(t, v[ch], a[ch][u]);
  }
}

template <int C>
struct HhSmem {
  float s[C], b[C], v[3][C], n[3];
};

template <int C>
__device__ __forceinline__ void load_hh_smem(HhSmem<C>& sm, const float* s, const float* b,
                                             const float* v1, const float* v2, const float* v3) {
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    sm.s[i] = s ? s[i] : 1.f;
    sm.b[i] = b ? b[i] : 0.f;
    sm.v[0][i] = v1 ? v1[i] : 0.f;
    sm.v[1][i] = v1 ? v2[i] : 0.f;
    sm.v[2][i] = v1 ? v3[i] : 0.f;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    float d = 0.f;
    for (int i = 0; i < C; ++i) d = fmaf(sm.v[threadIdx.x][i], sm.v[threadIdx.x][i], d);
    sm.n[threadIdx.x] = -2.f / d;  // compute_utils.jl:24
  }
  __syncthreads();
}

template <int C, int V>
__global__ void __launch_bounds__(256)
k_an_hh_fwd(const float* __restrict__ x, long long xbs, float* __restrict__ y, long long ybs,
            const float* __restrict__ s, const float* __restrict__ b, const float* __restrict__ v1,
            const float* __restrict__ v2, const float* __restrict__ v3, long long px, long long total,
            double* __restrict__ ld) {
  __shared__ HhSmem<C> sm;
  load_hh_smem<C>(sm, s, b, v1, v2, v3);
  if (ld && s && blockIdx.x == 0 && threadIdx.x == 0) {
    float acc = 0.f;
    for (int i = 0; i < C; ++i) acc += logf(fabsf(sm.s[i]));
    atomicAdd(ld, (double)((float)px * acc));  // actnorm.jl:188-189
  }
  // persistent blocks (one wave of resident CTAs, grid-stride): no partial last wave, parameters staged once per CTA
  const long long pxv = px / V;
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
    const long long bi = g / pxv, pix = (g - bi * pxv) * V;
    const float* xp = x + bi * xbs + pix;
    float a[C][V];
#pragma unroll
    for (int ch = 0; ch < C; ++ch) ldv<V>(xp + ch * px, a[ch]);
    if (s) {
#pragma unroll
      for (int ch = 0; ch < C; ++ch)
#pragma unroll
        for (int u = 0; u < V; ++u) a[ch][u] = a[ch][u] * sm.s[ch] + sm.b[ch];
    }
    if (v1) {
      reflect<C, V>(a, sm.v[0], sm.n[0]);
      reflect<C, V>(a, sm.v[1], sm.n[1]);
      reflect<C, V>(a, sm.v[2], sm.n[2]);
    }
    float* yp = y + bi * ybs + pix;
#pragma unroll
    for (int ch = 0; ch < C; ++ch) stv<V>(yp + ch * px, a[ch]);
  }
}

template <int C, int V>
__global__ void __launch_bounds__(256)
k_hh_an_inv(const float* __restrict__ y, long long ybs, float* __restrict__ x, long long xbs,
            const float* __restrict__ s, const float* __restrict__ b, const float* __restrict__ v1,
            const float* __restrict__ v2, const float* __restrict__ v3, long long px, long long total) {
  __shared__ HhSmem<C> sm;
  load_hh_smem<C>(sm, s, b, v1, v2, v3);
  const long long pxv = px / V;
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
    const long long bi = g / pxv, pix = (g - bi * pxv) * V;
    const float* yp = y + bi * ybs + pix;
    float a[C][V];
#pragma unroll
    for (int ch = 0; ch < C; ++ch) ldv<V>(yp + ch * px, a[ch]);
    if (v1) {
      reflect<C, V>(a, sm.v[2], sm.n[2]);
      reflect<C, V>(a, sm.v[1], sm.n[1]);
      reflect<C, V>(a, sm.v[0], sm.n[0]);
    }
    if (s) {
#pragma unroll
      for (int ch = 0; ch < C; ++ch)
#pragma unroll
        for (int u = 0; u < V; ++u) a[ch][u] = (a[ch][u] - sm.b[ch]) / sm.s[ch];  // actnorm.jl:93
    }
    float* xp = x + bi * xbs + pix;
#pragma unroll
    for (int ch = 0; ch < C; ++ch) stv<V>(xp + ch * px, a[ch]);
  }
}

// ---------------------------------------------------------------- backward of both
// Per pixel: dA = dY.H3H2H1, A = Y.H3H2H1 (conv1x1.jl:230-231), X = (A-b)/s, dX = dA*s
// (actnorm.jl:105-106).  Reductions: gram += A^T dY (the only data-dependent input of the
// Householder gradients), ds += sum dA*X, db += sum dA (actnorm.jl:107,111).
// A tile of 128 pixels is staged through shared memory as [C][128+4] rows for the reductions.
constexpr int BWD_T = 128;        // threads == pixels per tile
constexpr int BWD_LD = BWD_T + 4; // row pitch: 16B aligned, conflict-free for 128-bit reads

template <int C>
__global__ void __launch_bounds__(BWD_T)
k_hh_an_bwd(const float* __restrict__ dy, long long dybs, const float* __restrict__ y, long long ybs,
            float* __restrict__ dx, long long dxbs, float* __restrict__ x, long long xbs,
            const float* __restrict__ s, const float* __restrict__ b, const float* __restrict__ v1,
            const float* __restrict__ v2, const float* __restrict__ v3, long long px, long long total,
            double* __restrict__ gram, double* __restrict__ dsdb) {
  extern __shared__ __align__(16) float smem[];
  __shared__ HhSmem<C> sm;
  float* As = smem;                 // A   [C][BWD_LD]
  float* Ds = As + C * BWD_LD;      // dY  [C][BWD_LD]
  float* Es = Ds + C * BWD_LD;      // dA  [C][BWD_LD]
  load_hh_smem<C>(sm, s, b, v1, v2, v3);
  const int tid = threadIdx.x;
  // gram sub-block owned by this thread: rows ti + 8*ri, cols tj + 16*rj
  constexpr int RI = (C + 7) / 8, RJ = (C + 15) / 16;
  const int ti = tid >> 4, tj = tid & 15;
  float gacc[RI][RJ];
#pragma unroll
  for (int i = 0; i < RI; ++i)
#pragma unroll
    for (int j = 0; j < RJ; ++j) gacc[i][j] = 0.f;
  // register-tiled variant for the common channel counts: a thread owns a TB x TB block of the Gram matrix over one
  // of GP pixel groups of the tile (TB + TB shared-memory vectors per TB*TB*4 FMAs instead of 3 per 8: the scattered
  // variant above is bound by shared-memory bandwidth)
  constexpr bool kTiled = (C == 6 || C == 8 || C == 12 || C == 16 || C == 24);
  constexpr int TB = (C == 6) ? 3 : (C == 8) ? 2 : (C == 12) ? 3 : (C == 16) ? 4 : (C == 24) ? 6 : 1;
  constexpr int NB = kTiled ? C / TB : 1, NBLK = NB * NB;
  constexpr int GP = kTiled ? BWD_T / NBLK : 1, PG = BWD_T / GP;  // pixel groups per tile, pixels per group
  const int blk = tid % NBLK, grp = tid / NBLK;
  const int bi = blk / NB, bj = blk % NB;
  float tacc[TB][TB];
#pragma unroll
  for (int i = 0; i < TB; ++i)
#pragma unroll
    for (int j = 0; j < TB; ++j) tacc[i][j] = 0.f;
  constexpr int NG = BWD_T / C;   // threads per channel in the ds / db walk
  float sacc = 0.f, bacc = 0.f;   // threads < C * NG: partial ds * s, db of channel tid % C

  const long long ntiles = (total + BWD_T - 1) / BWD_T;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long g = tile * BWD_T + tid;
    const bool live = g < total;
    {
      float a[C][1], d[C][1];
      long long bi = 0, pix = 0;
      if (live) {
        bi = g / px;
        pix = g - bi * px;
        const float* yp = y + bi * ybs + pix;
        const float* dp = dy + bi * dybs + pix;
#pragma unroll
        for (int ch = 0; ch < C; ++ch) { a[ch][0] = yp[ch * px]; d[ch][0] = dp[ch * px]; }
      } else {
#pragma unroll
        for (int ch = 0; ch < C; ++ch) { a[ch][0] = 0.f; d[ch][0] = 0.f; }
      }
      if (v1 && gram) {
#pragma unroll
        for (int ch = 0; ch < C; ++ch) Ds[ch * BWD_LD + tid] = d[ch][0];
      }
      if (v1) {
        reflect<C, 1>(a, sm.v[2], sm.n[2]);
        reflect<C, 1>(a, sm.v[1], sm.n[1]);
        reflect<C, 1>(a, sm.v[0], sm.n[0]);
        reflect<C, 1>(d, sm.v[2], sm.n[2]);
        reflect<C, 1>(d, sm.v[1], sm.n[1]);
        reflect<C, 1>(d, sm.v[0], sm.n[0]);
      }
#pragma unroll
      for (int ch = 0; ch < C; ++ch) {
        As[ch * BWD_LD + tid] = a[ch][0];
        Es[ch * BWD_LD + tid] = d[ch][0];
      }
      if (live) {
        float* xp = x + bi * xbs + pix;
        float* dxp = dx + bi * dxbs + pix;
        if (s) {
#pragma unroll
          for (int ch = 0; ch < C; ++ch) {
            xp[ch * px] = (a[ch][0] - sm.b[ch]) / sm.s[ch];
            dxp[ch * px] = d[ch][0] * sm.s[ch];
          }
        } else {
#pragma unroll
          for (int ch = 0; ch < C; ++ch) { xp[ch * px] = a[ch][0]; dxp[ch * px] = d[ch][0]; }
        }
      }
    }
    __syncthreads();
    if (kTiled && v1 && gram) {
#pragma unroll
      for (int p = grp * PG; p < grp * PG + PG; p += 4) {
        float4 av[TB], dv[TB];
#pragma unroll
        for (int i = 0; i < TB; ++i) av[i] = *reinterpret_cast<const float4*>(As + (bi * TB + i) * BWD_LD + p);
#pragma unroll
        for (int j = 0; j < TB; ++j) dv[j] = *reinterpret_cast<const float4*>(Ds + (bj * TB + j) * BWD_LD + p);
#pragma unroll
        for (int i = 0; i < TB; ++i)
#pragma unroll
          for (int j = 0; j < TB; ++j) {
            tacc[i][j] = fmaf(av[i].x, dv[j].x, tacc[i][j]);
            tacc[i][j] = fmaf(av[i].y, dv[j].y, tacc[i][j]);
            tacc[i][j] = fmaf(av[i].z, dv[j].z, tacc[i][j]);
            tacc[i][j] = fmaf(av[i].w, dv[j].w, tacc[i][j]);
          }
      }
    } else if (v1 && gram) {
      for (int p = 0; p < BWD_T; p += 4) {
        float4 av[RI], dv[RJ];
#pragma unroll
        for (int i = 0; i < RI; ++i) {
          int r = ti + 8 * i;
          av[i] = (r < C) ? *reinterpret_cast<const float4*>(As + r * BWD_LD + p) : make_float4(0, 0, 0, 0);
        }
#pragma unroll
        for (int j = 0; j < RJ; ++j) {
          int r = tj + 16 * j;
          dv[j] = (r < C) ? *reinterpret_cast<const float4*>(Ds + r * BWD_LD + p) : make_float4(0, 0, 0, 0);
        }
#pragma unroll
        for (int i = 0; i < RI; ++i)
#pragma unroll
          for (int j = 0; j < RJ; ++j) {
            gacc[i][j] = fmaf(av[i].x, dv[j].x, gacc[i][j]);
            gacc[i][j] = fmaf(av[i].y, dv[j].y, gacc[i][j]);
            gacc[i][j] = fmaf(av[i].z, dv[j].z, gacc[i][j]);
            gacc[i][j] = fmaf(av[i].w, dv[j].w, gacc[i][j]);
          }
      }
    }
    if (s && dsdb && tid < C * NG) {
      // sum_p dA (A - b) / s = (sum_p dA A - b sum_p dA) / s: the division leaves the loop, and the walk over the tile is
      // split over NG = 128 / C threads per channel (C threads alone made this the longest phase of a tile)
      const int ch = tid % C, j = tid / C;
      const float bc = sm.b[ch];
      float s1 = 0.f, b1 = 0.f;
      for (int p = 4 * j; p < BWD_T; p += 4 * NG) {
        float4 av = *reinterpret_cast<const float4*>(As + ch * BWD_LD + p);
        float4 ev = *reinterpret_cast<const float4*>(Es + ch * BWD_LD + p);
        // dead pixels carry dA = 0, so they add nothing
        s1 = fmaf(ev.x, av.x - bc, s1);
        s1 = fmaf(ev.y, av.y - bc, s1);
        s1 = fmaf(ev.z, av.z - bc, s1);
        s1 = fmaf(ev.w, av.w - bc, s1);
        b1 += (ev.x + ev.y) + (ev.z + ev.w);
      }
      sacc += s1;
      bacc += b1;
    }
    __syncthreads();
  }
  if (kTiled && v1 && gram) {
    // the GP partial blocks of every Gram entry meet in shared memory (the tiles are free now): one atomic per entry
    __syncthreads();
    float* red = smem;  // [GP][C*C]
#pragma unroll
    for (int i = 0; i < TB; ++i)
#pragma unroll
      for (int j = 0; j < TB; ++j) red[grp * C * C + (bi * TB + i) * C + bj * TB + j] = tacc[i][j];
    __syncthreads();
    for (int e = tid; e < C * C; e += BWD_T) {
      float u = 0.f;
      for (int k = 0; k < GP; ++k) u += red[k * C * C + e];
      atomicAdd(gram + e, (double)u);
    }
  } else if (v1 && gram) {
#pragma unroll
    for (int i = 0; i < RI; ++i)
#pragma unroll
      for (int j = 0; j < RJ; ++j) {
        int r = ti + 8 * i, cc = tj + 16 * j;
        if (r < C && cc < C) atomicAdd(gram + r * C + cc, (double)gacc[i][j]);
      }
  }
  if (s && dsdb) {
    // the NG partial sums of a channel meet in shared memory (the A tile is free now), one pair of atomics per channel
    __syncthreads();
    if (tid < C * NG) {
      As[tid] = sacc;
      As[C * NG + tid] = bacc;
    }
    __syncthreads();
    if (tid < C) {
      float u = 0.f, w = 0.f;
      for (int j = 0; j < NG; ++j) { u += As[j * C + tid]; w += As[C * NG + j * C + tid]; }
      atomicAdd(dsdb + tid, (double)(u / sm.s[tid]));
      atomicAdd(dsdb + C + tid, (double)w);
    }
  }
}

// gram -> dv.  With G = sum x^T dy, L = <dY, X H1 H2 H3>:  dL/dH1 = G (H2 H3)^T, dL/dH2 = H1 G H3,
// dL/dH3 = H2 H1 G, and for H = I - 2 v v^T / (v^T v) and any M = dL/dH:
//   dL/dv = -2 [ (M v + M^T v)/n - 2 (v^T M v) v / n^2 ],  n = v^T v.
// This is conv1x1.jl:118-170 (d/dv of the chain through partial_derivative_outer :69-87)
// contracted with G instead of looping over batch elements.  fp64, one CTA.
__device__ void hh_apply_left(double* M, const double* v, double n, int C, double* tmp) {
  // M <- H M,  H = I - 2 v v^T / n
  for (int j = threadIdx.x; j < C; j += blockDim.x) {
    double t = 0;
    for (int i = 0; i < C; ++i) t += v[i] * M[i * C + j];
    tmp[j] = 2.0 * t / n;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < C * C; e += blockDim.x) M[e] -= v[e / C] * tmp[e % C];
  __syncthreads();
}
__device__ void hh_apply_right(double* M, const double* v, double n, int C, double* tmp) {
  // M <- M H
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    double t = 0;
    for (int j = 0; j < C; ++j) t += M[i * C + j] * v[j];
    tmp[i] = 2.0 * t / n;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < C * C; e += blockDim.x) M[e] -= tmp[e / C] * v[e % C];
  __syncthreads();
}
__device__ void hh_vgrad(const double* M, const double* v, double n, int C, double* tmp, float* out) {
  // tmp[0..C) = M v + M^T v ; tmp[C] = v^T M v
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    double a = 0, b = 0;
    for (int j = 0; j < C; ++j) { a += M[i * C + j] * v[j]; b += M[j * C + i] * v[j]; }
    tmp[i] = a + b;
    tmp[C + 1 + i] = a * v[i];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double q = 0;
    for (int i = 0; i < C; ++i) q += tmp[C + 1 + i];
    tmp[C] = q;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x)
    out[i] = (float)(-2.0 * (tmp[i] / n - 2.0 * tmp[C] * v[i] / (n * n)));
  __syncthreads();
}

// ActNorm's part of the same step (actnorm.jl:108-114,190), folded into this launch when `s` is given
struct AnFinish {
  const double* dsdb;
  const float* s;
  double px;
  int logdet;
  float *ds, *db;
};
__global__ void k_hh_grad_finish(const double* __restrict__ gram, const float* __restrict__ v1f,
                                 const float* __restrict__ v2f, const float* __restrict__ v3f, int C,
                                 int freeze, float* dv1, float* dv2, float* dv3, const AnFinish an) {
  extern __shared__ double dsm[];
  if (an.s) {
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
      float v = (float)an.dsdb[i];
      if (an.logdet) v -= (float)an.px / an.s[i];  // actnorm.jl:108-110,190
      an.ds[i] = v;
      an.db[i] = (float)an.dsdb[C + i];
    }
  }
  if (freeze) {  // conv1x1.jl:132-134
    for (int i = threadIdx.x; i < C; i += blockDim.x) { dv1[i] = 0.f; dv2[i] = 0.f; dv3[i] = 0.f; }
    return;
  }
  double* M = dsm;              // C*C
  double* v = M + C * C;        // 3*C
  double* tmp = v + 3 * C;      // 2*C+2
  __shared__ double n[3];
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    v[i] = v1f[i]; v[C + i] = v2f[i]; v[2 * C + i] = v3f[i];
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double d = 0;
    for (int i = 0; i < C; ++i) d += v[threadIdx.x * C + i] * v[threadIdx.x * C + i];
    n[threadIdx.x] = d;
  }
  __syncthreads();
  // dv1: M = G H3 H2
  for (int e = threadIdx.x; e < C * C; e += blockDim.x) M[e] = gram[e];
  __syncthreads();
  hh_apply_right(M, v + 2 * C, n[2], C, tmp);
  hh_apply_right(M, v + C, n[1], C, tmp);
  hh_vgrad(M, v, n[0], C, tmp, dv1);
  // dv2: M = H1 G H3
  for (int e = threadIdx.x; e < C * C; e += blockDim.x) M[e] = gram[e];
  __syncthreads();
  hh_apply_left(M, v, n[0], C, tmp);
  hh_apply_right(M, v + 2 * C, n[2], C, tmp);
  hh_vgrad(M, v + C, n[1], C, tmp, dv2);
  // dv3: M = H2 H1 G
  for (int e = threadIdx.x; e < C * C; e += blockDim.x) M[e] = gram[e];
  __syncthreads();
  hh_apply_left(M, v, n[0], C, tmp);
  hh_apply_left(M, v + C, n[1], C, tmp);
  hh_vgrad(M, v + 2 * C, n[2], C, tmp, dv3);
}

static void launch_hh_grad_finish(Ctx& c, int C, const double* gram, const float* v1, const float* v2, const float* v3,
                                  int freeze, float* dv1, float* dv2, float* dv3, const AnFinish& an) {
  Prof pf(c, F_GRAD_FINISH, 1, 0, 0);
  size_t sh = ((size_t)C * C + 3 * C + 2 * C + 2) * sizeof(double);
  INB_CHECK(sh <= 200 * 1024, "Conv1x1 with %d channels is not supported", C);
  if (sh > 48 * 1024)
    INB_CUDA(cudaFuncSetAttribute(k_hh_grad_finish, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
  k_hh_grad_finish<<<1, 256, sh, c.st>>>(gram, v1, v2, v3, C, freeze, dv1, dv2, dv3, an);
  INB_CUDA(cudaGetLastError());
}
void op_hh_grad_finish(Ctx& c, int C, const double* gram, const float* v1, const float* v2,
                       const float* v3, int freeze, float* dv1, float* dv2, float* dv3) {
  if (c.dry()) return;
  launch_hh_grad_finish(c, C, gram, v1, v2, v3, freeze, dv1, dv2, dv3, AnFinish{nullptr, nullptr, 0.0, 0, nullptr, nullptr});
}
// the Householder and the ActNorm gradients of one flow step in ONE launch
void op_hh_an_grad_finish(Ctx& c, int C, long long px, const double* gram, const float* v1, const float* v2, const float* v3,
                          int freeze, float* dv1, float* dv2, float* dv3, const double* dsdb, const float* s, int logdet,
                          float* ds, float* db) {
  if (c.dry()) return;
  launch_hh_grad_finish(c, C, gram, v1, v2, v3, freeze, dv1, dv2, dv3, AnFinish{dsdb, s, (double)px, logdet, ds, db});
}

__global__ void k_an_grad_finish(const double* __restrict__ dsdb, const float* __restrict__ s, int C,
                                 double px, int logdet, float* __restrict__ ds, float* __restrict__ db) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C) return;
  float v = (float)dsdb[i];
  if (logdet) v -= (float)px / s[i];  // actnorm.jl:108-110,190
  ds[i] = v;
  db[i] = (float)dsdb[C + i];
}
void op_an_grad_finish(Ctx& c, int C, long long px, const double* dsdb, const float* s, int logdet,
                       float* ds, float* db) {
  if (c.dry()) return;
  Prof pf(c, F_GRAD_FINISH, 1, 0, 0);
  k_an_grad_finish<<<(C + 127) / 128, 128, 0, c.st>>>(dsdb, s, C, (double)px, logdet, ds, db);
  INB_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------- dispatch on the channel count
#define INB_FOR_C(C_, MACRO)                                                                    \
  switch (C_) {                                                                                 \
    case 1: MACRO(1) break; case 2: MACRO(2) break; case 3: MACRO(3) break; case 4: MACRO(4) break; \
    case 6: MACRO(6) break; case 8: MACRO(8) break; case 12: MACRO(12) break;                   \
    case 16: MACRO(16) break; case 24: MACRO(24) break; case 32: MACRO(32) break;               \
    case 48: MACRO(48) break; case 64: MACRO(64) break;                                         \
    default:                                                                                    \
      fail(1, "channel count %d is not supported by the per-pixel kernels "                     \
              "(supported: 1,2,3,4,6,8,12,16,24,32,48,64)", C_);                                \
  }

static int pick_vec(int C, long long px, std::initializer_list<const View*> vs) {
  int v = (C <= 12) ? 4 : (C <= 24 ? 2 : 1);
  while (v > 1) {
    bool ok = (px % v == 0);
    for (const View* w : vs) ok = ok && (w->bs % v == 0) && (((uintptr_t)w->p) % (4 * v) == 0);
    if (ok) break;
    v >>= 1;
  }
  return v;
}

// one wave of resident CTAs for a grid-stride kernel of 256 threads (occupancy queried once per instantiation)
template <class K>
static int persistent_grid(K kern, long long blocks) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, 0) != cudaSuccess || per_sm < 1) {
    cudaGetLastError();
    per_sm = 2;
  }
  const long long cap = 148LL * per_sm;
  return (int)std::max<long long>(1, std::min(blocks, cap));
}
#define INB_PGRID(KERN) ([&] { static int per = 0; if (!per) per = persistent_grid(KERN, 1LL << 40) / 148; \
                               return (int)std::max<long long>(1, std::min<long long>(nblk, 148LL * per)); }())

void op_an_hh_fwd(Ctx& c, long long px, int B, int C, View x, View y, const float* s, const float* b,
                  const float* v1, const float* v2, const float* v3, double* ld) {
  if (c.dry()) return;
  Prof pf(c, F_AN_HH_FWD, 1, 12.0 * B * C * px, 8.0 * B * C * px);
  int V = pick_vec(C, px, {&x, &y});
  long long total = px / V * B;
  const long long nblk = cdiv(total, 256);
#define L_(CC)                                                                                       \
  if (V == 4) { auto k = k_an_hh_fwd<CC, (CC <= 12 ? 4 : 1)>; k<<<INB_PGRID(k), 256, 0, c.st>>>(x.p, x.bs, y.p, y.bs, s, b, v1, v2, v3, px, total, ld); } \
  else if (V == 2) { auto k = k_an_hh_fwd<CC, (CC <= 24 ? 2 : 1)>; k<<<INB_PGRID(k), 256, 0, c.st>>>(x.p, x.bs, y.p, y.bs, s, b, v1, v2, v3, px, total, ld); } \
  else { auto k = k_an_hh_fwd<CC, 1>; k<<<INB_PGRID(k), 256, 0, c.st>>>(x.p, x.bs, y.p, y.bs, s, b, v1, v2, v3, px, total, ld); }
  INB_FOR_C(C, L_)
#undef L_
  INB_CUDA(cudaGetLastError());
}

void op_hh_an_inv(Ctx& c, long long px, int B, int C, View y, View x, const float* s, const float* b,
                  const float* v1, const float* v2, const float* v3) {
  if (c.dry()) return;
  Prof pf(c, F_HH_AN_INV, 1, 12.0 * B * C * px, 8.0 * B * C * px);
  int V = pick_vec(C, px, {&x, &y});
  long long total = px / V * B;
  const long long nblk = cdiv(total, 256);
#define L_(CC)                                                                                       \
  if (V == 4) { auto k = k_hh_an_inv<CC, (CC <= 12 ? 4 : 1)>; k<<<INB_PGRID(k), 256, 0, c.st>>>(y.p, y.bs, x.p, x.bs, s, b, v1, v2, v3, px, total); } \
  else if (V == 2) { auto k = k_hh_an_inv<CC, (CC <= 24 ? 2 : 1)>; k<<<INB_PGRID(k), 256, 0, c.st>>>(y.p, y.bs, x.p, x.bs, s, b, v1, v2, v3, px, total); } \
  else { auto k = k_hh_an_inv<CC, 1>; k<<<INB_PGRID(k), 256, 0, c.st>>>(y.p, y.bs, x.p, x.bs, s, b, v1, v2, v3, px, total); }
  INB_FOR_C(C, L_)
#undef L_
  INB_CUDA(cudaGetLastError());
}

void op_hh_an_bwd(Ctx& c, long long px, int B, int C, View dy, View y, View dx, View x, const float* s,
                  const float* b, const float* v1, const float* v2, const float* v3, double* gram,
                  double* dsdb) {
  if (c.dry()) return;
  Prof pf(c, F_HH_AN_BWD, 1, (24.0 + 2.0 * C) * B * C * px, 16.0 * B * C * px);
  long long total = px * B;
  long long ntiles = cdiv(total, BWD_T);
  size_t sh = 3 * (size_t)C * BWD_LD * sizeof(float);
  int per_sm = (int)(200 * 1024 / (sh + 2048));
  if (per_sm > 8) per_sm = 8;
  if (per_sm < 1) per_sm = 1;
  long long grid = 148LL * per_sm;
  if (grid > ntiles) grid = ntiles;
#define L_(CC)                                                                                       \
  {                                                                                                  \
    if (sh > 48 * 1024)                                                                              \
      INB_CUDA(cudaFuncSetAttribute(k_hh_an_bwd<CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh)); \
    k_hh_an_bwd<CC><<<(int)grid, BWD_T, sh, c.st>>>(dy.p, dy.bs, y.p, y.bs, dx.p, dx.bs, x.p, x.bs, s, b, \
                                                    v1, v2, v3, px, total, gram, dsdb);              \
  }
  INB_FOR_C(C, L_)
#undef L_
  INB_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------- affine coupling
// idx -> (b, ch, pix) over a (B, C1, px) range, V pixels per thread
template <int V>
__global__ void k_coupling_fwd(const float* __restrict__ x1, long long xbs, float* __restrict__ y1,
                               long long ybs, const float* __restrict__ rb, long long px, int C1,
                               float low, float high, double* __restrict__ ld, float invB) {
  // grid (pixel vectors, channel, sample): no index divisions, a few vectors per thread
  const int ch = blockIdx.y;
  const long long b = blockIdx.z;
  const long long pxv = px / V;
  const float* xp = x1 + b * xbs + ch * px;
  const float* lp = rb + (b * 2 * C1 + ch) * px;
  const float* tp = rb + (b * 2 * C1 + C1 + ch) * px;
  float* yp = y1 + b * ybs + ch * px;
  float lsum = 0.f;
  for (long long pv = blockIdx.x * (long long)blockDim.x + threadIdx.x; pv < pxv; pv += (long long)gridDim.x * blockDim.x) {
    const long long pix = pv * V;
    float xv[V], ls[V], tv[V], out[V];
    ldv<V>(xp + pix, xv);
    ldv<V>(lp + pix, ls);
    ldv<V>(tp + pix, tv);
#pragma unroll
    for (int u = 0; u < V; ++u) {
      float S = sigmoid_lh(fmaxf(ls[u], 0.f), low, high);  // RB output ReLU (layer_residual_block.jl:133)
      out[u] = S * xv[u] + fmaxf(tv[u], 0.f);              // glow.jl:112
      lsum += logf(fabsf(S));                              // glow.jl:210
    }
    stv<V>(yp + pix, out);
  }
  if (ld) {
    double r = block_sum((double)lsum);
    if (threadIdx.x == 0) atomicAdd(ld, r * invB);
  }
}

template <int V>
__global__ void k_coupling_inv(const float* __restrict__ y1, long long ybs, float* __restrict__ x1,
                               long long xbs, const float* __restrict__ rb, long long px, int C1,
                               float low, float high) {
  const int ch = blockIdx.y;
  const long long b = blockIdx.z;
  const long long pxv = px / V;
  const float* yp = y1 + b * ybs + ch * px;
  const float* lp = rb + (b * 2 * C1 + ch) * px;
  const float* tp = rb + (b * 2 * C1 + C1 + ch) * px;
  float* xp = x1 + b * xbs + ch * px;
  for (long long pv = blockIdx.x * (long long)blockDim.x + threadIdx.x; pv < pxv; pv += (long long)gridDim.x * blockDim.x) {
    const long long pix = pv * V;
    float yv[V], ls[V], tv[V], out[V];
    ldv<V>(yp + pix, yv);
    ldv<V>(lp + pix, ls);
    ldv<V>(tp + pix, tv);
#pragma unroll
    for (int u = 0; u < V; ++u) {
      float S = sigmoid_lh(fmaxf(ls[u], 0.f), low, high);
      out[u] = (yv[u] - fmaxf(tv[u], 0.f)) / (S + 1.1920929e-07f);  // glow.jl:127, eps(Float32)
    }
    stv<V>(xp + pix, out);
  }
}

template <int V>
__global__ void k_coupling_bwd(const float* __restrict__ y1, long long ybs, float* __restrict__ x1,
                               long long xbs, const float* __restrict__ dy1, long long dybs,
                               float* __restrict__ dx1, long long dxbs, float* __restrict__ rb,
                               long long px, int C1, float low, float high, float invB, unsigned* __restrict__ amax) {
  const int ch = blockIdx.y;
  float gmax = 0.f;  // max |gradient of the block output| seen by this thread (INB_PREC_FP16X3 scales by it)
  const long long b = blockIdx.z;
  const long long pxv = px / V;
  const float* yp = y1 + b * ybs + ch * px;
  const float* dyp = dy1 + b * dybs + ch * px;
  float* plsb = rb + (b * 2 * C1 + ch) * px;
  float* ptvb = rb + (b * 2 * C1 + C1 + ch) * px;
  float* xp = x1 + b * xbs + ch * px;
  float* dxp = dx1 + b * dxbs + ch * px;
  for (long long pv = blockIdx.x * (long long)blockDim.x + threadIdx.x; pv < pxv; pv += (long long)gridDim.x * blockDim.x) {
    const long long pix = pv * V;
    float yv[V], dyv[V], ls[V], tv[V], xo[V], dxo[V], gl[V], gt[V];
    ldv<V>(yp + pix, yv);
    ldv<V>(dyp + pix, dyv);
    ldv<V>(plsb + pix, ls);
    ldv<V>(ptvb + pix, tv);
#pragma unroll
    for (int u = 0; u < V; ++u) {
      float S = sigmoid_lh(fmaxf(ls[u], 0.f), low, high);
      float X1 = (yv[u] - fmaxf(tv[u], 0.f)) / (S + 1.1920929e-07f);  // glow.jl:127
      float dS = dyv[u] * X1;                                         // glow.jl:144
      dS -= invB / S;                                                 // glow.jl:145-147,211 (invB = 0 w/o logdet)
      xo[u] = X1;
      dxo[u] = dyv[u] * S;                                            // glow.jl:149
      // activation_functions.jl:213-217: gradient from the output through the logit x = log(S - low) - log(high - S);
      // e^-x is the ratio itself, no logarithms needed
      float e = (high - S) / (S - low);
      float dl = (high - low) * dS * e / ((1.f + e) * (1.f + e));
      // _relugrad of the block's output ReLU (activation_functions.jl:84): pass when pre-activation >= 0
      gl[u] = (ls[u] < 0.f) ? 0.f : dl;
      gt[u] = (tv[u] < 0.f) ? 0.f : dyv[u];  // dT = dY1 (glow.jl:143)
      gmax = fmaxf(gmax, fmaxf(fabsf(gl[u]), fabsf(gt[u])));
    }
    stv<V>(xp + pix, xo);
    stv<V>(dxp + pix, dxo);
    stv<V>(plsb + pix, gl);
    stv<V>(ptvb + pix, gt);
  }
  if (amax) {  // bit patterns of non-negative floats are ordered like the values
#pragma unroll
    for (int o = 16; o; o >>= 1) gmax = fmaxf(gmax, __shfl_xor_sync(0xFFFFFFFFu, gmax, o));
    if ((threadIdx.x & 31) == 0 && gmax > 0.f) atomicMax(amax, __float_as_uint(gmax));
  }
}

// grid of the coupling kernels: (pixel vectors / 512, channel, sample) - two vectors per thread on large planes
static dim3 coupling_grid(long long pxv, int C1, int B) {
  INB_CHECK(C1 <= 65535 && B <= 65535, "coupling: channel / batch count exceeds the grid limits");
  return dim3((unsigned)std::max<long long>(1, cdiv(pxv, 512)), (unsigned)C1, (unsigned)B);
}
static int pick_vec_ew(long long px, std::initializer_list<const View*> vs, const float* rb) {
  int v = 4;
  while (v > 1) {
    bool ok = (px % v == 0) && (((uintptr_t)rb) % (4 * v) == 0);
    for (const View* w : vs) ok = ok && (w->bs % v == 0) && (((uintptr_t)w->p) % (4 * v) == 0);
    if (ok) break;
    v >>= 1;
  }
  return v;
}

void op_coupling_fwd(Ctx& c, long long px, int B, int C1, View x1, View y1, const float* rb, float low,
                     float high, double* ld, int ld_batch) {
  if (c.dry()) return;
  Prof pf(c, F_COUPLING_FWD, 1, 0, 16.0 * B * C1 * px);
  int V = pick_vec_ew(px, {&x1, &y1}, rb);
  const dim3 grid = coupling_grid(px / V, C1, B);
  float invB = 1.f / (float)(ld_batch > 0 ? ld_batch : B);
  if (V == 4) k_coupling_fwd<4><<<grid, 256, 0, c.st>>>(x1.p, x1.bs, y1.p, y1.bs, rb, px, C1, low, high, ld, invB);
  else if (V == 2) k_coupling_fwd<2><<<grid, 256, 0, c.st>>>(x1.p, x1.bs, y1.p, y1.bs, rb, px, C1, low, high, ld, invB);
  else k_coupling_fwd<1><<<grid, 256, 0, c.st>>>(x1.p, x1.bs, y1.p, y1.bs, rb, px, C1, low, high, ld, invB);
  INB_CUDA(cudaGetLastError());
}
void op_coupling_inv(Ctx& c, long long px, int B, int C1, View y1, View x1, const float* rb, float low,
                     float high) {
  if (c.dry()) return;
  Prof pf(c, F_COUPLING_INV, 1, 0, 16.0 * B * C1 * px);
  int V = pick_vec_ew(px, {&x1, &y1}, rb);
  const dim3 grid = coupling_grid(px / V, C1, B);
  if (V == 4) k_coupling_inv<4><<<grid, 256, 0, c.st>>>(y1.p, y1.bs, x1.p, x1.bs, rb, px, C1, low, high);
  else if (V == 2) k_coupling_inv<2><<<grid, 256, 0, c.st>>>(y1.p, y1.bs, x1.p, x1.bs, rb, px, C1, low, high);
  else k_coupling_inv<1><<<grid, 256, 0, c.st>>>(y1.p, y1.bs, x1.p, x1.bs, rb, px, C1, low, high);
  INB_CUDA(cudaGetLastError());
}
void op_coupling_bwd(Ctx& c, long long px, int B, int C1, View y1, View x1, View dy1, View dx1,
                     float* rb, float low, float high, int logdet, unsigned* amax) {
  if (c.dry()) return;
  Prof pf(c, F_COUPLING_BWD, 1, 0, 32.0 * B * C1 * px);
  int V = pick_vec_ew(px, {&x1, &y1, &dy1, &dx1}, rb);
  const dim3 grid = coupling_grid(px / V, C1, B);
  float invB = logdet ? 1.f / (float)B : 0.f;
  if (V == 4) k_coupling_bwd<4><<<grid, 256, 0, c.st>>>(y1.p, y1.bs, x1.p, x1.bs, dy1.p, dy1.bs, dx1.p, dx1.bs, rb, px, C1, low, high, invB, amax);
  else if (V == 2) k_coupling_bwd<2><<<grid, 256, 0, c.st>>>(y1.p, y1.bs, x1.p, x1.bs, dy1.p, dy1.bs, dx1.p, dx1.bs, rb, px, C1, low, high, invB, amax);
  else k_coupling_bwd<1><<<grid, 256, 0, c.st>>>(y1.p, y1.bs, x1.p, x1.bs, dy1.p, dy1.bs, dx1.p, dx1.bs, rb, px, C1, low, high, invB, amax);
  INB_CUDA(cudaGetLastError());
}

__global__ void k_relu_copy(const float* __restrict__ in, float* __restrict__ out, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    out[i] = fmaxf(in[i], 0.f);
}
void op_relu_copy(Ctx& c, long long n, const float* in, float* out) {
  if (c.dry()) return;
  Prof pf(c, F_MISC, 1, 0, 8.0 * n);
  k_relu_copy<<<grid_for(n, 256), 256, 0, c.st>>>(in, out, n);
  INB_CUDA(cudaGetLastError());
}

__global__ void k_relu_grad(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ out, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    out[i] = (y[i] < 0.f) ? 0.f : dy[i];
}
void op_relu_grad(Ctx& c, long long n, const float* dy, const float* y, float* out) {
  if (c.dry()) return;
  Prof pf(c, F_MISC, 1, 0, 12.0 * n);
  k_relu_grad<<<grid_for(n, 256), 256, 0, c.st>>>(dy, y, out, n);
  INB_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------- objective
__global__ void k_nll_grad(const float* __restrict__ z, float* __restrict__ dz, long long n, float invB,
                           double* __restrict__ acc) {
  float part = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    float v = z[i];
    part = fmaf(v, v, part);
    if (dz) dz[i] = v * invB;  // objective_functions.jl:65 with mu=0, sigma=1
  }
  if (acc) {
    double r = block_sum((double)part);
    if (threadIdx.x == 0) atomicAdd(acc, 0.5 * r * invB);  // objective_functions.jl:54 (negated)
  }
}
__global__ void k_ld_finish(const double* __restrict__ acc, float* __restrict__ out) { *out = (float)*acc; }

void op_nll_grad(Ctx& c, long long n, int B, const float* z, float* dz, double* acc, float* loss) {
  if (c.dry()) return;
  Prof pf(c, F_NLL, 2, 0, 8.0 * n);
  if (acc) INB_CUDA(cudaMemsetAsync(acc, 0, sizeof(double), c.st));
  k_nll_grad<<<grid_for(n, 256, 4), 256, 0, c.st>>>(z, dz, n, 1.f / (float)B, acc);
  if (acc && loss) k_ld_finish<<<1, 1, 0, c.st>>>(acc, loss);
  INB_CUDA(cudaGetLastError());
}
// Flux.Optimise.ADAM over a flat buffer (the `update!(opt, p.data, p.grad)` loop of examples/networks/network_glow.jl:38-42
// for all parameters in one launch):  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;
// x -= lr * (m / (1 - b1^t)) / (sqrt(v / (1 - b2^t)) + eps)
__global__ void k_adam(float* __restrict__ x, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                       long long n, float lr, float b1, float b2, float eps, float c1, float c2) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i];
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    x[i] -= lr * (mi * c1) / (sqrtf(vi * c2) + eps);
  }
}
void op_adam(Ctx& c, long long n, float* x, const float* g, float* m, float* v, float lr, float b1, float b2, float eps,
             float b1t, float b2t) {
  if (c.dry()) return;
  Prof pf(c, F_MISC, 1, 0, 28.0 * n);
  k_adam<<<grid_for(n, 256, 4), 256, 0, c.st>>>(x, g, m, v, n, lr, b1, b2, eps, 1.f / (1.f - b1t), 1.f / (1.f - b2t));
  INB_CUDA(cudaGetLastError());
}
void op_ld_finish(Ctx& c, const double* acc, float* out) {
  if (c.dry()) return;
  Prof pf(c, F_MISC, 1, 0, 0);
  k_ld_finish<<<1, 1, 0, c.st>>>(acc, out);
  INB_CUDA(cudaGetLastError());
}

}  // namespace inb
