// conv_simt.cu - INB_PREC_FP32: the ResidualBlock contractions as fp32 implicit GEMMs on the
// CUDA cores (exact fp32 FMA arithmetic; the parity yardstick for the tcgen05 path and the
// path used for tiny, latency-bound networks).
//
// Forward / dgrad:  out[m, o] = epi( sum_{tap,c} in[m + off(tap), c] * Wm[tap*Cin + c][o] )
//   m = global pixel index over (b, z, y, x); the A operand is gathered on the fly (im2col is
//   never materialised), coalesced along x.
// Wgrad:            dWm[tap*Cin + c][o] = sum_m in[m + off(tap), c] * dy[m, o]
#include "ops.cuh"
#include <algorithm>

namespace inb {

// ---------------------------------------------------------------- weight packing
__global__ void k_pack_w(int mode, int d0, int d1, int T, const float* __restrict__ w, float* __restrict__ Wm) {
  long long n = (long long)d0 * d1 * T;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    // i indexes Wm[tap][c][o]
    int Cc = mode == PACK_CONV ? d1 : d0;
    int O = mode == PACK_CONV ? d0 : d1;
    int o = (int)(i % O);
    long long t = i / O;
    int cc = (int)(t % Cc);
    int tap = (int)(t / Cc);
    float v;
    if (mode == PACK_CONV) v = w[((long long)o * d1 + cc) * T + (T - 1 - tap)];
    else v = w[((long long)cc * d1 + o) * T + tap];
    Wm[i] = v;
  }
}
void op_pack_w(Ctx& c, int mode, int d0, int d1, int T, const float* w, float* Wm) {
  if (c.dry()) return;
  Prof pf(c, F_PACK, 1, 0, 8.0 * d0 * d1 * T);
  long long n = (long long)d0 * d1 * T;
  int grid = (int)std::min<long long>(cdiv(n, 256), 148 * 8);
  k_pack_w<<<grid, 256, 0, c.st>>>(mode, d0, d1, T, w, Wm);
  INB_CUDA(cudaGetLastError());
}
__global__ void k_unpack_dw(int O, int Cc, int T, const float* __restrict__ dWm, float* __restrict__ dw) {
  long long n = (long long)O * Cc * T;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    // i indexes dw[o][c][t]
    int t = (int)(i % T);
    long long r = i / T;
    int cc = (int)(r % Cc);
    int o = (int)(r / Cc);
    dw[i] = dWm[((long long)(T - 1 - t) * Cc + cc) * O + o];
  }
}
void op_unpack_dw(Ctx& c, int O, int Cc, int T, const float* dWm, float* dw) {
  if (c.dry()) return;
  Prof pf(c, F_PACK, 1, 0, 8.0 * O * Cc * T);
  long long n = (long long)O * Cc * T;
  int grid = (int)std::min<long long>(cdiv(n, 256), 148 * 8);
  k_unpack_dw<<<grid, 256, 0, c.st>>>(O, Cc, T, dWm, dw);
  INB_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------- shared A-gather
struct GatherGeo {
  int W, H, D, k;
  long long px;
};
// decode the pixel once per row; returns false when the row is past the end
struct RowPos {
  long long b;
  int x, y, z;
  long long pix;
  bool live;
};
__device__ __forceinline__ RowPos decode_row(long long m, long long M, const GatherGeo& g) {
  RowPos r;
  r.live = m < M;
  long long mm = r.live ? m : 0;
  r.b = mm / g.px;
  r.pix = mm - r.b * g.px;
  long long t = r.pix;
  r.x = (int)(t % g.W); t /= g.W;
  r.y = (int)(t % g.H);
  r.z = (int)(t / g.H);
  return r;
}
// value of in[row + off(tap), c]  (zero padding), c indexes the concatenated input
__device__ __forceinline__ float gather(const RowPos& r, int kk, int K, const GatherGeo& g, int Cin,
                                        const float* __restrict__ in0, long long in0_bs, int c0,
                                        const float* __restrict__ in1, long long in1_bs, int relu_in) {
  if (!r.live || kk >= K) return 0.f;
  int tap = kk / Cin, c = kk - tap * Cin;
  long long off = 0;
  if (g.k == 3) {
    int dx = tap % 3 - 1, dy = (tap / 3) % 3 - 1, dz = (g.D > 1) ? tap / 9 - 1 : 0;
    int xx = r.x + dx, yy = r.y + dy, zz = r.z + dz;
    if (xx < 0 || xx >= g.W || yy < 0 || yy >= g.H || zz < 0 || zz >= g.D) return 0.f;
    off = ((long long)dz * g.H + dy) * g.W + dx;
  }
  float v = (c < c0) ? in0[r.b * in0_bs + (long long)c * g.px + r.pix + off]
                     : in1[r.b * in1_bs + (long long)(c - c0) * g.px + r.pix + off];
  return relu_in ? fmaxf(v, 0.f) : v;
}

// ---------------------------------------------------------------- forward / dgrad implicit GEMM
// CTA tile 128 pixels x (16*TN) channels, K step 8, 256 threads, thread tile 8 x TN.
template <int TN>
__global__ void __launch_bounds__(256)
k_conv_simt(ConvSpec s, long long M, int K) {
  constexpr int BM = 128, BN = 16 * TN, BK = 8;
  __shared__ __align__(16) float As[BK][BM];
  __shared__ __align__(16) float Bs[BK][BN];
  const GatherGeo g{s.g.W, s.g.H, s.g.D, s.k, s.g.px};
  const int tid = threadIdx.x;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  // A loader: thread owns row (tid % 128), k rows (tid / 128) + 2*r
  const int am = tid & 127, ak = tid >> 7;
  const RowPos arow = decode_row(m0 + am, M, g);
  // compute mapping: tm = tid % 16 (8 pixels each), tn = tid / 16 (TN channels each)
  const int tm = tid & 15, tn = tid >> 4;
  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float areg[4], breg[(BK * BN + 255) / 256];
  auto load_tile = [&](int k0) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
      areg[r] = gather(arow, k0 + ak + 2 * r, K, g, s.Cin, s.in0, s.in0_bs, s.c0, s.in1, s.in1_bs, s.relu_in);
#pragma unroll
    for (int r = 0; r < (BK * BN + 255) / 256; ++r) {
      int e = tid + r * 256;
      int kk = e / BN, n = e - kk * BN;
      breg[r] = (e < BK * BN && k0 + kk < K && n0 + n < s.N) ? s.Wm[(long long)(k0 + kk) * s.N + n0 + n] : 0.f;
    }
  };
  load_tile(0);
  for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
    for (int r = 0; r < 4; ++r) As[ak + 2 * r][am] = areg[r];
#pragma unroll
    for (int r = 0; r < (BK * BN + 255) / 256; ++r) {
      int e = tid + r * 256;
      if (e < BK * BN) Bs[e / BN][e % BN] = breg[r];
    }
    __syncthreads();
    if (k0 + BK < K) load_tile(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[8], b[TN];
      float4 a0 = *reinterpret_cast<const float4*>(&As[kk][tm * 8]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[kk][tm * 8 + 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tn * TN + j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  // epilogue: 8 consecutive pixels per thread -> still coalesced in 32-byte runs per channel
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    const int o = n0 + tn * TN + j;
    if (o >= s.N) continue;
    const float bias = s.bias ? s.bias[o] : 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const long long m = m0 + tm * 8 + i;
      if (m >= M) continue;
      const long long b = m / g.px, pix = m - b * g.px;
      float v = acc[i][j] + bias;
      if (s.add && o < s.add_n) {
        float a = s.add[b * s.add_bs + (long long)o * g.px + pix];
        v += s.add_relu ? fmaxf(a, 0.f) : a;
      }
      if (s.mask && s.mask[b * s.mask_bs + (long long)o * g.px + pix] < 0.f) v = 0.f;  // _relugrad
      if (o < s.n0) {
        s.out0[b * s.out0_bs + (long long)o * g.px + pix] = v;
      } else {
        float* q = s.out1 + b * s.out1_bs + (long long)(o - s.n0) * g.px + pix;
        *q = s.out1_accum ? (*q + v) : v;
      }
    }
  }
}

static int taps_of(const Geo& g, int k) {
  if (k == 1) return 1;
  return g.nd == 3 ? 27 : 9;
}

void op_conv_simt(Ctx& c, const ConvSpec& s) {
  INB_CHECK(s.k == 1 || s.k == 3, "ResidualBlock kernel size %d is not supported (1 or 3)", s.k);
  if (c.dry()) return;
  Prof pf(c, F_CONV_SIMT, 1, 2.0 * s.g.px * s.B * taps_of(s.g, s.k) * s.Cin * s.N, 4.0 * s.g.px * s.B * (s.Cin + s.N));
  const long long M = s.g.px * s.B;
  const int K = taps_of(s.g, s.k) * s.Cin;
  dim3 grid((unsigned)cdiv(M, 128), 1);
  if (s.N <= 16) { grid.y = 1; k_conv_simt<1><<<grid, 256, 0, c.st>>>(s, M, K); }
  else if (s.N <= 32) { grid.y = 1; k_conv_simt<2><<<grid, 256, 0, c.st>>>(s, M, K); }
  else if (s.N <= 48) { grid.y = 1; k_conv_simt<3><<<grid, 256, 0, c.st>>>(s, M, K); }
  else { grid.y = (unsigned)cdiv(s.N, 64); k_conv_simt<4><<<grid, 256, 0, c.st>>>(s, M, K); }
  INB_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------- wgrad
// CTA: 64 k-rows x 64 channels of dWm, reduction over a chunk of pixels, 32 pixels per step.
// grid (k tiles, n tiles, pixel chunks); partial sums are atomically added (dWm zeroed first).
__global__ void __launch_bounds__(256)
k_wgrad_simt(WgradSpec s, long long M, int K, long long chunk) {
  constexpr int BK = 64, BN = 64, BM = 32, LD = BK + 1;
  __shared__ float As[BM][LD];
  __shared__ float Bs[BM][LD];
  const GatherGeo g{s.g.W, s.g.H, s.g.D, s.k, s.g.px};
  const int tid = threadIdx.x;
  const int k0 = blockIdx.x * BK, n0 = blockIdx.y * BN;
  const long long mbeg = (long long)blockIdx.z * chunk;
  const long long mend = (mbeg + chunk < M) ? mbeg + chunk : M;
  const int lm = tid & 31, lr = tid >> 5;  // loader: pixel lm, rows lr + 8*r
  const int tk = tid >> 4, tn = tid & 15;  // compute: k rows tk*4.., channels tn*4..
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (long long mb = mbeg; mb < mend; mb += BM) {
    const long long m = mb + lm;
    const RowPos row = decode_row(m, mend, g);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      int kk = lr + 8 * r;
      As[lm][kk] = gather(row, k0 + kk, K, g, s.Cin, s.in0, s.in0_bs, s.c0, s.in1, s.in1_bs, s.relu_in);
      int n = n0 + kk;
      float v = 0.f;
      if (row.live && n < s.N) {
        v = s.dy[row.b * s.dy_bs + (long long)n * g.px + row.pix];
        if (s.relu_dy) v = fmaxf(v, 0.f);
      }
      Bs[lm][kk] = v;
    }
    __syncthreads();
#pragma unroll 8
    for (int mm = 0; mm < BM; ++mm) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[mm][tk * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[mm][tn * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int kk = k0 + tk * 4 + i, n = n0 + tn * 4 + j;
      if (kk < K && n < s.N) atomicAdd(s.dWm + (long long)kk * s.N + n, acc[i][j]);
    }
}

void op_wgrad_simt(Ctx& c, const WgradSpec& s) {
  INB_CHECK(s.k == 1 || s.k == 3, "ResidualBlock kernel size %d is not supported (1 or 3)", s.k);
  if (c.dry()) return;
  Prof pf(c, F_WGRAD_SIMT, 1, 2.0 * s.g.px * s.B * taps_of(s.g, s.k) * s.Cin * s.N, 4.0 * s.g.px * s.B * (s.Cin + s.N));
  const long long M = s.g.px * s.B;
  const int K = taps_of(s.g, s.k) * s.Cin;
  INB_CUDA(cudaMemsetAsync(s.dWm, 0, (size_t)K * s.N * sizeof(float), c.st));
  const int kt = (int)cdiv(K, 64), nt = (int)cdiv(s.N, 64);
  long long want = cdiv(148LL * 4, (long long)kt * nt);
  long long chunk = cdiv(cdiv(M, want), 32) * 32;
  if (chunk < 256) chunk = 256;
  const int zc = (int)cdiv(M, chunk);
  dim3 grid(kt, nt, zc);
  k_wgrad_simt<<<grid, 256, 0, c.st>>>(s, M, K, chunk);
  INB_CUDA(cudaGetLastError());
}

}  // namespace inb
