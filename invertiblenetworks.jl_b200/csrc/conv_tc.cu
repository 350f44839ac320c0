// conv_tc.cu - implicit-GEMM convolutions and weight gradients on the 5th-generation tensor cores:
// TMA (cp.async.bulk.tensor) stages the operands in shared memory, one elected thread issues
// tcgen05.mma with the fp32 accumulator in TMEM, four epilogue warps read it back with tcgen05.ld.
//
// Forward / dgrad (k_conv_tc):  D[128 pixels][N] = sum_{tap, chunk} A_tap[128][ck] * W_tap[N][ck]^T
//   A_tap is a 5-D TMA box (ck channels, wt, ht, dt, bt) of the pixel-major activation shifted by the
//   tap offset; out-of-image elements are zero-filled by the TMA unit = the conv's zero padding.
// Wgrad (k_wgrad_tc):  D_tap[128 chan of P][cq] = sum_{pixels} P[pix][chan] * Q[pix + off(tap)][q]
//   both operands MN-major (pixel = K is the strided dimension), split over pixel ranges across CTAs,
//   partial sums added atomically straight into the reference's weight layout.
#include "conv_tc.cuh"
#include "tc_common.cuh"

#include <cuda.h>
#include <algorithm>
#include <mutex>

namespace inb {
using namespace tc;

// ---------------------------------------------------------------- tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  });
  if (!fn) fail(2, "cuTensorMapEncodeTiled is not available from the CUDA driver");
  return fn;
}
static CUtensorMapSwizzle swizzle_for(int row_bytes) {
  return row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}
// pixel-major activation planes [B][D][H][W][pitch] bf16, box (ck, wt, ht, dt, bt)
static CUtensorMap make_act_map(const __nv_bfloat16* base, int pitch, const Geo& g, int B, int ck, const TileBox& tb) {
  CUtensorMap m;
  cuuint64_t dims[5] = {(cuuint64_t)pitch, (cuuint64_t)g.W, (cuuint64_t)g.H, (cuuint64_t)g.D, (cuuint64_t)B};
  cuuint64_t str[4] = {(cuuint64_t)pitch * 2, (cuuint64_t)pitch * 2 * g.W, (cuuint64_t)pitch * 2 * g.W * g.H,
                       (cuuint64_t)pitch * 2 * g.W * g.H * g.D};
  cuuint32_t box[5] = {(cuuint32_t)ck, (cuuint32_t)tb.wt, (cuuint32_t)tb.ht, (cuuint32_t)tb.dt, (cuuint32_t)tb.bt};
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, (void*)base, dims, str, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(ck * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) fail(2, "cuTensorMapEncodeTiled(activation) failed with %d", (int)r);
  return m;
}
// weight planes [rows][ktot] bf16 (K-major rows), box (ck, rows)
static CUtensorMap make_w_map(const __nv_bfloat16* base, int ktot, int rows, int ck) {
  CUtensorMap m;
  cuuint64_t dims[2] = {(cuuint64_t)ktot, (cuuint64_t)rows};
  cuuint64_t str[1] = {(cuuint64_t)ktot * 2};
  cuuint32_t box[2] = {(cuuint32_t)ck, (cuuint32_t)rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)base, dims, str, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(ck * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) fail(2, "cuTensorMapEncodeTiled(weights) failed with %d", (int)r);
  return m;
}

TileBox make_tile_box(const Geo& g, int B, int P) {
  TileBox t{1, 1, 1, 1, false};
  int rem = P;
  if (g.W >= rem) {
    if (g.W % rem) return t;
    t.wt = rem;
    t.ok = true;
    return t;
  }
  if (rem % g.W) return t;
  t.wt = g.W;
  rem /= g.W;
  if (g.H >= rem) {
    if (g.H % rem) return t;
    t.ht = rem;
    t.ok = true;
    return t;
  }
  if (rem % g.H) return t;
  t.ht = g.H;
  rem /= g.H;
  if (g.D >= rem) {
    if (g.D % rem) return t;
    t.dt = rem;
    t.ok = true;
    return t;
  }
  if (rem % g.D) return t;
  t.dt = g.D;
  rem /= g.D;
  t.bt = rem;  // whole samples per tile; a ragged last tile is zero-filled by the TMA unit
  t.ok = rem <= 256;
  (void)B;
  return t;
}
bool tc_geometry_ok(const Geo& g, int B) { return make_tile_box(g, B, 128).ok && make_tile_box(g, B, 64).ok; }

// ---------------------------------------------------------------- layout / packing kernels
__global__ void k_nchw_to_tc(const float* __restrict__ in0, long long in0_bs, int c0, const float* __restrict__ in1,
                             long long in1_bs, int Cin, int cpad, long long px, long long M,
                             __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  const long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (m >= M) return;
  const long long b = m / px, pix = m - b * px;
  for (int cb = 0; cb < cpad; cb += 8) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      __nv_bfloat16 hh[2], ll[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        int ch = cb + 2 * j + u;
        float v = 0.f;
        if (ch < Cin)
          v = (ch < c0) ? in0[b * in0_bs + (long long)ch * px + pix] : in1[b * in1_bs + (long long)(ch - c0) * px + pix];
        split_bf16(v, hh[u], ll[u]);
      }
      h[j] = pack2(hh[0], hh[1]);
      l[j] = pack2(ll[0], ll[1]);
    }
    *reinterpret_cast<uint4*>(hi + m * cpad + cb) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(lo + m * cpad + cb) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}
void op_nchw_to_tc(Ctx& c, const Geo& g, int B, const float* in0, long long in0_bs, int c0, const float* in1,
                   long long in1_bs, int Cin, int cpad, Planes out) {
  if (c.dry()) return;
  const long long M = g.px * B;
  Prof pf(c, F_LAYOUT_TC, 1, 0, (4.0 * Cin + 4.0 * cpad) * M);
  k_nchw_to_tc<<<(unsigned)cdiv(M, 256), 256, 0, c.st>>>(in0, in0_bs, c0, in1, in1_bs, Cin, cpad, g.px, M, out.hi, out.lo);
  INB_CUDA(cudaGetLastError());
}

__global__ void k_pack_w_tc(int mode, int d0, int d1, int T, const float* __restrict__ w, int npad, int cpad,
                            __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  const long long n_el = (long long)npad * T * cpad;
  const int O = mode == PACK_CONV ? d0 : d1, Cc = mode == PACK_CONV ? d1 : d0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_el;
       i += (long long)gridDim.x * blockDim.x) {
    int cc = (int)(i % cpad);
    long long t = i / cpad;
    int tap = (int)(t % T);
    int n = (int)(t / T);
    float v = 0.f;
    if (n < O && cc < Cc) {
      if (mode == PACK_CONV) v = w[((long long)n * d1 + cc) * T + (T - 1 - tap)];
      else v = w[((long long)cc * d1 + n) * T + tap];
    }
    __nv_bfloat16 h, l;
    split_bf16(v, h, l);
    hi[i] = h;
    lo[i] = l;
  }
}
void op_pack_w_tc(Ctx& c, int mode, int d0, int d1, int T, const float* w, int npad, int cpad, Planes out) {
  if (c.dry()) return;
  const long long n = (long long)npad * T * cpad;
  Prof pf(c, F_PACK, 1, 0, 8.0 * n);
  k_pack_w_tc<<<(unsigned)std::min<long long>(cdiv(n, 256), 148 * 8), 256, 0, c.st>>>(mode, d0, d1, T, w, npad, cpad,
                                                                                    out.hi, out.lo);
  INB_CUDA(cudaGetLastError());
}

// per-channel sums of [M][C] planes: thread = (row lane, channel pair)
__global__ void k_colsum_tc(const uint32_t* __restrict__ hi, const uint32_t* __restrict__ lo, long long M, int C2,
                            long long rows_per_block, float* __restrict__ out) {
  const int cp = threadIdx.x % C2;
  const int rl = threadIdx.x / C2, nrl = blockDim.x / C2;
  if (rl >= nrl) return;
  const long long r0 = blockIdx.x * rows_per_block;
  const long long r1 = (r0 + rows_per_block < M) ? r0 + rows_per_block : M;
  float s0 = 0.f, s1 = 0.f;
  for (long long r = r0 + rl; r < r1; r += nrl) {
    uint32_t h = hi[r * C2 + cp], l = lo[r * C2 + cp];
    s0 += bf16lo_to_f(h) + bf16lo_to_f(l);
    s1 += bf16hi_to_f(h) + bf16hi_to_f(l);
  }
  atomicAdd(out + 2 * cp, s0);
  atomicAdd(out + 2 * cp + 1, s1);
}
void op_colsum_tc(Ctx& c, long long M, int C, Planes in, float* out) {
  if (c.dry()) return;
  INB_CHECK(C % 2 == 0 && C / 2 <= 256 && in.pitch == C, "colsum: unsupported channel count %d", C);
  Prof pf(c, F_CHANNEL_SUM, 2, 0, 4.0 * M * C);
  INB_CUDA(cudaMemsetAsync(out, 0, C * sizeof(float), c.st));
  const int C2 = C / 2;
  const int threads = (256 / C2) * C2 > 0 ? (256 / C2) * C2 : C2;
  long long blocks = std::min<long long>(cdiv(M, 64), 148 * 4);
  long long rpb = cdiv(M, blocks);
  k_colsum_tc<<<(unsigned)cdiv(M, rpb), threads, 0, c.st>>>((const uint32_t*)in.hi, (const uint32_t*)in.lo, M, C2, rpb, out);
  INB_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------- forward / dgrad kernel
struct ConvTcArgs {
  int taps, ksz, nchunks, ck, cpad;
  int W, H, D;
  long long M, px;
  int N, n_real;
  int stages;
  uint32_t tmem_cols;
  uint32_t a_bytes, b_bytes, b_tx;  // tile sizes in shared memory (b rounded to 1 KB) and B's TMA bytes
  int mode;
  const float* bias;
  __nv_bfloat16 *out_hi, *out_lo;
  int out_pitch, relu_encode;
  const __nv_bfloat16 *skip_hi, *skip_lo;
  int skip_pitch;
  const __nv_bfloat16* mask_hi;
  int mask_pitch;
  float* out0; long long out0_bs; int n0;
  float* out1; long long out1_bs; int out1_accum;
  const float* add; long long add_bs; int add_n;
};

__device__ __forceinline__ void tap_offset(int tap, int ksz, int D, int& dx, int& dy, int& dz) {
  if (ksz == 1) { dx = dy = dz = 0; return; }
  dx = tap % 3 - 1;
  dy = (tap / 3) % 3 - 1;
  dz = (D > 1) ? tap / 9 - 1 : 0;
}

// epilogue helper: NC (16 or 32) accumulator columns [c0, c0+NC) of pixel row m
template <int NC>
__device__ __forceinline__ void conv_epilogue_cols(const ConvTcArgs& a, const uint32_t (&r)[NC], int c0, long long m) {
  if (m >= a.M) return;
  if (a.mode == 0) {
#pragma unroll
    for (int j0 = 0; j0 < NC; j0 += 8) {
      const int n = c0 + j0;
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[j0 + j]) + (a.bias ? __ldg(a.bias + n + j) : 0.f);
      if (a.skip_hi) {
        uint4 sh = *reinterpret_cast<const uint4*>(a.skip_hi + m * a.skip_pitch + n);
        uint4 sl = *reinterpret_cast<const uint4*>(a.skip_lo + m * a.skip_pitch + n);
        const uint32_t hh[4] = {sh.x, sh.y, sh.z, sh.w}, ll[4] = {sl.x, sl.y, sl.z, sl.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          v[2 * j] += bf16lo_to_f(hh[j]) + bf16lo_to_f(ll[j]);
          v[2 * j + 1] += bf16hi_to_f(hh[j]) + bf16hi_to_f(ll[j]);
        }
      }
      if (a.mask_hi) {
        uint4 mh = *reinterpret_cast<const uint4*>(a.mask_hi + m * a.mask_pitch + n);
        const uint32_t mm[4] = {mh.x, mh.y, mh.z, mh.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (mm[j] & 0x00008000u) v[2 * j] = 0.f;
          if (mm[j] & 0x80000000u) v[2 * j + 1] = 0.f;
        }
      }
      uint32_t ph[4], pl[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        __nv_bfloat16 h[2], l[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          float x = v[2 * j + u];
          if (a.relu_encode) {
            if (x < 0.f) { h[u] = __ushort_as_bfloat16(0x8000); l[u] = __ushort_as_bfloat16(0); continue; }
            if (x == 0.f) x = 0.f;  // +0
          }
          split_bf16(x, h[u], l[u]);
          if (a.relu_encode && __bfloat16_as_ushort(h[u]) == 0x8000) h[u] = __ushort_as_bfloat16(0);
        }
        ph[j] = pack2(h[0], h[1]);
        pl[j] = pack2(l[0], l[1]);
      }
      *reinterpret_cast<uint4*>(a.out_hi + m * a.out_pitch + n) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
      *reinterpret_cast<uint4*>(a.out_lo + m * a.out_pitch + n) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
    }
  } else {
    const long long b = m / a.px, pix = m - b * a.px;
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      const int n = c0 + j;
      if (n >= a.n_real) continue;
      float v = __uint_as_float(r[j]) + (a.bias ? __ldg(a.bias + n) : 0.f);
      if (a.add && n < a.add_n) v += a.add[b * a.add_bs + (long long)n * a.px + pix];
      if (n < a.n0) {
        a.out0[b * a.out0_bs + (long long)n * a.px + pix] = v;
      } else {
        float* q = a.out1 + b * a.out1_bs + (long long)(n - a.n0) * a.px + pix;
        *q = a.out1_accum ? (*q + v) : v;
      }
    }
  }
}

// NT = 1 (bf16) or 3 (bf16x3).  192 threads: warp 0 TMA producer, warp 1 MMA issuer (+TMEM alloc),
// warps 2..5 epilogue (warp w owns TMEM lanes 32*(w%4)..+31).
template <int NT>
__global__ void __launch_bounds__(192)
k_conv_tc(const __grid_constant__ CUtensorMap mA0, const __grid_constant__ CUtensorMap mA1,
          const __grid_constant__ CUtensorMap mB0, const __grid_constant__ CUtensorMap mB1, const ConvTcArgs a) {
  constexpr int NP = (NT == 1) ? 1 : 2;  // operand planes staged
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t stage_bytes = NP * (a.a_bytes + a.b_bytes);
  uint64_t* full = reinterpret_cast<uint64_t*>(base + (size_t)a.stages * stage_bytes);
  uint64_t* empty = full + a.stages;
  uint64_t* tfull = empty + a.stages;
  uint32_t* tslot = reinterpret_cast<uint32_t*>(tfull + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&mA0);
    prefetch_tmap(&mB0);
    if (NP == 2) { prefetch_tmap(&mA1); prefetch_tmap(&mB1); }
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < a.stages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
      mbar_init(tfull, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tslot, a.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tslot;

  const int nkb = a.taps * a.nchunks;
  const long long m0 = (long long)blockIdx.x * 128;

  if (warp == 0) {
    if (elect_one()) {
      long long t = m0;
      const int x0 = (int)(t % a.W); t /= a.W;
      const int y0 = (int)(t % a.H); t /= a.H;
      const int z0 = (int)(t % a.D); t /= a.D;
      const int b0 = (int)t;
      const uint32_t tx = NP * (a.a_bytes + a.b_tx);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % a.stages;
        const uint32_t ph = (kb / a.stages) & 1;
        mbar_wait(empty + s, ph ^ 1);
        mbar_expect_tx(full + s, tx);
        const int tap = kb / a.nchunks, chunk = kb - tap * a.nchunks;
        int dx, dy, dz;
        tap_offset(tap, a.ksz, a.D, dx, dy, dz);
        uint8_t* sa = base + (size_t)s * stage_bytes;
        uint8_t* sb = sa + NP * a.a_bytes;
        tma_load_5d(&mA0, full + s, sa, chunk * a.ck, x0 + dx, y0 + dy, z0 + dz, b0);
        tma_load_2d(&mB0, full + s, sb, tap * a.cpad + chunk * a.ck, 0);
        if (NP == 2) {
          tma_load_5d(&mA1, full + s, sa + a.a_bytes, chunk * a.ck, x0 + dx, y0 + dy, z0 + dz, b0);
          tma_load_2d(&mB1, full + s, sb + a.b_bytes, tap * a.cpad + chunk * a.ck, 0);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t row_bytes = a.ck * 2;
      const uint32_t layout = layout_for_row(row_bytes);
      const uint32_t sbo = 8 * row_bytes;
      const uint32_t idesc = make_idesc_bf16(128, a.N, 0, 0);
      const int ksteps = a.ck / 16;
      uint32_t acc = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % a.stages;
        const uint32_t ph = (kb / a.stages) & 1;
        mbar_wait(full + s, ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(base + (size_t)s * stage_bytes);
        const uint32_t sb = sa + NP * a.a_bytes;
#pragma unroll
        for (int term = 0; term < NT; ++term) {
          // terms: (a_hi,b_hi), (a_hi,b_lo), (a_lo,b_hi)
          const uint32_t ta = sa + ((term == 2) ? a.a_bytes : 0);
          const uint32_t tb = sb + ((term == 1) ? a.b_bytes : 0);
          for (int k = 0; k < ksteps; ++k) {
            const uint64_t ad = make_smem_desc(ta + k * 32, 0, sbo, layout);
            const uint64_t bd = make_smem_desc(tb + k * 32, 0, sbo, layout);
            umma_f16(tmem, ad, bd, idesc, acc);
            acc = 1;
          }
        }
        umma_commit(empty + s);  // frees the stage once these MMAs have read it
      }
      umma_commit(tfull);
    }
  } else {
    mbar_wait(tfull, 0);
    tc_fence_after();
    const int q = warp & 3;
    const long long m = m0 + q * 32 + lane;
    const uint32_t tbase = tmem + ((uint32_t)(q * 32) << 16);
    int c0 = 0;
    for (; c0 + 32 <= a.N; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(tbase + c0, r);
      tmem_ld_wait();
      conv_epilogue_cols<32>(a, r, c0, m);
    }
    if (c0 < a.N) {
      uint32_t r[16];
      tmem_ld16(tbase + c0, r);
      tmem_ld_wait();
      conv_epilogue_cols<16>(a, r, c0, m);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, a.tmem_cols);
}

static uint32_t tmem_cols_for(int n) {
  uint32_t c = 32;
  while ((int)c < n) c <<= 1;
  return c;
}
static int pick_ck(int cpad) { return (cpad % 64 == 0) ? 64 : ((cpad % 32 == 0) ? 32 : 16); }

void op_conv_tc(Ctx& c, const ConvTcSpec& s) {
  INB_CHECK(s.k == 1 || s.k == 3, "ResidualBlock kernel size %d is not supported (1 or 3)", s.k);
  INB_CHECK(s.N % 16 == 0 && s.N >= 16 && s.N <= 256, "tensor-core conv: N=%d must be a multiple of 16 <= 256", s.N);
  INB_CHECK(s.cpad_in % 16 == 0, "tensor-core conv: padded input channels must be a multiple of 16");
  const TileBox tb = make_tile_box(s.g, s.B, 128);
  INB_CHECK(tb.ok, "spatial size %dx%dx%d cannot be tiled for the tensor-core path; use precision fp32", s.g.W,
            s.g.H, s.g.D);
  if (c.dry()) return;
  const int NT = (c.prec == 1) ? 3 : 1;
  const int NP = NT == 1 ? 1 : 2;
  ConvTcArgs a{};
  a.taps = s.k == 1 ? 1 : (s.g.nd == 3 ? 27 : 9);
  a.ksz = s.k;
  a.ck = pick_ck(s.cpad_in);
  a.nchunks = s.cpad_in / a.ck;
  a.cpad = s.cpad_in;
  a.W = s.g.W; a.H = s.g.H; a.D = s.g.D;
  a.px = s.g.px;
  a.M = s.g.px * s.B;
  a.N = s.N;
  a.n_real = s.n_real;
  a.tmem_cols = tmem_cols_for(s.N);
  a.a_bytes = 128 * a.ck * 2;
  a.b_tx = s.N * a.ck * 2;
  a.b_bytes = (a.b_tx + 1023) & ~1023u;
  const uint32_t stage_bytes = NP * (a.a_bytes + a.b_bytes);
  // aim at two resident CTAs per SM (their epilogues overlap each other's main loops)
  int stages = (int)((100 * 1024) / stage_bytes);
  if (stages < 2) stages = 2;
  if (stages > 6) stages = 6;
  const int nkb = a.taps * a.nchunks;
  if (stages > nkb) stages = nkb < 1 ? 1 : nkb;
  a.stages = stages;
  a.mode = s.mode;
  a.bias = s.bias;
  a.out_hi = s.out.hi; a.out_lo = s.out.lo; a.out_pitch = s.out.pitch;
  a.relu_encode = s.relu_encode;
  a.skip_hi = s.skip.hi; a.skip_lo = s.skip.lo; a.skip_pitch = s.skip.pitch;
  a.mask_hi = s.mask.hi; a.mask_pitch = s.mask.pitch;
  a.out0 = s.out0; a.out0_bs = s.out0_bs; a.n0 = s.n0;
  a.out1 = s.out1; a.out1_bs = s.out1_bs; a.out1_accum = s.out1_accum;
  a.add = s.add; a.add_bs = s.add_bs; a.add_n = s.add_n;
  const size_t smem = (size_t)stages * stage_bytes + (2 * stages + 1) * 8 + 16 + 1024;
  INB_CHECK(smem <= 227 * 1024, "tensor-core conv: shared memory %zu too large", smem);
  CUtensorMap mA0 = make_act_map(s.in.hi, s.in.pitch, s.g, s.B, a.ck, tb);
  CUtensorMap mA1 = make_act_map(s.in.lo, s.in.pitch, s.g, s.B, a.ck, tb);
  CUtensorMap mB0 = make_w_map(s.w.hi, a.taps * s.cpad_in, s.N, a.ck);
  CUtensorMap mB1 = make_w_map(s.w.lo, a.taps * s.cpad_in, s.N, a.ck);
  const unsigned grid = (unsigned)cdiv(a.M, 128);
  Prof pf(c, F_CONV_TC, 1, 2.0 * a.M * a.taps * s.cpad_in * s.N * NT, 0);
  if (NT == 3) {
    INB_CUDA(cudaFuncSetAttribute(k_conv_tc<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_conv_tc<3><<<grid, 192, smem, c.st>>>(mA0, mA1, mB0, mB1, a);
  } else {
    INB_CUDA(cudaFuncSetAttribute(k_conv_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_conv_tc<1><<<grid, 192, smem, c.st>>>(mA0, mA1, mB0, mB1, a);
  }
  INB_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------- wgrad kernel
struct WgradTcArgs {
  int ksz, T, tap0, ntap;   // taps [tap0, tap0+ntap) handled by this launch slice (blockIdx.z selects)
  int W, H, D;
  long long M;
  int nblocks, blocks_per_cta;  // 64-pixel blocks
  int cq, qa, nqa, cq_real;
  int stages;
  uint32_t tmem_cols;
  uint32_t p_bytes, q_tap_bytes;  // per plane: P tile (128 ch x 64 px), Q tile of one tap
  int taps_per_group;
  float* dw;
};

template <int NT>
__global__ void __launch_bounds__(192)
k_wgrad_tc(const __grid_constant__ CUtensorMap mP0, const __grid_constant__ CUtensorMap mP1,
           const __grid_constant__ CUtensorMap mQ0, const __grid_constant__ CUtensorMap mQ1, const WgradTcArgs a) {
  constexpr int NP = (NT == 1) ? 1 : 2;
  constexpr int PB = 64;  // pixels per k-block
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tap_begin = blockIdx.z * a.taps_per_group;
  const int ntap = min(a.taps_per_group, a.T - tap_begin);
  const uint32_t stage_bytes = NP * (a.p_bytes + a.taps_per_group * a.q_tap_bytes);
  uint64_t* full = reinterpret_cast<uint64_t*>(base + (size_t)a.stages * stage_bytes);
  uint64_t* empty = full + a.stages;
  uint64_t* tfull = empty + a.stages;
  uint32_t* tslot = reinterpret_cast<uint32_t*>(tfull + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&mP0);
    prefetch_tmap(&mQ0);
    if (NP == 2) { prefetch_tmap(&mP1); prefetch_tmap(&mQ1); }
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < a.stages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
      mbar_init(tfull, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tslot, a.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tslot;

  const int blk0 = blockIdx.x * a.blocks_per_cta;
  const int blk1 = min(blk0 + a.blocks_per_cta, a.nblocks);
  const int nkb = max(blk1 - blk0, 0);
  const int pch0 = blockIdx.y * 128;

  if (warp == 0) {
    if (elect_one()) {
      const uint32_t q_rows_bytes = PB * a.qa * 2;  // one channel atom of one tap
      const uint32_t tx = NP * (a.p_bytes + ntap * a.nqa * q_rows_bytes);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % a.stages;
        const uint32_t ph = (kb / a.stages) & 1;
        mbar_wait(empty + s, ph ^ 1);
        mbar_expect_tx(full + s, tx);
        long long t = (long long)(blk0 + kb) * PB;
        const int x0 = (int)(t % a.W); t /= a.W;
        const int y0 = (int)(t % a.H); t /= a.H;
        const int z0 = (int)(t % a.D); t /= a.D;
        const int b0 = (int)t;
        uint8_t* sp = base + (size_t)s * stage_bytes;
        uint8_t* sq = sp + NP * a.p_bytes;
        for (int pl = 0; pl < NP; ++pl) {
          const CUtensorMap* mp = pl ? &mP1 : &mP0;
          const CUtensorMap* mq = pl ? &mQ1 : &mQ0;
          tma_load_5d(mp, full + s, sp + pl * a.p_bytes, pch0, x0, y0, z0, b0);
          tma_load_5d(mp, full + s, sp + pl * a.p_bytes + PB * 128, pch0 + 64, x0, y0, z0, b0);
          for (int tp = 0; tp < ntap; ++tp) {
            int dx, dy, dz;
            tap_offset(tap_begin + tp, a.ksz, a.D, dx, dy, dz);
            uint8_t* dst = sq + (size_t)(pl * a.taps_per_group + tp) * a.q_tap_bytes;
            for (int qa_i = 0; qa_i < a.nqa; ++qa_i)
              tma_load_5d(mq, full + s, dst + (size_t)qa_i * q_rows_bytes, qa_i * a.qa, x0 + dx, y0 + dy, z0 + dz, b0);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t q_row = a.qa * 2;
      const uint32_t q_layout = layout_for_row(q_row);
      const uint32_t idesc = make_idesc_bf16(128, a.cq, 1, 1);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % a.stages;
        const uint32_t ph = (kb / a.stages) & 1;
        mbar_wait(full + s, ph);
        tc_fence_after();
        const uint32_t sp = smem_u32(base + (size_t)s * stage_bytes);
        const uint32_t sq = sp + NP * a.p_bytes;
        for (int tp = 0; tp < ntap; ++tp) {
#pragma unroll
          for (int term = 0; term < NT; ++term) {
            const uint32_t tp_ = sp + ((term == 2) ? a.p_bytes : 0);
            const uint32_t tq_ = sq + (uint32_t)(((term == 1) ? a.taps_per_group : 0) + tp) * a.q_tap_bytes;
#pragma unroll
            for (int k = 0; k < PB / 16; ++k) {
              // MN-major operands: 16 pixels (K) = 16 rows; LBO = stride between channel atoms, SBO = 8 rows
              const uint64_t ad = make_smem_desc(tp_ + k * 16 * 128, PB * 128, 8 * 128, LAYOUT_SW128);
              const uint64_t bd = make_smem_desc(tq_ + k * 16 * q_row, PB * q_row, 8 * q_row, q_layout);
              umma_f16(tmem + tp * a.cq, ad, bd, idesc, (kb > 0 || term > 0 || k > 0) ? 1u : 0u);
            }
          }
        }
        umma_commit(empty + s);
      }
      umma_commit(tfull);
    }
  } else if (nkb > 0) {
    mbar_wait(tfull, 0);
    tc_fence_after();
    const int q = warp & 3;
    const int p = pch0 + q * 32 + lane;
    const uint32_t tbase = tmem + ((uint32_t)(q * 32) << 16);
    for (int tp = 0; tp < ntap; ++tp) {
      const int tap = tap_begin + tp;
      float* dst = a.dw + (long long)p * a.cq_real * a.T + (a.T - 1 - tap);
      for (int c0 = 0; c0 < a.cq; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(tbase + tp * a.cq + c0, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (c0 + j < a.cq_real) atomicAdd(dst + (long long)(c0 + j) * a.T, __uint_as_float(r[j]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, a.tmem_cols);
}

void op_wgrad_tc(Ctx& c, const WgradTcSpec& s) {
  INB_CHECK(s.k == 1 || s.k == 3, "ResidualBlock kernel size %d is not supported (1 or 3)", s.k);
  INB_CHECK(s.np % 128 == 0, "tensor-core wgrad needs n_hidden to be a multiple of 128 (got %d)", s.np);
  INB_CHECK(s.cq % 16 == 0 && s.cq <= 256, "tensor-core wgrad: bad channel count %d", s.cq);
  const TileBox tb = make_tile_box(s.g, s.B, 64);
  INB_CHECK(tb.ok, "spatial size %dx%dx%d cannot be tiled for the tensor-core path; use precision fp32", s.g.W,
            s.g.H, s.g.D);
  if (c.dry()) return;
  const int NT = (c.prec == 1) ? 3 : 1;
  const int NP = NT == 1 ? 1 : 2;
  const int T = s.k == 1 ? 1 : (s.g.nd == 3 ? 27 : 9);
  WgradTcArgs a{};
  a.ksz = s.k;
  a.T = T;
  a.W = s.g.W; a.H = s.g.H; a.D = s.g.D;
  a.M = s.g.px * s.B;
  a.nblocks = (int)cdiv(a.M, 64);
  a.cq = s.cq;
  a.qa = pick_ck(s.cq);
  a.nqa = s.cq / a.qa;
  a.cq_real = s.cq_real;
  a.p_bytes = 128 * 64 * 2;
  a.q_tap_bytes = 64 * s.cq * 2;
  // taps per CTA: TMEM (512 columns) and a >= 2-stage pipeline within shared memory
  int tpg = std::min(T, 512 / s.cq);
  while (tpg > 1 && 2 * NP * (a.p_bytes + tpg * a.q_tap_bytes) > 200 * 1024) --tpg;
  a.taps_per_group = tpg;
  const int groups = (int)cdiv(T, tpg);
  const uint32_t stage_bytes = NP * (a.p_bytes + tpg * a.q_tap_bytes);
  int stages = (int)((200 * 1024) / stage_bytes);
  if (stages > 4) stages = 4;
  INB_CHECK(stages >= 1, "tensor-core wgrad: stage of %u bytes does not fit", stage_bytes);
  a.stages = stages;
  a.tmem_cols = tmem_cols_for(tpg * s.cq);
  a.dw = s.dw;
  const int halves = s.np / 128;
  long long want = std::max<long long>(1, (148LL * 1) / ((long long)halves * groups));
  a.blocks_per_cta = (int)cdiv(a.nblocks, want);
  if (a.blocks_per_cta < 4) a.blocks_per_cta = std::min(4, a.nblocks);
  const unsigned gx = (unsigned)cdiv(a.nblocks, a.blocks_per_cta);
  const size_t smem = (size_t)stages * stage_bytes + (2 * stages + 1) * 8 + 16 + 1024;
  INB_CHECK(smem <= 227 * 1024, "tensor-core wgrad: shared memory %zu too large", smem);
  CUtensorMap mP0 = make_act_map(s.P.hi, s.P.pitch, s.g, s.B, 64, tb);
  CUtensorMap mP1 = make_act_map(s.P.lo, s.P.pitch, s.g, s.B, 64, tb);
  CUtensorMap mQ0 = make_act_map(s.Q.hi, s.Q.pitch, s.g, s.B, a.qa, tb);
  CUtensorMap mQ1 = make_act_map(s.Q.lo, s.Q.pitch, s.g, s.B, a.qa, tb);
  Prof pf(c, F_WGRAD_TC, 2, 2.0 * a.M * T * s.cq * s.np * NT, 0);
  INB_CUDA(cudaMemsetAsync(s.dw, 0, (size_t)s.np * s.cq_real * T * sizeof(float), c.st));
  dim3 grid(gx, halves, groups);
  if (NT == 3) {
    INB_CUDA(cudaFuncSetAttribute(k_wgrad_tc<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_wgrad_tc<3><<<grid, 192, smem, c.st>>>(mP0, mP1, mQ0, mQ1, a);
  } else {
    INB_CUDA(cudaFuncSetAttribute(k_wgrad_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_wgrad_tc<1><<<grid, 192, smem, c.st>>>(mP0, mP1, mQ0, mQ1, a);
  }
  INB_CUDA(cudaGetLastError());
}

}  // namespace inb
