// conv_tc_chain.cu - the three contractions of a ResidualBlock pass as ONE persistent tcgen05 kernel.
//
// Forward  (layer_residual_block.jl:122-129):  X -> relu(conv(X,W1)+b1) -> relu((W2+I) X2 + b2) -> \nabla conv_data(X3, W3)
// Backward (layer_residual_block.jl:151-162):  dY3 -> relugrad(conv(dY3,W3),Y2) -> relugrad((W2+I)^T dY2,Y1) -> \nabla conv_data(dY1, W1)
//
// Both are  im2col-GEMM (K = taps*16k)  ->  per-pixel GEMM (K = nh)  ->  per-pixel GEMM with TAP-EXPANDED
// output columns (N = taps*Cn) followed by a col2im gather (k_col2im): \nabla conv_data IS "GEMM, then
// col2im", so the 3x3 stencil of the last contraction costs one read of its 256-channel operand instead
// of nine, and every hidden tensor stays on the SM:
//
//   GEMM1: A = nine 5-D TMA boxes of the (padded, bf16 hi/lo) input shifted by the tap (zero fill = padding),
//          D1[128 x nh] in TMEM columns R0
//   E1   : tcgen05.ld -> +bias, ReLU (sign bit of -0.0 keeps the _relugrad mask) | relu-grad masking ->
//          bf16 hi/lo -> shared memory in the K-major SWIZZLE_128B operand layout, 64-channel chunk by chunk
//   GEMM2: A = those chunks as they become ready, B = (W2 + I) streamed through the TMA ring, D2 in R1
//   E2   : same as E1, overwrites the chunks
//   GEMM3: A = chunks, B = tap-expanded weights, D3[128 x taps*Cn] in R0 (+ R1 for > 256 columns)
//   E3   : D3 -> fp32 rows of P[M][taps*Cn] in HBM (the only per-pixel output; col2im finishes it)
//
// The two TMEM halves swap roles every tile, so GEMM1 of tile t+1 runs under E3 of tile t.  When the pass
// is the recompute/backward one, a store warp writes the hidden chunks to HBM with TMA (bulk tensor store
// straight from the operand layout) for the weight-gradient kernels; the plain forward writes nothing but P.
//
// Warps: 0 TMA producer | 1 MMA issuer | 2..9 epilogue (two per TMEM lane quadrant) | 10 TMA store.
#include "conv_tc.cuh"
#include "tc_common.cuh"
#include "tc_maps.cuh"

#include <algorithm>

namespace inb {
using namespace tc;

struct ChainMaps {
  CUtensorMap A[2], W1[2], W2[2], W3a[2], W3b[2], O1[2], O2[2];
};

struct ChainArgs {
  int W, H, D;
  long long M;
  int ntiles;
  int taps1, ksz1, nch1;  // GEMM1 k-blocks: taps1 x nch1 blocks of 16 channels
  int nh, nchunk;         // hidden channels (128 | 256), nh / 64
  int n3a, n3b, n3pad;    // GEMM3 column parts (multiples of 16, <= 256) and the pitch of P
  int stages;
  int mode;               // 0: bias + ReLU (forward)   1: relu-grad masks (backward)
  int store;              // write both hidden tensors to HBM
  const float *bias1, *bias2;
  const __nv_bfloat16 *mask1, *mask2;  // hi planes [M][nh] whose sign bits are the masks of E1 / E2
  float* P;
};

constexpr int kChainThreads = 352;
constexpr uint32_t kPlane = 16384;      // one plane of a 128 x 64 chunk / of a weight k-block
constexpr uint32_t kATap = 4096;        // 128 pixels x 16 channels

__device__ __forceinline__ void chain_tap_offset(int tap, int ksz, int D, int& dx, int& dy, int& dz) {
  if (ksz == 1) { dx = dy = dz = 0; return; }
  dx = tap % 3 - 1;
  dy = (tap / 3) % 3 - 1;
  dz = (D > 1) ? tap / 9 - 1 : 0;
}

// 8 accumulator columns -> packed bf16 hi / lo words
template <int MODE, int NT>
__device__ __forceinline__ void chain_pack8(const uint32_t* r, const float* sb, uint4 mh, uint4& oh, uint4& ol) {
  const uint32_t mm[4] = {mh.x, mh.y, mh.z, mh.w};
  uint32_t ph[4], pl[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float a = __uint_as_float(r[2 * j]), b = __uint_as_float(r[2 * j + 1]);
    uint32_t sign = 0;
    if (MODE == 0) {
      a += sb[2 * j];
      b += sb[2 * j + 1];
      // relu(x) = 0 is stored as -0.0 when x < 0: the sign bit of the hi plane is the _relugrad mask
      sign = ((__float_as_uint(a) >> 16) & 0x8000u) | (__float_as_uint(b) & 0x80000000u);
      a = fmaxf(a, 0.f);
      b = fmaxf(b, 0.f);
    } else {
      if (mm[j] & 0x00008000u) a = 0.f;
      if (mm[j] & 0x80000000u) b = 0.f;
    }
    __nv_bfloat162 h2 = __floats2bfloat162_rn(a, b);
    uint32_t h = *reinterpret_cast<uint32_t*>(&h2);
    if (NT == 3) {
      const float ra = a - __uint_as_float(h << 16), rb = b - __uint_as_float(h & 0xFFFF0000u);
      __nv_bfloat162 l2 = __floats2bfloat162_rn(ra, rb);
      pl[j] = *reinterpret_cast<uint32_t*>(&l2);
    } else {
      pl[j] = 0;
    }
    ph[j] = h | sign;
  }
  oh = make_uint4(ph[0], ph[1], ph[2], ph[3]);
  ol = make_uint4(pl[0], pl[1], pl[2], pl[3]);
}

template <int NT>
__global__ void __launch_bounds__(kChainThreads, 1)
k_rb_chain(const __grid_constant__ ChainMaps maps, const ChainArgs a) {
  constexpr int NP = (NT == 1) ? 1 : 2;
  constexpr uint32_t CHUNK = NP * kPlane;        // hi (+ lo) of one 128 x 64 chunk
  constexpr uint32_t STAGE = NP * kPlane;        // one ring stage
  constexpr uint32_t W1OFF = NP * kATap;         // weights of a GEMM1 k-block follow the A planes
  constexpr uint32_t W1PLANE = (NT == 1) ? 8192 : 8192;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* hbuf = smem;
  uint8_t* ring = hbuf + (size_t)a.nchunk * CHUNK;
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + (size_t)a.stages * STAGE);
  uint64_t* empty = full + 8;
  uint64_t* dfull = empty + 8;    // [3]
  uint64_t* hready = dfull + 3;   // [4]
  uint64_t* stdone = hready + 4;  // [4]
  uint64_t* e3done = stdone + 4;  // [1]
  uint32_t* tslot = reinterpret_cast<uint32_t*>(e3done + 1);
  float* sbias = reinterpret_cast<float*>(tslot + 2);  // [2][256]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    prefetch_tmap(&maps.A[0]);
    prefetch_tmap(&maps.W1[0]);
    prefetch_tmap(&maps.W2[0]);
    prefetch_tmap(&maps.W3a[0]);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < a.stages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
      for (int s = 0; s < 3; ++s) mbar_init(dfull + s, 1);
      for (int s = 0; s < 4; ++s) { mbar_init(hready + s, 256); mbar_init(stdone + s, 1); }
      mbar_init(e3done, 8);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tslot, 512);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < 512; i += blockDim.x) {
    const float* bp = (i < 256) ? a.bias1 : a.bias2;
    const int j = i & 255;
    sbias[i] = (bp && j < a.nh) ? bp[j] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tslot;
  const int nkb1 = a.taps1 * a.nch1;
  const int nkb2 = a.nh / 32;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (elect_one()) {
      uint32_t it = 0;
      auto acquire = [&](uint32_t tx) -> uint8_t* {
        const int s = it % a.stages;
        const uint32_t ph = (it / a.stages) & 1;
        mbar_wait(empty + s, ph ^ 1);
        mbar_expect_tx(full + s, tx);
        return ring + (size_t)s * STAGE;
      };
      for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
        long long t = (long long)tile * 128;
        const int x0 = (int)(t % a.W); t /= a.W;
        const int y0 = (int)(t % a.H); t /= a.H;
        const int z0 = (int)(t % a.D); t /= a.D;
        const int b0 = (int)t;
        for (int kb = 0; kb < nkb1; ++kb, ++it) {
          const int tap = kb / a.nch1, ch = kb - tap * a.nch1;
          int dx, dy, dz;
          chain_tap_offset(tap, a.ksz1, a.D, dx, dy, dz);
          uint8_t* st = acquire(NP * (kATap + a.nh * 32));
          uint64_t* fb = full + it % a.stages;
#pragma unroll
          for (int pl = 0; pl < NP; ++pl) {
            tma_load_5d(&maps.A[pl], fb, st + pl * kATap, ch * 16, x0 + dx, y0 + dy, z0 + dz, b0);
            tma_load_2d(&maps.W1[pl], fb, st + W1OFF + pl * W1PLANE, kb * 16, 0);
          }
        }
        for (int kb = 0; kb < nkb2; ++kb, ++it) {
          uint8_t* st = acquire(NP * a.nh * 64);
          uint64_t* fb = full + it % a.stages;
#pragma unroll
          for (int pl = 0; pl < NP; ++pl) tma_load_2d(&maps.W2[pl], fb, st + pl * kPlane, kb * 32, 0);
        }
        for (int kb = 0; kb < nkb2; ++kb, ++it) {
          uint8_t* st = acquire(NP * a.n3a * 64);
          uint64_t* fb = full + it % a.stages;
#pragma unroll
          for (int pl = 0; pl < NP; ++pl) tma_load_2d(&maps.W3a[pl], fb, st + pl * kPlane, kb * 32, 0);
        }
        if (a.n3b) {
          for (int kb = 0; kb < nkb2; ++kb, ++it) {
            uint8_t* st = acquire(NP * a.n3b * 64);
            uint64_t* fb = full + it % a.stages;
#pragma unroll
            for (int pl = 0; pl < NP; ++pl) tma_load_2d(&maps.W3b[pl], fb, st + pl * kPlane, kb * 32, a.n3a);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      const uint32_t idesc_h = make_idesc_bf16(128, a.nh, 0, 0);
      const uint32_t idesc_a = make_idesc_bf16(128, a.n3a, 0, 0);
      const uint32_t idesc_b = make_idesc_bf16(128, a.n3b ? a.n3b : 16, 0, 0);
      const uint32_t hb = smem_u32(hbuf);
      uint32_t it = 0, tl = 0;
      auto stage_wait = [&]() -> uint32_t {
        const int s = it % a.stages;
        const uint32_t ph = (it / a.stages) & 1;
        mbar_wait(full + s, ph);
        tc_fence_after();
        return smem_u32(ring + (size_t)s * STAGE);
      };
      // one K=32 weight block against k-steps (2j, 2j+1) of chunk c
      auto chunk_block = [&](uint32_t d_tmem, uint32_t idesc, int c, int j, uint32_t sa, uint32_t& acc) {
#pragma unroll
        for (int term = 0; term < NT; ++term) {
          const uint32_t ta = hb + c * CHUNK + ((term == 2) ? kPlane : 0);
          const uint32_t tb = sa + ((term == 1) ? kPlane : 0);
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const uint64_t ad = make_smem_desc(ta + (2 * j + k) * 32, 0, 1024, LAYOUT_SW128);
            const uint64_t bd = make_smem_desc(tb + k * 32, 0, 512, LAYOUT_SW64);
            umma_f16(d_tmem, ad, bd, idesc, acc);
            acc = 1;
          }
        }
      };
      for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++tl) {
        const uint32_t R0 = tmem + (tl & 1) * 256, R1 = tmem + ((tl & 1) ^ 1) * 256;
        if (a.n3b && tl > 0) {  // D3b of the previous tile lives where D1 goes
          mbar_wait(e3done, (tl - 1) & 1);
          tc_fence_after();
        }
        uint32_t acc = 0;
        for (int kb = 0; kb < nkb1; ++kb, ++it) {
          const uint32_t sa = stage_wait();
#pragma unroll
          for (int term = 0; term < NT; ++term) {
            const uint64_t ad = make_smem_desc(sa + ((term == 2) ? kATap : 0), 0, 256, LAYOUT_SW32);
            const uint64_t bd = make_smem_desc(sa + W1OFF + ((term == 1) ? W1PLANE : 0), 0, 256, LAYOUT_SW32);
            umma_f16(R0, ad, bd, idesc_h, acc);
            acc = 1;
          }
          umma_commit(empty + it % a.stages);
        }
        umma_commit(dfull + 0);
        acc = 0;
        for (int c = 0; c < a.nchunk; ++c) {
          mbar_wait(hready + c, 0);
          tc_fence_after();
          for (int j = 0; j < 2; ++j, ++it) {
            const uint32_t sa = stage_wait();
            chunk_block(R1, idesc_h, c, j, sa, acc);
            umma_commit(empty + it % a.stages);
          }
        }
        umma_commit(dfull + 1);
        acc = 0;
        for (int c = 0; c < a.nchunk; ++c) {
          mbar_wait(hready + c, 1);
          tc_fence_after();
          for (int j = 0; j < 2; ++j, ++it) {
            const uint32_t sa = stage_wait();
            chunk_block(R0, idesc_a, c, j, sa, acc);
            umma_commit(empty + it % a.stages);
          }
        }
        if (a.n3b) {
          acc = 0;
          for (int c = 0; c < a.nchunk; ++c) {
            for (int j = 0; j < 2; ++j, ++it) {
              const uint32_t sa = stage_wait();
              chunk_block(R1, idesc_b, c, j, sa, acc);
              umma_commit(empty + it % a.stages);
            }
          }
        }
        umma_commit(dfull + 2);
      }
    }
  } else if (warp < 10) {
    // ------------------------------------------------------------ epilogue warps
    const int e = warp - 2;
    const int q = warp & 3;   // TMEM lane quadrant this warp may read
    const int half = e >> 2;  // which 32 columns of a 64-channel chunk
    const int row = q * 32 + lane;
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    uint32_t tl = 0;
    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++tl) {
      const uint32_t R0 = tmem + (tl & 1) * 256, R1 = tmem + ((tl & 1) ^ 1) * 256;
      const long long m = (long long)tile * 128 + row;
      const bool live = m < a.M;
#pragma unroll 1
      for (int stg = 0; stg < 2; ++stg) {
        const uint32_t dsrc = (stg ? R1 : R0) + lane_sel;
        const float* sb = sbias + stg * 256;
        const __nv_bfloat16* mk = stg ? a.mask2 : a.mask1;
        mbar_wait(dfull + stg, tl & 1);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < a.nchunk; ++c) {
          const int col = 64 * c + 32 * half;
          uint4 msk[4];
          if (a.mode == 1) {
            if (live) {
              const uint4* mp = reinterpret_cast<const uint4*>(mk + m * a.nh + col);
#pragma unroll
              for (int j = 0; j < 4; ++j) msk[j] = __ldg(mp + j);
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j) msk[j] = make_uint4(0, 0, 0, 0);
            }
          }
          uint32_t r[32];
          tmem_ld32(dsrc + col, r);
          if (a.store && (tl > 0 || stg > 0)) mbar_wait(stdone + c, stg ^ 1);  // the chunk's previous store has read it
          tmem_ld_wait();
          uint8_t* dst = hbuf + (size_t)c * CHUNK + row * 128;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 oh, ol;
            if (a.mode == 0) chain_pack8<0, NT>(r + 8 * g, sb + col + 8 * g, make_uint4(0, 0, 0, 0), oh, ol);
            else chain_pack8<1, NT>(r + 8 * g, sb, msk[g], oh, ol);
            const uint32_t off = (uint32_t)(((half * 4 + g) ^ (row & 7)) << 4);
            *reinterpret_cast<uint4*>(dst + off) = oh;
            if (NT == 3) *reinterpret_cast<uint4*>(dst + kPlane + off) = ol;
          }
          fence_proxy_async();  // generic-proxy writes -> visible to the MMA / TMA (async proxy)
          tc_fence_before();
          mbar_arrive(hready + c);
        }
      }
      // E3: tap-expanded columns -> P
      mbar_wait(dfull + 2, tl & 1);
      tc_fence_after();
#pragma unroll 1
      for (int part = 0; part < 2; ++part) {
        const int n = part ? a.n3b : a.n3a;
        if (n == 0) break;
        const uint32_t dsrc = (part ? R1 : R0) + lane_sel;
        const int nsplit = ((n / 16 + 1) / 2) * 16;
        const int cbeg = half ? nsplit : 0, cend = half ? n : nsplit;
        float* prow = a.P + m * a.n3pad + (part ? a.n3a : 0);
        int c0 = cbeg;
        for (; c0 + 32 <= cend; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(dsrc + c0, r);
          tmem_ld_wait();
          if (live) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<uint4*>(prow + c0 + 4 * j) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
          }
        }
        if (c0 < cend) {
          uint32_t r[16];
          tmem_ld16(dsrc + c0, r);
          tmem_ld_wait();
          if (live) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              *reinterpret_cast<uint4*>(prow + c0 + 4 * j) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(e3done);
    }
  } else if (a.store) {
    // ------------------------------------------------------------ TMA store warp
    if (lane == 0) {
      for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
        for (int stg = 0; stg < 2; ++stg) {
          for (int c = 0; c < a.nchunk; ++c) {
            mbar_wait(hready + c, stg);
#pragma unroll
            for (int pl = 0; pl < NP; ++pl)
              tma_store_2d(stg ? &maps.O2[pl] : &maps.O1[pl], hbuf + (size_t)c * CHUNK + pl * kPlane, 64 * c, tile * 128);
            bulk_commit();
            bulk_wait_read0();
            mbar_arrive(stdone + c);
          }
        }
      }
      bulk_wait0();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------- col2im
// out[b][n][pix] = sum_tap P[pix + off(tap)][tap*Cn + n]  (+ passthrough add), coalesced both ways through a
// shared-memory transpose: phase 1 walks (pixel, n) with n fastest (contiguous in P), phase 2 walks pixels.
struct Col2imArgs {
  int W, H, D, ksz, taps, Cn, n3pad;
  long long px, M;
  const float* P;
  float* out0; long long out0_bs; int n0;
  float* out1; long long out1_bs; int out1_accum;
  const float* add; long long add_bs; int add_n;
};
constexpr int kC2iPix = 64;

__global__ void __launch_bounds__(256) k_col2im(const Col2imArgs a) {
  extern __shared__ float c2i_s[];  // [Cn][kC2iPix + 1]
  const long long m0 = (long long)blockIdx.x * kC2iPix;
  const int tot = kC2iPix * a.Cn;
  for (int i = threadIdx.x; i < tot; i += blockDim.x) {
    const int p = i / a.Cn, n = i - p * a.Cn;
    const long long m = m0 + p;
    float acc = 0.f;
    if (m < a.M) {
      long long t = m;
      const int x = (int)(t % a.W); t /= a.W;
      const int y = (int)(t % a.H); t /= a.H;
      const int z = (int)(t % a.D);
      for (int tap = 0; tap < a.taps; ++tap) {
        int dx, dy, dz;
        chain_tap_offset(tap, a.ksz, a.D, dx, dy, dz);
        const int xx = x + dx, yy = y + dy, zz = z + dz;
        if (xx < 0 || xx >= a.W || yy < 0 || yy >= a.H || zz < 0 || zz >= a.D) continue;
        const long long mm = m + dx + (long long)dy * a.W + (long long)dz * a.W * a.H;
        acc += __ldg(a.P + mm * a.n3pad + tap * a.Cn + n);
      }
    }
    c2i_s[n * (kC2iPix + 1) + p] = acc;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < tot; i += blockDim.x) {
    const int n = i / kC2iPix, p = i - n * kC2iPix;
    const long long m = m0 + p;
    if (m >= a.M) continue;
    const long long b = m / a.px, pix = m - b * a.px;
    float v = c2i_s[n * (kC2iPix + 1) + p];
    if (a.add && n < a.add_n) v += a.add[b * a.add_bs + (long long)n * a.px + pix];
    if (n < a.n0) {
      a.out0[b * a.out0_bs + (long long)n * a.px + pix] = v;
    } else {
      float* qq = a.out1 + b * a.out1_bs + (long long)(n - a.n0) * a.px + pix;
      *qq = a.out1_accum ? (*qq + v) : v;
    }
  }
}

// ---------------------------------------------------------------- tap-expanded weight packing
// rows r = tap*Cn + n (zero rows up to n3pad), K = nh columns:  Wexp[r][c] = coefficient of hidden
// channel c in tap `tap` of output channel n of the \nabla conv_data contraction (PACK_DATA order of
// op_pack_w_tc: w[c][n][tap] for the reference weight w[d0 = nh][d1 = Cn][T]).
__global__ void k_pack_wexp_tc(int nh, int Cn, int T, const float* __restrict__ w, int n3pad,
                               __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  const long long n_el = (long long)n3pad * nh;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_el;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % nh);
    const int r = (int)(i / nh);
    float v = 0.f;
    if (r < T * Cn) {
      const int tap = r / Cn, n = r - tap * Cn;
      v = w[((long long)c * Cn + n) * T + tap];
    }
    __nv_bfloat16 h, l;
    split_bf16(v, h, l);
    hi[i] = h;
    lo[i] = l;
  }
}
void op_pack_wexp_tc(Ctx& c, int nh, int Cn, int T, const float* w, int n3pad, Planes out) {
  if (c.dry()) return;
  const long long n = (long long)n3pad * nh;
  Prof pf(c, F_PACK, 1, 0, 8.0 * n);
  k_pack_wexp_tc<<<(unsigned)std::min<long long>(cdiv(n, 256), 148 * 8), 256, 0, c.st>>>(nh, Cn, T, w, n3pad, out.hi, out.lo);
  INB_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------- host side
int chain_n3pad(int taps, int Cn) { return (taps * Cn + 15) / 16 * 16; }

bool chain_supported(const Geo& g, int B, int k1, int k2, int nh, int c_in_pad, int Cn) {
  if (k2 != 1) return false;
  if (nh != 128 && nh != 256) return false;
  if (c_in_pad % 16 || c_in_pad > 256) return false;
  const int taps = k1 == 1 ? 1 : (g.nd == 3 ? 27 : 9);
  const int n3 = chain_n3pad(taps, Cn);
  if (n3 > 480 || Cn > 128) return false;
  return make_tile_box(g, B, 128).ok;
}

void op_rb_chain(Ctx& c, const ChainSpec& s) {
  const int taps = s.k1 == 1 ? 1 : (s.g.nd == 3 ? 27 : 9);
  INB_CHECK(chain_supported(s.g, s.B, s.k1, 1, s.nh, s.in.pitch, s.Cn), "fused ResidualBlock chain: unsupported shape");
  const TileBox tb = make_tile_box(s.g, s.B, 128);
  if (c.dry()) return;
  const int NT = (c.prec == 1) ? 3 : 1;
  const int NP = NT == 1 ? 1 : 2;
  ChainArgs a{};
  a.W = s.g.W; a.H = s.g.H; a.D = s.g.D;
  a.M = s.g.px * s.B;
  a.ntiles = (int)cdiv(a.M, 128);
  a.taps1 = taps;
  a.ksz1 = s.k1;
  a.nch1 = s.in.pitch / 16;
  a.nh = s.nh;
  a.nchunk = s.nh / 64;
  a.n3pad = chain_n3pad(taps, s.Cn);
  if (a.n3pad <= 256) { a.n3a = a.n3pad; a.n3b = 0; }
  else { a.n3a = (a.n3pad / 16 + 1) / 2 * 16; a.n3b = a.n3pad - a.n3a; }
  a.mode = s.mode;
  a.store = (s.o1.hi != nullptr) ? 1 : 0;
  a.bias1 = s.bias1; a.bias2 = s.bias2;
  a.mask1 = s.mask1.hi; a.mask2 = s.mask2.hi;
  a.P = s.P;
  const size_t stage = (size_t)NP * kPlane, chunk = (size_t)NP * kPlane;
  const size_t aux = 32 * 8 + 16 + 512 * 4;
  const size_t cap = 227 * 1024;
  int stages = (int)((cap - aux - a.nchunk * chunk) / stage);
  if (stages > 8) stages = 8;
  INB_CHECK(stages >= 2, "fused ResidualBlock chain: shared memory does not fit");
  a.stages = stages;
  const size_t smem = a.nchunk * chunk + stages * stage + aux;
  ChainMaps mp{};
  for (int pl = 0; pl < 2; ++pl) {
    const __nv_bfloat16* in = pl ? s.in.lo : s.in.hi;
    mp.A[pl] = make_act_map(in, s.in.pitch, s.g, s.B, 16, tb);
    mp.W1[pl] = make_w_map(pl ? s.w1.lo : s.w1.hi, taps * s.in.pitch, s.nh, 16);
    mp.W2[pl] = make_w_map(pl ? s.w2.lo : s.w2.hi, s.nh, s.nh, 32);
    mp.W3a[pl] = make_rows_map(pl ? s.w3.lo : s.w3.hi, s.nh, a.n3pad, 32, a.n3a);
    mp.W3b[pl] = make_rows_map(pl ? s.w3.lo : s.w3.hi, s.nh, a.n3pad, 32, a.n3b ? a.n3b : 16);
    if (a.store) {
      mp.O1[pl] = make_rows_map(pl ? s.o1.lo : s.o1.hi, s.nh, a.M, 64, 128);
      mp.O2[pl] = make_rows_map(pl ? s.o2.lo : s.o2.hi, s.nh, a.M, 64, 128);
    } else {
      mp.O1[pl] = mp.W2[pl];
      mp.O2[pl] = mp.W2[pl];
    }
  }
  const unsigned grid = (unsigned)std::min(a.ntiles, 148);
  const double flops = 2.0 * a.M * ((double)taps * s.in.pitch * s.nh + (double)s.nh * s.nh + (double)s.nh * a.n3pad) * NT;
  {
    Prof pf(c, F_CONV_TC, 1, flops, 0);
    if (NT == 3) {
      INB_CUDA(cudaFuncSetAttribute(k_rb_chain<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_rb_chain<3><<<grid, kChainThreads, smem, c.st>>>(mp, a);
    } else {
      INB_CUDA(cudaFuncSetAttribute(k_rb_chain<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_rb_chain<1><<<grid, kChainThreads, smem, c.st>>>(mp, a);
    }
    INB_CUDA(cudaGetLastError());
  }
  {
    Col2imArgs ca{};
    ca.W = s.g.W; ca.H = s.g.H; ca.D = s.g.D; ca.ksz = s.k1; ca.taps = taps; ca.Cn = s.Cn; ca.n3pad = a.n3pad;
    ca.px = s.g.px; ca.M = a.M; ca.P = s.P;
    ca.out0 = s.out0; ca.out0_bs = s.out0_bs; ca.n0 = s.n0;
    ca.out1 = s.out1; ca.out1_bs = s.out1_bs; ca.out1_accum = s.out1_accum;
    ca.add = s.add; ca.add_bs = s.add_bs; ca.add_n = s.add_n;
    Prof pf(c, F_COL2IM, 1, 0, (4.0 * a.n3pad + 4.0 * s.Cn) * a.M);
    const size_t sm = (size_t)s.Cn * (kC2iPix + 1) * sizeof(float);
    k_col2im<<<(unsigned)cdiv(a.M, kC2iPix), 256, sm, c.st>>>(ca);
    INB_CUDA(cudaGetLastError());
  }
}

}  // namespace inb
