// conv_tc_fwd.cu - forward / dgrad implicit-GEMM convolution on tcgen05 (see conv_tc.cuh).
//
//   D[128 pixels][N] = sum_{tap, chunk} A_tap[128][ck] * W_tap[N][ck]^T        (x3: three hi/lo terms)
//
// Persistent, warp-specialised CTA (one per SM, 320 threads):
//   warp 0      TMA producer: per k-block one 5-D box of the activation planes shifted by the tap offset
//               (out-of-image rows are zero-filled by the TMA unit = zero padding) + one 2-D box of weights
//   warp 1      MMA issuer: tcgen05.mma M=128, N<=256, K=16 per instruction, fp32 accumulator in TMEM;
//               two accumulator buffers so the epilogue of tile i overlaps the main loop of tile i+1
//   warps 2..9  epilogue: tcgen05.ld (warp w reads TMEM lanes 32*(w%4).., two warps per lane quadrant
//               split the columns), bias from shared memory, relu-grad masking, ReLU + sign-bit mask
//               encoding, hi/lo bf16 split and 16-byte stores (mode 0) or coalesced fp32 NCHW stores with
//               the passthrough add (mode 1).
#include "conv_tc.cuh"
#include "tc_common.cuh"
#include "tc_maps.cuh"

#include <algorithm>

namespace inb {
using namespace tc;

struct ConvTcArgs {
  int taps, ksz, nchunks, ck, cpad;
  int W, H, D;
  long long M, px;
  int ntiles;
  int N, n_real, nsplit;  // columns [0,nsplit) go to the first epilogue warp of a quadrant, the rest to the second
  int stages;
  uint32_t tmem_cols, acc_stride;
  uint32_t a_bytes, b_bytes, b_tx;
  int mode;
  const float* bias;
  __nv_bfloat16 *out_hi, *out_lo;
  int out_pitch, relu_encode;
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

// 8 consecutive columns -> packed bf16 hi / lo words (bias, relu-grad mask, ReLU + sign-bit encoding)
__device__ __forceinline__ void pack8(const ConvTcArgs& a, const uint32_t* r, const float* sb, bool use_mask,
                                      uint4 mh, uint4& out_hi, uint4& out_lo) {
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[j]) + sb[j];
  if (use_mask) {
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
      if (a.relu_encode && x < 0.f) {  // relu(x) = 0, remembered as -0.0: the sign bit is the mask
        h[u] = __ushort_as_bfloat16(0x8000);
        l[u] = __ushort_as_bfloat16(0);
      } else {
        if (a.relu_encode) x = x + 0.f;  // -0.0 -> +0.0 (a pre-activation of exactly zero passes gradients)
        split_bf16(x, h[u], l[u]);
      }
    }
    ph[j] = pack2(h[0], h[1]);
    pl[j] = pack2(l[0], l[1]);
  }
  out_hi = make_uint4(ph[0], ph[1], ph[2], ph[3]);
  out_lo = make_uint4(pl[0], pl[1], pl[2], pl[3]);
}

// mode 0, 32 columns [c0, c0+32) of the warp's 32 rows: each thread packs its row, the warp transposes
// through a private 4 KB staging tile (16-byte chunks XOR-swizzled by the row -> conflict free both ways)
// so that every global store instruction writes 8 rows x 64 contiguous bytes (full 32-byte sectors).
__device__ __forceinline__ void epilogue_store32(const ConvTcArgs& a, const uint32_t (&r)[32], const uint4* msk,
                                                 bool use_mask, const float* sbias, int c0, long long m_warp,
                                                 uint8_t* stg, int lane) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    uint4 h, l;
    pack8(a, r + 8 * g, sbias + c0 + 8 * g, use_mask, use_mask ? msk[g] : make_uint4(0, 0, 0, 0), h, l);
    *reinterpret_cast<uint4*>(stg + lane * 128 + ((g ^ (lane & 7)) << 4)) = h;
    *reinterpret_cast<uint4*>(stg + lane * 128 + (((4 + g) ^ (lane & 7)) << 4)) = l;
  }
  __syncwarp();
  const int sub = lane & 3, rsel = lane >> 2;
#pragma unroll
  for (int rr = 0; rr < 4; ++rr) {
    const int row = rr + 4 * rsel;
    const long long m = m_warp + row;
#pragma unroll
    for (int pl = 0; pl < 2; ++pl) {
      const uint4 v = *reinterpret_cast<const uint4*>(stg + row * 128 + (((pl * 4 + sub) ^ (row & 7)) << 4));
      if (m < a.M) {
        __nv_bfloat16* dst = (pl ? a.out_lo : a.out_hi) + m * a.out_pitch + c0 + sub * 8;
        *reinterpret_cast<uint4*>(dst) = v;
      }
    }
  }
  __syncwarp();
}

// mode 1: NC accumulator columns of pixel row m -> fp32 (B, C, px): coalesced along the pixels
template <int NC>
__device__ __forceinline__ void epilogue_nchw(const ConvTcArgs& a, const uint32_t (&r)[NC], const float* sbias,
                                              int c0, long long m) {
  const long long b = m / a.px, pix = m - b * a.px;
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    const int n = c0 + j;
    if (n >= a.n_real) continue;
    float v = __uint_as_float(r[j]) + sbias[n];
    if (a.add && n < a.add_n) v += a.add[b * a.add_bs + (long long)n * a.px + pix];
    if (n < a.n0) {
      a.out0[b * a.out0_bs + (long long)n * a.px + pix] = v;
    } else {
      float* q = a.out1 + b * a.out1_bs + (long long)(n - a.n0) * a.px + pix;
      *q = a.out1_accum ? (*q + v) : v;
    }
  }
}

constexpr int kConvThreads = 320;

template <int NT>
__global__ void __launch_bounds__(kConvThreads, 1)
k_conv_tc(const __grid_constant__ CUtensorMap mA0, const __grid_constant__ CUtensorMap mA1,
          const __grid_constant__ CUtensorMap mB0, const __grid_constant__ CUtensorMap mB1, const ConvTcArgs a) {
  constexpr int NP = (NT == 1) ? 1 : 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t stage_bytes = NP * (a.a_bytes + a.b_bytes);
  uint64_t* full = reinterpret_cast<uint64_t*>(base + (size_t)a.stages * stage_bytes);
  uint64_t* empty = full + a.stages;
  uint64_t* tfull = empty + a.stages;   // [2]
  uint64_t* tempty = tfull + 2;         // [2]
  uint32_t* tslot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* sbias = reinterpret_cast<float*>(tslot + 4);  // [256]
  uint8_t* stg_all = reinterpret_cast<uint8_t*>(sbias + 256);  // 8 x 4 KB epilogue staging (16-byte aligned)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&mA0);
    prefetch_tmap(&mB0);
    if (NP == 2) { prefetch_tmap(&mA1); prefetch_tmap(&mB1); }
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < a.stages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
      for (int s = 0; s < 2; ++s) { mbar_init(tfull + s, 1); mbar_init(tempty + s, 8); }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tslot, a.tmem_cols);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < 256; i += blockDim.x) sbias[i] = (a.bias && i < a.n_real) ? a.bias[i] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tslot;
  const int nkb = a.taps * a.nchunks;

  if (warp == 0) {
    if (elect_one()) {
      const uint32_t tx = NP * (a.a_bytes + a.b_tx);
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
        long long t = (long long)tile * 128;
        const int x0 = (int)(t % a.W); t /= a.W;
        const int y0 = (int)(t % a.H); t /= a.H;
        const int z0 = (int)(t % a.D); t /= a.D;
        const int b0 = (int)t;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % a.stages;
          const uint32_t ph = (it / a.stages) & 1;
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
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t row_bytes = a.ck * 2;
      const uint32_t layout = layout_for_row(row_bytes);
      const uint32_t sbo = 8 * row_bytes;
      const uint32_t idesc = make_idesc_bf16(128, a.N, 0, 0);
      const int ksteps = a.ck / 16;
      uint32_t it = 0, tl = 0;
      for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++tl) {
        const uint32_t acc_i = tl & 1;
        mbar_wait(tempty + acc_i, ((tl >> 1) & 1) ^ 1);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem + acc_i * a.acc_stride;
        uint32_t acc = 0;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % a.stages;
          const uint32_t ph = (it / a.stages) & 1;
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
              umma_f16(d_tmem, ad, bd, idesc, acc);
              acc = 1;
            }
          }
          umma_commit(empty + s);  // frees the stage once these MMAs have read it
        }
        umma_commit(tfull + acc_i);
      }
    }
  } else {
    const int e = warp - 2;           // 0..7
    const int q = warp & 3;           // TMEM lane quadrant this warp may read
    const int half = e >> 2;          // which column range
    uint8_t* stg = stg_all + e * 4096;
    const int cbeg = half ? a.nsplit : 0, cend = half ? a.N : a.nsplit;
    uint32_t tl = 0;
    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++tl) {
      const uint32_t acc_i = tl & 1;
      const long long m = (long long)tile * 128 + q * 32 + lane;
      const bool live = m < a.M;
      mbar_wait(tfull + acc_i, (tl >> 1) & 1);
      tc_fence_after();
      const uint32_t tbase = tmem + acc_i * a.acc_stride + ((uint32_t)(q * 32) << 16);
      const long long m_warp = (long long)tile * 128 + q * 32;
      int c0 = cbeg;
      for (; c0 + 32 <= cend; c0 += 32) {
        uint4 msk[4];
        const bool use_mask = a.mode == 0 && a.mask_hi != nullptr;
        if (use_mask) {
          if (live) {
            const uint4* mp = reinterpret_cast<const uint4*>(a.mask_hi + m * a.mask_pitch + c0);
#pragma unroll
            for (int j = 0; j < 4; ++j) msk[j] = mp[j];
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) msk[j] = make_uint4(0, 0, 0, 0);
          }
        }
        uint32_t r[32];
        tmem_ld32(tbase + c0, r);
        tmem_ld_wait();
        if (a.mode == 0) epilogue_store32(a, r, msk, use_mask, sbias, c0, m_warp, stg, lane);
        else if (live) epilogue_nchw<32>(a, r, sbias, c0, m);
      }
      if (c0 < cend) {  // 16-column tail: only the fp32 NCHW outputs have N % 32 != 0
        uint32_t r[16];
        tmem_ld16(tbase + c0, r);
        tmem_ld_wait();
        if (live) epilogue_nchw<16>(a, r, sbias, c0, m);
      }
      // this warp is done reading the accumulator: let the MMA warp reuse it
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty + acc_i);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, a.tmem_cols);
}

static uint32_t pow2_cols(int n) {
  uint32_t c = 32;
  while ((int)c < n) c <<= 1;
  return c;
}

void op_conv_tc(Ctx& c, const ConvTcSpec& s) {
  INB_CHECK(s.k == 1 || s.k == 3, "ResidualBlock kernel size %d is not supported (1 or 3)", s.k);
  INB_CHECK(s.N % 16 == 0 && s.N >= 16 && s.N <= 256, "tensor-core conv: N=%d must be a multiple of 16 <= 256", s.N);
  INB_CHECK(s.cpad_in % 16 == 0, "tensor-core conv: padded input channels must be a multiple of 16");
  INB_CHECK(s.mode != 0 || s.N % 64 == 0, "tensor-core conv: plane outputs need N to be a multiple of 64");
  const TileBox tb = make_tile_box(s.g, s.B, 128);
  INB_CHECK(tb.ok, "spatial size %dx%dx%d cannot be tiled for the tensor-core path; use precision fp32", s.g.W,
            s.g.H, s.g.D);
  if (c.dry()) return;
  const int NT = prec_terms(c.prec);
  const int NP = NT == 1 ? 1 : 2;
  ConvTcArgs a{};
  a.taps = s.k == 1 ? 1 : (s.g.nd == 3 ? 27 : 9);
  a.ksz = s.k;
  a.cpad = s.cpad_in;
  a.W = s.g.W; a.H = s.g.H; a.D = s.g.D;
  a.px = s.g.px;
  a.M = s.g.px * s.B;
  a.ntiles = (int)cdiv(a.M, 128);
  a.N = s.N;
  a.n_real = s.n_real;
  a.nsplit = ((s.N / 16 + 1) / 2) * 16;
  a.acc_stride = pow2_cols(s.N);
  a.tmem_cols = 2 * a.acc_stride;
  // channel chunk: the largest of 64/32/16 that still leaves a >= 4-stage ring in ~200 KB
  const size_t budget = 188 * 1024;
  int ck = 16;
  for (int cand : {64, 32, 16}) {
    if (s.cpad_in % cand) continue;
    const size_t bb = ((size_t)s.N * cand * 2 + 1023) & ~size_t(1023);
    const size_t st = NP * ((size_t)128 * cand * 2 + bb);
    ck = cand;
    if (budget / st >= 4) break;
  }
  a.ck = ck;
  a.nchunks = s.cpad_in / ck;
  a.a_bytes = 128 * ck * 2;
  a.b_tx = s.N * ck * 2;
  a.b_bytes = (a.b_tx + 1023) & ~1023u;
  const uint32_t stage_bytes = NP * (a.a_bytes + a.b_bytes);
  int stages = (int)(budget / stage_bytes);
  if (stages > 8) stages = 8;
  if (stages < 2) stages = 2;
  a.stages = stages;
  a.mode = s.mode;
  a.bias = s.bias;
  a.out_hi = s.out.hi; a.out_lo = s.out.lo; a.out_pitch = s.out.pitch;
  a.relu_encode = s.relu_encode;
  a.mask_hi = s.mask.hi; a.mask_pitch = s.mask.pitch;
  a.out0 = s.out0; a.out0_bs = s.out0_bs; a.n0 = s.n0;
  a.out1 = s.out1; a.out1_bs = s.out1_bs; a.out1_accum = s.out1_accum;
  a.add = s.add; a.add_bs = s.add_bs; a.add_n = s.add_n;
  const size_t smem = (size_t)stages * stage_bytes + (2 * stages + 4) * 8 + 16 + 256 * 4 + 8 * 4096 + 1024;
  INB_CHECK(smem <= 227 * 1024, "tensor-core conv: shared memory %zu too large", smem);
  CUtensorMap mA0 = make_act_map(s.in.hi, s.in.pitch, s.g, s.B, ck, tb);
  CUtensorMap mA1 = make_act_map(s.in.lo, s.in.pitch, s.g, s.B, ck, tb);
  CUtensorMap mB0 = make_w_map(s.w.hi, a.taps * s.cpad_in, s.N, ck);
  CUtensorMap mB1 = make_w_map(s.w.lo, a.taps * s.cpad_in, s.N, ck);
  const unsigned grid = (unsigned)std::min(a.ntiles, 148);
  Prof pf(c, F_CONV_TC, 1, 2.0 * a.M * a.taps * s.cpad_in * s.N * NT, 0);
  if (NT == 3) {
    INB_CUDA(cudaFuncSetAttribute(k_conv_tc<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_conv_tc<3><<<grid, kConvThreads, smem, c.st>>>(mA0, mA1, mB0, mB1, a);
  } else {
    INB_CUDA(cudaFuncSetAttribute(k_conv_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_conv_tc<1><<<grid, kConvThreads, smem, c.st>>>(mA0, mA1, mB0, mB1, a);
  }
  INB_CUDA(cudaGetLastError());
}

}  // namespace inb
