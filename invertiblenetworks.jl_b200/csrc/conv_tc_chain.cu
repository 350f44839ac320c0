// conv_tc_chain.cu - the three contractions of a ResidualBlock pass as ONE persistent tcgen05 kernel.
//
// Forward  (layer_residual_block.jl:122-129):  X -> relu(conv(X,W1)+b1) -> relu((W2+I) X2 + b2) -> \nabla conv_data(X3, W3)
// Backward (layer_residual_block.jl:151-162):  dY3 -> relugrad(conv(dY3,W3),Y2) -> relugrad((W2+I)^T dY2,Y1) -> \nabla conv_data(dY1, W1)
//
// Both are  im2col-GEMM (K = taps*C)  ->  per-pixel GEMM (K = nh)  ->  per-pixel GEMM with TAP-EXPANDED
// output columns (N = taps*Cn) followed by a col2im gather (k_col2im): \nabla conv_data IS "GEMM, then
// col2im", so the 3x3 stencil of the last contraction costs one read of its 256-channel operand instead
// of nine, and every hidden tensor stays on the SM:
//
//   GEMM1: A = im2col rows [M][taps*C -> 64k] of the block input (bf16 hi/lo, k_im2col_tc), D1[128 x nh] in
//          TMEM columns R0
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
#include <cstdlib>

namespace inb {
using namespace tc;

struct ChainMaps {
  CUtensorMap A[2], W1[2], W2[2], W3[2], O1[2], O2[2], P;
};

struct ChainArgs {
  int W, H, D;
  long long M;
  int ntiles;
  int nkb1;               // GEMM1 k-blocks of 64 im2col columns
  int nh, nchunk;         // hidden channels (128 | 256), nh / 64
  int nh_bias;            // entries of the bias vectors (the block's n_hidden; channels beyond it are padding)
  int n3a, n3b, n3pad;    // GEMM3 columns in R0 / R1 and the pitch of P (multiple of 16)
  int np3;                // 128-column pieces of GEMM3
  int stages;
  // k_rb_chain2 with qsum: E3 sums the three horizontal taps of every (dy[,dz]) tap row on the SM, so P shrinks from
  // taps*Cn to (taps/3)*cq columns per pixel (Pq[m][tg*cq + n], cq = Cn rounded up to 4, pitch nq = (taps/3)*cq)
  int qsum, Cn, cq, nq, ntg;   // qsum: 1 = whole tile staged at once (n3pad <= 128), 2 = one tap row at a time (<= 256)
  int rpw;                // qsum == 2: row pitch (floats) of the per-tap-row staging: 32 * chunks + 4
  int pstag_bytes;        // k_rb_chain2: size of the P staging area
  int lo8;                // k_rb_chain2, store: the lo planes are written as one byte per value (Planes::lo8)
  int hints;              // k_rb_chain2 L2 eviction hints of the bulk stores: 1 hidden planes evict_first, 2 P / Pq evict_last
  int npiece, n3piece;    // k_rb_chain2: GEMM3 runs in npiece passes of n3piece (<= 256) columns through the same TMEM region
  int mode;               // 0: bias + ReLU (forward)   1: relu-grad masks (backward)
  float wsinv;            // accumulators of GEMM1 / GEMM2 are multiplied by this before the epilogue's bias / mask:
                          // 1 / kF16WScale when the weights were packed scaled (INB_PREC_FP16X3), else 1 (exact)
  int store;              // write both hidden tensors to HBM
  const float *bias1, *bias2;
  // relu-grad masks as bit planes [M][nh/32] uint32 in channel order: bit (ch & 31) of word (ch >> 5) of a row
  // (bit set = pre-activation negative).  mode 1 reads mask1 / mask2 in E1 / E2; mode 0 with store writes bits1 / bits2.
  const uint32_t *mask1, *mask2;
  uint32_t *bits1, *bits2;
  float* P;
  int exp;                // experiment bits (INB_CHAIN_EXP; timing studies only, results are garbage): 1 no weight TMA, 2 no mask loads
  long long* trace;       // diagnostics: per-tile phase timestamps of CTA 0 (inb_debug_chain_trace), nullable
};

constexpr int kChainThreads = 352;
constexpr uint32_t kPlane = 16384;      // one plane of a 128 x 64 chunk / of a weight k-block

__device__ __forceinline__ void chain_tap_offset(int tap, int ksz, int D, int& dx, int& dy, int& dz) {
  if (ksz == 1) { dx = dy = dz = 0; return; }
  dx = tap % 3 - 1;
  dy = (tap / 3) % 3 - 1;
  dz = (D > 1) ? tap / 9 - 1 : 0;
}

// Bit position of value j (0..7) of 8-column group g inside a 16- or 32-bit mask word: the signs are shifted in with one
// funnel shift per value, so within each group of 8 the FIRST value lands in the HIGHEST bit.  Groups 0 / 1 are the
// high / low byte of the low half-word, groups 2 / 3 of the high half-word (16-wide and 32-wide kernels agree).
__device__ __forceinline__ constexpr int chain_bit_base(int g) { return ((g & 1) ? 0 : 8) + (g >> 1) * 16; }

// 8 accumulator columns -> packed bf16 hi / lo words.  MODE 0: + bias, ReLU; when BITS the signs of the 8
// pre-activations are merged into `bits` (the _relugrad mask of activation_functions.jl:84, kept as a bit plane).
// MODE 1: columns whose bit in `mbits` is set are zeroed.  g = index of the 8-column group inside the mask word.
template <int MODE, int NT, bool BITS, bool F16 = false>
__device__ __forceinline__ void chain_pack8(const uint32_t* r, const float* sb, uint32_t mbits, int g, uint32_t& bits,
                                            uint4& oh, uint4& ol, float wsinv = 1.f) {
  float bias[8];
  if (MODE == 0) {
    const float4 b0 = *reinterpret_cast<const float4*>(sb), b1 = *reinterpret_cast<const float4*>(sb + 4);
    bias[0] = b0.x; bias[1] = b0.y; bias[2] = b0.z; bias[3] = b0.w;
    bias[4] = b1.x; bias[5] = b1.y; bias[6] = b1.z; bias[7] = b1.w;
  }
  const int base = chain_bit_base(g);
  uint32_t ph[4], pl[4];
  uint32_t b8 = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float a = __uint_as_float(r[2 * j]), b = __uint_as_float(r[2 * j + 1]);
    if (MODE == 0) {
      a = fmaf(a, wsinv, bias[2 * j]);      // wsinv == 1: the same single rounding as a + bias
      b = fmaf(b, wsinv, bias[2 * j + 1]);
      if (BITS) {
        b8 = __funnelshift_l(__float_as_uint(a), b8, 1);  // (b8 << 1) | sign(a)
        b8 = __funnelshift_l(__float_as_uint(b), b8, 1);
      }
      a = fmaxf(a, 0.f);
      b = fmaxf(b, 0.f);
    } else {
      a *= wsinv;
      b *= wsinv;
      if ((mbits >> (base + 7 - 2 * j)) & 1u) a = 0.f;
      if ((mbits >> (base + 6 - 2 * j)) & 1u) b = 0.f;
    }
    if (NT == 3) {
      split2<F16>(a, b, ph[j], pl[j]);
    } else {
      __nv_bfloat162 h2 = __floats2bfloat162_rn(a, b);
      ph[j] = *reinterpret_cast<uint32_t*>(&h2);
      pl[j] = 0;
    }
  }
  if (MODE == 0 && BITS) bits |= b8 << base;
  oh = make_uint4(ph[0], ph[1], ph[2], ph[3]);
  ol = make_uint4(pl[0], pl[1], pl[2], pl[3]);
}

template <int NT>
__global__ void __launch_bounds__(kChainThreads, 1)
k_rb_chain(const __grid_constant__ ChainMaps maps, const ChainArgs a) {
  constexpr int NP = (NT == 1) ? 1 : 2;
  constexpr uint32_t CHUNK = NP * kPlane;        // hi (+ lo) of one 128 x 64 chunk
  constexpr uint32_t STAGE = NP * kPlane;        // one ring stage
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
  float* sbias = reinterpret_cast<float*>(tslot + 4);  // [2][256], 16-byte aligned
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    prefetch_tmap(&maps.A[0]);
    prefetch_tmap(&maps.W1[0]);
    prefetch_tmap(&maps.W2[0]);
    prefetch_tmap(&maps.W3[0]);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < a.stages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
      for (int s = 0; s < 3; ++s) mbar_init(dfull + s, 1);
      for (int s = 0; s < 4; ++s) { mbar_init(hready + s, 8); mbar_init(stdone + s, 1); }
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
    sbias[i] = (bp && j < a.nh_bias) ? bp[j] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tslot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    // every box is [<=128 rows] x [64 bf16 = 128 bytes]: the TMA unit is request-rate bound (a 32-byte row costs
    // what a 128-byte one does), which is why the first GEMM reads im2col rows instead of nine shifted 32-byte taps
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
        for (int kb = 0; kb < a.nkb1; ++kb) {
          {  // 128 pixels x 64 im2col columns (rows beyond M are zero-filled)
            uint8_t* st = acquire(NP * kPlane);
            uint64_t* fb = full + it % a.stages;
#pragma unroll
            for (int pl = 0; pl < NP; ++pl) tma_load_2d(&maps.A[pl], fb, st + pl * kPlane, kb * 64, tile * 128);
            ++it;
          }
          for (int nhf = 0; nhf < a.nh / 128; ++nhf, ++it) {
            uint8_t* st = acquire(NP * kPlane);
            uint64_t* fb = full + it % a.stages;
#pragma unroll
            for (int pl = 0; pl < NP; ++pl) tma_load_2d(&maps.W1[pl], fb, st + pl * kPlane, kb * 64, nhf * 128);
          }
        }
        for (int c = 0; c < a.nchunk; ++c) {
          for (int nhf = 0; nhf < a.nh / 128; ++nhf, ++it) {
            uint8_t* st = acquire(NP * kPlane);
            uint64_t* fb = full + it % a.stages;
#pragma unroll
            for (int pl = 0; pl < NP; ++pl) tma_load_2d(&maps.W2[pl], fb, st + pl * kPlane, c * 64, nhf * 128);
          }
        }
        for (int pc = 0; pc < a.np3; ++pc) {
          for (int c = 0; c < a.nchunk; ++c, ++it) {
            uint8_t* st = acquire(NP * kPlane);
            uint64_t* fb = full + it % a.stages;
#pragma unroll
            for (int pl = 0; pl < NP; ++pl) tma_load_2d(&maps.W3[pl], fb, st + pl * kPlane, c * 64, pc * 128);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      const uint32_t idesc_2 = make_idesc_bf16(128, 128, 0, 0);    // GEMM1 / GEMM2: 128-column halves
      const uint32_t hb = smem_u32(hbuf);
      uint32_t it = 0, tl = 0;
      long long twait = 0;
      auto stage_wait = [&]() -> uint32_t {
        const int s = it % a.stages;
        const uint32_t ph = (it / a.stages) & 1;
        const long long t0 = a.trace ? clock64() : 0;
        mbar_wait(full + s, ph);
        if (a.trace) twait += clock64() - t0;
        tc_fence_after();
        return smem_u32(ring + (size_t)s * STAGE);
      };
      // K-major SWIZZLE_128B descriptors differ only in the start-address field: the constant high word is built
      // once and every MMA costs two integer adds (descriptor arithmetic in the single issuing thread was
      // what paced the tensor pipe before)
      const uint32_t dhi = (uint32_t)(make_smem_desc(0, 0, 1024, LAYOUT_SW128) >> 32);
      // one [128 rows x 64 k] A block at a_addr against one [<=128 rows x 64 k] B block at b_addr: 4 k-steps x NT terms
      auto mma_block = [&](uint32_t d_tmem, uint32_t idesc, uint32_t a_addr, uint32_t b_addr, uint32_t first) {
        const uint32_t alo = (a_addr >> 4) & 0x3FFF, blo = (b_addr >> 4) & 0x3FFF;
#pragma unroll
        for (int term = 0; term < NT; ++term) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t ad = ((uint64_t)dhi << 32) | (alo + ((term == 2) ? (kPlane >> 4) : 0) + 2 * k);
            const uint64_t bd = ((uint64_t)dhi << 32) | (blo + ((term == 1) ? (kPlane >> 4) : 0) + 2 * k);
            umma_f16(d_tmem, ad, bd, idesc, (term == 0 && k == 0) ? (first ? 0u : 1u) : 1u);
          }
        }
      };
      for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++tl) {
        const uint32_t R0 = tmem + (tl & 1) * 256, R1 = tmem + ((tl & 1) ^ 1) * 256;
        long long* tr = (a.trace && blockIdx.x == 0 && tl < 16) ? a.trace + tl * 16 : nullptr;
        if (tr) tr[0] = clock64();
        if (a.n3b && tl > 0) {  // D3b of the previous tile lives where D1 goes
          mbar_wait(e3done, (tl - 1) & 1);
          tc_fence_after();
        }
        for (int kb = 0; kb < a.nkb1; ++kb) {
          const uint32_t sA = stage_wait();
          const int slotA = it % a.stages;
          ++it;
          for (int nhf = 0; nhf < a.nh / 128; ++nhf, ++it) {
            const uint32_t sb = stage_wait();
            mma_block(R0 + nhf * 128, idesc_2, sA, sb, kb == 0);
            umma_commit(empty + it % a.stages);
          }
          umma_commit(empty + slotA);  // the im2col block is released after both column halves have read it
        }
        umma_commit(dfull + 0);
        if (tr) { tr[1] = clock64(); tr[12] = twait; twait = 0; }
        for (int c = 0; c < a.nchunk; ++c) {
          mbar_wait(hready + c, 0);
          tc_fence_after();
          if (tr && c == 0) tr[2] = clock64();
          for (int nhf = 0; nhf < a.nh / 128; ++nhf, ++it) {
            const uint32_t sa = stage_wait();
            mma_block(R1 + nhf * 128, idesc_2, hb + c * CHUNK, sa, c == 0);
            umma_commit(empty + it % a.stages);
          }
        }
        umma_commit(dfull + 1);
        if (tr) { tr[3] = clock64(); tr[13] = twait; twait = 0; }
        for (int pc = 0; pc < a.np3; ++pc) {
          const int n = min(128, a.n3pad - pc * 128);
          const uint32_t idesc_3 = make_idesc_bf16(128, n, 0, 0);
          const uint32_t dst = (pc < 2) ? (R0 + pc * 128) : (R1 + (pc - 2) * 128);
          for (int c = 0; c < a.nchunk; ++c, ++it) {
            mbar_wait(hready + c, 1);  // pieces in R1 come after all of E2 (piece-outer order)
            tc_fence_after();
            if (tr && pc == 0 && c == 0) tr[4] = clock64();
            const uint32_t sa = stage_wait();
            mma_block(dst, idesc_3, hb + c * CHUNK, sa, c == 0);
            umma_commit(empty + it % a.stages);
          }
        }
        umma_commit(dfull + 2);
        if (tr) { tr[5] = clock64(); tr[14] = twait; }
        twait = 0;
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
      long long* tr = (a.trace && blockIdx.x == 0 && tl < 16 && e == 0 && lane == 0) ? a.trace + tl * 16 + 6 : nullptr;
#pragma unroll 1
      for (int stg = 0; stg < 2; ++stg) {
        const uint32_t dsrc = (stg ? R1 : R0) + lane_sel + 32 * half;
        const float* sb = sbias + stg * 256 + 32 * half;
        const long long mrow = m * (a.nh >> 5) + half;  // word 2c + half = channels 64c + 32half .. +31
        const uint32_t* mk = (stg ? a.mask2 : a.mask1) + mrow;
        uint32_t* bo = (stg ? a.bits2 : a.bits1) + mrow;
        const bool wait_store = a.store && (tl > 0 || stg > 0);
        mbar_wait(dfull + stg, tl & 1);
        tc_fence_after();
        if (stg == 0 && tl > 0) {  // the P stores of the previous tile have read the operand buffer
          if (e == 0 && lane == 0) bulk_wait_read0();
          asm volatile("bar.sync 1, 256;" ::: "memory");
        }
        if (tr) tr[2 * stg] = clock64();
        // software pipeline over the 64-channel chunks: the TMEM (and mask) loads of chunk c+1 are in flight
        // while chunk c is converted and written to shared memory
        uint32_t rA[32], rB[32];
        uint32_t mA = 0, mB = 0;
        auto fetch = [&](int c, uint32_t (&r)[32], uint32_t& mm) {
          if (a.mode == 1) mm = live ? __ldg(mk + 2 * c) : 0u;
          tmem_ld32(dsrc + 64 * c, r);
        };
        auto emit = [&](int c, const uint32_t (&r)[32], const uint32_t mm) {
          if (wait_store) mbar_wait(stdone + c, stg ^ 1);  // the chunk's previous TMA store has read it
          uint8_t* dst = hbuf + (size_t)c * CHUNK + row * 128;
          uint32_t bits = 0;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 oh, ol;
            if (a.mode == 0) {
              if (a.store) chain_pack8<0, NT, true>(r + 8 * g, sb + 64 * c + 8 * g, 0u, g, bits, oh, ol);
              else chain_pack8<0, NT, false>(r + 8 * g, sb + 64 * c + 8 * g, 0u, g, bits, oh, ol);
            } else {
              chain_pack8<1, NT, false>(r + 8 * g, sb, mm, g, bits, oh, ol);
            }
            const uint32_t off = (uint32_t)(((half * 4 + g) ^ (row & 7)) << 4);
            *reinterpret_cast<uint4*>(dst + off) = oh;
            if (NT == 3) *reinterpret_cast<uint4*>(dst + kPlane + off) = ol;
          }
          if (a.mode == 0 && a.store && live) bo[2 * c] = bits;
          fence_proxy_async();  // generic-proxy writes -> visible to the MMA / TMA (async proxy)
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(hready + c);
        };
        fetch(0, rA, mA);
#pragma unroll 1
        for (int c = 0; c < a.nchunk; c += 2) {
          tmem_ld_wait();
          fetch(c + 1, rB, mB);
          emit(c, rA, mA);
          tmem_ld_wait();
          if (c + 2 < a.nchunk) fetch(c + 2, rA, mA);
          emit(c + 1, rB, mB);
        }
        if (tr) tr[2 * stg + 1] = clock64();
      }
      // E3: tap-expanded columns -> P
      mbar_wait(dfull + 2, tl & 1);
      tc_fence_after();
      if (tr) tr[4] = clock64();
      // The operand buffer is free now (GEMM3 has read it): stage the fp32 rows there, 128 columns at a time,
      // so that P is written with coalesced 16-byte stores (a row per thread straight from registers costs one
      // memory transaction per lane and was the longest phase of the tile).
      if (a.store) {
#pragma unroll 1
        for (int c = 0; c < a.nchunk; ++c) mbar_wait(stdone + c, 1);  // the TMA stores of this tile's chunks have read them
      }
      const int tid = e * 32 + lane;
#pragma unroll 1
      for (int slab = 0; slab * 128 < a.n3pad; ++slab) {
        const int ncols = min(128, a.n3pad - slab * 128);
        const uint32_t dsrc = ((slab < 2) ? (R0 + slab * 128) : (R1 + (slab - 2) * 128)) + lane_sel;
        const int nsplit = ((ncols / 16 + 1) / 2) * 16;
        const int cbeg = half ? nsplit : 0, cend = half ? ncols : nsplit;
        if (slab > 0) {  // the previous slab's stores have read the staging area
          if (tid == 0) bulk_wait_read0();
          asm volatile("bar.sync 1, 256;" ::: "memory");
        }
        // staging = the TMA store's SWIZZLE_128B box layout: 32-column group g at g*16 KB, 128-byte rows
        auto stage16 = [&](int col, const uint32_t* r) {
          uint8_t* g = hbuf + (col >> 5) * kPlane + row * 128;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int ch = ((col & 31) >> 2) + j;
            *reinterpret_cast<uint4*>(g + ((ch ^ (row & 7)) << 4)) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
          }
        };
        int c0 = cbeg;
        for (; c0 + 32 <= cend; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(dsrc + c0, r);
          tmem_ld_wait();
          stage16(c0, r);
          stage16(c0 + 16, r + 16);
        }
        if (c0 < cend) {
          uint32_t r[16];
          tmem_ld16(dsrc + c0, r);
          tmem_ld_wait();
          stage16(c0, r);
        }
        fence_proxy_async();
        if (tr && slab == 0) tr[16 * 16 + 0 - 6] = clock64();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (tid == 0) {
          for (int g = 0; g * 32 < ncols; ++g) tma_store_2d(&maps.P, hbuf + g * kPlane, slab * 128 + g * 32, tile * 128);
          bulk_commit();
        }
        if (tr && slab == 0) tr[16 * 16 + 1 - 6] = clock64();
      }
      if (tr) tr[5] = clock64();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(e3done);
    }
    if (e == 0 && lane == 0) bulk_wait0();
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

// ---------------------------------------------------------------- chain with the hidden operand in tensor memory
// Same pipeline as k_rb_chain, but E1 / E2 write the bf16 hi/lo hidden tile back into TENSOR MEMORY (tcgen05.st, in
// place of the fp32 accumulator they just read) and GEMM2 / GEMM3 take their A operand from there
// (tcgen05.mma [d], [a_tmem], b_desc).  k_rb_chain is shared-memory-bandwidth bound: per tile the MMAs read ~1.3 MB of
// operands next to ~0.5 MB of TMA writes and ~0.4 MB of epilogue traffic against 128 B/clk.  With A in TMEM the
// MMAs read only the streamed weights, the 128 KB operand buffer disappears and the ring gets deeper.
// Column map of a 64-channel chunk c of the 256-column region R:  warp half h (channels 32h..32h+31) reads fp32
// columns R+64c+32h..+31 and overwrites them with 16 packed hi words (R+64c+32h..+15) and 16 packed lo words
// (+16..+31); k-step kk = 2h + j of the chunk starts at column R+64c+32h+8j (hi) / +16 (lo).
// Shared memory: [64 KB staging: two 32 KB slots for the TMA stores of the hidden chunks | the P slab of E3]
//                [ring of NP x 16 KB weight / im2col stages][barriers, biases].
template <int NT>
__global__ void __launch_bounds__(kChainThreads, 1)
k_rb_chain_t(const __grid_constant__ ChainMaps maps, const ChainArgs a) {
  constexpr int NP = (NT == 1) ? 1 : 2;
  constexpr uint32_t STAGE = NP * kPlane;
  constexpr uint32_t SLOT = 2 * kPlane;  // one staging slot: hi + lo planes of a 128 x 64 chunk
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* stag = smem;
  uint8_t* ring = stag + 2 * SLOT;
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + (size_t)a.stages * STAGE);
  uint64_t* empty = full + 8;
  uint64_t* dfull = empty + 8;    // [3]
  uint64_t* hready = dfull + 3;   // [4]  chunk c of the hidden operand is in tensor memory
  uint64_t* sready = hready + 4;  // [2]  staging slot written (store mode)
  uint64_t* sfree = sready + 2;   // [2]  staging slot read by its TMA store
  uint32_t* tslot = reinterpret_cast<uint32_t*>(sfree + 3);  // 28 barrier slots (one spare) keep sbias 16-byte aligned
  float* sbias = reinterpret_cast<float*>(tslot + 4);        // [2][256]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    prefetch_tmap(&maps.A[0]);
    prefetch_tmap(&maps.W1[0]);
    prefetch_tmap(&maps.W2[0]);
    prefetch_tmap(&maps.W3[0]);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < a.stages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
      for (int s = 0; s < 3; ++s) mbar_init(dfull + s, 1);
      for (int s = 0; s < 4; ++s) mbar_init(hready + s, 8);
      for (int s = 0; s < 2; ++s) { mbar_init(sready + s, 8); mbar_init(sfree + s, 1); }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tslot, 512);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < 512; i += blockDim.x) {
    const float* bp = (i < 256) ? a.bias1 : a.bias2;
    const int j = i & 255;
    sbias[i] = (bp && j < a.nh_bias) ? bp[j] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tslot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (identical to k_rb_chain)
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
        for (int kb = 0; kb < a.nkb1; ++kb) {
          {
            uint8_t* st = acquire(NP * kPlane);
            uint64_t* fb = full + it % a.stages;
#pragma unroll
            for (int pl = 0; pl < NP; ++pl) tma_load_2d(&maps.A[pl], fb, st + pl * kPlane, kb * 64, tile * 128);
            ++it;
          }
          for (int nhf = 0; nhf < a.nh / 128; ++nhf, ++it) {
            uint8_t* st = acquire(NP * kPlane);
            uint64_t* fb = full + it % a.stages;
#pragma unroll
            for (int pl = 0; pl < NP; ++pl) tma_load_2d(&maps.W1[pl], fb, st + pl * kPlane, kb * 64, nhf * 128);
          }
        }
        for (int c = 0; c < a.nchunk; ++c) {
          for (int nhf = 0; nhf < a.nh / 128; ++nhf, ++it) {
            uint8_t* st = acquire((a.exp & 1) ? 0 : NP * kPlane);
            uint64_t* fb = full + it % a.stages;
            if (a.exp & 1) continue;
#pragma unroll
            for (int pl = 0; pl < NP; ++pl) tma_load_2d(&maps.W2[pl], fb, st + pl * kPlane, c * 64, nhf * 128);
          }
        }
        for (int pc = 0; pc < a.np3; ++pc) {
          for (int c = 0; c < a.nchunk; ++c, ++it) {
            uint8_t* st = acquire((a.exp & 1) ? 0 : NP * kPlane);
            uint64_t* fb = full + it % a.stages;
            if (a.exp & 1) continue;
#pragma unroll
            for (int pl = 0; pl < NP; ++pl) tma_load_2d(&maps.W3[pl], fb, st + pl * kPlane, c * 64, pc * 128);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      const uint32_t idesc_2 = make_idesc_bf16(128, 128, 0, 0);
      const uint32_t dhi = (uint32_t)(make_smem_desc(0, 0, 1024, LAYOUT_SW128) >> 32);
      uint32_t it = 0, tl = 0;
      long long twait = 0;
      auto stage_wait = [&]() -> uint32_t {
        const int s = it % a.stages;
        const uint32_t ph = (it / a.stages) & 1;
        const long long t0 = a.trace ? clock64() : 0;
        mbar_wait(full + s, ph);
        if (a.trace) twait += clock64() - t0;
        tc_fence_after();
        return smem_u32(ring + (size_t)s * STAGE);
      };
      // GEMM1: A (im2col block) and B both in shared memory
      auto mma_block_ss = [&](uint32_t d_tmem, uint32_t idesc, uint32_t a_addr, uint32_t b_addr, uint32_t first) {
        const uint32_t alo = (a_addr >> 4) & 0x3FFF, blo = (b_addr >> 4) & 0x3FFF;
#pragma unroll
        for (int term = 0; term < NT; ++term) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t ad = ((uint64_t)dhi << 32) | (alo + ((term == 2) ? (kPlane >> 4) : 0) + 2 * k);
            const uint64_t bd = ((uint64_t)dhi << 32) | (blo + ((term == 1) ? (kPlane >> 4) : 0) + 2 * k);
            umma_f16(d_tmem, ad, bd, idesc, (term == 0 && k == 0) ? (first ? 0u : 1u) : 1u);
          }
        }
      };
      // GEMM2 / GEMM3: A = chunk c of the hidden operand in tensor memory (region ra), B = weight block in smem
      auto mma_block_ts = [&](uint32_t d_tmem, uint32_t idesc, uint32_t ra, int c, uint32_t b_addr, uint32_t first) {
        const uint32_t blo = (b_addr >> 4) & 0x3FFF;
#pragma unroll
        for (int term = 0; term < NT; ++term) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t at = ra + 64 * c + 32 * (k >> 1) + 8 * (k & 1) + ((term == 2) ? 16 : 0);
            const uint64_t bd = ((uint64_t)dhi << 32) | (blo + ((term == 1) ? (kPlane >> 4) : 0) + 2 * k);
            umma_f16_ts(d_tmem, at, bd, idesc, (term == 0 && k == 0) ? (first ? 0u : 1u) : 1u);
          }
        }
      };
      for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++tl) {
        const uint32_t R0 = tmem + (tl & 1) * 256, R1 = tmem + ((tl & 1) ^ 1) * 256;
        long long* tr = (a.trace && blockIdx.x == 0 && tl < 16) ? a.trace + tl * 16 : nullptr;
        if (tr) tr[0] = clock64();
        if (tl > 0) {  // R0 held the A operand of the previous tile's GEMM3: let those MMAs retire first
          mbar_wait(dfull + 2, (tl - 1) & 1);
          tc_fence_after();
        }
        for (int kb = 0; kb < a.nkb1; ++kb) {
          const uint32_t sA = stage_wait();
          const int slotA = it % a.stages;
          ++it;
          for (int nhf = 0; nhf < a.nh / 128; ++nhf, ++it) {
            const uint32_t sb = stage_wait();
            mma_block_ss(R0 + nhf * 128, idesc_2, sA, sb, kb == 0);
            umma_commit(empty + it % a.stages);
          }
          umma_commit(empty + slotA);
        }
        umma_commit(dfull + 0);
        if (tr) { tr[1] = clock64(); tr[12] = twait; }
        twait = 0;
        for (int c = 0; c < a.nchunk; ++c) {
          mbar_wait(hready + c, 0);
          tc_fence_after();
          if (tr && c == 0) tr[2] = clock64();
          for (int nhf = 0; nhf < a.nh / 128; ++nhf, ++it) {
            const uint32_t sb = stage_wait();
            mma_block_ts(R1 + nhf * 128, idesc_2, R0, c, sb, c == 0);
            umma_commit(empty + it % a.stages);
          }
        }
        umma_commit(dfull + 1);
        if (tr) { tr[3] = clock64(); tr[13] = twait; }
        twait = 0;
        for (int pc = 0; pc < a.np3; ++pc) {
          const int n = min(128, a.n3pad - pc * 128);
          const uint32_t idesc_3 = make_idesc_bf16(128, n, 0, 0);
          for (int c = 0; c < a.nchunk; ++c, ++it) {
            mbar_wait(hready + c, 1);
            tc_fence_after();
            if (tr && pc == 0 && c == 0) tr[4] = clock64();
            const uint32_t sb = stage_wait();
            mma_block_ts(R0 + pc * 128, idesc_3, R1, c, sb, c == 0);
            umma_commit(empty + it % a.stages);
          }
        }
        umma_commit(dfull + 2);
        if (tr) { tr[5] = clock64(); tr[14] = twait; }
        twait = 0;
      }
    }
  } else if (warp < 10) {
    // ------------------------------------------------------------ epilogue warps
    const int e = warp - 2;
    const int q = warp & 3;
    const int half = e >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    const int tid = e * 32 + lane;
    uint32_t tl = 0;
    uint32_t use0 = 0, use1 = 0;  // completed uses of the two staging slots (store mode)
    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++tl) {
      const uint32_t R0 = tmem + (tl & 1) * 256, R1 = tmem + ((tl & 1) ^ 1) * 256;
      const long long m = (long long)tile * 128 + row;
      const bool live = m < a.M;
      long long* tr = (a.trace && blockIdx.x == 0 && tl < 16 && e == 0 && lane == 0) ? a.trace + tl * 16 + 6 : nullptr;
#pragma unroll 1
      for (int stg = 0; stg < 2; ++stg) {
        const uint32_t reg = (stg ? R1 : R0) + lane_sel + 32 * half;
        const float* sb = sbias + stg * 256 + 32 * half;
        const long long mrow = m * (a.nh >> 5) + half;  // word 2c + half = channels 64c + 32half .. +31
        const uint32_t* mk = (stg ? a.mask2 : a.mask1) + mrow;
        uint32_t* bo = (stg ? a.bits2 : a.bits1) + mrow;
        mbar_wait(dfull + stg, tl & 1);
        tc_fence_after();
        if (stg == 0 && tl > 0) {  // the P stores of the previous tile have read the staging area
          if (tid == 0) bulk_wait_read0();
          asm volatile("bar.sync 1, 256;" ::: "memory");
        }
        if (tr) tr[2 * stg] = clock64();
        uint32_t rA[32], rB[32];
        uint32_t mA = 0, mB = 0;
        auto fetch = [&](int c, uint32_t (&r)[32], uint32_t& mm) {
          if (a.mode == 1) mm = (live && !(a.exp & 2)) ? __ldg(mk + 2 * c) : 0u;
          tmem_ld32(reg + 64 * c, r);
        };
        auto emit = [&](int c, const uint32_t (&r)[32], const uint32_t mm) {
          uint32_t wh[16], wl[16];
          uint32_t bits = 0;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 oh, ol;
            if (a.mode == 0) {
              if (a.store) chain_pack8<0, NT, true>(r + 8 * g, sb + 64 * c + 8 * g, 0u, g, bits, oh, ol);
              else chain_pack8<0, NT, false>(r + 8 * g, sb + 64 * c + 8 * g, 0u, g, bits, oh, ol);
            } else {
              chain_pack8<1, NT, false>(r + 8 * g, sb, mm, g, bits, oh, ol);
            }
            wh[4 * g] = oh.x; wh[4 * g + 1] = oh.y; wh[4 * g + 2] = oh.z; wh[4 * g + 3] = oh.w;
            wl[4 * g] = ol.x; wl[4 * g + 1] = ol.y; wl[4 * g + 2] = ol.z; wl[4 * g + 3] = ol.w;
          }
          // in place: the 32 fp32 columns just read become 16 packed hi + 16 packed lo columns
          tmem_st16(reg + 64 * c, wh);
          if (NT == 3) tmem_st16(reg + 64 * c + 16, wl);
          if (a.mode == 0 && a.store && live) bo[2 * c] = bits;
          if (a.store) {
            const int slot = c & 1;
            uint32_t& use = slot ? use1 : use0;
            if (use > 0) mbar_wait(sfree + slot, (use - 1) & 1);  // the slot's previous TMA store has read it
            uint8_t* dst = stag + slot * SLOT + row * 128;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const uint32_t off = (uint32_t)(((half * 4 + g) ^ (row & 7)) << 4);
              *reinterpret_cast<uint4*>(dst + off) = make_uint4(wh[4 * g], wh[4 * g + 1], wh[4 * g + 2], wh[4 * g + 3]);
              if (NT == 3)
                *reinterpret_cast<uint4*>(dst + kPlane + off) = make_uint4(wl[4 * g], wl[4 * g + 1], wl[4 * g + 2], wl[4 * g + 3]);
            }
            fence_proxy_async();
            ++use;
          }
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(hready + c);
            if (a.store) mbar_arrive(sready + (c & 1));
          }
        };
        fetch(0, rA, mA);
#pragma unroll 1
        for (int c = 0; c < a.nchunk; c += 2) {
          tmem_ld_wait();
          fetch(c + 1, rB, mB);
          emit(c, rA, mA);
          tmem_ld_wait();
          if (c + 2 < a.nchunk) fetch(c + 2, rA, mA);
          emit(c + 1, rB, mB);
        }
        if (tr) tr[2 * stg + 1] = clock64();
      }
      // E3: tap-expanded columns (region R0, <= 256 of them) -> staging -> TMA store of P
      mbar_wait(dfull + 2, tl & 1);
      tc_fence_after();
      if (tr) tr[4] = clock64();
      if (a.store) {  // both staging slots: their last TMA stores have read them
        if (use0 > 0) mbar_wait(sfree + 0, (use0 - 1) & 1);
        if (use1 > 0) mbar_wait(sfree + 1, (use1 - 1) & 1);
      }
#pragma unroll 1
      for (int slab = 0; slab * 128 < a.n3pad; ++slab) {
        const int ncols = min(128, a.n3pad - slab * 128);
        const uint32_t dsrc = R0 + slab * 128 + lane_sel;
        const int nsplit = ((ncols / 16 + 1) / 2) * 16;
        const int cbeg = half ? nsplit : 0, cend = half ? ncols : nsplit;
        if (slab > 0) {
          if (tid == 0) bulk_wait_read0();
          asm volatile("bar.sync 1, 256;" ::: "memory");
        }
        auto stage16 = [&](int col, const uint32_t* r) {
          uint8_t* g = stag + (col >> 5) * kPlane + row * 128;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int ch = ((col & 31) >> 2) + j;
            *reinterpret_cast<uint4*>(g + ((ch ^ (row & 7)) << 4)) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
          }
        };
        int c0 = cbeg;
        for (; c0 + 32 <= cend; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(dsrc + c0, r);
          tmem_ld_wait();
          stage16(c0, r);
          stage16(c0 + 16, r + 16);
        }
        if (c0 < cend) {
          uint32_t r[16];
          tmem_ld16(dsrc + c0, r);
          tmem_ld_wait();
          stage16(c0, r);
        }
        fence_proxy_async();
        tc_fence_before();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (tid == 0) {
          for (int g = 0; g * 32 < ncols; ++g) tma_store_2d(&maps.P, stag + g * kPlane, slab * 128 + g * 32, tile * 128);
          bulk_commit();
        }
      }
      if (tr) tr[5] = clock64();
    }
    if (tid == 0) bulk_wait0();
  } else if (a.store) {
    // ------------------------------------------------------------ TMA store warp: hidden chunks -> HBM
    if (lane == 0) {
      uint32_t use[2] = {0, 0};
      for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
        for (int stg = 0; stg < 2; ++stg) {
          for (int c = 0; c < a.nchunk; ++c) {
            const int slot = c & 1;
            mbar_wait(sready + slot, use[slot] & 1);
#pragma unroll
            for (int pl = 0; pl < NP; ++pl)
              tma_store_2d(stg ? &maps.O2[pl] : &maps.O1[pl], stag + slot * SLOT + pl * kPlane, 64 * c, tile * 128);
            bulk_commit();
            bulk_wait_read0();
            mbar_arrive(sfree + slot);
            ++use[slot];
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

// ---------------------------------------------------------------- chain on CTA pairs (cta_group::2)
// k_rb_chain_t is bound by shared-memory bandwidth, not by the tensor pipe: an M = 128 MMA reads 64 B/clk of its B
// operand while TMA writes the next weight stage (42 B/clk) against a 128 B/clk port, so every MMA runs ~1.35x
// (GEMM2) to 2x (GEMM1 / GEMM3) over its floor (chain_trace.py; skipping the weight loads altogether changes
// nothing, so it is not L2).  k_rb_chain2 runs the same pipeline on a CLUSTER OF TWO CTAs: each CTA owns one
// 128-pixel tile (its own TMEM, epilogue and stores), the leader CTA issues tcgen05.mma.cta_group::2 with
// M = 256 over both tiles, and each CTA holds only HALF of every weight stage (N/2 rows).  B reads and TMA
// writes per SM halve (32 + 21 B/clk).  The epilogue is spread over 16 warps (16 channels x 32 rows per chunk
// each, i.e. one k-step of the next GEMM per warp) so that it keeps up with the faster MMAs.
//
// Cross-CTA protocol: weight / im2col stages - both CTAs' TMA loads complete_tx on the LEADER's full barrier
// (cp.async.bulk.tensor ... cta_group::2), the leader's producer posts the expect_tx for both; stage release and
// accumulator-ready signals are tcgen05.commit.cta_group::2 multicast to both CTAs; "hidden operand ready" is
// counted on the leader's barrier by the epilogue warps of both CTAs (remote mbarrier.arrive).
// The hidden tensors of a storing pass are staged chunk by chunk in two shared-memory slots and leave with TMA stores.
// (Writing them straight from the registers - one 32-byte sector per thread and plane with 256-bit stores - was tried:
// a warp's 32 sectors are 32 separate L1 transactions, and the storing passes became 1.5x slower.)
// Warps: 0 TMA producer | 1 MMA issuer (leader only) | 2..17 epilogue | 18 TMA store of the hidden tensors.
constexpr int kEpi2Warps = 16;
constexpr int kChain2Threads = 32 * (2 + kEpi2Warps + 1);
constexpr int kChain2Slots = 2;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `saddr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  // default semantics (.release.cta) as for a local arrive: what the leader's MMA consumes lives in tensor memory and is
  // ordered by tcgen05.wait::st + tcgen05.fence::before_thread_sync; a .release.cluster here costs ~1000 cycles per arrive
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  const long long t0 = clock64();
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) return;
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
// TMA load into THIS CTA's shared memory whose completion is counted on a barrier given as a shared::cluster
// address (the leader CTA's barrier)
__device__ __forceinline__ void tma_load_2d_cg2(const void* tmap, uint32_t bar_cluster_addr, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_2d(const void* tmap, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0),
               "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma2_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this shared-memory offset in BOTH CTAs once all MMAs issued so far have completed
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }

template <int NT, bool TRACE, bool F16 = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kChain2Threads, 1)
k_rb_chain2(const __grid_constant__ ChainMaps maps, const ChainArgs a) {
  constexpr int NP = (NT == 1) ? 1 : 2;
  constexpr uint32_t STAGE = kPlane;       // one ring granule (16 KB): a plane of an im2col block or of a GEMM3 weight
                                           // block, or both planes of a GEMM1 / GEMM2 weight half
  // one staging slot: hi + lo planes of a 128 x 64 chunk (16 + 16 KB), or hi + the 1-byte lo plane (16 + 8 KB)
  const uint32_t SLOT = a.lo8 ? 3 * kPlane / 2 : 2 * kPlane;
  constexpr int EPI = kEpi2Warps * 32;
  // shared memory: [2 staging slots of the hidden-tensor stores (store mode only)][P staging][ring][barriers, biases]
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* stag = smem;
  uint8_t* pstag = stag + (a.store ? kChain2Slots * SLOT : 0);
  uint8_t* ring = pstag + a.pstag_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + (size_t)a.stages * STAGE);
  uint64_t* empty = full + 12;
  uint64_t* d1h = empty + 12;     // [2]  column half h of D1 is complete
  uint64_t* d2h = d1h + 2;        // [2]  column half h of D2 is complete
  uint64_t* d3f = d2h + 2;        // [1]
  uint64_t* hready = d3f + 1;     // [4]  (leader's copy is the live one)
  uint64_t* sready = hready + 4;  // [2]
  uint64_t* sfree = sready + 2;   // [2]
  uint64_t* e3free = sfree + 2;   // [1]  the epilogues have read a GEMM3 piece out of tensor memory (leader's copy)
  uint32_t* tslot = reinterpret_cast<uint32_t*>(full + 48);  // 48 barrier slots -> 384 bytes
  float* sbias = reinterpret_cast<float*>(tslot + 4);        // [2][256]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int npairs = (a.ntiles + 1) >> 1;
  const int pair0 = blockIdx.x >> 1, pstride = gridDim.x >> 1;
  const int nhh = a.nh >> 1;          // columns of one half of D1 / D2
  const int qrows = a.nh >> 2;        // weight rows of one half held by this CTA
  const int n3half = a.n3piece >> 1;  // weight rows of a GEMM3 piece held by this CTA
  const int chalf = a.nchunk >> 1;    // chunks per column half
  const bool w3_one = NP * n3half * 128 <= (int)kPlane;  // a GEMM3 weight block (hi + lo) fits one ring granule

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    prefetch_tmap(&maps.A[0]);
    prefetch_tmap(&maps.W1[0]);
    prefetch_tmap(&maps.W2[0]);
    prefetch_tmap(&maps.W3[0]);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < a.stages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
      for (int s = 0; s < 2; ++s) { mbar_init(d1h + s, 1); mbar_init(d2h + s, 1); }
      mbar_init(d3f, 1);
      mbar_init(e3free, 2 * kEpi2Warps);
      for (int s = 0; s < 4; ++s) mbar_init(hready + s, 2 * kEpi2Warps);
      for (int s = 0; s < kChain2Slots; ++s) { mbar_init(sready + s, kEpi2Warps); mbar_init(sfree + s, 1); }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc2(tslot, 512);
    tmem_relinquish2();
  }
  for (int i = threadIdx.x; i < 512; i += blockDim.x) {
    const float* bp = (i < 256) ? a.bias1 : a.bias2;
    const int j = i & 255;
    sbias[i] = (bp && j < a.nh_bias) ? bp[j] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers are initialised and its tensor memory is allocated
  tc_fence_after();
  const uint32_t tmem = *tslot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs)
    if (elect_one()) {
      uint32_t it = 0;
      // wait for the local stage to be free, (leader) post the bytes of BOTH CTAs, return the stage and the
      // leader's full barrier
      auto acquire = [&](uint32_t bytes_per_cta, uint32_t& bar) -> uint8_t* {
        const int s = it % a.stages;
        const uint32_t ph = (it / a.stages) & 1;
        mbar_wait(empty + s, ph ^ 1);
        if (leader) mbar_expect_tx(full + s, 2 * bytes_per_cta);
        bar = mapa_u32(smem_u32(full + s), 0);
        ++it;
        return ring + (size_t)s * STAGE;
      };
      for (int tp = pair0; tp < npairs; tp += pstride) {
        const int tile = 2 * tp + (int)rank;
        uint32_t bar;
        if (tp + pstride < npairs) {
          // the im2col rows are the only operand that comes from HBM: pull the next tile's blocks into L2 a whole
          // tile ahead (the ring's FIFO order cannot look that far)
          const int ntile = 2 * (tp + pstride) + (int)rank;
          for (int kb = 0; kb < a.nkb1; ++kb)
#pragma unroll
            for (int pl = 0; pl < NP; ++pl) tma_prefetch_2d(&maps.A[pl], kb * 64, ntile * 128);
        }
        for (int kb = 0; kb < a.nkb1; ++kb) {
#pragma unroll
          for (int pl = 0; pl < NP; ++pl) {  // 128 pixels x 64 im2col columns of this CTA's tile, one plane per granule
            uint8_t* st = acquire(kPlane, bar);
            tma_load_2d_cg2(&maps.A[pl], bar, st, kb * 64, tile * 128);
          }
          for (int h = 0; h < 2; ++h) {
            uint8_t* st = acquire(NP * qrows * 128, bar);
#pragma unroll
            for (int pl = 0; pl < NP; ++pl)
              tma_load_2d_cg2(&maps.W1[pl], bar, st + pl * qrows * 128, kb * 64, h * nhh + (int)rank * qrows);
          }
        }
        for (int h = 0; h < 2; ++h) {
          for (int c = 0; c < a.nchunk; ++c) {
            uint8_t* st = acquire(NP * qrows * 128, bar);
#pragma unroll
            for (int pl = 0; pl < NP; ++pl)
              tma_load_2d_cg2(&maps.W2[pl], bar, st + pl * qrows * 128, c * 64, h * nhh + (int)rank * qrows);
          }
        }
        for (int pc = 0; pc < a.npiece; ++pc) {
          const int w3row = pc * a.n3piece + (int)rank * n3half;
          for (int c = 0; c < a.nchunk; ++c) {
            if (w3_one) {  // both planes of the block fit one granule
              uint8_t* st = acquire(NP * n3half * 128, bar);
#pragma unroll
              for (int pl = 0; pl < NP; ++pl) tma_load_2d_cg2(&maps.W3[pl], bar, st + pl * n3half * 128, c * 64, w3row);
            } else {
#pragma unroll
              for (int pl = 0; pl < NP; ++pl) {
                uint8_t* st = acquire(n3half * 128, bar);
                tma_load_2d_cg2(&maps.W3[pl], bar, st, c * 64, w3row);
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA, both tiles of the pair)
    if (leader && elect_one()) {
      const uint32_t idesc_12 = make_idesc_16(256, nhh, 0, 0, F16);   // GEMM1 / GEMM2 run one column half at a time
      const uint32_t idesc_3 = make_idesc_16(256, a.n3piece, 0, 0, F16);
      const uint32_t dhi = (uint32_t)(make_smem_desc(0, 0, 1024, LAYOUT_SW128) >> 32);
      uint32_t it = 0, tl = 0;
      long long twait = 0;
      auto stage_wait = [&]() -> uint32_t {
        const int s = it % a.stages;
        const uint32_t ph = (it / a.stages) & 1;
        const long long t0 = (TRACE && a.trace) ? clock64() : 0;
        mbar_wait(full + s, ph);
        if (TRACE && a.trace) twait += clock64() - t0;
        tc_fence_after();
        return smem_u32(ring + (size_t)s * STAGE);
      };
      // operands as (hi plane, lo plane) shared-memory addresses
      auto mma_block_ss = [&](uint32_t d_tmem, uint32_t idesc, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo,
                              uint32_t first) {
        const uint32_t ah = (a_hi >> 4) & 0x3FFF, al = (a_lo >> 4) & 0x3FFF;
        const uint32_t bh = (b_hi >> 4) & 0x3FFF, bl = (b_lo >> 4) & 0x3FFF;
#pragma unroll
        for (int term = 0; term < NT; ++term) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t ad = ((uint64_t)dhi << 32) | (((term == 2) ? al : ah) + 2 * k);
            const uint64_t bd = ((uint64_t)dhi << 32) | (((term == 1) ? bl : bh) + 2 * k);
            umma2_f16(d_tmem, ad, bd, idesc, (term == 0 && k == 0) ? (first ? 0u : 1u) : 1u);
          }
        }
      };
      // A = chunk c of the hidden operand in tensor memory: k-step k at column ra + 64c + 16k (hi), + 8 (lo)
      auto mma_block_ts = [&](uint32_t d_tmem, uint32_t idesc, uint32_t ra, int c, uint32_t b_hi, uint32_t b_lo,
                              uint32_t first) {
        const uint32_t bh = (b_hi >> 4) & 0x3FFF, bl = (b_lo >> 4) & 0x3FFF;
#pragma unroll
        for (int term = 0; term < NT; ++term) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t at = ra + 64 * c + 16 * k + ((term == 2) ? 8 : 0);
            const uint64_t bd = ((uint64_t)dhi << 32) | (((term == 1) ? bl : bh) + 2 * k);
            umma2_f16_ts(d_tmem, at, bd, idesc, (term == 0 && k == 0) ? (first ? 0u : 1u) : 1u);
          }
        }
      };
      for (int tp = pair0; tp < npairs; tp += pstride, ++tl) {
        const uint32_t R0 = tmem + (tl & 1) * 256, R1 = tmem + ((tl & 1) ^ 1) * 256;
        long long* tr = (TRACE && a.trace && blockIdx.x == 0 && tl < 16) ? a.trace + tl * 16 : nullptr;
        if (tr) tr[0] = clock64();
        if (tl > 0) {  // R0 held the A operand of the previous pair's GEMM3: let those MMAs retire first
          mbar_wait(d3f, (tl * a.npiece - 1) & 1);
          tc_fence_after();
        }
        // GEMM1: both column halves per im2col block (the block stays resident while its two weight halves pass)
        const uint32_t wlo = (uint32_t)qrows * 128;  // lo plane of a GEMM1 / GEMM2 weight half inside its granule
        for (int kb = 0; kb < a.nkb1; ++kb) {
          const uint32_t aH = stage_wait();
          const int slotH = it % a.stages;
          ++it;
          uint32_t aL = aH;
          int slotL = slotH;
          if (NT == 3) {
            aL = stage_wait();
            slotL = it % a.stages;
            ++it;
          }
          for (int h = 0; h < 2; ++h, ++it) {
            const uint32_t sb = stage_wait();
            mma_block_ss(R0 + h * nhh, idesc_12, aH, aL, sb, sb + wlo, kb == 0);
            umma2_commit_mc(empty + it % a.stages);
            if (kb == a.nkb1 - 1) umma2_commit_mc(d1h + h);
          }
          umma2_commit_mc(empty + slotH);
          if (NT == 3) umma2_commit_mc(empty + slotL);
        }
        if (tr) { tr[1] = clock64(); tr[12] = twait; }
        twait = 0;
        // GEMM2: column half 0 over all chunks (paced by E1), then half 1 (its epilogue E2 of half 0 runs under it)
        for (int h = 0; h < 2; ++h) {
          for (int c = 0; c < a.nchunk; ++c, ++it) {
            if (h == 0) {
              mbar_wait(hready + c, 0);
              tc_fence_after();
              if (tr && c == 0) tr[2] = clock64();
            }
            const uint32_t sb = stage_wait();
            mma_block_ts(R1 + h * nhh, idesc_12, R0, c, sb, sb + wlo, c == 0);
            umma2_commit_mc(empty + it % a.stages);
          }
          umma2_commit_mc(d2h + h);
        }
        if (tr) { tr[3] = clock64(); tr[13] = twait; }
        twait = 0;
        for (int pc = 0; pc < a.npiece; ++pc) {
        if (pc > 0) {  // the epilogues of both CTAs have read the previous piece out of R0
          mbar_wait(e3free, (tl * (a.npiece - 1) + pc - 1) & 1);
          tc_fence_after();
        }
        for (int c = 0; c < a.nchunk; ++c) {
          if (pc == 0) {
            mbar_wait(hready + c, 1);
            tc_fence_after();
          }
          if (tr && c == 0) tr[4] = clock64();
          const uint32_t bH = stage_wait();
          const int slotH = it % a.stages;
          ++it;
          uint32_t bL = bH + (uint32_t)n3half * 128;
          int slotL = slotH;
          if (NT == 3 && !w3_one) {
            bL = stage_wait();
            slotL = it % a.stages;
            ++it;
          }
          mma_block_ts(R0, idesc_3, R1, c, bH, bL, c == 0);
          umma2_commit_mc(empty + slotH);
          if (NT == 3 && !w3_one) umma2_commit_mc(empty + slotL);
        }
        umma2_commit_mc(d3f);
        }
        if (tr) { tr[5] = clock64(); tr[14] = twait; }
        twait = 0;
      }
    }
  } else if (warp < 2 + kEpi2Warps) {
    // ------------------------------------------------------------ epilogue warps (both CTAs, own tile)
    const int e = warp - 2;
    const int q = warp & 3;   // TMEM lane quadrant this warp may access
    const int kk = e >> 2;    // which 16 channels of a 64-channel chunk (= k-step of the next GEMM)
    const int row = q * 32 + lane;
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    const int tid = e * 32 + lane;
    const uint32_t hready_leader = mapa_u32(smem_u32(hready), 0);
    const uint32_t e3free_leader = mapa_u32(smem_u32(e3free), 0);
    uint32_t tl = 0;
    uint32_t cs = 0;  // chunk stores issued so far (store mode): slot cs % 2, use cs / 2
    // qsum, second half of E3 (deferred): horizontal tap sums of tile `ptile` from the raw rows in shared memory into the
    // dense Q tile, then its TMA store.  It touches no tensor memory, so it runs while the epilogue warps would otherwise
    // wait for the tail of GEMM2's first half of the NEXT tile.
    float* const raw = reinterpret_cast<float*>(pstag);
    const int rp = a.n3pad + 4;  // row pitch in floats: 16-byte accesses of consecutive rows hit distinct banks
    float* const qst = raw + 128 * rp;  // dense [128][nq], the box of the (unswizzled) TMA store
    int qpending = -1;
    auto qsum_finish = [&](int ptile) {
      // items (pixel r, tap row tg, 4-channel group g), decoded with shifts and one multiply (W is a power of two,
      // tg = rest / ngrp through a 16-bit reciprocal): nothing is kept in registers across tiles
      const int nstep = rp + a.Cn;  // neighbour row, neighbour tap
      const int ngrp = a.cq >> 2, items = 128 * a.ntg * ngrp;
      const uint32_t inv = 65536u / (uint32_t)ngrp + 1u;
      const int vmode = ((a.Cn & 3) == 0) ? 4 : (((a.Cn & 1) == 0) ? 2 : 1);
#pragma unroll 1
      for (int i = tid; i < items; i += EPI) {
        const int r = i & 127, rest = i >> 7;
        const int tg = (int)(((uint32_t)rest * inv) >> 16), g = rest - tg * ngrp;
        const int x = r & (a.W - 1);
        const bool left = x > 0, right = x + 1 < a.W;
        const float* ctr = raw + r * rp + (tg * 3 + 1) * a.Cn + 4 * g;  // centre tap of this item's tap row
        float4 acc;
        if (vmode == 4) {
          acc = *reinterpret_cast<const float4*>(ctr);
          if (left) {
            const float4 v = *reinterpret_cast<const float4*>(ctr - nstep);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
          }
          if (right) {
            const float4 v = *reinterpret_cast<const float4*>(ctr + nstep);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
          }
        } else if (vmode == 2) {
          const bool hi2 = 4 * g + 2 < a.Cn;
          float2 a0 = *reinterpret_cast<const float2*>(ctr);
          float2 a1 = hi2 ? *reinterpret_cast<const float2*>(ctr + 2) : make_float2(0.f, 0.f);
          if (left) {
            const float2 v0 = *reinterpret_cast<const float2*>(ctr - nstep);
            a0.x += v0.x; a0.y += v0.y;
            if (hi2) { const float2 v1 = *reinterpret_cast<const float2*>(ctr - nstep + 2); a1.x += v1.x; a1.y += v1.y; }
          }
          if (right) {
            const float2 v0 = *reinterpret_cast<const float2*>(ctr + nstep);
            a0.x += v0.x; a0.y += v0.y;
            if (hi2) { const float2 v1 = *reinterpret_cast<const float2*>(ctr + nstep + 2); a1.x += v1.x; a1.y += v1.y; }
          }
          acc = make_float4(a0.x, a0.y, a1.x, a1.y);
        } else {
          float t4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int dxi = 0; dxi < 3; ++dxi) {
            if ((dxi == 0 && !left) || (dxi == 2 && !right)) continue;
            const float* src = ctr + (dxi - 1) * nstep;
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (4 * g + u < a.Cn) t4[u] += src[u];
          }
          acc = make_float4(t4[0], t4[1], t4[2], t4[3]);
        }
        *reinterpret_cast<float4*>(qst + (tg * 128 + r) * a.cq + 4 * g) = acc;  // [tap row][pixel][cq]
      }
      fence_proxy_async();
      asm volatile("bar.sync 1, %0;" ::"n"(EPI) : "memory");
      if (tid == 0) {
        // one dense [128][cq] box per tap row into that row's PLANE of Pq: the col2im gather then reads every plane as
        // contiguous pixels x cq floats (no sector shared with a segment another pixel row wants)
        if (a.hints & 2) {
          const uint64_t pol = l2_policy_evict_last();
          for (int tg = 0; tg < a.ntg; ++tg) tma_store_3d_hint(&maps.P, qst + tg * 128 * a.cq, 0, ptile * 128, tg, pol);
        } else {
          for (int tg = 0; tg < a.ntg; ++tg) tma_store_3d(&maps.P, qst + tg * 128 * a.cq, 0, ptile * 128, tg);
        }
        bulk_commit();
      }
    };
    for (int tp = pair0; tp < npairs; tp += pstride, ++tl) {
      const int tile = 2 * tp + (int)rank;
      const uint32_t R0 = tmem + (tl & 1) * 256, R1 = tmem + ((tl & 1) ^ 1) * 256;
      const long long m = (long long)tile * 128 + row;
      const bool live = m < a.M;
      long long* tr = (TRACE && a.trace && blockIdx.x == 0 && tl < 16 && e == 0 && lane == 0) ? a.trace + tl * 16 + 6 : nullptr;
#pragma unroll 1
      for (int stg = 0; stg < 2; ++stg) {
        const uint32_t reg = (stg ? R1 : R0) + lane_sel + 16 * kk;
        const float* sb = sbias + stg * 256 + 16 * kk;
        // this kernel's bit-plane layout: the 16-bit words of a row are ordered [kk][c], so the nchunk words a thread
        // produces (or needs) in a stage are contiguous: ONE 8-byte store (load) per thread and stage.  (Per-chunk 2-byte
        // stores were 32 L1 transactions per warp and chunk and cost ~290 cycles per chunk.)
        const long long hrow = m * (a.nh >> 4) + kk * a.nchunk;
        const uint16_t* mk = reinterpret_cast<const uint16_t*>(stg ? a.mask2 : a.mask1) + hrow;
        uint16_t* bo = reinterpret_cast<uint16_t*>(stg ? a.bits2 : a.bits1) + hrow;
        unsigned long long macc = 0, bacc = 0;  // masks of the stage's chunks (16 bits each): read / produced

        uint64_t* dh = stg ? d2h : d1h;
        if (a.mode == 1 && live && !(a.exp & 2)) {  // all of the stage's masks, before the wait
          if (a.nchunk == 4) macc = __ldg(reinterpret_cast<const unsigned long long*>(mk));
          else macc = __ldg(reinterpret_cast<const uint32_t*>(mk));
        }
        mbar_wait(dh + 0, tl & 1);
        tc_fence_after();
        if (tr) tr[2 * stg] = clock64();
        // one chunk at a time (no register double buffer: four epilogue warps per scheduler hide the TMEM load latency, and
        // the kernel has 96 registers per thread)
        uint32_t rA[16];
        auto emit = [&](int c, const uint32_t (&r)[16], const uint32_t mm) {
          uint32_t wh[8], wl[8];
          uint32_t bits = 0;
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            uint4 oh, ol;
            if (a.mode == 0) {
              if (a.store) chain_pack8<0, NT, true, F16>(r + 8 * g, sb + 64 * c + 8 * g, 0u, g, bits, oh, ol, a.wsinv);
              else chain_pack8<0, NT, false, F16>(r + 8 * g, sb + 64 * c + 8 * g, 0u, g, bits, oh, ol, a.wsinv);
            } else {
              chain_pack8<1, NT, false, F16>(r + 8 * g, sb, mm, g, bits, oh, ol, a.wsinv);
            }
            wh[4 * g] = oh.x; wh[4 * g + 1] = oh.y; wh[4 * g + 2] = oh.z; wh[4 * g + 3] = oh.w;
            wl[4 * g] = ol.x; wl[4 * g + 1] = ol.y; wl[4 * g + 2] = ol.z; wl[4 * g + 3] = ol.w;
          }
          // in place: the 16 fp32 columns just read become 8 packed hi + 8 packed lo columns
          tmem_st8(reg + 64 * c, wh);
          if (NT == 3) tmem_st8(reg + 64 * c + 8, wl);
          if (a.mode == 0 && a.store) bacc |= (unsigned long long)(bits & 0xFFFFu) << (16 * c);
          int slot = 0;
          if (a.store) {
            slot = (int)(cs % kChain2Slots);
            const uint32_t use = cs / kChain2Slots;
            if (use > 0) mbar_wait(sfree + slot, (use - 1) & 1);  // the slot's previous TMA store has read it
            uint8_t* dst = stag + slot * SLOT + row * 128;
            if (!(a.exp & 8))
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              const uint32_t off = (uint32_t)(((kk * 2 + g) ^ (row & 7)) << 4);
              *reinterpret_cast<uint4*>(dst + off) = make_uint4(wh[4 * g], wh[4 * g + 1], wh[4 * g + 2], wh[4 * g + 3]);
              if (NT == 3 && !a.lo8)
                *reinterpret_cast<uint4*>(dst + kPlane + off) = make_uint4(wl[4 * g], wl[4 * g + 1], wl[4 * g + 2], wl[4 * g + 3]);
            }
            if (NT == 3 && a.lo8 && !(a.exp & 8)) {
              // 16 residuals -> 16 bytes: the upper byte of each half, rounded to nearest (a carry out of the lower
              // half-word only touches the last mantissa bit of its neighbour, which the truncation drops).  Rows of
              // 64 bytes in the SWIZZLE_64B pattern of the byte plane's store: 16-byte chunk ^ ((row >> 1) & 3)
              constexpr uint32_t R = 0x00800080u;
              const uint4 o = make_uint4(__byte_perm(wl[0] + R, wl[1] + R, 0x7531), __byte_perm(wl[2] + R, wl[3] + R, 0x7531),
                                         __byte_perm(wl[4] + R, wl[5] + R, 0x7531), __byte_perm(wl[6] + R, wl[7] + R, 0x7531));
              *reinterpret_cast<uint4*>(stag + slot * SLOT + kPlane + row * 64 + ((kk ^ ((row >> 1) & 3)) << 4)) = o;
            }
            if (!(a.exp & 16)) fence_proxy_async();
            ++cs;
          }
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive_cluster(hready_leader + 8 * c);
            if (a.store) mbar_arrive(sready + slot);
          }
        };
#pragma unroll 1
        for (int c = 0; c < a.nchunk; ++c) {
          if (c == chalf) {  // first chunk of the second column half
            mbar_wait(dh + 1, tl & 1);
            tc_fence_after();
          }
          tmem_ld16(reg + 64 * c, rA);
          tmem_ld_wait();
          emit(c, rA, (uint32_t)(macc >> (16 * c)) & 0xFFFFu);
        }
        if (a.mode == 0 && a.store && live) {
          if (a.nchunk == 4) *reinterpret_cast<unsigned long long*>(bo) = bacc;
          else *reinterpret_cast<uint32_t*>(bo) = (uint32_t)bacc;
        }
        if (tr) tr[2 * stg + 1] = clock64();
        if (stg == 0 && qpending >= 0) {
          qsum_finish(qpending);
          qpending = -1;
        }
      }
      // E3: tap-expanded columns (region R0, <= 256 of them per piece) -> P staging -> TMA store of P
      for (int pc = 0; pc < a.npiece; ++pc) {
      mbar_wait(d3f, (tl * a.npiece + pc) & 1);
      tc_fence_after();
      if (tr) tr[4] = clock64();
      if (a.qsum == 2) {
        // Wide GEMM3 (129..256 columns, 2-D): the same tap-row planes, one tap row at a time.  Round tg stages the aligned
        // 32-column chunks that cover columns [3 tg Cn, 3 tg Cn + 3 Cn) of D3 as fp32 rows, sums the three horizontal
        // taps into plane tg of the dense Q tile and frees the rows for the next round (the rows of a whole 224-column
        // tile would not fit beside the ring).  Not deferred: the rows are reused.
        if (tl > 0 && tid == 0) bulk_wait_read0();  // the previous tile's Pq store has read qst (the barrier below orders it)
        const int ngrp = a.cq >> 2;
        float* const qsw = raw + 128 * a.rpw;
        for (int tg = 0; tg < a.ntg; ++tg) {
          const int col0 = 3 * tg * a.Cn, cbase = col0 & ~31, o = col0 - cbase;
          if (32 * kk < o + 3 * a.Cn) {
            uint32_t r[32];
            tmem_ld32(R0 + lane_sel + cbase + 32 * kk, r);
            tmem_ld_wait();
            float* dst = raw + row * a.rpw + 32 * kk;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<uint4*>(dst + 4 * j) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
          }
          tc_fence_before();
          asm volatile("bar.sync 1, %0;" ::"n"(EPI) : "memory");
          const int nstep = a.rpw + a.Cn;
#pragma unroll 1
          for (int i = tid; i < 128 * ngrp; i += EPI) {
            const int r = i & 127, g = i >> 7;
            const int x = r & (a.W - 1);
            const float* ctr = raw + r * a.rpw + o + a.Cn + 4 * g;
            float4 acc = *reinterpret_cast<const float4*>(ctr);
            if (x > 0) {
              const float4 v = *reinterpret_cast<const float4*>(ctr - nstep);
              acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
            if (x + 1 < a.W) {
              const float4 v = *reinterpret_cast<const float4*>(ctr + nstep);
              acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
            *reinterpret_cast<float4*>(qsw + (tg * 128 + r) * a.cq + 4 * g) = acc;
          }
          if (tg + 1 < a.ntg) asm volatile("bar.sync 1, %0;" ::"n"(EPI) : "memory");  // the rows may be overwritten
        }
        fence_proxy_async();
        asm volatile("bar.sync 1, %0;" ::"n"(EPI) : "memory");
        if (tid == 0) {
          for (int tg = 0; tg < a.ntg; ++tg) tma_store_3d(&maps.P, qsw + tg * 128 * a.cq, 0, tile * 128, tg);
          bulk_commit();
        }
      } else if (a.qsum) {
        // D3 rows -> plain fp32 rows in shared memory; then every (pixel, tap row, 4 channels) item adds its three
        // horizontal taps: Q[(y', x)][tg][n] = sum_dx D3[(y', x + dx)][(3 tg + dx + 1) Cn + n] (x + dx inside the image
        // row; a tile is whole image rows because 128 % W == 0); col2im then only sums over tap rows
        if (tl > 0 && tid == 0) bulk_wait_read0();  // the previous tile's Pq store has read qst
        {
          const int c0 = 32 * kk;
          float* dst = raw + row * rp + c0;
          if (c0 + 32 <= a.n3pad) {
            uint32_t r[32];
            tmem_ld32(R0 + lane_sel + c0, r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<uint4*>(dst + 4 * j) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
          } else if (c0 < a.n3pad) {
            uint32_t r[16];
            tmem_ld16(R0 + lane_sel + c0, r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 4; ++j)
              *reinterpret_cast<uint4*>(dst + 4 * j) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
          }
        }
        tc_fence_before();
        asm volatile("bar.sync 1, %0;" ::"n"(EPI) : "memory");
        qpending = tile;  // the tap sums and the store run after E1 of the next tile (see qsum_finish)
      } else
#pragma unroll 1
      for (int slab = 0; slab * 128 < a.n3piece; ++slab) {
        const int ncols = min(128, a.n3piece - slab * 128);
        const uint32_t dsrc = R0 + slab * 128 + lane_sel;
        if (slab > 0 || tl > 0 || pc > 0) {  // the previous P stores have read the staging area
          if (tid == 0) bulk_wait_read0();
          asm volatile("bar.sync 1, %0;" ::"n"(EPI) : "memory");
        }
        // warp group kk stages the 32-column group kk of the slab (the TMA store's SWIZZLE_128B box layout)
        auto stage16 = [&](int col, const uint32_t* r) {
          uint8_t* g = pstag + (col >> 5) * kPlane + row * 128;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int ch = ((col & 31) >> 2) + j;
            *reinterpret_cast<uint4*>(g + ((ch ^ (row & 7)) << 4)) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
          }
        };
        const int c0 = 32 * kk;
        if (c0 + 32 <= ncols) {
          uint32_t r[32];
          tmem_ld32(dsrc + c0, r);
          tmem_ld_wait();
          stage16(c0, r);
          stage16(c0 + 16, r + 16);
        } else if (c0 < ncols) {
          uint32_t r[16];
          tmem_ld16(dsrc + c0, r);
          tmem_ld_wait();
          stage16(c0, r);
        }
        fence_proxy_async();
        tc_fence_before();
        asm volatile("bar.sync 1, %0;" ::"n"(EPI) : "memory");
        if (tid == 0) {
          for (int g = 0; g * 32 < ncols; ++g)
            tma_store_2d(&maps.P, pstag + g * kPlane, pc * a.n3piece + slab * 128 + g * 32, tile * 128);
          bulk_commit();
        }
      }
      if (pc + 1 < a.npiece) {  // R0 may take the next piece (all of this warp's TMEM loads have completed)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(e3free_leader);
      }
      }
      if (tr) tr[5] = clock64();
    }
    if (qpending >= 0) qsum_finish(qpending);
    if (tid == 0) bulk_wait0();
  } else if (a.store) {
    // ------------------------------------------------------------ TMA store warp: hidden chunks -> HBM
    if (lane == 0) {
      uint32_t cs = 0;
      const uint64_t pol = l2_policy_evict_first();
      for (int tp = pair0; tp < npairs; tp += pstride) {
        const int tile = 2 * tp + (int)rank;
        for (int stg = 0; stg < 2; ++stg) {
          for (int c = 0; c < a.nchunk; ++c, ++cs) {
            const int slot = (int)(cs % kChain2Slots);
            mbar_wait(sready + slot, (cs / kChain2Slots) & 1);
            if (!(a.exp & 4)) {
#pragma unroll
              for (int pl = 0; pl < NP; ++pl) {
                // (with lo8 the second map is the byte plane: box of 64 bytes x 128 rows, same coordinates)
                if (a.hints & 1)
                  tma_store_2d_hint(stg ? &maps.O2[pl] : &maps.O1[pl], stag + slot * SLOT + pl * kPlane, 64 * c, tile * 128, pol);
                else
                  tma_store_2d(stg ? &maps.O2[pl] : &maps.O1[pl], stag + slot * SLOT + pl * kPlane, 64 * c, tile * 128);
              }
            }
            bulk_commit();
            if (cs > 0) {  // one store stays in flight; the one before it has read its slot
              bulk_wait_read1();
              mbar_arrive(sfree + (int)((cs - 1) % kChain2Slots));
            }
          }
        }
      }
      if (cs > 0) {
        bulk_wait_read0();
        mbar_arrive(sfree + (int)((cs - 1) % kChain2Slots));
      }
      bulk_wait0();
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // no CTA leaves (or frees tensor memory) while its peer may still signal or read it
  if (warp == 1) tmem_dealloc2(tmem, 512);
}

// ---------------------------------------------------------------- col2im
// out[b][n][pix] = sum_tap P[pix + off(tap)][tap*Cn + n]  (+ passthrough add), coalesced both ways through a
// shared-memory transpose: phase 1 walks (pixel, n) with n fastest (contiguous in P), phase 2 walks pixels.
struct Col2imArgs {
  // element (pixel row m, tap t, channel n) of P sits at P[m * n3pad + t * cstride + n].  Full P: n3pad = the padded row,
  // cstride = Cn.  qsum (k_rb_chain2's tap-ROW sums): `taps` counts tap rows and every tap row is a dense PLANE
  // [M][cq] of its own: n3pad = cq, cstride = M * cq.
  int qsum;
  long long cstride;
  int W, H, D, ksz, taps, Cn, n3pad;
  long long px, M;
  const float* P;
  float* out0; long long out0_bs; int n0;
  float* out1; long long out1_bs; int out1_accum;
  const float* add; long long add_bs; int add_n;
  // INB_PREC_FP16X3: P carries the weight scale of the last contraction (oscale = 1 / kF16WScale) and, in the backward
  // pass, the gradient scale derived from *smax; both are powers of two, removed here before the passthrough add
  float oscale;
  const uint32_t* smax;
};
__device__ __forceinline__ float col2im_scale(const Col2imArgs& a) {
  float s = a.oscale;
  if (a.smax) { float sc, inv; f16_scale_from_max(__ldg(a.smax), sc, inv); s *= inv; }
  return s;
}
// thread = pixel: every tap (or tap row) contributes V-wide vectors of the pixel's P row, the sums stay in registers
// (16 channels at a time) and each channel is stored with the warp's 32 consecutive pixels (coalesced both ways, no
// shared memory; the 144-byte row stride of the loads is absorbed by L1, every byte of a row is used)
template <int V>
__global__ void __launch_bounds__(128) k_col2im(const Col2imArgs a) {
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= a.M) return;
  const long long b = m / a.px, pix = m - b * a.px;
  long long t = pix;
  const int x = (int)(t % a.W); t /= a.W;
  const int y = (int)(t % a.H); t /= a.H;
  const int z = (int)t;
  const float osc = col2im_scale(a);
  for (int c0 = 0; c0 < a.Cn; c0 += 16) {
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = 0.f;
    for (int tap = 0; tap < a.taps; ++tap) {
      int dx, dy, dz;
      if (a.qsum) { dx = 0; dy = tap % 3 - 1; dz = (a.D > 1) ? tap / 3 - 1 : 0; }
      else chain_tap_offset(tap, a.ksz, a.D, dx, dy, dz);
      const int xx = x + dx, yy = y + dy, zz = z + dz;
      if (xx < 0 || xx >= a.W || yy < 0 || yy >= a.H || zz < 0 || zz >= a.D) continue;
      const float* src = a.P + (m + dx + (long long)dy * a.W + (long long)dz * a.W * a.H) * a.n3pad + tap * a.cstride + c0;
#pragma unroll
      for (int j = 0; j < 16; j += V) {
        if (c0 + j < a.Cn) {
          if (V == 4) {
            const float4 w = __ldg(reinterpret_cast<const float4*>(src + j));
            acc[j] += w.x; acc[(j + 1) % 16] += w.y; acc[(j + 2) % 16] += w.z; acc[(j + 3) % 16] += w.w;
          } else if (V == 2) {
            const float2 w = __ldg(reinterpret_cast<const float2*>(src + j));
            acc[j] += w.x; acc[(j + 1) % 16] += w.y;
          } else {
            acc[j] += __ldg(src + j);
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int n = c0 + j;
      if (n < a.Cn) {
        float v = acc[j] * osc;
        if (a.add && n < a.add_n) v += a.add[b * a.add_bs + (long long)n * a.px + pix];
        if (n < a.n0) {
          a.out0[b * a.out0_bs + (long long)n * a.px + pix] = v;
        } else {
          float* qq = a.out1 + b * a.out1_bs + (long long)(n - a.n0) * a.px + pix;
          *qq = a.out1_accum ? (*qq + v) : v;
        }
      }
    }
  }
}

// thread = (pixel, group of 4 channels), G = ceil(Cn / 4) threads per pixel: the G float4 loads of a pixel's tap (row)
// are adjacent lanes, so a warp instruction touches ~11 pixels x G chunks = a third of the cache lines the
// thread-per-pixel walk does (that one is L1-transaction bound: 32 lines per instruction).  Needs 16-byte aligned
// groups: cstride and the pitch multiples of 4 (always true for the tap-row sums Pq).
template <int G, bool Q2 = false>
__global__ void __launch_bounds__(32 * G) k_col2im_g(const Col2imArgs a) {
  // grid (pixels of a sample / 32, sample): the pixel decode is 32-bit (the kernel is instruction-bound; the 64-bit
  // divisions of a flat pixel index were most of its instructions)
  const int pix = (int)(blockIdx.x * 32 + threadIdx.x / G);
  const int g = threadIdx.x % G;
  if (pix >= (int)a.px) return;
  const long long b = blockIdx.y;
  const long long m = b * a.px + pix;
  const int x = pix % a.W;
  const int yz = pix / a.W;
  const int y = yz % a.H;
  const int z = yz / a.H;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if constexpr (Q2) {
    // 2-D tap-row sums (three rows dy = -1, 0, 1): the three loads are issued together (the rolled loop below keeps one
    // load in flight per thread and the kernel is latency-bound)
    const float* row = a.P + m * a.n3pad + 4 * g;
    const long long rs = (long long)a.W * a.n3pad;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 w0 = (y > 0) ? __ldg(reinterpret_cast<const float4*>(row - rs)) : zero;
    const float4 w1 = __ldg(reinterpret_cast<const float4*>(row + a.cstride));
    const float4 w2 = (y + 1 < a.H) ? __ldg(reinterpret_cast<const float4*>(row + rs + 2 * a.cstride)) : zero;
    acc.x = (w0.x + w1.x) + w2.x; acc.y = (w0.y + w1.y) + w2.y;
    acc.z = (w0.z + w1.z) + w2.z; acc.w = (w0.w + w1.w) + w2.w;
  } else if (!a.qsum && a.D == 1 && a.taps == 9) {
    // full 3x3 P: nine independent loads, added in the rolled loop's order (a skipped tap adds +0)
    const float* base = a.P + m * a.n3pad + 4 * g;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 w[9];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int dx = tap % 3 - 1, dy = tap / 3 - 1;
      const bool ok = x + dx >= 0 && x + dx < a.W && y + dy >= 0 && y + dy < a.H;
      w[tap] = ok ? __ldg(reinterpret_cast<const float4*>(base + (long long)(dx + dy * a.W) * a.n3pad + tap * a.cstride)) : zero;
    }
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) { acc.x += w[tap].x; acc.y += w[tap].y; acc.z += w[tap].z; acc.w += w[tap].w; }
  } else
  for (int tap = 0; tap < a.taps; ++tap) {
    int dx, dy, dz;
    if (a.qsum) { dx = 0; dy = tap % 3 - 1; dz = (a.D > 1) ? tap / 3 - 1 : 0; }
    else chain_tap_offset(tap, a.ksz, a.D, dx, dy, dz);
    const int xx = x + dx, yy = y + dy, zz = z + dz;
    if (xx < 0 || xx >= a.W || yy < 0 || yy >= a.H || zz < 0 || zz >= a.D) continue;
    const float4 w = __ldg(reinterpret_cast<const float4*>(
        a.P + (m + dx + (long long)dy * a.W + (long long)dz * a.W * a.H) * a.n3pad + tap * a.cstride + 4 * g));
    acc.x += w.x; acc.y += w.y; acc.z += w.z; acc.w += w.w;
  }
  const float osc = col2im_scale(a);
  const float v4[4] = {acc.x * osc, acc.y * osc, acc.z * osc, acc.w * osc};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int n = 4 * g + j;
    if (n >= a.Cn) break;
    float v = v4[j];
    if (a.add && n < a.add_n) v += a.add[b * a.add_bs + (long long)n * a.px + pix];
    if (n < a.n0) {
      a.out0[b * a.out0_bs + (long long)n * a.px + pix] = v;
    } else {
      float* qq = a.out1 + b * a.out1_bs + (long long)(n - a.n0) * a.px + pix;
      *qq = a.out1_accum ? (*qq + v) : v;
    }
  }
}

// col2im + affine coupling (CouplingFuse) in two phases per block of 32 pixels.  Phase 1 is k_col2im_g: thread =
// (pixel, group of 4 output channels), 16-byte loads of the tap (row) sums, the same order of additions (bit-identical
// Y3), results into a [Cn][32 + 1] shared-memory tile.  Phase 2 walks (channel of the transformed half, pixel) with the
// warp's 32 lanes on 32 consecutive pixels, so every access to X1 / Y1 / dY1 / dY3 is one full 128-byte line.
// (A first version that kept the (pixel, channel pair) mapping for both phases was as slow as the two kernels it
// replaced: its plane accesses covered ~11 pixels per warp instruction.)  J = C1 / 2 = Cn / 4 groups per pixel.
__device__ __forceinline__ float cf_sigmoid(float x, float low, float high) { return low + (high - low) / (1.f + expf(-x)); }
template <int J, int MODE>
__global__ void __launch_bounds__(32 * J) k_col2im_coupling(const Col2imArgs a, const CouplingFuse f) {
  constexpr int CN = 4 * J;
  __shared__ float tile[CN][33];
  __shared__ float red[32];
  const int pl = threadIdx.x / J, g = threadIdx.x % J;
  const int pix0 = (int)(blockIdx.x * 32);
  const int pix = pix0 + pl;
  const long long b = blockIdx.y;
  if (pix < (int)a.px) {
    const long long m = b * a.px + pix;
    const int x = pix % a.W;
    const int yz = pix / a.W;
    const int y = yz % a.H;
    const int z = yz / a.H;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a.qsum && a.D == 1 && a.taps == 3) {
      const float* row = a.P + m * a.n3pad + 4 * g;
      const long long rs = (long long)a.W * a.n3pad;
      const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 w0 = (y > 0) ? __ldg(reinterpret_cast<const float4*>(row - rs)) : zero;
      const float4 w1 = __ldg(reinterpret_cast<const float4*>(row + a.cstride));
      const float4 w2 = (y + 1 < a.H) ? __ldg(reinterpret_cast<const float4*>(row + rs + 2 * a.cstride)) : zero;
      acc.x = (w0.x + w1.x) + w2.x; acc.y = (w0.y + w1.y) + w2.y;
      acc.z = (w0.z + w1.z) + w2.z; acc.w = (w0.w + w1.w) + w2.w;
    } else if (!a.qsum && a.D == 1 && a.taps == 9) {
      // the full 3x3 P (GEMM3 wider than 128 columns): all nine loads in flight, added in the rolled loop's order
      const float* base = a.P + m * a.n3pad + 4 * g;
      const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
      float4 w[9];
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int dx = tap % 3 - 1, dy = tap / 3 - 1;
        const bool ok = x + dx >= 0 && x + dx < a.W && y + dy >= 0 && y + dy < a.H;
        w[tap] = ok ? __ldg(reinterpret_cast<const float4*>(base + (long long)(dx + dy * a.W) * a.n3pad + tap * a.cstride)) : zero;
      }
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) { acc.x += w[tap].x; acc.y += w[tap].y; acc.z += w[tap].z; acc.w += w[tap].w; }
    } else {
      for (int tap = 0; tap < a.taps; ++tap) {
        int dx, dy, dz;
        if (a.qsum) { dx = 0; dy = tap % 3 - 1; dz = (a.D > 1) ? tap / 3 - 1 : 0; }
        else chain_tap_offset(tap, a.ksz, a.D, dx, dy, dz);
        const int xx = x + dx, yy = y + dy, zz = z + dz;
        if (xx < 0 || xx >= a.W || yy < 0 || yy >= a.H || zz < 0 || zz >= a.D) continue;
        const float4 w = __ldg(reinterpret_cast<const float4*>(
            a.P + (m + dx + (long long)dy * a.W + (long long)dz * a.W * a.H) * a.n3pad + tap * a.cstride + 4 * g));
        acc.x += w.x; acc.y += w.y; acc.z += w.z; acc.w += w.w;
      }
    }
    const float osc = col2im_scale(a);
    tile[4 * g][pl] = acc.x * osc;
    tile[4 * g + 1][pl] = acc.y * osc;
    tile[4 * g + 2][pl] = acc.z * osc;
    tile[4 * g + 3][pl] = acc.w * osc;
  }
  __syncthreads();
  float lsum = 0.f, gmax = 0.f;
  const int C1 = 2 * J;
  for (int i = threadIdx.x; i < C1 * 32; i += 32 * J) {
    const int ch = i >> 5, p2 = i & 31;
    const int px2 = pix0 + p2;
    if (px2 >= (int)a.px) continue;
    const float lsv = tile[ch][p2], tvv = tile[C1 + ch][p2];
    float* ap = f.a1 + b * f.a1_bs + (long long)ch * a.px + px2;
    const float S = cf_sigmoid(fmaxf(lsv, 0.f), f.low, f.high);  // RB output ReLU (layer_residual_block.jl:133)
    const float T = fmaxf(tvv, 0.f);
    if (MODE == 0) {
      *ap = S * *ap + T;                 // invertible_layer_glow.jl:112
      lsum += logf(fabsf(S));            // :210
    } else if (MODE == 1) {
      *ap = (*ap - T) / (S + 1.1920929e-07f);  // :127, eps(Float32)
    } else {
      float* dp = f.d1 + b * f.d1_bs + (long long)ch * a.px + px2;
      const float dyv = *dp;
      const float X1 = (*ap - T) / (S + 1.1920929e-07f);
      float dS = dyv * X1;               // :144
      dS -= f.invB / S;                  // :145-147, 211
      *ap = X1;
      *dp = dyv * S;                     // :149
      const float e = (f.high - S) / (S - f.low);  // activation_functions.jl:213-217 through the logit
      const float dl = (f.high - f.low) * dS * e / ((1.f + e) * (1.f + e));
      const float gl = (lsv < 0.f) ? 0.f : dl;  // _relugrad of the block's output ReLU (:84)
      const float gt = (tvv < 0.f) ? 0.f : dyv; // dT = dY1 (:143)
      f.dY3[(b * 2 * C1 + ch) * a.px + px2] = gl;
      f.dY3[(b * 2 * C1 + C1 + ch) * a.px + px2] = gt;
      gmax = fmaxf(gmax, fmaxf(fabsf(gl), fabsf(gt)));
    }
  }
  if (MODE == 0 && f.ld) {
#pragma unroll
    for (int o = 16; o; o >>= 1) lsum += __shfl_xor_sync(0xFFFFFFFFu, lsum, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = lsum;
    __syncthreads();
    if (threadIdx.x == 0) {
      double r = 0.0;
      for (int w = 0; w < J; ++w) r += (double)red[w];
      atomicAdd(f.ld, r * (double)f.invB);
    }
  }
  if (MODE == 2 && f.amax) {
#pragma unroll
    for (int o = 16; o; o >>= 1) gmax = fmaxf(gmax, __shfl_xor_sync(0xFFFFFFFFu, gmax, o));
    if ((threadIdx.x & 31) == 0 && gmax > 0.f) atomicMax(f.amax, __float_as_uint(gmax));
  }
}
// The same fused step with ONE THREAD PER PIXEL, for the 2-D tap-row planes of Pq (every plane is [M][cq] dense, so the
// cq floats of consecutive pixels are contiguous: the warp's 16-byte loads cover whole lines, and with lanes along the
// pixels every access to the (B, C, px) tensors X1 / Y1 / dY1 / dY3 is a full 128-byte line too).  No shared-memory
// transpose, no block barrier, 3 * CN / 4 independent 16-byte loads in flight per thread (the two-phase kernel above
// has 3 and is latency-bound at a third of the HBM rate).  Same order of additions and the same formulas: bit-identical
// to k_col2im_coupling.  C1 = channels of the transformed half (CN = 2 C1 = cq columns per plane).
template <int C1, int MODE>
__global__ void __launch_bounds__(128) k_col2im_coupling_px(const Col2imArgs a, const CouplingFuse f) {
  constexpr int CN = 2 * C1, NV = CN / 4;
  __shared__ float red[4];
  const int pix = (int)(blockIdx.x * 128 + threadIdx.x);
  const long long b = blockIdx.y;
  float lsum = 0.f, gmax = 0.f;
  if (pix < (int)a.px) {
    const long long m = b * a.px + pix;
    const int y = (pix / a.W) % a.H;
    const float* row = a.P + m * CN;
    const long long rs = (long long)a.W * CN;
    const bool up = y > 0, dn = y + 1 < a.H;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 w0[NV], w1[NV], w2[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) w1[v] = __ldg(reinterpret_cast<const float4*>(row + a.cstride) + v);
#pragma unroll
    for (int v = 0; v < NV; ++v) w0[v] = up ? __ldg(reinterpret_cast<const float4*>(row - rs) + v) : zero;
#pragma unroll
    for (int v = 0; v < NV; ++v) w2[v] = dn ? __ldg(reinterpret_cast<const float4*>(row + rs + 2 * a.cstride) + v) : zero;
    float* ap = f.a1 + b * f.a1_bs + pix;
    float* dp = (MODE == 2) ? f.d1 + b * f.d1_bs + pix : nullptr;
    float av[C1], dv[C1];
#pragma unroll
    for (int ch = 0; ch < C1; ++ch) {
      av[ch] = ap[(long long)ch * a.px];
      if (MODE == 2) dv[ch] = dp[(long long)ch * a.px];
    }
    const float osc = col2im_scale(a);
    float y3[CN];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      y3[4 * v] = ((w0[v].x + w1[v].x) + w2[v].x) * osc;
      y3[4 * v + 1] = ((w0[v].y + w1[v].y) + w2[v].y) * osc;
      y3[4 * v + 2] = ((w0[v].z + w1[v].z) + w2[v].z) * osc;
      y3[4 * v + 3] = ((w0[v].w + w1[v].w) + w2[v].w) * osc;
    }
#pragma unroll
    for (int ch = 0; ch < C1; ++ch) {
      const float lsv = y3[ch], tvv = y3[C1 + ch];
      const float S = cf_sigmoid(fmaxf(lsv, 0.f), f.low, f.high);  // RB output ReLU (layer_residual_block.jl:133)
      const float T = fmaxf(tvv, 0.f);
      if (MODE == 0) {
        ap[(long long)ch * a.px] = S * av[ch] + T;       // invertible_layer_glow.jl:112
        lsum += logf(fabsf(S));                          // :210
      } else if (MODE == 1) {
        ap[(long long)ch * a.px] = (av[ch] - T) / (S + 1.1920929e-07f);  // :127, eps(Float32)
      } else {
        const float dyv = dv[ch];
        const float X1 = (av[ch] - T) / (S + 1.1920929e-07f);
        float dS = dyv * X1;               // :144
        dS -= f.invB / S;                  // :145-147, 211
        ap[(long long)ch * a.px] = X1;
        dp[(long long)ch * a.px] = dyv * S;  // :149
        const float e = (f.high - S) / (S - f.low);  // activation_functions.jl:213-217 through the logit
        const float dl = (f.high - f.low) * dS * e / ((1.f + e) * (1.f + e));
        const float gl = (lsv < 0.f) ? 0.f : dl;  // _relugrad of the block's output ReLU (:84)
        const float gt = (tvv < 0.f) ? 0.f : dyv; // dT = dY1 (:143)
        f.dY3[(b * CN + ch) * a.px + pix] = gl;
        f.dY3[(b * CN + C1 + ch) * a.px + pix] = gt;
        gmax = fmaxf(gmax, fmaxf(fabsf(gl), fabsf(gt)));
      }
    }
  }
  if (MODE == 0 && f.ld) {
#pragma unroll
    for (int o = 16; o; o >>= 1) lsum += __shfl_xor_sync(0xFFFFFFFFu, lsum, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = lsum;
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd(f.ld, ((double)red[0] + (double)red[1] + (double)red[2] + (double)red[3]) * (double)f.invB);
  }
  if (MODE == 2 && f.amax) {
#pragma unroll
    for (int o = 16; o; o >>= 1) gmax = fmaxf(gmax, __shfl_xor_sync(0xFFFFFFFFu, gmax, o));
    if ((threadIdx.x & 31) == 0 && gmax > 0.f) atomicMax(f.amax, __float_as_uint(gmax));
  }
}
template <int C1>
static void launch_col2im_coupling_px(cudaStream_t st, const Col2imArgs& ca, const CouplingFuse& f, int B) {
  const dim3 grid((unsigned)cdiv(ca.px, 128), (unsigned)B, 1);
  if (f.mode == 0) k_col2im_coupling_px<C1, 0><<<grid, 128, 0, st>>>(ca, f);
  else if (f.mode == 1) k_col2im_coupling_px<C1, 1><<<grid, 128, 0, st>>>(ca, f);
  else k_col2im_coupling_px<C1, 2><<<grid, 128, 0, st>>>(ca, f);
}

template <int J>
static void launch_col2im_coupling(cudaStream_t st, const dim3& grid, const Col2imArgs& ca, const CouplingFuse& f) {
  if (f.mode == 0) k_col2im_coupling<J, 0><<<grid, 32 * J, 0, st>>>(ca, f);
  else if (f.mode == 1) k_col2im_coupling<J, 1><<<grid, 32 * J, 0, st>>>(ca, f);
  else k_col2im_coupling<J, 2><<<grid, 32 * J, 0, st>>>(ca, f);
}
static bool col2im_coupling_enabled() {
  static const bool on = [] { const char* e = getenv("INB_FUSE_COUPLING"); return !(e && e[0] == '0'); }();
  return on;
}

// ---------------------------------------------------------------- weight packing for one chain pass
// One launch packs the three operands of a pass into bf16 hi/lo planes:
//   w1 [nh][kp]     dense-K rows against the im2col operand:  column tap*C1 + c  <-  wa[n][c][T-1-tap]   (NNlib conv)
//   w2 [nh][nh]     the 1x1 contraction + I (residual skip / '+ dY2'):  conv: wb[n][c],  data: wb[c][n]
//   w3 [n3pad][nh]  tap-expanded rows of the \nabla conv_data contraction:  row tap*Cn + n  <-  wc[c][n][tap]
// (wa = W1, wc = W3 in the forward pass; wa = W3, wc = W1 in the backward pass; reference layout w[d0][d1][T])
struct PackChainArgs {
  int f16;       // INB_PREC_FP16X3: IEEE-half planes of kF16WScale * w
  int nhr;       // the block's n_hidden: the reference arrays have nhr hidden channels, the planes nh (zero padded)
  int nh, T, C1, kp, Cn, n3pad, w2_data;
  const float *wa, *wb, *wc;
  __nv_bfloat16 *w1h, *w1l, *w2h, *w2l, *w3h, *w3l;
};
__global__ void k_pack_chain_tc(const PackChainArgs a) {
  const long long n1 = (long long)a.nh * a.kp, n2 = (long long)a.nh * a.nh, n3 = (long long)a.n3pad * a.nh;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n1 + n2 + n3;
       i += (long long)gridDim.x * blockDim.x) {
    float v = 0.f;
    __nv_bfloat16 *dh, *dl;
    long long o;
    if (i < n1) {
      o = i;
      const int k = (int)(i % a.kp), n = (int)(i / a.kp);
      if (k < a.T * a.C1 && n < a.nhr) {
        const int tap = k / a.C1, cc = k - tap * a.C1;
        v = a.wa[((long long)n * a.C1 + cc) * a.T + (a.T - 1 - tap)];
      }
      dh = a.w1h; dl = a.w1l;
    } else if (i < n1 + n2) {
      o = i - n1;
      const int cc = (int)(o % a.nh), n = (int)(o / a.nh);
      if (n < a.nhr && cc < a.nhr) {
        v = a.w2_data ? a.wb[(long long)cc * a.nhr + n] : a.wb[(long long)n * a.nhr + cc];
        if (n == cc) v += 1.f;
      }
      dh = a.w2h; dl = a.w2l;
    } else {
      o = i - n1 - n2;
      const int c = (int)(o % a.nh), r = (int)(o / a.nh);
      if (r < a.T * a.Cn && c < a.nhr) {
        const int tap = r / a.Cn, n = r - tap * a.Cn;
        v = a.wc[((long long)c * a.Cn + n) * a.T + tap];
      }
      dh = a.w3h; dl = a.w3l;
    }
    if (a.f16) {
      uint32_t hw, lw;
      split2<true>(v * kF16WScale, 0.f, hw, lw);
      reinterpret_cast<unsigned short*>(dh)[o] = (unsigned short)(hw & 0xFFFFu);
      reinterpret_cast<unsigned short*>(dl)[o] = (unsigned short)(lw & 0xFFFFu);
    } else {
      __nv_bfloat16 h, l;
      split_bf16(v, h, l);
      dh[o] = h;
      dl[o] = l;
    }
  }
}
constexpr int kPackMulti = 24;
struct PackMultiArgs {
  int f16;
  int nhr;
  int nh, T, C1, kp, Cn, n3pad, w2_data, n, blocks_per_item;
  const float* wa[kPackMulti];
  const float* wb[kPackMulti];
  const float* wc[kPackMulti];
  __nv_bfloat16* dst[kPackMulti][6];  // w1h, w1l, w2h, w2l, w3h, w3l
};
__global__ void k_pack_chain_multi(const __grid_constant__ PackMultiArgs a) {
  const int item = blockIdx.x / a.blocks_per_item, lb = blockIdx.x - item * a.blocks_per_item;
  const float *wa = a.wa[item], *wb = a.wb[item], *wc = a.wc[item];
  const long long n1 = (long long)a.nh * a.kp, n2 = (long long)a.nh * a.nh, n3 = (long long)a.n3pad * a.nh;
  for (long long i = lb * (long long)blockDim.x + threadIdx.x; i < n1 + n2 + n3; i += (long long)a.blocks_per_item * blockDim.x) {
    float v = 0.f;
    int which;
    long long o;
    if (i < n1) {
      o = i;
      const int k = (int)(i % a.kp), n = (int)(i / a.kp);
      if (k < a.T * a.C1 && n < a.nhr) {
        const int tap = k / a.C1, cc = k - tap * a.C1;
        v = wa[((long long)n * a.C1 + cc) * a.T + (a.T - 1 - tap)];
      }
      which = 0;
    } else if (i < n1 + n2) {
      o = i - n1;
      const int cc = (int)(o % a.nh), n = (int)(o / a.nh);
      if (n < a.nhr && cc < a.nhr) {
        v = a.w2_data ? wb[(long long)cc * a.nhr + n] : wb[(long long)n * a.nhr + cc];
        if (n == cc) v += 1.f;
      }
      which = 2;
    } else {
      o = i - n1 - n2;
      const int c = (int)(o % a.nh), r = (int)(o / a.nh);
      if (r < a.T * a.Cn && c < a.nhr) {
        const int tap = r / a.Cn, n = r - tap * a.Cn;
        v = wc[((long long)c * a.Cn + n) * a.T + tap];
      }
      which = 4;
    }
    if (a.f16) {
      uint32_t hw, lw;
      split2<true>(v * kF16WScale, 0.f, hw, lw);
      reinterpret_cast<unsigned short*>(a.dst[item][which])[o] = (unsigned short)(hw & 0xFFFFu);
      reinterpret_cast<unsigned short*>(a.dst[item][which + 1])[o] = (unsigned short)(lw & 0xFFFFu);
    } else {
      __nv_bfloat16 h, l;
      split_bf16(v, h, l);
      a.dst[item][which][o] = h;
      a.dst[item][which + 1][o] = l;
    }
  }
}
void op_pack_chain_multi(Ctx& c, int nh, int nhr, int T, int C1, int kp, int w2_data, int Cn, int n3pad,
                         const PackChainItem* items, int n) {
  if (c.dry()) return;
  const long long per = (long long)nh * kp + (long long)nh * nh + (long long)n3pad * nh;
  for (int i0 = 0; i0 < n; i0 += kPackMulti) {
    PackMultiArgs a{};
    a.f16 = prec_f16(c.prec) ? 1 : 0;
    a.nhr = nhr;
    a.nh = nh; a.T = T; a.C1 = C1; a.kp = kp; a.Cn = Cn; a.n3pad = n3pad; a.w2_data = w2_data;
    a.n = std::min(kPackMulti, n - i0);
    a.blocks_per_item = (int)std::min<long long>(cdiv(per, 256), 64);
    for (int j = 0; j < a.n; ++j) {
      const PackChainItem& it = items[i0 + j];
      a.wa[j] = it.wa; a.wb[j] = it.wb; a.wc[j] = it.wc;
      a.dst[j][0] = it.w1.hi; a.dst[j][1] = it.w1.lo; a.dst[j][2] = it.w2.hi; a.dst[j][3] = it.w2.lo;
      a.dst[j][4] = it.w3.hi; a.dst[j][5] = it.w3.lo;
    }
    Prof pf(c, F_PACK, 1, 0, 8.0 * per * a.n);
    k_pack_chain_multi<<<(unsigned)(a.n * a.blocks_per_item), 256, 0, c.st>>>(a);
    INB_CUDA(cudaGetLastError());
  }
}
void op_pack_chain_tc(Ctx& c, int nh, int nhr, int T, int C1, int kp, const float* wa, const float* wb, int w2_data, int Cn,
                      int n3pad, const float* wc, Planes w1, Planes w2, Planes w3) {
  if (c.dry()) return;
  PackChainArgs a{prec_f16(c.prec) ? 1 : 0, nhr, nh, T, C1, kp, Cn, n3pad, w2_data, wa, wb, wc, w1.hi, w1.lo, w2.hi, w2.lo, w3.hi, w3.lo};
  const long long n = (long long)nh * kp + (long long)nh * nh + (long long)n3pad * nh;
  Prof pf(c, F_PACK, 1, 0, 8.0 * n);
  k_pack_chain_tc<<<(unsigned)std::min<long long>(cdiv(n, 256), 148 * 8), 256, 0, c.st>>>(a);
  INB_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------- host side
static long long* g_chain_trace = nullptr;
static int g_chain_trace_launch = 0;  // successive launches write successive 32 x 16 blocks (4 of them, cyclic)
void chain_set_trace(long long* p) { g_chain_trace = p; g_chain_trace_launch = 0; }

// columns of the tap-expanded GEMM: a multiple of 16; beyond 256 a multiple of 32, so that the CTA-pair kernel can run
// it as two equal passes of a multiple of 16 columns
// beyond 256 columns the CTA-pair kernel runs GEMM3 in ceil(n / 256) equal passes of a multiple of 16 columns each
// through the same TMEM region (cfg2 scale 3: 432 -> 2 x 224; 3-D blocks with 27 taps x 32 channels: 864 -> 4 x 224)
int chain_n3pad(int taps, int Cn) {
  const int n = (taps * Cn + 15) / 16 * 16;
  if (n <= 256) return n;
  const int pieces = (n + 255) / 256;
  return (n + 16 * pieces - 1) / (16 * pieces) * (16 * pieces);
}

int chain_kpad(int taps, int C, int extra) { return (taps * C + extra + 63) / 64 * 64; }

// INB_PLANE_LO8=0 keeps 2-byte lo planes.  The single-CTA kernels (INB_CHAIN_KERNEL=t|smem, diagnostics) write 2-byte planes.
bool chain_planes_lo8(int prec) {
  static const bool on = [] {
    const char* e = getenv("INB_PLANE_LO8");
    const char* k = getenv("INB_CHAIN_KERNEL");
    const char* sm = getenv("INB_CHAIN_SMEM");
    return !(e && e[0] == '0') && !(k && (k[0] == 't' || k[0] == 's')) && !(sm && sm[0] == '1');
  }();
  return on && prec_f16(prec);  // the upper byte of an IEEE half is sign + exponent + 2 mantissa bits; of a bfloat16 it is not
}

bool chain_supported(const Geo& g, int B, int k1, int k2, int nh, int C_in, int Cn) {
  if (k2 != 1) return false;
  if (nh < 1 || nh > 256) return false;  // runs at 128 or 256 hidden channels (chain_nh_pad), smaller blocks zero padded
  const int taps = k1 == 1 ? 1 : (g.nd == 3 ? 27 : 9);
  if (chain_kpad(taps, C_in, 1) > 1024) return false;
  const int n3 = chain_n3pad(taps, Cn);
  if (n3 > 1024 || Cn > 128) return false;
  (void)B;
  return true;
}

void op_rb_chain(Ctx& c, const ChainSpec& s) {
  const int taps = s.k1 == 1 ? 1 : (s.g.nd == 3 ? 27 : 9);
  INB_CHECK(s.in.pitch % 64 == 0 && s.w1.pitch == s.in.pitch, "fused ResidualBlock chain: the im2col width must be a multiple of 64");
  if (c.dry()) return;
  const int NT = prec_terms(c.prec);
  const bool f16 = prec_f16(c.prec);
  const int NP = NT == 1 ? 1 : 2;
  ChainArgs a{};
  a.wsinv = f16 ? 1.f / kF16WScale : 1.f;
  a.W = s.g.W; a.H = s.g.H; a.D = s.g.D;
  a.M = s.g.px * s.B;
  a.ntiles = (int)cdiv(a.M, 128);
  a.nkb1 = s.in.pitch / 64;
  INB_CHECK(s.nh == 128 || s.nh == 256, "fused ResidualBlock chain: the planes have 128 or 256 hidden channels");
  a.nh = s.nh;
  a.nh_bias = s.nh_real > 0 ? s.nh_real : s.nh;
  a.nchunk = s.nh / 64;
  a.n3pad = chain_n3pad(taps, s.Cn);
  a.n3a = std::min(a.n3pad, 256);
  a.n3b = a.n3pad - a.n3a;
  a.np3 = (a.n3pad + 127) / 128;
  a.mode = s.mode;
  a.store = (s.o1.hi != nullptr) ? 1 : 0;
  a.bias1 = s.bias1; a.bias2 = s.bias2;
  a.mask1 = s.mask1; a.mask2 = s.mask2;
  a.bits1 = s.bits1; a.bits2 = s.bits2;
  INB_CHECK(s.mode == 0 || (s.mask1 && s.mask2), "fused ResidualBlock chain: the backward pass needs both masks");
  INB_CHECK(!(s.mode == 0 && s.o1.hi) || (s.bits1 && s.bits2), "fused ResidualBlock chain: a storing forward pass writes the mask bit planes");
  a.P = s.P;
  static const int chain_exp = [] { const char* e = getenv("INB_CHAIN_EXP"); return e ? atoi(e) : 0; }();
  a.exp = chain_exp;
  a.trace = g_chain_trace ? g_chain_trace + (size_t)(g_chain_trace_launch++ % 4) * 512 : nullptr;
  const size_t stage = (size_t)NP * kPlane, chunk = (size_t)NP * kPlane;
  const size_t aux = 48 * 8 + 16 + 512 * 4;
  const size_t cap = 227 * 1024;
  // kernel choice: CTA pairs (k_rb_chain2) whenever GEMM3 fits one 256-column region; INB_CHAIN_KERNEL=t|smem forces
  // the single-CTA kernels (hidden operand in tensor memory / in shared memory)
  static const int force = [] {
    const char* e = getenv("INB_CHAIN_KERNEL");
    if (!e) e = getenv("INB_CHAIN_SMEM") && getenv("INB_CHAIN_SMEM")[0] == '1' ? "smem" : "";
    return e[0] == 't' ? 1 : (e[0] == 's' ? 2 : 0);
  }();
  a.npiece = (a.n3pad + 255) / 256;
  a.n3piece = a.n3pad / a.npiece;
  const bool pair = (a.n3pad % (16 * a.npiece) == 0) && force == 0;
  if (a.store) {
    INB_CHECK(s.o1.lo8 == s.o2.lo8, "fused ResidualBlock chain: both stored tensors must use the same lo-plane format");
    INB_CHECK(!s.o1.lo8 || (pair && f16), "fused ResidualBlock chain: 1-byte lo planes need the CTA-pair kernel and fp16x3");
    a.lo8 = s.o1.lo8;
  }
  INB_CHECK(pair || (!f16 && a.n3pad <= 480), "fused ResidualBlock chain: %d expanded columns need the CTA-pair kernel", a.n3pad);
  static const bool no_qsum = [] { const char* e = getenv("INB_CHAIN_QSUM"); return e && e[0] == '0'; }();
  a.Cn = s.Cn;
  a.cq = (s.Cn + 3) / 4 * 4;
  a.ntg = taps / 3;
  a.nq = a.ntg * a.cq;
  const bool q_geo = pair && !no_qsum && taps >= 9 && 128 % s.g.W == 0 && (s.g.W & (s.g.W - 1)) == 0;
  a.qsum = (q_geo && a.n3pad <= 128) ? 1 : 0;
  a.pstag_bytes = a.qsum ? (int)((((size_t)128 * (a.n3pad + 4) + (size_t)128 * a.nq) * 4 + 1023) / 1024 * 1024) : 4 * (int)kPlane;
  static const bool no_qwide = [] { const char* e = getenv("INB_CHAIN_QWIDE"); return e && e[0] == '0'; }();
  if (q_geo && !no_qwide && !a.qsum && taps == 9 && a.npiece == 1 && s.Cn % 4 == 0) {
    // one tap row at a time: the aligned 32-column chunks covering 3 Cn columns at any of the three offsets
    int nck = 0;
    for (int tg = 0; tg < 3; ++tg) nck = std::max(nck, ((3 * tg * s.Cn) % 32 + 3 * s.Cn + 31) / 32);
    const int rpw = 32 * nck + 4;
    const int bytes = (int)((((size_t)128 * rpw + (size_t)128 * a.nq) * 4 + 1023) / 1024 * 1024);
    const size_t fixed_w = (size_t)(a.store ? kChain2Slots : 0) * (a.lo8 ? 3 * (size_t)kPlane / 2 : 2 * (size_t)kPlane) + bytes;
    if (nck <= 4 && cap > aux + fixed_w && (cap - aux - fixed_w) / kPlane >= 4) {
      a.qsum = 2;
      a.rpw = rpw;
      a.pstag_bytes = bytes;
    }
  }
  const bool tmem_a = a.n3pad <= 256 && force != 2;
  const size_t slot_bytes = a.lo8 ? 3 * (size_t)kPlane / 2 : 2 * (size_t)kPlane;
  const size_t fixed = pair ? (size_t)(a.store ? kChain2Slots : 0) * slot_bytes + a.pstag_bytes
                            : (tmem_a ? (size_t)4 * kPlane : a.nchunk * chunk);
  // k_rb_chain2 runs its ring in 16 KB granules (up to 12 of them)
  int stages = (int)((cap - aux - fixed) / (pair ? (size_t)kPlane : stage));
  if (stages > (pair ? 12 : 8)) stages = pair ? 12 : 8;
  if (const char* e = getenv("INB_CHAIN_MAXSTAGES")) {  // sensitivity studies: cap the ring depth
    const int v = atoi(e);
    if (v >= 4 && v < stages) stages = v;
  }
  // the im2col block of GEMM1 stays resident while the weight blocks stream past it
  INB_CHECK(stages >= (pair ? 4 : 1 + s.nh / 128), "fused ResidualBlock chain: shared memory does not fit");
  a.stages = stages;
  const size_t smem = fixed + stages * (pair ? (size_t)kPlane : stage) + aux;
  ChainMaps mp{};
  const int wrows = pair ? s.nh / 4 : 128, w3rows = pair ? a.n3piece / 2 : 128;
  for (int pl = 0; pl < 2; ++pl) {
    mp.A[pl] = make_rows_map(pl ? s.in.lo : s.in.hi, s.in.pitch, a.M, 64, 128);
    mp.W1[pl] = make_rows_map(pl ? s.w1.lo : s.w1.hi, s.in.pitch, s.nh, 64, wrows);
    mp.W2[pl] = make_rows_map(pl ? s.w2.lo : s.w2.hi, s.nh, s.nh, 64, wrows);
    mp.W3[pl] = make_rows_map(pl ? s.w3.lo : s.w3.hi, s.nh, a.n3pad, 64, w3rows);
    if (a.store && pl == 1 && s.o1.lo8) {
      mp.O1[1] = make_rows_map_u8(s.o1.lo, s.nh, a.M, 64, 128, true);
      mp.O2[1] = make_rows_map_u8(s.o2.lo, s.nh, a.M, 64, 128, true);
    } else if (a.store) {
      mp.O1[pl] = make_rows_map(pl ? s.o1.lo : s.o1.hi, s.nh, a.M, 64, 128);
      mp.O2[pl] = make_rows_map(pl ? s.o2.lo : s.o2.hi, s.nh, a.M, 64, 128);
    } else {
      mp.O1[pl] = mp.W2[pl];
      mp.O2[pl] = mp.W2[pl];
    }
  }
  mp.P = a.qsum ? make_planes_map_f32_dense(s.P, a.cq, a.M, a.ntg, 128) : make_rows_map_f32(s.P, a.n3pad, a.M, 32, 128);
  static const int l2_hints = [] { const char* e = getenv("INB_L2_HINTS"); return e ? atoi(e) : 0; }();
  a.hints = l2_hints;
  unsigned grid = (unsigned)std::min(a.ntiles, 148);
  const double flops = 2.0 * a.M * ((double)s.in.pitch * s.nh + (double)s.nh * s.nh + (double)s.nh * a.n3pad) * NT;
  {
    Prof pf(c, F_CONV_TC, 1, flops, 0);
    if (pair) {
      auto kern = a.trace ? (f16 ? k_rb_chain2<3, true, true> : ((NT == 3) ? k_rb_chain2<3, true> : k_rb_chain2<1, true>))
                          : (f16 ? k_rb_chain2<3, false, true> : ((NT == 3) ? k_rb_chain2<3, false> : k_rb_chain2<1, false>));
      INB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      // resident CTA pairs: one CTA per SM, pairs are placed inside a GPC (the query accounts for odd GPCs)
      static int max_pairs[6] = {0, 0, 0, 0, 0, 0};
      int& mpairs = max_pairs[(NT == 3) + (f16 ? 1 : 0) + 3 * (a.trace != nullptr)];
      if (mpairs == 0) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(148);
        cfg.blockDim = dim3(kChain2Threads);
        cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at{};
        at.id = cudaLaunchAttributeClusterDimension;
        at.val.clusterDim.x = 2; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
        cfg.attrs = &at;
        cfg.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n <= 0) { cudaGetLastError(); n = 64; }
        mpairs = std::min(n, 74);
      }
      int use_pairs = mpairs;
      {  // SMs left to the weight-gradient kernels that run under this pass (conv_tc_wgrad2.cu: wgrad_overlap_ctas)
        const int w = wgrad_overlap_ctas();
        if (w > 0 && c.lane && a.ntiles >= 4 * 148) use_pairs = std::max(8, std::min(use_pairs, (148 - 3 * w) / 2));
      }
      if (const char* e = getenv("INB_CHAIN_MAXPAIRS")) {  // tests: few resident pairs = many tiles per pair on small inputs
        const int v = atoi(e);
        if (v > 0 && v < use_pairs) use_pairs = v;
      }
      grid = 2u * (unsigned)std::min((a.ntiles + 1) / 2, use_pairs);
      kern<<<grid, kChain2Threads, smem, c.st>>>(mp, a);
    } else if (tmem_a && NT == 3) {
      INB_CUDA(cudaFuncSetAttribute(k_rb_chain_t<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_rb_chain_t<3><<<grid, kChainThreads, smem, c.st>>>(mp, a);
    } else if (tmem_a) {
      INB_CUDA(cudaFuncSetAttribute(k_rb_chain_t<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_rb_chain_t<1><<<grid, kChainThreads, smem, c.st>>>(mp, a);
    } else if (NT == 3) {
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
    ca.qsum = a.qsum; ca.cstride = s.Cn;
    if (a.qsum) { ca.taps = a.ntg; ca.n3pad = a.cq; ca.cstride = a.M * a.cq; }
    ca.px = s.g.px; ca.M = a.M; ca.P = s.P;
    ca.out0 = s.out0; ca.out0_bs = s.out0_bs; ca.n0 = s.n0;
    ca.out1 = s.out1; ca.out1_bs = s.out1_bs; ca.out1_accum = s.out1_accum;
    ca.add = s.add; ca.add_bs = s.add_bs; ca.add_n = s.add_n;
    ca.oscale = f16 ? 1.f / kF16WScale : 1.f;
    ca.smax = f16 ? s.smax : nullptr;
    const int prow = a.qsum ? a.nq : a.n3pad;  // floats of P per pixel (byte accounting)
    if (s.fuse) s.fuse->done = false;
    if (s.fuse && col2im_coupling_enabled() && s.Cn == 2 * s.fuse->C1 && s.fuse->C1 % 2 == 0 && !s.out1 && !s.add &&
        ca.cstride % 4 == 0 && ca.n3pad % 4 == 0 && s.Cn % 4 == 0 && s.B <= 65535 && s.g.px < (1ll << 31)) {
      const int J = s.fuse->C1 / 2;
      const dim3 grid((unsigned)cdiv(s.g.px, 32), (unsigned)s.B, 1);
      // bytes: the P rows, the transformed half read + written (twice in the backward), dY3 written in the backward
      const double elems = (double)a.M * s.fuse->C1;
      Prof pf(c, s.fuse->mode == 2 ? F_COUPLING_BWD : (s.fuse->mode == 1 ? F_COUPLING_INV : F_COUPLING_FWD), 1, 0,
              4.0 * prow * a.M + (s.fuse->mode == 2 ? 24.0 : 8.0) * elems);
      bool ok = true;
      static const bool no_px = [] { const char* e = getenv("INB_COUPLING_PX"); return e && e[0] == '0'; }();
      // thread-per-pixel variant: 2-D tap-row planes whose cq equals the 2 C1 channels of the block output
      const bool px_ok = !no_px && a.qsum && s.g.D == 1 && ca.taps == 3 && ca.n3pad == s.Cn && (J == 1 || J == 2 || J == 3 || J == 4 || J == 6);
      if (px_ok) {
        switch (J) {
          case 1: launch_col2im_coupling_px<2>(c.st, ca, *s.fuse, s.B); break;
          case 2: launch_col2im_coupling_px<4>(c.st, ca, *s.fuse, s.B); break;
          case 3: launch_col2im_coupling_px<6>(c.st, ca, *s.fuse, s.B); break;
          case 4: launch_col2im_coupling_px<8>(c.st, ca, *s.fuse, s.B); break;
          default: launch_col2im_coupling_px<12>(c.st, ca, *s.fuse, s.B); break;
        }
      } else
      switch (J) {
        case 1: launch_col2im_coupling<1>(c.st, grid, ca, *s.fuse); break;
        case 2: launch_col2im_coupling<2>(c.st, grid, ca, *s.fuse); break;
        case 3: launch_col2im_coupling<3>(c.st, grid, ca, *s.fuse); break;
        case 4: launch_col2im_coupling<4>(c.st, grid, ca, *s.fuse); break;
        case 6: launch_col2im_coupling<6>(c.st, grid, ca, *s.fuse); break;
        case 8: launch_col2im_coupling<8>(c.st, grid, ca, *s.fuse); break;
        case 12: launch_col2im_coupling<12>(c.st, grid, ca, *s.fuse); break;
        case 16: launch_col2im_coupling<16>(c.st, grid, ca, *s.fuse); break;
        case 24: launch_col2im_coupling<24>(c.st, grid, ca, *s.fuse); break;
        default: ok = false;
      }
      if (ok) {
        INB_CUDA(cudaGetLastError());
        s.fuse->done = true;
        return;
      }
    }
    Prof pf(c, F_COL2IM, 1, 0, (4.0 * prow + 4.0 * s.Cn) * a.M);
    const unsigned nb = (unsigned)cdiv(a.M, 128);
    const bool v4 = s.Cn % 4 == 0 && ca.cstride % 4 == 0 && ca.n3pad % 4 == 0;
    const bool v2 = s.Cn % 2 == 0 && ca.cstride % 2 == 0 && ca.n3pad % 2 == 0;
    // 16-byte groups: the tap-row sums are padded to 4 channels per tap row; the full P needs Cn % 4 == 0
    const bool g4 = ca.cstride % 4 == 0 && ca.n3pad % 4 == 0 && (a.qsum || s.Cn % 4 == 0) && s.B <= 65535 &&
                    s.g.px < (1ll << 31);
    const int G = (s.Cn + 3) / 4;
    const dim3 nbg((unsigned)cdiv(s.g.px, 32), (unsigned)s.B, 1);

    static const bool no_g = [] { const char* e = getenv("INB_COL2IM_G"); return e && e[0] == '0'; }();
    static const bool no_q2 = [] { const char* e = getenv("INB_COL2IM_Q2"); return e && e[0] == '0'; }();
    const bool q2 = g4 && !no_g && !no_q2 && a.qsum && s.g.D == 1 && ca.taps == 3;
    if (q2 && G == 2) k_col2im_g<2, true><<<nbg, 64, 0, c.st>>>(ca);
    else if (q2 && G == 3) k_col2im_g<3, true><<<nbg, 96, 0, c.st>>>(ca);
    else if (q2 && G == 6) k_col2im_g<6, true><<<nbg, 192, 0, c.st>>>(ca);
    else if (q2 && G == 12) k_col2im_g<12, true><<<nbg, 384, 0, c.st>>>(ca);
    else if (g4 && !no_g && G == 2) k_col2im_g<2><<<nbg, 64, 0, c.st>>>(ca);
    else if (g4 && !no_g && G == 3) k_col2im_g<3><<<nbg, 96, 0, c.st>>>(ca);
    else if (g4 && !no_g && G == 6) k_col2im_g<6><<<nbg, 192, 0, c.st>>>(ca);
    else if (g4 && !no_g && G == 12) k_col2im_g<12><<<nbg, 384, 0, c.st>>>(ca);
    else if (v4) k_col2im<4><<<nb, 128, 0, c.st>>>(ca);
    else if (v2) k_col2im<2><<<nb, 128, 0, c.st>>>(ca);
    else k_col2im<1><<<nb, 128, 0, c.st>>>(ca);
    INB_CUDA(cudaGetLastError());
  }
}

}  // namespace inb
