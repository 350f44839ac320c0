// conv_tc_wgrad.cu - weight gradients (\nabla conv_filter) on tcgen05 (see conv_tc.cuh).
//
//   D[tap][p][q] = sum_{pixels} P[pix][p] * Q[pix + off(tap)][q]
//
// Both operands are MN-major for the MMA: the contraction index K is the pixel (the strided dimension of
// the pixel-major planes), so a TMA box of (64 channels x 64 pixels) lands in shared memory as the
// canonical MN-major SWIZZLE_128B tile.  P supplies M = 128 channels (two 64-channel atoms); the taps of
// Q are separate TMA boxes (shifted, zero-filled at the image border) stored back to back, so that ONE
// MMA with N = taps*cq (<= 256) walks all of them through the leading-dimension byte offset: the
// accumulator of tap t sits at TMEM columns [t*cq, (t+1)*cq).
// CTAs split the pixel range; partial sums are added with red.global.add.f32 straight into the
// reference's weight layout dw[p][q][T-1-tap] (the kernel flip of NNlib's conv).
#include "conv_tc.cuh"
#include "tc_common.cuh"
#include "tc_maps.cuh"

#include <algorithm>

namespace inb {
using namespace tc;

struct WgradTcArgs {
  int ksz, T;
  int W, H, D;
  long long M;
  int nblocks, blocks_per_cta;  // 64-pixel k-blocks
  int cq, qa, nqa, cq_real;
  int stages;
  uint32_t tmem_cols;
  uint32_t p_bytes, q_tap_bytes;  // per plane: P tile (128 ch x 64 px), Q tile of one tap
  int taps_per_cta;               // taps handled by one CTA (blockIdx.z selects the group)
  int taps_per_mma;               // taps merged into one MMA (N = taps_per_mma * cq <= 256)
  float* dw;
};

__device__ __forceinline__ void wtap_offset(int tap, int ksz, int D, int& dx, int& dy, int& dz) {
  if (ksz == 1) { dx = dy = dz = 0; return; }
  dx = tap % 3 - 1;
  dy = (tap / 3) % 3 - 1;
  dz = (D > 1) ? tap / 9 - 1 : 0;
}

template <int NT>
__global__ void __launch_bounds__(192, 1)
k_wgrad_tc(const __grid_constant__ CUtensorMap mP0, const __grid_constant__ CUtensorMap mP1,
           const __grid_constant__ CUtensorMap mQ0, const __grid_constant__ CUtensorMap mQ1, const WgradTcArgs a) {
  constexpr int NP = (NT == 1) ? 1 : 2;
  constexpr int PB = 64;  // pixels per k-block
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tap_begin = blockIdx.z * a.taps_per_cta;
  const int ntap = min(a.taps_per_cta, a.T - tap_begin);
  const uint32_t stage_bytes = NP * (a.p_bytes + a.taps_per_cta * a.q_tap_bytes);
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
      const uint32_t q_atom_bytes = PB * a.qa * 2;  // one channel atom of one tap
      const uint32_t tx = NP * (a.p_bytes + ntap * a.nqa * q_atom_bytes);
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
            wtap_offset(tap_begin + tp, a.ksz, a.D, dx, dy, dz);
            uint8_t* dst = sq + (size_t)(pl * a.taps_per_cta + tp) * a.q_tap_bytes;
            for (int qi = 0; qi < a.nqa; ++qi)
              tma_load_5d(mq, full + s, dst + (size_t)qi * q_atom_bytes, qi * a.qa, x0 + dx, y0 + dy, z0 + dz, b0);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t q_row = a.qa * 2;
      const uint32_t q_layout = layout_for_row(q_row);
      const uint32_t q_atom_bytes = PB * q_row;
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % a.stages;
        const uint32_t ph = (kb / a.stages) & 1;
        mbar_wait(full + s, ph);
        tc_fence_after();
        const uint32_t sp = smem_u32(base + (size_t)s * stage_bytes);
        const uint32_t sq = sp + NP * a.p_bytes;
        // groups of taps merged into one MMA: N = g*cq columns, tap atoms are back to back in smem
        for (int t0 = 0; t0 < ntap; t0 += a.taps_per_mma) {
          const int g = min(a.taps_per_mma, ntap - t0);
          const uint32_t idesc = make_idesc_bf16(128, g * a.cq, 1, 1);
#pragma unroll
          for (int term = 0; term < NT; ++term) {
            const uint32_t tp_ = sp + ((term == 2) ? a.p_bytes : 0);
            const uint32_t tq_ = sq + (uint32_t)(((term == 1) ? a.taps_per_cta : 0) + t0) * a.q_tap_bytes;
#pragma unroll
            for (int k = 0; k < PB / 16; ++k) {
              // MN-major: 16 pixels (K) = 16 rows; LBO = stride between channel atoms, SBO = 8 rows
              const uint64_t ad = make_smem_desc(tp_ + k * 16 * 128, PB * 128, 8 * 128, LAYOUT_SW128);
              const uint64_t bd = make_smem_desc(tq_ + k * 16 * q_row, q_atom_bytes, 8 * q_row, q_layout);
              umma_f16(tmem + t0 * a.cq, ad, bd, idesc, (kb > 0 || term > 0 || k > 0) ? 1u : 0u);
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

static uint32_t pow2_cols(int n) {
  uint32_t c = 32;
  while ((int)c < n) c <<= 1;
  return c;
}
static int pick_atom(int cq) { return (cq % 64 == 0) ? 64 : ((cq % 32 == 0) ? 32 : 16); }

void op_wgrad_tc(Ctx& c, const WgradTcSpec& s) {
  INB_CHECK(s.k == 1 || s.k == 3, "ResidualBlock kernel size %d is not supported (1 or 3)", s.k);
  INB_CHECK(s.np % 128 == 0, "tensor-core wgrad needs n_hidden to be a multiple of 128 (got %d)", s.np);
  INB_CHECK(s.cq % 16 == 0 && s.cq <= 256, "tensor-core wgrad: bad channel count %d", s.cq);
  const TileBox tb = make_tile_box(s.g, s.B, 64);
  INB_CHECK(tb.ok, "spatial size %dx%dx%d cannot be tiled for the tensor-core path; use precision fp32", s.g.W,
            s.g.H, s.g.D);
  if (c.dry()) return;
  const int NT = prec_terms(c.prec);
  const int NP = NT == 1 ? 1 : 2;
  const int T = s.k == 1 ? 1 : (s.g.nd == 3 ? 27 : 9);
  WgradTcArgs a{};
  a.ksz = s.k;
  a.T = T;
  a.W = s.g.W; a.H = s.g.H; a.D = s.g.D;
  a.M = s.g.px * s.B;
  a.nblocks = (int)cdiv(a.M, 64);
  a.cq = s.cq;
  a.qa = pick_atom(s.cq);
  a.nqa = s.cq / a.qa;
  a.cq_real = s.cq_real;
  a.p_bytes = 128 * 64 * 2;
  a.q_tap_bytes = 64 * s.cq * 2;
  // taps per CTA: TMEM (512 columns) and a >= 2-stage ring within shared memory
  int tpc = std::min(T, 512 / s.cq);
  while (tpc > 1 && 2 * NP * (a.p_bytes + tpc * a.q_tap_bytes) > 200 * 1024) --tpc;
  a.taps_per_cta = tpc;
  a.taps_per_mma = std::max(1, std::min(tpc, 256 / s.cq));
  const int groups = (int)cdiv(T, tpc);
  const uint32_t stage_bytes = NP * (a.p_bytes + tpc * a.q_tap_bytes);
  int stages = (int)((220 * 1024) / stage_bytes);
  if (stages > 6) stages = 6;
  INB_CHECK(stages >= 1, "tensor-core wgrad: stage of %u bytes does not fit", stage_bytes);
  a.stages = stages;
  a.tmem_cols = pow2_cols(tpc * s.cq);
  a.dw = s.dw;
  const int halves = s.np / 128;
  long long want = std::max<long long>(1, 148LL / ((long long)halves * groups));
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
