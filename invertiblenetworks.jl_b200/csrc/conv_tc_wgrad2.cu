// conv_tc_wgrad2.cu - weight (and bias) gradients of the fused ResidualBlock path on tcgen05.
//
//   D[p][q] = sum_{pixels} P[pix][p] * Q[pix][q]          (\nabla conv_filter, layer_residual_block.jl:152,156,163)
//
// P = a 128/256-channel hidden gradient or activation [M][np], Q = im2col rows [M][taps*C -> 64k] (3x3 convs: the
// taps are columns, so there is no tap loop and every TMA row is 128 bytes) or the other hidden tensor (1x1 conv).
// Both operands are MN-major for the MMA (the contraction index is the pixel = the strided dimension): a TMA box of
// (64 channels x 32 pixels) lands in shared memory as the canonical MN-major SWIZZLE_128B atom.
// One persistent CTA per SM owns a contiguous pixel range and ALL np rows (np/128 accumulators of N <= 256 columns
// in TMEM, so P and Q are read from HBM exactly once).  Each CTA stores its partial D tile (TMA store through a
// swizzled staging tile) and k_wgrad_reduce sums the partials in a fixed order straight into the reference's weight
// layout dw[p][c][T-1-tap] (the kernel flip of NNlib's conv): deterministic, and cheaper than np*nq atomics per CTA.
// The four epilogue warps are idle during the main loop: they sum the columns of the P tile while it sits in shared
// memory, which yields the bias gradient db[p] = sum_pix P[pix][p] (:157,164) without another pass over HBM.
//
// Warps: 0 TMA producer | 1 MMA issuer | 2..5 expansion of 1-byte lo planes during the loop, TMEM -> dw afterwards |
// 6..9 bias sums during the loop.
#include "conv_tc.cuh"
#include "tc_common.cuh"
#include "tc_maps.cuh"

#include <algorithm>
#include <cstdlib>

namespace inb {
using namespace tc;

constexpr int kWgPB = 32;  // pixels per k-block

struct Wgrad2Args {
  long long M;
  int nblocks, blocks_per_cta;
  int np, halves;   // rows of D (channels of P), np / 128
  int nq, ng;       // columns handled by this launch's groups: group g = columns [g*256, ...) of width <= 256
  int qtot;         // total columns of Q
  int C, T;         // column q = tap*C + c for q < T*C, other columns are dropped
  int stages;
  uint32_t stage_bytes, p_plane, q_plane;  // per plane: P tile, Q tile (of the widest group)
  float* dbpart;    // [gridDim.x][np] partial bias sums, nullable
  int f16;          // operand planes are IEEE halves (INB_PREC_FP16X3)
  int p8, q8;       // the lo plane of P / Q holds one byte per value (Planes::lo8): expanded to halves in shared memory
};

constexpr int kWg2Threads = 320;
template <int NT>
__global__ void __launch_bounds__(kWg2Threads, 1)
k_wgrad2_tc(const __grid_constant__ CUtensorMap mP0, const __grid_constant__ CUtensorMap mP1,
            const __grid_constant__ CUtensorMap mQ0, const __grid_constant__ CUtensorMap mQ1,
            const __grid_constant__ CUtensorMap mD, const Wgrad2Args a) {
  constexpr int NP = (NT == 1) ? 1 : 2;
  constexpr uint32_t ATOM = kWgPB * 128;  // 64 channels x 32 pixels
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)a.stages * a.stage_bytes);
  uint64_t* empty = full + 8;
  uint64_t* tfull = empty + 8;
  uint64_t* conv = tfull + 1;   // [8] the 1-byte lo planes of a stage have been expanded (p8 / q8)
  uint32_t* tslot = reinterpret_cast<uint32_t*>(conv + 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = blockIdx.y;
  const int q0 = grp * 256;
  const int nqg = min(256, a.qtot - q0);  // columns of this group (multiple of 64)
  const int qatoms = nqg / 64, patoms = a.np / 64;
  const bool do_bias = a.dbpart != nullptr && grp == 0;

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    prefetch_tmap(&mP0);
    prefetch_tmap(&mQ0);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < a.stages; ++s) {
        mbar_init(full + s, 1);
        mbar_init(empty + s, do_bias ? 1 + patoms : 1);
        mbar_init(conv + s, 4);
      }
      mbar_init(tfull, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tslot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tslot;
  const int blk0 = blockIdx.x * a.blocks_per_cta;
  const int nkb = max(min(blk0 + a.blocks_per_cta, a.nblocks) - blk0, 0);

  if (warp == 0) {
    if (elect_one()) {
      // a 1-byte lo plane arrives as ONE dense [32 pixels][np | nqg bytes] box at the start of the operand's lo region
      const uint32_t tx = patoms * ATOM * (NP == 2 ? (a.p8 ? 3 : 4) : 2) / 2 + qatoms * ATOM * (NP == 2 ? (a.q8 ? 3 : 4) : 2) / 2;
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % a.stages;
        mbar_wait(empty + s, ((kb / a.stages) & 1) ^ 1);
        mbar_expect_tx(full + s, tx);
        const int row0 = (blk0 + kb) * kWgPB;
        uint8_t* sp = smem + (size_t)s * a.stage_bytes;
        uint8_t* sq = sp + NP * a.p_plane;
#pragma unroll
        for (int pl = 0; pl < NP; ++pl) {
          if (pl && a.p8) tma_load_2d(&mP1, full + s, sp + a.p_plane, 0, row0);
          else
            for (int at = 0; at < patoms; ++at)
              tma_load_2d(pl ? &mP1 : &mP0, full + s, sp + pl * a.p_plane + at * ATOM, at * 64, row0);
          if (pl && a.q8) tma_load_2d(&mQ1, full + s, sq + a.q_plane, q0, row0);
          else
            for (int at = 0; at < qatoms; ++at)
              tma_load_2d(pl ? &mQ1 : &mQ0, full + s, sq + pl * a.q_plane + at * ATOM, q0 + at * 64, row0);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc_16(128, nqg, 1, 1, a.f16 != 0);
      // MN-major SWIZZLE_128B: LBO = byte distance between 64-element atoms along M / N, SBO = 8 pixel rows
      const uint32_t dhi = (uint32_t)(make_smem_desc(0, ATOM, 1024, LAYOUT_SW128) >> 32);
      const uint32_t dlo_lbo = ((ATOM >> 4) & 0x3FFF) << 16;
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % a.stages;
        mbar_wait(full + s, (kb / a.stages) & 1);
        if (a.p8 | a.q8) mbar_wait(conv + s, (kb / a.stages) & 1);
        tc_fence_after();
        const uint32_t sp = smem_u32(smem + (size_t)s * a.stage_bytes);
        const uint32_t sq = sp + NP * a.p_plane;
        for (int h = 0; h < a.halves; ++h) {
#pragma unroll
          for (int term = 0; term < NT; ++term) {
            const uint32_t tp = sp + ((term == 2) ? a.p_plane : 0) + h * 2 * ATOM;
            const uint32_t tq = sq + ((term == 1) ? a.q_plane : 0);
#pragma unroll
            for (int k = 0; k < kWgPB / 16; ++k) {
              const uint64_t ad = ((uint64_t)dhi << 32) | dlo_lbo | (((tp + k * 16 * 128) >> 4) & 0x3FFF);
              const uint64_t bd = ((uint64_t)dhi << 32) | dlo_lbo | (((tq + k * 16 * 128) >> 4) & 0x3FFF);
              umma_f16(tmem + h * 256, ad, bd, idesc, (kb > 0 || term > 0 || k > 0) ? 1u : 0u);
            }
          }
        }
        umma_commit(empty + s);
      }
      umma_commit(tfull);
    }
  } else if (warp >= 6) {
    // bias warps: column sums of the P tile (atom bw = channels [64 bw, 64 bw + 64)): lane owns channels 2*lane, 2*lane+1
    const int bw = warp - 6;
    if (do_bias && bw < patoms) {
      float s0 = 0.f, s1 = 0.f;
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % a.stages;
        mbar_wait(full + s, (kb / a.stages) & 1);
        if (a.p8) mbar_wait(conv + s, (kb / a.stages) & 1);  // the lo halves of P exist once they have been expanded
        const uint8_t* at = smem + (size_t)s * a.stage_bytes + bw * ATOM;
#pragma unroll 8
        for (int px = 0; px < kWgPB; ++px) {
          const uint32_t off = px * 128 + ((((uint32_t)lane >> 2) ^ (px & 7)) << 4) + (lane & 3) * 4;
          const uint32_t h = *reinterpret_cast<const uint32_t*>(at + off);
          s0 += half_word_lo_to_f(a.f16 != 0, h);
          s1 += half_word_hi_to_f(a.f16 != 0, h);
          if (NT == 3) {
            const uint32_t l = *reinterpret_cast<const uint32_t*>(at + a.p_plane + off);
            s0 += half_word_lo_to_f(a.f16 != 0, l);
            s1 += half_word_hi_to_f(a.f16 != 0, l);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + s);
      }
      *reinterpret_cast<float2*>(a.dbpart + (long long)blockIdx.x * a.np + bw * 64 + 2 * lane) = make_float2(s0, s1);
    }
    asm volatile("bar.sync 3, 256;" ::: "memory");  // the epilogue warps may now reuse the stages as their staging area
  } else {
    const int ew = warp - 2;  // 0..3
    // 1-byte lo planes: ONE dense box [32 pixels][np (nqg) bytes] per operand lands at the start of the operand's lo
    // region (the TMA unit is request-rate bound: a 256-byte row costs what a 64-byte one does).  Warp ew expands the
    // 64 channels of atom ew (byte b of a value -> half b << 8): every lane reads its eight 8-byte pieces, the four
    // warps synchronise (the box overlays the atoms of other warps), then each writes the eight 16-byte chunks of its
    // MN-major SWIZZLE_128B atom, and the MMA warp is told through conv[s].
    const bool conv_on = (a.p8 | a.q8) != 0;
    const int j8 = lane & 7, rb = lane >> 3;
    if (conv_on) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % a.stages;
        mbar_wait(full + s, (kb / a.stages) & 1);
        uint8_t* sp = smem + (size_t)s * a.stage_bytes;
        uint8_t* pl8 = sp + a.p_plane;                     // [32][np] bytes
        uint8_t* ql8 = sp + NP * a.p_plane + a.q_plane;    // [32][nqg] bytes
        const bool dp = a.p8 && ew < patoms, dq = a.q8 && ew < qatoms;
        uint2 srcp[8], srcq[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (dp) srcp[i] = *reinterpret_cast<const uint2*>(pl8 + (rb + 4 * i) * a.np + ew * 64 + j8 * 8);
          if (dq) srcq[i] = *reinterpret_cast<const uint2*>(ql8 + (rb + 4 * i) * nqg + ew * 64 + j8 * 8);
        }
        asm volatile("bar.sync 2, 128;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = rb + 4 * i;
          const uint32_t off = (uint32_t)(r * 128 + ((j8 ^ (r & 7)) << 4));
          if (dp)
            *reinterpret_cast<uint4*>(pl8 + ew * ATOM + off) =
                make_uint4(__byte_perm(srcp[i].x, 0u, 0x1404), __byte_perm(srcp[i].x, 0u, 0x3424),
                           __byte_perm(srcp[i].y, 0u, 0x1404), __byte_perm(srcp[i].y, 0u, 0x3424));
          if (dq)
            *reinterpret_cast<uint4*>(ql8 + ew * ATOM + off) =
                make_uint4(__byte_perm(srcq[i].x, 0u, 0x1404), __byte_perm(srcq[i].x, 0u, 0x3424),
                           __byte_perm(srcq[i].y, 0u, 0x1404), __byte_perm(srcq[i].y, 0u, 0x3424));
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(conv + s);
      }
    }
    {
      // partial tile -> scratch: the pipeline stages are free now and serve as the staging area of the TMA store
      // (SWIZZLE_128B boxes of 32 fp32 columns x 128 rows, as for P in the fused chain)
      mbar_wait(tfull, 0);
      tc_fence_after();
      asm volatile("bar.sync 3, 256;" ::: "memory");  // every bias warp has finished reading the last stages
      const int q = warp & 3;  // TMEM lane quadrant
      const int row = q * 32 + lane;
      const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
      const int tid = ew * 32 + lane;
      for (int h = 0; h < a.halves; ++h) {
        if (h > 0) {
          if (tid == 0) bulk_wait_read0();
          asm volatile("bar.sync 1, 128;" ::: "memory");
        }
        for (int c0 = 0; c0 < nqg; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(tmem + h * 256 + c0 + lane_sel, r);
          tmem_ld_wait();
          uint8_t* g = smem + (size_t)(c0 >> 5) * 16384 + row * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<uint4*>(g + ((j ^ (row & 7)) << 4)) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
        }
        fence_proxy_async();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (tid == 0) {
          const int drow = ((grp * (int)gridDim.x + (int)blockIdx.x) * a.np) + h * 128;
          for (int c0 = 0; c0 < nqg; c0 += 32) tma_store_2d(&mD, smem + (size_t)(c0 >> 5) * 16384, c0, drow);
          bulk_commit();
        }
      }
      if (tid == 0) bulk_wait0();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// dw[p][c][T-1-tap] = sum_cta part[g][cta][p][col], db[p] = sum_cta dbpart[cta][p] in a fixed order: four lanes share an
// output (each walks every fourth partial with two interleaved sums, then two shuffles) - the kernel is a latency-bound
// column walk, so the parallelism is what counts; consecutive outputs stay on consecutive lane groups (coalesced)
struct WgradReduceDesc {
  const float* part;
  int ncta, np, pitch, C, T;
  int np_real;  // rows p >= np_real are zero padding of the hidden width: not part of dw / db
  float* dw;
  const float* dbpart;
  float* db;
  int ones_col;  // >= 0: db[p] = column ones_col of the summed gradient tile (group 0) instead of the dbpart sums
  const uint32_t* smax;  // INB_PREC_FP16X3: the sums carry the gradient scale derived from *smax
};
struct WgradReduceArgs {
  WgradReduceDesc d[3];
};
__global__ void __launch_bounds__(256) k_wgrad_reduce(const WgradReduceArgs a) {
  const WgradReduceDesc& q = a.d[blockIdx.y];  // one weight gradient per grid row (the three of a block in one launch)
  const float* __restrict__ part = q.part;
  const float* __restrict__ dbpart = q.dbpart;
  const int ncta = q.ncta, np = q.np, pitch = q.pitch, C = q.C, T = q.T;
  const int ncol = T * C;
  const int npr = q.np_real;
  const bool has_db = dbpart != nullptr || q.ones_col >= 0;
  const long long total = (long long)npr * ncol, outs = total + (has_db ? npr : 0);
  float osc = 1.f;
  if (q.smax) { float sc; f16_scale_from_max(__ldg(q.smax), sc, osc); }
  const int sub = threadIdx.x >> 6;  // which quarter of the partials (a warp works on 32 consecutive outputs)
  const int ol = threadIdx.x & 63;
  __shared__ float sm[4][64];
  for (long long i0 = (long long)blockIdx.x * 64; i0 < outs; i0 += (long long)gridDim.x * 64) {
    const long long i = i0 + ol;
    float s0 = 0.f, s1 = 0.f;
    if (i < outs) {
      const float* src;
      long long step;
      if (i >= total && q.ones_col >= 0) {
        src = part + (i - total) * pitch + q.ones_col;
        step = (long long)np * pitch;
      } else if (i >= total) {
        src = dbpart + (i - total);
        step = np;
      } else {
        const int col = (int)(i % ncol), p = (int)(i / ncol);
        const int g = col >> 8, cg = col & 255;
        src = part + ((long long)g * ncta * np + p) * pitch + cg;
        step = (long long)np * pitch;
      }
      int k = sub;
      for (; k + 4 < ncta; k += 8) {
        s0 += __ldg(src + k * step);
        s1 += __ldg(src + (k + 4) * step);
      }
      if (k < ncta) s0 += __ldg(src + k * step);
    }
    sm[sub][ol] = s0 + s1;
    __syncthreads();
    if (sub == 0 && i < outs) {
      const float s = ((sm[0][ol] + sm[1][ol]) + (sm[2][ol] + sm[3][ol])) * osc;
      if (i >= total) {
        q.db[i - total] = s;
      } else {
        const int col = (int)(i % ncol), p = (int)(i / ncol);
        const int tap = col / C, cc = col - tap * C;
        q.dw[((long long)p * C + cc) * T + (T - 1 - tap)] = s;
      }
    }
    __syncthreads();
  }
}

// Large batches: the weight gradients are HBM-bound (they read every stored hidden plane once) while the chain passes
// around them are tensor-bound, so instead of taking turns on all SMs the three split-K kernels of a step run on
// `wgrad_overlap_ctas()` CTAs each, on the side streams, UNDER the chain passes of the next flow step, which leave
// those SMs free (op_rb_chain caps its resident pairs accordingly).  0 = off.  INB_WGRAD_OVERLAP=<ctas per kernel>.
int wgrad_overlap_ctas() {
  static const int n = [] {
    const char* e = getenv("INB_WGRAD_OVERLAP");
    const int v = e ? atoi(e) : 0;
    return v < 0 ? 0 : (v > 48 ? 48 : v);
  }();
  return n;
}

// launches the split-K kernel of one weight gradient; its partial tiles stay allocated (the caller releases the arena
// scope after the reduction) and are described in `rd`; returns the number of outputs
static long long wgrad2_launch(Ctx& c, const Wgrad2TcSpec& s, WgradReduceDesc& rd) {
  INB_CHECK(s.np == 128 || s.np == 256, "tensor-core wgrad: np = %d must be 128 or 256", s.np);
  INB_CHECK(s.Q.pitch % 64 == 0 && s.P.pitch == s.np, "tensor-core wgrad: operand pitches must be multiples of 64");
  INB_CHECK(s.T * s.C <= s.Q.pitch, "tensor-core wgrad: Q has %d columns, need %d", s.Q.pitch, s.T * s.C);
  const int NT = prec_terms(c.prec);
  const int NP = NT == 1 ? 1 : 2;
  Wgrad2Args a{};
  a.f16 = prec_f16(c.prec) ? 1 : 0;
  a.p8 = s.P.lo8;
  a.q8 = s.Q.lo8;
  INB_CHECK(!(a.p8 | a.q8) || (NT == 3 && a.f16), "tensor-core wgrad: 1-byte lo planes exist in fp16x3 only");
  a.M = s.M;
  a.nblocks = (int)cdiv(s.M, kWgPB);
  a.np = s.np;
  a.halves = s.np / 128;
  a.qtot = s.Q.pitch;
  a.C = s.C;
  a.T = s.T;
  const int ng = (int)cdiv(a.qtot, 256);
  const int nqmax = std::min(a.qtot, 256);
  a.p_plane = (uint32_t)s.np * kWgPB * 2;
  a.q_plane = (uint32_t)nqmax * kWgPB * 2;
  a.stage_bytes = NP * (a.p_plane + a.q_plane);
  int stages = (int)((225 * 1024) / a.stage_bytes);
  if (stages > 8) stages = 8;
  INB_CHECK(stages >= 2, "tensor-core wgrad: stage of %u bytes does not fit", a.stage_bytes);
  INB_CHECK((size_t)stages * a.stage_bytes >= (size_t)(nqmax / 32) * 16384, "tensor-core wgrad: staging does not fit");
  a.stages = stages;
  // split-K over the SMs, but at least 16 k-blocks (512 pixels) per CTA: every CTA costs a partial tile that the
  // reduction has to read back, which dominates when the batch shard is small
  int ctas = std::max(1, std::min(148 / ng, a.nblocks / 16));
  if (c.wg_defer && wgrad_overlap_ctas() > 0) ctas = std::max(1, std::min(ctas, wgrad_overlap_ctas() / ng));
  a.blocks_per_cta = (int)cdiv(a.nblocks, ctas);
  const unsigned gx = (unsigned)cdiv(a.nblocks, a.blocks_per_cta);  // every CTA owns at least one block
  // scratch: partial tiles [ng][gx][np][nqmax] and partial bias sums [gx][np]
  float* part = c.ar->f32((size_t)ng * gx * s.np * nqmax);
  float* dbpart = c.ar->f32((size_t)gx * s.np);  // sized in the dry run too (gradient pointers are fake there)
  const int ones_col = (s.db && s.ones_col >= 0 && s.ones_col < std::min(a.qtot, 256) && s.ones_col >= s.T * s.C) ? s.ones_col : -1;
  a.dbpart = (s.db && ones_col < 0) ? dbpart : nullptr;
  if (c.dry()) return 0;
  const size_t smem = (size_t)stages * a.stage_bytes + 25 * 8 + 16;
  CUtensorMap mP0 = make_rows_map(s.P.hi, s.P.pitch, s.M, 64, kWgPB);
  CUtensorMap mP1 = a.p8 ? make_rows_map_u8(s.P.lo, s.P.pitch, s.M, s.np, kWgPB, false) : make_rows_map(s.P.lo, s.P.pitch, s.M, 64, kWgPB);
  CUtensorMap mQ0 = make_rows_map(s.Q.hi, s.Q.pitch, s.M, 64, kWgPB);
  INB_CHECK(!a.q8 || a.qtot <= 256, "tensor-core wgrad: a 1-byte Q plane wider than 256 columns is not supported");
  CUtensorMap mQ1 = a.q8 ? make_rows_map_u8(s.Q.lo, s.Q.pitch, s.M, nqmax, kWgPB, false) : make_rows_map(s.Q.lo, s.Q.pitch, s.M, 64, kWgPB);
  CUtensorMap mD = make_rows_map_f32(part, nqmax, (long long)ng * gx * s.np, 32, 128);
  Prof pf(c, F_WGRAD_TC, 1, 2.0 * s.M * a.qtot * s.np * NT,
          (double)s.M * (s.np * (NP == 2 ? (a.p8 ? 3.0 : 4.0) : 2.0) + a.qtot * (NP == 2 ? (a.q8 ? 3.0 : 4.0) : 2.0)));
  dim3 grid(gx, ng, 1);
  if (NT == 3) {
    INB_CUDA(cudaFuncSetAttribute(k_wgrad2_tc<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_wgrad2_tc<3><<<grid, kWg2Threads, smem, c.st>>>(mP0, mP1, mQ0, mQ1, mD, a);
  } else {
    INB_CUDA(cudaFuncSetAttribute(k_wgrad2_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_wgrad2_tc<1><<<grid, kWg2Threads, smem, c.st>>>(mP0, mP1, mQ0, mQ1, mD, a);
  }
  INB_CUDA(cudaGetLastError());
  const int npr = s.np_real > 0 ? s.np_real : s.np;
  rd = WgradReduceDesc{part, (int)gx, s.np, nqmax, s.C, s.T, npr, s.dw, a.dbpart, s.db, ones_col, prec_f16(c.prec) ? s.smax : nullptr};
  return (long long)npr * s.T * s.C + (s.db ? npr : 0);
}

// n (<= 3) weight gradients: their split-K kernels back to back, then ONE reduction launch (grid row = gradient)
void op_wgrad2_tc_multi(Ctx& c, const Wgrad2TcSpec* specs, int n) {
  INB_CHECK(n >= 1 && n <= 3, "wgrad: 1 to 3 gradients per call");
  size_t mk = c.ar->mark();
  WgradReduceArgs ra{};
  long long outs = 0;
  // Small pixel counts (a batch shard of 8, the coarse scales): a split-K kernel then occupies at most half of the SMs
  // (>= 512 pixels per CTA) and is latency-bound, so the three of a block run side by side on three streams - parallel
  // branches of the captured graph - and the reduction follows their join.
  SideLane* L = c.lane;
  const bool side_by_side = !c.dry() && n == 3 && L && L->wst[0] && L->wnext + 3 <= L->nwev &&
                            (specs[0].M / (kWgPB * 16) <= 74 || (c.wg_defer && wgrad_overlap_ctas() > 0)) &&
                            specs[0].M == specs[1].M && specs[0].M == specs[2].M;
  if (side_by_side && c.wg_defer && L->wst[2] && L->wnext + 4 <= L->nwev) {
    // all three off the main stream; nothing on the main stream waits for them before the region is reused
    cudaEvent_t ef = L->wev[L->wnext], e1 = L->wev[L->wnext + 1], e2 = L->wev[L->wnext + 2], ed = L->wev[L->wnext + 3];
    L->wnext += 4;
    INB_CUDA(cudaEventRecord(ef, c.st));
    Ctx cc[3] = {c, c, c};
    for (int i = 0; i < 3; ++i) {
      cc[i].st = L->wst[i];
      INB_CUDA(cudaStreamWaitEvent(L->wst[i], ef, 0));
      outs = std::max(outs, wgrad2_launch(cc[i], specs[i], ra.d[i]));
    }
    INB_CUDA(cudaEventRecord(e1, L->wst[1]));
    INB_CUDA(cudaEventRecord(e2, L->wst[2]));
    INB_CUDA(cudaStreamWaitEvent(L->wst[0], e1, 0));
    INB_CUDA(cudaStreamWaitEvent(L->wst[0], e2, 0));
    {
      Prof pf(cc[0], F_WGRAD_TC, 1, 0, 0);
      dim3 grid((unsigned)std::min<long long>(cdiv(outs, 64), 148 * 4), (unsigned)n, 1);
      k_wgrad_reduce<<<grid, 256, 0, L->wst[0]>>>(ra);
      INB_CUDA(cudaGetLastError());
    }
    INB_CUDA(cudaEventRecord(ed, L->wst[0]));
    L->wdone[L->wparity] = ed;
    L->wpending[L->wparity] = true;
    c.ar->release(mk);  // the partial tiles stay untouched: nothing else is allocated in this step's region afterwards
    return;
  }
  if (side_by_side) {
    cudaEvent_t ef = L->wev[L->wnext], e1 = L->wev[L->wnext + 1], e2 = L->wev[L->wnext + 2];
    L->wnext += 3;
    INB_CUDA(cudaEventRecord(ef, c.st));
    INB_CUDA(cudaStreamWaitEvent(L->wst[0], ef, 0));
    INB_CUDA(cudaStreamWaitEvent(L->wst[1], ef, 0));
    Ctx c1 = c, c2 = c;
    c1.st = L->wst[0];
    c2.st = L->wst[1];
    outs = std::max(outs, wgrad2_launch(c, specs[0], ra.d[0]));
    outs = std::max(outs, wgrad2_launch(c1, specs[1], ra.d[1]));
    outs = std::max(outs, wgrad2_launch(c2, specs[2], ra.d[2]));
    INB_CUDA(cudaEventRecord(e1, L->wst[0]));
    INB_CUDA(cudaEventRecord(e2, L->wst[1]));
    INB_CUDA(cudaStreamWaitEvent(c.st, e1, 0));
    INB_CUDA(cudaStreamWaitEvent(c.st, e2, 0));
  } else {
    for (int i = 0; i < n; ++i) outs = std::max(outs, wgrad2_launch(c, specs[i], ra.d[i]));
  }
  if (!c.dry()) {
    Prof pf(c, F_WGRAD_TC, 1, 0, 0);
    dim3 grid((unsigned)std::min<long long>(cdiv(outs, 64), 148 * 4), (unsigned)n, 1);
    k_wgrad_reduce<<<grid, 256, 0, c.st>>>(ra);
    INB_CUDA(cudaGetLastError());
  }
  c.ar->release(mk);
}
void op_wgrad2_tc(Ctx& c, const Wgrad2TcSpec& s) { op_wgrad2_tc_multi(c, &s, 1); }

}  // namespace inb
