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
#include "tc_maps.cuh"

#include <cuda.h>
#include <algorithm>
#include <cstdlib>
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
CUtensorMap make_act_map(const __nv_bfloat16* base, int pitch, const Geo& g, int B, int ck, const TileBox& tb) {
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
CUtensorMap make_w_map(const __nv_bfloat16* base, int ktot, int rows, int ck) {
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

CUtensorMap make_rows_map(const __nv_bfloat16* base, int pitch, long long rows, int box_c, int box_rows) {
  CUtensorMap m;
  cuuint64_t dims[2] = {(cuuint64_t)pitch, (cuuint64_t)rows};
  cuuint64_t str[1] = {(cuuint64_t)pitch * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_c, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)base, dims, str, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(box_c * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) fail(2, "cuTensorMapEncodeTiled(rows) failed with %d", (int)r);
  return m;
}

// fp32 [rows][pitch] tensor, box (box_c <= 32 columns = 128 bytes, box_rows), SWIZZLE_128B (P of the fused chain)
CUtensorMap make_rows_map_f32(const float* base, int pitch, long long rows, int box_c, int box_rows) {
  CUtensorMap m;
  cuuint64_t dims[2] = {(cuuint64_t)pitch, (cuuint64_t)rows};
  cuuint64_t str[1] = {(cuuint64_t)pitch * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_c, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, str, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) fail(2, "cuTensorMapEncodeTiled(fp32 rows) failed with %d", (int)r);
  return m;
}

// fp32 [rows][pitch] tensor, box (box_c columns with box_c * 4 a multiple of 16, box_rows), no swizzle: the shared-memory
// side is a dense [box_rows][box_c] array (Pq of the fused chain)
CUtensorMap make_rows_map_f32_dense(const float* base, int pitch, long long rows, int box_c, int box_rows) {
  CUtensorMap m;
  cuuint64_t dims[2] = {(cuuint64_t)pitch, (cuuint64_t)rows};
  cuuint64_t str[1] = {(cuuint64_t)pitch * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_c, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, str, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) fail(2, "cuTensorMapEncodeTiled(dense fp32 rows) failed with %d", (int)r);
  return m;
}

CUtensorMap make_rows_map_u8(const void* base, int pitch, long long rows, int box_c, int box_rows, bool swizzle64) {
  CUtensorMap m;
  cuuint64_t dims[2] = {(cuuint64_t)pitch, (cuuint64_t)rows};
  cuuint64_t str[1] = {(cuuint64_t)pitch};
  cuuint32_t box[2] = {(cuuint32_t)box_c, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, str, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) fail(2, "cuTensorMapEncodeTiled(byte rows) failed with %d", (int)r);
  return m;
}

// fp32 planes [nplanes][rows][cols], box (cols, box_rows, 1), no swizzle: the tap-row planes of the fused chain's Pq (a
// ragged last tile is clipped at `rows`, so it cannot spill into the next plane)
CUtensorMap make_planes_map_f32_dense(const float* base, int cols, long long rows, int nplanes, int box_rows) {
  CUtensorMap m;
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)nplanes};
  cuuint64_t str[2] = {(cuuint64_t)cols * 4, (cuuint64_t)cols * 4 * (cuuint64_t)rows};
  cuuint32_t box[3] = {(cuuint32_t)cols, (cuuint32_t)box_rows, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, dims, str, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) fail(2, "cuTensorMapEncodeTiled(dense fp32 planes) failed with %d", (int)r);
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

// im2col of a (B,C,px) fp32 tensor [two sources, conditional cat] into pixel-major bf16 hi/lo rows
//   col[m][tap*C + c] = x[c][pix(m) + off(tap)]   (zero outside the image = the conv's zero padding),
// K padded with zeros to kp (multiple of 64) except column `ones_col` (>= 0), which is 1.0.
// warp = 32 consecutive pixels (lane = pixel) x one 64-column block at a time.  Gather phase: the 16 independent loads
// of a (pixel, 16-column group) are coalesced across the warp (32 consecutive pixels of one channel plane; the 3x3
// neighbours re-read L1); hi / lo words go to a [32 rows][128 B + 16] tile per plane in shared memory.  Store phase:
// the tile leaves as full 128-byte rows (4 rows per warp instruction).  Storing straight from the gather layout costs
// one L1 transaction per lane (32 per instruction) and made the kernel transaction-bound at a third of HBM speed.
// The per-column (tap, channel) decode is a table built once per block; the per-pixel tap validity is a bit mask.
constexpr int kI2cThreads = 128;
constexpr int kI2cPix = 128;
constexpr int kI2cMaxK = 1024;
constexpr int kI2cRow = 144;  // bytes per tile row: 16-byte chunks of consecutive rows fall into different banks
__global__ void __launch_bounds__(kI2cThreads)
k_im2col_tc(const float* __restrict__ in0, long long in0_bs, int c0, const float* __restrict__ in1,
            long long in1_bs, int C, int T, int ksz, int W, int H, int D, long long px, long long M, int kp,
            int ones_col, int f16, const uint32_t* __restrict__ smax, __nv_bfloat16* __restrict__ hi,
            __nv_bfloat16* __restrict__ lo) {
  float insc = 1.f;  // INB_PREC_FP16X3: a gradient operand is scaled by the power of two derived from its max|.|
  if (smax) { float inv; f16_scale_from_max(__ldg(smax), insc, inv); }
  // per column: element offset relative to the pixel inside its source tensor, and (tap, which source)
  __shared__ int s_off[kI2cMaxK];
  __shared__ unsigned char s_tap[kI2cMaxK];  // tap index | 0x80 for the second source | 0xFF: padding column
  __shared__ __align__(16) unsigned char s_tile[kI2cThreads / 32][2][32 * kI2cRow];
  for (int k = threadIdx.x; k < kp; k += blockDim.x) {
    const int tap = k / C, ch = k - tap * C;
    if (tap < T) {
      int dx = 0, dy = 0, dz = 0;
      if (ksz != 1) { dx = tap % 3 - 1; dy = (tap / 3) % 3 - 1; dz = (D > 1) ? tap / 9 - 1 : 0; }
      const bool second = ch >= c0;
      s_off[k] = dx + dy * W + dz * W * H + (second ? ch - c0 : ch) * (int)px;
      s_tap[k] = (unsigned char)(tap | (second ? 0x80 : 0));
    } else {
      s_off[k] = 0;
      s_tap[k] = 0xFF;
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const long long m0 = (long long)blockIdx.x * kI2cPix + wid * 32;  // first pixel of this warp
  if (m0 >= M) return;
  const long long m = m0 + lane;
  const bool live = m < M;
  const long long mc = live ? m : M - 1;
  const long long b = mc / px, pix = mc - b * px;
  long long t = pix;
  const int x = (int)(t % W); t /= W;
  const int y = (int)(t % H); t /= H;
  const int z = (int)t;
  uint32_t okmask = 0;  // bit tap: that neighbour is inside the image
  if (live) {
    for (int tap = 0; tap < T; ++tap) {
      int dx = 0, dy = 0, dz = 0;
      if (ksz != 1) { dx = tap % 3 - 1; dy = (tap / 3) % 3 - 1; dz = (D > 1) ? tap / 9 - 1 : 0; }
      const int xx = x + dx, yy = y + dy, zz = z + dz;
      if (xx >= 0 && xx < W && yy >= 0 && yy < H && zz >= 0 && zz < D) okmask |= 1u << tap;
    }
  }
  const float* p0 = in0 + b * in0_bs + pix;
  const float* p1 = in1 ? in1 + b * in1_bs + pix : p0;
  unsigned char* th = s_tile[wid][0];
  unsigned char* tl = s_tile[wid][1];
  for (int kb = 0; kb < kp; kb += 64) {
#pragma unroll 1
    for (int g = 0; g < 4; ++g) {
      const int k0 = kb + g * 16;
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const unsigned tp = s_tap[k0 + j];
        const int off = s_off[k0 + j];
        const bool ok = tp != 0xFF && ((okmask >> (tp & 31)) & 1u);
        const float* src = (tp & 0x80) ? p1 : p0;
        v[j] = ok ? __ldg(src + off) * insc : ((k0 + j == ones_col) ? 1.f : 0.f);
      }
      uint32_t wh[8], wl[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) split2_rt(f16 != 0, v[2 * q], v[2 * q + 1], wh[q], wl[q]);
      const int o = lane * kI2cRow + g * 32;
      *reinterpret_cast<uint4*>(th + o) = make_uint4(wh[0], wh[1], wh[2], wh[3]);
      *reinterpret_cast<uint4*>(th + o + 16) = make_uint4(wh[4], wh[5], wh[6], wh[7]);
      *reinterpret_cast<uint4*>(tl + o) = make_uint4(wl[0], wl[1], wl[2], wl[3]);
      *reinterpret_cast<uint4*>(tl + o + 16) = make_uint4(wl[4], wl[5], wl[6], wl[7]);
    }
    __syncwarp();
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int id = it * 32 + lane, row = id >> 3, ch = id & 7;
      if (m0 + row < M) {
        const long long o = (m0 + row) * kp + kb + ch * 8;
        *reinterpret_cast<uint4*>(hi + o) = *reinterpret_cast<const uint4*>(th + row * kI2cRow + ch * 16);
        *reinterpret_cast<uint4*>(lo + o) = *reinterpret_cast<const uint4*>(tl + row * kI2cRow + ch * 16);
      }
    }
    __syncwarp();
  }
}
// The same rows for the shapes the chain sees most (2-D, 3x3, one source tensor, C channels known at compile time): the
// column -> (tap, channel) map, the nine neighbour offsets and the nine border predicates live in registers, so an
// element costs an address add and a predicated load instead of two shared-memory lookups, a bit test and two selects
// (the generic kernel is issue-bound at half of the HBM rate).  Output bit-identical to k_im2col_tc.
template <int C>
__global__ void __launch_bounds__(kI2cThreads)
k_im2col9_tc(const float* __restrict__ in0, long long in0_bs, int W, int H, long long px, long long M, int ones_col,
             int f16, const uint32_t* __restrict__ smax, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  constexpr int KR = 9 * C, KP = (KR + 63) / 64 * 64;
  float insc = 1.f;
  if (smax) { float inv; f16_scale_from_max(__ldg(smax), insc, inv); }
  __shared__ __align__(16) unsigned char s_tile[kI2cThreads / 32][2][32 * kI2cRow];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const long long m0 = (long long)blockIdx.x * kI2cPix + wid * 32;  // first pixel of this warp
  if (m0 >= M) return;
  const long long m = m0 + lane;
  const bool live = m < M;
  const long long mc = live ? m : M - 1;
  const long long b = mc / px;
  const int pix = (int)(mc - b * px);
  const int y = pix / W, x = pix - y * W;
  bool ok[9];
  int doff[9];
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const int dx = tap % 3 - 1, dy = tap / 3 - 1;
    ok[tap] = live && x + dx >= 0 && x + dx < W && y + dy >= 0 && y + dy < H;
    doff[tap] = dy * W + dx;
  }
  const float* p0 = in0 + b * in0_bs + pix;
  const int ipx = (int)px;
  unsigned char* th = s_tile[wid][0];
  unsigned char* tl = s_tile[wid][1];
#pragma unroll
  for (int kb = 0; kb < KP; kb += 64) {
    // small pixel counts (a batch shard of 8 at the coarse scales: fewer than three blocks per SM): one 64-column block per
    // grid row, so that the launch has KP / 64 times as many warps in flight (the kernel is latency-bound there)
    if (gridDim.y > 1 && kb / 64 != (int)blockIdx.y) continue;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int k = kb + g * 16 + j;  // compile-time after unrolling
        if (k < KR) {
          const int tap = k / C, ch = k - tap * C;
          v[j] = ok[tap] ? __ldg(p0 + (ch * ipx + doff[tap])) * insc : 0.f;
        } else {
          v[j] = (k == ones_col) ? 1.f : 0.f;
        }
      }
      uint32_t wh[8], wl[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) split2_rt(f16 != 0, v[2 * q], v[2 * q + 1], wh[q], wl[q]);
      const int o = lane * kI2cRow + g * 32;
      *reinterpret_cast<uint4*>(th + o) = make_uint4(wh[0], wh[1], wh[2], wh[3]);
      *reinterpret_cast<uint4*>(th + o + 16) = make_uint4(wh[4], wh[5], wh[6], wh[7]);
      *reinterpret_cast<uint4*>(tl + o) = make_uint4(wl[0], wl[1], wl[2], wl[3]);
      *reinterpret_cast<uint4*>(tl + o + 16) = make_uint4(wl[4], wl[5], wl[6], wl[7]);
    }
    __syncwarp();
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int id = it * 32 + lane, row = id >> 3, ch = id & 7;
      if (m0 + row < M) {
        const long long o = (m0 + row) * KP + kb + ch * 8;
        *reinterpret_cast<uint4*>(hi + o) = *reinterpret_cast<const uint4*>(th + row * kI2cRow + ch * 16);
        *reinterpret_cast<uint4*>(lo + o) = *reinterpret_cast<const uint4*>(tl + row * kI2cRow + ch * 16);
      }
    }
    __syncwarp();
  }
}
template <int C>
static void launch_im2col9(Ctx& c, const Geo& g, long long M, const float* in0, long long in0_bs, int ones_col,
                           const uint32_t* smax, Planes out) {
  constexpr int KB = (9 * C + 63) / 64;
  const unsigned nb = (unsigned)cdiv(M, kI2cPix);
  // (measured: at 512 blocks the split already costs the 448-column kernel 7 us, at <= 256 blocks it halves its time)
  const dim3 grid(nb, (nb < 148u * 3u && KB > 1) ? KB : 1, 1);
  k_im2col9_tc<C><<<grid, kI2cThreads, 0, c.st>>>(in0, in0_bs, g.W, g.H, g.px, M, ones_col,
                                                                     prec_f16(c.prec) ? 1 : 0, smax, out.hi, out.lo);
}
static bool im2col_fast_enabled() {
  static const bool on = [] { const char* e = getenv("INB_IM2COL_FAST"); return !(e && e[0] == '0'); }();
  return on;
}

void op_im2col_tc(Ctx& c, const Geo& g, int B, int k, const float* in0, long long in0_bs, int c0, const float* in1,
                  long long in1_bs, int C, int kp, int ones_col, Planes out, const uint32_t* smax) {
  if (c.dry()) return;
  const long long M = g.px * B;
  const int T = k == 1 ? 1 : (g.nd == 3 ? 27 : 9);
  INB_CHECK(kp <= kI2cMaxK && kp % 64 == 0, "im2col: unsupported row width %d", kp);
  INB_CHECK((long long)C * g.px + 2 * g.px < (1ll << 31), "im2col: sample too large for 32-bit offsets");
  Prof pf(c, F_LAYOUT_TC, 1, 0, (4.0 * C + 4.0 * kp) * M);
  if (k == 3 && g.nd == 2 && (in1 == nullptr || c0 >= C) && kp == (9 * C + 63) / 64 * 64 && im2col_fast_enabled()) {
    bool done = true;
    switch (C) {
      case 2: launch_im2col9<2>(c, g, M, in0, in0_bs, ones_col, smax, out); break;
      case 4: launch_im2col9<4>(c, g, M, in0, in0_bs, ones_col, smax, out); break;
      case 6: launch_im2col9<6>(c, g, M, in0, in0_bs, ones_col, smax, out); break;
      case 8: launch_im2col9<8>(c, g, M, in0, in0_bs, ones_col, smax, out); break;
      case 12: launch_im2col9<12>(c, g, M, in0, in0_bs, ones_col, smax, out); break;
      case 16: launch_im2col9<16>(c, g, M, in0, in0_bs, ones_col, smax, out); break;
      case 24: launch_im2col9<24>(c, g, M, in0, in0_bs, ones_col, smax, out); break;
      case 32: launch_im2col9<32>(c, g, M, in0, in0_bs, ones_col, smax, out); break;
      case 48: launch_im2col9<48>(c, g, M, in0, in0_bs, ones_col, smax, out); break;
      default: done = false;
    }
    if (done) {
      INB_CUDA(cudaGetLastError());
      return;
    }
  }
  k_im2col_tc<<<(unsigned)cdiv(M, kI2cPix), kI2cThreads, 0, c.st>>>(in0, in0_bs, c0, in1, in1_bs, C, T, k, g.W, g.H, g.D, g.px, M,
                                                          kp, ones_col, prec_f16(c.prec) ? 1 : 0, smax, out.hi, out.lo);
  INB_CUDA(cudaGetLastError());
}

// bits of max|x| over a (B, C, px) tensor (samples `bs` elements apart, `per` contiguous elements each); the bit
// patterns of non-negative floats are ordered like the values, so one atomicMax per warp finishes the reduction
template <int V>
__global__ void __launch_bounds__(256) k_absmax(const float* __restrict__ x, long long bs, long long per, int B,
                                                uint32_t* __restrict__ out) {
  const long long nv = per / V, n = nv * B;
  float m = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / nv, r = i - b * nv;
    if (V == 4) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(x + b * bs) + r);
      m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
    } else {
      m = fmaxf(m, fabsf(__ldg(x + b * bs + r)));
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xFFFFFFFFu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out, __float_as_uint(m));
}
void op_absmax(Ctx& c, long long px, int B, int C, const float* x, long long bs, uint32_t* smax) {
  if (c.dry()) return;
  const long long per = (long long)C * px;
  Prof pf(c, F_LAYOUT_TC, 1, 0, 4.0 * per * B);
  const bool v4 = per % 4 == 0 && bs % 4 == 0 && ((uintptr_t)x & 15) == 0;
  const unsigned grid = (unsigned)std::min<long long>(cdiv(per * B / (v4 ? 4 : 1), 256 * 4), 148 * 8);
  if (v4) k_absmax<4><<<std::max(grid, 1u), 256, 0, c.st>>>(x, bs, per, B, smax);
  else k_absmax<1><<<std::max(grid, 1u), 256, 0, c.st>>>(x, bs, per, B, smax);
  INB_CUDA(cudaGetLastError());
}

__global__ void k_pack_w_tc(int mode, int d0, int d1, int T, const float* __restrict__ w, int npad, int cpad,
                            int add_identity, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
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
      // residual skip folded into the weights: x + conv(x, W) == conv(x, W + I at the centre tap)
      if (add_identity && n == cc && tap == T / 2) v += 1.f;
    }
    __nv_bfloat16 h, l;
    split_bf16(v, h, l);
    hi[i] = h;
    lo[i] = l;
  }
}
void op_pack_w_tc(Ctx& c, int mode, int d0, int d1, int T, const float* w, int npad, int cpad, Planes out,
                  int add_identity) {
  if (c.dry()) return;
  const long long n = (long long)npad * T * cpad;
  Prof pf(c, F_PACK, 1, 0, 8.0 * n);
  k_pack_w_tc<<<(unsigned)std::min<long long>(cdiv(n, 256), 148 * 8), 256, 0, c.st>>>(mode, d0, d1, T, w, npad, cpad,
                                                                                    add_identity, out.hi, out.lo);
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
    const uint32_t h = hi[r * C2 + cp], l = lo ? lo[r * C2 + cp] : 0u;
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
  // single-pass bf16 keeps only the hi planes
  k_colsum_tc<<<(unsigned)cdiv(M, rpb), threads, 0, c.st>>>((const uint32_t*)in.hi, prec_terms(c.prec) == 3 ? (const uint32_t*)in.lo : nullptr, M, C2, rpb, out);
  INB_CUDA(cudaGetLastError());
}

}  // namespace inb
