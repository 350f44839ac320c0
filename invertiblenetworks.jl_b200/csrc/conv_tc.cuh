// conv_tc.cuh - tcgen05 tensor-core path of the ResidualBlock contractions (INB_PREC_BF16X3 / BF16).
//
// Internal layout: pixel-major "NHWC" bf16, row = pixel m = ((b*D+z)*H+y)*W+x, `pitch` channels per
// row.  Every fp32 value v is held as two bf16 planes hi = RN(v), lo = RN(v - hi); BF16X3 evaluates
// a*b as a_hi*b_hi + a_hi*b_lo + a_lo*b_hi (fp32 accumulate in TMEM), BF16 uses the hi planes only.
// A hidden activation X = relu(Y) stores -0.0 where Y < 0 and +0.0 where Y == 0, so the sign bit of
// the hi plane is the _relugrad mask (activation_functions.jl:84) while the MMA still sees zero.
#pragma once
#include "ops.cuh"
#include <cuda_bf16.h>

namespace inb {

struct Planes {
  __nv_bfloat16* hi;
  __nv_bfloat16* lo;
  int pitch;  // channels per row
  // The lo plane holds ONE byte per value: the upper byte (sign, exponent, two mantissa bits) of the half-precision
  // residual, rounded.  Only the hidden planes that k_rb_chain2 stores for k_wgrad2_tc use it (chain_planes_lo8):
  // 3 instead of 4 bytes per stored value on the one tensor family that dominates the step's DRAM traffic.
  int lo8 = 0;
};

// tile geometry of a P-pixel TMA box over (W,H,D,B); ok == false when the shape cannot be tiled
struct TileBox {
  int wt, ht, dt, bt;
  bool ok;
};
TileBox make_tile_box(const Geo& g, int B, int P);
bool tc_geometry_ok(const Geo& g, int B);

// (B,C,px) fp32 [two sources, conditional cat] -> planes [M][cpad] (channels >= Cin are zero)
void op_nchw_to_tc(Ctx& c, const Geo& g, int B, const float* in0, long long in0_bs, int c0, const float* in1,
                   long long in1_bs, int Cin, int cpad, Planes out);
// reference weight (C order w[d0][d1][T]) -> planes [npad][T*cpad] (K-major rows), mode PACK_CONV/PACK_DATA
// add_identity: add 1 on the diagonal of the centre tap (folds the block's residual skip into conv2 / dgrad2)
void op_pack_w_tc(Ctx& c, int mode, int d0, int d1, int T, const float* w, int npad, int cpad, Planes out,
                  int add_identity = 0);
// im2col rows for the fused chain's first GEMM and for the weight gradients (conv_tc.cu: k_im2col_tc):
// out [M][kp] with kp = chain_kpad(T*C); `ones_col` (>= 0) is a column of ones
// smax (INB_PREC_FP16X3, gradient operands): device word holding the bits of max|in|; the rows are scaled by the
// power of two f16_scale_from_max derives from it
void op_im2col_tc(Ctx& c, const Geo& g, int B, int k, const float* in0, long long in0_bs, int c0, const float* in1,
                  long long in1_bs, int C, int kp, int ones_col, Planes out, const uint32_t* smax = nullptr);
// bits of max|x| over a (B, C, px) tensor -> *smax (atomicMax; the caller zeroes it)
void op_absmax(Ctx& c, long long px, int B, int C, const float* x, long long bs, uint32_t* smax);
// per-channel sum over rows of planes [M][C] -> out[C] (bias gradients)
void op_colsum_tc(Ctx& c, long long M, int C, Planes in, float* out);

struct ConvTcSpec {
  Geo g;
  int B;
  int k;          // 1 or 3
  Planes in;      // [M][cpad_in]
  int cpad_in;    // multiple of 16
  Planes w;       // [npad][T*cpad_in]
  int N;          // MMA N = npad (multiple of 16, <= 256)
  int n_real;
  const float* bias;
  // mode 0: planes output (hidden / gradient of hidden)
  int mode;
  Planes out;
  int relu_encode;
  Planes mask;    // sign bit of hi -> zero the value, nullable
  // mode 1: fp32 (B,C,px) output
  float* out0; long long out0_bs; int n0;
  float* out1; long long out1_bs; int out1_accum;
  const float* add; long long add_bs; int add_n;
};
void op_conv_tc(Ctx& c, const ConvTcSpec& s);

struct WgradTcSpec {
  Geo g;
  int B;
  int k;
  Planes P;       // unshifted operand, [M][np] with np = number of its channels (multiple of 128)
  int np;
  Planes Q;       // tap-shifted operand, [M][cq]
  int cq;         // padded channels of Q (multiple of 16)
  int cq_real;
  float* dw;      // += D[tap][p][q] at ((p*cq_real + q)*T + (T-1-tap)); zeroed by the op
};
void op_wgrad_tc(Ctx& c, const WgradTcSpec& s);

// Weight / bias gradients of the fused path (conv_tc_wgrad2.cu): dw[p][c][T-1-tap] = sum_pix P[pix][p] Q[pix][tap*C+c],
// db[p] = sum_pix P[pix][p] (nullable).  P [M][np], Q [M][pitch >= T*C, multiple of 64].
struct Wgrad2TcSpec {
  long long M;
  Planes P;
  int np;
  Planes Q;
  int C, T;
  float* dw;
  float* db;
  const uint32_t* smax = nullptr;  // INB_PREC_FP16X3: one operand carries the gradient scale; the reduction divides by it
  int np_real = 0;                 // rows of dw / db that exist (P padded with zero channels up to np); 0 = np
  // >= 0: column `ones_col` of Q is 1.0 in every pixel row (a spare padding column of the im2col rows), so the
  // contraction itself yields db[p] = sum_pix P[pix][p] as column ones_col of the gradient tile - no bias warps, no
  // second read of the P tile from shared memory (the kernel is bound by shared-memory bandwidth where Q is narrow)
  int ones_col = -1;
};
void op_wgrad2_tc(Ctx& c, const Wgrad2TcSpec& s);
int wgrad_overlap_ctas();  // CTAs per weight-gradient kernel when they run under the next step's chain passes (0 = off)
void op_wgrad2_tc_multi(Ctx& c, const Wgrad2TcSpec* specs, int n);  // one reduction launch for up to three gradients

// The affine coupling folded into the col2im that finishes the block's last contraction (north_star: "the affine
// coupling (sigmoid scale, shift, log|s| sum) ... fused into the conv epilogues"): the kernel that gathers the tap
// (row) sums of P into (logS, T) applies sigma / S.X1 + T / sum log S (forward), the inverse, or the whole backward
// of invertible_layer_glow.jl:142-151 to the transformed half in place, so the block output never reaches HBM.
struct CouplingFuse {
  int mode;               // 0: Y1 = S.X1 + T (+ logdet)   1: X1 = (Y1 - T) / (S + eps)   2: backward (X1, dX1, dY3)
  int C1;                 // channels of the transformed half; the block has 2*C1 outputs (logS | T)
  float* a1; long long a1_bs;   // X1 -> Y1 (mode 0), Y1 -> X1 (modes 1, 2), in place
  float* d1; long long d1_bs;   // mode 2: dY1 -> dX1 in place
  float* dY3;             // mode 2: masked gradient of the block output (B, 2*C1, px), compact
  float low, high, invB;  // SigmoidLayer bounds; 1/B of the logdet terms (0 without logdet)
  double* ld;             // mode 0: logdet accumulator (nullable)
  unsigned* amax;         // mode 2: bits of max|dY3| (nullable)
  bool done;              // set by op_rb_chain when the fused kernel ran (the caller then skips op_coupling_*)
};

// Fused pass of the block's three contractions (conv_tc_chain.cu): im2col-GEMM -> per-pixel GEMM ->
// tap-expanded GEMM + col2im.  mode 0 = forward (bias + ReLU epilogues), mode 1 = backward (relu-grad masks).
struct ChainSpec {
  Geo g;
  int B;
  int k1;
  int nh;             // hidden channels of the planes: 128 or 256 (n_hidden padded with zero channels, chain_nh_pad)
  int nh_real;        // the block's n_hidden (length of the bias vectors); 0 = nh
  Planes in;          // im2col rows [M][kp], pitch = kp (multiple of 64)
  Planes w1, w2, w3;  // [nh][kp], [nh][nh] (+I), tap-expanded [n3pad][nh]
  int Cn;             // real output channels of the last contraction
  int mode;
  const float *bias1, *bias2;
  const uint32_t *mask1, *mask2;  // mode 1: relu-grad masks of the two epilogues, bit planes [M][nh/32]
  uint32_t *bits1, *bits2;        // mode 0 with o1/o2: the bit planes this pass writes (signs of Y1 / Y2)
  Planes o1, o2;      // hidden outputs in HBM (hi == nullptr: not stored)
  float* P;           // scratch [M][chain_n3pad(taps, Cn)]
  // output of col2im, same meaning as ConvTcSpec mode 1
  float* out0; long long out0_bs; int n0;
  float* out1; long long out1_bs; int out1_accum;
  const float* add; long long add_bs; int add_n;
  const uint32_t* smax = nullptr;  // INB_PREC_FP16X3 backward pass: the scale of `in` (col2im divides by it)
  CouplingFuse* fuse = nullptr;    // fold the affine coupling into the col2im (out0 etc. are then unused)
};
int chain_n3pad(int taps, int Cn);
// the hidden planes a storing pass of the fused chain writes carry 1-byte lo planes (Planes::lo8) in this precision mode
bool chain_planes_lo8(int prec);
inline int chain_nh_pad(int nh) { return nh <= 128 ? 128 : 256; }  // hidden width the chain kernels run at
int chain_kpad(int taps, int C, int extra);  // im2col width: taps*C (+ extra columns) rounded up to 64
bool chain_supported(const Geo& g, int B, int k1, int k2, int nh, int C_in, int Cn);
// the three packed operands of one chain pass in a single launch (conv_tc_chain.cu: k_pack_chain_tc)
// nh = padded hidden width of the planes, nhr = the block's n_hidden (rows / columns beyond it are zero)
void op_pack_chain_tc(Ctx& c, int nh, int nhr, int T, int C1, int kp, const float* wa, const float* wb, int w2_data, int Cn,
                      int n3pad, const float* wc, Planes w1, Planes w2, Planes w3);
// the same packing for several blocks of one shape in ONE launch (the K flow steps of a scale: their weights do not
// change during a network-level call, so all of them are packed before the first step runs)
struct PackChainItem {
  const float *wa, *wb, *wc;
  Planes w1, w2, w3;
};
void op_pack_chain_multi(Ctx& c, int nh, int nhr, int T, int C1, int kp, int w2_data, int Cn, int n3pad,
                         const PackChainItem* items, int n);
void op_rb_chain(Ctx& c, const ChainSpec& s);
void chain_set_trace(long long* p);

}  // namespace inb
