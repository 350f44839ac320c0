// ops.cuh - launch wrappers of the sm_100a kernels (one family per .cu file).
#pragma once
#include "common.cuh"

namespace inb {

// ---------------------------------------------------------------- elementwise.cu
// dimensionality_operations.jl:79-107 / 137-166 (checkerboard), index maps only
void op_squeeze(Ctx& c, const Geo& gin, int B, int C, View in, View out);
void op_unsqueeze(Ctx& c, const Geo& gout, int B, int Cout, View in, View out);
void op_copy(Ctx& c, long long px, int B, int C, View in, View out);
// one-level Haar transform + patch squeeze, 2-D (dimensionality_operations.jl:199-258 with WT.db1: type 0;
// Haar_squeeze / invHaar_unsqueeze :318-371: type 1).  g, C: the FULL-resolution geometry / channel count.
void op_haar_squeeze(Ctx& c, const Geo& g, int B, int C, int type, View in, View out);
void op_haar_unsqueeze(Ctx& c, const Geo& g, int B, int C, int type, View in, View out);
// dst[i] += src[i] for up to five pairs in one launch
void op_accum(Ctx& c, int npairs, const float* const* src, float* const* dst, const long long* n);
void op_zero(Ctx& c, void* p, size_t bytes);

// invertible_layer_actnorm.jl:67-72
void op_actnorm_init(Ctx& c, long long px, int B, int C, View x, float* s, float* b);

// ActNorm (actnorm.jl:73) followed by the Householder mix (conv1x1.jl:174-189) in one HBM pass.
// s/b nullable (no ActNorm); v1 nullable (no Conv1x1).  ld: device double accumulator that
// receives px*sum(log|s|) (actnorm.jl:185-195), nullable.
void op_an_hh_fwd(Ctx& c, long long px, int B, int C, View x, View y, const float* s, const float* b,
                  const float* v1, const float* v2, const float* v3, double* ld);
// inverse mix (conv1x1.jl:209-224) followed by ActNorm inverse (actnorm.jl:93)
void op_hh_an_inv(Ctx& c, long long px, int B, int C, View y, View x, const float* s, const float* b,
                  const float* v1, const float* v2, const float* v3);
// backward of both (conv1x1.jl:227-245, actnorm.jl:100-123): (dY,Y) -> (dX,X), may run in place.
// gram: C*C doubles += A^T dY (A = layer input of Conv1x1); dsdb: 2*C doubles += (sum dA*X, sum dA).
void op_hh_an_bwd(Ctx& c, long long px, int B, int C, View dy, View y, View dx, View x, const float* s,
                  const float* b, const float* v1, const float* v2, const float* v3, double* gram,
                  double* dsdb);
// gram -> dv1,dv2,dv3 (conv1x1.jl:118-170 restructured); dsdb -> ds (- px/s when logdet), db
void op_hh_grad_finish(Ctx& c, int C, const double* gram, const float* v1, const float* v2,
                       const float* v3, int freeze, float* dv1, float* dv2, float* dv3);
void op_hh_an_grad_finish(Ctx& c, int C, long long px, const double* gram, const float* v1, const float* v2, const float* v3,
                          int freeze, float* dv1, float* dv2, float* dv3, const double* dsdb, const float* s, int logdet,
                          float* ds, float* db);
void op_an_grad_finish(Ctx& c, int C, long long px, const double* dsdb, const float* s, int logdet,
                       float* ds, float* db);

// affine coupling (invertible_layer_glow.jl:110-116,121-129,142-157).  rb: (B, 2*C1, px) compact
// pre-activation output Y3 of the ResidualBlock; x1: view of the C1 transformed channels.
// ld_batch: the batch size the logdet is divided by (0: B; the HINT level pass runs on B * groups virtual samples)
void op_coupling_fwd(Ctx& c, long long px, int B, int C1, View x1, View y1, const float* rb, float low,
                     float high, double* ld, int ld_batch = 0);
void op_coupling_inv(Ctx& c, long long px, int B, int C1, View y1, View x1, const float* rb, float low,
                     float high);
// y1 -> x1, dy1 -> dx1 (views, in place allowed), rb (Y3) -> dY3 in place
// amax (nullable): atomicMax of the bit patterns of |masked gradient of the block output| (zeroed by the caller)
void op_coupling_bwd(Ctx& c, long long px, int B, int C1, View y1, View x1, View dy1, View dx1,
                     float* rb, float low, float high, int logdet, unsigned* amax = nullptr);
void op_relu_copy(Ctx& c, long long n, const float* in, float* out);
// _relugrad (activation_functions.jl:84): out = y < 0 ? 0 : dy
void op_relu_grad(Ctx& c, long long n, const float* dy, const float* y, float* out);
// per-channel sum over (B, px) of a compact (B,C,px) tensor -> out[C] (bias gradients)
void op_channel_sum(Ctx& c, long long px, int B, int C, const float* in, float* out);
void op_nll_grad(Ctx& c, long long n, int B, const float* z, float* dz, double* acc, float* loss);
void op_ld_finish(Ctx& c, const double* acc, float* out);
void op_adam(Ctx& c, long long n, float* x, const float* g, float* m, float* v, float lr, float b1, float b2, float eps,
             float b1t, float b2t);

// ---------------------------------------------------------------- conv_simt.cu (INB_PREC_FP32)
enum { PACK_CONV = 0, PACK_DATA = 1 };
// reference weight w[d0][d1][taps] (C order) -> Wm[tap][c][o] used by the implicit GEMMs:
//  PACK_CONV: NNlib conv            (o=d0, c=d1), Wm[tap][c][o] = w[o][c][T-1-tap]
//  PACK_DATA: NNlib \nabla conv_data (c=d0, o=d1), Wm[tap][c][o] = w[c][o][tap]
void op_pack_w(Ctx& c, int mode, int d0, int d1, int T, const float* w, float* Wm);
// dWm[tap][c][o] -> dw[o][c][T-1-tap]  (\nabla conv_filter of the NNlib conv)
void op_unpack_dw(Ctx& c, int O, int Cc, int T, const float* dWm, float* dw);

struct ConvSpec {
  Geo g;
  int B;
  int k;  // 1 or 3 ("same" padding)
  // input: Cin channels; the first c0 come from in0, the rest from in1 (conditional cat)
  const float* in0;
  long long in0_bs;
  int c0;
  const float* in1;
  long long in1_bs;
  int Cin;
  int relu_in;
  const float* Wm;  // [T*Cin][N]
  int N;
  // output: first n0 channels to out0, rest to out1 (+= when out1_accum)
  float* out0;
  long long out0_bs;
  int n0;
  float* out1;
  long long out1_bs;
  int out1_accum;
  // epilogue: val = acc + bias[o] + (add_relu ? relu(add) : add)[o,pix]; val = mask[o,pix] < 0 ? 0 : val
  const float* bias;
  const float* add;
  long long add_bs;
  int add_relu;
  int add_n;  // `add` applies to output channels < add_n
  const float* mask;
  long long mask_bs;
};
void op_conv_simt(Ctx& c, const ConvSpec& s);

struct WgradSpec {
  Geo g;
  int B;
  int k;
  const float* in0;
  long long in0_bs;
  int c0;
  const float* in1;
  long long in1_bs;
  int Cin;
  int relu_in;
  const float* dy;  // (B, N, px) view
  long long dy_bs;
  int N;
  int relu_dy;
  float* dWm;  // [T*Cin][N], zeroed by the op
};
void op_wgrad_simt(Ctx& c, const WgradSpec& s);

}  // namespace inb
