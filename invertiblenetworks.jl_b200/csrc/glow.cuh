// glow.cuh - layer compositions shared by the layer-level and network-level entry points.
#pragma once
#include "ops.cuh"
#include "conv_tc.cuh"
#include <vector>

namespace inb {

struct RBShape {
  Geo g;
  int B;
  int c0;    // channels taken from the flow tensor (X2)
  int ccond; // channels taken from the condition (0 when unconditional)
  int nh;
  int Cout;
  int k1, k2;
  int Cin() const { return c0 + ccond; }
  int T1() const { return k1 == 1 ? 1 : (g.nd == 3 ? 27 : 9); }
  int T2() const { return k2 == 1 ? 1 : (g.nd == 3 ? 27 : 9); }
};
// chain operands of a block packed ahead of time (rb_prepack_chain): [0] the forward / recompute pass, [1] the backward pass
struct PackedW {
  Planes w1, w2, w3;
};
struct RBParams {
  const float *W1, *W2, *W3, *b1, *b2;
  const PackedW* pre[2] = {nullptr, nullptr};
};
struct RBGrads {
  float *W1, *W2, *W3, *b1, *b2;
};
// hidden activations of one block, owned by the caller's arena
struct RBHidden {
  float* Y1;  // (B, nh, px) pre-activation of conv1
  float* Y2;  // (B, nh, px) pre-activation of conv2 (+skip)
  float* G;   // (B, nh, px) gradient scratch (backward only)
  // tensor-core path: the same three buffers hold bf16 hi/lo planes [M][nh]; xin is the padded
  // bf16 copy of the block input made by rb_forward (lives in the caller's arena scope)
  Planes xin{nullptr, nullptr, 0};
  int xin_ones = -1;  // fused chain: column of the im2col rows that holds 1.0 (the bias gradient rides on dW1), -1 = none
  // fused-chain path: relu-grad masks of Y1 / Y2 as bit planes [M][nh/32], written by the storing forward pass
  uint32_t* bm1 = nullptr;
  uint32_t* bm2 = nullptr;
  // INB_PREC_FP16X3: device word with the bits of max|dY3|, when the producer of dY3 already reduced it (the coupling
  // backward kernel does); rb_backward computes it with a pass over dY3 otherwise
  uint32_t* dy_absmax = nullptr;
  // request to fold the affine coupling into the block's last kernel (rb_forward sets fuse->done when it did)
  CouplingFuse* fuse = nullptr;
};

// layer_residual_block.jl:119-134, output = PRE-activation Y3 (B, Cout, px) compact; the consumers
// apply the final ReLU (:133) themselves.
void rb_forward(Ctx& c, const RBShape& s, View x2, View cond, const RBParams& p, RBHidden& h, float* Y3);
// layer_residual_block.jl:137-178 given the already ReLU-masked dY3 (:150) and the hidden
// activations of rb_forward.  dX2 (+= passthrough `add` when non-null) and dCond (accumulated).
void rb_backward(Ctx& c, const RBShape& s, const float* dY3, View x2, View cond, const RBParams& p,
                 RBHidden& h, const RBGrads& gr, View dx2, const float* add, long long add_bs,
                 View dcond);

struct FlowShape {
  Geo g;
  int B, C, ccond, nh, k1, k2;
  float low, high;
  int logdet, freeze;
  int C1() const { return split_k(C); }
  RBShape rb() const { return RBShape{g, B, C - C1(), ccond, nh, 2 * C1(), k1, k2}; }
};
struct FlowParams {
  const float *s, *b;  // ActNorm (nullable: coupling layer alone)
  const float *v1, *v2, *v3;
  RBParams rb;
};
struct FlowGrads {
  float *s, *b, *v1, *v2, *v3;
  RBGrads rb;
};
size_t rb_hidden_elems(const RBShape& s);
// Packs the chain operands of n blocks of shape s (direction 0: forward / recompute, 1: backward) into planes taken
// from the caller's arena scope, one launch per 24 blocks; returns false (and packs nothing) when the blocks do not run
// on the fused chain.  The caller points RBParams::pre[direction] at out[i].
bool rb_prepack_chain(Ctx& c, const RBShape& s, const RBParams* prm, int n, int direction, PackedW* out);
// the same for a block on the UNFUSED tensor-core convolutions (k2 = 3: the reference's default HINT block):
// direction 0 = (W1 conv, W2 + I conv, W3 data), direction 1 = (W3 conv, W2 + I data, W1 data)
bool rb_prepack_unfused(Ctx& c, const RBShape& s, const RBParams& prm, int direction, PackedW* out);

// ActNorm -> CouplingLayerGlow forward: x -> y (y != x)
void flow_forward(Ctx& c, const FlowShape& f, View x, View y, View cond, const FlowParams& p, double* ld);
// CouplingLayerGlow.inverse -> ActNorm.inverse: y (clobbered) -> x
void flow_inverse(Ctx& c, const FlowShape& f, View y, View x, View cond, const FlowParams& p);
// backward of both; (dy, y) are clobbered and (dx, x) may alias them
void flow_backward(Ctx& c, const FlowShape& f, View dy, View y, View dx, View x, View cond, View dcond,
                   const FlowParams& p, const FlowGrads& g);

// ---------------------------------------------------------------- HINT family (hint.cu)
// CouplingLayerHINT (invertible_layer_hint.jl:52-300) over CouplingLayerBasic (invertible_layer_basic.jl:62-166),
// all on channel-range views of ONE tensor (no tensor_split / tensor_cat copies).
enum { HINT_PERMUTE_NONE = 0, HINT_PERMUTE_FULL = 1, HINT_PERMUTE_LOWER = 2, HINT_PERMUTE_BOTH = 3 };
struct HintShape {
  Geo g;
  int B, C, nh, k1, k2;
  float low, high;
  int logdet;
  int permute;       // HINT_PERMUTE_*
  int shared_last;   // 1: gradients of a coupling layer visited several times keep the LAST visit only (the
                     // reference's set_grad=true, layer_residual_block.jl:168-172); 0: they are summed (hint.jl:222)
  int depth() const;  // get_depth, hint.jl:63-71
};
struct HintParams {
  std::vector<RBParams> cl;  // CL[1..depth].RB
  const float *v1 = nullptr, *v2 = nullptr, *v3 = nullptr;  // C (permute full / lower)
  const float *s = nullptr, *b = nullptr;  // an ActNorm in front (NetworkMultiScaleHINT), fused with `full`
};
struct HintGrads {
  std::vector<RBGrads> cl;
  float *v1 = nullptr, *v2 = nullptr, *v3 = nullptr, *s = nullptr, *b = nullptr;
};
void hint_check(const HintShape& h);
// CouplingLayerBasic on views (h.C is ignored, Ca = channels of X1 = channels of X2; invertible_layer_basic.jl:90-149):
//   forward  xb <- S .* xb + T with (logS, T) = RB(xa)           inverse  yb <- (yb - T) ./ (S + eps)
//   backward (dyb, yb) <- (dX2, X2) in place, dxa += RB.backward(...) (the caller pre-loads dxa with dY1)
void basic_forward(Ctx& c, const HintShape& h, int Ca, View xa, View xb, const RBParams& p, double* ld, int ld_batch = 0);
void basic_inverse(Ctx& c, const HintShape& h, int Ca, View xa, View yb, const RBParams& p);
void basic_backward(Ctx& c, const HintShape& h, int Ca, View xa, View dxa, View yb, View dyb, const RBParams& p,
                    const RBGrads& g, bool accumulate);
// [ActNorm ->] CouplingLayerHINT.forward: x -> y (y != x); ld accumulates both logdets
void hint_forward(Ctx& c, const HintShape& h, View x, View y, const HintParams& p, double* ld);
// CouplingLayerHINT.inverse [-> ActNorm.inverse]: y (clobbered) -> x (may alias y)
void hint_inverse(Ctx& c, const HintShape& h, View y, View x, const HintParams& p);
// backward of both: (dy, y) clobbered, (dx, x) may alias them
void hint_backward(Ctx& c, const HintShape& h, View dy, View y, View dx, View x, const HintParams& p,
                   const HintGrads& g);

}  // namespace inb
