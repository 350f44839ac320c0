// layers.cu - ResidualBlock and flow-step (ActNorm -> Conv1x1 -> affine coupling) compositions.
#include "glow.cuh"
#include <algorithm>
#include <vector>

namespace inb {

// hidden buffers hold (B, nh, px) fp32 on the CUDA-core path and bf16 / half hi+lo planes [M][nh padded to 128 | 256]
// (same bytes per channel) on the fused tensor-core chain
size_t rb_hidden_elems(const RBShape& s) {
  const int nh = s.nh <= 256 ? chain_nh_pad(s.nh) : s.nh;
  return (size_t)s.B * nh * s.g.px;
}

static ConvSpec conv_base(const RBShape& s) {
  ConvSpec cs{};
  cs.g = s.g;
  cs.B = s.B;
  cs.add_n = 1 << 30;
  return cs;
}

// ---------------------------------------------------------------- ResidualBlock, fp32 SIMT
static void rb_forward_fp32(Ctx& c, const RBShape& s, View x2, View cond, const RBParams& p, RBHidden& h,
                            float* Y3) {
  const long long px = s.g.px;
  const int Cin = s.Cin(), nh = s.nh, T1 = s.T1(), T2 = s.T2();
  size_t m = c.ar->mark();
  float* Wm1 = c.ar->f32((size_t)T1 * Cin * nh);
  float* Wm2 = c.ar->f32((size_t)T2 * nh * nh);
  float* Wm3 = c.ar->f32((size_t)T1 * nh * s.Cout);
  op_pack_w(c, PACK_CONV, nh, Cin, T1, p.W1, Wm1);
  op_pack_w(c, PACK_CONV, nh, nh, T2, p.W2, Wm2);
  op_pack_w(c, PACK_DATA, nh, s.Cout, T1, p.W3, Wm3);
  {  // Y1 = conv(X, W1) + b1                                  layer_residual_block.jl:122
    ConvSpec cs = conv_base(s);
    cs.k = s.k1;
    cs.in0 = x2.p; cs.in0_bs = x2.bs; cs.c0 = s.c0;
    cs.in1 = cond.p; cs.in1_bs = cond.bs;
    cs.Cin = Cin; cs.Wm = Wm1; cs.N = nh;
    cs.out0 = h.Y1; cs.out0_bs = (long long)nh * px; cs.n0 = nh;
    cs.bias = p.b1;
    op_conv_simt(c, cs);
  }
  {  // Y2 = X2 + conv(X2, W2) + b2, X2 = relu(Y1)             :123-125
    ConvSpec cs = conv_base(s);
    cs.k = s.k2;
    cs.in0 = h.Y1; cs.in0_bs = (long long)nh * px; cs.c0 = nh; cs.Cin = nh; cs.relu_in = 1;
    cs.Wm = Wm2; cs.N = nh;
    cs.out0 = h.Y2; cs.out0_bs = (long long)nh * px; cs.n0 = nh;
    cs.bias = p.b2;
    cs.add = h.Y1; cs.add_bs = (long long)nh * px; cs.add_relu = 1;
    op_conv_simt(c, cs);
  }
  {  // Y3 = \nabla conv_data(X3, W3), X3 = relu(Y2)            :126-129
    ConvSpec cs = conv_base(s);
    cs.k = s.k1;
    cs.in0 = h.Y2; cs.in0_bs = (long long)nh * px; cs.c0 = nh; cs.Cin = nh; cs.relu_in = 1;
    cs.Wm = Wm3; cs.N = s.Cout;
    cs.out0 = Y3; cs.out0_bs = (long long)s.Cout * px; cs.n0 = s.Cout;
    op_conv_simt(c, cs);
  }
  c.ar->release(m);
}

static void rb_backward_fp32(Ctx& c, const RBShape& s, const float* dY3, View x2, View cond,
                             const RBParams& p, RBHidden& h, const RBGrads& gr, View dx2,
                             const float* add, long long add_bs, View dcond) {
  const long long px = s.g.px;
  const int Cin = s.Cin(), nh = s.nh, T1 = s.T1(), T2 = s.T2(), Cout = s.Cout;
  const long long hbs = (long long)nh * px;
  size_t m = c.ar->mark();
  size_t wmax = std::max((size_t)T1 * std::max(Cin, Cout) * nh, (size_t)T2 * nh * nh);
  float* Wm = c.ar->f32(wmax);
  float* dWm = c.ar->f32(wmax);
  float* G2 = h.G;   // dY2
  float* G1 = h.Y2;  // dY1 reuses the Y2 buffer once dW3 has consumed relu(Y2)

  // dX3 = conv(dY3, W3); dY2 = relugrad(dX3, Y2)               layer_residual_block.jl:151,154
  op_pack_w(c, PACK_CONV, nh, Cout, T1, p.W3, Wm);
  {
    ConvSpec cs = conv_base(s);
    cs.k = s.k1;
    cs.in0 = dY3; cs.in0_bs = (long long)Cout * px; cs.c0 = Cout; cs.Cin = Cout;
    cs.Wm = Wm; cs.N = nh;
    cs.out0 = G2; cs.out0_bs = hbs; cs.n0 = nh;
    cs.mask = h.Y2; cs.mask_bs = hbs;
    op_conv_simt(c, cs);
  }
  {  // dW3 = \nabla conv_filter(dY3, X3)                         :152
    WgradSpec ws{};
    ws.g = s.g; ws.B = s.B; ws.k = s.k1;
    ws.in0 = dY3; ws.in0_bs = (long long)Cout * px; ws.c0 = Cout; ws.Cin = Cout;
    ws.dy = h.Y2; ws.dy_bs = hbs; ws.N = nh; ws.relu_dy = 1;
    ws.dWm = dWm;
    op_wgrad_simt(c, ws);
    op_unpack_dw(c, nh, Cout, T1, dWm, gr.W3);
  }
  // dX2 = \nabla conv_data(dY2, W2) + dY2; dY1 = relugrad(dX2, Y1)   :155,161
  op_pack_w(c, PACK_DATA, nh, nh, T2, p.W2, Wm);
  {
    ConvSpec cs = conv_base(s);
    cs.k = s.k2;
    cs.in0 = G2; cs.in0_bs = hbs; cs.c0 = nh; cs.Cin = nh;
    cs.Wm = Wm; cs.N = nh;
    cs.out0 = G1; cs.out0_bs = hbs; cs.n0 = nh;
    cs.add = G2; cs.add_bs = hbs;
    cs.mask = h.Y1; cs.mask_bs = hbs;
    op_conv_simt(c, cs);
  }
  {  // dW2 = \nabla conv_filter(X2, dY2); db2 = sum dY2          :156-157
    WgradSpec ws{};
    ws.g = s.g; ws.B = s.B; ws.k = s.k2;
    ws.in0 = h.Y1; ws.in0_bs = hbs; ws.c0 = nh; ws.Cin = nh; ws.relu_in = 1;
    ws.dy = G2; ws.dy_bs = hbs; ws.N = nh;
    ws.dWm = dWm;
    op_wgrad_simt(c, ws);
    op_unpack_dw(c, nh, nh, T2, dWm, gr.W2);
    op_channel_sum(c, px, s.B, nh, G2, gr.b2);
  }
  // dX1 = \nabla conv_data(dY1, W1) (+ passthrough)              :162
  op_pack_w(c, PACK_DATA, nh, Cin, T1, p.W1, Wm);
  {
    ConvSpec cs = conv_base(s);
    cs.k = s.k1;
    cs.in0 = G1; cs.in0_bs = hbs; cs.c0 = nh; cs.Cin = nh;
    cs.Wm = Wm; cs.N = Cin;
    cs.out0 = dx2.p; cs.out0_bs = dx2.bs; cs.n0 = s.c0;
    cs.out1 = dcond.p; cs.out1_bs = dcond.bs; cs.out1_accum = 1;
    cs.add = add; cs.add_bs = add_bs; cs.add_n = s.c0;
    op_conv_simt(c, cs);
  }
  {  // dW1 = \nabla conv_filter(X1, dY1); db1 = sum dY1          :163-164
    WgradSpec ws{};
    ws.g = s.g; ws.B = s.B; ws.k = s.k1;
    ws.in0 = x2.p; ws.in0_bs = x2.bs; ws.c0 = s.c0; ws.in1 = cond.p; ws.in1_bs = cond.bs; ws.Cin = Cin;
    ws.dy = G1; ws.dy_bs = hbs; ws.N = nh;
    ws.dWm = dWm;
    op_wgrad_simt(c, ws);
    op_unpack_dw(c, nh, Cin, T1, dWm, gr.W1);
    op_channel_sum(c, px, s.B, nh, G1, gr.b1);
  }
  c.ar->release(m);
}

// ---------------------------------------------------------------- ResidualBlock, tcgen05 path
static int pad16(int n) { return (n + 15) / 16 * 16; }

static bool rb_chain_ok(const RBShape& s);
// Which kernels run a block under a tensor-core precision mode: the fused chain (k2 = 1, any n_hidden <= 256: blocks
// narrower than 128 / 256 hidden channels are zero padded), else the unfused tensor-core convolutions (n_hidden in
// {128, 256}; bf16 operands - under fp16x3 they compute in bf16x3), else - shapes neither can tile - the fp32
// CUDA-core kernels.  A pure function of the shape: forward, recompute and backward of a block agree on it.
enum { RB_PATH_SIMT = 0, RB_PATH_CHAIN = 1, RB_PATH_UNFUSED = 2 };
static int rb_path(const Ctx& c, const RBShape& s) {
  if (c.prec == 0) return RB_PATH_SIMT;
  static const bool strict = [] { const char* e = getenv("INB_STRICT_TC"); return e && e[0] == '1'; }();
  const bool geo = tc_geometry_ok(s.g, s.B);
  if (geo && rb_chain_ok(s)) return RB_PATH_CHAIN;
  const bool unfused = geo && s.nh % 128 == 0 && s.nh <= 256 && pad16(s.Cin()) <= 256 && pad16(s.Cout) <= 256;
  if (unfused) return RB_PATH_UNFUSED;
  if (strict) {
    INB_CHECK(s.nh % 128 == 0 && s.nh <= 256,
              "the tensor-core path needs n_hidden in {128, 256} (got %d) unless k2 = 1; use precision fp32", s.nh);
    INB_CHECK(pad16(s.Cin()) <= 256 && pad16(s.Cout) <= 256, "the tensor-core path supports up to 256 channels");
    INB_CHECK(geo, "spatial size %dx%dx%d cannot be tiled for the tensor-core path; use precision fp32", s.g.W, s.g.H, s.g.D);
  }
  return RB_PATH_SIMT;
}
static Planes planes_at(void* mem, long long M, int pitch) {
  Planes p;
  p.hi = reinterpret_cast<__nv_bfloat16*>(mem);
  p.lo = p.hi + M * pitch;
  p.pitch = pitch;
  return p;
}
static Planes planes_new(Ctx& c, long long rows, int pitch) {
  return planes_at(c.ar->alloc_bytes((size_t)rows * pitch * 2 * 2), rows, pitch);
}
// the fused chain runs both directions of the block or neither (the hidden tensors change layout owner)
static bool rb_chain_ok(const RBShape& s) {
  return chain_supported(s.g, s.B, s.k1, s.k2, s.nh, s.Cin(), s.Cout) &&
         chain_supported(s.g, s.B, s.k1, s.k2, s.nh, s.Cout, s.Cin());
}
bool rb_prepack_chain(Ctx& c, const RBShape& s, const RBParams* prm, int n, int direction, PackedW* out) {
  if (rb_path(c, s) != RB_PATH_CHAIN || n <= 0) return false;
  const int T1 = s.T1(), nh = chain_nh_pad(s.nh);
  const int C1 = direction == 0 ? s.Cin() : s.Cout, Cn = direction == 0 ? s.Cout : s.Cin();
  const int kp = chain_kpad(T1, C1, 0), n3pad = chain_n3pad(T1, Cn);
  std::vector<PackChainItem> items(n);
  for (int i = 0; i < n; ++i) {
    out[i].w1 = planes_new(c, nh, kp);
    out[i].w2 = planes_new(c, nh, nh);
    out[i].w3 = planes_new(c, n3pad, nh);
    items[i].wa = direction == 0 ? prm[i].W1 : prm[i].W3;
    items[i].wb = prm[i].W2;
    items[i].wc = direction == 0 ? prm[i].W3 : prm[i].W1;
    items[i].w1 = out[i].w1; items[i].w2 = out[i].w2; items[i].w3 = out[i].w3;
  }
  op_pack_chain_multi(c, nh, s.nh, T1, C1, kp, direction, Cn, n3pad, items.data(), n);
  return true;
}

bool rb_prepack_unfused(Ctx& c, const RBShape& s, const RBParams& p, int direction, PackedW* out) {
  if (rb_path(c, s) != RB_PATH_UNFUSED) return false;
  const int Cin = s.Cin(), T1 = s.T1(), T2 = s.T2(), nh = s.nh;
  const int cin_pad = pad16(Cin), cout_pad = pad16(s.Cout);
  if (direction == 0) {
    out->w1 = planes_new(c, nh, T1 * cin_pad);
    out->w2 = planes_new(c, nh, T2 * nh);
    out->w3 = planes_new(c, cout_pad, T1 * nh);
    op_pack_w_tc(c, PACK_CONV, nh, Cin, T1, p.W1, nh, cin_pad, out->w1);
    op_pack_w_tc(c, PACK_CONV, nh, nh, T2, p.W2, nh, nh, out->w2, 1);  // + I: the skip of :125
    op_pack_w_tc(c, PACK_DATA, nh, s.Cout, T1, p.W3, cout_pad, nh, out->w3);
  } else {
    out->w1 = planes_new(c, nh, T1 * cout_pad);
    out->w2 = planes_new(c, nh, T2 * nh);
    out->w3 = planes_new(c, cin_pad, T1 * nh);
    op_pack_w_tc(c, PACK_CONV, nh, s.Cout, T1, p.W3, nh, cout_pad, out->w1);
    op_pack_w_tc(c, PACK_DATA, nh, nh, T2, p.W2, nh, nh, out->w2, 1);  // + I: the '+ dY2' of :155
    op_pack_w_tc(c, PACK_DATA, nh, Cin, T1, p.W1, cin_pad, nh, out->w3);
  }
  return true;
}

static ConvTcSpec tc_base(const RBShape& s) {
  ConvTcSpec cs{};
  cs.g = s.g;
  cs.B = s.B;
  cs.add_n = 1 << 30;
  return cs;
}

static void rb_forward_tc(Ctx& c, const RBShape& s, View x2, View cond, const RBParams& p, RBHidden& h, float* Y3) {
  const long long px = s.g.px, M = px * s.B;
  const int Cin = s.Cin(), T1 = s.T1(), T2 = s.T2();
  const int cin_pad = pad16(Cin), cout_pad = pad16(s.Cout);
  const bool chain = rb_path(c, s) == RB_PATH_CHAIN;
  const int nh = chain ? chain_nh_pad(s.nh) : s.nh;
  if (chain) {
    // one fused kernel for the three contractions (conv_tc_chain.cu); the hidden tensors reach HBM only
    // when a backward pass follows (h.G != nullptr: the recompute of flow_backward).  The im2col rows of the
    // block input outlive this call: the weight gradient of conv1 reads them again.
    const int kp = chain_kpad(T1, Cin, 0);
    h.xin = planes_new(c, M, kp);
    if (h.G) {  // the mask bit planes outlive this call as well (rb_backward_tc reads them)
      h.bm1 = reinterpret_cast<uint32_t*>(c.ar->alloc_bytes((size_t)M * (nh / 32) * 4));
      h.bm2 = reinterpret_cast<uint32_t*>(c.ar->alloc_bytes((size_t)M * (nh / 32) * 4));
    }
    // a spare padding column of the rows carries 1.0: the W1 weight gradient then delivers db1 as well (its packed
    // weight column is zero, so the forward contraction does not see it)
    h.xin_ones = (T1 * Cin < kp) ? T1 * Cin : -1;
    op_im2col_tc(c, s.g, s.B, s.k1, x2.p, x2.bs, s.c0, cond.p, cond.bs, Cin, kp, h.xin_ones, h.xin);
    size_t m = c.ar->mark();
    Planes H1 = planes_at(h.Y1, M, nh), H2 = planes_at(h.Y2, M, nh);
    H1.lo8 = H2.lo8 = chain_planes_lo8(c.prec) ? 1 : 0;
    const int n3pad = chain_n3pad(T1, s.Cout);
    Planes W1, W2, W3;
    if (p.pre[0]) {  // packed by rb_prepack_chain before the scale's first flow step
      W1 = p.pre[0]->w1; W2 = p.pre[0]->w2; W3 = p.pre[0]->w3;
    } else {
      W1 = planes_new(c, nh, kp); W2 = planes_new(c, nh, nh); W3 = planes_new(c, n3pad, nh);
      // conv(X, W1) | W2 + I (the skip of :125) | \nabla conv_data(., W3) tap-expanded
      op_pack_chain_tc(c, nh, s.nh, T1, Cin, kp, p.W1, p.W2, 0, s.Cout, n3pad, p.W3, W1, W2, W3);
    }
    ChainSpec cs{};
    cs.g = s.g; cs.B = s.B; cs.k1 = s.k1; cs.nh = nh; cs.nh_real = s.nh;
    cs.in = h.xin; cs.w1 = W1; cs.w2 = W2; cs.w3 = W3; cs.Cn = s.Cout;
    cs.mode = 0; cs.bias1 = p.b1; cs.bias2 = p.b2;
    if (h.G) { cs.o1 = H1; cs.o2 = H2; cs.bits1 = h.bm1; cs.bits2 = h.bm2; }
    cs.P = c.ar->f32((size_t)M * n3pad);
    cs.out0 = Y3; cs.out0_bs = (long long)s.Cout * px; cs.n0 = s.Cout;
    cs.add_n = 1 << 30;
    cs.fuse = h.fuse;
    op_rb_chain(c, cs);
    c.ar->release(m);
    return;
  }
  // the padded bf16 copy of the block input outlives this call (wgrad1 reads it again)
  h.xin = planes_new(c, M, cin_pad);
  op_nchw_to_tc(c, s.g, s.B, x2.p, x2.bs, s.c0, cond.p, cond.bs, Cin, cin_pad, h.xin);
  size_t m = c.ar->mark();
  Planes H1 = planes_at(h.Y1, M, nh), H2 = planes_at(h.Y2, M, nh);
  Planes W1, W2, W3;
  if (p.pre[0]) {  // packed once per pass (rb_prepack_unfused): a HINT coupling layer is visited many times
    W1 = p.pre[0]->w1; W2 = p.pre[0]->w2; W3 = p.pre[0]->w3;
  } else {
    W1 = planes_new(c, nh, T1 * cin_pad); W2 = planes_new(c, nh, T2 * nh); W3 = planes_new(c, cout_pad, T1 * nh);
    op_pack_w_tc(c, PACK_CONV, nh, Cin, T1, p.W1, nh, cin_pad, W1);
    op_pack_w_tc(c, PACK_CONV, nh, nh, T2, p.W2, nh, nh, W2, 1);  // + I: the skip of :125
    op_pack_w_tc(c, PACK_DATA, nh, s.Cout, T1, p.W3, cout_pad, nh, W3);
  }
  {  // X2 = relu(conv(X, W1) + b1)                           layer_residual_block.jl:122-123
    ConvTcSpec cs = tc_base(s);
    cs.k = s.k1; cs.in = h.xin; cs.cpad_in = cin_pad; cs.w = W1; cs.N = nh; cs.n_real = nh; cs.bias = p.b1;
    cs.mode = 0; cs.out = H1; cs.relu_encode = 1;
    op_conv_tc(c, cs);
  }
  {  // X3 = relu(X2 + conv(X2, W2) + b2)                      :125-126
    ConvTcSpec cs = tc_base(s);
    cs.k = s.k2; cs.in = H1; cs.cpad_in = nh; cs.w = W2; cs.N = nh; cs.n_real = nh; cs.bias = p.b2;
    cs.mode = 0; cs.out = H2; cs.relu_encode = 1;
    op_conv_tc(c, cs);
  }
  {  // Y3 = \nabla conv_data(X3, W3)                           :128-129
    ConvTcSpec cs = tc_base(s);
    cs.k = s.k1; cs.in = H2; cs.cpad_in = nh; cs.w = W3; cs.N = cout_pad; cs.n_real = s.Cout;
    cs.mode = 1; cs.out0 = Y3; cs.out0_bs = (long long)s.Cout * px; cs.n0 = s.Cout;
    op_conv_tc(c, cs);
  }
  c.ar->release(m);
}

static void rb_backward_tc(Ctx& c, const RBShape& s, const float* dY3, View x2, View cond, const RBParams& p,
                           RBHidden& h, const RBGrads& gr, View dx2, const float* add, long long add_bs,
                           View dcond) {
  const long long px = s.g.px, M = px * s.B;
  const int Cin = s.Cin(), T1 = s.T1(), T2 = s.T2(), Cout = s.Cout;
  const int cin_pad = pad16(Cin), cout_pad = pad16(Cout);
  const bool chain = rb_path(c, s) == RB_PATH_CHAIN;
  const int nh = chain ? chain_nh_pad(s.nh) : s.nh;
  size_t m = c.ar->mark();
  Planes H1 = planes_at(h.Y1, M, nh), H2 = planes_at(h.Y2, M, nh), G2 = planes_at(h.G, M, nh);
  if (chain) {
    // dY3 -> dY2 -> dY1 -> dX in one fused kernel (conv_tc_chain.cu); dY2 / dY1 go to HBM for the weight
    // gradients, which read every hidden tensor exactly once and produce the bias gradients on the way
    Planes G1 = planes_new(c, M, nh);
    H1.lo8 = H2.lo8 = G1.lo8 = G2.lo8 = chain_planes_lo8(c.prec) ? 1 : 0;  // as the storing passes wrote / write them
    const int kp = chain_kpad(T1, Cout, 0);
    const int n3pad = chain_n3pad(T1, Cin);
    Planes W3c, W2d, W1e;
    if (p.pre[1]) {
      W3c = p.pre[1]->w1; W2d = p.pre[1]->w2; W1e = p.pre[1]->w3;
    } else {
      W3c = planes_new(c, nh, kp); W2d = planes_new(c, nh, nh); W1e = planes_new(c, n3pad, nh);
    }
    Planes dcol = planes_new(c, M, kp);
    // INB_PREC_FP16X3: every tensor of this pass is linear in dY3, so ONE power-of-two scale (from max|dY3|, found on
    // the device) keeps the half-precision planes in their normal range; col2im and the gradient reduction divide by it
    uint32_t* smax = nullptr;
    if (prec_f16(c.prec)) {
      smax = h.dy_absmax;
      if (!smax) {
        smax = reinterpret_cast<uint32_t*>(c.ar->alloc_bytes(256));
        op_zero(c, smax, 4);
        op_absmax(c, px, s.B, Cout, dY3, (long long)Cout * px, smax);
      }
    }
    op_im2col_tc(c, s.g, s.B, s.k1, dY3, (long long)Cout * px, Cout, nullptr, 0, Cout, kp, -1, dcol, smax);
    // conv(dY3, W3) :151 | \nabla conv_data(., W2) + I (the '+ dY2' of :155) | \nabla conv_data(., W1) tap-expanded :162
    if (!p.pre[1]) op_pack_chain_tc(c, nh, s.nh, T1, Cout, kp, p.W3, p.W2, 1, Cin, n3pad, p.W1, W3c, W2d, W1e);
    ChainSpec cs{};
    cs.g = s.g; cs.B = s.B; cs.k1 = s.k1; cs.nh = nh; cs.nh_real = s.nh;
    cs.in = dcol; cs.w1 = W3c; cs.w2 = W2d; cs.w3 = W1e; cs.Cn = Cin;
    cs.mode = 1; cs.mask1 = h.bm2; cs.mask2 = h.bm1;                     // :154, :161
    cs.o1 = G2; cs.o2 = G1;
    cs.P = c.ar->f32((size_t)M * n3pad);
    cs.out0 = dx2.p; cs.out0_bs = dx2.bs; cs.n0 = s.c0;
    cs.out1 = dcond.p; cs.out1_bs = dcond.bs; cs.out1_accum = 1;
    cs.add = add; cs.add_bs = add_bs; cs.add_n = s.c0;
    cs.smax = smax;
    op_rb_chain(c, cs);
    Wgrad2TcSpec wg[3] = {
        Wgrad2TcSpec{M, H2, nh, dcol, Cout, T1, gr.W3, nullptr, smax, s.nh},   // :152
        Wgrad2TcSpec{M, G2, nh, H1, s.nh, 1, gr.W2, gr.b2, smax, s.nh},         // :156-157 (columns beyond n_hidden dropped)
        Wgrad2TcSpec{M, G1, nh, h.xin, Cin, T1, gr.W1, gr.b1, smax, s.nh}};     // :163-164
    wg[2].ones_col = h.xin_ones;
    op_wgrad2_tc_multi(c, wg, 3);
    c.ar->release(m);
    return;
  }
  Planes G1 = H2;  // dY1 reuses X3's storage once dW3 and the dgrad3 mask have consumed it
  Planes dY3p = planes_new(c, M, cout_pad);
  op_nchw_to_tc(c, s.g, s.B, dY3, (long long)Cout * px, Cout, nullptr, 0, Cout, cout_pad, dY3p);
  const int kmax = std::max(T1 * std::max(cout_pad, nh), T2 * nh);
  Planes Wp = planes_new(c, std::max(nh, cin_pad), kmax);
  auto wview = [&](int rows, int k) { return planes_at(Wp.hi, rows, k); };
  {  // dY2 = relugrad(conv(dY3, W3), Y2)                        layer_residual_block.jl:151,154
    Planes W = p.pre[1] ? p.pre[1]->w1 : wview(nh, T1 * cout_pad);
    if (!p.pre[1]) op_pack_w_tc(c, PACK_CONV, nh, Cout, T1, p.W3, nh, cout_pad, W);
    ConvTcSpec cs = tc_base(s);
    cs.k = s.k1; cs.in = dY3p; cs.cpad_in = cout_pad; cs.w = W; cs.N = nh; cs.n_real = nh;
    cs.mode = 0; cs.out = G2; cs.mask = H2;
    op_conv_tc(c, cs);
  }
  {  // dW3 = \nabla conv_filter(dY3, X3)                        :152
    WgradTcSpec ws{};
    ws.g = s.g; ws.B = s.B; ws.k = s.k1; ws.P = H2; ws.np = nh; ws.Q = dY3p; ws.cq = cout_pad; ws.cq_real = Cout;
    ws.dw = gr.W3;
    op_wgrad_tc(c, ws);
  }
  {  // dY1 = relugrad(\nabla conv_data(dY2, W2) + dY2, Y1)      :155,161
    Planes W = p.pre[1] ? p.pre[1]->w2 : wview(nh, T2 * nh);
    if (!p.pre[1]) op_pack_w_tc(c, PACK_DATA, nh, nh, T2, p.W2, nh, nh, W, 1);  // + I: the '+ dY2' of :155
    ConvTcSpec cs = tc_base(s);
    cs.k = s.k2; cs.in = G2; cs.cpad_in = nh; cs.w = W; cs.N = nh; cs.n_real = nh;
    cs.mode = 0; cs.out = G1; cs.mask = H1;
    op_conv_tc(c, cs);
  }
  {  // dW2 = \nabla conv_filter(X2, dY2); db2 = sum dY2         :156-157
    WgradTcSpec ws{};
    ws.g = s.g; ws.B = s.B; ws.k = s.k2; ws.P = G2; ws.np = nh; ws.Q = H1; ws.cq = nh; ws.cq_real = nh;
    ws.dw = gr.W2;
    op_wgrad_tc(c, ws);
    op_colsum_tc(c, M, nh, G2, gr.b2);
  }
  {  // dX1 = \nabla conv_data(dY1, W1) (+ passthrough)          :162
    Planes W = p.pre[1] ? p.pre[1]->w3 : wview(cin_pad, T1 * nh);
    if (!p.pre[1]) op_pack_w_tc(c, PACK_DATA, nh, Cin, T1, p.W1, cin_pad, nh, W);
    ConvTcSpec cs = tc_base(s);
    cs.k = s.k1; cs.in = G1; cs.cpad_in = nh; cs.w = W; cs.N = cin_pad; cs.n_real = Cin;
    cs.mode = 1; cs.out0 = dx2.p; cs.out0_bs = dx2.bs; cs.n0 = s.c0;
    cs.out1 = dcond.p; cs.out1_bs = dcond.bs; cs.out1_accum = 1;
    cs.add = add; cs.add_bs = add_bs; cs.add_n = s.c0;
    op_conv_tc(c, cs);
  }
  {  // dW1 = \nabla conv_filter(X1, dY1); db1 = sum dY1         :163-164
    WgradTcSpec ws{};
    ws.g = s.g; ws.B = s.B; ws.k = s.k1; ws.P = G1; ws.np = nh; ws.Q = h.xin; ws.cq = cin_pad; ws.cq_real = Cin;
    ws.dw = gr.W1;
    op_wgrad_tc(c, ws);
    op_colsum_tc(c, M, nh, G1, gr.b1);
  }
  (void)x2; (void)cond;
  c.ar->release(m);
}

void rb_forward(Ctx& c, const RBShape& s, View x2, View cond, const RBParams& p, RBHidden& h, float* Y3) {
  if (rb_path(c, s) == RB_PATH_SIMT) rb_forward_fp32(c, s, x2, cond, p, h, Y3);
  else rb_forward_tc(c, s, x2, cond, p, h, Y3);
}
void rb_backward(Ctx& c, const RBShape& s, const float* dY3, View x2, View cond, const RBParams& p,
                 RBHidden& h, const RBGrads& gr, View dx2, const float* add, long long add_bs, View dcond) {
  if (rb_path(c, s) == RB_PATH_SIMT) rb_backward_fp32(c, s, dY3, x2, cond, p, h, gr, dx2, add, add_bs, dcond);
  else rb_backward_tc(c, s, dY3, x2, cond, p, h, gr, dx2, add, add_bs, dcond);
}

// ---------------------------------------------------------------- flow step
void flow_forward(Ctx& c, const FlowShape& f, View x, View y, View cond, const FlowParams& p, double* ld) {
  const long long px = f.g.px;
  const int C1 = f.C1();
  const RBShape rs = f.rb();
  size_t m = c.ar->mark();
  // Y = ActNorm(X) (actnorm.jl:73), X_ = C.forward(Y) (glow.jl:105)
  op_an_hh_fwd(c, px, f.B, f.C, x, y, p.s, p.b, p.v1, p.v2, p.v3, f.logdet ? ld : nullptr);
  RBHidden h;
  h.Y1 = c.ar->f32(rb_hidden_elems(rs));
  h.Y2 = c.ar->f32(rb_hidden_elems(rs));
  h.G = nullptr;
  float* Y3 = c.ar->f32((size_t)f.B * rs.Cout * px);
  CouplingFuse cf{0, C1, y.p, y.bs, nullptr, 0, nullptr, f.low, f.high, 1.f / (float)f.B, f.logdet ? ld : nullptr, nullptr, false};
  h.fuse = &cf;
  rb_forward(c, rs, sub(y, C1, px), cond, p.rb, h, Y3);  // glow.jl:109, X2 = second part
  if (!cf.done) op_coupling_fwd(c, px, f.B, C1, y, y, Y3, f.low, f.high, f.logdet ? ld : nullptr);  // :110-116
  c.ar->release(m);
}

void flow_inverse(Ctx& c, const FlowShape& f, View y, View x, View cond, const FlowParams& p) {
  const long long px = f.g.px;
  const int C1 = f.C1();
  const RBShape rs = f.rb();
  size_t m = c.ar->mark();
  RBHidden h;
  h.Y1 = c.ar->f32(rb_hidden_elems(rs));
  h.Y2 = c.ar->f32(rb_hidden_elems(rs));
  h.G = nullptr;
  float* Y3 = c.ar->f32((size_t)f.B * rs.Cout * px);
  CouplingFuse cf{1, C1, y.p, y.bs, nullptr, 0, nullptr, f.low, f.high, 0.f, nullptr, nullptr, false};
  h.fuse = &cf;
  rb_forward(c, rs, sub(y, C1, px), cond, p.rb, h, Y3);             // glow.jl:124
  if (!cf.done) op_coupling_inv(c, px, f.B, C1, y, y, Y3, f.low, f.high);        // :127
  op_hh_an_inv(c, px, f.B, f.C, y, x, p.s, p.b, p.v1, p.v2, p.v3); // :130, actnorm.jl:93
  c.ar->release(m);
}

void flow_backward(Ctx& c, const FlowShape& f, View dy, View y, View dx, View x, View cond, View dcond,
                   const FlowParams& p, const FlowGrads& g) {
  const long long px = f.g.px;
  const int C1 = f.C1(), C = f.C;
  const RBShape rs = f.rb();
  size_t m = c.ar->mark();
  RBHidden h;
  h.Y1 = c.ar->f32(rb_hidden_elems(rs));
  h.Y2 = c.ar->f32(rb_hidden_elems(rs));
  h.G = c.ar->f32(rb_hidden_elems(rs));
  float* Y3 = c.ar->f32((size_t)f.B * rs.Cout * px);
  // Gram matrix + ActNorm sums: from the side lane's pool when there is one (their consumers run on the lane)
  double* gram = nullptr;
  bool on_lane = false;
  if (c.lane && !c.dry()) {
    gram = reinterpret_cast<double*>(c.lane->take(((size_t)C * C + 2 * C) * sizeof(double)));
    on_lane = gram != nullptr;
  }
  if (!gram) gram = c.ar->f64((size_t)C * C + 2 * C);
  double* dsdb = gram + (size_t)C * C;
  op_zero(c, gram, ((size_t)C * C + 2 * C) * sizeof(double));
  View y2 = sub(y, C1, px), dy2 = sub(dy, C1, px);
  // recompute the block once (glow.jl:139 -> :124; the reference recomputes it again at
  // layer_residual_block.jl:143 - same values)
  if (prec_f16(c.prec)) {  // max|dY3| falls out of the kernel that writes dY3
    h.dy_absmax = reinterpret_cast<uint32_t*>(c.ar->alloc_bytes(256));
    op_zero(c, h.dy_absmax, 4);
  }
  // X1, dX1, and the masked gradient of the block output        glow.jl:127,142-151 - folded into the block's last
  // kernel on the fused tensor-core chain, a kernel of its own otherwise
  CouplingFuse cf{2, C1, y.p, y.bs, dy.p, dy.bs, Y3, f.low, f.high, f.logdet ? 1.f / (float)f.B : 0.f, nullptr,
                  h.dy_absmax, false};
  h.fuse = &cf;
  rb_forward(c, rs, y2, cond, p.rb, h, Y3);
  h.fuse = nullptr;
  if (!cf.done) op_coupling_bwd(c, px, f.B, C1, y, y, dy, dy, Y3, f.low, f.high, f.logdet, h.dy_absmax);
  // dX2 = RB.backward(...) + dY2                                 glow.jl:151
  rb_backward(c, rs, Y3, y2, cond, p.rb, h, g.rb, dy2, dy2.p, dy2.bs, dcond);
  // Conv1x1 inverse on (dX_, X_) + ActNorm backward              glow.jl:159, actnorm.jl:100-123
  op_hh_an_bwd(c, px, f.B, C, dy, y, dx, x, p.s, p.b, p.v1, p.v2, p.v3, gram, p.s ? dsdb : nullptr);
  Ctx fc = c;
  if (on_lane) {
    c.lane->fork(c.st);
    fc.st = c.lane->st;
  }
  if (p.s) op_hh_an_grad_finish(fc, C, px, gram, p.v1, p.v2, p.v3, f.freeze, g.v1, g.v2, g.v3, dsdb, p.s, f.logdet, g.s, g.b);
  else op_hh_grad_finish(fc, C, gram, p.v1, p.v2, p.v3, f.freeze, g.v1, g.v2, g.v3);
  c.ar->release(m);
}

}  // namespace inb
