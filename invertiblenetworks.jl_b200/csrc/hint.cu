// hint.cu - the HINT family on the kernels of the Glow path: CouplingLayerBasic (invertible_layer_basic.jl:90-149)
// and the recursive CouplingLayerHINT (invertible_layer_hint.jl:105-297).
//
// The reference recursion splits and concatenates tensors at every level; here every level works on channel-range
// views of ONE tensor in place.  That needs a different execution order inside a level than the reference's
// (whose statements are independent where reordered):
//   forward   Xb <- HINT(Xb) ; Xb <- CL[s](Xa, Xb) ; Xa <- HINT(Xa)     (hint.jl:131-133: the coupling reads the
//                                                                        UNTRANSFORMED Xa, so Xa goes last)
//   inverse   Ya <- HINT^-1(Ya) ; Yb <- CL[s]^-1(Xa, Yb) ; Yb <- HINT^-1(Yb)                       (:172-178)
//   backward  (dYa,Ya) <- HINT.bwd ; CL[s].bwd adds into dXa and maps (dYb,Yb) ; (dYb,Yb) <- HINT.bwd  (:230-232,248)
// The affine coupling itself is the Glow coupling with the roles of the halves exchanged (the FIRST half conditions,
// the SECOND is transformed), so the same three kernels serve (op_coupling_fwd / inv / bwd).
#include "glow.cuh"

namespace inb {

int HintShape::depth() const {
  int count = 0;
  double nc = C;
  while (nc > 4) { nc /= 2; ++count; }
  return count + 1;
}

void hint_check(const HintShape& h) {
  INB_CHECK(h.C >= 2, "CouplingLayerHINT needs at least 2 channels (got %d)", h.C);
  INB_CHECK(h.permute >= 0 && h.permute <= 3, "permute must be none (0), full (1), lower (2) or both (3)");
  // Int(n_in/2^j) must be exact (hint.jl:86) and every level must split into equal halves
  const int n = h.depth();
  INB_CHECK(h.C % (1 << n) == 0 || (h.C <= 4 && h.C % 2 == 0),
            "CouplingLayerHINT: %d channels cannot be halved %d times (InexactError in the reference)", h.C, n);
}

static RBShape basic_rb(const HintShape& h, int Ca) {
  return RBShape{h.g, h.B, Ca, 0, h.nh, 2 * Ca, h.k1, h.k2};
}

// CL.forward(X1 = xa, X2 = xb): xb <- S .* xb + T                        basic.jl:96-98
void basic_forward(Ctx& c, const HintShape& h, int Ca, View xa, View xb, const RBParams& p, double* ld, int ld_batch) {
  const RBShape rs = basic_rb(h, Ca);
  size_t m = c.ar->mark();
  RBHidden hid;
  hid.Y1 = c.ar->f32(rb_hidden_elems(rs));
  hid.Y2 = c.ar->f32(rb_hidden_elems(rs));
  hid.G = nullptr;
  float* Y3 = c.ar->f32((size_t)h.B * rs.Cout * h.g.px);
  rb_forward(c, rs, xa, view(nullptr, 0), p, hid, Y3);
  op_coupling_fwd(c, h.g.px, h.B, Ca, xb, xb, Y3, h.low, h.high, h.logdet ? ld : nullptr, ld_batch);
  c.ar->release(m);
}
// CL.inverse(Y1 = xa, Y2 = yb): yb <- (yb - T) ./ (S + eps)               basic.jl:112-114
void basic_inverse(Ctx& c, const HintShape& h, int Ca, View xa, View yb, const RBParams& p) {
  const RBShape rs = basic_rb(h, Ca);
  size_t m = c.ar->mark();
  RBHidden hid;
  hid.Y1 = c.ar->f32(rb_hidden_elems(rs));
  hid.Y2 = c.ar->f32(rb_hidden_elems(rs));
  hid.G = nullptr;
  float* Y3 = c.ar->f32((size_t)h.B * rs.Cout * h.g.px);
  rb_forward(c, rs, xa, view(nullptr, 0), p, hid, Y3);
  op_coupling_inv(c, h.g.px, h.B, Ca, yb, yb, Y3, h.low, h.high);
  c.ar->release(m);
}
// CL.backward(dY1 = 0, dY2 = dyb, Y1 = xa, Y2 = yb) with its dX1 ADDED to dxa (hint.jl:231,248 / :253,264):
// (dyb, yb) <- (dX2, X2) in place                                          basic.jl:127-137
void basic_backward(Ctx& c, const HintShape& h, int Ca, View xa, View dxa, View yb, View dyb,
                           const RBParams& p, const RBGrads& g, bool accumulate) {
  const RBShape rs = basic_rb(h, Ca);
  size_t m = c.ar->mark();
  RBHidden hid;
  hid.Y1 = c.ar->f32(rb_hidden_elems(rs));
  hid.Y2 = c.ar->f32(rb_hidden_elems(rs));
  hid.G = c.ar->f32(rb_hidden_elems(rs));
  float* Y3 = c.ar->f32((size_t)h.B * rs.Cout * h.g.px);
  const int T1 = rs.T1(), T2 = rs.T2();
  const long long n1 = (long long)h.nh * Ca * T1, n2 = (long long)h.nh * h.nh * T2, n3 = (long long)h.nh * 2 * Ca * T1;
  RBGrads gw = g;
  if (accumulate) {  // a later visit of a shared layer: gradients into scratch, then added
    gw.W1 = c.ar->f32(n1); gw.W2 = c.ar->f32(n2); gw.W3 = c.ar->f32(n3);
    gw.b1 = c.ar->f32(h.nh); gw.b2 = c.ar->f32(h.nh);
  }
  rb_forward(c, rs, xa, view(nullptr, 0), p, hid, Y3);  // recompute (basic.jl:127 -> :112)
  op_coupling_bwd(c, h.g.px, h.B, Ca, yb, yb, dyb, dyb, Y3, h.low, h.high, h.logdet);  // :114,130-135
  rb_backward(c, rs, Y3, xa, view(nullptr, 0), p, hid, gw, dxa, dxa.p, dxa.bs, view(nullptr, 0));  // :137
  if (accumulate) {
    const float* src[5] = {gw.W1, gw.W2, gw.W3, gw.b1, gw.b2};
    float* dst[5] = {g.W1, g.W2, g.W3, g.b1, g.b2};
    const long long n[5] = {n1, n2, n3, h.nh, h.nh};
    op_accum(c, 5, src, dst, n);
  }
  c.ar->release(m);
}

// A CL[j] is applied 2^(j-1) times per pass with the same weights: its chain operands are packed once per pass
// (direction 0: forward / recompute, 1: backward) instead of once per visit.  Returns a copy of `p` whose RBParams
// point at the packed planes (allocated in the caller's arena scope, `store` keeps the descriptors alive).
static HintParams hint_prepack(Ctx& c, const HintShape& h, const HintParams& p, bool backward, std::vector<PackedW>& store) {
  HintParams q = p;
  const int n = (int)p.cl.size();
  store.resize(2 * n);
  for (int j = 0; j < n; ++j) {
    const RBShape rs = basic_rb(h, h.C >> (j + 1));
    if (rb_prepack_chain(c, rs, &p.cl[j], 1, 0, &store[2 * j]) || rb_prepack_unfused(c, rs, p.cl[j], 0, &store[2 * j]))
      q.cl[j].pre[0] = &store[2 * j];
    if (backward && (rb_prepack_chain(c, rs, &p.cl[j], 1, 1, &store[2 * j + 1]) ||
                     rb_prepack_unfused(c, rs, p.cl[j], 1, &store[2 * j + 1])))
      q.cl[j].pre[1] = &store[2 * j + 1];
  }
  return q;
}

static void rec_forward(Ctx& c, const HintShape& h, View x, int C, int scale, const HintParams& p, double* ld) {
  const int Ca = C / 2;
  View xa = x, xb = sub(x, Ca, h.g.px);
  INB_CHECK(scale < (int)p.cl.size() + 1, "internal: HINT recursion deeper than the coupling layers");
  if (C > 4) {
    rec_forward(c, h, xb, C - Ca, scale + 1, p, ld);
    basic_forward(c, h, Ca, xa, xb, p.cl[scale - 1], ld);
    rec_forward(c, h, xa, Ca, scale + 1, p, ld);
  } else {
    basic_forward(c, h, Ca, xa, xb, p.cl[scale - 1], ld);
  }
}
static void rec_inverse(Ctx& c, const HintShape& h, View y, int C, int scale, const HintParams& p) {
  const int Ca = C / 2;
  View ya = y, yb = sub(y, Ca, h.g.px);
  if (C > 4) {
    rec_inverse(c, h, ya, Ca, scale + 1, p);
    basic_inverse(c, h, Ca, ya, yb, p.cl[scale - 1]);
    rec_inverse(c, h, yb, C - Ca, scale + 1, p);
  } else {
    basic_inverse(c, h, Ca, ya, yb, p.cl[scale - 1]);
  }
}
static void rec_backward(Ctx& c, const HintShape& h, View dy, View y, int C, int scale, const HintParams& p,
                         const HintGrads& g, std::vector<char>& seen) {
  const int Ca = C / 2;
  View ya = y, yb = sub(y, Ca, h.g.px), dya = dy, dyb = sub(dy, Ca, h.g.px);
  auto coupling = [&] {
    const bool acc = seen[scale - 1] && !h.shared_last;
    basic_backward(c, h, Ca, ya, dya, yb, dyb, p.cl[scale - 1], g.cl[scale - 1], acc);
    seen[scale - 1] = 1;
  };
  if (C > 4) {
    rec_backward(c, h, dya, ya, Ca, scale + 1, p, g, seen);
    coupling();
    rec_backward(c, h, dyb, yb, C - Ca, scale + 1, p, g, seen);
  } else {
    coupling();
  }
}

void hint_forward(Ctx& c, const HintShape& h, View x, View y, const HintParams& p, double* ld) {
  const long long px = h.g.px;
  const bool both = h.permute == HINT_PERMUTE_BOTH;
  const bool full = h.permute == HINT_PERMUTE_FULL || both;
  // [ActNorm (actnorm.jl:73)] + C.forward (hint.jl:111-113) in one pass; a plain copy when neither applies
  if (p.s || full)
    op_an_hh_fwd(c, px, h.B, h.C, x, y, p.s, p.b, full ? p.v1 : nullptr, full ? p.v2 : nullptr, full ? p.v3 : nullptr,
                 (p.s && h.logdet) ? ld : nullptr);
  else
    op_copy(c, px, h.B, h.C, x, y);
  if (h.permute == HINT_PERMUTE_LOWER) {  // Xb = C.forward(Xb), :115 (in place: a per-pixel map)
    const int Ca = h.C / 2;
    View yb = sub(y, Ca, px);
    op_an_hh_fwd(c, px, h.B, h.C - Ca, yb, yb, nullptr, nullptr, p.v1, p.v2, p.v3, nullptr);
  }
  size_t m = c.ar->mark();
  std::vector<PackedW> store;
  const HintParams q = hint_prepack(c, h, p, false, store);
  static const bool no_levels = [] { const char* e = getenv("INB_HINT_LEVELS"); return e && e[0] == '0'; }();
  const int n = h.depth();
  if (n > 1 && !no_levels && y.bs == (long long)h.C * px && (long long)h.B * (1 << (n - 1)) <= 65535) {
    // Level-synchronous forward.  In the forward direction every coupling conditions on the UNTRANSFORMED first half
    // of its group (hint.jl:131-133 pass the input slices down), so all 2^(s-1) groups of level s are independent
    // given the levels below, and they share CL[s].  The groups tile the channel axis, hence (sample, group) is one
    // uniform batch stride of 2*Ca*px elements: a level is ONE coupling call on B * 2^(s-1) virtual samples, reading
    // the conditioning halves from a pristine copy and transforming the second halves of the working tensor in place
    // (n calls instead of 2^n - 1).  The inverse and backward directions are inherently depth-first (rec_*).
    float* y0 = c.ar->f32((size_t)h.B * h.C * px);
    op_copy(c, px, h.B, h.C, y, view(y0, (long long)h.C * px));
    for (int s = n; s >= 1; --s) {
      const int Ca = h.C >> s, G = 1 << (s - 1);
      HintShape hv = h;
      hv.B = h.B * G;
      basic_forward(c, hv, Ca, view(y0, 2LL * Ca * px), view(y.p + (long long)Ca * px, 2LL * Ca * px), q.cl[s - 1], ld, h.B);
    }
  } else {
    rec_forward(c, h, y, h.C, 1, q, ld);
  }
  c.ar->release(m);
  if (both) op_hh_an_inv(c, px, h.B, h.C, y, y, nullptr, nullptr, p.v1, p.v2, p.v3);  // Y = C.inverse(Y), :149
}

void hint_inverse(Ctx& c, const HintShape& h, View y, View x, const HintParams& p) {
  const long long px = h.g.px;
  const bool both = h.permute == HINT_PERMUTE_BOTH;
  const bool full = h.permute == HINT_PERMUTE_FULL || both;
  if (both) op_an_hh_fwd(c, px, h.B, h.C, y, y, nullptr, nullptr, p.v1, p.v2, p.v3, nullptr);  // Y = C.forward(Y), :164
  {
    size_t m = c.ar->mark();
    std::vector<PackedW> store;
    const HintParams q = hint_prepack(c, h, p, false, store);
    rec_inverse(c, h, y, h.C, 1, q);
    c.ar->release(m);
  }
  if (h.permute == HINT_PERMUTE_LOWER) {  // :193
    const int Ca = h.C / 2;
    View yb = sub(y, Ca, px);
    op_hh_an_inv(c, px, h.B, h.C - Ca, yb, yb, nullptr, nullptr, p.v1, p.v2, p.v3);
  }
  if (p.s || full)  // C.inverse (:195-197) [+ ActNorm.inverse]
    op_hh_an_inv(c, px, h.B, h.C, y, x, p.s, p.b, full ? p.v1 : nullptr, full ? p.v2 : nullptr, full ? p.v3 : nullptr);
  else if (x.p != y.p)
    op_copy(c, px, h.B, h.C, y, x);
}

void hint_backward(Ctx& c, const HintShape& h, View dy, View y, View dx, View x, const HintParams& p,
                   const HintGrads& g) {
  const long long px = h.g.px;
  const bool both = h.permute == HINT_PERMUTE_BOTH;
  const bool full = h.permute == HINT_PERMUTE_FULL || both, lower = h.permute == HINT_PERMUTE_LOWER;
  size_t m = c.ar->mark();
  float* tv = nullptr;  // gradient of C through its second use (Y = C.inverse(Y) at the end of forward), as (v3, v2, v1)
  if (both) {
    // dY, Y = C.forward((dY, Y)), hint.jl:219-221.  The map undone here is Z = Y * H3 H2 H1 - a Conv1x1 whose vectors are
    // (v3, v2, v1) - so its (dZ, Z) -> (dY, Y) step and its Householder gradients are the inverse-tuple kernels with the
    // vectors in reverse order (equal to conv1x1_grad_v(.; adjoint = true), conv1x1.jl:196, asserted against autograd
    // in tests/test_oracle_hint.py); the two contributions to C's gradients are summed below (conv1x1.jl:198-200).
    double* gram2 = c.ar->f64((size_t)h.C * h.C);
    tv = c.ar->f32(3 * (size_t)h.C);
    op_zero(c, gram2, (size_t)h.C * h.C * sizeof(double));
    op_hh_an_bwd(c, px, h.B, h.C, dy, y, dy, y, nullptr, nullptr, p.v3, p.v2, p.v1, gram2, nullptr);
    op_hh_grad_finish(c, h.C, gram2, p.v3, p.v2, p.v1, 0, tv, tv + h.C, tv + 2 * h.C);
  }
  std::vector<char> seen(p.cl.size(), 0);
  std::vector<PackedW> store;
  const HintParams q = hint_prepack(c, h, p, true, store);
  rec_backward(c, h, dy, y, h.C, 1, q, g, seen);
  const int Cm = lower ? h.C - h.C / 2 : h.C;  // channels the Householder mix acts on
  double* gram = c.ar->f64((size_t)Cm * Cm + 2 * h.C);
  double* dsdb = gram + (size_t)Cm * Cm;
  op_zero(c, gram, ((size_t)Cm * Cm + 2 * h.C) * sizeof(double));
  if (lower) {  // dXb, Xb = C.inverse((dXb, Xb)), :268
    View yb = sub(y, h.C / 2, px), dyb = sub(dy, h.C / 2, px);
    op_hh_an_bwd(c, px, h.B, Cm, dyb, yb, dyb, yb, nullptr, nullptr, p.v1, p.v2, p.v3, gram, nullptr);
    op_hh_grad_finish(c, Cm, gram, p.v1, p.v2, p.v3, 0, g.v1, g.v2, g.v3);
  }
  if (p.s || full) {  // C.inverse((dX, X)) :278 [+ ActNorm.backward, actnorm.jl:100-123]
    op_hh_an_bwd(c, px, h.B, h.C, dy, y, dx, x, p.s, p.b, full ? p.v1 : nullptr, full ? p.v2 : nullptr,
                 full ? p.v3 : nullptr, full ? gram : nullptr, p.s ? dsdb : nullptr);
    if (full) op_hh_grad_finish(c, h.C, gram, p.v1, p.v2, p.v3, 0, g.v1, g.v2, g.v3);
    if (both) {
      const float* src[3] = {tv + 2 * h.C, tv + h.C, tv};
      float* dst[3] = {g.v1, g.v2, g.v3};
      const long long n[3] = {h.C, h.C, h.C};
      op_accum(c, 3, src, dst, n);
    }
    if (p.s) op_an_grad_finish(c, h.C, px, dsdb, p.s, h.logdet, g.s, g.b);
  } else {
    if (x.p != y.p) op_copy(c, px, h.B, h.C, y, x);
    if (dx.p != dy.p) op_copy(c, px, h.B, h.C, dy, dx);
  }
  c.ar->release(m);
}

}  // namespace inb
