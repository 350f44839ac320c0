/* inb200.h - C ABI of libinb200.so: the B200 (sm_100a) implementation of the Glow training
 * hot path of slimgroup/InvertibleNetworks.jl (reference v2.3.1).
 *
 * This header is the drop-in boundary.  The reference has no FFI of its
 * own: its "plugin API" is Julia multiple dispatch (`forward(X, G)`, `inverse`, `backward`,
 * src/utils/neuralnet.jl:20-41).  A Julia shim adds methods specialised on CuArray{Float32}
 * that `ccall` the symbols below (julia/InvertibleNetworksB200.jl); in this
 * repository the same symbols are exercised from Python ctypes.
 *
 * Conventions
 *  - plain C types only: device pointers are `float*`, sizes are int / long long.
 *  - every tensor is float32 in the reference's memory layout: a Julia array (nx,ny[,nz],C,B)
 *    in column-major order == C order (B,C[,nz],ny,nx); x fastest, channel stride = nx*ny*nz.
 *  - conv weights are the reference's arrays W1 (k1..,Cin,nh), W2 (k2..,nh,nh), W3 (k1..,Cout,nh)
 *    (src/layers/layer_residual_block.jl:92-94), unmodified bytes.
 *  - every entry point returns 0 on success, non-zero on error; inb_last_error() gives the text
 *    (thread local).  No C++ exception crosses this boundary, the library never calls exit().
 *  - every entry point takes the CUDA stream to enqueue on (a cudaStream_t passed as void*);
 *    nothing here synchronises the device except where stated.  The caller owns every buffer it
 *    passes; the library only owns the per-plan workspace.
 *  - `params` / `grads` are host arrays of device pointers in the reference's get_params order
 *    (src/utils/neuralnet.jl:72-88): NetworkGlow: AN[i,j].(s,b) for i=1..L, j=1..K (i outer),
 *    then CL[i,j].(v1,v2,v3,W1,W2,W3,b1,b2); NetworkConditionalGlow: AN, then AN_C.(s,b), then CL.
 *  - gradients are WRITTEN (not accumulated); the reference's "+=" rule for Conv1x1 grads
 *    (invertible_layer_conv1x1.jl:237-239) and `nothing` handling live in the host shim.
 */
#ifndef INB200_H
#define INB200_H

#ifdef __cplusplus
extern "C" {
#endif

#define INB_OK 0
#define INB_ERR_INVALID 1      /* bad argument / unsupported configuration */
#define INB_ERR_CUDA 2         /* a CUDA runtime call failed */
#define INB_ERR_NOMEM 3        /* workspace allocation failed */

/* arithmetic used by the ResidualBlock contractions (everything else is always fp32) */
#define INB_PREC_FP32 0        /* fp32 FMA on CUDA cores: bit-for-bit fp32 arithmetic        */
#define INB_PREC_BF16X3 1      /* tcgen05 bf16 tensor cores, 3-term split (fp32-equivalent)  */
#define INB_PREC_BF16 2        /* tcgen05 bf16 tensor cores, single pass, fp32 accumulate    */
#define INB_PREC_FP16X3 3      /* tcgen05 kind::f16 with IEEE-half operands, 3-term split: 2 x 11 significant bits per
                                  operand (float32-level products) at the bf16x3 rate; weights and gradients are
                                  pre-scaled by exact powers of two (the gradient scale is derived on the device from
                                  max|dY| of each ResidualBlock backward).  Runs on the fused chain (k2 = 1, any n_hidden
                                  <= 256, zero padded to 128 / 256); blocks with k2 = 3 compute in bf16x3.  The hidden
                                  tensors handed to the weight gradients keep one byte of their lo half (14 significant
                                  bits per stored value; the weight gradients stay at 4-7e-6 of the float64 result). */

typedef struct inb_plan inb_plan;

/* Mirrors the constructor arguments of NetworkGlow / NetworkConditionalGlow
 * (invertible_network_glow.jl:78, invertible_network_conditional_glow.jl:78). */
typedef struct inb_glow_desc {
  int ndims;          /* 2 or 3 spatial dimensions (NetworkGlow3D: invertible_network_glow.jl:106) */
  int nx, ny, nz;     /* input spatial size, Julia order (nx fastest); nz = 1 when ndims == 2 */
  int n_in;           /* input channels */
  int n_cond;         /* condition channels; 0 = NetworkGlow, > 0 = NetworkConditionalGlow */
  int n_hidden;       /* ResidualBlock hidden channels */
  int L, K;           /* scales, flow steps per scale */
  int batch;          /* batch size the workspace is sized for (smaller batches are accepted) */
  int split_scales;   /* squeeze + split per scale (forced on when n_in == 1, :79) */
  int logdet;         /* compute logdet and bake its gradient into backward (:116, actnorm :108) */
  int k1, k2, p1, p2; /* ResidualBlock kernel sizes / paddings; supported: k1 in {1,3}, k2 in {1,3}, "same" padding, stride 1 */
  float sig_low, sig_high; /* SigmoidLayer(low, high) (activation_functions.jl:30-35); default 0,1 */
  int freeze_conv;    /* Conv1x1 freeze: zero Householder grads (invertible_layer_conv1x1.jl:132-134) */
  int precision;      /* INB_PREC_* */
} inb_glow_desc;

const char* inb_last_error(void);
int inb_version(void);
/* 1 when a CUDA device of compute capability 10.x is present */
int inb_device_ok(void);

/* ------------------------------------------------------------------ network level */
int inb_glow_plan_create(const inb_glow_desc* desc, inb_plan** plan);
int inb_glow_plan_destroy(inb_plan* plan);
/* number of parameter tensors (10*L*K, +2 when conditional) and the element count of each */
int inb_glow_num_params(const inb_plan* plan);
int inb_glow_param_numel(const inb_plan* plan, int index, long long* numel);
/* workspace bytes held by the plan */
long long inb_glow_workspace_bytes(const inb_plan* plan);
/* Z_dims bookkeeping the reference mutates in forward (invertible_network_glow.jl:123):
 * for scale i (0-based) writes (B,C,[nz,]ny,nx) of the latent split off there; returns count */
int inb_glow_zdims(const inb_plan* plan, int batch, int scale, int* dims5);

/* replaces forward(X, G::NetworkGlow), invertible_network_glow.jl:109-129.
 * X: (batch, n_in, spatial); Z: flat latent of X's element count ([vec(Z_1);..;vec(X_L)] when
 * split_scales, else X's shape); logdet: device pointer to ONE float (may be NULL when
 * desc.logdet == 0).  init_actnorm != 0 performs the data-dependent ActNorm initialisation
 * (invertible_layer_actnorm.jl:67-72) layer by layer, writing s and b into `params`. */
int inb_glow_forward(inb_plan* plan, int batch, const float* X, float* const* params, float* Z,
                     float* logdet, int init_actnorm, void* stream);
/* replaces inverse(Z, G::NetworkGlow), invertible_network_glow.jl:132-147 */
int inb_glow_inverse(inb_plan* plan, int batch, const float* Z, float* const* params, float* X,
                     void* stream);
/* replaces backward(dZ, Z, G::NetworkGlow) with set_grad=true, invertible_network_glow.jl:150-191.
 * Writes dX, X (recomputed by inversion) and every gradient in `grads` (get_params order). */
int inb_glow_backward(inb_plan* plan, int batch, const float* dZ, const float* Z,
                      float* const* params, float* const* grads, float* dX, float* X, void* stream);

/* ------------------------------------------------------------------ data-parallel training
 * One process per GPU, the batch sharded along its outermost dimension, parameters replicated.  The reference has no
 * multi-GPU path of its own; these entry points are what a Julia caller binds next to NCCL.jl / MPI.jl.
 * NCCL is resolved at run time (dlopen of libnccl.so.2, preferring the instance already loaded by the host framework;
 * INB200_NCCL_LIB overrides); a process that never calls these needs no NCCL.
 *
 * Canonical flat layout: offsets[i] = element offset of parameter i (get_params order) in ONE buffer of `total` floats,
 * every tensor on a 256-byte boundary.  A caller that keeps its gradients (and parameters) in such a buffer gets one
 * collective per contiguous range instead of one per tensor. */
typedef struct inb_comm inb_comm;
int inb_glow_flat_layout(const inb_plan* plan, long long* offsets /* [num_params], nullable */, long long* total);
/* rank 0 calls inb_comm_unique_id and ships the 128 bytes to the other ranks by any means (MPI.bcast, a file, torch's
 * store); every rank then calls inb_comm_create on its own device (ncclCommInitRank). */
int inb_comm_unique_id(char id[128]);
int inb_comm_create(int nranks, int rank, const char id[128], inb_comm** comm);
/* wrap an existing ncclComm_t (NCCL.jl: Communicator.handle; torch: ProcessGroupNCCL._comm_ptr()); not destroyed with the handle */
int inb_comm_wrap(void* nccl_comm, inb_comm** comm);
/* detach the communicator from every plan first (inb_glow_plan_set_comm(plan, NULL)): a plan's captured CUDA graphs hold
 * the communicator's collectives, and NCCL requires them to be destroyed before ncclCommDestroy */
int inb_comm_destroy(inb_comm* comm);
int inb_comm_info(const inb_comm* comm, int* nranks, int* rank, long long* allreduce_calls, long long* allreduce_bytes);
/* Attach a communicator to a plan (NULL detaches).  With a communicator attached
 *  - inb_glow_forward / inb_cglow_forward with init_actnorm != 0 initialise every ActNorm from the statistics of the
 *    GLOBAL batch (per layer: all-reduce of the per-shard sums, then of the squared deviations from the global mean =
 *    invertible_layer_actnorm.jl:67-72 on the concatenated shards), so every rank ends with identical s, b;
 *  - inb_glow_backward / inb_cglow_backward average the gradients over the ranks: as soon as the flow steps of a scale
 *    are done its 10*K tensors are all-reduced (ncclAvg) on the communicator's own stream, overlapped with the
 *    remaining scales; the call's stream waits for the last collective before it returns control to later work.
 *    Averaging the per-shard gradients of f = ||Z||^2/(2B) - logdet reproduces the single-process gradient of the
 *    global batch (the only batch-independent term, ActNorm's prod(spatial)*sum(log|s|), is invariant under averaging). */
int inb_glow_plan_set_comm(inb_plan* plan, inb_comm* comm);
/* the same average as one explicit call (all tensors of `grads`, on `stream`) for callers that do not attach */
int inb_allreduce_grads(inb_plan* plan, float* const* grads, inb_comm* comm, void* stream);
/* every rank receives rank `root`'s parameters */
int inb_broadcast_params(inb_plan* plan, float* const* params, inb_comm* comm, int root, void* stream);

/* NetworkConditionalGlow (invertible_network_conditional_glow.jl:107-181).  forward returns
 * ZX (X's shape), ZC (the ActNorm'ed condition squeezed L times when split_scales) and logdet;
 * inverse / backward consume that ZC. */
int inb_cglow_forward(inb_plan* plan, int batch, const float* X, const float* C,
                      float* const* params, float* ZX, float* ZC, float* logdet, int init_actnorm,
                      void* stream);
int inb_cglow_inverse(inb_plan* plan, int batch, const float* ZX, const float* ZC,
                      float* const* params, float* X, void* stream);
int inb_cglow_backward(inb_plan* plan, int batch, const float* dZX, const float* ZX,
                       const float* ZC, float* const* params, float* const* grads, float* dX,
                       float* X, float* dC, void* stream);

/* ------------------------------------------------------------------ layer level
 * All tensors (B, C, spatial) contiguous.  `sp` = nx*ny*nz. */

/* ActNorm, invertible_layer_actnorm.jl:60-123.  init: s = 1/sqrt(var), b = -mean/sqrt(var) with the
 * unbiased variance over (spatial, batch) (:67-72). */
int inb_actnorm_init(int B, int C, long long sp, const float* X, float* s, float* b, void* stream);
int inb_actnorm_forward(int B, int C, long long sp, const float* X, const float* s, const float* b,
                        float* Y, float* logdet /* nullable, device */, void* stream);
int inb_actnorm_inverse(int B, int C, long long sp, const float* Y, const float* s, const float* b,
                        float* X, void* stream);
int inb_actnorm_backward(int B, int C, long long sp, const float* dY, const float* Y, const float* s,
                         const float* b, int logdet, float* dX, float* X, float* ds, float* db,
                         void* stream);

/* Conv1x1 Householder mix, invertible_layer_conv1x1.jl:174-245 (logdet == 0). */
int inb_conv1x1_forward(int B, int C, long long sp, const float* X, const float* v1, const float* v2,
                        const float* v3, float* Y, void* stream);
int inb_conv1x1_inverse(int B, int C, long long sp, const float* Y, const float* v1, const float* v2,
                        const float* v3, float* X, void* stream);
/* inverse((dY, Y), C): dX, X and the gradients w.r.t. v1, v2, v3 (:227-245) */
int inb_conv1x1_backward(int B, int C, long long sp, const float* dY, const float* Y, const float* v1,
                         const float* v2, const float* v3, int freeze, float* dX, float* X,
                         float* dv1, float* dv2, float* dv3, void* stream);

/* ResidualBlock (fan = true), layer_residual_block.jl:119-178.  X: (B,Cin,spatial) -> (B,Cout,spatial) */
int inb_resblock_forward(int ndims, int nx, int ny, int nz, int B, int Cin, int nh, int Cout, int k1,
                         int k2, int precision, const float* X, const float* W1, const float* W2,
                         const float* W3, const float* b1, const float* b2, float* Y, void* stream);
int inb_resblock_backward(int ndims, int nx, int ny, int nz, int B, int Cin, int nh, int Cout, int k1,
                          int k2, int precision, const float* dY, const float* X, const float* W1,
                          const float* W2, const float* W3, const float* b1, const float* b2,
                          float* dX, float* dW1, float* dW2, float* dW3, float* db1, float* db2,
                          void* stream);

/* CouplingLayerGlow / ConditionalLayerGlow, invertible_layer_glow.jl:104-170,
 * conditional_layer_glow.jl:94-158.  cparams = {v1,v2,v3,W1,W2,W3,b1,b2}; Cond nullable (n_cond=0). */
int inb_coupling_forward(int ndims, int nx, int ny, int nz, int B, int C, int n_cond, int nh, int k1,
                         int k2, float low, float high, int precision, const float* X,
                         const float* Cond, float* const* cparams, float* Y,
                         float* logdet /* nullable */, void* stream);
int inb_coupling_inverse(int ndims, int nx, int ny, int nz, int B, int C, int n_cond, int nh, int k1,
                         int k2, float low, float high, int precision, const float* Y,
                         const float* Cond, float* const* cparams, float* X, void* stream);
int inb_coupling_backward(int ndims, int nx, int ny, int nz, int B, int C, int n_cond, int nh, int k1,
                          int k2, float low, float high, int logdet, int freeze, int precision,
                          const float* dY, const float* Y, const float* Cond, float* const* cparams,
                          float* const* cgrads, float* dX, float* X, float* dCond /* nullable */,
                          void* stream);

/* checkerboard squeeze / unsqueeze, dimensionality_operations.jl:40-47,79-107,137-166.
 * X: (B,C,[nz,]ny,nx) -> Y: (B,C*2^ndims,[nz/2,]ny/2,nx/2) */
int inb_squeeze(int ndims, int nx, int ny, int nz, int B, int C, const float* X, float* Y, void* stream);
int inb_unsqueeze(int ndims, int nx, int ny, int nz, int B, int C, const float* Y, float* X, void* stream);

/* ------------------------------------------------------------------ HINT family
 * Haar squeezes, 2-D: type 0 = wavelet_squeeze / wavelet_unsqueeze with WT.db1 (dimensionality_operations.jl:199-258,
 * channel 4c+q, q = approximation, x detail, y detail, diagonal), type 1 = Haar_squeeze / invHaar_unsqueeze
 * (:318-371, channel q*C+c, q = a, v, h, d).  (nx, ny, C) describe X for the squeeze and Y for the unsqueeze. */
#define INB_SQUEEZE_WAVELET 0
#define INB_SQUEEZE_HAAR 1
int inb_haar_squeeze(int nx, int ny, int B, int C, int type, const float* X, float* Y, void* stream);
int inb_haar_unsqueeze(int nx, int ny, int B, int C, int type, const float* Y, float* X, void* stream);

/* CouplingLayerHINT (invertible_layer_hint.jl:52-297) over CouplingLayerBasic (invertible_layer_basic.jl:62-149).
 * hparams = {CL[1].RB.(W1,W2,W3,b1,b2), ..., CL[n].RB.(...), [C.v1, C.v2, C.v3]} - the layer's get_params order,
 * n = inb_hint_depth(C) (get_depth, :63-71), the Conv1x1 entries present when permute != none; CL[j] acts on
 * C/2^j channels.  permute: 0 none, 1 full, 2 lower, 3 both.
 * shared_grads: a CL[j], j > 1, is applied 2^(j-1) times per pass; 0 = its gradient is the sum over the visits (the true
 * gradient; the reference's set_grad=false path, :222), 1 = only the last visit survives (what the reference's
 * set_grad=true path leaves in .grad, layer_residual_block.jl:168-172). */
/* CouplingLayerBasic (invertible_layer_basic.jl:62-149): X1 conditions, X2 is transformed, both (B, C1, spatial);
 * rbparams = RB.(W1 (k1..,C1,nh), W2, W3 (k1..,2*C1,nh), b1, b2).  forward: Y2 = S.*X2 + T (Y1 = X1 is the caller's);
 * inverse: X2; backward: dX1 = RB.backward(...) + dY1, dX2, X2 recomputed, the five gradients written. */
int inb_basic_coupling_forward(int ndims, int nx, int ny, int nz, int B, int C1, int nh, int k1, int k2, float low,
                               float high, int precision, const float* X1, const float* X2, float* const* rbparams,
                               float* Y2, float* logdet /* nullable */, void* stream);
int inb_basic_coupling_inverse(int ndims, int nx, int ny, int nz, int B, int C1, int nh, int k1, int k2, float low,
                               float high, int precision, const float* Y1, const float* Y2, float* const* rbparams,
                               float* X2, void* stream);
int inb_basic_coupling_backward(int ndims, int nx, int ny, int nz, int B, int C1, int nh, int k1, int k2, float low,
                                float high, int logdet, int precision, const float* dY1, const float* dY2,
                                const float* Y1, const float* Y2, float* const* rbparams, float* const* rbgrads,
                                float* dX1, float* dX2, float* X2, void* stream);

#define INB_PERMUTE_NONE 0
#define INB_PERMUTE_FULL 1
#define INB_PERMUTE_LOWER 2
#define INB_PERMUTE_BOTH 3
int inb_hint_depth(int C);
int inb_hint_coupling_forward(int ndims, int nx, int ny, int nz, int B, int C, int nh, int k1, int k2, float low,
                              float high, int permute, int precision, const float* X, float* const* hparams,
                              float* Y, float* logdet /* nullable */, void* stream);
int inb_hint_coupling_inverse(int ndims, int nx, int ny, int nz, int B, int C, int nh, int k1, int k2, float low,
                              float high, int permute, int precision, const float* Y, float* const* hparams,
                              float* X, void* stream);
int inb_hint_coupling_backward(int ndims, int nx, int ny, int nz, int B, int C, int nh, int k1, int k2, float low,
                               float high, int permute, int logdet, int shared_grads, int precision,
                               const float* dY, const float* Y, float* const* hparams, float* const* hgrads,
                               float* dX, float* X, void* stream);

/* NetworkMultiScaleHINT (invertible_network_hint_multiscale.jl:55-174), 2-D: per scale a Haar squeeze, then K x
 * (ActNorm, CouplingLayerHINT with permute = "full", logdet = true).  params / grads in get_params order:
 * AN[i,j].(s,b) for i = 1..L, j = 1..K, then CL[i,j].(CL[1..n_i].RB.(W1,W2,W3,b1,b2), C.(v1,v2,v3)). */
typedef struct inb_hint_plan inb_hint_plan;
typedef struct inb_hint_desc {
  int nx, ny;         /* input spatial size */
  int n_in;           /* input channels */
  int n_hidden;
  int L, K;
  int batch;          /* batch size the workspace is sized for */
  int split_scales;   /* :74-82: split half the channels off after every scale but the last */
  int k1, k2, p1, p2; /* ResidualBlock kernels ("same" padding); the reference's default is 3,3,1,1 */
  float sig_low, sig_high;
  int squeeze_type;   /* INB_SQUEEZE_*; the reference uses wavelet_squeeze (:102) */
  int shared_grads;   /* see inb_hint_coupling_backward */
  int precision;      /* INB_PREC_* */
} inb_hint_desc;
int inb_hint_plan_create(const inb_hint_desc* desc, inb_hint_plan** plan);
int inb_hint_plan_destroy(inb_hint_plan* plan);
int inb_hint_num_params(const inb_hint_plan* plan);
int inb_hint_param_numel(const inb_hint_plan* plan, int index, long long* numel);
long long inb_hint_workspace_bytes(const inb_hint_plan* plan);
/* Z: flat, X's element count ([vec(Z_1);..;vec(X_L)] when split_scales, else the (B, n_in*4^L, ny/2^L, nx/2^L) tensor) */
int inb_hint_forward(inb_hint_plan* plan, int batch, const float* X, float* const* params, float* Z,
                     float* logdet, int init_actnorm, void* stream);
int inb_hint_inverse(inb_hint_plan* plan, int batch, const float* Z, float* const* params, float* X, void* stream);
int inb_hint_backward(inb_hint_plan* plan, int batch, const float* dZ, const float* Z, float* const* params,
                      float* const* grads, float* dX, float* X, void* stream);

/* Gaussian negative log-likelihood value and gradient, objective_functions.jl:54,65:
 * loss = sum(Z^2)/(2B) (device float, nullable), dZ = Z/B. */
int inb_nll_grad(long long n, int B, const float* Z, float* dZ, float* loss, void* stream);

/* Flux.Optimise.ADAM over a flat parameter / gradient buffer (replaces the per-parameter `update!(opt, p.data, p.grad)`
 * loop of examples/networks/network_glow.jl:38-42): m, v are the caller's moment buffers (zero before the first step),
 * beta1_pow_t = beta1^t, beta2_pow_t = beta2^t for this step t >= 1 (Flux keeps the same running products). */
int inb_adam_update(long long n, float* params, const float* grads, float* m, float* v, float lr, float beta1, float beta2,
                    float eps, float beta1_pow_t, float beta2_pow_t, void* stream);

/* ------------------------------------------------------------------ accounting (bench / tests)
 * inb_launch_count: kernels launched by this library since load (process wide).
 * Profiling: when enabled every kernel family is timed with CUDA events on the launching stream
 * and its algorithmic flops / bytes are summed; inb_prof_get synchronises on the recorded events. */
long long inb_launch_count(void);
int inb_prof_enable(int on);
/* diagnostics: device buffer of 4 x 32 x 16 int64; launch i of the fused ResidualBlock kernel after this call
   writes block i % 4: per-tile phase timestamps (SM clock) of CTA 0 (MMA issuer: slots 0-5, first epilogue
   warp: slots 6-11); NULL switches it off */
int inb_debug_chain_trace(void* dev_buf);
/* CUDA-graph replay accounting of a plan's network-level calls: graphs captured, graph launches, and calls
   launched kernel by kernel because the caller's buffer addresses kept changing (see api.cu run_graphed) */
int inb_glow_graph_stats(const inb_plan* plan, long long* captures, long long* replays, long long* direct);
int inb_prof_reset(void);
int inb_prof_num(void);
int inb_prof_get(int index, char* name, int name_len, long long* launches, long long* scopes, double* ms,
                 double* flops, double* bytes);

#ifdef __cplusplus
}
#endif
#endif /* INB200_H */
