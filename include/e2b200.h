/*
 * e2b200.h -- C ABI of libe2b200.so, the B200 (sm_100a) implementation of
 * ELEKTRONN2's volumetric-CNN hot path.
 *
 * The reference has no FFI for this path: it calls Theano, which calls
 * cuDNN/BLAS.  Each entry point below therefore names the
 * reference call site it replaces (paths relative to /root/reference/elektronn2);
 * INTEGRATION.md shows the theano.Op / ctypes stub a maintainer would add at
 * those sites.
 *
 * Conventions
 *  - plain C, no torch types; every pointer is a DEVICE pointer owned by the
 *    caller (the library never allocates, frees or retains them);
 *  - every call is asynchronous on the given cudaStream_t (passed as void*);
 *  - return value: E2_OK or a negative e2_status; e2_last_error(h) holds the text;
 *  - activations are channels-last "NDHWC": element (n,z,x,y,c) lives at
 *    ((((n*Z + z)*X + x)*Y + y) * c_pitch + c), c_pitch >= c.  The reference's
 *    (b,f,z,x,y) arrays enter/leave through e2_ncdhw_to_ndhwc / e2_ndhwc_to_ncdhw;
 *  - conv weights are exchanged in the reference's own layout
 *    (f_out, f_in, kz, kx, ky) float32 (neural.py:618-620) and re-packed on the
 *    device by e2_conv3d_pack_weights;
 *  - a handle is bound to one device; calls on different handles are independent;
 *    the library keeps no mutable state outside the handle.  The only process-wide inputs are read-once
 *    E2_* environment switches for A/B profiling (kernel selection, split plans); they never change results
 *    beyond fp32 summation order and are listed in DESIGN.md.
 */
#ifndef E2B200_H
#define E2B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define E2B200_VERSION 100

typedef struct e2_handle e2_handle;

typedef enum {
  E2_OK = 0,
  E2_ERR_INVALID = -1,     /* bad descriptor / null pointer / shape mismatch (ValueError in the reference) */
  E2_ERR_UNSUPPORTED = -2, /* valid request this build cannot serve (NotImplementedError) */
  E2_ERR_CUDA = -3,        /* a CUDA runtime/driver call failed */
  E2_ERR_WORKSPACE = -4    /* workspace too small, see e2_*_workspace_size */
} e2_status;

/* apply_activation, computations.py:57-134 (the parameter-free subset) */
/* apply_activation, computations.py:57-134: 'soft+' = log(1+exp(x)), 'elu' = T.nnet.elu(x, 1), 'selu' = scale * elu(x, alpha) */
typedef enum { E2_ACT_LIN = 0, E2_ACT_RELU = 1, E2_ACT_TANH = 2, E2_ACT_SIGMOID = 3, E2_ACT_ABS = 4,
               E2_ACT_SOFTPLUS = 5, E2_ACT_ELU = 6, E2_ACT_SELU = 7,
               E2_ACT_PRELU = 8 /* T.nnet.relu(x, alpha): only through e2_affine_act_* (needs the slope parameter) */
} e2_act;

/* arithmetic used inside the conv GEMMs (accumulation is always fp32) */
typedef enum {
  E2_COMPUTE_F32 = 0,  /* CUDA-core FFMA, exact fp32 products                          */
  E2_COMPUTE_TF32 = 1, /* tcgen05.mma kind::tf32, operands rounded to tf32 (rel 2^-11) */
  E2_COMPUTE_BF16 = 2  /* reserved */
} e2_compute;

typedef enum { E2_TIE_FIRST = 0, E2_TIE_ALL = 1 } e2_tie_mode;
/* computations.pooling modes (computations.py:556-561, 589-590): dnn_pool's 'average_inc_pad' and 'average_exc_pad'
 * coincide because the path never pads */
typedef enum { E2_POOL_MAX = 0, E2_POOL_AVERAGE = 1, E2_POOL_SUM = 2 } e2_pool_mode;

typedef struct {
  int32_t n, z, x, y, c; /* logical extents */
  int32_t c_pitch;       /* floats between consecutive positions, >= c */
} e2_tensor;

/* ---------------------------------------------------------------- handle */
int e2_version(void);
int e2_create(e2_handle** out, int device);
int e2_destroy(e2_handle* h);
const char* e2_last_error(const e2_handle* h);
/* number of kernel launches issued through this handle since creation */
int64_t e2_launch_count(const e2_handle* h);

/* --------------------------------------------------------- layout / boundary
 * Node.__call__ (node_basic.py:464-494) hands numpy (b,f,z,x,y) arrays to Theano;
 * here they are converted once at the edge.  ncdhw is dense C-contiguous. */
int e2_ncdhw_to_ndhwc(e2_handle* h, const e2_tensor* t, const float* ncdhw, float* ndhwc, void* stream);
int e2_ndhwc_to_ncdhw(e2_handle* h, const e2_tensor* t, const float* ndhwc, float* ncdhw, void* stream);
/* same tensor, different channel pitch (pad lanes zero); round_tf32 != 0 rounds the values to tf32.
 * Hands the dense single-channel raw input to the TMA/tcgen05 path, which needs 16-byte rows. */
int e2_repitch(e2_handle* h, const e2_tensor* src_t, const float* src, const e2_tensor* dst_t, float* dst,
               int32_t round_tf32, void* stream);
/* predict_dense input scaling, node_basic.py:904-910: uint8 -> float32 / 255 (count elements) */
int e2_u8_to_f32(e2_handle* h, const uint8_t* src, float* dst, int64_t count, float scale, void* stream);
/* predict_dense as_uint8 output, node_basic.py:990-996: trunc(p*255) */
int e2_f32_to_u8(e2_handle* h, const float* src, uint8_t* dst, int64_t count, float scale, void* stream);

/* ------------------------------------------------------------------ conv3d
 * computations.conv 3-D 'valid' branch, computations.py:364-428 (true convolution,
 * kernel flipped; stride 1), the 1x1x1 tensordot shortcut :330-335/:377-384, and the
 * bias+activation epilogue of Conv._make_output, neural.py:711-712.
 * Backward: what T.grad (model.py:182) derives for them. */
typedef struct {
  e2_tensor x;        /* input  (n, Z, X, Y, c_in)                               */
  e2_tensor y;        /* output (n, Z-kz+1, X-kx+1, Y-ky+1, c_out)               */
  int32_t kz, kx, ky;
  int32_t act;        /* e2_act applied after +bias (fwd only)                   */
  int32_t has_bias;
  int32_t compute;    /* e2_compute                                              */
  int32_t accumulate; /* dgrad only: dx += result instead of dx = result          */
} e2_conv_desc;

/* floats needed for the packed forward / dgrad weight images */
int e2_conv3d_packed_floats(const e2_conv_desc* d, size_t* fwd_floats, size_t* dgrad_floats);
/* w (f_out,f_in,kz,kx,ky) -> wf[o][tap][c] (flipped taps, fwd) and wd[c][tap][o] (dgrad);
 * either output may be NULL.  With compute==TF32 values are rounded to tf32 (rna). */
int e2_conv3d_pack_weights(e2_handle* h, const e2_conv_desc* d, const float* w, float* wf, float* wd, void* stream);
int e2_conv3d_workspace_size(const e2_conv_desc* d, size_t* bytes);
/* y = act(conv(x, w) + bias) */
int e2_conv3d_fwd(e2_handle* h, const e2_conv_desc* d, const float* x, const float* wf, const float* bias,
                  float* y, void* ws, size_t ws_bytes, void* stream);
/* dx (=|+=) full-correlation(dy, w).
 * relu_gate (nullable, all four *_dgrad / *_bwd entry points that take it): the post-ReLU
 * output of the layer that produced x (same geometry and pitch as dx).  When given, the
 * contribution computed by this call is zeroed wherever relu_gate <= 0, i.e. the ReLU
 * backward of the upstream layer is fused into this kernel's epilogue. */
int e2_conv3d_dgrad(e2_handle* h, const e2_conv_desc* d, const float* dy, const float* wd, float* dx,
                    const float* relu_gate, void* ws, size_t ws_bytes, void* stream);
/* dw (reference layout, overwritten) and optional db = sum dy */
int e2_conv3d_wgrad(e2_handle* h, const e2_conv_desc* d, const float* x, const float* dy, float* dw, float* db,
                    void* ws, size_t ws_bytes, void* stream);

/* ----------------------------------------------------------------- upconv3d
 * UpConv._make_output, neural.py:989-1072: CPU path unpooling+conv
 * (computations.py:749-756, neural.py:1013-1020), GPU path computations.upconv
 * (computations.py:216-255).  kernel == pool == stride (neural.py:968), so
 *   y[n, z*pz+i, x*px+j, y*py+k, o] = act(sum_c x[n,z,x,y,c] * w[o,c,i,j,k] + b[o]). */
typedef struct {
  e2_tensor x; /* (n, Z, X, Y, c_in)            */
  e2_tensor y; /* (n, Z*pz, X*px, Y*py, c_out)  */
  int32_t pz, px, py;
  int32_t act, has_bias, compute, accumulate;
} e2_upconv_desc;

int e2_upconv3d_packed_floats(const e2_upconv_desc* d, size_t* fwd_floats, size_t* dgrad_floats);
/* scratch bytes the fwd / dgrad / wgrad entry points can use (split-K partial tiles); 0 is always accepted */
int e2_upconv3d_workspace_size(const e2_upconv_desc* d, size_t* bytes);
int e2_upconv3d_pack_weights(e2_handle* h, const e2_upconv_desc* d, const float* w, float* wf, float* wd, void* stream);
int e2_upconv3d_fwd(e2_handle* h, const e2_upconv_desc* d, const float* x, const float* wf, const float* bias,
                    float* y, void* ws, size_t ws_bytes, void* stream);
int e2_upconv3d_dgrad(e2_handle* h, const e2_upconv_desc* d, const float* dy, const float* wd, float* dx,
                      const float* relu_gate, void* ws, size_t ws_bytes, void* stream);
int e2_upconv3d_wgrad(e2_handle* h, const e2_upconv_desc* d, const float* x, const float* dy, float* dw, float* db,
                      void* ws, size_t ws_bytes, void* stream);

/* -------------------------------------------------- activation backward / bias
 * dpre = dy * act'(y) evaluated from the stored post-activation y; may run in place
 * (dpre == dy). */
int e2_act_bwd(e2_handle* h, const e2_tensor* t, int32_t act, const float* y, const float* dy, float* dpre, void* stream);

/* ----------------------------------------------------------------- pooling
 * computations.pooling 3-D 'max', stride == pool, ignore_border
 * (computations.py:538-649; dnn_pool :600 / pool_2d + z-maximum :617-631).
 * Optional fused "+bias -> act" for Conv nodes that carry a pool: the reference
 * pools BEFORE bias and activation (neural.py:678, 711-712).
 * argmax (may be NULL): int32 z*X*Y + x*Y + y of the FIRST maximum in (z,x,y)
 * row-major scan order of the window. */
typedef struct {
  e2_tensor x; /* (n, Z, X, Y, c)                */
  e2_tensor y; /* (n, Z/pz, X/px, Y/py, c)       */
  int32_t pz, px, py;
  int32_t act, has_bias; /* fused epilogue, fwd only                 */
  int32_t tie_mode;      /* e2_tie_mode, bwd only                    */
  int32_t accumulate;    /* bwd only: dx += instead of dx =          */
  int32_t round_tf32;    /* fwd only: round the output to tf32 (it feeds a kind::tf32 MMA) */
  int32_t gate_pooled;   /* bwd only: relu_gate has the geometry of y (the pooled tensor) instead of x.  Valid
                          * for a pool without bias/activation whose input is a post-ReLU tensor: the only
                          * element of a window that receives gradient is the argmax, whose value IS the
                          * pooled value, so gate(x)[argmax] == y -- one eighth of the gate traffic */
  int32_t mode;          /* e2_pool_mode; AVERAGE / SUM take no bias / activation / argmax / gate */
} e2_pool_desc;

int e2_maxpool3d_fwd(e2_handle* h, const e2_pool_desc* d, const float* x, const float* bias, float* y,
                     int32_t* argmax, void* stream);
/* The conv and the max-pool that consumes it as ONE launch: exactly e2_conv3d_fwd(d, ..., y) followed by
 * e2_maxpool3d_fwd(p, y, pool_bias, yp, argmax) -- same values, same argmax, bit for bit -- with the window reduced in
 * the conv kernel's epilogue, so the unpooled tensor is not read back (Conv nodes that carry a pool, neural.py:662-712:
 * conv -> pool -> +bias -> act; and the Conv -> Pool pairs of the U-Nets, examples/unet3d.py:63-74).  p->x must describe
 * the conv's output.  y may be NULL when the caller has no use for the unpooled tensor (fused problems only); y_keep
 * (nullable = everything) names the window of y the caller will read -- the skip connection's Crop, neural.py:1152-1168
 * -- and only that part of y is guaranteed to be written (the fused kernel stores the tiles that intersect it).
 * e2_conv3d_fwd_pool_supported: 1 if the pair is fused (TF32 mode, windows of 1 or 2 per axis, a layer the z-stack
 * kernel runs without a K split), 0 if e2_conv3d_fwd_pool would issue the two launches instead, < 0 on a bad descriptor.
 * Whether fusing PAYS is the caller's policy: layers with a short reduction (c_in * taps below ~800) are bound by their
 * epilogue already and ran 2 % slower per step with the pool inside it (elektronn2_b200/neuromancer/executor.py). */
int e2_conv3d_fwd_pool_supported(e2_handle* h, const e2_conv_desc* d, const e2_pool_desc* p);
typedef struct { int32_t z0, z1, x0, x1, y0, y1; } e2_window; /* half-open ranges of positions */
int e2_conv3d_fwd_pool(e2_handle* h, const e2_conv_desc* d, const e2_pool_desc* p, const float* x, const float* wf,
                       const float* bias, const float* pool_bias, float* y, const e2_window* y_keep, float* yp,
                       int32_t* argmax, void* ws, size_t ws_bytes, void* stream);
/* E2_TIE_FIRST routes dy through argmax (x may be NULL); E2_TIE_ALL (Theano-CPU
 * semantics) needs x and the pooled pre-bias maximum is recomputed from it. */
int e2_maxpool3d_bwd(e2_handle* h, const e2_pool_desc* d, const float* dy, const int32_t* argmax, const float* x,
                     float* dx, const float* relu_gate, void* stream);

/* ------------------------------------------------- unfused epilogue (next-row 8f-4)
 * Conv._make_output / UpConv._make_output between pooling and dropout, neural.py:681-712:
 *     y = act((gamma / std) * v + b - gamma * mean / std)
 * for the configurations the fused conv / pool epilogues do not cover: batch normalisation ('train': batch
 * statistics over all axes but f, std = T.std + 1e-6, running averages 0.9995/0.0005, neural.py:681-698;
 * 'predict': stored mean / std / gamma, :700-703), 'prelu' (b is (f_out, 2): bias b[:,0], slope b[:,1],
 * neural.py:655-657), and activations whose derivative needs the pre-activation ('abs').
 * scale / shift / mean / std are DEVICE float[c]; NULL scale == 1, NULL shift == 0. */
typedef struct {
  e2_tensor t;          /* geometry shared by v, y, dy, dv                                   */
  int32_t act;          /* e2_act, E2_ACT_PRELU included                                    */
  int32_t batch_stats;  /* bwd: scale / shift were derived from the batch statistics of v   */
  int32_t round_tf32;   /* round y (fwd) / dv (bwd) to tf32                                 */
  int32_t param_stride; /* floats between consecutive channels of alpha / dbias / dalpha (2 for prelu's (f,2) b) */
} e2_affine_desc;

/* batch statistics of v -> mean, std (= sqrt(mean((v-mean)^2)) + 1e-6), scale = gamma/std, shift = b - gamma*mean/std;
 * run_mean / run_std (nullable) <- keep * run + (1-keep) * batch.  scratch: DEVICE double[2*c]. */
int e2_bn_batch_stats(e2_handle* h, const e2_tensor* t, const float* v, const float* gamma, const float* bias,
                      int32_t bias_stride, float* mean, float* std_out, float* scale, float* shift, float* run_mean,
                      float* run_std, float keep, double* scratch, void* stream);
/* scale = gamma/std, shift = b - gamma*mean/std from stored parameters (any pointer may be NULL: gamma=1, b=0, mean=0, std=1) */
int e2_bn_fold(e2_handle* h, int32_t c, const float* gamma, const float* bias, int32_t bias_stride, const float* mean,
               const float* std_in, float* scale, float* shift, void* stream);
int e2_affine_act_fwd(e2_handle* h, const e2_affine_desc* d, const float* v, const float* scale, const float* shift,
                      const float* alpha, float* y, void* stream);
/* dv and the parameter gradients (each nullable): dbias = sum dpre, dgamma = sum dpre * (v-mean)/std,
 * dalpha = sum dy * min(pre, 0).  With batch_stats the gradient flows through mean and std as T.grad derives it.
 * scratch: e2_affine_scratch_bytes(c) DEVICE bytes, 8-byte aligned. */
int e2_affine_act_bwd(e2_handle* h, const e2_affine_desc* d, const float* v, const float* dy, const float* scale,
                      const float* shift, const float* alpha, const float* mean, const float* std_in, float* dv,
                      float* dgamma, float* dbias, float* dalpha, double* scratch, void* stream);
int e2_affine_scratch_bytes(int32_t c, size_t* bytes);

/* computations.maxout, computations.py:455-495: y[o, f, i] = max_k x[o, f*factor + k, i] on a dense
 * (outer, f_out*factor, inner) array; the backward routes dy to the first maximal slice. */
int e2_maxout_fwd(e2_handle* h, const float* x, float* y, int64_t outer, int32_t f_out, int64_t inner, int32_t factor,
                  void* stream);
int e2_maxout_bwd(e2_handle* h, const float* x, const float* dy, float* dx, int64_t outer, int32_t f_out, int64_t inner,
                  int32_t factor, void* stream);

/* ------------------------------------------------------ max-fragment-pooling
 * computations.fragmentpool, computations.py:652-678.  Output batch = prod(p) * n,
 * new-offset-major / old-fragment-minor, offsets in itertools.product order (last
 * spatial axis fastest).  y spatial = (S - p + 1) / p per axis. */
typedef struct {
  e2_tensor x; /* (n, Z, X, Y, c)                                    */
  e2_tensor y; /* (n*pz*px*py, (Z-pz+1)/pz, (X-px+1)/px, (Y-py+1)/py, c) */
  int32_t pz, px, py;
  int32_t act, has_bias;
  int32_t round_tf32; /* fwd only, as in e2_pool_desc */
} e2_mfp_desc;

int e2_mfp_fwd(e2_handle* h, const e2_mfp_desc* d, const float* x, const float* bias, float* y, int32_t* argmax,
               void* stream);
int e2_mfp_bwd(e2_handle* h, const e2_mfp_desc* d, const float* dy, const int32_t* argmax, float* dx, void* stream);

/* ------------------------------------------------------- fragments -> dense
 * computations.fragments2dense, computations.py:681-701:
 *   dense[0, off_k[0]::s0, off_k[1]::s1, off_k[2]::s2, :] = frag[k].
 * offsets: DEVICE int32 [n_frag][3]. */
typedef struct {
  e2_tensor frag;  /* (n_frag, Z, X, Y, c), n_frag == sz*sx*sy */
  e2_tensor dense; /* (1, Z*sz, X*sx, Y*sy, c)                 */
  int32_t sz, sx, sy;
} e2_f2d_desc;

int e2_frag2dense_fwd(e2_handle* h, const e2_f2d_desc* d, const float* frag, const int32_t* offsets, float* dense,
                      void* stream);
int e2_frag2dense_bwd(e2_handle* h, const e2_f2d_desc* d, const float* ddense, const int32_t* offsets, float* dfrag,
                      void* stream);

/* ----------------------------------------------------------- crop + concat
 * Crop (neural.py:1152-1168) fused with the channel placement of Concat(axis='f')
 * (node_basic.py:1403-1451) as AutoMerge builds them (neural.py:1393-1399):
 *   dst[n, z, x, y, dst_c0 + c] = src[n, z+oz, x+ox, y+oy, c]. */
typedef struct {
  e2_tensor src; /* (n, Z, X, Y, c)                        */
  e2_tensor dst; /* (n, Z-2oz, X-2ox, Y-2oy, c_total)      */
  int32_t oz, ox, oy;
  int32_t dst_c0;
  int32_t accumulate; /* bwd only */
} e2_crop_desc;

int e2_crop_concat_fwd(e2_handle* h, const e2_crop_desc* d, const float* src, float* dst, void* stream);
/* dsrc[.. cropped region ..] (=|+=) ddst[..., dst_c0 + c]; with accumulate==0 the border of dsrc is zeroed */
int e2_crop_concat_bwd(e2_handle* h, const e2_crop_desc* d, const float* ddst, float* dsrc, const float* relu_gate,
                       void* stream);

/* ------------------------------------------------ loss head (next-row 8f-2)
 * Softmax (computations.py:170-177) -> MultinoulliNLL sparse target
 * (loss.py:261-347, EPS=1e-5) -> AggregateLoss mean (loss.py:1346-1363) and
 * Classification/_Errors (loss.py:736-737, 805-807) in one pass.
 * out_scalars (device float[4]): loss_sum (sum of -log(p_t+EPS) over labelled voxels), n_labelled, n_errors, unused.
 * The caller finishes: loss = loss_sum / (n_labelled + EPS).
 * target == NULL computes the Softmax node alone (inference). */
int e2_softmax_nll_fwd(e2_handle* h, const e2_tensor* logits, const float* x, const float* target, float* probs,
                       float* out_scalars, void* stream);
/* dlogits given scalars from the forward pass; grad_scale multiplies the result (1/N for data parallel means) */
int e2_softmax_nll_bwd(e2_handle* h, const e2_tensor* logits, const float* probs, const float* target,
                       const float* scalars, float grad_scale, float* dlogits, void* stream);

/* ---------------------------------------------------- optimiser (next-row 8f-2)
 * Adam exactly as optimiser.py:301-324: eps=1e-5 inside the sqrt, factor =
 * sqrt(1-beta2^t)/(1-mom^t), L2 term wd*wd_mult*p where wd_mult is the parameter's apply_reg (0: none, 1, or a
 * multiplier > 1 such as batch-norm gamma's 3.0, neural.py:213; optimiser.py:312-318).  t is the 1-based step. */
int e2_adam_step(e2_handle* h, float* p, const float* g, float* m, float* s, int64_t count, float lr, float mom,
                 float beta2, float wd, float wd_mult, int32_t t, void* stream);
/* The same update with its hyper-parameters in DEVICE memory, so that the step can live inside a CUDA graph
 * (graph nodes bake their scalar arguments): hyper = float[8] {lr, mom, beta2, wd, factor, -, -, -}, t_dev = the
 * number of steps taken so far.  e2_adam_prepare increments *t_dev and refreshes hyper[4] = factor(t); it is the
 * first node of a training-step graph, e2_adam_step_dev may then run on any stream ordered behind it. */
int e2_adam_prepare(e2_handle* h, float* hyper, int32_t* t_dev, void* stream);
int e2_adam_step_dev(e2_handle* h, float* p, const float* g, float* m, float* s, int64_t count, const float* hyper,
                     float wd_mult, void* stream);
/* e2_adam_step_dev for the weights of ONE Conv / UpConv layer that also writes the packed forward / dgrad images of the
 * updated weights (what e2_conv3d_pack_weights / e2_upconv3d_pack_weights would produce from them, bit for bit) in the
 * same pass: the weights cross HBM once per step instead of three times.  w, g, m, s: the layer's (f_out,f_in,k..)
 * slices of the parameter / gradient / moment buffers.  wf / wd must have been packed once before (pad lanes are not
 * rewritten); either may be NULL.  E2_ERR_UNSUPPORTED for filters with more than ~150 taps (use the two calls). */
int e2_conv3d_adam_pack_dev(e2_handle* h, const e2_conv_desc* d, float* w, const float* g, float* m, float* s,
                            const float* hyper, float wd_mult, float* wf, float* wd, void* stream);
int e2_upconv3d_adam_pack_dev(e2_handle* h, const e2_upconv_desc* d, float* w, const float* g, float* m, float* s,
                              const float* hyper, float wd_mult, float* wf, float* wd, void* stream);
/* SGD with momentum, optimiser.py:146-160 */
int e2_sgd_step(e2_handle* h, float* p, const float* g, float* last_dir, int64_t count, float lr, float mom, float wd,
                float wd_mult, void* stream);

/* ---------------------------------------------------------------- introspection
 * Host-only (no GPU needed): the tile plan the z-stack conv kernel would use for a conv of K input / N output
 * channels with a (kz,kx,ky) filter and (Oz,Ox,Oy) output positions on a device with sm_count SMs.
 * out[8] = {BN, TZ, ksplit, channel blocks per split, weight slots, plane slots, units, N tiles}.
 * Returns 1 if the kernel takes the problem, 0 if it is left to the tap kernel.  (No reference counterpart:
 * cuDNN's algorithm choice is opaque; used by tests/test_host_api.py to pin the planner.) */
int e2_debug_zstack_plan(int sm_count, int K, int N, int Oz, int Ox, int Oy, int kz, int kx, int ky, int may_split,
                         int* out);
/* The same for a conv whose max-pool (window qz,qx,qy) runs in the kernel's epilogue (e2_conv3d_fwd_pool): returns 1 and
 * the plan if the pair is fused, 0 if it runs as two launches. */
int e2_debug_zstack_pool_plan(int sm_count, int K, int N, int Oz, int Ox, int Oy, int kz, int kx, int ky, int qz, int qx,
                              int qy, int* out);

#ifdef __cplusplus
}
#endif
#endif /* E2B200_H */
