// Loss head (Softmax -> MultinoulliNLL -> AggregateLoss, Errors) and optimiser updates:
// the memory-bound ops on either side of the conv stack.
#include "e2_common.cuh"

#define E2_EPS 1e-5f  // loss.py:30

// One thread per position; the class axis (c <= 8) is walked in registers.
// scalars[0] += sum -log(p_t+EPS), scalars[1] += #labelled, scalars[2] += #errors
__global__ void __launch_bounds__(256) k_softmax_nll_fwd(const float* __restrict__ x, const float* __restrict__ target,
                                                         float* __restrict__ probs, float* __restrict__ scalars,
                                                         int64_t P, int C, int pitch) {
  float l_sum = 0.f, l_lab = 0.f, l_err = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (int64_t)gridDim.x * blockDim.x) {
    const float* r = x + i * pitch;
    float mx = r[0];
    int am = 0;
    for (int c = 1; c < C; ++c) {
      float v = r[c];
      if (v > mx) mx = v, am = c;  // first maximum, like T.argmax
    }
    float den = 0.f;
    for (int c = 0; c < C; ++c) den += expf(r[c] - mx);  // computations.py:175-176
    // target == NULL: plain Softmax node (inference), no loss / error statistics
    float tf = target ? target[i] : -1.f;
    int tc = (int)tf;
    bool labelled = target && (tf == (float)tc) && tc >= 0 && tc < C;  // T.eq(target, classes), loss.py:271-276
    for (int c = 0; c < C; ++c) {
      float p = expf(r[c] - mx) / den;
      probs[i * pitch + c] = p;
      if (labelled && c == tc) l_sum += -logf(p + E2_EPS);  // -xlogy0(target, pred + EPS), loss.py:318
    }
    if (labelled) l_lab += 1.f;
    // _Errors: mean(int16(target) != argmax) over ALL positions (loss.py:803-807)
    if (target && (int)(short)tf != am) l_err += 1.f;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    l_sum += __shfl_xor_sync(0xffffffffu, l_sum, o);
    l_lab += __shfl_xor_sync(0xffffffffu, l_lab, o);
    l_err += __shfl_xor_sync(0xffffffffu, l_err, o);
  }
  __shared__ float red[3][8];
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) red[0][w] = l_sum, red[1][w] = l_lab, red[2][w] = l_err;
  __syncthreads();
  if (threadIdx.x < 3) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += red[threadIdx.x][i];
    atomicAdd(scalars + threadIdx.x, s);
  }
}

// loss = loss_sum / (n_lab + EPS)   [nll * size/(n_tot+EPS)/n_class summed over classes, then mean: loss.py:343-346,1357-1362]
// dlogit_c = -(p_t/(p_t+EPS)) * ([c==t] - p_c) / (n_lab+EPS) * grad_scale   for labelled positions, else 0
__global__ void __launch_bounds__(256) k_softmax_nll_bwd(const float* __restrict__ probs, const float* __restrict__ target,
                                                         const float* __restrict__ scalars, float grad_scale,
                                                         float* __restrict__ dlogits, int64_t P, int C, int pitch) {
  const float inv = grad_scale / (scalars[1] + E2_EPS);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (int64_t)gridDim.x * blockDim.x) {
    float tf = target[i];
    int tc = (int)tf;
    bool labelled = (tf == (float)tc) && tc >= 0 && tc < C;
    const float* p = probs + i * pitch;
    float pt = labelled ? p[tc] : 0.f;
    float k = labelled ? -(pt / (pt + E2_EPS)) * inv : 0.f;
    for (int c = 0; c < C; ++c) dlogits[i * pitch + c] = k * ((c == tc ? 1.f : 0.f) - p[c]);
  }
}

extern "C" int e2_softmax_nll_fwd(e2_handle* h, const e2_tensor* t, const float* x, const float* target, float* probs,
                                  float* out_scalars, void* stream) {
  E2_REQUIRE(h, e2_tensor_ok(t) && x && probs && out_scalars, "softmax_nll_fwd: bad arguments");
  E2_REQUIRE(h, t->c <= 64, "softmax_nll_fwd: at most 64 classes");
  cudaStream_t s = (cudaStream_t)stream;
  cudaMemsetAsync(out_scalars, 0, 4 * sizeof(float), s);
  int64_t P = e2_positions(t);
  k_softmax_nll_fwd<<<e2_grid_1d(P, 256, h->sm_count, 8), 256, 0, s>>>(x, target, probs, out_scalars, P, t->c, t->c_pitch);
  e2_count_launch(h);
  E2_CUDA_CHECK(h, "softmax_nll_fwd");
  return E2_OK;
}

extern "C" int e2_softmax_nll_bwd(e2_handle* h, const e2_tensor* t, const float* probs, const float* target,
                                  const float* scalars, float grad_scale, float* dlogits, void* stream) {
  E2_REQUIRE(h, e2_tensor_ok(t) && probs && target && scalars && dlogits, "softmax_nll_bwd: bad arguments");
  int64_t P = e2_positions(t);
  k_softmax_nll_bwd<<<e2_grid_1d(P, 256, h->sm_count, 8), 256, 0, (cudaStream_t)stream>>>(probs, target, scalars,
                                                                                         grad_scale, dlogits, P, t->c,
                                                                                         t->c_pitch);
  e2_count_launch(h);
  E2_CUDA_CHECK(h, "softmax_nll_bwd");
  return E2_OK;
}

// ------------------------------------------------------------------------ optimisers
// one element of the update (optimiser.py:306-319); shared by the flat kernel and the per-layer update + re-pack kernel so
// that both produce the same bits
__device__ __forceinline__ float e2_adam_update(float pi, float gi, float& mi, float& si, float lr, float mom, float beta2,
                                                float wd, float factor, float wd_mult) {
  // explicit roundings (no compiler-chosen fma contraction): every kernel that inlines this produces the same bits
  const float nm = __fmaf_rn(mom, mi, __fmul_rn(1.0f - mom, gi));                          // optimiser.py:306
  const float ns = __fmaf_rn(beta2, si, __fmul_rn(__fmul_rn(1.0f - beta2, gi), gi));       // :307
  const float dir = __fdiv_rn(__fmul_rn(factor, nm), __fsqrt_rn(__fadd_rn(ns, 1e-5f)));    // :309, epsilon inside the sqrt
  pi = wd_mult != 0.f ? __fmaf_rn(-lr, __fmaf_rn(__fmul_rn(wd, wd_mult), pi, dir), pi)     // :310-319
                      : __fmaf_rn(-lr, dir, pi);
  mi = nm, si = ns;
  return pi;
}

__global__ void __launch_bounds__(256) k_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                              float* __restrict__ s, int64_t n, float lr, float mom, float beta2, float wd,
                                              float wd_mult, float factor) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float mi = m[i], si = s[i];
    const float pi = e2_adam_update(p[i], g[i], mi, si, lr, mom, beta2, wd, factor, wd_mult);
    m[i] = mi, s[i] = si, p[i] = pi;
  }
}

extern "C" int e2_adam_step(e2_handle* h, float* p, const float* g, float* m, float* s, int64_t count, float lr,
                            float mom, float beta2, float wd, float wd_mult, int32_t t, void* stream) {
  E2_REQUIRE(h, p && g && m && s && count >= 0 && t >= 1, "adam_step: bad arguments");
  if (count == 0) return E2_OK;
  double factor = sqrt(1.0 - pow((double)beta2, (double)t)) / (1.0 - pow((double)mom, (double)t));  // optimiser.py:304
  k_adam<<<e2_grid_1d(count, 256, h->sm_count, 8), 256, 0, (cudaStream_t)stream>>>(p, g, m, s, count, lr, mom, beta2, wd,
                                                                                  wd_mult, (float)factor);
  e2_count_launch(h);
  E2_CUDA_CHECK(h, "adam_step");
  return E2_OK;
}

// ---- optimiser step inside a CUDA graph: hyper-parameters and the step counter live in device memory ----
// hyper = [lr, mom, beta2, wd, factor, -, -, -]; e2_adam_prepare advances the counter and refreshes `factor`.
__global__ void k_adam_prepare(float* __restrict__ hyper, int* __restrict__ t_dev) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const int t = t_dev[0] + 1;
    t_dev[0] = t;
    const double beta2 = (double)hyper[2], mom = (double)hyper[1];
    hyper[4] = (float)(sqrt(1.0 - pow(beta2, (double)t)) / (1.0 - pow(mom, (double)t)));   // optimiser.py:304
  }
}

__global__ void __launch_bounds__(256) k_adam_dev(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                  float* __restrict__ s, int64_t n, const float* __restrict__ hyper,
                                                  float wd_mult) {
  const float lr = hyper[0], mom = hyper[1], beta2 = hyper[2], wd = hyper[3], factor = hyper[4];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float mi = m[i], si = s[i];
    const float pi = e2_adam_update(p[i], g[i], mi, si, lr, mom, beta2, wd, factor, wd_mult);
    m[i] = mi, s[i] = si, p[i] = pi;
  }
}

// ---- Adam update of ONE layer's weights + the re-pack of the updated weights, in one pass ----
// The step used to end with e2_adam_step_dev over the flat buffer and then two pack kernels per layer that read the
// weights again.  Here a block owns an (8 output channels x 32 input channels x all taps) brick of w (f_out,f_in,taps):
// it updates the brick (p, m, s read and written once, coalesced: the (c,tap) slab of one o is contiguous), keeps the
// tf32-rounded new weights in shared memory and writes both packed images from there --
//   conv   (mode 0): wf[(o*T + flip(t))*cp + c]   wd[(c*T + t)*op + o]      (e2_conv3d_pack_weights)
//   upconv (mode 1): wf[(t*O + o)*cp + c]         wd[c*op + t*O + o]        (e2_upconv3d_pack_weights, op = row pitch)
// wf in 128-byte runs over c, wd in 32-byte runs over o.  Pad lanes of wf / wd are never written (they hold the zeros of
// the first e2_*_pack_weights call).
struct AdamPackP {
  int O, C, T, kz, kx, ky, cp, op, mode, tf32, ST, SO;
  int vec;   // 16-byte path: every (o, 32-channel block) slab of w starts on a 16-byte boundary and is whole float4s
};
constexpr int AP_OT = 8, AP_CT = 32;

__global__ void __launch_bounds__(256) k_adam_pack(AdamPackP q, E2FastDiv dT, float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ s,
                                                   const float* __restrict__ hyper, float wd_mult, float* __restrict__ wf,
                                                   float* __restrict__ wd) {
  extern __shared__ float slab[];   // [AP_OT][SO]: per o the [nc][ST] slab, ST odd, SO % 32 == 4 (conflict-free phases)
  const float lr = hyper[0], mom = hyper[1], beta2 = hyper[2], wdec = hyper[3], factor = hyper[4];
  const int c0 = blockIdx.x * AP_CT, o0 = blockIdx.y * AP_OT;
  const int nc = min(AP_CT, q.C - c0), no = min(AP_OT, q.O - o0);
  const int n = nc * q.T;
  const int64_t ostride = (int64_t)q.C * q.T;                 // floats between consecutive output channels of w
  const int64_t base0 = ((int64_t)o0 * q.C + c0) * q.T;
  // The update: every thread owns the same offsets r of FOUR output channels at a time and issues all of their loads
  // before the arithmetic (16 independent 16-byte loads in flight; a loop over o with one element per step was bound by
  // memory latency: 0.64 ms per unet3d step instead of 0.34 for the kernels it replaced).
  if (q.vec) {
    for (int r = threadIdx.x * 4; r < n; r += 1024) {
#pragma unroll
      for (int og = 0; og < AP_OT; og += 4) {
        float4 P[4], G[4], M[4], S[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (og + u < no) {
            const int64_t i = base0 + (og + u) * ostride + r;
            P[u] = *reinterpret_cast<const float4*>(p + i), G[u] = __ldg(reinterpret_cast<const float4*>(g + i));
            M[u] = *reinterpret_cast<const float4*>(m + i), S[u] = *reinterpret_cast<const float4*>(s + i);
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (og + u < no) {
            const int64_t i = base0 + (og + u) * ostride + r;
            P[u].x = e2_adam_update(P[u].x, G[u].x, M[u].x, S[u].x, lr, mom, beta2, wdec, factor, wd_mult);
            P[u].y = e2_adam_update(P[u].y, G[u].y, M[u].y, S[u].y, lr, mom, beta2, wdec, factor, wd_mult);
            P[u].z = e2_adam_update(P[u].z, G[u].z, M[u].z, S[u].z, lr, mom, beta2, wdec, factor, wd_mult);
            P[u].w = e2_adam_update(P[u].w, G[u].w, M[u].w, S[u].w, lr, mom, beta2, wdec, factor, wd_mult);
            *reinterpret_cast<float4*>(p + i) = P[u];
            *reinterpret_cast<float4*>(m + i) = M[u];
            *reinterpret_cast<float4*>(s + i) = S[u];
            const float e[4] = {P[u].x, P[u].y, P[u].z, P[u].w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int rr = r + k;
              const int cl = (int)dT.div((uint32_t)rr);
              slab[(og + u) * q.SO + cl * q.ST + (rr - cl * q.T)] = q.tf32 ? e2_round_tf32(e[k]) : e[k];
            }
          }
        }
      }
    }
  } else {
    for (int r = threadIdx.x; r < n; r += 256) {
      const int cl = (int)dT.div((uint32_t)r);
      const int so = cl * q.ST + (r - cl * q.T);
#pragma unroll
      for (int og = 0; og < AP_OT; og += 4) {
        float P[4], G[4], M[4], S[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (og + u < no) {
            const int64_t i = base0 + (og + u) * ostride + r;
            P[u] = p[i], G[u] = __ldg(g + i), M[u] = m[i], S[u] = s[i];
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (og + u < no) {
            const int64_t i = base0 + (og + u) * ostride + r;
            P[u] = e2_adam_update(P[u], G[u], M[u], S[u], lr, mom, beta2, wdec, factor, wd_mult);
            p[i] = P[u], m[i] = M[u], s[i] = S[u];
            slab[(og + u) * q.SO + so] = q.tf32 ? e2_round_tf32(P[u]) : P[u];
          }
        }
      }
    }
  }
  __syncthreads();
  if (wf) {
    const int items = no * q.T * AP_CT;
    for (int i = threadIdx.x; i < items; i += 256) {
      const int cl = i & (AP_CT - 1), rest = i >> 5;
      const int ol = (int)dT.div((uint32_t)rest), t = rest - ol * q.T;
      if (cl >= nc) continue;
      const float v = slab[ol * q.SO + cl * q.ST + t];
      const int o = o0 + ol;
      int64_t row;
      if (q.mode == 0) {
        const int k = t % q.ky, j = (t / q.ky) % q.kx, ii = t / (q.ky * q.kx);
        row = (int64_t)o * q.T + ((q.kz - 1 - ii) * q.kx + (q.kx - 1 - j)) * q.ky + (q.ky - 1 - k);
      } else {
        row = (int64_t)t * q.O + o;
      }
      wf[row * q.cp + c0 + cl] = v;
    }
  }
  if (wd) {
    const int items = nc * q.T * AP_OT;
    for (int i = threadIdx.x; i < items; i += 256) {
      const int ol = i & (AP_OT - 1), rest = i >> 3;
      const int cl = (int)dT.div((uint32_t)rest), t = rest - cl * q.T;
      if (ol >= no) continue;
      const float v = slab[ol * q.SO + cl * q.ST + t];
      const int o = o0 + ol, c = c0 + cl;
      if (q.mode == 0) wd[((int64_t)c * q.T + t) * q.op + o] = v;
      else wd[(int64_t)c * q.op + (int64_t)t * q.O + o] = v;
    }
  }
}

static int launch_adam_pack(e2_handle* h, AdamPackP q, float* w, const float* g, float* m, float* s, const float* hyper,
                            float wd_mult, float* wf, float* wd, cudaStream_t st, const char* what) {
  E2_REQUIRE(h, w && g && m && s && hyper && (wf || wd), "%s: null pointer", what);
  q.ST = q.T | 1;
  q.SO = AP_CT * q.ST;
  q.SO += (36 - (q.SO & 31)) & 31;
  q.vec = ((int64_t)q.C * q.T) % 4 == 0 && (AP_CT * q.T) % 4 == 0 &&
          !((reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
             reinterpret_cast<uintptr_t>(s)) & 15);
  const size_t smem = sizeof(float) * AP_OT * q.SO;
  if (smem > 160 * 1024 || (int64_t)AP_OT * q.T * AP_CT >= (1 << 20) || (q.O + AP_OT - 1) / AP_OT > 65535)
    return e2_fail(h, E2_ERR_UNSUPPORTED, "%s: %d filter taps exceed the fused kernel's brick (use adam_step_dev + pack_weights)",
                   what, q.T);
  if (smem > 48 * 1024 &&
      cudaFuncSetAttribute(k_adam_pack, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024) != cudaSuccess)
    return e2_fail(h, E2_ERR_CUDA, "%s: cudaFuncSetAttribute(max dynamic smem) failed", what);
  dim3 grid((unsigned)((q.C + AP_CT - 1) / AP_CT), (unsigned)((q.O + AP_OT - 1) / AP_OT));
  k_adam_pack<<<grid, 256, smem, st>>>(q, e2_fastdiv((uint32_t)q.T, (uint64_t)AP_OT * q.T * AP_CT), w, g, m, s, hyper, wd_mult,
                                       wf, wd);
  e2_count_launch(h);
  E2_CUDA_CHECK(h, what);
  return E2_OK;
}

extern "C" int e2_conv3d_adam_pack_dev(e2_handle* h, const e2_conv_desc* d, float* w, const float* g, float* m, float* s,
                                       const float* hyper, float wd_mult, float* wf, float* wd, void* stream) {
  E2_REQUIRE(h, d && e2_tensor_ok(&d->x) && e2_tensor_ok(&d->y) && d->kz >= 1 && d->kx >= 1 && d->ky >= 1,
             "conv3d_adam_pack_dev: bad descriptor");
  AdamPackP q;
  memset(&q, 0, sizeof(q));
  q.O = d->y.c, q.C = d->x.c, q.T = d->kz * d->kx * d->ky, q.kz = d->kz, q.kx = d->kx, q.ky = d->ky;
  q.cp = (d->x.c + 3) / 4 * 4, q.op = (d->y.c + 3) / 4 * 4, q.mode = 0, q.tf32 = d->compute == E2_COMPUTE_TF32;
  return launch_adam_pack(h, q, w, g, m, s, hyper, wd_mult, wf, wd, (cudaStream_t)stream, "conv3d_adam_pack_dev");
}

extern "C" int e2_upconv3d_adam_pack_dev(e2_handle* h, const e2_upconv_desc* d, float* w, const float* g, float* m,
                                         float* s, const float* hyper, float wd_mult, float* wf, float* wd, void* stream) {
  E2_REQUIRE(h, d && e2_tensor_ok(&d->x) && e2_tensor_ok(&d->y) && d->pz >= 1 && d->px >= 1 && d->py >= 1,
             "upconv3d_adam_pack_dev: bad descriptor");
  AdamPackP q;
  memset(&q, 0, sizeof(q));
  q.O = d->y.c, q.C = d->x.c, q.T = d->pz * d->px * d->py, q.kz = d->pz, q.kx = d->px, q.ky = d->py;
  q.cp = (d->x.c + 3) / 4 * 4, q.op = (q.T * d->y.c + 3) / 4 * 4, q.mode = 1, q.tf32 = d->compute == E2_COMPUTE_TF32;
  return launch_adam_pack(h, q, w, g, m, s, hyper, wd_mult, wf, wd, (cudaStream_t)stream, "upconv3d_adam_pack_dev");
}

extern "C" int e2_adam_prepare(e2_handle* h, float* hyper, int32_t* t_dev, void* stream) {
  E2_REQUIRE(h, hyper && t_dev, "adam_prepare: null pointer");
  k_adam_prepare<<<1, 32, 0, (cudaStream_t)stream>>>(hyper, t_dev);
  e2_count_launch(h);
  E2_CUDA_CHECK(h, "adam_prepare");
  return E2_OK;
}

extern "C" int e2_adam_step_dev(e2_handle* h, float* p, const float* g, float* m, float* s, int64_t count,
                                const float* hyper, float wd_mult, void* stream) {
  E2_REQUIRE(h, p && g && m && s && hyper && count >= 0, "adam_step_dev: bad arguments");
  if (count == 0) return E2_OK;
  k_adam_dev<<<e2_grid_1d(count, 256, h->sm_count, 8), 256, 0, (cudaStream_t)stream>>>(p, g, m, s, count, hyper, wd_mult);
  e2_count_launch(h);
  E2_CUDA_CHECK(h, "adam_step_dev");
  return E2_OK;
}

__global__ void __launch_bounds__(256) k_sgd(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ d,
                                             int64_t n, float lr, float mom, float wd, float wd_mult) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float nd = g[i] + mom * d[i];  // optimiser.py:148
    float pi = p[i];
    pi = wd_mult != 0.f ? pi - lr * (nd + wd * wd_mult * pi) : pi - lr * nd;
    d[i] = nd, p[i] = pi;
  }
}

extern "C" int e2_sgd_step(e2_handle* h, float* p, const float* g, float* last_dir, int64_t count, float lr, float mom,
                           float wd, float wd_mult, void* stream) {
  E2_REQUIRE(h, p && g && last_dir && count >= 0, "sgd_step: bad arguments");
  if (count == 0) return E2_OK;
  k_sgd<<<e2_grid_1d(count, 256, h->sm_count, 8), 256, 0, (cudaStream_t)stream>>>(p, g, last_dir, count, lr, mom, wd,
                                                                                 wd_mult);
  e2_count_launch(h);
  E2_CUDA_CHECK(h, "sgd_step");
  return E2_OK;
}
