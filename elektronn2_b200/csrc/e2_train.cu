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
__global__ void __launch_bounds__(256) k_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                              float* __restrict__ s, int64_t n, float lr, float mom, float beta2, float wd,
                                              float wd_mult, float factor) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float gi = g[i];
    float nm = mom * m[i] + (1.0f - mom) * gi;             // optimiser.py:306
    float ns = beta2 * s[i] + (1.0f - beta2) * gi * gi;    // :307
    float dir = factor * nm / sqrtf(ns + 1e-5f);           // :309, epsilon inside the sqrt (:283)
    float pi = p[i];
    pi = wd_mult != 0.f ? pi - lr * (dir + wd * wd_mult * pi) : pi - lr * dir;  // :310-319
    m[i] = nm, s[i] = ns, p[i] = pi;
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
    float gi = g[i];
    float nm = mom * m[i] + (1.0f - mom) * gi;
    float ns = beta2 * s[i] + (1.0f - beta2) * gi * gi;
    float dir = factor * nm / sqrtf(ns + 1e-5f);
    float pi = p[i];
    pi = wd_mult != 0.f ? pi - lr * (dir + wd * wd_mult * pi) : pi - lr * dir;
    m[i] = nm, s[i] = ns, p[i] = pi;
  }
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
