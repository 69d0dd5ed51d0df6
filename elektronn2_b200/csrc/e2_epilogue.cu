// Unfused epilogue family: what Conv._make_output / UpConv._make_output do between the
// pooling and the dropout gate when the configuration is not one of the BASELINE ones
// (neural.py:681-712):
//
//     y = act( (gamma / std) * v + b - gamma * mean / std ),       v = pool(conv(x, w))
//
//   * batch normalisation 'train'  : mean / std are batch statistics over all axes but f, std = T.std + 1e-6
//                                    (neural.py:681-685), running averages 0.9995 / 0.0005 (:695-698)
//   * batch normalisation 'predict': mean / std / gamma are stored parameters (:700-703)
//   * 'prelu'                      : b is (f_out, 2), T.nnet.relu(x + b[:,0], alpha = b[:,1])
//                                    (neural.py:655-657, computations.py:83-85)
//   * any activation whose backward needs the pre-activation ('abs')
// plus computations.maxout (computations.py:455-495) and the average / sum pooling modes of
// computations.pooling (:589-590, 600).
//
// All of it is HBM-bound elementwise / per-channel-reduction work on channels-last tensors: a thread block walks
// rows of C contiguous channels, so every global access is a full line; per-channel sums are accumulated in
// double (block partials -> one double atomic per channel and block) so that the statistics do not depend on
// the grid size to more than an ulp of fp32.
#include "e2_common.cuh"

namespace {

constexpr float BN_EPS = 1e-6f;   // neural.py:684

__device__ __forceinline__ float act_fwd(float pre, int act, float alpha) {
  if (act == E2_ACT_PRELU) return pre > 0.f ? pre : alpha * pre;   // T.nnet.relu(x, alpha)
  return e2_apply_act(pre, act);
}

// d act(pre) / d pre
__device__ __forceinline__ float act_deriv(float pre, int act, float alpha) {
  switch (act) {
    case E2_ACT_LIN: return 1.f;
    case E2_ACT_RELU: return pre > 0.f ? 1.f : 0.f;
    case E2_ACT_PRELU: return pre > 0.f ? 1.f : alpha;
    case E2_ACT_TANH: { float t = tanhf(pre); return 1.f - t * t; }
    case E2_ACT_SIGMOID: { float s = 1.f / (1.f + __expf(-pre)); return s * (1.f - s); }
    case E2_ACT_ABS: return pre > 0.f ? 1.f : (pre < 0.f ? -1.f : 0.f);
    case E2_ACT_SOFTPLUS: return 1.f / (1.f + __expf(-pre));
    case E2_ACT_ELU: return pre > 0.f ? 1.f : __expf(pre);
    case E2_ACT_SELU: return 1.0507009873554805f * (pre > 0.f ? 1.f : 1.6732632423543772f * __expf(pre));
    default: return 1.f;
  }
}

// ---------------------------------------------------------------- per-channel reductions
// out[k][c] += sum over positions of f_k(...).  Block = 256 threads = 8 position lanes x 32 channels.
// MODE 0: {v}            MODE 1: {(v - mean)^2}          MODE 2: {dpre, dpre * v, dy * min(pre, 0)}
template <int MODE>
__global__ void __launch_bounds__(256) k_channel_reduce(const float* __restrict__ v, const float* __restrict__ dy,
                                                        int64_t P, int C, int pitch, const float* __restrict__ a0,
                                                        const float* __restrict__ a1, const float* __restrict__ a2,
                                                        int a2_stride, int act, double* __restrict__ out) {
  constexpr int NK = MODE == 2 ? 3 : 1;
  __shared__ double red[NK][8][33];
  const int cl = threadIdx.x & 31, lane = threadIdx.x >> 5;
  const int c = blockIdx.y * 32 + cl;
  double acc[NK];
#pragma unroll
  for (int k = 0; k < NK; ++k) acc[k] = 0.0;
  if (c < C) {
    const float mean = (MODE == 1) ? a0[c] : 0.f;
    const float scale = (MODE == 2 && a0) ? a0[c] : 1.f;
    const float shift = (MODE == 2 && a1) ? a1[c] : 0.f;
    const float alpha = (MODE == 2 && a2) ? a2[(int64_t)c * a2_stride] : 0.f;
    for (int64_t m = (int64_t)blockIdx.x * 8 + lane; m < P; m += (int64_t)gridDim.x * 8) {
      const float x = __ldg(v + m * pitch + c);
      if (MODE == 0) {
        acc[0] += (double)x;
      } else if (MODE == 1) {
        const float d = x - mean;
        acc[0] += (double)d * (double)d;
      } else {
        const float pre = scale * x + shift;
        const float g = __ldg(dy + m * pitch + c);
        const float dpre = g * act_deriv(pre, act, alpha);
        acc[0] += (double)dpre;
        acc[1] += (double)dpre * (double)x;
        acc[2] += (double)g * (double)fminf(pre, 0.f);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < NK; ++k) red[k][lane][cl] = acc[k];
  __syncthreads();
  if (lane == 0 && c < C) {
#pragma unroll
    for (int k = 0; k < NK; ++k) {
      double s = 0.0;
#pragma unroll
      for (int i = 0; i < 8; ++i) s += red[k][i][cl];
      atomicAdd(out + (int64_t)k * C + c, s);
    }
  }
}

dim3 reduce_grid(const e2_handle* h, int64_t P, int C) {
  int gx = (int)((P + 8 * 32 - 1) / (8 * 32));
  const int cap = 4 * h->sm_count;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  return dim3((unsigned)gx, (unsigned)((C + 31) / 32));
}

// sums -> mean ; centred squares -> std (+eps), scale = gamma / std, shift = b - gamma * mean / std
__global__ void k_bn_mean(const double* __restrict__ sum, int C, double inv_n, float* __restrict__ mean) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) mean[c] = (float)(sum[c] * inv_n);
}

__global__ void k_bn_finish(const double* __restrict__ sq, int C, double inv_n, const float* __restrict__ mean,
                            const float* __restrict__ gamma, const float* __restrict__ bias, int bias_stride,
                            float* __restrict__ std_out, float* __restrict__ scale, float* __restrict__ shift,
                            float* __restrict__ run_mean, float* __restrict__ run_std, float keep) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float sd = (float)sqrt(sq[c] * inv_n) + BN_EPS;          // T.std (population) + 1e-6, neural.py:683-684
  const float g = gamma ? gamma[c] : 1.f, b = bias ? bias[(int64_t)c * bias_stride] : 0.f, m = mean[c];
  std_out[c] = sd;
  scale[c] = g / sd;
  shift[c] = b - g * m / sd;                                     // neural.py:711
  if (run_mean) run_mean[c] = keep * run_mean[c] + (1.f - keep) * m;    // neural.py:695-696
  if (run_std) run_std[c] = keep * run_std[c] + (1.f - keep) * sd;      // neural.py:697-698
}

__global__ void k_bn_fold(int C, const float* __restrict__ gamma, const float* __restrict__ bias, int bias_stride,
                          const float* __restrict__ mean, const float* __restrict__ sd, float* __restrict__ scale,
                          float* __restrict__ shift) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float g = gamma ? gamma[c] : 1.f, b = bias ? bias[(int64_t)c * bias_stride] : 0.f;
  const float m = mean ? mean[c] : 0.f, s = sd ? sd[c] : 1.f;
  scale[c] = g / s;
  shift[c] = b - g * m / s;
}

// ------------------------------------------------------------------- elementwise passes
__global__ void __launch_bounds__(256) k_affine_act_fwd(const float* __restrict__ v, float* __restrict__ y, int64_t P,
                                                        int C, int pitch, const float* __restrict__ scale,
                                                        const float* __restrict__ shift, const float* __restrict__ alpha,
                                                        int alpha_stride, int act, int round_tf32) {
  const int64_t total = P * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int64_t ofs = (i / C) * pitch + c;
    const float pre = (scale ? scale[c] : 1.f) * v[ofs] + (shift ? shift[c] : 0.f);
    float r = act_fwd(pre, act, alpha ? alpha[(int64_t)c * alpha_stride] : 0.f);
    y[ofs] = round_tf32 ? e2_round_tf32(r) : r;
  }
}

// dv = scale * (dpre - k1[c] - vhat * k2[c]),  vhat = (v - mean) / std   (k1 = k2 = 0 without batch statistics)
__global__ void __launch_bounds__(256) k_affine_act_bwd(const float* __restrict__ v, const float* __restrict__ dy,
                                                        float* __restrict__ dv, int64_t P, int C, int pitch,
                                                        const float* __restrict__ scale, const float* __restrict__ shift,
                                                        const float* __restrict__ alpha, int alpha_stride, int act,
                                                        const float* __restrict__ mean, const float* __restrict__ sd,
                                                        const float* __restrict__ k12, int round_tf32) {
  const int64_t total = P * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int64_t ofs = (i / C) * pitch + c;
    const float sc = scale ? scale[c] : 1.f;
    const float x = v[ofs];
    const float pre = sc * x + (shift ? shift[c] : 0.f);
    float d = dy[ofs] * act_deriv(pre, act, alpha ? alpha[(int64_t)c * alpha_stride] : 0.f);
    if (k12) d -= k12[c] + (x - mean[c]) / sd[c] * k12[C + c];
    d *= sc;
    dv[ofs] = round_tf32 ? e2_round_tf32(d) : d;
  }
}

// parameter gradients and the two batch-statistics constants from the MODE-2 sums
//   red = {S0 = sum dpre, S1 = sum dpre * v, S2 = sum dy * min(pre, 0)}
__global__ void k_affine_param_grads(const double* __restrict__ red, int C, double inv_n, const float* __restrict__ mean,
                                     const float* __restrict__ sd, int batch_stats, float* __restrict__ dgamma,
                                     float* __restrict__ dbias, int dbias_stride, float* __restrict__ dalpha,
                                     int dalpha_stride, float* __restrict__ k12) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double S0 = red[c], S1 = red[C + c], S2 = red[2 * C + c];
  if (dbias) dbias[(int64_t)c * dbias_stride] = (float)S0;
  if (dalpha) dalpha[(int64_t)c * dalpha_stride] = (float)S2;
  double dg = 0.0;
  if (mean && sd) dg = (S1 - (double)mean[c] * S0) / (double)sd[c];          // sum dpre * vhat
  if (dgamma) dgamma[c] = (float)dg;
  if (k12) {
    if (batch_stats) {
      // y = gamma * (v - mu) / s + b with s = sigma + eps:  dv_i = (gamma/s) [dpre_i - mean(dpre) - vhat_i (s/sigma) mean(dpre vhat)]
      const double s = (double)sd[c], sigma = s - (double)BN_EPS;
      k12[c] = (float)(S0 * inv_n);
      k12[C + c] = (float)(sigma > 0.0 ? (s / sigma) * dg * inv_n : 0.0);
    } else {
      k12[c] = 0.f, k12[C + c] = 0.f;
    }
  }
}

// ------------------------------------------------------------------ average / sum pooling
struct APool {
  int n, Z, X, Y, C, xp, Zo, Xo, Yo, yp, pz, px, py, accumulate;
  float w;   // 1 / prod(pool) (average) or 1 (sum)
};

__global__ void __launch_bounds__(256) k_avgpool_fwd(APool p, const float* __restrict__ x, float* __restrict__ y) {
  const int64_t total = (int64_t)p.n * p.Zo * p.Xo * p.Yo * p.C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % p.C);
    int64_t pos = i / p.C;
    const int yo = (int)(pos % p.Yo);
    int64_t t = pos / p.Yo;
    const int xo = (int)(t % p.Xo);
    t /= p.Xo;
    const int zo = (int)(t % p.Zo);
    const int n = (int)(t / p.Zo);
    float s = 0.f;
    for (int dz = 0; dz < p.pz; ++dz)
      for (int dx = 0; dx < p.px; ++dx)
        for (int dy = 0; dy < p.py; ++dy) {
          const int64_t lin = (((int64_t)n * p.Z + zo * p.pz + dz) * p.X + xo * p.px + dx) * p.Y + yo * p.py + dy;
          s += __ldg(x + lin * p.xp + c);
        }
    y[pos * p.yp + c] = s * p.w;
  }
}

__global__ void __launch_bounds__(256) k_avgpool_bwd(APool p, const float* __restrict__ dy, float* __restrict__ dx) {
  const int64_t total = (int64_t)p.n * p.Z * p.X * p.Y * p.C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % p.C);
    int64_t pos = i / p.C;
    const int yy = (int)(pos % p.Y);
    int64_t t = pos / p.Y;
    const int xx = (int)(t % p.X);
    t /= p.X;
    const int zz = (int)(t % p.Z);
    const int n = (int)(t / p.Z);
    const int64_t opos = (((int64_t)n * p.Zo + zz / p.pz) * p.Xo + xx / p.px) * p.Yo + yy / p.py;
    const float g = __ldg(dy + opos * p.yp + c) * p.w;
    const int64_t ofs = pos * p.xp + c;
    dx[ofs] = p.accumulate ? dx[ofs] + g : g;
  }
}

// ------------------------------------------------------------------------------- maxout
// y[o, c, i] = max_k x[o, c * factor + k, i]   on a dense (outer, F, inner) array (computations.py:481-493)
__global__ void __launch_bounds__(256) k_maxout_fwd(const float* __restrict__ x, float* __restrict__ y, int64_t outer,
                                                    int Fo, int64_t inner, int factor) {
  const int64_t total = outer * Fo * inner;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t in = i % inner;
    int64_t t = i / inner;
    const int c = (int)(t % Fo);
    const int64_t o = t / Fo;
    const float* src = x + ((o * Fo + c) * factor) * inner + in;
    float m = src[0];
    for (int k = 1; k < factor; ++k) m = fmaxf(m, src[(int64_t)k * inner]);   // T.maximum
    y[i] = m;
  }
}

// gradient of the chain of T.maximum: routed to the first maximal slice (ties measure-zero on continuous data)
__global__ void __launch_bounds__(256) k_maxout_bwd(const float* __restrict__ x, const float* __restrict__ dy,
                                                    float* __restrict__ dx, int64_t outer, int Fo, int64_t inner,
                                                    int factor) {
  const int64_t total = outer * Fo * inner;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t in = i % inner;
    int64_t t = i / inner;
    const int c = (int)(t % Fo);
    const int64_t o = t / Fo;
    const int64_t base = ((o * Fo + c) * factor) * inner + in;
    int best = 0;
    float m = x[base];
    for (int k = 1; k < factor; ++k) {
      const float v = x[base + (int64_t)k * inner];
      if (v > m) m = v, best = k;
    }
    const float g = dy[i];
    for (int k = 0; k < factor; ++k) dx[base + (int64_t)k * inner] = (k == best) ? g : 0.f;
  }
}

int check_affine(e2_handle* h, const e2_affine_desc* d) {
  E2_REQUIRE(h, d && e2_tensor_ok(&d->t), "affine_act: bad descriptor");
  E2_REQUIRE(h, d->act >= E2_ACT_LIN && d->act <= E2_ACT_PRELU, "affine_act: unknown activation %d", d ? d->act : -1);
  E2_REQUIRE(h, d->param_stride >= 1, "affine_act: param_stride must be >= 1");
  return E2_OK;
}

}  // namespace

// ======================================================================== C ABI
extern "C" int e2_bn_batch_stats(e2_handle* h, const e2_tensor* t, const float* v, const float* gamma, const float* bias,
                                 int32_t bias_stride, float* mean, float* std_out, float* scale, float* shift,
                                 float* run_mean, float* run_std, float keep, double* scratch, void* stream) {
  E2_REQUIRE(h, e2_tensor_ok(t) && v && mean && std_out && scale && shift && scratch, "bn_batch_stats: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t P = e2_positions(t);
  const int C = t->c;
  cudaMemsetAsync(scratch, 0, sizeof(double) * 2 * C, s);
  const dim3 grid = reduce_grid(h, P, C);
  k_channel_reduce<0><<<grid, 256, 0, s>>>(v, nullptr, P, C, t->c_pitch, nullptr, nullptr, nullptr, 1, 0, scratch);
  k_bn_mean<<<(C + 127) / 128, 128, 0, s>>>(scratch, C, 1.0 / (double)P, mean);
  k_channel_reduce<1><<<grid, 256, 0, s>>>(v, nullptr, P, C, t->c_pitch, mean, nullptr, nullptr, 1, 0, scratch + C);
  k_bn_finish<<<(C + 127) / 128, 128, 0, s>>>(scratch + C, C, 1.0 / (double)P, mean, gamma, bias, bias_stride > 0 ? bias_stride : 1,
                                              std_out, scale, shift, run_mean, run_std, keep);
  for (int i = 0; i < 4; ++i) e2_count_launch(h);
  E2_CUDA_CHECK(h, "bn_batch_stats");
  return E2_OK;
}

extern "C" int e2_bn_fold(e2_handle* h, int32_t C, const float* gamma, const float* bias, int32_t bias_stride,
                          const float* mean, const float* std_in, float* scale, float* shift, void* stream) {
  E2_REQUIRE(h, C > 0 && scale && shift, "bn_fold: bad arguments");
  k_bn_fold<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(C, gamma, bias, bias_stride > 0 ? bias_stride : 1, mean, std_in,
                                                               scale, shift);
  e2_count_launch(h);
  E2_CUDA_CHECK(h, "bn_fold");
  return E2_OK;
}

extern "C" int e2_affine_act_fwd(e2_handle* h, const e2_affine_desc* d, const float* v, const float* scale,
                                 const float* shift, const float* alpha, float* y, void* stream) {
  int rc = check_affine(h, d);
  if (rc) return rc;
  E2_REQUIRE(h, v && y && (d->act != E2_ACT_PRELU || alpha), "affine_act_fwd: null pointer");
  const int64_t P = e2_positions(&d->t);
  k_affine_act_fwd<<<e2_grid_1d(P * d->t.c, 256, h->sm_count), 256, 0, (cudaStream_t)stream>>>(
      v, y, P, d->t.c, d->t.c_pitch, scale, shift, alpha, d->param_stride, d->act, d->round_tf32);
  e2_count_launch(h);
  E2_CUDA_CHECK(h, "affine_act_fwd");
  return E2_OK;
}

extern "C" int e2_affine_act_bwd(e2_handle* h, const e2_affine_desc* d, const float* v, const float* dy,
                                 const float* scale, const float* shift, const float* alpha, const float* mean,
                                 const float* std_in, float* dv, float* dgamma, float* dbias, float* dalpha,
                                 double* scratch, void* stream) {
  int rc = check_affine(h, d);
  if (rc) return rc;
  E2_REQUIRE(h, v && dy && dv && scratch && (d->act != E2_ACT_PRELU || alpha), "affine_act_bwd: null pointer");
  E2_REQUIRE(h, !d->batch_stats || (mean && std_in), "affine_act_bwd: batch statistics need mean and std");
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t P = e2_positions(&d->t);
  const int C = d->t.c;
  // scratch: double[3*C] sums, then float[2*C] k1 / k2
  float* k12 = reinterpret_cast<float*>(scratch + 3 * (size_t)C);
  cudaMemsetAsync(scratch, 0, sizeof(double) * 3 * C, s);
  k_channel_reduce<2><<<reduce_grid(h, P, C), 256, 0, s>>>(v, dy, P, C, d->t.c_pitch, scale, shift, alpha, d->param_stride,
                                                          d->act, scratch);
  k_affine_param_grads<<<(C + 127) / 128, 128, 0, s>>>(scratch, C, 1.0 / (double)P, mean, std_in, d->batch_stats, dgamma, dbias,
                                                       d->param_stride, dalpha, d->param_stride, d->batch_stats ? k12 : nullptr);
  k_affine_act_bwd<<<e2_grid_1d(P * C, 256, h->sm_count), 256, 0, s>>>(v, dy, dv, P, C, d->t.c_pitch, scale, shift, alpha,
                                                                      d->param_stride, d->act, mean, std_in,
                                                                      d->batch_stats ? k12 : nullptr, d->round_tf32);
  for (int i = 0; i < 3; ++i) e2_count_launch(h);
  E2_CUDA_CHECK(h, "affine_act_bwd");
  return E2_OK;
}

extern "C" int e2_affine_scratch_bytes(int32_t C, size_t* bytes) {
  if (C <= 0 || !bytes) return E2_ERR_INVALID;
  *bytes = sizeof(double) * 3 * (size_t)C + sizeof(float) * 2 * (size_t)C;
  return E2_OK;
}

int e2_launch_avgpool_fwd(e2_handle* h, const e2_pool_desc* d, const float* x, float* y, cudaStream_t s) {
  APool p = {d->x.n, d->x.z, d->x.x, d->x.y, d->x.c, d->x.c_pitch, d->y.z, d->y.x, d->y.y, d->y.c_pitch,
             d->pz, d->px, d->py, 0, d->mode == E2_POOL_SUM ? 1.f : 1.f / (float)(d->pz * d->px * d->py)};
  k_avgpool_fwd<<<e2_grid_1d(e2_positions(&d->y) * d->y.c, 256, h->sm_count), 256, 0, s>>>(p, x, y);
  e2_count_launch(h);
  E2_CUDA_CHECK(h, "avgpool_fwd");
  return E2_OK;
}

int e2_launch_avgpool_bwd(e2_handle* h, const e2_pool_desc* d, const float* dy, float* dx, cudaStream_t s) {
  APool p = {d->x.n, d->x.z, d->x.x, d->x.y, d->x.c, d->x.c_pitch, d->y.z, d->y.x, d->y.y, d->y.c_pitch,
             d->pz, d->px, d->py, d->accumulate, d->mode == E2_POOL_SUM ? 1.f : 1.f / (float)(d->pz * d->px * d->py)};
  k_avgpool_bwd<<<e2_grid_1d(e2_positions(&d->x) * d->x.c, 256, h->sm_count), 256, 0, s>>>(p, dy, dx);
  e2_count_launch(h);
  E2_CUDA_CHECK(h, "avgpool_bwd");
  return E2_OK;
}

extern "C" int e2_maxout_fwd(e2_handle* h, const float* x, float* y, int64_t outer, int32_t f_out, int64_t inner,
                             int32_t factor, void* stream) {
  E2_REQUIRE(h, x && y && outer > 0 && f_out > 0 && inner > 0 && factor >= 1, "maxout_fwd: bad arguments");
  k_maxout_fwd<<<e2_grid_1d(outer * f_out * inner, 256, h->sm_count), 256, 0, (cudaStream_t)stream>>>(x, y, outer, f_out, inner,
                                                                                                     factor);
  e2_count_launch(h);
  E2_CUDA_CHECK(h, "maxout_fwd");
  return E2_OK;
}

extern "C" int e2_maxout_bwd(e2_handle* h, const float* x, const float* dy, float* dx, int64_t outer, int32_t f_out,
                             int64_t inner, int32_t factor, void* stream) {
  E2_REQUIRE(h, x && dy && dx && outer > 0 && f_out > 0 && inner > 0 && factor >= 1, "maxout_bwd: bad arguments");
  k_maxout_bwd<<<e2_grid_1d(outer * f_out * inner, 256, h->sm_count), 256, 0, (cudaStream_t)stream>>>(x, dy, dx, outer, f_out,
                                                                                                     inner, factor);
  e2_count_launch(h);
  E2_CUDA_CHECK(h, "maxout_bwd");
  return E2_OK;
}
