// CUDA-core (exact fp32 FFMA) implementations of the conv-shaped ops, plus weight
// packing, activation backward and bias gradient.  These are the E2_COMPUTE_F32
// algorithm and also serve the layers that are not tensor-core shaped (c_in == 1,
// c_out == 2; SURVEY.md 8d "report as HBM-bound").  The tcgen05 path (e2_conv_tc.cu)
// takes exactly the same operands.
//
// Two GEMM-shaped kernels cover everything:
//   gather-GEMM   C[m, n]     = sum_{tap,k} A[pos(m)*s + tap + org][k] * B[n][tap][k]
//                 conv fwd (A=x), conv dgrad (A=dy, org=-(k-1)), upconv dgrad (s=p),
//                 upconv fwd (taps=1, N=(tap,o), pixel-shuffle epilogue)
//   reduce-GEMM   W[r][tap][s] = sum_m P[m][r] * Q[pos(m)*st + tap + org][s]
//                 conv wgrad (P=dy,Q=x), upconv wgrad (P=x,Q=dy,st=p)
#include <stdlib.h>
#include <algorithm>
#include "e2_common.cuh"
#include "e2_conv_internal.cuh"

// ------------------------------------------------------------------ weight packing
// wf[o][tap][c_pitch]  : tap index t=(i,j,k) addresses x[pos + t]; holds w[o][c][kz-1-i,kx-1-j,ky-1-k]
// wd[c][tap][o_pitch]  : tap index t' addresses dy[pos - (k-1) + t']; holds w[o][c][t'] (unflipped)
__global__ void __launch_bounds__(256) k_pack_conv(const float* __restrict__ w, float* __restrict__ wf,
                                                   float* __restrict__ wd, int O, int C, int kz, int kx, int ky, int cp,
                                                   int op, int tf32) {
  const int T = kz * kx * ky;
  const int64_t total = (int64_t)O * C * T;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int t = (int)(i % T);
    int c = (int)((i / T) % C);
    int o = (int)(i / ((int64_t)T * C));
    float v = w[i];
    if (tf32) v = e2_round_tf32(v);
    int k = t % ky, j = (t / ky) % kx, ii = t / (ky * kx);
    int tflip = ((kz - 1 - ii) * kx + (kx - 1 - j)) * ky + (ky - 1 - k);
    if (wf) wf[((int64_t)o * T + tflip) * cp + c] = v;
    if (wd) wd[((int64_t)c * T + t) * op + o] = v;
  }
}

// Tiled versions of the conv packing (the element-wise kernel above scatters 4-byte stores at pitch stride:
// 0.24 ms per unet3d step).  Both are shared-memory transposes with coalesced reads and writes.
//   wd: the [O][C*T] matrix transposed to [C*T][op]
__global__ void __launch_bounds__(256) k_pack_conv_wd(const float* __restrict__ w, float* __restrict__ wd, int O, int M,
                                                      int op, int tf32) {
  __shared__ float tile[32][33];
  const int m0 = blockIdx.x * 32, o0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int j = ty; j < 32; j += 8) {
    const int o = o0 + j, m = m0 + tx;
    float v = (o < O && m < M) ? w[(int64_t)o * M + m] : 0.f;
    tile[j][tx] = tf32 ? e2_round_tf32(v) : v;
  }
  __syncthreads();
#pragma unroll
  for (int j = ty; j < 32; j += 8) {
    const int m = m0 + j, o = o0 + tx;
    if (m < M && o < O) wd[(int64_t)m * op + o] = tile[tx][j];
  }
}
//   wf: per output channel the [C][T] slab transposed to [flip(T)][cp]; one block = one o x WF_CH input channels
constexpr int WF_CH = 128;
__global__ void __launch_bounds__(256) k_pack_conv_wf(const float* __restrict__ w, float* __restrict__ wf, int C, int kz,
                                                      int kx, int ky, int cp, int tf32, E2FastDiv dT) {
  extern __shared__ float slab[];   // [WF_CH][T + 1]
  const int T = kz * kx * ky;
  const int o = blockIdx.y, c0 = blockIdx.x * WF_CH;
  const int nc = min(WF_CH, C - c0);
  const float* src = w + ((int64_t)o * C + c0) * T;
  const int n = nc * T;
  if ((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (n & 3) == 0) {
    for (int i4 = threadIdx.x; i4 < n / 4; i4 += 256) {
      const float4 v4 = __ldg(reinterpret_cast<const float4*>(src) + i4);
      const float v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int i = i4 * 4 + e;
        const int c = (int)dT.div((uint32_t)i);
        slab[c * (T + 1) + (i - c * T)] = tf32 ? e2_round_tf32(v[e]) : v[e];
      }
    }
  } else {
    for (int i = threadIdx.x; i < n; i += 256) {
      const float v = src[i];
      const int c = (int)dT.div((uint32_t)i);
      slab[c * (T + 1) + (i - c * T)] = tf32 ? e2_round_tf32(v) : v;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < T * WF_CH; i += 256) {
    const int t = i / WF_CH, c = i % WF_CH;
    if (c >= nc) continue;
    const int k = t % ky, j = (t / ky) % kx, ii = t / (ky * kx);
    const int tflip = ((kz - 1 - ii) * kx + (kx - 1 - j)) * ky + (ky - 1 - k);
    wf[((int64_t)o * T + tflip) * cp + c0 + c] = slab[c * (T + 1) + t];
  }
}

// upconv: wf[(tap,o)][c_pitch] = w[o][c][tap];  wd[c][(tap,o) pitch] = w[o][c][tap]
__global__ void __launch_bounds__(256) k_pack_upconv(const float* __restrict__ w, float* __restrict__ wf,
                                                     float* __restrict__ wd, int O, int C, int T, int cp, int np,
                                                     int tf32) {
  const int64_t total = (int64_t)O * C * T;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int t = (int)(i % T);
    int c = (int)((i / T) % C);
    int o = (int)(i / ((int64_t)T * C));
    float v = w[i];
    if (tf32) v = e2_round_tf32(v);
    if (wf) wf[((int64_t)t * O + o) * cp + c] = v;
    if (wd) wd[(int64_t)c * np + t * O + o] = v;
  }
}

static inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

extern "C" int e2_conv3d_packed_floats(const e2_conv_desc* d, size_t* fwd_floats, size_t* dgrad_floats) {
  if (!d) return E2_ERR_INVALID;
  int T = d->kz * d->kx * d->ky;
  if (fwd_floats) *fwd_floats = (size_t)d->y.c * T * round_up(d->x.c, 4);
  if (dgrad_floats) *dgrad_floats = (size_t)d->x.c * T * round_up(d->y.c, 4);
  return E2_OK;
}

static int check_conv(e2_handle* h, const e2_conv_desc* d) {
  E2_REQUIRE(h, d && e2_tensor_ok(&d->x) && e2_tensor_ok(&d->y), "conv3d: bad tensor descriptor");
  E2_REQUIRE(h, d->kz >= 1 && d->kx >= 1 && d->ky >= 1, "conv3d: filter extents must be >= 1");
  // 'valid' only, stride 1 only (computations.py:375-376 raises for strides)
  E2_REQUIRE(h, d->y.n == d->x.n && d->y.z == d->x.z - d->kz + 1 && d->y.x == d->x.x - d->kx + 1 &&
                    d->y.y == d->x.y - d->ky + 1,
             "conv3d: output extents (%d,%d,%d) != input (%d,%d,%d) - filter (%d,%d,%d) + 1", d->y.z, d->y.x, d->y.y,
             d->x.z, d->x.x, d->x.y, d->kz, d->kx, d->ky);
  E2_REQUIRE(h, d->compute == E2_COMPUTE_F32 || d->compute == E2_COMPUTE_TF32, "conv3d: unsupported compute type %d",
             d->compute);
  E2_REQUIRE(h, e2_positions(&d->x) * d->x.c_pitch < (1ll << 40), "conv3d: tensor too large");
  return E2_OK;
}

extern "C" int e2_conv3d_pack_weights(e2_handle* h, const e2_conv_desc* d, const float* w, float* wf, float* wd,
                                      void* stream) {
  int rc = check_conv(h, d);
  if (rc) return rc;
  E2_REQUIRE(h, w && (wf || wd), "conv3d_pack_weights: null pointer");
  int T = d->kz * d->kx * d->ky, cp = round_up(d->x.c, 4), op = round_up(d->y.c, 4);
  cudaStream_t s = (cudaStream_t)stream;
  // pad lanes must be finite zeros (they are multiplied into accumulators)
  if (wf && cp != d->x.c) cudaMemsetAsync(wf, 0, sizeof(float) * (size_t)d->y.c * T * cp, s);
  if (wd && op != d->y.c) cudaMemsetAsync(wd, 0, sizeof(float) * (size_t)d->x.c * T * op, s);
  int64_t total = (int64_t)d->y.c * d->x.c * T;
  const int tf32 = d->compute == E2_COMPUTE_TF32;
  if (total >= 4096 && T <= 343 && d->y.c <= 65535) {
    const int M = d->x.c * T;
    if (wf) {
      dim3 grid((unsigned)((d->x.c + WF_CH - 1) / WF_CH), (unsigned)d->y.c);
      const size_t slab_bytes = sizeof(float) * WF_CH * (T + 1);
      if (slab_bytes > 48 * 1024) {
        static bool configured = false;
        if (!configured) cudaFuncSetAttribute(k_pack_conv_wf, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        configured = true;
      }
      k_pack_conv_wf<<<grid, 256, slab_bytes, s>>>(w, wf, d->x.c, d->kz, d->kx, d->ky, cp, tf32,
                                                  e2_fastdiv((uint32_t)T, (uint64_t)WF_CH * T));
      e2_count_launch(h);
    }
    if (wd) {
      dim3 grid((unsigned)((M + 31) / 32), (unsigned)((d->y.c + 31) / 32));
      k_pack_conv_wd<<<grid, 256, 0, s>>>(w, wd, d->y.c, M, op, tf32);
      e2_count_launch(h);
    }
  } else {
    k_pack_conv<<<e2_grid_1d(total, 256, h->sm_count), 256, 0, s>>>(w, wf, wd, d->y.c, d->x.c, d->kz, d->kx, d->ky, cp, op, tf32);
    e2_count_launch(h);
  }
  E2_CUDA_CHECK(h, "conv3d_pack_weights");
  return E2_OK;
}

static int check_upconv(e2_handle* h, const e2_upconv_desc* d) {
  E2_REQUIRE(h, d && e2_tensor_ok(&d->x) && e2_tensor_ok(&d->y), "upconv3d: bad tensor descriptor");
  E2_REQUIRE(h, d->pz >= 1 && d->px >= 1 && d->py >= 1, "upconv3d: pool factors must be >= 1");
  E2_REQUIRE(h, d->y.n == d->x.n && d->y.z == d->x.z * d->pz && d->y.x == d->x.x * d->px && d->y.y == d->x.y * d->py,
             "upconv3d: output extents must be input * pool");
  E2_REQUIRE(h, d->compute == E2_COMPUTE_F32 || d->compute == E2_COMPUTE_TF32, "upconv3d: unsupported compute type %d",
             d->compute);
  return E2_OK;
}

extern "C" int e2_upconv3d_packed_floats(const e2_upconv_desc* d, size_t* fwd_floats, size_t* dgrad_floats) {
  if (!d) return E2_ERR_INVALID;
  int T = d->pz * d->px * d->py;
  if (fwd_floats) *fwd_floats = (size_t)T * d->y.c * round_up(d->x.c, 4);
  if (dgrad_floats) *dgrad_floats = (size_t)d->x.c * round_up(T * d->y.c, 4);
  return E2_OK;
}

extern "C" int e2_upconv3d_pack_weights(e2_handle* h, const e2_upconv_desc* d, const float* w, float* wf, float* wd,
                                        void* stream) {
  int rc = check_upconv(h, d);
  if (rc) return rc;
  E2_REQUIRE(h, w && (wf || wd), "upconv3d_pack_weights: null pointer");
  int T = d->pz * d->px * d->py, cp = round_up(d->x.c, 4), np = round_up(T * d->y.c, 4);
  cudaStream_t s = (cudaStream_t)stream;
  if (wf && cp != d->x.c) cudaMemsetAsync(wf, 0, sizeof(float) * (size_t)T * d->y.c * cp, s);
  if (wd && np != T * d->y.c) cudaMemsetAsync(wd, 0, sizeof(float) * (size_t)d->x.c * np, s);
  int64_t total = (int64_t)d->y.c * d->x.c * T;
  k_pack_upconv<<<e2_grid_1d(total, 256, h->sm_count), 256, 0, s>>>(w, wf, wd, d->y.c, d->x.c, T, cp, np,
                                                                    d->compute == E2_COMPUTE_TF32);
  e2_count_launch(h);
  E2_CUDA_CHECK(h, "upconv3d_pack_weights");
  return E2_OK;
}

// -------------------------------------------------------------------- gather-GEMM
// 64x64 output tile, K chunks of 16, 256 threads, 4x4 register micro-tile.
constexpr int GBM = 64, GBN = 64, GBK = 16;

__global__ void __launch_bounds__(256) k_gather_gemm(GatherGemm g) {
  __shared__ __align__(16) float As[GBK][GBM + 4];
  __shared__ __align__(16) float Bs[GBK][GBN + 4];
  const int tid = threadIdx.x;
  const int64_t M = (int64_t)g.On * g.Oz * g.Ox * g.Oy;
  const int64_t m0 = (int64_t)blockIdx.x * GBM;
  const int n0 = blockIdx.y * GBN;
  // loader role: one row, 4 consecutive k
  const int lrow = tid >> 2, lk = (tid & 3) * 4;
  int64_t lm = m0 + lrow;
  bool lvalid = lm < M;
  int ly = 0, lx = 0, lz = 0, ln = 0;
  if (lvalid) {
    ly = (int)(lm % g.Oy);
    int64_t t = lm / g.Oy;
    lx = (int)(t % g.Ox);
    t /= g.Ox;
    lz = (int)(t % g.Oz);
    ln = (int)(t / g.Oz);
  }
  const int bn = n0 + lrow;
  const bool bvalid = bn < g.N;
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int T = g.tz * g.tx * g.ty;
  for (int tap = 0; tap < T; ++tap) {
    int k3 = tap % g.ty, j3 = (tap / g.ty) % g.tx, i3 = tap / (g.ty * g.tx);
    int az = lz * g.sz + g.oz + i3, ax = lx * g.sx + g.ox + j3, ay = ly * g.sy + g.oy + k3;
    bool ain = lvalid && az >= 0 && az < g.Az && ax >= 0 && ax < g.Ax && ay >= 0 && ay < g.Ay;
    const float* arow = g.A + ((((int64_t)ln * g.Az + az) * g.Ax + ax) * g.Ay + ay) * g.a_pitch;
    const float* brow = g.B + (int64_t)bn * g.b_row + (int64_t)tap * g.b_tap;
    for (int k0 = 0; k0 < g.K; k0 += GBK) {
      float av[4], bv[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        int k = k0 + lk + q;
        av[q] = (ain && k < g.K) ? __ldg(arow + k) : 0.f;
        bv[q] = (bvalid && k < g.K) ? __ldg(brow + k) : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        As[lk + q][lrow] = av[q];
        Bs[lk + q][lrow] = bv[q];
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < GBK; ++kk) {
        float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
        float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
        float aa[4] = {a.x, a.y, a.z, a.w}, bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
      }
    }
  }
  // epilogue
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
    int oy = (int)(m % g.Oy);
    int64_t t = m / g.Oy;
    int ox = (int)(t % g.Ox);
    t /= g.Ox;
    int oz = (int)(t % g.Oz);
    int on = (int)(t / g.Oz);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= g.N) continue;
      float v = acc[i][j];
      int64_t ofs;
      int ch;
      if (g.shuffle) {
        int tp = n / g.Fo;
        ch = n - tp * g.Fo;
        int k3 = tp % g.py, j3 = (tp / g.py) % g.px, i3 = tp / (g.py * g.px);
        ofs = ((((int64_t)on * (g.Oz * g.pz) + oz * g.pz + i3) * (g.Ox * g.px) + ox * g.px + j3) * (g.Oy * g.py) +
               oy * g.py + k3) * g.c_pitch + ch;
      } else {
        ch = n;
        ofs = m * g.c_pitch + n;
      }
      if (g.bias) v += __ldg(g.bias + ch);
      v = e2_apply_act(v, g.act);
      if (g.gate && !(__ldg(g.gate + ofs) > 0.f)) v = 0.f;
      if (g.accumulate) v += g.C[ofs];
      if (g.round_tf32) v = e2_round_tf32(v);
      g.C[ofs] = v;
    }
  }
}

int e2_launch_gather_gemm_ffma(e2_handle* h, const GatherGemm& g, cudaStream_t s) {
  int64_t M = (int64_t)g.On * g.Oz * g.Ox * g.Oy;
  dim3 grid((unsigned)((M + GBM - 1) / GBM), (unsigned)((g.N + GBN - 1) / GBN));
  k_gather_gemm<<<grid, 256, 0, s>>>(g);
  e2_count_launch(h);
  E2_CUDA_CHECK(h, "gather_gemm_ffma");
  return E2_OK;
}

// ---------------------------------------------------- first layer (c_in == 1) forward
// K per tap is 1, so the GEMM tile above would waste 15/16 of its k-loop.  One thread
// computes NB output channels of one position; taps*NB weights live in shared memory.
template <int NB>
__global__ void __launch_bounds__(128) k_conv_c1_fwd(GatherGemm g) {
  extern __shared__ float wsm[];  // [T][NB]
  const int T = g.tz * g.tx * g.ty;
  const int n0 = blockIdx.y * NB;
  for (int i = threadIdx.x; i < T * NB; i += blockDim.x) {
    int t = i / NB, j = i % NB;
    wsm[i] = (n0 + j < g.N) ? g.B[(int64_t)(n0 + j) * g.b_row + (int64_t)t * g.b_tap] : 0.f;
  }
  __syncthreads();
  const int64_t M = (int64_t)g.On * g.Oz * g.Ox * g.Oy;
  for (int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < M; m += (int64_t)gridDim.x * blockDim.x) {
    int oy = (int)(m % g.Oy);
    int64_t t = m / g.Oy;
    int ox = (int)(t % g.Ox);
    t /= g.Ox;
    int oz = (int)(t % g.Oz);
    int on = (int)(t / g.Oz);
    float acc[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) acc[j] = 0.f;
    int tap = 0;
    for (int i3 = 0; i3 < g.tz; ++i3)
      for (int j3 = 0; j3 < g.tx; ++j3)
        for (int k3 = 0; k3 < g.ty; ++k3, ++tap) {
          float xv = __ldg(g.A + ((((int64_t)on * g.Az + oz + i3) * g.Ax + ox + j3) * g.Ay + oy + k3) * g.a_pitch);
          const float* wr = wsm + tap * NB;
#pragma unroll
          for (int j = 0; j < NB; ++j) acc[j] = fmaf(xv, wr[j], acc[j]);
        }
    float* out = g.C + m * g.c_pitch + n0;
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      if (n0 + j < g.N) {
        float v = acc[j];
        if (g.bias) v += __ldg(g.bias + n0 + j);
        v = e2_apply_act(v, g.act);
        if (g.round_tf32) v = e2_round_tf32(v);
        out[j] = v;
      }
    }
  }
}

int e2_launch_conv_c1_fwd(e2_handle* h, const GatherGemm& g, cudaStream_t s) {
  const int T = g.tz * g.tx * g.ty;
  int64_t M = (int64_t)g.On * g.Oz * g.Ox * g.Oy;
  constexpr int NB = 16;
  dim3 grid((unsigned)e2_grid_1d(M, 128, h->sm_count, 16), (unsigned)((g.N + NB - 1) / NB));
  size_t smem = sizeof(float) * T * NB;
  if (smem > 48 * 1024) return e2_fail(h, E2_ERR_UNSUPPORTED, "conv c_in==1: filter too large");
  k_conv_c1_fwd<NB><<<grid, 128, smem, s>>>(g);
  e2_count_launch(h);
  E2_CUDA_CHECK(h, "conv_c1_fwd");
  return E2_OK;
}

// -------------------------------------------------------------------- reduce-GEMM
// W[r][tap][s] += sum over this block's slice of positions.  64x64 tile, 16 positions / step.
__global__ void __launch_bounds__(256) k_reduce_gemm(ReduceGemm g) {
  __shared__ __align__(16) float Ps[GBK][GBM + 4];
  __shared__ __align__(16) float Qs[GBK][GBN + 4];
  const int tid = threadIdx.x;
  const int tiles_s = (g.S + GBN - 1) / GBN;
  const int r0 = (blockIdx.x / tiles_s) * GBM, s0 = (blockIdx.x % tiles_s) * GBN;
  const int tap = blockIdx.y;
  const int k3 = tap % g.ty, j3 = (tap / g.ty) % g.tx, i3 = tap / (g.ty * g.tx);
  const int64_t M = (int64_t)g.Mn * g.Mz * g.Mx * g.My;
  const int64_t per = ((M + gridDim.z - 1) / gridDim.z + GBK - 1) / GBK * GBK;
  const int64_t mb = (int64_t)blockIdx.z * per, me = (mb + per < M) ? mb + per : M;
  const int lpos = tid >> 4, lc = (tid & 15) * 4;  // loader: position lpos of the step, 4 channels
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int64_t mm = mb; mm < me; mm += GBK) {
    int64_t m = mm + lpos;
    float pv[4] = {0.f, 0.f, 0.f, 0.f}, qv[4] = {0.f, 0.f, 0.f, 0.f};
    if (m < me) {
      int my = (int)(m % g.My);
      int64_t t = m / g.My;
      int mx = (int)(t % g.Mx);
      t /= g.Mx;
      int mz = (int)(t % g.Mz);
      int mn = (int)(t / g.Mz);
      const float* prow = g.P + m * g.p_pitch;
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (r0 + lc + q < g.R) pv[q] = __ldg(prow + r0 + lc + q);
      int qz = mz * g.sz + g.oz + i3, qx = mx * g.sx + g.ox + j3, qy = my * g.sy + g.oy + k3;
      if (qz >= 0 && qz < g.Qz && qx >= 0 && qx < g.Qx && qy >= 0 && qy < g.Qy) {
        const float* qrow = g.Q + ((((int64_t)mn * g.Qz + qz) * g.Qx + qx) * g.Qy + qy) * g.q_pitch;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (s0 + lc + q < g.S) qv[q] = __ldg(qrow + s0 + lc + q);
      }
    }
    __syncthreads();
    *reinterpret_cast<float4*>(&Ps[lpos][lc]) = make_float4(pv[0], pv[1], pv[2], pv[3]);
    *reinterpret_cast<float4*>(&Qs[lpos][lc]) = make_float4(qv[0], qv[1], qv[2], qv[3]);
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GBK; ++kk) {
      float4 a = *reinterpret_cast<const float4*>(&Ps[kk][ty * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Qs[kk][tx * 4]);
      float aa[4] = {a.x, a.y, a.z, a.w}, bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
    }
  }
  const int T = g.tz * g.tx * g.ty;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int r = r0 + ty * 4 + i;
    if (r >= g.R) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int s = s0 + tx * 4 + j;
      if (s >= g.S) continue;
      int64_t ofs;
      if (g.out_mode == 0) {  // conv: r=o, s=c, reference layout dw[o][c][flip(tap)]
        int tflip = ((g.tz - 1 - i3) * g.tx + (g.tx - 1 - j3)) * g.ty + (g.ty - 1 - k3);
        ofs = ((int64_t)r * g.S + s) * T + tflip;
      } else {  // upconv: r=c, s=o, dw[o][c][tap]
        ofs = ((int64_t)s * g.R + r) * T + tap;
      }
      atomicAdd(g.W + ofs, acc[i][j]);
    }
  }
}

int e2_launch_reduce_gemm_ffma(e2_handle* h, const ReduceGemm& g, cudaStream_t s) {
  const int T = g.tz * g.tx * g.ty;
  const int tiles = ((g.R + GBM - 1) / GBM) * ((g.S + GBN - 1) / GBN);
  const int64_t M = (int64_t)g.Mn * g.Mz * g.Mx * g.My;
  int64_t splits = (4ll * h->sm_count + (int64_t)tiles * T - 1) / ((int64_t)tiles * T);
  int64_t max_splits = (M + 255) / 256;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  cudaMemsetAsync(g.W, 0, sizeof(float) * (size_t)g.R * g.S * T, s);
  dim3 grid((unsigned)tiles, (unsigned)T, (unsigned)splits);
  k_reduce_gemm<<<grid, 256, 0, s>>>(g);
  e2_count_launch(h);
  E2_CUDA_CHECK(h, "reduce_gemm_ffma");
  return E2_OK;
}

// ------------------------------------------------- first layer (c_in == 1) weight grad
// dw[o][tap] = sum_m dy[m][o] * x[pos(m)+tap].  Block = 32 channels x 8 tap groups; the
// x loads are warp-uniform (broadcast), dy is staged through shared memory.
__global__ void __launch_bounds__(256) k_conv_c1_wgrad(ReduceGemm g) {
  constexpr int PB = 128;
  __shared__ float dys[PB][33];
  __shared__ int64_t qbase[PB];
  const int T = g.tz * g.tx * g.ty;
  const int o = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int r0 = blockIdx.y * 32;
  const int64_t M = (int64_t)g.Mn * g.Mz * g.Mx * g.My;
  float acc[8];  // taps grp, grp+8, ...  (T <= 64)
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  for (int64_t mb = (int64_t)blockIdx.x * PB; mb < M; mb += (int64_t)gridDim.x * PB) {
    __syncthreads();
    for (int i = threadIdx.x; i < PB * 32; i += 256) {
      int p = i >> 5, c = i & 31;
      int64_t m = mb + p;
      dys[p][c] = (m < M && r0 + c < g.R) ? __ldg(g.P + m * g.p_pitch + r0 + c) : 0.f;
    }
    if (threadIdx.x < PB) {
      int64_t m = mb + threadIdx.x;
      int64_t b = -1;
      if (m < M) {
        int my = (int)(m % g.My);
        int64_t t = m / g.My;
        int mx = (int)(t % g.Mx);
        t /= g.Mx;
        int mz = (int)(t % g.Mz);
        int mn = (int)(t / g.Mz);
        b = (((int64_t)mn * g.Qz + mz) * g.Qx + mx) * g.Qy + my;
      }
      qbase[threadIdx.x] = b;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int tap = grp + i * 8;
      if (tap >= T) break;
      int k3 = tap % g.ty, j3 = (tap / g.ty) % g.tx, i3 = tap / (g.ty * g.tx);
      int64_t toff = ((int64_t)i3 * g.Qx + j3) * g.Qy + k3;
      float a = 0.f;
      for (int p = 0; p < PB; ++p) {
        int64_t b = qbase[p];
        if (b < 0) break;
        a = fmaf(dys[p][o], __ldg(g.Q + (b + toff) * g.q_pitch), a);
      }
      acc[i] += a;
    }
  }
  if (r0 + o < g.R) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int tap = grp + i * 8;
      if (tap >= T) break;
      int k3 = tap % g.ty, j3 = (tap / g.ty) % g.tx, i3 = tap / (g.ty * g.tx);
      int tflip = ((g.tz - 1 - i3) * g.tx + (g.tx - 1 - j3)) * g.ty + (g.ty - 1 - k3);
      atomicAdd(g.W + (int64_t)(r0 + o) * T + tflip, acc[i]);
    }
  }
}

int e2_launch_conv_c1_wgrad(e2_handle* h, const ReduceGemm& g, cudaStream_t s) {
  const int T = g.tz * g.tx * g.ty;
  if (T > 64) return e2_fail(h, E2_ERR_UNSUPPORTED, "conv c_in==1 wgrad: more than 64 taps");
  const int64_t M = (int64_t)g.Mn * g.Mz * g.Mx * g.My;
  cudaMemsetAsync(g.W, 0, sizeof(float) * (size_t)g.R * T, s);
  int gx = (int)((M + 127) / 128);
  int cap = 4 * h->sm_count;
  if (gx > cap) gx = cap;
  dim3 grid((unsigned)gx, (unsigned)((g.R + 31) / 32));
  k_conv_c1_wgrad<<<grid, 256, 0, s>>>(g);
  e2_count_launch(h);
  E2_CUDA_CHECK(h, "conv_c1_wgrad");
  return E2_OK;
}

// ----------------------------------------------------------------------- bias grad
// db[c] = sum over positions of dy[pos][c].  32 channels x 8 position lanes per block.
__global__ void __launch_bounds__(256) k_bias_grad(const float* __restrict__ dy, int64_t M, int C, int pitch,
                                                   float* __restrict__ db) {
  __shared__ float red[8][33];
  const int c = blockIdx.y * 32 + (threadIdx.x & 31);
  const int lane = threadIdx.x >> 5;
  float a = 0.f;
  if (c < C) {
    // four independent loads in flight per thread: the serial a += load chain was latency-bound
    const int64_t step = (int64_t)gridDim.x * 8;
    int64_t m = (int64_t)blockIdx.x * 8 + lane;
    float a1 = 0.f, a2 = 0.f, a3 = 0.f;
    for (; m + 3 * step < M; m += 4 * step) {
      a += __ldg(dy + m * pitch + c);
      a1 += __ldg(dy + (m + step) * pitch + c);
      a2 += __ldg(dy + (m + 2 * step) * pitch + c);
      a3 += __ldg(dy + (m + 3 * step) * pitch + c);
    }
    for (; m < M; m += step) a += __ldg(dy + m * pitch + c);
    a += a1 + a2 + a3;
  }
  red[lane][threadIdx.x & 31] = a;
  __syncthreads();
  if (lane == 0 && c < C) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[i][threadIdx.x & 31];
    atomicAdd(db + c, s);
  }
}

int e2_launch_bias_grad(e2_handle* h, const float* dy, int64_t M, int C, int pitch, float* db, cudaStream_t s) {
  cudaMemsetAsync(db, 0, sizeof(float) * C, s);
  int gx = (int)((M + 8 * 64 - 1) / (8 * 64));
  int cap = 2 * h->sm_count;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  dim3 grid((unsigned)gx, (unsigned)((C + 31) / 32));
  k_bias_grad<<<grid, 256, 0, s>>>(dy, M, C, pitch, db);
  e2_count_launch(h);
  E2_CUDA_CHECK(h, "bias_grad");
  return E2_OK;
}

// --------------------------------------------------------------- activation backward
__global__ void __launch_bounds__(256) k_act_bwd(const float* __restrict__ y, const float* __restrict__ dy,
                                                 float* __restrict__ dpre, int64_t P, int C, int pitch, int act) {
  const int64_t total = P * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t ofs = (i / C) * pitch + (i % C);
    float yv = y[ofs], g = dy[ofs], r;
    switch (act) {
      case E2_ACT_RELU: r = yv > 0.f ? g : 0.f; break;
      case E2_ACT_TANH: r = g * (1.f - yv * yv); break;
      case E2_ACT_SIGMOID: r = g * yv * (1.f - yv); break;
      // derivatives expressed through the OUTPUT y (the pre-activation is not kept):
      case E2_ACT_SOFTPLUS: r = g * (1.f - __expf(-yv)); break;                        // sigmoid(x) = 1 - exp(-softplus(x))
      case E2_ACT_ELU: r = yv > 0.f ? g : g * (yv + 1.f); break;                        // exp(x) = y + 1 for x <= 0
      case E2_ACT_SELU: r = yv > 0.f ? g * 1.0507009873554805f : g * (yv + 1.0507009873554805f * 1.6732632423543772f); break;
      default: r = g; break;
    }
    dpre[ofs] = r;
  }
}

extern "C" int e2_act_bwd(e2_handle* h, const e2_tensor* t, int32_t act, const float* y, const float* dy, float* dpre,
                          void* stream) {
  E2_REQUIRE(h, e2_tensor_ok(t) && y && dy && dpre, "act_bwd: bad arguments");
  E2_REQUIRE(h, act != E2_ACT_ABS, "act_bwd: 'abs' needs the pre-activation sign and is not supported here");
  if (act == E2_ACT_LIN && dy == dpre) return E2_OK;
  int64_t P = e2_positions(t);
  k_act_bwd<<<e2_grid_1d(P * t->c, 256, h->sm_count), 256, 0, (cudaStream_t)stream>>>(y, dy, dpre, P, t->c, t->c_pitch,
                                                                                      act);
  e2_count_launch(h);
  E2_CUDA_CHECK(h, "act_bwd");
  return E2_OK;
}

// ------------------------------------------------------------------- op front-ends
int e2_dispatch_gather_gemm(e2_handle* h, const GatherGemm& g, int compute, cudaStream_t s) {
  static const bool no_zstack = getenv("E2_DISABLE_ZSTACK") != nullptr;
  if (compute == E2_COMPUTE_TF32) {
    if (!no_zstack && e2_gather_gemm_tc_ok(h, g) && e2_conv_zstack_tc_ok(h, g)) return e2_launch_conv_zstack_tc(h, g, s);
    if (e2_gather_gemm_tc_ok(h, g)) return e2_launch_gather_gemm_tc(h, g, s);
  }
  return e2_launch_gather_gemm_ffma(h, g, s);
}

static void conv_fwd_problem(const e2_conv_desc* d, const float* x, const float* wf, const float* bias, float* y,
                             GatherGemm* g) {
  memset(g, 0, sizeof(*g));
  int T = d->kz * d->kx * d->ky, cp = round_up(d->x.c, 4);
  g->A = x, g->a_pitch = d->x.c_pitch, g->K = d->x.c;
  g->An = d->x.n, g->Az = d->x.z, g->Ax = d->x.x, g->Ay = d->x.y;
  g->B = wf, g->b_row = (int64_t)T * cp, g->b_tap = cp;
  g->C = y, g->c_pitch = d->y.c_pitch, g->N = d->y.c;
  g->On = d->y.n, g->Oz = d->y.z, g->Ox = d->y.x, g->Oy = d->y.y;
  g->tz = d->kz, g->tx = d->kx, g->ty = d->ky;
  g->sz = g->sx = g->sy = 1;
  g->bias = d->has_bias ? bias : nullptr;
  g->act = d->act;
  // tensors that feed a kind::tf32 MMA are rounded to tf32 (round-to-nearest) by their
  // producer, so the hardware's truncation of the low 13 mantissa bits is exact
  g->round_tf32 = (d->compute == E2_COMPUTE_TF32);
}

static int current_sm_count() {
  static int sm_count = 0;
  if (!sm_count) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
      sm_count = 148;
  }
  return sm_count;
}

static void conv_dgrad_problem(const e2_conv_desc* d, const float* dy, const float* wd, float* dx, const float* relu_gate,
                               GatherGemm* g) {
  memset(g, 0, sizeof(*g));
  g->gate = relu_gate;
  int T = d->kz * d->kx * d->ky, op = round_up(d->y.c, 4);
  g->A = dy, g->a_pitch = d->y.c_pitch, g->K = d->y.c;
  g->An = d->y.n, g->Az = d->y.z, g->Ax = d->y.x, g->Ay = d->y.y;
  g->B = wd, g->b_row = (int64_t)T * op, g->b_tap = op;
  g->C = dx, g->c_pitch = d->x.c_pitch, g->N = d->x.c;
  g->On = d->x.n, g->Oz = d->x.z, g->Ox = d->x.x, g->Oy = d->x.y;
  g->tz = d->kz, g->tx = d->kx, g->ty = d->ky;
  g->oz = -(d->kz - 1), g->ox = -(d->kx - 1), g->oy = -(d->ky - 1);
  g->sz = g->sx = g->sy = 1;
  g->accumulate = d->accumulate;
  g->round_tf32 = (d->compute == E2_COMPUTE_TF32);
}

static void conv_wgrad_problem(const e2_conv_desc* d, const float* x, const float* dy, float* dw, ReduceGemm* gp) {
  ReduceGemm& g = *gp;
  memset(&g, 0, sizeof(g));
  g.P = dy, g.p_pitch = d->y.c_pitch, g.R = d->y.c;
  g.Mn = d->y.n, g.Mz = d->y.z, g.Mx = d->y.x, g.My = d->y.y;
  g.Q = x, g.q_pitch = d->x.c_pitch, g.S = d->x.c;
  g.Qn = d->x.n, g.Qz = d->x.z, g.Qx = d->x.x, g.Qy = d->x.y;
  g.tz = d->kz, g.tx = d->kx, g.ty = d->ky;
  g.sz = g.sx = g.sy = 1;
  g.W = dw, g.out_mode = 0;
}

extern "C" int e2_conv3d_workspace_size(const e2_conv_desc* d, size_t* bytes) {
  if (!d || !bytes) return E2_ERR_INVALID;
  *bytes = 0;
  // only the tcgen05 wgrad uses a workspace (per-CTA partial weight tiles, summed by a second kernel)
  if (d->compute == E2_COMPUTE_TF32 && d->x.c > 1) {
    const int sm_count = current_sm_count();
    ReduceGemm g;
    alignas(16) static float dummy[4];
    e2_handle fake;
    memset(&fake, 0, sizeof(fake));
    fake.sm_count = sm_count;
    conv_wgrad_problem(d, dummy, dummy, dummy, &g);
    if (e2_wgrad_halo_tc_ok(&fake, g)) *bytes = e2_wgrad_halo_workspace_bytes(sm_count, g);
    else if (e2_reduce_gemm_tc_ok(&fake, g)) *bytes = e2_reduce_gemm_tc_workspace_bytes(sm_count, g);
    if (e2_wgrad_zs_tc_ok(&fake, g)) *bytes = std::max(*bytes, e2_wgrad_zs_workspace_bytes(sm_count, g));
    // forward / dgrad on the tap kernel: split-K partial tiles for layers with few output positions
    GatherGemm f, b;
    conv_fwd_problem(d, dummy, dummy, nullptr, dummy, &f);
    conv_dgrad_problem(d, dummy, dummy, dummy, nullptr, &b);
    for (GatherGemm* q : {&f, &b}) {
      if (!e2_gather_gemm_tc_ok(&fake, *q)) continue;
      q->ws = dummy;                      // "a workspace will be supplied": lets the planners consider K splits
      if (e2_conv_zstack_tc_ok(&fake, *q))
        *bytes = std::max(*bytes, e2_conv_zstack_workspace_bytes(sm_count, *q));
      else
        *bytes = std::max(*bytes, e2_gather_gemm_tc_workspace_bytes(sm_count, *q));
    }
  }
  return E2_OK;
}

extern "C" int e2_conv3d_fwd(e2_handle* h, const e2_conv_desc* d, const float* x, const float* wf, const float* bias,
                             float* y, void* ws, size_t ws_bytes, void* stream) {
  int rc = check_conv(h, d);
  if (rc) return rc;
  E2_REQUIRE(h, x && wf && y && (!d->has_bias || bias), "conv3d_fwd: null pointer");
  GatherGemm g;
  conv_fwd_problem(d, x, wf, bias, y, &g);
  g.ws = ws, g.ws_bytes = ws_bytes;
  cudaStream_t s = (cudaStream_t)stream;
  if (e2_conv_pw_fwd_ok(g)) return e2_launch_conv_pw_fwd(h, g, s);
  if (d->x.c == 1) {
    // a single input channel is HBM-bound either way; with 16-byte rows (pitch % 4 == 0) the halo-plane
    // tcgen05 kernel takes it (one K = 8 MMA per tap and plane, TMA zero-fills the 7 missing channels) and
    // its TMA-store epilogue writes the output at full speed
    // (measured: TMA walks every 16-byte chunk of a box, valid or zero-filled, so the 1-of-32-channel boxes
    // load at 1/8 speed and the path is slower than the CUDA-core kernel -- opt-in via E2_C1_TC=1)
    static const bool tc1 = getenv("E2_C1_TC") != nullptr;
    if (tc1 && d->compute == E2_COMPUTE_TF32 && d->x.c_pitch % 4 == 0 && d->y.c >= 8 &&
        !(reinterpret_cast<uintptr_t>(x) & 15) && !(reinterpret_cast<uintptr_t>(wf) & 15) && e2_conv_zstack_tc_ok(h, g))
      return e2_launch_conv_zstack_tc(h, g, s);
    // TF32 mode: the threads build the im2col tile, the tensor core does the 27 MACs per output (e2_conv_c1_tc.cu)
    if (d->compute == E2_COMPUTE_TF32 && e2_conv_c1_fwd_ws_ok(g)) return e2_launch_conv_c1_fwd_ws(h, g, s);
    if (d->compute == E2_COMPUTE_TF32 && e2_conv_c1_fwd_tc_ok(g)) return e2_launch_conv_c1_fwd_tc(h, g, s);
    if (e2_conv_c1_fwd_reg_ok(g)) return e2_launch_conv_c1_fwd_reg(h, g, s);
    return e2_conv_c1_fwd_line_ok(g) ? e2_launch_conv_c1_fwd_line(h, g, s) : e2_launch_conv_c1_fwd(h, g, s);
  }
  return e2_dispatch_gather_gemm(h, g, d->compute, s);
}

// ---- conv with the max-pool in its epilogue
static int conv_pool_problem(e2_handle* h, const e2_conv_desc* d, const e2_pool_desc* p, const float* x, const float* wf,
                             const float* bias, const float* pool_bias, float* y, const e2_window* y_keep, float* yp,
                             int32_t* argmax, GatherGemm* g) {
  int rc = check_conv(h, d);
  if (rc) return rc;
  E2_REQUIRE(h, p && e2_tensor_ok(&p->x) && e2_tensor_ok(&p->y), "conv3d_fwd_pool: bad pool descriptor");
  E2_REQUIRE(h, p->x.n == d->y.n && p->x.z == d->y.z && p->x.x == d->y.x && p->x.y == d->y.y && p->x.c == d->y.c,
             "conv3d_fwd_pool: the pool's input is not the conv's output");
  E2_REQUIRE(h, p->pz >= 1 && p->px >= 1 && p->py >= 1 && p->x.z % p->pz == 0 && p->x.x % p->px == 0 && p->x.y % p->py == 0 &&
                    p->y.z == p->x.z / p->pz && p->y.x == p->x.x / p->px && p->y.y == p->x.y / p->py && p->y.n == p->x.n &&
                    p->y.c == p->x.c,
             "conv3d_fwd_pool: pooled extents do not match input/pool");
  conv_fwd_problem(d, x, wf, bias, y, g);
  g->fuse_pool = 1, g->qz = p->pz, g->qx = p->px, g->qy = p->py;
  g->Cp = yp, g->Ci = argmax, g->cp_pitch = p->y.c_pitch;
  g->pbias = p->has_bias ? pool_bias : nullptr, g->pact = p->act, g->pround = p->round_tf32;
  if (y_keep) {
    E2_REQUIRE(h, y_keep->z0 <= y_keep->z1 && y_keep->x0 <= y_keep->x1 && y_keep->y0 <= y_keep->y1,
               "conv3d_fwd_pool: y_keep is not a half-open window");
    g->keep[0] = y_keep->z0, g->keep[1] = y_keep->z1, g->keep[2] = y_keep->x0, g->keep[3] = y_keep->x1;
    g->keep[4] = y_keep->y0, g->keep[5] = y_keep->y1;
  } else {
    g->keep[0] = g->keep[2] = g->keep[4] = 0;
    g->keep[1] = d->y.z, g->keep[3] = d->y.x, g->keep[5] = d->y.y;
  }
  return E2_OK;
}

static bool conv_pool_fusable(const e2_handle* h, const e2_conv_desc* d, const e2_pool_desc* p, const GatherGemm& g) {
  static const bool off = getenv("E2_DISABLE_ZSTACK") != nullptr || getenv("E2_NO_POOL_FUSION") != nullptr;
  if (off || d->compute != E2_COMPUTE_TF32 || d->x.c == 1 || p->mode != E2_POOL_MAX) return false;
  if (p->x.c_pitch != d->y.c_pitch || e2_conv_pw_fwd_ok(g)) return false;
  return e2_gather_gemm_tc_ok(h, g) && e2_conv_zstack_pool_ok(h, g);
}

extern "C" int e2_conv3d_fwd_pool_supported(e2_handle* h, const e2_conv_desc* d, const e2_pool_desc* p) {
  if (!h) return E2_ERR_INVALID;
  alignas(16) static float dummy[4];
  GatherGemm g;
  int rc = conv_pool_problem(h, d, p, dummy, dummy, dummy, dummy, dummy, nullptr, dummy, reinterpret_cast<int32_t*>(dummy), &g);
  if (rc) return rc;
  return conv_pool_fusable(h, d, p, g) ? 1 : 0;
}

extern "C" int e2_conv3d_fwd_pool(e2_handle* h, const e2_conv_desc* d, const e2_pool_desc* p, const float* x,
                                  const float* wf, const float* bias, const float* pool_bias, float* y,
                                  const e2_window* y_keep, float* yp, int32_t* argmax, void* ws, size_t ws_bytes,
                                  void* stream) {
  GatherGemm g;
  int rc = conv_pool_problem(h, d, p, x, wf, bias, pool_bias, y, y_keep, yp, argmax, &g);
  if (rc) return rc;
  E2_REQUIRE(h, x && wf && yp && (!d->has_bias || bias) && (!p->has_bias || pool_bias), "conv3d_fwd_pool: null pointer");
  if (conv_pool_fusable(h, d, p, g)) return e2_launch_conv_zstack_tc(h, g, (cudaStream_t)stream);
  // not fusable: the two calls this entry point stands for (they need the unpooled tensor)
  E2_REQUIRE(h, y, "conv3d_fwd_pool: this problem is not fused (e2_conv3d_fwd_pool_supported), so y is required");
  rc = e2_conv3d_fwd(h, d, x, wf, bias, y, ws, ws_bytes, stream);
  if (rc) return rc;
  return e2_maxpool3d_fwd(h, p, y, pool_bias, yp, argmax, stream);
}

extern "C" int e2_conv3d_dgrad(e2_handle* h, const e2_conv_desc* d, const float* dy, const float* wd, float* dx,
                               const float* relu_gate, void* ws, size_t ws_bytes, void* stream) {
  int rc = check_conv(h, d);
  if (rc) return rc;
  E2_REQUIRE(h, dy && wd && dx, "conv3d_dgrad: null pointer");
  GatherGemm g;
  conv_dgrad_problem(d, dy, wd, dx, relu_gate, &g);
  g.ws = ws, g.ws_bytes = ws_bytes;
  cudaStream_t s = (cudaStream_t)stream;
  if (e2_conv_pw_dgrad_ok(g)) return e2_launch_conv_pw_dgrad(h, g, s);
  return e2_dispatch_gather_gemm(h, g, d->compute, s);
}

extern "C" int e2_conv3d_wgrad(e2_handle* h, const e2_conv_desc* d, const float* x, const float* dy, float* dw,
                               float* db, void* ws, size_t ws_bytes, void* stream) {
  int rc = check_conv(h, d);
  if (rc) return rc;
  E2_REQUIRE(h, x && dy && dw, "conv3d_wgrad: null pointer");
  ReduceGemm g;
  conv_wgrad_problem(d, x, dy, dw, &g);
  bool db_done = false;
  cudaStream_t s = (cudaStream_t)stream;
  if (e2_conv_pw_wgrad_ok(g))
    rc = e2_launch_conv_pw_wgrad(h, g, db, s), db_done = (db != nullptr);
  else if (d->x.c == 1 && d->compute == E2_COMPUTE_TF32 && e2_wgrad_c1_tc_ok(g))
    rc = e2_launch_wgrad_c1_tc(h, g, db, s), db_done = (db != nullptr);
  else if (d->x.c == 1 && e2_conv_c1_wgrad_reg_ok(g))
    rc = e2_launch_conv_c1_wgrad_reg(h, g, db, s), db_done = (db != nullptr);
  else if (d->x.c == 1 && e2_conv_c1_wgrad_line_ok(g))
    rc = e2_launch_conv_c1_wgrad_line(h, g, s);
  else if (d->x.c == 1 && d->kz * d->kx * d->ky <= 64)
    rc = e2_launch_conv_c1_wgrad(h, g, s);
  else if (d->compute == E2_COMPUTE_TF32 && ws && !(reinterpret_cast<uintptr_t>(ws) & 15) && e2_wgrad_zs_tc_ok(h, g) &&
           ws_bytes >= e2_wgrad_zs_workspace_bytes(h->sm_count, g))
    rc = e2_launch_wgrad_zs_tc(h, g, ws, ws_bytes, db, &db_done, s);
  else if (d->compute == E2_COMPUTE_TF32 && e2_wgrad_halo_tc_ok(h, g))
    rc = e2_launch_wgrad_halo_tc(h, g, ws, ws_bytes, db, &db_done, s);
  else if (d->compute == E2_COMPUTE_TF32 && e2_reduce_gemm_tc_ok(h, g))
    rc = e2_launch_reduce_gemm_tc(h, g, ws, ws_bytes, s);
  else
    rc = e2_launch_reduce_gemm_ffma(h, g, s);
  if (rc) return rc;
  if (db && !db_done) return e2_launch_bias_grad(h, dy, e2_positions(&d->y), d->y.c, d->y.c_pitch, db, s);
  return E2_OK;
}

static void upconv_fwd_problem(const e2_upconv_desc* d, const float* x, const float* wf, const float* bias, float* y,
                               GatherGemm* g) {
  memset(g, 0, sizeof(*g));
  int T = d->pz * d->px * d->py, cp = round_up(d->x.c, 4);
  g->A = x, g->a_pitch = d->x.c_pitch, g->K = d->x.c;
  g->An = d->x.n, g->Az = d->x.z, g->Ax = d->x.x, g->Ay = d->x.y;
  g->B = wf, g->b_row = cp, g->b_tap = 0;
  g->C = y, g->c_pitch = d->y.c_pitch, g->N = T * d->y.c;
  g->On = d->x.n, g->Oz = d->x.z, g->Ox = d->x.x, g->Oy = d->x.y;
  g->tz = g->tx = g->ty = 1;
  g->sz = g->sx = g->sy = 1;
  g->bias = d->has_bias ? bias : nullptr;
  g->act = d->act;
  g->shuffle = 1, g->pz = d->pz, g->px = d->px, g->py = d->py, g->Fo = d->y.c;
  g->round_tf32 = (d->compute == E2_COMPUTE_TF32);
}

static void upconv_dgrad_problem(const e2_upconv_desc* d, const float* dy, const float* wd, float* dx,
                                 const float* relu_gate, GatherGemm* g) {
  memset(g, 0, sizeof(*g));
  g->gate = relu_gate;
  int T = d->pz * d->px * d->py, np = round_up(T * d->y.c, 4);
  g->A = dy, g->a_pitch = d->y.c_pitch, g->K = d->y.c;
  g->An = d->y.n, g->Az = d->y.z, g->Ax = d->y.x, g->Ay = d->y.y;
  g->B = wd, g->b_row = np, g->b_tap = d->y.c;
  g->C = dx, g->c_pitch = d->x.c_pitch, g->N = d->x.c;
  g->On = d->x.n, g->Oz = d->x.z, g->Ox = d->x.x, g->Oy = d->x.y;
  g->tz = d->pz, g->tx = d->px, g->ty = d->py;
  g->sz = d->pz, g->sx = d->px, g->sy = d->py;
  g->accumulate = d->accumulate;
  g->round_tf32 = (d->compute == E2_COMPUTE_TF32);
}

// dw[o][c][tap] = sum_m x[m][c] * dy[m*p + tap][o]
static void upconv_wgrad_problem(const e2_upconv_desc* d, const float* x, const float* dy, float* dw, ReduceGemm* gp) {
  ReduceGemm& g = *gp;
  memset(&g, 0, sizeof(g));
  g.P = x, g.p_pitch = d->x.c_pitch, g.R = d->x.c;
  g.Mn = d->x.n, g.Mz = d->x.z, g.Mx = d->x.x, g.My = d->x.y;
  g.Q = dy, g.q_pitch = d->y.c_pitch, g.S = d->y.c;
  g.Qn = d->y.n, g.Qz = d->y.z, g.Qx = d->y.x, g.Qy = d->y.y;
  g.tz = d->pz, g.tx = d->px, g.ty = d->py;
  g.sz = d->pz, g.sx = d->px, g.sy = d->py;
  g.W = dw, g.out_mode = 1;
}

extern "C" int e2_upconv3d_workspace_size(const e2_upconv_desc* d, size_t* bytes) {
  if (!d || !bytes) return E2_ERR_INVALID;
  *bytes = 0;
  if (d->compute != E2_COMPUTE_TF32) return E2_OK;
  const int sm_count = current_sm_count();
  alignas(16) static float dummy[4];
  e2_handle fake;
  memset(&fake, 0, sizeof(fake));
  fake.sm_count = sm_count;
  GatherGemm f, b;
  upconv_fwd_problem(d, dummy, dummy, nullptr, dummy, &f);
  upconv_dgrad_problem(d, dummy, dummy, dummy, nullptr, &b);
  for (const GatherGemm* q : {&f, &b})
    if (e2_gather_gemm_tc_ok(&fake, *q)) *bytes = std::max(*bytes, e2_gather_gemm_tc_workspace_bytes(sm_count, *q));
  ReduceGemm w;
  upconv_wgrad_problem(d, dummy, dummy, dummy, &w);
  if (e2_reduce_gemm_tc_ok(&fake, w)) *bytes = std::max(*bytes, e2_reduce_gemm_tc_workspace_bytes(sm_count, w));
  return E2_OK;
}

extern "C" int e2_upconv3d_fwd(e2_handle* h, const e2_upconv_desc* d, const float* x, const float* wf,
                               const float* bias, float* y, void* ws, size_t ws_bytes, void* stream) {
  int rc = check_upconv(h, d);
  if (rc) return rc;
  E2_REQUIRE(h, x && wf && y && (!d->has_bias || bias), "upconv3d_fwd: null pointer");
  GatherGemm g;
  upconv_fwd_problem(d, x, wf, bias, y, &g);
  g.ws = ws, g.ws_bytes = ws_bytes;
  cudaStream_t s = (cudaStream_t)stream;
  return e2_dispatch_gather_gemm(h, g, d->compute, s);
}

extern "C" int e2_upconv3d_dgrad(e2_handle* h, const e2_upconv_desc* d, const float* dy, const float* wd, float* dx,
                                 const float* relu_gate, void* ws, size_t ws_bytes, void* stream) {
  int rc = check_upconv(h, d);
  if (rc) return rc;
  E2_REQUIRE(h, dy && wd && dx, "upconv3d_dgrad: null pointer");
  // dx[m][c] = sum_{tap,o} dy[m*p + tap][o] * w[o][c][tap]  -- a strided gather-GEMM
  GatherGemm g;
  upconv_dgrad_problem(d, dy, wd, dx, relu_gate, &g);
  g.ws = ws, g.ws_bytes = ws_bytes;
  cudaStream_t s = (cudaStream_t)stream;
  return e2_dispatch_gather_gemm(h, g, d->compute, s);
}

extern "C" int e2_upconv3d_wgrad(e2_handle* h, const e2_upconv_desc* d, const float* x, const float* dy, float* dw,
                                 float* db, void* ws, size_t ws_bytes, void* stream) {
  int rc = check_upconv(h, d);
  if (rc) return rc;
  E2_REQUIRE(h, x && dy && dw, "upconv3d_wgrad: null pointer");
  ReduceGemm g;
  upconv_wgrad_problem(d, x, dy, dw, &g);
  cudaStream_t s = (cudaStream_t)stream;
  if (d->compute == E2_COMPUTE_TF32 && e2_reduce_gemm_tc_ok(h, g))
    rc = e2_launch_reduce_gemm_tc(h, g, ws, ws_bytes, s);
  else
    rc = e2_launch_reduce_gemm_ffma(h, g, s);
  if (rc) return rc;
  if (db) return e2_launch_bias_grad(h, dy, e2_positions(&d->y), d->y.c, d->y.c_pitch, db, s);
  return E2_OK;
}
