// 1x1x1 convolutions with at most 4 output channels: the last layer of every BASELINE net (64 -> 2 "barr",
// 200 -> 2).  Reference: the tensordot shortcut computations.py:330-335, 377-384.  Arithmetic intensity is
// < 2 FLOP/B, so these are streaming kernels in exact fp32 (both compute modes): the generic 64x64 GEMM tiles they
// replace spent 25-50 us per pass on a 14 MB tensor.
//   fwd    y[m][o]  = act(sum_c x[m][c] * w[o][c] + b[o])        one position per lane group, shuffle reduce
//   dgrad  dx[m][c] = sum_o dy[m][o] * w[o][c]  (gate / accumulate)   one thread per position x 4 channels
//   wgrad  dw[o][c] = sum_m dy[m][o] * x[m][c],  db[o] = sum_m dy[m][o]    per-thread register sums over a
//          position stripe, block reduction in shared memory, one fp32 atomic per block and output
#include <algorithm>
#include "e2_common.cuh"
#include "e2_conv_internal.cuh"

namespace {

constexpr int PW_MAXN = 4;

__global__ void __launch_bounds__(256) k_pw_fwd(GatherGemm g, int64_t positions, int lpp) {
  extern __shared__ float w_s[];   // [N][K]
  for (int i = threadIdx.x; i < g.N * g.K; i += blockDim.x) w_s[i] = __ldg(g.B + (int64_t)(i / g.K) * g.b_row + (i % g.K));
  __syncthreads();
  const int sub = threadIdx.x % lpp;
  const int k4n = g.K / 4;
  const int64_t gstride = (int64_t)gridDim.x * (blockDim.x / lpp);
  const int64_t iters = (positions + gstride - 1) / gstride;     // uniform trip count: every lane joins the shuffles
  for (int64_t it = 0; it < iters; ++it) {
    const int64_t m = it * gstride + (int64_t)blockIdx.x * (blockDim.x / lpp) + threadIdx.x / lpp;
    const bool valid = m < positions;
    float acc[PW_MAXN] = {0.f, 0.f, 0.f, 0.f};
    const float4* row = reinterpret_cast<const float4*>(g.A + (valid ? m : 0) * g.a_pitch);
    for (int c4 = sub; c4 < k4n; c4 += lpp) {
      const float4 x4 = __ldg(row + c4);
#pragma unroll
      for (int n = 0; n < PW_MAXN; ++n) {
        if (n >= g.N) break;
        const float4 w4 = *reinterpret_cast<const float4*>(w_s + n * g.K + c4 * 4);
        acc[n] += x4.x * w4.x + x4.y * w4.y + x4.z * w4.z + x4.w * w4.w;
      }
    }
    for (int o = lpp >> 1; o > 0; o >>= 1) {
#pragma unroll
      for (int n = 0; n < PW_MAXN; ++n) acc[n] += __shfl_xor_sync(0xffffffffu, acc[n], o);
    }
    if (sub == 0 && valid) {
#pragma unroll
      for (int n = 0; n < PW_MAXN; ++n) {
        if (n >= g.N) break;
        float v = acc[n] + (g.bias ? __ldg(g.bias + n) : 0.f);
        v = e2_apply_act(v, g.act);
        const int64_t ofs = m * g.c_pitch + n;
        if (g.gate && !(__ldg(g.gate + ofs) > 0.f)) v = 0.f;
        if (g.accumulate) v += g.C[ofs];
        g.C[ofs] = g.round_tf32 ? e2_round_tf32(v) : v;
      }
    }
  }
}

// dx[m][c..c+3] = sum_k dy[m][k] * B[c*b_row + k]   (K = the layer's output channels, <= 4)
__global__ void __launch_bounds__(256) k_pw_dgrad(GatherGemm g, int64_t positions) {
  extern __shared__ float w_s[];   // [K][N4*4]  transposed: w_s[k*Np + c]
  const int Np = (g.N + 3) / 4 * 4;
  for (int i = threadIdx.x; i < g.K * Np; i += blockDim.x) {
    const int k = i / Np, c = i % Np;
    w_s[i] = c < g.N ? __ldg(g.B + (int64_t)c * g.b_row + k) : 0.f;
  }
  __syncthreads();
  const int n4 = Np / 4;
  const int64_t total = positions * n4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = i / n4;
    const int c = (int)(i - m * n4) * 4;
    float a[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < g.K; ++k) {
      const float d = __ldg(g.A + m * g.a_pitch + k);
      const float4 w4 = *reinterpret_cast<const float4*>(w_s + k * Np + c);
      a[0] += d * w4.x, a[1] += d * w4.y, a[2] += d * w4.z, a[3] += d * w4.w;
    }
    const int64_t ofs = m * g.c_pitch + c;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (c + e >= g.N) break;
      float v = e2_apply_act(a[e] + (g.bias ? __ldg(g.bias + c + e) : 0.f), g.act);
      if (g.gate && !(__ldg(g.gate + ofs + e) > 0.f)) v = 0.f;
      if (g.accumulate) v += g.C[ofs + e];
      g.C[ofs + e] = g.round_tf32 ? e2_round_tf32(v) : v;
    }
  }
}

// W[r][s] = sum_m P[m][r] * Q[m][s], db[r] = sum_m P[m][r]    (R <= 4; W and db zeroed by the launcher)
__global__ void __launch_bounds__(256) k_pw_wgrad(ReduceGemm g, int64_t positions, float* __restrict__ db) {
  __shared__ float red[256][PW_MAXN * 4 + PW_MAXN];
  const int s4n = (g.S + 3) / 4;
  const int plan = 256 / s4n;                 // position lanes per block
  const int sl = threadIdx.x % s4n, pl = threadIdx.x / s4n;
  const bool active = pl < plan;
  float acc[PW_MAXN][4];
  float bsum[PW_MAXN];
#pragma unroll
  for (int r = 0; r < PW_MAXN; ++r) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = bsum[r] = 0.f;
  if (active) {
    const int s = sl * 4;
    for (int64_t m = (int64_t)blockIdx.x * plan + pl; m < positions; m += (int64_t)gridDim.x * plan) {
      float q[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) q[e] = s + e < g.S ? __ldg(g.Q + m * g.q_pitch + s + e) : 0.f;
#pragma unroll
      for (int r = 0; r < PW_MAXN; ++r) {
        if (r >= g.R) break;
        const float d = __ldg(g.P + m * g.p_pitch + r);
        acc[r][0] += d * q[0], acc[r][1] += d * q[1], acc[r][2] += d * q[2], acc[r][3] += d * q[3];
        bsum[r] += d;
      }
    }
  }
#pragma unroll
  for (int r = 0; r < PW_MAXN; ++r) {
#pragma unroll
    for (int e = 0; e < 4; ++e) red[threadIdx.x][r * 4 + e] = acc[r][e];
    red[threadIdx.x][PW_MAXN * 4 + r] = bsum[r];
  }
  __syncthreads();
  // one thread per (s lane, value): sum over the position lanes
  for (int i = threadIdx.x; i < s4n * (PW_MAXN * 4 + PW_MAXN); i += 256) {
    const int l = i % s4n, v = i / s4n;
    float t = 0.f;
    for (int p2 = 0; p2 < plan; ++p2) t += red[p2 * s4n + l][v];
    if (v < PW_MAXN * 4) {
      const int r = v / 4, s = l * 4 + (v & 3);
      if (r < g.R && s < g.S) atomicAdd(g.W + (int64_t)r * g.S + s, t);
    } else if (l == 0 && db) {
      const int r = v - PW_MAXN * 4;
      if (r < g.R) atomicAdd(db + r, t);
    }
  }
}

}  // namespace

bool e2_conv_pw_fwd_ok(const GatherGemm& g) {
  return g.tz * g.tx * g.ty == 1 && g.sz == 1 && g.sx == 1 && g.sy == 1 && !g.shuffle && g.N >= 1 && g.N <= PW_MAXN &&
         g.K >= 4 && (g.K % 4) == 0 && g.K <= 2048 && (g.a_pitch % 4) == 0 && !(reinterpret_cast<uintptr_t>(g.A) & 15) &&
         g.oz == 0 && g.ox == 0 && g.oy == 0 && !getenv("E2_NO_PW");
}

int e2_launch_conv_pw_fwd(e2_handle* h, const GatherGemm& g, cudaStream_t s) {
  const int64_t positions = (int64_t)g.On * g.Oz * g.Ox * g.Oy;
  int lpp = 1;
  while (lpp < 32 && lpp < g.K / 4) lpp *= 2;
  const int groups = 256 / lpp;
  k_pw_fwd<<<e2_grid_1d((positions + groups - 1) / groups * 256, 256, h->sm_count, 16), 256, sizeof(float) * g.N * g.K, s>>>(
      g, positions, lpp);
  e2_count_launch(h);
  E2_CUDA_CHECK(h, "conv_pw_fwd");
  return E2_OK;
}

bool e2_conv_pw_dgrad_ok(const GatherGemm& g) {
  return g.tz * g.tx * g.ty == 1 && g.sz == 1 && g.sx == 1 && g.sy == 1 && !g.shuffle && g.K >= 1 && g.K <= PW_MAXN &&
         g.N >= 4 && g.N <= 2048 && g.oz == 0 && g.ox == 0 && g.oy == 0 && !getenv("E2_NO_PW");
}

int e2_launch_conv_pw_dgrad(e2_handle* h, const GatherGemm& g, cudaStream_t s) {
  const int64_t positions = (int64_t)g.On * g.Oz * g.Ox * g.Oy;
  const int Np = (g.N + 3) / 4 * 4;
  k_pw_dgrad<<<e2_grid_1d(positions * (Np / 4), 256, h->sm_count, 16), 256, sizeof(float) * g.K * Np, s>>>(g, positions);
  e2_count_launch(h);
  E2_CUDA_CHECK(h, "conv_pw_dgrad");
  return E2_OK;
}

bool e2_conv_pw_wgrad_ok(const ReduceGemm& g) {
  return g.tz * g.tx * g.ty == 1 && g.sz == 1 && g.sx == 1 && g.sy == 1 && g.R >= 1 && g.R <= PW_MAXN && g.S >= 4 &&
         g.S <= 1024 && g.out_mode == 0 && g.oz == 0 && g.ox == 0 && g.oy == 0 && !getenv("E2_NO_PW");
}

int e2_launch_conv_pw_wgrad(e2_handle* h, const ReduceGemm& g, float* db, cudaStream_t s) {
  const int64_t positions = (int64_t)g.Mn * g.Mz * g.Mx * g.My;
  cudaMemsetAsync(g.W, 0, sizeof(float) * (size_t)g.R * g.S, s);
  if (db) cudaMemsetAsync(db, 0, sizeof(float) * (size_t)g.R, s);
  const int plan = 256 / ((g.S + 3) / 4);
  int grid = (int)std::min<int64_t>((positions + plan - 1) / plan, (int64_t)h->sm_count * 4);
  if (grid < 1) grid = 1;
  k_pw_wgrad<<<grid, 256, 0, s>>>(g, positions, db);
  e2_count_launch(h);
  E2_CUDA_CHECK(h, "conv_pw_wgrad");
  return E2_OK;
}
