// First-layer convolutions (one input channel): K per tap is 1, so these layers are HBM-bound and run
// on CUDA cores.  Both kernels work on whole output lines (fixed n, z, x; all y) staged through shared
// memory so that every global access is a coalesced 128-byte run:
//   forward  y[pos][o] = act(sum_tap x[pos+tap] * w[o][tap] + b[o])    8 lanes x float4 = 32 channels / position
//   wgrad    dw[o][tap] = sum_pos dy[pos][o] * x[pos+tap]              warp = one (z,x) tap row, registers hold
//                                                                       [4 channels x ky taps], persistent blocks
// Algorithmic traffic: forward 4*(N_in + N_out) bytes, wgrad 4*(N_dy + N_x) bytes.
#include <algorithm>
#include "e2_common.cuh"
#include "e2_conv_internal.cuh"

namespace {

constexpr int C1_MAX_ROW = 640;   // max (Oy + ky - 1) handled by the line kernels
constexpr int C1_MAX_ROWS = 16;   // max kz*kx

// ------------------------------------------------------------------------------------ forward
__global__ void __launch_bounds__(256) k_c1_fwd_line(GatherGemm g, int lines) {
  extern __shared__ float sm[];
  const int T = g.tz * g.tx * g.ty;
  const int rows = g.tz * g.tx, roww = g.Oy + g.ty - 1;
  float* ws = sm;                       // [T][32]
  float* bs = ws + T * 32;              // [32]
  float* xs = bs + 32;                  // [rows][roww]
  const int n0 = blockIdx.y * 32;
  for (int i = threadIdx.x; i < T * 32; i += blockDim.x) {
    const int t = i >> 5, c = i & 31;
    ws[i] = (n0 + c < g.N) ? __ldg(g.B + (int64_t)(n0 + c) * g.b_row + (int64_t)t * g.b_tap) : 0.f;
  }
  if (threadIdx.x < 32) bs[threadIdx.x] = (g.bias && n0 + threadIdx.x < g.N) ? __ldg(g.bias + n0 + threadIdx.x) : 0.f;
  const int quad = threadIdx.x & 7, p0 = threadIdx.x >> 3;
  const bool vec_ok = (g.c_pitch % 4 == 0) && ((reinterpret_cast<uintptr_t>(g.C) & 15) == 0);
  for (int line = blockIdx.x; line < lines; line += gridDim.x) {
    const int ox = line % g.Ox;
    const int t2 = line / g.Ox;
    const int oz = t2 % g.Oz, on = t2 / g.Oz;
    __syncthreads();
    for (int i = threadIdx.x; i < rows * roww; i += blockDim.x) {
      const int r = i / roww, y = i - r * roww;
      const int az = oz + r / g.tx, ax = ox + r % g.tx;
      xs[i] = __ldg(g.A + ((((int64_t)on * g.Az + az) * g.Ax + ax) * g.Ay + y) * g.a_pitch);
    }
    __syncthreads();
    for (int y = p0; y < g.Oy; y += 32) {
      float4 acc = *reinterpret_cast<const float4*>(bs + quad * 4);
      for (int r = 0; r < rows; ++r) {
        const float* xr = xs + r * roww + y;
        const float* wr = ws + (r * g.ty) * 32 + quad * 4;
        for (int k = 0; k < g.ty; ++k) {
          const float xv = xr[k];
          const float4 w4 = *reinterpret_cast<const float4*>(wr + k * 32);
          acc.x = fmaf(xv, w4.x, acc.x), acc.y = fmaf(xv, w4.y, acc.y);
          acc.z = fmaf(xv, w4.z, acc.z), acc.w = fmaf(xv, w4.w, acc.w);
        }
      }
      if (g.act == E2_ACT_RELU) {
        acc.x = fmaxf(acc.x, 0.f), acc.y = fmaxf(acc.y, 0.f), acc.z = fmaxf(acc.z, 0.f), acc.w = fmaxf(acc.w, 0.f);
      } else if (g.act != E2_ACT_LIN) {
        acc.x = e2_apply_act(acc.x, g.act), acc.y = e2_apply_act(acc.y, g.act);
        acc.z = e2_apply_act(acc.z, g.act), acc.w = e2_apply_act(acc.w, g.act);
      }
      if (g.round_tf32) {
        acc.x = e2_round_tf32(acc.x), acc.y = e2_round_tf32(acc.y), acc.z = e2_round_tf32(acc.z), acc.w = e2_round_tf32(acc.w);
      }
      const int64_t pos = (((int64_t)on * g.Oz + oz) * g.Ox + ox) * g.Oy + y;
      float* out = g.C + pos * g.c_pitch + n0 + quad * 4;
      if (vec_ok && n0 + quad * 4 + 3 < g.N) {
        *reinterpret_cast<float4*>(out) = acc;
      } else {
        const float v[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (n0 + quad * 4 + j < g.N) out[j] = v[j];
      }
    }
  }
}

// -------------------------------------------------------------------------------------- wgrad
// blockDim = 32 * rows (one warp per (i,j) tap row); lane = yph*8 + og: 4 y phases x 8 channel quads.
template <int KY>
__global__ void __launch_bounds__(512) k_c1_wgrad_line(ReduceGemm g, int lines) {
  extern __shared__ float sm[];
  const int rows = g.tz * g.tx, roww = g.My + KY - 1;
  float* dys = sm;                         // [My][32]
  float* xs = dys + g.My * 32;             // [rows][roww]
  const int r0 = blockIdx.y * 32;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int og = lane & 7, yph = lane >> 3;
  float acc[4][KY];
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int k = 0; k < KY; ++k) acc[c][k] = 0.f;
  const bool vec_ok = (g.p_pitch % 4 == 0) && ((reinterpret_cast<uintptr_t>(g.P) & 15) == 0) && (r0 + 32 <= g.R);
  for (int line = blockIdx.x; line < lines; line += gridDim.x) {
    const int mx = line % g.Mx;
    const int t2 = line / g.Mx;
    const int mz = t2 % g.Mz, mn = t2 / g.Mz;
    const int64_t pos0 = (((int64_t)mn * g.Mz + mz) * g.Mx + mx) * g.My;
    __syncthreads();
    if (vec_ok) {
      for (int i = threadIdx.x; i < g.My * 8; i += blockDim.x) {
        const int y = i >> 3, q = i & 7;
        *reinterpret_cast<float4*>(dys + y * 32 + q * 4) =
            __ldg(reinterpret_cast<const float4*>(g.P + (pos0 + y) * g.p_pitch + r0 + q * 4));
      }
    } else {
      for (int i = threadIdx.x; i < g.My * 32; i += blockDim.x) {
        const int y = i >> 5, c = i & 31;
        dys[i] = (r0 + c < g.R) ? __ldg(g.P + (pos0 + y) * g.p_pitch + r0 + c) : 0.f;
      }
    }
    for (int i = threadIdx.x; i < rows * roww; i += blockDim.x) {
      const int r = i / roww, y = i - r * roww;
      const int qz = mz + g.oz + r / g.tx, qx = mx + g.ox + r % g.tx, qy = y + g.oy;
      float v = 0.f;
      if (qz >= 0 && qz < g.Qz && qx >= 0 && qx < g.Qx && qy >= 0 && qy < g.Qy)
        v = __ldg(g.Q + ((((int64_t)mn * g.Qz + qz) * g.Qx + qx) * g.Qy + qy) * g.q_pitch);
      xs[i] = v;
    }
    __syncthreads();
    const float* xr = xs + w * roww;
    for (int y = yph; y < g.My; y += 4) {
      const float4 d4 = *reinterpret_cast<const float4*>(dys + y * 32 + og * 4);
#pragma unroll
      for (int k = 0; k < KY; ++k) {
        const float xv = xr[y + k];
        acc[0][k] = fmaf(d4.x, xv, acc[0][k]);
        acc[1][k] = fmaf(d4.y, xv, acc[1][k]);
        acc[2][k] = fmaf(d4.z, xv, acc[2][k]);
        acc[3][k] = fmaf(d4.w, xv, acc[3][k]);
      }
    }
  }
  // reduce the 4 y phases, then one atomic per (channel, tap) and block
  const int T = g.tz * g.tx * KY;
  const int i3 = w / g.tx, j3 = w % g.tx;
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int k = 0; k < KY; ++k) {
      float v = acc[c][k];
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      const int r = r0 + og * 4 + c;
      if (yph == 0 && r < g.R) {
        const int tflip = ((g.tz - 1 - i3) * g.tx + (g.tx - 1 - j3)) * KY + (KY - 1 - k);
        atomicAdd(g.W + (int64_t)r * T + tflip, v);
      }
    }
}

template <int KY>
int launch_wgrad(e2_handle* h, const ReduceGemm& g, cudaStream_t s) {
  const int rows = g.tz * g.tx;
  const int lines = g.Mn * g.Mz * g.Mx;
  const size_t smem = sizeof(float) * ((size_t)g.My * 32 + (size_t)rows * (g.My + KY - 1));
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(k_c1_wgrad_line<KY>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024) != cudaSuccess)
      return e2_fail(h, E2_ERR_CUDA, "cudaFuncSetAttribute(c1 wgrad) failed");
    configured = true;
  }
  int per_sm = (int)std::min<size_t>(4, (200 * 1024) / std::max<size_t>(smem, 1));
  per_sm = std::max(1, std::min(per_sm, 2048 / (32 * rows)));
  const int gx = std::max(1, std::min(lines, per_sm * h->sm_count));
  dim3 grid((unsigned)gx, (unsigned)((g.R + 31) / 32));
  k_c1_wgrad_line<KY><<<grid, 32 * rows, smem, s>>>(g, lines);
  h->launches++;
  E2_CUDA_CHECK(h, "conv_c1_wgrad_line");
  return E2_OK;
}

}  // namespace

bool e2_conv_c1_fwd_line_ok(const GatherGemm& g) {
  if (g.K != 1 || g.sz != 1 || g.sx != 1 || g.sy != 1 || g.shuffle || g.gate || g.accumulate) return false;
  if (g.oz != 0 || g.ox != 0 || g.oy != 0) return false;
  if (g.Oz + g.tz - 1 > g.Az || g.Ox + g.tx - 1 > g.Ax || g.Oy + g.ty - 1 > g.Ay) return false;
  if (g.tz * g.tx > C1_MAX_ROWS || g.Oy + g.ty - 1 > C1_MAX_ROW) return false;
  return true;
}

int e2_launch_conv_c1_fwd_line(e2_handle* h, const GatherGemm& g, cudaStream_t s) {
  const int T = g.tz * g.tx * g.ty;
  const int lines = g.On * g.Oz * g.Ox;
  const size_t smem = sizeof(float) * ((size_t)T * 32 + 32 + (size_t)g.tz * g.tx * (g.Oy + g.ty - 1));
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(k_c1_fwd_line, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024) != cudaSuccess)
      return e2_fail(h, E2_ERR_CUDA, "cudaFuncSetAttribute(c1 fwd) failed");
    configured = true;
  }
  if (smem > 100 * 1024) return e2_fail(h, E2_ERR_UNSUPPORTED, "conv c_in==1 fwd: line does not fit shared memory");
  const int gx = std::max(1, std::min(lines, 8 * h->sm_count));
  dim3 grid((unsigned)gx, (unsigned)((g.N + 31) / 32));
  k_c1_fwd_line<<<grid, 256, smem, s>>>(g, lines);
  h->launches++;
  E2_CUDA_CHECK(h, "conv_c1_fwd_line");
  return E2_OK;
}

bool e2_conv_c1_wgrad_line_ok(const ReduceGemm& g) {
  if (g.S != 1 || g.sz != 1 || g.sx != 1 || g.sy != 1 || g.out_mode != 0) return false;
  if (g.ty < 1 || g.ty > 6 || g.tz * g.tx > C1_MAX_ROWS || g.My + g.ty - 1 > C1_MAX_ROW) return false;
  return true;
}

int e2_launch_conv_c1_wgrad_line(e2_handle* h, const ReduceGemm& g, cudaStream_t s) {
  const int T = g.tz * g.tx * g.ty;
  cudaMemsetAsync(g.W, 0, sizeof(float) * (size_t)g.R * T, s);
  switch (g.ty) {
    case 1: return launch_wgrad<1>(h, g, s);
    case 2: return launch_wgrad<2>(h, g, s);
    case 3: return launch_wgrad<3>(h, g, s);
    case 4: return launch_wgrad<4>(h, g, s);
    case 5: return launch_wgrad<5>(h, g, s);
    default: return launch_wgrad<6>(h, g, s);
  }
}
