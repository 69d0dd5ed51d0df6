// First-layer convolutions (one input channel): K per tap is 1, so these layers are HBM-bound and run
// on CUDA cores.  Both kernels work on whole output lines (fixed n, z, x; all y) staged through shared
// memory so that every global access is a coalesced 128-byte run:
//   forward  y[pos][o] = act(sum_tap x[pos+tap] * w[o][tap] + b[o])    8 lanes x float4 = 32 channels / position
//   wgrad    dw[o][tap] = sum_pos dy[pos][o] * x[pos+tap]              warp = one (z,x) tap row, registers hold
//                                                                       [4 channels x ky taps], persistent blocks
// Algorithmic traffic: forward 4*(N_in + N_out) bytes, wgrad 4*(N_dy + N_x) bytes.
#include <algorithm>
#include <stdlib.h>
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


// ------------------------------------------------------------------------------ register-tiled kernels
// Same line decomposition, but the filter taps are compile-time so the loop nest unrolls completely:
//   forward  each thread keeps the weights of its 4 channels for ALL taps in registers (27 x float4) and
//            computes 8 consecutive y positions x 4 channels -> ~9 FMA per shared-memory load
//   wgrad    each thread keeps [taps x 4 channels] accumulators (+ the bias gradient) in registers over
//            all the tiles it visits; dy comes straight from global as coalesced float4
// warp = one (z, x) output line x 32 y positions, lane = yg*8 + q (y group of 8, channel quad);
// block = 8 consecutive x lines; persistent blocks walk (n, z, x-tile, y-tile); edge tiles are shifted
// back inside the tensor (overlapping tiles recompute identical values).
constexpr int RT_X = 8, RT_Y = 32, RT_YT = 8;

// packed fp32x2 FMA (sm_100: FFMA2): d = a * b + c on two lanes per instruction
__device__ __forceinline__ uint64_t pk2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}


template <int KZ, int KX, int KY>
__global__ void __launch_bounds__(256, 3) k_c1_fwd_reg(GatherGemm g, int ntx, int nty, int tiles) {
  constexpr int XW = RT_Y + KY - 1, XR = RT_X + KX - 1;
  __shared__ float xs[2][KZ * XR * XW];
  const int n0 = blockIdx.y * 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = lane & 7, yg = lane >> 3;
  // weights of the block's 32 channels in shared memory: one broadcast LDS.128 per tap feeds the
  // 8 positions x 4 channels of a thread (keeping them in registers costs 108 registers -> 8 warps per SM,
  // and the kernel was latency-bound at 31 % FMA utilisation)
  __shared__ __align__(16) float wsm[KZ * KX * KY][32];
  for (int i = threadIdx.x; i < KZ * KX * KY * 32; i += 256) {
    const int t = i >> 5, n = n0 + (i & 31);
    wsm[t][i & 31] = n < g.N ? __ldg(g.B + (int64_t)n * g.b_row + (int64_t)t * g.b_tap) : 0.f;
  }
  __syncthreads();
  float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (g.bias) {
    float v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) v[e] = (n0 + q * 4 + e < g.N) ? __ldg(g.bias + n0 + q * 4 + e) : 0.f;
    b4 = make_float4(v[0], v[1], v[2], v[3]);
  }
  const bool vec_ok = (g.c_pitch % 4 == 0) && ((reinterpret_cast<uintptr_t>(g.C) & 15) == 0) && (n0 + q * 4 + 3 < g.N);

  auto tile_origin = [&](int t, int& on, int& oz, int& x0, int& y0) {
    const int ity = t % nty;
    t /= nty;
    const int itx = t % ntx;
    t /= ntx;
    oz = t % g.Oz;
    on = t / g.Oz;
    x0 = min(itx * RT_X, g.Ox - RT_X);
    y0 = min(ity * RT_Y, g.Oy - RT_Y);
  };
  auto load_tile = [&](int t, int buf) {
    int on, oz, x0, y0;
    tile_origin(t, on, oz, x0, y0);
    for (int i = threadIdx.x; i < KZ * XR * XW; i += 256) {
      const int yy = i % XW, r = i / XW;
      const int xx = r % XR, zz = r / XR;
      const float* src = g.A + ((((int64_t)on * g.Az + oz + zz) * g.Ax + x0 + xx) * g.Ay + y0 + yy) * g.a_pitch;
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(&xs[buf][i])), "l"(src)
                   : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  int buf = 0;
  if ((int)blockIdx.x < tiles) load_tile(blockIdx.x, 0);
  for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
    const int tn = t + gridDim.x;
    if (tn < tiles) {
      load_tile(tn, buf ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    int on, oz, x0, y0;
    tile_origin(t, on, oz, x0, y0);
    uint64_t alo[RT_YT], ahi[RT_YT];
#pragma unroll
    for (int i = 0; i < RT_YT; ++i) alo[i] = pk2(b4.x, b4.y), ahi[i] = pk2(b4.z, b4.w);
    const float* xb = xs[buf];
#pragma unroll
    for (int zz = 0; zz < KZ; ++zz)
#pragma unroll
      for (int xx = 0; xx < KX; ++xx) {
        const float* row = xb + (zz * XR + warp + xx) * XW + yg * RT_YT;
        uint64_t xv[RT_YT + KY - 1];
#pragma unroll
        for (int i = 0; i < RT_YT + KY - 1; ++i) {
          const float v = row[i];
          xv[i] = pk2(v, v);
        }
#pragma unroll
        for (int k = 0; k < KY; ++k) {
          const float4 w4 = *reinterpret_cast<const float4*>(&wsm[(zz * KX + xx) * KY + k][q * 4]);
          const uint64_t w0 = pk2(w4.x, w4.y), w1 = pk2(w4.z, w4.w);
#pragma unroll
          for (int i = 0; i < RT_YT; ++i) alo[i] = ffma2(xv[i + k], w0, alo[i]), ahi[i] = ffma2(xv[i + k], w1, ahi[i]);
        }
      }
    const int64_t pos0 = (((int64_t)on * g.Oz + oz) * g.Ox + x0 + warp) * g.Oy + y0 + yg * RT_YT;
#pragma unroll
    for (int i = 0; i < RT_YT; ++i) {
      float4 a;
      upk2(alo[i], a.x, a.y);
      upk2(ahi[i], a.z, a.w);
      if (g.act == E2_ACT_RELU) {
        a.x = fmaxf(a.x, 0.f), a.y = fmaxf(a.y, 0.f), a.z = fmaxf(a.z, 0.f), a.w = fmaxf(a.w, 0.f);
      } else if (g.act != E2_ACT_LIN) {
        a.x = e2_apply_act(a.x, g.act), a.y = e2_apply_act(a.y, g.act);
        a.z = e2_apply_act(a.z, g.act), a.w = e2_apply_act(a.w, g.act);
      }
      if (g.round_tf32) a.x = e2_round_tf32(a.x), a.y = e2_round_tf32(a.y), a.z = e2_round_tf32(a.z), a.w = e2_round_tf32(a.w);
      float* out = g.C + (pos0 + i) * g.c_pitch + n0 + q * 4;
      if (vec_ok) {
        *reinterpret_cast<float4*>(out) = a;
      } else {
        const float v[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (n0 + q * 4 + e < g.N) out[e] = v[e];
      }
    }
    __syncthreads();   // the buffer is refilled two iterations later
    buf ^= 1;
  }
}

template <int KZ, int KX, int KY>
__global__ void __launch_bounds__(256, 1) k_c1_wgrad_reg(ReduceGemm g, int ntx, int nty, int tiles, float* db) {
  constexpr int XW = RT_Y + KY - 1, XR = RT_X + KX - 1, T = KZ * KX * KY;
  __shared__ float xs[2][KZ * XR * XW];
  __shared__ float red[8][T + 1][33];     // per warp: [tap | bias][channel]
  const int r0 = blockIdx.y * 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = lane & 7, yg = lane >> 3;
  const bool vec_ok = (g.p_pitch % 4 == 0) && ((reinterpret_cast<uintptr_t>(g.P) & 15) == 0) && (r0 + q * 4 + 3 < g.R);
  uint64_t alo[T], ahi[T];     // packed accumulators: channels (0,1) and (2,3) of the thread's quad
#pragma unroll
  for (int t = 0; t < T; ++t) alo[t] = 0ull, ahi[t] = 0ull;
  float4 accb = make_float4(0.f, 0.f, 0.f, 0.f);

  // tiles are NOT shifted here (a position must be counted once): ragged edges are masked instead
  auto tile_origin = [&](int t, int& mn, int& mz, int& x0, int& y0) {
    const int ity = t % nty;
    t /= nty;
    const int itx = t % ntx;
    t /= ntx;
    mz = t % g.Mz;
    mn = t / g.Mz;
    x0 = itx * RT_X;
    y0 = ity * RT_Y;
  };
  auto load_tile = [&](int t, int buf) {
    int mn, mz, x0, y0;
    tile_origin(t, mn, mz, x0, y0);
    for (int i = threadIdx.x; i < KZ * XR * XW; i += 256) {
      const int yy = i % XW, r = i / XW;
      const int xx = r % XR, zz = r / XR;
      const int qz = mz + zz + g.oz, qx = x0 + xx + g.ox, qy = y0 + yy + g.oy;
      const bool ok = qz >= 0 && qz < g.Qz && qx >= 0 && qx < g.Qx && qy >= 0 && qy < g.Qy;
      const float* src = g.Q + (ok ? ((((int64_t)mn * g.Qz + qz) * g.Qx + qx) * g.Qy + qy) * g.q_pitch : 0);
      const uint32_t nbytes = ok ? 4u : 0u;     // src-size 0: zero fill
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(&xs[buf][i])),
                   "l"(src), "r"(nbytes)
                   : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  auto load_dy = [&](int t, float4* d) {
    int mn, mz, x0, y0;
    tile_origin(t, mn, mz, x0, y0);
    const int mx = x0 + warp;
#pragma unroll
    for (int i = 0; i < RT_YT; ++i) {
      const int my = y0 + yg * RT_YT + i;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (mx < g.Mx && my < g.My) {
        const float* src = g.P + ((((int64_t)mn * g.Mz + mz) * g.Mx + mx) * g.My + my) * g.p_pitch + r0 + q * 4;
        if (vec_ok) {
          v = __ldg(reinterpret_cast<const float4*>(src));
        } else {
          float e[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) e[j] = (r0 + q * 4 + j < g.R) ? __ldg(src + j) : 0.f;
          v = make_float4(e[0], e[1], e[2], e[3]);
        }
      }
      d[i] = v;
    }
  };

  int buf = 0;
  float4 dcur[RT_YT], dnext[RT_YT];
  if ((int)blockIdx.x < tiles) {
    load_tile(blockIdx.x, 0);
    load_dy(blockIdx.x, dcur);
  }
  for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
    const int tn = t + gridDim.x;
    if (tn < tiles) {
      load_tile(tn, buf ^ 1);
      load_dy(tn, dnext);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const float* xb = xs[buf];
#pragma unroll
    for (int i = 0; i < RT_YT; ++i) accb.x += dcur[i].x, accb.y += dcur[i].y, accb.z += dcur[i].z, accb.w += dcur[i].w;
    uint64_t dlo[RT_YT], dhi[RT_YT];
#pragma unroll
    for (int i = 0; i < RT_YT; ++i) dlo[i] = pk2(dcur[i].x, dcur[i].y), dhi[i] = pk2(dcur[i].z, dcur[i].w);
#pragma unroll
    for (int zz = 0; zz < KZ; ++zz)
#pragma unroll
      for (int xx = 0; xx < KX; ++xx) {
        const float* row = xb + (zz * XR + warp + xx) * XW + yg * RT_YT;
        uint64_t xv[RT_YT + KY - 1];
#pragma unroll
        for (int i = 0; i < RT_YT + KY - 1; ++i) {
          const float v = row[i];
          xv[i] = pk2(v, v);
        }
#pragma unroll
        for (int k = 0; k < KY; ++k) {
          uint64_t a0 = alo[(zz * KX + xx) * KY + k], a1 = ahi[(zz * KX + xx) * KY + k];
#pragma unroll
          for (int i = 0; i < RT_YT; ++i) a0 = ffma2(dlo[i], xv[i + k], a0), a1 = ffma2(dhi[i], xv[i + k], a1);
          alo[(zz * KX + xx) * KY + k] = a0, ahi[(zz * KX + xx) * KY + k] = a1;
        }
      }
    __syncthreads();
    buf ^= 1;
#pragma unroll
    for (int i = 0; i < RT_YT; ++i) dcur[i] = dnext[i];
  }
  // reduce: 4 y groups by shuffle, 8 warps through shared memory, then one atomic per (tap, channel) and block
  auto lane_sum = [](float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    return v;
  };
#pragma unroll
  for (int t = 0; t <= T; ++t) {
    float4 a = accb;
    if (t < T) {
      upk2(alo[t < T ? t : 0], a.x, a.y);
      upk2(ahi[t < T ? t : 0], a.z, a.w);
    }
    const float sx = lane_sum(a.x), sy = lane_sum(a.y), sz = lane_sum(a.z), sw = lane_sum(a.w);
    if (yg == 0) red[warp][t][q * 4 + 0] = sx, red[warp][t][q * 4 + 1] = sy, red[warp][t][q * 4 + 2] = sz, red[warp][t][q * 4 + 3] = sw;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < (T + 1) * 32; i += 256) {
    const int t = i >> 5, c = i & 31;
    float v = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) v += red[w8][t][c];
    const int r = r0 + c;
    if (r >= g.R) continue;
    if (t < T) {
      const int k3 = t % KY, j3 = (t / KY) % KX, i3 = t / (KY * KX);
      const int tflip = ((KZ - 1 - i3) * KX + (KX - 1 - j3)) * KY + (KY - 1 - k3);
      atomicAdd(g.W + (int64_t)r * T + tflip, v);
    } else if (db) {
      atomicAdd(db + r, v);
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
  e2_count_launch(h);
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
  e2_count_launch(h);
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

// ---------------------------------------------------------------------- register-tiled launchers
static int reg_variant(int kz, int kx, int ky) {
  if (kz == 3 && kx == 3 && ky == 3) return 1;
  if (kz == 1 && kx == 3 && ky == 3) return 2;
  if (kz == 1 && kx == 4 && ky == 4) return 3;
  return 0;
}

bool e2_conv_c1_fwd_reg_ok(const GatherGemm& g) {
  if (getenv("E2_C1_NOREG")) return false;
  if (!e2_conv_c1_fwd_line_ok(g)) return false;
  if (g.Ox < RT_X || g.Oy < RT_Y) return false;
  return reg_variant(g.tz, g.tx, g.ty) != 0;
}

int e2_launch_conv_c1_fwd_reg(e2_handle* h, const GatherGemm& g, cudaStream_t s) {
  const int ntx = (g.Ox + RT_X - 1) / RT_X, nty = (g.Oy + RT_Y - 1) / RT_Y;
  const int tiles = g.On * g.Oz * ntx * nty;
  dim3 grid((unsigned)std::min(tiles, 3 * h->sm_count), (unsigned)((g.N + 31) / 32));
  switch (reg_variant(g.tz, g.tx, g.ty)) {
    case 1: k_c1_fwd_reg<3, 3, 3><<<grid, 256, 0, s>>>(g, ntx, nty, tiles); break;
    case 2: k_c1_fwd_reg<1, 3, 3><<<grid, 256, 0, s>>>(g, ntx, nty, tiles); break;
    case 3: k_c1_fwd_reg<1, 4, 4><<<grid, 256, 0, s>>>(g, ntx, nty, tiles); break;
    default: return e2_fail(h, E2_ERR_UNSUPPORTED, "conv c_in==1 fwd: no register-tiled variant");
  }
  e2_count_launch(h);
  E2_CUDA_CHECK(h, "conv_c1_fwd_reg");
  return E2_OK;
}

bool e2_conv_c1_wgrad_reg_ok(const ReduceGemm& g) {
  if (getenv("E2_C1_NOREG")) return false;
  if (g.S != 1 || g.sz != 1 || g.sx != 1 || g.sy != 1 || g.out_mode != 0) return false;
  return reg_variant(g.tz, g.tx, g.ty) != 0;
}

// db (nullable): fused bias gradient = column sums of P
int e2_launch_conv_c1_wgrad_reg(e2_handle* h, const ReduceGemm& g, float* db, cudaStream_t s) {
  const int T = g.tz * g.tx * g.ty;
  const int ntx = (g.Mx + RT_X - 1) / RT_X, nty = (g.My + RT_Y - 1) / RT_Y;
  const int tiles = g.Mn * g.Mz * ntx * nty;
  cudaMemsetAsync(g.W, 0, sizeof(float) * (size_t)g.R * T, s);
  if (db) cudaMemsetAsync(db, 0, sizeof(float) * (size_t)g.R, s);
  dim3 grid((unsigned)std::min(tiles, h->sm_count), (unsigned)((g.R + 31) / 32));
  switch (reg_variant(g.tz, g.tx, g.ty)) {
    case 1: k_c1_wgrad_reg<3, 3, 3><<<grid, 256, 0, s>>>(g, ntx, nty, tiles, db); break;
    case 2: k_c1_wgrad_reg<1, 3, 3><<<grid, 256, 0, s>>>(g, ntx, nty, tiles, db); break;
    case 3: k_c1_wgrad_reg<1, 4, 4><<<grid, 256, 0, s>>>(g, ntx, nty, tiles, db); break;
    default: return e2_fail(h, E2_ERR_UNSUPPORTED, "conv c_in==1 wgrad: no register-tiled variant");
  }
  e2_count_launch(h);
  E2_CUDA_CHECK(h, "conv_c1_wgrad_reg");
  return E2_OK;
}
