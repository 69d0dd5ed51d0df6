// tcgen05 weight-gradient kernel with halo reuse (E2_COMPUTE_TF32, unit position stride).
//
// reduce-GEMM  W[r][tap][s] = sum_m P[m][r] * Q[pos(m) + tap + org][s]      (conv: P = dy, Q = x)
//
// The first tcgen05 wgrad kernel (e2_wgrad_tc.cu) loads one shifted copy of the x tile per filter tap:
// 27 taps -> 27 x the x bytes through L2 -> shared memory, and the kernel sat at ~2.4 KB/clk of L2
// traffic (11 % of the tensor peak).  Here the x tile is loaded ONCE with its halo and every tap reads it
// through a shifted UMMA descriptor:
//
//   operands  both MN-major (a shared-memory row = one position x 32 channels = 128 B, the channels-last
//             TMA box as it lands), UMMA layout "128B swizzle / 32B atom"; the swizzle is a function of
//             the shared-memory address bits, so a view may start at any 128-B row
//   tile      TZ x TX x 8 positions; one MMA (K = 8) consumes one y-line of 8 positions
//   x halo    box [TZ][TX+kx-1][8+ky-1] rows of the CTA's z-tap i and 32-channel block
//   M = 128   4 chunks of 32 s-channels with LBO = 128 B: chunk c is the view shifted by c rows = y-tap
//             k = c (ky = 3: chunk 3 is a dummy whose lanes are never stored)
//   N         r-channel chunk of dy (<= 256), dy tile = N/32 boxes, LBO = box bytes
//   D         one TMEM accumulator [128 x N] per (x-tap j, y-tap group): kx * ceil(ky/4) blocks
//   unit      (r chunk, s block, z-tap i) x split over position tiles; partial sums are added to dw with
//             fp32 atomics (reference layout, taps flipped)
// Shared-memory traffic per K step: one dy line (N*32 B) + ~1.9 x-lines, against kx MMAs.
#include <algorithm>
#include <stdlib.h>
#include "e2_common.cuh"
#include "e2_conv_internal.cuh"
#include "e2_tc_ptx.cuh"

namespace {

constexpr int TY = 8;
constexpr int WH_THREADS = 192;
constexpr int MAX_STAGES = 8;

struct WhParams {
  int Mn, Mz, Mx, My;
  int TZ, TX, ntz, ntx, nty;
  int XP, YP;
  int kz, kx, ky, oz, ox, oy;
  int R, S;
  int n_cols, n_chunks, n_rc, n_sc;
  int kgroups, n_acc;
  int stages, dy_chunk_bytes, dy_bytes, x_bytes, x_stride, stage_bytes;
  int tmem_cols;
  int tiles_total, tiles_per_split;
  float* W;
  float* db_ws;   // fused bias gradient: per-(split, r chunk) column sums of dy [splits][n_rc][n_cols], NULL: off
  float* db;      // final bias gradient (reduce kernel)
  float* ws;   // partial-sum workspace [cta][acc][col/4][lane][4]; NULL: fp32 atomics straight into W
  int out_mode;
  uint32_t idesc;
  int dbg;   // bit 0: skip the TMA loads, bit 1: skip the MMAs (E2_WGRAD_DBG, bottleneck experiments only)
};

__device__ __forceinline__ void wait_bar(uint64_t* bar, uint32_t parity) {
  uint32_t n = 0;
  while (!tc::mbar_try_wait(bar, parity)) {
    if (++n > (1u << 24)) {
      printf("e2b200: wgrad_halo mbarrier wait timed out (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y, threadIdx.x);
      __trap();
    }
  }
}

// KX_ / KG_: compile-time x-tap count and y-tap groups (0: run-time values from the parameters)
template <int KX_, int KG_>
__global__ void __launch_bounds__(WH_THREADS) k_wgrad_halo_tc(const __grid_constant__ CUtensorMap tmP,
                                                              const __grid_constant__ CUtensorMap tmQ,
                                                              const WhParams p) {
  extern __shared__ uint8_t smem_raw[];
  // aligned base by OFFSET arithmetic on smem_raw (no integer round trip), so the pointer keeps its address space and the
  // warps' own accesses compile to LDS / STS instead of generic LD / ST
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.stages * p.stage_bytes);
  uint64_t* full = bars;                    // [stages]
  uint64_t* empty = full + MAX_STAGES;      // [stages]
  uint64_t* acc_full = empty + MAX_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  // unit decode: blockIdx.x = ((rc * n_sc) + sc) * kz + i
  int u = blockIdx.x;
  const int i3 = u % p.kz;
  u /= p.kz;
  const int sc = u % p.n_sc;
  const int rc = u / p.n_sc;
  const int r0 = rc * p.n_cols, s0 = sc * 32;
  const int t_begin = blockIdx.y * p.tiles_per_split;
  const int t_end = min(t_begin + p.tiles_per_split, p.tiles_total);
  const int ntiles = max(t_end - t_begin, 0);

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmP);
    tc::prefetch_tmap(&tmQ);
    // CTAs that also sum dy for the bias gradient have a second consumer of each stage (the epilogue warps)
    const uint32_t consumers = (p.db_ws && i3 == 0 && sc == 0) ? 2u : 1u;
    for (int i = 0; i < p.stages; ++i) tc::mbar_init(&full[i], 1), tc::mbar_init(&empty[i], consumers);
    tc::mbar_init(acc_full, 1);
    tc::fence_barrier_init();
  }
  if (warp == 2) {
    tc::tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    tc::tmem_relinquish();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int s = 0;
      uint32_t par = 1;
      for (int ti = 0; ti < ntiles; ++ti) {
        int t = t_begin + ti;
        const int ity = t % p.nty;
        t /= p.nty;
        const int itx = t % p.ntx;
        t /= p.ntx;
        const int itz = t % p.ntz;
        const int in_ = t / p.ntz;
        const int z0 = itz * p.TZ, x0 = itx * p.TX, y0 = ity * TY;
        tc::mbar_wait(&empty[s], par);
        if (p.dbg & 1) {
          tc::mbar_arrive(&full[s]);
          if (++s == p.stages) s = 0, par ^= 1u;
          continue;
        }
        tc::mbar_arrive_expect_tx(&full[s], (uint32_t)(p.dy_bytes + p.x_bytes));
        uint8_t* st = smem + s * p.stage_bytes;
        tc::tma_load_5d(st, &tmQ, &full[s], s0, y0 + p.oy, x0 + p.ox, z0 + i3 + p.oz, in_);
        for (int c = 0; c < p.n_chunks; ++c)
          tc::tma_load_5d(st + p.x_stride + c * p.dy_chunk_bytes, &tmP, &full[s], r0 + c * 32, y0, x0, z0, in_);
        if (++s == p.stages) s = 0, par ^= 1u;
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------------- MMA issuer
    // whole warp walks the tile ring; per tile ONE elected lane issues all MMAs from an unrolled body whose
    // descriptors advance by constants (uniform registers only, ~3 instructions per MMA)
    const int KX = KX_ ? KX_ : p.kx, KG = KX_ ? KG_ : p.kgroups;
    const uint64_t a_tmpl = tc::make_smem_desc(0, 128, 512, 1);
    const uint64_t b_tmpl = tc::make_smem_desc(0, (uint32_t)p.dy_chunk_bytes, 512, 1);
    const uint32_t smem_enc = tc::smem_u32(smem) >> 4;
    const uint32_t stage_enc = (uint32_t)p.stage_bytes >> 4, xs_enc = (uint32_t)p.x_stride >> 4;
    const uint32_t yp_enc = (uint32_t)p.YP * 8;   // one halo x-line = YP rows of 128 B, >> 4
    const uint32_t zskip_enc = (uint32_t)(p.kx - 1) * yp_enc;   // from the last line of a z-plane to the next plane
    const uint32_t ncols = (uint32_t)p.n_cols;
    const uint32_t idesc = p.idesc;
    int s = 0;
    uint32_t par = 0;
    uint32_t accf = 0u;
    for (int ti = 0; ti < ntiles; ++ti) {
      wait_bar(&full[s], par);
      tc::tc_fence_after();
      if (!(p.dbg & 2) && tc::elect_one()) {
        const uint32_t a0 = smem_enc + (uint32_t)s * stage_enc;
        uint64_t bd = b_tmpl + (uint64_t)(a0 + xs_enc);
        uint64_t ad0 = a_tmpl + (uint64_t)a0;
        for (int z = 0; z < p.TZ; ++z) {
          for (int x = 0; x < p.TX; ++x) {
            uint32_t acc = tmem_base;
            uint64_t adj = ad0;
#pragma unroll
            for (int j = 0; j < KX; ++j) {
#pragma unroll
              for (int kg = 0; kg < KG; ++kg) {
                tc::mma_tf32_ss(acc, adj + (uint64_t)(kg * 32), bd, idesc, accf);   // y-tap group: 4 rows = 512 B
                acc += ncols;
              }
              adj += yp_enc;
            }
            accf = 1u;
            bd += 64;          // next dy line: 8 rows = 1024 B
            ad0 += yp_enc;
          }
          ad0 += zskip_enc;
        }
        tc::mma_commit(&empty[s]);
      }
      accf = 1u;
      __syncwarp();
      if (++s == p.stages) s = 0, par ^= 1u;
    }
    if (tc::elect_one()) tc::mma_commit(acc_full);
    __syncwarp();
  } else {
    // ---------------------------------------------------------------------- epilogue
    const int q = warp & 3;
    const int T = p.kz * p.kx * p.ky;
    const int s = s0 + lane;                        // TMEM lane = (chunk q = y-tap within group, s)
    if (p.db_ws && i3 == 0 && sc == 0) {
      // fused bias gradient: while the MMAs run, the epilogue warps sum the dy tile of every stage over its
      // 64 positions (thread = channel; the 32B-atom swizzle keeps a warp's 32 channels on distinct banks)
      const int et = (int)threadIdx.x - 64;         // 0..127
      float sum0 = 0.f, sum1 = 0.f;
      const int c0 = et, c1 = et + 128;
      const uint32_t off0 = (uint32_t)(c0 >> 5) * (uint32_t)p.dy_chunk_bytes, off1 = (uint32_t)(c1 >> 5) * (uint32_t)p.dy_chunk_bytes;
      const uint32_t a0 = (uint32_t)((c0 >> 3) & 3), a1 = (uint32_t)((c1 >> 3) & 3), w0 = (uint32_t)(c0 & 7) * 4u, w1 = (uint32_t)(c1 & 7) * 4u;
      const int rows = p.TZ * p.TX * TY;
      int st = 0;
      uint32_t par = 0;
      for (int ti = 0; ti < ntiles; ++ti) {
        wait_bar(&full[st], par);
        const uint8_t* dyb = smem + st * p.stage_bytes + p.x_stride;
        if (c0 < p.n_cols) {
#pragma unroll 8
          for (int r = 0; r < rows; ++r)
            sum0 += *reinterpret_cast<const float*>(dyb + off0 + r * 128 + ((a0 ^ (uint32_t)(r & 3)) << 5) + w0);
        }
        if (c1 < p.n_cols) {
#pragma unroll 8
          for (int r = 0; r < rows; ++r)
            sum1 += *reinterpret_cast<const float*>(dyb + off1 + r * 128 + ((a1 ^ (uint32_t)(r & 3)) << 5) + w1);
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");      // all four epilogue warps are done with the stage
        if (et == 0) tc::mbar_arrive(&empty[st]);
        if (++st == p.stages) st = 0, par ^= 1u;
      }
      float* dst = p.db_ws + ((size_t)blockIdx.y * p.n_rc + rc) * p.n_cols;
      if (c0 < p.n_cols) dst[c0] = sum0;
      if (c1 < p.n_cols) dst[c1] = sum1;
    }
    tc::mbar_wait(acc_full, 0);
    tc::tc_fence_after();
    if (p.ws) {
      // partial sums of this CTA -> workspace, one coalesced float4 per lane and 4 columns
      float4* dst = reinterpret_cast<float4*>(p.ws) +
                    (size_t)(blockIdx.y * gridDim.x + blockIdx.x) * p.n_acc * (p.n_cols / 4) * 128 + q * 32 + lane;
      for (int a = 0; a < p.n_acc; ++a) {
        for (int c0 = 0; c0 < p.n_cols; c0 += 16) {
          uint32_t v[16];
          tc::tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * p.n_cols + c0), v);
          tc::tmem_ld_wait();
          if (ntiles == 0) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = 0u;
          }
#pragma unroll
          for (int j = 0; j < 4; ++j)
            dst[(size_t)(a * (p.n_cols / 4) + c0 / 4 + j) * 128] =
                make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                            __uint_as_float(v[4 * j + 3]));
        }
      }
    } else if (ntiles > 0) {
      for (int a = 0; a < p.n_acc; ++a) {
        const int j3 = a / p.kgroups, kg = a % p.kgroups;
        const int k3 = kg * 4 + q;
        const bool row_ok = k3 < p.ky && s < p.S;
        const int kq = min(k3, p.ky - 1);
        const int tq = (i3 * p.kx + j3) * p.ky + kq;
        const int tflip = ((p.kz - 1 - i3) * p.kx + (p.kx - 1 - j3)) * p.ky + (p.ky - 1 - kq);
        for (int c0 = 0; c0 < p.n_cols; c0 += 16) {
          uint32_t v[16];
          tc::tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * p.n_cols + c0), v);
          tc::tmem_ld_wait();
          if (!row_ok) continue;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int r = r0 + c0 + j;
            if (r >= p.R) continue;
            const int64_t ofs = (p.out_mode == 0) ? ((int64_t)r * p.S + s) * T + tflip : ((int64_t)s * p.R + r) * T + tq;
            atomicAdd(p.W + ofs, __uint_as_float(v[j]));
          }
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// Sum the per-CTA partial tiles over the position splits and scatter into the reference weight layout.
// One thread per (unit, accumulator, 4 columns, lane); reads are coalesced float4, every dw element is
// written exactly once (no memset, no atomics, deterministic).
__global__ void __launch_bounds__(128) k_wgrad_halo_reduce(const WhParams p, int units, int splits, int w_blocks) {
  if ((int)blockIdx.x >= w_blocks) {
    // bias gradient: sum the per-split column sums
    const int i = ((int)blockIdx.x - w_blocks) * 128 + (int)threadIdx.x;
    if (i >= p.n_rc * p.n_cols) return;
    const int rc = i / p.n_cols, c = i % p.n_cols;
    const int r = rc * p.n_cols + c;
    if (r >= p.R) return;
    float acc = 0.f;
    for (int sp = 0; sp < splits; ++sp) acc += __ldcg(p.db_ws + ((size_t)sp * p.n_rc + rc) * p.n_cols + c);
    p.db[r] = acc;
    return;
  }
  const int lane = threadIdx.x;                       // TMEM lane: (y-tap within group, s)
  int b = blockIdx.x;
  const int c4 = b % (p.n_cols / 4);
  b /= (p.n_cols / 4);
  const int a = b % p.n_acc;
  const int unit = b / p.n_acc;
  int u = unit;
  const int i3 = u % p.kz;
  u /= p.kz;
  const int sc = u % p.n_sc;
  const int rc = u / p.n_sc;
  const int j3 = a / p.kgroups, kg = a % p.kgroups;
  const int k3 = kg * 4 + (lane >> 5);
  const int s = sc * 32 + (lane & 31);
  if (k3 >= p.ky || s >= p.S) return;
  const size_t per_cta = (size_t)p.n_acc * (p.n_cols / 4) * 128;
  const float4* src = reinterpret_cast<const float4*>(p.ws) + (size_t)unit * per_cta + (size_t)(a * (p.n_cols / 4) + c4) * 128 + lane;
  const size_t stride = (size_t)units * per_cta;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  int sp = 0;
  for (; sp + 4 <= splits; sp += 4) {
    const float4 v0 = __ldcg(src + (size_t)sp * stride), v1 = __ldcg(src + (size_t)(sp + 1) * stride);
    const float4 v2 = __ldcg(src + (size_t)(sp + 2) * stride), v3 = __ldcg(src + (size_t)(sp + 3) * stride);
    acc.x += (v0.x + v1.x) + (v2.x + v3.x), acc.y += (v0.y + v1.y) + (v2.y + v3.y);
    acc.z += (v0.z + v1.z) + (v2.z + v3.z), acc.w += (v0.w + v1.w) + (v2.w + v3.w);
  }
  for (; sp < splits; ++sp) {
    const float4 v = __ldcg(src + (size_t)sp * stride);
    acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
  }
  const int T = p.kz * p.kx * p.ky;
  const int tq = (i3 * p.kx + j3) * p.ky + k3;
  const int tflip = ((p.kz - 1 - i3) * p.kx + (p.kx - 1 - j3)) * p.ky + (p.ky - 1 - k3);
  const float v[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int r = rc * p.n_cols + c4 * 4 + e;
    if (r >= p.R) continue;
    const int64_t ofs = (p.out_mode == 0) ? ((int64_t)r * p.S + s) * T + tflip : ((int64_t)s * p.R + r) * T + tq;
    p.W[ofs] = v[e];
  }
}

// Same reduction for the conv layout (out_mode 0), staged through shared memory so that the global writes are
// contiguous: a block owns 4 r-channels x one 32-channel s block and ALL taps, i.e. for each r one run of
// 32*T consecutive floats of dw.  blockDim = 128 * kz (thread = TMEM lane x z-tap).
__global__ void __launch_bounds__(512) k_wgrad_halo_reduce_rows(const WhParams p, int units, int splits, int w_blocks) {
  extern __shared__ float outs[];                    // [4][32][T]
  if ((int)blockIdx.x >= w_blocks) {
    // bias gradient: sum the per-split column sums
    const int i = ((int)blockIdx.x - w_blocks) * (int)blockDim.x + (int)threadIdx.x;
    if (i >= p.n_rc * p.n_cols) return;
    const int rc = i / p.n_cols, c = i % p.n_cols;
    const int r = rc * p.n_cols + c;
    if (r >= p.R) return;
    float acc = 0.f;
    for (int sp = 0; sp < splits; ++sp) acc += __ldcg(p.db_ws + ((size_t)sp * p.n_rc + rc) * p.n_cols + c);
    p.db[r] = acc;
    return;
  }
  const int T = p.kz * p.kx * p.ky;
  const int lane = threadIdx.x & 127, i3 = threadIdx.x >> 7;
  int b = blockIdx.x;
  const int c4 = b % (p.n_cols / 4);
  b /= (p.n_cols / 4);
  const int sc = b % p.n_sc;
  const int rc = b / p.n_sc;
  const int unit = (rc * p.n_sc + sc) * p.kz + i3;
  const size_t per_cta = (size_t)p.n_acc * (p.n_cols / 4) * 128;
  const size_t stride = (size_t)units * per_cta;
  const int sl = lane & 31;
  for (int a = 0; a < p.n_acc; ++a) {
    const int j3 = a / p.kgroups, kg = a % p.kgroups;
    const int k3 = kg * 4 + (lane >> 5);
    if (k3 >= p.ky) continue;
    const float4* src = reinterpret_cast<const float4*>(p.ws) + (size_t)unit * per_cta + (size_t)(a * (p.n_cols / 4) + c4) * 128 + lane;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int sp = 0;
    for (; sp + 4 <= splits; sp += 4) {
      const float4 v0 = __ldcg(src + (size_t)sp * stride), v1 = __ldcg(src + (size_t)(sp + 1) * stride);
      const float4 v2 = __ldcg(src + (size_t)(sp + 2) * stride), v3 = __ldcg(src + (size_t)(sp + 3) * stride);
      acc.x += (v0.x + v1.x) + (v2.x + v3.x), acc.y += (v0.y + v1.y) + (v2.y + v3.y);
      acc.z += (v0.z + v1.z) + (v2.z + v3.z), acc.w += (v0.w + v1.w) + (v2.w + v3.w);
    }
    for (; sp < splits; ++sp) {
      const float4 v = __ldcg(src + (size_t)sp * stride);
      acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
    }
    const int tflip = ((p.kz - 1 - i3) * p.kx + (p.kx - 1 - j3)) * p.ky + (p.ky - 1 - k3);
    outs[(0 * 32 + sl) * T + tflip] = acc.x;
    outs[(1 * 32 + sl) * T + tflip] = acc.y;
    outs[(2 * 32 + sl) * T + tflip] = acc.z;
    outs[(3 * 32 + sl) * T + tflip] = acc.w;
  }
  __syncthreads();
  const int s0 = sc * 32;
  const int ns = min(32, p.S - s0);
  for (int e = 0; e < 4; ++e) {
    const int r = rc * p.n_cols + c4 * 4 + e;
    if (r >= p.R) break;
    float* dst = p.W + ((int64_t)r * p.S + s0) * T;
    for (int i = threadIdx.x; i < ns * T; i += blockDim.x) dst[i] = outs[e * 32 * T + i];
  }
}

int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

bool plan_halo(const e2_handle* h, const ReduceGemm& g, WhParams* p) {
  if (!e2_get_tmap_encode()) return false;
  if (g.sz != 1 || g.sx != 1 || g.sy != 1) return false;
  if (g.R < 8 || g.S < 8) return false;
  if (g.p_pitch % 4 || g.q_pitch % 4) return false;
  if ((reinterpret_cast<uintptr_t>(g.P) & 15) || (reinterpret_cast<uintptr_t>(g.Q) & 15)) return false;
  if (g.ty > 8 || g.tx > 8 || g.tz > 8) return false;
  if (g.tz * g.tx * g.ty < 2) return false;   // 1x1x1: nothing to reuse
  memset(p, 0, sizeof(*p));
  p->Mn = g.Mn, p->Mz = g.Mz, p->Mx = g.Mx, p->My = g.My;
  p->kz = g.tz, p->kx = g.tx, p->ky = g.ty, p->oz = g.oz, p->ox = g.ox, p->oy = g.oy;
  p->R = g.R, p->S = g.S;
  p->kgroups = (g.ty + 3) / 4;
  p->n_acc = g.tx * p->kgroups;
  // N: r chunk, multiple of 16, all accumulators of the CTA must fit 512 TMEM columns
  int nmax = (512 / p->n_acc) / 16 * 16;
  if (nmax > 256) nmax = 256;
  if (nmax < 16) return false;
  p->n_rc = (g.R + nmax - 1) / nmax;
  p->n_cols = ((g.R + p->n_rc - 1) / p->n_rc + 15) / 16 * 16;
  p->n_chunks = (p->n_cols + 31) / 32;
  p->n_sc = (g.S + 31) / 32;
  int cols = 32;
  while (cols < p->n_acc * p->n_cols) cols *= 2;
  p->tmem_cols = cols;
  // position tile: TZ x TX x 8 with TZ*TX = 8 (64 positions), least padded volume
  static const int opts[][2] = {{1, 8}, {2, 4}, {4, 2}, {8, 1}};
  int64_t best = -1;
  const int force_tx = env_int("E2_WGRAD_TX", 0);
  for (auto& o : opts) {
    if (force_tx && o[1] != force_tx) continue;
    // cost: padded positions, weighted by the halo overhead of the x tile
    const int64_t vol = (int64_t)((g.Mz + o[0] - 1) / o[0]) * o[0] * ((g.Mx + o[1] - 1) / o[1]) * o[1];
    const double halo = (double)(o[1] + g.tx - 1) / o[1];
    const int64_t cost = (int64_t)(vol * (p->n_cols * 4.0 + 128.0 * halo * (TY + g.ty - 1) / TY));
    if (best < 0 || cost < best) best = cost, p->TZ = o[0], p->TX = o[1];
  }
  p->ntz = (g.Mz + p->TZ - 1) / p->TZ, p->ntx = (g.Mx + p->TX - 1) / p->TX, p->nty = (g.My + TY - 1) / TY;
  p->XP = p->TX + g.tx - 1;
  p->YP = TY + g.ty - 1;
  p->x_bytes = p->TZ * p->XP * p->YP * 128;
  // the dummy chunks (y-taps >= ky) read up to 7 rows past the last line: keep them inside the stage
  p->x_stride = (p->x_bytes + 8 * 128 + 1023) / 1024 * 1024;
  p->dy_chunk_bytes = p->TZ * p->TX * TY * 128;
  p->dy_bytes = p->n_chunks * p->dy_chunk_bytes;
  p->stage_bytes = p->x_stride + p->dy_bytes;
  p->tiles_total = g.Mn * p->ntz * p->ntx * p->nty;
  return true;
}

}  // namespace

bool e2_wgrad_halo_tc_ok(const e2_handle* h, const ReduceGemm& g) {
  if (env_int("E2_WGRAD_HALO", 1) == 0) return false;
  WhParams p;
  return plan_halo(h, g, &p);
}

// grid plan shared by the launcher and the workspace query
static bool plan_grid(int sm_count, WhParams* p, int* units_out, int* splits_out) {
  // two CTAs per SM when TMEM and shared memory allow (one CTA's epilogue overlaps the other's main loop)
  const int ctas_per_sm = (p->tmem_cols <= 256 && 3 * p->stage_bytes + 2048 <= 112 * 1024) ? 2 : 1;
  const int budget = (ctas_per_sm == 2 ? 112 : 226) * 1024 - 2048;
  p->stages = std::min(MAX_STAGES, budget / p->stage_bytes);
  if (p->stages < 2) return false;
  const int units = p->n_rc * p->n_sc * p->kz;
  int splits = (ctas_per_sm * sm_count) / units;   // floor: never spill one CTA into a second wave
  if (splits > p->tiles_total) splits = p->tiles_total;
  if (splits < 1) splits = 1;
  p->tiles_per_split = (p->tiles_total + splits - 1) / splits;
  splits = (p->tiles_total + p->tiles_per_split - 1) / p->tiles_per_split;
  *units_out = units, *splits_out = splits;
  return true;
}

size_t e2_wgrad_halo_workspace_bytes(int sm_count, const ReduceGemm& g) {
  WhParams p;
  int units, splits;
  if (!plan_halo(nullptr, g, &p) || !plan_grid(sm_count, &p, &units, &splits)) return 0;
  return (size_t)units * splits * p.n_acc * p.n_cols * 128 * sizeof(float) + (size_t)splits * p.n_rc * p.n_cols * sizeof(float);
}

int e2_launch_wgrad_halo_tc(e2_handle* h, const ReduceGemm& g, void* ws, size_t ws_bytes, float* db, bool* db_done,
                            cudaStream_t s) {
  if (db_done) *db_done = false;
  EncodeTiledFn enc = e2_get_tmap_encode();
  if (!enc) return e2_fail(h, E2_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled entry point not available");
  WhParams p;
  if (!plan_halo(h, g, &p)) return e2_fail(h, E2_ERR_UNSUPPORTED, "wgrad_halo_tc: problem does not qualify");
  const int T = g.tz * g.tx * g.ty;
  int units, splits;
  if (!plan_grid(h->sm_count, &p, &units, &splits)) return e2_fail(h, E2_ERR_UNSUPPORTED, "wgrad_halo_tc: shared memory budget");
  const size_t w_part = (size_t)units * splits * p.n_acc * p.n_cols * 128 * sizeof(float);
  const size_t ws_need = w_part + (size_t)splits * p.n_rc * p.n_cols * sizeof(float);
  // with a workspace: per-CTA partial tiles + a reduce kernel (deterministic); without: fp32 atomics
  p.ws = (ws && ws_bytes >= ws_need && !(reinterpret_cast<uintptr_t>(ws) & 15) && env_int("E2_WGRAD_ATOMIC", 0) == 0)
             ? static_cast<float*>(ws) : nullptr;
  p.W = g.W, p.out_mode = g.out_mode;
  // fused bias gradient (sum of P = dy over all positions): only with the workspace path, conv layout
  if (p.ws && db && g.out_mode == 0 && env_int("E2_WGRAD_FUSE_DB", 1)) {
    p.db_ws = reinterpret_cast<float*>(static_cast<uint8_t*>(ws) + w_part);
    p.db = db;
    if (db_done) *db_done = true;
  }
  p.dbg = env_int("E2_WGRAD_DBG", 0);
  p.idesc = tc::make_idesc(2 /*TF32*/, 1, 1, 128, (uint32_t)p.n_cols);

  CUtensorMap tmP, tmQ;
  {
    cuuint64_t dims[5] = {(cuuint64_t)g.R, (cuuint64_t)g.My, (cuuint64_t)g.Mx, (cuuint64_t)g.Mz, (cuuint64_t)g.Mn};
    cuuint64_t pitch = (cuuint64_t)g.p_pitch * 4;
    cuuint64_t strides[4] = {pitch, pitch * g.My, pitch * g.My * g.Mx, pitch * g.My * g.Mx * g.Mz};
    cuuint32_t box[5] = {32, (cuuint32_t)TY, (cuuint32_t)p.TX, (cuuint32_t)p.TZ, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&tmP, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(g.P), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return e2_fail(h, E2_ERR_CUDA, "cuTensorMapEncodeTiled(P) failed: %d", (int)r);
  }
  {
    cuuint64_t dims[5] = {(cuuint64_t)g.S, (cuuint64_t)g.Qy, (cuuint64_t)g.Qx, (cuuint64_t)g.Qz, (cuuint64_t)g.Qn};
    cuuint64_t pitch = (cuuint64_t)g.q_pitch * 4;
    cuuint64_t strides[4] = {pitch, pitch * g.Qy, pitch * g.Qy * g.Qx, pitch * g.Qy * g.Qx * g.Qz};
    cuuint32_t box[5] = {32, (cuuint32_t)p.YP, (cuuint32_t)p.XP, (cuuint32_t)p.TZ, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&tmQ, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(g.Q), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return e2_fail(h, E2_ERR_CUDA, "cuTensorMapEncodeTiled(Q) failed: %d", (int)r);
  }
  if (!p.ws) cudaMemsetAsync(g.W, 0, sizeof(float) * (size_t)g.R * g.S * T, s);
  const size_t smem = 1024 + (size_t)p.stages * p.stage_bytes + (2 * MAX_STAGES + 1) * 8 + 16;
  typedef void (*KernelFn)(const CUtensorMap, const CUtensorMap, const WhParams);
  KernelFn fn = k_wgrad_halo_tc<0, 0>;
  if (p.kgroups == 1) {
    switch (p.kx) {
      case 1: fn = k_wgrad_halo_tc<1, 1>; break;
      case 2: fn = k_wgrad_halo_tc<2, 1>; break;
      case 3: fn = k_wgrad_halo_tc<3, 1>; break;
      case 4: fn = k_wgrad_halo_tc<4, 1>; break;
      default: break;
    }
  } else if (p.kgroups == 2) {
    switch (p.kx) {
      case 1: fn = k_wgrad_halo_tc<1, 2>; break;
      case 5: fn = k_wgrad_halo_tc<5, 2>; break;
      case 6: fn = k_wgrad_halo_tc<6, 2>; break;
      default: break;
    }
  }
  if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)) != cudaSuccess)
    return e2_fail(h, E2_ERR_CUDA, "cudaFuncSetAttribute(max dynamic smem) failed");
  dim3 grid((unsigned)units, (unsigned)splits);
  fn<<<grid, WH_THREADS, smem, s>>>(tmP, tmQ, p);
  e2_count_launch(h);
  E2_CUDA_CHECK(h, "wgrad_halo_tc");
  if (p.ws) {
    const int db_blocks = p.db_ws ? (p.n_rc * p.n_cols + 127) / 128 : 0;
    // few splits: the scattered 4-byte stores dominate -> row-staged variant; many splits (small dw, long
    // sums): keep the variant with one block per (unit, accumulator, 4 columns)
    if (p.out_mode == 0 && p.kz <= 4 && splits <= 8 && (size_t)4 * 32 * T * sizeof(float) <= 48 * 1024 &&
        env_int("E2_WGRAD_REDUCE_ROWS", 1)) {
      const int w_blocks = p.n_rc * p.n_sc * (p.n_cols / 4);
      const int dbb = p.db_ws ? (p.n_rc * p.n_cols + 128 * p.kz - 1) / (128 * p.kz) : 0;
      k_wgrad_halo_reduce_rows<<<(unsigned)(w_blocks + dbb), 128 * p.kz, (size_t)4 * 32 * T * sizeof(float), s>>>(
          p, units, splits, w_blocks);
      e2_count_launch(h);
      E2_CUDA_CHECK(h, "wgrad_halo_reduce_rows");
    } else {
      const int w_blocks = units * p.n_acc * (p.n_cols / 4);
      k_wgrad_halo_reduce<<<(unsigned)(w_blocks + db_blocks), 128, 0, s>>>(p, units, splits, w_blocks);
      e2_count_launch(h);
      E2_CUDA_CHECK(h, "wgrad_halo_reduce");
    }
  }
  return E2_OK;
}
