// First-layer convolution (one input channel, E2_COMPUTE_TF32), warp-specialised.
//
// k_c1_fwd_tc (e2_conv_c1_tc.cu) lets the same 128 threads load the halo, build the im2col tile, issue the MMAs and
// copy the result out, three block-wide barriers per tile; it reaches 1.84 TB/s on unet3d conv0 (0.28 of the HBM
// roofline for a layer that has nothing to do but write 246 MB).  Here each phase has its own warps and they overlap:
//
//   warp 0        TMA producer: the x halo of a tile [kz][8+kx-1][32+ky-1 (+pad)] into a 3-stage ring
//   warps 2-5     builders: im2col rows in the K-major 128B-swizzle layout (row = position, 128 B = 32 taps), a quarter
//                 warp writes the 8 chunks of one row (conflict-free); values tf32-rounded
//   warp 1        MMA issuer: per tile 2 position blocks x ceil(T/8) MMAs M128 x N32 x K8 against the weight tile
//                 (built once per CTA), accumulators double-buffered in TMEM
//   warps 6-13    epilogue: tcgen05.ld -> +bias -> act -> tf32 round -> the thread's 128-byte row into a swizzled staging
//                 tile -> ONE TMA store of the [8 lines][32 y][32 ch] box per tile (clipped at the tensor edges)
// tile = 256 positions (8 x-lines x 32 y of one z plane).
#include <algorithm>
#include <stdlib.h>
#include "e2_common.cuh"
#include "e2_conv_internal.cuh"
#include "e2_tc_ptx.cuh"

namespace {

constexpr int WX = 8, WY = 32, ROWS = WX * WY;
constexpr int WS_THREADS = 448;              // producer, MMA issuer, 4 builder warps, 8 epilogue warps
constexpr int STAGES = 3;
constexpr int A_BYTES = ROWS * 128, OUT_BYTES = ROWS * 128;

struct C1sParams {
  int On, Oz, Ox, Oy;
  int kz, kx, ky, oz, ox, oy;
  int N, T, KS;                 // output channels, taps, K steps (ceil(T/8))
  int XH, YHP, halo_bytes, stage_bytes;
  int ntx, nty, tiles;
  const float* B;               // packed weights [N][b_row], tap stride b_tap
  int64_t b_row, b_tap;
  const float* bias;
  int act, round_tf32;
  uint32_t idesc;
  int dbg;   // E2_C1S_DBG (bottleneck experiments): 1 no halo TMA, 2 no build, 4 no MMA, 8 no epilogue math, 16 no TMA store
};

__device__ __forceinline__ void wait_bar(uint64_t* bar, uint32_t parity) {
  uint32_t n = 0;
  while (!tc::mbar_try_wait(bar, parity)) {
    __nanosleep(64);                 // 13 of the 14 warps wait most of the time: leave the issue slots to the one that works
    if (++n > (1u << 22)) {
      printf("e2b200: conv_c1_ws mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(tc::smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(tc::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }

__global__ void __launch_bounds__(WS_THREADS, 1) k_c1_fwd_ws(const __grid_constant__ CUtensorMap tmX,
                                                             const __grid_constant__ CUtensorMap tmY, const C1sParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  // [STAGES x (A 32 KB | halo)] [2 x out staging 32 KB] [B 4 KB] [bias] [barriers] [tap table]
  uint8_t* smOut = smem + STAGES * p.stage_bytes;
  uint8_t* smB = smOut + 2 * OUT_BYTES;
  float* bias_s = reinterpret_cast<float*>(smB + 32 * 128);
  uint64_t* full = reinterpret_cast<uint64_t*>(bias_s + 32);
  uint64_t* built = full + STAGES;
  uint64_t* empty = built + STAGES;
  uint64_t* acc_full = empty + STAGES;      // [2]
  uint64_t* acc_empty = acc_full + 2;       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  int* tapofs = reinterpret_cast<int*>(tmem_slot + 4);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  int my_tiles = 0;
  for (int t = blockIdx.x; t < p.tiles; t += gridDim.x) ++my_tiles;

  if (threadIdx.x == 0) {
    tc::prefetch_tmap(&tmX);
    tc::prefetch_tmap(&tmY);
    for (int i = 0; i < STAGES; ++i) tc::mbar_init(&full[i], 1), tc::mbar_init(&built[i], 4), tc::mbar_init(&empty[i], 1);
    for (int i = 0; i < 2; ++i) tc::mbar_init(&acc_full[i], 1), tc::mbar_init(&acc_empty[i], 8);
    tc::fence_barrier_init();
  }
  if (threadIdx.x < 32) {
    int o = -1;
    if ((int)threadIdx.x < p.T) {
      const int k = threadIdx.x % p.ky, j = (threadIdx.x / p.ky) % p.kx, i = threadIdx.x / (p.ky * p.kx);
      o = (i * p.XH + j) * p.YHP + k;
    }
    tapofs[threadIdx.x] = o;
    bias_s[threadIdx.x] = (p.bias && (int)threadIdx.x < p.N) ? __ldg(p.bias + threadIdx.x) : 0.f;
  }
  // weights: row n = the 32 (zero-padded) tap values of channel n, K-major 128B swizzle (chunk ^ (row & 7))
  for (int i = threadIdx.x; i < 32 * 32; i += WS_THREADS) {
    const int e = i & 31, n = i >> 5;
    float v = 0.f;
    if (n < p.N && e < p.T) v = __ldg(p.B + (int64_t)n * p.b_row + (int64_t)e * p.b_tap);
    *reinterpret_cast<float*>(smB + n * 128 + ((((e >> 2) ^ (n & 7)) << 4) | ((e & 3) << 2))) = v;
  }
  if (warp == 1) {
    tc::tmem_alloc(tmem_slot, 128u);       // 2 buffers x 2 position blocks x 32 columns
    tc::tmem_relinquish();
  }
  tc::fence_proxy_async();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto tile_coords = [&](int t, int& n, int& z, int& x0, int& y0) {
    const int ity = t % p.nty;
    t /= p.nty;
    const int itx = t % p.ntx;
    t /= p.ntx;
    z = t % p.Oz, n = t / p.Oz;
    x0 = itx * WX, y0 = ity * WY;
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (halo)
    if (lane == 0) {
      int s = 0;
      uint32_t par = 1;
      for (int t = blockIdx.x; t < p.tiles; t += gridDim.x) {
        int n, z, x0, y0;
        tile_coords(t, n, z, x0, y0);
        tc::mbar_wait(&empty[s], par);
        tc::mbar_arrive_expect_tx(&full[s], (p.dbg & 1) ? 0u : (uint32_t)p.halo_bytes);
        if (!(p.dbg & 1)) tma_load_4d(smem + s * p.stage_bytes + A_BYTES, &tmX, &full[s], y0 + p.oy, x0 + p.ox, z + p.oz, n);
        if (++s == STAGES) s = 0, par ^= 1u;
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------------- MMA issuer
    const uint64_t tmpl = tc::make_smem_desc(0, 16, 1024, 2);           // K-major, 128B swizzle, 8-row groups 1024 B apart
    const uint64_t bd0 = tmpl + (uint64_t)(tc::smem_u32(smB) >> 4);
    int s = 0, buf = 0;
    uint32_t par = 0, bpar = 1;
    for (int i = 0; i < my_tiles; ++i) {
      wait_bar(&acc_empty[buf], bpar);                                   // the epilogue has drained this accumulator pair
      wait_bar(&built[s], par);
      tc::tc_fence_after();
      if (tc::elect_one()) {
        const uint32_t a_addr = tc::smem_u32(smem + s * p.stage_bytes);
#pragma unroll
        for (int mb = 0; mb < 2; ++mb) {
          const uint64_t ad0 = tmpl + (uint64_t)((a_addr + (uint32_t)(mb * 128 * 128)) >> 4);
          const uint32_t acc = tmem_base + (uint32_t)((buf * 2 + mb) * 32);
          for (int k = 0; k < ((p.dbg & 4) ? 0 : p.KS); ++k) tc::mma_tf32_ss(acc, ad0 + 2 * k, bd0 + 2 * k, p.idesc, k > 0 ? 1u : 0u);
        }
        tc::mma_commit(&empty[s]);
        tc::mma_commit(&acc_full[buf]);
      }
      __syncwarp();
      if (++s == STAGES) s = 0, par ^= 1u;
      if (++buf == 2) buf = 0, bpar ^= 1u;
    }
  } else if (warp < 6) {
    // ---------------------------------------------------------------------- builders
    const int bt = (int)threadIdx.x - 64;           // 0..127
    const int cj = bt & 7, rg = bt >> 3;             // 16-byte chunk (taps 4cj..4cj+3) of rows rg, rg+16, ...
    int to0 = tapofs[4 * cj], to1 = tapofs[4 * cj + 1], to2 = tapofs[4 * cj + 2], to3 = tapofs[4 * cj + 3];
    const bool t0 = to0 >= 0, t1 = to1 >= 0, t2 = to2 >= 0, t3 = to3 >= 0;
    to0 = max(to0, 0), to1 = max(to1, 0), to2 = max(to2, 0), to3 = max(to3, 0);
    int s = 0;
    uint32_t par = 0;
    for (int i = 0; i < my_tiles; ++i) {
      wait_bar(&full[s], par);
      uint8_t* A = smem + s * p.stage_bytes;
      const float* halo = reinterpret_cast<const float*>(A + A_BYTES);
      if (!(p.dbg & 2))
#pragma unroll
      for (int m0 = 0; m0 < ROWS / 16; m0 += 8) {
        float4 v[8];
#pragma unroll
        for (int m = 0; m < 8; ++m) {
          const int r = rg + 16 * (m0 + m);
          const float* hp = halo + (r >> 5) * p.YHP + (r & 31);
          v[m].x = hp[to0], v[m].y = hp[to1], v[m].z = hp[to2], v[m].w = hp[to3];
        }
#pragma unroll
        for (int m = 0; m < 8; ++m) {
          const int r = rg + 16 * (m0 + m);
          float4 o;
          o.x = t0 ? e2_round_tf32(v[m].x) : 0.f;
          o.y = t1 ? e2_round_tf32(v[m].y) : 0.f;
          o.z = t2 ? e2_round_tf32(v[m].z) : 0.f;
          o.w = t3 ? e2_round_tf32(v[m].w) : 0.f;
          *reinterpret_cast<float4*>(A + r * 128 + ((cj ^ (r & 7)) << 4)) = o;
        }
      }
      tc::fence_proxy_async();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&built[s]);
      if (++s == STAGES) s = 0, par ^= 1u;
    }
  } else {
    // ---------------------------------------------------------------------- epilogue
    // 8 warps: warps 6..9 take position block 0 of a tile, warps 10..13 block 1 (a warp reads the TMEM lane quarter
    // warp % 4, whatever its block)
    const int ew = warp & 3;
    const int mb = (warp - 6) >> 2;
    const int et = (int)threadIdx.x - 192;            // 0..255
    int buf = 0, ob = 0;
    uint32_t fpar = 0;
    int tl = 0;
    for (int t = blockIdx.x; t < p.tiles; t += gridDim.x, ++tl) {
      int n, z, x0, y0;
      tile_coords(t, n, z, x0, y0);
      // the TMA store that last read this staging buffer (two tiles ago) must have finished reading it
      if (et == 0) bulk_wait_read1();
      asm volatile("bar.sync 2, 256;" ::: "memory");
      wait_bar(&acc_full[buf], fpar);
      tc::tc_fence_after();
      uint8_t* out = smOut + ob * OUT_BYTES;
      if (!(p.dbg & 8)) {
        uint32_t r[32];
        tc::tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)((buf * 2 + mb) * 32), r);
        tc::tmem_ld_wait();
        const int row = mb * 128 + ew * 32 + lane;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 b4 = *reinterpret_cast<const float4*>(bias_s + q * 4);
          float v0 = __uint_as_float(r[4 * q]) + b4.x, v1 = __uint_as_float(r[4 * q + 1]) + b4.y;
          float v2 = __uint_as_float(r[4 * q + 2]) + b4.z, v3 = __uint_as_float(r[4 * q + 3]) + b4.w;
          if (p.act == E2_ACT_RELU) {
            v0 = fmaxf(v0, 0.f), v1 = fmaxf(v1, 0.f), v2 = fmaxf(v2, 0.f), v3 = fmaxf(v3, 0.f);
          } else if (p.act != E2_ACT_LIN) {
            v0 = e2_apply_act(v0, p.act), v1 = e2_apply_act(v1, p.act), v2 = e2_apply_act(v2, p.act), v3 = e2_apply_act(v3, p.act);
          }
          if (p.round_tf32) v0 = e2_round_tf32(v0), v1 = e2_round_tf32(v1), v2 = e2_round_tf32(v2), v3 = e2_round_tf32(v3);
          *reinterpret_cast<float4*>(out + row * 128 + ((q ^ (row & 7)) << 4)) = make_float4(v0, v1, v2, v3);
        }
      }
      // accumulators drained: hand the pair back to the MMA warp; staging complete: one TMA store for the tile
      tc::tc_fence_before();
      tc::fence_proxy_async();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&acc_empty[buf]);
      asm volatile("bar.sync 2, 256;" ::: "memory");
      if (et == 0) {
        if (!(p.dbg & 16)) tc::tma_store_5d(&tmY, out, 0, y0, x0, z, n);
        tc::bulk_commit();
      }
      ob ^= 1;
      if (++buf == 2) buf = 0, fpar ^= 1u;
    }
    if (et == 0) tc::bulk_wait0();
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, 128u);
  }
}

bool plan_c1s(const GatherGemm& g, C1sParams* p) {
  if (!e2_get_tmap_encode()) return false;
  if (g.K != 1 || g.a_pitch != 1 || g.shuffle || g.sz != 1 || g.sx != 1 || g.sy != 1) return false;
  if (g.gate || g.accumulate) return false;
  const int T = g.tz * g.tx * g.ty;
  if (T < 9 || T > 32) return false;
  if (g.N < 8 || g.N > 32 || (g.c_pitch & 3)) return false;
  if ((reinterpret_cast<uintptr_t>(g.C) & 15) || (reinterpret_cast<uintptr_t>(g.A) & 15)) return false;
  if (g.Ay % 4) return false;                       // TMA: global strides are multiples of 16 bytes
  if ((int64_t)g.Oz * g.Ox * g.Oy < 4096) return false;
  memset(p, 0, sizeof(*p));
  p->On = g.On, p->Oz = g.Oz, p->Ox = g.Ox, p->Oy = g.Oy;
  p->kz = g.tz, p->kx = g.tx, p->ky = g.ty, p->oz = g.oz, p->ox = g.ox, p->oy = g.oy;
  p->N = g.N, p->T = T, p->KS = (T + 7) / 8;
  p->XH = WX + g.tx - 1;
  p->YHP = (WY + g.ty - 1 + 3) / 4 * 4;
  if (p->YHP > 256 || p->XH > 256 || g.tz > 16) return false;
  p->halo_bytes = g.tz * p->XH * p->YHP * 4;
  p->stage_bytes = (A_BYTES + p->halo_bytes + 1023) / 1024 * 1024;
  p->ntx = (g.Ox + WX - 1) / WX, p->nty = (g.Oy + WY - 1) / WY;
  const int64_t tiles = (int64_t)g.On * g.Oz * p->ntx * p->nty;
  if (tiles >= (1ll << 31)) return false;
  p->tiles = (int)tiles;
  p->B = g.B, p->b_row = g.b_row, p->b_tap = g.b_tap;
  p->bias = g.bias, p->act = g.act, p->round_tf32 = g.round_tf32;
  return STAGES * p->stage_bytes + 2 * OUT_BYTES + 8192 <= 226 * 1024;
}

}  // namespace

bool e2_conv_c1_fwd_ws_ok(const GatherGemm& g) {
  static const bool off = getenv("E2_C1_FWD_OLD") != nullptr;      // A/B switch: keep k_c1_fwd_tc / the CUDA-core kernels
  if (off) return false;
  C1sParams p;
  return plan_c1s(g, &p);
}

int e2_launch_conv_c1_fwd_ws(e2_handle* h, const GatherGemm& g, cudaStream_t s) {
  EncodeTiledFn enc = e2_get_tmap_encode();
  C1sParams p;
  if (!enc || !plan_c1s(g, &p)) return e2_fail(h, E2_ERR_UNSUPPORTED, "conv_c1_fwd_ws: problem does not qualify");
  p.idesc = tc::make_idesc(2 /*TF32*/, 0, 0, 128, 32);
  {
    const char* dv = getenv("E2_C1S_DBG");
    p.dbg = dv ? atoi(dv) : 0;
  }
  CUtensorMap tmX, tmY;
  {
    cuuint64_t dims[4] = {(cuuint64_t)g.Ay, (cuuint64_t)g.Ax, (cuuint64_t)g.Az, (cuuint64_t)g.An};
    cuuint64_t strides[3] = {(cuuint64_t)g.Ay * 4, (cuuint64_t)g.Ay * g.Ax * 4, (cuuint64_t)g.Ay * g.Ax * g.Az * 4};
    cuuint32_t box[4] = {(cuuint32_t)p.YHP, (cuuint32_t)p.XH, (cuuint32_t)p.kz, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&tmX, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(g.A), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return e2_fail(h, E2_ERR_CUDA, "cuTensorMapEncodeTiled(x) failed: %d", (int)r);
  }
  {
    const cuuint64_t pitch = (cuuint64_t)g.c_pitch * 4;
    cuuint64_t dims[5] = {(cuuint64_t)g.N, (cuuint64_t)g.Oy, (cuuint64_t)g.Ox, (cuuint64_t)g.Oz, (cuuint64_t)g.On};
    cuuint64_t strides[4] = {pitch, pitch * g.Oy, pitch * g.Oy * g.Ox, pitch * g.Oy * g.Ox * g.Oz};
    cuuint32_t box[5] = {32, (cuuint32_t)WY, (cuuint32_t)WX, 1, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&tmY, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, g.C, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return e2_fail(h, E2_ERR_CUDA, "cuTensorMapEncodeTiled(y) failed: %d", (int)r);
  }
  const size_t smem = 1024 + (size_t)STAGES * p.stage_bytes + 2 * OUT_BYTES + 32 * 128 + 32 * 4 + (3 * STAGES + 4) * 8 + 16 +
                      32 * 4 + 64;
  if (cudaFuncSetAttribute(k_c1_fwd_ws, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return e2_fail(h, E2_ERR_CUDA, "cudaFuncSetAttribute(max dynamic smem) failed");
  const int grid = std::min(p.tiles, h->sm_count);
  k_c1_fwd_ws<<<grid, WS_THREADS, smem, s>>>(tmX, tmY, p);
  e2_count_launch(h);
  E2_CUDA_CHECK(h, "conv_c1_fwd_ws");
  return E2_OK;
}
