// First-layer weight gradient (one input channel, E2_COMPUTE_TF32) on the tensor cores.
//
//   dw[r][tap] = sum_pos dy[pos][r] * x[pos + tap]          db[r] = sum_pos dy[pos][r]
//
// The register-tiled CUDA-core kernel (k_c1_wgrad_reg, e2_conv_c1.cu) issues ~1600 instructions per 8x32-position
// tile and warp, of which 432 are FFMA2: it is issue-bound at 219 us for unet3d conv0 (0.18 of the HBM roofline) and
// runs alone at the very end of the training step.  The work itself is tiny for the tensor pipe; what the layer has to
// do is stream dy (246 MB) once.  Here the threads only BUILD the im2col rows and the MMA does the arithmetic:
//
//   tile   256 positions = 8 x-lines x 32 y of one z plane, seen as 4 blocks of 64 positions (2 lines each)
//   B      dy tile, one TMA box [8 lines][32 y][32 ch]: rows of 128 B, "128B swizzle / 32B atom" (MN-major, N = r)
//   halo   x[kz][8+kx-1][32+ky-1 (+pad)] by a second TMA box (no swizzle; out-of-range voxels arrive as zeros)
//   A      built by 128 threads: row p = the T tap values of position p (tf32-rounded) and zeros behind them; same
//          swizzled layout (MN-major, M = tap).  The same threads sum the dy tile's columns in exact fp32 (bias gradient)
//   MMA    M128 x N128 x K8, kind::tf32: the four 32-row chunks of A and the four 32-column chunks of B are the four
//          position blocks (LBO = 64 rows), so one MMA multiplies 4 x 8 positions; only the diagonal 32x32 blocks of
//          the accumulator (same block on both sides) are meaningful -- 8 MMAs (512 cycles) per tile, against ~1000
//          cycles of shared-memory traffic and ~800 of HBM time: the kernel is bandwidth-bound, as it should be
//   out    per CTA: the four diagonal blocks are added and sent to dw (reference layout, taps flipped) / db with fp32
//          atomics -- the same accumulation the kernel it replaces uses
// 3-stage ring (dy 32 KB + A 32 KB + halo), warp roles: TMA producer / MMA issuer / 4 builder-epilogue warps.
#include <algorithm>
#include <stdlib.h>
#include "e2_common.cuh"
#include "e2_conv_internal.cuh"
#include "e2_tc_ptx.cuh"

namespace {

constexpr int WX = 8, WY = 32;                 // tile: 8 x-lines of 32 y
constexpr int ROWS = WX * WY;                  // 256 positions
constexpr int C1W_THREADS = 192;
constexpr int STAGES = 3;
constexpr int DY_BYTES = ROWS * 128, A_BYTES = ROWS * 128;

struct C1wParams {
  int Mn, Mz, Mx, My;
  int kz, kx, ky, oz, ox, oy;
  int R, T;
  int XH, YHP;                                  // halo lines, padded halo row length (floats, multiple of 4)
  int halo_bytes, stage_bytes;
  int ntx, nty, tiles;
  float* W;
  float* db;
  uint32_t idesc;
  int dbg;   // E2_C1W_DBG (bottleneck experiments only): 1 skip the dy load, 2 skip the halo load, 4 skip the MMAs, 8 skip the build
};

__device__ __forceinline__ void wait_bar(uint64_t* bar, uint32_t parity) {
  uint32_t n = 0;
  while (!tc::mbar_try_wait(bar, parity)) {
    if (++n > (1u << 24)) {
      printf("e2b200: wgrad_c1_tc mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
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

__global__ void __launch_bounds__(C1W_THREADS, 1) k_c1_wgrad_tc(const __grid_constant__ CUtensorMap tmP,
                                                               const __grid_constant__ CUtensorMap tmX, const C1wParams p) {
  extern __shared__ uint8_t smem_raw[];
  // the aligned base keeps its address space for the compiler (offset arithmetic on smem_raw, no integer round trip):
  // the builders' loads / stores must be LDS / STS, not generic LD / ST (3x the latency, measured)
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * p.stage_bytes);   // TMA landed
  uint64_t* built = full + STAGES;                                               // A rows written
  uint64_t* empty = built + STAGES;                                              // MMAs of the stage complete
  uint64_t* acc_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
  int* tapofs = reinterpret_cast<int*>(tmem_slot + 4);                           // [32] halo offset (floats) of tap t

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  int my_tiles = 0;
  for (int t = blockIdx.x; t < p.tiles; t += gridDim.x) ++my_tiles;

  if (threadIdx.x == 0) {
    tc::prefetch_tmap(&tmP);
    tc::prefetch_tmap(&tmX);
    for (int i = 0; i < STAGES; ++i) tc::mbar_init(&full[i], 1), tc::mbar_init(&built[i], 4), tc::mbar_init(&empty[i], 1);
    tc::mbar_init(acc_full, 1);
    tc::fence_barrier_init();
  }
  if (threadIdx.x < 32) {
    int o = -1;
    if ((int)threadIdx.x < p.T) {
      const int k = threadIdx.x % p.ky, j = (threadIdx.x / p.ky) % p.kx, i = threadIdx.x / (p.ky * p.kx);
      o = (i * p.XH + j) * p.YHP + k;
    }
    tapofs[threadIdx.x] = o;
  }
  if (warp == 1) {
    tc::tmem_alloc(tmem_slot, 128u);
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
      for (int t = blockIdx.x; t < p.tiles; t += gridDim.x) {
        int u = t;
        const int ity = u % p.nty;
        u /= p.nty;
        const int itx = u % p.ntx;
        u /= p.ntx;
        const int z = u % p.Mz, n = u / p.Mz;
        tc::mbar_wait(&empty[s], par);
        tc::mbar_arrive_expect_tx(&full[s], (uint32_t)(((p.dbg & 1) ? 0 : DY_BYTES) + ((p.dbg & 2) ? 0 : p.halo_bytes)));
        uint8_t* st = smem + s * p.stage_bytes;
        if (!(p.dbg & 1)) tc::tma_load_5d(st, &tmP, &full[s], 0, ity * WY, itx * WX, z, n);
        if (!(p.dbg & 2)) tma_load_4d(st + DY_BYTES + A_BYTES, &tmX, &full[s], ity * WY + p.oy, itx * WX + p.ox, z + p.oz, n);
        if (++s == STAGES) s = 0, par ^= 1u;
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------------- MMA issuer
    const uint64_t tmpl = tc::make_smem_desc(0, 64 * 128, 512, 1);   // chunk c = position block c: 64 rows further
    const uint32_t smem_enc = tc::smem_u32(smem) >> 4;
    const uint32_t stage_enc = (uint32_t)p.stage_bytes >> 4;
    int s = 0;
    uint32_t par = 0, accf = 0u;
    for (int i = 0; i < my_tiles; ++i) {
      wait_bar(&built[s], par);
      tc::tc_fence_after();
      if (tc::elect_one()) {
        const uint32_t base = smem_enc + (uint32_t)s * stage_enc;
        uint64_t bd = tmpl + (uint64_t)base;                              // dy rows
        uint64_t ad = tmpl + (uint64_t)(base + (DY_BYTES >> 4));         // im2col rows
#pragma unroll
        for (int k = 0; k < 8; ++k) {                                     // 8 rows (K) of each of the 4 blocks per MMA
          if (!(p.dbg & 4)) tc::mma_tf32_ss(tmem_base, ad, bd, p.idesc, accf);
          accf = 1u;
          ad += 64, bd += 64;                                             // 8 rows = 1024 B
        }
        tc::mma_commit(&empty[s]);
      }
      accf = 1u;
      __syncwarp();
      if (++s == STAGES) s = 0, par ^= 1u;
    }
    if (tc::elect_one()) tc::mma_commit(acc_full);
    __syncwarp();
  } else {
    // ------------------------------------------------------- im2col builders, then the epilogue
    const int bt = (int)threadIdx.x - 64;           // 0..127: rows bt and bt + 128
    int s = 0;
    uint32_t par = 0;
    // Thread (row group g = bt >> 3, chunk j = bt & 7) writes the 16-byte chunk j (taps 4j .. 4j+3) of rows g, g+16, ...:
    // the 8 threads of a quarter-warp store the 8 chunks of ONE row = 128 contiguous (swizzled) bytes, conflict-free,
    // and a thread needs only its own 4 tap offsets.  ~10 instructions per chunk, 16 chunks per thread and tile.
    const int cj = bt & 7, rg = bt >> 3;
    int to0 = tapofs[4 * cj], to1 = tapofs[4 * cj + 1], to2 = tapofs[4 * cj + 2], to3 = tapofs[4 * cj + 3];
    const bool t0 = to0 >= 0, t1 = to1 >= 0, t2 = to2 >= 0, t3 = to3 >= 0;
    to0 = max(to0, 0), to1 = max(to1, 0), to2 = max(to2, 0), to3 = max(to3, 0);
    float4 dbsum = make_float4(0.f, 0.f, 0.f, 0.f);    // channels 4cj .. 4cj+3 over this thread's rows
    for (int i = 0; i < my_tiles; ++i) {
      wait_bar(&full[s], par);
      uint8_t* st = smem + s * p.stage_bytes;
      uint8_t* A = st + DY_BYTES;
      const float* halo = reinterpret_cast<const float*>(st + DY_BYTES + A_BYTES);
      if (!(p.dbg & 8)) {
        // 16 chunks per thread, two batches of 8: all 32 halo loads of a batch are issued before the first is used
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
            *reinterpret_cast<float4*>(A + r * 128 + ((((cj >> 1) ^ (r & 3)) << 5) | ((cj & 1) << 4))) = o;
          }
        }
      }
      if (p.db) {
        // bias gradient in exact fp32 from the dy tile (the MMA sees dy through tf32 eyes): the thread's 16-byte chunk
        // (channels 4cj .. 4cj+3) of rows rg, rg+16, ... -- same addressing as the A rows, 16 LDS.128 per tile
        float4 d[16];
#pragma unroll
        for (int m = 0; m < 16; ++m) {
          const int r = rg + 16 * m;
          d[m] = *reinterpret_cast<const float4*>(st + r * 128 + ((((cj >> 1) ^ (r & 3)) << 5) | ((cj & 1) << 4)));
        }
#pragma unroll
        for (int m = 0; m < 16; ++m) dbsum.x += d[m].x, dbsum.y += d[m].y, dbsum.z += d[m].z, dbsum.w += d[m].w;
      }
      tc::fence_proxy_async();                      // generic-proxy stores -> visible to the tensor core's async proxy
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&built[s]);    // one arrival per builder warp
      if (++s == STAGES) s = 0, par ^= 1u;
    }
    // epilogue: warp w (TMEM lanes 32w..) reads its diagonal block = columns 32w..32w+31; lane = tap, register = r
    // (all MMAs have completed and no TMA is in flight once acc_full fires: stage 0 is free to hold the 4 blocks)
    float (*red)[32][33] = reinterpret_cast<float (*)[32][33]>(smem);
    const int w = warp & 3;
    tc::mbar_wait(acc_full, 0);
    tc::tc_fence_after();
    uint32_t rr[32];
    tc::tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(w * 32) << 16) + (uint32_t)(w * 32), rr);
    tc::tmem_ld_wait();
#pragma unroll
    for (int c = 0; c < 32; ++c) red[w][lane][c] = my_tiles > 0 ? __uint_as_float(rr[c]) : 0.f;
    asm volatile("bar.sync 1, 128;" ::: "memory");
    for (int i = bt; i < 32 * 32; i += 128) {
      const int t = i >> 5, c = i & 31;             // tap row, channel
      if (c >= p.R || t >= p.T) continue;
      const float v = (red[0][t][c] + red[1][t][c]) + (red[2][t][c] + red[3][t][c]);
      const int k3 = t % p.ky, j3 = (t / p.ky) % p.kx, i3 = t / (p.ky * p.kx);
      const int tflip = ((p.kz - 1 - i3) * p.kx + (p.kx - 1 - j3)) * p.ky + (p.ky - 1 - k3);
      atomicAdd(p.W + (int64_t)c * p.T + tflip, v);
    }
    if (p.db) {
      float* dbs = &red[0][0][0] + 4 * 32 * 33;       // [16 row groups][32 channels], behind the weight blocks
      *reinterpret_cast<float4*>(dbs + rg * 32 + 4 * cj) = dbsum;
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (bt < p.R) {
        float v = 0.f;
#pragma unroll
        for (int g2 = 0; g2 < 16; ++g2) v += dbs[g2 * 32 + bt];
        atomicAdd(p.db + bt, v);
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, 128u);
  }
}

bool plan_c1w(const ReduceGemm& g, C1wParams* p) {
  if (!e2_get_tmap_encode()) return false;
  if (g.S != 1 || g.q_pitch != 1 || g.out_mode != 0) return false;
  if (g.sz != 1 || g.sx != 1 || g.sy != 1) return false;
  const int T = g.tz * g.tx * g.ty;
  if (T > 32 || g.R > 32 || g.R < 1) return false;            // one 32-row block of taps, one 32-column block of channels
  if (g.p_pitch % 4 || (reinterpret_cast<uintptr_t>(g.P) & 15) || (reinterpret_cast<uintptr_t>(g.Q) & 15)) return false;
  if (g.Qy % 4) return false;                                  // TMA: global strides are multiples of 16 bytes
  memset(p, 0, sizeof(*p));
  p->Mn = g.Mn, p->Mz = g.Mz, p->Mx = g.Mx, p->My = g.My;
  p->kz = g.tz, p->kx = g.tx, p->ky = g.ty, p->oz = g.oz, p->ox = g.ox, p->oy = g.oy;
  p->R = g.R, p->T = T;
  p->XH = WX + g.tx - 1;
  p->YHP = (WY + g.ty - 1 + 3) / 4 * 4;
  if (p->YHP > 256 || p->XH > 256 || g.tz > 16) return false;
  p->halo_bytes = g.tz * p->XH * p->YHP * 4;
  p->stage_bytes = (DY_BYTES + A_BYTES + p->halo_bytes + 1023) / 1024 * 1024;
  if (STAGES * p->stage_bytes + 4096 > 226 * 1024) return false;
  p->ntx = (g.Mx + WX - 1) / WX, p->nty = (g.My + WY - 1) / WY;
  const int64_t tiles = (int64_t)g.Mn * g.Mz * p->ntx * p->nty;
  if (tiles >= (1ll << 31)) return false;
  p->tiles = (int)tiles;
  return true;
}

}  // namespace

bool e2_wgrad_c1_tc_ok(const ReduceGemm& g) {
  static const bool off = getenv("E2_C1_WGRAD_REG") != nullptr;     // A/B switch: keep the CUDA-core kernel
  if (off) return false;
  C1wParams p;
  return plan_c1w(g, &p);
}

int e2_launch_wgrad_c1_tc(e2_handle* h, const ReduceGemm& g, float* db, cudaStream_t s) {
  EncodeTiledFn enc = e2_get_tmap_encode();
  C1wParams p;
  if (!enc || !plan_c1w(g, &p)) return e2_fail(h, E2_ERR_UNSUPPORTED, "wgrad_c1_tc: problem does not qualify");
  p.W = g.W, p.db = db;
  p.idesc = tc::make_idesc(2 /*TF32*/, 1, 1, 128, 128);
  {
    const char* dv = getenv("E2_C1W_DBG");
    p.dbg = dv ? atoi(dv) : 0;
  }
  CUtensorMap tmP, tmX;
  {
    const cuuint64_t pitch = (cuuint64_t)g.p_pitch * 4;
    cuuint64_t dims[5] = {(cuuint64_t)g.R, (cuuint64_t)g.My, (cuuint64_t)g.Mx, (cuuint64_t)g.Mz, (cuuint64_t)g.Mn};
    cuuint64_t strides[4] = {pitch, pitch * g.My, pitch * g.My * g.Mx, pitch * g.My * g.Mx * g.Mz};
    cuuint32_t box[5] = {32, (cuuint32_t)WY, (cuuint32_t)WX, 1, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&tmP, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(g.P), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return e2_fail(h, E2_ERR_CUDA, "cuTensorMapEncodeTiled(dy) failed: %d", (int)r);
  }
  {
    cuuint64_t dims[4] = {(cuuint64_t)g.Qy, (cuuint64_t)g.Qx, (cuuint64_t)g.Qz, (cuuint64_t)g.Qn};
    cuuint64_t strides[3] = {(cuuint64_t)g.Qy * 4, (cuuint64_t)g.Qy * g.Qx * 4, (cuuint64_t)g.Qy * g.Qx * g.Qz * 4};
    cuuint32_t box[4] = {(cuuint32_t)p.YHP, (cuuint32_t)p.XH, (cuuint32_t)p.kz, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&tmX, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(g.Q), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return e2_fail(h, E2_ERR_CUDA, "cuTensorMapEncodeTiled(x) failed: %d", (int)r);
  }
  cudaMemsetAsync(g.W, 0, sizeof(float) * (size_t)g.R * p.T, s);
  if (db) cudaMemsetAsync(db, 0, sizeof(float) * (size_t)g.R, s);
  const size_t smem = 1024 + (size_t)STAGES * p.stage_bytes + (3 * STAGES + 1) * 8 + 16 + 32 * 4 + 64;
  if (cudaFuncSetAttribute(k_c1_wgrad_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return e2_fail(h, E2_ERR_CUDA, "cudaFuncSetAttribute(max dynamic smem) failed");
  const int grid = std::min(p.tiles, h->sm_count);
  k_c1_wgrad_tc<<<grid, C1W_THREADS, smem, s>>>(tmP, tmX, p);
  e2_count_launch(h);
  E2_CUDA_CHECK(h, "wgrad_c1_tc");
  return E2_OK;
}
