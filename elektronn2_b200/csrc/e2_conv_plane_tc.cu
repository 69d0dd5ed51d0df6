// tcgen05 implicit-GEMM convolution with shared-memory halo reuse ("plane" kernel).
//
// The first tcgen05 kernel (e2_conv_tc.cu) re-loads the shifted activation tile for every
// filter tap: ~27x redundant L2->SMEM traffic for a 3x3x3 filter, which makes it L2-bound at
// ~17-30 % of the tensor peak.  This kernel loads each input z-plane of the tile's halo ONCE:
//
//   tile     = TZ output z-planes x (16 x-lines x 8 y) positions x BN channels; one accumulator
//              of [128 lanes x BN columns] per output plane, all TZ of them live in TMEM
//   plane    = one TMA box (32 channels, YP = 8+ky-1 y, 16+kx-1 x, 1 z): rows of 128 B in (x,y)
//              order, YP rows per x-line, 128B-swizzled
//   A view   = for tap (i,j,k) and output plane zl the 128 rows needed are 16 groups (x-lines) of
//              8 consecutive rows starting at row (j*YP + k) of input plane zl+i.  A UMMA K-major
//              descriptor expresses exactly that: start address = plane + (j*YP+k)*128 B, stride
//              between 8-row groups (SBO) = YP*128 B.  The 128B swizzle of both TMA and UMMA is a
//              pure function of the shared-memory address bits (measured on B200: views starting
//              at any 128-B row work with base_offset = 0), so no data movement per tap.
//   weights  = [BN x 32] block per (channel block, tap) through a small TMA ring; each block is
//              used by TZ MMAs groups (one per output plane), dividing the weight traffic by TZ
//   schedule = channel block (outer) -> z tap i -> (j,k) taps -> output planes; an input plane is
//              released to the producer as soon as the last z-tap phase that needs it is done
//   roles    = warp 0 plane TMA, warp 1 MMA issuer, warp 2 weight TMA (+TMEM alloc), warps 4-7
//              epilogue.  Persistent CTAs (one per SM) loop over tiles; with 2*TZ*BN <= 512 the
//              accumulators are double-buffered so the epilogue of tile t overlaps tile t+1.
#include <algorithm>
#include <stdlib.h>
#include "e2_common.cuh"
#include "e2_conv_internal.cuh"
#include "e2_tc_ptx.cuh"

namespace {

constexpr int TX = 16, TY = 8;   // tile x/y extent
constexpr int PL_THREADS = 256;
constexpr int MAX_SLOTS = 8;

struct PlaneParams {
  int On, Oz, Ox, Oy;
  int ntz, ntx, nty, ntn;     // tiles per axis, N tiles
  int TZ, NP, XH, YP;
  int kz, kx, ky, oz, ox, oy;
  int K, N, BN, CB;
  int nslot, wslot, acc_bufs;
  int plane_bytes, plane_stride, w_bytes;
  int tmem_cols;
  int num_tiles;
  float* C;
  int c_pitch;
  const float* bias;
  const float* gate;
  int act, accumulate, round_tf32;
  uint32_t idesc;
};

__global__ void __launch_bounds__(PL_THREADS, 1) k_conv_plane_tc(const __grid_constant__ CUtensorMap tmA,
                                                                 const __grid_constant__ CUtensorMap tmB,
                                                                 const PlaneParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smP = smem;                                   // plane slots
  uint8_t* smW = smem + p.nslot * p.plane_stride;        // weight ring
  uint64_t* bars = reinterpret_cast<uint64_t*>(smW + p.wslot * p.w_bytes);
  uint64_t* pl_full = bars;
  uint64_t* pl_empty = pl_full + MAX_SLOTS;
  uint64_t* w_full = pl_empty + MAX_SLOTS;
  uint64_t* w_empty = w_full + MAX_SLOTS;
  uint64_t* acc_full = w_empty + MAX_SLOTS;              // [2]
  uint64_t* acc_empty = acc_full + 2;                    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmA);
    tc::prefetch_tmap(&tmB);
    for (int i = 0; i < p.nslot; ++i) tc::mbar_init(&pl_full[i], 1), tc::mbar_init(&pl_empty[i], 1);
    for (int i = 0; i < p.wslot; ++i) tc::mbar_init(&w_full[i], 1), tc::mbar_init(&w_empty[i], 1);
    for (int i = 0; i < 2; ++i) tc::mbar_init(&acc_full[i], 1), tc::mbar_init(&acc_empty[i], 4);
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
  const int T9 = p.kx * p.ky;

  auto tile_coords = [&](int t, int& in_, int& z0, int& x0, int& y0, int& n0) {
    const int nt = t % p.ntn;
    t /= p.ntn;
    const int ity = t % p.nty;
    t /= p.nty;
    const int itx = t % p.ntx;
    t /= p.ntx;
    const int itz = t % p.ntz;
    in_ = t / p.ntz;
    z0 = itz * p.TZ, x0 = itx * TX, y0 = ity * TY, n0 = nt * p.BN;
  };

  if (warp == 0) {
    // ----------------------------------------------------------- plane producer
    if (lane == 0) {
      int pc = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
        int in_, z0, x0, y0, n0;
        tile_coords(t, in_, z0, x0, y0, n0);
        for (int cb = 0; cb < p.CB; ++cb)
          for (int pl = 0; pl < p.NP; ++pl, ++pc) {
            const int s = pc % p.nslot;
            tc::mbar_wait(&pl_empty[s], ((uint32_t)(pc / p.nslot) & 1u) ^ 1u);
            tc::mbar_arrive_expect_tx(&pl_full[s], (uint32_t)p.plane_bytes);
            tc::tma_load_5d(smP + s * p.plane_stride, &tmA, &pl_full[s], cb * 32, y0 + p.oy, x0 + p.ox,
                            z0 + p.oz + pl, in_);
          }
      }
    }
  } else if (warp == 2) {
    // ---------------------------------------------------------- weight producer
    if (lane == 0) {
      int wc = 0;
      const int T = p.kz * T9;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
        int in_, z0, x0, y0, n0;
        tile_coords(t, in_, z0, x0, y0, n0);
        for (int cb = 0; cb < p.CB; ++cb)
          for (int tap = 0; tap < T; ++tap, ++wc) {
            const int s = wc % p.wslot;
            tc::mbar_wait(&w_empty[s], ((uint32_t)(wc / p.wslot) & 1u) ^ 1u);
            tc::mbar_arrive_expect_tx(&w_full[s], (uint32_t)p.w_bytes);
            tc::tma_load_3d(smW + s * p.w_bytes, &tmB, &w_full[s], cb * 32, tap, n0);
          }
      }
    }
  } else if (warp == 1) {
    // --------------------------------------------------------------- MMA issuer
    // All ring indices / parities are kept incrementally: no integer division in the hot loop.
    if (lane == 0) {
      const uint32_t smP_addr = tc::smem_u32(smP), smW_addr = tc::smem_u32(smW);
      // K-major 128B-swizzle descriptor templates (start address added per MMA):
      // A: 8-row groups YP*128 B apart (shifted view into a halo plane); B: dense, 1024 B apart
      const uint64_t a_tmpl = tc::make_smem_desc(0, 16, (uint32_t)(p.YP * 128), 2);
      const uint64_t b_tmpl = tc::make_smem_desc(0, 16, 1024, 2);
      int pslot = 0;            // slot of plane 0 of the current unit
      uint32_t ppar = 0;        // full-barrier parity of that slot
      int ws = 0;
      uint32_t wpar = 0;
      int buf = 0;
      uint32_t bpar = 1;        // acc_empty parity to wait for (first use passes)
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
        tc::mbar_wait(&acc_empty[buf], bpar);
        tc::tc_fence_after();
        const uint32_t acc0 = tmem_base + (uint32_t)(buf * p.TZ * p.BN);
        for (int cb = 0; cb < p.CB; ++cb) {
          // slot / parity / encoded address of every plane of this unit
          uint32_t pl_enc[MAX_SLOTS], pl_par[MAX_SLOTS];
          int pl_slot[MAX_SLOTS];
          {
            int s = pslot;
            uint32_t par = ppar;
#pragma unroll
            for (int pl = 0; pl < MAX_SLOTS; ++pl) {
              pl_slot[pl] = s, pl_par[pl] = par;
              pl_enc[pl] = (smP_addr + (uint32_t)(s * p.plane_stride)) >> 4;
              if (++s == p.nslot) s = 0, par ^= 1u;
              if (pl + 1 == p.NP) pslot = s, ppar = par;   // first plane of the next unit
            }
          }
          uint32_t waited = 0;
          for (int i = 0; i < p.kz; ++i) {
            int j = 0, k = 0;
            for (int jk = 0; jk < T9; ++jk) {
              tc::mbar_wait(&w_full[ws], wpar);
              const uint64_t bd0 = b_tmpl + (uint64_t)((smW_addr + (uint32_t)(ws * p.w_bytes)) >> 4);
              const uint32_t row_enc = (uint32_t)((j * p.YP + k) * 8);      // (rows * 128 B) >> 4
              const uint32_t first = (cb == 0 && i == 0 && jk == 0) ? 0u : 1u;
#pragma unroll
              for (int zl = 0; zl < 4; ++zl) {
                if (zl < p.TZ) {
                  const int pl = zl + i;
                  if (!(waited & (1u << pl))) {
                    tc::mbar_wait(&pl_full[pl_slot[pl]], pl_par[pl]);
                    waited |= 1u << pl;
                  }
                  tc::tc_fence_after();
                  const uint64_t ad0 = a_tmpl + (uint64_t)(pl_enc[pl] + row_enc);
                  const uint32_t acc = acc0 + (uint32_t)(zl * p.BN);
                  tc::mma_tf32_ss(acc, ad0, bd0, p.idesc, first);
                  tc::mma_tf32_ss(acc, ad0 + 2, bd0 + 2, p.idesc, 1u);
                  tc::mma_tf32_ss(acc, ad0 + 4, bd0 + 4, p.idesc, 1u);
                  tc::mma_tf32_ss(acc, ad0 + 6, bd0 + 6, p.idesc, 1u);
                }
              }
              tc::mma_commit(&w_empty[ws]);
              if (++ws == p.wslot) ws = 0, wpar ^= 1u;
              if (++k == p.ky) k = 0, ++j;
            }
            // input planes whose last z-tap phase was i can go back to the producer
            if (i < p.kz - 1) {
              tc::mma_commit(&pl_empty[pl_slot[i]]);
            } else {
              for (int pl = p.kz - 1; pl < p.NP; ++pl) tc::mma_commit(&pl_empty[pl_slot[pl]]);
            }
          }
        }
        tc::mma_commit(&acc_full[buf]);
        if (++buf == p.acc_bufs) buf = 0, bpar ^= 1u;
      }
    }
  } else if (warp >= 4) {
    // ----------------------------------------------------------------- epilogue
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int lx = row >> 3, ly = row & 7;
    int tcount = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++tcount) {
      int in_, z0, x0, y0, n0;
      tile_coords(t, in_, z0, x0, y0, n0);
      const int buf = tcount % p.acc_bufs;
      tc::mbar_wait(&acc_full[buf], (uint32_t)(tcount / p.acc_bufs) & 1u);
      tc::tc_fence_after();
      const int ox = x0 + lx, oy = y0 + ly;
      for (int zl = 0; zl < p.TZ; ++zl) {
        const int oz = z0 + zl;
        const bool row_ok = oz < p.Oz && ox < p.Ox && oy < p.Oy;
        const int64_t pos = (((int64_t)in_ * p.Oz + oz) * p.Ox + ox) * p.Oy + oy;
        const uint32_t acc = tmem_base + (uint32_t)((buf * p.TZ + zl) * p.BN) + ((uint32_t)(q * 32) << 16);
        for (int c0 = 0; c0 < p.BN; c0 += 32) {
          uint32_t r[32];
          if (p.BN - c0 >= 32) {
            tc::tmem_ld_32x32b_x32(acc + (uint32_t)c0, r);
          } else {
            tc::tmem_ld_32x32b_x16(acc + (uint32_t)c0, r);
#pragma unroll
            for (int jj = 16; jj < 32; ++jj) r[jj] = 0u;
          }
          tc::tmem_ld_wait();
          if (!row_ok) continue;
          float* out = p.C + pos * p.c_pitch + n0 + c0;
          const bool vec = ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
#pragma unroll
          for (int j4 = 0; j4 < 32; j4 += 4) {
            float v[4];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              const int n = n0 + c0 + j4 + jj;
              float a = __uint_as_float(r[j4 + jj]);
              if (n < p.N) {
                if (p.bias) a += __ldg(p.bias + n);
                a = e2_apply_act(a, p.act);
                if (p.gate && !(__ldg(p.gate + (out - p.C) + j4 + jj) > 0.f)) a = 0.f;
                if (p.accumulate) a += out[j4 + jj];
                if (p.round_tf32) a = e2_round_tf32(a);
              }
              v[jj] = a;
            }
            if (vec && n0 + c0 + j4 + 3 < p.N) {
              *reinterpret_cast<float4*>(out + j4) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
              for (int jj = 0; jj < 4; ++jj)
                if (n0 + c0 + j4 + jj < p.N) out[j4 + jj] = v[jj];
            }
          }
        }
      }
      // all TMEM reads of this warp are complete (tcgen05.wait::ld above): hand the buffer back
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&acc_empty[buf]);
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

}  // namespace

// Decide whether the plane kernel applies and is worthwhile for this problem; fill the geometry.
static bool plan_plane(const e2_handle* h, const GatherGemm& g, PlaneParams* p) {
  if (g.sz != 1 || g.sx != 1 || g.sy != 1 || g.shuffle) return false;
  if (g.ty > 9 || g.tx > 9 || g.tz > 8) return false;
  const int T = g.tz * g.tx * g.ty;
  if (T < 2) return false;                       // 1x1x1: nothing to reuse, the tap kernel is fine
  if (g.Ox < 12 || g.Oy < 6) return false;       // tile quantisation would waste too much
  memset(p, 0, sizeof(*p));
  p->XH = TX + g.tx - 1;
  p->YP = TY + g.ty - 1;
  p->plane_bytes = p->XH * p->YP * 128;
  p->plane_stride = (p->plane_bytes + 1023) / 1024 * 1024;   // slots stay 1024-B aligned (swizzle period)
  int bn = (g.N + 15) / 16 * 16;
  if (bn > 256) bn = 256;
  // prefer an N tile that divides N into equal parts
  const int ntn = (g.N + bn - 1) / bn;
  bn = ((g.N + ntn - 1) / ntn + 15) / 16 * 16;
  p->BN = bn, p->ntn = ntn;
  p->w_bytes = bn * 128;
  const int budget = 225 * 1024 - 2048;
  int best_tz = 0;
  for (int tz = 4; tz >= 1; --tz) {
    if (tz * bn > 512) continue;
    const int np = tz + g.tz - 1;
    if (np > MAX_SLOTS) continue;
    const int need = np * p->plane_stride + 3 * p->w_bytes;
    if (need <= budget) {
      best_tz = tz;
      break;
    }
  }
  if (best_tz == 0) return false;
  if (best_tz > g.Oz) best_tz = g.Oz;
  p->TZ = best_tz;
  p->NP = best_tz + g.tz - 1;
  int rest = budget - p->NP * p->plane_stride;
  // spare memory: extra plane slots first (the next unit's first z-tap phase needs TZ planes
  // in flight while the current unit finishes), then weight slots
  p->nslot = p->NP;
  while (p->nslot < MAX_SLOTS && p->nslot < p->NP + p->TZ && rest - p->plane_stride >= 4 * p->w_bytes) {
    p->nslot += 1;
    rest -= p->plane_stride;
  }
  p->wslot = std::min(MAX_SLOTS, rest / p->w_bytes);
  if (p->wslot < 2) return false;
  p->acc_bufs = (2 * p->TZ * bn <= 512) ? 2 : 1;
  int cols = 32;
  while (cols < p->acc_bufs * p->TZ * bn) cols *= 2;
  p->tmem_cols = cols;
  p->On = g.On, p->Oz = g.Oz, p->Ox = g.Ox, p->Oy = g.Oy;
  p->ntz = (g.Oz + p->TZ - 1) / p->TZ, p->ntx = (g.Ox + TX - 1) / TX, p->nty = (g.Oy + TY - 1) / TY;
  p->kz = g.tz, p->kx = g.tx, p->ky = g.ty, p->oz = g.oz, p->ox = g.ox, p->oy = g.oy;
  p->K = g.K, p->N = g.N, p->CB = (g.K + 31) / 32;
  p->num_tiles = p->On * p->ntz * p->ntx * p->nty * p->ntn;
  // tile-quantisation efficiency: useful positions / computed positions
  const double eff = (double)g.Oz * g.Ox * g.Oy / ((double)p->ntz * p->TZ * p->ntx * TX * p->nty * TY);
  if (eff < 0.55) return false;
  return true;
}

bool e2_conv_plane_tc_ok(const e2_handle* h, const GatherGemm& g) {
  PlaneParams p;
  return plan_plane(h, g, &p);
}

int e2_launch_conv_plane_tc(e2_handle* h, const GatherGemm& g, cudaStream_t s) {
  EncodeTiledFn enc = e2_get_tmap_encode();
  if (!enc) return e2_fail(h, E2_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled entry point not available");
  PlaneParams p;
  if (!plan_plane(h, g, &p)) return e2_fail(h, E2_ERR_UNSUPPORTED, "conv_plane_tc: problem does not qualify");
  p.C = g.C, p.c_pitch = g.c_pitch, p.bias = g.bias, p.gate = g.gate;
  p.act = g.act, p.accumulate = g.accumulate, p.round_tf32 = g.round_tf32;
  p.idesc = tc::make_idesc(2 /*TF32*/, 0, 0, 128, (uint32_t)p.BN);
  CUtensorMap tmA, tmB;
  {
    cuuint64_t dims[5] = {(cuuint64_t)g.K, (cuuint64_t)g.Ay, (cuuint64_t)g.Ax, (cuuint64_t)g.Az, (cuuint64_t)g.An};
    cuuint64_t pitch = (cuuint64_t)g.a_pitch * 4;
    cuuint64_t strides[4] = {pitch, pitch * g.Ay, pitch * g.Ay * g.Ax, pitch * g.Ay * g.Ax * g.Az};
    cuuint32_t box[5] = {32, (cuuint32_t)p.YP, (cuuint32_t)p.XH, 1, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(g.A), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return e2_fail(h, E2_ERR_CUDA, "cuTensorMapEncodeTiled(plane) failed: %d", (int)r);
  }
  {
    const int T = g.tz * g.tx * g.ty;
    cuuint64_t dims[3] = {(cuuint64_t)g.K, (cuuint64_t)T, (cuuint64_t)g.N};
    cuuint64_t strides[2] = {(cuuint64_t)g.b_tap * 4, (cuuint64_t)g.b_row * 4};
    cuuint32_t box[3] = {32, 1, (cuuint32_t)p.BN};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(g.B), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return e2_fail(h, E2_ERR_CUDA, "cuTensorMapEncodeTiled(weights) failed: %d", (int)r);
  }
  const size_t smem = 1024 + (size_t)p.nslot * p.plane_stride + (size_t)p.wslot * p.w_bytes + (4 * MAX_SLOTS + 4) * 8 + 16;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(k_conv_plane_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)) !=
        cudaSuccess)
      return e2_fail(h, E2_ERR_CUDA, "cudaFuncSetAttribute(max dynamic smem) failed");
    configured = true;
  }
  const int grid = std::min(p.num_tiles, h->sm_count);
  k_conv_plane_tc<<<grid, PL_THREADS, smem, s>>>(tmA, tmB, p);
  h->launches++;
  E2_CUDA_CHECK(h, "conv_plane_tc");
  return E2_OK;
}
