// tcgen05 weight-gradient kernel with the z-taps stacked along MMA N (E2_COMPUTE_TF32, unit stride, batch 1).
//
//   dw[r][tap][s] = sum_m dy[m][r] * x[pos(m) + tap][s]
//
// e2_wgrad_halo_tc.cu stacks the y-taps along M and takes N = the dy channel chunk.  For layers with few output
// channels (R <= 64: unet3d conv1 / conv2 / mconv4 / mconv5, 59 % of the wgrad FLOPs) that MMA is M128 x N64 x K8:
// it reads (128 + 64) * 32 B of shared memory for 32 cycles of math, and the operand fetch (128 B/clk/SM) needs
// 48 -- measured 48.0 cyc/MMA (profiles/r1_*mma*), i.e. the tensor pipe can be at most 67 % busy.
// Here one MMA multiplies an x line with the dy lines of ALL kz z-taps at once:
//
//   A (M = 128)  x plane q, line X: 4 chunks of 32 s-channels, chunk c = the view shifted by c rows = y-tap c
//                (LBO = 128 B; ky = 3 leaves chunk 3 unused)
//   B (N = kz*NR) dy planes q-kz+1 .. q, line X - j: the dy tile lies in shared memory as [plane][r block][line][8]
//                so the chunks (plane, r block) of one MMA are an arithmetic progression (LBO = one box);
//                column block d holds z-tap i = kz-1-d
//   D            ONE TMEM accumulator [128 x N] per CTA; N = 192 for kz = 3, R = 64: 96 cycles of math against
//                (128 + 192) * 32 B = 80 cycles of operand fetch -> math-bound
//   unit         (r chunk, s block, x-tap j) x split over position tiles (partial tiles + reduce kernel)
// The x tile needs no x halo (j is fixed per CTA: a coordinate offset), the dy tile carries a z halo of kz-1 planes.
// Both operands come straight from the channels-last tensors by TMA ("128B swizzle / 32B atom", MN-major); the dy
// tensor map views the channel axis as (32, r block) so one box lands in the [plane][r block] order.
#include <algorithm>
#include <stdlib.h>
#include "e2_common.cuh"
#include "e2_conv_internal.cuh"
#include "e2_tc_ptx.cuh"

namespace {

constexpr int TY = 8;
constexpr int ZS_THREADS = 192;
constexpr int MAX_STAGES = 8;

struct ZsParams {
  int Mz, Mx, My;            // dy extents (batch 1)
  int TQ, TX, ntq, ntx, nty; // tile: TQ x-planes x TX lines x 8 positions
  int YP;                    // x rows per line: 8 + ky - 1
  int kz, kx, ky, oz, ox, oy;
  int R, S;
  int n_rb, n_rc, n_sc;      // r blocks (of 32) per CTA, r chunks, s blocks
  int N;                     // MMA N = kz * n_rb * 32
  int stages, x_bytes, x_stride, box_bytes, dy_bytes, stage_bytes;
  int tmem_cols;
  int tiles_total, tiles_per_split;
  float* W;
  float* ws;                 // partial tiles [split][unit][col/4][lane][4]
  float* db_ws;              // fused bias gradient [split][rc][sc*kx + j][n_rb*32] (NULL: off)
  float* db;
  int out_mode;
  uint32_t idesc;
};

__device__ __forceinline__ void wait_bar(uint64_t* bar, uint32_t parity) {
  uint32_t n = 0;
  while (!tc::mbar_try_wait(bar, parity)) {
    if (++n > (1u << 24)) {
      printf("e2b200: wgrad_zs mbarrier wait timed out (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y, threadIdx.x);
      __trap();
    }
  }
}

__global__ void __launch_bounds__(ZS_THREADS) k_wgrad_zs_tc(const __grid_constant__ CUtensorMap tmP,
                                                            const __grid_constant__ CUtensorMap tmQ, const ZsParams p) {
  extern __shared__ uint8_t smem_raw[];
  // aligned base by OFFSET arithmetic on smem_raw (no integer round trip), so the pointer keeps its address space and the
  // warps' own accesses compile to LDS / STS instead of generic LD / ST
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.stages * p.stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = full + MAX_STAGES;
  uint64_t* acc_full = empty + MAX_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  // unit decode: blockIdx.x = ((rc * n_sc) + sc) * kx + j
  int u = blockIdx.x;
  const int j3 = u % p.kx;
  u /= p.kx;
  const int sc = u % p.n_sc;
  const int rc = u / p.n_sc;
  const int rb0 = rc * p.n_rb, s0 = sc * 32;
  const int t_begin = blockIdx.y * p.tiles_per_split;
  const int t_end = min(t_begin + p.tiles_per_split, p.tiles_total);
  const int ntiles = max(t_end - t_begin, 0);
  // fused bias gradient: every CTA of an r chunk sums its share of the dy planes (plane index mod #CTAs): the extra
  // work is the same for all CTAs, so they keep walking the tile list in step and share the dy / x tiles through L2
  const bool do_db = p.db_ws != nullptr;
  const int db_grp = p.n_sc * p.kx, db_idx = sc * p.kx + j3;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmP);
    tc::prefetch_tmap(&tmQ);
    for (int i = 0; i < p.stages; ++i) tc::mbar_init(&full[i], 1), tc::mbar_init(&empty[i], do_db ? 2u : 1u);
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
        const int itq = t / p.ntx;
        const int q0 = itq * p.TQ, x0 = itx * p.TX, y0 = ity * TY;
        tc::mbar_wait(&empty[s], par);
        tc::mbar_arrive_expect_tx(&full[s], (uint32_t)(p.x_bytes + p.dy_bytes));
        uint8_t* st = smem + s * p.stage_bytes;
        // x: planes q0 .. q0+TQ-1, lines x0+j .. , rows y0 .. y0+YP-1 of the CTA's 32-channel s block
        tc::tma_load_5d(st, &tmQ, &full[s], s0, y0 + p.oy, x0 + j3 + p.ox, q0 + p.oz, 0);
        // dy: planes q0-kz+1 .. q0+TQ-1 (out-of-range planes / lines / rows arrive as zeros), all r blocks of the chunk
        tc::tma_load_5d(st + p.x_stride, &tmP, &full[s], 0, y0, x0, rb0, q0 - (p.kz - 1));
        if (++s == p.stages) s = 0, par ^= 1u;
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------------- MMA issuer
    const uint64_t a_tmpl = tc::make_smem_desc(0, 128, 512, 1);
    const uint64_t b_tmpl = tc::make_smem_desc(0, (uint32_t)p.box_bytes, 512, 1);
    const uint32_t smem_enc = tc::smem_u32(smem) >> 4;
    const uint32_t stage_enc = (uint32_t)p.stage_bytes >> 4, xs_enc = (uint32_t)p.x_stride >> 4;
    const uint32_t yp_enc = (uint32_t)p.YP * 8;                        // one x line = YP rows of 128 B, >> 4
    const uint32_t plane_enc = (uint32_t)(p.n_rb * p.box_bytes) >> 4;  // one dy plane = n_rb boxes
    const uint32_t idesc = p.idesc;
    int s = 0;
    uint32_t par = 0;
    uint32_t accf = 0u;
    for (int ti = 0; ti < ntiles; ++ti) {
      wait_bar(&full[s], par);
      tc::tc_fence_after();
      if (tc::elect_one()) {
        const uint32_t a0 = smem_enc + (uint32_t)s * stage_enc;
        uint64_t ad = a_tmpl + (uint64_t)a0;
        uint64_t bp = b_tmpl + (uint64_t)(a0 + xs_enc);
        for (int q = 0; q < p.TQ; ++q) {
          uint64_t bd = bp;
          for (int x = 0; x < p.TX; ++x) {
            tc::mma_tf32_ss(tmem_base, ad, bd, idesc, accf);
            accf = 1u;
            ad += yp_enc;
            bd += 64;               // next dy line: 8 rows = 1024 B
          }
          bp += plane_enc;
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
    const int q4 = warp & 3;
    if (do_db) {
      // fused bias gradient: column sums of the dy tile's OWN planes (local planes kz-1 .. kz-1+TQ-1: each dy plane
      // belongs to exactly one tile), thread = channel
      const int et = (int)threadIdx.x - 64;          // 0..127
      const int nch = p.n_rb * 32;
      float sum = 0.f;
      const uint32_t a = (uint32_t)((et >> 3) & 3), w = (uint32_t)(et & 7) * 4u;
      const int rows = p.TX * TY;
      int st = 0;
      uint32_t par = 0;
      for (int ti = 0; ti < ntiles; ++ti) {
        const int q0 = ((t_begin + ti) / (p.nty * p.ntx)) * p.TQ;
        wait_bar(&full[st], par);
        if (et < nch) {
          const uint8_t* dyb = smem + st * p.stage_bytes + p.x_stride + (size_t)(et >> 5) * p.box_bytes;
          for (int o = 0; o < p.TQ; ++o) {
            if ((q0 + o) % db_grp != db_idx) continue;       // dy plane q0 + o is summed by exactly one CTA of the group
            const uint8_t* bx = dyb + (size_t)(p.kz - 1 + o) * p.n_rb * p.box_bytes;
#pragma unroll 8
            for (int r = 0; r < rows; ++r)
              sum += *reinterpret_cast<const float*>(bx + r * 128 + ((a ^ (uint32_t)(r & 3)) << 5) + w);
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (et == 0) tc::mbar_arrive(&empty[st]);
        if (++st == p.stages) st = 0, par ^= 1u;
      }
      if (et < nch) p.db_ws[(((size_t)blockIdx.y * p.n_rc + rc) * db_grp + db_idx) * nch + et] = sum;
    }
    tc::mbar_wait(acc_full, 0);
    tc::tc_fence_after();
    float4* dst = reinterpret_cast<float4*>(p.ws) + (size_t)(blockIdx.y * gridDim.x + blockIdx.x) * (p.N / 4) * 128 + q4 * 32 + lane;
    for (int c0 = 0; c0 < p.N; c0 += 16) {
      uint32_t v[16];
      tc::tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)c0, v);
      tc::tmem_ld_wait();
      if (ntiles == 0) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 0u;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j)
        dst[(size_t)(c0 / 4 + j) * 128] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                                      __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// Sum the partial tiles over the position splits and scatter into the weight layout.  One thread per
// (unit, 4 columns, TMEM lane); every dw element is written exactly once (no memset, no atomics, deterministic).
__global__ void __launch_bounds__(128) k_wgrad_zs_reduce(const ZsParams p, int units, int splits, int w_blocks) {
  if ((int)blockIdx.x >= w_blocks) {
    const int nch = p.n_rb * 32;
    const int i = ((int)blockIdx.x - w_blocks) * 128 + (int)threadIdx.x;
    if (i >= p.n_rc * nch) return;
    const int rc = i / nch, c = i % nch;
    const int r = rc * nch + c;
    if (r >= p.R) return;
    float acc = 0.f;
    const int grp = p.n_sc * p.kx;
    for (int sp = 0; sp < splits; ++sp)
      for (int gi = 0; gi < grp; ++gi) acc += __ldcg(p.db_ws + (((size_t)sp * p.n_rc + rc) * grp + gi) * nch + c);
    p.db[r] = acc;
    return;
  }
  const int lane = threadIdx.x;                      // TMEM lane: (y-tap k, s)
  int b = blockIdx.x;
  const int c4 = b % (p.N / 4);
  const int unit = b / (p.N / 4);
  int u = unit;
  const int j3 = u % p.kx;
  u /= p.kx;
  const int sc = u % p.n_sc;
  const int rc = u / p.n_sc;
  const int k3 = lane >> 5;
  const int s = sc * 32 + (lane & 31);
  if (k3 >= p.ky || s >= p.S) return;
  const size_t per_cta = (size_t)(p.N / 4) * 128;
  const float4* src = reinterpret_cast<const float4*>(p.ws) + (size_t)unit * per_cta + (size_t)c4 * 128 + lane;
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
  const int nch = p.n_rb * 32;
  const int col = c4 * 4;
  const int d = col / nch;                           // column block = dy plane offset -> z-tap kz-1-d
  const int i3 = p.kz - 1 - d;
  const int T = p.kz * p.kx * p.ky;
  const int tq = (i3 * p.kx + j3) * p.ky + k3;
  const int tflip = ((p.kz - 1 - i3) * p.kx + (p.kx - 1 - j3)) * p.ky + (p.ky - 1 - k3);
  const float v[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int r = rc * nch + (col - d * nch) + e;
    if (r >= p.R) continue;
    const int64_t ofs = (p.out_mode == 0) ? ((int64_t)r * p.S + s) * T + tflip : ((int64_t)s * p.R + r) * T + tq;
    p.W[ofs] = v[e];
  }
}

int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

bool plan_zs(const ReduceGemm& g, ZsParams* p) {
  if (!e2_get_tmap_encode()) return false;
  if (g.sz != 1 || g.sx != 1 || g.sy != 1) return false;
  if (g.Mn != 1 || g.Qn != 1) return false;          // the dy map spends its 5th dimension on the r blocks
  if (g.tz < 2 || g.tz > 8 || g.ty > 4 || g.tx > 8) return false;
  if (g.R < 8 || g.S < 8) return false;
  // the dy map reads 32-channel blocks at a 32-float stride: a partial LAST block would read past the channel pitch
  // (into the next position -- harmless -- and, at the last position, past the tensor); R < 32 is one clipped block
  if (g.R > 32 && g.R % 32) return false;
  if (g.p_pitch % 4 || g.q_pitch % 4) return false;
  if ((reinterpret_cast<uintptr_t>(g.P) & 15) || (reinterpret_cast<uintptr_t>(g.Q) & 15)) return false;
  memset(p, 0, sizeof(*p));
  p->Mz = g.Mz, p->Mx = g.Mx, p->My = g.My;
  p->kz = g.tz, p->kx = g.tx, p->ky = g.ty, p->oz = g.oz, p->ox = g.ox, p->oy = g.oy;
  p->R = g.R, p->S = g.S;
  const int rblocks = (g.R + 31) / 32;
  const int max_rb = 256 / (32 * g.tz);
  if (max_rb < 1) return false;
  p->n_rb = std::min(rblocks, max_rb);
  p->n_rc = (rblocks + p->n_rb - 1) / p->n_rb;
  p->n_rb = (rblocks + p->n_rc - 1) / p->n_rc;       // even chunks
  p->N = g.tz * p->n_rb * 32;
  p->n_sc = (g.S + 31) / 32;
  int cols = 32;
  while (cols < p->N) cols *= 2;
  p->tmem_cols = cols;
  p->YP = TY + g.ty - 1;
  // tile: TQ planes x TX lines with TQ * TX = 8; cost = padded lines x bytes per line (dy carries the z halo)
  const int Qz = g.Mz + g.tz - 1;
  static const int opts[][2] = {{8, 1}, {4, 2}, {2, 4}, {1, 8}};
  double best = -1;
  const int force_tq = env_int("E2_WGRAD_ZS_TQ", 0);
  for (auto& o : opts) {
    if (force_tq && o[0] != force_tq) continue;
    const double lines = (double)((Qz + o[0] - 1) / o[0]) * o[0] * ((g.Mx + o[1] - 1) / o[1]) * o[1];
    const double per_line = 128.0 * p->YP + 1024.0 * p->n_rb * (o[0] + g.tz - 1) / o[0];
    const double cost = lines * per_line;
    if (best < 0 || cost < best) best = cost, p->TQ = o[0], p->TX = o[1];
  }
  p->ntq = (Qz + p->TQ - 1) / p->TQ, p->ntx = (g.Mx + p->TX - 1) / p->TX, p->nty = (g.My + TY - 1) / TY;
  p->x_bytes = p->TQ * p->TX * p->YP * 128;
  p->x_stride = (p->x_bytes + 8 * 128 + 1023) / 1024 * 1024;      // the unused y-tap chunk reads up to 3 rows past a line
  p->box_bytes = p->TX * TY * 128;
  p->dy_bytes = (p->TQ + g.tz - 1) * p->n_rb * p->box_bytes;
  p->stage_bytes = p->x_stride + p->dy_bytes;
  p->tiles_total = p->ntq * p->ntx * p->nty;
  return true;
}

bool plan_grid(int sm_count, ZsParams* p, int* units_out, int* splits_out) {
  const int ctas_per_sm = (p->tmem_cols <= 256 && 3 * p->stage_bytes + 2048 <= 112 * 1024) ? 2 : 1;
  const int budget = (ctas_per_sm == 2 ? 112 : 226) * 1024 - 2048;
  p->stages = std::min(MAX_STAGES, budget / p->stage_bytes);
  if (p->stages < 2) return false;
  const int units = p->n_rc * p->n_sc * p->kx;
  int splits = (ctas_per_sm * sm_count) / units;
  if (splits > p->tiles_total) splits = p->tiles_total;
  if (splits < 1) splits = 1;
  p->tiles_per_split = (p->tiles_total + splits - 1) / splits;
  splits = (p->tiles_total + p->tiles_per_split - 1) / p->tiles_per_split;
  *units_out = units, *splits_out = splits;
  return true;
}

}  // namespace

// Taken when the halo kernel's MMA would be operand-fetch-bound (N = R <= 64) and a z extent exists to stack.
bool e2_wgrad_zs_tc_ok(const e2_handle* h, const ReduceGemm& g) {
  if (env_int("E2_WGRAD_ZS", 0) == 0) return false;   // opt-in until it beats the halo kernel (profiles/r2_wgrad_zs.md)
  if (g.R > env_int("E2_WGRAD_ZS_MAXR", 64)) return false;
  ZsParams p;
  int units, splits;
  return plan_zs(g, &p) && plan_grid(h ? h->sm_count : 148, &p, &units, &splits);
}

size_t e2_wgrad_zs_workspace_bytes(int sm_count, const ReduceGemm& g) {
  ZsParams p;
  int units, splits;
  if (!plan_zs(g, &p) || !plan_grid(sm_count, &p, &units, &splits)) return 0;
  return (size_t)units * splits * p.N * 128 * sizeof(float) + (size_t)splits * units * p.n_rb * 32 * sizeof(float);
}

int e2_launch_wgrad_zs_tc(e2_handle* h, const ReduceGemm& g, void* ws, size_t ws_bytes, float* db, bool* db_done,
                          cudaStream_t s) {
  if (db_done) *db_done = false;
  EncodeTiledFn enc = e2_get_tmap_encode();
  if (!enc) return e2_fail(h, E2_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled entry point not available");
  ZsParams p;
  int units, splits;
  if (!plan_zs(g, &p) || !plan_grid(h->sm_count, &p, &units, &splits))
    return e2_fail(h, E2_ERR_UNSUPPORTED, "wgrad_zs_tc: problem does not qualify");
  const size_t w_part = (size_t)units * splits * p.N * 128 * sizeof(float);
  const size_t ws_need = w_part + (size_t)splits * units * p.n_rb * 32 * sizeof(float);
  if (!ws || ws_bytes < ws_need || (reinterpret_cast<uintptr_t>(ws) & 15))
    return e2_fail(h, E2_ERR_WORKSPACE, "wgrad_zs_tc: needs %zu bytes of workspace", ws_need);
  p.ws = static_cast<float*>(ws);
  p.W = g.W, p.out_mode = g.out_mode;
  if (db && g.out_mode == 0) {
    p.db_ws = reinterpret_cast<float*>(static_cast<uint8_t*>(ws) + w_part);
    p.db = db;
    if (db_done) *db_done = true;
  }
  p.idesc = tc::make_idesc(2 /*TF32*/, 1, 1, 128, (uint32_t)p.N);

  CUtensorMap tmP, tmQ;
  {
    // dy: (32 channels, y, x, r block, z); the r-block axis strides 32 floats, so a box is [planes][r blocks][lines][8][32]
    const cuuint64_t pitch = (cuuint64_t)g.p_pitch * 4;
    cuuint64_t dims[5] = {(cuuint64_t)std::min(32, g.R), (cuuint64_t)g.My, (cuuint64_t)g.Mx, (cuuint64_t)((g.R + 31) / 32),
                          (cuuint64_t)g.Mz};
    cuuint64_t strides[4] = {pitch, pitch * g.My, 128, pitch * g.My * g.Mx};
    cuuint32_t box[5] = {32, (cuuint32_t)TY, (cuuint32_t)p.TX, (cuuint32_t)p.n_rb, (cuuint32_t)(p.TQ + p.kz - 1)};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&tmP, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(g.P), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return e2_fail(h, E2_ERR_CUDA, "cuTensorMapEncodeTiled(dy) failed: %d", (int)r);
  }
  {
    cuuint64_t dims[5] = {(cuuint64_t)g.S, (cuuint64_t)g.Qy, (cuuint64_t)g.Qx, (cuuint64_t)g.Qz, (cuuint64_t)g.Qn};
    const cuuint64_t pitch = (cuuint64_t)g.q_pitch * 4;
    cuuint64_t strides[4] = {pitch, pitch * g.Qy, pitch * g.Qy * g.Qx, pitch * g.Qy * g.Qx * g.Qz};
    cuuint32_t box[5] = {32, (cuuint32_t)p.YP, (cuuint32_t)p.TX, (cuuint32_t)p.TQ, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&tmQ, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(g.Q), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return e2_fail(h, E2_ERR_CUDA, "cuTensorMapEncodeTiled(x) failed: %d", (int)r);
  }
  const size_t smem = 1024 + (size_t)p.stages * p.stage_bytes + (2 * MAX_STAGES + 1) * 8 + 16;
  if (cudaFuncSetAttribute(k_wgrad_zs_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)) != cudaSuccess)
    return e2_fail(h, E2_ERR_CUDA, "cudaFuncSetAttribute(max dynamic smem) failed");
  dim3 grid((unsigned)units, (unsigned)splits);
  k_wgrad_zs_tc<<<grid, ZS_THREADS, smem, s>>>(tmP, tmQ, p);
  e2_count_launch(h);
  E2_CUDA_CHECK(h, "wgrad_zs_tc");
  const int w_blocks = units * (p.N / 4);
  const int db_blocks = p.db_ws ? (p.n_rc * p.n_rb * 32 + 127) / 128 : 0;
  k_wgrad_zs_reduce<<<(unsigned)(w_blocks + db_blocks), 128, 0, s>>>(p, units, splits, w_blocks);
  e2_count_launch(h);
  E2_CUDA_CHECK(h, "wgrad_zs_reduce");
  return E2_OK;
}
