// tcgen05 weight-gradient kernel (E2_COMPUTE_TF32).
//
// reduce-GEMM  W[r][tap][s] = sum_m P[m][r] * Q[pos(m) + tap + org][s]      (conv: P = dy, Q = x)
//
// Per filter tap this is D_tap[s, r] = Q_tap^T P with the reduction running over positions, i.e.
// both operands are "MN-major": a shared-memory row is one position holding 32 consecutive
// channels (128 B).  That is exactly what a channels-last TMA box load produces, so the dy tile
// and the tap-shifted x tiles go from L2 to the tensor core without any transpose.
//
//   M (128 TMEM lanes) = [taps of one MMA group] x [s-channel block]: several taps are stacked
//                        along M when the layer has few input channels (S <= 32: 4 taps,
//                        S <= 64: 2 taps), so small layers still issue full 128-row MMAs
//   N (<= 128 columns) = r-channel chunk of dy
//   A = 4 chunks of (32 channels x 64 positions), one TMA box each (tap shift = box coordinates)
//   B = dy tile, N/32 chunks, loaded once per position tile and reused by every tap group
//   D = TMEM, [group][N] column blocks, fp32, accumulated over the CTA's slice of position tiles
// MN-major tf32 operands exist only in the "128B swizzle, 32B atom" layout (UMMA layout type 1 ==
// TMA SWIZZLE_128B_ATOM_32B).  Partial sums go to dw (reference layout) with fp32 atomics.
#include <algorithm>
#include "e2_common.cuh"
#include "e2_conv_internal.cuh"
#include "e2_tc_ptx.cuh"

namespace {

constexpr int KP = 64;                    // positions per tile (reduction depth per stage)
constexpr int CHUNK_BYTES = KP * 128;     // 32 channels x KP positions
constexpr int A_CHUNKS = 4;               // M = 128 lanes
constexpr int A_BYTES = A_CHUNKS * CHUNK_BYTES;
constexpr int B_SLOTS = 2;
constexpr int WG_THREADS = 192;

struct WgParams {
  int Mn, Mz, Mx, My;      // position grid of P
  int tz, tx, ty;          // position tile box (product KP)
  int ntz, ntx, nty;
  int kz, kx, ky, oz, ox, oy;
  int sz, sx, sy;          // position stride of the Q gather (upconv wgrad: pool factors)
  int R, S;
  int sw;                  // s-channels per tap inside one MMA group: 32, 64 or 128
  int tpm;                 // taps per MMA group = 128 / sw
  int n_cols;              // N = r-chunk width (multiple of 32, <= 128)
  int gpc;                 // groups per CTA = 512 / n_cols (TMEM capacity)
  int n_gs, n_rc, n_sc;    // group sets, r chunks, s chunks
  int a_slots;
  int tmem_cols;
  int tiles_total, tiles_per_split;
  float* W;
  int out_mode;
  uint32_t idesc;
  float* part;             // per-CTA partial tiles [split][unit][group][col/4][row][4]; NULL: fp32 atomics into W
};

__global__ void __launch_bounds__(WG_THREADS) k_wgrad_tc(const __grid_constant__ CUtensorMap tmP,
                                                         const __grid_constant__ CUtensorMap tmQ, const WgParams p) {
  extern __shared__ uint8_t smem_raw[];
  // aligned base by OFFSET arithmetic on smem_raw (no integer round trip), so the pointer keeps its address space and the
  // warps' own accesses compile to LDS / STS instead of generic LD / ST
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  const int b_chunks = p.n_cols / 32;
  const int b_bytes = b_chunks * CHUNK_BYTES;
  uint8_t* smB = smem;
  uint8_t* smA = smem + B_SLOTS * b_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smA + p.a_slots * A_BYTES);
  uint64_t* b_full = bars;                       // [B_SLOTS]
  uint64_t* b_empty = b_full + B_SLOTS;          // [B_SLOTS]
  uint64_t* a_full = b_empty + B_SLOTS;          // [a_slots]
  uint64_t* a_empty = a_full + p.a_slots;        // [a_slots]
  uint64_t* acc_full = a_empty + p.a_slots;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // unit decode: blockIdx.x = ((rc * n_sc) + sc) * n_gs + gs
  int u = blockIdx.x;
  const int gs = u % p.n_gs;
  u /= p.n_gs;
  const int sc = u % p.n_sc;
  const int rc = u / p.n_sc;
  const int T = p.kz * p.kx * p.ky;
  const int tap0 = gs * p.gpc * p.tpm;                       // first tap of this CTA
  const int ntap = min(p.gpc * p.tpm, T - tap0);
  const int ngroups = (ntap + p.tpm - 1) / p.tpm;
  const int r0 = rc * p.n_cols, s0 = sc * p.sw;
  const int spb = p.sw / 32;                                 // 32-channel blocks per tap
  const int t_begin = blockIdx.y * p.tiles_per_split;
  const int t_end = min(t_begin + p.tiles_per_split, p.tiles_total);
  const int ntiles = max(t_end - t_begin, 0);

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmP);
    tc::prefetch_tmap(&tmQ);
    for (int i = 0; i < B_SLOTS; ++i) tc::mbar_init(&b_full[i], 1), tc::mbar_init(&b_empty[i], 1);
    for (int i = 0; i < p.a_slots; ++i) tc::mbar_init(&a_full[i], 1), tc::mbar_init(&a_empty[i], 1);
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
      int ait = 0;
      for (int ti = 0; ti < ntiles; ++ti) {
        int t = t_begin + ti;
        const int ity = t % p.nty;
        t /= p.nty;
        const int itx = t % p.ntx;
        t /= p.ntx;
        const int itz = t % p.ntz;
        const int in_ = t / p.ntz;
        const int z0 = itz * p.tz, x0 = itx * p.tx, y0 = ity * p.ty;
        const int bs = ti % B_SLOTS;
        tc::mbar_wait(&b_empty[bs], ((uint32_t)(ti / B_SLOTS) & 1u) ^ 1u);
        tc::mbar_arrive_expect_tx(&b_full[bs], (uint32_t)b_bytes);
        for (int c = 0; c < b_chunks; ++c)
          tc::tma_load_5d(smB + bs * b_bytes + c * CHUNK_BYTES, &tmP, &b_full[bs], r0 + c * 32, y0, x0, z0, in_);
        for (int g = 0; g < ngroups; ++g, ++ait) {
          const int as = ait % p.a_slots;
          tc::mbar_wait(&a_empty[as], ((uint32_t)(ait / p.a_slots) & 1u) ^ 1u);
          tc::mbar_arrive_expect_tx(&a_full[as], (uint32_t)A_BYTES);
          for (int ci = 0; ci < A_CHUNKS; ++ci) {
            const int tl = g * p.tpm + ci / spb;
            const int sb = ci % spb;
            // taps past the end of the filter: any in-range box, its rows are never stored
            const int tap = min(tap0 + tl, T - 1);
            const int k3 = tap % p.ky, j3 = (tap / p.ky) % p.kx, i3 = tap / (p.ky * p.kx);
            tc::tma_load_5d(smA + as * A_BYTES + ci * CHUNK_BYTES, &tmQ, &a_full[as], s0 + sb * 32,
                            y0 * p.sy + k3 + p.oy, x0 * p.sx + j3 + p.ox, z0 * p.sz + i3 + p.oz, in_);
          }
        }
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      int ait = 0;
      const uint64_t desc_tmpl = tc::make_smem_desc(0, CHUNK_BYTES, 512, 1);
      for (int ti = 0; ti < ntiles; ++ti) {
        const int bs = ti % B_SLOTS;
        tc::mbar_wait(&b_full[bs], (uint32_t)(ti / B_SLOTS) & 1u);
        const uint32_t b_addr = tc::smem_u32(smB + bs * b_bytes);
        for (int g = 0; g < ngroups; ++g, ++ait) {
          const int as = ait % p.a_slots;
          tc::mbar_wait(&a_full[as], (uint32_t)(ait / p.a_slots) & 1u);
          tc::tc_fence_after();
          const uint32_t a_addr = tc::smem_u32(smA + as * A_BYTES);
          // rows (positions) 128 B apart, swizzle period 4 rows = 512 B (SBO), 32-channel chunks
          // CHUNK_BYTES apart (LBO); one MMA consumes K = 8 positions = 1024 B (+64 in the
          // descriptor's 16-byte start-address field)
          const uint64_t ad = desc_tmpl + (uint64_t)(a_addr >> 4);
          const uint64_t bd = desc_tmpl + (uint64_t)(b_addr >> 4);
          const uint32_t acc = tmem_base + (uint32_t)(g * p.n_cols);
#pragma unroll
          for (int k = 0; k < KP / 8; ++k)
            tc::mma_tf32_ss(acc, ad + (uint64_t)(k * 64), bd + (uint64_t)(k * 64), p.idesc, (ti > 0 || k > 0) ? 1u : 0u);
          tc::mma_commit(&a_empty[as]);
        }
        tc::mma_commit(&b_empty[bs]);
      }
      tc::mma_commit(acc_full);
    }
  } else {
    // ---------------------------------------------------------------------- epilogue
    const int q = warp & 3;
    const int row = q * 32 + lane;                  // TMEM lane = (tap_local, s)
    const int tl_in_group = row / p.sw;
    const int s = s0 + row % p.sw;
    tc::mbar_wait(acc_full, 0);
    tc::tc_fence_after();
    if (ntiles > 0) {
      for (int g = 0; g < ngroups; ++g) {
        const int tap = tap0 + g * p.tpm + tl_in_group;
        const bool row_ok = tap < T && tap < tap0 + ntap && s < p.S;
        const int tq = min(tap, T - 1);
        const int k3 = tq % p.ky, j3 = (tq / p.ky) % p.kx, i3 = tq / (p.ky * p.kx);
        const int tflip = ((p.kz - 1 - i3) * p.kx + (p.kx - 1 - j3)) * p.ky + (p.ky - 1 - k3);
        for (int c0 = 0; c0 < p.n_cols; c0 += 32) {
          uint32_t v[32];
          tc::tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * p.n_cols + c0), v);
          tc::tmem_ld_wait();
          if (p.part) {
            // raw accumulators, one float4 per row and 4 columns (coalesced over the rows of a warp)
            float4* dst = reinterpret_cast<float4*>(p.part) +
                          (((size_t)blockIdx.y * gridDim.x + blockIdx.x) * p.gpc + g) * (size_t)(p.n_cols / 4) * 128 + row;
#pragma unroll
            for (int j4 = 0; j4 < 32; j4 += 4)
              dst[(size_t)((c0 + j4) / 4) * 128] = make_float4(__uint_as_float(v[j4]), __uint_as_float(v[j4 + 1]),
                                                              __uint_as_float(v[j4 + 2]), __uint_as_float(v[j4 + 3]));
            continue;
          }
          if (!row_ok) continue;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
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

// Second pass of the partial-tile mode: sums the position splits and writes dw in the reference layout.  Every
// valid (r, s, tap) belongs to exactly one (unit, group, row, column), so W needs no memset and the result
// does not depend on the split order (deterministic).  One thread per (unit, group, 4 columns, row).
__global__ void __launch_bounds__(128) k_wgrad_tc_reduce(const WgParams p, int splits, int units) {
  const int row = threadIdx.x;
  int b = blockIdx.x;
  const int c4 = b % (p.n_cols / 4);
  b /= (p.n_cols / 4);
  const int g = b % p.gpc;
  int u = b / p.gpc;
  const int unit = u;
  const int gs = u % p.n_gs;
  u /= p.n_gs;
  const int sc = u % p.n_sc;
  const int rc = u / p.n_sc;
  const int T = p.kz * p.kx * p.ky;
  const int tap = gs * p.gpc * p.tpm + g * p.tpm + row / p.sw;
  const int s = sc * p.sw + row % p.sw;
  if (tap >= T || s >= p.S) return;
  const size_t tile_f4 = (size_t)(p.n_cols / 4) * 128;
  const float4* src = reinterpret_cast<const float4*>(p.part) + ((size_t)unit * p.gpc + g) * tile_f4 + (size_t)c4 * 128 + row;
  const size_t zstride = (size_t)units * p.gpc * tile_f4;
  float4 a4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int z = 0; z < splits; ++z) {
    const float4 v = __ldcg(src + (size_t)z * zstride);
    a4.x += v.x, a4.y += v.y, a4.z += v.z, a4.w += v.w;
  }
  const int k3 = tap % p.ky, j3 = (tap / p.ky) % p.kx, i3 = tap / (p.ky * p.kx);
  const int tflip = ((p.kz - 1 - i3) * p.kx + (p.kx - 1 - j3)) * p.ky + (p.ky - 1 - k3);
  const float v[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int r = rc * p.n_cols + c4 * 4 + e;
    if (r >= p.R) continue;
    const int64_t ofs = (p.out_mode == 0) ? ((int64_t)r * p.S + s) * T + tflip : ((int64_t)s * p.R + r) * T + tap;
    p.W[ofs] = v[e];
  }
}

bool pick_tile64(int Oz, int Ox, int Oy, int sz, int sx, int sy, int* tz, int* tx, int* ty) {
  static const int opts[][3] = {{1, 8, 8},  {2, 4, 8}, {1, 4, 16}, {4, 4, 4},  {1, 2, 32}, {2, 2, 16}, {1, 16, 4},
                                {1, 1, 64}, {2, 8, 4}, {4, 2, 8},  {2, 1, 32}, {4, 1, 16}, {8, 2, 4},  {8, 1, 8},
                                {1, 32, 2}, {2, 16, 2}, {4, 8, 2}, {8, 4, 2},  {16, 2, 2}, {1, 64, 1}, {2, 32, 1},
                                {4, 16, 1}, {8, 8, 1}, {16, 4, 1}, {16, 1, 4}, {32, 1, 2}, {32, 2, 1}, {64, 1, 1}};
  int64_t best = -1;
  for (auto& o : opts) {
    if (o[0] * sz > 256 || o[1] * sx > 256 || o[2] * sy > 256) continue;   // TMA box extent limit (strided gather)
    int64_t v = (int64_t)((Oz + o[0] - 1) / o[0]) * o[0] * ((Ox + o[1] - 1) / o[1]) * o[1] * ((Oy + o[2] - 1) / o[2]) * o[2];
    if (best < 0 || v < best) best = v, *tz = o[0], *tx = o[1], *ty = o[2];
  }
  return best > 0;
}

}  // namespace

bool e2_reduce_gemm_tc_ok(const e2_handle* h, const ReduceGemm& g) {
  if (!e2_get_tmap_encode()) return false;
  if (g.sz < 1 || g.sx < 1 || g.sy < 1 || g.sz > 8 || g.sx > 8 || g.sy > 8) return false;   // TMA element-stride limit
  if (g.R < 8 || g.S < 8) return false;
  if (g.p_pitch % 4 || g.q_pitch % 4) return false;
  if ((reinterpret_cast<uintptr_t>(g.P) & 15) || (reinterpret_cast<uintptr_t>(g.Q) & 15)) return false;
  return true;
}

// Geometry shared by the launcher and the workspace query.  `partials`: per-CTA partial tiles + reduce kernel
// (a split costs |dw| of coalesced traffic) instead of fp32 atomics (~32 per clock chip-wide when scattered).
static bool plan_wg(int sm_count, const ReduceGemm& g, bool partials, WgParams* pp, int* units_out, int* splits_out,
                    int* a_slots_out, int* b_bytes_out) {
  WgParams& p = *pp;
  memset(&p, 0, sizeof(p));
  p.Mn = g.Mn, p.Mz = g.Mz, p.Mx = g.Mx, p.My = g.My;
  if (!pick_tile64(g.Mz, g.Mx, g.My, g.sz, g.sx, g.sy, &p.tz, &p.tx, &p.ty)) return false;
  p.sz = g.sz, p.sx = g.sx, p.sy = g.sy;
  p.ntz = (g.Mz + p.tz - 1) / p.tz, p.ntx = (g.Mx + p.tx - 1) / p.tx, p.nty = (g.My + p.ty - 1) / p.ty;
  p.kz = g.tz, p.kx = g.tx, p.ky = g.ty, p.oz = g.oz, p.ox = g.ox, p.oy = g.oy;
  p.R = g.R, p.S = g.S;
  const int T = g.tz * g.tx * g.ty;
  p.sw = g.S <= 32 ? 32 : (g.S <= 64 ? 64 : 128);
  p.tpm = 128 / p.sw;
  p.n_sc = (g.S + p.sw - 1) / p.sw;
  int ncols = (g.R + 31) / 32 * 32;
  if (ncols > 128) ncols = 128;
  p.n_cols = ncols;
  p.n_rc = (g.R + ncols - 1) / ncols;
  p.gpc = 512 / ncols;
  const int taps_per_cta = p.gpc * p.tpm;
  p.n_gs = (T + taps_per_cta - 1) / taps_per_cta;
  const int groups_max = (std::min(taps_per_cta, T) + p.tpm - 1) / p.tpm;
  int cols = 32;
  while (cols < groups_max * ncols) cols *= 2;
  p.tmem_cols = cols;
  const int b_bytes = (ncols / 32) * CHUNK_BYTES;
  int a_slots = (int)((208 * 1024 - B_SLOTS * b_bytes) / A_BYTES);
  if (a_slots > 5) a_slots = 5;
  if (a_slots < 2) return false;
  p.a_slots = a_slots;
  p.tiles_total = g.Mn * p.ntz * p.ntx * p.nty;
  const int units = p.n_rc * p.n_sc * p.n_gs;
  // Position splits: more CTAs shorten the MMA phase, every split adds |dw| of output traffic; pick the
  // split count that minimises the sum (cycles).
  int splits = 1;
  {
    const double per_tile = groups_max * (KP / 8) * 70.0;                       // MMA cycles per position tile
    const double out_per_split = (double)units * groups_max * 128.0 * ncols / (partials ? 350.0 : 32.0);
    double best = -1;
    const int smax = std::max(1, std::min(p.tiles_total, (2 * sm_count + units - 1) / units));
    for (int sp = 1; sp <= smax; ++sp) {
      const int tps = (p.tiles_total + sp - 1) / sp;
      const int waves = (units * sp + sm_count - 1) / sm_count;
      const double t = waves * (tps * per_tile + 3000.0) + sp * out_per_split;
      if (best < 0 || t < best) best = t, splits = sp;
    }
  }
  p.tiles_per_split = (p.tiles_total + splits - 1) / splits;
  splits = (p.tiles_total + p.tiles_per_split - 1) / p.tiles_per_split;
  *units_out = units, *splits_out = splits, *a_slots_out = a_slots, *b_bytes_out = b_bytes;
  return true;
}

size_t e2_reduce_gemm_tc_workspace_bytes(int sm_count, const ReduceGemm& g) {
  WgParams p;
  int units, splits, a_slots, b_bytes;
  if (!plan_wg(sm_count, g, true, &p, &units, &splits, &a_slots, &b_bytes)) return 0;
  return (size_t)splits * units * p.gpc * p.n_cols * 128 * sizeof(float);
}

int e2_launch_reduce_gemm_tc(e2_handle* h, const ReduceGemm& g, void* ws, size_t ws_bytes, cudaStream_t s) {
  EncodeTiledFn enc = e2_get_tmap_encode();
  if (!enc) return e2_fail(h, E2_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled entry point not available");
  WgParams p;
  int units, splits, a_slots, b_bytes;
  const int T = g.tz * g.tx * g.ty;
  bool partials = ws && !(reinterpret_cast<uintptr_t>(ws) & 15) && !getenv("E2_WGRAD_ATOMIC");
  if (partials) {
    if (!plan_wg(h->sm_count, g, true, &p, &units, &splits, &a_slots, &b_bytes))
      return e2_fail(h, E2_ERR_UNSUPPORTED, "wgrad_tc: no position tile / shared memory budget");
    if ((size_t)splits * units * p.gpc * p.n_cols * 128 * sizeof(float) > ws_bytes) partials = false;
  }
  if (!partials && !plan_wg(h->sm_count, g, false, &p, &units, &splits, &a_slots, &b_bytes))
    return e2_fail(h, E2_ERR_UNSUPPORTED, "wgrad_tc: no position tile / shared memory budget");
  const int ncols = p.n_cols;
  p.part = partials ? static_cast<float*>(ws) : nullptr;
  p.W = g.W, p.out_mode = g.out_mode;
  p.idesc = tc::make_idesc(2 /*TF32*/, 1, 1, 128, (uint32_t)ncols);

  CUtensorMap tmP, tmQ;
  {
    cuuint64_t dims[5] = {(cuuint64_t)g.R, (cuuint64_t)g.My, (cuuint64_t)g.Mx, (cuuint64_t)g.Mz, (cuuint64_t)g.Mn};
    cuuint64_t pitch = (cuuint64_t)g.p_pitch * 4;
    cuuint64_t strides[4] = {pitch, pitch * g.My, pitch * g.My * g.Mx, pitch * g.My * g.Mx * g.Mz};
    cuuint32_t box[5] = {32, (cuuint32_t)p.ty, (cuuint32_t)p.tx, (cuuint32_t)p.tz, 1};
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
    // strided gather (upconv): traverse ty*sy elements, keep every sy-th -> ty rows
    cuuint32_t box[5] = {32, (cuuint32_t)(p.ty * g.sy), (cuuint32_t)(p.tx * g.sx), (cuuint32_t)(p.tz * g.sz), 1};
    cuuint32_t es[5] = {1, (cuuint32_t)g.sy, (cuuint32_t)g.sx, (cuuint32_t)g.sz, 1};
    CUresult r = enc(&tmQ, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(g.Q), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return e2_fail(h, E2_ERR_CUDA, "cuTensorMapEncodeTiled(Q) failed: %d", (int)r);
  }
  if (!p.part) cudaMemsetAsync(g.W, 0, sizeof(float) * (size_t)g.R * g.S * T, s);
  const size_t smem = 1024 + (size_t)B_SLOTS * b_bytes + (size_t)a_slots * A_BYTES + (2 * B_SLOTS + 2 * a_slots + 1) * 8 + 16;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(k_wgrad_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)) != cudaSuccess)
      return e2_fail(h, E2_ERR_CUDA, "cudaFuncSetAttribute(max dynamic smem) failed");
    configured = true;
  }
  dim3 grid((unsigned)units, (unsigned)splits);
  k_wgrad_tc<<<grid, WG_THREADS, smem, s>>>(tmP, tmQ, p);
  e2_count_launch(h);
  E2_CUDA_CHECK(h, "wgrad_tc");
  if (p.part) {
    k_wgrad_tc_reduce<<<(unsigned)(units * p.gpc * (ncols / 4)), 128, 0, s>>>(p, splits, units);
    e2_count_launch(h);
    E2_CUDA_CHECK(h, "wgrad_tc_reduce");
  }
  return E2_OK;
}
