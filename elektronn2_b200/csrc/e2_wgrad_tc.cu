// tcgen05 weight-gradient kernel (E2_COMPUTE_TF32).
//
// reduce-GEMM  W[r][tap][s] = sum_m P[m][r] * Q[pos(m) + tap + org][s]      (conv: P = dy, Q = x)
//
// Per filter tap this is D[r, s] = P^T Q_tap with the reduction running over positions, i.e.
// both operands are "MN-major": a shared-memory row is one position holding 32 consecutive
// channels (128 B).  That is exactly what a channels-last TMA box load produces, so P and the
// tap-shifted Q tiles go from HBM/L2 to the tensor core without any transpose:
//   A = P tile   : 4 chunks of 32 r-channels x 64 positions  (MN-major, 128B swizzle / 32B atom)
//   B = Q_tap    : NCH chunks of 32 s-channels x 64 positions
//   D            : TMEM, 128 lanes (r) x [taps of this CTA][NCH*32] columns, fp32
// The P tile is loaded once per position tile and reused for every tap of the CTA's tap group.
// Each CTA owns (r-chunk, s-chunk, tap-group) and a slice of the position tiles (split-K over
// positions); partial sums are added to dw with fp32 atomics in the reference's weight layout.
#include "e2_common.cuh"
#include "e2_conv_internal.cuh"
#include "e2_tc_ptx.cuh"

namespace {

constexpr int KP = 64;                    // positions per tile (reduction depth per stage)
constexpr int CHUNK_BYTES = KP * 128;     // 32 channels x KP positions
constexpr int A_CHUNKS = 4;               // M = 128 r-channels
constexpr int A_BYTES = A_CHUNKS * CHUNK_BYTES;
constexpr int A_SLOTS = 2;
constexpr int WG_THREADS = 192;

struct WgParams {
  int Mn, Mz, Mx, My;      // P position grid
  int tz, tx, ty;          // position tile box (product KP)
  int ntz, ntx, nty;
  int kz, kx, ky, oz, ox, oy;
  int R, S;
  int nch;                 // B chunks (N = 32*nch)
  int tg;                  // taps per CTA
  int n_tg, n_rc, n_sc;    // tap groups, r chunks, s chunks
  int b_slots;
  int tmem_cols;
  int tiles_total, tiles_per_split;
  float* W;
  int out_mode;
  uint32_t idesc;
};

__global__ void __launch_bounds__(WG_THREADS) k_wgrad_tc(const __grid_constant__ CUtensorMap tmP,
                                                         const __grid_constant__ CUtensorMap tmQ, const WgParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int b_bytes = p.nch * CHUNK_BYTES;
  uint8_t* smA = smem;
  uint8_t* smB = smem + A_SLOTS * A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smB + p.b_slots * b_bytes);
  uint64_t* a_full = bars;                       // [A_SLOTS]
  uint64_t* a_empty = a_full + A_SLOTS;          // [A_SLOTS]
  uint64_t* b_full = a_empty + A_SLOTS;          // [b_slots]
  uint64_t* b_empty = b_full + p.b_slots;        // [b_slots]
  uint64_t* acc_full = b_empty + p.b_slots;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // unit decode: blockIdx.x = ((rc * n_sc) + sc) * n_tg + tgi
  int u = blockIdx.x;
  const int tgi = u % p.n_tg;
  u /= p.n_tg;
  const int sc = u % p.n_sc;
  const int rc = u / p.n_sc;
  const int T = p.kz * p.kx * p.ky;
  const int tap0 = tgi * p.tg;
  const int ntap = min(p.tg, T - tap0);
  const int r0 = rc * 128, s0 = sc * p.nch * 32;
  const int t_begin = blockIdx.y * p.tiles_per_split;
  const int t_end = min(t_begin + p.tiles_per_split, p.tiles_total);
  const int ntiles = max(t_end - t_begin, 0);

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmP);
    tc::prefetch_tmap(&tmQ);
    for (int i = 0; i < A_SLOTS; ++i) tc::mbar_init(&a_full[i], 1), tc::mbar_init(&a_empty[i], 1);
    for (int i = 0; i < p.b_slots; ++i) tc::mbar_init(&b_full[i], 1), tc::mbar_init(&b_empty[i], 1);
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
    if (lane == 0) {
      int bit = 0;
      for (int ti = 0; ti < ntiles; ++ti) {
        int t = t_begin + ti;
        const int ity = t % p.nty;
        t /= p.nty;
        const int itx = t % p.ntx;
        t /= p.ntx;
        const int itz = t % p.ntz;
        const int in_ = t / p.ntz;
        const int z0 = itz * p.tz, x0 = itx * p.tx, y0 = ity * p.ty;
        const int as = ti % A_SLOTS;
        tc::mbar_wait(&a_empty[as], ((uint32_t)(ti / A_SLOTS) & 1u) ^ 1u);
        tc::mbar_arrive_expect_tx(&a_full[as], (uint32_t)A_BYTES);
        for (int c = 0; c < A_CHUNKS; ++c)
          tc::tma_load_5d(smA + as * A_BYTES + c * CHUNK_BYTES, &tmP, &a_full[as], r0 + c * 32, y0, x0, z0, in_);
        for (int tl = 0; tl < ntap; ++tl, ++bit) {
          const int tap = tap0 + tl;
          const int k3 = tap % p.ky, j3 = (tap / p.ky) % p.kx, i3 = tap / (p.ky * p.kx);
          const int bs = bit % p.b_slots;
          tc::mbar_wait(&b_empty[bs], ((uint32_t)(bit / p.b_slots) & 1u) ^ 1u);
          tc::mbar_arrive_expect_tx(&b_full[bs], (uint32_t)b_bytes);
          for (int c = 0; c < p.nch; ++c)
            tc::tma_load_5d(smB + bs * b_bytes + c * CHUNK_BYTES, &tmQ, &b_full[bs], s0 + c * 32, y0 + k3 + p.oy,
                            x0 + j3 + p.ox, z0 + i3 + p.oz, in_);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int bit = 0;
      const uint32_t ncols = (uint32_t)p.nch * 32u;
      for (int ti = 0; ti < ntiles; ++ti) {
        const int as = ti % A_SLOTS;
        tc::mbar_wait(&a_full[as], (uint32_t)(ti / A_SLOTS) & 1u);
        const uint32_t a_addr = tc::smem_u32(smA + as * A_BYTES);
        for (int tl = 0; tl < ntap; ++tl, ++bit) {
          const int bs = bit % p.b_slots;
          tc::mbar_wait(&b_full[bs], (uint32_t)(bit / p.b_slots) & 1u);
          tc::tc_fence_after();
          const uint32_t b_addr = tc::smem_u32(smB + bs * b_bytes);
#pragma unroll
          for (int k = 0; k < KP / 8; ++k) {
            // MN-major tf32 operands exist only in the "128B swizzle, 32B atom" layout (UMMA layout
            // type 1 == TMA SWIZZLE_128B_ATOM_32B): rows (positions) 128 B apart, swizzle period
            // 4 rows = 512 B (SBO), 32-channel chunks CHUNK_BYTES apart (LBO); K = 8 rows per MMA
            const uint64_t ad = tc::make_smem_desc(a_addr + k * 1024, CHUNK_BYTES, 512, 1);
            const uint64_t bd = tc::make_smem_desc(b_addr + k * 1024, CHUNK_BYTES, 512, 1);
            tc::mma_tf32_ss(tmem_base + (uint32_t)tl * ncols, ad, bd, p.idesc, (ti > 0 || k > 0) ? 1u : 0u);
          }
          tc::mma_commit(&b_empty[bs]);
        }
        tc::mma_commit(&a_empty[as]);
      }
      tc::mma_commit(acc_full);
    }
  } else {
    const int q = warp & 3;
    const int r = r0 + q * 32 + lane;
    tc::mbar_wait(acc_full, 0);
    tc::tc_fence_after();
    if (ntiles > 0) {
      const int ncols = p.nch * 32;
      for (int tl = 0; tl < ntap; ++tl) {
        const int tap = tap0 + tl;
        const int k3 = tap % p.ky, j3 = (tap / p.ky) % p.kx, i3 = tap / (p.ky * p.kx);
        const int tflip = ((p.kz - 1 - i3) * p.kx + (p.kx - 1 - j3)) * p.ky + (p.ky - 1 - k3);
        for (int c0 = 0; c0 < ncols; c0 += 32) {
          uint32_t v[32];
          tc::tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(tl * ncols + c0), v);
          tc::tmem_ld_wait();
          if (r >= p.R) continue;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int s = s0 + c0 + j;
            if (s >= p.S) continue;
            const int64_t ofs = (p.out_mode == 0) ? ((int64_t)r * p.S + s) * T + tflip : ((int64_t)s * p.R + r) * T + tap;
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

bool pick_tile64(int Oz, int Ox, int Oy, int* tz, int* tx, int* ty) {
  static const int opts[][3] = {{1, 8, 8},  {2, 4, 8}, {1, 4, 16}, {4, 4, 4},  {1, 2, 32}, {2, 2, 16}, {1, 16, 4},
                                {1, 1, 64}, {2, 8, 4}, {4, 2, 8},  {2, 1, 32}, {4, 1, 16}, {8, 2, 4},  {8, 1, 8},
                                {1, 32, 2}, {2, 16, 2}, {4, 8, 2}, {8, 4, 2},  {16, 2, 2}, {1, 64, 1}, {2, 32, 1},
                                {4, 16, 1}, {8, 8, 1}, {16, 4, 1}, {16, 1, 4}, {32, 1, 2}, {32, 2, 1}, {64, 1, 1}};
  int64_t best = -1;
  for (auto& o : opts) {
    int64_t v = (int64_t)((Oz + o[0] - 1) / o[0]) * o[0] * ((Ox + o[1] - 1) / o[1]) * o[1] * ((Oy + o[2] - 1) / o[2]) * o[2];
    if (best < 0 || v < best) best = v, *tz = o[0], *tx = o[1], *ty = o[2];
  }
  return best > 0;
}

}  // namespace

bool e2_reduce_gemm_tc_ok(const e2_handle* h, const ReduceGemm& g) {
  if (!e2_get_tmap_encode()) return false;
  if (g.sz != 1 || g.sx != 1 || g.sy != 1) return false;  // strided (upconv) wgrad stays on CUDA cores
  if (g.R < 8 || g.S < 8) return false;
  if (g.p_pitch % 4 || g.q_pitch % 4) return false;
  if ((reinterpret_cast<uintptr_t>(g.P) & 15) || (reinterpret_cast<uintptr_t>(g.Q) & 15)) return false;
  return true;
}

int e2_launch_reduce_gemm_tc(e2_handle* h, const ReduceGemm& g, cudaStream_t s) {
  EncodeTiledFn enc = e2_get_tmap_encode();
  if (!enc) return e2_fail(h, E2_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled entry point not available");
  WgParams p;
  memset(&p, 0, sizeof(p));
  p.Mn = g.Mn, p.Mz = g.Mz, p.Mx = g.Mx, p.My = g.My;
  pick_tile64(g.Mz, g.Mx, g.My, &p.tz, &p.tx, &p.ty);
  p.ntz = (g.Mz + p.tz - 1) / p.tz, p.ntx = (g.Mx + p.tx - 1) / p.tx, p.nty = (g.My + p.ty - 1) / p.ty;
  p.kz = g.tz, p.kx = g.tx, p.ky = g.ty, p.oz = g.oz, p.ox = g.ox, p.oy = g.oy;
  p.R = g.R, p.S = g.S;
  const int T = g.tz * g.tx * g.ty;
  int nch = (g.S + 31) / 32;
  if (nch > 4) nch = 4;                      // N <= 128 per tap keeps the B ring at 3-4 slots
  p.nch = nch;
  p.n_sc = (g.S + nch * 32 - 1) / (nch * 32);
  p.n_rc = (g.R + 127) / 128;
  int tg = 512 / (nch * 32);
  if (tg > T) tg = T;
  p.tg = tg;
  p.n_tg = (T + tg - 1) / tg;
  int cols = 32;
  while (cols < tg * nch * 32) cols *= 2;
  p.tmem_cols = cols;
  const int b_bytes = nch * CHUNK_BYTES;
  int b_slots = (int)((200 * 1024 - A_SLOTS * A_BYTES) / b_bytes);
  if (b_slots > 6) b_slots = 6;
  if (b_slots < 2) return e2_fail(h, E2_ERR_UNSUPPORTED, "wgrad_tc: shared memory budget");
  p.b_slots = b_slots;
  p.tiles_total = g.Mn * p.ntz * p.ntx * p.nty;
  const int units = p.n_rc * p.n_sc * p.n_tg;
  int splits = (2 * h->sm_count + units - 1) / units;
  if (splits > p.tiles_total) splits = p.tiles_total;
  if (splits < 1) splits = 1;
  p.tiles_per_split = (p.tiles_total + splits - 1) / splits;
  splits = (p.tiles_total + p.tiles_per_split - 1) / p.tiles_per_split;
  p.W = g.W, p.out_mode = g.out_mode;
  p.idesc = tc::make_idesc(2 /*TF32*/, 1, 1, 128, (uint32_t)(nch * 32));

  CUtensorMap tmP, tmQ;
  {
    cuuint64_t dims[5] = {(cuuint64_t)g.R, (cuuint64_t)g.My, (cuuint64_t)g.Mx, (cuuint64_t)g.Mz, (cuuint64_t)g.Mn};
    cuuint64_t pitch = (cuuint64_t)g.p_pitch * 4;
    cuuint64_t strides[4] = {pitch, pitch * g.My, pitch * g.My * g.Mx, pitch * g.My * g.Mx * g.Mz};
    cuuint32_t box[5] = {32, (cuuint32_t)p.ty, (cuuint32_t)p.tx, (cuuint32_t)p.tz, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&tmP, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(g.P), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return e2_fail(h, E2_ERR_CUDA, "cuTensorMapEncodeTiled(P) failed: %d", (int)r);
  }
  {
    cuuint64_t dims[5] = {(cuuint64_t)g.S, (cuuint64_t)g.Qy, (cuuint64_t)g.Qx, (cuuint64_t)g.Qz, (cuuint64_t)g.Qn};
    cuuint64_t pitch = (cuuint64_t)g.q_pitch * 4;
    cuuint64_t strides[4] = {pitch, pitch * g.Qy, pitch * g.Qy * g.Qx, pitch * g.Qy * g.Qx * g.Qz};
    cuuint32_t box[5] = {32, (cuuint32_t)p.ty, (cuuint32_t)p.tx, (cuuint32_t)p.tz, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&tmQ, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(g.Q), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return e2_fail(h, E2_ERR_CUDA, "cuTensorMapEncodeTiled(Q) failed: %d", (int)r);
  }
  cudaMemsetAsync(g.W, 0, sizeof(float) * (size_t)g.R * g.S * T, s);
  const size_t smem = 1024 + (size_t)A_SLOTS * A_BYTES + (size_t)b_slots * b_bytes + (2 * A_SLOTS + 2 * b_slots + 1) * 8 + 16;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(k_wgrad_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)) != cudaSuccess)
      return e2_fail(h, E2_ERR_CUDA, "cudaFuncSetAttribute(max dynamic smem) failed");
    configured = true;
  }
  dim3 grid((unsigned)units, (unsigned)splits);
  k_wgrad_tc<<<grid, WG_THREADS, smem, s>>>(tmP, tmQ, p);
  h->launches++;
  E2_CUDA_CHECK(h, "wgrad_tc");
  return E2_OK;
}
