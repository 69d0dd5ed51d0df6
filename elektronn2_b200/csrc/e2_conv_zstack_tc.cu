// tcgen05 implicit-GEMM convolution, halo planes in shared memory, z-taps stacked along the MMA N axis.
//
// Why this shape (measured on B200 with scripts/mma_bench.cu): a kind::tf32 M128 x N x K8 MMA costs
// max(48, N/2) cycles and at least ~55-85 cycles of issue work when every MMA needs fresh descriptors,
// so MMAs with N = F_out = 64 cap the tensor pipe at <= 67 % (in practice ~35 %).  N >= 128 per MMA is
// needed.  F_out of the big U-Net layers is 64, therefore N is widened with the filter's z-taps:
//
//   tile     = TZ output z-planes x (16 x-lines x 8 y) positions x BN output channels.  One TMEM
//              accumulator block of [128 lanes x BN columns] per output plane, blocks in DESCENDING
//              plane order (plane zl lives at column (TZ-1-zl)*BN).
//   plane    = one TMA box of the input (32 channels, YP = 8+ky-1 y, XH = 16+kx-1 x, 1 z), rows of 128 B,
//              128B-swizzled.  A tile needs NP = TZ+kz-1 input planes per 32-channel block.
//   A view   = for an in-plane tap (j,k): 16 groups of 8 consecutive rows starting at row j*YP+k, groups
//              YP rows apart -> a K-major UMMA descriptor with SBO = YP*128 B (the 128B swizzle is a
//              function of the smem address bits, so any 128-B-row start works).
//   B        = weights of ALL kz z-taps of in-plane tap (j,k), stacked [i][n] -> [kz*BN rows x 32 ch].
//   MMA      = input plane q times the stacked weights gives, in one instruction, the contributions of
//              plane q to output planes q, q-1, ..., q-kz+1 (z-tap i = q - zl): D = accumulator blocks of
//              those planes (contiguous because of the descending order), N = (#planes)*BN, B rows start
//              at block i_lo.  For kz = 3, BN = 64: N = 192 -> 96 cycles per MMA = full tensor rate, and
//              6 MMA groups per (j,k) instead of 12.
//   schedule = channel block -> in-plane tap (j,k) -> input plane q.  All NP planes of a unit stay
//              resident; the weight block of (cb, j, k) is streamed once per unit (2-stage ring).
//   roles    = warp 0 plane TMA, warp 1 MMA issuer (whole warp runs the loop with uniform values so the
//              descriptor arithmetic stays in uniform registers; one elected lane issues), warp 2 weight
//              TMA (+TMEM alloc), warps 4-7 epilogue.  Persistent CTAs, double-buffered accumulators.
#include <algorithm>
#include <stdlib.h>
#include "e2_common.cuh"
#include "e2_conv_internal.cuh"
#include "e2_tc_ptx.cuh"

namespace {

constexpr int TX = 16, TY = 8;   // tile x/y extent (16 groups of 8 rows = 128 MMA rows)
constexpr int EPI_WARPS = 8;      // two epilogue warps per TMEM lane quadrant, alternating over the chunks of a tile
constexpr int ZS_THREADS = (4 + EPI_WARPS) * 32;
constexpr int MAX_PSLOTS = 24;   // NP <= 11 planes per unit, single or double buffered
constexpr int MAX_WSLOTS = 8;

struct ZsParams {
  int On, Oz, Ox, Oy;
  int ntz, ntx, nty, ntn;     // tiles per axis, N tiles
  int TZ, NP, XH, YP;
  int kz, kx, ky, oz, ox, oy;
  int K, N, BN, CB;
  int nslot, wslot, acc_bufs;
  int plane_bytes, plane_stride, w_bytes, wblk_bytes;
  int tmem_cols;
  int num_tiles;
  int ksplit, cb_per;            // K split over channel blocks: unit = (tile, split); raw partial tiles go to a workspace
  float* C;
  int c_pitch;
  const float* bias;
  const float* gate;
  int act, accumulate, round_tf32;
  uint32_t idesc0, idesc_step;   // idesc for N = nblk*BN is idesc0 + nblk*idesc_step
  int epi_off;                   // byte offset of the epilogue staging area (1024-aligned)
  int bias_copies;               // EPI_WARPS private copies of the bias tile, or 1 shared copy when every tile has the same n0
  long long* trace;              // E2_ZS_TRACE: per-role cycle counters of CTA 0 (debug)
  int dbg;                       // E2_ZS_DBG bottleneck experiments: 1 no plane TMA, 2 no weight TMA, 4 no MMA, 8 no stores
  // max-pool fused into the epilogue (e2_conv3d_fwd_pool): window (ppz,ppx,ppy) in {1,2}^3 over the values the
  // epilogue above produces, then +pbias -> pact -> tf32 round; argmax = first maximum in (z,x,y) scan order
  int pool, ppz, ppx, ppy;
  int store_full;                // also store the unpooled tile through tmC (0: the caller does not want that tensor)
  int kz0, kz1, kx0, kx1, ky0, ky1;   // ... but only the [32 ch x 8 y x 4 x] boxes that intersect this window (half-open)
  int pool_amax;                 // store the argmax tile through tmI
  const float* pbias;
  int pact, pround;
};

// Fused max-pool, argmax part.  Every lane of a 2x2 (x,y) window group knows the window maximum m[j]; `e` = bit j set
// if this lane's column (its own position, best of the two planes) attains it, `mz` = bit j set if the column's best
// came from the second plane.  Seen from the lane at the window's origin the other columns sit at lane ^ 1 (dy = 1),
// lane ^ 8 (dx = 1) and lane ^ 9, so six shuffles bring all masks together and the FIRST maximum in (z,x,y) scan order
// -- candidates ordered (dz,dx,dy) -- is picked for all 32 channels at once with bitwise logic.
__device__ __forceinline__ void zs_pool_first(uint32_t e, uint32_t mz, bool has_x, bool has_y, uint32_t& rz, uint32_t& rx,
                                              uint32_t& ry) {
  const uint32_t e01 = __shfl_xor_sync(0xffffffffu, e, 1), z01 = __shfl_xor_sync(0xffffffffu, mz, 1);
  const uint32_t e10 = __shfl_xor_sync(0xffffffffu, e, 8), z10 = __shfl_xor_sync(0xffffffffu, mz, 8);
  const uint32_t e11 = __shfl_xor_sync(0xffffffffu, e, 9), z11 = __shfl_xor_sync(0xffffffffu, mz, 9);
  const uint32_t E00 = e, E01 = has_y ? e01 : 0u, E10 = has_x ? e10 : 0u, E11 = (has_x && has_y) ? e11 : 0u;
  uint32_t found, c;
  rz = rx = ry = 0u;
  found = E00 & ~mz;                                                  // (0,0,0)
  c = E01 & ~z01 & ~found, ry |= c, found |= c;                       // (0,0,1)
  c = E10 & ~z10 & ~found, rx |= c, found |= c;                       // (0,1,0)
  c = E11 & ~z11 & ~found, rx |= c, ry |= c, found |= c;              // (0,1,1)
  c = E00 & mz & ~found, rz |= c, found |= c;                         // (1,0,0)
  c = E01 & z01 & ~found, rz |= c, ry |= c, found |= c;               // (1,0,1)
  c = E10 & z10 & ~found, rz |= c, rx |= c, found |= c;               // (1,1,0)
  c = E11 & z11 & ~found, rz |= c, rx |= c, ry |= c;                  // (1,1,1)
}

// lean bounded wait for the issuing warp (all lanes poll; try_wait suspends in hardware)
__device__ __forceinline__ void wait_bar(uint64_t* bar, uint32_t parity) {
  uint32_t n = 0;
  while (!tc::mbar_try_wait(bar, parity)) {
    if (++n > (1u << 24)) {
      printf("e2b200: zstack mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}

// one 128-byte K block = up to 4 MMAs of K = 8; nk < 4 when the channel count ends inside the block
__device__ __forceinline__ void mma4(uint32_t acc, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t first_acc, int nk) {
  tc::mma_tf32_ss(acc, ad, bd, idesc, first_acc);
  if (nk > 1) tc::mma_tf32_ss(acc, ad + 2, bd + 2, idesc, 1u);
  if (nk > 2) tc::mma_tf32_ss(acc, ad + 4, bd + 4, idesc, 1u);
  if (nk > 3) tc::mma_tf32_ss(acc, ad + 6, bd + 6, idesc, 1u);
}

// One stage = one (channel block, in-plane tap): every input plane of the unit times the stacked weights.
// TZ / KZ are compile-time so the plane loop unrolls into straight-line code whose per-plane constants
// (accumulator column, first weight block, MMA width) fold away; the elected lane issues ~6 uniform
// instructions per MMA instead of ~19 (the run-time loop was issue-bound: 122 cycles per MMA).
template <int TZ, int KZ, bool WAIT>
__device__ __forceinline__ void zs_issue_stage(uint64_t a_desc0, uint32_t pstride_enc, uint64_t bd0, uint32_t wblk_enc,
                                               uint32_t acc0, uint32_t bn, uint32_t idesc0, uint32_t idesc_step,
                                               bool first_stage, bool last_stage, uint64_t* pl_full, uint64_t* pl_empty,
                                               uint32_t par_full, int nk) {
  constexpr int NP = TZ + KZ - 1;
  uint64_t ad = a_desc0;
#pragma unroll
  for (int q = 0; q < NP; ++q) {
    const int zl_hi = q < TZ - 1 ? q : TZ - 1;
    const int zl_lo = q - (KZ - 1) > 0 ? q - (KZ - 1) : 0;
    const int i_lo = q - zl_hi, nblk = zl_hi - zl_lo + 1;
    if (WAIT) {
      wait_bar(&pl_full[q], par_full);
      tc::tc_fence_after();
    }
    const uint64_t bd = bd0 + (uint64_t)((uint32_t)i_lo * wblk_enc);
    const uint32_t dcol = acc0 + (uint32_t)(TZ - 1 - zl_hi) * bn;
    if (q < TZ && first_stage) {
      // output plane q is touched for the first time (z-tap 0 = block i_lo = 0): overwrite
      mma4(dcol, ad, bd, idesc0 + idesc_step, 0u, nk);
      if (nblk > 1) mma4(dcol + bn, ad, bd + wblk_enc, idesc0 + (uint32_t)(nblk - 1) * idesc_step, 1u, nk);
    } else {
      mma4(dcol, ad, bd, idesc0 + (uint32_t)nblk * idesc_step, 1u, nk);
    }
    if (last_stage) tc::mma_commit(&pl_empty[q]);   // plane slot free once these MMAs have read it
    ad += pstride_enc;
  }
}

// POOL_: the fused max-pool epilogue is a separate instantiation -- with it compiled into every kernel the epilogue of
// the plain convolutions doubled in size and the small, epilogue-bound layers (unet3d_litelite, neuro3d_lite) lost 4-9 %
template <int TZ_, int KZ_, int POOL_>
__global__ void __launch_bounds__(ZS_THREADS, 1) k_conv_zstack_tc(const __grid_constant__ CUtensorMap tmA,
                                                                  const __grid_constant__ CUtensorMap tmB,
                                                                  const __grid_constant__ CUtensorMap tmC,
                                                                  const __grid_constant__ CUtensorMap tmG,
                                                                  const __grid_constant__ CUtensorMap tmP,
                                                                  const __grid_constant__ CUtensorMap tmI,
                                                                  const ZsParams p) {
  extern __shared__ uint8_t smem_raw[];
  // aligned base by OFFSET arithmetic on smem_raw (no integer round trip), so the pointer keeps its address space and the
  // warps' own accesses compile to LDS / STS instead of generic LD / ST
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smP = smem;                                   // plane slots
  uint8_t* smW = smem + p.nslot * p.plane_stride;        // weight ring
  uint8_t* smE = smem + p.epi_off;                       // epilogue: per warp one 4 KB staging/aux tile, then bias
  uint64_t* bars = reinterpret_cast<uint64_t*>(smE + EPI_WARPS * 4096 + p.bias_copies * p.BN * 4);
  uint64_t* pl_full = bars;
  uint64_t* pl_empty = pl_full + MAX_PSLOTS;
  uint64_t* w_full = pl_empty + MAX_PSLOTS;
  uint64_t* w_empty = w_full + MAX_WSLOTS;
  uint64_t* acc_full = w_empty + MAX_WSLOTS;             // [2]
  uint64_t* acc_empty = acc_full + 2;                    // [2]
  uint64_t* aux_bar = acc_empty + 2;                     // [EPI_WARPS] one per epilogue warp
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux_bar + EPI_WARPS);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform
  const int lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmA);
    tc::prefetch_tmap(&tmB);
    tc::prefetch_tmap(&tmC);
    if (p.gate) tc::prefetch_tmap(&tmG);
    if (p.pool) tc::prefetch_tmap(&tmP);
    if (p.pool && p.pool_amax) tc::prefetch_tmap(&tmI);
    for (int i = 0; i < p.nslot; ++i) tc::mbar_init(&pl_full[i], 1), tc::mbar_init(&pl_empty[i], 1);
    for (int i = 0; i < p.wslot; ++i) tc::mbar_init(&w_full[i], 1), tc::mbar_init(&w_empty[i], 1);
    for (int i = 0; i < 2; ++i) tc::mbar_init(&acc_full[i], 1), tc::mbar_init(&acc_empty[i], EPI_WARPS);
    for (int i = 0; i < EPI_WARPS; ++i) tc::mbar_init(&aux_bar[i], 1);
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

  // unit index = (((n, tile z, x, y), N tile), K split) with the split fastest: the CTAs working on one tile's
  // channel-block ranges run at the same time and write partial tile `ks` of the workspace
  auto tile_coords = [&](int t, int& in_, int& z0, int& x0, int& y0, int& n0, int& cb_lo, int& cb_hi) {
    const int ks = t % p.ksplit;
    t /= p.ksplit;
    cb_lo = ks * p.cb_per, cb_hi = min(p.CB, cb_lo + p.cb_per);
    const int nt = t % p.ntn;
    t /= p.ntn;
    const int ity = t % p.nty;
    t /= p.nty;
    const int itx = t % p.ntx;
    t /= p.ntx;
    const int itz = t % p.ntz;
    in_ = t / p.ntz + ks * p.On;     // + split * On: batch coordinate of the partial tensor (only the store uses it)
    z0 = itz * p.TZ, x0 = itx * TX, y0 = ity * TY, n0 = nt * p.BN;
  };

  if (warp == 0) {
    // ----------------------------------------------------------- plane producer
    // A unit (tile, channel block) owns NP consecutive slots: group 0, or groups 0/1 alternately when the
    // ring is double buffered (nslot == 2*NP).
    if (lane == 0) {
      const int dbl = (p.nslot >= 2 * p.NP) ? 1 : 0;
      int ucount = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
        int in_, z0, x0, y0, n0, cb_lo, cb_hi;
        tile_coords(t, in_, z0, x0, y0, n0, cb_lo, cb_hi);
        in_ %= p.On;
        for (int cb = cb_lo; cb < cb_hi; ++cb, ++ucount) {
          const int s0 = dbl ? (ucount & 1) * p.NP : 0;
          const uint32_t par = (((uint32_t)(dbl ? (ucount >> 1) : ucount)) & 1u) ^ 1u;
          for (int pl = 0; pl < p.NP; ++pl) {
            const int s = s0 + pl;
            tc::mbar_wait(&pl_empty[s], par);
            if (p.dbg & 1) {
              tc::mbar_arrive(&pl_full[s]);
            } else {
              tc::mbar_arrive_expect_tx(&pl_full[s], (uint32_t)p.plane_bytes);
              tc::tma_load_5d(smP + s * p.plane_stride, &tmA, &pl_full[s], cb * 32, y0 + p.oy, x0 + p.ox,
                              z0 + p.oz + pl, in_);
            }
          }
        }
      }
    }
  } else if (warp == 2) {
    // ---------------------------------------------------------- weight producer
    if (lane == 0) {
      int s = 0;
      uint32_t par = 1;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
        int in_, z0, x0, y0, n0, cb_lo, cb_hi;
        tile_coords(t, in_, z0, x0, y0, n0, cb_lo, cb_hi);
        for (int cb = cb_lo; cb < cb_hi; ++cb)
          for (int jk = 0; jk < T9; ++jk) {
            tc::mbar_wait(&w_empty[s], par);
            if (p.dbg & 2) {
              tc::mbar_arrive(&w_full[s]);
            } else {
              tc::mbar_arrive_expect_tx(&w_full[s], (uint32_t)p.w_bytes);
              for (int i = 0; i < p.kz; ++i)
                tc::tma_load_3d(smW + s * p.w_bytes + i * p.wblk_bytes, &tmB, &w_full[s], cb * 32, i * T9 + jk, n0);
            }
            if (++s == p.wslot) s = 0, par ^= 1u;
          }
      }
    }
  } else if (warp == 1) {
    // --------------------------------------------------------------- MMA issuer
    // The warp walks tiles / channel blocks / in-plane taps with warp-uniform values; per stage ONE elected
    // lane issues the unrolled plane sequence (zs_issue_stage).
    const uint32_t smP_enc = tc::smem_u32(smP) >> 4, smW_enc = tc::smem_u32(smW) >> 4;
    const uint32_t pstride_enc = (uint32_t)p.plane_stride >> 4, w_enc = (uint32_t)p.w_bytes >> 4;
    const uint32_t wblk_enc = (uint32_t)p.wblk_bytes >> 4;
    const uint64_t a_tmpl = tc::make_smem_desc(0, 16, (uint32_t)(p.YP * 128), 2);
    const uint64_t b_tmpl = tc::make_smem_desc(0, 16, 1024, 2);
    const uint32_t bn = (uint32_t)p.BN, idesc0 = p.idesc0, idesc_step = p.idesc_step;
    const int dbl = (p.nslot >= 2 * p.NP) ? 1 : 0;
    int ucount = 0;
    int ws = 0;
    uint32_t wpar = 0;
    int buf = 0;
    uint32_t bpar = 1;
    long long tr_acc = 0, tr_w = 0, tr_issue = 0, tr_t0 = clock64();
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
      long long c0 = clock64();
      wait_bar(&acc_empty[buf], bpar);
      tr_acc += clock64() - c0;
      tc::tc_fence_after();
      const uint32_t acc0 = tmem_base + (uint32_t)(buf * TZ_) * bn;
      const int ks = t % p.ksplit;
      const int cb_lo = ks * p.cb_per, cb_hi = min(p.CB, cb_lo + p.cb_per);
      for (int cb = cb_lo; cb < cb_hi; ++cb, ++ucount) {
        const int s0 = dbl ? (ucount & 1) * p.NP : 0;
        const uint32_t par_full = ((uint32_t)(dbl ? (ucount >> 1) : ucount)) & 1u;
        const uint32_t unit_enc = smP_enc + (uint32_t)s0 * pstride_enc;
        const int nk = min(4, (p.K - cb * 32 + 7) >> 3);
        int j = 0, k = 0;
        for (int jk = 0; jk < T9; ++jk) {
          long long c1 = clock64();
          wait_bar(&w_full[ws], wpar);
          long long c2 = clock64();
          tr_w += c2 - c1;
          tc::tc_fence_after();
          if (tc::elect_one()) {
            const uint64_t bd0 = b_tmpl + (uint64_t)(smW_enc + (uint32_t)ws * w_enc);
            const uint64_t ad0 = a_tmpl + (uint64_t)(unit_enc + (uint32_t)((j * p.YP + k) * 8));   // (rows * 128 B) >> 4
            const bool last_stage = (jk == T9 - 1);
            if (p.dbg & 4) {
              if (jk == 0)
                for (int q = 0; q < p.NP; ++q) wait_bar(&pl_full[s0 + q], par_full);
              if (last_stage)
                for (int q = 0; q < p.NP; ++q) tc::mma_commit(&pl_empty[s0 + q]);
            } else if (jk == 0) {
              zs_issue_stage<TZ_, KZ_, true>(ad0, pstride_enc, bd0, wblk_enc, acc0, bn, idesc0, idesc_step, cb == cb_lo,
                                             last_stage, pl_full + s0, pl_empty + s0, par_full, nk);
            } else {
              zs_issue_stage<TZ_, KZ_, false>(ad0, pstride_enc, bd0, wblk_enc, acc0, bn, idesc0, idesc_step, false,
                                              last_stage, pl_full + s0, pl_empty + s0, par_full, nk);
            }
            tc::mma_commit(&w_empty[ws]);
          }
          __syncwarp();
          tr_issue += clock64() - c2;
          if (++ws == p.wslot) ws = 0, wpar ^= 1u;
          if (++k == p.ky) k = 0, ++j;
        }
      }
      if (tc::elect_one()) tc::mma_commit(&acc_full[buf]);
      __syncwarp();
      if (++buf == p.acc_bufs) buf = 0, bpar ^= 1u;
    }
    if (p.trace && blockIdx.x == 0 && lane == 0) {
      p.trace[0] = clock64() - tr_t0, p.trace[1] = tr_acc, p.trace[2] = tr_w, p.trace[3] = tr_issue;
    }
  } else if (warp >= 4) {
    // ----------------------------------------------------------------- epilogue
    // Warps q and q+4 own TMEM lanes 32q..32q+31 = x-lines 4q..4q+3 of the tile and alternate over its
    // (plane, 32-column chunk) list.  Per chunk: tcgen05.ld -> +bias -> act -> (gate / accumulate from a
    // TMA-loaded tile) -> tf32 round -> swizzled st.shared -> ONE TMA store of the [32 ch x 8 y x 4 x] box.
    // The 4 KB buffer serves as aux tile first and as store staging afterwards.  No per-thread global
    // stores: with ~220 KB of shared memory in use the L1 is gone and scattered 16-byte STGs ran at ~300 GB/s.
    const int ew = warp - 4;
    const int q = warp & 3, half = ew >> 2;
    uint8_t* stage = smE + ew * 4096;
    // one N tile (n0 == 0 for every tile): all warps write the same values, so one copy serves them all (frees 7 x BN x 4
    // bytes -- what the N = 100 layers of neuro3d need to run BN = 112 with two output planes per tile)
    float* bias_w = reinterpret_cast<float*>(smE + EPI_WARPS * 4096) + (p.bias_copies > 1 ? ew * p.BN : 0);
    uint64_t* abar = &aux_bar[ew];
    uint32_t apar = 0;
    const int r8 = lane & 7;
    const bool has_aux = p.gate || p.accumulate;
    long long tr_full = 0, tr_work = 0, tr_e0 = clock64();
    int buf = 0;
    uint32_t fpar = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
      int in_, z0, x0, y0, n0, cb_lo, cb_hi;
      tile_coords(t, in_, z0, x0, y0, n0, cb_lo, cb_hi);
      for (int i = lane; i < p.BN; i += 32) bias_w[i] = (p.bias && n0 + i < p.N) ? __ldg(p.bias + n0 + i) : 0.f;
      __syncwarp();
      long long e0 = clock64();
      tc::mbar_wait(&acc_full[buf], fpar);
      long long e1 = clock64();
      tr_full += e1 - e0;
      tc::tc_fence_after();
      int ci = 0;
      constexpr bool POOL = POOL_ != 0;
      const int PZ = POOL ? p.ppz : 1;                    // planes per unit of epilogue work (the pool window's z extent)
      for (int zp = 0; zp < p.TZ; zp += PZ) {
        if (z0 + zp >= p.Oz) break;                       // uniform: whole plane (pair) outside the tensor
        for (int c0 = 0; c0 < p.BN; c0 += 32) {
          if (n0 + c0 >= p.N) break;                      // uniform: chunk entirely past the last channel
          if ((ci++ & 1) != half) continue;               // the sibling warp's chunk
          float pv[32];                                   // fused pool: best value per channel so far ...
          uint32_t mz = 0;                                // ... and bit j = it came from the window's second plane
#pragma unroll
          for (int dz = 0; dz < 2; ++dz) {
            if (dz < PZ) {
              const int zl = zp + dz;
              const uint32_t acc =
                  tmem_base + (uint32_t)((buf * p.TZ + (p.TZ - 1 - zl)) * p.BN) + ((uint32_t)(q * 32) << 16);
              // the previous TMA store must have finished reading the buffer
              if (lane == 0) {
                tc::bulk_wait_read0();
                if (has_aux) {
                  tc::mbar_arrive_expect_tx(abar, 4096u);
                  tc::tma_load_5d(stage, p.gate ? &tmG : &tmC, abar, n0 + c0, y0, x0 + 4 * q, z0 + zl, in_);
                }
              }
              __syncwarp();
              uint32_t r[32];
              if (p.BN - c0 >= 32) {
                tc::tmem_ld_32x32b_x32(acc + (uint32_t)c0, r);
              } else {
                tc::tmem_ld_32x32b_x16(acc + (uint32_t)c0, r);
#pragma unroll
                for (int jj = 16; jj < 32; ++jj) r[jj] = 0u;
              }
              tc::tmem_ld_wait();
              float v[32];
#pragma unroll
              for (int j4 = 0; j4 < 32; j4 += 4) {
                float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (c0 + j4 < p.BN) b4 = *reinterpret_cast<const float4*>(bias_w + c0 + j4);
                v[j4 + 0] = __uint_as_float(r[j4 + 0]) + b4.x;
                v[j4 + 1] = __uint_as_float(r[j4 + 1]) + b4.y;
                v[j4 + 2] = __uint_as_float(r[j4 + 2]) + b4.z;
                v[j4 + 3] = __uint_as_float(r[j4 + 3]) + b4.w;
              }
              // one (uniform) branch per chunk, not one switch per value
              if (p.act == E2_ACT_RELU) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
              } else if (p.act != E2_ACT_LIN) {
                // fully unrolled on purpose: a partially unrolled loop indexes v[] dynamically, which puts the
                // whole array in local memory (L1 is carved out for shared memory here -> every access goes to L2)
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = e2_apply_act(v[j], p.act);
              }
              if (has_aux) {
                tc::mbar_wait(abar, apar);
                apar ^= 1u;
                if (p.gate) {
#pragma unroll
                  for (int j = 0; j < 8; ++j) {
                    const float4 g4 = *reinterpret_cast<const float4*>(stage + lane * 128 + ((j ^ r8) << 4));
                    if (!(g4.x > 0.f)) v[4 * j + 0] = 0.f;
                    if (!(g4.y > 0.f)) v[4 * j + 1] = 0.f;
                    if (!(g4.z > 0.f)) v[4 * j + 2] = 0.f;
                    if (!(g4.w > 0.f)) v[4 * j + 3] = 0.f;
                  }
                  if (p.accumulate) {      // rare: both -> second aux load of the destination tile
                    tc::fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                      tc::mbar_arrive_expect_tx(abar, 4096u);
                      tc::tma_load_5d(stage, &tmC, abar, n0 + c0, y0, x0 + 4 * q, z0 + zl, in_);
                    }
                    tc::mbar_wait(abar, apar);
                    apar ^= 1u;
                  }
                }
                if (p.accumulate) {
#pragma unroll
                  for (int j = 0; j < 8; ++j) {
                    const float4 c4 = *reinterpret_cast<const float4*>(stage + lane * 128 + ((j ^ r8) << 4));
                    v[4 * j + 0] += c4.x, v[4 * j + 1] += c4.y, v[4 * j + 2] += c4.z, v[4 * j + 3] += c4.w;
                  }
                }
              }
              if (p.round_tf32) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = e2_round_tf32(v[j]);
              }
              // fused pool: the caller may need only a window of the unpooled tensor (the skip connection's crop)
              const bool keep = !POOL || (p.store_full && z0 + zl >= p.kz0 && z0 + zl < p.kz1 && x0 + 4 * q + 4 > p.kx0 &&
                                                             x0 + 4 * q < p.kx1 && y0 + TY > p.ky0 && y0 < p.ky1);
              if (keep) {
                // each lane reads and writes only its own 128-byte row of the buffer: no cross-lane hazard
#pragma unroll
                for (int j = 0; j < 8; ++j)
                  *reinterpret_cast<float4*>(stage + lane * 128 + ((j ^ r8) << 4)) =
                      make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                tc::fence_proxy_async();
                __syncwarp();
                if (lane == 0 && !(p.dbg & 8)) {
                  tc::tma_store_5d(&tmC, stage, n0 + c0, y0, x0 + 4 * q, z0 + zl, in_);
                  tc::bulk_commit();
                }
              }
              if (POOL) {
                if (dz == 0) {
#pragma unroll
                  for (int j = 0; j < 32; ++j) pv[j] = v[j];
                } else {
#pragma unroll
                  for (int j = 0; j < 32; ++j)
                    if (v[j] > pv[j]) pv[j] = v[j], mz |= 1u << j;   // strict '>': the first plane keeps ties
                }
              }
            }
          }
          if (POOL) {
            // ------------------------------------------------------------ fused max-pool of this unit
            // lane = x-line (lane >> 3) x y (lane & 7) of the 4 x 8 position patch; window partners are lane ^ 1
            // (y) and lane ^ 8 (x).  After the merges the lane at a window's origin holds (max, first argmax).
            const int xl = lane >> 3, yy = lane & 7;
            // window maximum: two butterfly steps of plain maxima (relu outputs / accumulators: no -0 vs +0 question
            // in practice -- max.f32 orders -0 below +0, the pool kernel keeps the first of equal values)
            float m[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) m[j] = pv[j];
            if (p.ppy == 2) {
#pragma unroll
              for (int j = 0; j < 32; ++j) m[j] = fmaxf(m[j], __shfl_xor_sync(0xffffffffu, m[j], 1));
            }
            if (p.ppx == 2) {
#pragma unroll
              for (int j = 0; j < 32; ++j) m[j] = fmaxf(m[j], __shfl_xor_sync(0xffffffffu, m[j], 8));
            }
            uint32_t mx = 0, my = 0;
            if (p.pool_amax) {
              uint32_t e = 0;
#pragma unroll
              for (int j = 0; j < 32; ++j) e |= (pv[j] == m[j]) ? (1u << j) : 0u;
              uint32_t rz;
              zs_pool_first(e, mz, p.ppx == 2, p.ppy == 2, rz, mx, my);
              mz = rz;
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) pv[j] = m[j];
            if (p.pbias) {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (n0 + c0 + j < p.N) pv[j] += __ldg(p.pbias + n0 + c0 + j);
            }
            if (p.pact == E2_ACT_RELU) {
#pragma unroll
              for (int j = 0; j < 32; ++j) pv[j] = fmaxf(pv[j], 0.f);
            } else if (p.pact != E2_ACT_LIN) {
#pragma unroll
              for (int j = 0; j < 32; ++j) pv[j] = e2_apply_act(pv[j], p.pact);
            }
            if (p.pround) {
#pragma unroll
              for (int j = 0; j < 32; ++j) pv[j] = e2_round_tf32(pv[j]);
            }
            const bool owner = (((xl & (p.ppx - 1)) | (yy & (p.ppy - 1))) == 0);
            const int rows_y = TY / p.ppy;
            const int row = (xl / p.ppx) * rows_y + yy / p.ppy;          // row of the pooled box [4/ppx][8/ppy][32 ch]
            const int nrows = 32 / (p.ppx * p.ppy);
            uint8_t* st_v = stage;
            uint8_t* st_i = stage + (nrows * 256 <= 4096 ? nrows * 128 : 0);   // 1024-aligned either way (swizzle period)
            const int pzc = (z0 + zp) / p.ppz, pxc = (x0 + 4 * q) / p.ppx, pyc = y0 / p.ppy;
            if (lane == 0) tc::bulk_wait_read0();
            __syncwarp();
            if (owner) {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                *reinterpret_cast<float4*>(st_v + row * 128 + ((j ^ (row & 7)) << 4)) =
                    make_float4(pv[4 * j], pv[4 * j + 1], pv[4 * j + 2], pv[4 * j + 3]);
            }
            tc::fence_proxy_async();
            __syncwarp();
            if (lane == 0 && !(p.dbg & 8)) {
              tc::tma_store_5d(&tmP, st_v, n0 + c0, pyc, pxc, pzc, in_);
              tc::bulk_commit();
            }
            if (p.pool_amax) {
              if (st_i == st_v) {
                if (lane == 0) tc::bulk_wait_read0();
                __syncwarp();
              }
              if (owner) {
                // int32 z*X*Y + x*Y + y in the geometry of the unpooled tensor (include/e2b200.h, e2_pool_desc)
                const int base = ((z0 + zp) * p.Ox + (x0 + 4 * q + xl)) * p.Oy + y0 + yy;
                const int sz = p.Ox * p.Oy, sx = p.Oy;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  int4 a;
                  a.x = base + (int)((mz >> (4 * j + 0)) & 1u) * sz + (int)((mx >> (4 * j + 0)) & 1u) * sx + (int)((my >> (4 * j + 0)) & 1u);
                  a.y = base + (int)((mz >> (4 * j + 1)) & 1u) * sz + (int)((mx >> (4 * j + 1)) & 1u) * sx + (int)((my >> (4 * j + 1)) & 1u);
                  a.z = base + (int)((mz >> (4 * j + 2)) & 1u) * sz + (int)((mx >> (4 * j + 2)) & 1u) * sx + (int)((my >> (4 * j + 2)) & 1u);
                  a.w = base + (int)((mz >> (4 * j + 3)) & 1u) * sz + (int)((mx >> (4 * j + 3)) & 1u) * sx + (int)((my >> (4 * j + 3)) & 1u);
                  *reinterpret_cast<int4*>(st_i + row * 128 + ((j ^ (row & 7)) << 4)) = a;
                }
              }
              tc::fence_proxy_async();
              __syncwarp();
              if (lane == 0 && !(p.dbg & 8)) {
                tc::tma_store_5d(&tmI, st_i, n0 + c0, pyc, pxc, pzc, in_);
                tc::bulk_commit();
              }
            }
          }
        }
      }
      // all TMEM reads of this warp are complete (tcgen05.wait::ld above): hand the buffer back
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&acc_empty[buf]);
      tr_work += clock64() - e1;
      if (++buf == p.acc_bufs) buf = 0, fpar ^= 1u;
    }
    if (lane == 0) tc::bulk_wait0();
    if (p.trace && blockIdx.x == 0 && lane == 0 && (ew == 0 || ew == 4)) {
      long long* tr = p.trace + 8 + (ew >> 2) * 4;
      tr[0] = clock64() - tr_e0, tr[1] = tr_full, tr[2] = tr_work;
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}



// Split-K second pass: sums the raw partial tiles [split][n][z][x][y][Np] and applies the fused epilogue
// (+bias -> act -> ReLU gate -> accumulate -> tf32 round).  One thread per position x 4 channels.
__global__ void __launch_bounds__(256) k_zstack_reduce(const float* __restrict__ part, int ksplit, int64_t positions, int Np,
                                                       int N, float* __restrict__ C, int c_pitch,
                                                       const float* __restrict__ bias, const float* __restrict__ gate, int act,
                                                       int accumulate, int round_tf32) {
  const uint32_t q4 = (uint32_t)Np / 4;
  const uint32_t total = (uint32_t)positions * q4;            // < 2^31 (checked by the launcher)
  const int64_t zstride = positions * Np;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const uint32_t pos = i / q4;                              // 32-bit division: cheap next to the ksplit loads
    const int c = (int)(i - pos * q4) * 4;
    const float* src = part + (int64_t)pos * Np + c;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int z = 0; z < ksplit; ++z) {
      const float4 v = __ldcs(reinterpret_cast<const float4*>(src + (int64_t)z * zstride));
      a.x += v.x, a.y += v.y, a.z += v.z, a.w += v.w;
    }
    float v[4] = {a.x, a.y, a.z, a.w};
    const int64_t ofs = (int64_t)pos * c_pitch + c;
    const bool full = c + 3 < N;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (c + e >= N) continue;
      if (bias) v[e] += __ldg(bias + c + e);
      v[e] = e2_apply_act(v[e], act);
    }
    if (full) {
      if (gate) {
        const float4 g4 = __ldg(reinterpret_cast<const float4*>(gate + ofs));
        v[0] = g4.x > 0.f ? v[0] : 0.f, v[1] = g4.y > 0.f ? v[1] : 0.f;
        v[2] = g4.z > 0.f ? v[2] : 0.f, v[3] = g4.w > 0.f ? v[3] : 0.f;
      }
      if (accumulate) {
        const float4 c4 = *reinterpret_cast<const float4*>(C + ofs);
        v[0] += c4.x, v[1] += c4.y, v[2] += c4.z, v[3] += c4.w;
      }
      if (round_tf32) v[0] = e2_round_tf32(v[0]), v[1] = e2_round_tf32(v[1]), v[2] = e2_round_tf32(v[2]), v[3] = e2_round_tf32(v[3]);
      *reinterpret_cast<float4*>(C + ofs) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (c + e >= N) continue;
        float x = v[e];
        if (gate && !(__ldg(gate + ofs + e) > 0.f)) x = 0.f;
        if (accumulate) x += C[ofs + e];
        if (round_tf32) x = e2_round_tf32(x);
        C[ofs + e] = x;
      }
    }
  }
}

}  // namespace

// Decide whether the kernel applies and is worthwhile for this problem; fill the geometry.
// The N tile, the number of output planes per tile and the K split are chosen together by a cycle model:
//   * an MMA of width n costs max(48, 32 + n/4, n/2) cycles (measured, scripts/mma_bench.cu) and a stage
//     ((channel block, in-plane tap): NP plane MMA groups) carries ~250 cycles of wait / descriptor / commit work,
//     so layers whose tiles have a single output plane want wide N tiles (up to 256 / min(TZ, kz));
//   * layers with few output positions have fewer tiles than SMs and a long serial K loop: their channel blocks
//     are split over CTAs (raw partial tiles in the caller's workspace + k_zstack_reduce).
static bool plan_zstack(const e2_handle* h, const GatherGemm& g, ZsParams* p, bool may_split) {
  if (g.sz != 1 || g.sx != 1 || g.sy != 1 || g.shuffle) return false;
  if (g.ty > 9 || g.tx > 9 || g.tz > 4) return false;
  const int T = g.tz * g.tx * g.ty;
  if (T < 2) return false;                       // 1x1x1: nothing to reuse, the tap kernel is fine
  if (g.Ox < 12 || g.Oy < 6) return false;       // tile quantisation would waste too much
  memset(p, 0, sizeof(*p));
  const int S = g.tz;
  p->XH = TX + g.tx - 1;
  p->YP = TY + g.ty - 1;
  p->plane_bytes = p->XH * p->YP * 128;
  p->plane_stride = (p->plane_bytes + 1023) / 1024 * 1024;   // slots stay 1024-B aligned (swizzle period)
  if (g.c_pitch % 4 || (reinterpret_cast<uintptr_t>(g.C) & 15) || (reinterpret_cast<uintptr_t>(g.gate) & 15)) return false;
  if (g.fuse_pool) {
    // pooled tiles must be whole windows of whole tiles: windows of 1 or 2 per axis (tile = TZ x 16 x 8, TZ even below),
    // extents divisible by the window (Pool._calc_shape, neural.py:1543-1546), no K split (the epilogue lives here)
    if (g.qz < 1 || g.qz > 2 || g.qx < 1 || g.qx > 2 || g.qy < 1 || g.qy > 2) return false;
    if (g.Oz % g.qz || g.Ox % g.qx || g.Oy % g.qy) return false;
    if (g.gate || g.accumulate || !g.Cp || g.cp_pitch % 4) return false;
    if ((reinterpret_cast<uintptr_t>(g.Cp) & 15) || (reinterpret_cast<uintptr_t>(g.Ci) & 15)) return false;
    may_split = false;
  }
  const int ntx = (g.Ox + TX - 1) / TX, nty = (g.Oy + TY - 1) / TY;
  const double kLatency = 1200.0, kStageOvh = 250.0, kEpiChunk = 1200.0, kReduceFixed = 8000.0, kReduceBpc = 2500.0;
  const int T9 = g.tx * g.ty, CBn = (g.K + 31) / 32;
  const double nk_avg = (double)((g.K + 7) / 8) / CBn;   // K = 8 slices per 32-channel block
  const double c_bytes = (double)g.On * g.Oz * g.Ox * g.Oy * ((g.N + 3) / 4 * 4) * 4.0;
  static const int ks_opts[] = {1, 2, 3, 4, 5, 6, 8, 10, 12, 16, 24};
  if (c_bytes / 16.0 >= 2147483648.0) may_split = false;     // k_zstack_reduce indexes float4s with 32 bits
  static const bool no_split = getenv("E2_ZS_NOSPLIT") != nullptr;
  static const bool narrow = getenv("E2_ZS_NARROW") != nullptr;     // A/B switch: N tiles as before (kz * BN <= 256)
  int best_tz = 0, best_w = 0, best_dbl = 0, best_bn = 0, best_ntn = 0, best_ks = 1;
  double best_cost = 0;
  const int bn_cap = std::min(256, (g.N + 15) / 16 * 16);
  for (int b = bn_cap; b >= 16; b -= 16) {
    const int ntn = (g.N + b - 1) / b;
    if (ntn > 1 && (b % 32)) continue;   // the epilogue stores 32-channel boxes: inner N tiles must be whole boxes
    if (narrow && S * b > 256) continue;
    const int w_bytes = S * b * 128;
    const int epi_bytes = EPI_WARPS * 4096 + (ntn == 1 ? 1 : EPI_WARPS) * b * 4;  // per epilogue warp: one 4 KB staging/aux tile; bias copies
    const int budget = 227 * 1024 - 1024 - 1024 - epi_bytes;   // alignment slack + barriers + epilogue
    static const int force_tz = getenv("E2_ZS_TZ") ? atoi(getenv("E2_ZS_TZ")) : 0;   // experiments only
    for (int tz = 8; tz >= 1; --tz) {
      if (force_tz && tz != force_tz) continue;
      if (g.fuse_pool && g.qz == 2 && (tz & 1)) continue;   // a tile holds whole pool windows
      if (std::min(tz, S) * b > 256) continue;   // widest stacked MMA
      if (2 * tz * b > 512) continue;            // double-buffered accumulators in TMEM
      const int np = tz + S - 1;
      if (np > 11) continue;
      if (np * p->plane_stride + 2 * w_bytes > budget) continue;
      int rest = budget - np * p->plane_stride;
      int wsl = std::min(MAX_WSLOTS, rest / w_bytes);
      int dbl = 0;
      // double buffering the planes is worth more than weight slots beyond the third
      if (wsl >= 3 && rest - 3 * w_bytes >= np * p->plane_stride) {
        dbl = 1;
        wsl = std::min(MAX_WSLOTS, (rest - np * p->plane_stride) / w_bytes);
      }
      const int ntz = (g.Oz + tz - 1) / tz;
      const int64_t tiles = (int64_t)g.On * ntz * ntx * nty * ntn;
      double mma = 0;
      for (int q = 0; q < np; ++q) {
        const int nblk = std::min(q, tz - 1) - std::max(0, q - (S - 1)) + 1;
        const double n = nblk * b;
        mma += std::max(std::max(48.0, 32.0 + n / 4.0), n / 2.0);   // operand fetch (128 B/clk) or math bound
      }
      const double stage = std::max(mma * nk_avg + kStageOvh, kLatency / (wsl - 1));
      for (int ks : ks_opts) {
        if (ks > 1 && (!may_split || no_split || ks > CBn)) break;
        const int cb_per = (CBn + ks - 1) / ks;
        if ((CBn + cb_per - 1) / cb_per != ks) continue;   // same split as a smaller ks
        const int64_t waves = (tiles * ks + h->sm_count - 1) / h->sm_count;
        double cost = (double)waves * (cb_per * T9 * stage + (dbl ? 0.0 : kLatency) + 200.0) + tz * (b / 32.0) / 2.0 * kEpiChunk;
        if (ks > 1) cost += kReduceFixed + (ks + 1.0) * c_bytes / kReduceBpc;
        if (best_tz == 0 || cost < best_cost * 0.97)
          best_tz = tz, best_cost = cost, best_w = wsl, best_dbl = dbl, best_bn = b, best_ntn = ntn, best_ks = ks;
      }
    }
  }
  if (best_tz == 0) return false;
  const int bn = best_bn;
  p->BN = bn, p->ntn = best_ntn;
  p->bias_copies = best_ntn == 1 ? 1 : EPI_WARPS;
  p->wblk_bytes = bn * 128;
  p->w_bytes = S * bn * 128;
  p->TZ = best_tz;
  p->NP = best_tz + S - 1;
  p->nslot = best_dbl ? 2 * p->NP : p->NP;
  p->wslot = best_w;
  p->acc_bufs = 2;
  int cols = 32;
  while (cols < p->acc_bufs * p->TZ * bn) cols *= 2;
  p->tmem_cols = cols;
  p->On = g.On, p->Oz = g.Oz, p->Ox = g.Ox, p->Oy = g.Oy;
  p->ntz = (g.Oz + p->TZ - 1) / p->TZ, p->ntx = ntx, p->nty = nty;
  p->kz = g.tz, p->kx = g.tx, p->ky = g.ty, p->oz = g.oz, p->ox = g.ox, p->oy = g.oy;
  p->K = g.K, p->N = g.N, p->CB = CBn;
  p->cb_per = (CBn + best_ks - 1) / best_ks;
  p->ksplit = (CBn + p->cb_per - 1) / p->cb_per;
  p->num_tiles = p->On * p->ntz * p->ntx * p->nty * p->ntn * p->ksplit;
  p->epi_off = (p->nslot * p->plane_stride + p->wslot * p->w_bytes + 1023) / 1024 * 1024;
  // tile-quantisation efficiency: useful outputs / computed outputs
  const double eff = (double)g.Oz * g.Ox * g.Oy * g.N /
                     ((double)p->ntz * p->TZ * p->ntx * TX * p->nty * TY * p->ntn * bn);
  // Big layers with badly fitting planes are left to the tap kernel (its tiles are free-form position boxes).  Small
  // ones stay here even at low efficiency: their time is launch / pipeline latency either way, and the alternative
  // halo-plane kernel needed 75-115 us for the decoder layers of unet3d_litelite (here: ~19 us).
  if (eff < 0.5 && !(eff >= 0.25 && best_cost < 120000.0)) return false;
  return true;
}

static size_t zs_ws_bytes(const GatherGemm& g, const ZsParams& p) {
  if (p.ksplit <= 1) return 0;
  return (size_t)p.ksplit * g.On * g.Oz * g.Ox * g.Oy * ((g.N + 3) / 4 * 4) * sizeof(float);
}

size_t e2_conv_zstack_workspace_bytes(int sm_count, const GatherGemm& g) {
  e2_handle fake;
  memset(&fake, 0, sizeof(fake));
  fake.sm_count = sm_count;
  ZsParams p;
  if (!plan_zstack(&fake, g, &p, true)) return 0;
  return zs_ws_bytes(g, p);
}

// debug / test hook: the plan for a problem as 8 ints {BN, TZ, ksplit, cb_per, wslot, nslot, num_tiles, ntn}
extern "C" int e2_debug_zstack_plan(int sm_count, int K, int N, int Oz, int Ox, int Oy, int kz, int kx, int ky, int may_split,
                                    int* out) {
  e2_handle fake;
  memset(&fake, 0, sizeof(fake));
  fake.sm_count = sm_count;
  GatherGemm g;
  memset(&g, 0, sizeof(g));
  g.K = K, g.N = N, g.On = 1, g.Oz = Oz, g.Ox = Ox, g.Oy = Oy, g.tz = kz, g.tx = kx, g.ty = ky, g.sz = g.sx = g.sy = 1;
  g.c_pitch = (N + 3) / 4 * 4;
  ZsParams p;
  if (!plan_zstack(&fake, g, &p, may_split != 0)) return 0;
  out[0] = p.BN, out[1] = p.TZ, out[2] = p.ksplit, out[3] = p.cb_per, out[4] = p.wslot, out[5] = p.nslot, out[6] = p.num_tiles,
  out[7] = p.ntn;
  return 1;
}

// same for the conv + fused max-pool plan (window qz,qx,qy): 1 if e2_conv3d_fwd_pool would fuse the pair
extern "C" int e2_debug_zstack_pool_plan(int sm_count, int K, int N, int Oz, int Ox, int Oy, int kz, int kx, int ky, int qz,
                                         int qx, int qy, int* out) {
  e2_handle fake;
  memset(&fake, 0, sizeof(fake));
  fake.sm_count = sm_count;
  GatherGemm g;
  memset(&g, 0, sizeof(g));
  alignas(16) static float dummy[4];
  g.K = K, g.N = N, g.On = 1, g.Oz = Oz, g.Ox = Ox, g.Oy = Oy, g.tz = kz, g.tx = kx, g.ty = ky, g.sz = g.sx = g.sy = 1;
  g.c_pitch = g.cp_pitch = (N + 3) / 4 * 4;
  g.fuse_pool = 1, g.qz = qz, g.qx = qx, g.qy = qy, g.Cp = dummy;
  if (!e2_conv_zstack_pool_ok(&fake, g)) return 0;
  ZsParams p;
  if (!plan_zstack(&fake, g, &p, false)) return 0;
  out[0] = p.BN, out[1] = p.TZ, out[2] = p.ksplit, out[3] = p.cb_per, out[4] = p.wslot, out[5] = p.nslot, out[6] = p.num_tiles,
  out[7] = p.ntn;
  return 1;
}

bool e2_conv_zstack_tc_ok(const e2_handle* h, const GatherGemm& g) {
  ZsParams p;
  return plan_zstack(h, g, &p, g.ws != nullptr);
}

bool e2_conv_zstack_pool_ok(const e2_handle* h, const GatherGemm& g) {
  if (!g.fuse_pool) return false;
  // the plan without the pool: if it splits K (few tiles, long K loop) the epilogue runs in k_zstack_reduce and
  // fusing would cost the split
  GatherGemm plain = g;
  plain.fuse_pool = 0;
  alignas(16) static float dummy_ws[4];
  plain.ws = dummy_ws;
  if (!plain.C) plain.C = g.Cp;
  ZsParams p;
  if (!plan_zstack(h, plain, &p, true) || p.ksplit > 1) return false;
  return plan_zstack(h, g, &p, false);
}

int e2_launch_conv_zstack_tc(e2_handle* h, const GatherGemm& g, cudaStream_t s) {
  EncodeTiledFn enc = e2_get_tmap_encode();
  if (!enc) return e2_fail(h, E2_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled entry point not available");
  ZsParams p;
  bool may_split = g.ws && !(reinterpret_cast<uintptr_t>(g.ws) & 15);
  if (!plan_zstack(h, g, &p, may_split)) return e2_fail(h, E2_ERR_UNSUPPORTED, "conv_zstack_tc: problem does not qualify");
  if (zs_ws_bytes(g, p) > g.ws_bytes && !plan_zstack(h, g, &p, false))
    return e2_fail(h, E2_ERR_UNSUPPORTED, "conv_zstack_tc: problem does not qualify");
  const bool split = p.ksplit > 1;
  // split-K: the kernel stores raw partial tiles, the epilogue moves to k_zstack_reduce
  p.C = split ? static_cast<float*>(g.ws) : g.C;
  p.c_pitch = split ? (g.N + 3) / 4 * 4 : g.c_pitch;
  p.bias = split ? nullptr : g.bias, p.gate = split ? nullptr : g.gate;
  p.act = split ? E2_ACT_LIN : g.act, p.accumulate = split ? 0 : g.accumulate, p.round_tf32 = split ? 0 : g.round_tf32;
  p.idesc0 = tc::make_idesc(2 /*TF32*/, 0, 0, 128, 0);
  {
    const char* e = getenv("E2_ZS_DBG");
    p.dbg = e ? atoi(e) : 0;
    if (getenv("E2_ZS_INFO")) fprintf(stderr, "zstack: TZ %d NP %d nslot %d wslot %d BN %d CB %d ksplit %d units %d smem plane %d w %d\n", p.TZ, p.NP, p.nslot, p.wslot, p.BN, p.CB, p.ksplit, p.num_tiles, p.plane_stride, p.w_bytes);
  }
  p.idesc_step = (uint32_t)(p.BN >> 3) << 17;
  if (g.fuse_pool) {
    if (split) return e2_fail(h, E2_ERR_UNSUPPORTED, "conv_zstack_tc: fused pool with a K split");
    p.pool = 1, p.ppz = g.qz, p.ppx = g.qx, p.ppy = g.qy;
    p.store_full = g.C != nullptr, p.pool_amax = g.Ci != nullptr;
    p.kz0 = g.keep[0], p.kz1 = g.keep[1], p.kx0 = g.keep[2], p.kx1 = g.keep[3], p.ky0 = g.keep[4], p.ky1 = g.keep[5];
    p.pbias = g.pbias, p.pact = g.pact, p.pround = g.pround;
  } else {
    p.store_full = 1;
  }
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
  // C (and the gate, same geometry) as 32-channel x 8 y x 4 x boxes for the epilogue's TMA store / aux loads
  CUtensorMap tmC, tmG;
  for (int which = 0; which < 2; ++which) {
    const float* base = which == 0 ? p.C : p.gate;
    if (which == 0 && !base) base = g.Cp;   // fused pool without the unpooled tensor: tmC is never used, keep it valid
    if (!base) {
      tmG = tmC;
      continue;
    }
    cuuint64_t dims[5] = {(cuuint64_t)g.N, (cuuint64_t)g.Oy, (cuuint64_t)g.Ox, (cuuint64_t)g.Oz, (cuuint64_t)g.On * p.ksplit};
    cuuint64_t pitch = (cuuint64_t)p.c_pitch * 4;
    cuuint64_t strides[4] = {pitch, pitch * g.Oy, pitch * g.Oy * g.Ox, pitch * g.Oy * g.Ox * g.Oz};
    cuuint32_t box[5] = {32, 8, 4, 1, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(which == 0 ? &tmC : &tmG, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(base), dims, strides,
                     box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return e2_fail(h, E2_ERR_CUDA, "cuTensorMapEncodeTiled(output) failed: %d", (int)r);
  }
  // pooled values / argmax as [32 ch x 8/qy x 4/qx] boxes (one per epilogue warp and plane pair)
  CUtensorMap tmP = tmC, tmI = tmC;
  if (g.fuse_pool) {
    for (int which = 0; which < 2; ++which) {
      void* base = which == 0 ? static_cast<void*>(g.Cp) : static_cast<void*>(g.Ci);
      if (!base) continue;
      cuuint64_t dims[5] = {(cuuint64_t)g.N, (cuuint64_t)(g.Oy / g.qy), (cuuint64_t)(g.Ox / g.qx), (cuuint64_t)(g.Oz / g.qz),
                            (cuuint64_t)g.On};
      cuuint64_t pitch = (cuuint64_t)g.cp_pitch * 4;
      cuuint64_t strides[4] = {pitch, pitch * dims[1], pitch * dims[1] * dims[2], pitch * dims[1] * dims[2] * dims[3]};
      cuuint32_t box[5] = {32, (cuuint32_t)(TY / g.qy), (cuuint32_t)(4 / g.qx), 1, 1};
      cuuint32_t es[5] = {1, 1, 1, 1, 1};
      CUresult r = enc(which == 0 ? &tmP : &tmI, which == 0 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_INT32,
                       5, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return e2_fail(h, E2_ERR_CUDA, "cuTensorMapEncodeTiled(pooled) failed: %d", (int)r);
    }
  }
  const size_t smem = 1024 + (size_t)p.epi_off + EPI_WARPS * 4096 + (size_t)p.bias_copies * p.BN * 4 + (2 * MAX_PSLOTS + 2 * MAX_WSLOTS + 4 + EPI_WARPS) * 8 + 16;
  typedef void (*KernelFn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap,
                           const CUtensorMap, const ZsParams);
#define ZS_ROW(KZ, P)                                                                                            \
  {k_conv_zstack_tc<1, KZ, P>, k_conv_zstack_tc<2, KZ, P>, k_conv_zstack_tc<3, KZ, P>, k_conv_zstack_tc<4, KZ, P>,     \
   k_conv_zstack_tc<5, KZ, P>, k_conv_zstack_tc<6, KZ, P>, k_conv_zstack_tc<7, KZ, P>, k_conv_zstack_tc<8, KZ, P>}
  static const KernelFn table[2][4][8] = {{ZS_ROW(1, 0), ZS_ROW(2, 0), ZS_ROW(3, 0), ZS_ROW(4, 0)},
                                          {ZS_ROW(1, 1), ZS_ROW(2, 1), ZS_ROW(3, 1), ZS_ROW(4, 1)}};
#undef ZS_ROW
  if (p.kz < 1 || p.kz > 4 || p.TZ < 1 || p.TZ > 8) return e2_fail(h, E2_ERR_UNSUPPORTED, "conv_zstack_tc: no kernel for TZ %d kz %d", p.TZ, p.kz);
  KernelFn fn = table[p.pool ? 1 : 0][p.kz - 1][p.TZ - 1];
  if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)) != cudaSuccess)
    return e2_fail(h, E2_ERR_CUDA, "cudaFuncSetAttribute(max dynamic smem) failed");
  if (smem > 227 * 1024) return e2_fail(h, E2_ERR_UNSUPPORTED, "conv_zstack_tc: shared memory plan exceeds 227 KB");
  const int grid = std::min(p.num_tiles, h->sm_count);
  if (getenv("E2_ZS_TRACE")) {
    cudaMalloc(&p.trace, 32 * sizeof(long long));
    cudaMemset(p.trace, 0, 32 * sizeof(long long));
  }
  fn<<<grid, ZS_THREADS, smem, s>>>(tmA, tmB, tmC, tmG, tmP, tmI, p);
  if (p.trace) {
    long long tr[32];
    cudaMemcpy(tr, p.trace, sizeof(tr), cudaMemcpyDeviceToHost);
    cudaFree(p.trace);
    const int tiles0 = (p.num_tiles + grid - 1) / grid;
    fprintf(stderr, "zstack trace (CTA 0, %d tiles): mma warp total %lld  wait acc_empty %lld  wait w_full %lld  issue %lld | epi0 total %lld wait acc_full %lld work %lld | epi4 total %lld wait %lld work %lld\n",
            tiles0, tr[0], tr[1], tr[2], tr[3], tr[8], tr[9], tr[10], tr[12], tr[13], tr[14]);
  }
  e2_count_launch(h);
  E2_CUDA_CHECK(h, "conv_zstack_tc");
  if (split) {
    const int64_t positions = (int64_t)g.On * g.Oz * g.Ox * g.Oy;
    const int Np = p.c_pitch;
    k_zstack_reduce<<<e2_grid_1d(positions * (Np / 4), 256, h->sm_count, 16), 256, 0, s>>>(
        static_cast<const float*>(g.ws), p.ksplit, positions, Np, g.N, g.C, g.c_pitch, g.bias, g.gate,
        g.act, g.accumulate, g.round_tf32);
    e2_count_launch(h);
    E2_CUDA_CHECK(h, "zstack_reduce");
  }
  return E2_OK;
}
