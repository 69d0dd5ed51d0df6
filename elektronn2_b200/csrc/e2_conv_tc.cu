// tcgen05 / TMEM / TMA implicit-GEMM kernels (E2_COMPUTE_TF32).
//
// gather-GEMM  C[m, n] = sum_{tap, k} A[pos(m) + tap + org][k] * B[n][tap][k]
//   M tile  = 128 output positions forming a (tz, tx, ty) box of the output grid, so that for
//             every filter tap the A operand is ONE 5-D TMA box load of the activation tensor
//             shifted by the tap offset: rows land in shared memory in (z,x,y) order, 32 fp32
//             channels (128 B) per row, 128B-swizzled -> exactly the canonical K-major UMMA
//             layout.  Out-of-range coordinates (dgrad halo, ragged edge tiles, channel tail)
//             are zero-filled by the TMA unit: no padding copies, no im2col buffer.
//   N tile  = up to 256 output channels, B operand = packed weights [n][tap][k], 3-D TMA box.
//   K loop  = taps x ceil(K/32) blocks through a multi-stage mbarrier ring.
//   MMA     = tcgen05.mma.cta_group::1.kind::tf32, M=128, N=BN, K=8, fp32 accumulators in TMEM.
//   roles   = warp 0: TMA producer, warp 1: MMA issuer, warps 2-5: epilogue
//             (tcgen05.ld -> +bias -> activation -> (accumulate) -> store).
// Several CTAs share an SM (smem permitting), so one CTA's epilogue overlaps another's MMAs.
#include <algorithm>
#include <stdlib.h>
#include "e2_common.cuh"
#include "e2_conv_internal.cuh"
#include "e2_tc_ptx.cuh"

namespace {

constexpr int BM = 128;          // rows per tile (TMEM lanes)
constexpr int BK = 32;           // fp32 per K block = 128 bytes = one swizzle row
constexpr int A_STAGE_BYTES = BM * BK * 4;
constexpr int EPI_WARPS = 8;     // two per TMEM lane quadrant, alternating over the 32-column chunks of the tile
constexpr int NUM_THREADS = (2 + EPI_WARPS) * 32;

struct TcParams {
  int On, Oz, Ox, Oy;        // output position grid
  int tz, tx, ty;            // tile box (tz*tx*ty <= 128 rows; the rest of the 128-row operand tile is masked)
  int ntz, ntx, nty;         // tiles per axis
  int kz, kx, ky;            // taps
  int oz, ox, oy;            // origin
  int K, N;                  // reduction channels per tap, output channels (GEMM N)
  int BN;                    // N tile (multiple of 16, <= 256)
  int stages;
  int tmem_cols;
  float* C;
  int c_pitch;
  const float* bias;
  int act, accumulate, round_tf32;
  int shuffle, pz, px, py, Fo;
  int sz, sx, sy;            // position stride of the gather (upconv dgrad: pool factors)
  const float* gate;         // fused ReLU backward
  uint32_t idesc;
  int its_per_split;         // K iterations (tap x channel block) per blockIdx.z
  float* part;               // split-K: raw partial tiles [z][n tile][m tile][col/4][row][4]; NULL: fused epilogue
};

// Epilogue of one transposed 32 x 32 chunk: this lane serves rows rsub, rsub+4, ... (8 lanes per 128-byte row).
// ACT: 0 lin, 1 relu, 2 anything else -- compile-time so the row loop carries no per-value activation branches.
template <int ACT>
__device__ __forceinline__ void tc_store_rows(const TcParams& p, const uint8_t* st, const int64_t (&rowofs)[8], uint32_t rowmask,
                                              int64_t colofs, float4 b4, int rsub, int pc) {
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    if (!((rowmask >> it) & 1u)) continue;
    const int rl = it * 4 + rsub;
    const int64_t ofs = rowofs[it] + colofs;
    const float4 raw = *reinterpret_cast<const float4*>(st + rl * 128 + ((pc ^ (rl & 7)) << 4));
    float4 v = make_float4(raw.x + b4.x, raw.y + b4.y, raw.z + b4.z, raw.w + b4.w);
    if (ACT == 1) {
      v.x = fmaxf(v.x, 0.f), v.y = fmaxf(v.y, 0.f), v.z = fmaxf(v.z, 0.f), v.w = fmaxf(v.w, 0.f);
    } else if (ACT == 2) {
      v.x = e2_apply_act(v.x, p.act), v.y = e2_apply_act(v.y, p.act), v.z = e2_apply_act(v.z, p.act), v.w = e2_apply_act(v.w, p.act);
    }
    if (p.gate) {
      const float4 g4 = __ldg(reinterpret_cast<const float4*>(p.gate + ofs));
      v.x = g4.x > 0.f ? v.x : 0.f, v.y = g4.y > 0.f ? v.y : 0.f;
      v.z = g4.z > 0.f ? v.z : 0.f, v.w = g4.w > 0.f ? v.w : 0.f;
    }
    if (p.accumulate) {
      const float4 c4 = *reinterpret_cast<const float4*>(p.C + ofs);
      v.x += c4.x, v.y += c4.y, v.z += c4.z, v.w += c4.w;
    }
    if (p.round_tf32) v.x = e2_round_tf32(v.x), v.y = e2_round_tf32(v.y), v.z = e2_round_tf32(v.z), v.w = e2_round_tf32(v.w);
    *reinterpret_cast<float4*>(p.C + ofs) = v;
  }
}

__global__ void __launch_bounds__(NUM_THREADS) k_gather_gemm_tc(const __grid_constant__ CUtensorMap tmA,
                                                                const __grid_constant__ CUtensorMap tmB,
                                                                const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 128B swizzle needs 1024-byte aligned stage bases
  // aligned base by OFFSET arithmetic on smem_raw (no integer round trip), so the pointer keeps its address space and the
  // warps' own accesses compile to LDS / STS instead of generic LD / ST
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  const int b_stage_bytes = p.BN * BK * 4;
  uint8_t* smA = smem;
  uint8_t* smB = smem + p.stages * A_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smB + p.stages * b_stage_bytes);
  uint64_t* full = bars;                 // [stages]
  uint64_t* empty = bars + p.stages;     // [stages]
  uint64_t* acc_full = bars + 2 * p.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // tile coordinates
  int tile = blockIdx.x;
  const int ity = tile % p.nty;
  tile /= p.nty;
  const int itx = tile % p.ntx;
  tile /= p.ntx;
  const int itz = tile % p.ntz;
  const int in_ = tile / p.ntz;
  const int z0 = itz * p.tz, x0 = itx * p.tx, y0 = ity * p.ty;
  const int n0 = blockIdx.y * p.BN;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmA);
    tc::prefetch_tmap(&tmB);
    for (int s = 0; s < p.stages; ++s) {
      tc::mbar_init(&full[s], 1);
      tc::mbar_init(&empty[s], 1);
    }
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

  const int kblocks = (p.K + BK - 1) / BK;
  const int T = p.kz * p.kx * p.ky;
  const int it_begin = blockIdx.z * p.its_per_split;
  const int it_end = min(T * kblocks, it_begin + p.its_per_split);

  if (warp == 0) {
    // ------------------------------------------------------------- TMA producer
    if (lane == 0) {
      for (int it = it_begin; it < it_end; ++it) {
        const int tap = it / kblocks, kb = it - tap * kblocks;
        const int k3 = tap % p.ky, j3 = (tap / p.ky) % p.kx, i3 = tap / (p.ky * p.kx);
        const int li = it - it_begin;
        const int s = li % p.stages;
        const uint32_t ph = (uint32_t)(li / p.stages) & 1u;
        tc::mbar_wait(&empty[s], ph ^ 1u);
        tc::mbar_arrive_expect_tx(&full[s], (uint32_t)(p.tz * p.tx * p.ty * BK * 4 + b_stage_bytes));
        tc::tma_load_5d(smA + s * A_STAGE_BYTES, &tmA, &full[s], kb * BK, y0 * p.sy + k3 + p.oy,
                        x0 * p.sx + j3 + p.ox, z0 * p.sz + i3 + p.oz, in_);
        tc::tma_load_3d(smB + s * b_stage_bytes, &tmB, &full[s], kb * BK, tap, n0);
      }
    }
  } else if (warp == 1) {
    // --------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      const uint64_t desc_tmpl = tc::make_smem_desc(0, 16, 1024, 2);
      const uint32_t smA_addr = tc::smem_u32(smA), smB_addr = tc::smem_u32(smB);
      for (int li = 0; li < it_end - it_begin; ++li) {
        const int s = li % p.stages;
        const uint32_t ph = (uint32_t)(li / p.stages) & 1u;
        tc::mbar_wait(&full[s], ph);
        tc::tc_fence_after();
        // K-major, 128B swizzle: rows 128 B apart, 8-row groups 1024 B apart; advancing K by
        // 8 tf32 = 32 bytes inside the swizzle row is a plain start-address advance (+2 in the
        // descriptor's 16-byte start-address field)
        const uint64_t ad = desc_tmpl + (uint64_t)((smA_addr + (uint32_t)(s * A_STAGE_BYTES)) >> 4);
        const uint64_t bd = desc_tmpl + (uint64_t)((smB_addr + (uint32_t)(s * b_stage_bytes)) >> 4);
        tc::mma_tf32_ss(tmem_base, ad, bd, p.idesc, li > 0 ? 1u : 0u);
        tc::mma_tf32_ss(tmem_base, ad + 2, bd + 2, p.idesc, 1u);
        tc::mma_tf32_ss(tmem_base, ad + 4, bd + 4, p.idesc, 1u);
        tc::mma_tf32_ss(tmem_base, ad + 6, bd + 6, p.idesc, 1u);
        tc::mma_commit(&empty[s]);   // frees the smem stage when these MMAs have read it
      }
      tc::mma_commit(acc_full);      // accumulator complete
    }
  } else {
    // ----------------------------------------------------------------- epilogue
    // A lane of tcgen05.ld owns one ROW of the tile (32 channels = 128 B); storing that directly makes a warp
    // instruction touch 32 different lines.  Every chunk is therefore transposed through the (now idle) operand
    // ring so that 8 lanes cover one row: each store / gate / accumulate access is a full 128-byte line, and bias /
    // activation run in the transposed domain.  Row offsets are decoded once per tile; the chunk loop is kept
    // rolled -- an unrolled version of this epilogue was 17 k instructions and instruction-fetch bound.
    const int q = warp & 3;                 // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;       // warps w and w+4 share a quadrant and alternate over the chunks
    tc::mbar_wait(acc_full, 0);
    tc::tc_fence_after();
    uint8_t* st = smA + (warp - 2) * 4096;
    const int r8 = lane & 7, pc = lane & 7, rsub = lane >> 3;
    const bool vec_ok = (p.N & 3) == 0 && (p.c_pitch & 3) == 0 && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) &&
                        (!p.gate || (reinterpret_cast<uintptr_t>(p.gate) & 15) == 0) && (!p.shuffle || (p.Fo & 3) == 0);
    // vector path: this lane serves rows rsub, rsub+4, ... of the warp's 32; scalar path: its own row `lane`
    int64_t rowofs[8];
    uint32_t rowmask = 0;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int rr = q * 32 + (vec_ok ? it * 4 + rsub : lane);
      const int ry = rr % p.ty, rx = (rr / p.ty) % p.tx, rz = rr / (p.ty * p.tx);
      const int pz_ = z0 + rz, px_ = x0 + rx, py_ = y0 + ry;
      if (rz < p.tz && pz_ < p.Oz && px_ < p.Ox && py_ < p.Oy) rowmask |= 1u << it;   // rz >= tz: row behind the box
      rowofs[it] = p.shuffle ? ((((int64_t)in_ * (p.Oz * p.pz) + pz_ * p.pz) * (p.Ox * p.px) + px_ * p.px) * (p.Oy * p.py) +
                                py_ * p.py) * p.c_pitch
                             : ((((int64_t)in_ * p.Oz + pz_) * p.Ox + px_) * p.Oy + py_) * p.c_pitch;
      if (!vec_ok && it == 0) break;
    }
#pragma unroll 1
    for (int c0 = half * 32; c0 < p.BN; c0 += 64) {
      uint32_t r[32];
      if (p.BN - c0 >= 32) {
        tc::tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
      } else {
        tc::tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
#pragma unroll
        for (int j = 16; j < 32; ++j) r[j] = 0u;
      }
      tc::tmem_ld_wait();
      if (p.part) {
        // split-K: raw accumulators, one coalesced float4 per row and 4 columns; the reduce kernel sums
        // the splits and applies the epilogue
        float4* dst = reinterpret_cast<float4*>(p.part) +
                      ((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * (p.BN / 4) * BM + q * 32 + lane;
#pragma unroll
        for (int j4 = 0; j4 < 32; j4 += 4)
          if (c0 + j4 < p.BN)
            dst[(size_t)((c0 + j4) / 4) * BM] = make_float4(__uint_as_float(r[j4]), __uint_as_float(r[j4 + 1]),
                                                           __uint_as_float(r[j4 + 2]), __uint_as_float(r[j4 + 3]));
        continue;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<uint4*>(st + lane * 128 + ((j ^ r8) << 4)) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
      __syncwarp();
      if (vec_ok) {
        const int nb = n0 + c0 + pc * 4;                 // first of this lane's 4 columns
        if (nb < p.N) {
          int ch = nb;
          int64_t colofs = nb;
          if (p.shuffle) {
            const int tp = nb / p.Fo;
            ch = nb - tp * p.Fo;
            const int k3 = tp % p.py, j3 = (tp / p.py) % p.px, i3 = tp / (p.py * p.px);
            colofs = (((int64_t)i3 * (p.Ox * p.px) + j3) * (p.Oy * p.py) + k3) * p.c_pitch + ch;
          }
          float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (p.bias) b4 = make_float4(__ldg(p.bias + ch), __ldg(p.bias + ch + 1), __ldg(p.bias + ch + 2), __ldg(p.bias + ch + 3));
          if (p.act == E2_ACT_RELU) tc_store_rows<1>(p, st, rowofs, rowmask, colofs, b4, rsub, pc);
          else if (p.act == E2_ACT_LIN) tc_store_rows<0>(p, st, rowofs, rowmask, colofs, b4, rsub, pc);
          else tc_store_rows<2>(p, st, rowofs, rowmask, colofs, b4, rsub, pc);
        }
      } else if (rowmask & 1u) {
        // scalar path (channel counts / pointers that are not 16-byte friendly): one element at a time
#pragma unroll 1
        for (int j = 0; j < 32; ++j) {
          const int n = n0 + c0 + j;
          if (n >= p.N) break;
          int ch = n;
          int64_t ofs = rowofs[0] + n;
          if (p.shuffle) {
            const int tp = n / p.Fo;
            ch = n - tp * p.Fo;
            const int k3 = tp % p.py, j3 = (tp / p.py) % p.px, i3 = tp / (p.py * p.px);
            ofs = rowofs[0] + (((int64_t)i3 * (p.Ox * p.px) + j3) * (p.Oy * p.py) + k3) * p.c_pitch + ch;
          }
          float a = *reinterpret_cast<const float*>(st + lane * 128 + (((j >> 2) ^ r8) << 4) + ((j & 3) << 2));
          if (p.bias) a += __ldg(p.bias + ch);
          a = e2_apply_act(a, p.act);
          if (p.gate && !(__ldg(p.gate + ofs) > 0.f)) a = 0.f;
          if (p.accumulate) a += p.C[ofs];
          if (p.round_tf32) a = e2_round_tf32(a);
          p.C[ofs] = a;
        }
      }
      __syncwarp();     // the staging tile is rewritten by the next chunk
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// ------------------------------------------------------------------ host side
EncodeTiledFn get_encode() { return e2_get_tmap_encode(); }

// Split-K second pass: sum the partial tiles over the K splits and apply the epilogue of the fused path
// (+bias -> act -> ReLU gate -> accumulate -> tf32 round; pixel shuffle for upconv forward).
// One thread per (tile, 4 columns, row): the partial reads are coalesced float4.
__global__ void __launch_bounds__(BM) k_gather_gemm_reduce(const TcParams p, int ksplit, int tiles_m, int tiles_n) {
  const int row = threadIdx.x;
  int b = blockIdx.x;
  const int c4 = b % (p.BN / 4);
  b /= (p.BN / 4);
  const int tm = b % tiles_m;
  const int tn = b / tiles_m;
  const size_t tile_f4 = (size_t)(p.BN / 4) * BM;
  const float4* src = reinterpret_cast<const float4*>(p.part) + ((size_t)tn * tiles_m + tm) * tile_f4 + (size_t)c4 * BM + row;
  const size_t zstride = (size_t)tiles_n * tiles_m * tile_f4;
  float4 a4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int z = 0; z < ksplit; ++z) {
    const float4 v = __ldcg(src + (size_t)z * zstride);
    a4.x += v.x, a4.y += v.y, a4.z += v.z, a4.w += v.w;
  }
  int tile = tm;
  const int ity = tile % p.nty;
  tile /= p.nty;
  const int itx = tile % p.ntx;
  tile /= p.ntx;
  const int itz = tile % p.ntz;
  const int in_ = tile / p.ntz;
  const int ly = row % p.ty, lx = (row / p.ty) % p.tx, lz = row / (p.ty * p.tx);
  const int oz = itz * p.tz + lz, ox = itx * p.tx + lx, oy = ity * p.ty + ly;
  if (lz >= p.tz || oz >= p.Oz || ox >= p.Ox || oy >= p.Oy) return;
  const int64_t pos = (((int64_t)in_ * p.Oz + oz) * p.Ox + ox) * p.Oy + oy;
  const float v[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int n = tn * p.BN + c4 * 4 + e;
    if (n >= p.N) continue;
    float a = v[e];
    int64_t ofs;
    int ch = n;
    if (p.shuffle) {
      const int tp = n / p.Fo;
      ch = n - tp * p.Fo;
      const int k3 = tp % p.py, j3 = (tp / p.py) % p.px, i3 = tp / (p.py * p.px);
      ofs = ((((int64_t)in_ * (p.Oz * p.pz) + oz * p.pz + i3) * (p.Ox * p.px) + ox * p.px + j3) * (p.Oy * p.py) +
             oy * p.py + k3) * p.c_pitch + ch;
    } else {
      ofs = pos * p.c_pitch + n;
    }
    if (p.bias) a += __ldg(p.bias + ch);
    a = e2_apply_act(a, p.act);
    if (p.gate && !(__ldg(p.gate + ofs) > 0.f)) a = 0.f;
    if (p.accumulate) a += p.C[ofs];
    if (p.round_tf32) a = e2_round_tf32(a);
    p.C[ofs] = a;
  }
}

}  // namespace

EncodeTiledFn e2_get_tmap_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  }
  return fn;
}

namespace {

bool pick_tile(int Oz, int Ox, int Oy, int sz, int sx, int sy, int* tz, int* tx, int* ty) {
  // The M tile is a (tz, tx, ty) box of up to 128 output positions -- ANY extents, not only powers of two: the TMA box
  // fills tz*tx*ty rows of the 128-row operand tile and the rows behind them are masked in the epilogue (MMA rows are
  // independent).  Every tile costs a full M = 128 MMA whatever its fill, so the box that needs the fewest tiles wins
  // (unet3d conv7, 7x9x9 positions: five (7,2,9) boxes at 90 % fill instead of nine (8,4,4) boxes at 49 %);
  // ties go to the longer y run (contiguous stores), then to the smaller padded volume.
  static const bool pow2_only = getenv("E2_TC_POW2_TILES") != nullptr;     // A/B switch: round-1 behaviour
  int64_t best_tiles = -1, best_vol = 0;
  int best_ty = 0;
  for (int z = 1; z <= 128 && z <= Oz; ++z) {
    if (z * sz > 256) break;                                              // TMA box extent limit (strided gather)
    for (int x = 1; x * z <= 128 && x <= Ox; ++x) {
      if (x * sx > 256) break;
      const int ymax = std::min(std::min(128 / (z * x), Oy), 256 / sy);
      if (ymax < 1) continue;
      const int ny = (Oy + ymax - 1) / ymax;
      const int cand[2] = {ymax, (Oy + ny - 1) / ny};                     // widest, and the balanced split of the y axis
      for (int y : cand) {
        if (pow2_only && ((z & (z - 1)) || (x & (x - 1)) || (y & (y - 1)) || z * x * y != 128)) continue;
        const int64_t tiles = (int64_t)((Oz + z - 1) / z) * ((Ox + x - 1) / x) * ((Oy + y - 1) / y);
        const int64_t vol = tiles * z * x * y;
        if (best_tiles < 0 || tiles < best_tiles || (tiles == best_tiles && (y > best_ty || (y == best_ty && vol < best_vol)))) {
          best_tiles = tiles, best_vol = vol, best_ty = y;
          *tz = z, *tx = x, *ty = y;
        }
      }
    }
  }
  if (best_tiles < 0 && pow2_only) {   // shapes smaller than every power-of-two box: fall back to the free search
    *tz = std::min(Oz, 128), *tx = std::min(Ox, 128 / *tz), *ty = std::min(Oy, 128 / (*tz * *tx));
    return true;
  }
  return best_tiles > 0;
}

}  // namespace

bool e2_gather_gemm_tc_ok(const e2_handle* h, const GatherGemm& g) {
  if (!get_encode()) return false;
  if (g.sz < 1 || g.sx < 1 || g.sy < 1 || g.sz > 8 || g.sx > 8 || g.sy > 8) return false;   // TMA element-stride limit
  if (g.K < 8 || g.N < 8) return false;                    // not tensor-core shaped (HBM-bound layers)
  if (g.a_pitch % 4 || (reinterpret_cast<uintptr_t>(g.A) & 15) || (reinterpret_cast<uintptr_t>(g.B) & 15)) return false;
  if ((g.b_row % 4) || (g.b_tap % 4)) return false;
  return true;
}

// N tile and K split.  Layers with few output positions give few CTAs (conv7 of unet3d: 5 M tiles), so the
// N tile shrinks and the (tap x channel block) loop is split over blockIdx.z when scratch is available.
static void plan_tc(int sm_count, const GatherGemm& g, int tiles_m, bool may_split, int* bn_out, int* ksplit_out,
                    int* its_out) {
  int bn = (g.N + 15) / 16 * 16;
  if (bn > 256) bn = 256;
  // the shuffle epilogue wants whole 32-channel runs; keep multiples of 32 when N allows
  while (bn > 64 && (int64_t)tiles_m * ((g.N + bn - 1) / bn) < sm_count) bn = (bn / 2 + 31) / 32 * 32;
  const int total = g.tz * g.tx * g.ty * ((g.K + BK - 1) / BK);
  // short K loops (upconv forward: K = F_in, one tap) are epilogue-dominated: with N tiles of 128 two CTAs fit an SM
  // and one's epilogue runs under the other's loads and MMAs
  static const int short_bn = getenv("E2_TC_SHORT_BN") ? atoi(getenv("E2_TC_SHORT_BN")) : 128;
  if (total <= 8 && bn > short_bn && short_bn >= 32) bn = short_bn;
  const int tiles_n = (g.N + bn - 1) / bn;
  int ksplit = 1;
  if (may_split) {
    const int64_t ctas = (int64_t)tiles_m * tiles_n;
    ksplit = (int)((2 * (int64_t)sm_count) / ctas);
    if (ksplit > total / 6) ksplit = total / 6;      // at least 6 K iterations (24 MMAs) per CTA
    if (ksplit < 1) ksplit = 1;
  }
  int its = (total + ksplit - 1) / ksplit;
  ksplit = (total + its - 1) / its;
  *bn_out = bn, *ksplit_out = ksplit, *its_out = its;
}

static int tc_tiles_m(const GatherGemm& g, int* tz, int* tx, int* ty) {
  pick_tile(g.Oz, g.Ox, g.Oy, g.sz, g.sx, g.sy, tz, tx, ty);
  return g.On * ((g.Oz + *tz - 1) / *tz) * ((g.Ox + *tx - 1) / *tx) * ((g.Oy + *ty - 1) / *ty);
}

size_t e2_gather_gemm_tc_workspace_bytes(int sm_count, const GatherGemm& g) {
  int tz, tx, ty, bn, ksplit, its;
  const int tiles_m = tc_tiles_m(g, &tz, &tx, &ty);
  plan_tc(sm_count, g, tiles_m, true, &bn, &ksplit, &its);
  if (ksplit <= 1) return 0;
  return (size_t)ksplit * tiles_m * ((g.N + bn - 1) / bn) * BM * bn * sizeof(float);
}

int e2_launch_gather_gemm_tc(e2_handle* h, const GatherGemm& g, cudaStream_t s) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return e2_fail(h, E2_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled entry point not available");
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.On = g.On, p.Oz = g.Oz, p.Ox = g.Ox, p.Oy = g.Oy;
  const int tiles_m = tc_tiles_m(g, &p.tz, &p.tx, &p.ty);
  p.sz = g.sz, p.sx = g.sx, p.sy = g.sy;
  p.gate = g.gate;
  p.ntz = (g.Oz + p.tz - 1) / p.tz, p.ntx = (g.Ox + p.tx - 1) / p.tx, p.nty = (g.Oy + p.ty - 1) / p.ty;
  p.kz = g.tz, p.kx = g.tx, p.ky = g.ty;
  p.oz = g.oz, p.ox = g.ox, p.oy = g.oy;
  p.K = g.K, p.N = g.N;
  int bn, ksplit, its;
  const bool have_ws = g.ws && !(reinterpret_cast<uintptr_t>(g.ws) & 15) && !getenv("E2_NO_SPLITK");
  plan_tc(h->sm_count, g, tiles_m, have_ws, &bn, &ksplit, &its);
  if (ksplit > 1 && (size_t)ksplit * tiles_m * ((g.N + bn - 1) / bn) * BM * bn * sizeof(float) > g.ws_bytes)
    plan_tc(h->sm_count, g, tiles_m, false, &bn, &ksplit, &its);
  p.BN = bn;
  p.its_per_split = its;
  p.part = ksplit > 1 ? static_cast<float*>(g.ws) : nullptr;
  const int b_stage = bn * BK * 4;
  p.stages = (bn <= 64) ? 4 : (bn <= 128 ? 3 : 4);
  int cols = 32;
  while (cols < bn) cols *= 2;
  p.tmem_cols = cols;
  p.C = g.C, p.c_pitch = g.c_pitch, p.bias = g.bias, p.act = g.act, p.accumulate = g.accumulate;
  p.round_tf32 = g.round_tf32;
  p.shuffle = g.shuffle, p.pz = g.pz, p.px = g.px, p.py = g.py, p.Fo = g.Fo;
  p.idesc = tc::make_idesc(2 /*TF32*/, 0, 0, BM, (uint32_t)bn);

  // A: activations (c, y, x, z, n), box (32, ty, tx, tz, 1), 128B swizzle, zero OOB fill
  CUtensorMap tmA, tmB;
  {
    cuuint64_t dims[5] = {(cuuint64_t)g.K, (cuuint64_t)g.Ay, (cuuint64_t)g.Ax, (cuuint64_t)g.Az, (cuuint64_t)g.An};
    cuuint64_t pitch = (cuuint64_t)g.a_pitch * 4;
    cuuint64_t strides[4] = {pitch, pitch * g.Ay, pitch * g.Ay * g.Ax, pitch * g.Ay * g.Ax * g.Az};
    // strided gather (upconv dgrad): traverse ty*sy elements, keep every sy-th -> ty rows
    cuuint32_t box[5] = {(cuuint32_t)BK, (cuuint32_t)(p.ty * g.sy), (cuuint32_t)(p.tx * g.sx), (cuuint32_t)(p.tz * g.sz), 1};
    cuuint32_t es[5] = {1, (cuuint32_t)g.sy, (cuuint32_t)g.sx, (cuuint32_t)g.sz, 1};
    CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(g.A), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return e2_fail(h, E2_ERR_CUDA, "cuTensorMapEncodeTiled(A) failed: %d", (int)r);
  }
  {
    const int T = g.tz * g.tx * g.ty;
    // B: packed weights [n][tap][k]: (k, tap, n) with byte strides (b_tap, b_row); b_tap == 0 when T == 1
    cuuint64_t dims[3] = {(cuuint64_t)g.K, (cuuint64_t)T, (cuuint64_t)g.N};
    cuuint64_t tap_stride = (cuuint64_t)(T > 1 ? g.b_tap : g.b_row) * 4;
    cuuint64_t strides[2] = {tap_stride, (cuuint64_t)g.b_row * 4};
    cuuint32_t box[3] = {(cuuint32_t)BK, 1, (cuuint32_t)bn};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(g.B), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return e2_fail(h, E2_ERR_CUDA, "cuTensorMapEncodeTiled(B) failed: %d", (int)r);
  }
  const size_t smem = 1024 + (size_t)p.stages * (A_STAGE_BYTES + b_stage) + (2 * p.stages + 1) * 8 + 16;
  static size_t configured = 0;
  if (smem > configured) {
    if (cudaFuncSetAttribute(k_gather_gemm_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)) !=
        cudaSuccess)
      return e2_fail(h, E2_ERR_CUDA, "cudaFuncSetAttribute(max dynamic smem) failed");
    configured = 227 * 1024;
  }
  const int tiles_n = (g.N + bn - 1) / bn;
  dim3 grid((unsigned)tiles_m, (unsigned)tiles_n, (unsigned)ksplit);
  k_gather_gemm_tc<<<grid, NUM_THREADS, smem, s>>>(tmA, tmB, p);
  e2_count_launch(h);
  E2_CUDA_CHECK(h, "gather_gemm_tc");
  if (p.part) {
    k_gather_gemm_reduce<<<(unsigned)(tiles_m * tiles_n * (bn / 4)), BM, 0, s>>>(p, ksplit, tiles_m, tiles_n);
    e2_count_launch(h);
    E2_CUDA_CHECK(h, "gather_gemm_reduce");
  }
  return E2_OK;
}

