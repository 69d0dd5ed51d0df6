// First-layer convolution (one input channel) on the tensor cores.
//
// With F_in == 1 the implicit GEMM has K = #taps (27 for 3x3x3), so the layer is HBM-bound (it writes F_out
// channels per voxel and reads one); the CUDA-core kernels in e2_conv_c1.cu nevertheless ran at the FFMA limit
// (27 FMA per output).  Here the threads only BUILD the im2col tile and the tensor core does the arithmetic:
//
//   tile   128 output positions = 8 x-lines x 16 y of one z plane; its input halo (kz x (8+kx-1) x (16+ky-1)
//          voxels) is staged in shared memory
//   A      thread r writes row r = the T tap values of position r (tf32-rounded, zero-padded to a multiple of 32)
//          as 16-byte chunks in the K-major 128B-swizzle layout (chunk ^ (row & 7)) -- what TMA would have produced
//   B      packed weights [N][taps] in the same layout, built once per CTA
//   MMA    tcgen05.mma kind::tf32 M128 x N x K8, ceil(T/8) of them per tile, accumulators double-buffered in TMEM so
//          the epilogue of tile i overlaps the build of tile i+1
//   out    tcgen05.ld -> +bias -> act -> tf32 round -> swizzled staging tile -> coalesced float4 stores
// Several CTAs share an SM (about 54 KB of shared memory each) and hide each other's load / build / store phases.
#include <algorithm>
#include "e2_common.cuh"
#include "e2_conv_internal.cuh"
#include "e2_tc_ptx.cuh"

namespace {

constexpr int C1_TXL = 8, C1_TYL = 16;     // tile: 8 x-lines of 16 y
constexpr int C1_THREADS = 128;
constexpr int C1_MAXT = 64;                // taps (two 128-byte K blocks)

struct C1TcParams {
  const float* x;
  int An, Az, Ax, Ay;
  int kz, kx, ky, oz, ox, oy;
  const float* B;
  int64_t b_row, b_tap;
  float* C;
  int c_pitch, N, NP, SP;       // channels, MMA N (multiple of 16), staging row pitch in floats (multiple of 32)
  int On, Oz, Ox, Oy;
  int ntx, nty, tiles;
  int T, KB;                    // taps, 32-tap K blocks
  int XH, YH, halo;             // halo extents / floats
  const float* bias;
  int act, round_tf32;
  int tmem_cols;
  int qw, qshift;               // copy-out: 16-byte chunks per row rounded up to a power of two, log2
  int stage_in_a;               // the staging tile aliases the A buffer of the tile being stored
  uint32_t idesc;
  E2FastDiv d_nty, d_ntx, d_oz;
  int off_b, off_halo, off_tap, off_hrel, off_hxyz, off_bias, off_stage, off_bar;   // bytes from the 1024-aligned base
};

// FAST: relu + tf32 rounding (every TF32-mode conv layer of the configs) without per-value branches
template <bool FAST>
__global__ void __launch_bounds__(C1_THREADS) k_c1_fwd_tc(const C1TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  // aligned base by OFFSET arithmetic on smem_raw (no integer round trip), so the pointer keeps its address space and the
  // warps' own accesses compile to LDS / STS instead of generic LD / ST
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smA = smem;                                               // [2][KB][128 rows][128 B]
  uint8_t* smB = smem + p.off_b;                                     // [KB][NP rows][128 B]
  float* halo = reinterpret_cast<float*>(smem + p.off_halo);         // [halo] + XH*YH zeros behind it for the padding taps
  int* tapofs = reinterpret_cast<int*>(smem + p.off_tap);            // [KB*32]: halo offset of tap kk
  int* hrel = reinterpret_cast<int*>(smem + p.off_hrel);             // [halo]: global offset relative to the tile origin
  int* hxyz = reinterpret_cast<int*>(smem + p.off_hxyz);             // [halo]: hz | hx << 8 | hy << 16
  float* bias_s = reinterpret_cast<float*>(smem + p.off_bias);       // [NP]
  uint64_t* acc_full = reinterpret_cast<uint64_t*>(smem + p.off_bar);   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 2);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int a_buf_bytes = p.KB * 16384;

  // ---- one-time setup: barriers, TMEM, tap / halo tables, bias, weights
  if (tid == 0) {
    tc::mbar_init(&acc_full[0], 1);
    tc::mbar_init(&acc_full[1], 1);
    tc::fence_barrier_init();
  }
  if (warp == 0) {
    tc::tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    tc::tmem_relinquish();
  }
  for (int kk = tid; kk < p.KB * 32; kk += C1_THREADS) {
    int o = p.halo;                                    // zero region behind the halo
    if (kk < p.T) {
      const int k = kk % p.ky, j = (kk / p.ky) % p.kx, i = kk / (p.ky * p.kx);
      o = (i * p.XH + j) * p.YH + k;
    }
    tapofs[kk] = o * 4;                                // byte offset
  }
  for (int i = tid; i < p.halo; i += C1_THREADS) {
    const int hy = i % p.YH, hx = (i / p.YH) % p.XH, hz = i / (p.YH * p.XH);
    hrel[i] = (hz * p.Ax + hx) * p.Ay + hy;
    hxyz[i] = hz | (hx << 8) | (hy << 16);
  }
  for (int i = tid; i < p.NP; i += C1_THREADS) bias_s[i] = (p.bias && i < p.N) ? __ldg(p.bias + i) : 0.f;
  for (int i = tid; i < p.XH * p.YH; i += C1_THREADS) halo[p.halo + i] = 0.f;   // padding taps read zeros (any row base)
  for (int i = tid; i < p.KB * p.NP * 32; i += C1_THREADS) {
    const int e = i & 31, n = (i >> 5) % p.NP, kb = i / (32 * p.NP);
    const int kk = kb * 32 + e;
    float v = 0.f;
    if (n < p.N && kk < p.T) v = __ldg(p.B + (int64_t)n * p.b_row + (int64_t)kk * p.b_tap);
    *reinterpret_cast<float*>(smB + (size_t)kb * p.NP * 128 + n * 128 + ((((e >> 2) ^ (n & 7)) << 4) | ((e & 3) << 2))) = v;
  }
  tc::fence_proxy_async();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int xl = tid / C1_TYL, yl = tid % C1_TYL;       // this thread's row of the tile
  const int r7 = tid & 7;
  const float* halo_row = halo + xl * p.YH + yl;
  // copy-out role: chunk cq of rows cr, cr + 128/qw, ...
  const int cq = tid & (p.qw - 1), cr0 = tid >> p.qshift, crstep = C1_THREADS >> p.qshift;

  auto tile_coords = [&](int t, int& in_, int& z, int& x0, int& y0) {
    const int t1 = (int)p.d_nty.div((uint32_t)t);
    const int ity = t - t1 * p.nty;
    const int t2 = (int)p.d_ntx.div((uint32_t)t1);
    const int itx = t1 - t2 * p.ntx;
    in_ = (int)p.d_oz.div((uint32_t)t2);
    z = t2 - in_ * p.Oz;
    x0 = itx * C1_TXL, y0 = ity * C1_TYL;
  };

  auto epilogue = [&](int in_, int z, int x0, int y0, int buf, uint32_t parity) {
    uint8_t* stage = p.stage_in_a ? smA + buf * a_buf_bytes : smem + p.off_stage;   // [128][SP] floats, chunks swizzled
    tc::mbar_wait(&acc_full[buf], parity);     // the tile's MMAs are complete: accumulators ready, A[buf] free
    tc::tc_fence_after();
    const uint32_t acc = tmem_base + (uint32_t)(buf * p.NP) + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < p.NP; c0 += 32) {
      uint32_t r[32];
      if (p.NP - c0 >= 32) {
        tc::tmem_ld_32x32b_x32(acc + (uint32_t)c0, r);
      } else {
        tc::tmem_ld_32x32b_x16(acc + (uint32_t)c0, r);
#pragma unroll
        for (int j = 16; j < 32; ++j) r[j] = 0u;
      }
      tc::tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c0 + q * 4 < p.NP) b4 = *reinterpret_cast<const float4*>(bias_s + c0 + q * 4);
        float v[4] = {__uint_as_float(r[q * 4]) + b4.x, __uint_as_float(r[q * 4 + 1]) + b4.y,
                      __uint_as_float(r[q * 4 + 2]) + b4.z, __uint_as_float(r[q * 4 + 3]) + b4.w};
        if (FAST || p.act == E2_ACT_RELU) {
          v[0] = fmaxf(v[0], 0.f), v[1] = fmaxf(v[1], 0.f), v[2] = fmaxf(v[2], 0.f), v[3] = fmaxf(v[3], 0.f);
        } else if (p.act != E2_ACT_LIN) {
          v[0] = e2_apply_act(v[0], p.act), v[1] = e2_apply_act(v[1], p.act), v[2] = e2_apply_act(v[2], p.act),
          v[3] = e2_apply_act(v[3], p.act);
        }
        if (FAST || p.round_tf32) v[0] = e2_round_tf32(v[0]), v[1] = e2_round_tf32(v[1]), v[2] = e2_round_tf32(v[2]), v[3] = e2_round_tf32(v[3]);
        const int chunk = (c0 >> 2) + q;       // 16-byte chunk of the row; the swizzle permutes inside 128-byte groups
        *reinterpret_cast<float4*>(stage + (size_t)tid * p.SP * 4 + (((chunk & ~7) | ((chunk & 7) ^ r7)) << 4)) =
            make_float4(v[0], v[1], v[2], v[3]);
      }
    }
    tc::tc_fence_before();
    __syncthreads();
    // coalesced copy-out: qw consecutive threads take the 16-byte chunks of a row, rows in position order
    if (cq < (p.N >> 2)) {
      float* out0 = p.C + ((((int64_t)in_ * p.Oz + z) * p.Ox + x0) * p.Oy + y0) * p.c_pitch + cq * 4;
      const bool full = x0 + C1_TXL <= p.Ox && y0 + C1_TYL <= p.Oy;
      const uint32_t st_s = tc::smem_u32(stage);
      const int sp4 = p.SP * 4, cq_hi = (cq & ~7) << 4, cq_lo = cq & 7;
#pragma unroll 4
      for (int row = cr0; row < 128; row += crstep) {
        const int rx = row >> 4, ry = row & 15;
        if (!full && (x0 + rx >= p.Ox || y0 + ry >= p.Oy)) continue;
        float4 v4;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(v4.x), "=f"(v4.y), "=f"(v4.z), "=f"(v4.w)
                     : "r"(st_s + (uint32_t)(row * sp4 + cq_hi + ((cq_lo ^ (row & 7)) << 4))));
        *reinterpret_cast<float4*>(out0 + (rx * p.Oy + ry) * p.c_pitch) = v4;
      }
    }
  };

  int it = 0;
  int pin = 0, pz = 0, px0 = 0, py0 = 0;
  bool have_prev = false;
  for (int t = blockIdx.x; t < p.tiles; t += gridDim.x, ++it) {
    const int buf = it & 1;
    int in_, z, x0, y0;
    tile_coords(t, in_, z, x0, y0);
    // ---- input halo -> shared memory (zero outside the tensor: only rows that are never stored read it)
    {
      const int gz0 = z + p.oz, gx0 = x0 + p.ox, gy0 = y0 + p.oy;
      const float* src = p.x + (((int64_t)in_ * p.Az + gz0) * p.Ax + gx0) * p.Ay + gy0;
      const bool interior = gz0 >= 0 && gz0 + p.kz <= p.Az && gx0 >= 0 && gx0 + p.XH <= p.Ax && gy0 >= 0 && gy0 + p.YH <= p.Ay;
      if (interior) {
        for (int i = tid; i < p.halo; i += C1_THREADS) halo[i] = e2_round_tf32(__ldg(src + hrel[i]));
      } else {
        for (int i = tid; i < p.halo; i += C1_THREADS) {
          const int c = hxyz[i];
          const int gz = gz0 + (c & 255), gx = gx0 + ((c >> 8) & 255), gy = gy0 + (c >> 16);
          float v = 0.f;
          if (gz >= 0 && gz < p.Az && gx >= 0 && gx < p.Ax && gy >= 0 && gy < p.Ay) v = __ldg(src + hrel[i]);
          halo[i] = e2_round_tf32(v);
        }
      }
    }
    __syncthreads();
    // ---- im2col row of this thread into A[buf]
    const uint32_t a_dst = tc::smem_u32(smA + buf * a_buf_bytes + tid * 128);
    const uint32_t hrow_s = tc::smem_u32(halo_row), tap_s = tc::smem_u32(tapofs);
    for (int kb = 0; kb < p.KB; ++kb) {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        int o0, o1, o2, o3;      // byte offsets of the 4 taps inside the halo
        asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(o0), "=r"(o1), "=r"(o2), "=r"(o3) : "r"(tap_s + (uint32_t)((kb * 32 + c * 4) * 4)));
        float v0, v1, v2, v3;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v0) : "r"(hrow_s + (uint32_t)o0));
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v1) : "r"(hrow_s + (uint32_t)o1));
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v2) : "r"(hrow_s + (uint32_t)o2));
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v3) : "r"(hrow_s + (uint32_t)o3));
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a_dst + (uint32_t)(kb * 16384 + ((c ^ r7) << 4))), "f"(v0), "f"(v1), "f"(v2), "f"(v3) : "memory");
      }
    }
    tc::fence_proxy_async();     // generic-proxy writes -> visible to the tensor core (async proxy)
    tc::tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc::tc_fence_after();
      const uint64_t tmpl = tc::make_smem_desc(0, 16, 1024, 2);
      const uint32_t a_addr = tc::smem_u32(smA + buf * a_buf_bytes), b_addr = tc::smem_u32(smB);
      const uint32_t acc = tmem_base + (uint32_t)(buf * p.NP);
      for (int kb = 0; kb < p.KB; ++kb) {
        const int nk = min(4, (p.T - kb * 32 + 7) >> 3);
        const uint64_t ad = tmpl + (uint64_t)((a_addr + (uint32_t)(kb * 16384)) >> 4);
        const uint64_t bd = tmpl + (uint64_t)((b_addr + (uint32_t)(kb * p.NP * 128)) >> 4);
        for (int k = 0; k < nk; ++k) tc::mma_tf32_ss(acc, ad + 2 * k, bd + 2 * k, p.idesc, (kb > 0 || k > 0) ? 1u : 0u);
      }
      tc::mma_commit(&acc_full[buf]);
    }
    // ---- epilogue of the previous tile while this tile's MMAs run
    if (have_prev) epilogue(pin, pz, px0, py0, buf ^ 1, (uint32_t)(((it - 1) >> 1) & 1));
    pin = in_, pz = z, px0 = x0, py0 = y0, have_prev = true;
  }
  if (have_prev) epilogue(pin, pz, px0, py0, (it - 1) & 1, (uint32_t)(((it - 1) >> 1) & 1));
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

bool fill_c1(const GatherGemm& g, C1TcParams* pp) {
  C1TcParams& p = *pp;
  memset(&p, 0, sizeof(p));
  p.x = g.A, p.An = g.An, p.Az = g.Az, p.Ax = g.Ax, p.Ay = g.Ay;
  p.kz = g.tz, p.kx = g.tx, p.ky = g.ty, p.oz = g.oz, p.ox = g.ox, p.oy = g.oy;
  p.B = g.B, p.b_row = g.b_row, p.b_tap = g.b_tap;
  p.C = g.C, p.c_pitch = g.c_pitch, p.N = g.N;
  p.NP = (g.N + 15) / 16 * 16;
  p.SP = (p.NP + 31) / 32 * 32;
  p.On = g.On, p.Oz = g.Oz, p.Ox = g.Ox, p.Oy = g.Oy;
  p.ntx = (g.Ox + C1_TXL - 1) / C1_TXL, p.nty = (g.Oy + C1_TYL - 1) / C1_TYL;
  const int64_t tiles = (int64_t)g.On * g.Oz * p.ntx * p.nty;
  if (tiles >= (1ll << 31)) return false;
  p.tiles = (int)tiles;
  p.T = g.tz * g.tx * g.ty;
  p.KB = (p.T + 31) / 32;
  p.XH = C1_TXL + g.tx - 1, p.YH = C1_TYL + g.ty - 1;
  p.halo = g.tz * p.XH * p.YH;
  p.bias = g.bias, p.act = g.act, p.round_tf32 = g.round_tf32;
  int cols = 32;
  while (cols < 2 * p.NP) cols *= 2;
  p.tmem_cols = cols;
  p.idesc = tc::make_idesc(2 /*TF32*/, 0, 0, 128, (uint32_t)p.NP);
  p.d_nty = e2_fastdiv((uint32_t)p.nty, (uint64_t)p.tiles);
  p.d_ntx = e2_fastdiv((uint32_t)p.ntx, (uint64_t)p.tiles);
  p.d_oz = e2_fastdiv((uint32_t)g.Oz, (uint64_t)p.tiles);
  p.qw = 1, p.qshift = 0;
  while (p.qw < p.N / 4) p.qw *= 2, p.qshift++;
  p.stage_in_a = (128 * p.SP * 4 <= p.KB * 16384) ? 1 : 0;
  if (g.tz > 255 || p.XH > 255 || p.YH > 255) return false;
  int off = 2 * p.KB * 16384;
  p.off_b = off, off += p.KB * p.NP * 128;
  p.off_stage = off, off += p.stage_in_a ? 0 : 128 * p.SP * 4;
  p.off_halo = off, off += ((p.halo + p.XH * p.YH) * 4 + 15) / 16 * 16;
  p.off_tap = off, off += p.KB * 32 * 4;
  p.off_hrel = off, off += (p.halo * 4 + 15) / 16 * 16;
  p.off_hxyz = off, off += (p.halo * 4 + 15) / 16 * 16;
  p.off_bias = off, off += p.NP * 4;
  p.off_bar = off, off += 32;
  return off + 1024 <= 200 * 1024;
}

}  // namespace

bool e2_conv_c1_fwd_tc_ok(const GatherGemm& g) {
  if (getenv("E2_C1_NOTC")) return false;
  if (g.K != 1 || g.a_pitch != 1 || g.shuffle || g.sz != 1 || g.sx != 1 || g.sy != 1) return false;
  if (g.gate || g.accumulate) return false;
  const int T = g.tz * g.tx * g.ty;
  // measured on B200: ahead of the CUDA-core kernel from ~16 taps on (27 taps: 136 vs 153 us, 36 taps: 87 vs 146 us),
  // behind it for 9 taps where 27 FMA per output were never the limit
  if (T < 16 || T > C1_MAXT) return false;
  if (g.N < 8 || g.N > 64 || (g.N & 3) || (g.c_pitch & 3) || (reinterpret_cast<uintptr_t>(g.C) & 15)) return false;
  if ((int64_t)g.Oz * g.Ox * g.Oy < 4096) return false;      // tiny layers: not worth a TMEM allocation
  C1TcParams p;
  return fill_c1(g, &p);
}

int e2_launch_conv_c1_fwd_tc(e2_handle* h, const GatherGemm& g, cudaStream_t s) {
  C1TcParams p;
  if (!fill_c1(g, &p)) return e2_fail(h, E2_ERR_UNSUPPORTED, "conv_c1_fwd_tc: problem does not qualify");
  const size_t smem = 1024 + (size_t)p.off_bar + 32;
  static size_t configured = 0;
  if (smem > configured) {
    if (cudaFuncSetAttribute(k_c1_fwd_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024)) != cudaSuccess ||
        cudaFuncSetAttribute(k_c1_fwd_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024)) != cudaSuccess)
      return e2_fail(h, E2_ERR_CUDA, "cudaFuncSetAttribute(max dynamic smem) failed");
    configured = 200 * 1024;
  }
  // resident CTAs per SM: shared memory and TMEM columns
  int per_sm = (int)std::min<size_t>((220 * 1024) / smem, (size_t)(512 / p.tmem_cols));
  per_sm = std::max(1, std::min(per_sm, 8));
  const int grid = std::min(p.tiles, h->sm_count * per_sm);
  if (p.act == E2_ACT_RELU && p.round_tf32)
    k_c1_fwd_tc<true><<<grid, C1_THREADS, smem, s>>>(p);
  else
    k_c1_fwd_tc<false><<<grid, C1_THREADS, smem, s>>>(p);
  e2_count_launch(h);
  E2_CUDA_CHECK(h, "conv_c1_fwd_tc");
  return E2_OK;
}
