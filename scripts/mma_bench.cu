// Micro-benchmark: issue rate of tcgen05.mma (SS mode) for the operand views the conv kernels use.
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../elektronn2_b200/csrc mma_bench.cu -o mma_bench
// Each CTA (one per SM) issues `reps` groups of 4 MMAs (one 128-byte K block) and reports cycles/MMA.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "e2_tc_ptx.cuh"

EncodeTiledFn e2_get_tmap_encode() { return nullptr; }

struct Cfg {
  int N;          // MMA N
  int fmt;        // 2 tf32, 1 bf16
  int sbo_a;      // bytes between 8-row groups of A
  int shift;      // 1: cycle the A start row through the 9 (j,k) taps of a YP=sbo_a/128 plane
  int nacc;       // accumulators rotated (each N columns)
  int a_bufs;     // distinct A buffers rotated
  int kmma;       // MMAs per 128-B K block (4)
};

__global__ void __launch_bounds__(128, 1) k_bench(Cfg c, int reps, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  float* f = reinterpret_cast<float*>(smem);
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) f[i] = (float)((i * 2654435761u) >> 20) * 1e-4f;
  if (threadIdx.x == 0) {
    tc::mbar_init(&bar, 1);
    tc::fence_barrier_init();
  }
  if (threadIdx.x < 32) {
    tc::tmem_alloc(&tmem_slot, 512);
    tc::tmem_relinquish();
  }
  tc::fence_proxy_async();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t base = tc::smem_u32(smem);
    const uint32_t a_region = 24 * 1024;            // per A buffer
    const uint32_t b_addr = base + 6 * a_region;    // B after 6 A buffers
    const uint64_t a_tmpl = tc::make_smem_desc(0, 16, (uint32_t)c.sbo_a, 2);
    const uint64_t b_tmpl = tc::make_smem_desc(b_addr, 16, 1024, 2);
    const uint32_t idesc = tc::make_idesc((uint32_t)c.fmt, 0, 0, 128, (uint32_t)c.N);
    const int YP = c.sbo_a / 128;
    long long t0 = clock64();
    int acc = 0, ab = 0, tap = 0;
    for (int r = 0; r < reps; ++r) {
      uint32_t row = 0;
      if (c.shift) row = (uint32_t)((tap / 3) * YP + (tap % 3));
      const uint64_t ad = a_tmpl + (uint64_t)((base + ab * a_region + row * 128) >> 4);
      const uint32_t d = tmem + (uint32_t)(acc * c.N);
#pragma unroll 4
      for (int k = 0; k < c.kmma; ++k) tc::mma_tf32_ss(d, ad + 2 * k, b_tmpl + 2 * k, idesc, 1u);
      if (++acc == c.nacc) acc = 0;
      if (++ab == c.a_bufs) ab = 0;
      if (++tap == 9) tap = 0;
    }
    tc::mma_commit(&bar);
    tc::mbar_wait(&bar, 0);
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  tc::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tc::tmem_dealloc(tmem, 512);
}


// Generic variant: KIND 0 = tf32, 1 = f16(bf16 inputs); TS = A operand from TMEM.
template <int KIND, int TS>
__global__ void __launch_bounds__(128, 1) k_bench2(int M, int N, int nacc, int reps, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  uint32_t* f = reinterpret_cast<uint32_t*>(smem);
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) f[i] = KIND ? 0x3c003c01u : 0x3f800000u + (i & 1023);
  if (threadIdx.x == 0) {
    tc::mbar_init(&bar, 1);
    tc::fence_barrier_init();
  }
  if (threadIdx.x < 32) {
    tc::tmem_alloc(&tmem_slot, 512);
    tc::tmem_relinquish();
  }
  tc::fence_proxy_async();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t base = tc::smem_u32(smem);
    const uint32_t b_addr = base + 64 * 1024;
    const uint64_t a_desc = tc::make_smem_desc(base, 16, 1024, 2);
    const uint64_t b_desc = tc::make_smem_desc(b_addr, 16, 1024, 2);
    const uint32_t idesc = tc::make_idesc(KIND ? 1u : 2u, 0, 0, (uint32_t)M, (uint32_t)N);
    const uint32_t a_tmem = tmem + 480;   // 32 columns of A at the end of TMEM (junk content)
    long long t0 = clock64();
    int acc = 0;
    for (int r = 0; r < reps; ++r) {
      const uint32_t d = tmem + (uint32_t)(acc * N);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (TS) {
          if (KIND)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
                         "r"(a_tmem + 8 * k), "l"(b_desc + 2 * k), "r"(idesc), "r"(1u) : "memory");
          else
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
                         "r"(a_tmem + 8 * k), "l"(b_desc + 2 * k), "r"(idesc), "r"(1u) : "memory");
        } else {
          if (KIND)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
                         "l"(a_desc + 2 * k), "l"(b_desc + 2 * k), "r"(idesc), "r"(1u) : "memory");
          else
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
                         "l"(a_desc + 2 * k), "l"(b_desc + 2 * k), "r"(idesc), "r"(1u) : "memory");
        }
      }
      if (++acc == nacc) acc = 0;
    }
    tc::mma_commit(&bar);
    tc::mbar_wait(&bar, 0);
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  tc::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tc::tmem_dealloc(tmem, 512);
}

template <int KIND, int TS>
static void run2(int M, int N, long long* d_out) {
  const int reps = 4096;
  int nacc = 448 / N;   // keep the last 32+ columns for the TMEM A operand
  if (nacc > 4) nacc = 4;
  if (nacc < 1) nacc = 1;
  cudaFuncSetAttribute(k_bench2<KIND, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  k_bench2<KIND, TS><<<148, 128, 220 * 1024>>>(M, N, nacc, reps, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("error %s\n", cudaGetErrorString(e));
    exit(1);
  }
  long long h[148];
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
  const double per = (double)mx / (reps * 4);
  const double ideal = (double)M * N / 256.0 * (KIND ? 1.0 : 1.0);   // cycles at 128xN/256 per 32-byte K slice
  printf("%s %s M=%3d N=%3d nacc=%d : %.1f cyc/MMA  (M*N/256 = %.0f) -> %.0f%%\n", KIND ? "bf16" : "tf32", TS ? "TS" : "SS", M, N,
         nacc, per, ideal, 100.0 * ideal / per);
}


// Issue-loop shapes.  MODE 0: single-lane branch (as the first plane kernel); MODE 1: whole warp runs the
// loop with uniform values, the MMA / commit are predicated on elect_one.
struct LoopP { int TZ, N, YP, ky, T9, kz, nslot, plane_stride, wslot, w_bytes; uint32_t idesc; };
template <int MODE>
__global__ void __launch_bounds__(128, 1) k_issue(LoopP p, int reps, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar, ready, sink;
  __shared__ uint32_t tmem_slot;
  uint32_t* f = reinterpret_cast<uint32_t*>(smem);
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) f[i] = 0x3f800000u + (i & 1023);
  if (threadIdx.x == 0) {
    tc::mbar_init(&bar, 1);
    tc::mbar_init(&ready, 1);
    tc::mbar_init(&sink, 1 << 20);
    tc::fence_barrier_init();
  }
  if (threadIdx.x < 32) {
    tc::tmem_alloc(&tmem_slot, 512);
    tc::tmem_relinquish();
  }
  tc::fence_proxy_async();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  if (warp == 0 && (MODE == 1 || threadIdx.x == 0)) {
    const uint32_t smP_addr = tc::smem_u32(smem), smW_addr = smP_addr + p.nslot * p.plane_stride;
    const uint64_t a_tmpl = tc::make_smem_desc(0, 16, (uint32_t)(p.YP * 128), 2);
    const uint64_t b_tmpl = tc::make_smem_desc(0, 16, 1024, 2);
    long long t0 = clock64();
    int ws = 0;
    for (int r = 0; r < reps; ++r) {
      for (int i = 0; i < p.kz; ++i) {
        int j = 0, k = 0;
        for (int jk = 0; jk < p.T9; ++jk) {
          tc::mbar_wait(&ready, 1);
          const uint64_t bd0 = b_tmpl + (uint64_t)((smW_addr + (uint32_t)(ws * p.w_bytes)) >> 4);
          const uint32_t row_enc = (uint32_t)((j * p.YP + k) * 8);
          for (int zl = 0; zl < p.TZ; ++zl) {
            int s = zl + i;
            if (s >= p.nslot) s -= p.nslot;
            tc::tc_fence_after();
            const uint64_t ad0 = a_tmpl + (uint64_t)(((smP_addr + (uint32_t)(s * p.plane_stride)) >> 4) + row_enc);
            const uint32_t acc = tmem + (uint32_t)(zl * p.N);
            if (MODE == 0 || tc::elect_one()) {
              tc::mma_tf32_ss(acc, ad0, bd0, p.idesc, 1u);
              tc::mma_tf32_ss(acc, ad0 + 2, bd0 + 2, p.idesc, 1u);
              tc::mma_tf32_ss(acc, ad0 + 4, bd0 + 4, p.idesc, 1u);
              tc::mma_tf32_ss(acc, ad0 + 6, bd0 + 6, p.idesc, 1u);
            }
            if (MODE == 1) __syncwarp();
          }
          if (MODE == 0 || tc::elect_one()) tc::mma_commit(&sink);
          if (MODE == 1) __syncwarp();
          if (++ws == p.wslot) ws = 0;
          if (++k == p.ky) k = 0, ++j;
        }
      }
    }
    if (MODE == 0 || tc::elect_one()) tc::mma_commit(&bar);
    if (MODE == 1) __syncwarp();
    tc::mbar_wait(&bar, 0);
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) out[blockIdx.x] = t1 - t0;
  }
  tc::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tc::tmem_dealloc(tmem, 512);
}

template <int MODE>
static void run3(int TZ, int N, long long* d_out) {
  LoopP p;
  p.TZ = TZ, p.N = N, p.YP = 10, p.ky = 3, p.T9 = 9, p.kz = 3, p.nslot = 6, p.plane_stride = 23 * 1024, p.wslot = 3;
  p.w_bytes = N * 128;
  p.idesc = tc::make_idesc(2, 0, 0, 128, (uint32_t)N);
  const int reps = 64;
  cudaFuncSetAttribute(k_issue<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  k_issue<MODE><<<148, 128, 220 * 1024>>>(p, reps, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("error %s\n", cudaGetErrorString(e));
    exit(1);
  }
  long long h[148];
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
  const double per = (double)mx / (reps * 27.0 * TZ * 4);
  printf("issue-loop mode %d TZ=%d N=%3d : %.1f cyc/MMA  (floor %d)\n", MODE, TZ, N, per, N < 96 ? 48 : N / 2);
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 148 * sizeof(long long));
  for (int N : {32, 64, 128}) {
    run3<0>(4, N, d_out);
    run3<1>(4, N, d_out);
  }
  run3<0>(2, 256, d_out);
  run3<1>(2, 256, d_out);
  return 0;
}
