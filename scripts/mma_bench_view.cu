// Micro-benchmark: tcgen05.mma kind::tf32 K-major rate for the shifted / strided A views of the halo-plane
// conv kernels, issued from a fully unrolled single-thread loop (issue cost ~3 instructions per MMA).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../elektronn2_b200/csrc mma_bench_view.cu -o mma_bench_view
#include <cstdio>
#include <cstdlib>
#include "e2_tc_ptx.cuh"

EncodeTiledFn e2_get_tmap_encode() { return nullptr; }

// MODE 0: A dense (SBO 1024), aligned      1: SBO = 1280 (YP = 10 rows), aligned start
//      2: SBO 1280, start row shifted by 1  3: SBO 1280, start row cycles through the 9 (j,k) taps
//      4: dense SBO 1024 but start row shifted by 1
template <int N, int MODE, int NACC>
__global__ void __launch_bounds__(128, 1) k_bench(int reps, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  uint32_t* f = reinterpret_cast<uint32_t*>(smem);
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) f[i] = 0x3f800000u + (i & 1023);
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
    const uint32_t sbo = (MODE == 0 || MODE == 4) ? 1024u : 1280u;
    const uint64_t a_tmpl = tc::make_smem_desc(0, 16, sbo, 2);
    const uint64_t b_desc = tc::make_smem_desc(base + 128 * 1024, 16, 1024, 2);
    const uint32_t idesc = tc::make_idesc(2, 0, 0, 128, (uint32_t)N);
    const uint32_t plane = 24 * 1024;
    unsigned long long g0, g1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int g = 0; g < 9; ++g) {          // 9 "taps"
        uint32_t row = 0;
        if (MODE == 2 || MODE == 4) row = 1;
        if (MODE == 3) row = (uint32_t)((g / 3) * 10 + (g % 3));
#pragma unroll
        for (int a = 0; a < NACC; ++a) {     // planes -> accumulators
          const uint64_t ad = a_tmpl + (uint64_t)((base + (uint32_t)a * plane + row * 128u) >> 4);
          const uint32_t d = tmem + (uint32_t)(a * N);
          tc::mma_tf32_ss(d, ad, b_desc, idesc, 1u);
          tc::mma_tf32_ss(d, ad + 2, b_desc + 2, idesc, 1u);
          tc::mma_tf32_ss(d, ad + 4, b_desc + 4, idesc, 1u);
          tc::mma_tf32_ss(d, ad + 6, b_desc + 6, idesc, 1u);
        }
      }
    }
    tc::mma_commit(&bar);
    tc::mbar_wait(&bar, 0);
    long long t1 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    out[blockIdx.x] = t1 - t0;
    out[148 + blockIdx.x] = (long long)(g1 - g0);
  }
  tc::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tc::tmem_dealloc(tmem, 512);
}

template <int N, int MODE, int NACC>
static void run(long long* d_out, int reps = 200) {
  cudaFuncSetAttribute(k_bench<N, MODE, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  k_bench<N, MODE, NACC><<<148, 128, 220 * 1024>>>(reps, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("error %s\n", cudaGetErrorString(e));
    exit(1);
  }
  long long h[296];
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0, ns = 0;
  for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx, ns = h[148 + i] > ns ? h[148 + i] : ns;
  static const char* names[] = {"dense SBO1024 aligned", "SBO1280 aligned", "SBO1280 row+1", "SBO1280 tap rows", "dense SBO1024 row+1"};
  printf("N=%3d nacc=%d %-22s : %6.1f cyc/MMA (N/2 = %d)\n", N, NACC, names[MODE], (double)mx / ((double)reps * 9 * NACC * 4), N / 2);
  if (reps > 1000) printf("    sustained: %.2f ms, SM clock %.0f MHz, %.0f TF/s\n", ns * 1e-6, (double)mx / ns * 1e3,
                          2.0 * 128 * N * 8 * (double)reps * 9 * NACC * 4 * 148 / ns / 1e3);
}

template <int N, int NACC>
static void run_all(long long* d_out) {
  run<N, 0, NACC>(d_out);
  run<N, 1, NACC>(d_out);
  run<N, 2, NACC>(d_out);
  run<N, 3, NACC>(d_out);
  run<N, 4, NACC>(d_out);
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 296 * sizeof(long long));
  run_all<64, 4>(d_out);
  run_all<96, 4>(d_out);
  run_all<128, 4>(d_out);
  run_all<192, 2>(d_out);
  run_all<256, 2>(d_out);
  // sustained tensor load: what SM clock does the chip hold?
  run<256, 0, 2>(d_out, 3000);
  run<256, 0, 2>(d_out, 30000);
  run<192, 3, 2>(d_out, 30000);
  run<64, 3, 4>(d_out, 30000);
  return 0;
}
