// Micro-benchmark: tcgen05.mma kind::tf32 rate with MN-major operands (the wgrad kernels' layout).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../elektronn2_b200/csrc mma_bench_mn.cu -o mma_bench_mn
#include <cstdio>
#include <cstdlib>
#include "e2_tc_ptx.cuh"

EncodeTiledFn e2_get_tmap_encode() { return nullptr; }

struct Cfg {
  int a_mn, b_mn;       // 1: MN-major (layout type 1, 128B swizzle / 32B atom), 0: K-major SW128
  int a_lbo, b_lbo;     // bytes (MN-major: 32-channel chunk stride)
  int N, nacc;
  int a_step, b_step;   // descriptor increment per MMA (bytes)
  int nbuf;             // distinct K steps cycled
  int a_row_shift;      // extra start-row shift cycled 0..2 (x128 B) per accumulator (tap views)
};

__global__ void __launch_bounds__(128, 1) k_bench(Cfg c, int reps, long long* out) {
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
    const uint32_t b_addr = base + 96 * 1024;
    const uint64_t a_tmpl = c.a_mn ? tc::make_smem_desc(base, (uint32_t)c.a_lbo, 512, 1) : tc::make_smem_desc(base, 16, 1024, 2);
    const uint64_t b_tmpl = c.b_mn ? tc::make_smem_desc(b_addr, (uint32_t)c.b_lbo, 512, 1) : tc::make_smem_desc(b_addr, 16, 1024, 2);
    const uint32_t idesc = tc::make_idesc(2, (uint32_t)c.a_mn, (uint32_t)c.b_mn, 128, (uint32_t)c.N);
    long long t0 = clock64();
    int buf = 0;
    for (int r = 0; r < reps; ++r) {
      const uint64_t ad0 = a_tmpl + (uint64_t)((buf * c.a_step) >> 4);
      const uint64_t bd = b_tmpl + (uint64_t)((buf * c.b_step) >> 4);
#pragma unroll 1
      for (int a = 0; a < c.nacc; ++a)
        tc::mma_tf32_ss(tmem + (uint32_t)(a * c.N), ad0 + (uint64_t)(a * c.a_row_shift * 8), bd, idesc, 1u);
      if (++buf == c.nbuf) buf = 0;
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

static void run(const char* name, Cfg c, long long* d_out) {
  const int reps = 4096;
  cudaFuncSetAttribute(k_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  k_bench<<<148, 128, 220 * 1024>>>(c, reps, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("%s: error %s\n", name, cudaGetErrorString(e));
    exit(1);
  }
  long long h[148];
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("%-46s N=%3d : %6.1f cyc/MMA (ideal %d)\n", name, c.N, (double)mx / ((double)reps * c.nacc), c.N / 2);
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 148 * sizeof(long long));
  for (int N : {64, 128, 256}) {
    const int nacc = N == 256 ? 2 : 3;
    // K-major both (reference point): one MMA = 32 B of each 128-B row; step 32 B, 4 steps then next buffer
    run("A K-major, B K-major", Cfg{0, 0, 0, 0, N, nacc, 32, 32, 4, 0}, d_out);
    run("A MN (chunks 8 KB apart), B MN (8 KB)", Cfg{1, 1, 8192, 8192, N, nacc, 1024, 1024, 8, 0}, d_out);
    run("A MN (LBO 128: stacked y-taps), B MN (8 KB)", Cfg{1, 1, 128, 8192, N, nacc, 1280, 1024, 8, 0}, d_out);
    run("A MN (LBO 128, tap row shifts), B MN (8 KB)", Cfg{1, 1, 128, 8192, N, nacc, 1280, 1024, 8, 10}, d_out);
    run("A MN (8 KB), B K-major", Cfg{1, 0, 8192, 0, N, nacc, 1024, 32, 4, 0}, d_out);
    run("A K-major, B MN (8 KB)", Cfg{0, 1, 0, 8192, N, nacc, 32, 1024, 4, 0}, d_out);
  }
  return 0;
}
