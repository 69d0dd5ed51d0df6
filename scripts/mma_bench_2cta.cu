// Micro-benchmark + semantics check of tcgen05.mma.cta_group::2 (CTA pair, M = 256) for the operand views the conv
// kernels use.  Two questions: (1) which half of B does each CTA of the pair supply, and where do its
// columns land in D -- checked numerically against a host product; (2) what does an MMA cost per CTA when the pair
// shares B (operand fetch through each SM's 128 B/clk shared-memory port: 4 KB of A + N/2 x 32 B of B instead of N x 32 B).
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../elektronn2_b200/csrc mma_bench_2cta.cu -o mma_bench_2cta
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "e2_tc_ptx.cuh"

EncodeTiledFn e2_get_tmap_encode() { return nullptr; }

namespace t2 {
__device__ __forceinline__ uint32_t cta_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(dst_smem)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void mma_tf32_ss2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// arrive on the barrier at the same shared-memory offset in every CTA of `mask` once all MMAs issued so far are done
__device__ __forceinline__ void mma_commit2(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(tc::smem_u32(bar)), "h"(mask)
               : "memory");
}
}  // namespace t2

struct Cfg {
  int N;         // MMA N of the pair (each CTA holds N/2 rows of B)
  int mn;        // 0: K-major SW128 operands (z-stack conv), 1: MN-major 128B swizzle / 32B atom (wgrad)
  int pair;      // 1: cta_group::2 M = 256, 0: cta_group::1 M = 128 (both CTAs of the cluster run their own MMAs)
  int nacc;
  int check;     // 1: one K = 32 product, D dumped for the host check (K-major only)
};

// K-major SW128 tile: element (row r, k) of a [rows x 32 tf32] tile
__device__ __forceinline__ uint32_t sw128_off(int r, int k) { return (uint32_t)(r * 128 + ((((k >> 2) ^ (r & 7)) & 7) << 4) + (k & 3) * 4); }

template <int PAIR, int MN, int CHECK>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k_bench(Cfg c, int reps, long long* out, float* dump) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t rank = t2::cta_rank();
  uint8_t* smA = smem;                  // 64 KB: A tiles
  uint8_t* smB = smem + 64 * 1024;      // 64 KB: B tiles
  if (CHECK) {
    // A[m][k], m = rank * 128 + row; B[n][k], n = rank * N/2 + row (the hypothesis under test): small integers, exact in tf32
    for (int i = threadIdx.x; i < 128 * 32; i += blockDim.x) {
      const int r = i >> 5, k = i & 31, m = (int)rank * 128 + r;
      *reinterpret_cast<float*>(smA + sw128_off(r, k)) = (float)(((m * 7 + k * 3) % 11) - 5);
    }
    for (int i = threadIdx.x; i < (c.N / 2) * 32; i += blockDim.x) {
      const int r = i >> 5, k = i & 31, n = (int)rank * (c.N / 2) + r;
      *reinterpret_cast<float*>(smB + sw128_off(r, k)) = (float)(((n * 5 + k * 2) % 9) - 4);
    }
  } else {
    uint32_t* f = reinterpret_cast<uint32_t*>(smem);
    for (int i = threadIdx.x; i < 128 * 1024 / 4; i += blockDim.x) f[i] = 0x3f800000u + (i & 1023);
  }
  if (threadIdx.x == 0) {
    tc::mbar_init(&bar, 1);
    tc::fence_barrier_init();
  }
  if (threadIdx.x < 32) {
    if (PAIR) {
      t2::tmem_alloc2(&tmem_slot, 256);   // 256, not 512: were the two CTAs' allocations independent, both still fit
      t2::tmem_relinquish2();
    } else {
      tc::tmem_alloc(&tmem_slot, 256);
      tc::tmem_relinquish();
    }
  }
  tc::fence_proxy_async();
  tc::tc_fence_before();
  __syncthreads();
  t2::cluster_sync();                    // the peer's operands and barrier are in place before the leader issues
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t nloc = c.pair ? (uint32_t)c.N : (uint32_t)c.N;   // idesc N is the pair's N
  if (threadIdx.x == 0 && (rank == 0 || !PAIR)) {
    const uint32_t a_addr = tc::smem_u32(smA), b_addr = tc::smem_u32(smB);
    uint64_t a_desc, b_desc;
    uint32_t idesc;
    if (MN) {
      // wgrad views: A = 4 chunks of 32 channels one row apart (LBO 128), B = N/32 chunks of 8 KB
      a_desc = tc::make_smem_desc(a_addr, 128, 512, 1);
      b_desc = tc::make_smem_desc(b_addr, 8192, 512, 1);
      idesc = tc::make_idesc(2, 1, 1, PAIR ? 256u : 128u, nloc);
    } else {
      a_desc = tc::make_smem_desc(a_addr, 16, 1024, 2);
      b_desc = tc::make_smem_desc(b_addr, 16, 1024, 2);
      idesc = tc::make_idesc(2, 0, 0, PAIR ? 256u : 128u, nloc);
    }
    const long long t0 = clock64();
    int acc = 0;
    for (int r = 0; r < reps; ++r) {
      const uint32_t d = tmem + (uint32_t)(acc * c.N);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        // K-major: the next K = 8 slice is 32 B further in the row; MN-major: 8 rows = 1024 B further
        const uint64_t step = MN ? (uint64_t)(64 * k) : (uint64_t)(2 * k);
        const uint32_t accf = CHECK ? (k > 0 ? 1u : 0u) : 1u;
        if (PAIR)
          t2::mma_tf32_ss2(d, a_desc + step, b_desc + step, idesc, accf);
        else
          tc::mma_tf32_ss(d, a_desc + step, b_desc + step, idesc, accf);
      }
      if (++acc == c.nacc) acc = 0;
    }
    if (PAIR)
      t2::mma_commit2(&bar, 3);
    else
      tc::mma_commit(&bar);
    tc::mbar_wait(&bar, 0);
    const long long t1 = clock64();
    out[blockIdx.x] = CHECK ? (long long)tmem : t1 - t0;
  } else if (threadIdx.x == 0) {
    tc::mbar_wait(&bar, 0);              // follower: the leader's commit arrives here too
    out[blockIdx.x] = CHECK ? (long long)tmem : 0;   // check run: the follower reports its TMEM base address
  }
  __syncthreads();
  tc::tc_fence_after();
  if (CHECK && blockIdx.x < 2) {
    // every warp reads its 32 lanes x N columns
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int c0 = 0; c0 < c.N; c0 += 16) {
      uint32_t v[16];
      tc::tmem_ld_32x32b_x16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
      tc::tmem_ld_wait();
      for (int j = 0; j < 16; ++j) dump[((size_t)rank * 128 + warp * 32 + lane) * c.N + c0 + j] = __uint_as_float(v[j]);
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  t2::cluster_sync();
  if (threadIdx.x < 32) {
    tc::tc_fence_after();
    if (PAIR)
      t2::tmem_dealloc2(tmem, 256);
    else
      tc::tmem_dealloc(tmem, 256);
  }
}

static int launch(const Cfg& c, int reps, int grid, long long* d_out, float* d_dump, long long* h) {
  void (*fn)(Cfg, int, long long*, float*) = c.check ? k_bench<1, 0, 1>
                                             : c.pair ? (c.mn ? k_bench<1, 1, 0> : k_bench<1, 0, 0>)
                                                      : (c.mn ? k_bench<0, 1, 0> : k_bench<0, 0, 0>);
  cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  fn<<<grid, 128, 200 * 1024>>>(c, reps, d_out, d_dump);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("error %s (N %d mn %d pair %d)\n", cudaGetErrorString(e), c.N, c.mn, c.pair);
    return 1;
  }
  cudaMemcpy(h, d_out, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
  return 0;
}

int main() {
  long long* d_out;
  float* d_dump;
  cudaMalloc(&d_out, 148 * sizeof(long long));
  cudaMalloc(&d_dump, 2 * 128 * 256 * sizeof(float));
  long long h[148];

  // ---- (1) semantics: D = A B^T with M = 256 over the pair, N = 64 and 192
  for (int N : {64, 192}) {
    Cfg c = {N, 0, 1, 1, 1};
    cudaMemset(d_dump, 0, 2 * 128 * 256 * sizeof(float));
    if (launch(c, 1, 2, d_out, d_dump, h)) return 1;
    printf("TMEM base addresses: leader 0x%llx follower 0x%llx\n", h[0], h[1]);
    std::vector<float> D(2 * 128 * N);
    cudaMemcpy(D.data(), d_dump, D.size() * sizeof(float), cudaMemcpyDeviceToHost);
    auto A = [](int m, int k) { return (float)(((m * 7 + k * 3) % 11) - 5); };
    auto B = [](int n, int k) { return (float)(((n * 5 + k * 2) % 9) - 4); };
    long bad = 0;
    for (int m = 0; m < 256; ++m)
      for (int n = 0; n < N; ++n) {
        float ref = 0.f;
        for (int k = 0; k < 32; ++k) ref += A(m, k) * B(n, k);
        if (D[(size_t)m * N + n] != ref) {
          if (bad < 6) printf("  mismatch m %d n %d: got %g want %g\n", m, n, D[(size_t)m * N + n], ref);
          ++bad;
        }
      }
    printf("check N=%d: D[rank*128 + lane][n] = sum_k A[rank*128 + lane][k] * B[n][k], B rows n < N/2 from CTA 0, the rest from CTA 1: %s (%ld mismatches)\n",
           N, bad ? "NO" : "yes", bad);
  }

  // ---- (2) cycles per MMA per CTA
  const int reps = 4096;
  for (int mn = 0; mn < 2; ++mn)
    for (int N : {64, 96, 128, 192, 256}) {
      if (mn && N > 128) continue;
      for (int pair = 0; pair < 2; ++pair) {
        int nacc = 256 / N;
        if (nacc < 1) nacc = 1;
        Cfg c = {N, mn, pair, nacc, 0};
        if (launch(c, reps, 148, d_out, d_dump, h)) return 1;
        long long mx = 0;
        for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
        const double per = (double)mx / (reps * 4);
        printf("%s tf32 %s N=%3d: %6.1f cyc/MMA per CTA  (math N/2 = %d; fetch model %s = %d)\n", mn ? "MN-major" : "K-major ",
               pair ? "cta_group::2 M=256" : "cta_group::1 M=128", N, per, N / 2, pair ? "32 + N/8" : "32 + N/4",
               pair ? 32 + N / 8 : 32 + N / 4);
      }
    }
  return 0;
}
