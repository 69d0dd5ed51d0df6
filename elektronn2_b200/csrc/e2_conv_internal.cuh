// Problem descriptions shared by the CUDA-core and tcgen05 conv kernels (internal).
#pragma once
#include "e2_common.cuh"

// C[m, n] = sum_{tap,k} A[pos(m)*s + tap + org][k] * B[n*b_row + tap*b_tap + k]
struct GatherGemm {
  const float* A;
  int a_pitch, K;
  int An, Az, Ax, Ay;
  const float* B;
  int64_t b_row, b_tap;
  float* C;
  int c_pitch, N;
  int On, Oz, Ox, Oy;  // output position grid (GEMM M)
  int tz, tx, ty;      // taps
  int oz, ox, oy;      // origin
  int sz, sx, sy;      // position stride
  const float* bias;
  int act, accumulate, round_tf32;
  int shuffle, pz, px, py, Fo;  // pixel-shuffle epilogue (upconv fwd)
  const float* gate;            // fused ReLU backward: zero the result where gate <= 0 (same layout as C)
  void* ws;                     // optional caller-owned scratch (split-K partial tiles)
  size_t ws_bytes;
  // max-pool fused into the z-stack kernel's epilogue (e2_conv3d_fwd_pool).  C may be null then (the caller does not
  // want the unpooled tensor).  Cp / Ci: pooled values and int32 argmax, (On, Oz/qz, Ox/qx, Oy/qy, N) at pitch cp_pitch;
  // the pooled values go through +pbias -> pact -> tf32 round (pround) after the maximum.
  int fuse_pool, qz, qx, qy;
  float* Cp;
  int32_t* Ci;
  int cp_pitch;
  const float* pbias;
  int pact, pround;
  int keep[6];   // window {z0,z1,x0,x1,y0,y1} of C the caller needs (tiles outside it are not stored)
};

// W[r][tap][s] = sum_m P[m][r] * Q[pos(m)*st + tap + org][s]
struct ReduceGemm {
  const float* P;
  int p_pitch, R;
  int Mn, Mz, Mx, My;
  const float* Q;
  int q_pitch, S;
  int Qn, Qz, Qx, Qy;
  int tz, tx, ty, oz, ox, oy, sz, sx, sy;
  float* W;
  int out_mode;  // 0: dw[r][s][flip(tap)] (conv)   1: dw[s][r][tap] (upconv)
};

int e2_launch_gather_gemm_ffma(e2_handle* h, const GatherGemm& g, cudaStream_t s);
int e2_launch_reduce_gemm_ffma(e2_handle* h, const ReduceGemm& g, cudaStream_t s);
int e2_launch_conv_c1_fwd(e2_handle* h, const GatherGemm& g, cudaStream_t s);
int e2_launch_conv_c1_wgrad(e2_handle* h, const ReduceGemm& g, cudaStream_t s);
// line-staged first-layer kernels (e2_conv_c1.cu); preferred when they qualify
bool e2_conv_c1_fwd_line_ok(const GatherGemm& g);
int e2_launch_conv_c1_fwd_line(e2_handle* h, const GatherGemm& g, cudaStream_t s);
bool e2_conv_c1_wgrad_line_ok(const ReduceGemm& g);
int e2_launch_conv_c1_wgrad_line(e2_handle* h, const ReduceGemm& g, cudaStream_t s);
// register-tiled first-layer kernels (compile-time taps); wgrad fuses the bias gradient
bool e2_conv_c1_fwd_reg_ok(const GatherGemm& g);
int e2_launch_conv_c1_fwd_reg(e2_handle* h, const GatherGemm& g, cudaStream_t s);
bool e2_conv_c1_wgrad_reg_ok(const ReduceGemm& g);
int e2_launch_conv_c1_wgrad_reg(e2_handle* h, const ReduceGemm& g, float* db, cudaStream_t s);
int e2_launch_bias_grad(e2_handle* h, const float* dy, int64_t M, int C, int pitch, float* db, cudaStream_t s);

// first layer on the tensor cores: thread-built im2col tile + tcgen05 (e2_conv_c1_tc.cu), TF32 mode
bool e2_conv_c1_fwd_tc_ok(const GatherGemm& g);
int e2_launch_conv_c1_fwd_tc(e2_handle* h, const GatherGemm& g, cudaStream_t s);
// 1x1x1 convolutions with <= 4 output channels (e2_conv_pw.cu): streaming fp32 kernels; wgrad fuses the bias gradient
bool e2_conv_pw_fwd_ok(const GatherGemm& g);
int e2_launch_conv_pw_fwd(e2_handle* h, const GatherGemm& g, cudaStream_t s);
bool e2_conv_pw_dgrad_ok(const GatherGemm& g);
int e2_launch_conv_pw_dgrad(e2_handle* h, const GatherGemm& g, cudaStream_t s);
bool e2_conv_pw_wgrad_ok(const ReduceGemm& g);
int e2_launch_conv_pw_wgrad(e2_handle* h, const ReduceGemm& g, float* db, cudaStream_t s);

// tcgen05 path (e2_conv_tc.cu)
bool e2_gather_gemm_tc_ok(const e2_handle* h, const GatherGemm& g);
int e2_launch_gather_gemm_tc(e2_handle* h, const GatherGemm& g, cudaStream_t s);
// scratch bytes the tap kernel wants for its split-K path (0: it will not split)
size_t e2_gather_gemm_tc_workspace_bytes(int sm_count, const GatherGemm& g);
// halo planes + z-taps stacked along N (e2_conv_zstack_tc.cu); preferred over both when it qualifies
bool e2_conv_zstack_tc_ok(const e2_handle* h, const GatherGemm& g);
// would the z-stack kernel take g with the max-pool (g.fuse_pool, g.qz/qx/qy) in its epilogue -- and without losing a
// K split it would otherwise use?
bool e2_conv_zstack_pool_ok(const e2_handle* h, const GatherGemm& g);
int e2_launch_conv_zstack_tc(e2_handle* h, const GatherGemm& g, cudaStream_t s);
// scratch bytes its split-K plan wants (0: no split); g.ws / g.ws_bytes carry the scratch at launch
size_t e2_conv_zstack_workspace_bytes(int sm_count, const GatherGemm& g);
// picks zstack / tap kernel / CUDA cores for a TF32 request
int e2_dispatch_gather_gemm(e2_handle* h, const GatherGemm& g, int compute, cudaStream_t s);
bool e2_reduce_gemm_tc_ok(const e2_handle* h, const ReduceGemm& g);
// ws (nullable): per-CTA partial tiles + deterministic reduce kernel; without it fp32 atomics
int e2_launch_reduce_gemm_tc(e2_handle* h, const ReduceGemm& g, void* ws, size_t ws_bytes, cudaStream_t s);
size_t e2_reduce_gemm_tc_workspace_bytes(int sm_count, const ReduceGemm& g);
// halo-reuse wgrad (e2_wgrad_halo_tc.cu): x tile loaded once, taps are shifted descriptor views
bool e2_wgrad_halo_tc_ok(const e2_handle* h, const ReduceGemm& g);
// db (nullable): bias gradient = column sums of P; *db_done tells the caller whether the kernel produced it
int e2_launch_wgrad_halo_tc(e2_handle* h, const ReduceGemm& g, void* ws, size_t ws_bytes, float* db, bool* db_done,
                            cudaStream_t s);
size_t e2_wgrad_halo_workspace_bytes(int sm_count, const ReduceGemm& g);

// z-taps stacked along MMA N (e2_wgrad_zs_tc.cu): for few dy channels (R <= 64), where the halo kernel's N = R MMA
// is operand-fetch-bound.  Needs the workspace (partial tiles + deterministic reduce).
bool e2_wgrad_zs_tc_ok(const e2_handle* h, const ReduceGemm& g);
int e2_launch_wgrad_zs_tc(e2_handle* h, const ReduceGemm& g, void* ws, size_t ws_bytes, float* db, bool* db_done,
                          cudaStream_t s);
size_t e2_wgrad_zs_workspace_bytes(int sm_count, const ReduceGemm& g);

// first-layer wgrad on the tensor cores (e2_wgrad_c1_tc.cu): thread-built im2col rows + TMA-fed dy, bias gradient as an
// extra im2col column; TF32 mode, <= 31 taps, <= 32 output channels
bool e2_wgrad_c1_tc_ok(const ReduceGemm& g);
int e2_launch_wgrad_c1_tc(e2_handle* h, const ReduceGemm& g, float* db, cudaStream_t s);

// first-layer forward, warp-specialised (e2_conv_c1_ws.cu): TMA halo -> builder warps -> MMA -> epilogue warps -> TMA store
bool e2_conv_c1_fwd_ws_ok(const GatherGemm& g);
int e2_launch_conv_c1_fwd_ws(e2_handle* h, const GatherGemm& g, cudaStream_t s);
