// Bandwidth-bound kernels: 3-D max-pool (+argmax, optional fused +bias->act), its
// backward, max-fragment-pooling fwd/bwd, fragments->dense fwd/bwd, crop+concat fwd/bwd.
// All tensors are channels-last; one thread owns V consecutive channels of one output
// position, so a warp reads/writes 32*V contiguous floats (V=4: 512 B) per window tap.
#include <initializer_list>
#include <stdlib.h>
#include "e2_common.cuh"

template <int V>
struct Vec;
template <>
struct Vec<1> {
  float v[1];
  __device__ __forceinline__ void load(const float* p) { v[0] = __ldg(p); }
  __device__ __forceinline__ void store(float* p) const { p[0] = v[0]; }
};
template <>
struct Vec<4> {
  float v[4];
  __device__ __forceinline__ void load(const float* p) {
    float4 t = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = t.x, v[1] = t.y, v[2] = t.z, v[3] = t.w;
  }
  __device__ __forceinline__ void store(float* p) const {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <int V>
struct IVec;
template <>
struct IVec<1> {
  int v[1];
  __device__ __forceinline__ void load(const int* p) { v[0] = __ldg(p); }
  __device__ __forceinline__ void store(int* p) const { p[0] = v[0]; }
};
template <>
struct IVec<4> {
  int v[4];
  __device__ __forceinline__ void load(const int* p) {
    int4 t = __ldg(reinterpret_cast<const int4*>(p));
    v[0] = t.x, v[1] = t.y, v[2] = t.z, v[3] = t.w;
  }
  __device__ __forceinline__ void store(int* p) const { *reinterpret_cast<int4*>(p) = make_int4(v[0], v[1], v[2], v[3]); }
};

static inline bool vec4_ok(std::initializer_list<const void*> ptrs, std::initializer_list<int> counts) {
  for (const void* p : ptrs)
    if (p && (reinterpret_cast<uintptr_t>(p) & 15)) return false;
  for (int c : counts)
    if (c & 3) return false;
  return true;
}

// float4 paths also serve channel counts that are not a multiple of 4 when the lanes up to the next multiple are real
// padding of the tensor (pitch - C < 4, never the neighbour slice of a concat buffer): pad lanes are read (finite:
// buffers start zeroed and pads only ever receive zeros) and written as zeros.
static inline bool pad4_ok(int C, int pitch) { return (pitch & 3) == 0 && pitch - C < 4; }

struct PoolP {
  int n, Z, X, Y, C, xp;  // input dims, input pitch
  int Zo, Xo, Yo, yp;     // output dims, output pitch
  int pz, px, py;
  int act, has_bias, tie, accumulate, round_tf32, gate_pooled;
};

// ------------------------------------------------------------------ max-pool forward
// One block per OUTPUT row (n, zo, xo): the row's Yo*C outputs and each of its window taps are contiguous in
// HBM, the block decomposes its row index once, and a thread needs a single multiply-high to split its item
// into (yo, channel group) -- the 64-bit div/mod chain of a flat index cost more than the memory traffic.
// PZ/PX/PY > 0: compile-time window (all loads of a window are issued before the compares); 0: run-time window
template <int V, int PZ, int PX, int PY>
__global__ void __launch_bounds__(256) k_maxpool_fwd(PoolP p, E2FastDiv dcv, const float* __restrict__ x,
                                                     const float* __restrict__ bias, float* __restrict__ y,
                                                     int* __restrict__ amax) {
  const int cv = (p.C + V - 1) / V;
  int r = blockIdx.x;
  const int xo = r % p.Xo;
  r /= p.Xo;
  const int zo = r % p.Zo;
  const int n = r / p.Zo;
  const int rowlen = p.Yo * cv;
  const float* xn = x + (int64_t)n * p.Z * p.X * p.Y * p.xp;
  const int64_t orow = (((int64_t)n * p.Zo + zo) * p.Xo + xo) * p.Yo;
  for (int t = threadIdx.x; t < rowlen; t += blockDim.x) {
    const int yo = (int)dcv.div((uint32_t)t);
    const int c = (t - yo * cv) * V;
    float best[V];
    int bi[V];
#pragma unroll
    for (int j = 0; j < V; ++j) best[j] = -INFINITY, bi[j] = 0;
    if (PZ > 0) {
      Vec<V> w[PZ > 0 ? PZ * PX * PY : 1];
#pragma unroll
      for (int dz = 0; dz < PZ; ++dz)
#pragma unroll
        for (int dx = 0; dx < PX; ++dx)
#pragma unroll
          for (int dy = 0; dy < PY; ++dy) {
            const int lin = ((zo * PZ + dz) * p.X + xo * PX + dx) * p.Y + yo * PY + dy;
            w[(dz * PX + dx) * PY + dy].load(xn + (int64_t)lin * p.xp + c);
          }
#pragma unroll
      for (int dz = 0; dz < PZ; ++dz)
#pragma unroll
        for (int dx = 0; dx < PX; ++dx)
#pragma unroll
          for (int dy = 0; dy < PY; ++dy) {
            const int lin = ((zo * PZ + dz) * p.X + xo * PX + dx) * p.Y + yo * PY + dy;
            const int wi = (dz * PX + dx) * PY + dy;
#pragma unroll
            for (int j = 0; j < V; ++j)
              if (wi == 0 || w[wi].v[j] > best[j]) best[j] = w[wi].v[j], bi[j] = lin;   // strict '>': FIRST maximum
          }
    } else {
      bool first = true;
      for (int dz = 0; dz < p.pz; ++dz)
        for (int dx = 0; dx < p.px; ++dx) {
          const int lin0 = ((zo * p.pz + dz) * p.X + xo * p.px + dx) * p.Y + yo * p.py;
          const float* src = xn + (int64_t)lin0 * p.xp + c;
#pragma unroll 2
          for (int dy = 0; dy < p.py; ++dy) {
            Vec<V> v;
            v.load(src + (int64_t)dy * p.xp);
#pragma unroll
            for (int j = 0; j < V; ++j) {
              // strict '>' keeps the FIRST maximum in (z,x,y) scan order
              if (first || v.v[j] > best[j]) best[j] = v.v[j], bi[j] = lin0 + dy;
            }
            first = false;
          }
        }
    }
    Vec<V> o;
    IVec<V> oi;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      float r2 = best[j];
      if (c + j < p.C) {
        if (p.has_bias) r2 += __ldg(bias + c + j);  // reference order: pool -> +bias -> act (neural.py:678,711-712)
        r2 = e2_apply_act(r2, p.act);
        o.v[j] = p.round_tf32 ? e2_round_tf32(r2) : r2;
      } else {
        o.v[j] = 0.f;                               // pad lane
      }
      oi.v[j] = bi[j];
    }
    const int64_t oofs = (orow + yo) * p.yp + c;
    o.store(y + oofs);
    if (amax) oi.store(amax + oofs);
  }
}

// ----------------------------------------------------------------- max-pool backward
// Windows do not overlap and tile the input, so each dx element is written exactly once.
template <int V>
__global__ void __launch_bounds__(256) k_maxpool_bwd(PoolP p, const float* __restrict__ dy, const int* __restrict__ amax,
                                                     const float* __restrict__ x, float* __restrict__ dx,
                                                     const float* __restrict__ gate) {
  const int cv = p.C / V;
  const int64_t total = (int64_t)p.n * p.Zo * p.Xo * p.Yo * cv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % cv) * V;
    int64_t pos = i / cv;
    int yo = (int)(pos % p.Yo);
    int64_t t = pos / p.Yo;
    int xo = (int)(t % p.Xo);
    t /= p.Xo;
    int zo = (int)(t % p.Zo);
    int n = (int)(t / p.Zo);
    Vec<V> g;
    g.load(dy + pos * p.yp + c);
    IVec<V> am;
    float mx[V];
    if (p.tie == E2_TIE_FIRST) {
      am.load(amax + pos * p.yp + c);
    } else {
#pragma unroll
      for (int j = 0; j < V; ++j) mx[j] = -INFINITY;
      for (int dz = 0; dz < p.pz; ++dz)
        for (int dxx = 0; dxx < p.px; ++dxx)
          for (int dyy = 0; dyy < p.py; ++dyy) {
            int lin = ((zo * p.pz + dz) * p.X + xo * p.px + dxx) * p.Y + yo * p.py + dyy;
            Vec<V> v;
            v.load(x + ((int64_t)n * p.Z * p.X * p.Y + lin) * p.xp + c);
#pragma unroll
            for (int j = 0; j < V; ++j) mx[j] = fmaxf(mx[j], v.v[j]);
          }
    }
    for (int dz = 0; dz < p.pz; ++dz)
      for (int dxx = 0; dxx < p.px; ++dxx)
        for (int dyy = 0; dyy < p.py; ++dyy) {
          int lin = ((zo * p.pz + dz) * p.X + xo * p.px + dxx) * p.Y + yo * p.py + dyy;
          int64_t ofs = ((int64_t)n * p.Z * p.X * p.Y + lin) * p.xp + c;
          Vec<V> o;
          if (p.tie == E2_TIE_FIRST) {
#pragma unroll
            for (int j = 0; j < V; ++j) o.v[j] = (am.v[j] == lin) ? g.v[j] : 0.f;
          } else {
            Vec<V> v;
            v.load(x + ofs);
#pragma unroll
            for (int j = 0; j < V; ++j) o.v[j] = (v.v[j] == mx[j]) ? g.v[j] : 0.f;
          }
          if (gate) {  // fused ReLU backward of the layer that produced the pooled tensor
            Vec<V> gt;
            gt.load(gate + ofs);
#pragma unroll
            for (int j = 0; j < V; ++j)
              if (!(gt.v[j] > 0.f)) o.v[j] = 0.f;
          }
          if (p.accumulate) {
            Vec<V> old;
            old.load(dx + ofs);
#pragma unroll
            for (int j = 0; j < V; ++j) o.v[j] += old.v[j];
          }
          o.store(dx + ofs);
        }
  }
}

// Gather form for E2_TIE_FIRST: one block per INPUT row (n, z, x); a thread owns one pooled position x V
// channels of that row and writes the py input positions below it, so the two big streams (dx write, ReLU-gate
// read) are fully coalesced 16-byte accesses and dy / argmax are read once per window row.
template <int V>
__global__ void __launch_bounds__(256) k_maxpool_bwd_gather(PoolP p, E2FastDiv dcv, const float* __restrict__ dy,
                                                            const int* __restrict__ amax, float* __restrict__ dx,
                                                            const float* __restrict__ gate) {
  const int cv = (p.C + V - 1) / V;
  int r = blockIdx.x;
  const int xx = r % p.X;
  r /= p.X;
  const int zz = r % p.Z;
  const int n = r / p.Z;
  const int rowlen = p.Yo * cv;
  const int lin_row = (zz * p.X + xx) * p.Y;
  const int64_t irow = ((int64_t)n * p.Z * p.X * p.Y + lin_row) * p.xp;
  const int64_t orow = ((((int64_t)n * p.Zo + zz / p.pz) * p.Xo + xx / p.px) * p.Yo) * p.yp;
  for (int t = threadIdx.x; t < rowlen; t += blockDim.x) {
    const int yo = (int)dcv.div((uint32_t)t);
    const int c = (t - yo * cv) * V;
    Vec<V> g;
    IVec<V> am;
    const int64_t oofs = orow + (int64_t)yo * p.yp + c;
    g.load(dy + oofs);
    am.load(amax + oofs);
    if (gate && p.gate_pooled) {
      Vec<V> gt;
      gt.load(gate + oofs);
#pragma unroll
      for (int j = 0; j < V; ++j)
        if (!(gt.v[j] > 0.f)) g.v[j] = 0.f;
    }
    for (int k = 0; k < p.py; ++k) {
      const int yy = yo * p.py + k;
      const int lin = lin_row + yy;
      const int64_t ofs = irow + (int64_t)yy * p.xp + c;
      Vec<V> o;
#pragma unroll
      for (int j = 0; j < V; ++j) o.v[j] = (am.v[j] == lin) ? g.v[j] : 0.f;
      if (gate && !p.gate_pooled) {
        Vec<V> gt;
        gt.load(gate + ofs);
#pragma unroll
        for (int j = 0; j < V; ++j)
          if (!(gt.v[j] > 0.f)) o.v[j] = 0.f;
      }
      if (p.accumulate) {
        Vec<V> old;
        old.load(dx + ofs);
#pragma unroll
        for (int j = 0; j < V; ++j) o.v[j] += old.v[j];
      }
      o.store(dx + ofs);
    }
  }
}

static inline int pool_block(int rowlen) { return rowlen >= 256 ? 256 : ((rowlen + 31) / 32) * 32; }

static int fill_pool(e2_handle* h, const e2_pool_desc* d, PoolP* p) {
  E2_REQUIRE(h, d && e2_tensor_ok(&d->x) && e2_tensor_ok(&d->y), "maxpool3d: bad descriptor");
  E2_REQUIRE(h, d->pz >= 1 && d->px >= 1 && d->py >= 1, "maxpool3d: pool factors must be >= 1");
  // Pool._calc_shape (neural.py:1543-1546) / Conv._calc_shape (:746-750): axes must divide
  E2_REQUIRE(h, d->x.z % d->pz == 0 && d->x.x % d->px == 0 && d->x.y % d->py == 0,
             "maxpool3d: cannot downsample (%d,%d,%d) by (%d,%d,%d)", d->x.z, d->x.x, d->x.y, d->pz, d->px, d->py);
  E2_REQUIRE(h, d->y.z == d->x.z / d->pz && d->y.x == d->x.x / d->px && d->y.y == d->x.y / d->py && d->y.n == d->x.n &&
                    d->y.c == d->x.c,
             "maxpool3d: output extents do not match input/pool");
  E2_REQUIRE(h, (int64_t)d->x.z * d->x.x * d->x.y < (1ll << 31), "maxpool3d: volume too large for int32 argmax");
  p->n = d->x.n, p->Z = d->x.z, p->X = d->x.x, p->Y = d->x.y, p->C = d->x.c, p->xp = d->x.c_pitch;
  p->Zo = d->y.z, p->Xo = d->y.x, p->Yo = d->y.y, p->yp = d->y.c_pitch;
  p->pz = d->pz, p->px = d->px, p->py = d->py;
  p->act = d->act, p->has_bias = d->has_bias, p->tie = d->tie_mode, p->accumulate = d->accumulate;
  p->round_tf32 = d->round_tf32, p->gate_pooled = d->gate_pooled;
  return E2_OK;
}

extern "C" int e2_maxpool3d_fwd(e2_handle* h, const e2_pool_desc* d, const float* x, const float* bias, float* y,
                                int32_t* argmax, void* stream) {
  PoolP p;
  int rc = fill_pool(h, d, &p);
  if (rc) return rc;
  E2_REQUIRE(h, x && y && (!d->has_bias || bias), "maxpool3d_fwd: null pointer");
  if (d->mode != E2_POOL_MAX) {
    E2_REQUIRE(h, d->mode == E2_POOL_AVERAGE || d->mode == E2_POOL_SUM, "pool3d_fwd: unknown mode %d", d->mode);
    E2_REQUIRE(h, !d->has_bias && d->act == E2_ACT_LIN && !argmax, "pool3d_fwd: average / sum pooling has no fused epilogue or argmax");
    return e2_launch_avgpool_fwd(h, d, x, y, (cudaStream_t)stream);
  }
  const int64_t rows = (int64_t)p.n * p.Zo * p.Xo;
  E2_REQUIRE(h, rows < (1ll << 31), "maxpool3d_fwd: too many rows");
  if (vec4_ok({x, y, argmax}, {p.xp, p.yp}) && pad4_ok(p.C, p.xp) && pad4_ok(p.C, p.yp)) {
    const int rowlen = p.Yo * ((p.C + 3) / 4);
    const E2FastDiv dv = e2_fastdiv((p.C + 3) / 4, rowlen);
    cudaStream_t s = (cudaStream_t)stream;
    if (p.pz == 2 && p.px == 2 && p.py == 2)
      k_maxpool_fwd<4, 2, 2, 2><<<(int)rows, pool_block(rowlen), 0, s>>>(p, dv, x, bias, y, argmax);
    else if (p.pz == 1 && p.px == 2 && p.py == 2)
      k_maxpool_fwd<4, 1, 2, 2><<<(int)rows, pool_block(rowlen), 0, s>>>(p, dv, x, bias, y, argmax);
    else if (p.pz == 2 && p.px == 1 && p.py == 1)
      k_maxpool_fwd<4, 2, 1, 1><<<(int)rows, pool_block(rowlen), 0, s>>>(p, dv, x, bias, y, argmax);
    else
      k_maxpool_fwd<4, 0, 0, 0><<<(int)rows, pool_block(rowlen), 0, s>>>(p, dv, x, bias, y, argmax);
  } else {
    const int rowlen = p.Yo * p.C;
    k_maxpool_fwd<1, 0, 0, 0><<<(int)rows, pool_block(rowlen), 0, (cudaStream_t)stream>>>(p, e2_fastdiv(p.C, rowlen), x, bias, y, argmax);
  }
  e2_count_launch(h);
  E2_CUDA_CHECK(h, "maxpool3d_fwd");
  return E2_OK;
}

extern "C" int e2_maxpool3d_bwd(e2_handle* h, const e2_pool_desc* d, const float* dy, const int32_t* argmax,
                                const float* x, float* dx, const float* relu_gate, void* stream) {
  PoolP p;
  int rc = fill_pool(h, d, &p);
  if (rc) return rc;
  E2_REQUIRE(h, dy && dx, "maxpool3d_bwd: null pointer");
  if (d->mode != E2_POOL_MAX) {
    E2_REQUIRE(h, d->mode == E2_POOL_AVERAGE || d->mode == E2_POOL_SUM, "pool3d_bwd: unknown mode %d", d->mode);
    E2_REQUIRE(h, !relu_gate, "pool3d_bwd: average / sum pooling takes no ReLU gate");
    return e2_launch_avgpool_bwd(h, d, dy, dx, (cudaStream_t)stream);
  }
  E2_REQUIRE(h, d->tie_mode == E2_TIE_FIRST ? argmax != nullptr : x != nullptr,
             "maxpool3d_bwd: tie_mode FIRST needs argmax, tie_mode ALL needs x");
  E2_REQUIRE(h, !d->gate_pooled || (d->tie_mode == E2_TIE_FIRST && !d->has_bias && d->act == E2_ACT_LIN),
             "maxpool3d_bwd: gate_pooled needs tie_mode FIRST and a pool without bias/activation");
  int64_t work = e2_positions(&d->y) * d->y.c;
  if (p.tie == E2_TIE_FIRST) {
    const int64_t rows = (int64_t)p.n * p.Z * p.X;
    E2_REQUIRE(h, rows < (1ll << 31), "maxpool3d_bwd: too many rows");
    if (vec4_ok({dy, dx, argmax, relu_gate}, {p.xp, p.yp}) && pad4_ok(p.C, p.xp) && pad4_ok(p.C, p.yp)) {
      const int rowlen = p.Yo * ((p.C + 3) / 4);
      k_maxpool_bwd_gather<4><<<(int)rows, pool_block(rowlen), 0, (cudaStream_t)stream>>>(p, e2_fastdiv((p.C + 3) / 4, rowlen), dy, argmax, dx, relu_gate);
    } else {
      const int rowlen = p.Yo * p.C;
      k_maxpool_bwd_gather<1><<<(int)rows, pool_block(rowlen), 0, (cudaStream_t)stream>>>(p, e2_fastdiv(p.C, rowlen), dy, argmax, dx, relu_gate);
    }
  } else if (vec4_ok({dy, dx, argmax, x}, {p.C, p.xp, p.yp})) {
    k_maxpool_bwd<4><<<e2_grid_1d(work / 4, 256, h->sm_count), 256, 0, (cudaStream_t)stream>>>(p, dy, argmax, x, dx,
                                                                                               relu_gate);
  } else {
    k_maxpool_bwd<1><<<e2_grid_1d(work, 256, h->sm_count), 256, 0, (cudaStream_t)stream>>>(p, dy, argmax, x, dx,
                                                                                           relu_gate);
  }
  e2_count_launch(h);
  E2_CUDA_CHECK(h, "maxpool3d_bwd");
  return E2_OK;
}

// --------------------------------------------------------------- MFP forward / backward
// One pass produces all prod(p) fragments (the reference issues prod(p) pool ops + a
// concat copy, computations.py:665-674).  Fragment index = off_idx * n_in + n_old with
// off_idx = (iz*px + ix)*py + iy  (itertools.product order, last axis fastest).
template <int V>
__global__ void __launch_bounds__(256) k_mfp_fwd(PoolP p, E2FastDiv dcv, const float* __restrict__ x,
                                                 const float* __restrict__ bias, float* __restrict__ y,
                                                 int* __restrict__ amax) {
  // one block per output row (fragment, zo, xo), see k_maxpool_fwd.  The fragment offset is the FASTEST block
  // coordinate: the prod(p) blocks that read the same input rows are neighbours in launch order, so those rows cross
  // HBM once and are shared through L2 (fragment-major order re-read the input prod(p) times: 1.4 TB/s algorithmic)
  const int cv = (p.C + V - 1) / V;
  const int nfr = p.pz * p.px * p.py;
  int r = blockIdx.x;
  const int off = r % nfr;
  r /= nfr;
  const int xo = r % p.Xo;
  r /= p.Xo;
  const int zo = r % p.Zo;
  const int n = r / p.Zo;
  const int fr = off * p.n + n;
  const int iy = off % p.py, ix = (off / p.py) % p.px, iz = off / (p.py * p.px);
  const int rowlen = p.Yo * cv;
  const float* xn = x + (int64_t)n * p.Z * p.X * p.Y * p.xp;
  const int64_t orow = (((int64_t)fr * p.Zo + zo) * p.Xo + xo) * p.Yo;
  for (int t = threadIdx.x; t < rowlen; t += blockDim.x) {
    const int yo = (int)dcv.div((uint32_t)t);
    const int c = (t - yo * cv) * V;
    float best[V];
    int bi[V];
    bool first = true;
#pragma unroll
    for (int j = 0; j < V; ++j) best[j] = -INFINITY, bi[j] = 0;
    for (int dz = 0; dz < p.pz; ++dz)
      for (int dx = 0; dx < p.px; ++dx) {
        const int lin0 = ((iz + zo * p.pz + dz) * p.X + ix + xo * p.px + dx) * p.Y + iy + yo * p.py;
        const float* src = xn + (int64_t)lin0 * p.xp + c;
#pragma unroll 2
        for (int dy = 0; dy < p.py; ++dy) {
          Vec<V> v;
          v.load(src + (int64_t)dy * p.xp);
#pragma unroll
          for (int j = 0; j < V; ++j)
            if (first || v.v[j] > best[j]) best[j] = v.v[j], bi[j] = lin0 + dy;
          first = false;
        }
      }
    Vec<V> o;
    IVec<V> oi;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      float r2 = best[j];
      if (c + j < p.C) {
        if (p.has_bias) r2 += __ldg(bias + c + j);
        r2 = e2_apply_act(r2, p.act);
        o.v[j] = p.round_tf32 ? e2_round_tf32(r2) : r2;
      } else {
        o.v[j] = 0.f;
      }
      oi.v[j] = bi[j];
    }
    const int64_t oofs = (orow + yo) * p.yp + c;
    o.store(y + oofs);
    if (amax) oi.store(amax + oofs);
  }
}

// Shared-memory form of the forward pass.  k_mfp_fwd above launches one block per OUTPUT row, fragment-major: the
// prod(p) fragments read the same input rows at launch positions far apart, so for a tensor larger than L2 every input
// byte crosses HBM prod(p) times (measured 1.4 - 2.4 TB/s of algorithmic traffic on the dense-prediction tile).
// Here a block owns the outputs (zo, xo, y chunk) of ALL fragments: it stages the (2pz-1) x (2px-1) input rows those
// windows touch once (coalesced float4), computes every fragment's maxima from shared memory and writes each
// fragment row with full-line stores.  Neighbouring blocks share input rows through L2 (adjacent in launch order).
struct MfpTile {
  int rows_z, rows_x;      // staged input rows: 2pz-1, 2px-1
  int yc, n_chunks;        // output positions per fragment and chunk, chunks per row
  int in_pos;              // staged positions per row: yc*py + py - 1
};

__global__ void __launch_bounds__(256) k_mfp_fwd_tile(PoolP p, MfpTile m, E2FastDiv dcv, const float* __restrict__ x,
                                                      const float* __restrict__ bias, float* __restrict__ y,
                                                      int* __restrict__ amax) {
  extern __shared__ float4 tile[];                 // [rows_z*rows_x][in_pos][cv]
  const int cv = (p.C + 3) / 4;
  int b = blockIdx.x;
  const int ch = b % m.n_chunks;
  b /= m.n_chunks;
  const int xo = b % p.Xo;
  b /= p.Xo;
  const int zo = b % p.Zo;
  const int n = b / p.Zo;
  const int yo0 = ch * m.yc;
  const int nyo = min(m.yc, p.Yo - yo0);           // output positions of this chunk
  const int y_in0 = yo0 * p.py;
  const int npos = nyo * p.py + p.py - 1;          // input positions needed (<= in_pos, inside the row by the MFP rule)
  const int nrows = m.rows_z * m.rows_x;
  const float* xn = x + (int64_t)n * p.Z * p.X * p.Y * p.xp;
  // ---- stage: a row segment is npos * cv consecutive float4 in HBM (pitch == 4 * cv on this path)
  const int per_row = npos * cv;
  for (int r = 0; r < nrows; ++r) {
    const int rz = r / m.rows_x, rx = r - rz * m.rows_x;
    const int64_t lin = ((int64_t)(zo * p.pz + rz) * p.X + xo * p.px + rx) * p.Y + y_in0;
    const float4* src = reinterpret_cast<const float4*>(xn + lin * p.xp);
    float4* dst = tile + r * m.in_pos * cv;
    for (int q = threadIdx.x; q < per_row; q += blockDim.x) dst[q] = __ldg(src + q);
  }
  __syncthreads();
  // ---- compute: for each fragment offset, item = (yo, channel group), channel group fastest
  const int nfr = p.pz * p.px * p.py;
  const int per_frag = nyo * cv;
  for (int off = 0; off < nfr; ++off) {
    const int iy = off % p.py, ix = (off / p.py) % p.px, iz = off / (p.py * p.px);
    const int fr = off * p.n + n;
    const int64_t obase = ((((int64_t)fr * p.Zo + zo) * p.Xo + xo) * p.Yo + yo0) * p.yp;
    for (int q = threadIdx.x; q < per_frag; q += blockDim.x) {
      const int yl = (int)dcv.div((uint32_t)q);
      const int c4 = q - yl * cv;
      float best[4];
      int bi[4];
      bool first = true;
      for (int dz = 0; dz < p.pz; ++dz)
        for (int dx = 0; dx < p.px; ++dx) {
          const int r = (iz + dz) * m.rows_x + ix + dx;
          const int lin0 = ((zo * p.pz + iz + dz) * p.X + xo * p.px + ix + dx) * p.Y + y_in0 + iy + yl * p.py;
          const float4* src = tile + (r * m.in_pos + iy + yl * p.py) * cv + c4;
          for (int dy = 0; dy < p.py; ++dy) {
            const float4 v4 = src[dy * cv];
            const float v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (first || v[j] > best[j]) best[j] = v[j], bi[j] = lin0 + dy;     // strict '>': FIRST maximum in scan order
            first = false;
          }
        }
      const int c = c4 * 4;
      float o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float r2 = best[j];
        if (c + j < p.C) {
          if (p.has_bias) r2 += __ldg(bias + c + j);
          r2 = e2_apply_act(r2, p.act);
          o[j] = p.round_tf32 ? e2_round_tf32(r2) : r2;
        } else {
          o[j] = 0.f;
        }
      }
      const int64_t oofs = obase + (int64_t)q * 4;      // (yl * cv + c4) * 4 == yl * pitch + c
      *reinterpret_cast<float4*>(y + oofs) = make_float4(o[0], o[1], o[2], o[3]);
      if (amax) *reinterpret_cast<int4*>(amax + oofs) = make_int4(bi[0], bi[1], bi[2], bi[3]);
    }
  }
}

// plan the shared-memory tile; false: rows too long even for one output position per chunk
static bool plan_mfp_tile(const PoolP& p, MfpTile* m, size_t* smem) {
  const int cv = (p.C + 3) / 4;
  m->rows_z = 2 * p.pz - 1, m->rows_x = 2 * p.px - 1;
  const size_t budget = 64 * 1024;                 // <= 3 blocks per SM
  const size_t per_pos = (size_t)m->rows_z * m->rows_x * cv * 16;
  const int64_t max_pos = (int64_t)(budget / per_pos);
  int64_t yc = (max_pos - (p.py - 1)) / p.py;
  if (yc < 1) return false;
  if (yc > p.Yo) yc = p.Yo;
  m->n_chunks = (int)((p.Yo + yc - 1) / yc);
  yc = (p.Yo + m->n_chunks - 1) / m->n_chunks;     // even chunks
  m->yc = (int)yc;
  m->in_pos = m->yc * p.py + p.py - 1;
  *smem = per_pos * m->in_pos;
  return true;
}

// Sliding form of the forward pass (round 2, default for windows of 1 or 2 per axis).  Max-fragment-pooling is a
// STRIDE-1 max-pool whose outputs are de-interleaved: the window starting at (sz,sx,sy) goes to fragment
// (sz % pz, sx % px, sy % py) at position (sz / pz, sx / px, sy / py).  The row kernel above computes every fragment's
// windows from scratch, so each input element travels from L2 prod(p) times (2.0 - 2.6 TB/s algorithmic on the
// dense-prediction tile: the L2, not HBM, was the limit).  Here a thread owns (sy, 4 channels) at a fixed coordinate of
// one outer axis and WALKS the other one (x for in-plane windows, z when the window spans z): per step it loads the
// cross-section of the window once, keeps the previous step's partial maximum in registers, and emits one output --
// every input element is loaded once per thread that needs it (py times, neighbours in L1), not prod(p) times.
// Ties: each partial result is (max value, smallest linear index); merges compare the index on equal values, so the
// argmax is the first maximum in (z,x,y) scan order whatever the merge order.
template <int PZ, int PX, int PY, bool WALKZ, bool AMAX>
__global__ void __launch_bounds__(256) k_mfp_fwd_slide(PoolP p, int seg, int nseg, int nchunks, E2FastDiv dcv,
                                                       const float* __restrict__ x, const float* __restrict__ bias,
                                                       float* __restrict__ y, int* __restrict__ amax) {
  constexpr int PW = WALKZ ? PZ : PX;      // window extent along the walking axis
  constexpr int PA = WALKZ ? PX : PZ;      // cross-section: the fixed outer axis ... and y (PY)
  static_assert(PW == 2, "the walking axis is one the window spans");
  const int cv = (p.C + 3) / 4;
  const int Sz = p.Z - PZ + 1, Sx = p.X - PX + 1, Sy = p.Y - PY + 1;   // window start positions per axis
  const int Sf = WALKZ ? Sx : Sz, Sw = WALKZ ? Sz : Sx;
  int b = blockIdx.x;
  const int chunk = b % nchunks;
  b /= nchunks;
  const int f = b % Sf;                    // neighbouring blocks share cross-section rows through L2
  b /= Sf;
  const int sg = b % nseg;
  const int n = b / nseg;
  const int t = chunk * 256 + threadIdx.x;
  if (t >= Sy * cv) return;
  const int sy = (int)dcv.div((uint32_t)t);
  const int c = (t - sy * cv) * 4;
  const int w0 = sg * seg, w1 = min(Sw, w0 + seg);
  const int sw = WALKZ ? p.X * p.Y : p.Y;  // input positions per step of the walking axis
  const int sf = WALKZ ? p.Y : p.X * p.Y;  // ... and of the fixed axis
  const float* xn = x + (int64_t)n * p.Z * p.X * p.Y * p.xp + c;
  float4 bz = make_float4(0.f, 0.f, 0.f, 0.f);
  if (p.has_bias) {
    bz.x = c + 0 < p.C ? __ldg(bias + c + 0) : 0.f, bz.y = c + 1 < p.C ? __ldg(bias + c + 1) : 0.f;
    bz.z = c + 2 < p.C ? __ldg(bias + c + 2) : 0.f, bz.w = c + 3 < p.C ? __ldg(bias + c + 3) : 0.f;
  }
  // maximum over the window's cross-section at walking coordinate wp; loads in ascending linear index, strict '>'
  auto cross = [&](int wp, float (&v)[4], int (&li)[4]) {
#pragma unroll
    for (int da = 0; da < PA; ++da)
#pragma unroll
      for (int dy = 0; dy < PY; ++dy) {
        const int lin = wp * sw + (f + da) * sf + sy + dy;
        const float4 q = __ldg(reinterpret_cast<const float4*>(xn + (int64_t)lin * p.xp));
        const float e[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if ((da == 0 && dy == 0) || e[j] > v[j]) {
            v[j] = e[j];
            if (AMAX) li[j] = lin;
          }
      }
  };
  float g0[4], g1[4];
  int l0[4], l1[4];
  cross(w0, g0, l0);
#pragma unroll 2
  for (int w = w0; w < w1; ++w) {
    cross(w + 1, g1, l1);
    float o[4];
    int oi[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      // walking along z the later cross-section has the larger index: strict '>' is enough; walking along x with a
      // window over z it has not (dz outranks dx), so equal values compare their indices
      const bool take = g1[j] > g0[j] || (AMAX && !WALKZ && PA > 1 && g1[j] == g0[j] && l1[j] < l0[j]);
      float r2 = take ? g1[j] : g0[j];
      if (AMAX) oi[j] = take ? l1[j] : l0[j];
      if (c + j < p.C) {
        r2 += j == 0 ? bz.x : j == 1 ? bz.y : j == 2 ? bz.z : bz.w;     // reference order: pool -> +bias -> act
        r2 = e2_apply_act(r2, p.act);
        o[j] = p.round_tf32 ? e2_round_tf32(r2) : r2;
      } else {
        o[j] = 0.f;                                                      // pad lane
      }
      g0[j] = g1[j];
      if (AMAX) l0[j] = l1[j];
    }
    const int sz = WALKZ ? w : f, sx = WALKZ ? f : w;
    const int off = ((sz % PZ) * PX + sx % PX) * PY + sy % PY;           // itertools.product order, last axis fastest
    const int fr = off * p.n + n;
    const int64_t oofs = ((((int64_t)fr * p.Zo + sz / PZ) * p.Xo + sx / PX) * p.Yo + sy / PY) * p.yp + c;
    *reinterpret_cast<float4*>(y + oofs) = make_float4(o[0], o[1], o[2], o[3]);
    if (AMAX) *reinterpret_cast<int4*>(amax + oofs) = make_int4(oi[0], oi[1], oi[2], oi[3]);
  }
}

template <int PZ, int PX, int PY, bool WALKZ>
static void launch_mfp_slide(const PoolP& p, int sm_count, const float* x, const float* bias, float* y, int* amax,
                             cudaStream_t s) {
  const int cv = (p.C + 3) / 4;
  const int Sz = p.Z - PZ + 1, Sx = p.X - PX + 1, Sy = p.Y - PY + 1;
  const int Sf = WALKZ ? Sx : Sz, Sw = WALKZ ? Sz : Sx;
  const int nchunks = (Sy * cv + 255) / 256;
  // segments of the walking axis: long enough that the one re-read cross-section per segment is small change, short
  // enough for a few waves of blocks
  int seg = 16;
  while (seg > 4 && (int64_t)p.n * ((Sw + seg - 1) / seg) * Sf * nchunks < 8ll * sm_count) seg /= 2;
  const int nseg = (Sw + seg - 1) / seg;
  const unsigned grid = (unsigned)((int64_t)p.n * nseg * Sf * nchunks);
  const E2FastDiv dv = e2_fastdiv((uint32_t)cv, (uint64_t)Sy * cv);
  if (amax)
    k_mfp_fwd_slide<PZ, PX, PY, WALKZ, true><<<grid, 256, 0, s>>>(p, seg, nseg, nchunks, dv, x, bias, y, amax);
  else
    k_mfp_fwd_slide<PZ, PX, PY, WALKZ, false><<<grid, 256, 0, s>>>(p, seg, nseg, nchunks, dv, x, bias, y, amax);
}

// Gather form (no atomics, deterministic): each input element checks the prod(p) windows
// (one per fragment offset) that contain it.
template <int V>
__global__ void __launch_bounds__(256) k_mfp_bwd(PoolP p, const float* __restrict__ dy, const int* __restrict__ amax,
                                                 float* __restrict__ dx) {
  const int cv = p.C / V;
  const int64_t total = (int64_t)p.n * p.Z * p.X * p.Y * cv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % cv) * V;
    int64_t pos = i / cv;
    int yy = (int)(pos % p.Y);
    int64_t t = pos / p.Y;
    int xx = (int)(t % p.X);
    t /= p.X;
    int zz = (int)(t % p.Z);
    int n = (int)(t / p.Z);
    int lin = (zz * p.X + xx) * p.Y + yy;
    float acc[V];
#pragma unroll
    for (int j = 0; j < V; ++j) acc[j] = 0.f;
    for (int iz = 0; iz < p.pz; ++iz) {
      int rz = zz - iz;
      if (rz < 0) continue;
      int zo = rz / p.pz;
      if (zo >= p.Zo) continue;
      for (int ix = 0; ix < p.px; ++ix) {
        int rx = xx - ix;
        if (rx < 0) continue;
        int xo = rx / p.px;
        if (xo >= p.Xo) continue;
        for (int iy = 0; iy < p.py; ++iy) {
          int ry = yy - iy;
          if (ry < 0) continue;
          int yo = ry / p.py;
          if (yo >= p.Yo) continue;
          int fr = ((iz * p.px + ix) * p.py + iy) * p.n + n;
          int64_t o = ((((int64_t)fr * p.Zo + zo) * p.Xo + xo) * p.Yo + yo) * p.yp + c;
          IVec<V> am;
          am.load(amax + o);
          Vec<V> g;
          g.load(dy + o);
#pragma unroll
          for (int j = 0; j < V; ++j)
            if (am.v[j] == lin) acc[j] += g.v[j];
        }
      }
    }
    Vec<V> o;
#pragma unroll
    for (int j = 0; j < V; ++j) o.v[j] = acc[j];
    o.store(dx + pos * p.xp + c);
  }
}

static int fill_mfp(e2_handle* h, const e2_mfp_desc* d, PoolP* p) {
  E2_REQUIRE(h, d && e2_tensor_ok(&d->x) && e2_tensor_ok(&d->y), "mfp: bad descriptor");
  E2_REQUIRE(h, d->pz >= 1 && d->px >= 1 && d->py >= 1, "mfp: pool factors must be >= 1");
  // Conv._calc_shape MFP rule (neural.py:739-744): (S - p + 1) % p == 0
  E2_REQUIRE(h, (d->x.z - d->pz + 1) % d->pz == 0 && (d->x.x - d->px + 1) % d->px == 0 && (d->x.y - d->py + 1) % d->py == 0,
             "mfp: cannot pool (%d,%d,%d) by (%d,%d,%d) using MFP", d->x.z, d->x.x, d->x.y, d->pz, d->px, d->py);
  E2_REQUIRE(h, d->y.n == d->x.n * d->pz * d->px * d->py && d->y.z == (d->x.z - d->pz + 1) / d->pz &&
                    d->y.x == (d->x.x - d->px + 1) / d->px && d->y.y == (d->x.y - d->py + 1) / d->py && d->y.c == d->x.c,
             "mfp: output extents do not match input/pool");
  E2_REQUIRE(h, (int64_t)d->x.z * d->x.x * d->x.y < (1ll << 31), "mfp: volume too large for int32 argmax");
  p->n = d->x.n, p->Z = d->x.z, p->X = d->x.x, p->Y = d->x.y, p->C = d->x.c, p->xp = d->x.c_pitch;
  p->Zo = d->y.z, p->Xo = d->y.x, p->Yo = d->y.y, p->yp = d->y.c_pitch;
  p->pz = d->pz, p->px = d->px, p->py = d->py;
  p->act = d->act, p->has_bias = d->has_bias, p->tie = 0, p->accumulate = 0;
  p->round_tf32 = d->round_tf32, p->gate_pooled = 0;
  return E2_OK;
}

extern "C" int e2_mfp_fwd(e2_handle* h, const e2_mfp_desc* d, const float* x, const float* bias, float* y,
                          int32_t* argmax, void* stream) {
  PoolP p;
  int rc = fill_mfp(h, d, &p);
  if (rc) return rc;
  E2_REQUIRE(h, x && y && (!d->has_bias || bias), "mfp_fwd: null pointer");
  const int64_t rows = (int64_t)p.n * p.pz * p.px * p.py * p.Zo * p.Xo;
  E2_REQUIRE(h, rows < (1ll << 31), "mfp_fwd: too many rows");
  MfpTile mt;
  size_t smem = 0;
  static const bool use_tile = getenv("E2_MFP_TILE") != nullptr;      // A/B switches for profiling (default: sliding kernel)
  static const bool row_only = getenv("E2_MFP_ROW") != nullptr;       // the round-1 row kernel
  if (vec4_ok({x, y, argmax}, {p.xp, p.yp}) && pad4_ok(p.C, p.xp) && pad4_ok(p.C, p.yp) && use_tile &&
      p.xp == 4 * ((p.C + 3) / 4) && p.yp == p.xp && plan_mfp_tile(p, &mt, &smem) && (int64_t)p.n * p.Zo * p.Xo * mt.n_chunks < (1ll << 31)) {
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(k_mfp_fwd_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return e2_fail(h, E2_ERR_CUDA, "mfp_fwd: cudaFuncSetAttribute(max dynamic smem) failed");
    k_mfp_fwd_tile<<<(unsigned)((int64_t)p.n * p.Zo * p.Xo * mt.n_chunks), 256, smem, (cudaStream_t)stream>>>(
        p, mt, e2_fastdiv((p.C + 3) / 4, (uint64_t)mt.yc * ((p.C + 3) / 4)), x, bias, y, argmax);
  } else if (!row_only && vec4_ok({x, y, argmax}, {p.xp, p.yp}) && pad4_ok(p.C, p.xp) && pad4_ok(p.C, p.yp) &&
             ((p.pz == 1 && p.px == 2 && p.py == 2) || (p.pz == 2 && p.px == 1 && p.py == 1) ||
              (p.pz == 2 && p.px == 2 && p.py == 2) || (p.pz == 1 && p.px == 2 && p.py == 1)) &&
             (int64_t)p.n * p.Z * p.X * ((p.Y * ((p.C + 3) / 4) + 255) / 256) < (1ll << 31)) {
    cudaStream_t s = (cudaStream_t)stream;
    if (p.pz == 1 && p.py == 2) launch_mfp_slide<1, 2, 2, false>(p, h->sm_count, x, bias, y, argmax, s);
    else if (p.pz == 1) launch_mfp_slide<1, 2, 1, false>(p, h->sm_count, x, bias, y, argmax, s);
    else if (p.px == 1) launch_mfp_slide<2, 1, 1, true>(p, h->sm_count, x, bias, y, argmax, s);
    else launch_mfp_slide<2, 2, 2, true>(p, h->sm_count, x, bias, y, argmax, s);
  } else if (vec4_ok({x, y, argmax}, {p.xp, p.yp}) && pad4_ok(p.C, p.xp) && pad4_ok(p.C, p.yp)) {
    const int rowlen = p.Yo * ((p.C + 3) / 4);
    k_mfp_fwd<4><<<(int)rows, pool_block(rowlen), 0, (cudaStream_t)stream>>>(p, e2_fastdiv((p.C + 3) / 4, rowlen), x, bias, y, argmax);
  } else {
    const int rowlen = p.Yo * p.C;
    k_mfp_fwd<1><<<(int)rows, pool_block(rowlen), 0, (cudaStream_t)stream>>>(p, e2_fastdiv(p.C, rowlen), x, bias, y, argmax);
  }
  e2_count_launch(h);
  E2_CUDA_CHECK(h, "mfp_fwd");
  return E2_OK;
}

extern "C" int e2_mfp_bwd(e2_handle* h, const e2_mfp_desc* d, const float* dy, const int32_t* argmax, float* dx,
                          void* stream) {
  PoolP p;
  int rc = fill_mfp(h, d, &p);
  if (rc) return rc;
  E2_REQUIRE(h, dy && argmax && dx, "mfp_bwd: null pointer");
  int64_t work = e2_positions(&d->x) * d->x.c;
  if (vec4_ok({dy, dx, argmax}, {p.C, p.xp, p.yp}))
    k_mfp_bwd<4><<<e2_grid_1d(work / 4, 256, h->sm_count), 256, 0, (cudaStream_t)stream>>>(p, dy, argmax, dx);
  else
    k_mfp_bwd<1><<<e2_grid_1d(work, 256, h->sm_count), 256, 0, (cudaStream_t)stream>>>(p, dy, argmax, dx);
  e2_count_launch(h);
  E2_CUDA_CHECK(h, "mfp_bwd");
  return E2_OK;
}

// ------------------------------------------------------------- fragments <-> dense
struct F2DP {
  int nfr, Z, X, Y, C, fp, dp;
  int sz, sx, sy;
};

template <int V, bool FWD>
__global__ void __launch_bounds__(256) k_f2d(F2DP p, const float* __restrict__ src, const int* __restrict__ offs,
                                             float* __restrict__ dst) {
  const int cv = p.C / V;
  const int64_t total = (int64_t)p.nfr * p.Z * p.X * p.Y * cv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % cv) * V;
    int64_t pos = i / cv;
    int yy = (int)(pos % p.Y);
    int64_t t = pos / p.Y;
    int xx = (int)(t % p.X);
    t /= p.X;
    int zz = (int)(t % p.Z);
    int fr = (int)(t / p.Z);
    int oz = __ldg(offs + fr * 3), ox = __ldg(offs + fr * 3 + 1), oy = __ldg(offs + fr * 3 + 2);
    int64_t dpos = ((int64_t)(zz * p.sz + oz) * (p.X * p.sx) + (xx * p.sx + ox)) * (p.Y * p.sy) + (yy * p.sy + oy);
    Vec<V> v;
    if (FWD) {
      v.load(src + pos * p.fp + c);
      v.store(dst + dpos * p.dp + c);
    } else {
      v.load(src + dpos * p.dp + c);
      v.store(dst + pos * p.fp + c);
    }
  }
}

static int f2d_launch(e2_handle* h, const e2_f2d_desc* d, const float* src, const int32_t* offs, float* dst, bool fwd,
                      void* stream) {
  E2_REQUIRE(h, d && e2_tensor_ok(&d->frag) && e2_tensor_ok(&d->dense) && src && dst && offs, "frag2dense: bad arguments");
  // FragmentsToDense._make_output (neural.py:873-875)
  E2_REQUIRE(h, d->frag.n == d->sz * d->sx * d->sy, "frag2dense: need %d fragments on the batch axis, got %d",
             d->sz * d->sx * d->sy, d->frag.n);
  E2_REQUIRE(h, d->dense.n == 1 && d->dense.z == d->frag.z * d->sz && d->dense.x == d->frag.x * d->sx &&
                    d->dense.y == d->frag.y * d->sy && d->dense.c == d->frag.c,
             "frag2dense: dense extents do not match fragments*strides");
  F2DP p = {d->frag.n, d->frag.z, d->frag.x, d->frag.y, d->frag.c, d->frag.c_pitch, d->dense.c_pitch, d->sz, d->sx, d->sy};
  int64_t work = e2_positions(&d->frag) * d->frag.c;
  bool v4 = vec4_ok({src, dst}, {p.C, p.fp, p.dp});
  cudaStream_t s = (cudaStream_t)stream;
  if (fwd) {
    if (v4) k_f2d<4, true><<<e2_grid_1d(work / 4, 256, h->sm_count), 256, 0, s>>>(p, src, offs, dst);
    else k_f2d<1, true><<<e2_grid_1d(work, 256, h->sm_count), 256, 0, s>>>(p, src, offs, dst);
  } else {
    if (v4) k_f2d<4, false><<<e2_grid_1d(work / 4, 256, h->sm_count), 256, 0, s>>>(p, src, offs, dst);
    else k_f2d<1, false><<<e2_grid_1d(work, 256, h->sm_count), 256, 0, s>>>(p, src, offs, dst);
  }
  e2_count_launch(h);
  E2_CUDA_CHECK(h, "frag2dense");
  return E2_OK;
}

extern "C" int e2_frag2dense_fwd(e2_handle* h, const e2_f2d_desc* d, const float* frag, const int32_t* offsets,
                                 float* dense, void* stream) {
  return f2d_launch(h, d, frag, offsets, dense, true, stream);
}
extern "C" int e2_frag2dense_bwd(e2_handle* h, const e2_f2d_desc* d, const float* ddense, const int32_t* offsets,
                                 float* dfrag, void* stream) {
  return f2d_launch(h, d, ddense, offsets, dfrag, false, stream);
}

// -------------------------------------------------------------------- crop + concat
struct CropP {
  int n, Z, X, Y, C, sp;   // src
  int Zd, Xd, Yd, dp, c0;  // dst
  int oz, ox, oy, accumulate;
};

template <int V>
__global__ void __launch_bounds__(256) k_crop_fwd(CropP p, const float* __restrict__ src, float* __restrict__ dst) {
  const int cv = p.C / V;
  const int64_t total = (int64_t)p.n * p.Zd * p.Xd * p.Yd * cv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % cv) * V;
    int64_t pos = i / cv;
    int yy = (int)(pos % p.Yd);
    int64_t t = pos / p.Yd;
    int xx = (int)(t % p.Xd);
    t /= p.Xd;
    int zz = (int)(t % p.Zd);
    int n = (int)(t / p.Zd);
    int64_t spos = (((int64_t)n * p.Z + zz + p.oz) * p.X + xx + p.ox) * p.Y + yy + p.oy;
    Vec<V> v;
    v.load(src + spos * p.sp + c);
    v.store(dst + pos * p.dp + p.c0 + c);
  }
}

template <int V>
__global__ void __launch_bounds__(256) k_crop_bwd(CropP p, const float* __restrict__ ddst, float* __restrict__ dsrc,
                                                  const float* __restrict__ gate) {
  const int cv = p.C / V;
  if (p.accumulate) {
    // only the cropped region changes: walk the (small) gradient of the crop, not the whole source
    const int64_t total = (int64_t)p.n * p.Zd * p.Xd * p.Yd * cv;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
      int c = (int)(i % cv) * V;
      int64_t dpos = i / cv;
      int yd = (int)(dpos % p.Yd);
      int64_t t = dpos / p.Yd;
      int xd = (int)(t % p.Xd);
      t /= p.Xd;
      int zd = (int)(t % p.Zd);
      int n = (int)(t / p.Zd);
      int64_t pos = (((int64_t)n * p.Z + zd + p.oz) * p.X + xd + p.ox) * p.Y + yd + p.oy;
      Vec<V> v, old;
      v.load(ddst + dpos * p.dp + p.c0 + c);
      if (gate) {
        Vec<V> gt;
        gt.load(gate + pos * p.sp + c);
#pragma unroll
        for (int j = 0; j < V; ++j)
          if (!(gt.v[j] > 0.f)) v.v[j] = 0.f;
      }
      old.load(dsrc + pos * p.sp + c);
#pragma unroll
      for (int j = 0; j < V; ++j) v.v[j] += old.v[j];
      v.store(dsrc + pos * p.sp + c);
    }
    return;
  }
  const int64_t total = (int64_t)p.n * p.Z * p.X * p.Y * cv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % cv) * V;
    int64_t pos = i / cv;
    int yy = (int)(pos % p.Y);
    int64_t t = pos / p.Y;
    int xx = (int)(t % p.X);
    t /= p.X;
    int zz = (int)(t % p.Z);
    int n = (int)(t / p.Z);
    int zd = zz - p.oz, xd = xx - p.ox, yd = yy - p.oy;
    bool inside = zd >= 0 && zd < p.Zd && xd >= 0 && xd < p.Xd && yd >= 0 && yd < p.Yd;
    Vec<V> v;
    if (inside) {
      int64_t dpos = (((int64_t)n * p.Zd + zd) * p.Xd + xd) * p.Yd + yd;
      v.load(ddst + dpos * p.dp + p.c0 + c);
      if (gate) {  // fused ReLU backward of the layer that produced src
        Vec<V> gt;
        gt.load(gate + pos * p.sp + c);
#pragma unroll
        for (int j = 0; j < V; ++j)
          if (!(gt.v[j] > 0.f)) v.v[j] = 0.f;
      }
    } else {
#pragma unroll
      for (int j = 0; j < V; ++j) v.v[j] = 0.f;
    }
    v.store(dsrc + pos * p.sp + c);
  }
}

static int fill_crop(e2_handle* h, const e2_crop_desc* d, CropP* p) {
  E2_REQUIRE(h, d && e2_tensor_ok(&d->src) && e2_tensor_ok(&d->dst), "crop_concat: bad descriptor");
  E2_REQUIRE(h, d->oz >= 0 && d->ox >= 0 && d->oy >= 0, "crop_concat: negative crop");
  E2_REQUIRE(h, d->dst.z == d->src.z - 2 * d->oz && d->dst.x == d->src.x - 2 * d->ox && d->dst.y == d->src.y - 2 * d->oy &&
                    d->dst.n == d->src.n,
             "crop_concat: destination extents do not match source minus 2*crop");
  E2_REQUIRE(h, d->dst_c0 >= 0 && d->dst_c0 + d->src.c <= d->dst.c, "crop_concat: channel slice [%d,%d) outside %d",
             d->dst_c0, d->dst_c0 + d->src.c, d->dst.c);
  *p = CropP{d->src.n, d->src.z, d->src.x, d->src.y, d->src.c, d->src.c_pitch, d->dst.z, d->dst.x, d->dst.y,
             d->dst.c_pitch, d->dst_c0, d->oz, d->ox, d->oy, d->accumulate};
  return E2_OK;
}

extern "C" int e2_crop_concat_fwd(e2_handle* h, const e2_crop_desc* d, const float* src, float* dst, void* stream) {
  CropP p;
  int rc = fill_crop(h, d, &p);
  if (rc) return rc;
  E2_REQUIRE(h, src && dst, "crop_concat_fwd: null pointer");
  int64_t work = e2_positions(&d->dst) * d->src.c;
  if (vec4_ok({src, dst}, {p.C, p.sp, p.dp, p.c0}))
    k_crop_fwd<4><<<e2_grid_1d(work / 4, 256, h->sm_count), 256, 0, (cudaStream_t)stream>>>(p, src, dst);
  else
    k_crop_fwd<1><<<e2_grid_1d(work, 256, h->sm_count), 256, 0, (cudaStream_t)stream>>>(p, src, dst);
  e2_count_launch(h);
  E2_CUDA_CHECK(h, "crop_concat_fwd");
  return E2_OK;
}

extern "C" int e2_crop_concat_bwd(e2_handle* h, const e2_crop_desc* d, const float* ddst, float* dsrc,
                                  const float* relu_gate, void* stream) {
  CropP p;
  int rc = fill_crop(h, d, &p);
  if (rc) return rc;
  E2_REQUIRE(h, ddst && dsrc, "crop_concat_bwd: null pointer");
  int64_t work = (d->accumulate ? e2_positions(&d->dst) : e2_positions(&d->src)) * d->src.c;
  if (vec4_ok({ddst, dsrc}, {p.C, p.sp, p.dp, p.c0}))
    k_crop_bwd<4><<<e2_grid_1d(work / 4, 256, h->sm_count), 256, 0, (cudaStream_t)stream>>>(p, ddst, dsrc, relu_gate);
  else
    k_crop_bwd<1><<<e2_grid_1d(work, 256, h->sm_count), 256, 0, (cudaStream_t)stream>>>(p, ddst, dsrc, relu_gate);
  e2_count_launch(h);
  E2_CUDA_CHECK(h, "crop_concat_bwd");
  return E2_OK;
}
