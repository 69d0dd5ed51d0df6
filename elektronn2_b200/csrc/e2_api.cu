// Handle management + layout conversion at the (b,f,z,x,y) boundary.
#include "e2_common.cuh"

extern "C" int e2_version(void) { return E2B200_VERSION; }

extern "C" int e2_create(e2_handle** out, int device) {
  if (!out) return E2_ERR_INVALID;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) return E2_ERR_CUDA;
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, device) != cudaSuccess) return E2_ERR_CUDA;
  e2_handle* h = new e2_handle();
  h->device = device;
  h->sm_count = p.multiProcessorCount;
  h->launches = 0;
  h->err[0] = 0;
  h->tmap_cache = nullptr;
  if (p.major != 10) {
    snprintf(h->err, sizeof(h->err), "libe2b200 is built for sm_100a only; device %d is sm_%d%d", device, p.major,
             p.minor);
    // the handle is still returned so the caller can read the message; every op will fail to launch.
  }
  *out = h;
  return E2_OK;
}

extern "C" int e2_destroy(e2_handle* h) {
  if (!h) return E2_ERR_INVALID;
  delete h;
  return E2_OK;
}

extern "C" const char* e2_last_error(const e2_handle* h) { return h ? h->err : "null handle"; }
extern "C" int64_t e2_launch_count(const e2_handle* h) { return h ? __atomic_load_n(&h->launches, __ATOMIC_RELAXED) : -1; }

// ---------------------------------------------------------------- NCDHW <-> NDHWC
// A [C][P] <-> [P][c_pitch] transpose per batch item through a padded 32x32 smem tile;
// both sides are read / written along their contiguous axis.
template <bool TO_CL>
__global__ void __launch_bounds__(256) k_layout(const float* __restrict__ src, float* __restrict__ dst, int C, int64_t P,
                                                int pitch) {
  __shared__ float tile[32][33];
  const int64_t p0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int n = blockIdx.z;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const float* s = src + (TO_CL ? (int64_t)n * C * P : (int64_t)n * P * pitch);
  float* d = dst + (TO_CL ? (int64_t)n * P * pitch : (int64_t)n * C * P);
  if (TO_CL) {
#pragma unroll
    for (int j = ty; j < 32; j += 8) {
      int c = c0 + j;
      int64_t p = p0 + tx;
      tile[j][tx] = (c < C && p < P) ? s[(int64_t)c * P + p] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int j = ty; j < 32; j += 8) {
      int64_t p = p0 + j;
      int c = c0 + tx;
      if (p < P && c < C) d[p * pitch + c] = tile[tx][j];
    }
  } else {
#pragma unroll
    for (int j = ty; j < 32; j += 8) {
      int64_t p = p0 + j;
      int c = c0 + tx;
      tile[j][tx] = (p < P && c < C) ? s[p * pitch + c] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int j = ty; j < 32; j += 8) {
      int c = c0 + j;
      int64_t p = p0 + tx;
      if (c < C && p < P) d[(int64_t)c * P + p] = tile[tx][j];
    }
  }
}

template <bool TO_CL>
static int layout_launch(e2_handle* h, const e2_tensor* t, const float* src, float* dst, void* stream) {
  E2_REQUIRE(h, e2_tensor_ok(t) && src && dst, "layout_convert: bad tensor descriptor or null pointer");
  int64_t P = (int64_t)t->z * t->x * t->y;
  dim3 grid((unsigned)((P + 31) / 32), (unsigned)((t->c + 31) / 32), (unsigned)t->n);
  E2_REQUIRE(h, grid.y <= 65535 && grid.z <= 65535, "layout_convert: too many channels / batch items");
  k_layout<TO_CL><<<grid, 256, 0, (cudaStream_t)stream>>>(src, dst, t->c, P, t->c_pitch);
  e2_count_launch(h);
  E2_CUDA_CHECK(h, "layout_convert");
  return E2_OK;
}

extern "C" int e2_ncdhw_to_ndhwc(e2_handle* h, const e2_tensor* t, const float* ncdhw, float* ndhwc, void* stream) {
  return layout_launch<true>(h, t, ncdhw, ndhwc, stream);
}
extern "C" int e2_ndhwc_to_ncdhw(e2_handle* h, const e2_tensor* t, const float* ndhwc, float* ncdhw, void* stream) {
  return layout_launch<false>(h, t, ndhwc, ncdhw, stream);
}

// ------------------------------------------------------------- channel re-pitch
// Copies (n,z,x,y,c) between two channel pitches, optionally rounding to tf32.  Used to hand a dense
// single-channel input to the tcgen05 path (TMA needs 16-byte rows): pad lanes are written as zero.
__global__ void __launch_bounds__(256) k_repitch(const float* __restrict__ src, float* __restrict__ dst, int64_t P, int C,
                                                 int sp, int dp, int round_tf32) {
  const int64_t total = P * dp;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pos = i / dp;
    const int c = (int)(i - pos * dp);
    float v = c < C ? src[pos * sp + c] : 0.f;
    if (round_tf32) v = e2_round_tf32(v);
    dst[i] = v;
  }
}

extern "C" int e2_repitch(e2_handle* h, const e2_tensor* src_t, const float* src, const e2_tensor* dst_t, float* dst,
                          int32_t round_tf32, void* stream) {
  E2_REQUIRE(h, e2_tensor_ok(src_t) && e2_tensor_ok(dst_t) && src && dst, "repitch: bad tensor descriptor or null pointer");
  E2_REQUIRE(h, src_t->n == dst_t->n && src_t->z == dst_t->z && src_t->x == dst_t->x && src_t->y == dst_t->y &&
                    src_t->c == dst_t->c, "repitch: geometry mismatch");
  const int64_t P = e2_positions(src_t);
  k_repitch<<<e2_grid_1d(P * dst_t->c_pitch, 256, h->sm_count), 256, 0, (cudaStream_t)stream>>>(
      src, dst, P, src_t->c, src_t->c_pitch, dst_t->c_pitch, round_tf32);
  e2_count_launch(h);
  E2_CUDA_CHECK(h, "repitch");
  return E2_OK;
}

// ------------------------------------------------------------- uint8 <-> float32
__global__ void __launch_bounds__(256) k_u8_to_f32(const uint8_t* __restrict__ s, float* __restrict__ d, int64_t n,
                                                   float scale) {
  int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
  for (; i < n; i += stride) {
    if (i + 3 < n && ((reinterpret_cast<uintptr_t>(s + i) & 3) == 0) && ((reinterpret_cast<uintptr_t>(d + i) & 15) == 0)) {
      uchar4 v = *reinterpret_cast<const uchar4*>(s + i);
      // the reference divides: as_floatX(raw) / 255 (node_basic.py:910); keep the division for bit parity
      float4 o = make_float4((float)v.x / scale, (float)v.y / scale, (float)v.z / scale, (float)v.w / scale);
      *reinterpret_cast<float4*>(d + i) = o;
    } else {
      for (int j = 0; j < 4 && i + j < n; ++j) d[i + j] = (float)s[i + j] / scale;
    }
  }
}

__global__ void __launch_bounds__(256) k_f32_to_u8(const float* __restrict__ s, uint8_t* __restrict__ d, int64_t n,
                                                   float scale) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    float v = s[i] * scale;  // prob *= 255, then C truncation on the uint8 store (node_basic.py:990-996)
    d[i] = (uint8_t)(int)v;
  }
}

extern "C" int e2_u8_to_f32(e2_handle* h, const uint8_t* src, float* dst, int64_t count, float scale, void* stream) {
  E2_REQUIRE(h, src && dst && count >= 0 && scale != 0.f, "u8_to_f32: bad arguments");
  if (count == 0) return E2_OK;
  k_u8_to_f32<<<e2_grid_1d((count + 3) / 4, 256, h->sm_count), 256, 0, (cudaStream_t)stream>>>(src, dst, count, scale);
  e2_count_launch(h);
  E2_CUDA_CHECK(h, "u8_to_f32");
  return E2_OK;
}

extern "C" int e2_f32_to_u8(e2_handle* h, const float* src, uint8_t* dst, int64_t count, float scale, void* stream) {
  E2_REQUIRE(h, src && dst && count >= 0, "f32_to_u8: bad arguments");
  if (count == 0) return E2_OK;
  k_f32_to_u8<<<e2_grid_1d(count, 256, h->sm_count), 256, 0, (cudaStream_t)stream>>>(src, dst, count, scale);
  e2_count_launch(h);
  E2_CUDA_CHECK(h, "f32_to_u8");
  return E2_OK;
}
