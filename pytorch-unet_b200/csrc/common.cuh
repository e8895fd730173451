// Shared host/device helpers of libb200unet: error reporting, view validation, small device utilities.
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/b200unet.h"

namespace b200 {

typedef __nv_bfloat16 bf16;

// ------------------------------------------------------------------ errors (thread-local message)
char* err_buf();
int fail(int code, const char* fmt, ...);
int check_launch(const char* what);

#define B200_REQUIRE(cond, ...)                     \
  do {                                              \
    if (!(cond)) return ::b200::fail(-1, __VA_ARGS__); \
  } while (0)

inline bool view_ok(const b200_view* v) {
  return v && v->ptr && v->n > 0 && v->h > 0 && v->w > 0 && v->c > 0 && v->stride_w >= v->c;
}
inline bool same_extent(const b200_view& a, const b200_view& b) {
  return a.n == b.n && a.h == b.h && a.w == b.w && a.c == b.c;
}
inline int64_t view_pixels(const b200_view& v) { return (int64_t)v.n * v.h * v.w; }

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

constexpr int kNumSMsB200 = 148;

// ------------------------------------------------------------------ device view (POD copy usable in kernels)
struct DView {
  bf16* p;
  bf16* lo;  // split tier: low-order plane (same geometry), else nullptr
  int n, h, w, c;
  long long sn, sh, sw;
  __host__ __device__ long long off(int in, int ih, int iw) const { return in * sn + ih * sh + iw * sw; }
};
inline DView dview(const b200_view& v) {
  return DView{reinterpret_cast<bf16*>(v.ptr), reinterpret_cast<bf16*>(v.lo), v.n, v.h, v.w, v.c,
               v.stride_n, v.stride_h, v.stride_w};
}

// 16-byte vector path is usable on this view
inline bool vec8_ok(const b200_view& v) {
  return v.c % 8 == 0 && reinterpret_cast<uintptr_t>(v.ptr) % 16 == 0 && reinterpret_cast<uintptr_t>(v.lo) % 16 == 0 &&
         v.stride_w % 8 == 0 && v.stride_h % 8 == 0 && v.stride_n % 8 == 0;
}
inline int stream_grid(long long total) {
  long long g = (total + 255) / 256;
  const long long cap = (long long)kNumSMsB200 * 8;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

__device__ __forceinline__ float bf2f(bf16 v) { return __bfloat162float(v); }
__device__ __forceinline__ bf16 f2bf(float v) { return __float2bfloat16_rn(v); }

// 8 bf16 <-> 8 floats through one 16-byte access
// (a plain uint4 so that the compiler emits one LDG.128 / STG.128; a struct of four bfloat162 is copied word by word)
typedef uint4 bf16x8;
__device__ __forceinline__ float2 bf2x_to_f2(uint32_t w) {
  // bf16 -> fp32 is a 16-bit shift
  return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
__device__ __forceinline__ uint32_t f2_to_bf2x(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ void unpack8(const bf16x8 p, float (&f)[8]) {
  float2 t;
  t = bf2x_to_f2(p.x); f[0] = t.x; f[1] = t.y;
  t = bf2x_to_f2(p.y); f[2] = t.x; f[3] = t.y;
  t = bf2x_to_f2(p.z); f[4] = t.x; f[5] = t.y;
  t = bf2x_to_f2(p.w); f[6] = t.x; f[7] = t.y;
}
__device__ __forceinline__ bf16x8 pack8(const float (&f)[8]) {
  return make_uint4(f2_to_bf2x(f[0], f[1]), f2_to_bf2x(f[2], f[3]), f2_to_bf2x(f[4], f[5]), f2_to_bf2x(f[6], f[7]));
}

// Split tier (b200unet.h): value = hi + lo.  The sum of the two planes is exact in fp32 (8 + 8 significant bits).
__device__ __forceinline__ void load8s(const bf16* hi, const bf16* lo, long long off, float (&f)[8]) {
  unpack8(*reinterpret_cast<const bf16x8*>(hi + off), f);
  if (lo) {
    float t[8];
    unpack8(*reinterpret_cast<const bf16x8*>(lo + off), t);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] += t[j];
  }
}
__device__ __forceinline__ float split_lo(float v, float hi) {  // residual plane; 0 where hi is inf / NaN
  return fabsf(hi) < INFINITY ? v - hi : 0.f;
}
__device__ __forceinline__ void store8s(bf16* hi, bf16* lo, long long off, const float (&f)[8]) {
  const bf16x8 h = pack8(f);
  *reinterpret_cast<bf16x8*>(hi + off) = h;
  if (lo) {
    float t[8];
    unpack8(h, t);
#pragma unroll
    for (int j = 0; j < 8; ++j) t[j] = split_lo(f[j], t[j]);
    *reinterpret_cast<bf16x8*>(lo + off) = pack8(t);
  }
}

// packed fp32x2 FMA (FFMA2, sm_100): d = a * b + c on both halves — halves the FMA instruction count of the CUDA-core
// kernels that are bound by instruction issue
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                     rc = *reinterpret_cast<unsigned long long*>(&c), rd;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace b200
