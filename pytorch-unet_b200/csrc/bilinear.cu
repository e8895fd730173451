// nn.Upsample(scale_factor=2, mode='bilinear', align_corners=False) forward / backward on NHWC bf16
// (reference: unet.py:146, 176).  HBM-bound; 8 channels (16 B) per thread.
//
// Per axis (ATen area_pixel_compute_source_index with scale 0.5, clamp at 0):
//   out[2i]   = 0.25*in[i-1] + 0.75*in[i]   (out[0] = in[0])
//   out[2i+1] = 0.75*in[i]   + 0.25*in[min(i+1, L-1)]
#include "common.cuh"

namespace b200 {

__device__ __forceinline__ void src_taps(int o, int len, int& i0, int& i1, float& w0, float& w1) {
  float s = (o + 0.5f) * 0.5f - 0.5f;
  if (s < 0.f) s = 0.f;
  i0 = (int)s;
  const float lam = s - (float)i0;
  i1 = i0 + (i0 < len - 1 ? 1 : 0);
  w0 = 1.f - lam;
  w1 = lam;
}

template <int VEC>
__device__ __forceinline__ void load_vec(const bf16* p, float (&f)[VEC]) {
  if (VEC == 8) {
    float t[8];
    unpack8(*reinterpret_cast<const bf16x8*>(p), t);
#pragma unroll
    for (int j = 0; j < VEC; ++j) f[j] = t[j];
  } else {
    f[0] = bf2f(p[0]);
  }
}
// split tier (b200unet.h): value = hi + lo; lo == nullptr in the bf16 tier
template <int VEC>
__device__ __forceinline__ void load_vec_s(const bf16* hi, const bf16* lo, long long off, float (&f)[VEC]) {
  load_vec<VEC>(hi + off, f);
  if (VEC == 8 && lo) {
    float t[VEC];
    load_vec<VEC>(lo + off, t);
#pragma unroll
    for (int j = 0; j < VEC; ++j) f[j] += t[j];
  }
}
template <int VEC>
__device__ __forceinline__ void store_vec(bf16* p, const float (&f)[VEC]) {
  if (VEC == 8) {
    float t[8];
#pragma unroll
    for (int j = 0; j < VEC; ++j) t[j] = f[j];
    *reinterpret_cast<bf16x8*>(p) = pack8(t);
  } else {
    p[0] = f2bf(f[0]);
  }
}

template <int VEC>
__global__ void __launch_bounds__(256) bilinear_fwd_kernel(DView x, DView y) {
  const int lanes = y.c / VEC;
  const long long total = (long long)y.n * y.h * y.w * lanes;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int l = (int)(e % lanes);
    long long p = e / lanes;
    const int ow = (int)(p % y.w);
    p /= y.w;
    const int oh = (int)(p % y.h);
    const int n = (int)(p / y.h);
    int h0, h1, w0, w1;
    float a0, a1, b0, b1;
    src_taps(oh, x.h, h0, h1, a0, a1);
    src_taps(ow, x.w, w0, w1, b0, b1);
    float v00[VEC], v01[VEC], v10[VEC], v11[VEC], r[VEC];
    load_vec_s<VEC>(x.p, x.lo, x.off(n, h0, w0) + l * VEC, v00);
    load_vec_s<VEC>(x.p, x.lo, x.off(n, h0, w1) + l * VEC, v01);
    load_vec_s<VEC>(x.p, x.lo, x.off(n, h1, w0) + l * VEC, v10);
    load_vec_s<VEC>(x.p, x.lo, x.off(n, h1, w1) + l * VEC, v11);
#pragma unroll
    for (int j = 0; j < VEC; ++j) r[j] = a0 * (b0 * v00[j] + b1 * v01[j]) + a1 * (b0 * v10[j] + b1 * v11[j]);
    const long long oo = y.off(n, oh, ow) + l * VEC;
    store_vec<VEC>(y.p + oo, r);
    if (VEC == 8 && y.lo) {
      float t[VEC];
#pragma unroll
      for (int j = 0; j < VEC; ++j) t[j] = split_lo(r[j], bf2f(f2bf(r[j])));
      store_vec<VEC>(y.lo + oo, t);
    }
  }
}

// Transpose of the forward: input pixel (i, j) gathers from the <= 4x4 outputs that read it.
template <int VEC>
__global__ void __launch_bounds__(256) bilinear_bwd_kernel(DView dy, DView dx, const bf16* __restrict__ mask) {
  const int lanes = dx.c / VEC;
  const long long total = (long long)dx.n * dx.h * dx.w * lanes;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int l = (int)(e % lanes);
    long long p = e / lanes;
    const int iw = (int)(p % dx.w);
    p /= dx.w;
    const int ih = (int)(p % dx.h);
    const int n = (int)(p / dx.h);
    float wy[4], wx[4];
    int oy[4], ox[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      oy[k] = 2 * ih - 1 + k;
      ox[k] = 2 * iw - 1 + k;
      wy[k] = 0.f;
      wx[k] = 0.f;
      if (oy[k] >= 0 && oy[k] < dy.h) {
        int i0, i1;
        float a0, a1;
        src_taps(oy[k], dx.h, i0, i1, a0, a1);
        wy[k] = (i0 == ih ? a0 : 0.f) + (i1 == ih ? a1 : 0.f);
      }
      if (ox[k] >= 0 && ox[k] < dy.w) {
        int i0, i1;
        float a0, a1;
        src_taps(ox[k], dx.w, i0, i1, a0, a1);
        wx[k] = (i0 == iw ? a0 : 0.f) + (i1 == iw ? a1 : 0.f);
      }
    }
    float r[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) r[j] = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      if (wy[a] == 0.f) continue;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        if (wx[b] == 0.f) continue;
        float g[VEC];
        load_vec<VEC>(dy.p + dy.off(n, oy[a], ox[b]) + l * VEC, g);
        const float wgt = wy[a] * wx[b];
#pragma unroll
        for (int j = 0; j < VEC; ++j) r[j] += wgt * g[j];
      }
    }
    const long long o = dx.off(n, ih, iw) + l * VEC;
    if (mask) {
      float m[VEC];
      load_vec<VEC>(mask + o, m);
#pragma unroll
      for (int j = 0; j < VEC; ++j) r[j] = m[j] > 0.f ? r[j] : 0.f;
    }
    store_vec<VEC>(dx.p + o, r);
  }
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200unet_bilinear_up2x_fwd(const b200_view* x, const b200_view* y, void* stream) {
  B200_REQUIRE(view_ok(x) && view_ok(y), "bilinear_fwd: bad arguments");
  B200_REQUIRE(y->n == x->n && y->c == x->c && y->h == 2 * x->h && y->w == 2 * x->w,
               "bilinear_fwd: output extent must be 2x the input");
  const bool v8 = vec8_ok(*x) && vec8_ok(*y);
  B200_REQUIRE((x->lo == nullptr) == (y->lo == nullptr), "bilinear_fwd: x and y must be of the same precision tier");
  B200_REQUIRE(v8 || !x->lo, "bilinear_fwd: the split tier needs channel counts / strides that are multiples of 8");
  const long long total = view_pixels(*y) * (v8 ? y->c / 8 : y->c);
  if (v8)
    bilinear_fwd_kernel<8><<<stream_grid(total), 256, 0, as_stream(stream)>>>(dview(*x), dview(*y));
  else
    bilinear_fwd_kernel<1><<<stream_grid(total), 256, 0, as_stream(stream)>>>(dview(*x), dview(*y));
  return check_launch("bilinear_fwd");
}

int b200unet_bilinear_up2x_bwd(const b200_view* dy, const b200_view* dx, const void* mask, void* stream) {
  B200_REQUIRE(view_ok(dy) && view_ok(dx), "bilinear_bwd: bad arguments");
  B200_REQUIRE(dy->n == dx->n && dy->c == dx->c && dy->h == 2 * dx->h && dy->w == 2 * dx->w,
               "bilinear_bwd: dy extent must be 2x dx");
  const bool v8 = vec8_ok(*dy) && vec8_ok(*dx) && reinterpret_cast<uintptr_t>(mask) % 16 == 0;
  const long long total = view_pixels(*dx) * (v8 ? dx->c / 8 : dx->c);
  if (v8)
    bilinear_bwd_kernel<8><<<stream_grid(total), 256, 0, as_stream(stream)>>>(dview(*dy), dview(*dx), (const bf16*)mask);
  else
    bilinear_bwd_kernel<1><<<stream_grid(total), 256, 0, as_stream(stream)>>>(dview(*dy), dview(*dx), (const bf16*)mask);
  return check_launch("bilinear_bwd");
}
}
