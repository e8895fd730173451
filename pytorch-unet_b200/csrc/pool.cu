// F.max_pool2d(x, 2) forward with arg-max codes, and its backward fused with the skip-connection gradient add
// and the producer's ReLU mask (reference: unet.py:79 forward; unet.py:152-163 + autograd for the backward).
// HBM-bound: one thread moves 8 channels (16 B) of one 2x2 window; NHWC makes every access a full 16-byte
// vector and consecutive threads cover consecutive channels, then consecutive pixels -> fully coalesced.
#include "common.cuh"
#include "ptx.cuh"

namespace b200 {

// ATen rule (aten/src/ATen/native/cpu/MaxPoolKernel.cpp semantics): scan (0,0),(0,1),(1,0),(1,1); replace when
// val > max or val is NaN.  Start: max = -inf, index = first element.
__device__ __forceinline__ void pool_scan(float v, int code, float& best, int& arg) {
  if (v > best || v != v) {
    best = v;
    arg = code;
  }
}

template <int VEC>
__global__ void __launch_bounds__(256)
maxpool_fwd_kernel(DView x, DView y, uint8_t* __restrict__ idx8, long long* __restrict__ idx64) {
  // one block per output row (n, oh): only 32-bit index arithmetic per element
  const int lanes = y.c / VEC;
  const int n = blockIdx.x / y.h, oh = blockIdx.x - n * y.h;
  for (int e = threadIdx.x; e < y.w * lanes; e += blockDim.x) {
    const int ow = e / lanes, l = e - ow * lanes;
    float v[4][VEC];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const long long so = x.off(n, 2 * oh + a, 2 * ow + b) + l * VEC;
        if (VEC == 8) {
          float t[8];
          load8s(x.p, x.lo, so, t);  // split tier: compare and forward hi + lo
#pragma unroll
          for (int j = 0; j < VEC; ++j) v[a * 2 + b][j] = t[j];
        } else {
          v[a * 2 + b][0] = bf2f(x.p[so]);
        }
      }
    float best[VEC];
    int arg[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      best[j] = -INFINITY;
      arg[j] = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) pool_scan(v[k][j], k, best[j], arg[j]);
    }
    const long long opix = ((long long)n * y.h + oh) * y.w + ow;
    const long long oo = y.off(n, oh, ow) + l * VEC;
    bf16* o = y.p + oo;
    if (VEC == 8) {
      float t[8];
#pragma unroll
      for (int j = 0; j < VEC; ++j) t[j] = best[j];
      store8s(y.p, y.lo, oo, t);
      uint2 codes;
      codes.x = arg[0] | (arg[1] << 8) | (arg[2] << 16) | (arg[3] << 24);
      codes.y = arg[4 % VEC] | (arg[5 % VEC] << 8) | (arg[6 % VEC] << 16) | (arg[7 % VEC] << 24);
      *reinterpret_cast<uint2*>(idx8 + opix * y.c + l * 8) = codes;
    } else {
      o[0] = f2bf(best[0]);
      idx8[opix * y.c + l] = (uint8_t)arg[0];
    }
    if (idx64) {
#pragma unroll
      for (int j = 0; j < VEC; ++j)
        idx64[opix * y.c + l * VEC + j] = (long long)(2 * oh + (arg[j] >> 1)) * x.w + (2 * ow + (arg[j] & 1));
    }
  }
}

// One thread per 2x2 window position of dx (windows also cover a trailing odd row / column, where only the
// skip-gradient term exists).
template <int VEC>
__global__ void __launch_bounds__(256)
maxpool_bwd_kernel(DView dy, const uint8_t* __restrict__ idx8, DView dx, DView add, int has_add, int add_y, int add_x,
                   const bf16* __restrict__ mask, const bf16* __restrict__ ypool) {
  // ypool (pre-masked mode, laid out like dy): the pooled activation.  The ReLU mask of the scattered term is then
  // [pooled value > 0] (the arg-max position holds exactly that value) and the skip-gradient window arrives already
  // masked by the kernel that produced it, so the full-resolution mask tensor is not read at all.
  // one block per window row (n, oh): only 32-bit index arithmetic per element
  const int lanes = dx.c / VEC;
  const int wh = (dx.h + 1) / 2, ww = (dx.w + 1) / 2;
  const int n = blockIdx.x / wh, oh = blockIdx.x - n * wh;
  for (int e = threadIdx.x; e < ww * lanes; e += blockDim.x) {
    const int ow = e / lanes, l = e - ow * lanes;
    const bool win = oh < dy.h && ow < dy.w;
    float g[VEC];
    int code[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      g[j] = 0.f;
      code[j] = -1;
    }
    if (win) {
      const long long opix = ((long long)n * dy.h + oh) * dy.w + ow;
      const bf16* s = dy.p + dy.off(n, oh, ow) + l * VEC;
      if (VEC == 8) {
        float t[8];
        unpack8(*reinterpret_cast<const bf16x8*>(s), t);
        const uint2 cw = *reinterpret_cast<const uint2*>(idx8 + opix * dy.c + l * 8);
        if (ypool) {
          float yv[8];
          unpack8(*reinterpret_cast<const bf16x8*>(ypool + dy.off(n, oh, ow) + l * 8), yv);
#pragma unroll
          for (int j = 0; j < VEC; ++j) t[j] = yv[j] > 0.f ? t[j] : 0.f;
        }
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          g[j] = t[j];
          code[j] = ((j < 4 ? cw.x : cw.y) >> (8 * (j & 3))) & 0xff;
        }
      } else {
        g[0] = bf2f(s[0]);
        if (ypool && !(bf2f(ypool[dy.off(n, oh, ow) + l]) > 0.f)) g[0] = 0.f;
        code[0] = idx8[opix * dy.c + l];
      }
    }
    if (VEC == 8) {
      // every load of the window (4 skip-gradient vectors + 4 mask vectors) is issued before the first use: with the
      // loads behind the per-position branches this kernel ran at 0.71 of the copy bandwidth
      bf16x8 av[4], mv[4];
      bool live[4], has_a[4];
      long long off[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int a = k >> 1, b = k & 1;
        const int ih = 2 * oh + a, iw = 2 * ow + b;
        live[k] = ih < dx.h && iw < dx.w;
        off[k] = dx.off(n, ih, iw) + l * 8;
        const int ay = ih - add_y, ax = iw - add_x;
        has_a[k] = live[k] && has_add && ay >= 0 && ay < add.h && ax >= 0 && ax < add.w;
        av[k] = make_uint4(0, 0, 0, 0);
        mv[k] = make_uint4(0, 0, 0, 0);
        if (has_a[k]) av[k] = *reinterpret_cast<const bf16x8*>(add.p + add.off(n, ay, ax) + l * 8);
        if (live[k] && mask) mv[k] = *reinterpret_cast<const bf16x8*>(mask + off[k]);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (!live[k]) continue;
        float r[8], t[8];
        unpack8(av[k], t);
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = ((code[j] == k) ? g[j] : 0.f) + t[j];
        bf16x8 o = pack8(r);
        if (mask) {
          o.x &= bf16x2_gt0_mask(mv[k].x);
          o.y &= bf16x2_gt0_mask(mv[k].y);
          o.z &= bf16x2_gt0_mask(mv[k].z);
          o.w &= bf16x2_gt0_mask(mv[k].w);
        }
        *reinterpret_cast<bf16x8*>(dx.p + off[k]) = o;
      }
    } else {
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          const int ih = 2 * oh + a, iw = 2 * ow + b;
          if (ih >= dx.h || iw >= dx.w) continue;
          float r = (code[0] == a * 2 + b) ? g[0] : 0.f;
          const int ay = ih - add_y, ax = iw - add_x;
          if (has_add && ay >= 0 && ay < add.h && ax >= 0 && ax < add.w) r += bf2f(add.p[add.off(n, ay, ax) + l]);
          const long long o = dx.off(n, ih, iw) + l;
          if (mask) r = bf2f(mask[o]) > 0.f ? r : 0.f;
          dx.p[o] = f2bf(r);
        }
    }
  }
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200unet_maxpool2x2_fwd(const b200_view* x, const b200_view* y, uint8_t* idx8, int64_t* idx64, void* stream) {
  B200_REQUIRE(view_ok(x) && view_ok(y) && idx8, "maxpool_fwd: bad arguments");
  B200_REQUIRE(y->n == x->n && y->c == x->c && y->h == x->h / 2 && y->w == x->w / 2,
               "maxpool_fwd: output extent must be floor(input/2)");
  const bool v8 = vec8_ok(*x) && vec8_ok(*y) && reinterpret_cast<uintptr_t>(idx8) % 8 == 0;
  B200_REQUIRE((x->lo == nullptr) == (y->lo == nullptr), "maxpool_fwd: x and y must be of the same precision tier");
  B200_REQUIRE(v8 || !x->lo, "maxpool_fwd: the split tier needs channel counts / strides that are multiples of 8");
  if (v8)
    maxpool_fwd_kernel<8><<<(unsigned)(y->n * y->h), 256, 0, as_stream(stream)>>>(dview(*x), dview(*y), idx8,
                                                                            (long long*)idx64);
  else
    maxpool_fwd_kernel<1><<<(unsigned)(y->n * y->h), 256, 0, as_stream(stream)>>>(dview(*x), dview(*y), idx8,
                                                                            (long long*)idx64);
  return check_launch("maxpool_fwd");
}

static int maxpool_bwd_launch(const b200_view* dy, const uint8_t* idx8, const b200_view* dx, const b200_view* add, int add_y,
                              int add_x, const void* mask, const b200_view* ypool, void* stream) {
  B200_REQUIRE(view_ok(dy) && view_ok(dx) && idx8, "maxpool_bwd: bad arguments");
  B200_REQUIRE(dy->n == dx->n && dy->c == dx->c && dy->h == dx->h / 2 && dy->w == dx->w / 2,
               "maxpool_bwd: dy extent must be floor(dx/2)");
  if (add) {
    B200_REQUIRE(view_ok(add) && add->n == dx->n && add->c == dx->c && add_y >= 0 && add_x >= 0 &&
                     add_y + add->h <= dx->h && add_x + add->w <= dx->w,
                 "maxpool_bwd: skip-gradient window outside dx");
  }
  if (ypool) {
    B200_REQUIRE(view_ok(ypool) && same_extent(*ypool, *dy) && ypool->stride_n == dy->stride_n &&
                     ypool->stride_h == dy->stride_h && ypool->stride_w == dy->stride_w,
                 "maxpool_bwd: the pooled activation must be laid out like dy");
  }
  const bool v8 = vec8_ok(*dy) && vec8_ok(*dx) && (!add || vec8_ok(*add)) &&
                  reinterpret_cast<uintptr_t>(idx8) % 8 == 0 && reinterpret_cast<uintptr_t>(mask) % 16 == 0 &&
                  (!ypool || reinterpret_cast<uintptr_t>(ypool->ptr) % 16 == 0);
  DView dadd = add ? dview(*add) : dview(*dx);
  const bf16* yp = ypool ? (const bf16*)ypool->ptr : nullptr;
  if (v8)
    maxpool_bwd_kernel<8><<<(unsigned)(dx->n * ((dx->h + 1) / 2)), 256, 0, as_stream(stream)>>>(dview(*dy), idx8, dview(*dx), dadd,
                                                                            add ? 1 : 0, add_y, add_x, (const bf16*)mask, yp);
  else
    maxpool_bwd_kernel<1><<<(unsigned)(dx->n * ((dx->h + 1) / 2)), 256, 0, as_stream(stream)>>>(dview(*dy), idx8, dview(*dx), dadd,
                                                                            add ? 1 : 0, add_y, add_x, (const bf16*)mask, yp);
  return check_launch("maxpool_bwd");
}

int b200unet_maxpool2x2_bwd(const b200_view* dy, const uint8_t* idx8, const b200_view* dx, const b200_view* add,
                            int add_y, int add_x, const void* mask, void* stream) {
  return maxpool_bwd_launch(dy, idx8, dx, add, add_y, add_x, mask, nullptr, stream);
}

int b200unet_maxpool2x2_bwd_premasked(const b200_view* dy, const uint8_t* idx8, const b200_view* y_pooled, const b200_view* dx,
                                      const b200_view* add, int add_y, int add_x, void* stream) {
  B200_REQUIRE(y_pooled, "maxpool_bwd_premasked: the pooled activation is required");
  return maxpool_bwd_launch(dy, idx8, dx, add, add_y, add_x, nullptr, y_pooled, stream);
}
}
