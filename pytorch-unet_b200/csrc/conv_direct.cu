// CUDA-core ("direct") implementation of the convolution family.  It serves the layers the tensor-core kernels
// cannot take (channel counts that are not multiples of 8: the 1/3-channel input image, the 4-channel levels of
// the feature-extraction variant) and is the in-library cross-check for them.  One generic gather kernel covers
// conv 3x3 / 1x1 forward, their backward-data (flipped, transposed weights) and ConvTranspose2d backward-data
// (2x2 taps, input stride 2); one generic reduction kernel covers every backward-weights pass.
#include "chan_reduce.cuh"
#include "conv_impl.h"

namespace b200 {

constexpr int kDirectThreads = 256;
constexpr int kCinChunk = 8;

struct GatherArgs {
  DView src[2];
  int nsrc;
  int ky, kx, pad, in_stride;
  const float* w;
  long long w_so, w_sc;  // weight index = o * w_so + c * w_sc + tap_index * w_st
  int w_st, w_flip;
  const float* bias;
  int relu;
  DView dst[2];
  int ndst;
  const bf16* mask[2];
  int out_n, out_h, out_w, cout;
};

// thread = (pixel slot, group of 8 output channels); 8 lanes cover 64 output channels of one pixel so that a warp
// stores 4 pixels x 128 contiguous bytes.
__global__ void __launch_bounds__(kDirectThreads) gather_conv_kernel(GatherArgs a) {
  __shared__ __align__(16) float ws[9 * kCinChunk * 64];
  const int g = threadIdx.x & 7;
  const int slot = threadIdx.x >> 3;
  const int o_base = blockIdx.y * 64;
  const int o0 = o_base + g * 8;
  const long long npix = (long long)a.out_n * a.out_h * a.out_w;
  const long long p = (long long)blockIdx.x * (kDirectThreads / 8) + slot;
  const bool live = p < npix;
  int n = 0, oh = 0, ow = 0;
  if (live) {
    const long long hw = (long long)a.out_h * a.out_w;
    n = (int)(p / hw);
    const long long r = p - n * hw;
    oh = (int)(r / a.out_w);
    ow = (int)(r - (long long)oh * a.out_w);
  }
  const int taps = a.ky * a.kx;
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;

  int cbase = 0;
  for (int s = 0; s < a.nsrc; ++s) {
    const DView& src = a.src[s];
    const bool vec = (src.c % 8 == 0) && (reinterpret_cast<uintptr_t>(src.p) % 16 == 0) && (src.sw % 8 == 0) &&
                     (src.sh % 8 == 0) && (src.sn % 8 == 0);
    for (int c0 = 0; c0 < src.c; c0 += kCinChunk) {
      const int cc_n = min(kCinChunk, src.c - c0);
      __syncthreads();
      for (int i = threadIdx.x; i < taps * kCinChunk * 64; i += kDirectThreads) {
        const int o = i & 63;
        const int cc = (i >> 6) % kCinChunk;
        const int t = i / (64 * kCinChunk);
        float v = 0.f;
        if (cc < cc_n && o_base + o < a.cout) {
          const int ti = a.w_flip ? (taps - 1 - t) : t;
          v = a.w[(long long)(o_base + o) * a.w_so + (long long)(cbase + c0 + cc) * a.w_sc + (long long)ti * a.w_st];
        }
        ws[i] = v;
      }
      __syncthreads();
      if (live && o0 < a.cout) {
        for (int t = 0; t < taps; ++t) {
          const int ty = t / a.kx, tx = t - ty * a.kx;
          const int iy = oh * a.in_stride + ty - a.pad, ix = ow * a.in_stride + tx - a.pad;
          if (iy < 0 || iy >= src.h || ix < 0 || ix >= src.w) continue;
          const bf16* xp = src.p + src.off(n, iy, ix) + c0;
          float xv[kCinChunk];
          if (vec) {
            unpack8(*reinterpret_cast<const bf16x8*>(xp), xv);
          } else {
#pragma unroll
            for (int cc = 0; cc < kCinChunk; ++cc) xv[cc] = cc < cc_n ? bf2f(xp[cc]) : 0.f;
          }
          const float* wt = ws + (t * kCinChunk) * 64 + g * 8;
#pragma unroll
          for (int cc = 0; cc < kCinChunk; ++cc) {
            const float4 w0 = *reinterpret_cast<const float4*>(wt + cc * 64);
            const float4 w1 = *reinterpret_cast<const float4*>(wt + cc * 64 + 4);
            acc[0] += xv[cc] * w0.x;
            acc[1] += xv[cc] * w0.y;
            acc[2] += xv[cc] * w0.z;
            acc[3] += xv[cc] * w0.w;
            acc[4] += xv[cc] * w1.x;
            acc[5] += xv[cc] * w1.y;
            acc[6] += xv[cc] * w1.z;
            acc[7] += xv[cc] * w1.w;
          }
        }
      }
    }
    cbase += src.c;
  }
  if (!live || o0 >= a.cout) return;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if (a.bias && o0 + k < a.cout) acc[k] += a.bias[o0 + k];
    if (a.relu) acc[k] = fmaxf(acc[k], 0.f);
  }
  // destination split by output channel
  const int c_first = a.dst[0].c;
  const int d = (a.ndst > 1 && o0 >= c_first) ? 1 : 0;
  const DView& dst = a.dst[d];
  const int oc = o0 - (d ? c_first : 0);
  const long long off = dst.off(n, oh, ow) + oc;
  const bf16* mk = a.mask[d];
  const bool whole = (oc + 8 <= dst.c) && (dst.c % 8 == 0) && (reinterpret_cast<uintptr_t>(dst.p) % 16 == 0) &&
                     (dst.sw % 8 == 0) && (dst.sh % 8 == 0) && (dst.sn % 8 == 0) &&
                     (reinterpret_cast<uintptr_t>(mk) % 16 == 0) && (a.ndst == 1 || c_first % 8 == 0);
  if (whole) {
    if (mk) {
      float m[8];
      unpack8(*reinterpret_cast<const bf16x8*>(mk + off), m);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = m[k] > 0.f ? acc[k] : 0.f;
    }
    *reinterpret_cast<bf16x8*>(dst.p + off) = pack8(acc);
  } else {
    // scalar tail; a group may straddle the destination boundary when dst[0].c is not a multiple of 8
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int o = o0 + k;
      if (o >= a.cout) break;
      const int dd = (a.ndst > 1 && o >= c_first) ? 1 : 0;
      const DView& ds = a.dst[dd];
      const long long of = ds.off(n, oh, ow) + (o - (dd ? c_first : 0));
      float v = acc[k];
      if (a.mask[dd]) v = bf2f(a.mask[dd][of]) > 0.f ? v : 0.f;
      ds.p[of] = f2bf(v);
    }
  }
}

static int launch_gather(const GatherArgs& a, cudaStream_t st) {
  const long long npix = (long long)a.out_n * a.out_h * a.out_w;
  dim3 grid((unsigned)((npix + kDirectThreads / 8 - 1) / (kDirectThreads / 8)), (unsigned)((a.cout + 63) / 64));
  gather_conv_kernel<<<grid, kDirectThreads, 0, st>>>(a);
  return check_launch("gather_conv");
}

// ------------------------------------------------------------------ backward-weights
struct WgradArgs {
  DView dz;  // "rows": n, h, w, cout
  DView src[2];
  int nsrc;
  int ky, kx, pad, in_stride;
  int cin_total, cout;
  long long o_so, o_sc;  // output index = o * o_so + c * o_sc + tap
  int splits;
  float* partial;  // [splits][cout * cin_total * taps]
};

// warp = 8 consecutive output channels, lane = one (input channel, tap) pair; block = 64 couts x 32 pairs.
__global__ void __launch_bounds__(kDirectThreads) wgrad_direct_kernel(WgradArgs a) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int taps = a.ky * a.kx;
  const int q = blockIdx.y * 32 + lane;  // (c, tap)
  const int o0 = blockIdx.z * 64 + warp * 8;
  const bool qlive = q < a.cin_total * taps;
  int c = 0, t = 0, s = 0;
  if (qlive) {
    c = q / taps;
    t = q - c * taps;
    if (a.nsrc > 1 && c >= a.src[0].c) {
      s = 1;
    }
  }
  const DView& src = a.src[s];
  const int cs = c - (s ? a.src[0].c : 0);
  const int ty = t / a.kx, tx = t - ty * a.kx;
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  const long long npix = (long long)a.dz.n * a.dz.h * a.dz.w;
  const long long per = (npix + a.splits - 1) / a.splits;
  const long long p_begin = (long long)blockIdx.x * per;
  const long long p_end = min(npix, p_begin + per);
  const long long hw = (long long)a.dz.h * a.dz.w;
  const bool zvec = (a.dz.c % 8 == 0) && (reinterpret_cast<uintptr_t>(a.dz.p) % 16 == 0) && (a.dz.sw % 8 == 0) &&
                    (a.dz.sh % 8 == 0) && (a.dz.sn % 8 == 0);
  if (qlive && o0 < a.cout) {
    for (long long p = p_begin; p < p_end; ++p) {
      const int n = (int)(p / hw);
      const long long r = p - n * hw;
      const int oh = (int)(r / a.dz.w), ow = (int)(r - (long long)oh * a.dz.w);
      const int iy = oh * a.in_stride + ty - a.pad, ix = ow * a.in_stride + tx - a.pad;
      if (iy < 0 || iy >= src.h || ix < 0 || ix >= src.w) continue;
      const float xv = bf2f(src.p[src.off(n, iy, ix) + cs]);
      const bf16* zp = a.dz.p + a.dz.off(n, oh, ow) + o0;
      float zv[8];
      if (zvec && o0 + 8 <= a.cout) {
        unpack8(*reinterpret_cast<const bf16x8*>(zp), zv);
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) zv[k] = (o0 + k < a.cout) ? bf2f(zp[k]) : 0.f;
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += zv[k] * xv;
    }
    float* out = a.partial + (long long)blockIdx.x * a.cout * a.cin_total * taps;
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (o0 + k < a.cout) out[(long long)(o0 + k) * a.o_so + (long long)c * a.o_sc + t] = acc[k];
  }
}

__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, int splits, long long total, float* __restrict__ dw) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= total) return;
  float s = 0.f;
  for (int i = 0; i < splits; ++i) s += partial[(long long)i * total + e];
  dw[e] = s;
}

struct SumF2 {
  template <int VEC>
  __device__ void operator()(const float (&a)[VEC], const float (&)[VEC], float (&acc)[1][VEC], int) const {
#pragma unroll
    for (int j = 0; j < VEC; ++j) acc[0][j] += a[j];
  }
};
__global__ void finalize_sum2_kernel(const float* __restrict__ partial, int blocks, int c, float* __restrict__ out) {
  const int ch = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (ch >= c) return;
  const double s = warp_partial_sum(partial, blocks, c, ch);
  if ((threadIdx.x & 31) == 0) out[ch] = (float)s;
}

int bias_grad(const b200_view& dz, float* db, void* ws, cudaStream_t st) {
  ReducePlan pl;
  int r = launch_chan_reduce<1, false>(SumF2(), dz, nullptr, (float*)ws, &pl, st);
  if (r) return r;
  finalize_sum2_kernel<<<finalize_grid(dz.c), kFinalizeThreads, 0, st>>>((const float*)ws, pl.blocks, dz.c, db);
  return check_launch("bias_grad");
}

constexpr size_t kWgradPartialBudget = (size_t)64 << 20;

static int wgrad_splits(long long npix, long long outputs) {
  long long s = (npix + 255) / 256;
  if (s > 2 * kNumSMsB200) s = 2 * kNumSMsB200;
  const long long cap = (long long)(kWgradPartialBudget / (outputs * sizeof(float)));
  if (s > cap) s = cap;
  if (s < 1) s = 1;
  return (int)s;
}

size_t direct_wgrad_workspace(long long npix, int cout, int cin_total, int taps) {
  const long long outputs = (long long)cout * cin_total * taps;
  return (size_t)wgrad_splits(npix, outputs) * outputs * sizeof(float) + reduce_workspace_bytes(cout, 1);
}

static int launch_wgrad(WgradArgs a, float* dw, float* db, const b200_view& dzv, void* ws, size_t ws_bytes,
                        cudaStream_t st) {
  const int taps = a.ky * a.kx;
  const long long outputs = (long long)a.cout * a.cin_total * taps;
  const long long npix = (long long)a.dz.n * a.dz.h * a.dz.w;
  a.splits = wgrad_splits(npix, outputs);
  const size_t need = (size_t)a.splits * outputs * sizeof(float) + reduce_workspace_bytes(a.cout, 1);
  if (!ws || ws_bytes < need) return fail(-1, "wgrad (direct): workspace too small (%zu < %zu)", ws_bytes, need);
  a.partial = (float*)ws;
  dim3 grid(a.splits, (unsigned)((a.cin_total * taps + 31) / 32), (unsigned)((a.cout + 63) / 64));
  // lanes that are dead in the kernel leave their partial slot untouched: cover every slot
  cudaMemsetAsync(a.partial, 0, (size_t)a.splits * outputs * sizeof(float), st);
  wgrad_direct_kernel<<<grid, kDirectThreads, 0, st>>>(a);
  int r = check_launch("wgrad_direct");
  if (r) return r;
  wgrad_reduce_kernel<<<(unsigned)((outputs + 255) / 256), 256, 0, st>>>(a.partial, a.splits, outputs, dw);
  r = check_launch("wgrad_reduce");
  if (r) return r;
  if (db) return bias_grad(dzv, db, (char*)ws + (size_t)a.splits * outputs * sizeof(float), st);
  return 0;
}

// ------------------------------------------------------------------ entry points used by conv_api.cu
int direct_conv_fwd(const b200_conv_fwd_params* p, cudaStream_t st) {
  if (!p->w_f32) return fail(-1, "conv_fwd (direct): w_f32 is required");
  GatherArgs a{};
  a.nsrc = p->num_src;
  for (int i = 0; i < p->num_src; ++i) a.src[i] = dview(p->src[i]);
  const int k = p->taps == 9 ? 3 : 1;
  a.ky = a.kx = k;
  a.pad = p->pad;
  a.in_stride = 1;
  int cin = 0;
  for (int i = 0; i < p->num_src; ++i) cin += p->src[i].c;
  a.w = p->w_f32;
  a.w_so = (long long)cin * p->taps;
  a.w_sc = p->taps;
  a.w_st = 1;
  a.w_flip = 0;
  a.bias = p->bias;
  a.relu = p->relu;
  a.dst[0] = dview(p->dst);
  a.ndst = 1;
  a.out_n = p->dst.n;
  a.out_h = p->dst.h;
  a.out_w = p->dst.w;
  a.cout = p->dst.c;
  return launch_gather(a, st);
}

int direct_conv_dgrad(const b200_conv_dgrad_params* p, cudaStream_t st) {
  if (!p->w_f32) return fail(-1, "conv_dgrad (direct): w_f32 is required");
  GatherArgs a{};
  a.nsrc = 1;
  a.src[0] = dview(p->dz);
  const int k = p->taps == 9 ? 3 : 1;
  a.ky = a.kx = k;
  a.pad = k - 1 - p->pad;
  a.in_stride = 1;
  int cin = 0;
  for (int i = 0; i < p->num_dst; ++i) cin += p->dst[i].c;
  a.w = p->w_f32;  // [cout][cin][taps]: "output channel" of this gather is the forward input channel
  a.w_so = p->taps;
  a.w_sc = (long long)cin * p->taps;
  a.w_st = 1;
  a.w_flip = 1;
  a.ndst = p->num_dst;
  for (int i = 0; i < p->num_dst; ++i) {
    a.dst[i] = dview(p->dst[i]);
    a.mask[i] = (const bf16*)p->mask[i];
  }
  a.out_n = p->dst[0].n;
  a.out_h = p->dst[0].h;
  a.out_w = p->dst[0].w;
  a.cout = cin;
  return launch_gather(a, st);
}

int direct_conv_wgrad(const b200_conv_wgrad_params* p, void* ws, size_t ws_bytes, cudaStream_t st) {
  WgradArgs a{};
  a.dz = dview(p->dz);
  a.nsrc = p->num_src;
  int cin = 0;
  for (int i = 0; i < p->num_src; ++i) {
    a.src[i] = dview(p->src[i]);
    cin += p->src[i].c;
  }
  const int k = p->taps == 9 ? 3 : 1;
  a.ky = a.kx = k;
  a.pad = p->pad;
  a.in_stride = 1;
  a.cin_total = cin;
  a.cout = p->dz.c;
  a.o_so = (long long)cin * p->taps;
  a.o_sc = p->taps;
  return launch_wgrad(a, p->dw_f32, p->db_f32, p->dz, ws, ws_bytes, st);
}

// ConvTranspose2d forward: four 1x1 convolutions, one per (a, b), each writing the stride-2 sub-lattice of y.
int direct_convt_fwd(const b200_convt_fwd_params* p, cudaStream_t st) {
  if (!p->w_f32) return fail(-1, "convt_fwd (direct): w_f32 is required");
  for (int ab = 0; ab < 4; ++ab) {
    const int ay = ab >> 1, ax = ab & 1;
    GatherArgs a{};
    a.nsrc = 1;
    a.src[0] = dview(p->x);
    a.ky = a.kx = 1;
    a.pad = 0;
    a.in_stride = 1;
    a.w = p->w_f32 + ab;  // [cin][cout][2][2]
    a.w_so = 4;
    a.w_sc = (long long)p->y.c * 4;
    a.w_st = 0;
    a.bias = p->bias;
    DView y = dview(p->y);
    y.p += ay * y.sh + ax * y.sw;
    y.h = p->x.h;
    y.w = p->x.w;
    y.sh *= 2;
    y.sw *= 2;
    a.dst[0] = y;
    a.ndst = 1;
    a.out_n = p->x.n;
    a.out_h = p->x.h;
    a.out_w = p->x.w;
    a.cout = p->y.c;
    int r = launch_gather(a, st);
    if (r) return r;
  }
  return 0;
}

// dx[n,i,j,c] = sum_{a,b,o} dy[n,2i+a,2j+b,o] * W[c,o,a,b]: a 2x2-tap gather with input stride 2.
int direct_convt_dgrad(const b200_convt_dgrad_params* p, cudaStream_t st) {
  if (!p->w_f32) return fail(-1, "convt_dgrad (direct): w_f32 is required");
  GatherArgs a{};
  a.nsrc = 1;
  a.src[0] = dview(p->dy);
  a.ky = a.kx = 2;
  a.pad = 0;
  a.in_stride = 2;
  a.w = p->w_f32;  // index c*(cout*4) + o*4 + ab: gather-output channel = c, reduction channel = o
  a.w_so = (long long)p->dy.c * 4;
  a.w_sc = 4;
  a.w_st = 1;
  a.dst[0] = dview(p->dx);
  a.mask[0] = (const bf16*)p->mask;
  a.ndst = 1;
  a.out_n = p->dx.n;
  a.out_h = p->dx.h;
  a.out_w = p->dx.w;
  a.cout = p->dx.c;
  return launch_gather(a, st);
}

// dW[c,o,a,b] = sum x[n,i,j,c] * dy[n,2i+a,2j+b,o]: rows = x pixels ("cout" role = cin), gathered = dy.
int direct_convt_wgrad(const b200_convt_wgrad_params* p, void* ws, size_t ws_bytes, cudaStream_t st) {
  WgradArgs a{};
  a.dz = dview(p->x);
  a.nsrc = 1;
  a.src[0] = dview(p->dy);
  a.ky = a.kx = 2;
  a.pad = 0;
  a.in_stride = 2;
  a.cin_total = p->dy.c;  // gathered channels = convT output channels
  a.cout = p->x.c;        // row channels = convT input channels
  a.o_so = (long long)p->dy.c * 4;
  a.o_sc = 4;
  int r = launch_wgrad(a, p->dw_f32, nullptr, p->x, ws, ws_bytes, st);
  if (r) return r;
  if (p->db_f32) {
    const long long outputs = (long long)p->x.c * p->dy.c * 4;
    const long long npix = view_pixels(p->x);
    return bias_grad(p->dy, p->db_f32, (char*)ws + (size_t)wgrad_splits(npix, outputs) * outputs * sizeof(float), st);
  }
  return 0;
}

}  // namespace b200
