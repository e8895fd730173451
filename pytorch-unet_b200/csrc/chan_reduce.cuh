// Deterministic per-channel reductions over the pixels of NHWC bf16 views (BatchNorm statistics, BatchNorm
// backward sums, bias gradients).  Pass 1 writes one fp32 partial per block, pass 2 (the caller's finalize
// kernel) sums the partials in double in a fixed order, so results are run-to-run reproducible.
#pragma once
#include "common.cuh"

namespace b200 {

constexpr int kReduceMaxBlocks = 4 * kNumSMsB200;
constexpr int kReduceThreads = 256;

// Functor contract:  template<int VEC> __device__ void operator()(const float (&a)[VEC], const float (&b)[VEC],
//                                                                  float (&acc)[NV][VEC]) const;
// `b` is only loaded when HAS_B.
template <int VEC, bool HAS_B>
__device__ __forceinline__ void reduce_load(const DView& a, const DView& b, long long oa, long long ob, float (&fa)[VEC],
                                            float (&fb)[VEC]) {
  if (VEC == 8) {
    float t8[8];
    load8s(a.p, a.lo, oa, t8);  // `a` may be a split-tier tensor (forward statistics)
#pragma unroll
    for (int j = 0; j < VEC; ++j) fa[j] = t8[j];
    if (HAS_B) {
      unpack8(*reinterpret_cast<const bf16x8*>(b.p + ob), t8);
#pragma unroll
      for (int j = 0; j < VEC; ++j) fb[j] = t8[j];
    }
  } else {
    fa[0] = bf2f(a.p[oa]);
    if (HAS_B) fb[0] = bf2f(b.p[ob]);
  }
  if (!HAS_B) {
#pragma unroll
    for (int j = 0; j < VEC; ++j) fb[j] = 0.f;
  }
}

// `dense` (both views contiguous over their pixels: offset = pixel * stride_w): no index arithmetic beyond one multiply,
// and four pixels in flight per thread — one 16-byte load per dependent loop iteration left this kernel at ~2.8 TB/s.
template <int VEC, int NV, bool HAS_B, class F>
__global__ void __launch_bounds__(kReduceThreads)
chan_reduce_kernel(F f, DView a, DView b, int lanes, int tb, long long npix, int dense, float* __restrict__ partial) {
  extern __shared__ float red[];  // [NV*VEC][tb]
  const int t = threadIdx.x;
  float acc[NV][VEC];
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int j = 0; j < VEC; ++j) acc[v][j] = 0.f;

  if (t < tb) {
    const int ps = t / lanes, l = t - ps * lanes;
    const int pb = tb / lanes;
    const long long step = (long long)gridDim.x * pb;
    long long p = (long long)blockIdx.x * pb + ps;
    if (dense) {
      const long long lo = (long long)l * VEC;
      for (; p + 3 * step < npix; p += 4 * step) {
        float fa[4][VEC], fb[4][VEC];
#pragma unroll
        for (int u = 0; u < 4; ++u) reduce_load<VEC, HAS_B>(a, b, (p + u * step) * a.sw + lo, (p + u * step) * b.sw + lo, fa[u], fb[u]);
#pragma unroll
        for (int u = 0; u < 4; ++u) f(fa[u], fb[u], acc, l * VEC);
      }
      for (; p < npix; p += step) {
        float fa[VEC], fb[VEC];
        reduce_load<VEC, HAS_B>(a, b, p * a.sw + lo, p * b.sw + lo, fa, fb);
        f(fa, fb, acc, l * VEC);
      }
    } else {
      const long long hw = (long long)a.h * a.w;
      for (; p < npix; p += step) {
        const int n = (int)(p / hw);
        const int r = (int)(p - n * hw);
        const int ih = r / a.w, iw = r - ih * a.w;
        float fa[VEC], fb[VEC];
        reduce_load<VEC, HAS_B>(a, b, a.off(n, ih, iw) + l * VEC, b.off(n, ih, iw) + l * VEC, fa, fb);
        f(fa, fb, acc, l * VEC);
      }
    }
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int j = 0; j < VEC; ++j) red[(v * VEC + j) * tb + t] = acc[v][j];
  }
  __syncthreads();
  // thread q sums column (v, channel) over the pixel slots in fixed order
  const int c = lanes * VEC;
  for (int q = t; q < NV * c; q += blockDim.x) {
    const int v = q / c, ch = q - v * c;
    const int l = ch / VEC, j = ch - l * VEC;
    float s = 0.f;
    for (int ps = 0; ps < tb / lanes; ++ps) s += red[(v * VEC + j) * tb + ps * lanes + l];
    partial[((long long)blockIdx.x * NV + v) * c + ch] = s;
  }
}

// Second-level reduction helper: sum of partial[b * stride + idx] over b = 0..blocks-1 done by ONE WARP (lanes stride over
// the blocks, then a fixed butterfly), in double.  Deterministic, and ~20x faster than one thread walking all partials.
__device__ __forceinline__ double warp_partial_sum(const float* __restrict__ partial, int blocks, long long stride,
                                                    long long idx) {
  const int lane = threadIdx.x & 31;
  double s = 0.0;
  int b = lane;
  // four independent (strided, so uncoalesced) loads in flight per lane: one dependent load per iteration made the
  // finalize kernels 13-17 us each, 0.5 ms per training step of the BatchNorm configurations
  for (; b + 96 < blocks; b += 128) {
    const float v0 = partial[(long long)b * stride + idx], v1 = partial[(long long)(b + 32) * stride + idx];
    const float v2 = partial[(long long)(b + 64) * stride + idx], v3 = partial[(long long)(b + 96) * stride + idx];
    s += (double)v0;
    s += (double)v1;
    s += (double)v2;
    s += (double)v3;
  }
  for (; b < blocks; b += 32) s += (double)partial[(long long)b * stride + idx];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  return s;
}
constexpr int kFinalizeThreads = 256;  // 8 warps = 8 outputs per block
inline int finalize_grid(long long outputs) { return (int)((outputs + kFinalizeThreads / 32 - 1) / (kFinalizeThreads / 32)); }

struct ReducePlan {
  int vec, lanes, tb, blocks;
  size_t smem;
};

// Returns false if the channel count is not supported (c % 8 != 0 and c > 256, or c > 2048).
inline bool plan_reduce(const b200_view& a, const b200_view* b, int nv, ReducePlan* pl) {
  const bool al = (reinterpret_cast<uintptr_t>(a.ptr) % 16 == 0) && a.stride_w % 8 == 0 && a.stride_h % 8 == 0 &&
                  a.stride_n % 8 == 0 &&
                  (!b || ((reinterpret_cast<uintptr_t>(b->ptr) % 16 == 0) && b->stride_w % 8 == 0 &&
                          b->stride_h % 8 == 0 && b->stride_n % 8 == 0));
  pl->vec = (a.c % 8 == 0 && al) ? 8 : 1;
  pl->lanes = a.c / pl->vec;
  if (pl->lanes > kReduceThreads) return false;
  pl->tb = (kReduceThreads / pl->lanes) * pl->lanes;
  const long long npix = view_pixels(a);
  const long long pb = pl->tb / pl->lanes;
  long long blocks = (npix + pb * 8 - 1) / (pb * 8);  // >= 8 pixels per thread slot
  if (blocks < 1) blocks = 1;
  if (blocks > kReduceMaxBlocks) blocks = kReduceMaxBlocks;
  pl->blocks = (int)blocks;
  pl->smem = (size_t)nv * pl->vec * pl->tb * sizeof(float);
  return true;
}

inline size_t reduce_workspace_bytes(int c, int nv) { return (size_t)kReduceMaxBlocks * nv * c * sizeof(float); }

template <int NV, bool HAS_B, class F>
inline int launch_chan_reduce(F f, const b200_view& a, const b200_view* b, float* partial, ReducePlan* pl,
                              cudaStream_t st) {
  if (!plan_reduce(a, b, NV, pl)) return fail(-1, "channel reduction: unsupported channel count %d", a.c);
  DView da = dview(a), db = b ? dview(*b) : da;
  const long long npix = view_pixels(a);
  auto pix_dense = [](const b200_view& v) {
    return v.stride_h == (int64_t)v.w * v.stride_w && (v.n == 1 || v.stride_n == (int64_t)v.h * v.stride_h);
  };
  const int dense = pix_dense(a) && (!b || pix_dense(*b));
  if (pl->vec == 8)
    chan_reduce_kernel<8, NV, HAS_B, F><<<pl->blocks, kReduceThreads, pl->smem, st>>>(f, da, db, pl->lanes, pl->tb,
                                                                                     npix, dense, partial);
  else
    chan_reduce_kernel<1, NV, HAS_B, F><<<pl->blocks, kReduceThreads, pl->smem, st>>>(f, da, db, pl->lanes, pl->tb,
                                                                                     npix, dense, partial);
  return check_launch("chan_reduce");
}

}  // namespace b200
