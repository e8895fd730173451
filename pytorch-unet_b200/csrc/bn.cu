// nn.BatchNorm2d (eps, momentum, affine, running statistics) applied to the post-ReLU activation
// (reference: unet.py:95, 100 — Conv -> ReLU -> BatchNorm).  Training forward = per-channel statistics
// (deterministic two-level reduction) + one normalise pass; backward = one reduction pass + one apply pass that
// also applies the ReLU mask of the convolution in front (mask == BN input > 0).  All passes are HBM-bound.
#include "chan_reduce.cuh"

namespace b200 {

struct StatsF {  // acc[0] = sum x, acc[1] = sum x^2
  template <int VEC>
  __device__ void operator()(const float (&a)[VEC], const float (&)[VEC], float (&acc)[2][VEC], int) const {
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      acc[0][j] += a[j];
      acc[1][j] += a[j] * a[j];
    }
  }
};

struct BwdSumsF {  // a = x (BN input), b = dy: acc[0] = sum dy, acc[1] = sum dy * xhat
  const float* mean;
  const float* invstd;
  template <int VEC>
  __device__ void operator()(const float (&a)[VEC], const float (&b)[VEC], float (&acc)[2][VEC], int ch0) const {
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const float xh = (a[j] - mean[ch0 + j]) * invstd[ch0 + j];
      acc[0][j] += b[j];
      acc[1][j] += b[j] * xh;
    }
  }
};

__global__ void bn_finalize_stats_kernel(const float* __restrict__ partial, int blocks, int c, double count, float eps,
                                         float momentum, float* __restrict__ running_mean,
                                         float* __restrict__ running_var, float* __restrict__ save_mean,
                                         float* __restrict__ save_invstd) {
  const int ch = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);  // one warp per channel
  if (ch >= c) return;
  const double s = warp_partial_sum(partial, blocks, 2LL * c, ch);
  const double ss = warp_partial_sum(partial, blocks, 2LL * c, (long long)c + ch);
  if ((threadIdx.x & 31) != 0) return;
  const double mean = s / count;
  double var = ss / count - mean * mean;  // biased
  if (var < 0.0) var = 0.0;
  save_mean[ch] = (float)mean;
  save_invstd[ch] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean) running_mean[ch] = (1.f - momentum) * running_mean[ch] + momentum * (float)mean;
  if (running_var) {
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    running_var[ch] = (1.f - momentum) * running_var[ch] + momentum * (float)unbiased;
  }
}

__global__ void bn_finalize_bwd_kernel(const float* __restrict__ partial, int blocks, int c, float* __restrict__ dgamma,
                                       float* __restrict__ dbeta, float* __restrict__ sums /* [2][c] */) {
  const int ch = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);  // one warp per channel
  if (ch >= c) return;
  const double s = warp_partial_sum(partial, blocks, 2LL * c, ch);
  const double ss = warp_partial_sum(partial, blocks, 2LL * c, (long long)c + ch);
  if ((threadIdx.x & 31) != 0) return;
  dbeta[ch] = (float)s;
  dgamma[ch] = (float)ss;
  sums[ch] = (float)s;
  sums[c + ch] = (float)ss;
}

// y = x * scale[c] + shift[c]   (scale = gamma * invstd, shift = beta - mean * scale)
template <int VEC>
__global__ void __launch_bounds__(256)
bn_apply_kernel(DView x, DView y, const float* __restrict__ gamma, const float* __restrict__ beta,
                const float* __restrict__ mean, const float* __restrict__ invstd_or_var, float eps, int var_mode) {
  const int lanes = x.c / VEC;
  const long long total = (long long)x.n * x.h * x.w * lanes;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int l = (int)(e % lanes);
    long long p = e / lanes;
    const int iw = (int)(p % x.w);
    p /= x.w;
    const int ih = (int)(p % x.h);
    const int n = (int)(p / x.h);
    float v[VEC];
    if (VEC == 8) {
      float t[8];
      load8s(x.p, x.lo, x.off(n, ih, iw) + l * 8, t);
#pragma unroll
      for (int j = 0; j < VEC; ++j) v[j] = t[j];
    } else {
      v[0] = bf2f(x.p[x.off(n, ih, iw) + l]);
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const int ch = l * VEC + j;
      const float is = var_mode ? rsqrtf(invstd_or_var[ch] + eps) : invstd_or_var[ch];
      const float sc = gamma[ch] * is;
      v[j] = (v[j] - mean[ch]) * sc + beta[ch];
    }
    if (VEC == 8) {
      float t[8];
#pragma unroll
      for (int j = 0; j < VEC; ++j) t[j] = v[j];
      store8s(y.p, y.lo, y.off(n, ih, iw) + l * 8, t);
    } else {
      y.p[y.off(n, ih, iw) + l] = f2bf(v[0]);
    }
  }
}

// dx = gamma*invstd*(dy - sum_dy/M - xhat*sum_dy_xhat/M) [* (x > 0)]
template <int VEC>
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(DView x, DView dy, DView dx, const float* __restrict__ gamma, const float* __restrict__ mean,
                    const float* __restrict__ invstd, const float* __restrict__ sums, float inv_count, int relu_mask) {
  const int lanes = x.c / VEC;
  const int c = x.c;
  const long long total = (long long)x.n * x.h * x.w * lanes;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int l = (int)(e % lanes);
    long long p = e / lanes;
    const int iw = (int)(p % x.w);
    p /= x.w;
    const int ih = (int)(p % x.h);
    const int n = (int)(p / x.h);
    float xv[VEC], gv[VEC];
    if (VEC == 8) {
      float t[8];
      unpack8(*reinterpret_cast<const bf16x8*>(x.p + x.off(n, ih, iw) + l * 8), t);
#pragma unroll
      for (int j = 0; j < VEC; ++j) xv[j] = t[j];
      unpack8(*reinterpret_cast<const bf16x8*>(dy.p + dy.off(n, ih, iw) + l * 8), t);
#pragma unroll
      for (int j = 0; j < VEC; ++j) gv[j] = t[j];
    } else {
      xv[0] = bf2f(x.p[x.off(n, ih, iw) + l]);
      gv[0] = bf2f(dy.p[dy.off(n, ih, iw) + l]);
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const int ch = l * VEC + j;
      const float is = invstd[ch];
      const float xh = (xv[j] - mean[ch]) * is;
      float r = gamma[ch] * is * (gv[j] - sums[ch] * inv_count - xh * sums[c + ch] * inv_count);
      if (relu_mask && !(xv[j] > 0.f)) r = 0.f;
      gv[j] = r;
    }
    if (VEC == 8) {
      float t[8];
#pragma unroll
      for (int j = 0; j < VEC; ++j) t[j] = gv[j];
      *reinterpret_cast<bf16x8*>(dx.p + dx.off(n, ih, iw) + l * 8) = pack8(t);
    } else {
      dx.p[dx.off(n, ih, iw) + l] = f2bf(gv[0]);
    }
  }
}


// ------------------------------------------------------------------ dense fast paths of the two apply passes
// The generic kernels above decode (n, h, w, lane) with 64-bit divisions for every 16-byte vector and keep one load in
// flight per thread: 2.3-3.1 TB/s.  When the tensors are pixel-contiguous and the channel count is 8 * 2^k, the flat
// vector index is the address, the grid stride is a multiple of the vectors per pixel (so a thread owns the SAME eight
// channels for the whole kernel and its per-channel constants live in registers), and U vectors are in flight per thread.
constexpr int kBnU = 4;

template <bool SPLIT>
__global__ void __launch_bounds__(256)
bn_apply_dense_kernel(const bf16* __restrict__ x, const bf16* __restrict__ xlo, bf16* __restrict__ y, bf16* __restrict__ ylo,
                      long long nvec, int lanes, const float* __restrict__ gamma, const float* __restrict__ beta,
                      const float* __restrict__ mean, const float* __restrict__ invstd_or_var, float eps, int var_mode) {
  const int l = threadIdx.x % lanes;   // 256 % lanes == 0: fixed for every vector this thread touches
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int ch = l * 8 + j;
    const float is = var_mode ? rsqrtf(invstd_or_var[ch] + eps) : invstd_or_var[ch];
    sc[j] = gamma[ch] * is;
    sh[j] = beta[ch] - mean[ch] * sc[j];
  }
  const long long step = (long long)gridDim.x * (256 * kBnU);
  for (long long e0 = (long long)blockIdx.x * (256 * kBnU) + threadIdx.x; e0 < nvec; e0 += step) {
    bf16x8 v[kBnU], vl[SPLIT ? kBnU : 1];
#pragma unroll
    for (int u = 0; u < kBnU; ++u) {
      const long long e = e0 + u * 256;
      v[u] = make_uint4(0, 0, 0, 0);
      if (SPLIT) vl[u] = make_uint4(0, 0, 0, 0);
      if (e < nvec) {
        v[u] = *reinterpret_cast<const bf16x8*>(x + e * 8);
        if (SPLIT) vl[u] = *reinterpret_cast<const bf16x8*>(xlo + e * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < kBnU; ++u) {
      const long long e = e0 + u * 256;
      if (e >= nvec) break;
      float f[8];
      unpack8(v[u], f);
      if (SPLIT) {
        float t[8];
        unpack8(vl[u], t);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] += t[j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = f[j] * sc[j] + sh[j];
      store8s(y, SPLIT ? ylo : nullptr, e * 8, f);
    }
  }
}

// dx = A*dy + B*x + C per channel [* (x > 0)], with A = gamma*invstd, B = -A*invstd*sum_dy_xhat/M,
// C = -A*sum_dy/M - B*mean   (the training-mode formula expanded; M -> infinity for frozen statistics)
__global__ void __launch_bounds__(256)
bn_bwd_apply_dense_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy, bf16* __restrict__ dx, long long nvec,
                          int lanes, int c, const float* __restrict__ gamma, const float* __restrict__ mean,
                          const float* __restrict__ invstd, const float* __restrict__ sums, float inv_count, int relu_mask) {
  const int l = threadIdx.x % lanes;
  float A[8], B[8], C[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int ch = l * 8 + j;
    const float is = invstd[ch];
    A[j] = gamma[ch] * is;
    B[j] = -A[j] * is * sums[c + ch] * inv_count;
    C[j] = -A[j] * sums[ch] * inv_count - B[j] * mean[ch];
  }
  const long long step = (long long)gridDim.x * (256 * kBnU);
  for (long long e0 = (long long)blockIdx.x * (256 * kBnU) + threadIdx.x; e0 < nvec; e0 += step) {
    bf16x8 xv[kBnU], gv[kBnU];
#pragma unroll
    for (int u = 0; u < kBnU; ++u) {
      const long long e = e0 + u * 256;
      xv[u] = make_uint4(0, 0, 0, 0);
      gv[u] = make_uint4(0, 0, 0, 0);
      if (e < nvec) {
        xv[u] = *reinterpret_cast<const bf16x8*>(x + e * 8);
        gv[u] = *reinterpret_cast<const bf16x8*>(dy + e * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < kBnU; ++u) {
      const long long e = e0 + u * 256;
      if (e >= nvec) break;
      float xf[8], gf[8];
      unpack8(xv[u], xf);
      unpack8(gv[u], gf);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float r = A[j] * gf[j] + B[j] * xf[j] + C[j];
        if (relu_mask && !(xf[j] > 0.f)) r = 0.f;
        gf[j] = r;
      }
      *reinterpret_cast<bf16x8*>(dx + e * 8) = pack8(gf);
    }
  }
}

static bool bn_dense_ok(const b200_view& v) {
  const int lanes = v.c / 8;
  return v.c % 8 == 0 && lanes >= 1 && lanes <= 256 && 256 % lanes == 0 && vec8_ok(v) && v.stride_w == v.c &&
         v.stride_h == (int64_t)v.w * v.stride_w && (v.n == 1 || v.stride_n == (int64_t)v.h * v.stride_h);
}
static int bn_dense_grid(long long nvec) {
  long long g = (nvec + 256 * kBnU - 1) / (256 * kBnU);
  const long long cap = (long long)kNumSMsB200 * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace b200

using namespace b200;

extern "C" {

size_t b200unet_bn_workspace_bytes(int c) {
  return reduce_workspace_bytes(c, 2) + (size_t)2 * c * sizeof(float);
}

int b200unet_bn_fwd_train(const b200_view* x, const b200_view* y, const float* gamma, const float* beta,
                          float* running_mean, float* running_var, float momentum, float eps, float* save_mean,
                          float* save_invstd, void* workspace, size_t workspace_bytes, void* stream) {
  B200_REQUIRE(view_ok(x) && view_ok(y) && same_extent(*x, *y) && gamma && beta && save_mean && save_invstd && workspace,
               "bn_fwd_train: bad arguments");
  B200_REQUIRE(workspace_bytes >= reduce_workspace_bytes(x->c, 2), "bn_fwd_train: workspace too small");
  B200_REQUIRE((x->lo == nullptr) == (y->lo == nullptr), "bn_fwd_train: x and y must be of the same precision tier");
  B200_REQUIRE(!x->lo || (vec8_ok(*x) && vec8_ok(*y)), "bn_fwd_train: the split tier needs channels / strides % 8 == 0");
  cudaStream_t st = as_stream(stream);
  ReducePlan pl;
  int r = launch_chan_reduce<2, false>(StatsF(), *x, nullptr, (float*)workspace, &pl, st);
  if (r) return r;
  bn_finalize_stats_kernel<<<finalize_grid(x->c), kFinalizeThreads, 0, st>>>((const float*)workspace, pl.blocks, x->c,
                                                              (double)view_pixels(*x), eps, momentum, running_mean,
                                                              running_var, save_mean, save_invstd);
  r = check_launch("bn finalize");
  if (r) return r;
  if (bn_dense_ok(*x) && bn_dense_ok(*y)) {
    const long long nvec = view_pixels(*x) * (x->c / 8);
    if (x->lo)
      bn_apply_dense_kernel<true><<<bn_dense_grid(nvec), 256, 0, st>>>((const bf16*)x->ptr, (const bf16*)x->lo, (bf16*)y->ptr,
                                                                       (bf16*)y->lo, nvec, x->c / 8, gamma, beta, save_mean,
                                                                       save_invstd, eps, 0);
    else
      bn_apply_dense_kernel<false><<<bn_dense_grid(nvec), 256, 0, st>>>((const bf16*)x->ptr, nullptr, (bf16*)y->ptr, nullptr,
                                                                        nvec, x->c / 8, gamma, beta, save_mean, save_invstd,
                                                                        eps, 0);
    return check_launch("bn apply (dense)");
  }
  const bool v8 = vec8_ok(*x) && vec8_ok(*y);
  const long long total = view_pixels(*x) * (v8 ? x->c / 8 : x->c);
  if (v8)
    bn_apply_kernel<8><<<stream_grid(total), 256, 0, st>>>(dview(*x), dview(*y), gamma, beta, save_mean, save_invstd, eps, 0);
  else
    bn_apply_kernel<1><<<stream_grid(total), 256, 0, st>>>(dview(*x), dview(*y), gamma, beta, save_mean, save_invstd, eps, 0);
  return check_launch("bn apply");
}

int b200unet_bn_fwd_eval(const b200_view* x, const b200_view* y, const float* gamma, const float* beta,
                         const float* running_mean, const float* running_var, float eps, void* stream) {
  B200_REQUIRE(view_ok(x) && view_ok(y) && same_extent(*x, *y) && gamma && beta && running_mean && running_var,
               "bn_fwd_eval: bad arguments");
  B200_REQUIRE((x->lo == nullptr) == (y->lo == nullptr), "bn_fwd_eval: x and y must be of the same precision tier");
  B200_REQUIRE(!x->lo || (vec8_ok(*x) && vec8_ok(*y)), "bn_fwd_eval: the split tier needs channels / strides % 8 == 0");
  cudaStream_t st = as_stream(stream);
  if (bn_dense_ok(*x) && bn_dense_ok(*y)) {
    const long long nvec = view_pixels(*x) * (x->c / 8);
    if (x->lo)
      bn_apply_dense_kernel<true><<<bn_dense_grid(nvec), 256, 0, st>>>((const bf16*)x->ptr, (const bf16*)x->lo, (bf16*)y->ptr,
                                                                       (bf16*)y->lo, nvec, x->c / 8, gamma, beta, running_mean,
                                                                       running_var, eps, 1);
    else
      bn_apply_dense_kernel<false><<<bn_dense_grid(nvec), 256, 0, st>>>((const bf16*)x->ptr, nullptr, (bf16*)y->ptr, nullptr,
                                                                        nvec, x->c / 8, gamma, beta, running_mean, running_var,
                                                                        eps, 1);
    return check_launch("bn apply (eval, dense)");
  }
  const bool v8 = vec8_ok(*x) && vec8_ok(*y);
  const long long total = view_pixels(*x) * (v8 ? x->c / 8 : x->c);
  if (v8)
    bn_apply_kernel<8><<<stream_grid(total), 256, 0, st>>>(dview(*x), dview(*y), gamma, beta, running_mean, running_var, eps, 1);
  else
    bn_apply_kernel<1><<<stream_grid(total), 256, 0, st>>>(dview(*x), dview(*y), gamma, beta, running_mean, running_var, eps, 1);
  return check_launch("bn apply (eval)");
}

int b200unet_bn_bwd(const b200_view* x, const b200_view* dy, const b200_view* dx, const float* gamma,
                    const float* save_mean, const float* save_invstd, float* dgamma, float* dbeta, int flags,
                    void* workspace, size_t workspace_bytes, void* stream) {
  B200_REQUIRE(view_ok(x) && view_ok(dy) && view_ok(dx) && same_extent(*x, *dy) && same_extent(*x, *dx) && gamma &&
                   save_mean && save_invstd && dgamma && dbeta && workspace,
               "bn_bwd: bad arguments");
  // workspace: partials [blocks][2][c] followed by the reduced sums [2][c]
  const size_t need = reduce_workspace_bytes(x->c, 2) + (size_t)2 * x->c * sizeof(float);
  B200_REQUIRE(workspace_bytes >= need, "bn_bwd: workspace too small (need %zu)", need);
  cudaStream_t st = as_stream(stream);
  float* partial = (float*)workspace;
  float* sums = partial + (size_t)kReduceMaxBlocks * 2 * x->c;
  ReducePlan pl;
  BwdSumsF f{save_mean, save_invstd};
  int r = launch_chan_reduce<2, true>(f, *x, dy, partial, &pl, st);
  if (r) return r;
  bn_finalize_bwd_kernel<<<finalize_grid(x->c), kFinalizeThreads, 0, st>>>(partial, pl.blocks, x->c, dgamma, dbeta, sums);
  r = check_launch("bn bwd finalize");
  if (r) return r;
  const bool v8 = vec8_ok(*x) && vec8_ok(*dy) && vec8_ok(*dx);
  const long long total = view_pixels(*x) * (v8 ? x->c / 8 : x->c);
  // flags bit 1: mean / invstd are constants (eval mode: running statistics), so the batch-statistic correction terms
  // of the training-mode formula vanish: dx = gamma * invstd * dy
  const float inv_count = (flags & 2) ? 0.f : 1.f / (float)view_pixels(*x);
  const int relu_mask = flags & 1;
  if (bn_dense_ok(*x) && bn_dense_ok(*dy) && bn_dense_ok(*dx)) {
    const long long nvec = view_pixels(*x) * (x->c / 8);
    bn_bwd_apply_dense_kernel<<<bn_dense_grid(nvec), 256, 0, st>>>((const bf16*)x->ptr, (const bf16*)dy->ptr, (bf16*)dx->ptr, nvec,
                                                                   x->c / 8, x->c, gamma, save_mean, save_invstd, sums, inv_count,
                                                                   relu_mask);
    return check_launch("bn bwd apply (dense)");
  }
  if (v8)
    bn_bwd_apply_kernel<8><<<stream_grid(total), 256, 0, st>>>(dview(*x), dview(*dy), dview(*dx), gamma, save_mean,
                                                              save_invstd, sums, inv_count, relu_mask);
  else
    bn_bwd_apply_kernel<1><<<stream_grid(total), 256, 0, st>>>(dview(*x), dview(*dy), dview(*dx), gamma, save_mean,
                                                              save_invstd, sums, inv_count, relu_mask);
  return check_launch("bn bwd apply");
}
}
