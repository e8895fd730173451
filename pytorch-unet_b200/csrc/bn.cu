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
  if (v8)
    bn_bwd_apply_kernel<8><<<stream_grid(total), 256, 0, st>>>(dview(*x), dview(*dy), dview(*dx), gamma, save_mean,
                                                              save_invstd, sums, inv_count, relu_mask);
  else
    bn_bwd_apply_kernel<1><<<stream_grid(total), 256, 0, st>>>(dview(*x), dview(*dy), dview(*dx), gamma, save_mean,
                                                              save_invstd, sums, inv_count, relu_mask);
  return check_launch("bn bwd apply");
}
}
