// CVO kernel-Gramian loss: what the reference computes with three hand-written CUDA extensions whose
// source is absent (geometry.py:4 `cross_prod_cuda, cross_subtract_cuda, sub_norm_cuda_half_paral`) plus PyTorch glue:
//
//   kern_mat        geometry.py:47-136      K[b,i,j] = exp(-|x1_i - x2_j|^2 / (2 s^2)), zero below 8.315e-3
//   calc_gramian    network_modules.py:995-1015   one such N1 x N2 matrix per domain (xyz, colour, feature)
//   calc_inner_prod network_modules.py:1096-1149  sum_ij  prod_domains K_d[i,j]  (* w_i w_j in weight_map_mode)
//   calc_w_v        network_modules.py:1052-1094  sum_ij P[i,j] (x1_i x x2_j),  sum_ij P[i,j] (x1_i - x2_j)
//
// The reference materialises every N1 x N2 matrix (12 288^2 fp32 = 604 MB each, three domains x three frame pairs, kept
// for autograd).  Here ONE kernel evaluates the distance, the cut-off, the exponentials and the product for a pair and
// keeps only per-point sums, so nothing of size N1 x N2 ever reaches HBM; the backward pass recomputes the pair weights
// and reduces them straight into the point gradients.  The work is exp/FMA on the CUDA cores (2-16 channels per point:
// a 12 288 x 12 288 x 14 "GEMM" is 4 GFLOP, far too thin for tcgen05 and dominated by the exponential and the cut-off).
//
// Layout: point sets are fp32, channel-planar x[b][c][n] exactly as the reference passes them (B*C*N).
// A thread owns ONE point of set 1 (its channels live in registers); the CTA walks set 2 in tiles of 128 points staged
// in shared memory ([channel][point], read as broadcast float4 = four pairs per step).  Set 2 is split over
// gridDim.y so that the grid covers the 148 SMs a few times; the per-split sums go to a workspace and a finalize kernel
// folds them (fixed order: results are deterministic).  Pairs beyond the cut-off of the first domain are dropped before
// any exponential is evaluated — on real frames that is > 90 % of all pairs.
#include "common.cuh"

namespace b200 {

constexpr int kCvoThreads = 128;  // points of set 1 per CTA
constexpr int kCvoTile = 128;     // points of set 2 per shared-memory tile
constexpr int kCvoMaxC = 16;      // channels over all domains of one call
constexpr float kCvoThreT = 8.315e-3f;  // geometry.py:108

enum { CVO_FWD = 0, CVO_FWD_WV = 1, CVO_BWD = 2, CVO_MAT_BWD = 3 };

struct CvoArgs {
  const float* x1[kCvoMaxC];  // per channel: plane of batch 0 (set 1: owned by threads)
  const float* x2[kCvoMaxC];  // set 2: staged in shared memory
  long long bs1[kCvoMaxC], bs2[kCvoMaxC];  // batch strides of each plane
  float coef[kCvoMaxC];       // 1 / (2 s^2) of the channel's domain
  unsigned last_mask;         // bit c: channel c closes its domain
  unsigned dot_mask;          // bit c: channel c belongs to the plain inner-product domain (`not kernalize`)
  const float* w1;            // [B][N1] point weights or null
  const float* w2;            // [B][N2]
  const float* dy;            // CVO_MAT_BWD: upstream gradient of the materialised matrix, element (i of set 1, j of set 2) at
  long long dy_s1, dy_s2, dy_sb;  //   dy[b * dy_sb + i * dy_s1 + j * dy_s2]
  int plain_distance;         // CVO_MAT_BWD: 1 = sub_norm (G = dy), 0 = kern_mat (G = -coef * K * dy)
  int n1, n2, nsplit, geo0;   // geo0: first channel of the 3-channel geometry domain (CVO_FWD_WV), else -1
  float cut_s;                // pairs with scaled squared distance above this are zero for certain (-ln 8.315e-3, plus a margin)
  float* ws;                  // [nsplit][B][rows][N1]
};

template <int MODE, int CT>
struct CvoRows {
  static constexpr int value = MODE == CVO_FWD ? 1 : (MODE == CVO_FWD_WV ? 4 : 1 + CT);
};

template <int CT, int MODE>
__global__ void __launch_bounds__(kCvoThreads) cvo_pair_kernel(CvoArgs a) {
  constexpr int ROWS = CvoRows<MODE, CT>::value;
  constexpr int NT = MODE == CVO_FWD ? 1 : (MODE == CVO_FWD_WV ? 3 : CT);
  __shared__ __align__(16) float s2[CT + 1][kCvoTile];  // row CT: weight of the point (0 = padding)
  const int b = blockIdx.z;
  const int i = blockIdx.x * kCvoThreads + threadIdx.x;
  const bool live = i < a.n1;
  float xa[CT];
#pragma unroll
  for (int c = 0; c < CT; ++c) xa[c] = live ? a.x1[c][b * a.bs1[c] + i] : 0.f;
  float S = 0.f, T[NT];
#pragma unroll
  for (int c = 0; c < NT; ++c) T[c] = 0.f;
  const int per = ((a.n2 + a.nsplit - 1) / a.nsplit + kCvoTile - 1) / kCvoTile * kCvoTile;
  const int j_begin = blockIdx.y * per;
  const int j_end = min(a.n2, j_begin + per);
  const float* dyrow = MODE == CVO_MAT_BWD ? a.dy + b * a.dy_sb + (long long)i * a.dy_s1 : nullptr;
  for (int j0 = j_begin; j0 < j_end; j0 += kCvoTile) {
    __syncthreads();
    for (int e = threadIdx.x; e < (CT + 1) * kCvoTile; e += kCvoThreads) {
      const int c = e / kCvoTile, jj = e - c * kCvoTile;
      const int j = j0 + jj;
      float v = 0.f;
      if (j < j_end) v = c < CT ? a.x2[c][b * a.bs2[c] + j] : (a.w2 ? a.w2[(long long)b * a.n2 + j] : 1.f);
      s2[c][jj] = v;
    }
    __syncthreads();
#pragma unroll 1
    for (int q = 0; q < kCvoTile / 4; ++q) {
      float d[4] = {0.f, 0.f, 0.f, 0.f}, dot[4] = {0.f, 0.f, 0.f, 0.f};
      float E[4] = {1.f, 1.f, 1.f, 1.f};  // product of the RBF factors
      bool dead = false;
#pragma unroll
      for (int c = 0; c < CT; ++c) {
        const float4 xb = *reinterpret_cast<const float4*>(&s2[c][q * 4]);
        const float xv[4] = {xb.x, xb.y, xb.z, xb.w};
        if ((a.dot_mask >> c) & 1u) {
#pragma unroll
          for (int u = 0; u < 4; ++u) dot[u] = fmaf(xa[c], xv[u], dot[u]);
        } else {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float t = xa[c] - xv[u];
            d[u] = fmaf(t, t, d[u]);
          }
          if ((a.last_mask >> c) & 1u) {
            if (MODE == CVO_MAT_BWD && a.plain_distance) {
              // sub_norm backward: the pair weight is dy itself
            } else {
              bool any = false;
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const float s = d[u] * a.coef[c];
                float ev = 0.f;
                if (s <= a.cut_s) {              // the reference's test is on the VALUE (geometry.py:120): evaluate it
                  ev = expf(-s);
                  ev = ev >= kCvoThreT ? ev : 0.f;
                }
                E[u] *= ev;
                any |= E[u] != 0.f;
                d[u] = 0.f;
              }
              if (!any) {
                dead = true;
                break;  // every pair of this group is cut off in this domain: the product is zero whatever follows
              }
            }
          }
        }
      }
      if (dead) continue;
      const float4 wb = *reinterpret_cast<const float4*>(&s2[CT][q * 4]);
      const float wv[4] = {wb.x, wb.y, wb.z, wb.w};
      float G[4] = {0.f, 0.f, 0.f, 0.f};
      if (MODE == CVO_MAT_BWD) {
        if (live) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int j = j0 + q * 4 + u;
            if (j < j_end) G[u] = dyrow[(long long)j * a.dy_s2];
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float Pw, Ew;  // pair weight for the RBF channels / for the inner-product channels
        if (MODE == CVO_MAT_BWD) {
          Pw = a.plain_distance ? G[u] : -a.coef[CT - 1] * E[u] * G[u];
          Ew = Pw;
        } else {
          Ew = E[u] * wv[u];
          Pw = a.dot_mask ? Ew * dot[u] : Ew;
        }
        S += Pw;
        if (MODE == CVO_FWD_WV) {
#pragma unroll
          for (int c = 0; c < 3; ++c) T[c] = fmaf(Pw, s2[a.geo0 + c][q * 4 + u], T[c]);
        } else if (MODE == CVO_BWD || MODE == CVO_MAT_BWD) {
#pragma unroll
          for (int c = 0; c < CT; ++c) T[c] = fmaf(((a.dot_mask >> c) & 1u) ? Ew : Pw, s2[c][q * 4 + u], T[c]);
        }
      }
    }
  }
  if (live) {
    float* out = a.ws + ((long long)(blockIdx.y * gridDim.z + b) * ROWS) * a.n1 + i;
    out[0] = S;
    if (MODE != CVO_FWD) {
#pragma unroll
      for (int c = 0; c < NT; ++c) out[(long long)(1 + c) * a.n1] = T[c];
    }
  }
}

// out[b] = sum_i w1_i S_i ; wv[b][0..2] = sum_i w1_i (x1_i x T_i), wv[b][3..5] = sum_i w1_i (x1_i S_i - T_i)
__global__ void __launch_bounds__(1024) cvo_fwd_finalize_kernel(const float* __restrict__ ws, int nsplit, int rows, int n1,
                                                               const float* __restrict__ w1, const float* gx, const float* gy,
                                                               const float* gz, long long gbs, float* __restrict__ out,
                                                               float* __restrict__ wv) {
  const int b = blockIdx.x, nb = gridDim.x;
  double acc[7] = {0, 0, 0, 0, 0, 0, 0};
  for (int i = threadIdx.x; i < n1; i += blockDim.x) {
    float r[4] = {0.f, 0.f, 0.f, 0.f};
    for (int s = 0; s < nsplit; ++s)
      for (int k = 0; k < rows && k < 4; ++k) r[k] += ws[((long long)(s * nb + b) * rows + k) * n1 + i];
    const float w = w1 ? w1[(long long)b * n1 + i] : 1.f;
    acc[0] += (double)(w * r[0]);
    if (wv) {
      const float x = gx[b * gbs + i], y = gy[b * gbs + i], z = gz[b * gbs + i];
      acc[1] += (double)(w * (y * r[3] - z * r[2]));
      acc[2] += (double)(w * (z * r[1] - x * r[3]));
      acc[3] += (double)(w * (x * r[2] - y * r[1]));
      acc[4] += (double)(w * (x * r[0] - r[1]));
      acc[5] += (double)(w * (y * r[0] - r[2]));
      acc[6] += (double)(w * (z * r[0] - r[3]));
    }
  }
  __shared__ double red[7][32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    double v = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[k][wid] = v;
  }
  __syncthreads();
  if (threadIdx.x < 7) {
    double v = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += red[threadIdx.x][w];
    if (threadIdx.x == 0) out[b] = (float)v;
    else if (wv) wv[b * 6 + threadIdx.x - 1] = (float)v;
  }
}

// dx[c][i] (+)= g * w_i * 2 coef_c (T_c - x_c S)   (RBF channel)      g * w_i * T_c   (inner-product channel)
// dw[i] = g * S_i.   MAT_BWD: dx[c][i] = 2 (x_c S - T_c) with S, T built from G (the sign and coef are inside G).
struct CvoBwdOut {
  float* dx[kCvoMaxC];  // per channel plane of batch 0, or null
  long long bs[kCvoMaxC];
  float* dw;            // [B][N1] or null
  const float* grad;    // [B] upstream gradient of out[b] (device), null = 1
  int mat;              // CVO_MAT_BWD semantics
};
__global__ void __launch_bounds__(256) cvo_bwd_finalize_kernel(CvoArgs a, CvoBwdOut o, int ct, int nb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (i >= a.n1) return;
  const int rows = 1 + ct;
  float S = 0.f;
  for (int s = 0; s < a.nsplit; ++s) S += a.ws[((long long)(s * nb + b) * rows) * a.n1 + i];
  const float g = o.grad ? o.grad[b] : 1.f;
  const float w = a.w1 ? a.w1[(long long)b * a.n1 + i] : 1.f;
  if (o.dw) o.dw[(long long)b * a.n1 + i] = g * S;
  for (int c = 0; c < ct; ++c) {
    if (!o.dx[c]) continue;
    float T = 0.f;
    for (int s = 0; s < a.nsplit; ++s) T += a.ws[((long long)(s * nb + b) * rows + 1 + c) * a.n1 + i];
    const float x = a.x1[c][b * a.bs1[c] + i];
    float v;
    if (o.mat) v = 2.f * (x * S - T);
    else if ((a.dot_mask >> c) & 1u) v = g * w * T;
    else v = g * w * 2.f * a.coef[c] * (T - x * S);
    o.dx[c][b * o.bs[c] + i] = v;
  }
}

// ------------------------------------------------------------------ materialised matrices (drop-in for the extensions)
// out[b][i][j] = |x1_i - x2_j|^2 (sub_norm) or its thresholded exponential (kern_mat).  16 rows x 4 columns per thread,
// float4 stores along j; HBM-write-bound (4 bytes per pair).
constexpr int kMatRows = 16;
template <bool KERN>
__global__ void __launch_bounds__(128) cvo_mat_fwd_kernel(const float* __restrict__ x1, const float* __restrict__ x2, int C,
                                                         int n1, int n2, float coef, float* __restrict__ out) {
  extern __shared__ float s1[];  // [C][16]
  const int b = blockIdx.z;
  const int i0 = blockIdx.y * kMatRows;
  const int j = (blockIdx.x * 128 + threadIdx.x) * 4;
  for (int e = threadIdx.x; e < C * kMatRows; e += 128) {
    const int c = e / kMatRows, r = e - c * kMatRows;
    s1[e] = i0 + r < n1 ? x1[((long long)b * C + c) * n1 + i0 + r] : 0.f;
  }
  __syncthreads();
  if (j >= n2) return;
  float d[kMatRows][4];
#pragma unroll
  for (int r = 0; r < kMatRows; ++r)
#pragma unroll
    for (int u = 0; u < 4; ++u) d[r][u] = 0.f;
  const bool full = j + 3 < n2 && (n2 % 4 == 0);
  for (int c = 0; c < C; ++c) {
    const float* p = x2 + ((long long)b * C + c) * n2 + j;
    float xv[4];
    if (full) {
      const float4 t = *reinterpret_cast<const float4*>(p);
      xv[0] = t.x; xv[1] = t.y; xv[2] = t.z; xv[3] = t.w;
    } else {
#pragma unroll
      for (int u = 0; u < 4; ++u) xv[u] = j + u < n2 ? p[u] : 0.f;
    }
#pragma unroll
    for (int r = 0; r < kMatRows; ++r) {
      const float xa = s1[c * kMatRows + r];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float t = xa - xv[u];
        d[r][u] = fmaf(t, t, d[r][u]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < kMatRows; ++r) {
    if (i0 + r >= n1) break;
    float v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      v[u] = d[r][u];
      if (KERN) {
        const float ev = expf(-d[r][u] * coef);
        v[u] = ev >= kCvoThreT ? ev : 0.f;
      }
    }
    float* o = out + ((long long)b * n1 + i0 + r) * n2 + j;
    if (full) *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
    else
      for (int u = 0; u < 4 && j + u < n2; ++u) o[u] = v[u];
  }
}

// out[b][i][j][0..2] = x1_i x x2_j (cross product) or x1_i - x2_j
__global__ void __launch_bounds__(256) cvo_cross_kernel(const float* __restrict__ x1, const float* __restrict__ x2, int n1, int n2,
                                                       int subtract, float* __restrict__ out, long long total) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(e % n2);
    const long long t = e / n2;
    const int i = (int)(t % n1), b = (int)(t / n1);
    const float* p = x1 + (long long)b * 3 * n1 + i;
    const float* q = x2 + (long long)b * 3 * n2 + j;
    const float ax = p[0], ay = p[n1], az = p[2 * (long long)n1];
    const float bx = q[0], by = q[n2], bz = q[2 * (long long)n2];
    float* o = out + e * 3;
    if (subtract) {
      o[0] = ax - bx; o[1] = ay - by; o[2] = az - bz;
    } else {
      o[0] = ay * bz - az * by; o[1] = az * bx - ax * bz; o[2] = ax * by - ay * bx;
    }
  }
}

// ------------------------------------------------------------------ host side
inline int cvo_nsplit(int b, int n1, int n2) {
  const int rows_of_ctas = b * ((n1 + kCvoThreads - 1) / kCvoThreads);
  int want = (4 * kNumSMsB200 + rows_of_ctas - 1) / rows_of_ctas;  // ~4 CTAs of 128 threads per SM
  const int max_split = (n2 + kCvoTile - 1) / kCvoTile;
  if (want > max_split) want = max_split;
  if (want > 64) want = 64;
  return want < 1 ? 1 : want;
}

template <int MODE, int CT>
void cvo_launch_ct(const CvoArgs& a, int b, cudaStream_t st) {
  dim3 grid((a.n1 + kCvoThreads - 1) / kCvoThreads, a.nsplit, b);
  cvo_pair_kernel<CT, MODE><<<grid, kCvoThreads, 0, st>>>(a);
}
template <int MODE>
int cvo_launch(const CvoArgs& a, int ct, int b, cudaStream_t st) {
  switch (ct) {
#define B200_CVO_CASE(N) case N: cvo_launch_ct<MODE, N>(a, b, st); break;
    B200_CVO_CASE(1) B200_CVO_CASE(2) B200_CVO_CASE(3) B200_CVO_CASE(4) B200_CVO_CASE(5) B200_CVO_CASE(6) B200_CVO_CASE(7)
    B200_CVO_CASE(8) B200_CVO_CASE(9) B200_CVO_CASE(10) B200_CVO_CASE(11) B200_CVO_CASE(12) B200_CVO_CASE(13)
    B200_CVO_CASE(14) B200_CVO_CASE(15) B200_CVO_CASE(16)
#undef B200_CVO_CASE
    default: return fail(-1, "cvo: %d channels over all domains (1..%d supported)", ct, kCvoMaxC);
  }
  return check_launch("cvo pair kernel");
}

// fills the per-channel tables; swap = set 2 plays the thread-owned role (gradient of the second frame)
int cvo_fill(const b200_cvo_item* items, int num_items, const float* w1, const float* w2, int b, int n1, int n2, bool swap,
             float* ws, CvoArgs* a, int* ct_out) {
  B200_REQUIRE(items && num_items >= 1 && num_items <= 4, "cvo: 1..4 domains per call");
  B200_REQUIRE(b >= 1 && n1 >= 1 && n2 >= 1 && ws, "cvo: empty point set or null workspace");
  *a = CvoArgs{};
  int ct = 0, ndot = 0;
  for (int k = 0; k < num_items; ++k) {
    const b200_cvo_item& it = items[k];
    B200_REQUIRE(it.x1 && it.x2 && it.c >= 1, "cvo: domain %d has a null point set or no channels", k);
    B200_REQUIRE(ct + it.c <= kCvoMaxC, "cvo: more than %d channels over all domains", kCvoMaxC);
    const bool dot = !(it.dist_coef > 0.f);
    ndot += dot;
    for (int c = 0; c < it.c; ++c, ++ct) {
      a->x1[ct] = (swap ? it.x2 : it.x1) + (long long)c * (swap ? n2 : n1);
      a->x2[ct] = (swap ? it.x1 : it.x2) + (long long)c * (swap ? n1 : n2);
      a->bs1[ct] = (long long)it.c * (swap ? n2 : n1);
      a->bs2[ct] = (long long)it.c * (swap ? n1 : n2);
      a->coef[ct] = dot ? 0.f : 1.f / (2.f * it.dist_coef * it.dist_coef);
      if (dot) a->dot_mask |= 1u << ct;
    }
    a->last_mask |= 1u << (ct - 1);
  }
  B200_REQUIRE(ndot <= 1, "cvo: at most one plain inner-product domain per call (network_modules.py:1012 has one: 'feature')");
  a->w1 = swap ? w2 : w1;
  a->w2 = swap ? w1 : w2;
  a->n1 = swap ? n2 : n1;
  a->n2 = swap ? n1 : n2;
  a->nsplit = cvo_nsplit(b, a->n1, a->n2);
  a->geo0 = -1;
  a->cut_s = -logf(kCvoThreT) * 1.001f + 1e-3f;
  a->ws = ws;
  *ct_out = ct;
  return 0;
}

}  // namespace b200

using namespace b200;

extern "C" {

size_t b200unet_cvo_workspace_bytes(int b, int n1, int n2, int total_c) {
  if (b < 1 || n1 < 1 || n2 < 1 || total_c < 1) return 0;
  const int nmax = n1 > n2 ? n1 : n2;
  const int nsplit = 64;  // upper bound of cvo_nsplit
  return (size_t)nsplit * b * (1 + total_c) * nmax * sizeof(float);
}

int b200unet_cvo_inner_prod_fwd(const b200_cvo_item* items, int num_items, const float* w1, const float* w2, int b, int n1,
                                int n2, int geo_item, float* workspace, float* out, float* wv, void* stream) {
  CvoArgs a;
  int ct = 0;
  if (int rc = cvo_fill(items, num_items, w1, w2, b, n1, n2, false, workspace, &a, &ct)) return rc;
  B200_REQUIRE(out, "cvo_inner_prod_fwd: null output");
  cudaStream_t st = as_stream(stream);
  const float *gx = nullptr, *gy = nullptr, *gz = nullptr;
  long long gbs = 0;
  if (wv) {
    B200_REQUIRE(geo_item >= 0 && geo_item < num_items && items[geo_item].c == 3,
                 "cvo_inner_prod_fwd: wv needs the index of the 3-channel geometry domain");
    int c0 = 0;
    for (int k = 0; k < geo_item; ++k) c0 += items[k].c;
    a.geo0 = c0;
    gx = a.x1[c0]; gy = a.x1[c0 + 1]; gz = a.x1[c0 + 2];
    gbs = a.bs1[c0];
    if (int rc = cvo_launch<CVO_FWD_WV>(a, ct, b, st)) return rc;
  } else {
    if (int rc = cvo_launch<CVO_FWD>(a, ct, b, st)) return rc;
  }
  cvo_fwd_finalize_kernel<<<b, 1024, 0, st>>>(a.ws, a.nsplit, wv ? 4 : 1, a.n1, a.w1, gx, gy, gz, gbs, out, wv);
  return check_launch("cvo_fwd_finalize");
}

int b200unet_cvo_inner_prod_bwd(const b200_cvo_item* items, int num_items, const float* w1, const float* w2, int b, int n1,
                                int n2, const float* grad_out, float* workspace, float* const* dx1, float* const* dx2,
                                float* dw1, float* dw2, void* stream) {
  cudaStream_t st = as_stream(stream);
  for (int side = 0; side < 2; ++side) {
    float* const* dx = side == 0 ? dx1 : dx2;
    float* dw = side == 0 ? dw1 : dw2;
    bool any = dw != nullptr;
    for (int k = 0; k < num_items && dx; ++k) any |= dx[k] != nullptr;
    if (!any) continue;
    CvoArgs a;
    int ct = 0;
    if (int rc = cvo_fill(items, num_items, w1, w2, b, n1, n2, side == 1, workspace, &a, &ct)) return rc;
    if (int rc = cvo_launch<CVO_BWD>(a, ct, b, st)) return rc;
    CvoBwdOut o{};
    int c0 = 0;
    for (int k = 0; k < num_items; ++k) {
      for (int c = 0; c < items[k].c; ++c) {
        o.dx[c0 + c] = (dx && dx[k]) ? dx[k] + (long long)c * a.n1 : nullptr;
        o.bs[c0 + c] = (long long)items[k].c * a.n1;
      }
      c0 += items[k].c;
    }
    o.dw = dw;
    o.grad = grad_out;
    o.mat = 0;
    dim3 grid((a.n1 + 255) / 256, b);
    cvo_bwd_finalize_kernel<<<grid, 256, 0, st>>>(a, o, ct, b);
    if (int rc = check_launch("cvo_bwd_finalize")) return rc;
  }
  return 0;
}

static int cvo_mat_fwd(const float* x1, const float* x2, int b, int c, int n1, int n2, float dist_coef, bool kern, float* out,
                       void* stream) {
  B200_REQUIRE(x1 && x2 && out && b >= 1 && c >= 1 && n1 >= 1 && n2 >= 1, "cvo matrix forward: null or empty argument");
  B200_REQUIRE(c <= 512, "cvo matrix forward: at most 512 channels");
  B200_REQUIRE(!kern || dist_coef > 0.f, "kern_mat: dist_coef must be positive");
  B200_REQUIRE(reinterpret_cast<uintptr_t>(x2) % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0,
               "cvo matrix forward: x2 and out must be 16-byte aligned");
  dim3 grid((n2 + 511) / 512, (n1 + kMatRows - 1) / kMatRows, b);
  const size_t smem = (size_t)c * kMatRows * sizeof(float);
  const float coef = kern ? 1.f / (2.f * dist_coef * dist_coef) : 0.f;
  if (kern) cvo_mat_fwd_kernel<true><<<grid, 128, smem, as_stream(stream)>>>(x1, x2, c, n1, n2, coef, out);
  else cvo_mat_fwd_kernel<false><<<grid, 128, smem, as_stream(stream)>>>(x1, x2, c, n1, n2, coef, out);
  return check_launch("cvo matrix forward");
}

static int cvo_mat_bwd(const float* dy, const float* x1, const float* x2, int b, int c, int n1, int n2, float dist_coef,
                       bool kern, float* workspace, float* dx1, float* dx2, void* stream) {
  B200_REQUIRE(dy && (dx1 || dx2), "cvo matrix backward: null gradient");
  B200_REQUIRE(!kern || dist_coef > 0.f, "kern_mat: dist_coef must be positive");
  b200_cvo_item it{x1, x2, c, kern ? dist_coef : 1.f};
  cudaStream_t st = as_stream(stream);
  for (int side = 0; side < 2; ++side) {
    float* dx = side == 0 ? dx1 : dx2;
    if (!dx) continue;
    CvoArgs a;
    int ct = 0;
    if (int rc = cvo_fill(&it, 1, nullptr, nullptr, b, n1, n2, side == 1, workspace, &a, &ct)) return rc;
    a.dy = dy;
    a.dy_sb = (long long)n1 * n2;
    a.dy_s1 = side == 0 ? n2 : 1;
    a.dy_s2 = side == 0 ? 1 : n2;
    a.plain_distance = kern ? 0 : 1;
    if (int rc = cvo_launch<CVO_MAT_BWD>(a, ct, b, st)) return rc;
    CvoBwdOut o{};
    for (int k = 0; k < c; ++k) {
      o.dx[k] = dx + (long long)k * a.n1;
      o.bs[k] = (long long)c * a.n1;
    }
    o.mat = 1;
    dim3 grid((a.n1 + 255) / 256, b);
    cvo_bwd_finalize_kernel<<<grid, 256, 0, st>>>(a, o, ct, b);
    if (int rc = check_launch("cvo matrix backward finalize")) return rc;
  }
  return 0;
}

int b200unet_cvo_sub_norm_fwd(const float* x1, const float* x2, int b, int c, int n1, int n2, float* out, void* stream) {
  return cvo_mat_fwd(x1, x2, b, c, n1, n2, 0.f, false, out, stream);
}
int b200unet_cvo_sub_norm_bwd(const float* dy, const float* x1, const float* x2, int b, int c, int n1, int n2,
                              float* workspace, float* dx1, float* dx2, void* stream) {
  return cvo_mat_bwd(dy, x1, x2, b, c, n1, n2, 0.f, false, workspace, dx1, dx2, stream);
}
int b200unet_cvo_kern_mat_fwd(const float* x1, const float* x2, int b, int c, int n1, int n2, float dist_coef, float* out,
                              void* stream) {
  return cvo_mat_fwd(x1, x2, b, c, n1, n2, dist_coef, true, out, stream);
}
int b200unet_cvo_kern_mat_bwd(const float* dy, const float* x1, const float* x2, int b, int c, int n1, int n2,
                              float dist_coef, float* workspace, float* dx1, float* dx2, void* stream) {
  return cvo_mat_bwd(dy, x1, x2, b, c, n1, n2, dist_coef, true, workspace, dx1, dx2, stream);
}
int b200unet_cvo_cross_fwd(const float* x1, const float* x2, int b, int n1, int n2, int subtract, float* out, void* stream) {
  B200_REQUIRE(x1 && x2 && out && b >= 1 && n1 >= 1 && n2 >= 1, "cvo_cross_fwd: null or empty argument");
  const long long total = (long long)b * n1 * n2;
  cvo_cross_kernel<<<stream_grid(total), 256, 0, as_stream(stream)>>>(x1, x2, n1, n2, subtract, out, total);
  return check_launch("cvo_cross_fwd");
}

}  // extern "C"
