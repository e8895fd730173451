// The classifier head: nn.Conv2d(prev, n_classes, 1) [+ ReLU when non_neg] (unet.py:65-71, 84), alone or fused
// with F.cross_entropy(mean, ignore_index=-100) (README.md:58).  HBM-bound: the 64-channel full-resolution
// activation is read once per pass; logits never round-trip through HBM in the fused-loss path.
//
// Thread mapping: 8 lanes cooperate on one pixel, each lane owns one 16-byte vector of every 64-channel chunk,
// so a warp reads 4 pixels x 128 contiguous bytes per instruction.  blockIdx.y selects the 64-channel chunk whose
// dx / dW this block produces (the dot product itself always runs over all channels).
#include "chan_reduce.cuh"

namespace b200 {

constexpr int kHeadThreads = 256;
constexpr int kHeadMaxBlocks = 8 * kNumSMsB200;  // HBM-bound, one 16-byte load in flight per thread: needs many resident warps
constexpr long long kIgnoreIndex = -100;

enum { HEAD_FWD = 0, HEAD_CE_FWD = 1, HEAD_BWD = 2, HEAD_CE_BWD = 3 };

struct HeadArgs {
  DView x;
  const float* w;      // [k][c]
  const float* b;      // [k]
  int k, relu;
  float* logits;       // NCHW fp32 or null
  const long long* labels;
  const float* dlogits;  // NCHW fp32 (HEAD_BWD)
  const float* gscale;   // device scalar: upstream gradient of the loss (HEAD_CE_BWD), may be null (=1)
  const float* ce_state; // [0] loss, [1] 1/valid_count (written by the CE forward finalize)
  DView dx;
  const bf16* mask;    // laid out like dx
  float* ws;           // per-block partials
};

template <int KMAX, int VEC, int MODE>
__global__ void __launch_bounds__(kHeadThreads) head_kernel(HeadArgs a) {
  extern __shared__ float smem[];
  const int c = a.x.c, K = a.k;
  float* sw = smem;           // [K][c]
  float* sb = sw + K * c;     // [K]
  float* red = sb + 8;        // [8 warps][KMAX*VEC*8 + KMAX + 2]
  for (int i = threadIdx.x; i < K * c; i += blockDim.x) sw[i] = a.w[i];
  if (threadIdx.x < 8) sb[threadIdx.x] = (threadIdx.x < K && a.b) ? a.b[threadIdx.x] : 0.f;
  __syncthreads();

  const int lane8 = threadIdx.x & 7;
  const int slot = threadIdx.x >> 3;
  const int slots = kHeadThreads / 8;
  const int nvec = c / VEC;                    // vectors per pixel
  const int chunk_v0 = blockIdx.y * 8;         // first vector of this block's chunk
  const int myv = chunk_v0 + lane8;            // the vector whose dx/dW this lane produces
  const long long hw = (long long)a.x.h * a.x.w;
  const long long npix = hw * a.x.n;

  float dw_acc[KMAX][VEC];
  float db_acc[KMAX];
  float loss_acc = 0.f, cnt_acc = 0.f;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    db_acc[k] = 0.f;
#pragma unroll
    for (int j = 0; j < VEC; ++j) dw_acc[k][j] = 0.f;
  }
  float gs = 1.f;
  if (MODE == HEAD_CE_BWD) gs = (a.gscale ? a.gscale[0] : 1.f) * a.ce_state[1];

  // work unit = `slots` consecutive pixels of one image row: only 32-bit index arithmetic per iteration (the 64-bit
  // divisions of a flat pixel index sat in front of every load)
  const unsigned upr = (unsigned)(a.x.w + slots - 1) / slots;       // units per row
  const unsigned units = (unsigned)a.x.n * a.x.h * upr;
  (void)npix;
  for (unsigned u = blockIdx.x; u < units; u += gridDim.x) {
    const unsigned row = u / upr, cb = u - row * upr;
    const int n = (int)(row / a.x.h), ih = (int)(row - (unsigned)n * a.x.h);
    const int iw = (int)(cb * slots) + slot;
    const bool live = iw < a.x.w;
    const long long p = (long long)row * a.x.w + iw;
    const bf16* xp = a.x.p + a.x.off(n, ih, iw);
    float z[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) z[k] = 0.f;
    float xmine[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) xmine[j] = 0.f;
    // backward: request the ReLU mask of this lane's dx vector together with x, not after the softmax that depends on x
    // (two dependent memory round trips per iteration otherwise).  Without BatchNorm the mask IS x: nothing to load.
    const bool mask_is_x = a.mask == a.x.p && a.dx.sn == a.x.sn && a.dx.sh == a.x.sh && a.dx.sw == a.x.sw;
    bf16x8 mvec = make_uint4(0, 0, 0, 0);
    if ((MODE == HEAD_BWD || MODE == HEAD_CE_BWD) && VEC == 8 && a.mask && !mask_is_x && live && myv < nvec && a.dx.p)
      mvec = *reinterpret_cast<const bf16x8*>(a.mask + a.dx.off(n, ih, iw) + myv * VEC);
    if (live) {
      for (int v = lane8; v < nvec; v += 8) {
        float xv[VEC];
        if (VEC == 8) {
          float t[8];
          unpack8(*reinterpret_cast<const bf16x8*>(xp + v * 8), t);
#pragma unroll
          for (int j = 0; j < VEC; ++j) xv[j] = t[j];
          if (a.x.lo) {  // split tier (forward modes): x = hi + lo
            unpack8(*reinterpret_cast<const bf16x8*>(a.x.lo + (xp - a.x.p) + v * 8), t);
#pragma unroll
            for (int j = 0; j < VEC; ++j) xv[j] += t[j];
          }
        } else {
          xv[0] = bf2f(xp[v]);
        }
        if (v == myv) {
#pragma unroll
          for (int j = 0; j < VEC; ++j) xmine[j] = xv[j];
        }
#pragma unroll
        for (int k = 0; k < KMAX; ++k)
          if (k < K) {
#pragma unroll
            for (int j = 0; j < VEC; ++j) z[k] += xv[j] * sw[k * c + v * VEC + j];
          }
      }
    }
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      z[k] += __shfl_xor_sync(0xffffffffu, z[k], 1);
      z[k] += __shfl_xor_sync(0xffffffffu, z[k], 2);
      z[k] += __shfl_xor_sync(0xffffffffu, z[k], 4);
      z[k] += sb[k];
    }
    float zr[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) zr[k] = a.relu ? fmaxf(z[k], 0.f) : z[k];

    if ((MODE == HEAD_FWD || MODE == HEAD_CE_FWD) && a.logits && live) {
#pragma unroll
      for (int k = 0; k < KMAX; ++k)
        if (k == lane8 && k < K) a.logits[((long long)n * K + k) * hw + (p - n * hw)] = zr[k];
    }
    if (MODE == HEAD_FWD) continue;

    long long label = 0;
    bool valid = live;
    float lse = 0.f;
    if (MODE == HEAD_CE_FWD || MODE == HEAD_CE_BWD) {
      if (live) label = a.labels[p];
      valid = live && label != kIgnoreIndex;
      float m = -INFINITY;
#pragma unroll
      for (int k = 0; k < KMAX; ++k)
        if (k < K) m = fmaxf(m, zr[k]);
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < KMAX; ++k)
        if (k < K) s += __expf(zr[k] - m);
      lse = m + __logf(s);
    }
    if (MODE == HEAD_CE_FWD) {
      if (valid && lane8 == 0) {
        // a label outside [0, K) that is not ignore_index is an error (F.cross_entropy raises a device-side assert):
        // it poisons the loss with NaN instead of silently counting as some class
        float zy = 0.f;
#pragma unroll
        for (int k = 0; k < KMAX; ++k)
          if (k == (int)label) zy = zr[k];
        if (label < 0 || label >= K) zy = __int_as_float(0x7fc00000);
        loss_acc += lse - zy;
        cnt_acc += 1.f;
      }
      continue;
    }

    // ---- backward: dz, then dx / dW for this block's channel chunk
    float dz[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      float g = 0.f;
      if (k < K && valid) {
        if (MODE == HEAD_CE_BWD)
          g = gs * (__expf(zr[k] - lse) - (k == (int)label ? 1.f : 0.f));
        else
          g = a.dlogits[((long long)n * K + k) * hw + (p - n * hw)];
        if (a.relu && !(z[k] > 0.f)) g = 0.f;
      }
      dz[k] = g;
    }
    if (live && myv < nvec) {
      float r[VEC];
#pragma unroll
      for (int j = 0; j < VEC; ++j) r[j] = 0.f;
#pragma unroll
      for (int k = 0; k < KMAX; ++k)
        if (k < K) {
#pragma unroll
          for (int j = 0; j < VEC; ++j) {
            r[j] += dz[k] * sw[k * c + myv * VEC + j];
            dw_acc[k][j] += dz[k] * xmine[j];
          }
        }
      if (a.dx.p) {
        const long long o = a.dx.off(n, ih, iw) + myv * VEC;
        if (a.mask) {
          if (VEC == 8) {
            float t[8];
            unpack8(mvec, t);
#pragma unroll
            for (int j = 0; j < VEC; ++j) r[j] = (mask_is_x ? xmine[j] : t[j]) > 0.f ? r[j] : 0.f;
          } else {
            r[0] = bf2f(a.mask[o]) > 0.f ? r[0] : 0.f;
          }
        }
        if (VEC == 8) {
          float t[8];
#pragma unroll
          for (int j = 0; j < VEC; ++j) t[j] = r[j];
          *reinterpret_cast<bf16x8*>(a.dx.p + o) = pack8(t);
        } else {
          a.dx.p[o] = f2bf(r[0]);
        }
      }
    }
    if (lane8 == 0) {
#pragma unroll
      for (int k = 0; k < KMAX; ++k) db_acc[k] += dz[k];
    }
  }
  if (MODE == HEAD_FWD) return;

  // ---- deterministic block reduction: across the 4 pixel slots of a warp by shuffle, across warps in shared memory
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kRow = KMAX * VEC * 8 + KMAX + 2;
  if (MODE == HEAD_CE_FWD) {
    loss_acc += __shfl_xor_sync(0xffffffffu, loss_acc, 8);
    loss_acc += __shfl_xor_sync(0xffffffffu, loss_acc, 16);
    cnt_acc += __shfl_xor_sync(0xffffffffu, cnt_acc, 8);
    cnt_acc += __shfl_xor_sync(0xffffffffu, cnt_acc, 16);
    if (lane == 0) {
      red[warp * kRow + 0] = loss_acc;
      red[warp * kRow + 1] = cnt_acc;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      float l = 0.f, cn = 0.f;
      for (int wdx = 0; wdx < kHeadThreads / 32; ++wdx) {
        l += red[wdx * kRow + 0];
        cn += red[wdx * kRow + 1];
      }
      a.ws[blockIdx.x * 2 + 0] = l;
      a.ws[blockIdx.x * 2 + 1] = cn;
    }
    return;
  }
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      float v = dw_acc[k][j];
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if (lane < 8) red[warp * kRow + (k * 8 + lane) * VEC + j] = v;
    }
    float d = db_acc[k];
    d += __shfl_xor_sync(0xffffffffu, d, 8);
    d += __shfl_xor_sync(0xffffffffu, d, 16);
    if (lane == 0) red[warp * kRow + KMAX * VEC * 8 + k] = d;
  }
  __syncthreads();
  // partial layout: ws[blockIdx.x][K*c + K]
  float* out = a.ws + (long long)blockIdx.x * (K * c + K);
  for (int q = threadIdx.x; q < KMAX * 8 * VEC; q += blockDim.x) {
    const int k = q / (8 * VEC), rem = q - k * 8 * VEC;
    const int ch = chunk_v0 * VEC + rem;
    if (k < K && ch < c) {
      float s = 0.f;
      for (int wdx = 0; wdx < kHeadThreads / 32; ++wdx) s += red[wdx * kRow + q];
      out[k * c + ch] = s;
    }
  }
  if (blockIdx.y == 0 && threadIdx.x < K) {
    float s = 0.f;
    for (int wdx = 0; wdx < kHeadThreads / 32; ++wdx) s += red[wdx * kRow + KMAX * VEC * 8 + threadIdx.x];
    out[K * c + threadIdx.x] = s;
  }
}

__global__ void head_ce_finalize_kernel(const float* __restrict__ ws, int blocks, float* __restrict__ loss,
                                        float* __restrict__ ce_state) {
  if (threadIdx.x >= 32 || blockIdx.x != 0) return;
  const double l = warp_partial_sum(ws, blocks, 2, 0);
  const double cn = warp_partial_sum(ws, blocks, 2, 1);
  if (threadIdx.x != 0) return;
  const float v = (float)(l / cn);  // 0/0 = NaN when every label is ignored, like F.cross_entropy
  if (loss) loss[0] = v;
  ce_state[0] = v;
  ce_state[1] = cn > 0.0 ? (float)(1.0 / cn) : 0.f;
}

__global__ void head_bwd_finalize_kernel(const float* __restrict__ ws, int blocks, int kc, int k, float* __restrict__ dw,
                                         float* __restrict__ db) {
  const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);  // one warp per output
  if (q >= kc + k) return;
  const double s = warp_partial_sum(ws, blocks, kc + k, q);
  if ((threadIdx.x & 31) != 0) return;
  if (q < kc)
    dw[q] = (float)s;
  else if (db)
    db[q - kc] = (float)s;
}

template <int MODE>
int launch_head(const HeadArgs& a, int blocks, cudaStream_t st) {
  const int c = a.x.c, K = a.k;
  const bool v8 = c % 8 == 0 && reinterpret_cast<uintptr_t>(a.x.p) % 16 == 0 && a.x.sw % 8 == 0 && a.x.sh % 8 == 0 &&
                  a.x.sn % 8 == 0 &&
                  (!a.dx.p || (reinterpret_cast<uintptr_t>(a.dx.p) % 16 == 0 && a.dx.sw % 8 == 0 && a.dx.sh % 8 == 0 &&
                               a.dx.sn % 8 == 0 && reinterpret_cast<uintptr_t>(a.mask) % 16 == 0));
  if (a.x.lo && !(v8 && reinterpret_cast<uintptr_t>(a.x.lo) % 16 == 0))
    return fail(-1, "head: the split tier needs channel counts / strides that are multiples of 8");
  const int vec = v8 ? 8 : 1;
  const int nvec = c / vec;
  const int chunks = (MODE == HEAD_BWD || MODE == HEAD_CE_BWD) ? (nvec + 7) / 8 : 1;
  const int kmax = K <= 2 ? 2 : (K <= 4 ? 4 : 8);
  const size_t smem = (size_t)(K * c + 8 + (kHeadThreads / 32) * (kmax * vec * 8 + kmax + 2)) * sizeof(float);
  if (smem > 48 * 1024) return fail(-1, "head: n_classes*channels too large for shared memory (%d x %d)", K, c);
  dim3 grid(blocks, chunks);
#define B200_HEAD_LAUNCH(KM, V) head_kernel<KM, V, MODE><<<grid, kHeadThreads, smem, st>>>(a)
  if (v8) {
    if (kmax == 2) B200_HEAD_LAUNCH(2, 8);
    else if (kmax == 4) B200_HEAD_LAUNCH(4, 8);
    else B200_HEAD_LAUNCH(8, 8);
  } else {
    if (kmax == 2) B200_HEAD_LAUNCH(2, 1);
    else if (kmax == 4) B200_HEAD_LAUNCH(4, 1);
    else B200_HEAD_LAUNCH(8, 1);
  }
#undef B200_HEAD_LAUNCH
  return check_launch("head");
}

inline int head_blocks(long long npix) {
  long long b = (npix + (kHeadThreads / 8) * 4 - 1) / ((kHeadThreads / 8) * 4);
  if (b < 1) b = 1;
  if (b > kHeadMaxBlocks) b = kHeadMaxBlocks;
  return (int)b;
}

}  // namespace b200

using namespace b200;

extern "C" {

size_t b200unet_head_workspace_bytes(int c, int n_classes) {
  return (size_t)kHeadMaxBlocks * ((size_t)n_classes * c + n_classes + 2) * sizeof(float);
}

int b200unet_head_fwd(const b200_view* x, const float* w, const float* b, int n_classes, int relu, float* logits_nchw,
                      void* stream) {
  B200_REQUIRE(view_ok(x) && w && logits_nchw && n_classes >= 1 && n_classes <= 8, "head_fwd: bad arguments");
  HeadArgs a{};
  a.x = dview(*x);
  a.w = w;
  a.b = b;
  a.k = n_classes;
  a.relu = relu;
  a.logits = logits_nchw;
  return launch_head<HEAD_FWD>(a, head_blocks(view_pixels(*x)), as_stream(stream));
}

int b200unet_head_bwd(const b200_view* x, const float* w, const float* b, int n_classes, int relu,
                      const float* dlogits_nchw, const b200_view* dx, const void* mask, float* dw, float* db,
                      void* workspace, size_t workspace_bytes, void* stream) {
  B200_REQUIRE(view_ok(x) && w && dlogits_nchw && dw && workspace && n_classes >= 1 && n_classes <= 8,
               "head_bwd: bad arguments");
  B200_REQUIRE(!dx || (view_ok(dx) && same_extent(*x, *dx)), "head_bwd: dx extent differs from x");
  B200_REQUIRE(workspace_bytes >= b200unet_head_workspace_bytes(x->c, n_classes), "head_bwd: workspace too small");
  HeadArgs a{};
  a.x = dview(*x);
  a.w = w;
  a.b = b;
  a.k = n_classes;
  a.relu = relu;
  a.dlogits = dlogits_nchw;
  if (dx) a.dx = dview(*dx);
  a.mask = (const bf16*)mask;
  a.ws = (float*)workspace;
  const int blocks = head_blocks(view_pixels(*x));
  int r = launch_head<HEAD_BWD>(a, blocks, as_stream(stream));
  if (r) return r;
  const int kc = n_classes * x->c;
  head_bwd_finalize_kernel<<<finalize_grid(kc + n_classes), kFinalizeThreads, 0, as_stream(stream)>>>((const float*)workspace, blocks,
                                                                                        kc, n_classes, dw, db);
  return check_launch("head_bwd finalize");
}

int b200unet_head_ce_fwd(const b200_view* x, const float* w, const float* b, int n_classes, int relu,
                         const int64_t* labels, float* loss, float* logits_nchw, float* ce_state, void* workspace,
                         size_t workspace_bytes, void* stream) {
  B200_REQUIRE(view_ok(x) && w && labels && ce_state && workspace && n_classes >= 1 && n_classes <= 8,
               "head_ce_fwd: bad arguments");
  B200_REQUIRE(workspace_bytes >= b200unet_head_workspace_bytes(x->c, n_classes), "head_ce_fwd: workspace too small");
  HeadArgs a{};
  a.x = dview(*x);
  a.w = w;
  a.b = b;
  a.k = n_classes;
  a.relu = relu;
  a.logits = logits_nchw;
  a.labels = (const long long*)labels;
  a.ws = (float*)workspace;
  const int blocks = head_blocks(view_pixels(*x));
  int r = launch_head<HEAD_CE_FWD>(a, blocks, as_stream(stream));
  if (r) return r;
  head_ce_finalize_kernel<<<1, 32, 0, as_stream(stream)>>>((const float*)workspace, blocks, loss, ce_state);
  return check_launch("head_ce finalize");
}

int b200unet_head_ce_bwd(const b200_view* x, const float* w, const float* b, int n_classes, int relu,
                         const int64_t* labels, const float* grad_scale, const float* ce_state, const b200_view* dx,
                         const void* mask, float* dw, float* db, void* workspace, size_t workspace_bytes, void* stream) {
  B200_REQUIRE(view_ok(x) && w && labels && ce_state && dw && workspace && n_classes >= 1 && n_classes <= 8,
               "head_ce_bwd: bad arguments");
  B200_REQUIRE(!dx || (view_ok(dx) && same_extent(*x, *dx)), "head_ce_bwd: dx extent differs from x");
  B200_REQUIRE(workspace_bytes >= b200unet_head_workspace_bytes(x->c, n_classes), "head_ce_bwd: workspace too small");
  HeadArgs a{};
  a.x = dview(*x);
  a.w = w;
  a.b = b;
  a.k = n_classes;
  a.relu = relu;
  a.labels = (const long long*)labels;
  a.gscale = grad_scale;
  a.ce_state = ce_state;
  if (dx) a.dx = dview(*dx);
  a.mask = (const bf16*)mask;
  a.ws = (float*)workspace;
  const int blocks = head_blocks(view_pixels(*x));
  int r = launch_head<HEAD_CE_BWD>(a, blocks, as_stream(stream));
  if (r) return r;
  const int kc = n_classes * x->c;
  head_bwd_finalize_kernel<<<finalize_grid(kc + n_classes), kFinalizeThreads, 0, as_stream(stream)>>>((const float*)workspace, blocks,
                                                                                        kc, n_classes, dw, db);
  return check_launch("head_ce_bwd finalize");
}
}
