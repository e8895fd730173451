// The classifier head: nn.Conv2d(prev, n_classes, 1) [+ ReLU when non_neg] (unet.py:65-71, 84), alone or fused
// with F.cross_entropy(mean, ignore_index=-100) (README.md:58).  HBM-bound: the 64-channel full-resolution
// activation is read once per pass; logits never round-trip through HBM in the fused-loss path.
//
// Thread mapping: 8 lanes cooperate on one pixel, each lane owns one 16-byte vector of every 64-channel chunk,
// so a warp reads 4 pixels x 128 contiguous bytes per instruction.  blockIdx.y selects the 64-channel chunk whose
// dx / dW this block produces (the dot product itself always runs over all channels).
#include <cstdlib>

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


// ------------------------------------------------------------------ fast path (dense tensors, 8..64 channels)
// The first version of this kernel (head_kernel above, kept for strided views / wide heads) has ONE 16-byte load in
// flight per thread and iteration and reached 1.5-2.0 TB/s (ncu: long-scoreboard stalls, l1tex 52 %).  Here every thread
// owns U pixels per iteration: all U loads (and the labels) are issued before any arithmetic, LPP = C/8 lanes cooperate
// on a pixel (a warp instruction reads 512 contiguous bytes), the classifier weights of the thread's 8 channels live in
// registers, and the loop is grid-strided over contiguous chunks of (256/LPP)*U pixels.
constexpr int kHeadDenseU = 4;   // pixels per thread and stage
constexpr int kHeadStages = 4;   // cp.async ring depth
__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_8(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
inline size_t head_ring_bytes(int lpp, bool split) {
  return (size_t)kHeadStages * kHeadDenseU * (kHeadThreads * 16 * (split ? 2 : 1) + (kHeadThreads / lpp) * 8);
}
template <int KMAX, int LPP, int MODE, int U, bool SPLIT>
__global__ void __launch_bounds__(kHeadThreads, 2) head_dense_kernel(HeadArgs a, long long npix) {
  constexpr int PPB = kHeadThreads / LPP;   // pixels per block and sub-iteration
  constexpr bool BWD = MODE == HEAD_BWD || MODE == HEAD_CE_BWD;
  constexpr bool CE = MODE == HEAD_CE_FWD || MODE == HEAD_CE_BWD;
  __shared__ float red[kHeadThreads / 32][KMAX * LPP * 8 + KMAX + 2];
  const int K = a.k, c = LPP * 8;
  const int lip = threadIdx.x % LPP;        // lane in pixel: owns channels lip*8 .. +7
  const int slot = threadIdx.x / LPP;
  float w[KMAX][8], bias[KMAX];
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    bias[k] = (k < K && a.b) ? a.b[k] : 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) w[k][j] = k < K ? a.w[k * c + lip * 8 + j] : 0.f;
  }
  float dw_acc[BWD ? KMAX : 1][8], db_acc[KMAX];
  float loss_acc = 0.f, cnt_acc = 0.f;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    db_acc[k] = 0.f;
    if (BWD) {
#pragma unroll
      for (int j = 0; j < 8; ++j) dw_acc[k][j] = 0.f;
    }
  }
  float gs = 1.f;
  if (MODE == HEAD_CE_BWD) gs = (a.gscale ? a.gscale[0] : 1.f) * a.ce_state[1];
  const long long hw = (long long)a.x.h * a.x.w;
  const bf16* __restrict__ xp = a.x.p;
  const bf16* __restrict__ xlo = SPLIT ? a.x.lo : nullptr;

  // Multi-stage cp.async ring (thread-private 16-byte slots, so no block barrier): S stages of U vectors per thread
  // are in flight — S*U*16 B = 256 B per thread, 128 KB per SM at 2 blocks — without holding them in registers.  With one
  // round of register loads per iteration (512 threads x 64 B per SM and memory round trip) the kernel was latency-bound
  // at 1.6 TB/s; prefetching the next round into registers cost occupancy and was slower still.
  extern __shared__ __align__(16) uint8_t ring_raw[];
  uint4* xb = reinterpret_cast<uint4*>(ring_raw);                                        // [S][U][256]
  uint4* xlb = xb + (SPLIT ? kHeadStages * U * kHeadThreads : 0);                        // [S][U][256] (split tier)
  long long* lb = reinterpret_cast<long long*>(xlb + kHeadStages * U * kHeadThreads);    // [S][U][PPB]
  const long long stride = (long long)gridDim.x * (PPB * U);
  const long long base0 = (long long)blockIdx.x * (PPB * U);
  auto issue = [&](int stage, long long b) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long p = b + u * PPB + slot;
      if (p < npix) {
        cp_async_16(&xb[(stage * U + u) * kHeadThreads + threadIdx.x], xp + p * c + lip * 8);
        if (SPLIT) cp_async_16(&xlb[(stage * U + u) * kHeadThreads + threadIdx.x], xlo + p * c + lip * 8);
        if (CE && lip == 0) cp_async_8(&lb[(stage * U + u) * PPB + slot], a.labels + p);
      }
    }
    cp_async_commit();
  };
#pragma unroll
  for (int st = 0; st < kHeadStages - 1; ++st) issue(st, base0 + st * stride);
  int stage = 0;
  for (long long base = base0; base < npix; base += stride) {
    __syncwarp();  // every lane has read the label slots of the stage that is refilled now
    issue(stage == 0 ? kHeadStages - 1 : stage - 1, base + (kHeadStages - 1) * stride);
    cp_async_wait<kHeadStages - 1>();
    __syncwarp();  // labels are written by lane lip == 0 of each pixel
    bf16x8 xv[U], xl[SPLIT ? U : 1];
    int lab[U];
    bool live[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long p = base + u * PPB + slot;
      live[u] = p < npix;
      xv[u] = make_uint4(0, 0, 0, 0);
      if (SPLIT) xl[u] = make_uint4(0, 0, 0, 0);
      lab[u] = -1;  // class, or -1 = ignore_index, or -2 = out of range (poisons the loss)
      if (live[u]) {
        xv[u] = xb[(stage * U + u) * kHeadThreads + threadIdx.x];
        if (SPLIT) xl[u] = xlb[(stage * U + u) * kHeadThreads + threadIdx.x];
        if (CE) {
          const long long l64 = lb[(stage * U + u) * PPB + slot];
          lab[u] = l64 == kIgnoreIndex ? -1 : ((l64 < 0 || l64 >= K) ? -2 : (int)l64);
        }
      }
    }
    stage = stage + 1 == kHeadStages ? 0 : stage + 1;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long p = base + u * PPB + slot;
      float x8[8];
      unpack8(xv[u], x8);
      if (SPLIT) {  // split tier (forward modes): x = hi + lo
        float t[8];
        unpack8(xl[u], t);
#pragma unroll
        for (int j = 0; j < 8; ++j) x8[j] += t[j];
      }
      float z[KMAX];
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) s += x8[j] * w[k][j];
#pragma unroll
        for (int o = 1; o < LPP; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        z[k] = s + bias[k];
      }
      float zr[KMAX];
#pragma unroll
      for (int k = 0; k < KMAX; ++k) zr[k] = a.relu ? fmaxf(z[k], 0.f) : z[k];
      const long long n = p / hw, r = p - n * hw;
      if (!BWD && a.logits && live[u]) {
#pragma unroll
        for (int k = 0; k < KMAX; ++k)
          if (k < K && k % LPP == lip) a.logits[(n * K + k) * hw + r] = zr[k];
      }
      if (MODE == HEAD_FWD) continue;
      bool valid = live[u];
      float lse = 0.f;
      if (CE) {
        valid = live[u] && lab[u] != -1;
        float m = -INFINITY;
#pragma unroll
        for (int k = 0; k < KMAX; ++k)
          if (k < K) m = fmaxf(m, zr[k]);
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < KMAX; ++k)
          if (k < K) sum += __expf(zr[k] - m);
        lse = m + __logf(sum);
      }
      if (MODE == HEAD_CE_FWD) {
        if (valid && lip == 0) {
          float zy = 0.f;
#pragma unroll
          for (int k = 0; k < KMAX; ++k)
            if (k == lab[u]) zy = zr[k];
          if (lab[u] == -2) zy = __int_as_float(0x7fc00000);  // out-of-range label: poison the loss
          loss_acc += lse - zy;
          cnt_acc += 1.f;
        }
        continue;
      }
      if (BWD) {
        float dz[KMAX];
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
          float g = 0.f;
          if (k < K && valid) {
            if (MODE == HEAD_CE_BWD)
              g = gs * (__expf(zr[k] - lse) - (k == lab[u] ? 1.f : 0.f));
            else
              g = a.dlogits[(n * K + k) * hw + r];
            if (a.relu && !(z[k] > 0.f)) g = 0.f;
          }
          dz[k] = g;
        }
        float rr[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) rr[j] = 0.f;
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            rr[j] += dz[k] * w[k][j];
            dw_acc[k][j] += dz[k] * x8[j];
          }
          if (lip == 0) db_acc[k] += dz[k];
        }
        if (a.dx.p && live[u]) {
          if (a.mask) {  // fast path: the mask IS x (no BatchNorm): ReLU backward of the last convolution
#pragma unroll
            for (int j = 0; j < 8; ++j) rr[j] = x8[j] > 0.f ? rr[j] : 0.f;
          }
          *reinterpret_cast<bf16x8*>(a.dx.p + p * c + lip * 8) = pack8(rr);
        }
      }
    }
  }
  if (MODE == HEAD_FWD) return;

  // ---- deterministic block reduction: lanes with the same `lip` by shuffle, warps through shared memory
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kRowD = KMAX * LPP * 8;
  if (MODE == HEAD_CE_FWD) {
#pragma unroll
    for (int o = LPP; o < 32; o <<= 1) {
      loss_acc += __shfl_xor_sync(0xffffffffu, loss_acc, o);
      cnt_acc += __shfl_xor_sync(0xffffffffu, cnt_acc, o);
    }
    if (lane == 0) {
      red[warp][0] = loss_acc;
      red[warp][1] = cnt_acc;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      float l = 0.f, cn = 0.f;
      for (int wdx = 0; wdx < kHeadThreads / 32; ++wdx) {
        l += red[wdx][0];
        cn += red[wdx][1];
      }
      a.ws[blockIdx.x * 2 + 0] = l;
      a.ws[blockIdx.x * 2 + 1] = cn;
    }
    return;
  }
  if (BWD) {
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float v = dw_acc[k][j];
#pragma unroll
        for (int o = LPP; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane < LPP) red[warp][(k * LPP + lane) * 8 + j] = v;
      }
      float d = db_acc[k];
#pragma unroll
      for (int o = LPP; o < 32; o <<= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
      if (lane == 0) red[warp][kRowD + k] = d;
    }
    __syncthreads();
    float* out = a.ws + (long long)blockIdx.x * (K * c + K);   // partial layout of head_kernel: [K*c + K]
    for (int q = threadIdx.x; q < kRowD; q += blockDim.x) {
      const int k = q / (LPP * 8), ch = q - k * (LPP * 8);
      if (k < K) {
        float sacc = 0.f;
        for (int wdx = 0; wdx < kHeadThreads / 32; ++wdx) sacc += red[wdx][q];
        out[k * c + ch] = sacc;
      }
    }
    if (threadIdx.x < K) {
      float sacc = 0.f;
      for (int wdx = 0; wdx < kHeadThreads / 32; ++wdx) sacc += red[wdx][kRowD + threadIdx.x];
      out[K * c + threadIdx.x] = sacc;
    }
  }
}

// ------------------------------------------------------------------ pixel-per-thread path (bf16 tier, 16 / 32 / 64 channels, <= 4 classes)
// ncu on head_dense_kernel above (profiles/r02_hbm_metrics.txt): 210-268 M warp instructions per launch, DRAM 28-45 % busy —
// the kernel is bound by INSTRUCTION ISSUE, not by memory: with 8 lanes per pixel every lane repeats the bias / ReLU /
// softmax / label logic of its pixel, each dot product ends in a 3-step shuffle tree, and zr[label] went through local
// memory.  Here a block stages 256 pixels (256 x C bf16, cp.async, 3-stage ring) in shared memory with the 16-byte chunks
// of a pixel XOR-swizzled by the pixel index, and works on them in two mappings:
//   pass 1  thread = pixel: reads its own row (conflict-free thanks to the swizzle), one dot product per class with packed
//           FFMA2 against broadcast weights, softmax / loss / dz ONCE per pixel; dz goes to shared memory;
//   pass 2  (backward) thread = (16-byte chunk j, pixel subset): its 8 channels' weights and dW accumulators stay in
//           registers; per pixel one LDS.128 of x and the pixel's dz give dW += dz * x and dx = mask(dz . w), and the 8 lanes
//           of a pixel store its 128 contiguous bytes of dx.
// ~200 (forward) / ~700 (backward) thread instructions per pixel instead of ~1400 / ~1800.
constexpr int kPixPerBlock = 256;

// KMAX = 8 (up to 8 classes): the per-thread weights of pass 2 no longer fit in registers next to the dW accumulators and
// are read from shared memory.  SPLIT (forward modes of the split precision tier): x = hi + lo, a second ring carries the
// lo plane and the pipeline is two stages deep so that both still fit.
template <int KMAX, int C, int MODE, bool SPLIT>
__global__ void __launch_bounds__(kHeadThreads, (KMAX > 4 || SPLIT) ? 1 : 2) head_pix_kernel(HeadArgs a, long long npix) {
  constexpr int kPixStages = SPLIT ? 2 : 3;
  constexpr bool W2REG = KMAX <= 4;           // pass 2: this lane's weights in registers (else from shared memory)
  constexpr int CH = C / 8;                   // 16-byte chunks per pixel
  constexpr int NS = kHeadThreads / CH;       // pass 2: pixel subsets
  constexpr int PPS = kPixPerBlock / NS;      // pass 2: pixels per subset (= CH)
  constexpr bool BWD = MODE == HEAD_BWD || MODE == HEAD_CE_BWD;
  constexpr bool CE = MODE == HEAD_CE_FWD || MODE == HEAD_CE_BWD;
  extern __shared__ __align__(16) uint8_t pix_raw[];
  uint4* ring = reinterpret_cast<uint4*>(pix_raw);                             // [stages][256 pixels][CH]
  uint4* ring_lo = ring + (SPLIT ? kPixStages * kPixPerBlock * CH : 0);        // split tier: the lo plane, same layout
  float* sw = reinterpret_cast<float*>(ring_lo + kPixStages * kPixPerBlock * CH);  // [KMAX][C], zero for k >= K
  float2* sdz = reinterpret_cast<float2*>(sw + KMAX * C);                       // backward: [KMAX][256] as (dz, dz) pairs
  __shared__ float redw[kHeadThreads / 32][2];
  const int K = a.k;
  const int t = threadIdx.x;
  for (int i = t; i < KMAX * C; i += kHeadThreads) sw[i] = i < K * C ? a.w[i] : 0.f;
  float bias[KMAX];
#pragma unroll
  for (int k = 0; k < KMAX; ++k) bias[k] = (k < K && a.b) ? a.b[k] : 0.f;
  float gs = 1.f;
  if (MODE == HEAD_CE_BWD) gs = (a.gscale ? a.gscale[0] : 1.f) * a.ce_state[1];
  const long long hw = (long long)a.x.h * a.x.w;
  const uint4* __restrict__ xg = reinterpret_cast<const uint4*>(a.x.p);  // dense: pixel p, chunk j at xg[p * CH + j]
  const uint4* __restrict__ xgl = SPLIT ? reinterpret_cast<const uint4*>(a.x.lo) : nullptr;

  // pass-2 role
  const int j2 = t % CH, s2 = t / CH;
  float2 w2[(BWD && W2REG) ? KMAX : 1][4], dw2[BWD ? KMAX : 1][4];
  float db_acc[KMAX];
  float loss_acc = 0.f, cnt_acc = 0.f;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) db_acc[k] = 0.f;
  if (BWD) {
#pragma unroll
    for (int k = 0; k < KMAX; ++k)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (W2REG)
          w2[k][e] = k < K ? make_float2(a.w[k * C + j2 * 8 + 2 * e], a.w[k * C + j2 * 8 + 2 * e + 1]) : make_float2(0.f, 0.f);
        dw2[k][e] = make_float2(0.f, 0.f);
      }
  }

  const long long nchunks = (npix + kPixPerBlock - 1) / kPixPerBlock;
  auto issue = [&](long long chunk, int stage) {
    if (chunk < nchunks) {
      const long long p0 = chunk * kPixPerBlock;
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        const int g = i * kHeadThreads + t;        // 16-byte piece of the block's 256 * CH, coalesced
        const int pl = g / CH, j = g - pl * CH;
        if (p0 + pl < npix) {
          cp_async_16(&ring[(stage * kPixPerBlock + pl) * CH + (j ^ (pl & (CH - 1)))], xg + (p0 + pl) * CH + j);
          if (SPLIT) cp_async_16(&ring_lo[(stage * kPixPerBlock + pl) * CH + (j ^ (pl & (CH - 1)))], xgl + (p0 + pl) * CH + j);
        }
      }
    }
    cp_async_commit();
  };
  long long chunk = blockIdx.x;
#pragma unroll
  for (int st = 0; st < kPixStages - 1; ++st) issue(chunk + (long long)st * gridDim.x, st);
  int stage = 0;
  for (; chunk < nchunks; chunk += gridDim.x) {
    issue(chunk + (long long)(kPixStages - 1) * gridDim.x, stage == 0 ? kPixStages - 1 : stage - 1);
    cp_async_wait<kPixStages - 1>();
    __syncthreads();  // every thread's pieces of this stage have landed (and sw on the first iteration)
    const long long p0 = chunk * kPixPerBlock;
    const uint4* rows = ring + stage * kPixPerBlock * CH;
    // ---------------- pass 1: thread = pixel
    {
      const long long p = p0 + t;
      const bool live = p < npix;
      float2 acc[KMAX];
#pragma unroll
      for (int k = 0; k < KMAX; ++k) acc[k] = make_float2(0.f, 0.f);
      if (live) {
#pragma unroll
        for (int j = 0; j < CH; ++j) {
          const uint4 v = rows[t * CH + (j ^ (t & (CH - 1)))];
          float2 x01 = bf2x_to_f2(v.x), x23 = bf2x_to_f2(v.y), x45 = bf2x_to_f2(v.z), x67 = bf2x_to_f2(v.w);
          if (SPLIT) {  // x = hi + lo (exact in fp32)
            const uint4 vl = ring_lo[(stage * kPixPerBlock + t) * CH + (j ^ (t & (CH - 1)))];
            const float2 l01 = bf2x_to_f2(vl.x), l23 = bf2x_to_f2(vl.y), l45 = bf2x_to_f2(vl.z), l67 = bf2x_to_f2(vl.w);
            x01.x += l01.x; x01.y += l01.y; x23.x += l23.x; x23.y += l23.y;
            x45.x += l45.x; x45.y += l45.y; x67.x += l67.x; x67.y += l67.y;
          }
#pragma unroll
          for (int k = 0; k < KMAX; ++k) {
            const float4 wa = *reinterpret_cast<const float4*>(sw + k * C + j * 8);
            const float4 wb = *reinterpret_cast<const float4*>(sw + k * C + j * 8 + 4);
            acc[k] = ffma2(x01, make_float2(wa.x, wa.y), acc[k]);
            acc[k] = ffma2(x23, make_float2(wa.z, wa.w), acc[k]);
            acc[k] = ffma2(x45, make_float2(wb.x, wb.y), acc[k]);
            acc[k] = ffma2(x67, make_float2(wb.z, wb.w), acc[k]);
          }
        }
      }
      float z[KMAX], zr[KMAX];
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        z[k] = acc[k].x + acc[k].y + bias[k];
        zr[k] = a.relu ? fmaxf(z[k], 0.f) : z[k];
      }
      long long n = 0, r = 0;
      if ((!BWD && a.logits) || MODE == HEAD_BWD) {
        n = p / hw;
        r = p - n * hw;
      }
      if (!BWD && a.logits && live) {
#pragma unroll
        for (int k = 0; k < KMAX; ++k)
          if (k < K) a.logits[(n * K + k) * hw + r] = zr[k];
      }
      if (MODE != HEAD_FWD) {
        int lab = -1;  // class, -1 = ignore_index, -2 = out of range
        bool valid = live;
        float lse = 0.f;
        if (CE) {
          if (live) {
            const long long l64 = a.labels[p];
            lab = l64 == kIgnoreIndex ? -1 : ((l64 < 0 || l64 >= K) ? -2 : (int)l64);
          }
          valid = live && lab != -1;
          float m = -INFINITY;
#pragma unroll
          for (int k = 0; k < KMAX; ++k)
            if (k < K) m = fmaxf(m, zr[k]);
          float sum = 0.f;
#pragma unroll
          for (int k = 0; k < KMAX; ++k)
            if (k < K) sum += __expf(zr[k] - m);
          lse = m + __logf(sum);
        }
        if (MODE == HEAD_CE_FWD) {
          if (valid) {
            float zy = 0.f;
#pragma unroll
            for (int k = 0; k < KMAX; ++k) zy = k == lab ? zr[k] : zy;
            if (lab == -2) zy = __int_as_float(0x7fc00000);  // out-of-range label: poison the loss
            loss_acc += lse - zy;
            cnt_acc += 1.f;
          }
        } else {
#pragma unroll
          for (int k = 0; k < KMAX; ++k) {
            float g = 0.f;
            if (k < K && valid) {
              if (MODE == HEAD_CE_BWD) g = gs * (__expf(zr[k] - lse) - (k == lab ? 1.f : 0.f));
              else g = a.dlogits[(n * K + k) * hw + r];
              if (a.relu && !(z[k] > 0.f)) g = 0.f;
            }
            sdz[k * kPixPerBlock + t] = make_float2(g, g);
          }
        }
      }
    }
    if (BWD) {
      __syncthreads();  // dz of the 256 pixels
      // ---------------- pass 2: thread = (chunk j2, subset s2); pixel pl = i * NS + s2 -> the 32 lanes of a warp cover
      // 32 / CH consecutive pixels x all chunks: one contiguous run of dx per store instruction
      uint4* dxg = reinterpret_cast<uint4*>(a.dx.p);
#pragma unroll
      for (int i = 0; i < PPS; ++i) {
        const int pl = i * NS + s2;
        if (p0 + pl < npix) {
          const uint4 v = rows[pl * CH + (j2 ^ (pl & (CH - 1)))];
          float2 x2[4] = {bf2x_to_f2(v.x), bf2x_to_f2(v.y), bf2x_to_f2(v.z), bf2x_to_f2(v.w)};
          float2 r2[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
          for (int k = 0; k < KMAX; ++k) {
            const float2 d = sdz[k * kPixPerBlock + pl];
            if (j2 == 0) db_acc[k] += d.x;
            float2 wk[4];
            if (W2REG) {
#pragma unroll
              for (int e = 0; e < 4; ++e) wk[e] = w2[k][e];
            } else {
              const float4 wa = *reinterpret_cast<const float4*>(sw + k * C + j2 * 8);
              const float4 wb = *reinterpret_cast<const float4*>(sw + k * C + j2 * 8 + 4);
              wk[0] = make_float2(wa.x, wa.y); wk[1] = make_float2(wa.z, wa.w);
              wk[2] = make_float2(wb.x, wb.y); wk[3] = make_float2(wb.z, wb.w);
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              dw2[k][e] = ffma2(d, x2[e], dw2[k][e]);
              r2[e] = ffma2(d, wk[e], r2[e]);
            }
          }
          if (dxg) {
            if (a.mask) {  // the mask IS x here (no BatchNorm): ReLU backward of the last convolution
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                r2[e].x = x2[e].x > 0.f ? r2[e].x : 0.f;
                r2[e].y = x2[e].y > 0.f ? r2[e].y : 0.f;
              }
            }
            dxg[(p0 + pl) * CH + j2] = make_uint4(f2_to_bf2x(r2[0].x, r2[0].y), f2_to_bf2x(r2[1].x, r2[1].y),
                                                  f2_to_bf2x(r2[2].x, r2[2].y), f2_to_bf2x(r2[3].x, r2[3].y));
          }
        }
      }
    }
    __syncthreads();  // this stage (and sdz) may be overwritten
    stage = stage + 1 == kPixStages ? 0 : stage + 1;
  }
  cp_async_wait<0>();
  if (MODE == HEAD_FWD) return;
  __syncthreads();
  const int warp = t >> 5, lane = t & 31;
  if (MODE == HEAD_CE_FWD) {
    loss_acc = warp_sum(loss_acc);
    cnt_acc = warp_sum(cnt_acc);
    if (lane == 0) {
      redw[warp][0] = loss_acc;
      redw[warp][1] = cnt_acc;
    }
    __syncthreads();
    if (t == 0) {
      float l = 0.f, cn = 0.f;
      for (int wdx = 0; wdx < kHeadThreads / 32; ++wdx) {
        l += redw[wdx][0];
        cn += redw[wdx][1];
      }
      a.ws[blockIdx.x * 2 + 0] = l;
      a.ws[blockIdx.x * 2 + 1] = cn;
    }
    return;
  }
  if (BWD) {
    // reduce over the NS pixel subsets through the (now idle) ring: red[s2][k][c], then [s2][K] for db
    float* red = reinterpret_cast<float*>(pix_raw);
#pragma unroll
    for (int k = 0; k < KMAX; ++k)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        red[(s2 * KMAX + k) * C + j2 * 8 + 2 * e] = dw2[k][e].x;
        red[(s2 * KMAX + k) * C + j2 * 8 + 2 * e + 1] = dw2[k][e].y;
      }
    float* redb = red + NS * KMAX * C;
    if (j2 == 0) {
#pragma unroll
      for (int k = 0; k < KMAX; ++k) redb[s2 * KMAX + k] = db_acc[k];
    }
    __syncthreads();
    float* out = a.ws + (long long)blockIdx.x * (K * C + K);
    for (int q = t; q < K * C; q += kHeadThreads) {
      float sacc = 0.f;
      for (int s = 0; s < NS; ++s) sacc += red[s * KMAX * C + q];  // q = k * C + c with k < K <= KMAX
      out[q] = sacc;
    }
    if (t < K) {
      float sacc = 0.f;
      for (int s = 0; s < NS; ++s) sacc += redb[s * KMAX + t];
      out[K * C + t] = sacc;
    }
  }
}

inline size_t head_pix_smem(int c, int kmax, bool bwd, bool split) {
  const size_t ring = (size_t)(split ? 2 * 2 : 3) * kPixPerBlock * c * 2;
  const size_t tail = (size_t)kmax * c * 4 + (bwd ? (size_t)kmax * kPixPerBlock * 8 : 0);
  const size_t red = bwd ? (size_t)(kHeadThreads / (c / 8)) * kmax * (c + 1) * 4 : 0;  // aliases the ring after the loop
  return (ring > red ? ring : red) + tail;
}

inline bool dense_nhwc(const DView& v) {
  return v.sw == v.c && v.sh == (long long)v.w * v.sw && v.sn == (long long)v.h * v.sh;
}

template <int MODE, int KMAX, int LPP>
void launch_head_dense_inst(const HeadArgs& a, int blocks, long long npix, cudaStream_t st) {
  constexpr bool FWD = MODE == HEAD_FWD || MODE == HEAD_CE_FWD;  // only the forward modes read the lo plane
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(head_dense_kernel<KMAX, LPP, MODE, kHeadDenseU, FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)head_ring_bytes(LPP, true));
    cudaFuncSetAttribute(head_dense_kernel<KMAX, LPP, MODE, kHeadDenseU, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)head_ring_bytes(LPP, false));
    attr_done = true;
  }
  if (FWD && a.x.lo)
    head_dense_kernel<KMAX, LPP, MODE, kHeadDenseU, FWD><<<blocks, kHeadThreads, head_ring_bytes(LPP, true), st>>>(a, npix);
  else
    head_dense_kernel<KMAX, LPP, MODE, kHeadDenseU, false><<<blocks, kHeadThreads, head_ring_bytes(LPP, false), st>>>(a, npix);
}

// returns true if the fast kernel took the launch
template <int MODE>
bool launch_head_dense(const HeadArgs& a, int* blocks_io, cudaStream_t st) {
  constexpr bool BWD = MODE == HEAD_BWD || MODE == HEAD_CE_BWD;
  const int c = a.x.c, K = a.k;
  if (!(c == 8 || c == 16 || c == 32 || c == 64) || !dense_nhwc(a.x)) return false;
  if (K > 4) return false;  // 8 classes x 8 channels of weights (+ as many dW accumulators) in registers: measured slower
  if (reinterpret_cast<uintptr_t>(a.x.p) % 16 || reinterpret_cast<uintptr_t>(a.x.lo) % 16) return false;
  if (BWD) {
    if (a.x.lo) return false;  // backward entry points read the hi plane only, and the general kernel handles it
    if (a.dx.p && (!dense_nhwc(a.dx) || reinterpret_cast<uintptr_t>(a.dx.p) % 16)) return false;
    if (a.mask && a.mask != a.x.p) return false;  // a separate mask tensor: general kernel
  }
  const long long npix = (long long)a.x.n * a.x.h * a.x.w;
  const int lpp = c / 8;
  const long long per_block = (kHeadThreads / lpp) * kHeadDenseU;
  long long blocks = (npix + per_block - 1) / per_block;
  if (blocks > kHeadMaxBlocks) blocks = kHeadMaxBlocks;
  if (blocks < 1) blocks = 1;
  *blocks_io = (int)blocks;
  const int kmax = K <= 2 ? 2 : 4;
#define B200_HD(KM, L) launch_head_dense_inst<MODE, KM, L>(a, (int)blocks, npix, st)
#define B200_HD_K(L)            \
  do {                          \
    if (kmax == 2) B200_HD(2, L);      \
    else B200_HD(4, L);                \
  } while (0)
  if (lpp == 8) B200_HD_K(8);
  else if (lpp == 4) B200_HD_K(4);
  else if (lpp == 2) B200_HD_K(2);
  else B200_HD_K(1);
#undef B200_HD_K
#undef B200_HD
  return true;
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

template <int MODE, int KMAX, int CC, bool SPLIT>
void launch_head_pix_inst(const HeadArgs& a, int blocks, long long npix, cudaStream_t st) {
  constexpr bool BWD = MODE == HEAD_BWD || MODE == HEAD_CE_BWD;
  const size_t smem = head_pix_smem(CC, KMAX, BWD, SPLIT);
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(head_pix_kernel<KMAX, CC, MODE, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_done = true;
  }
  head_pix_kernel<KMAX, CC, MODE, SPLIT><<<blocks, kHeadThreads, smem, st>>>(a, npix);
}

static int g_head_pix = getenv("B200UNET_NO_HEAD_PIX") ? 0 : 1;

// returns true if the pixel-per-thread kernel took the launch
template <int MODE>
bool launch_head_pix(const HeadArgs& a, int* blocks_io, cudaStream_t st) {
  constexpr bool BWD = MODE == HEAD_BWD || MODE == HEAD_CE_BWD;
  const int c = a.x.c, K = a.k;
  if (!g_head_pix || !(c == 16 || c == 32 || c == 64) || K > 8 || !dense_nhwc(a.x)) return false;
  if (reinterpret_cast<uintptr_t>(a.x.p) % 16 || reinterpret_cast<uintptr_t>(a.x.lo) % 16) return false;
  const bool split = !BWD && a.x.lo != nullptr;  // the backward entry points read the hi plane only
  if (BWD) {
    if (a.dx.p && (!dense_nhwc(a.dx) || reinterpret_cast<uintptr_t>(a.dx.p) % 16)) return false;
    if (a.mask && a.mask != a.x.p) return false;  // a separate mask tensor: general kernel
  }
  const long long npix = (long long)a.x.n * a.x.h * a.x.w;
  long long blocks = (npix + kPixPerBlock - 1) / kPixPerBlock;
  const int per_sm = (K > 4 || split) ? 1 : 2;
  if (blocks > per_sm * kNumSMsB200) blocks = per_sm * kNumSMsB200;  // persistent
  if (blocks < 1) blocks = 1;
  *blocks_io = (int)blocks;
  const int kmax = K <= 2 ? 2 : (K <= 4 ? 4 : 8);
#define B200_HP(KM, CC)                                                        \
  do {                                                                         \
    if (split) launch_head_pix_inst<MODE, KM, CC, !BWD>(a, (int)blocks, npix, st); \
    else launch_head_pix_inst<MODE, KM, CC, false>(a, (int)blocks, npix, st);  \
  } while (0)
#define B200_HP_C(CC)                         \
  do {                                        \
    if (kmax == 2) B200_HP(2, CC);            \
    else if (kmax == 4) B200_HP(4, CC);       \
    else B200_HP(8, CC);                      \
  } while (0)
  if (c == 64) B200_HP_C(64);
  else if (c == 32) B200_HP_C(32);
  else B200_HP_C(16);
#undef B200_HP_C
#undef B200_HP
  return true;
}

template <int MODE>
int launch_head(const HeadArgs& a, int* blocks_io, cudaStream_t st) {
  if (launch_head_pix<MODE>(a, blocks_io, st)) return check_launch("head (pixel-per-thread)");
  if (launch_head_dense<MODE>(a, blocks_io, st)) return check_launch("head (dense)");
  const int blocks = *blocks_io;
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
  int blocks = head_blocks(view_pixels(*x));
  return launch_head<HEAD_FWD>(a, &blocks, as_stream(stream));
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
  int blocks = head_blocks(view_pixels(*x));
  int r = launch_head<HEAD_BWD>(a, &blocks, as_stream(stream));
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
  int blocks = head_blocks(view_pixels(*x));
  int r = launch_head<HEAD_CE_FWD>(a, &blocks, as_stream(stream));
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
  int blocks = head_blocks(view_pixels(*x));
  int r = launch_head<HEAD_CE_BWD>(a, &blocks, as_stream(stream));
  if (r) return r;
  const int kc = n_classes * x->c;
  head_bwd_finalize_kernel<<<finalize_grid(kc + n_classes), kFinalizeThreads, 0, as_stream(stream)>>>((const float*)workspace, blocks,
                                                                                        kc, n_classes, dw, db);
  return check_launch("head_ce_bwd finalize");
}
}
