// First-layer kernels: 3x3 convolution whose INPUT has 1..4 channels (the image: unet.py:49-52 with in_channels 1 / 3).
// K = 9*Cin <= 36 is far below one tensor-core K step worth feeding through TMA (2..8 bytes per pixel), and the layer
// is bound by writing (forward) or reading (backward-weights) the 64-channel full-resolution tensor: 128 B / pixel
// against 576*Cin FMAs / pixel.  So: CUDA cores, fp32 weights in shared memory, and a thread mapping in which the
// 4 lanes that share a pixel group cover its 64 output channels -> every global access is a full 32-byte sector and
// a warp touches whole 128-byte lines.
//   forward : thread = 4 consecutive pixels x 16 output channels (64 accumulators, 16 FMAs per LDS.128 of weights)
//   wgrad   : thread = CPT output channels x all 9*Cin taps, grid-strided over pixels; warp-shuffle + shared-memory
//             block reduction, per-block partials, fixed-order final reduction (reproducible); bias gradient for free
#include "conv_impl.h"

namespace b200 {

constexpr int kScThreads = 256;

template <int CIN>
__global__ void __launch_bounds__(kScThreads, 2)
smallc_fwd_kernel(DView src, DView dst, const float* __restrict__ w, const float* __restrict__ bias, int relu, int pad) {
  __shared__ __align__(16) float ws[9 * CIN * 64];
  __shared__ float bs[64];
  const int cout = dst.c;
  const int o_base = blockIdx.y * 64;
  for (int i = threadIdx.x; i < 9 * CIN * 64; i += kScThreads) {
    const int o = i & 63, tc = i >> 6;  // tc = tap * CIN + c
    const int tap = tc / CIN, c = tc - tap * CIN;
    ws[i] = (o_base + o < cout) ? w[((long long)(o_base + o) * CIN + c) * 9 + tap] : 0.f;
  }
  if (threadIdx.x < 64) bs[threadIdx.x] = (bias && o_base + threadIdx.x < cout) ? bias[o_base + threadIdx.x] : 0.f;
  __syncthreads();

  const int q4 = threadIdx.x & 3;
  const int groups_per_row = (dst.w + 3) >> 2;
  const long long total_groups = (long long)dst.n * dst.h * groups_per_row;
  if (o_base + q4 * 8 >= cout) return;
  const bool second_half = o_base + 32 + q4 * 8 < cout;  // cout is a multiple of 16, not necessarily of 64
  // grid-stride over pixel groups: the weights are staged once per block (one block per 256 pixels spent more time in
  // its prologue and in block scheduling than in the 576 FMAs per thread)
  for (long long pg = (long long)blockIdx.x * (kScThreads / 4) + (threadIdx.x >> 2); pg < total_groups;
       pg += (long long)gridDim.x * (kScThreads / 4)) {
  asm volatile("" ::: "memory");  // keep the weight LDS inside the iteration (hoisted, 9*CIN*16 values spill)
  const int xg = (int)(pg % groups_per_row);
  const long long t2 = pg / groups_per_row;
  const int oy = (int)(t2 % dst.h), n = (int)(t2 / dst.h);
  const int ox0 = xg * 4;

  float2 acc[4][8];  // 4 pixels x 16 channels as fp32 pairs: FFMA2 does two channels per instruction
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[p][k] = make_float2(0.f, 0.f);

#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int iy = oy + r - pad;
    const bool yok = iy >= 0 && iy < src.h;
#pragma unroll
    for (int c = 0; c < CIN; ++c) {
      float xin[6];
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        const int ix = ox0 + j - pad;
        float v = 0.f;
        if (yok && ix >= 0 && ix < src.w) {
          const long long so = src.off(n, iy, ix) + c;
          v = bf2f(src.p[so]);
          if (src.lo) v += bf2f(src.lo[so]);  // split tier: the image is carried as hi + lo
        }
        xin[j] = v;
      }
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        // lane q4 owns channels [8 q4, 8 q4 + 8) and [32 + 8 q4, 32 + 8 q4 + 8): the four lanes of a pixel then write two
        // contiguous 64-byte runs (full 32-byte sectors per store instruction; 16 channels in a row per lane left gaps)
        const float* wp = ws + ((r * 3 + s) * CIN + c) * 64 + q4 * 8;
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const float4 wv = *reinterpret_cast<const float4*>(wp + (j4 >> 1) * 32 + (j4 & 1) * 4);
          const float2 w01 = make_float2(wv.x, wv.y), w23 = make_float2(wv.z, wv.w);
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            const float2 xx = make_float2(xin[p + s], xin[p + s]);
            acc[p][j4 * 2 + 0] = ffma2(xx, w01, acc[p][j4 * 2 + 0]);
            acc[p][j4 * 2 + 1] = ffma2(xx, w23, acc[p][j4 * 2 + 1]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const int ox = ox0 + p;
    if (ox >= dst.w) break;
    float f[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      f[k] = ((k & 1) ? acc[p][k >> 1].y : acc[p][k >> 1].x) + bs[(k >> 3) * 32 + q4 * 8 + (k & 7)];
      if (relu) f[k] = fmaxf(f[k], 0.f);
    }
    const long long oo = dst.off(n, oy, ox) + o_base + q4 * 8;
    float f0[8], f1[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      f0[k] = f[k];
      f1[k] = f[8 + k];
    }
    store8s(dst.p, dst.lo, oo, f0);
    if (second_half) store8s(dst.p, dst.lo, oo + 32, f1);
  }
  }
}

// partial layout per block: [64 couts][9*CIN] weights followed by [64] bias sums.
// Thread = CPT output channels x all 9*CIN taps; one loop iteration = a strip of 4 consecutive pixels of one row, whose
// 3 x 6 input window is loaded once per input channel (sliding window) and whose index arithmetic is paid once.
template <int CIN, int CPT>
__global__ void __launch_bounds__(kScThreads)
smallc_wgrad_kernel(DView dz, DView src, int pad, float* __restrict__ partial) {
  constexpr int LPP = 64 / CPT;       // lanes per pixel strip
  constexpr int NT = 9 * CIN;         // taps x input channels
  constexpr int ROW = 64 * NT + 64;   // floats per block partial
  extern __shared__ float red[];      // [8 warps][ROW]
  const int cout = dz.c;
  const int o_base = blockIdx.y * 64;
  const int cg = threadIdx.x % LPP;
  const int slot = threadIdx.x / LPP;
  constexpr int slots = kScThreads / LPP;
  const int o0 = o_base + cg * CPT;
  float acc[CPT][NT];
  float accb[CPT];
#pragma unroll
  for (int i = 0; i < CPT; ++i) {
    accb[i] = 0.f;
#pragma unroll
    for (int j = 0; j < NT; ++j) acc[i][j] = 0.f;
  }
  const int groups_per_row = (dz.w + 3) >> 2;
  const unsigned total_groups = (unsigned)dz.n * dz.h * groups_per_row;
  if (o0 < cout) {
    for (unsigned g = blockIdx.x * slots + slot; g < total_groups; g += gridDim.x * slots) {
      const unsigned xg = g % groups_per_row;
      const unsigned t2 = g / groups_per_row;
      const int oy = (int)(t2 % dz.h), n = (int)(t2 / dz.h);
      const int ox0 = (int)xg * 4;
      float z[4][CPT];
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        if (ox0 + p < dz.w) {
          const bf16* zp = dz.p + dz.off(n, oy, ox0 + p) + o0;
          if (CPT == 16) {
            float t[8];
            unpack8(*reinterpret_cast<const bf16x8*>(zp), t);
#pragma unroll
            for (int i = 0; i < 8; ++i) z[p][i] = t[i];
            unpack8(*reinterpret_cast<const bf16x8*>(zp + 8), t);
#pragma unroll
            for (int i = 0; i < 8; ++i) z[p][(8 + i) % CPT] = t[i];
          } else if (CPT == 8) {
            float t[8];
            unpack8(*reinterpret_cast<const bf16x8*>(zp), t);
#pragma unroll
            for (int i = 0; i < CPT; ++i) z[p][i] = t[i % 8];
          } else {
            const uint2 u = *reinterpret_cast<const uint2*>(zp);
            const float2 a2 = bf2x_to_f2(u.x), b2 = bf2x_to_f2(u.y);
            z[p][0] = a2.x;
            z[p][1 % CPT] = a2.y;
            z[p][2 % CPT] = b2.x;
            z[p][3 % CPT] = b2.y;
          }
        } else {
#pragma unroll
          for (int i = 0; i < CPT; ++i) z[p][i] = 0.f;
        }
#pragma unroll
        for (int i = 0; i < CPT; ++i) accb[i] += z[p][i];
      }
#pragma unroll
      for (int c = 0; c < CIN; ++c) {
        float xin[3][6];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const int iy = oy + r - pad;
          const bool yok = iy >= 0 && iy < src.h;
#pragma unroll
          for (int j = 0; j < 6; ++j) {
            const int ix = ox0 + j - pad;
            xin[r][j] = (yok && ix >= 0 && ix < src.w) ? bf2f(src.p[src.off(n, iy, ix) + c]) : 0.f;
          }
        }
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int sx = 0; sx < 3; ++sx)
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
              for (int i = 0; i < CPT; ++i) acc[i][c * 9 + r * 3 + sx] += z[p][i] * xin[r][p + sx];
      }
    }
  }
  // reduce over the pixel slots: first inside the warp (lanes with equal cg), then across the 8 warps
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < CPT; ++i) {
#pragma unroll
    for (int off = LPP; off < 32; off <<= 1) accb[i] += __shfl_xor_sync(0xffffffffu, accb[i], off);
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      float v = acc[i][j];
#pragma unroll
      for (int off = LPP; off < 32; off <<= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
      acc[i][j] = v;
    }
  }
  if (lane < LPP) {
    float* rw = red + warp * ROW;
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
      rw[64 * NT + lane * CPT + i] = accb[i];
#pragma unroll
      for (int j = 0; j < NT; ++j) rw[(lane * CPT + i) * NT + j] = acc[i][j];
    }
  }
  __syncthreads();
  float* out = partial + ((long long)blockIdx.y * gridDim.x + blockIdx.x) * ROW;
  for (int q = threadIdx.x; q < ROW; q += kScThreads) {
    float s = 0.f;
#pragma unroll
    for (int wdx = 0; wdx < kScThreads / 32; ++wdx) s += red[wdx * ROW + q];
    out[q] = s;
  }
}

// dw[o][c][tap] = sum_b partial[yblk][b][(o%64)*NT + c*9 + tap];  db[o] likewise
__global__ void smallc_wgrad_reduce_kernel(const float* __restrict__ partial, int blocks, int nt, int cout,
                                           float* __restrict__ dw, float* __restrict__ db) {
  const int row = 64 * nt + 64;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= cout * nt + cout) return;
  int yb, q;
  if (e < cout * nt) {
    const int o = e / nt;
    yb = o >> 6;
    q = (o & 63) * nt + (e - o * nt);
  } else {
    const int o = e - cout * nt;
    yb = o >> 6;
    q = 64 * nt + (o & 63);
  }
  double s = 0.0;
  for (int b = 0; b < blocks; ++b) s += (double)partial[((long long)yb * blocks + b) * row + q];
  if (e < cout * nt)
    dw[e] = (float)s;
  else if (db)
    db[e - cout * nt] = (float)s;
}

constexpr int kScWgradBlocks = 2 * kNumSMsB200;

static bool sc_aligned_out(const b200_view& v) {
  return v.c % 16 == 0 && reinterpret_cast<uintptr_t>(v.ptr) % 16 == 0 && reinterpret_cast<uintptr_t>(v.lo) % 16 == 0 &&
         v.stride_w % 8 == 0 && v.stride_h % 8 == 0 && (v.n == 1 || v.stride_n % 8 == 0);
}

bool smallc_conv_fwd_ok(const b200_conv_fwd_params* p) {
  return p->num_src == 1 && p->taps == 9 && p->src[0].c >= 1 && p->src[0].c <= 4 && sc_aligned_out(p->dst) && p->w_f32;
}

int smallc_conv_fwd(const b200_conv_fwd_params* p, cudaStream_t st) {
  const long long groups = (long long)p->dst.n * p->dst.h * ((p->dst.w + 3) / 4);
  long long gx = (groups + kScThreads / 4 - 1) / (kScThreads / 4);
  if (gx > 8 * kNumSMsB200) gx = 8 * kNumSMsB200;
  dim3 grid((unsigned)gx, (unsigned)((p->dst.c + 63) / 64));
  const DView s = dview(p->src[0]), d = dview(p->dst);
  switch (p->src[0].c) {
    case 1: smallc_fwd_kernel<1><<<grid, kScThreads, 0, st>>>(s, d, p->w_f32, p->bias, p->relu, p->pad); break;
    case 2: smallc_fwd_kernel<2><<<grid, kScThreads, 0, st>>>(s, d, p->w_f32, p->bias, p->relu, p->pad); break;
    case 3: smallc_fwd_kernel<3><<<grid, kScThreads, 0, st>>>(s, d, p->w_f32, p->bias, p->relu, p->pad); break;
    default: smallc_fwd_kernel<4><<<grid, kScThreads, 0, st>>>(s, d, p->w_f32, p->bias, p->relu, p->pad); break;
  }
  return check_launch("smallc_fwd");
}

bool smallc_conv_wgrad_ok(const b200_conv_wgrad_params* p) {
  return p->num_src == 1 && p->taps == 9 && p->src[0].c >= 1 && p->src[0].c <= 4 && sc_aligned_out(p->dz);
}

size_t smallc_conv_wgrad_workspace(const b200_conv_wgrad_params* p) {
  const int nt = 9 * p->src[0].c;
  return (size_t)((p->dz.c + 63) / 64) * kScWgradBlocks * (64 * nt + 64) * sizeof(float);
}

template <int CIN, int CPT>
static int sc_wgrad_launch(const b200_conv_wgrad_params* p, float* ws, int blocks, cudaStream_t st) {
  constexpr int ROW = 64 * 9 * CIN + 64;
  const size_t smem = (size_t)(kScThreads / 32) * ROW * sizeof(float);
  auto kern = smallc_wgrad_kernel<CIN, CPT>;
  static bool attr_done = false;
  if (!attr_done && smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail((int)e, "smallc_wgrad: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_done = true;
  }
  dim3 grid(blocks, (unsigned)((p->dz.c + 63) / 64));
  kern<<<grid, kScThreads, smem, st>>>(dview(p->dz), dview(p->src[0]), p->pad, ws);
  return check_launch("smallc_wgrad");
}

int smallc_conv_wgrad(const b200_conv_wgrad_params* p, void* ws, size_t ws_bytes, cudaStream_t st) {
  const size_t need = smallc_conv_wgrad_workspace(p);
  if (!ws || ws_bytes < need) return fail(-1, "conv_wgrad (first layer): workspace too small (%zu < %zu)", ws_bytes, need);
  const long long npix = view_pixels(p->dz);
  int blocks = (int)((npix + 255) / 256);
  if (blocks > kScWgradBlocks) blocks = kScWgradBlocks;
  if (blocks < 1) blocks = 1;
  int r;
  switch (p->src[0].c) {
    case 1: r = sc_wgrad_launch<1, 16>(p, (float*)ws, blocks, st); break;
    case 2: r = sc_wgrad_launch<2, 8>(p, (float*)ws, blocks, st); break;
    case 3: r = sc_wgrad_launch<3, 4>(p, (float*)ws, blocks, st); break;
    default: r = sc_wgrad_launch<4, 4>(p, (float*)ws, blocks, st); break;
  }
  if (r) return r;
  const int nt = 9 * p->src[0].c;
  const int outs = p->dz.c * nt + p->dz.c;
  smallc_wgrad_reduce_kernel<<<(outs + 255) / 256, 256, 0, st>>>((const float*)ws, blocks, nt, p->dz.c, p->dw_f32, p->db_f32);
  return check_launch("smallc_wgrad_reduce");
}

}  // namespace b200
