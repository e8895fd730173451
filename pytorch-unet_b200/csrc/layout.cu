// Boundary transforms: NCHW fp32 <-> NHWC bf16, weight packing into GEMM operand layouts, ReLU masking,
// per-channel sums (bias gradients).
#include "chan_reduce.cuh"

namespace b200 {

// ------------------------------------------------------------------ NCHW fp32 -> NHWC bf16 (unet.py:73 input)
// One thread per pixel; reads are coalesced along w for every channel plane, writes are C contiguous bf16.
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, DView dst) {
  const long long hw = (long long)dst.h * dst.w;
  const long long total = hw * dst.n;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < total; p += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(p / hw);
    const long long r = p - n * hw;
    const int ih = (int)(r / dst.w), iw = (int)(r - (long long)ih * dst.w);
    bf16* o = dst.p + dst.off(n, ih, iw);
    const float* s = src + (long long)n * dst.c * hw + r;
    for (int c = 0; c < dst.c; ++c) {
      const float v = s[c * hw];
      const bf16 h = f2bf(v);
      o[c] = h;
      if (dst.lo) dst.lo[dst.off(n, ih, iw) + c] = f2bf(split_lo(v, bf2f(h)));
    }
  }
}

// Tiled transpose for wide tensors: block handles 32 pixels x 32 channels through shared memory.
__global__ void nhwc_to_nchw_kernel(DView src, float* __restrict__ dst) {
  __shared__ float tile[32][33];
  const long long hw = (long long)src.h * src.w;
  const long long total = hw * src.n;
  const long long p0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  {
    const long long p = p0 + threadIdx.y;
    const int c = c0 + threadIdx.x;
    float v = 0.f;
    if (p < total && c < src.c) {
      const int n = (int)(p / hw);
      const long long r = p - n * hw;
      const int ih = (int)(r / src.w), iw = (int)(r - (long long)ih * src.w);
      v = bf2f(src.p[src.off(n, ih, iw) + c]);
    }
    tile[threadIdx.y][threadIdx.x] = v;
  }
  __syncthreads();
  {
    const long long p = p0 + threadIdx.x;
    const int c = c0 + threadIdx.y;
    if (p < total && c < src.c) {
      const int n = (int)(p / hw);
      const long long r = p - n * hw;
      dst[((long long)n * src.c + c) * hw + r] = tile[threadIdx.x][threadIdx.y];
    }
  }
}

// ------------------------------------------------------------------ weight packing
// conv fprop: out[o][tap][kpad]; source i occupies columns [koff_i, koff_i + src_c[i]) of kpad, zero elsewhere.
// split = 1: the K axis is three copies of that layout, {hi(w) | hi(w) | lo(w)} (split precision tier, b200unet.h).
__global__ void pack_conv_fprop_kernel(const float* __restrict__ w, bf16* __restrict__ out, int cout, int cin_total,
                                       int taps, int kpad, int c0, int c0pad, int split) {
  const int kall = split ? 3 * kpad : kpad;
  const long long total = (long long)cout * taps * kall;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int kk = (int)(e % kall);
    const int seg = kk / kpad, k = kk - seg * kpad;
    const int tap = (int)((e / kall) % taps);
    const int o = (int)(e / ((long long)kall * taps));
    int c = -1;
    if (k < c0pad) {
      if (k < c0) c = k;
    } else {
      const int k1 = k - c0pad;
      if (k1 < cin_total - c0) c = c0 + k1;
    }
    const float v = c >= 0 ? w[((long long)o * cin_total + c) * taps + tap] : 0.f;
    const bf16 h = f2bf(v);
    out[e] = seg < 2 ? h : f2bf(split_lo(v, bf2f(h)));
  }
}

// conv dgrad: out[c][tap'][opad] = w[o][c][taps-1-tap'] (spatial flip), zero for o >= cout.
__global__ void pack_conv_dgrad_kernel(const float* __restrict__ w, bf16* __restrict__ out, int cout, int cin_total,
                                       int taps, int opad) {
  const long long total = (long long)cin_total * taps * opad;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int o = (int)(e % opad);
    const int tap = (int)((e / opad) % taps);
    const int c = (int)(e / ((long long)opad * taps));
    out[e] = f2bf(o < cout ? w[((long long)o * cin_total + c) * taps + (taps - 1 - tap)] : 0.f);
  }
}

// convT fwd: out[(ab*cout + o)][cpad] = w[c][o][ab];  convT dgrad: out[c][ab][opad] = w[c][o][ab]
__global__ void pack_convt_kernel(const float* __restrict__ w, bf16* __restrict__ out, int cin, int cout, int mode,
                                  int kp) {
  if (mode == 0 || mode == 2) {
    const int kall = mode == 2 ? 3 * kp : kp;  // mode 2: {hi | hi | lo} along K (split precision tier)
    const long long total = (long long)4 * cout * kall;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
      const int kk = (int)(e % kall);
      const int seg = kk / kp, c = kk - seg * kp;
      const int row = (int)(e / kall);
      const int ab = row / cout, o = row - ab * cout;
      const float v = c < cin ? w[((long long)c * cout + o) * 4 + ab] : 0.f;
      const bf16 h = f2bf(v);
      out[e] = seg < 2 ? h : f2bf(split_lo(v, bf2f(h)));
    }
  } else {
    const long long total = (long long)cin * 4 * kp;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
      const int o = (int)(e % kp);
      const int ab = (int)((e / kp) % 4);
      const int c = (int)(e / ((long long)kp * 4));
      out[e] = f2bf(o < cout ? w[((long long)c * cout + o) * 4 + ab] : 0.f);
    }
  }
}

// ------------------------------------------------------------------ first-layer im2col
// dst[n, y, x, c*9 + r*3 + s] = src[n, y + r - pad, x + s - pad, c] (zero outside / for the padding channels).
// With 1..7 input channels the 3x3 patch (9..63 values) fits one 64-wide K chunk, so the first convolution and its
// backward-weights become 1x1 problems for the tensor-core kernels; the patch tensor costs 2*Kp bytes per pixel.
__global__ void im2col3x3_kernel(DView src, DView dst, int pad) {
  const long long total = (long long)dst.n * dst.h * dst.w;
  const int kp = dst.c, cin = src.c;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < total; p += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(p % dst.w);
    const long long t = p / dst.w;
    const int oy = (int)(t % dst.h), n = (int)(t / dst.h);
    bf16* o = dst.p + dst.off(n, oy, ox);
    for (int j0 = 0; j0 < kp; j0 += 8) {
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int q = j0 + j;
        float v = 0.f;
        if (q < 9 * cin) {
          const int c = q / 9, tap = q - c * 9;
          const int iy = oy + tap / 3 - pad, ix = ox + tap % 3 - pad;
          if (iy >= 0 && iy < src.h && ix >= 0 && ix < src.w) v = bf2f(src.p[src.off(n, iy, ix) + c]);
        }
        f[j] = v;
      }
      *reinterpret_cast<bf16x8*>(o + j0) = pack8(f);
    }
  }
}

// One input channel, 16-wide patches (the paper network: in_channels = 1): the generic kernel above spends ~25 instructions
// per patch VALUE on index arithmetic (two divisions by 9 / 3 per value, a 64-bit offset and four bounds tests) and ran at
// 1.5 TB/s; here the three row pointers are formed once per pixel, the nine loads have constant offsets, and the two
// 16-byte stores of a pixel are adjacent.
__global__ void __launch_bounds__(256) im2col3x3_c1_kernel(DView src, DView dst, int pad) {
  const long long total = (long long)dst.n * dst.h * dst.w;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < total; p += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(p % dst.w);
    const long long t = p / dst.w;
    const int oy = (int)(t % dst.h), n = (int)(t / dst.h);
    const int iy0 = oy - pad, ix0 = ox - pad;
    uint32_t v[9];
    if (iy0 >= 0 && iy0 + 2 < src.h && ix0 >= 0 && ix0 + 2 < src.w) {
      const unsigned short* r0 = reinterpret_cast<const unsigned short*>(src.p) + src.off(n, iy0, ix0);
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int s2 = 0; s2 < 3; ++s2) v[r * 3 + s2] = r0[r * src.sh + s2 * src.sw];
    } else {
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int s2 = 0; s2 < 3; ++s2) {
          const int iy = iy0 + r, ix = ix0 + s2;
          v[r * 3 + s2] = (iy >= 0 && iy < src.h && ix >= 0 && ix < src.w)
                              ? reinterpret_cast<const unsigned short*>(src.p)[src.off(n, iy, ix)] : 0u;
        }
    }
    uint4* o = reinterpret_cast<uint4*>(dst.p + dst.off(n, oy, ox));
    o[0] = make_uint4(v[0] | (v[1] << 16), v[2] | (v[3] << 16), v[4] | (v[5] << 16), v[6] | (v[7] << 16));
    o[1] = make_uint4(v[8], 0u, 0u, 0u);
  }
}

// ------------------------------------------------------------------ uint8 HWC images -> the first convolution's operand
// Input pipeline at the boundary.  The reference divides the uint8 image by 255 on the host
// (dataloader.py:258-264), transposes it to CHW and uploads it as fp32 per sample (dataloader.py:553-579): 4 bytes per
// value over PCIe, then the module converts fp32 NCHW -> bf16 NHWC.  Here the batch travels as uint8 NHWC (1 byte per
// value, the layout image decoders produce) and ONE kernel writes dst[n,y,x,c] = src * scale[c] + shift[c] in the NHWC
// bf16 layout (hi + lo planes in the split tier) that down_path.0 reads; channels >= src_c of dst stay zero.
// scale = 1/255, shift = 0 reproduces the reference: the division is IEEE-rounded like fp32(u8 / 255.0).
__global__ void u8_nhwc_to_bf16_kernel(const uint8_t* __restrict__ src, int src_c, DView dst, const float* __restrict__ scale,
                                       const float* __restrict__ shift, int divide_255) {
  const long long total = (long long)dst.n * dst.h * dst.w;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < total; p += (long long)gridDim.x * blockDim.x) {
    const int ix = (int)(p % dst.w);
    const long long t = p / dst.w;
    const int iy = (int)(t % dst.h), n = (int)(t / dst.h);
    const long long o = dst.off(n, iy, ix);
    const uint8_t* s = src + p * src_c;
    for (int c = 0; c < dst.c; ++c) {
      float v = 0.f;
      if (c < src_c) {
        v = (float)s[c];
        v = divide_255 ? __fdiv_rn(v, 255.f) : v;
        if (scale) v *= scale[c];
        if (shift) v += shift[c];
      }
      const bf16 h = f2bf(v);
      dst.p[o + c] = h;
      if (dst.lo) dst.lo[o + c] = f2bf(split_lo(v, bf2f(h)));
    }
  }
}

// ------------------------------------------------------------------ y = mask > 0 ? x : 0
__global__ void relu_mask_kernel(DView x, DView m, DView y) {
  const long long total = (long long)x.n * x.h * x.w * x.c;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e % x.c);
    long long p = e / x.c;
    const int iw = (int)(p % x.w);
    p /= x.w;
    const int ih = (int)(p % x.h);
    const int n = (int)(p / x.h);
    const float mv = bf2f(m.p[m.off(n, ih, iw) + c]);
    y.p[y.off(n, ih, iw) + c] = mv > 0.f ? x.p[x.off(n, ih, iw) + c] : f2bf(0.f);
  }
}

struct SumF {
  template <int VEC>
  __device__ void operator()(const float (&a)[VEC], const float (&)[VEC], float (&acc)[1][VEC], int) const {
#pragma unroll
    for (int j = 0; j < VEC; ++j) acc[0][j] += a[j];
  }
};

__global__ void finalize_sum_kernel(const float* __restrict__ partial, int blocks, int c, float* __restrict__ out) {
  const int ch = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (ch >= c) return;
  const double s = warp_partial_sum(partial, blocks, c, ch);
  if ((threadIdx.x & 31) == 0) out[ch] = (float)s;
}

inline int grid_for(long long total, int threads) {
  long long g = (total + threads - 1) / threads;
  const long long cap = (long long)kNumSMsB200 * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200unet_nchw_f32_to_nhwc_bf16(const float* src, const b200_view* dst, void* stream) {
  B200_REQUIRE(src && view_ok(dst), "nchw_to_nhwc: bad arguments");
  nchw_to_nhwc_kernel<<<grid_for(view_pixels(*dst), 256), 256, 0, as_stream(stream)>>>(src, dview(*dst));
  return check_launch("nchw_to_nhwc");
}

int b200unet_nhwc_bf16_to_nchw_f32(const b200_view* src, float* dst, void* stream) {
  B200_REQUIRE(dst && view_ok(src), "nhwc_to_nchw: bad arguments");
  const long long total = view_pixels(*src);
  dim3 grid((unsigned)((total + 31) / 32), (unsigned)((src->c + 31) / 32));
  nhwc_to_nchw_kernel<<<grid, dim3(32, 32), 0, as_stream(stream)>>>(dview(*src), dst);
  return check_launch("nhwc_to_nchw");
}

int b200unet_u8_nhwc_to_bf16(const uint8_t* src, int src_c, const b200_view* dst, const float* scale, const float* shift,
                             int divide_255, void* stream) {
  B200_REQUIRE(src && view_ok(dst) && src_c >= 1 && src_c <= dst->c, "u8_nhwc_to_bf16: bad arguments");
  u8_nhwc_to_bf16_kernel<<<grid_for(view_pixels(*dst), 256), 256, 0, as_stream(stream)>>>(src, src_c, dview(*dst), scale, shift,
                                                                                      divide_255);
  return check_launch("u8_nhwc_to_bf16");
}

int b200unet_im2col3x3(const b200_view* src, const b200_view* dst, int pad, void* stream) {
  B200_REQUIRE(view_ok(src) && view_ok(dst), "im2col3x3: bad views");
  B200_REQUIRE(dst->n == src->n && dst->h == src->h + 2 * pad - 2 && dst->w == src->w + 2 * pad - 2,
               "im2col3x3: dst extent must be the 3x3 convolution's output extent");
  B200_REQUIRE(dst->c % 8 == 0 && dst->c >= 9 * src->c && vec8_ok(*dst), "im2col3x3: dst needs >= 9*cin channels, multiple of 8");
  if (src->c == 1 && dst->c == 16 && !src->lo)
    im2col3x3_c1_kernel<<<grid_for(view_pixels(*dst), 256), 256, 0, as_stream(stream)>>>(dview(*src), dview(*dst), pad);
  else
    im2col3x3_kernel<<<grid_for(view_pixels(*dst), 256), 256, 0, as_stream(stream)>>>(dview(*src), dview(*dst), pad);
  return check_launch("im2col3x3");
}

static int conv_kpad(int num_src, const int* src_c) {
  int k = 0;
  for (int i = 0; i < num_src; ++i) k += (src_c[i] + 63) / 64 * 64;
  return k;
}

size_t b200unet_pack_conv_weight_bytes(int cout, int num_src, const int* src_c, int taps, int mode) {
  if (!src_c || num_src < 1 || num_src > 2) return 0;
  int cin = 0;
  for (int i = 0; i < num_src; ++i) cin += src_c[i];
  if (mode == 0 || mode == 2) return (size_t)cout * taps * conv_kpad(num_src, src_c) * 2 * (mode == 2 ? 3 : 1);
  return (size_t)cin * taps * ((cout + 63) / 64 * 64) * 2;
}

int b200unet_pack_conv_weight(const float* w, int cout, int num_src, const int* src_c, int taps, int mode, void* out,
                              void* stream) {
  B200_REQUIRE(w && out && src_c && num_src >= 1 && num_src <= 2 && (taps == 1 || taps == 9) && cout > 0 && mode >= 0 &&
                   mode <= 2,
               "pack_conv_weight: bad arguments");
  int cin = 0;
  for (int i = 0; i < num_src; ++i) cin += src_c[i];
  if (mode == 0 || mode == 2) {
    const int kpad = conv_kpad(num_src, src_c);
    const int c0 = src_c[0], c0pad = (c0 + 63) / 64 * 64;
    const long long total = (long long)cout * taps * kpad * (mode == 2 ? 3 : 1);
    pack_conv_fprop_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(w, (bf16*)out, cout, cin, taps, kpad,
                                                                                c0, c0pad, mode == 2);
  } else {
    const int opad = (cout + 63) / 64 * 64;
    const long long total = (long long)cin * taps * opad;
    pack_conv_dgrad_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(w, (bf16*)out, cout, cin, taps, opad);
  }
  return check_launch("pack_conv_weight");
}

size_t b200unet_pack_convt_weight_bytes(int cin, int cout, int mode) {
  if (mode == 0 || mode == 2) return (size_t)4 * cout * ((cin + 63) / 64 * 64) * 2 * (mode == 2 ? 3 : 1);
  return (size_t)cin * 4 * ((cout + 63) / 64 * 64) * 2;
}

int b200unet_pack_convt_weight(const float* w, int cin, int cout, int mode, void* out, void* stream) {
  B200_REQUIRE(w && out && cin > 0 && cout > 0 && mode >= 0 && mode <= 2, "pack_convt_weight: bad arguments");
  const int kp = mode != 1 ? (cin + 63) / 64 * 64 : (cout + 63) / 64 * 64;
  const long long total = mode != 1 ? (long long)4 * cout * kp * (mode == 2 ? 3 : 1) : (long long)cin * 4 * kp;
  pack_convt_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(w, (bf16*)out, cin, cout, mode, kp);
  return check_launch("pack_convt_weight");
}

int b200unet_relu_mask(const b200_view* x, const void* mask, const b200_view* y, void* stream) {
  B200_REQUIRE(view_ok(x) && view_ok(y) && mask && same_extent(*x, *y), "relu_mask: bad arguments");
  DView m = dview(*y);
  m.p = (bf16*)mask;
  const long long total = view_pixels(*x) * x->c;
  relu_mask_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(dview(*x), m, dview(*y));
  return check_launch("relu_mask");
}

int b200unet_channel_sum(const b200_view* dz, float* out, void* workspace, size_t workspace_bytes, void* stream) {
  B200_REQUIRE(view_ok(dz) && out && workspace, "channel_sum: bad arguments");
  B200_REQUIRE(workspace_bytes >= reduce_workspace_bytes(dz->c, 1), "channel_sum: workspace too small");
  ReducePlan pl;
  int r = launch_chan_reduce<1, false>(SumF(), *dz, nullptr, (float*)workspace, &pl, as_stream(stream));
  if (r) return r;
  finalize_sum_kernel<<<finalize_grid(dz->c), kFinalizeThreads, 0, as_stream(stream)>>>((const float*)workspace, pl.blocks, dz->c, out);
  return check_launch("channel_sum finalize");
}
}
