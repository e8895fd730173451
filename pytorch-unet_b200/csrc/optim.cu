// torch.optim.Adam step (README.md:49, run.py:71) for ALL parameter tensors of the U-Net in ONE launch, fused with the
// refresh of the packed bf16 operand copies the tensor-core kernels read.
//
// Without it a training step ends with torch's multi-tensor Adam (6 launches) and begins with ~90 tiny pack launches
// (one per conv weight and direction).  Here a block owns either 2048 elements of a plain tensor (bias, BatchNorm,
// head) or a [16 x 32 x taps] tile of a conv weight: it reads param / grad / exp_avg / exp_avg_sq coalesced, applies the
// update, writes the three fp32 tensors back, keeps the new weights in shared memory and writes them as bf16 into the
// forward-operand layout and the backward-data layout (the two transposes go through shared memory so that both are
// written in runs of 32..64 bytes).  The job table lives in device memory; a block finds its job by binary search.
//
// Arithmetic = torch's fused Adam (aten/src/ATen/native/cuda/fused_adam_utils.cuh, amsgrad = false, maximize = false):
//   g += wd * p;  m += (g - m) * (1 - b1);  v = b2 * v + (1 - b2) * g * g;
//   p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
#include "common.cuh"

namespace b200 {

constexpr int kAdamThreads = 256;
constexpr int kAdamPlainPerBlock = 2048;
constexpr int kTileA = 16, kTileB = 32;  // tile of the two leading weight dimensions

struct AdamCoef {
  float b1, b2, eps, wd, step_size, inv_bc2_sqrt;
};

struct AdamHyper {
  float lr, b1, b2, eps, wd;
};

__device__ __forceinline__ AdamCoef load_coef(const AdamHyper& h, const float* __restrict__ step) {
  AdamCoef c;
  const float lr = h.lr;
  c.b1 = h.b1;
  c.b2 = h.b2;
  c.eps = h.eps;
  c.wd = h.wd;
  const double t = (double)step[0];  // device-resident counter, already counting this update (graph capturable)
  const double bc1 = 1.0 - pow((double)c.b1, t), bc2 = 1.0 - pow((double)c.b2, t);
  c.step_size = (float)((double)lr / bc1);
  c.inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
  return c;
}

__device__ __forceinline__ float adam_update(const AdamCoef& c, float p, float g, float& m, float& v) {
  if (c.wd != 0.f) g += c.wd * p;
  m += (g - m) * (1.f - c.b1);
  v = c.b2 * v + (1.f - c.b2) * g * g;
  const float denom = sqrtf(v) * c.inv_bc2_sqrt + c.eps;
  return p - c.step_size * (m / denom);
}

__device__ __forceinline__ void store_fwd(bf16* out, long long row_off, int k, int kpad, int split, float v) {
  const bf16 h = f2bf(v);
  out[row_off + k] = h;
  if (split) {  // {hi | hi | lo} along K (split precision tier)
    out[row_off + kpad + k] = h;
    out[row_off + 2 * kpad + k] = f2bf(split_lo(v, bf2f(h)));
  }
}

// One [16 x 32 x T] weight tile: update, then the two packed layouts.  T (filter taps) and FULL (a complete tile) are
// compile-time so that the index arithmetic is shifts and multiplies, not divisions by runtime values.
template <int T, bool FULL>
__device__ __forceinline__ void adam_tile(const b200_adam_job& J, const AdamCoef& c, int blk,
                                          float (*tile)[kTileB * 9 + 1]) {
  const int A = J.dim0, B = J.dim1;
  const int tiles_b = (B + kTileB - 1) / kTileB;
  const int a0 = (blk / tiles_b) * kTileA, b0 = (blk % tiles_b) * kTileB;
  const int nb = FULL ? kTileB : min(kTileB, B - b0), na = FULL ? kTileA : min(kTileA, A - a0);
  const int run = nb * T;  // contiguous floats per a-row
  const bool al16 = ((reinterpret_cast<uintptr_t>(J.param) | reinterpret_cast<uintptr_t>(J.grad) |
                      reinterpret_cast<uintptr_t>(J.exp_avg) | reinterpret_cast<uintptr_t>(J.exp_avg_sq)) & 15) == 0;
  if (FULL && B % 4 == 0 && al16) {
    // complete tile: a row is 32 * T contiguous floats starting at a multiple of 16 bytes -> four tensors x one 16-byte
    // access per thread and iteration instead of sixteen 4-byte ones (the kernel ran at 2.9 TB/s on scalar accesses)
    constexpr int RUN4 = kTileB * T / 4;
    for (int idx = threadIdx.x; idx < kTileA * RUN4; idx += kAdamThreads) {
      const int al = idx / RUN4, r4 = idx - al * RUN4;
      const long long i = ((long long)(a0 + al) * B + b0) * T + 4 * r4;
      const float4 p4 = *reinterpret_cast<const float4*>(J.param + i), g4 = *reinterpret_cast<const float4*>(J.grad + i);
      float4 m4 = *reinterpret_cast<const float4*>(J.exp_avg + i), v4 = *reinterpret_cast<const float4*>(J.exp_avg_sq + i);
      float4 o;
      o.x = adam_update(c, p4.x, g4.x, m4.x, v4.x);
      o.y = adam_update(c, p4.y, g4.y, m4.y, v4.y);
      o.z = adam_update(c, p4.z, g4.z, m4.z, v4.z);
      o.w = adam_update(c, p4.w, g4.w, m4.w, v4.w);
      *reinterpret_cast<float4*>(J.param + i) = o;
      *reinterpret_cast<float4*>(J.exp_avg + i) = m4;
      *reinterpret_cast<float4*>(J.exp_avg_sq + i) = v4;
      tile[al][4 * r4 + 0] = o.x;
      tile[al][4 * r4 + 1] = o.y;
      tile[al][4 * r4 + 2] = o.z;
      tile[al][4 * r4 + 3] = o.w;
    }
  } else {
    for (int idx = threadIdx.x; idx < na * run; idx += kAdamThreads) {
      const int al = idx / run, r = idx - al * run;
      const long long i = ((long long)(a0 + al) * B + b0) * T + r;
      float m = J.exp_avg[i], v = J.exp_avg_sq[i];
      const float p = adam_update(c, J.param[i], J.grad[i], m, v);
      J.param[i] = p;
      J.exp_avg[i] = m;
      J.exp_avg_sq[i] = v;
      tile[al][r] = p;
    }
  }
  __syncthreads();
  bf16* pf = reinterpret_cast<bf16*>(J.pack_fwd);
  bf16* pd = reinterpret_cast<bf16*>(J.pack_dgrad);
  if (J.kind == 1) {
    // conv w[o][c][tap]: forward operand [o][tap][kall] (K = source 0 | source 1, each padded to 64), backward-data
    // operand [c][taps-1-tap][opad]
    const int c0 = J.src0_c, c0pad = (c0 + 63) / 64 * 64;
    const int kpad = c0pad + (B > c0 ? (B - c0 + 63) / 64 * 64 : 0);
    const int kall = J.split ? 3 * kpad : kpad;
    if (pf) {
      for (int idx = threadIdx.x; idx < na * T * nb; idx += kAdamThreads) {
        const int bl = idx % nb, tap = (idx / nb) % T, al = idx / (nb * T);
        const int ch = b0 + bl;
        const int k = ch < c0 ? ch : c0pad + (ch - c0);
        store_fwd(pf, ((long long)(a0 + al) * T + tap) * kall, k, kpad, J.split, tile[al][bl * T + tap]);
      }
    }
    if (pd) {
      const int opad = (A + 63) / 64 * 64;
      for (int idx = threadIdx.x; idx < na * T * nb; idx += kAdamThreads) {
        const int al = idx % na, tap = (idx / na) % T, bl = idx / (na * T);
        pd[((long long)(b0 + bl) * T + (T - 1 - tap)) * opad + a0 + al] = f2bf(tile[al][bl * T + tap]);
      }
    }
  } else {
    // convT w[c][o][ab]: forward operand [(ab*O + o)][kall] (K = c), backward-data operand [c][ab][opad]
    const int kpad = (A + 63) / 64 * 64;
    const int kall = J.split ? 3 * kpad : kpad;
    if (pf) {
      for (int idx = threadIdx.x; idx < na * T * nb; idx += kAdamThreads) {
        const int al = idx % na, bl = (idx / na) % nb, ab = idx / (na * nb);
        store_fwd(pf, ((long long)ab * B + b0 + bl) * kall, a0 + al, kpad, J.split, tile[al][bl * T + ab]);
      }
    }
    if (pd) {
      const int opad = (B + 63) / 64 * 64;
      for (int idx = threadIdx.x; idx < na * T * nb; idx += kAdamThreads) {
        const int bl = idx % nb, ab = (idx / nb) % T, al = idx / (nb * T);
        pd[((long long)(a0 + al) * T + ab) * opad + b0 + bl] = f2bf(tile[al][bl * T + ab]);
      }
    }
  }
}

template <int T>
__device__ __forceinline__ void adam_tile_t(const b200_adam_job& J, const AdamCoef& c, int blk,
                                            float (*tile)[kTileB * 9 + 1]) {
  const int tiles_b = (J.dim1 + kTileB - 1) / kTileB;
  const int a0 = (blk / tiles_b) * kTileA, b0 = (blk % tiles_b) * kTileB;
  if (a0 + kTileA <= J.dim0 && b0 + kTileB <= J.dim1) adam_tile<T, true>(J, c, blk, tile);
  else adam_tile<T, false>(J, c, blk, tile);
}

__global__ void __launch_bounds__(kAdamThreads)
adam_pack_kernel(const b200_adam_job* __restrict__ jobs, int num_jobs, AdamHyper hyper, const float* __restrict__ step) {
  __shared__ float tile[kTileA][kTileB * 9 + 1];
  __shared__ AdamCoef s_coef;
  if (threadIdx.x == 0) s_coef = load_coef(hyper, step);  // two double-precision pow(): once per block, not per thread
  // job of this block: last j with block0 <= blockIdx.x
  int lo = 0, hi = num_jobs - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (jobs[mid].block0 <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const b200_adam_job J = jobs[lo];
  const int blk = (int)blockIdx.x - J.block0;
  __syncthreads();
  const AdamCoef c = s_coef;

  if (J.kind == 0) {
    const long long base = (long long)blk * kAdamPlainPerBlock;
#pragma unroll
    for (int u = 0; u < kAdamPlainPerBlock / kAdamThreads; ++u) {
      const long long i = base + u * kAdamThreads + threadIdx.x;
      if (i < J.numel) {
        float m = J.exp_avg[i], v = J.exp_avg_sq[i];
        J.param[i] = adam_update(c, J.param[i], J.grad[i], m, v);
        J.exp_avg[i] = m;
        J.exp_avg_sq[i] = v;
      }
    }
    return;
  }
  switch (J.taps) {
    case 9: adam_tile_t<9>(J, c, blk, tile); break;
    case 4: adam_tile_t<4>(J, c, blk, tile); break;
    case 1: adam_tile_t<1>(J, c, blk, tile); break;
    default: break;  // adam_plan only admits taps 1, 4, 9 for weight jobs
  }
}

// The job table travels to the device as kernel ARGUMENTS (32 jobs per launch): no pinned staging buffer, no copy to
// order against, and a CUDA graph captures the values themselves.
struct JobChunk {
  b200_adam_job j[32];
};
__global__ void adam_upload_kernel(const __grid_constant__ JobChunk c, b200_adam_job* __restrict__ dst, int n) {
  if ((int)threadIdx.x < n) dst[threadIdx.x] = c.j[threadIdx.x];
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200unet_adam_plan(b200_adam_job* jobs, int num_jobs) {
  B200_REQUIRE(jobs && num_jobs > 0, "adam_plan: bad arguments");
  long long blocks = 0;
  for (int j = 0; j < num_jobs; ++j) {
    b200_adam_job& J = jobs[j];
    B200_REQUIRE(J.numel > 0 && J.kind >= 0 && J.kind <= 2, "adam_plan: job %d: bad numel / kind", j);
    long long nb;
    if (J.kind == 0) {
      nb = (J.numel + kAdamPlainPerBlock - 1) / kAdamPlainPerBlock;
    } else {
      B200_REQUIRE(J.dim0 > 0 && J.dim1 > 0 && (J.taps == 1 || J.taps == 4 || J.taps == 9) &&
                       (long long)J.dim0 * J.dim1 * J.taps == J.numel,
                   "adam_plan: job %d: extents do not match numel", j);
      B200_REQUIRE(J.kind != 1 || (J.src0_c > 0 && J.src0_c <= J.dim1), "adam_plan: job %d: bad src0_c", j);
      nb = (long long)((J.dim0 + kTileA - 1) / kTileA) * ((J.dim1 + kTileB - 1) / kTileB);
    }
    B200_REQUIRE(blocks + nb < (1LL << 31), "adam_plan: too many blocks");
    J.block0 = (int)blocks;
    J.nblocks = (int)nb;
    blocks += nb;
  }
  return (int)blocks;
}

int b200unet_adam_upload(b200_adam_job* jobs_dev, const b200_adam_job* jobs_host, int num_jobs, void* stream) {
  B200_REQUIRE(jobs_dev && jobs_host && num_jobs > 0, "adam_upload: bad arguments");
  for (int j0 = 0; j0 < num_jobs; j0 += 32) {
    JobChunk c;
    const int n = num_jobs - j0 < 32 ? num_jobs - j0 : 32;
    for (int i = 0; i < n; ++i) c.j[i] = jobs_host[j0 + i];
    for (int i = n; i < 32; ++i) c.j[i] = b200_adam_job{};
    adam_upload_kernel<<<1, 32, 0, as_stream(stream)>>>(c, jobs_dev + j0, n);
    int r = check_launch("adam_upload");
    if (r) return r;
  }
  return 0;
}

int b200unet_adam_step(const b200_adam_job* jobs_dev, int num_jobs, int total_blocks, float lr, float beta1,
                       float beta2, float eps, float weight_decay, const float* step_dev, void* stream) {
  B200_REQUIRE(jobs_dev && step_dev && num_jobs > 0 && total_blocks > 0, "adam_step: bad arguments");
  const AdamHyper h{lr, beta1, beta2, eps, weight_decay};
  adam_pack_kernel<<<total_blocks, kAdamThreads, 0, as_stream(stream)>>>(jobs_dev, num_jobs, h, step_dev);
  return check_launch("adam_step");
}
}
