// Hardware probe (not part of the product): steady-state cycles per tcgen05.mma (M=128, K=16, bf16) as a function of
// N, operand majorness (K-major vs MN-major, SWIZZLE_128B) and A-row shift, with both operands resident in shared
// memory (no loads in the loop).  Prints cycles/MMA and the implied fraction of the 8192 FLOP/cycle/SM peak.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate_probe mma_rate_probe.cu
#include <cstdio>
#include <cstdlib>
#include "../pytorch-unet_b200/csrc/ptx.cuh"
using namespace b200;

struct P { int n, a_mn, b_mn, iters, shift; };

template <int N>
__global__ void __launch_bounds__(128, 1) probe(P p, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc<512>(&slot);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, N, p.a_mn, p.b_mn);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 64 * 1024);
    const uint64_t hiA = p.a_mn ? umma_desc_hi_sw128(32 * 1024, 1024) : umma_desc_hi_sw128(16, 1024);
    const uint64_t hiB = p.b_mn ? umma_desc_hi_sw128(16 * 1024, 1024) : umma_desc_hi_sw128(16, 1024);
    long long t0 = clock64();
    for (int i = 0; i < p.iters; ++i) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t aa = a0 + p.shift * 128 + (p.a_mn ? k * 2048 : k * 32);
        const uint32_t bb = b0 + (p.b_mn ? k * 2048 : k * 32);
        umma_bf16(tm + (i & 1) * N, umma_desc(hiA, aa), umma_desc(hiB, bb), idesc, 1);
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<512>(tm);
}

int main() {
  long long* d; cudaMalloc(&d, 148 * 8);
  const int smem = 161 * 1024 + 1024;
  int ns[6] = {64, 128, 256, 32, 192, 96};
  for (int ni = 0; ni < 6; ++ni)
    for (int maj = 0; maj < 4; ++maj)
      for (int shift = 0; shift < 2; ++shift) {
        P p{ns[ni], maj & 1, maj >> 1, 2000, shift * 33};
        auto launch = [&](auto kern) {
          cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
          kern<<<148, 128, smem>>>(p, d);
        };
        if (p.n == 64) launch(probe<64>); else if (p.n == 128) launch(probe<128>); else if (p.n == 256) launch(probe<256>); else if (p.n == 192) launch(probe<192>); else if (p.n == 96) launch(probe<96>); else launch(probe<32>);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
        const double cyc = avg / (p.iters * 4.0);
        printf("N=%3d A=%s B=%s shift=%2d : %7.1f cycles/MMA  (%5.1f%% of 8192 FLOP/cyc/SM) %s\n", p.n, p.a_mn ? "MN" : "K ", p.b_mn ? "MN" : "K ",
               p.shift, cyc, 100.0 * (2.0 * 128 * p.n * 16 / cyc) / 8192.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
      }
  return 0;
}
