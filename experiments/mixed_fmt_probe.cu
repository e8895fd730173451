// Hardware probe (not part of the product): does tcgen05.mma kind::f16 accept DIFFERENT 16-bit formats for A and B
// (instruction descriptor a_format / b_format: 0 = fp16, 1 = bf16)?  The split precision tier wants
// hi(x) [bf16] * lo(W) [fp16] and lo(x) [fp16] * hi(W) [bf16] in the same accumulator.
//   mixed_fmt_probe <a_fmt> <b_fmt>      prints max abs error against a CPU reference on the same rounded inputs
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_fp16.h>

#include "../pytorch-unet_b200/csrc/ptx.cuh"
#include "../pytorch-unet_b200/csrc/tmap.h"

using namespace b200;
constexpr int M = 128, N = 64, KTOT = 64;

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, int a_fmt, int b_fmt,
             float* __restrict__ d_out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + M * 128;
  __shared__ uint64_t bar_full, bar_mma;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(&bar_full, 1);
    mbar_init(&bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<64>(&tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    tma_load_2d(&map_a, &bar_full, sA, 0, 0);
    tma_load_2d(&map_b, &bar_full, sB, 0, 0);
    mbar_arrive_expect_tx(&bar_full, (M + N) * 128);
    mbar_wait(&bar_full, 0);
    tc_fence_after_sync();
    // umma_idesc_bf16 with the two format fields replaced
    uint32_t idesc = umma_idesc_bf16(M, N, 0, 0);
    idesc &= ~((7u << 7) | (7u << 10));
    idesc |= (uint32_t(a_fmt) << 7) | (uint32_t(b_fmt) << 10);
    const uint64_t hi = umma_desc_hi_sw128(16, 1024);
    for (int k = 0; k < KTOT / 16; ++k)
      umma_bf16(tmem, umma_desc(hi, smem_u32(sA) + k * 32), umma_desc(hi, smem_u32(sB) + k * 32), idesc, k > 0);
    umma_commit(&bar_mma);
  }
  __syncwarp();
  mbar_wait(&bar_mma, 0);
  tc_fence_after_sync();
  uint32_t v[32];
  for (int c = 0; c < N; c += 32) {
    tmem_ld_32x32(tmem + (uint32_t(warp * 32) << 16) + c, v);
    tmem_ld_wait();
    const int row = warp * 32 + (threadIdx.x & 31);
    for (int j = 0; j < 32; ++j) d_out[row * N + c + j] = __uint_as_float(v[j]);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<64>(tmem);
}

static uint16_t enc(float x, int fmt, float* back) {
  if (fmt == 1) {
    __nv_bfloat16 h = __float2bfloat16(x);
    *back = __bfloat162float(h);
    return *reinterpret_cast<uint16_t*>(&h);
  }
  __half h = __float2half(x);
  *back = __half2float(h);
  return *reinterpret_cast<uint16_t*>(&h);
}

int main(int argc, char** argv) {
  const int a_fmt = argc > 1 ? atoi(argv[1]) : 1, b_fmt = argc > 2 ? atoi(argv[2]) : 0;
  std::vector<float> hA(M * 64), hB(N * 64);
  std::vector<uint16_t> bA(M * 64), bB(N * 64);
  srand(99);
  // values with more than 8 significant bits so that an fp16 operand read as bf16 (or vice versa) is plainly wrong
  for (size_t i = 0; i < hA.size(); ++i) bA[i] = enc((rand() % 20001 - 10000) / 9973.0f, a_fmt, &hA[i]);
  for (size_t i = 0; i < hB.size(); ++i) bB[i] = enc((rand() % 20001 - 10000) / 9973.0f, b_fmt, &hB[i]);
  uint16_t *dA, *dB;
  float* dD;
  cudaMalloc(&dA, bA.size() * 2);
  cudaMalloc(&dB, bB.size() * 2);
  cudaMalloc(&dD, M * N * 4);
  cudaMemcpy(dA, bA.data(), bA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, bB.data(), bB.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap ma, mb;
  {
    uint64_t dims[2] = {64, (uint64_t)M};
    uint64_t strides[1] = {128};
    uint32_t box[2] = {64, (uint32_t)M};
    if (make_tmap_bf16(&ma, dA, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  }
  {
    uint64_t dims[2] = {64, (uint64_t)N};
    uint64_t strides[1] = {128};
    uint32_t box[2] = {64, (uint32_t)N};
    if (make_tmap_bf16(&mb, dB, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  }
  const int smem_bytes = (M + N) * 128 + 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  probe_kernel<<<1, 128, smem_bytes>>>(ma, mb, a_fmt, b_fmt, dD);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("a_fmt=%d b_fmt=%d: CUDA error %s\n", a_fmt, b_fmt, cudaGetErrorString(e));
    return 1;
  }
  std::vector<float> hD(M * N);
  cudaMemcpy(hD.data(), dD, M * N * 4, cudaMemcpyDeviceToHost);
  double max_err = 0, ref_max = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double acc = 0;
      for (int k = 0; k < KTOT; ++k) acc += double(hA[m * 64 + k]) * hB[n * 64 + k];
      max_err = fmax(max_err, fabs(acc - hD[m * N + n]));
      ref_max = fmax(ref_max, fabs(acc));
    }
  printf("a_fmt=%d b_fmt=%d (0 = fp16, 1 = bf16): max_abs_err %.3e ref_max %.3f %s\n", a_fmt, b_fmt, max_err, ref_max,
         max_err < 1e-4 * ref_max + 1e-5 ? "OK" : "MISMATCH");
  return 0;
}
