// Hardware probe (not part of the product): checks the UMMA shared-memory descriptor encodings this library
// relies on, in particular operand start addresses that are 128-byte-row shifted inside a SWIZZLE_128B tile
// (needed to reuse one haloed activation tile for all 3x3 filter taps).
//
//   umma_probe <a_major> <b_major> <a_shift> <b_shift> <base_offset_mode>
//     major: 0 = K-major (rows = M/N index, 64 bf16 of K contiguous), 1 = MN-major (rows = K index)
//     shift: operand start is moved by this many 128-byte rows
//     base_offset_mode: 0 -> descriptor base_offset field 0; 1 -> (start_addr >> 7) & 7
// Prints "max_abs_err <e> ref_norm <n>" comparing against a CPU fp32 reference on the same bf16 inputs.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../pytorch-unet_b200/csrc/ptx.cuh"
#include "../pytorch-unet_b200/csrc/tmap.h"

using namespace b200;

constexpr int M = 128, N = 64, KTOT = 64;
constexpr int A_ROWS_K = 160, B_ROWS_K = 96;  // K-major global row counts (rows = m / n index)
constexpr int PIX = 96;                       // MN-major global row count (rows = k index)

struct Params {
  int a_major, b_major, a_shift, b_shift, bo_mode;
};

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, Params p,
             float* __restrict__ d_out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;              // up to 2 * 12288 (MN) or 20480 (K) bytes
  uint8_t* sB = smem + 24576;      // 12288 bytes
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
    uint32_t bytes = 0;
    if (p.a_major == 0) {
      tma_load_2d(&map_a, &bar_full, sA, 0, 0);  // box 64 x 160
      bytes += A_ROWS_K * 128;
    } else {
      tma_load_2d(&map_a, &bar_full, sA, 0, 0);  // channels 0..63, box 64 x 96
      tma_load_2d(&map_a, &bar_full, sA + PIX * 128, 64, 0);
      bytes += 2 * PIX * 128;
    }
    tma_load_2d(&map_b, &bar_full, sB, 0, 0);
    bytes += (p.b_major == 0 ? B_ROWS_K : PIX) * 128;
    mbar_arrive_expect_tx(&bar_full, bytes);
    mbar_wait(&bar_full, 0);
    tc_fence_after_sync();

    const uint32_t idesc = umma_idesc_bf16(M, N, p.a_major, p.b_major);
    // K-major: SBO = 1024 (8 rows of 128 B), LBO unused (set to 16 B like CUTLASS).
    // MN-major: SBO = 1024 (8 k-rows of 128 B), LBO = byte distance between 64-element MN blocks.
    const uint64_t hiA = p.a_major == 0 ? umma_desc_hi_sw128(16, 1024) : umma_desc_hi_sw128(PIX * 128, 1024);
    const uint64_t hiB = p.b_major == 0 ? umma_desc_hi_sw128(16, 1024) : umma_desc_hi_sw128(PIX * 128, 1024);
    const uint32_t a0 = smem_u32(sA) + p.a_shift * 128;
    const uint32_t b0 = smem_u32(sB) + p.b_shift * 128;
    for (int k = 0; k < KTOT / 16; ++k) {
      const uint32_t aa = a0 + (p.a_major == 0 ? k * 32 : k * 2048);
      const uint32_t bb = b0 + (p.b_major == 0 ? k * 32 : k * 2048);
      const uint32_t boa = p.bo_mode ? ((aa >> 7) & 7) : 0;
      const uint32_t bob = p.bo_mode ? ((bb >> 7) & 7) : 0;
      umma_bf16(tmem, umma_desc(hiA, aa, boa), umma_desc(hiB, bb, bob), idesc, k > 0);
    }
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

static float bf16r(float x) { return __bfloat162float(__float2bfloat16(x)); }

int main(int argc, char** argv) {
  if (argc < 6) {
    fprintf(stderr, "usage: %s a_major b_major a_shift b_shift bo_mode\n", argv[0]);
    return 2;
  }
  Params p{atoi(argv[1]), atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), atoi(argv[5])};
  const int a_rows = p.a_major == 0 ? A_ROWS_K : PIX, a_cols = p.a_major == 0 ? 64 : 128;
  const int b_rows = p.b_major == 0 ? B_ROWS_K : PIX, b_cols = 64;
  std::vector<float> hA(a_rows * a_cols), hB(b_rows * b_cols);
  srand(1234);
  for (auto& x : hA) x = bf16r((rand() % 2001 - 1000) / 1000.0f);
  for (auto& x : hB) x = bf16r((rand() % 2001 - 1000) / 1000.0f);
  std::vector<__nv_bfloat16> bA(hA.size()), bB(hB.size());
  for (size_t i = 0; i < hA.size(); ++i) bA[i] = __float2bfloat16(hA[i]);
  for (size_t i = 0; i < hB.size(); ++i) bB[i] = __float2bfloat16(hB[i]);

  __nv_bfloat16 *dA, *dB;
  float* dD;
  cudaMalloc(&dA, bA.size() * 2);
  cudaMalloc(&dB, bB.size() * 2);
  cudaMalloc(&dD, M * N * 4);
  cudaMemcpy(dA, bA.data(), bA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, bB.data(), bB.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, M * N * 4);

  CUtensorMap ma, mb;
  {
    uint64_t dims[2] = {(uint64_t)a_cols, (uint64_t)a_rows};
    uint64_t strides[1] = {(uint64_t)a_cols * 2};
    uint32_t box[2] = {64, (uint32_t)a_rows};
    int r = make_tmap_bf16(&ma, dA, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (r) { printf("tmap A failed %d\n", r); return 1; }
  }
  {
    uint64_t dims[2] = {(uint64_t)b_cols, (uint64_t)b_rows};
    uint64_t strides[1] = {(uint64_t)b_cols * 2};
    uint32_t box[2] = {64, (uint32_t)b_rows};
    int r = make_tmap_bf16(&mb, dB, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (r) { printf("tmap B failed %d\n", r); return 1; }
  }
  const int smem_bytes = 24576 + 12288 + 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  probe_kernel<<<1, 128, smem_bytes>>>(ma, mb, p, dD);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("cfg %d %d %d %d %d: CUDA error %s\n", p.a_major, p.b_major, p.a_shift, p.b_shift, p.bo_mode,
           cudaGetErrorString(e));
    return 1;
  }
  std::vector<float> hD(M * N);
  cudaMemcpy(hD.data(), dD, M * N * 4, cudaMemcpyDeviceToHost);

  double max_err = 0, ref_norm = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double acc = 0;
      for (int k = 0; k < KTOT; ++k) {
        const float a = p.a_major == 0 ? hA[(m + p.a_shift) * 64 + k] : hA[(k + p.a_shift) * 128 + m];
        const float b = p.b_major == 0 ? hB[(n + p.b_shift) * 64 + k] : hB[(k + p.b_shift) * 64 + n];
        acc += double(a) * b;
      }
      max_err = fmax(max_err, fabs(acc - hD[m * N + n]));
      ref_norm = fmax(ref_norm, fabs(acc));
    }
  printf("cfg a_major=%d b_major=%d a_shift=%d b_shift=%d bo_mode=%d : max_abs_err %.5f ref_max %.3f %s\n", p.a_major,
         p.b_major, p.a_shift, p.b_shift, p.bo_mode, max_err, ref_norm, max_err < 1e-2 ? "OK" : "MISMATCH");
  return 0;
}
