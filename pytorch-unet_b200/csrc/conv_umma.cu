// tcgen05 implicit-GEMM kernel for the convolution family (forward and backward-data):
//   conv 3x3 / 1x1 forward (+bias, ReLU, two concatenated sources = folded crop + concat)   unet.py:92-98, 152-163
//   conv 3x3 / 1x1 backward-data (+ReLU mask of the producer, two destinations)             autograd of the above
//   ConvTranspose2d(2, 2) forward (GEMM with N = 4*Cout + pixel-shuffle store)              unet.py:143, 173
//   ConvTranspose2d(2, 2) backward-data (GEMM with K = 4*Cout over four stride-2 sub-lattices)
//
// GEMM view: D[pixels, Cout] = sum over (source, 64-channel chunk, filter tap) A[pixels, 64] * B[Cout, 64]^T.
//
// A operand ("haloed tile"): one TMA box {64 channels, P columns, TH+2 rows} of the NHWC activation is loaded ONCE
// per 64-channel chunk into SWIZZLE_128B shared memory; its rows are the pixels of the haloed tile in row-major
// order with pitch P.  The operand of filter tap (r, s) for M-block m is the same buffer read through a UMMA
// descriptor whose start address is advanced by (m*128 + r*P + s) 128-byte rows: output position q = y*P + x reads
// row q + r*P + s = (y+r)*P + (x+s).  Positions with x >= P-2 are wasted MMA rows (2 of P) and are discarded by
// the epilogue.  So the activation tile moves L2 -> SMEM once, not nine times, and zero padding / image borders
// come from TMA out-of-bounds fill.  (experiments/umma_probe.cu verified on B200 that row-shifted descriptor
// starts inside a 1024-byte-aligned SWIZZLE_128B buffer are addressed correctly with base_offset = 0.)
//
// B operand: packed weights [Cout][tap][kpad] (K-major), one TMA box {64, BN} per (tap, chunk).
//
// Warp roles (11 warps, 1 CTA / SM, persistent over tiles): 0 = A producer, 1 = B producer, 2 = TMEM allocator +
// tcgen05.mma issuer (warp-uniform control flow, one elected lane issues), 3..10 = epilogue (two warps per TMEM lane
// quarter; bias pre-loaded into the accumulator with tcgen05.st, tcgen05.ld -> cvt.relu.bf16x2 -> per-warp shared
// memory transpose -> mask -> full-line global stores).
// Accumulators: MB x BN fp32 columns in TMEM, double-buffered when 2*MB*BN <= 512 so the epilogue of tile i overlaps
// the MMAs of tile i+1.
// CTA pairs (CL = 2, tcgen05 cta_group::2): two CTAs of a cluster work on two pixel tiles that need the SAME weights.
// One MMA of M = 256 covers both: each CTA supplies its own 128 activation rows and HALF of the weight rows from its
// shared memory and keeps its 128 accumulator rows in its own TMEM.  Weight bytes moved L2 -> SMEM per CTA are halved,
// which is what bounds the deep layers: the L2 delivers ~43 B/cycle/SM chip-wide (B300_MICROARCH: LTS cap ~6300 B/cycle),
// while a single-CTA BN = 256 tile asks for ~36 B/cycle.  (Merely multicasting the weight tile into both CTAs, the first
// version of CL = 2, was perf-neutral: every SM still ingests the full tile.)  Only the leader CTA issues MMAs; TMA loads
// of both CTAs complete on the leader's "full" barriers; tcgen05.commit multicasts the "empty" / "accumulator ready"
// arrivals to both; the peer's epilogue warps hand their TMEM buffer back with a remote mbarrier arrive.
//
// Split precision tier (b200unet.h): the forward kernels accept activations carried as hi + lo bf16 planes.  The K loop
// then runs three operand passes into the same accumulator — hi(x)*hi(W), lo(x)*hi(W), hi(x)*lo(W); the dropped
// lo*lo term is 2^-18 relative — simply as three times as many A sources against a weight tensor packed
// {hi | hi | lo} along K, and the epilogue writes the fp32 result as two planes (hi = bf16(v), lo = bf16(v - hi)).
#include <cstdlib>

#include "conv_impl.h"
#include "ptx.cuh"
#include "tmap.h"

namespace b200 {

constexpr int kUmmaThreads = 352;  // 3 control warps + 8 epilogue warps (two per TMEM lane quarter)
constexpr int kEpiWarps = 8;
constexpr int kNA = 2;        // A (activation tile) stages of a 3x3 convolution: one stage feeds nine taps of MMAs
constexpr int kMaxNA = 6;     // 1x1 / transposed convolutions: a stage feeds ONE tap (512 MMA cycles at MB = 2), less than the
                              // L2 round trip of its 32 KB TMA box, so two stages left the tensor pipe 27-35 % busy (ncu,
                              // profiles/r02_conv_metrics.txt launches 9/12/15/18): they get as many stages as fit
constexpr int kMaxNB = 8;     // B (weight tile) stages
constexpr uint32_t kSmemBudget = 188 * 1024;   // A + B stages
constexpr uint32_t kStageOutBytes = kEpiWarps * 4096;  // epilogue transpose buffers: per warp [32 rows][64 cols] bf16
constexpr int kBiasSmemFloats = 1024;                  // bias of every N tile of the launch, staged once (when it fits)

constexpr int kMaxA = 6;  // A sources: 2 concat sources x 3 split-tier passes, or the 4 convT sub-lattices

struct TileMaps {
  CUtensorMap a[kMaxA];
  CUtensorMap b;
};

struct UmmaArgs {
  int num_a;
  int a_c[kMaxA];
  int a_koff[kMaxA];
  int taps, kx, kpad, pad;
  int P, TH, TW;
  int tiles_x, tiles_y, n_img, n_ntiles;
  int Ho, Wo, cout_total;
  DView dst[4];
  bf16* dst_lo[4];  // split tier: low-order output planes (same geometry as dst), else nullptr
  const bf16* mask[4];
  int ndst, dst_c0;
  const float* bias;
  int relu;
  uint32_t a_stage_bytes, a_tx_bytes;
  int nb_stages, na_stages;
};

struct TileCoord {
  int n, y0, x0, n0;
};

// CL = 1: work item = tile, N tile fastest.  CL = 2 (CTA pair, tcgen05 cta_group::2): work item =
// pair of neighbouring pixel tiles with the SAME N tile; CTA `rank` of the cluster takes pixel tile 2*pair + rank (an
// odd tail is duplicated: both CTAs then compute and store the same tile, which is harmless).
template <int CL>
__device__ __forceinline__ TileCoord decode_tile(const UmmaArgs& a, int work, int rank, int bn) {
  TileCoord t;
  const int nt = work % a.n_ntiles;
  int pt = work / a.n_ntiles;
  if (CL == 2) pt = min(2 * pt + rank, a.tiles_x * a.tiles_y * a.n_img - 1);
  const int txi = pt % a.tiles_x;
  pt /= a.tiles_x;
  const int tyi = pt % a.tiles_y;
  t.n = pt / a.tiles_y;
  t.y0 = tyi * a.TH;
  t.x0 = txi * a.TW;
  t.n0 = nt * bn;
  return t;
}

// 168 registers is the ceiling for 11 warps: an SM sub-partition (16384 registers) hosts three of them
template <int MB, int BN, int CL>
__global__ void __launch_bounds__(kUmmaThreads, 1)
umma_conv_kernel(const __grid_constant__ TileMaps maps, const __grid_constant__ UmmaArgs a) {
  constexpr int NBUF = (2 * MB * BN <= 512) ? 2 : 1;
  constexpr uint32_t TMEM_COLS = NBUF * MB * BN;
  static_assert(TMEM_COLS >= 32 && TMEM_COLS <= 512 && (TMEM_COLS & (TMEM_COLS - 1)) == 0, "TMEM columns");
  constexpr uint32_t B_STAGE_BYTES = (BN / CL) * 128;  // CL = 2: this CTA's half of the weight rows

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = sA + (uint32_t)a.na_stages * a.a_stage_bytes;
  uint8_t* sOut = sB + (uint32_t)a.nb_stages * B_STAGE_BYTES;
  float* sBias = reinterpret_cast<float*>(sOut + kStageOutBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + kBiasSmemFloats);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + kMaxNA;
  uint64_t* b_full = a_empty + kMaxNA;
  uint64_t* b_empty = b_full + kMaxNB;
  uint64_t* t_full = b_empty + kMaxNB;
  uint64_t* t_empty = t_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // work iteration space (see decode_tile)
  const int rank = CL == 2 ? (int)cluster_ctarank() : 0;
  const int pix_tiles = a.tiles_x * a.tiles_y * a.n_img;
  const int total_tiles = (CL == 2 ? (pix_tiles + 1) / 2 : pix_tiles) * a.n_ntiles;
  const int work0 = CL == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int work_step = CL == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kMaxNA; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < kMaxNB; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&t_full[i], 1);
      mbar_init(&t_empty[i], CL * kEpiWarps);  // CL = 2: the leader's barrier also collects the peer's epilogue warps
    }
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < a.num_a; ++s) tma_prefetch_desc(&maps.a[s]);
    tma_prefetch_desc(&maps.b);
  }
  // Bias of all N tiles in shared memory (column nc of the GEMM -> channel nc - d * dst_c0 of its destination d; zero past
  // cout_total).  The epilogue pre-loads it into the TMEM accumulators; fetching it from global memory there put an L2
  // round trip on the critical path of every (M-block, pass) item (ncu: 15 % of the epilogue warps' time).
  const bool bias_smem = a.bias != nullptr && a.n_ntiles * BN <= kBiasSmemFloats;
  if (bias_smem) {
    for (int i = threadIdx.x; i < a.n_ntiles * BN; i += kUmmaThreads) {
      const int d = a.ndst > 1 ? min(i / a.dst_c0, a.ndst - 1) : 0;
      sBias[i] = i < a.cout_total ? a.bias[i - d * a.dst_c0] : 0.f;
    }
  }
  if (warp == 2) {
    if (CL == 2) tmem_alloc2<TMEM_COLS>(tmem_slot); else tmem_alloc<TMEM_COLS>(tmem_slot);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (CL == 2) cluster_sync_all();  // the peer's barriers must be initialised before anything arrives on them
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= A producer: one haloed activation tile per (tile, source, 64-channel chunk)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = work0; tile < total_tiles; tile += work_step) {
        const TileCoord t = decode_tile<CL>(a, tile, rank, BN);
        for (int s = 0; s < a.num_a; ++s) {
          for (int c0 = 0; c0 < a.a_c[s]; c0 += 64) {
            mbar_wait(&a_empty[stage], phase ^ 1);
            if (CL == 2) {  // both CTAs' tiles complete on the leader's barrier
              if (rank == 0) mbar_arrive_expect_tx(&a_full[stage], 2 * a.a_tx_bytes);
              tma_load_4d_2cta(&maps.a[s], &a_full[stage], sA + stage * a.a_stage_bytes, c0, t.x0 - a.pad, t.y0 - a.pad, t.n);
            } else {
              mbar_arrive_expect_tx(&a_full[stage], a.a_tx_bytes);
              tma_load_4d(&maps.a[s], &a_full[stage], sA + stage * a.a_stage_bytes, c0, t.x0 - a.pad, t.y0 - a.pad, t.n);
            }
            if (++stage == a.na_stages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================= B producer: one [BN x 64] weight tile per (tile, source, chunk, tap)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = work0; tile < total_tiles; tile += work_step) {
        const TileCoord t = decode_tile<CL>(a, tile, rank, BN);
        for (int s = 0; s < a.num_a; ++s) {
          for (int c0 = 0; c0 < a.a_c[s]; c0 += 64) {
            for (int tap = 0; tap < a.taps; ++tap) {
              mbar_wait(&b_empty[stage], phase ^ 1);
              if (CL == 2) {
                // this CTA holds weight rows [rank*BN/2, +BN/2) only; the MMA reads the other half from the peer
                if (rank == 0) mbar_arrive_expect_tx(&b_full[stage], 2 * B_STAGE_BYTES);
                tma_load_2d_2cta(&maps.b, &b_full[stage], sB + stage * B_STAGE_BYTES, tap * a.kpad + a.a_koff[s] + c0,
                                 t.n0 + rank * (BN / 2));
              } else {
                mbar_arrive_expect_tx(&b_full[stage], B_STAGE_BYTES);
                tma_load_2d(&maps.b, &b_full[stage], sB + stage * B_STAGE_BYTES, tap * a.kpad + a.a_koff[s] + c0, t.n0);
              }
              if (++stage == a.nb_stages) {
                stage = 0;
                phase ^= 1;
              }
            }
          }
        }
      }
    }
  } else if (warp == 2) {
    // ================= MMA issuer.  The whole warp runs the control flow so that addresses and descriptors are
    // warp-uniform (uniform registers; a lane-0 branch costs an ELECT + 4 R2UR per MMA and halves the issue rate at
    // N = 64); one elected lane issues the tcgen05 instructions.
    constexpr uint32_t idesc = umma_idesc_bf16(128 * CL, BN, 0, 0);
    constexpr uint64_t desc_hi = umma_desc_hi_sw128(16, 1024);
    int astage = 0, bstage = 0;
    uint32_t aphase = 0, bphase = 0;
    int it = 0;
    const int n_taps = a.taps, kx = a.kx, pitch = a.P, n_bstages = a.nb_stages, n_astages = a.na_stages;
    const uint32_t a_stage_bytes = a.a_stage_bytes;
    // CL = 2: the leader issues for the pair; the peer's MMA warp only took part in the TMEM allocation
    for (int tile = (CL == 2 && rank != 0) ? total_tiles : work0; tile < total_tiles; tile += work_step, ++it) {
      const int buf = it % NBUF;
      mbar_wait(&t_empty[buf], (it / NBUF) & 1);  // epilogue hands every buffer over, also before its first use
      tc_fence_after_sync();
      const uint32_t acc = tmem_base + buf * (MB * BN);
      uint32_t accum = a.bias ? 1u : 0u;  // with a bias the epilogue warps pre-loaded it into the accumulator
      for (int s = 0; s < a.num_a; ++s) {
        for (int c0 = 0; c0 < a.a_c[s]; c0 += 64) {
          mbar_wait(&a_full[astage], aphase);
          // K steps of 16 channels that hold real data in this chunk: narrow layers (16 / 32 channels: the feature net, the
          // first levels of wf = 5 networks) load a 64-wide box whose tail is TMA zero fill — issuing MMAs on it was 2-4x
          // the necessary tensor work
          const int c_left = a.a_c[s] - c0;  // K step kk holds data iff 16 * kk < c_left
          const uint32_t a_base = smem_u32(sA + astage * a_stage_bytes);
          int tap_s = -1;
          uint32_t tap_row_off = 0;
          const uint64_t a_desc0 = umma_desc(desc_hi, a_base);
          for (int tap = 0; tap < n_taps; ++tap) {
            mbar_wait(&b_full[bstage], bphase);
            tc_fence_after_sync();
            // tap (r, s) reads the A tile r*P + s rows further: 128 B per row = 8 descriptor units
            if (++tap_s == kx) {
              tap_s = 0;
              tap_row_off += (uint32_t)(pitch - kx + 1) * 8;
            } else if (tap != 0) {
              tap_row_off += 8;
            }
            const uint64_t a_desc = a_desc0 + tap_row_off;
            const uint64_t b_desc = umma_desc(desc_hi, smem_u32(sB + bstage * B_STAGE_BYTES));
            if (elect_one()) {
              if (c_left >= 64) {  // full chunk: the unpredicated sequence (the issuing lane has 48 cycles per N = 64 MMA)
#pragma unroll
                for (int mb = 0; mb < MB; ++mb) {
                  // M-block m starts 128 rows (16 KiB -> 1024 in the >>4 address field) further; K step = 32 B -> 2.
                  // The first MMA into each M-block's accumulator overwrites, everything after accumulates.
                  if (CL == 2) {
                    umma_bf16_2cta(acc + mb * BN, a_desc + (uint64_t)mb * 1024, b_desc, idesc, accum);
#pragma unroll
                    for (int kk = 1; kk < 4; ++kk)
                      umma_bf16_2cta(acc + mb * BN, a_desc + (uint64_t)(mb * 1024 + kk * 2), b_desc + (uint64_t)(kk * 2), idesc, 1u);
                  } else {
                    umma_bf16(acc + mb * BN, a_desc + (uint64_t)mb * 1024, b_desc, idesc, accum);
#pragma unroll
                    for (int kk = 1; kk < 4; ++kk)
                      umma_bf16(acc + mb * BN, a_desc + (uint64_t)(mb * 1024 + kk * 2), b_desc + (uint64_t)(kk * 2), idesc, 1u);
                  }
                }
              } else {  // narrow chunk: only the K steps that hold data
#pragma unroll
                for (int mb = 0; mb < MB; ++mb) {
                  if (CL == 2) {
                    umma_bf16_2cta(acc + mb * BN, a_desc + (uint64_t)mb * 1024, b_desc, idesc, accum);
#pragma unroll
                    for (int kk = 1; kk < 4; ++kk)
                      if (16 * kk < c_left)
                        umma_bf16_2cta(acc + mb * BN, a_desc + (uint64_t)(mb * 1024 + kk * 2), b_desc + (uint64_t)(kk * 2), idesc, 1u);
                  } else {
                    umma_bf16(acc + mb * BN, a_desc + (uint64_t)mb * 1024, b_desc, idesc, accum);
#pragma unroll
                    for (int kk = 1; kk < 4; ++kk)
                      if (16 * kk < c_left)
                        umma_bf16(acc + mb * BN, a_desc + (uint64_t)(mb * 1024 + kk * 2), b_desc + (uint64_t)(kk * 2), idesc, 1u);
                  }
                }
              }
              if (CL == 2)
                umma_commit_2cta(&b_empty[bstage], (uint16_t)0x3);
              else
                umma_commit(&b_empty[bstage]);
            }
            __syncwarp();
            accum = 1;
            if (++bstage == n_bstages) {
              bstage = 0;
              bphase ^= 1;
            }
          }
          if (elect_one()) {
            if (CL == 2) umma_commit_2cta(&a_empty[astage], (uint16_t)0x3); else umma_commit(&a_empty[astage]);
          }
          __syncwarp();
          if (++astage == n_astages) {
            astage = 0;
            aphase ^= 1;
          }
        }
      }
      if (elect_one()) {
        if (CL == 2) umma_commit_2cta(&t_full[buf], (uint16_t)0x3); else umma_commit(&t_full[buf]);
      }
      __syncwarp();
    }
  } else {
    // ================= epilogue: 4 warps, warp w owns TMEM lanes 32*(w%4) .. +31 (= 32 output positions).
    // Measured: with a row-per-lane store, and then with fp32 bias/ReLU math on the store side, the 64-channel layers
    // were bound by this epilogue's instruction count (~1.4 TB/s of output), not by the MMAs.  So:
    //  * the bias is PRE-LOADED into the TMEM accumulator (tcgen05.st) before the tile's first MMA, which then
    //    accumulates from the start: no FADD per element;
    //  * ReLU is the .relu modifier of the fp32 -> bf16x2 conversion: no FMNMX per element;
    //  * the converted row goes through a per-warp shared-memory transpose (16-byte chunks XOR-swizzled by row) so that
    //    CP/8 consecutive lanes own one position's contiguous CP*2 bytes: mask loads and stores are full lines.
    constexpr int CP = BN < 64 ? BN : 64;   // columns per pass
    constexpr int LPR = CP / 8;             // lanes per row on the store side (one 16-byte bf16 chunk each)
    constexpr int RPI = 32 / LPR;           // rows per store iteration
    const int quarter = warp & 3;     // TMEM lane quarter this warp may access
    const int egrp = (warp - 3) >> 2;  // two warps share a quarter: they split the (M-block, column pass) work items
    uint4* stg = reinterpret_cast<uint4*>(sOut + (warp - 3) * 4096);   // [32 rows][LPR chunks] of 16 B
    const int rsub = lane / LPR, cch = lane % LPR;
    const uint32_t lane_base = uint32_t(quarter * 32) << 16;
    // position of this lane's accumulator row inside the tile, per M-block: fixed for the whole kernel (P is), so the
    // division is paid once here and not in every (tile, M-block, pass) iteration
    int pos_ty[MB], pos_tx[MB];
#pragma unroll
    for (int mb = 0; mb < MB; ++mb) {
      const int q = mb * 128 + quarter * 32 + lane;
      pos_ty[mb] = q / a.P;
      pos_tx[mb] = q - pos_ty[mb] * a.P;
      if (pos_ty[mb] >= a.TH || pos_tx[mb] >= a.TW) pos_ty[mb] = 1 << 20;  // halo column / beyond the tile: never stored
    }
    const uint32_t stg_s = smem_u32(stg);  // explicit shared-memory accesses (generic LD/ST cost an address-space check)

    // writes bias[n0 .. n0+BN) into the rows of accumulator buffer `buf` — each warp exactly the (M-block, column
    // pass) regions it drains itself, so a fast warp never overwrites a region its partner is still reading
    constexpr int PASSES = BN / CP;
    // one (M-block, column pass) region: CP columns of bias for the tile whose N origin is n0
    auto preload_item = [&](int n0, int buf, int mb, int c0) {
#pragma unroll 1
      for (int col0 = c0; col0 < c0 + CP; col0 += 32) {
        const int nc = n0 + col0;
        uint32_t bv[32];
        if (bias_smem) {  // warp-uniform broadcast reads
          const uint32_t bs = smem_u32(sBias + nc);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint4 f = ld_shared_v4(bs + 16 * j);
            bv[4 * j] = f.x;
            bv[4 * j + 1] = f.y;
            bv[4 * j + 2] = f.z;
            bv[4 * j + 3] = f.w;
          }
        } else {
          // several destinations (ConvTranspose quadrants) share one bias vector: column -> channel inside its destination
          const int d = a.ndst > 1 ? min(nc / a.dst_c0, a.ndst - 1) : 0;
          const float* bp = a.bias + (nc - d * a.dst_c0);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
            if (nc + 4 * j < a.cout_total) f = *reinterpret_cast<const float4*>(bp + 4 * j);
            bv[4 * j] = __float_as_uint(f.x);
            bv[4 * j + 1] = __float_as_uint(f.y);
            bv[4 * j + 2] = __float_as_uint(f.z);
            bv[4 * j + 3] = __float_as_uint(f.w);
          }
        }
        tmem_st_32x32(tmem_base + buf * (MB * BN) + mb * BN + col0 + lane_base, bv);
      }
    };
    auto preload_bias = [&](int tile, int buf) {
      if (a.bias == nullptr) return;
      const TileCoord t = decode_tile<CL>(a, tile, rank, BN);
#pragma unroll 1
      for (int wi = egrp; wi < MB * PASSES; wi += 2) {
        const int mb = wi / PASSES;
        preload_item(t.n0, buf, mb, (wi - mb * PASSES) * CP);
      }
      tmem_st_wait();
    };

    // hand the (pre-loaded) buffers of the first NBUF tiles to the MMA warp
    for (int b = 0; b < NBUF; ++b) {
      const int tile = work0 + b * work_step;
      if (tile < total_tiles) preload_bias(tile, b);
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) {
        if (CL == 2) mbar_arrive_remote(&t_empty[b], 0); else mbar_arrive(&t_empty[b]);
      }
    }

    int it = 0;
    for (int tile = work0; tile < total_tiles; tile += work_step, ++it) {
      const TileCoord t = decode_tile<CL>(a, tile, rank, BN);
      const int buf = it % NBUF;
      // this buffer's next user is tile + NBUF * work_step: its bias is stored region by region, right after a region
      // has been drained, so that the asynchronous tcgen05.st overlaps the rest of the epilogue (a single store + wait
      // block at the end of the tile cost ~20 % of the epilogue warps' time on the 64-channel layers)
      const int next = tile + NBUF * work_step;
      const bool pre = a.bias != nullptr && next < total_tiles;
      const int next_n0 = pre ? decode_tile<CL>(a, next, rank, BN).n0 : 0;
      mbar_wait(&t_full[buf], (it / NBUF) & 1);
      tc_fence_after_sync();
#pragma unroll 1  // keep the epilogue body small: fully unrolled it was ~3000 instructions and ran out of the I-cache
      for (int wi = egrp; wi < MB * PASSES; wi += 2) {
        const int mb = wi / PASSES;
        const int col0 = (wi - mb * PASSES) * CP;
        // destination of this 64-column pass (a tile may span several destinations, a pass never does)
        const int d = a.ndst > 1 ? min((t.n0 + col0) / a.dst_c0, a.ndst - 1) : 0;
        const int ch0 = t.n0 + col0 - d * a.dst_c0;
        const DView& dst = a.dst[d];
        const bf16* mk = a.mask[d];
        int ty = pos_ty[0], tx = pos_tx[0];
#pragma unroll
        for (int m = 1; m < MB; ++m)
          if (m == mb) {
            ty = pos_ty[m];
            tx = pos_tx[m];
          }
        const int oy = t.y0 + ty, ox = t.x0 + tx;
        const bool valid = oy < a.Ho && ox < a.Wo;
        // element offset inside image t.n of the destination; fits 32 bits (checked on the host); -1 = nothing to store
        const int off = valid ? (int)(oy * dst.sh + ox * dst.sw + ch0) : -1;
        const long long img = (long long)t.n * dst.sn;
        const uint32_t taddr = tmem_base + buf * (MB * BN) + mb * BN + lane_base;
        bf16* const dlo = a.dst_lo[d];
        // split tier: a second pass over the same accumulators writes the low-order plane (the TMEM read is repeated
        // rather than keeping 64 values live across the store phase)
#pragma unroll 1
        for (int plane = 0; plane < (dlo ? 2 : 1); ++plane) {
          uint32_t v[CP];
          {
            uint32_t (&v0)[32] = *reinterpret_cast<uint32_t (*)[32]>(&v[0]);
            tmem_ld_32x32(taddr + col0, v0);
            if (CP == 64) {
              uint32_t (&v1)[32] = *reinterpret_cast<uint32_t (*)[32]>(&v[CP - 32]);
              tmem_ld_32x32(taddr + col0 + 32, v1);
            }
          }
          tmem_ld_wait();
          // row-owner side: convert this position's CP accumulators (ReLU folded into the conversion) and park them as
          // LPR 16-byte chunks; chunk j goes to slot j ^ (lane & (LPR-1)) so the 32 rows do not collide on banks
          if (plane == 0) {
#pragma unroll
            for (int j = 0; j < LPR; ++j) {
              uint4 c;
              if (a.relu) {
                c.x = pack_bf16x2_relu(__uint_as_float(v[8 * j + 0]), __uint_as_float(v[8 * j + 1]));
                c.y = pack_bf16x2_relu(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3]));
                c.z = pack_bf16x2_relu(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5]));
                c.w = pack_bf16x2_relu(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7]));
              } else {
                c.x = pack_bf16x2(__uint_as_float(v[8 * j + 0]), __uint_as_float(v[8 * j + 1]));
                c.y = pack_bf16x2(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3]));
                c.z = pack_bf16x2(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5]));
                c.w = pack_bf16x2(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7]));
              }
              st_shared_v4(stg_s + (uint32_t)(lane * LPR + (j ^ (lane & (LPR - 1)))) * 16, c);
            }
          } else {
            // lo = bf16(v' - bf16(v')), v' = the (ReLU'd) fp32 result
#pragma unroll
            for (int j = 0; j < LPR; ++j) {
              float r[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                float x = __uint_as_float(v[8 * j + e]);
                if (a.relu) x = fmaxf(x, 0.f);
                r[e] = split_lo(x, bf2f(f2bf(x)));
              }
              st_shared_v4(stg_s + (uint32_t)(lane * LPR + (j ^ (lane & (LPR - 1)))) * 16, pack8(r));
            }
          }
          __syncwarp();
          // store side: lane (rsub, cch) moves chunk cch of rows rsub, rsub + RPI, ...
          bf16* const dp = (plane ? dlo : dst.p) + img + cch * 8;
          const bool col_ok = t.n0 + col0 + cch * 8 < a.cout_total;
          int offs[LPR];
#pragma unroll
          for (int i = 0; i < LPR; ++i) offs[i] = __shfl_sync(0xffffffffu, off, i * RPI + rsub);
          if (col_ok) {
            if (mk) {  // dgrad: zero where the producer's ReLU was inactive (mask <= 0)
              bf16x8 mv[LPR];
#pragma unroll
              for (int i = 0; i < LPR; ++i)
                if (offs[i] >= 0) mv[i] = *reinterpret_cast<const bf16x8*>(mk + img + cch * 8 + offs[i]);
#pragma unroll
              for (int i = 0; i < LPR; ++i) {
                const int R = i * RPI + rsub;
                if (offs[i] >= 0) {
                  uint4 c = ld_shared_v4(stg_s + (uint32_t)(R * LPR + (cch ^ (R & (LPR - 1)))) * 16);
                  c.x &= bf16x2_gt0_mask(mv[i].x);  // on the packed pairs: 2 instructions per pair
                  c.y &= bf16x2_gt0_mask(mv[i].y);
                  c.z &= bf16x2_gt0_mask(mv[i].z);
                  c.w &= bf16x2_gt0_mask(mv[i].w);
                  *reinterpret_cast<uint4*>(dp + offs[i]) = c;
                }
              }
            } else {
#pragma unroll
              for (int i = 0; i < LPR; ++i) {
                const int R = i * RPI + rsub;
                if (offs[i] >= 0)
                  *reinterpret_cast<uint4*>(dp + offs[i]) = ld_shared_v4(stg_s + (uint32_t)(R * LPR + (cch ^ (R & (LPR - 1)))) * 16);
              }
            }
          }
          __syncwarp();
        }
        if (pre) preload_item(next_n0, buf, mb, col0);  // this region is drained: store the next user's bias into it
      }
      if (pre) tmem_st_wait();  // the bias stores issued along the way; then hand the buffer back
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) {
        if (CL == 2) mbar_arrive_remote(&t_empty[buf], 0); else mbar_arrive(&t_empty[buf]);
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (CL == 2) cluster_sync_all();  // the pair's MMAs read this CTA's shared memory and arrive on its barriers
  if (warp == 2) {
    if (CL == 2) tmem_dealloc2<TMEM_COLS>(tmem_base); else tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// ------------------------------------------------------------------ host side
// 0 disables the CTA-pair (cta_group::2) variant; B200UNET_NO_CLUSTER=1 in the environment sets it (for A/B timing)
static int g_umma_cluster = getenv("B200UNET_NO_CLUSTER") ? 0 : 1;
static int g_umma_deep_a = getenv("B200UNET_NO_DEEP_A") ? 0 : 1;
static int g_umma_max_mb = getenv("B200UNET_MAX_MB") ? atoi(getenv("B200UNET_MAX_MB")) : 4;
// N tile: 128 by default.  BN = 256 halves the activation re-reads but needs all 512 TMEM columns for one MB = 2 tile, so
// the epilogue is not overlapped; once CTA pairs had halved the weight traffic, BN = 128 (double-buffered accumulators)
// measured 10..40 % faster on every layer with >= 256 output channels.  B200UNET_MAX_BN=256 restores the old choice.
static int g_umma_max_bn = getenv("B200UNET_MAX_BN") ? atoi(getenv("B200UNET_MAX_BN")) : 128;

struct Plan {
  int MB, BN, CL, P, TH, TW, halo;
  int tiles_x, tiles_y, n_ntiles;
  uint32_t a_stage_bytes, a_tx_bytes, smem_bytes;
  int nb_stages, na_stages;
};

static bool aligned_view(const b200_view& v) {
  // the epilogue addresses a destination with 32-bit element offsets inside one image
  const int64_t span = (int64_t)(v.h - 1) * v.stride_h + (int64_t)(v.w - 1) * v.stride_w + v.c;
  return v.c % 8 == 0 && reinterpret_cast<uintptr_t>(v.ptr) % 16 == 0 && reinterpret_cast<uintptr_t>(v.lo) % 16 == 0 &&
         v.stride_w % 8 == 0 && v.stride_h % 8 == 0 && (v.n == 1 || v.stride_n % 8 == 0) && span < (int64_t(1) << 31);
}

// 1x1 / transposed convolutions have no halo, so nothing separates the images of a batch: when every view involved is
// "row-contiguous across images" (stride_n == h * stride_h) the batch is folded into the row dimension and pixel tiles
// run across image boundaries.  With 28 x 28 images and 256-position tiles that is 100 tiles instead of 128 (the last
// tile of each image was 1 row of 9), and the kernel itself does not change.
static bool foldable(const b200_view& v) { return v.n == 1 || v.stride_n == (int64_t)v.h * v.stride_h; }
static b200_view fold_batch(const b200_view& v) {
  b200_view f = v;
  f.h = v.n * v.h;
  f.n = 1;
  f.stride_n = (int64_t)f.h * v.stride_h;
  return f;
}

static b200_view lo_plane(const b200_view& v) {  // the low-order plane of a split-tier view, as a plain view
  b200_view l = v;
  l.ptr = v.lo;
  l.lo = nullptr;
  return l;
}

static bool device_is_sm100() {
  static int cached = -1;
  if (cached < 0) cached = b200unet_device_ok();
  return cached == 1;
}

static int pick_bn(int cout_total, int ndst, int dst_c0) {
  const int cands[4] = {256, 128, 64, 32};
  for (int i = 0; i < 4; ++i) {
    const int bn = cands[i];
    if (bn > g_umma_max_bn) continue;
    // a tile may span destinations as long as no 64-column epilogue pass does
    if (cout_total % bn == 0 && (ndst == 1 || dst_c0 % bn == 0 || (bn >= 64 && dst_c0 % 64 == 0))) return bn;
  }
  // no exact tiling: a partial last N tile is fine (weight rows past Cout are zero-filled by TMA, the epilogue predicates
  // the columns) as long as no tile straddles two destinations
  for (int i = 2; i < 4; ++i) {
    const int bn = cands[i];
    if (ndst == 1 || dst_c0 % bn == 0) return bn;
  }
  return 0;
}

// Chooses tile geometry (MB M-blocks of 128 positions, pitch P, TH rows) minimising issued MMA rows.
// k_work = (64-channel K chunks) x (filter taps): MMA batches per tile
static bool make_plan(int Ho, int Wo, int n_img, int halo, int cout_total, int ndst, int dst_c0, int k_work, Plan* pl) {
  int bn = pick_bn(cout_total, ndst, dst_c0);
  if (bn == 0) return false;
  pl->halo = halo;
  double best = 1e30;
  bool found = false;
  for (int mb = 1; mb <= g_umma_max_mb; mb *= 2) {
    if (mb * bn > 512) continue;
    if (mb == 4 && bn > 64) continue;  // MB = 4 only where TMEM stays double-buffered
    for (int P = halo + 1; P <= 256 && P <= Wo + halo + 8; ++P) {
      const int TW = P - halo;
      int TH = (mb * 128) / P;
      if (TH < 1) break;
      if (TH > Ho) TH = Ho;
      if (TH + halo > 256) continue;
      const uint32_t rows = (uint32_t)max((TH + halo) * P, mb * 128 + halo * P + halo);
      const uint32_t a_stage = (rows * 128 + 1023) & ~1023u;
      if (kNA * a_stage + 2 * (uint32_t)bn * 128 + 1024 > kSmemBudget) continue;
      const long long tiles = (long long)((Wo + TW - 1) / TW) * ((Ho + TH - 1) / TH) * n_img;
      // cost: MMA rows issued, + fixed per-tile overhead, + mild preference for MB = 2 (halves weight traffic)
      double cost = (double)tiles * (mb * 128 + 24) * (mb == 1 ? 1.06 : (mb == 4 ? 0.94 : 1.0));
      if (cost < best) {
        best = cost;
        found = true;
        pl->MB = mb;
        pl->P = P;
        pl->TH = TH;
        pl->TW = TW;
        pl->a_stage_bytes = a_stage;
      }
    }
  }
  if (!found) return false;
  pl->BN = bn;
  // CTA pairs (cta_group::2) split the weight tile between two pixel tiles: needs enough pixel tiles for every pair
  // ... and enough MMA work per tile: a pair runs in lockstep (the leader's next tile needs BOTH epilogues done), which
  // costs more than the halved weight traffic saves when a tile is mostly epilogue (transposed conv with few input
  // channels: measured 0.29 -> 0.33 ms at 128 -> 64 channels)
  pl->CL = (g_umma_cluster && bn >= 32 && k_work >= 8 &&
            (long long)((Wo + pl->TW - 1) / pl->TW) * ((Ho + pl->TH - 1) / pl->TH) * n_img >= 4) ? 2 : 1;
  pl->tiles_x = (Wo + pl->TW - 1) / pl->TW;
  pl->tiles_y = (Ho + pl->TH - 1) / pl->TH;
  pl->n_ntiles = (cout_total + bn - 1) / bn;
  pl->a_tx_bytes = (uint32_t)(pl->TH + halo) * pl->P * 128;
  const uint32_t b_stage = (uint32_t)(bn / pl->CL) * 128;
  int na = kNA;
  if (halo == 0 && g_umma_deep_a)  // no taps to amortise an A stage over: deepen the A pipeline (B keeps >= 4 stages)
    while (na < kMaxNA && (uint32_t)(na + 1) * pl->a_stage_bytes + 4 * b_stage + 1024 <= kSmemBudget) ++na;
  pl->na_stages = na;
  int nb = (int)((kSmemBudget - 1024 - na * pl->a_stage_bytes) / b_stage);
  if (nb > kMaxNB) nb = kMaxNB;
  if (nb < 2) return false;
  pl->nb_stages = nb;
  pl->smem_bytes = na * pl->a_stage_bytes + nb * b_stage + kStageOutBytes + kBiasSmemFloats * 4 + 1024 /*align*/ +
                   512 /*barriers*/;
  return true;
}

static int make_a_map(CUtensorMap* m, const b200_view& v, int P, int rows) {
  uint64_t dims[4] = {(uint64_t)v.c, (uint64_t)v.w, (uint64_t)v.h, (uint64_t)v.n};
  uint64_t strides[3] = {(uint64_t)v.stride_w * 2, (uint64_t)v.stride_h * 2,
                         (uint64_t)(v.n > 1 ? v.stride_n : (int64_t)v.stride_h * v.h) * 2};
  uint32_t box[4] = {64, (uint32_t)P, (uint32_t)rows, 1};
  return make_tmap_bf16(m, v.ptr, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

static int make_b_map(CUtensorMap* m, const void* w, int rows, int cols, int bn) {
  uint64_t dims[2] = {(uint64_t)cols, (uint64_t)rows};
  uint64_t strides[1] = {(uint64_t)cols * 2};
  uint32_t box[2] = {64, (uint32_t)bn};
  return make_tmap_bf16(m, w, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

template <int MB, int BN, int CL>
static int launch_inst(const TileMaps& maps, const UmmaArgs& a, const Plan& pl, cudaStream_t st) {
  auto kern = umma_conv_kernel<MB, BN, CL>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(kSmemBudget + kStageOutBytes + kBiasSmemFloats * 4 + 2048));
    if (e != cudaSuccess) return fail((int)e, "umma_conv: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_done = true;
  }
  static int sms = 0;
  if (!sms) sms = b200unet_num_sms();
  const int pix = a.tiles_x * a.tiles_y * a.n_img;
  if (CL == 2) {
    const int pairs = (pix + 1) / 2 * a.n_ntiles;
    int clusters = pairs < sms / 2 ? pairs : sms / 2;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * clusters);
    cfg.blockDim = dim3(kUmmaThreads);
    cfg.dynamicSmemBytes = pl.smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, maps, a);
    if (e != cudaSuccess) return fail((int)e, "umma_conv: cluster launch: %s", cudaGetErrorString(e));
  } else {
    const int total = pix * a.n_ntiles;
    const int grid = total < sms ? total : sms;
    kern<<<grid, kUmmaThreads, pl.smem_bytes, st>>>(maps, a);
  }
  return check_launch("umma_conv");
}

static int launch_plan(const TileMaps& maps, UmmaArgs& a, const Plan& pl, cudaStream_t st) {
  a.P = pl.P;
  a.TH = pl.TH;
  a.TW = pl.TW;
  a.tiles_x = pl.tiles_x;
  a.tiles_y = pl.tiles_y;
  a.n_ntiles = pl.n_ntiles;
  a.a_stage_bytes = pl.a_stage_bytes;
  a.a_tx_bytes = pl.a_tx_bytes;
  a.nb_stages = pl.nb_stages;
  a.na_stages = pl.na_stages;
#define B200_INST(MBv, BNv)                                                                   \
  if (pl.MB == MBv && pl.BN == BNv)                                                           \
    return pl.CL == 2 ? launch_inst<MBv, BNv, 2>(maps, a, pl, st) : launch_inst<MBv, BNv, 1>(maps, a, pl, st);
  B200_INST(1, 32)
  B200_INST(1, 64)
  B200_INST(1, 128)
  B200_INST(1, 256)
  B200_INST(2, 32)
  B200_INST(2, 64)
  B200_INST(2, 128)
  B200_INST(2, 256)
  B200_INST(4, 32)
  B200_INST(4, 64)
#undef B200_INST
  return fail(-1, "umma_conv: no kernel instance for MB=%d BN=%d", pl.MB, pl.BN);
}

// ------------------------------------------------------------------ conv forward
static int fwd_k_work(const b200_conv_fwd_params* p) {
  int chunks = 0;
  for (int i = 0; i < p->num_src; ++i) chunks += (p->src[i].c + 63) / 64;
  return chunks * (p->dst.lo ? 3 : 1) * p->taps;
}

static bool folded_span_ok(const b200_view& v) {  // 32-bit element offsets inside the (folded) image
  return (int64_t)(v.n * (int64_t)v.h - 1) * v.stride_h + (int64_t)(v.w - 1) * v.stride_w + v.c < (int64_t(1) << 31);
}

// 1x1 convolution: fold the batch into the rows when every view allows it (see fold_batch)
static const b200_conv_fwd_params* fold_conv_fwd(const b200_conv_fwd_params* p, b200_conv_fwd_params* tmp) {
  if (p->taps != 1 || p->dst.n == 1 || !foldable(p->dst) || !folded_span_ok(p->dst)) return p;
  for (int i = 0; i < p->num_src; ++i)
    if (!foldable(p->src[i]) || !folded_span_ok(p->src[i])) return p;
  *tmp = *p;
  tmp->dst = fold_batch(p->dst);
  for (int i = 0; i < p->num_src; ++i) tmp->src[i] = fold_batch(p->src[i]);
  return tmp;
}
static const b200_conv_dgrad_params* fold_conv_dgrad(const b200_conv_dgrad_params* p, b200_conv_dgrad_params* tmp) {
  if (p->taps != 1 || p->dz.n == 1 || !foldable(p->dz) || !folded_span_ok(p->dz)) return p;
  for (int i = 0; i < p->num_dst; ++i)
    if (!foldable(p->dst[i]) || !folded_span_ok(p->dst[i])) return p;  // masks are laid out like their destinations
  *tmp = *p;
  tmp->dz = fold_batch(p->dz);
  for (int i = 0; i < p->num_dst; ++i) tmp->dst[i] = fold_batch(p->dst[i]);
  return tmp;
}
static const b200_convt_fwd_params* fold_convt_fwd(const b200_convt_fwd_params* p, b200_convt_fwd_params* tmp) {
  if (p->x.n == 1 || !foldable(p->x) || !foldable(p->y) || !folded_span_ok(p->x) || !folded_span_ok(p->y)) return p;
  *tmp = *p;
  tmp->x = fold_batch(p->x);
  tmp->y = fold_batch(p->y);  // rows 2i + a of the folded y: image n starts at row 2 * (n * h) = n * (2h)
  return tmp;
}
static const b200_convt_dgrad_params* fold_convt_dgrad(const b200_convt_dgrad_params* p, b200_convt_dgrad_params* tmp) {
  if (p->dx.n == 1 || !foldable(p->dx) || !foldable(p->dy) || !folded_span_ok(p->dx) || !folded_span_ok(p->dy)) return p;
  *tmp = *p;
  tmp->dx = fold_batch(p->dx);
  tmp->dy = fold_batch(p->dy);
  return tmp;
}

bool umma_conv_fwd_ok(const b200_conv_fwd_params* p) {
  if (!device_is_sm100()) return false;
  for (int i = 0; i < p->num_src; ++i)
    if (!aligned_view(p->src[i])) return false;
  if (!aligned_view(p->dst)) return false;
  if (p->bias && reinterpret_cast<uintptr_t>(p->bias) % 16 != 0) return false;
  Plan pl;
  return make_plan(p->dst.h, p->dst.w, p->dst.n, p->taps == 9 ? 2 : 0, p->dst.c, 1, p->dst.c, fwd_k_work(p), &pl);
}

int umma_conv_fwd(const b200_conv_fwd_params* p, cudaStream_t st) {
  if (!p->w_packed) return fail(-1, "conv_fwd (tcgen05): w_packed is required");
  b200_conv_fwd_params folded;
  p = fold_conv_fwd(p, &folded);
  const int halo = p->taps == 9 ? 2 : 0;
  Plan pl;
  if (!make_plan(p->dst.h, p->dst.w, p->dst.n, halo, p->dst.c, 1, p->dst.c, fwd_k_work(p), &pl))
    return fail(-1, "conv_fwd: no plan");
  TileMaps maps;
  UmmaArgs a{};
  // split tier: operand passes {hi(x) | lo(x) | hi(x)} against weights packed {hi(W) | hi(W) | lo(W)} along K
  const bool split = p->dst.lo != nullptr;
  const int passes = split ? 3 : 1;
  int kpad = 0;
  for (int i = 0; i < p->num_src; ++i) kpad += (p->src[i].c + 63) / 64 * 64;
  a.num_a = passes * p->num_src;
  for (int ps = 0; ps < passes; ++ps) {
    int koff = ps * kpad;
    for (int i = 0; i < p->num_src; ++i) {
      const int j = ps * p->num_src + i;
      int r = make_a_map(&maps.a[j], ps == 1 ? lo_plane(p->src[i]) : p->src[i], pl.P, pl.TH + halo);
      if (r) return fail(r, "conv_fwd: tensor map for src[%d] failed (%d)", i, r);
      a.a_c[j] = p->src[i].c;
      a.a_koff[j] = koff;
      koff += (p->src[i].c + 63) / 64 * 64;
    }
  }
  a.kpad = passes * kpad;
  a.dst_lo[0] = (bf16*)p->dst.lo;
  int r = make_b_map(&maps.b, p->w_packed, p->dst.c, p->taps * a.kpad, pl.BN / pl.CL);
  if (r) return fail(r, "conv_fwd: weight tensor map failed (%d)", r);
  a.taps = p->taps;
  a.kx = p->taps == 9 ? 3 : 1;
  a.pad = p->pad;
  a.n_img = p->dst.n;
  a.Ho = p->dst.h;
  a.Wo = p->dst.w;
  a.cout_total = p->dst.c;
  a.dst[0] = dview(p->dst);
  a.ndst = 1;
  a.dst_c0 = p->dst.c;
  a.bias = p->bias;
  a.relu = p->relu;
  return launch_plan(maps, a, pl, st);
}

// ------------------------------------------------------------------ conv backward-data
static int dgrad_cin(const b200_conv_dgrad_params* p) {
  int c = 0;
  for (int i = 0; i < p->num_dst; ++i) c += p->dst[i].c;
  return c;
}

bool umma_conv_dgrad_ok(const b200_conv_dgrad_params* p) {
  if (!device_is_sm100()) return false;
  if (!aligned_view(p->dz)) return false;
  for (int i = 0; i < p->num_dst; ++i) {
    if (!aligned_view(p->dst[i])) return false;
    if (reinterpret_cast<uintptr_t>(p->mask[i]) % 16 != 0) return false;
  }
  Plan pl;
  return make_plan(p->dst[0].h, p->dst[0].w, p->dst[0].n, p->taps == 9 ? 2 : 0, dgrad_cin(p), p->num_dst, p->dst[0].c,
                   (p->dz.c + 63) / 64 * p->taps, &pl);
}

int umma_conv_dgrad(const b200_conv_dgrad_params* p, cudaStream_t st) {
  if (!p->w_packed) return fail(-1, "conv_dgrad (tcgen05): w_packed is required");
  b200_conv_dgrad_params folded;
  p = fold_conv_dgrad(p, &folded);
  const int halo = p->taps == 9 ? 2 : 0;
  const int cin = dgrad_cin(p);
  Plan pl;
  if (!make_plan(p->dst[0].h, p->dst[0].w, p->dst[0].n, halo, cin, p->num_dst, p->dst[0].c, (p->dz.c + 63) / 64 * p->taps, &pl))
    return fail(-1, "conv_dgrad: no plan");
  TileMaps maps;
  UmmaArgs a{};
  a.num_a = 1;
  int r = make_a_map(&maps.a[0], p->dz, pl.P, pl.TH + halo);
  if (r) return fail(r, "conv_dgrad: tensor map for dz failed (%d)", r);
  a.a_c[0] = p->dz.c;
  a.a_koff[0] = 0;
  a.kpad = (p->dz.c + 63) / 64 * 64;
  r = make_b_map(&maps.b, p->w_packed, cin, p->taps * a.kpad, pl.BN / pl.CL);
  if (r) return fail(r, "conv_dgrad: weight tensor map failed (%d)", r);
  a.taps = p->taps;
  a.kx = p->taps == 9 ? 3 : 1;
  a.pad = halo - p->pad;  // full correlation with the flipped filter
  a.n_img = p->dst[0].n;
  a.Ho = p->dst[0].h;
  a.Wo = p->dst[0].w;
  a.cout_total = cin;
  a.ndst = p->num_dst;
  a.dst_c0 = p->dst[0].c;
  for (int i = 0; i < p->num_dst; ++i) {
    a.dst[i] = dview(p->dst[i]);
    a.mask[i] = (const bf16*)p->mask[i];
  }
  return launch_plan(maps, a, pl, st);
}

// ------------------------------------------------------------------ ConvTranspose2d forward / backward-data
static b200_view quadrant(const b200_view& big, int ab) {
  b200_view q = big;
  const int64_t shift = ((int64_t)(ab >> 1) * big.stride_h + (int64_t)(ab & 1) * big.stride_w) * 2;
  q.ptr = (char*)big.ptr + shift;
  if (big.lo) q.lo = (char*)big.lo + shift;
  q.h = big.h / 2;
  q.w = big.w / 2;
  q.stride_h = big.stride_h * 2;
  q.stride_w = big.stride_w * 2;
  return q;
}

bool umma_convt_fwd_ok(const b200_convt_fwd_params* p) {
  if (!device_is_sm100()) return false;
  if (!aligned_view(p->x) || !aligned_view(p->y)) return false;
  if (p->bias && reinterpret_cast<uintptr_t>(p->bias) % 16 != 0) return false;
  Plan pl;
  return make_plan(p->x.h, p->x.w, p->x.n, 0, 4 * p->y.c, 4, p->y.c, (p->x.c + 63) / 64 * (p->y.lo ? 3 : 1), &pl);
}

int umma_convt_fwd(const b200_convt_fwd_params* p, cudaStream_t st) {
  if (!p->w_packed) return fail(-1, "convt_fwd (tcgen05): w_packed is required");
  b200_convt_fwd_params folded;
  p = fold_convt_fwd(p, &folded);
  Plan pl;
  if (!make_plan(p->x.h, p->x.w, p->x.n, 0, 4 * p->y.c, 4, p->y.c, (p->x.c + 63) / 64 * (p->y.lo ? 3 : 1), &pl))
    return fail(-1, "convt_fwd: no plan");
  TileMaps maps;
  UmmaArgs a{};
  const bool split = p->y.lo != nullptr;  // see umma_conv_fwd
  const int cpad = (p->x.c + 63) / 64 * 64;
  a.num_a = split ? 3 : 1;
  int r = 0;
  for (int ps = 0; ps < a.num_a; ++ps) {
    r = make_a_map(&maps.a[ps], ps == 1 ? lo_plane(p->x) : p->x, pl.P, pl.TH);
    if (r) return fail(r, "convt_fwd: tensor map for x failed (%d)", r);
    a.a_c[ps] = p->x.c;
    a.a_koff[ps] = ps * cpad;
  }
  a.kpad = a.num_a * cpad;
  r = make_b_map(&maps.b, p->w_packed, 4 * p->y.c, a.kpad, pl.BN / pl.CL);
  if (r) return fail(r, "convt_fwd: weight tensor map failed (%d)", r);
  a.taps = 1;
  a.kx = 1;
  a.pad = 0;
  a.n_img = p->x.n;
  a.Ho = p->x.h;
  a.Wo = p->x.w;
  a.cout_total = 4 * p->y.c;
  a.ndst = 4;
  a.dst_c0 = p->y.c;
  for (int ab = 0; ab < 4; ++ab) {
    const b200_view q = quadrant(p->y, ab);
    a.dst[ab] = dview(q);
    a.dst_lo[ab] = (bf16*)q.lo;
  }
  a.bias = p->bias;
  return launch_plan(maps, a, pl, st);
}

bool umma_convt_dgrad_ok(const b200_convt_dgrad_params* p) {
  if (!device_is_sm100()) return false;
  if (!aligned_view(p->dx) || !aligned_view(p->dy)) return false;
  if (reinterpret_cast<uintptr_t>(p->mask) % 16 != 0) return false;
  Plan pl;
  return make_plan(p->dx.h, p->dx.w, p->dx.n, 0, p->dx.c, 1, p->dx.c, 4 * ((p->dy.c + 63) / 64), &pl);
}

int umma_convt_dgrad(const b200_convt_dgrad_params* p, cudaStream_t st) {
  if (!p->w_packed) return fail(-1, "convt_dgrad (tcgen05): w_packed is required");
  b200_convt_dgrad_params folded;
  p = fold_convt_dgrad(p, &folded);
  Plan pl;
  if (!make_plan(p->dx.h, p->dx.w, p->dx.n, 0, p->dx.c, 1, p->dx.c, 4 * ((p->dy.c + 63) / 64), &pl))
    return fail(-1, "convt_dgrad: no plan");
  TileMaps maps;
  UmmaArgs a{};
  a.num_a = 4;
  const int opad = (p->dy.c + 63) / 64 * 64;
  for (int ab = 0; ab < 4; ++ab) {
    int r = make_a_map(&maps.a[ab], quadrant(p->dy, ab), pl.P, pl.TH);
    if (r) return fail(r, "convt_dgrad: tensor map for dy quadrant %d failed (%d)", ab, r);
    a.a_c[ab] = p->dy.c;
    a.a_koff[ab] = ab * opad;
  }
  a.kpad = 4 * opad;
  int r = make_b_map(&maps.b, p->w_packed, p->dx.c, 4 * opad, pl.BN / pl.CL);
  if (r) return fail(r, "convt_dgrad: weight tensor map failed (%d)", r);
  a.taps = 1;
  a.kx = 1;
  a.pad = 0;
  a.n_img = p->dx.n;
  a.Ho = p->dx.h;
  a.Wo = p->dx.w;
  a.cout_total = p->dx.c;
  a.ndst = 1;
  a.dst_c0 = p->dx.c;
  a.dst[0] = dview(p->dx);
  a.mask[0] = (const bf16*)p->mask;
  return launch_plan(maps, a, pl, st);
}

// backward-weights on tensor cores: wgrad_umma.cu
}  // namespace b200
