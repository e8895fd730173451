// tcgen05 backward-weights kernel for the convolution family:
//   conv 3x3 / 1x1:   dW[o][c][r][s] = sum_{n,y,x} dz[n,y,x,o] * x[n,y+r-pad,x+s-pad,c]      (two concatenated sources)
//                     db[o]          = sum_{n,y,x} dz[n,y,x,o]
//   ConvTranspose2d:  dW[c][o][a][b] = sum_{n,i,j} x[n,i,j,c] * dy[n,2i+a,2j+b,o]
//
// GEMM view (per filter tap): D_tap[M = row-operand channels, N = gathered-operand channels] += R^T * G_tap, with the
// PIXELS as the contraction dimension.  NHWC makes both operands "MN-major" (channels contiguous, one 128-byte
// SWIZZLE_128B row per pixel), which tcgen05 consumes directly: no transpose pass anywhere.
//   R tile ("row operand": dz, or x for the transposed conv): TH x TW pixels stored with pitch P = TW + halo; the halo
//     columns and the rows past TH*P stay zero (zero-filled once, TMA never touches them), so they add nothing.
//   G tile ("gathered operand": x, or the four stride-2 sub-lattices of dy): (TH+halo) x P pixels, loaded once; tap
//     (r, s) is the same buffer read through a descriptor advanced by r*P + s rows — the shifted-descriptor trick of
//     the forward kernel, now on the K axis.
// Accumulators: one [128 x 64] fp32 block per tap in TMEM.  9 taps x 64 columns do not fit in 512, so
//   * Cout >= 128: a 3x3 runs as two tap groups (5 + 4) executed back to back by the same CTA;
//   * Cout <= 64 ("paired" mode): the second 64-lane half of the M = 128 MMA would be idle, so it is given the SAME dz
//     tile stored one pixel (one 128-byte shared-memory row) later: lanes 0..63 see filter column s+1, lanes 64..127
//     column s.  The N side is widened the same way: N = 192 = the x tile read at IMAGE-row shifts 0, 1, 2 — three
//     64-channel groups whose descriptor stride (LBO) is one tile pitch, P*128 bytes, i.e. three overlapping views of
//     the same buffer.  One MMA (full tensor rate) yields the 6 taps of filter columns 0 and 1, a second one at a
//     gathered shift of 2 pixels the column 2: 2 MMAs / 192 cycles per K step instead of 6 N = 64 MMAs / 288 cycles,
//     384 TMEM columns, one pass.
// Bias gradient: one extra N = 16 MMA per K step against a block of ones (column sums of dz on the tensor core).
// CTA pairs (cta2 mode, tcgen05 cta_group::2; >= 256 row channels and N = 128): two CTAs take the two 128-channel M
// blocks of the same (split, N block) and execute ONE MMA of M = 256 per tap and K step; each loads its own R tile and
// only ONE 64-channel half of the gathered tile (the MMA reads the other half from the peer).  The kernel is fed from L2
// (~43 B/cycle/SM chip-wide); this cuts its demand from ~47 to ~34 B/cycle per MMA cycle, and a stage shrinks enough
// for a third pipeline stage.
// Work item = (pixel split, 64-channel N block, 128-channel M block); fp32 partials [split][tap][M][N] (+ [split][M]
// for the bias); a second kernel reduces the splits in a fixed order and writes the state_dict layout, so results are
// run-to-run reproducible.
#include <cstdlib>

#include "chan_reduce.cuh"
#include "conv_impl.h"
#include "ptx.cuh"
#include "tmap.h"

namespace b200 {

constexpr int kWgThreads = 192;  // warp 0 producer, warp 1 MMA, warps 2..5 epilogue
constexpr int kWgMaxStages = 4;
constexpr uint32_t kWgSmemBudget = 200 * 1024;
constexpr uint32_t kWgOnesBytes = 2048;  // 16 K rows x 128 B of bf16 1.0
constexpr int kWgBiasCol = 448;          // TMEM column of the bias accumulator (taps use at most 6 * 64 = 384)

struct WgMaps {
  CUtensorMap r;     // row operand
  CUtensorMap g[4];  // gathered operand(s): conv = one per concat source, convT = one per (a, b) sub-lattice
};

struct WgArgs {
  int mode;    // 0 = conv (taps share one G tile), 1 = convT (one G tile per tap)
  int paired;  // conv 3x3 with <= 64 row channels: two taps per MMA (see header)
  int bias;    // accumulate column sums of the row operand (conv bias gradient)
  int m_total, n_total;
  int g_c[2];   // conv: channels per concat source
  int n_blks0;  // conv: number of N blocks of source 0
  int nbw;      // N block width: 64, or 128 (two 64-channel sub-tiles, one N = 128 MMA at the full tensor rate)
  int taps, kx, halo, pad;
  int P, TH, TW, kt_rows;
  int tiles_x, tiles_y, n_img;
  int m_blks, n_blks, tap_groups, splits;
  int m_pad, n_pad;
  float* ws;
  float* ws_bias;  // [splits][m_pad]
  uint32_t r_blk_bytes, g_tile_bytes, g_box_bytes, stage_bytes;
  uint32_t r_region_bytes;  // row-operand part of a stage: two 64-channel blocks, or (paired) one block + one more row
  int stages;
  int r_blocks;  // 64-channel blocks of the row operand actually loaded (1 or 2)
  int cta2;      // CTA pairs (see header): cluster of 2, rank = parity of the M block
  int m_units;   // M blocks (cta2: pairs of M blocks) enumerated by the work items
  int collector; // reuse the row-operand K slice across the MMAs of a K step (A collector buffer)
};

struct WgItem {
  int split, grp, nb, mb;
  int tap0, ntap;  // unpaired: taps of this group; paired: ntap = number of MMA groups (6)
  int tile0, tile1;
};

// The tap groups of a (split, N block, M block) are SEPARATE work items, the fastest-varying ones: neighbouring CTAs
// stream the same pixel tiles at the same time for different taps, so the tiles are fetched from HBM once and hit in L2
// for the other groups.  (Running the groups back to back in one CTA re-read every tile from HBM per group: the whole
// split does not stay in the 126 MB L2 — measured 2.8x the algorithmic DRAM traffic, 35 GB per training step.)
__device__ __forceinline__ WgItem decode_item(const WgArgs& a, int item, int rank) {
  WgItem w;
  const int grp = item % a.tap_groups;
  item /= a.tap_groups;
  w.mb = item % a.m_units;
  if (a.cta2) w.mb = 2 * w.mb + rank;
  item /= a.m_units;
  w.nb = item % a.n_blks;
  w.split = item / a.n_blks;
  w.grp = grp;
  if (a.paired) {
    w.tap0 = 0;
    w.ntap = 2;  // MMAs per K step (gathered shifts 0 and 2P)
  } else {
    const int per = (a.taps + a.tap_groups - 1) / a.tap_groups;
    w.tap0 = w.grp * per;
    w.ntap = min(per, a.taps - w.tap0);
  }
  const long long total = (long long)a.tiles_x * a.tiles_y * a.n_img;
  w.tile0 = (int)(total * w.split / a.splits);
  w.tile1 = (int)(total * (w.split + 1) / a.splits);
  return w;
}

// Issues the MMAs of one pipeline stage (one pixel tile): K step outermost, so that the NT filter taps [+ the bias column]
// of a K slice follow each other and share the slice of the row operand through the A collector buffer.  Everything about
// the sequence is a compile-time constant (the issuing lane has 48-96 cycles per MMA: a handful of uniform-datapath
// instructions, no branches).
template <int MODE, bool CTA2>
__device__ __forceinline__ void wg_mma(uint32_t d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  if (CTA2) {
    if (MODE == UMMA_A_FILL) umma_bf16_2cta_fill(d, a_desc, b_desc, idesc, acc);
    else if (MODE == UMMA_A_USE) umma_bf16_2cta_use(d, a_desc, b_desc, idesc, acc);
    else if (MODE == UMMA_A_LASTUSE) umma_bf16_2cta_lastuse(d, a_desc, b_desc, idesc, acc);
    else umma_bf16_2cta(d, a_desc, b_desc, idesc, acc);
  } else {
    if (MODE == UMMA_A_FILL) umma_bf16_fill(d, a_desc, b_desc, idesc, acc);
    else if (MODE == UMMA_A_USE) umma_bf16_use(d, a_desc, b_desc, idesc, acc);
    else if (MODE == UMMA_A_LASTUSE) umma_bf16_lastuse(d, a_desc, b_desc, idesc, acc);
    else umma_bf16(d, a_desc, b_desc, idesc, acc);
  }
}
template <int NT, int BIAS, bool CTA2>
__device__ __forceinline__ void wg_issue_tile(uint32_t tmem_base, int nbw, uint64_t r_desc, uint64_t g_desc0,
                                              const uint32_t (&goff)[6], uint64_t ones_desc, uint32_t idesc,
                                              uint32_t idesc_bias, int ksteps, uint32_t accum, bool coll_on) {
  constexpr int USERS = NT + BIAS;
  if (USERS > 1 && coll_on) {
#pragma unroll 2
    for (int k = 0; k < ksteps; ++k) {
      // one K step = 16 pixel rows = 2048 B = 128 in the descriptor's (address >> 4) field
      const uint64_t r_k = r_desc + (uint64_t)k * 128, g_k = g_desc0 + (uint64_t)k * 128;
      const uint32_t acc_k = k == 0 ? accum : 1u;
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        const uint32_t d = tmem_base + j * nbw;
        if (j == 0) wg_mma<UMMA_A_FILL, CTA2>(d, r_k, g_k + goff[j], idesc, acc_k);
        else if (j == USERS - 1) wg_mma<UMMA_A_LASTUSE, CTA2>(d, r_k, g_k + goff[j], idesc, acc_k);
        else wg_mma<UMMA_A_USE, CTA2>(d, r_k, g_k + goff[j], idesc, acc_k);
      }
      if (BIAS) wg_mma<UMMA_A_LASTUSE, CTA2>(tmem_base + kWgBiasCol, r_k, ones_desc, idesc_bias, acc_k);
    }
  } else {
    // tap outermost (consecutive MMAs accumulate into the same TMEM block), no operand reuse
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const uint64_t g_desc = g_desc0 + goff[j];
      const uint32_t d = tmem_base + j * nbw;
      wg_mma<UMMA_A_DISCARD, CTA2>(d, r_desc, g_desc, idesc, accum);
#pragma unroll 4
      for (int k = 1; k < ksteps; ++k)
        wg_mma<UMMA_A_DISCARD, CTA2>(d, r_desc + (uint64_t)k * 128, g_desc + (uint64_t)k * 128, idesc, 1u);
    }
    if (BIAS) {
      wg_mma<UMMA_A_DISCARD, CTA2>(tmem_base + kWgBiasCol, r_desc, ones_desc, idesc_bias, accum);
#pragma unroll 4
      for (int k = 1; k < ksteps; ++k)
        wg_mma<UMMA_A_DISCARD, CTA2>(tmem_base + kWgBiasCol, r_desc + (uint64_t)k * 128, ones_desc, idesc_bias, 1u);
    }
  }
}

// CTA2 is a template parameter: a kernel that contains cta_group::2 instructions can only be launched as a cluster.
template <bool CTA2>
__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_umma_kernel(const __grid_constant__ WgMaps maps, const __grid_constant__ WgArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ones = smem + (uint32_t)a.stages * a.stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ones + kWgOnesBytes);
  uint64_t* full = bars;
  uint64_t* empty = full + kWgMaxStages;
  uint64_t* t_full = empty + kWgMaxStages;
  uint64_t* t_empty = t_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_items = a.m_units * a.n_blks * a.splits * a.tap_groups;
  const int g_tiles = a.mode == 1 ? a.taps : 1;
  const int cl = CTA2 ? 2 : 1;
  const int rank = CTA2 ? (int)cluster_ctarank() : 0;
  const int item0 = CTA2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int item_step = CTA2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int g_sub = (a.nbw / 64) / cl;  // 64-channel sub-tiles of the gathered operand held by THIS CTA (per tap)

  // zero the whole pipeline once: halo columns / tail rows of the R tiles and the unused tail of the G tiles must
  // read as 0 (never NaN bit patterns); TMA only ever writes the box interiors afterwards.
  {
    uint4* p = reinterpret_cast<uint4*>(smem);
    const uint32_t n16 = (uint32_t)a.stages * a.stage_bytes / 16;
    for (uint32_t i = threadIdx.x; i < n16; i += kWgThreads) p[i] = make_uint4(0, 0, 0, 0);
    uint4* o = reinterpret_cast<uint4*>(ones);
    for (uint32_t i = threadIdx.x; i < kWgOnesBytes / 16; i += kWgThreads)
      o[i] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);  // bf16 1.0 pairs
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < kWgMaxStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(t_full, 1);
    mbar_init(t_empty, 4 * cl);  // cta2: the leader's barrier also collects the peer's epilogue warps
    fence_barrier_init();
  }
  if (warp == 1) {
    if (CTA2) tmem_alloc2<512>(tmem_slot); else tmem_alloc<512>(tmem_slot);
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  if (CTA2) cluster_sync_all();  // the peer's barriers must be initialised before anything arrives on them
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&maps.r);
      int stage = 0;
      uint32_t phase = 0;
      // paired: ONE copy of the dz tile, stored one row down; the second M half is the same buffer read one row later
      // (descriptor LBO = 128 B), not a second TMA copy: 27 % fewer bytes L2 -> SMEM per stage (the kernel is L2-fed)
      const int r_loads = a.paired ? 1 : a.r_blocks;
      for (int item = item0; item < total_items; item += item_step) {
          const WgItem w = decode_item(a, item, rank);
          // gathered-operand source of this N block
          int gsrc = 0, gc0 = w.nb * a.nbw;
          if (a.mode == 0 && w.nb >= a.n_blks0) {
            gsrc = 1;
            gc0 = (w.nb - a.n_blks0) * a.nbw;
          }
          for (int tile = w.tile0; tile < w.tile1; ++tile) {
            const int txi = tile % a.tiles_x;
            const int tyi = (tile / a.tiles_x) % a.tiles_y;
            const int n = tile / (a.tiles_x * a.tiles_y);
            const int y0 = tyi * a.TH, x0 = txi * a.TW;
            uint8_t* st = smem + stage * a.stage_bytes;
            mbar_wait(&empty[stage], phase ^ 1);
            const uint32_t tx_bytes = (uint32_t)r_loads * a.TH * a.TW * 128 + (uint32_t)g_tiles * g_sub * a.g_box_bytes;
            // cta2: the loads of both CTAs complete on the leader's barrier
            if (rank == 0) mbar_arrive_expect_tx(&full[stage], cl * tx_bytes);
            auto load = [&](const CUtensorMap* m, void* dst, int c0, int c1, int c2, int c3) {
              if (CTA2) tma_load_4d_2cta(m, &full[stage], dst, c0, c1, c2, c3);
              else tma_load_4d(m, &full[stage], dst, c0, c1, c2, c3);
            };
            // R tile: one TMA per tile row so that rows land with pitch P (halo columns stay zero).  Paired mode: the
            // tile is stored one row later (row q+1 = pixel q); M half 0 reads it from row 0, M half 1 from row 1.
            for (int rb = 0; rb < r_loads; ++rb) {
              const int ch = a.paired ? w.mb * 128 : w.mb * 128 + rb * 64;
              const int row_off = (a.paired && rb == 0) ? 1 : 0;
              for (int ty = 0; ty < a.TH; ++ty)
                load(&maps.r, st + rb * a.r_blk_bytes + (uint32_t)(ty * a.P + row_off) * 128, ch, x0, y0 + ty, n);
            }
            uint8_t* gt = st + a.r_region_bytes;
            // gathered operand: this CTA's 64-channel sub-tiles (cta2: sub-tile `rank` of the two)
            if (a.mode == 0) {
              for (int h = 0; h < g_sub; ++h)
                load(&maps.g[gsrc], gt + h * a.g_tile_bytes, gc0 + (rank * g_sub + h) * 64, x0 - a.pad, y0 - a.pad, n);
            } else {
              for (int t = 0; t < a.taps; ++t)
                for (int h = 0; h < g_sub; ++h)
                  load(&maps.g[t], gt + (t * g_sub + h) * a.g_tile_bytes, gc0 + (rank * g_sub + h) * 64, x0, y0, n);
            }
            if (++stage == a.stages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
    }
  } else if (warp == 1) {
    // The whole warp runs the control flow so that every address / descriptor is warp-uniform (lives in uniform
    // registers, no per-MMA R2UR/ELECT sequences); one elected lane issues the tcgen05 instructions.
    const int mma_n = a.paired ? 192 : a.nbw;                        // paired: three row-shifted views of the x tile
    const uint32_t idesc = umma_idesc_bf16(128 * cl, mma_n, 1, 1);   // both operands MN-major
    const uint32_t idesc_bias = umma_idesc_bf16(128 * cl, 16, 1, 1);
    // paired: the second 64-lane half of M starts one 128-byte row after the first, inside the same buffer
    const uint64_t hi_r = umma_desc_hi_sw128(a.paired ? 128u : a.r_blk_bytes, 1024);
    // LBO = distance of the next 64-channel group: the second sub-tile, or (paired) the same tile one image row later
    const uint64_t hi_g = umma_desc_hi_sw128(a.paired ? (uint32_t)a.P * 128u : a.g_tile_bytes, 1024);
    const uint64_t ones_desc = umma_desc(umma_desc_hi_sw128(a.r_blk_bytes, 1024), smem_u32(ones));
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    const int ksteps = a.kt_rows / 16;
    const int nbw = mma_n;
    const uint32_t stage_bytes = a.stage_bytes, r_region_bytes = a.r_region_bytes;
    const int n_stages = a.stages;
    const bool coll_on = a.collector != 0;
    // cta2: the leader issues for the pair; the peer's warp only took part in the TMEM allocation
    for (int item = (CTA2 && rank != 0) ? total_items : item0; item < total_items; item += item_step, ++it) {
        const WgItem w = decode_item(a, item, rank);
        const bool do_bias = a.bias && w.nb == 0 && w.grp == 0;
        // per-item table of gathered-operand start offsets (in descriptor units of 16 B): no division / constant
        // loads inside the issue loop — with N = 64 MMAs (48 cycles each) the issuing lane is otherwise the bottleneck
        uint32_t goff[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) {
          uint32_t o = 0;
          if (j < w.ntap) {
            if (a.paired) {
              o = (uint32_t)(2 * j) * 8;  // gathered shift of 0 / 2 pixels; 128 B = 8 descriptor units
            } else if (a.mode == 0) {
              const int tap = w.tap0 + j;
              const int r = tap / a.kx, sx = tap - r * a.kx;
              o = (uint32_t)(r * a.P + sx) * 8;
            } else {
              o = (uint32_t)(w.tap0 + j) * g_sub * (a.g_tile_bytes >> 4);
            }
          }
          goff[j] = o;
        }
        const int ntap = w.ntap;
        const int issue_variant = ntap * 2 + (do_bias ? 1 : 0);
        mbar_wait(t_empty, (it & 1) ^ 1);
        tc_fence_after_sync();
        uint32_t accum = 0;
        for (int tile = w.tile0; tile < w.tile1; ++tile) {
          mbar_wait(&full[stage], phase);
          tc_fence_after_sync();
          const uint32_t r_base = smem_u32(smem + stage * stage_bytes);
          const uint64_t r_desc = umma_desc(hi_r, r_base);
          const uint64_t g_desc0 = umma_desc(hi_g, r_base + r_region_bytes);
          if (elect_one()) {
            switch (issue_variant) {
#define B200_WG_CASE(NT, BIAS) case NT * 2 + BIAS: wg_issue_tile<NT, BIAS, CTA2>(tmem_base, nbw, r_desc, g_desc0, goff, ones_desc, idesc, idesc_bias, ksteps, accum, coll_on); break;
              B200_WG_CASE(1, 0) B200_WG_CASE(1, 1) B200_WG_CASE(2, 0) B200_WG_CASE(2, 1) B200_WG_CASE(3, 0) B200_WG_CASE(3, 1)
              B200_WG_CASE(4, 0) B200_WG_CASE(4, 1) B200_WG_CASE(5, 0) B200_WG_CASE(5, 1) B200_WG_CASE(6, 0) B200_WG_CASE(6, 1)
#undef B200_WG_CASE
              default: break;
            }
            if (CTA2) umma_commit_2cta(&empty[stage], (uint16_t)0x3); else umma_commit(&empty[stage]);
          }
          __syncwarp();
          accum = 1;
          if (++stage == n_stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (elect_one()) {
          if (CTA2) umma_commit_2cta(t_full, (uint16_t)0x3); else umma_commit(t_full);
        }
        __syncwarp();
      }
  } else {
    // epilogue: TMEM -> fp32 partials ws[split][tap][m][n]
    const int quarter = warp & 3;
    int it = 0;
    for (int item = item0; item < total_items; item += item_step, ++it) {
        const WgItem w = decode_item(a, item, rank);
        mbar_wait(t_full, it & 1);
        tc_fence_after_sync();
        const int row = quarter * 32 + lane;           // TMEM lane = accumulator row
        const bool empty_item = w.tile1 <= w.tile0;    // no MMA was issued: the accumulator holds stale data
        const int nacc = a.paired ? 6 : w.ntap;  // [128 x 64|128] accumulator blocks to drain
        for (int j = 0; j < nacc; ++j) {
          int tap, m, tcol;
          bool live = true;
          if (a.paired) {
            // MMA jm (gathered shift of 2*jm pixels), column group g = filter row g; lanes 0..63 hold filter column
            // 2*jm + 1 (the dz copy stored one pixel later), lanes 64..127 filter column 2*jm
            const int jm = j / 3, g = j - jm * 3;
            const int fc = 2 * jm + (row < 64 ? 1 : 0);
            live = fc <= 2;
            tap = g * 3 + fc;
            m = row & 63;
            tcol = jm * 192 + g * 64;
          } else {
            tap = w.tap0 + j;
            m = w.mb * 128 + row;
            tcol = j * a.nbw;
          }
          float* out = a.ws + (((long long)w.split * a.taps + tap) * a.m_pad + m) * a.n_pad + w.nb * a.nbw;
#pragma unroll 1
          for (int col0 = 0; col0 < a.nbw; col0 += 32) {
            uint32_t v[32];
            tmem_ld_32x32(tmem_base + tcol + col0 + (uint32_t(quarter * 32) << 16), v);
            tmem_ld_wait();
            if (live) {
#pragma unroll
              for (int q4 = 0; q4 < 8; ++q4) {
                float4 f;
                f.x = empty_item ? 0.f : __uint_as_float(v[q4 * 4 + 0]);
                f.y = empty_item ? 0.f : __uint_as_float(v[q4 * 4 + 1]);
                f.z = empty_item ? 0.f : __uint_as_float(v[q4 * 4 + 2]);
                f.w = empty_item ? 0.f : __uint_as_float(v[q4 * 4 + 3]);
                *reinterpret_cast<float4*>(out + col0 + q4 * 4) = f;
              }
            }
          }
        }
        if (a.bias && w.nb == 0 && w.grp == 0) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + kWgBiasCol + (uint32_t(quarter * 32) << 16), v);  // columns 0..15 are the sums
          tmem_ld_wait();
          const bool live = !a.paired || row < 64;
          if (live) a.ws_bias[(long long)w.split * a.m_pad + w.mb * 128 + row] = empty_item ? 0.f : __uint_as_float(v[0]);
        }
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) {
          if (CTA2) mbar_arrive_cluster(t_empty, 0); else mbar_arrive(t_empty);
        }
      }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (CTA2) cluster_sync_all();  // the pair's MMAs read this CTA's shared memory and arrive on its barriers
  if (warp == 1) {
    if (CTA2) tmem_dealloc2<512>(tmem_base); else tmem_dealloc<512>(tmem_base);
  }
}

// dw[(m * n_total + n) * taps + tap] = sum_s ws[s][tap][m][npad(n)].  Block = one m, 128 consecutive n, all taps (one
// warp per tap, one float4 = 4 consecutive n per lane): 512-byte coalesced reads with the split loop unrolled so that
// eight loads are in flight per thread, and the transposed result goes through shared memory so that the store is
// contiguous.  (The first version — 32 n per block, scalar loads, one dependent load per iteration — ran at a third
// of the HBM rate and cost ~1 ms per training step.)  Summation order over the splits is fixed: reproducible.
constexpr int kRedN = 128;
__global__ void __launch_bounds__(32 * 9)
wgrad_umma_reduce_kernel(const float* __restrict__ ws, int splits, int taps, int m_total, int n_total, int m_pad,
                         int n_pad, int c_src0, int n_blks0, float* __restrict__ dw,
                         const float* __restrict__ ws_bias, float* __restrict__ db) {
  __shared__ float tile[kRedN * 9];
  const int m = blockIdx.y;
  const int n0 = blockIdx.x * kRedN;
  const int lane = threadIdx.x & 31, t = threadIdx.x >> 5;
  const int n = n0 + 4 * lane;  // channel counts are multiples of 8: the four n of a lane belong to the same source
  if (t < taps && n < n_total) {
    const int np = n < c_src0 ? n : n_blks0 + (n - c_src0);  // n_blks0 = padded width of source 0 here
    const long long stride = (long long)taps * m_pad * n_pad;
    const float* p = ws + ((long long)t * m_pad + m) * n_pad + np;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    int i = 0;
    for (; i + 8 <= splits; i += 8) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = *reinterpret_cast<const float4*>(p + (i + u) * stride);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        s.x += v[u].x;
        s.y += v[u].y;
        s.z += v[u].z;
        s.w += v[u].w;
      }
    }
    for (; i < splits; ++i) {
      const float4 v = *reinterpret_cast<const float4*>(p + i * stride);
      s.x += v.x;
      s.y += v.y;
      s.z += v.z;
      s.w += v.w;
    }
    const int nl = 4 * lane;
    tile[(nl + 0) * taps + t] = s.x;
    tile[(nl + 1) * taps + t] = s.y;
    tile[(nl + 2) * taps + t] = s.z;
    tile[(nl + 3) * taps + t] = s.w;
  }
  __syncthreads();
  const int cnt = min(kRedN, n_total - n0) * taps;
  for (int j = threadIdx.x; j < cnt; j += blockDim.x) dw[((long long)m * n_total + n0) * taps + j] = tile[j];
  if (db && blockIdx.x == 0 && t == 0) {  // bias gradient: lanes stride over the splits, fixed butterfly
    float s = 0.f;
    for (int i = lane; i < splits; i += 32) s += ws_bias[(long long)i * m_pad + m];
    s = warp_sum(s);
    if (lane == 0) db[m] = s;
  }
}

// ------------------------------------------------------------------ host side
// N = 128 variant and pipeline depth, measured on B200 at batch 32 (ms: 256->256 @136^2 / 128->128 @280^2 / 1024->512 @54^2):
//   N = 64,  >= 3 stages : 0.705 / 0.885 / 1.156      N = 128, >= 3 stages : 0.767 / 0.895 / 1.297
//   N = 64,  >= 2 stages : 0.702 / 0.863 / 1.033      N = 128, >= 2 stages : 0.688 / 0.843 / 0.898   <- default
// (three passes need big K tiles to pay off, and big tiles only fit twice into shared memory).
// B200UNET_WGRAD_N64=1 / B200UNET_WGRAD_MIN_STAGES=k override for experiments.
static int g_wgrad_n128 = getenv("B200UNET_WGRAD_N64") ? 0 : 1;
static int g_wgrad_min_stages = getenv("B200UNET_WGRAD_MIN_STAGES") ? atoi(getenv("B200UNET_WGRAD_MIN_STAGES")) : 2;
static int g_wgrad_cta2 = getenv("B200UNET_WGRAD_NO_PAIRS") ? 0 : 1;
static int g_wgrad_collector = getenv("B200UNET_WGRAD_NO_COLLECTOR") ? 0 : 1;
// cost-model cycles per TMA instruction of a stage (the row operand arrives one tile row per instruction): measured on
// the whole training step 0 -> 31.1 ms, 30 -> 30.7, 70 -> 30.2, 100/150 -> +0.3; without it the planner picked 13 x 19
// pixel tiles for the first layer (19 tiny row loads per stage) and 0.46 ms where 114 x 2 tiles take 0.38
static double g_wgrad_tma_cost = getenv("B200UNET_WGRAD_TMA_COST") ? atof(getenv("B200UNET_WGRAD_TMA_COST")) : 70.0;

struct WgPlan {
  WgArgs a;
  size_t ws_bytes, ws_bias_bytes;
  uint32_t smem_bytes;
  int grid;
};

static bool wg_aligned(const b200_view& v) {
  return v.c % 8 == 0 && reinterpret_cast<uintptr_t>(v.ptr) % 16 == 0 && v.stride_w % 8 == 0 && v.stride_h % 8 == 0 &&
         (v.n == 1 || v.stride_n % 8 == 0);
}

static bool wg_device_ok() {
  static int cached = -1;
  if (cached < 0) cached = (getenv("B200UNET_PLAN_ONLY") != nullptr) ? 1 : b200unet_device_ok();  // plan dumps without a GPU
  return cached == 1;
}

static int wg_sms() {
  static int sms = 0;
  if (!sms) sms = b200unet_num_sms();
  return sms > 0 ? sms : kNumSMsB200;
}

// row operand extent Ho x Wo x n_img; M = row channels, N = gathered channels (per source for conv)
static bool wg_make_plan(int mode, int Ho, int Wo, int n_img, int m_total, int n_total, int c_src0, int taps, int pad,
                         bool want_bias, WgPlan* pl) {
  WgArgs& a = pl->a;
  a = WgArgs{};
  a.mode = mode;
  a.m_total = m_total;
  a.n_total = n_total;
  a.taps = taps;
  a.kx = taps == 9 ? 3 : (taps == 4 ? 2 : 1);
  a.halo = (mode == 0 && taps == 9) ? 2 : 0;
  a.pad = pad;
  a.n_img = n_img;
  a.g_c[0] = c_src0;
  a.g_c[1] = n_total - c_src0;
  a.paired = (mode == 0 && taps == 9 && m_total <= 64) ? 1 : 0;
  // N = 128 blocks (MMA at the full rate instead of the shared-memory-bound 2/3 of N = 64) when every source is a
  // multiple of 128 channels; the 3x3 then runs as three tap groups (one filter row each: 3 x 128 TMEM columns)
  a.nbw = (g_wgrad_n128 && mode == 0 && taps == 9 && !a.paired && c_src0 % 128 == 0 && (n_total - c_src0) % 128 == 0) ? 128 : 64;
  // ConvTranspose: every dy element feeds exactly one tap, so operand reuse is low (43 MAC per loaded byte with N = 64)
  // and the kernel is fed from L2: N = 128 (4 taps x 128 = all 512 TMEM columns) more than doubles the reuse
  if (mode == 1 && n_total % 128 == 0) a.nbw = 128;
  a.n_blks0 = (c_src0 + a.nbw - 1) / a.nbw;
  a.n_blks = a.n_blks0 + (mode == 0 ? (a.g_c[1] + a.nbw - 1) / a.nbw : 0);
  a.m_blks = (m_total + 127) / 128;
  a.r_blocks = m_total > 64 ? 2 : 1;
  a.bias = (want_bias && mode == 0) ? 1 : 0;
  a.tap_groups = (taps == 9 && !a.paired) ? (a.nbw == 128 ? 3 : 2) : 1;
  a.m_pad = a.m_blks * 128;
  a.n_pad = a.n_blks * a.nbw;
  // CTA pairs (cta_group::2): two M blocks per work item, each CTA loads one of the two 64-channel gathered sub-tiles
  a.cta2 = (g_wgrad_cta2 && a.nbw == 128 && !a.paired && a.m_blks % 2 == 0 && m_total % 128 == 0) ? 1 : 0;
  a.m_units = a.cta2 ? a.m_blks / 2 : a.m_blks;
  // A-collector reuse across the taps of a K step: measured (B200, batch 32) 0.51 -> 0.40 ms on the 64-wide N blocks with 4-5
  // taps per group, 0.85 -> 0.83 ms in paired mode, 1-3 % SLOWER with N = 128 blocks (there the tap-outermost order, whose
  // consecutive MMAs accumulate into the same TMEM block, wins)
  a.collector = (g_wgrad_collector && a.nbw == 64) ? 1 : 0;
  const int cl = a.cta2 ? 2 : 1;
  const int g_tiles = (mode == 1 ? taps : 1) * (a.nbw / 64) / cl;  // gathered sub-tiles per CTA and stage
  const int r_loads = a.paired ? 1 : a.r_blocks;
  const double mma_groups = a.paired ? 2.0 : (taps == 9 ? (a.nbw == 128 ? 3.0 : 4.5) : (double)taps);
  const double mma_cycles = a.paired ? 96.0 : (a.nbw == 128 ? 64.0 : 48.0);
  const double passes = a.tap_groups;
  // tile geometry: kt_rows (multiple of 16) K rows per tile
  double best = 1e30;
  bool found = false;
  for (int kt = 32; kt <= 256; kt += 16) {
    for (int P = a.halo + 1; P <= 256 && P <= Wo + a.halo + 8; ++P) {
      const int TW = P - a.halo;
      int TH = (kt - a.paired) / P;
      if (TH < 1) break;
      if (TH > Ho) TH = Ho;
      const uint32_t r_blk = (uint32_t)kt * 128;
      const uint32_t g_rows = (uint32_t)max((TH + a.halo) * P, kt + a.halo * P + a.halo);
      const uint32_t g_tile = (g_rows * 128 + 1023) & ~1023u;
      const uint32_t r_region = a.paired ? r_blk + 1024 : 2 * r_blk;  // paired: rows 0 .. kt (the second M half starts at row 1)
      const uint32_t stage = r_region + g_tiles * g_tile;
      if ((uint32_t)g_wgrad_min_stages * stage + kWgOnesBytes + 1024 > kWgSmemBudget) continue;  // >= 3 stages: fed from L2, latency must hide
      const long long tiles = (long long)((Wo + TW - 1) / TW) * ((Ho + TH - 1) / TH) * n_img;
      // per tile and pass: MMA time ~ MMA groups * K steps * ~48 cycles (shared-memory-bound N = 64 MMA), load time ~
      // bytes moved L2 -> SMEM at ~32 B/cycle/SM; whichever is larger, plus a fixed per-tile cost
      const double t_mma = mma_groups * (kt / 16) * mma_cycles;
      const double t_load = ((double)r_loads * TH * TW * 128 + (double)g_tiles * (TH + a.halo) * P * 128) / 32.0;
      const double t_ops = g_wgrad_tma_cost * ((double)r_loads * TH + g_tiles);  // per-TMA-instruction cost (row loads)
      const double cost = (double)tiles * passes * ((t_mma > t_load ? t_mma : t_load) + 300.0 + t_ops);
      if (cost < best) {
        best = cost;
        found = true;
        a.kt_rows = kt;
        a.P = P;
        a.TH = TH;
        a.TW = TW;
        a.r_blk_bytes = r_blk;
        a.r_region_bytes = r_region;
        a.g_tile_bytes = g_tile;
        a.g_box_bytes = (uint32_t)(TH + a.halo) * P * 128;
        a.stage_bytes = stage;
      }
    }
  }
  if (!found) return false;
  a.tiles_x = (Wo + a.TW - 1) / a.TW;
  a.tiles_y = (Ho + a.TH - 1) / a.TH;
  int stages = (int)((kWgSmemBudget - 1024 - kWgOnesBytes) / a.stage_bytes);
  if (stages > kWgMaxStages) stages = kWgMaxStages;
  a.stages = stages;
  const long long tiles = (long long)a.tiles_x * a.tiles_y * n_img;
  const long long base_items = (long long)a.m_units * a.n_blks * a.tap_groups;
  const int slots = wg_sms() / cl;  // work items resident at a time (cta2: CTA pairs)
  // pixel splits: make the number of work items fill whole waves of the persistent grid
  long long splits = 1;
  double best_eff = 0.0;
  for (int k = 1; k <= 3; ++k) {
    long long sp = (long long)slots * k / base_items;
    if (sp < 1) sp = 1;
    if (sp > tiles) sp = tiles;
    const long long items = base_items * sp;
    const long long waves = (items + slots - 1) / slots;
    const double eff = (double)items / (double)(waves * slots);
    if (eff > best_eff + 0.03) {
      best_eff = eff;
      splits = sp;
    }
  }
  // keep the fp32 partials bounded (<= 256 MiB)
  const long long per_split = (long long)taps * a.m_pad * a.n_pad * 4;
  while (splits > 1 && splits * per_split > (256LL << 20)) --splits;
  a.splits = (int)splits;
  pl->ws_bytes = (size_t)splits * per_split;
  pl->ws_bias_bytes = a.bias ? (size_t)splits * a.m_pad * 4 : 0;
  pl->smem_bytes = (uint32_t)a.stages * a.stage_bytes + kWgOnesBytes + 1024 + 256;
  const long long items = base_items * splits;
  pl->grid = cl * (int)(items < slots ? items : slots);
  static const bool dbg = getenv("B200UNET_DEBUG_PLAN") != nullptr;
  if (dbg)
    fprintf(stderr, "[wgrad plan] mode %d %dx%dx%d m %d n %d taps %d: paired %d cta2 %d nbw %d groups %d kt %d P %d TH %d TW %d "
            "stages %d stage_bytes %u splits %d items %lld grid %d tiles %lld\n", mode, n_img, Ho, Wo, m_total, n_total, taps,
            a.paired, a.cta2, a.nbw, a.tap_groups, a.kt_rows, a.P, a.TH, a.TW, a.stages, a.stage_bytes, a.splits, items,
            pl->grid, tiles);
  return true;
}

static int wg_map(CUtensorMap* m, const b200_view& v, int bw, int bh) {
  uint64_t dims[4] = {(uint64_t)v.c, (uint64_t)v.w, (uint64_t)v.h, (uint64_t)v.n};
  uint64_t strides[3] = {(uint64_t)v.stride_w * 2, (uint64_t)v.stride_h * 2,
                         (uint64_t)(v.n > 1 ? v.stride_n : (int64_t)v.stride_h * v.h) * 2};
  uint32_t box[4] = {64, (uint32_t)bw, (uint32_t)bh, 1};
  return make_tmap_bf16(m, v.ptr, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

static int wg_launch(const WgMaps& maps, WgPlan& pl, float* dw, float* db, int c_src0, void* ws, size_t ws_bytes,
                     cudaStream_t st) {
  const size_t need = pl.ws_bytes + pl.ws_bias_bytes;
  if (!ws || ws_bytes < need) return fail(-1, "wgrad (tcgen05): workspace too small (%zu < %zu)", ws_bytes, need);
  pl.a.ws = (float*)ws;
  pl.a.ws_bias = (float*)((char*)ws + pl.ws_bytes);
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_umma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)kWgSmemBudget + 4096);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(wgrad_umma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)kWgSmemBudget + 4096);
    if (e != cudaSuccess) return fail((int)e, "wgrad: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_done = true;
  }
  if (pl.a.cta2) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(pl.grid);
    cfg.blockDim = dim3(kWgThreads);
    cfg.dynamicSmemBytes = pl.smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, wgrad_umma_kernel<true>, maps, pl.a);
    if (e != cudaSuccess) return fail((int)e, "wgrad: cluster launch: %s", cudaGetErrorString(e));
  } else {
    wgrad_umma_kernel<false><<<pl.grid, kWgThreads, pl.smem_bytes, st>>>(maps, pl.a);
  }
  int r = check_launch("wgrad_umma");
  if (r) return r;
  dim3 grid((unsigned)((pl.a.n_total + kRedN - 1) / kRedN), (unsigned)pl.a.m_total);
  wgrad_umma_reduce_kernel<<<grid, 32 * pl.a.taps, 0, st>>>(pl.a.ws, pl.a.splits, pl.a.taps, pl.a.m_total, pl.a.n_total,
                                                   pl.a.m_pad, pl.a.n_pad, c_src0, pl.a.n_blks0 * pl.a.nbw, dw,
                                                   pl.a.bias ? pl.a.ws_bias : nullptr, pl.a.bias ? db : nullptr);
  return check_launch("wgrad_umma_reduce");
}

static int conv_cin(const b200_conv_wgrad_params* p) {
  int c = 0;
  for (int i = 0; i < p->num_src; ++i) c += p->src[i].c;
  return c;
}

static bool conv_plan(const b200_conv_wgrad_params* p, WgPlan* pl) {
  return wg_make_plan(0, p->dz.h, p->dz.w, p->dz.n, p->dz.c, conv_cin(p), p->src[0].c, p->taps, p->pad,
                      p->db_f32 != nullptr, pl);
}

bool umma_conv_wgrad_ok(const b200_conv_wgrad_params* p) {
  if (!wg_device_ok()) return false;
  if (!wg_aligned(p->dz)) return false;
  for (int i = 0; i < p->num_src; ++i)
    if (!wg_aligned(p->src[i])) return false;
  if (p->num_src == 2 && p->src[0].c % 8 != 0) return false;
  WgPlan pl;
  return conv_plan(p, &pl);
}

size_t umma_conv_wgrad_workspace(const b200_conv_wgrad_params* p) {
  WgPlan pl;
  if (!conv_plan(p, &pl)) return 0;
  return pl.ws_bytes + pl.ws_bias_bytes;
}

int umma_conv_wgrad(const b200_conv_wgrad_params* p, void* ws, size_t ws_bytes, cudaStream_t st) {
  WgPlan pl;
  if (!conv_plan(p, &pl)) return fail(-1, "conv_wgrad: no plan");
  WgMaps maps;
  int r = wg_map(&maps.r, p->dz, pl.a.TW, 1);
  if (r) return fail(r, "conv_wgrad: tensor map for dz failed (%d)", r);
  for (int i = 0; i < p->num_src; ++i) {
    r = wg_map(&maps.g[i], p->src[i], pl.a.P, pl.a.TH + pl.a.halo);
    if (r) return fail(r, "conv_wgrad: tensor map for src[%d] failed (%d)", i, r);
  }
  return wg_launch(maps, pl, p->dw_f32, p->db_f32, p->src[0].c, ws, ws_bytes, st);
}

// ConvTranspose2d: row operand = x (M = cin), gathered operands = the four stride-2 sub-lattices of dy (N = cout)
static b200_view wg_quadrant(const b200_view& big, int ab) {
  b200_view q = big;
  q.ptr = (char*)big.ptr + ((int64_t)(ab >> 1) * big.stride_h + (int64_t)(ab & 1) * big.stride_w) * 2;
  q.h = big.h / 2;
  q.w = big.w / 2;
  q.stride_h = big.stride_h * 2;
  q.stride_w = big.stride_w * 2;
  return q;
}

bool umma_convt_wgrad_ok(const b200_convt_wgrad_params* p) {
  if (!wg_device_ok()) return false;
  if (!wg_aligned(p->x) || !wg_aligned(p->dy)) return false;
  WgPlan pl;
  return wg_make_plan(1, p->x.h, p->x.w, p->x.n, p->x.c, p->dy.c, p->dy.c, 4, 0, false, &pl);
}

size_t umma_convt_wgrad_workspace(const b200_convt_wgrad_params* p) {
  WgPlan pl;
  if (!wg_make_plan(1, p->x.h, p->x.w, p->x.n, p->x.c, p->dy.c, p->dy.c, 4, 0, false, &pl)) return 0;
  return pl.ws_bytes + reduce_workspace_bytes(p->dy.c, 1);
}

int umma_convt_wgrad(const b200_convt_wgrad_params* p, void* ws, size_t ws_bytes, cudaStream_t st) {
  WgPlan pl;
  if (!wg_make_plan(1, p->x.h, p->x.w, p->x.n, p->x.c, p->dy.c, p->dy.c, 4, 0, false, &pl))
    return fail(-1, "convt_wgrad: no plan");
  WgMaps maps;
  int r = wg_map(&maps.r, p->x, pl.a.TW, 1);
  if (r) return fail(r, "convt_wgrad: tensor map for x failed (%d)", r);
  for (int ab = 0; ab < 4; ++ab) {
    r = wg_map(&maps.g[ab], wg_quadrant(p->dy, ab), pl.a.P, pl.a.TH);
    if (r) return fail(r, "convt_wgrad: tensor map for dy sub-lattice %d failed (%d)", ab, r);
  }
  const size_t need = pl.ws_bytes + (p->db_f32 ? reduce_workspace_bytes(p->dy.c, 1) : 0);
  if (ws_bytes < need) return fail(-1, "convt_wgrad (tcgen05): workspace too small (%zu < %zu)", ws_bytes, need);
  r = wg_launch(maps, pl, p->dw_f32, nullptr, p->dy.c, ws, ws_bytes, st);
  if (r) return r;
  if (p->db_f32) return bias_grad(p->dy, p->db_f32, (char*)ws + pl.ws_bytes, st);
  return 0;
}

}  // namespace b200
