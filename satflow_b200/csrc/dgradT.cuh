// Transposed data-gradient GEMM:  D^T[c][pixel] = sum_k Wd[c][k] * dz_shift[pixel][k]
//
// The pixel-major dgrad of convgemm.cuh has only N = Cx + hid = 128 output columns, and a tcgen05 SS-mode
// MMA with M = 128, N = 128 reads A (4 KB) + B (4 KB) from shared memory every 64 cycles = 128 B/clk, the
// whole shared-memory read bandwidth of an SM: measured 1027 TFLOP/s with NO global traffic at all
// (CLSTM_NOTMA=1 CLSTM_SKIP=1), against 1608 for the N = 256 forward shape (96 B/clk).  Swapping the roles —
// the 128 output channels become M (weights as the A operand), 256 PIXELS become N — restores the N = 256
// shape with the same per-k-block traffic as the forward kernel (16 KB weights + 2 x 16 KB pixel tiles).
//
// The accumulator is then [128 channels (TMEM lanes)] x [256 pixels (columns)]; the epilogue transposes
// 16-pixel groups through shared memory into [16 px][64 ch] fp32 blocks (256-byte rows) and writes them with
// TMA stores to dx [pix][CIP] (channels < split) and dh_prev [pix][HP] (channels >= split).
//
// Roles (384 threads): warp 0 = TMA producer (lane 0: weight box, lanes 1-2: the two pixel tiles), warp 1 =
// MMA issuer, warp 2 = TMEM allocator, warps 4..11 = epilogue (quadrant = 32 channels, two warps per
// quadrant split the 256 pixel columns).
#pragma once
#include "convgemm.cuh"

namespace clstm {

constexpr int kDtStageBytes = 16384 + 2 * kABytes;  // weights [128 x 64] + two pixel tiles [128 x 64]
constexpr int kDtStgHalf = 8192;                    // [16 px][64 ch] fp32 x 2 outputs (x part | h part)

struct DgradTParams {
  int B, H, W;
  int BW, BH, tiles_w, tiles_h, num_m_tiles;
  ConvSeg seg;     // dz: chunks = 4HP/64, kh x kw taps
  int m_tiles;     // 128-row blocks of output channels
  int stages;
  int split_col;   // output channels [0, split) -> X0 (dx), [split, ...) -> X1 (dh_prev); multiple of 64
  int lbw;         // log2(BW)
};

inline size_t dgradT_smem_bytes(int stages) {
  return 1024 + static_cast<size_t>(stages) * kDtStageBytes + 2 * kDtStgHalf + (2 * kMaxStages + 4) * 8 + 16 + 64;
}

template <typename E>
__global__ void __launch_bounds__(kGemmThreads, 1)
dgradT_kernel(const __grid_constant__ CUtensorMap tmDz, const __grid_constant__ CUtensorMap tmW,
              const __grid_constant__ CUtensorMap tmX0, const __grid_constant__ CUtensorMap tmX1,
              const DgradTParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_stg = smem + p.stages * kDtStageBytes;
  uint8_t* tail = smem_stg + 2 * kDtStgHalf;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tmem_full = empty_bar + kMaxStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_pairs = (p.num_m_tiles + 1) >> 1;
  const int total_units = num_pairs * p.m_tiles;
  const int taps = p.seg.kh * p.seg.kw;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmDz);
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 8);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, kTmemCols);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // pixel tile `t` (0/1) of unit: origin (w0, h0, b); b == p.B for the padding tile of an odd tile count
  auto tile_origin = [&](int unit, int t, int& w0, int& h0, int& b) {
    const int mt = 2 * (unit / p.m_tiles) + t;
    w0 = (mt % p.tiles_w) * p.BW;
    h0 = ((mt / p.tiles_w) % p.tiles_h) * p.BH;
    b = mt / (p.tiles_w * p.tiles_h);
  };

  if (warp == 0) {
    // ===================== TMA producer: one lane per box =====================
    if (lane < 3) {
      int stage = 0;
      uint32_t phase = 0;
      for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
        const int mtile = unit % p.m_tiles;
        int w0 = 0, h0 = 0, b = 0;
        if (lane > 0) tile_origin(unit, lane - 1, w0, h0, b);
        int kb = 0;
        for (int dy = 0; dy < p.seg.kh; ++dy)
          for (int dx = 0; dx < p.seg.kw; ++dx)
            for (int ch = 0; ch < p.seg.chunks; ++ch, ++kb) {
              mbar_wait(&empty_bar[stage], phase ^ 1);
              uint8_t* dst = smem + stage * kDtStageBytes;
              if (lane == 0) {
                mbar_expect_tx(&full_bar[stage], kDtStageBytes);
                tma_load_2d(dst, &tmW, &full_bar[stage], kb * kBlockK, mtile * 128);
              } else {
                tma_load_4d(dst + 16384 + (lane - 1) * kABytes, &tmDz, &full_bar[stage], ch * kBlockK,
                            w0 + dx - p.seg.kw / 2, h0 + dy - p.seg.kh / 2, b + p.seg.b_off);
              }
              if (++stage == p.stages) {
                stage = 0;
                phase ^= 1;
              }
            }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: M = 128 channels, N = 256 pixels =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(Elem<E>::kFmt, 128, 256, 0, 0);
      const int kblocks = taps * p.seg.chunks;
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t d = tmem_base + acc * 256;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t base = smem_u32(smem + stage * kDtStageBytes);
          const uint64_t adesc = make_smem_desc_sw128(base, 16, 1024);
          const uint64_t bdesc = make_smem_desc_sw128(base + 16384, 16, 1024);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) umma_f16(d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          umma_commit(&empty_bar[stage]);
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tmem_full[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int q = warp & 3;            // TMEM lanes 32q .. 32q+31 == output channels of this m-tile
    const int half = (warp - 4) >> 2;  // pixel tile (columns 128*half .. 128*half+127)
    const int cl = q * 32 + lane;      // channel within the m-tile
    float* stg = reinterpret_cast<float*>(smem_stg + half * kDtStgHalf);  // [2][16 px][64 ch]
    const int bar_id = 1 + half;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
      const int mtile = unit % p.m_tiles;
      int w0, h0, b;
      tile_origin(unit, half, w0, h0, b);
      mbar_wait(&tmem_full[acc], acc_phase);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + acc * 256 + half * 128 + (static_cast<uint32_t>(q * 32) << 16);
      // channels of this m-tile: global channel = mtile*128 + cl; 64-channel block index decides the destination
#pragma unroll 1
      for (int g = 0; g < 8; ++g) {
        uint32_t v[16];
        tmem_ld16(taddr + g * 16, v);
        if (q == 0 && lane < 2) tma_store_wait_read();
        named_bar_sync(bar_id, 128);  // the staging of this half is free again
        tmem_ld_wait();
        if (g == 7) {
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
        float* dst = stg + (cl >> 6) * (16 * 64) + (cl & 63);
#pragma unroll
        for (int j = 0; j < 16; ++j) dst[j * 64] = __uint_as_float(v[j]);
        fence_proxy_async_smem();
        named_bar_sync(bar_id, 128);
        if (q == 0 && lane < 2) {
          // lane 0: channels [0,64) of the m-tile, lane 1: channels [64,128); 16 consecutive tile pixels
          const int px = g * 16;
          const int wx = w0 + (px & (p.BW - 1)), hy = h0 + (px >> p.lbw);
          const int chan = mtile * 128 + lane * 64;
          if (chan < p.split_col)
            tma_store_4d(&tmX0, stg + lane * (16 * 64), chan, wx, hy, b);
          else
            tma_store_4d(&tmX1, stg + lane * (16 * 64), chan - p.split_col, wx, hy, b);
          tma_store_commit();
        }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (q == 0 && lane < 2) tma_store_wait_all();
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace clstm

// ======================================================================================================
// Halo-row variant of the transposed dgrad (3x3 filters, W > 128): a unit is 256 consecutive pixels of ONE image
// row.  For each 64-channel chunk of dz the three image rows h-1, h, h+1 are loaded once as [258 px x 64 ch] rows
// (a 256-pixel box plus an 8-pixel box) into a ring of row slots; tap (dy, dx) is the B descriptor of row dy
// shifted by dx pixels (N = 256 rows are contiguous inside a row slot).  Pixel-operand traffic drops from
// 9 x 32 KB to 3 x 33 KB per chunk; with the 16 KB weight tile per tap that is 27 KB instead of 48 KB per k-block.
// ======================================================================================================
namespace clstm {

constexpr int kDtRowPx = 264;                       // row slot pitch in pixels (258 used)
constexpr int kDtRowBytes = kDtRowPx * 128;         // 33792
constexpr int kDtMaxRows = 6;

struct DgradTHaloParams {
  int B, H, W;
  int segs_w;      // ceil(W / 256)
  int chunks;      // dz chunks (4HP / 64)
  int m_tiles;
  int w_stages;    // weight ring
  int rows;        // row ring
  int split_col;
};

inline size_t dgradTh_smem_bytes(int w_stages, int rows) {
  return 1024 + static_cast<size_t>(w_stages) * 16384 + static_cast<size_t>(rows) * kDtRowBytes + 2 * kDtStgHalf +
         (2 * kMaxStages + 2 * kDtMaxRows + 4) * 8 + 16 + 64;
}

template <typename E>
__global__ void __launch_bounds__(kGemmThreads, 1)
dgradT_halo_kernel(const __grid_constant__ CUtensorMap tmRow256, const __grid_constant__ CUtensorMap tmRow8,
                   const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmX0,
                   const __grid_constant__ CUtensorMap tmX1, const DgradTHaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_w = smem;
  uint8_t* smem_r = smem_w + p.w_stages * 16384;
  uint8_t* smem_stg = smem_r + p.rows * kDtRowBytes;
  uint8_t* tail = smem_stg + 2 * kDtStgHalf;
  uint64_t* w_full = reinterpret_cast<uint64_t*>(tail);
  uint64_t* w_empty = w_full + kMaxStages;
  uint64_t* r_full = w_empty + kMaxStages;
  uint64_t* r_empty = r_full + kDtMaxRows;
  uint64_t* tmem_full = r_empty + kDtMaxRows;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_units = p.B * p.H * p.segs_w * p.m_tiles;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmRow256);
    tma_prefetch_desc(&tmRow8);
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < p.w_stages; ++s) {
      mbar_init(&w_full[s], 1);
      mbar_init(&w_empty[s], 1);
    }
    for (int s = 0; s < p.rows; ++s) {
      mbar_init(&r_full[s], 1);
      mbar_init(&r_empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 8);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, kTmemCols);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto decode = [&](int unit, int& mtile, int& w0, int& h, int& b) {
    mtile = unit % p.m_tiles;
    int rest = unit / p.m_tiles;
    w0 = (rest % p.segs_w) * 256;
    rest /= p.segs_w;
    h = rest % p.H;
    b = rest / p.H;
  };

  if (warp == 0) {
    // ===================== weight producer =====================
    if (lane == 0) {
      uint32_t idx = 0;
      for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
        const int mtile = unit % p.m_tiles;
        for (int ch = 0; ch < p.chunks; ++ch)
          for (int tap = 0; tap < 9; ++tap, ++idx) {
            const uint32_t stage = idx % p.w_stages;
            mbar_wait(&w_empty[stage], ((idx / p.w_stages) & 1) ^ 1);
            mbar_expect_tx(&w_full[stage], 16384);
            tma_load_2d(smem_w + stage * 16384, &tmW, &w_full[stage], (tap * p.chunks + ch) * kBlockK, mtile * 128);
          }
      }
    }
  } else if (warp == 3) {
    // ===================== halo-row producer: lane 0 the 256-pixel box, lane 1 the 8-pixel tail =====================
    if (lane < 2) {
      uint32_t idx = 0;
      for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
        int mtile, w0, h, b;
        decode(unit, mtile, w0, h, b);
        for (int ch = 0; ch < p.chunks; ++ch)
          for (int r = 0; r < 3; ++r, ++idx) {
            const uint32_t slot = idx % p.rows;
            mbar_wait(&r_empty[slot], ((idx / p.rows) & 1) ^ 1);
            uint8_t* dst = smem_r + slot * kDtRowBytes;
            if (lane == 0) {
              mbar_expect_tx(&r_full[slot], 264 * 128);
              tma_load_4d(dst, &tmRow256, &r_full[slot], ch * kBlockK, w0 - 1, h + r - 1, b);
            } else {
              tma_load_4d(dst + 256 * 128, &tmRow8, &r_full[slot], ch * kBlockK, w0 + 255, h + r - 1, b);
            }
          }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(Elem<E>::kFmt, 128, 256, 0, 0);
      uint32_t w_idx = 0, r_idx = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t d = tmem_base + acc * 256;
        uint32_t first = 1;
        for (int ch = 0; ch < p.chunks; ++ch)
          for (int dy = 0; dy < 3; ++dy, ++r_idx) {
            const uint32_t slot = r_idx % p.rows;
            mbar_wait(&r_full[slot], (r_idx / p.rows) & 1);
            const uint32_t row = smem_u32(smem_r + slot * kDtRowBytes);
            for (int dx = 0; dx < 3; ++dx, ++w_idx) {
              const uint32_t stage = w_idx % p.w_stages;
              mbar_wait(&w_full[stage], (w_idx / p.w_stages) & 1);
              tcgen05_fence_after();
              const uint64_t adesc = make_smem_desc_sw128(smem_u32(smem_w + stage * 16384), 16, 1024);
              const uint64_t bdesc = make_smem_desc_sw128(row + dx * 128, 16, 1024);
#pragma unroll
              for (int k = 0; k < kBlockK / 16; ++k)
                umma_f16(d, adesc + 2 * k, bdesc + 2 * k, idesc, (first && k == 0) ? 0u : 1u);
              first = 0;
              umma_commit(&w_empty[stage]);
            }
            umma_commit(&r_empty[slot]);
          }
        umma_commit(&tmem_full[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (as dgradT_kernel; the unit's 256 pixels are one row segment) ==============
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    const int cl = q * 32 + lane;
    float* stg = reinterpret_cast<float*>(smem_stg + half * kDtStgHalf);
    const int bar_id = 1 + half;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
      int mtile, w0, h, b;
      decode(unit, mtile, w0, h, b);
      mbar_wait(&tmem_full[acc], acc_phase);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + acc * 256 + half * 128 + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
      for (int g = 0; g < 8; ++g) {
        uint32_t v[16];
        tmem_ld16(taddr + g * 16, v);
        if (q == 0 && lane < 2) tma_store_wait_read();
        named_bar_sync(bar_id, 128);
        tmem_ld_wait();
        if (g == 7) {
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
        float* dst = stg + (cl >> 6) * (16 * 64) + (cl & 63);
#pragma unroll
        for (int j = 0; j < 16; ++j) dst[j * 64] = __uint_as_float(v[j]);
        fence_proxy_async_smem();
        named_bar_sync(bar_id, 128);
        if (q == 0 && lane < 2) {
          const int wx = w0 + half * 128 + g * 16;
          const int chan = mtile * 128 + lane * 64;
          if (chan < p.split_col)
            tma_store_4d(&tmX0, stg + lane * (16 * 64), chan, wx, h, b);
          else
            tma_store_4d(&tmX1, stg + lane * (16 * 64), chan - p.split_col, wx, h, b);
          tma_store_commit();
        }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (q == 0 && lane < 2) tma_store_wait_all();
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace clstm
