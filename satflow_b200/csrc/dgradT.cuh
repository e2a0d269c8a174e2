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
#include "pointwise.cuh"

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
  int rotate;      // 1: unit u starts its K loop at k-block u % kblocks (spreads the concurrent weight reads over L2)
  int extra;       // dgradT_fused2_kernel only: k-blocks of a SECOND, un-shifted K segment (tmG / tmW2) appended to the
                   // K loop — the head's dgrad for the consumer's output frame, accumulated into the same TMEM tile
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
        // rotated start: (dy, dx, ch) decoded once per unit, then stepped with wrap-around
        const int kblocks = taps * p.seg.chunks;
        int kb = p.rotate ? unit % kblocks : 0;
        int ch = kb % p.seg.chunks, dx = (kb / p.seg.chunks) % p.seg.kw, dy = kb / (p.seg.chunks * p.seg.kw);
        for (int i = 0; i < kblocks; ++i) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* dst = smem + stage * kDtStageBytes;
          if (lane == 0) {
            mbar_expect_tx(&full_bar[stage], kDtStageBytes);
            tma_load_2d(dst, &tmW, &full_bar[stage], kb * kBlockK, mtile * 128);
          } else {
            tma_load_4d(dst + 16384 + (lane - 1) * kABytes, &tmDz, &full_bar[stage], ch * kBlockK,
                        w0 + dx - p.seg.kw / 2, h0 + dy - p.seg.kh / 2, b + p.seg.b_off);
          }
          ++kb;
          if (++ch == p.seg.chunks) {
            ch = 0;
            if (++dx == p.seg.kw) {
              dx = 0;
              if (++dy == p.seg.kh) dy = 0, kb = 0;
            }
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
      uint64_t adesc = make_smem_desc_sw128(smem_u32(smem), 16, 1024);
      uint64_t bdesc = make_smem_desc_sw128(smem_u32(smem) + 16384, 16, 1024);
      for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t d = tmem_base + acc * 256;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) umma_f16(d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          umma_commit(&empty_bar[stage]);
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
          const uint32_t base = smem_u32(smem + stage * kDtStageBytes);  // next stage's descriptors, off the wait path
          adesc = make_smem_desc_sw128(base, 16, 1024);
          bdesc = make_smem_desc_sw128(base + 16384, 16, 1024);
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
// dgradT with the NEXT gate-gradient fused into its epilogue.
//
// In BPTT the x-part of dgrad(cell k+1, step t) is the last-arriving gradient source of gate-grad(cell k, step t)
// (likewise the decoder feedback).  gate-grad is a pure HBM stream (2.4 GB, 0.43 ms) during which the tensor
// pipes idle, and dgradT is tensor bound with DRAM 25 % busy — so the epilogue of dgradT consumes its own x-part
// tile straight from shared memory and runs the gate-gradient math of the consumer cell for those pixels:
// it adds the consumer's other dh sources, reads gates / c_prev / c_next / dc, writes dz (16-bit, into the OTHER dz
// buffer — this kernel is still reading its own dz through TMA), updates dc in place and accumulates the bias
// partial sums.  dx is never written or re-read, the separate gate-grad launch disappears, and its memory traffic
// overlaps the MMAs.  Pixel groups are 32 wide; thread mapping of the fused part = gate_grad_kernel's
// (one pixel x 8 channels per item).
// ======================================================================================================
namespace clstm {

constexpr int kDfStgHalf = 16384;  // [2 blocks (x|h)][32 px][64 ch] fp32

struct GateFuse {
  const void* gates;     // E [pix][4*HP] of the consumer cell / step
  const float* c_prev;   // nullable (zeros)
  const float* c_next;
  const float* src0;     // dh sources of the consumer read from global memory (nullable), fp32 [pix][HP], scaled by S
  const float* src1;
  const float* src2;
  int use_stg;           // 1: staging block 0 (this launch's dx tile, 64 channels) is one more dh source
  int h_block;           // staging block holding dh_prev of the producing cell (1 with an x part, else 0)
  int pf_dist;           // L2 prefetch distance in 32-pixel groups (0 = off)
  float* dc;             // in/out
  void* dz_out;          // E [pix][4*HP]
  float* bias_partial;   // [gridDim.x][4*HP], accumulated
  unsigned int* dz_absmax;  // range statistics of the 16-bit dz this pass writes (float bits, see fold_absmax)
  int HP;
  int c16;               // 1: c_prev (and c_next) are E arrays holding c * kCScale (CLSTM_C16; worker-warp kernel only)
  int fuse_units;        // hybrid schedule: only units < fuse_units run the gate gradient here; for the others the dx
                         // block is written to HBM (tmX0) and the worker warps of the following wgrad launch finish
                         // the job (clstm.cu "hybrid").  INT_MAX = everything here.
};

inline size_t dgradTf_smem_bytes(int stages) {
  return 1024 + static_cast<size_t>(stages) * kDtStageBytes + 2 * kDfStgHalf + (2 * kMaxStages + 4) * 8 + 16 + 64;
}

// TEAMS = epilogue teams of 4 warps (one warp per TMEM lane quadrant); a team owns 256/TEAMS accumulator columns.
// TEAMS = 2: 384 threads, 32-pixel groups, two register sets of loads in flight per thread.
// TEAMS = 4: 640 threads (<= 96 registers), 16-pixel groups, one set per thread — twice the warps per scheduler.
// SETS (TEAMS == 2 only) = register sets of raw gate-gradient loads a thread keeps in flight.  The pass is latency
// bound (ncu source view, profiles/r2_fused_dgrad_ncu.md: a third of all warp samples sit on the first use of a
// loaded value), and registers are what holds loads in flight.  SETS = 4 rebalances them with setmaxnreg — the four
// control warps need 56 registers, not the 168 of the launch bound: 128 x 112 registers move to the 256 epilogue
// threads (224 each) — and every item's loads are then issued a WHOLE 32-pixel group ahead of their use.
template <typename E, int TEAMS, int SETS = 2>
__global__ void __launch_bounds__(128 + TEAMS * 128, 1)
dgradT_fused_kernel(const __grid_constant__ CUtensorMap tmDz, const __grid_constant__ CUtensorMap tmW,
                    const __grid_constant__ CUtensorMap tmX0, const __grid_constant__ CUtensorMap tmX1,
                    const DgradTParams p, const GateFuse f) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_stg = smem + p.stages * kDtStageBytes;
  uint8_t* tail = smem_stg + 2 * kDfStgHalf;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tmem_full = empty_bar + kMaxStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_units = (p.num_m_tiles + 1) >> 1;  // m_tiles == 1: 64 x-channels + 64 h-channels
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
      mbar_init(&tmem_empty[a], 4 * TEAMS);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, kTmemCols);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr int GP = 64 / TEAMS;          // pixels per epilogue group
  constexpr int TCOLS = 256 / TEAMS;      // accumulator columns (pixels) per team
  constexpr int STG_TEAM = 2 * GP * 64;   // floats of staging per team: [2 blocks (x|h)][GP px][64 ch]

  auto tile_origin = [&](int unit, int t, int& w0, int& h0, int& b) {
    const int mt = 2 * unit + t;
    w0 = (mt % p.tiles_w) * p.BW;
    h0 = ((mt / p.tiles_w) % p.tiles_h) * p.BH;
    b = mt / (p.tiles_w * p.tiles_h);
  };

  if (warp < 4) {
  // ---- control warpgroup (TMA producer, MMA issuer, TMEM allocator, idle): gives registers back when SETS == 4
  if constexpr (SETS == 4) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;" ::: "memory");
  if (warp == 0) {
    if (lane < 3) {
      int stage = 0;
      uint32_t phase = 0;
      for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
        int w0 = 0, h0 = 0, b = 0;
        if (lane > 0) tile_origin(unit, lane - 1, w0, h0, b);
        const int kblocks = taps * p.seg.chunks;
        int kb = p.rotate ? unit % kblocks : 0;  // rotated start, see dgradT_kernel
        int ch = kb % p.seg.chunks, dx = (kb / p.seg.chunks) % p.seg.kw, dy = kb / (p.seg.chunks * p.seg.kw);
        for (int i = 0; i < kblocks; ++i) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* dst = smem + stage * kDtStageBytes;
          if (lane == 0) {
            mbar_expect_tx(&full_bar[stage], kDtStageBytes);
            tma_load_2d(dst, &tmW, &full_bar[stage], kb * kBlockK, 0);
          } else {
            tma_load_4d(dst + 16384 + (lane - 1) * kABytes, &tmDz, &full_bar[stage], ch * kBlockK,
                        w0 + dx - p.seg.kw / 2, h0 + dy - p.seg.kh / 2, b + p.seg.b_off);
          }
          ++kb;
          if (++ch == p.seg.chunks) {
            ch = 0;
            if (++dx == p.seg.kw) {
              dx = 0;
              if (++dy == p.seg.kh) dy = 0, kb = 0;
            }
          }
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc(Elem<E>::kFmt, 128, 256, 0, 0);
      const int kblocks = taps * p.seg.chunks;
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      uint64_t adesc = make_smem_desc_sw128(smem_u32(smem), 16, 1024);
      uint64_t bdesc = make_smem_desc_sw128(smem_u32(smem) + 16384, 16, 1024);
      for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t d = tmem_base + acc * 256;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) umma_f16(d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          umma_commit(&empty_bar[stage]);
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
          const uint32_t base = smem_u32(smem + stage * kDtStageBytes);  // next stage's descriptors, off the wait path
          adesc = make_smem_desc_sw128(base, 16, 1024);
          bdesc = make_smem_desc_sw128(base + 16384, 16, 1024);
        }
        umma_commit(&tmem_full[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  }
  } else {
    if constexpr (SETS == 4) asm volatile("setmaxnreg.inc.sync.aligned.u32 224;" ::: "memory");
    // ===================== epilogue: transpose 32-pixel groups, h-part -> TMA store, x-part -> fused gate-grad ======
    // The gate-gradient part is latency bound unless its global loads are kept in flight while the thread drains
    // TMEM and waits at barriers: items are 1 pixel x 4 channels (32 registers of raw loads), two register sets
    // ping-pong, and the loads of an item are issued two items ahead — across group and unit boundaries.
    const int q = warp & 3;
    const int team = (warp - 4) >> 2;
    const int half = team / (TEAMS / 2);          // pixel tile of the unit
    const int pcol0 = (team % (TEAMS / 2)) * TCOLS;  // first tile pixel of this team
    const int cl = q * 32 + lane;                 // channel within the 128 output rows == thread inside the team
    float* stg = reinterpret_cast<float*>(smem_stg) + team * STG_TEAM;  // [2][GP px][64 ch]
    const int bar_id = 1 + team;
    const int HP = f.HP;
    const int chunk = cl & 15;                    // 4-channel chunk of the consumer's hidden channels
    const int pxl = cl >> 4;                      // 0..7: pixel inside an 8-pixel pass
    const E* gates = static_cast<const E*>(f.gates);
    E* dzo = static_cast<E*>(f.dz_out);
    float bsum[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int e = 0; e < 4; ++e) bsum[a][e] = 0.f;
    uint32_t zmax = 0;

    struct Raw {
      uint2 g[4];
      float4 cp, cn, dc, s0, s1, s2;
      unsigned pix;  // global pixel index (B*H*W < 2^32 / (4*HP) is checked on the host)
      bool valid;
    };
    // Per-unit tile base (integer divisions once per unit, not per item): first pixel index of this half's pixel
    // tile, and the number of valid rows / columns inside it (0 rows = padding tile or no such unit).
    struct TileBase {
      unsigned pix0;
      int rows, cols;
    };
    auto tile_base = [&](int unit) -> TileBase {
      TileBase t{0u, 0, 0};
      if (unit >= total_units || unit >= f.fuse_units) return t;  // no gate-gradient items for this unit
      int w0, h0, b;
      tile_origin(unit, half, w0, h0, b);
      if (b >= p.B) return t;
      t.pix0 = (static_cast<unsigned>(b) * p.H + h0) * p.W + w0;
      t.rows = p.H - h0;
      t.cols = p.W - w0;
      return t;
    };
    // thread-constant bases (the chunk offset folded in)
    const E* gates_c = gates + chunk * 4;
    E* dzo_c = dzo + chunk * 4;
    const float* cp_c = f.c_prev ? f.c_prev + chunk * 4 : nullptr;
    const float* cn_c = f.c_next + chunk * 4;
    float* dc_c = f.dc + chunk * 4;
    const float* s0_c = f.src0 ? f.src0 + chunk * 4 : nullptr;
    const float* s1_c = f.src1 ? f.src1 + chunk * 4 : nullptr;
    const float* s2_c = f.src2 ? f.src2 + chunk * 4 : nullptr;
    const int bwm = p.BW - 1;

    auto issue = [&](Raw& r, const TileBase& tb, int g, int s) {
      const int px = pcol0 + g * GP + s * 8 + pxl;
      const int lx = px & bwm, ly = px >> p.lbw;
      r.valid = ly < tb.rows && lx < tb.cols;
      if (!r.valid) return;
      r.pix = tb.pix0 + ly * p.W + lx;
      const unsigned o4 = r.pix * (4 * 64), o1 = r.pix * 64;  // HP == 64
      r.g[0] = __ldg(reinterpret_cast<const uint2*>(gates_c + o4));
      r.g[1] = __ldg(reinterpret_cast<const uint2*>(gates_c + o4 + 64));
      r.g[2] = __ldg(reinterpret_cast<const uint2*>(gates_c + o4 + 128));
      r.g[3] = __ldg(reinterpret_cast<const uint2*>(gates_c + o4 + 192));
      r.cp = cp_c ? __ldg(reinterpret_cast<const float4*>(cp_c + o1)) : make_float4(0.f, 0.f, 0.f, 0.f);
      r.cn = __ldg(reinterpret_cast<const float4*>(cn_c + o1));
      r.dc = *reinterpret_cast<const float4*>(dc_c + o1);  // read and written by this thread only
      if (s0_c) r.s0 = __ldg(reinterpret_cast<const float4*>(s0_c + o1));
      if (s1_c) r.s1 = __ldg(reinterpret_cast<const float4*>(s1_c + o1));
      if (s2_c) r.s2 = __ldg(reinterpret_cast<const float4*>(s2_c + o1));
    };
    const float* stg_c = stg + pxl * 64 + chunk * 4;
    auto consume = [&](const Raw& r, int s) {
      if (!r.valid) return;
      float dhv[4] = {0.f, 0.f, 0.f, 0.f};
      if (f.use_stg) {
        const float4 a0 = *reinterpret_cast<const float4*>(stg_c + s * (8 * 64));
        dhv[0] = a0.x, dhv[1] = a0.y, dhv[2] = a0.z, dhv[3] = a0.w;
      }
      if (s0_c) dhv[0] += r.s0.x, dhv[1] += r.s0.y, dhv[2] += r.s0.z, dhv[3] += r.s0.w;
      if (s1_c) dhv[0] += r.s1.x, dhv[1] += r.s1.y, dhv[2] += r.s1.z, dhv[3] += r.s1.w;
      if (s2_c) dhv[0] += r.s2.x, dhv[1] += r.s2.y, dhv[2] += r.s2.z, dhv[3] += r.s2.w;
      float4 dcn;
      uint2 dzp[4];
      gate_grad_item4<E>(r.g, r.cp, r.cn, r.dc, dhv, bsum, zmax, dcn, dzp);
      const unsigned o4 = r.pix * (4 * 64), o1 = r.pix * 64;
      *reinterpret_cast<float4*>(dc_c + o1) = dcn;
#pragma unroll
      for (int a = 0; a < 4; ++a) *reinterpret_cast<uint2*>(dzo_c + o4 + a * 64) = dzp[a];
    };

    // L2 prefetch of the lines the NEXT group will load (costs no registers: the memory system holds the requests).
    // Thread -> (pixel = cl / 4, quarter = cl % 4): one gate line (64 ch x 2 B) and every 4th fp32 half-row line.
    auto prefetch_group = [&](const TileBase& tb, int g) {
      if ((cl >> 2) >= GP) return;
      const int px = pcol0 + g * GP + (cl >> 2), sub = cl & 3;
      const int lx = px & bwm, ly = px >> p.lbw;
      if (ly >= tb.rows || lx >= tb.cols) return;
      const unsigned pix = tb.pix0 + ly * p.W + lx;
      prefetch_l2(gates + pix * (4 * 64) + sub * 64);
      const float* arr[6] = {f.c_prev, f.c_next, f.dc, f.src0, f.src1, f.src2};
#pragma unroll
      for (int j = 0; j < 6; ++j)
#pragma unroll
        for (int l = 0; l < 2; ++l)
          if (((j * 2 + l) & 3) == sub && arr[j] != nullptr) prefetch_l2(arr[j] + pix * 64 + l * 32);
    };

    TileBase tb_cur = tile_base(blockIdx.x), tb_next;
    Raw ra, rb, rc, rd;
    issue(ra, tb_cur, 0, 0);
    if (TEAMS == 2) issue(rb, tb_cur, 0, 1);
    if (TEAMS == 2 && SETS == 4) {
      issue(rc, tb_cur, 0, 2);
      issue(rd, tb_cur, 0, 3);
    }
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
      int w0, h0, b;
      tile_origin(unit, half, w0, h0, b);
      tb_next = tile_base(unit + static_cast<int>(gridDim.x));
      mbar_wait(&tmem_full[acc], acc_phase);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + acc * 256 + team * TCOLS + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
      for (int g = 0; g < 4; ++g) {
        {
          uint32_t v[GP];
#pragma unroll
          for (int i = 0; i < GP / 16; ++i) tmem_ld16(taddr + g * GP + i * 16, *reinterpret_cast<uint32_t(*)[16]>(v + i * 16));
          if (q == 0 && lane < 2) tma_store_wait_read();
          named_bar_sync(bar_id, 128);  // staging free: the previous group's TMA store and fused reads are done
          tmem_ld_wait();
          if (g == 3) {
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
          }
          float* dst = stg + (cl >> 6) * (GP * 64) + (cl & 63);
#pragma unroll
          for (int j = 0; j < GP; ++j) dst[j * 64] = __uint_as_float(v[j]);
        }
        fence_proxy_async_smem();
        named_bar_sync(bar_id, 128);
        if (q == 0 && lane == 1) {  // dh_prev of this cell (64 channels), 16-pixel boxes
#pragma unroll
          for (int s2 = 0; s2 < GP / 16; ++s2) {
            const int px = pcol0 + g * GP + s2 * 16;
            tma_store_4d(&tmX1, stg + f.h_block * (GP * 64) + s2 * 16 * 64, 0, w0 + (px & (p.BW - 1)), h0 + (px >> p.lbw),
                         b);
          }
          tma_store_commit();
        }
        if (q == 0 && lane == 0 && unit >= f.fuse_units && f.h_block == 1) {
          // hybrid schedule: the gate gradient of these pixels runs in the next wgrad launch, which reads dx from HBM
#pragma unroll
          for (int s2 = 0; s2 < GP / 16; ++s2) {
            const int px = pcol0 + g * GP + s2 * 16;
            tma_store_4d(&tmX0, stg + s2 * 16 * 64, 0, w0 + (px & (p.BW - 1)), h0 + (px >> p.lbw), b);
          }
          tma_store_commit();
        }
        // fused gate gradient of the consumer cell: passes of 8 pixels, loads issued ahead of their use
        const int ng = (g + 1) & 3;
        const TileBase& tbn = (g == 3) ? tb_next : tb_cur;
        if (f.pf_dist && SETS != 4) prefetch_group(tbn, ng);
        if (TEAMS == 2 && SETS == 4) {
          // every set is re-issued for the NEXT group right after its item of this group is consumed
          consume(ra, 0);
          issue(ra, tbn, ng, 0);
          consume(rb, 1);
          issue(rb, tbn, ng, 1);
          consume(rc, 2);
          issue(rc, tbn, ng, 2);
          consume(rd, 3);
          issue(rd, tbn, ng, 3);
        } else if (TEAMS == 2) {
          consume(ra, 0);
          issue(ra, tb_cur, g, 2);
          consume(rb, 1);
          issue(rb, tb_cur, g, 3);
          consume(ra, 2);
          issue(ra, tbn, ng, 0);
          consume(rb, 3);
          issue(rb, tbn, ng, 1);
        } else {
          consume(ra, 0);
          issue(ra, tb_cur, g, 1);
          consume(ra, 1);
          issue(ra, tbn, ng, 0);
        }
      }
      tb_cur = tb_next;
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (q == 0 && lane < 2) tma_store_wait_all();
    fold_absmax<E>(zmax, f.dz_absmax);
    // bias partial sums: reduce over the 16 threads (both halves) that share a channel chunk, one gate at a time
    const int te = threadIdx.x - 128;  // over the epilogue warps; te & 15 == chunk
    float* red = reinterpret_cast<float*>(smem_stg);
#pragma unroll
    for (int a = 0; a < 4; ++a) {  // unrolled: bsum must stay in registers
      named_bar_sync(1 + TEAMS, TEAMS * 128);
#pragma unroll
      for (int e = 0; e < 4; ++e) red[te * 5 + e] = bsum[a][e];
      named_bar_sync(1 + TEAMS, TEAMS * 128);
      if (te < 16) {
        for (int e = 0; e < 4; ++e) {
          float sum = 0.f;
          for (int pl = 0; pl < TEAMS * 8; ++pl) sum += red[(pl * 16 + te) * 5 + e];
          f.bias_partial[static_cast<size_t>(blockIdx.x) * 4 * HP + a * HP + te * 4 + e] += sum;
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace clstm
