// Output head, forward:  y = sigmoid(Conv3d(1,3,3)(h) + b)  (conv_lstm.py:198-201), row-marching formulation.
//
// The implicit-GEMM head of convgemm.cuh (EPI_HEAD) fetches every 128-pixel h tile nine times (once per tap) for
// N = 16 output columns: it is bound by L2 -> shared-memory traffic (7.95 ms for the 24 x 16 frames of the bench
// workload, 28 GB of TMA loads for 3.2 GB of h).  This kernel loads every h pixel ONCE and turns the taps into
// output columns:
//     z[p][tap*16 + o] = sum_c h[p][c] * Wh[o][c][tap]            one [128 px x 144] x K=hid MMA per tile
//     y[r][p][o]       = sigmoid(b[o] + sum_{dy,dx} z[r+dy][p+dx][(dy,dx)*16 + o])
// The shift-and-add never touches shared memory: an epilogue thread owns one image COLUMN of a 256-pixel-wide
// strip and marches down the rows of a band.  The dx = +-1 terms come from the neighbouring lanes by warp
// shuffles (warp-boundary lanes exchange 2 x 36 floats through a small shared buffer, one named barrier per row),
// the dy terms are accumulated in three rotating register accumulators (rows R-1, R, R+1).  Out-of-image pixels
// need no special case: TMA zero-fills them, so their z is zero.
//
// Roles (384 threads): warp 0 lane 0 = TMA producer (Wz once, then one 16 KB box per tile and 64-channel chunk),
// warp 1 lane 0 = MMA issuer (M = 128, N = 144), warp 2 = TMEM allocator, warps 4..11 = two epilogue teams
// (left / right 128-pixel tile of the strip; warp % 4 = TMEM lane quadrant).  TMEM: 3 accumulators of 144 columns.
// Units: (image, strip, band of rows); a band recomputes one halo row above and below (6 % for 32-row bands).
#pragma once
#include "convgemm.cuh"

namespace clstm {

constexpr int kHrN = 144;                 // 9 taps x 16 output-channel slots
constexpr int kHrWzBytes = kHrN * 128;    // one 64-channel chunk of Wz (K-major, 128-byte rows)
constexpr int kHrMaxStages = 12;
constexpr int kHrBordFloats = 2 * 8 * 2 * 3 * 16;  // [row parity][warp slot][side][dy][o]

struct HeadRowsParams {
  int H, W;
  int images;   // T * B_img images of this launch
  int img_off;  // first image inside the h tensor map
  int b_img, t0, t_out, c_out;
  int chunks;   // padded hidden / 64
  int strips, strip_w, x_halo;  // strips per row; output pixels per strip; 1 if strips overlap by one pixel each side
  int bands, band_rows;
  int stages;
  const float* bias;
  float* y;     // (B_img, C_out, T_out, H, W)
};

inline size_t head_rows_smem_bytes(int chunks, int stages) {
  return 1024 + static_cast<size_t>(chunks) * kHrWzBytes + static_cast<size_t>(stages) * kABytes +
         kHrBordFloats * 4 + (2 * kHrMaxStages + 8) * 8 + 64;
}

template <typename E, int CO>  // CO = output-channel slots processed per tap (c_out rounded up to 4)
__global__ void __launch_bounds__(kGemmThreads, 1)
head_rows_kernel(const __grid_constant__ CUtensorMap tmH, const __grid_constant__ CUtensorMap tmWz,
                 const HeadRowsParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_wz = smem;
  uint8_t* smem_a = smem_wz + p.chunks * kHrWzBytes;
  float* bord = reinterpret_cast<float*>(smem_a + p.stages * kABytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bord + kHrBordFloats);
  uint64_t* empty_bar = full_bar + kHrMaxStages;
  uint64_t* tfull = empty_bar + kHrMaxStages;  // [3]
  uint64_t* tempty = tfull + 3;                // [3]
  uint64_t* wz_bar = tempty + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wz_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_units = p.images * p.strips * p.bands;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmH);
    tma_prefetch_desc(&tmWz);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 3; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 4);
    }
    mbar_init(wz_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, kTmemCols);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // unit -> (image, first pixel of the strip, row band); the same decomposition in every role
  auto unit_geom = [&](int unit, int& img, int& x0, int& r0, int& r1, int& ntile) {
    const int band = unit % p.bands;
    const int strip = (unit / p.bands) % p.strips;
    img = unit / (p.bands * p.strips);
    x0 = strip * p.strip_w - p.x_halo;
    r0 = band * p.band_rows;
    r1 = min(r0 + p.band_rows, p.H);
    ntile = (x0 + 128 < p.W) ? 2 : 1;
  };

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(wz_bar, p.chunks * kHrWzBytes);
      for (int ch = 0; ch < p.chunks; ++ch) tma_load_2d(smem_wz + ch * kHrWzBytes, &tmWz, wz_bar, 0, ch * kHrN);
      int stage = 0;
      uint32_t phase = 0;
      for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
        int img, x0, r0, r1, ntile;
        unit_geom(unit, img, x0, r0, r1, ntile);
        for (int R = max(r0 - 1, 0); R <= min(r1, p.H - 1); ++R)
          for (int t = 0; t < ntile; ++t)
            for (int ch = 0; ch < p.chunks; ++ch) {
              mbar_wait(&empty_bar[stage], phase ^ 1);
              mbar_expect_tx(&full_bar[stage], kABytes);
              tma_load_4d(smem_a + stage * kABytes, &tmH, &full_bar[stage], ch * kBlockK, x0 + t * 128, R,
                          p.img_off + img);
              if (++stage == p.stages) {
                stage = 0;
                phase ^= 1;
              }
            }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc(Elem<E>::kFmt, 128, kHrN, 0, 0);
      mbar_wait(wz_bar, 0);
      int stage = 0;
      uint32_t phase = 0;
      int ti = 0;  // running tile index -> accumulator ti % 3
      for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
        int img, x0, r0, r1, ntile;
        unit_geom(unit, img, x0, r0, r1, ntile);
        for (int R = max(r0 - 1, 0); R <= min(r1, p.H - 1); ++R)
          for (int t = 0; t < ntile; ++t, ++ti) {
            const int buf = ti % 3;
            mbar_wait(&tempty[buf], ((ti / 3) & 1) ^ 1);
            tcgen05_fence_after();
            const uint32_t d = tmem_base + buf * kHrN;
            for (int ch = 0; ch < p.chunks; ++ch) {
              mbar_wait(&full_bar[stage], phase);
              tcgen05_fence_after();
              const uint64_t adesc = make_smem_desc_sw128(smem_u32(smem_a + stage * kABytes), 16, 1024);
              const uint64_t bdesc = make_smem_desc_sw128(smem_u32(smem_wz + ch * kHrWzBytes), 16, 1024);
#pragma unroll
              for (int k = 0; k < kBlockK / 16; ++k) umma_f16(d, adesc + 2 * k, bdesc + 2 * k, idesc, (ch | k) != 0);
              umma_commit(&empty_bar[stage]);
              if (++stage == p.stages) {
                stage = 0;
                phase ^= 1;
              }
            }
            umma_commit(&tfull[buf]);
          }
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int team = (warp - 4) >> 2;
    const int wslot = team * 4 + q;               // position of this warp along the 256-pixel strip
    const int pl = wslot * 32 + lane;             // pixel inside the strip window
    const size_t plane = static_cast<size_t>(p.H) * p.W;
    float bias_r[CO];
#pragma unroll
    for (int o = 0; o < CO; ++o) bias_r[o] = (o < p.c_out) ? __ldg(p.bias + o) : 0.f;
    int ti = 0;   // tiles issued before the current row (all teams count alike)
    int par = 0;  // parity of the border exchange buffer
    for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
      int img, x0, r0, r1, ntile;
      unit_geom(unit, img, x0, r0, r1, ntile);
      const int x = x0 + pl;
      const bool store_px = pl >= p.x_halo && pl < p.x_halo + p.strip_w && x < p.W;
      const int t = p.t0 + img / p.b_img, bi = img % p.b_img;
      float* ycol = p.y + ((static_cast<size_t>(bi) * p.c_out) * p.t_out + t) * plane + x;
      float acc_a[CO], acc_b[CO], acc_c[CO];  // output rows R-1, R, R+1
#pragma unroll
      for (int o = 0; o < CO; ++o) acc_a[o] = acc_b[o] = 0.f;
#pragma unroll 1
      for (int R = r0 - 1; R <= r1; ++R) {
        if (R >= 0 && R < p.H) {
          float* bw = bord + (par * 8 + wslot) * (2 * 3 * 16);
          if (team < ntile) {
            const int tix = ti + team;
            const int buf = tix % 3;
            mbar_wait(&tfull[buf], (tix / 3) & 1);
            tcgen05_fence_after();
            const uint32_t taddr = tmem_base + buf * kHrN + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll
            for (int dyi = 0; dyi < 3; ++dyi) {
              uint32_t zl[16], zc[16], zr[16];
              tmem_ld16(taddr + (dyi * 3 + 0) * 16, zl);
              tmem_ld16(taddr + (dyi * 3 + 1) * 16, zc);
              tmem_ld16(taddr + (dyi * 3 + 2) * 16, zr);
              tmem_ld_wait();
              if (lane == 0) {  // my (dy,+1) taps are the left neighbour's dx = +1 term
#pragma unroll
                for (int o = 0; o < CO; ++o) bw[(0 * 3 + dyi) * 16 + o] = __uint_as_float(zr[o]);
              }
              if (lane == 31) {  // my (dy,-1) taps are the right neighbour's dx = -1 term
#pragma unroll
                for (int o = 0; o < CO; ++o) bw[(1 * 3 + dyi) * 16 + o] = __uint_as_float(zl[o]);
              }
#pragma unroll
              for (int o = 0; o < CO; ++o) {
                float l = __shfl_up_sync(0xffffffffu, __uint_as_float(zl[o]), 1);
                float r = __shfl_down_sync(0xffffffffu, __uint_as_float(zr[o]), 1);
                if (lane == 0) l = 0.f;
                if (lane == 31) r = 0.f;
                const float s = __uint_as_float(zc[o]) + l + r;
                if (dyi == 0) acc_c[o] = s;        // dy = -1: first contribution to row R+1
                else if (dyi == 1) acc_b[o] += s;  // dy = 0
                else acc_a[o] += s;                // dy = +1: last contribution to row R-1
              }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[buf]);
          } else {
#pragma unroll
            for (int o = 0; o < CO; ++o) acc_c[o] = 0.f;
          }
          named_bar_sync(1, 256);
          // warp-boundary lanes: the neighbour pixel lives in another warp (or outside the strip: zero)
          if (team < ntile) {
            if (lane == 0 && wslot > 0) {
              const float* nb = bord + (par * 8 + wslot - 1) * (2 * 3 * 16) + 1 * 3 * 16;
#pragma unroll
              for (int o = 0; o < CO; ++o) {
                acc_c[o] += nb[0 * 16 + o];
                acc_b[o] += nb[1 * 16 + o];
                acc_a[o] += nb[2 * 16 + o];
              }
            }
            if (lane == 31 && wslot + 1 < ntile * 4) {
              const float* nb = bord + (par * 8 + wslot + 1) * (2 * 3 * 16);
#pragma unroll
              for (int o = 0; o < CO; ++o) {
                acc_c[o] += nb[0 * 16 + o];
                acc_b[o] += nb[1 * 16 + o];
                acc_a[o] += nb[2 * 16 + o];
              }
            }
          }
          ti += ntile;
          par ^= 1;
        } else {
#pragma unroll
          for (int o = 0; o < CO; ++o) acc_c[o] = 0.f;
        }
        // row R-1 has now received its three dy contributions
        if (R - 1 >= r0 && R - 1 < r1 && store_px) {
          float* yrow = ycol + static_cast<size_t>(R - 1) * p.W;
#pragma unroll
          for (int o = 0; o < CO; ++o)
            if (o < p.c_out) yrow[static_cast<size_t>(o) * p.t_out * plane] = fast_sigmoid(fmaf(acc_a[o], kHScaleInv, bias_r[o]));
        }
#pragma unroll
        for (int o = 0; o < CO; ++o) {
          acc_a[o] = acc_b[o];
          acc_b[o] = acc_c[o];
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
