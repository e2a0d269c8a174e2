// Third-generation pixel-major implicit-GEMM convolution (single CTA, no cluster):
//
//   * a unit = TWO vertically adjacent 128-pixel row tiles (rows h0, h0+1) x one N sub-tile of <= 128
//     columns.  Both tiles are multiplied by the SAME weight stage, so weight traffic per pixel halves, and
//     with N <= 128 the two accumulators (2 x 128 TMEM columns) still double-buffer against the epilogue;
//   * the A operand is halo-stationary: per 64-channel chunk the kh+1 image rows h0-kh/2 .. h0+1+kh/2 are
//     loaded ONCE (130-pixel rows for a 3-wide filter) into a ring of row slots; tap (dy, dx) of tile t is
//     the tcgen05 descriptor of ring row dy+t with its start address shifted by dx*128 B (validated by
//     clstm_selftest_shifted_desc).  A row is released to the producer as soon as its last tap row retires.
//   L2 -> SMEM bytes per 128-pixel tile, K = (64+64)*9: fused cell step (N = 256 as two sub-tiles)
//   133 + 288 = 421 KB instead of 864 KB; dgrad (N = 128) 133 + 144 = 277 KB instead of 1152 KB.
//   The per-tap kernel of convgemm.cuh saturates the SM's L2 ingest port (~64 B/clk): DESIGN.md §4.
//
// Roles (384 threads): warp 0 = A-row TMA producer, warp 3 = weight TMA producer (one lane per 32-row box),
// warp 1 = MMA issuer, warp 2 = TMEM allocator, warps 4..11 = epilogue: warps 4-7 own tile 0, warps 8-11 tile 1,
// staged through swizzled shared memory and written with TMA stores (one lane per output tensor).
#pragma once
#include "convgemm.cuh"

namespace clstm {

constexpr int kH2MaxRows = 10;
constexpr int kH2MaxBStages = 8;
constexpr int kH2StgLstm = 28672;  // per tile: c 8 KB | h 4 KB | gates 4 x 4 KB
constexpr int kH2StgStore = 8192;  // per tile: one fp32 [128 x 16] group
constexpr int kH2C = 0, kH2H = 8192, kH2G = 12288;

struct Halo2Params {
  ConvGemmParams g;  // geometry (BW == 128, BH == 1), segments, epilogue pointers / image offsets
  int n_sub;         // accumulator columns per tile and unit (64 or 128)
  int n_subs;        // sub-tiles covering N
  int row_pairs;     // ceil(H / 2)
  int a_rows;        // ring slots
  int a_row_bytes;   // slot size (max pitch * 128)
  int b_stages;
  int pitch[2], halo_w[2];
};

__host__ __device__ constexpr int h2_stg_bytes(int epi) { return epi == 0 ? kH2StgLstm : (epi == 1 ? kH2StgStore : 0); }

inline size_t halo2_smem_bytes(int a_rows, int a_row_bytes, int b_stages, int n_sub, int stg, int bias_floats) {
  return 1024 + static_cast<size_t>(a_rows) * a_row_bytes + static_cast<size_t>(b_stages) * n_sub * 128 + 2 * stg +
         (2 * kH2MaxRows + 2 * kH2MaxBStages + 4) * 8 + 16 + static_cast<size_t>(bias_floats) * 4 + 64;
}

template <typename E, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
halo2_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
             const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmX0,
             const __grid_constant__ CUtensorMap tmX1, const __grid_constant__ CUtensorMap tmX2,
             const Halo2Params hp) {
  const ConvGemmParams& p = hp.g;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int b_stage_bytes = hp.n_sub * 128;
  constexpr int kStg = h2_stg_bytes(EPI);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem_a + hp.a_rows * hp.a_row_bytes;
  uint8_t* smem_stg = smem_b + hp.b_stages * b_stage_bytes;
  uint8_t* tail = smem_stg + 2 * kStg;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(tail);
  uint64_t* a_empty = a_full + kH2MaxRows;
  uint64_t* b_full = a_empty + kH2MaxRows;
  uint64_t* b_empty = b_full + kH2MaxBStages;
  uint64_t* tmem_full = b_empty + kH2MaxBStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* bias_s = reinterpret_cast<float*>(tmem_slot + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_units = p.B * hp.row_pairs * p.tiles_w * hp.n_subs;
  const int bias_n = p.n_tiles * p.n_tile;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA0);
    if (p.nseg > 1) tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < hp.a_rows; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < hp.b_stages; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 8);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, kTmemCols);
  if (p.bias != nullptr) {
    for (int i = threadIdx.x; i < bias_n; i += blockDim.x) bias_s[i] = p.bias[i];
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto decode = [&](int unit, int& ns, int& w0, int& h0, int& b) {
    ns = unit % hp.n_subs;
    int rest = unit / hp.n_subs;
    w0 = (rest % p.tiles_w) * 128;
    rest /= p.tiles_w;
    h0 = (rest % hp.row_pairs) * 2;
    b = rest / hp.row_pairs;
  };

  if (warp == 0) {
    // ===================== A-row producer =====================
    if (lane == 0) {
      uint32_t idx = 0;  // running ring-row counter
      for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
        int ns, w0, h0, b;
        decode(unit, ns, w0, h0, b);
        for (int s = 0; s < p.nseg; ++s) {
          const CUtensorMap* tmA = (s == 0) ? &tmA0 : &tmA1;
          const ConvSeg sg = p.seg[s];
          const int row_bytes = hp.halo_w[s] * 128;
          for (int ch = 0; ch < sg.chunks; ++ch)
            for (int r = 0; r <= sg.kh; ++r, ++idx) {
              const uint32_t slot = idx % hp.a_rows, phase = (idx / hp.a_rows) & 1;
              mbar_wait(&a_empty[slot], phase ^ 1);
              mbar_expect_tx(&a_full[slot], row_bytes);
              tma_load_4d(smem_a + slot * hp.a_row_bytes, tmA, &a_full[slot], ch * kBlockK, w0 - sg.kw / 2,
                          h0 - sg.kh / 2 + r, b + sg.b_off);
            }
        }
      }
    }
  } else if (warp == 3) {
    // ===================== weight producer: n_sub / 32 boxes of 32 rows per stage, one lane each =============
    const int box_rows = hp.n_sub < 32 ? hp.n_sub : 32;
    const int boxes = hp.n_sub / box_rows;
    if (lane < boxes) {
      uint32_t idx = 0;
      for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
        const int ns = unit % hp.n_subs;
        // global weight row of this lane's box
        int row;
        if constexpr (EPI == EPI_LSTM)
          row = (ns >> 1) * 256 + lane * 64 + (ns & 1) * 32;  // gate `lane`, 32-channel half of the 64-channel block
        else
          row = ns * hp.n_sub + lane * box_rows;
        int kb = 0;
        for (int s = 0; s < p.nseg; ++s) {
          const ConvSeg sg = p.seg[s];
          const int taps = sg.kh * sg.kw;
          for (int ch = 0; ch < sg.chunks; ++ch)
            for (int tap = 0; tap < taps; ++tap, ++idx) {
              const uint32_t stage = idx % hp.b_stages, phase = (idx / hp.b_stages) & 1;
              mbar_wait(&b_empty[stage], phase ^ 1);
              if (lane == 0) mbar_expect_tx(&b_full[stage], b_stage_bytes);
              // packed K order is tap-major inside a segment: k-block = seg base + tap * chunks + ch
              tma_load_2d(smem_b + stage * b_stage_bytes + lane * box_rows * 128, &tmB, &b_full[stage],
                          (kb + tap * sg.chunks + ch) * kBlockK, row);
            }
          kb += sg.chunks * taps;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(Elem<E>::kFmt, kTileM, hp.n_sub, 0, 0);
      uint32_t a_idx = 0, b_idx = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t d0 = tmem_base + acc * 256, d1 = d0 + 128;
        uint32_t first = 1;
        for (int s = 0; s < p.nseg; ++s) {
          const ConvSeg sg = p.seg[s];
          for (int ch = 0; ch < sg.chunks; ++ch) {
            const uint32_t row0 = a_idx;  // ring index of halo row 0 of this chunk
            uint32_t waited = 0;          // rows of this chunk already waited for
            for (int dy = 0; dy < sg.kh; ++dy) {
              while (waited <= static_cast<uint32_t>(dy) + 1) {  // tile 0 needs row dy, tile 1 row dy + 1
                const uint32_t ri = row0 + waited;
                mbar_wait(&a_full[ri % hp.a_rows], (ri / hp.a_rows) & 1);
                ++waited;
              }
              const uint32_t ra = smem_u32(smem_a + ((row0 + dy) % hp.a_rows) * hp.a_row_bytes);
              const uint32_t rb = smem_u32(smem_a + ((row0 + dy + 1) % hp.a_rows) * hp.a_row_bytes);
              for (int dx = 0; dx < sg.kw; ++dx, ++b_idx) {
                const uint32_t stage = b_idx % hp.b_stages;
                mbar_wait(&b_full[stage], (b_idx / hp.b_stages) & 1);
                tcgen05_fence_after();
                const uint64_t bdesc = make_smem_desc_sw128(smem_u32(smem_b + stage * b_stage_bytes), 16, 1024);
                const uint64_t a0 = make_smem_desc_sw128(ra + dx * 128, 16, 1024);
                const uint64_t a1 = make_smem_desc_sw128(rb + dx * 128, 16, 1024);
#pragma unroll
                for (int k = 0; k < kBlockK / 16; ++k)
                  umma_f16(d0, a0 + 2 * k, bdesc + 2 * k, idesc, (first && k == 0) ? 0u : 1u);
#pragma unroll
                for (int k = 0; k < kBlockK / 16; ++k)
                  umma_f16(d1, a1 + 2 * k, bdesc + 2 * k, idesc, (first && k == 0) ? 0u : 1u);
                first = 0;
                umma_commit(&b_empty[stage]);
              }
              umma_commit(&a_empty[(row0 + dy) % hp.a_rows]);  // row dy: last used by this tap row
            }
            umma_commit(&a_empty[(row0 + sg.kh) % hp.a_rows]);
            a_idx += sg.kh + 1;
          }
        }
        umma_commit(&tmem_full[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: warps 4-7 drain tile 0, warps 8-11 tile 1 =====================
    const int q = warp & 3;
    const int tile = (warp - 4) >> 2;
    const int r = q * 32 + lane;
    uint8_t* stg = smem_stg + tile * kStg;
    const int bar_id = 1 + tile;
    const uint32_t x64 = (static_cast<uint32_t>(r) >> 1) & 3u, x32 = (static_cast<uint32_t>(r) >> 2) & 1u;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
      int ns, w0, h0, b;
      decode(unit, ns, w0, h0, b);
      const int hy = h0 + tile, wx = w0 + r;
      const bool valid = (hy < p.H) && (wx < p.W);
      const size_t pix = (static_cast<size_t>(b) * p.H + hy) * p.W + wx;
      mbar_wait(&tmem_full[acc], acc_phase);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + acc * 256 + tile * 128 + (static_cast<uint32_t>(q * 32) << 16);
      if constexpr (EPI == EPI_LSTM) {
        const int nt = ns >> 1, hh = ns & 1;
        const float* bs = bias_s + nt * 256;
#pragma unroll 1
        for (int g2 = 0; g2 < 2; ++g2) {
          const int jl = g2 * 16;            // column offset inside a 32-column gate slice
          const int j0 = hh * 32 + jl;       // hidden channel offset inside the 64-channel block
          const int chan = nt * 64 + j0;
          uint32_t vi[16], vf[16], vo[16], vg[16];
          tmem_ld16(taddr + 0 + jl, vi);
          tmem_ld16(taddr + 32 + jl, vf);
          tmem_ld16(taddr + 64 + jl, vo);
          tmem_ld16(taddr + 96 + jl, vg);
          float cp[16];
          if (p.c_prev != nullptr && valid) {
            const float4* src = reinterpret_cast<const float4*>(p.c_prev + pix * p.ldc + chan);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float4 t = __ldg(src + e);
              cp[4 * e + 0] = t.x, cp[4 * e + 1] = t.y, cp[4 * e + 2] = t.z, cp[4 * e + 3] = t.w;
            }
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) cp[e] = 0.f;
          }
          if (q == 0 && lane < 6) tma_store_wait_read();
          named_bar_sync(bar_id, 128);  // staging of this tile is free again
          tmem_ld_wait();
          float cn[16], hn[16], gi[16], gf[16], go[16], gg[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            lstm_gates_shared_rcp(__uint_as_float(vi[e]) + bs[0 + j0 + e], __uint_as_float(vf[e]) + bs[64 + j0 + e],
                                  __uint_as_float(vo[e]) + bs[128 + j0 + e], __uint_as_float(vg[e]) + bs[192 + j0 + e],
                                  gi[e], gf[e], go[e], gg[e]);
            cn[e] = fmaf(gf[e], cp[e], gi[e] * gg[e]);
          }
#pragma unroll
          for (int e = 0; e < 16; e += 2) {
            float ta, tb;
            tanh_pair_shared_rcp(cn[e], cn[e + 1], ta, tb);
            hn[e] = go[e] * ta;
            hn[e + 1] = go[e + 1] * tb;
          }
          if (g2 == 1) {
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
          }
#pragma unroll
          for (uint32_t j = 0; j < 4; ++j)
            *reinterpret_cast<float4*>(stg + kH2C + r * 64 + ((j ^ x64) << 4)) =
                make_float4(cn[4 * j], cn[4 * j + 1], cn[4 * j + 2], cn[4 * j + 3]);
          auto pack8 = [](const float* v) {
            return make_uint4(Elem<E>::pack2(v[0], v[1]), Elem<E>::pack2(v[2], v[3]), Elem<E>::pack2(v[4], v[5]),
                              Elem<E>::pack2(v[6], v[7]));
          };
#pragma unroll
          for (uint32_t j = 0; j < 2; ++j)
            *reinterpret_cast<uint4*>(stg + kH2H + r * 32 + ((j ^ x32) << 4)) = pack8(hn + 8 * j);
          if (p.gates_boff >= 0) {
#pragma unroll
            for (uint32_t j = 0; j < 2; ++j) {
              *reinterpret_cast<uint4*>(stg + kH2G + 0 * 4096 + r * 32 + ((j ^ x32) << 4)) = pack8(gi + 8 * j);
              *reinterpret_cast<uint4*>(stg + kH2G + 1 * 4096 + r * 32 + ((j ^ x32) << 4)) = pack8(gf + 8 * j);
              *reinterpret_cast<uint4*>(stg + kH2G + 2 * 4096 + r * 32 + ((j ^ x32) << 4)) = pack8(go + 8 * j);
              *reinterpret_cast<uint4*>(stg + kH2G + 3 * 4096 + r * 32 + ((j ^ x32) << 4)) = pack8(gg + 8 * j);
            }
          }
          fence_proxy_async_smem();
          named_bar_sync(bar_id, 128);
          if (q == 0 && lane < 6) {  // one lane per output box; rows beyond the image are clipped by TMA
            if (lane == 0) {
              tma_store_4d(&tmX0, stg + kH2C, chan, w0, hy, b + p.cnext_boff);
            } else if (lane == 1) {
              tma_store_4d(&tmX1, stg + kH2H, chan, w0, hy, b + p.hnext_boff);
            } else if (p.gates_boff >= 0) {
              const int gt = lane - 2;
              tma_store_4d(&tmX2, stg + kH2G + gt * 4096, gt * p.ldc + chan, w0, hy, b + p.gates_boff);
            }
            tma_store_commit();
          }
        }
      } else if constexpr (EPI == EPI_HEAD) {
        // bias + sigmoid, written straight into y (B, C_out, T, H, W): 32 consecutive pixels per warp and channel
        const int groups = hp.n_sub / 16;
        const int t = p.t0 + b / p.b_img, bi = b % p.b_img;
        const size_t plane = static_cast<size_t>(p.H) * p.W;
#pragma unroll 1
        for (int g = 0; g < groups; ++g) {
          uint32_t v[16];
          tmem_ld16(taddr + g * 16, v);
          tmem_ld_wait();
          if (g == groups - 1) {
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
          }
          if (valid) {
            float* ybase = p.y + ((static_cast<size_t>(bi) * p.c_out) * p.t_out + t) * plane +
                           static_cast<size_t>(hy) * p.W + wx;
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const int co = ns * hp.n_sub + g * 16 + e;
              if (co < p.c_out)
                ybase[static_cast<size_t>(co) * p.t_out * plane] = fast_sigmoid(__uint_as_float(v[e]) + bias_s[co]);
            }
          }
        }
      } else {  // EPI_STORE
        const int groups = hp.n_sub / 16;
#pragma unroll 1
        for (int g = 0; g < groups; ++g) {
          uint32_t v[16];
          tmem_ld16(taddr + g * 16, v);
          if (q == 0 && lane == 0) tma_store_wait_read();
          named_bar_sync(bar_id, 128);
          tmem_ld_wait();
          if (g == groups - 1) {
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
          }
#pragma unroll
          for (uint32_t j = 0; j < 4; ++j)
            *reinterpret_cast<float4*>(stg + r * 64 + ((j ^ x64) << 4)) =
                make_float4(__uint_as_float(v[4 * j]) * p.out_scale, __uint_as_float(v[4 * j + 1]) * p.out_scale,
                            __uint_as_float(v[4 * j + 2]) * p.out_scale, __uint_as_float(v[4 * j + 3]) * p.out_scale);
          fence_proxy_async_smem();
          named_bar_sync(bar_id, 128);
          if (q == 0 && lane == 0) {
            const int col = ns * hp.n_sub + g * 16;
            if (col < p.split_col)
              tma_store_4d(&tmX0, stg, col, w0, hy, b);
            else
              tma_store_4d(&tmX1, stg, col - p.split_col, w0, hy, b);
            tma_store_commit();
          }
        }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (q == 0 && lane < 6) tma_store_wait_all();
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace clstm
