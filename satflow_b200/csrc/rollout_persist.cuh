// Persistent ConvLSTM rollout for shapes whose recurrent state fits on chip (sm_100a).
//
// conv_lstm.py:176-196 is a strictly sequential chain of cell steps; launched one kernel per step (convgemm.cuh) a
// small image pays a launch, a TMEM allocation, barrier initialisation and a c_prev / c_next round trip through HBM
// for a few microseconds of tensor work.  When every (128-pixel tile, 64-hidden-channel slice) of the image batch
// can be owned by its own CTA (tiles <= SMs), ONE cooperative launch runs the whole encoder + decoder chain:
//
//   * a CTA owns its tile for ALL steps of ALL cells;
//   * the cell state c never leaves the SM: cell k's [128 px x 64 ch] fp32 tile lives in TMEM columns
//     [256 + 64k, 256 + 64k + 64) next to the 256-column gate accumulator, read with tcgen05.ld and written back
//     with tcgen05.st by the thread that owns the pixel (c needs no halo, so no other CTA ever wants it);
//   * h does need a halo (the next step's 3x3 taps reach into neighbouring tiles): it is written with TMA stores to
//     its slot of the plan's h stack (L2 resident at these sizes) and a grid-wide hand-off — a release-add on a
//     global counter by the two store lanes of every CTA, an acquire-spin by the TMA producer lane — orders step
//     s + 1's loads behind step s's stores.  The weight loads of step s + 1 do not wait for it;
//   * training plans also stream the gates and c of every step to HBM (the saved activations of BPTT); inference
//     plans write nothing but h.
//
// Warp roles as in convgemm.cuh (384 threads): warp 0 TMA producer (lane 0 activations, lane 1 weights), warp 1 MMA
// issuer, warp 2 TMEM allocator, warps 4..11 epilogue (two per lane quadrant, splitting the 64 hidden channels).
#pragma once
#include "convgemm.cuh"

namespace clstm {

constexpr int kPersistMaxCells = 4;   // 4 x 64 TMEM columns of cell state next to the 256-column accumulator
constexpr int kPersistMaxMaps = 32;

struct PersistCell {
  ConvSeg seg[2];   // [0] the cell's input (x im2col or the h of another cell), [1] its own h; b_off unused here
  int kblocks;
  int kb_first;     // k-blocks of seg[0] alone: a cell's first step sees h == 0, so its h segment is skipped (the same
                    // K range and rotation as the per-step launch with nseg = 1 -> bit-identical results)
  int map_a1, map_b, map_xc, map_xh, map_xg;  // indices into the tensor-map table (map_xc / map_xg: training only)
  const float* bias;                          // packed [n_tiles * 256]
};

struct PersistStep {
  int cell;
  int map_a0;       // tensor map of the input segment
  int a0_boff;      // image offsets (time slot * B) into the stacks
  int a1_boff;
  int hnext_boff;
  int cnext_boff;   // training only
  int gates_boff;   // training only
  int first;        // 1: the cell's state is still zero (its first step)
  int store_c;      // 1: also write c' to its HBM slot (every step of a training plan; else the cell's last step)
};

struct PersistParams {
  int B, H, W, BW, BH, tiles_w, tiles_h, num_m_tiles, n_tiles;
  int ldc;          // padded hidden channels
  int stages, nsteps, ncell, training, rotate;
  const CUtensorMap* maps;   // device memory, 64-byte aligned entries
  const PersistStep* steps;  // device memory
  unsigned int* counter;     // zeroed before the launch; += 2 per CTA per step
  PersistCell cells[kPersistMaxCells];
};

inline size_t persist_smem_bytes(int stages, int ncell) {
  return 1024 + static_cast<size_t>(stages) * (kABytes + 256 * 128) + 2 * static_cast<size_t>(kStgHalfLstm) +
         (2 * kMaxStages + 8) * 8 + 16 + static_cast<size_t>(ncell) * 256 * 4 + 64 +
         static_cast<size_t>(ncell) * kKtabMax * 16 + 16;
}

template <typename E>
__global__ void __launch_bounds__(kGemmThreads, 1) rollout_persist_kernel(const PersistParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int stage_bytes = kABytes + 256 * 128;
  uint8_t* smem_stg = smem + p.stages * stage_bytes;
  uint8_t* tail = smem_stg + 2 * kStgHalfLstm;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tmem_full = empty_bar + kMaxStages;   // [1] accumulator complete
  uint64_t* tmem_empty = tmem_full + 2;           // [1] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 6);
  float* bias_s = reinterpret_cast<float*>(tmem_slot + 4);  // [ncell][256]: this CTA's N tile of every cell
  int4* ktab = reinterpret_cast<int4*>((reinterpret_cast<uintptr_t>(bias_s + p.ncell * 256) + 15) & ~uintptr_t(15));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // this CTA's tile, fixed for the whole rollout
  const int mt = blockIdx.x / p.n_tiles, nt = blockIdx.x % p.n_tiles;
  const int w0 = (mt % p.tiles_w) * p.BW;
  const int h0 = ((mt / p.tiles_w) % p.tiles_h) * p.BH;
  const int b = mt / (p.tiles_w * p.tiles_h);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&tmem_full[0], 1);
    mbar_init(&tmem_empty[0], 8);  // one arrive per epilogue warp
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, kTmemCols);
  for (int i = threadIdx.x; i < p.ncell * 256; i += blockDim.x) {
    const PersistCell& c = p.cells[i >> 8];
    bias_s[i] = c.bias ? c.bias[nt * 256 + (i & 255)] : 0.f;
  }
  for (int i = threadIdx.x; i < p.ncell * kKtabMax; i += blockDim.x) {
    const PersistCell& c = p.cells[i / kKtabMax];
    int kb = i % kKtabMax;
    if (kb < c.kblocks) {
      int sgi = 0;
      while (kb >= c.seg[sgi].chunks * c.seg[sgi].kh * c.seg[sgi].kw) kb -= c.seg[sgi].chunks * c.seg[sgi].kh * c.seg[sgi].kw, ++sgi;
      const ConvSeg sg = c.seg[sgi];
      const int ch = kb % sg.chunks, tap = kb / sg.chunks;
      ktab[i] = make_int4(sgi, tap % sg.kw - sg.kw / 2, tap / sg.kw - sg.kh / 2, ch * kBlockK);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const unsigned int per_step = 2u * gridDim.x;  // hand-off arrivals per step (two store lanes per CTA)

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane < 2) {
      int stage = 0;
      uint32_t phase = 0;
      for (int s = 0; s < p.nsteps; ++s) {
        const PersistStep st = p.steps[s];
        const PersistCell& c = p.cells[st.cell];
        const int4* kt = ktab + st.cell * kKtabMax;
        const CUtensorMap* mapA0 = p.maps + st.map_a0;
        const CUtensorMap* mapA1 = p.maps + c.map_a1;
        const CUtensorMap* mapB = p.maps + c.map_b;
        if (lane == 0 && s > 0) {
          // step s reads h tiles (with their halo) that other CTAs stored during step s - 1
          const unsigned int target = per_step * static_cast<unsigned int>(s);
          uint32_t spins = 0;
          while (ld_acquire_gpu(p.counter) < target) {
            if (++spins > (1u << 23)) {
              printf("clstm: persistent rollout hand-off timed out (block %d step %d)\n", blockIdx.x, s);
              __trap();
            }
          }
          fence_proxy_async_all();
        }
        const int nkb = st.first ? c.kb_first : c.kblocks;
        int kb = (p.rotate && nkb > 1) ? (((mt % (p.tiles_w * p.tiles_h)) & ~1) + nt) % nkb : 0;  // as convgemm.cuh
        for (int i = 0; i < nkb; ++i) {
          const int4 e = kt[kb];
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* a_dst = smem + stage * stage_bytes;
          if (lane == 0) {
            mbar_expect_tx(&full_bar[stage], stage_bytes);
            tma_load_4d(a_dst, e.x ? mapA1 : mapA0, &full_bar[stage], e.w, w0 + e.y, h0 + e.z,
                        b + (e.x ? st.a1_boff : st.a0_boff));
          } else {
            tma_load_2d(a_dst + kABytes, mapB, &full_bar[stage], kb * kBlockK, nt * 256);
          }
          if (++kb == nkb) kb = 0;
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one accumulator: the chain is serial) =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(Elem<E>::kFmt, kTileM, 256, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      uint64_t adesc = make_smem_desc_sw128(smem_u32(smem), 16, 1024);
      uint64_t bdesc = make_smem_desc_sw128(smem_u32(smem) + kABytes, 16, 1024);
      for (int s = 0; s < p.nsteps; ++s) {
        const int kblocks = p.steps[s].first ? p.cells[p.steps[s].cell].kb_first : p.cells[p.steps[s].cell].kblocks;
        mbar_wait(&tmem_empty[0], (static_cast<uint32_t>(s) & 1) ^ 1);
        tcgen05_fence_after();
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) umma_f16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          umma_commit(&empty_bar[stage]);
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
          const uint32_t a_addr = smem_u32(smem + stage * stage_bytes);
          adesc = make_smem_desc_sw128(a_addr, 16, 1024);
          bdesc = make_smem_desc_sw128(a_addr + kABytes, 16, 1024);
        }
        umma_commit(&tmem_full[0]);
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    const int r = q * 32 + lane;  // pixel within the tile == TMEM lane
    uint8_t* stg = smem_stg + half * kStgHalfLstm;
    const int bar_id = 1 + half;
    const uint32_t x64 = (static_cast<uint32_t>(r) >> 1) & 3u;
    const uint32_t x32 = (static_cast<uint32_t>(r) >> 2) & 1u;
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    const bool store_lane = (q == 0) && (lane < 6);
    for (int s = 0; s < p.nsteps; ++s) {
      const PersistStep st = p.steps[s];
      const PersistCell& c = p.cells[st.cell];
      const float* bs = bias_s + st.cell * 256;
      const uint32_t c_cols = tmem_base + 256 + st.cell * 64 + lane_base;
      const CUtensorMap* mapC = p.maps + c.map_xc;
      const CUtensorMap* mapH = p.maps + c.map_xh;
      const CUtensorMap* mapG = p.maps + c.map_xg;
      mbar_wait(&tmem_full[0], static_cast<uint32_t>(s) & 1);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + lane_base;
#pragma unroll 1
      for (int g2 = 0; g2 < 2; ++g2) {
        const int j0 = half * 32 + g2 * 16;
        uint32_t vi[16], vf[16], vo[16], vg[16], vc[16];
        tmem_ld16(taddr + 0 + j0, vi);
        tmem_ld16(taddr + 64 + j0, vf);
        tmem_ld16(taddr + 128 + j0, vo);
        tmem_ld16(taddr + 192 + j0, vg);
        if (!st.first) tmem_ld16(c_cols + j0, vc);
        if (store_lane) tma_store_wait_read();  // this lane's previous store has finished reading the staging
        named_bar_sync(bar_id, 128);
        tmem_ld_wait();
        float cn[16], hn[16], gi[16], gf[16], go[16], gg[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          lstm_gates_shared_rcp(fmaf(__uint_as_float(vi[e]), kHScaleInv, bs[0 + j0 + e]),
                                fmaf(__uint_as_float(vf[e]), kHScaleInv, bs[64 + j0 + e]),
                                fmaf(__uint_as_float(vo[e]), kHScaleInv, bs[128 + j0 + e]),
                                fmaf(__uint_as_float(vg[e]), kHScaleInv, bs[192 + j0 + e]), gi[e], gf[e], go[e], gg[e]);
          const float cp = st.first ? 0.f : __uint_as_float(vc[e]);
          cn[e] = fmaf(gf[e], cp, gi[e] * gg[e]);
        }
#pragma unroll
        for (int e = 0; e < 16; e += 2) {
          float ta, tb;
          tanh_pair_shared_rcp(cn[e], cn[e + 1], ta, tb);
          hn[e] = go[e] * ta * kHScale;  // stored scaled, see kHScale
          hn[e + 1] = go[e + 1] * tb * kHScale;
        }
        // the state stays on chip: c' back into this cell's TMEM columns (same thread reads it at the cell's next step)
#pragma unroll
        for (int e = 0; e < 16; ++e) vc[e] = __float_as_uint(cn[e]);
        tmem_st16(c_cols + j0, vc);
        if (g2 == 1) {  // every TMEM read of the accumulator is done: hand it back to the MMA warp
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[0]);
        }
        auto pack8 = [](const float* v) {
          return make_uint4(Elem<E>::pack2(v[0], v[1]), Elem<E>::pack2(v[2], v[3]), Elem<E>::pack2(v[4], v[5]),
                            Elem<E>::pack2(v[6], v[7]));
        };
#pragma unroll
        for (uint32_t j = 0; j < 2; ++j)
          *reinterpret_cast<uint4*>(stg + kStgH + r * 32 + ((j ^ x32) << 4)) = pack8(hn + 8 * j);
        if (st.store_c) {
#pragma unroll
          for (uint32_t j = 0; j < 4; ++j)
            *reinterpret_cast<float4*>(stg + kStgC + r * 64 + ((j ^ x64) << 4)) =
                make_float4(cn[4 * j], cn[4 * j + 1], cn[4 * j + 2], cn[4 * j + 3]);
        }
        if (p.training) {
#pragma unroll
          for (int e = 0; e < 16; ++e) gi[e] -= kGateCenter, gf[e] -= kGateCenter, go[e] -= kGateCenter;  // stored centred
#pragma unroll
          for (uint32_t j = 0; j < 2; ++j) {
            *reinterpret_cast<uint4*>(stg + kStgG + 0 * 4096 + r * 32 + ((j ^ x32) << 4)) = pack8(gi + 8 * j);
            *reinterpret_cast<uint4*>(stg + kStgG + 1 * 4096 + r * 32 + ((j ^ x32) << 4)) = pack8(gf + 8 * j);
            *reinterpret_cast<uint4*>(stg + kStgG + 2 * 4096 + r * 32 + ((j ^ x32) << 4)) = pack8(go + 8 * j);
            *reinterpret_cast<uint4*>(stg + kStgG + 3 * 4096 + r * 32 + ((j ^ x32) << 4)) = pack8(gg + 8 * j);
          }
        }
        fence_proxy_async_smem();
        named_bar_sync(bar_id, 128);
        if (store_lane) {
          const int chan = nt * 64 + j0;
          if (lane == 1) {
            tma_store_4d(mapH, stg + kStgH, chan, w0, h0, b + st.hnext_boff);
          } else if (lane == 0) {
            if (st.store_c) tma_store_4d(mapC, stg + kStgC, chan, w0, h0, b + st.cnext_boff);
          } else if (p.training) {
            const int gt = lane - 2;
            tma_store_4d(mapG, stg + kStgG + gt * 4096, gt * p.ldc + chan, w0, h0, b + st.gates_boff);
          }
          tma_store_commit();
        }
        tmem_st_wait();
      }
      if (q == 0 && lane == 1) {
        // hand-off: this half's h stores of step s are complete and visible device-wide
        tma_store_wait_all();
        fence_proxy_async_all();
        __threadfence();
        red_release_gpu_add(p.counter, 1u);
      }
    }
    if (store_lane) tma_store_wait_all();
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace clstm
