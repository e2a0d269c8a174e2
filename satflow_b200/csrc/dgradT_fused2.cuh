// Transposed dgrad with the NEXT gate gradient fused in, second generation: the gate gradient runs on its OWN warps.
//
// dgradT_fused_kernel (dgradT.cuh) lets the eight epilogue warps do everything in turn: drain 32 accumulator columns
// from TMEM, transpose them through shared memory, store dh_prev, then run the gate-gradient items of those pixels.
// ncu (profiles/r2_fused_dgrad_ncu.md) shows what that costs: the launch is paced by those warps (27 us per 256-pixel
// unit against 14 us of tensor work); while a warp waits for a barrier or drains TMEM its gate-gradient loads are
// not being issued, and while it waits for loads the accumulator is not being drained.
//
// Here the two jobs are split between warp groups that only meet at a shared-memory ring (640 threads per CTA):
//   warps  0..3   control: TMA producer, MMA issuer, TMEM allocator            (setmaxnreg -> 56 registers)
//   warps  4..11  DRAIN   (2 teams x 4 lane quadrants): TMEM -> transposed [32 px][64 ch] fp32 blocks in a
//                 two-buffer staging ring per team, TMA store of dh_prev        (setmaxnreg -> 72 registers)
//   warps 12..19  WORKERS (2 teams x 128 threads): gate gradient of the consumer cell for the staged pixels — they
//                 read dx from the ring, stream gates / c / dc / dh sources from HBM with WSETS register sets of
//                 loads in flight, write dz and dc, and never touch TMEM or a named barrier
//                                                                               (setmaxnreg -> 136 registers)
// Hand-off per (team, buffer): stg_full (the drain issuer arrives after the team's writes) / stg_empty (one arrive per
// worker warp after its reads).  The operand ring has 3 stages of 48 KB (the staging ring takes the fourth's place).
#pragma once
#include "dgradT.cuh"

namespace clstm {

constexpr int kDf2Threads = 640;
constexpr int kDf2StgBuf = 16384;     // [2 blocks (x|h)][32 px][64 ch] fp32
constexpr int kDf2RedBytes = 256 * 5 * 4;

inline size_t dgradTf2_smem_bytes(int stages) {
  return 1024 + static_cast<size_t>(stages) * kDtStageBytes + 4 * static_cast<size_t>(kDf2StgBuf) + kDf2RedBytes +
         (2 * kMaxStages + 4 + 8) * 8 + 16 + 64;
}

// RC: c' of the consumer step is recomputed from its saved gates and c_prev instead of being read (gate_grad_item4).
// S16: the recurrent gradient states live in HBM as 16-bit values times kStateDown (ptx.cuh): the producing cell's dh_prev
// (tmX1 is then a 16-bit map) and the consumer's dc and own dh (f.dc / f.src0 point at E arrays) — half the bytes of four
// of the launch's streams; every sum and the dc recurrence itself stay fp32 in registers.
template <typename E, int WSETS, bool RC = false, bool S16 = false>  // WSETS == 2 (a 4-set variant spilled: removed)
__global__ void __launch_bounds__(kDf2Threads, 1)
dgradT_fused2_kernel(const __grid_constant__ CUtensorMap tmDz, const __grid_constant__ CUtensorMap tmW,
                     const __grid_constant__ CUtensorMap tmX1, const __grid_constant__ CUtensorMap tmG,
                     const __grid_constant__ CUtensorMap tmW2, const DgradTParams p, const GateFuse f) {
  static_assert(WSETS == 2, "the worker loop below is written for two register sets of loads in flight");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_stg = smem + p.stages * kDtStageBytes;           // [team][buffer][kDf2StgBuf]
  float* red = reinterpret_cast<float*>(smem_stg + 4 * kDf2StgBuf);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_stg + 4 * kDf2StgBuf + kDf2RedBytes);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tmem_full = empty_bar + kMaxStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* stg_full = tmem_empty + 2;   // [team * 2 + buffer]
  uint64_t* stg_empty = stg_full + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(stg_empty + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_units = (p.num_m_tiles + 1) >> 1;  // 64 x-channels + 64 h-channels, two pixel tiles per unit
  const int taps = p.seg.kh * p.seg.kw;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmDz);
    tma_prefetch_desc(&tmW);
    if (p.extra) {
      tma_prefetch_desc(&tmG);
      tma_prefetch_desc(&tmW2);
    }
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 8);  // one arrive per drain warp
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&stg_full[i], 1);    // the team's drain issuer
      mbar_init(&stg_empty[i], 4);   // one arrive per worker warp of the team
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, kTmemCols);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto tile_origin = [&](int unit, int t, int& w0, int& h0, int& b) {
    const int mt = 2 * unit + t;
    w0 = (mt % p.tiles_w) * p.BW;
    h0 = ((mt / p.tiles_w) % p.tiles_h) * p.BH;
    b = mt / (p.tiles_w * p.tiles_h);
  };

  if (warp < 4) {
    // ================================================================= control warpgroup
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;" ::: "memory");
    if (warp == 0) {
      if (lane < 3) {  // lane 0: weight box, lanes 1-2: the two pixel tiles
        int stage = 0;
        uint32_t phase = 0;
        for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
          int w0 = 0, h0 = 0, b = 0;
          if (lane > 0) tile_origin(unit, lane - 1, w0, h0, b);
          const int kblocks = taps * p.seg.chunks;
          int kb = p.rotate ? unit % kblocks : 0;
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
          // second K segment (p.extra k-blocks, no taps): the head's dlogit "col" tensor of the consumer's output
          // frame against the head dgrad weights, so dx already contains the head's share of the consumer's dh
          for (int j = 0; j < p.extra; ++j) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* dst = smem + stage * kDtStageBytes;
            if (lane == 0) {
              mbar_expect_tx(&full_bar[stage], kDtStageBytes);
              tma_load_2d(dst, &tmW2, &full_bar[stage], j * kBlockK, 0);
            } else {
              tma_load_4d(dst + 16384 + (lane - 1) * kABytes, &tmG, &full_bar[stage], j * kBlockK, w0, h0, b);
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
        const int kblocks = taps * p.seg.chunks + p.extra;
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
            const uint32_t base = smem_u32(smem + stage * kDtStageBytes);
            adesc = make_smem_desc_sw128(base, 16, 1024);
            bdesc = make_smem_desc_sw128(base + 16384, 16, 1024);
          }
          umma_commit(&tmem_full[acc]);
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1;
        }
      }
    }
  } else if (warp < 12) {
    // ================================================================= drain warps: TMEM -> staging ring -> dh_prev
    asm volatile("setmaxnreg.dec.sync.aligned.u32 72;" ::: "memory");
    const int q = warp & 3;
    const int team = (warp - 4) >> 2;  // == pixel tile of the unit
    const int cl = q * 32 + lane;      // output channel (TMEM lane): 0..63 x part, 64..127 h part
    const int bar_id = 1 + team;
    const bool issuer = (q == 0) && (lane == 0);
    uint8_t* stg_team = smem_stg + team * 2 * kDf2StgBuf;
    uint32_t gcount = 0;  // groups handled so far by this team (selects the buffer and the barrier parities)
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
      int w0, h0, b;
      tile_origin(unit, team, w0, h0, b);
      mbar_wait(&tmem_full[acc], acc_phase);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + acc * 256 + team * 128 + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
      for (int g = 0; g < 4; ++g, ++gcount) {
        const uint32_t bsel = gcount & 1, use = gcount >> 1;
        float* stg = reinterpret_cast<float*>(stg_team + bsel * kDf2StgBuf);
        uint32_t v[32];
        tmem_ld16(taddr + g * 32, *reinterpret_cast<uint32_t(*)[16]>(v));
        tmem_ld16(taddr + g * 32 + 16, *reinterpret_cast<uint32_t(*)[16]>(v + 16));
        // the buffer is free when the workers have read it (stg_empty) and the TMA store issued from it two groups
        // ago has finished reading it (only the issuer can know: it tells the team through the named barrier)
        mbar_wait(&stg_empty[team * 2 + bsel], (use & 1) ^ 1);
        if (issuer) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        named_bar_sync(bar_id, 128);
        tmem_ld_wait();
        if (g == 3) {
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
        if (S16 && (cl >> 6) == f.h_block) {  // dh_prev: [32 px][64 ch] 16-bit in the first half of its block
          E* dst = reinterpret_cast<E*>(stg + f.h_block * (32 * 64)) + (cl & 63);
#pragma unroll
          for (int j = 0; j < 32; ++j) dst[j * 64] = Elem<E>::from_float(__uint_as_float(v[j]) * kStateDown);
        } else {
          float* dst = stg + (cl >> 6) * (32 * 64) + (cl & 63);
#pragma unroll
          for (int j = 0; j < 32; ++j) dst[j * 64] = __uint_as_float(v[j]);
        }
        fence_proxy_async_smem();
        named_bar_sync(bar_id, 128);
        if (issuer) {
#pragma unroll
          for (int s2 = 0; s2 < 2; ++s2) {  // dh_prev of the producing cell (64 channels), 16-pixel boxes
            const int px = g * 32 + s2 * 16;
            const float* hb = stg + f.h_block * (32 * 64);
            const void* src = S16 ? static_cast<const void*>(reinterpret_cast<const E*>(hb) + s2 * 16 * 64)
                                  : static_cast<const void*>(hb + s2 * 16 * 64);
            tma_store_4d(&tmX1, src, 0, w0 + (px & (p.BW - 1)), h0 + (px >> p.lbw), b);
          }
          tma_store_commit();
          mbar_arrive(&stg_full[team * 2 + bsel]);  // ordered after every drain thread's writes by the barrier above
        }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (issuer) tma_store_wait_all();
  } else {
    // ================================================================= gate-gradient workers
    asm volatile("setmaxnreg.inc.sync.aligned.u32 136;" ::: "memory");
    const int wt = threadIdx.x - 384;   // 0..255
    const int team = wt >> 7;           // pixel tile of the unit
    const int cl = wt & 127;            // thread inside the team
    const int chunk = cl & 15;          // 4-channel chunk of the consumer's hidden channels
    const int pxl = cl >> 4;            // 0..7: pixel inside an 8-pixel pass
    const uint8_t* stg_team = smem_stg + team * 2 * kDf2StgBuf;
    const E* gates_c = static_cast<const E*>(f.gates) + chunk * 4;
    E* dzo_c = static_cast<E*>(f.dz_out) + chunk * 4;
    const float* cp_c = f.c_prev ? f.c_prev + chunk * 4 : nullptr;
    const float* cn_c = f.c_next + chunk * 4;
    float* dc_c = f.dc + chunk * 4;
    const float* s0_c = f.src0 ? f.src0 + chunk * 4 : nullptr;
    const float* s1_c = f.src1 ? f.src1 + chunk * 4 : nullptr;
    const int bwm = p.BW - 1;
    float bsum[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int e = 0; e < 4; ++e) bsum[a][e] = 0.f;
    uint32_t zmax = 0;

    struct Raw {
      uint2 g[4];
      float4 cp, cn, dc, s0, s1;
      uint2 dc16, s016;  // S16: the packed forms of dc / s0
      uint2 cp16;        // f.c16: the packed form of c_prev
      unsigned pix;  // global pixel index, or 0xFFFFFFFF for an item outside the image / past the last unit
    };
    struct TileBase {
      unsigned pix0;
      int rows, cols;
    };
    auto tile_base = [&](int unit) -> TileBase {
      TileBase t{0u, 0, 0};
      if (unit >= total_units) return t;
      int w0, h0, b;
      tile_origin(unit, team, w0, h0, b);
      if (b >= p.B) return t;
      t.pix0 = (static_cast<unsigned>(b) * p.H + h0) * p.W + w0;
      t.rows = p.H - h0;
      t.cols = p.W - w0;
      return t;
    };
    auto issue = [&](Raw& r, const TileBase& tb, int g, int s) {
      const int px = g * 32 + s * 8 + pxl;
      const int lx = px & bwm, ly = px >> p.lbw;
      if (!(ly < tb.rows && lx < tb.cols)) {
        r.pix = 0xFFFFFFFFu;
        return;
      }
      r.pix = tb.pix0 + ly * p.W + lx;
      const unsigned o4 = r.pix * (4 * 64), o1 = r.pix * 64;  // HP == 64
      r.g[0] = __ldg(reinterpret_cast<const uint2*>(gates_c + o4));
      r.g[1] = __ldg(reinterpret_cast<const uint2*>(gates_c + o4 + 64));
      r.g[2] = __ldg(reinterpret_cast<const uint2*>(gates_c + o4 + 128));
      r.g[3] = __ldg(reinterpret_cast<const uint2*>(gates_c + o4 + 192));
      if (f.c16)
        r.cp16 = f.c_prev ? __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const E*>(f.c_prev) + chunk * 4 + o1))
                          : make_uint2(0u, 0u);
      else
        r.cp = cp_c ? __ldg(reinterpret_cast<const float4*>(cp_c + o1)) : make_float4(0.f, 0.f, 0.f, 0.f);
      if constexpr (!RC) r.cn = __ldg(reinterpret_cast<const float4*>(cn_c + o1));
      if constexpr (S16) {
        r.dc16 = *reinterpret_cast<const uint2*>(reinterpret_cast<const E*>(f.dc) + chunk * 4 + o1);
        if (s0_c) r.s016 = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const E*>(f.src0) + chunk * 4 + o1));
      } else {
        r.dc = *reinterpret_cast<const float4*>(dc_c + o1);  // read and written by this thread only
        if (s0_c) r.s0 = __ldg(reinterpret_cast<const float4*>(s0_c + o1));
      }
      if (s1_c) r.s1 = __ldg(reinterpret_cast<const float4*>(s1_c + o1));
    };
    auto consume = [&](const Raw& r, const float* stg, int s) {
      if (r.pix == 0xFFFFFFFFu) return;
      float dhv[4] = {0.f, 0.f, 0.f, 0.f};
      if (f.use_stg) {
        const float4 a0 = *reinterpret_cast<const float4*>(stg + (s * 8 + pxl) * 64 + chunk * 4);
        dhv[0] = a0.x, dhv[1] = a0.y, dhv[2] = a0.z, dhv[3] = a0.w;
      }
      float4 cpv = r.cp;
      if (f.c16) {
        const float2 a = Elem<E>::unpack2(r.cp16.x), b2 = Elem<E>::unpack2(r.cp16.y);
        cpv = make_float4(a.x * kCScaleInv, a.y * kCScaleInv, b2.x * kCScaleInv, b2.y * kCScaleInv);
      }
      float4 dcin;
      if constexpr (S16) {
        const float2 a = Elem<E>::unpack2(r.dc16.x), b2 = Elem<E>::unpack2(r.dc16.y);
        dcin = make_float4(a.x * kStateUp, a.y * kStateUp, b2.x * kStateUp, b2.y * kStateUp);
        if (s0_c) {
          const float2 c0 = Elem<E>::unpack2(r.s016.x), c1 = Elem<E>::unpack2(r.s016.y);
          dhv[0] = fmaf(c0.x, kStateUp, dhv[0]), dhv[1] = fmaf(c0.y, kStateUp, dhv[1]);
          dhv[2] = fmaf(c1.x, kStateUp, dhv[2]), dhv[3] = fmaf(c1.y, kStateUp, dhv[3]);
        }
      } else {
        dcin = r.dc;
        if (s0_c) dhv[0] += r.s0.x, dhv[1] += r.s0.y, dhv[2] += r.s0.z, dhv[3] += r.s0.w;
      }
      if (s1_c) dhv[0] += r.s1.x, dhv[1] += r.s1.y, dhv[2] += r.s1.z, dhv[3] += r.s1.w;
      float4 dcn;
      uint2 dzp[4];
      gate_grad_item4<E, RC>(r.g, cpv, r.cn, dcin, dhv, bsum, zmax, dcn, dzp);
      const unsigned o4 = r.pix * (4 * 64), o1 = r.pix * 64;
      if constexpr (S16)
        *reinterpret_cast<uint2*>(reinterpret_cast<E*>(f.dc) + chunk * 4 + o1) =
            make_uint2(Elem<E>::pack2(dcn.x * kStateDown, dcn.y * kStateDown), Elem<E>::pack2(dcn.z * kStateDown, dcn.w * kStateDown));
      else
        *reinterpret_cast<float4*>(dc_c + o1) = dcn;
#pragma unroll
      for (int a = 0; a < 4; ++a) *reinterpret_cast<uint2*>(dzo_c + o4 + a * 64) = dzp[a];
    };

    TileBase tb_cur = tile_base(blockIdx.x), tb_next;
    Raw rs[WSETS];
#pragma unroll
    for (int i = 0; i < WSETS; ++i) issue(rs[i], tb_cur, 0, i);  // the first items of group 0
    uint32_t gcount = 0;
    for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
      tb_next = tile_base(unit + static_cast<int>(gridDim.x));
#pragma unroll 1
      for (int g = 0; g < 4; ++g, ++gcount) {
        const uint32_t bsel = gcount & 1, use = gcount >> 1;
        const float* stg = reinterpret_cast<const float*>(stg_team + bsel * kDf2StgBuf);
        const int ng = (g + 1) & 3;
        const TileBase& tbn = (g == 3) ? tb_next : tb_cur;
        mbar_wait(&stg_full[team * 2 + bsel], use & 1);  // this group's dx block is staged
        {  // two sets: item s uses set s & 1, loads run two items ahead
          consume(rs[0], stg, 0);
          issue(rs[0], tb_cur, g, 2);
          consume(rs[1], stg, 1);
          issue(rs[1], tb_cur, g, 3);
          consume(rs[0], stg, 2);
          issue(rs[0], tbn, ng, 0);
          consume(rs[1], stg, 3);
          __syncwarp();
          if (lane == 0) mbar_arrive(&stg_empty[team * 2 + bsel]);
          issue(rs[1], tbn, ng, 1);
        }
      }
      tb_cur = tb_next;
    }
    fold_absmax<E>(zmax, f.dz_absmax);
    // bias partial sums: reduce over the 16 threads (both teams) that share a channel chunk, one gate at a time
    const int HP = f.HP;
#pragma unroll
    for (int a = 0; a < 4; ++a) {  // unrolled: bsum must stay in registers
      named_bar_sync(3, 256);
#pragma unroll
      for (int e = 0; e < 4; ++e) red[wt * 5 + e] = bsum[a][e];
      named_bar_sync(3, 256);
      if (wt < 16) {
        for (int e = 0; e < 4; ++e) {
          float sum = 0.f;
          for (int pl = 0; pl < 16; ++pl) sum += red[(pl * 16 + wt) * 5 + e];
          f.bias_partial[static_cast<size_t>(blockIdx.x) * 4 * HP + a * HP + wt * 4 + e] += sum;
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
