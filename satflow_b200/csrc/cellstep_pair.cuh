// Fused ConvLSTM cell step on CTA PAIRS (tcgen05 cta_group::2): the same math, epilogue and K order as
// convgemm_kernel<E, EPI_LSTM> (staged TMA-store epilogue), but two CTAs of a cluster form ONE M = 256 MMA over their
// two 128-pixel tiles and each of them loads only HALF of every [256 x 64] weight k-block; the pair's tensor cores
// read both halves.
//
// Why: the step is bounded by board power (DESIGN.md §4 "board power"), and after the multipliers the biggest
// reducible terms of the cell step are the L2 -> shared-memory operand feed (48 KB per 128x256x64 MMA block, two thirds
// of it weights that every CTA fetches again for every tile) and the shared-memory reads of the MMA itself (A 4 KB +
// B 8 KB per K = 16 instruction).  A pair moves 32 KB per CTA per k-block and reads A 4 KB + B 4 KB per instruction
// and CTA — what cuBLAS's 2-SM kernels do.  A stage is 32 KB instead of 48, so the ring is one stage deeper.
//
// Roles per CTA (384 threads) as in convgemm.cuh: warp 0 = TMA producer (lane 0: the CTA's own pixel tile, lane 1: its
// half of the weight k-block), warp 1 = MMA issuer (LEADER CTA only), warp 2 = TMEM allocator, warps 4..11 = epilogue.
// Barriers: `full` lives in the leader and collects the transaction bytes of BOTH CTAs' loads; `empty` and `tmem_full`
// are signalled in both CTAs by multicast tcgen05.commit; both CTAs' epilogue warps release an accumulator with
// (remote) arrives on the leader's `tmem_empty`.
#pragma once
#include "convgemm.cuh"

namespace clstm {

// ------------------------------------------------------------------ cluster / cta_group::2 PTX
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_smem_addr` in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// TMA loads into the executing CTA's shared memory whose completion bytes are credited to a barrier given as a
// shared::cluster address (the leader's).
__device__ __forceinline__ void tma_load_4d_pair(void* dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0,
                                                 int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, "
      "%5, %6}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0,
                                                 int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
      "[%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrives (once all previously issued MMAs retire) on the barrier at this offset in BOTH CTAs of the pair.
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(static_cast<uint16_t>(3))
      : "memory");
}

constexpr int kPairStageBytes = kABytes + 128 * 128;  // the CTA's [128 px x 64] tile + its [128 x 64] weight half

inline size_t cellstep_pair_smem_bytes(int stages, int n_tiles) {
  return 1024 + static_cast<size_t>(stages) * kPairStageBytes + 2 * static_cast<size_t>(kStgHalfLstm) +
         (2 * kMaxStages + 8) * 8 + 16 + static_cast<size_t>(n_tiles) * 256 * 4 + 64 + kKtabMax * 16 + 16;
}

// tmB: the packed forward weights [4HP rows][K] with 128-row boxes.  tmX0 / tmX1 / tmX2: c, h and gate stacks (16-channel
// epilogue boxes), tmX3: the c stack again (c_prev loads).  p as for convgemm_kernel<E, EPI_LSTM> with staged == 1,
// n_tile == 256 and an EVEN number of tiles per image (a pair never straddles two images, so the K order of a tile
// depends only on where it lies in its image).
template <typename E>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
cellstep_pair_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                     const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmX0,
                     const __grid_constant__ CUtensorMap tmX1, const __grid_constant__ CUtensorMap tmX2,
                     const __grid_constant__ CUtensorMap tmX3, const ConvGemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_stg = smem + p.stages * kPairStageBytes;
  uint8_t* tail = smem_stg + 2 * kStgHalfLstm;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);   // used in the leader only
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tmem_full = empty_bar + kMaxStages;
  uint64_t* tmem_empty = tmem_full + 2;                     // used in the leader only
  uint64_t* cprev_full = tmem_empty + 2;                    // [2], one per epilogue half
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(cprev_full + 4);
  float* bias_s = reinterpret_cast<float*>(tmem_slot + 4);
  int4* ktab = reinterpret_cast<int4*>((reinterpret_cast<uintptr_t>(bias_s + p.n_tiles * 256) + 15) & ~uintptr_t(15));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = (rank == 0);
  const int n_clusters = gridDim.x >> 1;
  const int cluster_id = blockIdx.x >> 1;
  const int num_pairs = (p.num_m_tiles + 1) >> 1;
  const int total_units = num_pairs * p.n_tiles;
  const int tiles_img = p.tiles_w * p.tiles_h;
  int kblocks = 0;
  for (int s = 0; s < p.nseg; ++s) kblocks += p.seg[s].chunks * p.seg[s].kh * p.seg[s].kw;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA0);
    if (p.nseg > 1) tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 16);  // 8 epilogue warps x 2 CTAs
      mbar_init(&cprev_full[a], 1);
    }
    fence_barrier_init();
  }
  // a cta_group::2 allocation is performed by one warp of EACH CTA of the pair (blackwell guide, 'tensor_allocator')
  if (warp == 2) tmem_alloc_pair(tmem_slot, kTmemCols);
  for (int i = threadIdx.x; i < p.n_tiles * 256; i += blockDim.x) bias_s[i] = p.bias[i];
  if (threadIdx.x < kblocks) {
    int kb = threadIdx.x, sgi = 0;
    while (kb >= p.seg[sgi].chunks * p.seg[sgi].kh * p.seg[sgi].kw) kb -= p.seg[sgi].chunks * p.seg[sgi].kh * p.seg[sgi].kw, ++sgi;
    const ConvSeg sg = p.seg[sgi];
    const int ch = kb % sg.chunks, tap = kb / sg.chunks;
    ktab[threadIdx.x] = make_int4(sgi, tap % sg.kw - sg.kw / 2, tap / sg.kw - sg.kh / 2, ch * kBlockK);
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();  // the barriers of both CTAs exist before any remote arrive / multicast commit / peer TMA
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // this CTA's pixel tile of unit u (tile index num_m_tiles of an odd tile count lies past the last image: its loads
  // are out of bounds -> zeros, its stores are clipped)
  auto coords = [&](int unit, int& nt, int& w0, int& h0, int& b, int& kb0) {
    const int pair = unit / p.n_tiles;
    nt = unit % p.n_tiles;
    const int mt = 2 * pair + static_cast<int>(rank);
    w0 = (mt % p.tiles_w) * p.BW;
    h0 = ((mt / p.tiles_w) % p.tiles_h) * p.BH;
    b = mt / tiles_img;
    // one K order per pair, keyed on the position of the pair's first tile inside its image (batch-index independent)
    kb0 = p.rotate ? ((2 * pair) % tiles_img + nt) % kblocks : 0;
  };

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane < 2) {
      int stage = 0;
      uint32_t phase = 0;
      const int boff0 = p.seg[0].b_off, boff1 = p.seg[1].b_off;
      for (int unit = cluster_id; unit < total_units; unit += n_clusters) {
        int nt, w0, h0, b, kb;
        coords(unit, nt, w0, h0, b, kb);
        for (int i = 0; i < kblocks; ++i) {
          const int4 e = ktab[kb];
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* a_dst = smem + stage * kPairStageBytes;
          const uint32_t bar = mapa_u32(smem_u32(&full_bar[stage]), 0);
          if (lane == 0) {
            if (leader) mbar_expect_tx(&full_bar[stage], 2 * kPairStageBytes);
            tma_load_4d_pair(a_dst, e.x ? &tmA1 : &tmA0, bar, e.w, w0 + e.y, h0 + e.z, b + (e.x ? boff1 : boff0));
          } else {
            tma_load_2d_pair(a_dst + kABytes, &tmB, bar, kb * kBlockK, nt * 256 + static_cast<int>(rank) * 128);
          }
          if (++kb == kblocks) kb = 0;
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA) =====================
    if (leader && lane == 0) {
      const uint32_t idesc = make_idesc(Elem<E>::kFmt, 256, 256, 0, 0);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      uint64_t adesc = make_smem_desc_sw128(smem_u32(smem), 16, 1024);
      uint64_t bdesc = make_smem_desc_sw128(smem_u32(smem) + kABytes, 16, 1024);
      for (int unit = cluster_id; unit < total_units; unit += n_clusters) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t d = tmem_base + acc * 256;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) umma_f16_pair(d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          umma_commit_pair(&empty_bar[stage]);  // frees the slot in both CTAs when these MMAs retire
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
          const uint32_t a_addr = smem_u32(smem + stage * kPairStageBytes);
          adesc = make_smem_desc_sw128(a_addr, 16, 1024);
          bdesc = make_smem_desc_sw128(a_addr + kABytes, 16, 1024);
        }
        umma_commit_pair(&tmem_full[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (each CTA drains its own 128 accumulator lanes) =====================
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    const int r = q * 32 + lane;
    uint8_t* stg = smem_stg + half * kStgHalfLstm;
    const bool issuer = (q == 0) && (lane == 0);
    const int bar_id = 1 + half;
    const uint32_t x64 = (static_cast<uint32_t>(r) >> 1) & 3u;
    const uint32_t x32 = (static_cast<uint32_t>(r) >> 2) & 1u;
    const uint32_t empty_remote = mapa_u32(smem_u32(&tmem_empty[0]), 0);
    const bool has_cprev = p.cprev_boff >= 0;
    int acc = 0;
    uint32_t acc_phase = 0, cp_phase = 0;
    if (issuer && has_cprev && cluster_id < total_units) {
      int nt, w0, h0, b, kb;
      coords(cluster_id, nt, w0, h0, b, kb);
      mbar_expect_tx(&cprev_full[half], p.c16 ? 4096 : 8192);
      tma_load_4d(stg + kStgCprev, &tmX3, &cprev_full[half], nt * 64 + half * 32, w0, h0, b + p.cprev_boff);
    }
    for (int unit = cluster_id; unit < total_units; unit += n_clusters) {
      int nt, w0, h0, b, kb_unused;
      coords(unit, nt, w0, h0, b, kb_unused);
      mbar_wait(&tmem_full[acc], acc_phase);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + acc * 256 + (static_cast<uint32_t>(q * 32) << 16);
      const float* bs = bias_s + nt * 256;
#pragma unroll 1
      for (int g2 = 0; g2 < 2; ++g2) {
        const int j0 = half * 32 + g2 * 16;
        uint32_t vi[16], vf[16], vo[16], vg[16];
        tmem_ld16(taddr + 0 + j0, vi);
        tmem_ld16(taddr + 64 + j0, vf);
        tmem_ld16(taddr + 128 + j0, vo);
        tmem_ld16(taddr + 192 + j0, vg);
        float cp[16];
        if (has_cprev) {
          mbar_wait(&cprev_full[half], cp_phase);
          cp_phase ^= 1;
          if (p.c16) {
#pragma unroll
            for (uint32_t j = 0; j < 2; ++j) {  // 32-byte rows, SWIZZLE_32B (like h)
              const uint4 t = *reinterpret_cast<const uint4*>(stg + kStgCprev + r * 32 + ((j ^ x32) << 4));
              const float2 a0 = Elem<E>::unpack2(t.x), a1 = Elem<E>::unpack2(t.y), a2 = Elem<E>::unpack2(t.z),
                           a3 = Elem<E>::unpack2(t.w);
              cp[8 * j + 0] = a0.x * kCScaleInv, cp[8 * j + 1] = a0.y * kCScaleInv, cp[8 * j + 2] = a1.x * kCScaleInv;
              cp[8 * j + 3] = a1.y * kCScaleInv, cp[8 * j + 4] = a2.x * kCScaleInv, cp[8 * j + 5] = a2.y * kCScaleInv;
              cp[8 * j + 6] = a3.x * kCScaleInv, cp[8 * j + 7] = a3.y * kCScaleInv;
            }
          } else {
#pragma unroll
            for (uint32_t j = 0; j < 4; ++j) {
              const float4 t = *reinterpret_cast<const float4*>(stg + kStgCprev + r * 64 + ((j ^ x64) << 4));
              cp[4 * j + 0] = t.x, cp[4 * j + 1] = t.y, cp[4 * j + 2] = t.z, cp[4 * j + 3] = t.w;
            }
          }
        } else {
#pragma unroll
          for (int e = 0; e < 16; ++e) cp[e] = 0.f;
        }
        if (q == 0 && lane < 6) tma_store_wait_read();
        named_bar_sync(bar_id, 128);
        if (issuer && has_cprev) {  // prefetch the next group's c_prev behind this group's math
          int un = unit, ntn = nt, w0n = w0, h0n = h0, bn = b, j0n = j0 + 16, kbn;
          if (g2 == 1) {
            un = unit + n_clusters;
            j0n = half * 32;
            if (un < total_units) coords(un, ntn, w0n, h0n, bn, kbn);
          }
          if (un < total_units) {
            mbar_expect_tx(&cprev_full[half], p.c16 ? 4096 : 8192);
            tma_load_4d(stg + kStgCprev, &tmX3, &cprev_full[half], ntn * 64 + j0n, w0n, h0n, bn + p.cprev_boff);
          }
        }
        tmem_ld_wait();
        float cn[16], hn[16], gi[16], gf[16], go[16], gg[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          lstm_gates_shared_rcp(fmaf(__uint_as_float(vi[e]), kHScaleInv, bs[0 + j0 + e]),
                                fmaf(__uint_as_float(vf[e]), kHScaleInv, bs[64 + j0 + e]),
                                fmaf(__uint_as_float(vo[e]), kHScaleInv, bs[128 + j0 + e]),
                                fmaf(__uint_as_float(vg[e]), kHScaleInv, bs[192 + j0 + e]), gi[e], gf[e], go[e], gg[e]);
          cn[e] = fmaf(gf[e], cp[e], gi[e] * gg[e]);
        }
#pragma unroll
        for (int e = 0; e < 16; e += 2) {
          float ta, tb;
          tanh_pair_shared_rcp(cn[e], cn[e + 1], ta, tb);
          hn[e] = go[e] * ta * kHScale;
          hn[e + 1] = go[e + 1] * tb * kHScale;
        }
        if (g2 == 1) {  // all TMEM reads of this accumulator are done: hand it back to the leader's MMA warp
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(empty_remote + acc * 8);
        }
        auto pack8 = [](const float* v) {
          return make_uint4(Elem<E>::pack2(v[0], v[1]), Elem<E>::pack2(v[2], v[3]), Elem<E>::pack2(v[4], v[5]),
                            Elem<E>::pack2(v[6], v[7]));
        };
        if (p.c16) {
          float cs16[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) cs16[e] = cn[e] * kCScale;
#pragma unroll
          for (uint32_t j = 0; j < 2; ++j)
            *reinterpret_cast<uint4*>(stg + kStgC + r * 32 + ((j ^ x32) << 4)) = pack8(cs16 + 8 * j);
        } else {
#pragma unroll
          for (uint32_t j = 0; j < 4; ++j)
            *reinterpret_cast<float4*>(stg + kStgC + r * 64 + ((j ^ x64) << 4)) =
                make_float4(cn[4 * j], cn[4 * j + 1], cn[4 * j + 2], cn[4 * j + 3]);
        }
#pragma unroll
        for (uint32_t j = 0; j < 2; ++j)
          *reinterpret_cast<uint4*>(stg + kStgH + r * 32 + ((j ^ x32) << 4)) = pack8(hn + 8 * j);
        if (p.gates_boff >= 0) {
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
        if (q == 0 && lane < 6) {
          const int chan = nt * 64 + j0;
          if (lane == 0) {
            tma_store_4d(&tmX0, stg + kStgC, chan, w0, h0, b + p.cnext_boff);
          } else if (lane == 1) {
            tma_store_4d(&tmX1, stg + kStgH, chan, w0, h0, b + p.hnext_boff);
          } else if (p.gates_boff >= 0) {
            const int gt = lane - 2;
            tma_store_4d(&tmX2, stg + kStgG + gt * 4096, gt * p.ldc + chan, w0, h0, b + p.gates_boff);
          }
          tma_store_commit();
        }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (q == 0 && lane < 6) tma_store_wait_all();
  }

  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer may still read this CTA's shared memory / write its TMEM until here
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc_pair(tmem_base, kTmemCols);
  }
}

}  // namespace clstm
