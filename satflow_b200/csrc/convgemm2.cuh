// Second-generation pixel-major implicit-GEMM convolution: CTA pairs (tcgen05 cta_group::2) with a
// halo-stationary A operand.
//
// convgemm.cuh loads one [128 px x 64 ch] A tile AND one [n_tile x 64] weight tile per (tap, chunk)
// k-block: 48 KB of L2->SMEM traffic per 512 tensor cycles, which is what bounds it.
// Here
//   * a CTA's pixel tile is 128 consecutive pixels of ONE image row; for each 64-channel chunk the
//     kh halo rows ([128 + kw - 1] px each) are loaded ONCE and every tap (dy, dx) is a tcgen05
//     shared-memory descriptor whose start address is shifted by (dy * pitch + dx) * 128 B
//     (validated by clstm_selftest_shifted_desc: the 128-B swizzle is a function of the absolute
//     shared-memory address, so row-shifted views of a TMA-written tile read correctly);
//   * two CTAs of a cluster form one M = 256 MMA (their two pixel tiles) and each loads only HALF of
//     every weight tile (N/2 rows); the pair's tensor cores read both halves (cta_group::2).
// L2->SMEM bytes per 128-px tile for K = (64 + 64) * 9, N = 256: 100 KB (A halos) + 288 KB (B halves)
// instead of 288 KB + 576 KB.
//
// Roles per CTA (384 threads): warp 0 = A-halo TMA producer, warp 3 = B TMA producer, warp 1 = MMA
// issuer (leader CTA only), warp 2 = TMEM allocator, warps 4..11 = epilogue (shared with convgemm.cuh).
// Full barriers live in the LEADER CTA and collect the transaction bytes of both CTAs' TMA loads;
// empty / tmem_full barriers are signalled in both CTAs by multicast tcgen05.commit; the non-leader's
// epilogue releases accumulators by remote arrives on the leader's tmem_empty barrier.
#pragma once
#include "convgemm.cuh"

namespace clstm {

constexpr int kPairMaxAStages = 8;
constexpr int kPairMaxBStages = 8;

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
// TMA loads whose completion bytes are credited to a barrier given as a shared::cluster address
// (the leader CTA's), destination in the executing CTA's shared memory.
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

struct PairGemmParams {
  ConvGemmParams g;    // geometry (BW == 128, BH == 1), segments, epilogue pointers
  int a_stages;        // halo slots
  int b_stages;        // weight half-tile stages
  int a_slot_bytes;    // max over segments of kh * pitch * 128
  int pitch[2];        // halo row pitch in pixels per segment (multiple of 8)
  int halo_w[2];       // 128 + kw - 1 per segment (TMA box width)
  int halo;            // 1: halo-stationary A (one load per chunk, taps = shifted descriptors);
                       // 0: one [128 px x 64 ch] A load per (tap, chunk) like convgemm.cuh
};

inline size_t pairgemm_smem_bytes(int a_stages, int a_slot_bytes, int b_stages, int n_tile, int n_tiles) {
  return 1024 + static_cast<size_t>(a_stages) * a_slot_bytes + static_cast<size_t>(b_stages) * (n_tile / 2) * 128 +
         (2 * kPairMaxAStages + 2 * kPairMaxBStages + 4) * 8 + 16 + static_cast<size_t>(n_tiles) * n_tile * 4 + 64;
}

template <typename E, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
pairgemm_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                const __grid_constant__ CUtensorMap tmB, const PairGemmParams pp) {
  const ConvGemmParams& p = pp.g;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int b_stage_bytes = (p.n_tile / 2) * 128;
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + pp.a_stages * pp.a_slot_bytes;
  uint8_t* tail = smem_b + pp.b_stages * b_stage_bytes;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(tail);
  uint64_t* a_empty = a_full + kPairMaxAStages;
  uint64_t* b_full = a_empty + kPairMaxAStages;
  uint64_t* b_empty = b_full + kPairMaxBStages;
  uint64_t* tmem_full = b_empty + kPairMaxBStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* bias_s = reinterpret_cast<float*>(tmem_slot + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = (rank == 0);
  const int n_clusters = gridDim.x >> 1;
  const int cluster_id = blockIdx.x >> 1;
  const int num_pairs = (p.num_m_tiles + 1) >> 1;
  const int total_units = num_pairs * p.n_tiles;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA0);
    if (p.nseg > 1) tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < pp.a_stages; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < pp.b_stages; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 16);  // 8 epilogue warps x 2 CTAs (only the leader's is waited on)
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_pair(tmem_slot, kTmemCols);
  if (p.bias != nullptr) {
    for (int i = threadIdx.x; i < p.n_tiles * p.n_tile; i += blockDim.x) bias_s[i] = p.bias[i];
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();  // barriers of both CTAs are initialised before any remote arrive / multicast commit
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // this CTA's pixel tile of unit u
  auto tile_of = [&](int unit, int& nt, int& w0, int& h0, int& b) {
    const int pair = unit / p.n_tiles;
    nt = unit % p.n_tiles;
    const int mt = 2 * pair + static_cast<int>(rank);
    const int tw = mt % p.tiles_w;
    h0 = (mt / p.tiles_w) % p.tiles_h;
    b = mt / (p.tiles_w * p.tiles_h);  // == p.B for the padding tile of an odd tile count: all loads OOB -> zeros
    w0 = tw * 128;
  };

  if (warp == 0) {
    // ===================== A-halo producer =====================
    if (lane == 0) {
      int slot = 0;
      uint32_t phase = 0;
      for (int unit = cluster_id; unit < total_units; unit += n_clusters) {
        int nt, w0, h0, b;
        tile_of(unit, nt, w0, h0, b);
        for (int s = 0; s < p.nseg; ++s) {
          const CUtensorMap* tmA = (s == 0) ? &tmA0 : &tmA1;
          const ConvSeg sg = p.seg[s];
          if (pp.halo) {
            const int row_bytes = pp.halo_w[s] * 128;
            for (int ch = 0; ch < sg.chunks; ++ch) {
              mbar_wait(&a_empty[slot], phase ^ 1);
              uint8_t* dst = smem_a + slot * pp.a_slot_bytes;
              if (leader) mbar_expect_tx(&a_full[slot], 2 * sg.kh * row_bytes);
              const uint32_t bar = mapa_u32(smem_u32(&a_full[slot]), 0);
              for (int r = 0; r < sg.kh; ++r)
                tma_load_4d_pair(dst + r * pp.pitch[s] * 128, tmA, bar, ch * kBlockK, w0 - sg.kw / 2,
                                 h0 + r - sg.kh / 2, b + sg.b_off);
              if (++slot == pp.a_stages) {
                slot = 0;
                phase ^= 1;
              }
            }
          } else {
            for (int dy = 0; dy < sg.kh; ++dy)
              for (int dx = 0; dx < sg.kw; ++dx)
                for (int ch = 0; ch < sg.chunks; ++ch) {
                  mbar_wait(&a_empty[slot], phase ^ 1);
                  if (leader) mbar_expect_tx(&a_full[slot], 2 * kABytes);
                  const uint32_t bar = mapa_u32(smem_u32(&a_full[slot]), 0);
                  tma_load_4d_pair(smem_a + slot * pp.a_slot_bytes, tmA, bar, ch * kBlockK, w0 + dx - sg.kw / 2,
                                   h0 + dy - sg.kh / 2, b + sg.b_off);
                  if (++slot == pp.a_stages) {
                    slot = 0;
                    phase ^= 1;
                  }
                }
          }
        }
      }
    }
  } else if (warp == 3) {
    // ===================== B (weight half-tile) producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const int half_rows = p.n_tile / 2;
      for (int unit = cluster_id; unit < total_units; unit += n_clusters) {
        const int nt = unit % p.n_tiles;
        int kb = 0;
        for (int s = 0; s < p.nseg; ++s) {
          const ConvSeg sg = p.seg[s];
          // k-block order inside a segment: halo mode = chunk-major then taps (the halo of a chunk serves all
          // its taps); per-tap mode = the packed order (tap-major).  Packed K index = seg_base + tap*chunks + ch.
          const int nkb = sg.chunks * sg.kh * sg.kw;
          for (int i = 0; i < nkb; ++i) {
            const int kidx = pp.halo ? ((i % (sg.kh * sg.kw)) * sg.chunks + i / (sg.kh * sg.kw)) : i;
            mbar_wait(&b_empty[stage], phase ^ 1);
            if (leader) mbar_expect_tx(&b_full[stage], 2 * b_stage_bytes);
            const uint32_t bar = mapa_u32(smem_u32(&b_full[stage]), 0);
            tma_load_2d_pair(smem_b + stage * b_stage_bytes, &tmB, bar, (kb + kidx) * kBlockK,
                             nt * p.n_tile + static_cast<int>(rank) * half_rows);
            if (++stage == pp.b_stages) {
              stage = 0;
              phase ^= 1;
            }
          }
          kb += sg.chunks * sg.kh * sg.kw;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA) =====================
    if (leader && lane == 0) {
      const uint32_t idesc = make_idesc(Elem<E>::kFmt, 256, p.n_tile, 0, 0);
      int slot = 0, stage = 0, acc = 0;
      uint32_t a_phase = 0, b_phase = 0, acc_phase = 0;
      for (int unit = cluster_id; unit < total_units; unit += n_clusters) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t d = tmem_base + acc * 256;
        uint32_t first = 1;
        for (int s = 0; s < p.nseg; ++s) {
          const ConvSeg sg = p.seg[s];
          const int taps = sg.kh * sg.kw;
          const int items = pp.halo ? sg.chunks : sg.chunks * taps;
          for (int it = 0; it < items; ++it) {
            mbar_wait(&a_full[slot], a_phase);
            const uint32_t a_base = smem_u32(smem_a + slot * pp.a_slot_bytes);
            const int ntap = pp.halo ? taps : 1;
            for (int tp = 0; tp < ntap; ++tp) {
              const int dy = tp / sg.kw, dx = tp % sg.kw;
              mbar_wait(&b_full[stage], b_phase);
              tcgen05_fence_after();
              const uint64_t adesc = make_smem_desc_sw128(a_base + (dy * pp.pitch[s] + dx) * 128, 16, 1024);
              const uint64_t bdesc = make_smem_desc_sw128(smem_u32(smem_b + stage * b_stage_bytes), 16, 1024);
#pragma unroll
              for (int k = 0; k < kBlockK / 16; ++k) {
                umma_f16_pair(d, adesc + 2 * k, bdesc + 2 * k, idesc, first ? 0u : 1u);
                first = 0;
              }
              umma_commit_pair(&b_empty[stage]);
              if (++stage == pp.b_stages) {
                stage = 0;
                b_phase ^= 1;
              }
            }
            umma_commit_pair(&a_empty[slot]);
            if (++slot == pp.a_stages) {
              slot = 0;
              a_phase ^= 1;
            }
          }
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
    const uint32_t empty_remote = mapa_u32(smem_u32(&tmem_empty[0]), 0);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int unit = cluster_id; unit < total_units; unit += n_clusters) {
      int nt, w0, h0, b;
      tile_of(unit, nt, w0, h0, b);
      const int hy = h0, wx = w0 + r;
      const bool valid = (b < p.B) && (hy < p.H) && (wx < p.W);
      const size_t pix = (static_cast<size_t>(b) * p.H + hy) * p.W + wx;
      mbar_wait(&tmem_full[acc], acc_phase);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + acc * 256 + (static_cast<uint32_t>(q * 32) << 16);
      convgemm_epilogue_tile<E, EPI>(p, bias_s, taddr, nt, b, hy, wx, valid, pix, half);
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(empty_remote + acc * 8);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer may still be reading this CTA's shared memory / writing its TMEM until here
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc_pair(tmem_base, kTmemCols);
  }
}

}  // namespace clstm
