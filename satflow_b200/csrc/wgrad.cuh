// Weight-gradient GEMM on tcgen05 tensor cores (sm_100a).
//
//   D[n, (j, c)] = sum over pixels p:  A[p, n] * B_j[p + shift_j, c]
//
// A  = gate pre-activation gradients dz (or the head's dlogit "col" tensor), NHWC fp16/bf16.
// B_j = one 64-channel chunk of an activation tensor, shifted by tap j (again a TMA coordinate shift
//       with out-of-bounds zero fill == the convolution's zero padding).
// Both operands are read from their NHWC layout, i.e. the GEMM's K dimension (pixels) is the slow
// one in shared memory, so both are described to the tensor core as MN-major (128-B swizzle).
// Each CTA owns (one 128-row block of n) x (one group of <= 8 column blocks) x (one split of the
// pixel range): its [128 x 64*nb] fp32 accumulator stays in TMEM for the whole K loop and is then
// added into its private slice partial[split] -- deterministic, no atomics; a finalize kernel
// reduces the splits (pointwise.cuh).
#pragma once
#include "pointwise.cuh"

namespace clstm {

constexpr int kWgTileP = 64;  // pixels per K step
constexpr int kWgMaxGroupBlocks = 8;
constexpr int kWgThreads = 256;
constexpr int kWgHaloRowBytes = 9216;  // one halo row slot: 66 pixels x 128 B, padded to a multiple of 1024

// Column blocks of one operand tensor: block j -> tap j / chunks (row-major over (kh, kw)), 64-channel
// chunk j % chunks.  A "direct" segment has kw == 1, cy == cx == 0 and a single tap.
struct WgradSeg {
  int nblk;    // taps * chunks
  int chunks;  // 64-channel chunks of the tensor
  int kw;      // taps per filter row
  int cy, cx;  // filter centre (kh/2, kw/2)
  int b_off;   // image offset into the segment's tensor map
};

struct WgradParams {
  int B, H, W;
  int BW, BH;  // pixel tile of a K step, BW*BH == 64
  int tiles_w, tiles_h, num_p_tiles;
  int n_blocks;     // 128-channel blocks of A
  int group_size;   // column blocks per CTA (<= kWgMaxGroupBlocks)
  int total_blocks;
  WgradSeg seg[2];  // seg[1].nblk may be 0
  int a_b_off;      // image offset into the A map
  int splits;
  int stages;
  float* partial;  // [splits][n_blocks*128][total_blocks*64] fp32
  int accumulate;  // 0: overwrite, 1: +=
  int halo;        // 1: 3x3 taps served from halo rows (see wgrad_kernel); requires chunks == 1, BW == 64, BH == 1
};

// Gate-gradient work that can ride along with a wgrad launch (GATE = true): wgrad is tensor bound and leaves the CUDA
// cores and 50 K of the SM's 64 K registers idle, the gate gradient of the NEXT step of the BPTT chain is an
// independent HBM stream (its dz goes to the other dz buffer).  16 extra warps per CTA run it as a grid-stride loop
// over pixels; they share nothing with the GEMM roles (no barrier, no shared memory).
constexpr int kWgGateThreads = 512;   // GATE == 1: 16 worker warps, <= 80 registers each
constexpr int kWgGateThreads2 = 384;  // GATE == 2: 12 worker warps with 128 registers each after the rebalance
struct WgGateWork {
  const void* gates;     // E [pix][4*HP] of the consumer cell / step
  const float* c_prev;   // nullable (zeros)
  const float* c_next;
  const float* src0;     // dh sources (nullable), fp32 [pix][HP], scaled by S
  const float* src1;
  const float* src2;
  float* dc;             // in/out
  void* dz_out;          // E [pix][4*HP]
  float* bias_partial;   // [gridDim.x * 16][4*HP]: one row per worker warp, accumulated
  unsigned int* dz_absmax;  // range statistics of the dz written here (float bits, see fold_absmax)
  unsigned npix;         // HP == 64 and npix * 256 < 2^32 (checked on the host)
  unsigned pix_begin;    // the workers handle pixels [pix_begin, npix) (hybrid schedule: the rest ran in the dgrad epilogue)
};

// GATE: 0 = plain wgrad (256 threads); 1 = + 16 gate-gradient worker warps, one item in flight per worker (validated,
// CLSTM_WG_GATE=1); 2 = 12 worker warps after a register rebalance (setmaxnreg: 48 for the GEMM roles, 128 for the
// workers) with two items in flight per worker.  GATE = 2 compiles but HAS NOT RUN ON HARDWARE YET (CLSTM_WG_GATE=2);
// it is the first thing to measure next.
template <typename E, int GATE>
__global__ void __launch_bounds__(GATE == 2 ? kWgThreads + kWgGateThreads2 : (GATE ? kWgThreads + kWgGateThreads : kWgThreads), 1)
wgrad_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB0,
             const __grid_constant__ CUtensorMap tmB1, const WgradParams p, const WgGateWork gw) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  int idx = blockIdx.x;
  const int split = idx % p.splits;
  idx /= p.splits;
  const int nb = idx % p.n_blocks;
  const int grp = idx / p.n_blocks;
  const int blk0 = grp * p.group_size;
  const int nblk = min(p.group_size, p.total_blocks - blk0);

  constexpr int kBoxBytes = kWgTileP * 128;  // one 64-channel x 64-pixel box
  // halo mode: a group is two filter rows (6 taps) = two [66 px x 64 ch] halo rows instead of six 64-pixel boxes;
  // the three horizontal taps of a row are ONE N = 192 MMA whose 64-channel groups are LBO = 128 B (one pixel)
  // apart, i.e. overlapping views of the same row shifted by one pixel each.
  const int stage_bytes = p.halo ? (2 * kBoxBytes + 2 * kWgHaloRowBytes) : (2 + nblk) * kBoxBytes;
  uint8_t* tail = smem + p.stages * stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty_bar = full_bar + 8;
  uint64_t* done_bar = empty_bar + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB0);
    tma_prefetch_desc(&tmB1);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // this CTA's K range: pixel tiles split, split+splits, ...
  const int my_tiles = (p.num_p_tiles - split + p.splits - 1) / p.splits;

  // GATE == 2: the register rebalance has to sit at the top of two disjoint code regions (ptxas sizes each region by the
  // setmaxnreg that dominates it; after a merge it falls back to the launch bound and spills).
  if (GATE == 2 && warp >= 8) {
    // launch: 640 threads x 96 registers.  Warps 0-7 (two warpgroups) give back 48 each, the three worker warpgroups
    // take 32 each: 256 x 48 + 384 x 128 = 61 440 = the launch allocation.
    asm volatile("setmaxnreg.inc.sync.aligned.u32 128;" ::: "memory");
    // ===================== gate-gradient workers, two items in flight (NOT YET RUN ON HARDWARE) =====================
    const int wt = threadIdx.x - kWgThreads;
    const int chunk = wt & 15, plane = wt >> 4;
    const E* gates_c = static_cast<const E*>(gw.gates) + chunk * 4;
    E* dzo_c = static_cast<E*>(gw.dz_out) + chunk * 4;
    float bsum[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int e = 0; e < 4; ++e) bsum[a][e] = 0.f;
    uint32_t zmax = 0;
    struct Raw {
      uint2 g[4];
      float4 cp, cn, dc, s0, s1, s2;
      unsigned pix;
      bool valid;
    };
    auto issue = [&](Raw& r, unsigned pix) {
      r.pix = pix;
      r.valid = pix < gw.npix;
      if (!r.valid) return;
      const unsigned o4 = pix * 256u, o1 = pix * 64u + chunk * 4;
#pragma unroll
      for (int a = 0; a < 4; ++a) r.g[a] = __ldg(reinterpret_cast<const uint2*>(gates_c + o4 + a * 64));
      r.cp = gw.c_prev ? __ldg(reinterpret_cast<const float4*>(gw.c_prev + o1)) : make_float4(0.f, 0.f, 0.f, 0.f);
      r.cn = __ldg(reinterpret_cast<const float4*>(gw.c_next + o1));
      r.dc = *reinterpret_cast<const float4*>(gw.dc + o1);
      if (gw.src0) r.s0 = __ldg(reinterpret_cast<const float4*>(gw.src0 + o1));
      if (gw.src1) r.s1 = __ldg(reinterpret_cast<const float4*>(gw.src1 + o1));
      if (gw.src2) r.s2 = __ldg(reinterpret_cast<const float4*>(gw.src2 + o1));
    };
    auto consume = [&](const Raw& r) {
      float dhv[4] = {0.f, 0.f, 0.f, 0.f};
      if (gw.src0) dhv[0] += r.s0.x, dhv[1] += r.s0.y, dhv[2] += r.s0.z, dhv[3] += r.s0.w;
      if (gw.src1) dhv[0] += r.s1.x, dhv[1] += r.s1.y, dhv[2] += r.s1.z, dhv[3] += r.s1.w;
      if (gw.src2) dhv[0] += r.s2.x, dhv[1] += r.s2.y, dhv[2] += r.s2.z, dhv[3] += r.s2.w;
      float4 dcn;
      uint2 dzp[4];
      gate_grad_item4<E>(r.g, r.cp, r.cn, r.dc, dhv, bsum, zmax, dcn, dzp);
      const unsigned o4 = r.pix * 256u, o1 = r.pix * 64u + chunk * 4;
      *reinterpret_cast<float4*>(gw.dc + o1) = dcn;
#pragma unroll
      for (int a = 0; a < 4; ++a) *reinterpret_cast<uint2*>(dzo_c + o4 + a * 64) = dzp[a];
    };
    const unsigned stride = gridDim.x * 24u;  // 384 workers = 24 pixels per CTA pass
    unsigned p0 = gw.pix_begin + blockIdx.x * 24u + plane, p1 = p0 + stride;
    Raw ra, rb;
    issue(ra, p0);
    issue(rb, p1);
    while (ra.valid) {
      consume(ra);
      p0 += 2 * stride;
      issue(ra, p0);
      if (!rb.valid) break;
      consume(rb);
      p1 += 2 * stride;
      issue(rb, p1);
    }
    fold_absmax<E>(zmax, gw.dz_absmax);
    float* row = gw.bias_partial + (static_cast<size_t>(blockIdx.x) * 16 + (warp - 8)) * 256;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float v = bsum[a][e] + __shfl_xor_sync(0xffffffffu, bsum[a][e], 16);
        if (lane < 16) row[a * 64 + chunk * 4 + e] += v;
      }
  } else {
  if constexpr (GATE == 2) asm volatile("setmaxnreg.dec.sync.aligned.u32 48;" ::: "memory");
  if (warp == 0) {
    // Lanes 0 .. nblk+1 each own one box of the stage (lane 0/1: the two A boxes, lane 2+j: column block j), so
    // the boxes of a stage are issued in parallel instead of serially by one thread.
    if (p.halo) {
      if (lane < 4) {
        // lanes 0,1: the two A boxes; lanes 2,3: halo rows rho = 2*grp + (lane-2): segment rho/3, filter row rho%3
        const CUtensorMap* map = &tmA;
        int c0 = nb * 128 + lane * 64, dxs = 0, dys = 0, boff = p.a_b_off;
        uint32_t bytes_off = lane * kBoxBytes;
        if (lane >= 2) {
          const int rho = 2 * grp + (lane - 2);
          const int sidx = rho / 3;
          map = sidx == 0 ? &tmB0 : &tmB1;
          c0 = 0;
          dxs = -1;
          dys = rho % 3 - 1;
          boff = p.seg[sidx].b_off;
          bytes_off = 2 * kBoxBytes + (lane - 2) * kWgHaloRowBytes;
        }
        int stage = 0;
        uint32_t phase = 0;
        for (int it = 0; it < my_tiles; ++it) {
          const int pt = split + it * p.splits;
          const int tw = pt % p.tiles_w;
          const int th = (pt / p.tiles_w) % p.tiles_h;
          const int b = pt / (p.tiles_w * p.tiles_h);
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (lane == 0) mbar_expect_tx(&full_bar[stage], 2 * kBoxBytes + 2 * 66 * 128);
          tma_load_4d(smem + stage * stage_bytes + bytes_off, map, &full_bar[stage], c0, tw * p.BW + dxs, th * p.BH + dys,
                      b + boff);
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    } else if (lane < 2 + nblk) {
      // per-lane constants of the box this lane loads
      const CUtensorMap* map = &tmA;
      int c0 = nb * 128 + lane * 64, dxs = 0, dys = 0, boff = p.a_b_off;
      if (lane >= 2) {
        int jj = blk0 + (lane - 2);
        const int sidx = (jj < p.seg[0].nblk) ? 0 : 1;
        if (sidx) jj -= p.seg[0].nblk;
        const WgradSeg sg = p.seg[sidx];
        const int tap = jj / sg.chunks;
        map = sidx == 0 ? &tmB0 : &tmB1;
        c0 = (jj % sg.chunks) * 64;
        dxs = tap % sg.kw - sg.cx;
        dys = tap / sg.kw - sg.cy;
        boff = sg.b_off;
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < my_tiles; ++it) {
        const int pt = split + it * p.splits;
        const int tw = pt % p.tiles_w;
        const int th = (pt / p.tiles_w) % p.tiles_h;
        const int b = pt / (p.tiles_w * p.tiles_h);
        const int w0 = tw * p.BW, h0 = th * p.BH;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* dst = smem + stage * stage_bytes;
        if (lane == 0) mbar_expect_tx(&full_bar[stage], stage_bytes);
        tma_load_4d(dst + lane * kBoxBytes, map, &full_bar[stage], c0, w0 + dxs, h0 + dys, b + boff);
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // M = 128 (two 64-channel MN groups, LBO apart); both operands MN-major.  Up to four adjacent
      // 64-column blocks (LBO apart) form ONE N = 256 instruction, so A is read from shared memory
      // once per four blocks instead of once per block (N = 64 MMAs are smem-read bound).
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < my_tiles; ++it) {
        mbar_wait(&full_bar[stage], phase);
        tcgen05_fence_after();
        const uint32_t base = smem_u32(smem + stage * stage_bytes);
        const uint64_t adesc = make_smem_desc_sw128(base, kBoxBytes, 1024);
        if (p.halo) {
          const uint32_t idesc = make_idesc(Elem<E>::kFmt, 128, 192, 1, 1);
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const uint64_t bdesc = make_smem_desc_sw128(base + 2 * kBoxBytes + r * kWgHaloRowBytes, 128, 1024);
#pragma unroll
            for (int k = 0; k < kWgTileP / 16; ++k)
              umma_f16(tmem_base + r * 192, adesc + k * 128, bdesc + k * 128, idesc, (it | k) != 0);
          }
        } else
        for (int j = 0; j < nblk; j += 4) {
          const int nb4 = min(4, nblk - j);
          const uint32_t idesc = make_idesc(Elem<E>::kFmt, 128, 64 * nb4, 1, 1);
          const uint64_t bdesc = make_smem_desc_sw128(base + (2 + j) * kBoxBytes, kBoxBytes, 1024);
#pragma unroll
          for (int k = 0; k < kWgTileP / 16; ++k) {
            // K = 16 pixels = 16 rows of 128 B = 2048 B per step (start-address field in 16-B units)
            umma_f16(tmem_base + j * 64, adesc + k * 128, bdesc + k * 128, idesc, (it | k) != 0);
          }
        }
        umma_commit(&empty_bar[stage]);
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
      umma_commit(done_bar);
    }
  } else if (GATE == 1 && warp >= 8) {
    // ===================== gate-gradient workers (see WgGateWork) =====================
    // one item = 1 pixel x 4 channels; a warp covers 2 pixels x 16 chunks; same math as gate_grad_kernel
    const int wt = threadIdx.x - kWgThreads;
    const int chunk = wt & 15, plane = wt >> 4;  // 32 pixels per CTA pass
    const E* gates_c = static_cast<const E*>(gw.gates) + chunk * 4;
    E* dzo_c = static_cast<E*>(gw.dz_out) + chunk * 4;
    float bsum[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int e = 0; e < 4; ++e) bsum[a][e] = 0.f;
    uint32_t zmax = 0;
    for (unsigned pix = gw.pix_begin + blockIdx.x * 32u + plane; pix < gw.npix; pix += gridDim.x * 32u) {
      const unsigned o4 = pix * 256u + 0u, o1 = pix * 64u + chunk * 4;
      uint2 g[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) g[a] = __ldg(reinterpret_cast<const uint2*>(gates_c + o4 + a * 64));
      const float4 cp = gw.c_prev ? __ldg(reinterpret_cast<const float4*>(gw.c_prev + o1)) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 cn4 = __ldg(reinterpret_cast<const float4*>(gw.c_next + o1));
      const float4 dc4 = *reinterpret_cast<const float4*>(gw.dc + o1);
      float dhv[4] = {0.f, 0.f, 0.f, 0.f};
      if (gw.src0) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(gw.src0 + o1));
        dhv[0] += t.x, dhv[1] += t.y, dhv[2] += t.z, dhv[3] += t.w;
      }
      if (gw.src1) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(gw.src1 + o1));
        dhv[0] += t.x, dhv[1] += t.y, dhv[2] += t.z, dhv[3] += t.w;
      }
      if (gw.src2) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(gw.src2 + o1));
        dhv[0] += t.x, dhv[1] += t.y, dhv[2] += t.z, dhv[3] += t.w;
      }
      float4 dcn;
      uint2 dzp[4];
      gate_grad_item4<E>(g, cp, cn4, dc4, dhv, bsum, zmax, dcn, dzp);
      *reinterpret_cast<float4*>(gw.dc + o1) = dcn;
#pragma unroll
      for (int a = 0; a < 4; ++a) *reinterpret_cast<uint2*>(dzo_c + o4 + a * 64) = dzp[a];
    }
    fold_absmax<E>(zmax, gw.dz_absmax);
    // bias partial sums: add the warp's two pixel lanes, one row of bias_partial per worker warp (fixed order)
    float* row = gw.bias_partial + (static_cast<size_t>(blockIdx.x) * 16 + (warp - 8)) * 256;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float v = bsum[a][e] + __shfl_xor_sync(0xffffffffu, bsum[a][e], 16);
        if (lane < 16) row[a * 64 + chunk * 4 + e] += v;
      }
  } else if (warp >= 4 && warp < 8) {
    const int q = warp & 3;
    const int row = q * 32 + lane;  // n within the block
    if (my_tiles > 0) {
      mbar_wait(done_bar, 0);
      tcgen05_fence_after();
    }
    const size_t ld = static_cast<size_t>(p.total_blocks) * 64;
    float* dst_row = p.partial + (static_cast<size_t>(split) * p.n_blocks * 128 + nb * 128 + row) * ld +
                     static_cast<size_t>(blk0) * 64;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
    for (int g = 0; g < nblk * 4; ++g) {
      uint32_t v[16];
      if (my_tiles > 0) {
        tmem_ld16(taddr + g * 16, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int e = 0; e < 16; ++e) v[e] = 0;
      }
      float4* d4 = reinterpret_cast<float4*>(dst_row + g * 16);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float4 o = make_float4(__uint_as_float(v[4 * e]), __uint_as_float(v[4 * e + 1]), __uint_as_float(v[4 * e + 2]),
                               __uint_as_float(v[4 * e + 3]));
        if (p.accumulate) {
          float4 old = d4[e];
          o.x += old.x, o.y += old.y, o.z += old.z, o.w += old.w;
        }
        d4[e] = o;
      }
    }
  }

  }  // GEMM roles / GATE == 1 workers

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

inline size_t wgrad_smem_bytes(int stages, int max_group_blocks) {
  return 1024 + static_cast<size_t>(stages) * (2 + max_group_blocks) * (kWgTileP * 128) + 17 * 8 + 16 + 64;
}

}  // namespace clstm
