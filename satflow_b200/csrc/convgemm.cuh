// Pixel-major implicit-GEMM convolution on tcgen05 tensor cores (sm_100a).
//
//   D[pixel, n] = sum over K-segments s, taps (dy,dx) of s, channels c of s:
//                     A_s[pixel + (dy - kh/2, dx - kw/2), c] * Wp[n, k(s,dy,dx,c)]
//
// A_s are NHWC activations (fp16/bf16, channel count a multiple of 64) reached through 4-D TMA
// tensor maps (C, W, H, B): a tap is a coordinate shift, padding is TMA out-of-bounds zero fill,
// so no im2col buffer and no concat([x,h]) is ever materialised (the reference does both:
// layers/ConvLSTM.py:45-47).  Wp is the packed weight matrix [N][K] (K-major).  Accumulators live
// in TMEM (two buffers, so the epilogue of tile i overlaps the MMAs of tile i+1).
//
// One kernel serves every pixel-major GEMM of the path through its epilogue:
//   EPI_LSTM  : fused gates -> c' = f*c + i*g, h' = o*tanh(c')   (layers/ConvLSTM.py:48-55)
//   EPI_STORE : plain fp32 store of up to two column ranges       (dgrad: dx | dh_prev)
//   EPI_HEAD  : bias + sigmoid, written as (B, C_out, T, H, W)    (conv_lstm.py:198-201)
//
// Warp roles (384 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator,
// warp 3 = idle, warps 4..11 = epilogue (two warps per TMEM lane quadrant, splitting columns).
#pragma once
#include "ptx.cuh"

namespace clstm {

enum { EPI_LSTM = 0, EPI_STORE = 1, EPI_HEAD = 2 };

constexpr int kTileM = 128;        // pixels per tile == UMMA M
constexpr int kBlockK = 64;        // channels per k-block (128 B of fp16/bf16 == one swizzle row)
constexpr int kABytes = kTileM * 128;
constexpr int kMaxStages = 8;
constexpr int kGemmThreads = 384;
constexpr int kTmemCols = 512;
constexpr int kKtabMax = 64;       // k-block table entries (rotated K loop)
// Epilogue staging, per half (4 warps = 128 pixel rows x one 16-channel group):
//   [c_prev 8 KB][c 8 KB][h 4 KB][gates 4 x 4 KB]   (EPI_STORE uses the c slot only)
constexpr int kStgCprev = 0, kStgC = 8192, kStgH = 16384, kStgG = 20480;
constexpr int kStgHalfLstm = 36864;   // EPI_LSTM
constexpr int kStgHalfStore = 8192;   // EPI_STORE: one fp32 [128 x 16] group
// EPI_LSTM with 8-channel groups (staged == 2): [c 4 KB][h 2 KB][gates 4 x 2 KB]; c_prev goes straight to registers.
// 28 KB of staging instead of 72 KB is what a FOURTH 48 KB operand stage needs (DESIGN.md finding 8: the 3-stage ring
// has no slack).
constexpr int kStg8C = 0, kStg8H = 4096, kStg8G = 6144;
constexpr int kStgHalfLstm8 = 14336;
__host__ __device__ constexpr int stg_half_bytes(int epi, int staged = 1) {
  return epi == 0 /*EPI_LSTM*/ ? (staged == 2 ? kStgHalfLstm8 : kStgHalfLstm) : (epi == 1 /*EPI_STORE*/ ? kStgHalfStore : 0);
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

struct ConvSeg {
  int chunks;  // 64-channel chunks of this segment's activation tensor
  int kh, kw;  // taps; (1,1) means "direct" (no shift)
  int b_off;   // image offset into the segment's tensor map (selects a time slot of a [T*B] stack)
};

struct ConvGemmParams {
  // geometry of the activation tensors (all segments share it)
  int B, H, W;
  int BW, BH;  // pixel tile, BW * BH == 128
  int tiles_w, tiles_h;
  int num_m_tiles;
  int n_tiles;  // N tiles of n_tile columns each
  int n_tile;   // UMMA N: multiple of 16, <= 256
  int nseg;
  ConvSeg seg[2];
  int stages;
  int rotate;       // 1: tile t starts its K loop at k-block t % kblocks (needs kblocks <= kKtabMax), see producer
  // ---- EPI_LSTM (n_tile == 256: gate-interleaved [i|f|o|g] x 64 hidden channels per N tile)
  const float* bias;  // [n_tiles * n_tile] in packed row order (all epilogues)
  const float* c_prev;  // fp32 [pixel][ldc] or nullptr (== zeros)
  float* c_next;        // fp32 [pixel][ldc]
  void* h_next;         // E    [pixel][ldc]
  void* gates;          // E    [pixel][4*ldc] ([i|f|o|g] blocks of ldc) or nullptr
  int ldc;              // padded hidden channels (multiple of 64)
  // staged epilogue (TMA stores / c_prev TMA load): image offsets into the epilogue tensor maps
  int staged;           // 1: outputs go through shared memory + TMA (tmX0..tmX2), 0: direct per-thread stores,
                        // 2 (EPI_LSTM): the same in 8-channel groups with c_prev read straight into registers
  int cprev_boff, cnext_boff, hnext_boff, gates_boff;
  int c16;              // staged == 1 only: the c stack holds E values times kCScale (tmX0 / tmX3 are then 16-bit maps)
  // ---- EPI_STORE
  float* out0;
  float* out1;
  int split_col;  // columns [0,split) -> out0, [split, N) -> out1
  int ld0, ld1;
  float out_scale;
  // ---- EPI_HEAD (n_tile == C_out rounded up to 16)
  float* y;  // (Bimg, C_out, T, H, W); the maps' batch index is t * Bimg + b
  int c_out, t_out, b_img, t0;  // output frame of image b is t0 + b / b_img
};

// Epilogue of one 128-pixel x n_tile accumulator tile for the calling warp: TMEM -> registers -> fused
// pointwise math -> global.  `taddr` already carries the warp's lane quadrant and the accumulator's column
// base; `half` selects which half of the column groups this warp handles (two warps share a quadrant).
template <typename E, int EPI>
__device__ __forceinline__ void convgemm_epilogue_tile(const ConvGemmParams& p, const float* bias_s, uint32_t taddr,
                                                       int nt, int b, int hy, int wx, bool valid, size_t pix,
                                                       int half) {
  if constexpr (EPI == EPI_LSTM) {
    // N tile = [i(64) | f(64) | o(64) | g(64)] for hidden channels nt*64 .. nt*64+63
    const float* bs = bias_s + nt * 256;
#pragma unroll 1
    for (int g2 = 0; g2 < 2; ++g2) {
      const int j0 = half * 32 + g2 * 16;
      uint32_t vi[16], vf[16], vo[16], vg[16];
      tmem_ld16(taddr + 0 + j0, vi);
      tmem_ld16(taddr + 64 + j0, vf);
      tmem_ld16(taddr + 128 + j0, vo);
      tmem_ld16(taddr + 192 + j0, vg);
      tmem_ld_wait();
      if (valid) {
        const size_t off = pix * p.ldc + nt * 64 + j0;
        float cp[16];
        if (p.c_prev != nullptr) {
          const float4* src = reinterpret_cast<const float4*>(p.c_prev + off);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float4 t = __ldg(src + e);
            cp[4 * e + 0] = t.x, cp[4 * e + 1] = t.y, cp[4 * e + 2] = t.z, cp[4 * e + 3] = t.w;
          }
        } else {
#pragma unroll
          for (int e = 0; e < 16; ++e) cp[e] = 0.f;
        }
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
          hn[e] = go[e] * ta * kHScale;  // stored scaled, see kHScale
          hn[e + 1] = go[e + 1] * tb * kHScale;
        }
        float4* cdst = reinterpret_cast<float4*>(p.c_next + off);
#pragma unroll
        for (int e = 0; e < 4; ++e)
          cdst[e] = make_float4(cn[4 * e], cn[4 * e + 1], cn[4 * e + 2], cn[4 * e + 3]);
        uint4* hdst = reinterpret_cast<uint4*>(reinterpret_cast<E*>(p.h_next) + off);
#pragma unroll
        for (int e = 0; e < 2; ++e)
          hdst[e] = make_uint4(Elem<E>::pack2(hn[8 * e + 0], hn[8 * e + 1]), Elem<E>::pack2(hn[8 * e + 2], hn[8 * e + 3]),
                               Elem<E>::pack2(hn[8 * e + 4], hn[8 * e + 5]), Elem<E>::pack2(hn[8 * e + 6], hn[8 * e + 7]));
        if (p.gates != nullptr) {
          E* gbase = reinterpret_cast<E*>(p.gates) + pix * (4 * static_cast<size_t>(p.ldc)) + nt * 64 + j0;
          const float* gsrc[4] = {gi, gf, go, gg};
#pragma unroll
          for (int gt = 0; gt < 4; ++gt) {
            uint4* gd = reinterpret_cast<uint4*>(gbase + static_cast<size_t>(gt) * p.ldc);
            const float* gv = gsrc[gt];
            const float ctr = gt < 3 ? kGateCenter : 0.f;  // sigmoid gates are stored centred (ptx.cuh kGateCenter)
#pragma unroll
            for (int e = 0; e < 2; ++e)
              gd[e] = make_uint4(Elem<E>::pack2(gv[8 * e + 0] - ctr, gv[8 * e + 1] - ctr),
                                 Elem<E>::pack2(gv[8 * e + 2] - ctr, gv[8 * e + 3] - ctr),
                                 Elem<E>::pack2(gv[8 * e + 4] - ctr, gv[8 * e + 5] - ctr),
                                 Elem<E>::pack2(gv[8 * e + 6] - ctr, gv[8 * e + 7] - ctr));
          }
        }
      }
    }
  } else if constexpr (EPI == EPI_STORE) {
    const int groups = p.n_tile / 16;
#pragma unroll 1
    for (int g = half; g < groups; g += 2) {
      uint32_t v[16];
      tmem_ld16(taddr + g * 16, v);
      tmem_ld_wait();
      if (valid) {
        const int col = nt * p.n_tile + g * 16;
        float* dst = (col < p.split_col) ? (p.out0 + pix * p.ld0 + col)
                                         : (p.out1 + pix * p.ld1 + (col - p.split_col));
        float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
        for (int e = 0; e < 4; ++e)
          d4[e] = make_float4(__uint_as_float(v[4 * e]) * p.out_scale, __uint_as_float(v[4 * e + 1]) * p.out_scale,
                              __uint_as_float(v[4 * e + 2]) * p.out_scale, __uint_as_float(v[4 * e + 3]) * p.out_scale);
      }
    }
  } else {  // EPI_HEAD
    const int groups = p.n_tile / 16;
    const int t = p.t0 + b / p.b_img, bi = b % p.b_img;
    const size_t plane = static_cast<size_t>(p.H) * p.W;
#pragma unroll 1
    for (int g = half; g < groups; g += 2) {
      uint32_t v[16];
      tmem_ld16(taddr + g * 16, v);
      tmem_ld_wait();
      if (valid) {
        float* ybase = p.y + ((static_cast<size_t>(bi) * p.c_out) * p.t_out + t) * plane +
                       static_cast<size_t>(hy) * p.W + wx;
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const int co = nt * p.n_tile + g * 16 + e;
          if (co < p.c_out)
            ybase[static_cast<size_t>(co) * p.t_out * plane] = fast_sigmoid(fmaf(__uint_as_float(v[e]), kHScaleInv, bias_s[co]));
        }
      }
    }
  }
}

template <typename E, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
convgemm_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmX0,
                const __grid_constant__ CUtensorMap tmX1, const __grid_constant__ CUtensorMap tmX2,
                const __grid_constant__ CUtensorMap tmX3, const ConvGemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stage_bytes = kABytes + p.n_tile * 128;
  uint8_t* smem_stg = smem + p.stages * stage_bytes;  // epilogue staging (1024-aligned: stage_bytes is)
  const int stg_half = p.staged ? stg_half_bytes(EPI, p.staged) : 0;
  uint8_t* tail = smem_stg + 2 * stg_half;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tmem_full = empty_bar + kMaxStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* cprev_full = tmem_empty + 2;  // [2], one per epilogue half
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(cprev_full + 4);
  float* bias_s = reinterpret_cast<float*>(tmem_slot + 4);  // n_tiles * n_tile floats

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.num_m_tiles * p.n_tiles;
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
      mbar_init(&tmem_empty[a], 8);  // one arrive per epilogue warp
      mbar_init(&cprev_full[a], 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, kTmemCols);
  if (p.bias != nullptr) {
    for (int i = threadIdx.x; i < p.n_tiles * p.n_tile; i += blockDim.x) bias_s[i] = p.bias[i];
  }
  // k-block table for the rotated K loop: {segment, dx offset, dy offset, channel offset}
  int4* ktab = reinterpret_cast<int4*>((reinterpret_cast<uintptr_t>(bias_s + p.n_tiles * p.n_tile) + 15) & ~uintptr_t(15));
  if (p.rotate && threadIdx.x < kblocks) {
    int kb = threadIdx.x, sgi = 0;
    while (kb >= p.seg[sgi].chunks * p.seg[sgi].kh * p.seg[sgi].kw) kb -= p.seg[sgi].chunks * p.seg[sgi].kh * p.seg[sgi].kw, ++sgi;
    const ConvSeg sg = p.seg[sgi];
    const int ch = kb % sg.chunks, tap = kb / sg.chunks;
    ktab[threadIdx.x] = make_int4(sgi, tap % sg.kw - sg.kw / 2, tap / sg.kw - sg.kh / 2, ch * kBlockK);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    // One lane per box: lane 0 loads the A tile, lane 1 the weight tile.  A single
    // thread issuing every cp.async.bulk.tensor of a stage serialises on the issue latency (measured: the wgrad
    // kernel went from 555 to 1250 TFLOP/s when its 8 boxes per stage were spread over 8 lanes).
    if (p.rotate) {
      // Rotated K loop.  Every CTA walks the k-blocks of its tile in the same order, so at any moment all 148 SMs
      // ask L2 for the SAME 32 KB weight k-block: those ~256 lines live in a few L2 slices, which serialise the
      // requests while the other slices idle.  Starting tile t at k-block t % kblocks spreads the concurrent weight
      // reads over the whole [n_tile x K] matrix.  (The accumulation order of a tile depends only on its position inside
      // the image, so a sample's result does not depend on where it sits in the batch.)
      if (lane < 2) {
        int stage = 0;
        uint32_t phase = 0;
        const int boff0 = p.seg[0].b_off, boff1 = p.seg[1].b_off;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
          const int mt = tile / p.n_tiles, nt = tile % p.n_tiles;
          const int w0 = (mt % p.tiles_w) * p.BW;
          const int h0 = ((mt / p.tiles_w) % p.tiles_h) * p.BH;
          const int b = mt / (p.tiles_w * p.tiles_h);
          // position inside the image: batch-index independent.  Keyed on the EVEN tile of a pair so that the CTA-pair
          // kernel (cellstep_pair.cuh: one K order per pair) and this one accumulate every tile in the same order —
          // which kernel runs depends on the batch size, a sample's bits must not.
          int kb = (((mt % (p.tiles_w * p.tiles_h)) & ~1) + nt) % kblocks;
          for (int i = 0; i < kblocks; ++i) {
            const int4 e = ktab[kb];
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* a_dst = smem + stage * stage_bytes;
            if (lane == 0) {
              mbar_expect_tx(&full_bar[stage], stage_bytes);
              tma_load_4d(a_dst, e.x ? &tmA1 : &tmA0, &full_bar[stage], e.w, w0 + e.y, h0 + e.z,
                          b + (e.x ? boff1 : boff0));
            } else {
              tma_load_2d(a_dst + kABytes, &tmB, &full_bar[stage], kb * kBlockK, nt * p.n_tile);
            }
            if (++kb == kblocks) kb = 0;
            if (++stage == p.stages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    } else if (lane < 2) {
      // plain K order (more k-blocks than the rotation table holds, or CLSTM_ROTATE=0)
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int mt = tile / p.n_tiles, nt = tile % p.n_tiles;
        const int tw = mt % p.tiles_w;
        const int th = (mt / p.tiles_w) % p.tiles_h;
        const int b = mt / (p.tiles_w * p.tiles_h);
        const int w0 = tw * p.BW, h0 = th * p.BH;
        int kb = 0;
        for (int s = 0; s < p.nseg; ++s) {
          const CUtensorMap* tmA = (s == 0) ? &tmA0 : &tmA1;
          const ConvSeg sg = p.seg[s];
          for (int dy = 0; dy < sg.kh; ++dy)
            for (int dx = 0; dx < sg.kw; ++dx)
              for (int ch = 0; ch < sg.chunks; ++ch, ++kb) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* a_dst = smem + stage * stage_bytes;
                if (lane == 0) {
                  mbar_expect_tx(&full_bar[stage], stage_bytes);
                  tma_load_4d(a_dst, tmA, &full_bar[stage], ch * kBlockK, w0 + dx - sg.kw / 2, h0 + dy - sg.kh / 2,
                              b + sg.b_off);
                } else {
                  tma_load_2d(a_dst + kABytes, &tmB, &full_bar[stage], kb * kBlockK, nt * p.n_tile);
                }
                if (++stage == p.stages) {
                  stage = 0;
                  phase ^= 1;
                }
              }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(Elem<E>::kFmt, kTileM, p.n_tile, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      // The descriptors of the NEXT stage are built right after the MMAs of the current one are issued: with a
      // 3-stage ring nothing between "data landed" and "first MMA issued" is hidden (DESIGN.md finding 8).
      uint64_t adesc = make_smem_desc_sw128(smem_u32(smem), 16, 1024);
      uint64_t bdesc = make_smem_desc_sw128(smem_u32(smem) + kABytes, 16, 1024);
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t d = tmem_base + acc * 256;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            // +32 B per K=16 step inside the 128-B swizzle row (start-address field is in 16-B units)
            umma_f16(d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
          const uint32_t a_addr = smem_u32(smem + stage * stage_bytes);
          adesc = make_smem_desc_sw128(a_addr, 16, 1024);
          bdesc = make_smem_desc_sw128(a_addr + kABytes, 16, 1024);
        }
        umma_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int q = warp & 3;             // TMEM lane quadrant this warp may access
    const int half = (warp - 4) >> 2;   // column half handled by this warp
    const int r = q * 32 + lane;        // tile row == pixel within tile
    const int hl = r / p.BW, wl = r % p.BW;
    int acc = 0;
    uint32_t acc_phase = 0;
    bool staged_path = false;
    if constexpr (EPI == EPI_LSTM || EPI == EPI_STORE) staged_path = (p.staged != 0);
    if (staged_path) {
      // ---- staged epilogue: registers -> swizzled shared memory -> TMA stores (and c_prev by TMA load).
      // The four warps of a half own one [128 px x 16 ch] group at a time; two named-barrier syncs per group.
      uint8_t* stg = smem_stg + half * stg_half;
      const bool issuer = (q == 0) && (lane == 0);
      const int bar_id = 1 + half;
      const uint32_t x64 = (static_cast<uint32_t>(r) >> 1) & 3u;  // SWIZZLE_64B: 16-B chunk ^= row bits [1,2]
      const uint32_t x32 = (static_cast<uint32_t>(r) >> 2) & 1u;  // SWIZZLE_32B: 16-B chunk ^= row bit 2
      auto coords = [&](int tile, int& nt, int& w0, int& h0, int& b) {
        const int mt = tile / p.n_tiles;
        nt = tile % p.n_tiles;
        w0 = (mt % p.tiles_w) * p.BW;
        h0 = ((mt / p.tiles_w) % p.tiles_h) * p.BH;
        b = mt / (p.tiles_w * p.tiles_h);
      };
      if (EPI == EPI_LSTM && p.staged == 2) {
        // ---- 8-channel groups.  Per half: 4 groups per tile; c_prev of the whole tile (32 floats per thread) is
        // requested from global memory BEFORE the wait for the accumulator, so its latency hides behind the MMAs.
        const bool has_cprev = p.c_prev != nullptr;
        const int lbw = 31 - __clz(p.BW);
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
          int nt, w0, h0, b;
          coords(tile, nt, w0, h0, b);
          float4 cpv[4][2];
          {
            const int hy = h0 + (r >> lbw), wx = w0 + (r & (p.BW - 1));
            const bool valid = has_cprev && hy < p.H && wx < p.W;
            const float* src = p.c_prev + ((static_cast<size_t>(b) * p.H + hy) * p.W + wx) * p.ldc + nt * 64 + half * 32;
#pragma unroll
            for (int g4 = 0; g4 < 4; ++g4)
#pragma unroll
              for (int j = 0; j < 2; ++j)
                cpv[g4][j] = valid ? __ldg(reinterpret_cast<const float4*>(src + g4 * 8 + j * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
          mbar_wait(&tmem_full[acc], acc_phase);
          tcgen05_fence_after();
          const uint32_t taddr = tmem_base + acc * 256 + (static_cast<uint32_t>(q * 32) << 16);
          const float* bs = bias_s + nt * 256;
#pragma unroll
          for (int g4 = 0; g4 < 4; ++g4) {
            const int j0 = half * 32 + g4 * 8;
            uint32_t vi[8], vf[8], vo[8], vg[8];
            tmem_ld8(taddr + 0 + j0, vi);
            tmem_ld8(taddr + 64 + j0, vf);
            tmem_ld8(taddr + 128 + j0, vo);
            tmem_ld8(taddr + 192 + j0, vg);
            if (q == 0 && lane < 6) tma_store_wait_read();  // this lane's previous store has finished reading staging
            named_bar_sync(bar_id, 128);
            tmem_ld_wait();
            const float cp[8] = {cpv[g4][0].x, cpv[g4][0].y, cpv[g4][0].z, cpv[g4][0].w,
                                 cpv[g4][1].x, cpv[g4][1].y, cpv[g4][1].z, cpv[g4][1].w};
            float cn[8], hn[8], gi[8], gf[8], go[8], gg[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              lstm_gates_shared_rcp(fmaf(__uint_as_float(vi[e]), kHScaleInv, bs[0 + j0 + e]),
                                    fmaf(__uint_as_float(vf[e]), kHScaleInv, bs[64 + j0 + e]),
                                    fmaf(__uint_as_float(vo[e]), kHScaleInv, bs[128 + j0 + e]),
                                    fmaf(__uint_as_float(vg[e]), kHScaleInv, bs[192 + j0 + e]), gi[e], gf[e], go[e], gg[e]);
              cn[e] = fmaf(gf[e], cp[e], gi[e] * gg[e]);
            }
#pragma unroll
            for (int e = 0; e < 8; e += 2) {
              float ta, tb;
              tanh_pair_shared_rcp(cn[e], cn[e + 1], ta, tb);
              hn[e] = go[e] * ta * kHScale;
              hn[e + 1] = go[e + 1] * tb * kHScale;
            }
            if (g4 == 3) {  // all TMEM reads of this accumulator are done
              tcgen05_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&tmem_empty[acc]);
            }
            auto pack8 = [](const float* v) {
              return make_uint4(Elem<E>::pack2(v[0], v[1]), Elem<E>::pack2(v[2], v[3]), Elem<E>::pack2(v[4], v[5]),
                                Elem<E>::pack2(v[6], v[7]));
            };
            // c: 32-byte rows, SWIZZLE_32B; h / gates: 16-byte rows, dense (thread r owns row r: conflict free)
#pragma unroll
            for (uint32_t j = 0; j < 2; ++j)
              *reinterpret_cast<float4*>(stg + kStg8C + r * 32 + ((j ^ x32) << 4)) =
                  make_float4(cn[4 * j], cn[4 * j + 1], cn[4 * j + 2], cn[4 * j + 3]);
            *reinterpret_cast<uint4*>(stg + kStg8H + r * 16) = pack8(hn);
            if (p.gates_boff >= 0) {
#pragma unroll
              for (int e = 0; e < 8; ++e) gi[e] -= kGateCenter, gf[e] -= kGateCenter, go[e] -= kGateCenter;  // stored centred
              *reinterpret_cast<uint4*>(stg + kStg8G + 0 * 2048 + r * 16) = pack8(gi);
              *reinterpret_cast<uint4*>(stg + kStg8G + 1 * 2048 + r * 16) = pack8(gf);
              *reinterpret_cast<uint4*>(stg + kStg8G + 2 * 2048 + r * 16) = pack8(go);
              *reinterpret_cast<uint4*>(stg + kStg8G + 3 * 2048 + r * 16) = pack8(gg);
            }
            fence_proxy_async_smem();
            named_bar_sync(bar_id, 128);
            if (q == 0 && lane < 6) {
              const int chan = nt * 64 + j0;
              if (lane == 0) {
                tma_store_4d(&tmX0, stg + kStg8C, chan, w0, h0, b + p.cnext_boff);
              } else if (lane == 1) {
                tma_store_4d(&tmX1, stg + kStg8H, chan, w0, h0, b + p.hnext_boff);
              } else if (p.gates_boff >= 0) {
                const int gt = lane - 2;
                tma_store_4d(&tmX2, stg + kStg8G + gt * 2048, gt * p.ldc + chan, w0, h0, b + p.gates_boff);
              }
              tma_store_commit();
            }
          }
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1;
        }
      } else if constexpr (EPI == EPI_LSTM) {
        const bool has_cprev = p.cprev_boff >= 0;
        uint32_t cp_phase = 0;
        if (issuer && has_cprev && static_cast<int>(blockIdx.x) < total_tiles) {
          int nt, w0, h0, b;
          coords(blockIdx.x, nt, w0, h0, b);
          mbar_expect_tx(&cprev_full[half], p.c16 ? 4096 : 8192);
          tma_load_4d(stg + kStgCprev, &tmX3, &cprev_full[half], nt * 64 + half * 32, w0, h0, b + p.cprev_boff);
        }
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
          int nt, w0, h0, b;
          coords(tile, nt, w0, h0, b);
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
            if (q == 0 && lane < 6) tma_store_wait_read();  // this lane's previous store has finished reading staging
            named_bar_sync(bar_id, 128);                    // staging is free; everyone has consumed c_prev
            if (issuer && has_cprev) {          // prefetch the next group's c_prev behind this group's math
              int tn = tile, ntn = nt, w0n = w0, h0n = h0, bn = b, j0n = j0 + 16;
              if (g2 == 1) {
                tn = tile + gridDim.x;
                j0n = half * 32;
                if (tn < total_tiles) coords(tn, ntn, w0n, h0n, bn);
              }
              if (tn < total_tiles) {
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
              hn[e] = go[e] * ta * kHScale;  // stored scaled, see kHScale
              hn[e + 1] = go[e + 1] * tb * kHScale;
            }
            if (g2 == 1) {  // all TMEM reads of this accumulator are done: hand it back to the MMA warp early
              tcgen05_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&tmem_empty[acc]);
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
              // one lane per output box (c, h, 4 gates): parallel TMA issue
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
      } else {  // EPI_STORE
        const int groups = p.n_tile / 16;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
          int nt, w0, h0, b;
          coords(tile, nt, w0, h0, b);
          mbar_wait(&tmem_full[acc], acc_phase);
          tcgen05_fence_after();
          const uint32_t taddr = tmem_base + acc * 256 + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
          for (int g = half; g < groups; g += 2) {
            uint32_t v[16];
            tmem_ld16(taddr + g * 16, v);
            if (issuer) tma_store_wait_read();
            named_bar_sync(bar_id, 128);
            tmem_ld_wait();
            if (g + 2 >= groups) {
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
            if (issuer) {
              const int col = nt * p.n_tile + g * 16;
              if (col < p.split_col)
                tma_store_4d(&tmX0, stg, col, w0, h0, b);
              else
                tma_store_4d(&tmX1, stg, col - p.split_col, w0, h0, b);
              tma_store_commit();
            }
          }
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1;
        }
      }
      if (q == 0 && lane < 6) tma_store_wait_all();  // outstanding bulk stores complete before the CTA retires
    } else {
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int mt = tile / p.n_tiles, nt = tile % p.n_tiles;
        const int tw = mt % p.tiles_w;
        const int th = (mt / p.tiles_w) % p.tiles_h;
        const int b = mt / (p.tiles_w * p.tiles_h);
        const int hy = th * p.BH + hl, wx = tw * p.BW + wl;
        const bool valid = (hy < p.H) && (wx < p.W);
        const size_t pix = (static_cast<size_t>(b) * p.H + hy) * p.W + wx;

        mbar_wait(&tmem_full[acc], acc_phase);
        tcgen05_fence_after();
        const uint32_t taddr = tmem_base + acc * 256 + (static_cast<uint32_t>(q * 32) << 16);

        convgemm_epilogue_tile<E, EPI>(p, bias_s, taddr, nt, b, hy, wx, valid, pix, half);
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
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

inline size_t convgemm_smem_bytes(int stages, int n_tile, int n_tiles, int stg_half = 0) {
  return 1024 + static_cast<size_t>(stages) * (kABytes + n_tile * 128) + 2 * static_cast<size_t>(stg_half) +
         (2 * kMaxStages + 8) * 8 + 16 + static_cast<size_t>(n_tiles) * n_tile * 4 + 64 + kKtabMax * 16 + 16;
}

}  // namespace clstm
