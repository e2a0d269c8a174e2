// HBM-bound helper kernels of the ConvLSTM path: layout packing, weight repacking, the fused
// gate-gradient pointwise backward, the head's loss-gradient "col" tensor, split reductions.
// All are plain coalesced/vectorised SIMT kernels; grids are multiples of the SM count.
#pragma once
#include "ptx.cuh"

namespace clstm {

// ------------------------------------------------------------------------------------------
// x (B,T,C,H,W) fp32  ->  xcol[t][b][h][w][KX]  (im2col of the 12-channel input, k = tap*C + c).
// The encoder-1 input never changes during the recurrence, so its taps are gathered once per
// forward; inside the cell kernel the x part is then a plain ("direct") K segment.
// ------------------------------------------------------------------------------------------
template <typename E>
__global__ void pack_xcol_kernel(const float* __restrict__ x, E* __restrict__ xcol, int B, int T, int C, int H,
                                 int W, int kh, int kw, int KX, int channels_last = 0) {
  const int chunks = KX / 8;
  const size_t total = static_cast<size_t>(T) * B * H * W * chunks;
  const int kreal = kh * kw * C;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int ck = static_cast<int>(i % chunks);
    size_t pix = i / chunks;
    const int w = static_cast<int>(pix % W);
    pix /= W;
    const int h = static_cast<int>(pix % H);
    pix /= H;
    const int b = static_cast<int>(pix % B);
    const int t = static_cast<int>(pix / B);
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = ck * 8 + e;
      float val = 0.f;
      if (k < kreal) {
        const int tap = k / C, c = k % C;
        const int hy = h + tap / kw - kh / 2, wx = w + tap % kw - kw / 2;
        if (hy >= 0 && hy < H && wx >= 0 && wx < W)
          val = channels_last ? __ldg(x + ((((static_cast<size_t>(b) * T + t) * H + hy) * W + wx) * C + c))
                              : __ldg(x + (((static_cast<size_t>(b) * T + t) * C + c) * H + hy) * W + wx);
      }
      v[e] = val;
    }
    uint4 o = make_uint4(Elem<E>::pack2(v[0], v[1]), Elem<E>::pack2(v[2], v[3]), Elem<E>::pack2(v[4], v[5]),
                         Elem<E>::pack2(v[6], v[7]));
    reinterpret_cast<uint4*>(xcol)[i] = o;
  }
}

// ------------------------------------------------------------------------------------------
// Row-tiled im2col (the fast path of pack_xcol_kernel and head_grad_col_kernel): one block owns one
// image row (t, b, h).  It stages the kh source rows of all C channels in shared memory once
// (coalesced along W, converted to E), then writes the row's [W][KP] output in 16-byte chunks, fully
// coalesced.  MODE 0: src0 = x (B,T,C,H,W), taps look forward (h + dy - kh/2).  MODE 1: the head's
// loss gradient dy*y*(1-y)*S from src0 = dy, src1 = y (B,C,T,H,W), taps look backward (h - (dy - kh/2)),
// which is what conv_transpose / the head wgrad need.  out is [(nt*B)][H][W][KP], k = tap*C + c.
// ------------------------------------------------------------------------------------------
template <typename E, int MODE>
__global__ void __launch_bounds__(256)
row_im2col_kernel(const float* __restrict__ src0, const float* __restrict__ src1, E* __restrict__ out, int B, int T,
                  int C, int H, int W, int kh, int kw, int KP, int t0, const float* __restrict__ scale_ptr) {
  extern __shared__ uint8_t rim_smem[];
  const int WP = W + kw - 1;
  int* lut = reinterpret_cast<int*>(rim_smem);                         // [KP] smem offset of (tap, c), or -1
  E* sm = reinterpret_cast<E*>(rim_smem + static_cast<size_t>(KP) * 4);  // [kh][C][WP]
  int idx = blockIdx.x;
  const int h = idx % H;
  idx /= H;
  const int b = idx % B;
  const int tt = idx / B;
  const int t = t0 + tt;
  const float scale = (MODE == 1) ? *scale_ptr : 1.f;
  for (int k = threadIdx.x; k < KP; k += blockDim.x) {
    int off = -1;
    if (k < kh * kw * C) {
      const int tap = k / C, c = k % C;
      const int dyi = tap / kw, dxi = tap % kw;
      const int r = (MODE != 1) ? dyi : kh - 1 - dyi;
      const int col = (MODE != 1) ? dxi : kw - 1 - dxi;
      off = (r * C + c) * WP + col;
    }
    lut[k] = off;
  }
  // Stage the source rows: one warp per (row, channel) line, lanes along W -> no per-element divisions and
  // up to WP/32 independent, fully coalesced loads in flight per lane.
  const int lines = kh * C;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int ln = warp; ln < lines; ln += nwarps) {
    const int c = ln % C, r = ln / C;
    const int hy = h + r - kh / 2;
    E* dst = sm + static_cast<size_t>(ln) * WP;
    if (hy < 0 || hy >= H) {
      for (int wcol = lane; wcol < WP; wcol += 32) dst[wcol] = Elem<E>::from_float(0.f);
      continue;
    }
    if (MODE == 2) {  // channels-last source (B,T,H,W,C): the reference datasets' on-wire layout (data/datasets.py:70-106)
      const size_t rowbase = (((static_cast<size_t>(b) * T + t) * H + hy) * W) * C + c;
#pragma unroll 4
      for (int wcol = lane; wcol < WP; wcol += 32) {
        const int wx = wcol - kw / 2;
        dst[wcol] = Elem<E>::from_float((wx >= 0 && wx < W) ? __ldg(src0 + rowbase + static_cast<size_t>(wx) * C) : 0.f);
      }
      continue;
    }
    const size_t base = (MODE == 0) ? (((static_cast<size_t>(b) * T + t) * C + c) * H + hy) * W
                                    : (((static_cast<size_t>(b) * C + c) * T + t) * H + hy) * W;
#pragma unroll 4
    for (int wcol = lane; wcol < WP; wcol += 32) {
      const int wx = wcol - kw / 2;
      float val = 0.f;
      if (wx >= 0 && wx < W) {
        if (MODE == 0) {
          val = __ldg(src0 + base + wx);
        } else {
          const float yy = __ldg(src1 + base + wx);
          val = __ldg(src0 + base + wx) * yy * (1.f - yy) * scale;
        }
      }
      dst[wcol] = Elem<E>::from_float(val);
    }
  }
  __syncthreads();
  const int chunks = KP / 8;
  uint4* orow = reinterpret_cast<uint4*>(out + ((static_cast<size_t>(tt) * B + b) * H + h) * static_cast<size_t>(W) * KP);
  const E zero = Elem<E>::from_float(0.f);
  if (blockDim.x % chunks == 0) {
    // fast path: a thread owns one 8-element chunk index for every pixel it visits -> its 8 tap offsets live in
    // registers and the inner loop is 8 shared-memory reads + one 16-byte store
    const int ck = threadIdx.x % chunks;
    int off[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) off[e] = lut[ck * 8 + e];
    const int wstep = blockDim.x / chunks;
    for (int w = threadIdx.x / chunks; w < W; w += wstep) {
      alignas(16) E v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = off[e] >= 0 ? sm[off[e] + w] : zero;
      orow[w * chunks + ck] = *reinterpret_cast<const uint4*>(v);
    }
  } else {
    for (int i = threadIdx.x; i < W * chunks; i += blockDim.x) {
      const int ck = i % chunks, w = i / chunks;
      alignas(16) E v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int o = lut[ck * 8 + e];
        v[e] = o >= 0 ? sm[o + w] : zero;
      }
      orow[i] = *reinterpret_cast<const uint4*>(v);
    }
  }
}

inline size_t row_im2col_smem_bytes(int C, int W, int kh, int kw, int KP) {
  return static_cast<size_t>(KP) * 4 + static_cast<size_t>(kh) * C * (W + kw - 1) * 2 + 16;
}

// NCHW fp32 (optionally a (B,T,C,H,W) time slice) -> NHWC E with channel padding (zeros).
template <typename E>
__global__ void pack_nhwc_kernel(const float* __restrict__ src, E* __restrict__ dst, int B, int C, int H, int W,
                                 int CP, size_t src_batch_stride, float scale) {
  const int chunks = CP / 8;
  const size_t total = static_cast<size_t>(B) * H * W * chunks;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int ck = static_cast<int>(i % chunks);
    size_t pix = i / chunks;
    const size_t hw = pix % (static_cast<size_t>(H) * W);
    const int b = static_cast<int>(pix / (static_cast<size_t>(H) * W));
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = ck * 8 + e;
      v[e] = (c < C) ? scale * __ldg(src + b * src_batch_stride + static_cast<size_t>(c) * H * W + hw) : 0.f;
    }
    reinterpret_cast<uint4*>(dst)[i] = make_uint4(Elem<E>::pack2(v[0], v[1]), Elem<E>::pack2(v[2], v[3]),
                                                  Elem<E>::pack2(v[4], v[5]), Elem<E>::pack2(v[6], v[7]));
  }
}

// NCHW fp32 -> NHWC fp32 with channel padding.
__global__ void pack_nhwc_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, int B, int C, int H,
                                     int W, int CP, const float* __restrict__ scale_ptr) {
  const float scale = scale_ptr ? *scale_ptr : 1.f;
  const size_t total = static_cast<size_t>(B) * H * W * CP;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % CP);
    const size_t pix = i / CP;
    const size_t hw = pix % (static_cast<size_t>(H) * W);
    const size_t b = pix / (static_cast<size_t>(H) * W);
    dst[i] = (c < C) ? scale * __ldg(src + (b * C + c) * H * W + hw) : 0.f;
  }
}

// NHWC (E or fp32, channel stride CP) -> NCHW fp32 (C real channels).  scale applied.
template <typename S>
__global__ void unpack_nchw_kernel(const S* __restrict__ src, float* __restrict__ dst, int B, int C, int H, int W,
                                   int CP, const float* __restrict__ scale_ptr, int accumulate, float extra_scale = 1.f) {
  const float scale = (scale_ptr ? *scale_ptr : 1.f) * extra_scale;
  const size_t total = static_cast<size_t>(B) * C * H * W;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t hw = i % (static_cast<size_t>(H) * W);
    const size_t bc = i / (static_cast<size_t>(H) * W);
    const int c = static_cast<int>(bc % C);
    const size_t b = bc / C;
    float v;
    if constexpr (sizeof(S) == 4)
      v = src[(b * H * W + hw) * CP + c];
    else
      v = Elem<S>::to_float(src[(b * H * W + hw) * CP + c]);
    v *= scale;
    dst[i] = accumulate ? dst[i] + v : v;
  }
}

// ------------------------------------------------------------------------------------------
// Weight repacking (reference layout fp32 -> packed E).  See DESIGN.md "Packed weights".
// ------------------------------------------------------------------------------------------
struct CellGeom {
  int cin;     // real input channels of the cell
  int hid;     // real hidden channels
  int HP;      // hidden channels padded to a multiple of 64
  int kh, kw;
  int in_col;  // 1: input segment is the im2col'd x (k = tap*cin + c, KIN = round_up(kh*kw*cin, 64))
               // 0: input segment is a conv over a CIP-channel NHWC tensor (k = tap*CIP + c, KIN = kh*kw*CIP)
  int CIP;     // input channels padded to a multiple of 64 (in_col == 0)
  int KIN;     // K extent of the input segment
  int in_scaled;  // 1: the input tensor is a hidden state, i.e. stored times kHScale (ptx.cuh); 0: plain data (x)
};

// forward: Wp[row = nt*256 + gate*64 + jj][k], bias_p[row]
template <typename E>
__global__ void pack_cell_weights_fwd_kernel(const float* __restrict__ w, const float* __restrict__ bias,
                                             E* __restrict__ wp, float* __restrict__ bias_p, CellGeom g) {
  const int K = g.KIN + g.kh * g.kw * g.HP;
  const int rows = 4 * g.HP;
  const int ctot = g.cin + g.hid;
  const size_t total = static_cast<size_t>(rows) * K;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(i % K);
    const int row = static_cast<int>(i / K);
    const int nt = row / 256, gate = (row % 256) / 64, jj = row % 64;
    const int j = nt * 64 + jj;
    float val = 0.f;
    if (j < g.hid) {
      const int n_ref = gate * g.hid + j;
      int c_full = -1, tap = 0;
      if (k < g.KIN) {
        if (g.in_col) {
          if (k < g.kh * g.kw * g.cin) tap = k / g.cin, c_full = k % g.cin;
        } else {
          tap = k / g.CIP;
          const int c = k % g.CIP;
          if (c < g.cin) c_full = c;
        }
      } else {
        const int kk = k - g.KIN;
        tap = kk / g.HP;
        const int c = kk % g.HP;
        if (c < g.hid) c_full = g.cin + c;
      }
      if (c_full >= 0) val = w[(static_cast<size_t>(n_ref) * ctot + c_full) * (g.kh * g.kw) + tap];
      // the accumulator is z * kHScale (the h segment is stored scaled): a plain input gets the factor in its weights
      if (k < g.KIN && !g.in_scaled) val *= kHScale;
      if (k == 0) bias_p[row] = bias ? bias[n_ref] : 0.f;
    } else if (k == 0) {
      bias_p[row] = 0.f;
    }
    wp[i] = Elem<E>::from_float(val);
  }
}

// dgrad: Wd[row][k = tap'*(4HP) + gate*HP + j] = W[gate*hid + j][c_full(row)][flipped tap']
//   rows: with_x ? [x part: CIP rows | h part: HP rows] : [h part: HP rows]
template <typename E>
__global__ void pack_cell_weights_dgrad_kernel(const float* __restrict__ w, E* __restrict__ wd, CellGeom g,
                                               int with_x) {
  const int taps = g.kh * g.kw;
  const int K = taps * 4 * g.HP;
  const int xrows = with_x ? g.CIP : 0;
  const int rows = xrows + g.HP;
  const int ctot = g.cin + g.hid;
  const size_t total = static_cast<size_t>(rows) * K;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(i % K);
    const int row = static_cast<int>(i / K);
    const int tapf = k / (4 * g.HP);
    const int n = k % (4 * g.HP);
    const int gate = n / g.HP, j = n % g.HP;
    int c_full = -1;
    if (row < xrows) {
      if (row < g.cin) c_full = row;
    } else {
      const int c = row - xrows;
      if (c < g.hid) c_full = g.cin + c;
    }
    float val = 0.f;
    if (c_full >= 0 && j < g.hid) {
      const int tap = taps - 1 - tapf;  // (kh-1-dy', kw-1-dx') in row-major tap index
      val = w[(static_cast<size_t>(gate * g.hid + j) * ctot + c_full) * taps + tap];
    }
    wd[i] = Elem<E>::from_float(val);
  }
}

// head forward: Wh[row = co (padded to NT)][k = tap*HP + c] = Whead[co][c][0][tap]; bias padded
template <typename E>
__global__ void pack_head_weights_fwd_kernel(const float* __restrict__ w, const float* __restrict__ bias,
                                             E* __restrict__ wp, float* __restrict__ bias_p, int c_out, int hid,
                                             int HP, int NT) {
  const int K = 9 * HP;
  const size_t total = static_cast<size_t>(NT) * K;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(i % K), row = static_cast<int>(i / K);
    const int tap = k / HP, c = k % HP;
    float val = 0.f;
    if (row < c_out && c < hid) val = w[(static_cast<size_t>(row) * hid + c) * 9 + tap];
    wp[i] = Elem<E>::from_float(val);
    if (k == 0) bias_p[row] = (row < c_out) ? bias[row] : 0.f;
  }
}

// row-marching head (head_rows.cuh): Wz[chunk][n = tap*16 + co (144 rows)][k = c % 64] = Whead[co][c][0][tap]
template <typename E>
__global__ void pack_head_weights_rows_kernel(const float* __restrict__ w, E* __restrict__ wz, int c_out, int hid,
                                              int HP) {
  const size_t total = static_cast<size_t>(HP / 64) * 144 * 64;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int kk = static_cast<int>(i % 64), n = static_cast<int>((i / 64) % 144), ch = static_cast<int>(i / (64 * 144));
    const int tap = n / 16, co = n % 16, c = ch * 64 + kk;
    float val = 0.f;
    if (co < c_out && c < hid) val = w[(static_cast<size_t>(co) * hid + c) * 9 + tap];
    wz[i] = Elem<E>::from_float(val);
  }
}

// head dgrad: Whd[row = c (HP rows)][k = tap*c_out + co (padded to KG)] = Whead[co][c][0][tap]
template <typename E>
__global__ void pack_head_weights_dgrad_kernel(const float* __restrict__ w, E* __restrict__ wd, int c_out, int hid,
                                               int HP, int KG) {
  const size_t total = static_cast<size_t>(HP) * KG;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(i % KG), c = static_cast<int>(i / KG);
    float val = 0.f;
    if (c < hid && k < 9 * c_out) {
      const int tap = k / c_out, co = k % c_out;
      val = w[(static_cast<size_t>(co) * hid + c) * 9 + tap];
    }
    wd[i] = Elem<E>::from_float(val);
  }
}

// ------------------------------------------------------------------------------------------
// Fused gate-gradient pointwise backward of layers/ConvLSTM.py:49-55 for one cell step.
//   dh = dh0 + dh1 + dh2 (nullable sources; fp32 NHWC, stride HP)
//   tc = tanh(c'); do = dh*tc; dct = dc + dh*o*(1-tc^2)
//   dz = [dct*g*i(1-i), dct*c*f(1-f), do*o(1-o), dct*i*(1-g^2)]  -> E, layout [pixel][4*HP]
//   dc <- dct*f (in place);  bias-gradient partial sums per block (deterministic two-stage)
// One thread = one pixel x 8 channels; consecutive threads = consecutive channel groups.
// ------------------------------------------------------------------------------------------
template <typename E>
__global__ void __launch_bounds__(256)
gate_grad_kernel(const E* __restrict__ gates, const float* __restrict__ c_prev, const float* __restrict__ c_next,
                 const float* __restrict__ dh0, const float* __restrict__ dh1, const float* __restrict__ dh2,
                 float* __restrict__ dc, E* __restrict__ dz, float* __restrict__ bias_partial, int bias_accumulate,
                 size_t npix, int HP, unsigned int* __restrict__ dz_absmax, int state16 = 0, int c16 = 0) {
  // c16: c_prev / c_next are E arrays holding c * kCScale
  // state16: dc and dh0 (the cell's own recurrent dh) are E arrays holding value * kStateDown (see ptx.cuh)
  extern __shared__ float red[];  // [256][9] padded
  uint32_t zmax = 0;  // packed running max |dz| of this thread (range statistics, see fold_absmax)
  const int groups = HP / 8;
  const int grp = threadIdx.x % groups;
  const int plane = threadIdx.x / groups;
  const int ppb = blockDim.x / groups;  // pixels per block iteration
  float bsum[4][8];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int e = 0; e < 8; ++e) bsum[a][e] = 0.f;

  for (size_t pix = static_cast<size_t>(blockIdx.x) * ppb + plane; pix < npix;
       pix += static_cast<size_t>(gridDim.x) * ppb) {
    const size_t off = pix * HP + grp * 8;
    float gv[4][8];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const uint4 u = *reinterpret_cast<const uint4*>(gates + pix * 4 * HP + a * HP + grp * 8);
      const float2 p0 = Elem<E>::unpack2(u.x), p1 = Elem<E>::unpack2(u.y), p2 = Elem<E>::unpack2(u.z),
                   p3 = Elem<E>::unpack2(u.w);
      const float ctr = a < 3 ? kGateCenter : 0.f;  // sigmoid gates are stored centred (ptx.cuh kGateCenter)
      gv[a][0] = p0.x + ctr, gv[a][1] = p0.y + ctr, gv[a][2] = p1.x + ctr, gv[a][3] = p1.y + ctr;
      gv[a][4] = p2.x + ctr, gv[a][5] = p2.y + ctr, gv[a][6] = p3.x + ctr, gv[a][7] = p3.y + ctr;
    }
    float cp[8], cn[8], dhv[8], dcv[8];
    auto ld8 = [&](const float* src, float* out) {
      const float4 a = *reinterpret_cast<const float4*>(src + off);
      const float4 b = *reinterpret_cast<const float4*>(src + off + 4);
      out[0] = a.x, out[1] = a.y, out[2] = a.z, out[3] = a.w, out[4] = b.x, out[5] = b.y, out[6] = b.z, out[7] = b.w;
    };
    auto ld8h = [&](const float* src, float* out, float up) {  // 8 packed 16-bit values times a power of two
      const uint4 u = *reinterpret_cast<const uint4*>(reinterpret_cast<const E*>(src) + off);
      const float2 p0 = Elem<E>::unpack2(u.x), p1 = Elem<E>::unpack2(u.y), p2 = Elem<E>::unpack2(u.z),
                   p3 = Elem<E>::unpack2(u.w);
      out[0] = p0.x * up, out[1] = p0.y * up, out[2] = p1.x * up, out[3] = p1.y * up;
      out[4] = p2.x * up, out[5] = p2.y * up, out[6] = p3.x * up, out[7] = p3.y * up;
    };
    if (c_prev) {
      if (c16) ld8h(c_prev, cp, kCScaleInv); else ld8(c_prev, cp);
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) cp[e] = 0.f;
    }
    if (c16) ld8h(c_next, cn, kCScaleInv); else ld8(c_next, cn);
    if (state16) ld8h(dc, dcv, kStateUp); else ld8(dc, dcv);
#pragma unroll
    for (int e = 0; e < 8; ++e) dhv[e] = 0.f;
    float tmp[8];
    if (dh0) {
      if (state16) ld8h(dh0, tmp, kStateUp); else ld8(dh0, tmp);
#pragma unroll
      for (int e = 0; e < 8; ++e) dhv[e] += tmp[e];
    }
    if (dh1) {
      ld8(dh1, tmp);
#pragma unroll
      for (int e = 0; e < 8; ++e) dhv[e] += tmp[e];
    }
    if (dh2) {
      ld8(dh2, tmp);
#pragma unroll
      for (int e = 0; e < 8; ++e) dhv[e] += tmp[e];
    }
    float dzv[4][8], dcn[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float i = gv[0][e], f = gv[1][e], o = gv[2][e], g = gv[3][e];
      const float tc = fast_tanh(cn[e]);
      const float d_o = dhv[e] * tc;
      const float dct = fmaf(dhv[e] * o, 1.f - tc * tc, dcv[e]);
      dzv[0][e] = dct * g * i * (1.f - i);
      dzv[1][e] = dct * cp[e] * f * (1.f - f);
      dzv[2][e] = d_o * o * (1.f - o);
      dzv[3][e] = dct * i * (1.f - g * g);
      dcn[e] = dct * f;
#pragma unroll
      for (int a = 0; a < 4; ++a) bsum[a][e] += dzv[a][e];
    }
    if (state16) {
      *reinterpret_cast<uint4*>(reinterpret_cast<E*>(dc) + off) =
          make_uint4(Elem<E>::pack2(dcn[0] * kStateDown, dcn[1] * kStateDown), Elem<E>::pack2(dcn[2] * kStateDown, dcn[3] * kStateDown),
                     Elem<E>::pack2(dcn[4] * kStateDown, dcn[5] * kStateDown), Elem<E>::pack2(dcn[6] * kStateDown, dcn[7] * kStateDown));
    } else {
      *reinterpret_cast<float4*>(dc + off) = make_float4(dcn[0], dcn[1], dcn[2], dcn[3]);
      *reinterpret_cast<float4*>(dc + off + 4) = make_float4(dcn[4], dcn[5], dcn[6], dcn[7]);
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const uint4 o = make_uint4(Elem<E>::pack2(dzv[a][0], dzv[a][1]), Elem<E>::pack2(dzv[a][2], dzv[a][3]),
                                 Elem<E>::pack2(dzv[a][4], dzv[a][5]), Elem<E>::pack2(dzv[a][6], dzv[a][7]));
      zmax = Elem<E>::absmax2(Elem<E>::absmax2(Elem<E>::absmax2(Elem<E>::absmax2(zmax, o.x), o.y), o.z), o.w);
      *reinterpret_cast<uint4*>(dz + pix * 4 * HP + a * HP + grp * 8) = o;
    }
  }
  fold_absmax<E>(zmax, dz_absmax);
  // block reduction of the bias partial sums over the `ppb` pixel lanes (fixed order -> deterministic), one
  // gate at a time so the kernel needs only 256 x 9 floats of shared memory: it must be able to co-reside with
  // a weight-gradient CTA (198 KB of shared memory) on the same SM (DESIGN.md "backward overlap").
#pragma unroll
  for (int a = 0; a < 4; ++a) {  // unrolled so that bsum stays in registers
    float* mine = red + threadIdx.x * 9;
#pragma unroll
    for (int e = 0; e < 8; ++e) mine[e] = bsum[a][e];
    __syncthreads();
    if (threadIdx.x < groups) {
      for (int e = 0; e < 8; ++e) {
        float s = 0.f;
        for (int pl = 0; pl < ppb; ++pl) s += red[(pl * groups + threadIdx.x) * 9 + e];
        float* dst = bias_partial + static_cast<size_t>(blockIdx.x) * 4 * HP + a * HP + threadIdx.x * 8 + e;
        *dst = bias_accumulate ? *dst + s : s;
      }
    }
    __syncthreads();
  }
}

// The same math for one item of 1 pixel x 4 channels held in registers — the unit of work of the gate-gradient passes
// that ride inside the GEMM kernels (dgradT_fused_kernel's epilogue, the worker warps of wgrad_kernel).
//   g[a]: the four saved gates (packed 16-bit x4), cp / cn: c_prev / c_next, dcin: incoming dc, dhv: summed dh sources
//   -> dz_out[a] packed 16-bit, dc_out = dct * f, bsum += dz (bias-gradient partial), zmax = running packed max |dz|
// RC (recompute c'): cn4 is ignored and c' = f * c_prev + i * g is rebuilt from the saved (16-bit) gates — the same
// rounded gates every other term of the gradient already uses — which removes one fp32 stream (8 % of the fused
// launch's HBM bytes).  Only with 11-bit gates (fp16); bf16 gates would put 2^-9 of c' into tanh(c').
template <typename E, bool RC = false>
__device__ __forceinline__ void gate_grad_item4(const uint2 (&g)[4], const float4& cp4, const float4& cn4,
                                                const float4& dcin, const float (&dhv)[4], float (&bsum)[4][4],
                                                uint32_t& zmax, float4& dc_out, uint2 (&dz_out)[4]) {
  float gv[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const float2 p0 = Elem<E>::unpack2(g[a].x), p1 = Elem<E>::unpack2(g[a].y);
    const float ctr = a < 3 ? kGateCenter : 0.f;  // sigmoid gates are stored centred
    gv[a][0] = p0.x + ctr, gv[a][1] = p0.y + ctr, gv[a][2] = p1.x + ctr, gv[a][3] = p1.y + ctr;
  }
  const float cp[4] = {cp4.x, cp4.y, cp4.z, cp4.w};
  const float cn[4] = {cn4.x, cn4.y, cn4.z, cn4.w};
  const float dcv[4] = {dcin.x, dcin.y, dcin.z, dcin.w};
  float dzv[4][4], dcn[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float gi = gv[0][e], gf = gv[1][e], go = gv[2][e], gg = gv[3][e];
    const float tc = fast_tanh(RC ? fmaf(gf, cp[e], gi * gg) : cn[e]);
    const float d_o = dhv[e] * tc;
    const float dct = fmaf(dhv[e] * go, 1.f - tc * tc, dcv[e]);
    dzv[0][e] = dct * gg * gi * (1.f - gi);
    dzv[1][e] = dct * cp[e] * gf * (1.f - gf);
    dzv[2][e] = d_o * go * (1.f - go);
    dzv[3][e] = dct * gi * (1.f - gg * gg);
    dcn[e] = dct * gf;
#pragma unroll
    for (int a = 0; a < 4; ++a) bsum[a][e] += dzv[a][e];
  }
  dc_out = make_float4(dcn[0], dcn[1], dcn[2], dcn[3]);
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    dz_out[a] = make_uint2(Elem<E>::pack2(dzv[a][0], dzv[a][1]), Elem<E>::pack2(dzv[a][2], dzv[a][3]));
    zmax = Elem<E>::absmax2(Elem<E>::absmax2(zmax, dz_out[a].x), dz_out[a].y);
  }
}

// out[0..3] = {S, 1/S, max |dlogit| of the head, max |S * dz| over every gate-gradient pass of the last backward}
// (clstm_plan_grad_status; stats[0] / stats[1] hold float bits).
__global__ void grad_status_kernel(const float* __restrict__ scale, const unsigned int* __restrict__ stats,
                                   float* __restrict__ out) {
  out[0] = scale[0];
  out[1] = scale[1];
  out[2] = __uint_as_float(stats[0]);
  out[3] = __uint_as_float(stats[1]);
}

// ------------------------------------------------------------------------------------------
// Head backward pointwise: dlogit = dy * y * (1 - y) * scale, gathered into the "col" tensor
//   G[(t*B + b)][h][w][k = tap*C + co] = dlogit[b][co][t][h - (dy-1)][w - (dx-1)]   (zero outside)
// which is the A operand of both the head dgrad (plain GEMM) and the head wgrad.
// ------------------------------------------------------------------------------------------
template <typename E>
__global__ void head_grad_col_kernel(const float* __restrict__ dyv, const float* __restrict__ y,
                                     E* __restrict__ G, int B, int C, int T, int H, int W, int KG, int t0, int nt,
                                     const float* __restrict__ scale_ptr) {
  // Only time steps [t0, t0 + nt) are gathered: G is [nt*B][H][W][KG].
  const float scale = *scale_ptr;
  const int chunks = KG / 8;
  const size_t total = static_cast<size_t>(nt) * B * H * W * chunks;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int ck = static_cast<int>(i % chunks);
    size_t pix = i / chunks;
    const int w = static_cast<int>(pix % W);
    pix /= W;
    const int h = static_cast<int>(pix % H);
    pix /= H;
    const int b = static_cast<int>(pix % B);
    const int t = t0 + static_cast<int>(pix / B);
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = ck * 8 + e;
      float val = 0.f;
      if (k < 9 * C) {
        const int tap = k / C, co = k % C;
        const int hy = h - (tap / 3 - 1), wx = w - (tap % 3 - 1);
        if (hy >= 0 && hy < H && wx >= 0 && wx < W) {
          const size_t idx = (((static_cast<size_t>(b) * C + co) * T + t) * H + hy) * W + wx;
          const float yy = __ldg(y + idx);
          val = __ldg(dyv + idx) * yy * (1.f - yy) * scale;
        }
      }
      v[e] = val;
    }
    reinterpret_cast<uint4*>(G)[i] = make_uint4(Elem<E>::pack2(v[0], v[1]), Elem<E>::pack2(v[2], v[3]),
                                                Elem<E>::pack2(v[4], v[5]), Elem<E>::pack2(v[6], v[7]));
  }
}

// One pass over dy and y (B, C, T, H, W) for both statistics the head backward needs before anything else:
//   partial[co][b * chunks + chunk] = sum over the chunk of dlogit = dy * y * (1 - y)     (head bias gradient)
//   *amax_bits = max |dlogit|                                                            (loss scale; optional)
// Block (chunk, b*C + co) walks a contiguous piece of one (b, co) run of T*H*W floats: no per-element index
// arithmetic, float4 loads when the run length allows.
__global__ void __launch_bounds__(256)
head_grad_stats_kernel(const float* __restrict__ dyv, const float* __restrict__ y, float* __restrict__ partial,
                       unsigned int* __restrict__ amax_bits, int B, int C, size_t per_b, int chunks) {
  const int bc = blockIdx.y, chunk = blockIdx.x;
  const int b = bc / C, co = bc % C;
  const size_t base = static_cast<size_t>(bc) * per_b;
  float s = 0.f, m = 0.f;
  if ((per_b & 3) == 0) {
    const size_t n4 = per_b >> 2;
    const size_t lo = n4 * chunk / chunks, hi = n4 * (chunk + 1) / chunks;
    const float4* d4 = reinterpret_cast<const float4*>(dyv + base);
    const float4* y4 = reinterpret_cast<const float4*>(y + base);
    for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
      const float4 a = __ldg(d4 + i), yy = __ldg(y4 + i);
      const float v0 = a.x * yy.x * (1.f - yy.x), v1 = a.y * yy.y * (1.f - yy.y);
      const float v2 = a.z * yy.z * (1.f - yy.z), v3 = a.w * yy.w * (1.f - yy.w);
      s += (v0 + v1) + (v2 + v3);
      m = fmaxf(fmaxf(m, fmaxf(fabsf(v0), fabsf(v1))), fmaxf(fabsf(v2), fabsf(v3)));
    }
  } else {
    const size_t lo = per_b * chunk / chunks, hi = per_b * (chunk + 1) / chunks;
    for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
      const float yy = __ldg(y + base + i);
      const float v = __ldg(dyv + base + i) * yy * (1.f - yy);
      s += v;
      m = fmaxf(m, fabsf(v));
    }
  }
  __shared__ float red_s[256];
  __shared__ float red_m[256];
  red_s[threadIdx.x] = s;
  red_m[threadIdx.x] = m;
  __syncthreads();
  for (int st = 128; st > 0; st >>= 1) {
    if (threadIdx.x < st) {
      red_s[threadIdx.x] += red_s[threadIdx.x + st];
      red_m[threadIdx.x] = fmaxf(red_m[threadIdx.x], red_m[threadIdx.x + st]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    partial[static_cast<size_t>(co) * (static_cast<size_t>(B) * chunks) + static_cast<size_t>(b) * chunks + chunk] = red_s[0];
    const float mm = red_m[0];
    // a non-finite dy (NaN compares false everywhere) is recorded as +Inf so that the host can see it
    const float ms = (red_s[0] - red_s[0] == 0.f && mm < 3.0e38f) ? mm : __uint_as_float(0x7F800000u);
    if (amax_bits != nullptr && ms > 0.f) atomicMax(amax_bits, __float_as_uint(ms));
  }
}

// out[i] = scale * sum_r partial[r*stride_r + idx(i)]   generic strided split reduction
__global__ void reduce_rows_kernel(const float* __restrict__ partial, float* __restrict__ out, int n, int rows,
                                   size_t row_stride, size_t elem_stride, float scale, int accumulate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int r = 0; r < rows; ++r) s += partial[r * row_stride + i * elem_stride];
  s *= scale;
  out[i] = accumulate ? out[i] + s : s;
}

// Cell bias gradient: db_ref[gate*hid + j] = scale * sum_blocks partial[blk][gate*HP + j]
// One block reduces 32 output columns: 8 row lanes per column walk the partial rows with stride 8 (coalesced 128-byte
// row segments), then a fixed-order tree over the lanes in shared memory — deterministic, and ~6x shorter than one
// thread per column walking all rows (which cost 51 us per cell on launch-bound shapes).
__global__ void __launch_bounds__(256)
cell_bias_finalize_kernel(const float* __restrict__ partial, float* __restrict__ db, int nblocks, int hid, int HP,
                          const float* __restrict__ scale_ptr, int accumulate) {
  __shared__ float red[8][33];
  const int col = threadIdx.x & 31, lane = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + col;
  float s = 0.f;
  if (i < 4 * hid) {
    const int gate = i / hid, j = i % hid;
    const float* src = partial + gate * HP + j;
    for (int b = lane; b < nblocks; b += 8) s += src[static_cast<size_t>(b) * 4 * HP];
  }
  red[lane][col] = s;
  __syncthreads();
  if (lane == 0 && i < 4 * hid) {
    float t = 0.f;
#pragma unroll
    for (int l = 0; l < 8; ++l) t += red[l][col];
    t *= *scale_ptr;
    db[i] = accumulate ? db[i] + t : t;
  }
}

// Cell weight gradient: dW_ref[n_ref][c_full][tap] = scale * sum_splits D[s][gate*HP + j][k(c_full,tap)]
__global__ void cell_wgrad_finalize_kernel(const float* __restrict__ partial, float* __restrict__ dw, CellGeom g,
                                           int splits, const float* __restrict__ scale_ptr, int accumulate) {
  const float scale = *scale_ptr;
  const int taps = g.kh * g.kw;
  const int ctot = g.cin + g.hid;
  const size_t total = static_cast<size_t>(4 * g.hid) * ctot * taps;
  const size_t K = static_cast<size_t>(g.KIN) + taps * g.HP;
  const size_t split_stride = static_cast<size_t>(4 * g.HP) * K;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int tap = static_cast<int>(i % taps);
    const int c_full = static_cast<int>((i / taps) % ctot);
    const int n_ref = static_cast<int>(i / (static_cast<size_t>(taps) * ctot));
    const int gate = n_ref / g.hid, j = n_ref % g.hid;
    size_t k;
    float unscale = kHScaleInv;  // the wgrad operand of these columns was a hidden state stored times kHScale
    if (c_full < g.cin) {
      k = g.in_col ? static_cast<size_t>(tap) * g.cin + c_full : static_cast<size_t>(tap) * g.CIP + c_full;
      if (!g.in_scaled) unscale = 1.f;
    } else {
      k = static_cast<size_t>(g.KIN) + static_cast<size_t>(tap) * g.HP + (c_full - g.cin);
    }
    const float* src = partial + static_cast<size_t>(gate * g.HP + j) * K + k;
    float s = 0.f;
    for (int sp = 0; sp < splits; ++sp) s += src[sp * split_stride];
    s *= scale * unscale;
    dw[i] = accumulate ? dw[i] + s : s;
  }
}

// Head weight gradient: dWhead[co][c][tap] = scale * sum_splits D[s][tap*C + co][c]   (D rows KG, cols HP)
__global__ void head_wgrad_finalize_kernel(const float* __restrict__ partial, float* __restrict__ dw, int c_out,
                                           int hid, int HP, int KG, int splits, const float* __restrict__ scale_ptr,
                                           int accumulate) {
  const int total = c_out * hid * 9;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const float scale = *scale_ptr;
  const int tap = i % 9, c = (i / 9) % hid, co = i / (9 * hid);
  const float* src = partial + static_cast<size_t>(tap * c_out + co) * HP + c;
  const size_t split_stride = static_cast<size_t>(KG) * HP;
  float s = 0.f;
  for (int sp = 0; sp < splits; ++sp) s += src[sp * split_stride];
  s *= scale * kHScaleInv;  // the h-stack operand is stored times kHScale
  dw[i] = accumulate ? dw[i] + s : s;
}

// ------------------------------------------------------------------------------------------
// Fused MSE loss + gradient (training_step, conv_lstm.py:55-69): y is the rollout output (B,C,T,H,W), target is
// (B,T,C,H,W) as the reference's batches are; one block per (b, t, c) plane (contiguous in both tensors).
//   partial[(b*T + t)*C + c] = sum over the plane of (y - target)^2        (ordered -> deterministic)
//   dy = (2 / N) * (y - target)   in y's layout, N = B*C*T*H*W
// mse_finalize_kernel turns the partials into the mean loss and the T per-frame means (the reference fetches each
// of those with its own .item() host sync, conv_lstm.py:66-69).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
mse_loss_grad_kernel(const float* __restrict__ y, const float* __restrict__ target, float* __restrict__ dy,
                     float* __restrict__ partial, int B, int C, int T, int HW, float two_over_n) {
  int idx = blockIdx.x;
  const int c = idx % C;
  idx /= C;
  const int t = idx % T;
  const int b = idx / T;
  const size_t yo = ((static_cast<size_t>(b) * C + c) * T + t) * HW;
  const size_t to = ((static_cast<size_t>(b) * T + t) * C + c) * HW;
  float s = 0.f;
  if ((HW & 3) == 0) {
    const float4* y4 = reinterpret_cast<const float4*>(y + yo);
    const float4* t4 = reinterpret_cast<const float4*>(target + to);
    float4* d4 = dy ? reinterpret_cast<float4*>(dy + yo) : nullptr;
    for (int i = threadIdx.x; i < HW / 4; i += blockDim.x) {
      const float4 a = __ldg(y4 + i), g = __ldg(t4 + i);
      const float4 d = make_float4(a.x - g.x, a.y - g.y, a.z - g.z, a.w - g.w);
      s += d.x * d.x + d.y * d.y + d.z * d.z + d.w * d.w;
      if (d4) d4[i] = make_float4(d.x * two_over_n, d.y * two_over_n, d.z * two_over_n, d.w * two_over_n);
    }
  } else {
    for (int i = threadIdx.x; i < HW; i += blockDim.x) {
      const float d = __ldg(y + yo + i) - __ldg(target + to + i);
      s += d * d;
      if (dy) dy[yo + i] = d * two_over_n;
    }
  }
  __shared__ float red[256];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int st = 128; st > 0; st >>= 1) {
    if (threadIdx.x < st) red[threadIdx.x] += red[threadIdx.x + st];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}

// out[0] = mean loss, out[1 + t] = mean loss of frame t
__global__ void mse_finalize_kernel(const float* __restrict__ partial, float* __restrict__ out, int B, int C, int T,
                                    float inv_n_total, float inv_n_frame) {
  const int t = threadIdx.x;
  __shared__ float frame[1024];
  float s = 0.f;
  if (t < T)
    for (int b = 0; b < B; ++b)
      for (int c = 0; c < C; ++c) s += partial[(static_cast<size_t>(b) * T + t) * C + c];
  if (t < T) {
    frame[t] = s;
    out[1 + t] = s * inv_n_frame;
  }
  __syncthreads();
  if (t == 0) {
    float tot = 0.f;
    for (int i = 0; i < T; ++i) tot += frame[i];
    out[0] = tot * inv_n_total;
  }
}

// amax_bits <- max(amax_bits, max |a|)
__global__ void __launch_bounds__(256)
abs_amax_kernel(const float* __restrict__ a, size_t n, unsigned int* __restrict__ amax_bits) {
  float m = 0.f;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    m = fmaxf(m, fabsf(__ldg(a + i)));
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int wv = 1; wv < 8; ++wv) m = fmaxf(m, red[wv]);
    if (m > 0.f && m < 3.0e38f) atomicMax(amax_bits, __float_as_uint(m));
  }
}

// Loss-scale selection for 16-bit gradient operands.  amax_bits = max |dy*y*(1-y)| as float bits (non-negative floats
// order like unsigned ints; written by head_grad_stats_kernel), then scale[0] = S = 2^floor(log2(target/amax)) and
// scale[1] = 1/S; a positive `fixed` overrides it (clstm_config_t.grad_scale).  S is a power of two, so scaling and
// un-scaling are exact.
__global__ void choose_scale_kernel(const unsigned int* __restrict__ amax_bits, float* __restrict__ scale,
                                    float target, float fixed) {
  float s = fixed;
  if (!(fixed > 0.f)) {
    const float amax = __uint_as_float(*amax_bits);
    s = 1.f;
    if (amax > 0.f && amax < 3.0e38f) s = exp2f(floorf(log2f(target / amax)));
    s = fminf(fmaxf(s, 1.0e-30f), 1.0e30f);
  }
  scale[0] = s;
  scale[1] = 1.f / s;
}

}  // namespace clstm
