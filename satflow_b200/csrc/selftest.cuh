// Bring-up experiment (not on the product path): can a tcgen05 shared-memory descriptor address a
// K-major, 128-byte-swizzled operand that starts at a row which is NOT a multiple of 8 (i.e. a start
// address that is not 1024-byte aligned)?  If yes, a convolution's horizontal taps can be served as
// shifted views of ONE halo tile in shared memory instead of one TMA load per tap (DESIGN.md
// "Halo-stationary A").  Variant 0 leaves the descriptor's base_offset field at 0, variant 1 sets it to
// (start_address >> 7) & 7.  The kernel reports max |D - expected| per (variant, shift).
#pragma once
#include <atomic>

#include "ptx.cuh"

namespace clstm {

__device__ __forceinline__ uint32_t sw128_offset(int row, int col) {  // fp16 element (row, col) of a 64-wide tile
  return static_cast<uint32_t>(row) * 128u + ((static_cast<uint32_t>(col >> 3) ^ (static_cast<uint32_t>(row) & 7u)) << 4) +
         (static_cast<uint32_t>(col) & 7u) * 2u;
}

__global__ void __launch_bounds__(128, 1) shifted_desc_kernel(float* out, int n_variants, int n_shifts) {
  extern __shared__ uint8_t st_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(st_smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int kRowsA = 128 + 16;
  uint8_t* sA = smem;                    // kRowsA x 128 B
  uint8_t* sB = smem + 20 * 1024;        // 64 x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 28 * 1024);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < kRowsA * 64; i += blockDim.x) {
    const int r = i / 64, c = i % 64;
    *reinterpret_cast<__half*>(sA + sw128_offset(r, c)) = __float2half_rn(static_cast<float>((r * 7 + c * 3) % 17 - 8));
  }
  for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) {
    const int n = i / 64, c = i % 64;
    *reinterpret_cast<__half*>(sB + sw128_offset(n, c)) = __float2half_rn(static_cast<float>((n * 5 + c) % 13 - 6));
  }
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 64);
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t idesc = make_idesc(FMT_F16, 128, 64, 0, 0);
  uint32_t phase = 0;

  for (int v = 0; v < n_variants; ++v) {
    for (int s = 0; s < n_shifts; ++s) {
      if (threadIdx.x == 0) {
        const uint32_t a_addr = smem_u32(sA) + static_cast<uint32_t>(s) * 128u;
        uint64_t adesc = make_smem_desc_sw128(a_addr, 16, 1024);
        if (v == 1) adesc |= static_cast<uint64_t>((a_addr >> 7) & 7u) << 49;
        const uint64_t bdesc = make_smem_desc_sw128(smem_u32(sB), 16, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, k != 0);
        umma_commit(bar);
      }
      mbar_wait(bar, phase);
      phase ^= 1;
      tcgen05_fence_after();
      const int m = warp * 32 + lane;
      float err = 0.f;
      for (int g = 0; g < 4; ++g) {
        uint32_t vv[16];
        tmem_ld16(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + g * 16, vv);
        tmem_ld_wait();
        for (int e = 0; e < 16; ++e) {
          const int n = g * 16 + e;
          float ref = 0.f;
          for (int c = 0; c < 64; ++c)
            ref += __half2float(*reinterpret_cast<const __half*>(sA + sw128_offset(m + s, c))) *
                   __half2float(*reinterpret_cast<const __half*>(sB + sw128_offset(n, c)));
          err = fmaxf(err, fabsf(ref - __uint_as_float(vv[e])));
        }
      }
      for (int o = 16; o > 0; o >>= 1) err = fmaxf(err, __shfl_xor_sync(0xffffffffu, err, o));
      if (lane == 0) atomicMax(reinterpret_cast<unsigned int*>(out + v * n_shifts + s), __float_as_uint(err));
      tcgen05_fence_before();
      __syncthreads();
      tcgen05_fence_after();
    }
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 64);
}

// out must hold n_variants * n_shifts floats (device memory).
inline int run_shifted_desc_selftest(float* out, int n_variants, int n_shifts, cudaStream_t st,
                                     std::atomic<uint64_t>* launches) {
  if (n_variants > 2) n_variants = 2;
  if (n_shifts > 16) n_shifts = 16;
  if (cudaMemsetAsync(out, 0, sizeof(float) * n_variants * n_shifts, st) != cudaSuccess) return -1;
  shifted_desc_kernel<<<1, 128, 32 * 1024, st>>>(out, n_variants, n_shifts);
  launches->fetch_add(1);
  return cudaPeekAtLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace clstm
