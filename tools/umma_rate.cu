// Probe: issue rate of tcgen05.mma kind::tf32 (M=128, K=8) as a function of N and of the shared-memory operand
// layout (swizzle mode / row pitch).  One CTA, one issuing thread, back-to-back MMAs on resident operands.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o umma_rate umma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../dahitra_b200/csrc/tc_common.cuh"
using namespace dhtc;

// layout_type: 0 none(interleave), 6 = 32B, 4 = 64B, 2 = 128B swizzle.  row_bytes = bytes of one K-major row in the
// swizzle atom (16 for "none": core matrices 8 rows x 16 B).
__device__ uint64_t make_desc(uint32_t addr, int layout_type, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)layout_type << 61);
}

__global__ void rate(int N, int layout_type, int iters, int kadv_bytes, int same_k, long long* out) {
  extern __shared__ uint8_t raw[];
  __shared__ __align__(8) uint64_t bar, bar2;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  float* f = reinterpret_cast<float*>(raw + (base - smem_u32(raw)));
  for (int i = tid; i < 56 * 1024 / 4; i += blockDim.x) f[i] = 1.0f;       // A (halo) at 0 (23 KB), B at 24 KB (32 KB)
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); mbar_init(smem_u32(&bar2), 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&slot), 256);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {
    uint32_t lbo, sbo;
    switch (layout_type) {
      case 2: lbo = 16; sbo = 1024; break;      // 128B swizzle: rows of 128 B, 8-row atoms of 1024 B
      case 4: lbo = 16; sbo = 512; break;       // 64B swizzle: rows of 64 B
      case 6: lbo = 16; sbo = 256; break;       // 32B swizzle: rows of 32 B
      default: lbo = 128; sbo = 256; break;     // interleave: core matrix 8x16B = 128 B; K-adjacent core +128, M-adjacent +256
    }
    const uint64_t ad = make_desc(base, layout_type, lbo, sbo), bd = make_desc(base + 24576, layout_type, lbo, sbo);
    const uint32_t idesc = umma_idesc_tf32(128, N);
    long long t0 = clock64();
    if (same_k == 2) {
      // conv-like: A = shifted windows of a halo (SBO = 1280 B, start offset (r*10+s)*128 B), 9 taps x 4 k-steps, a
      // commit every 12 MMAs (never waited on), accumulator reset every 36 MMAs
      const uint64_t ah = make_desc(base, 2, 16, 1280);
      int i = 0;
      while (i < iters) {
        for (int tap = 0; tap < 9 && i < iters; ++tap) {
          const uint64_t a = ah + (uint64_t)((((tap / 3) * 10 + tap % 3) * 128) >> 4);
          for (int k = 0; k < 4; ++k, ++i) umma_tf32(slot, a + 2 * k, bd + 2 * k, idesc, (tap | k) ? 1u : 0u);
          if (tap % 3 == 2) umma_commit(smem_u32(&bar2));
        }
      }
    } else {
      for (int i = 0; i < iters; ++i) {
        const uint64_t adv = same_k ? 0 : (uint64_t)(((i & 3) * kadv_bytes) >> 4);
        umma_tf32(slot, ad + adv, bd + adv, idesc, 1u);
      }
    }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    long long t1 = clock64();
    out[0] = t1 - t0;
  } else if (tid == 32) {
    // nothing: keeps a second warp alive
  }
  __syncthreads();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(slot, 256); }
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int iters = 2048;
  const int Ns[] = {32, 64, 128, 256};
  struct { int lt; int kadv; const char* name; } lays[] = {{2, 32, "SW128 (k advance 32B in row)"}, {4, 32, "SW64"}, {6, 0, "SW32 (one row per MMA)"}, {0, 0, "INTERLEAVE"}};
  for (auto& L : lays)
    for (int N : Ns) {
      for (int same = 0; same < 3; ++same) {
        if (same == 2 && L.lt != 2) continue;
        rate<<<1, 128, 64 * 1024>>>(N, L.lt, iters, L.kadv, same, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long c = 0; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
        printf("%-30s N=%3d %s : %s  %.1f cycles/MMA (floor %d)\n", L.name, N, same == 2 ? "convlike" : (same ? "same-k " : "k-cycle"), cudaGetErrorString(e),
               (double)c / iters, 64 * N / 128);
        if (e != cudaSuccess) return 1;
      }
    }
  return 0;
}
