// Probe: tcgen05.mma kind::f16 with BF16 operands in K-major SWIZZLE_64B tiles whose rows are 64 bytes (32 bf16 channels
// of one pixel), read through shifted-window descriptors (start = any pixel, stride-byte-offset = halo pitch x 64 B) —
// the bf16 twin of the fp32 halo-reuse trick (umma_probe.cu).  The tile is written by threads with the XOR taken from
// ABSOLUTE shared-memory address bits: 16-byte chunk index ^= (address >> 7) & 3.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o umma_bf16_probe umma_bf16_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include "../dahitra_b200/csrc/tc_common.cuh"
using namespace dhtc;

__device__ __forceinline__ float val(int q, int c) { return (float)((q % 8) * 32 + c); }     // < 256: exact in bf16

// byte offset (from a 1024-aligned base) of bf16 element (row q, channel c) in the SW64 layout
__device__ __forceinline__ uint32_t sw64_off(uint32_t row_byte, int c) {
  const uint32_t chunk = (uint32_t)(c >> 3) ^ ((row_byte >> 7) & 3u);
  return row_byte + (chunk << 4) + (uint32_t)(c & 7) * 2u;
}
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__global__ void probe(int q0, int pitch, int* mism) {
  extern __shared__ uint8_t raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* A = raw + (base - smem_u32(raw));                 // 256 pixels x 64 B = 16 KB
  uint8_t* Bm = A + 256 * 64;                                // 32 rows x 64 B
  for (int i = tid; i < 256 * 32; i += 128) {
    const int q = i / 32, c = i % 32;
    *reinterpret_cast<__nv_bfloat16*>(A + sw64_off((uint32_t)q * 64u, c)) = __float2bfloat16(val(q, c));
  }
  for (int i = tid; i < 32 * 32; i += 128) {
    const int n = i / 32, k = i % 32;
    *reinterpret_cast<__nv_bfloat16*>(Bm + sw64_off((uint32_t)n * 64u, k)) = __float2bfloat16(n == k ? 1.f : 0.f);
  }
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&slot), 32);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {
    const uint32_t a_addr = base + (uint32_t)q0 * 64u, b_addr = base + 256 * 64;
    const uint64_t ad = (uint64_t)((a_addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)((pitch * 64) >> 4) << 32) |
                        ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
    const uint64_t bd = (uint64_t)((b_addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(512 >> 4) << 32) |
                        ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
    for (int k = 0; k < 2; ++k)                              // K = 16 bf16 = 32 bytes per MMA
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                   "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                   ::"r"(slot), "l"(ad + (uint64_t)(2 * k)), "l"(bd + (uint64_t)(2 * k)), "r"(idesc_bf16(128, 32)), "r"(k ? 1u : 0u) : "memory");
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  uint32_t u[32];
  tmem_ld32(slot + ((uint32_t)(warp * 32) << 16), u);
  const int m = tid, q = q0 + (m / 8) * pitch + (m % 8);
  int bad = 0;
  for (int n = 0; n < 32; ++n) if (__uint_as_float(u[n]) != val(q, n)) ++bad;
  if (bad) atomicAdd(mism, bad);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(slot, 32); }
}

int main() {
  int* d; cudaMalloc(&d, 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024);
  const int cases[][2] = {{0, 8}, {8, 8}, {0, 10}, {1, 10}, {2, 10}, {10, 10}, {11, 10}, {12, 10}, {20, 10}, {21, 10}, {22, 10}, {3, 10}, {5, 9}};
  int rc = 0;
  for (auto& c : cases) {
    cudaMemset(d, 0, 4);
    probe<<<1, 128, 32 * 1024>>>(c[0], c[1], d);
    cudaError_t e = cudaDeviceSynchronize();
    int bad = -1; cudaMemcpy(&bad, d, 4, cudaMemcpyDeviceToHost);
    printf("bf16 SW64 window: first pixel %2d, pitch %2d pixels (SBO %4d B): %s, mismatches %d / 4096\n", c[0], c[1], c[1] * 64,
           cudaGetErrorString(e), bad);
    if (e != cudaSuccess) return 1;
    rc |= bad != 0;
  }
  return rc;
}
