// Probe: how does tcgen05.mma interpret a K-major SWIZZLE_128B descriptor whose start address is NOT 1024-byte
// aligned and whose stride-byte-offset is not 1024?  (Needed for reusing a TMA-loaded activation halo across the
// 9 filter taps of a 3x3 convolution.)  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o umma_probe umma_probe.cu
//
// Shared memory holds "pixels" q = 0..255, 32 floats each, written the way TMA writes them into a 1024-aligned
// buffer: pixel q at byte q*128, its 16-byte chunk c stored at chunk position c ^ (q % 8).
// A = 128 rows: atom g (8 rows) starts at pixel q0 + g*pitch.  B = identity (32x32).  So D[m][n] must equal
// value(pixel(m), channel n) if the hardware applies the XOR on absolute address bits (or with the base_offset
// we give it).  Prints the mismatch count per variant.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../dahitra_b200/csrc/tc_common.cuh"
using namespace dhtc;

__device__ __forceinline__ float val(int q, int c) { return (float)((q % 64) * 32 + c); }

__global__ void probe(int q0, int pitch, int base_off, int* mism, float* dump) {
  extern __shared__ uint8_t raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  float* A = reinterpret_cast<float*>(raw + (base - smem_u32(raw)));       // 256 pixels * 128 B = 32 KB
  float* Bm = A + 256 * 32;                                                  // 32 rows * 128 B
  for (int i = tid; i < 256 * 32; i += 128) {
    const int q = i / 32, c = i % 32;
    A[q * 32 + ((((c >> 2) ^ (q & 7)) << 2) | (c & 3))] = val(q, c);
  }
  for (int i = tid; i < 32 * 32; i += 128) {
    const int n = i / 32, k = i % 32;
    Bm[sw128_idx(n, k)] = (n == k) ? 1.f : 0.f;
  }
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&slot), 32);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {
    const uint32_t a_addr = base + (uint32_t)q0 * 128u;
    uint64_t ad = (uint64_t)((a_addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)((pitch * 128) >> 4) << 32) |
                  ((uint64_t)1 << 46) | ((uint64_t)(base_off & 7) << 49) | ((uint64_t)2 << 61);
    const uint64_t bd = umma_desc_sw128(base + 256 * 128);
    for (int k = 0; k < 4; ++k) umma_tf32(slot, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), umma_idesc_tf32(128, 32), k ? 1u : 0u);
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  uint32_t u[32];
  tmem_ld32(slot + ((uint32_t)(warp * 32) << 16), u);
  const int m = tid, g = m / 8, i = m % 8;
  const int q = q0 + g * pitch + i;
  int bad = 0;
  for (int n = 0; n < 32; ++n) {
    if (__uint_as_float(u[n]) != val(q, n)) ++bad;
    if (dump) dump[m * 32 + n] = __uint_as_float(u[n]);
  }
  if (bad) atomicAdd(mism, bad);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(slot, 32); }
}

int main() {
  int* d_m; float* d_dump;
  cudaMalloc(&d_m, 4); cudaMalloc(&d_dump, 128 * 32 * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024);
  const int cases[][3] = {{0, 8, 0},            // sanity: aligned start, SBO = 1024
                          {0, 10, 0},           // aligned start, SBO = 1280
                          {0, 16, 0},           // aligned start, SBO = 2048
                          {3, 8, 0}, {3, 8, 3},     // start shifted by 3 rows, SBO 1024, base_offset 0 / 3
                          {3, 10, 0}, {3, 10, 3},   // dense 10-pixel halo pitch
                          {3, 16, 0}, {3, 16, 3},   // padded 16-pixel pitch
                          {11, 10, 0}, {11, 10, 3}, {21, 10, 5}, {21, 10, 0}};
  for (auto& c : cases) {
    cudaMemset(d_m, 0, 4);
    probe<<<1, 128, 40 * 1024>>>(c[0], c[1], c[2], d_m, d_dump);
    cudaError_t e = cudaDeviceSynchronize();
    int m = -1;
    cudaMemcpy(&m, d_m, 4, cudaMemcpyDeviceToHost);
    float h[128 * 32];
    cudaMemcpy(h, d_dump, sizeof(h), cudaMemcpyDeviceToHost);
    printf("q0=%2d pitch=%2d base_offset=%d : %s mismatches=%d   row8[0..3]=%.0f %.0f %.0f %.0f (expect %d..)  row1[0]=%.0f (expect %d)\n",
           c[0], c[1], c[2], cudaGetErrorString(e), m, h[8 * 32], h[8 * 32 + 1], h[8 * 32 + 2], h[8 * 32 + 3],
           ((c[0] + c[1]) % 64) * 32, h[32], ((c[0] + 1) % 64) * 32);
    if (e != cudaSuccess) break;
  }
  return 0;
}
