// Probe 2: what makes the "shifted halo window" A operand slower than an aligned tile?  Separates the effects of
// (a) stride-byte-offset != 1024, (b) a start address that is not 1024-B aligned, (c) tcgen05.commit frequency,
// (d) cycling over several accumulators, for tcgen05.mma kind::tf32 M=128 K=8.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o umma_rate2 umma_rate2.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../dahitra_b200/csrc/tc_common.cuh"
using namespace dhtc;

__device__ uint64_t make_desc(uint32_t addr, uint32_t sbo) {
  return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

// MODE: 0 = aligned, 1 = +128 B, 3 = +512 B, 4 = conv taps (r*pitch+s)*128 with pitch = SBO/128, 6 = taps with s = 0 only
// Everything is a template parameter so that the issuing thread does no index arithmetic between MMAs.
template <int N, int SBO, int MODE, int COMMIT_EVERY, int NACC>
__global__ void rate(int iters, long long* out) {
  extern __shared__ uint8_t raw[];
  __shared__ __align__(8) uint64_t bar, bar2;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  float* f = reinterpret_cast<float*>(raw + (base - smem_u32(raw)));
  for (int i = tid; i < 100 * 1024 / 4; i += blockDim.x) f[i] = 1.0f;      // A region at 0 (64 KB), B at 64 KB (32 KB)
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); mbar_init(smem_u32(&bar2), 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&slot), 512);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {
    const uint64_t bd = make_desc(base + 65536, 1024);
    constexpr uint32_t idesc = umma_idesc_tf32(128, N);
    constexpr int pitch = SBO / 128;
    const uint64_t a0 = make_desc(base, SBO);
    const uint32_t d0 = slot;
    long long t0 = clock64();
    for (int it = 0; it < iters; it += 36) {
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        constexpr int dummy = 0; (void)dummy;
        const uint32_t sh = MODE == 1 ? 128u : MODE == 3 ? 512u : MODE == 4 ? (uint32_t)((tap / 3) * pitch + tap % 3) * 128u
                          : MODE == 6 ? (uint32_t)((tap % 3) * pitch) * 128u : 0u;
        const uint64_t a = a0 + (uint64_t)(sh >> 4);
        const uint32_t d = d0 + (uint32_t)((tap % NACC) * N);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_tf32(d, a + 2 * k, bd + 2 * k, idesc, 1u);
        if (COMMIT_EVERY && ((tap + 1) * 4) % COMMIT_EVERY == 0) umma_commit(smem_u32(&bar2));
      }
    }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    out[0] = clock64() - t0;
  }
  __syncthreads();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(slot, 512); }
}

template <int N, int SBO, int MODE, int COMMIT_EVERY, int NACC>
void run(long long* d) {
  const int iters = 2304;
  cudaFuncSetAttribute(rate<N, SBO, MODE, COMMIT_EVERY, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
  rate<N, SBO, MODE, COMMIT_EVERY, NACC><<<1, 128, 112 * 1024>>>(iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long cy = 0; cudaMemcpy(&cy, d, 8, cudaMemcpyDeviceToHost);
  const char* smn[] = {"aligned", "+128B", "", "+512B", "conv taps", "", "row taps"};
  printf("N=%3d SBO=%4d A-start %-9s commit/%-2d accs=%d : %s  %.1f cycles/MMA\n", N, SBO, smn[MODE], COMMIT_EVERY, NACC,
         cudaGetErrorString(e), (double)cy / iters);
  if (e != cudaSuccess) exit(1);
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  run<32, 1024, 0, 0, 1>(d); run<64, 1024, 0, 0, 1>(d); run<128, 1024, 0, 0, 1>(d);
  run<32, 1280, 0, 0, 1>(d); run<64, 1280, 0, 0, 1>(d); run<128, 1280, 0, 0, 1>(d);
  run<32, 2048, 0, 0, 1>(d); run<128, 2048, 0, 0, 1>(d);
  run<32, 1024, 1, 0, 1>(d); run<32, 1024, 3, 0, 1>(d); run<128, 1024, 1, 0, 1>(d); run<128, 1024, 3, 0, 1>(d);
  run<32, 1280, 4, 0, 1>(d); run<64, 1280, 4, 0, 1>(d); run<128, 1280, 4, 0, 1>(d);
  run<32, 1280, 6, 0, 1>(d); run<128, 1280, 6, 0, 1>(d);
  run<32, 1024, 4, 0, 1>(d); run<128, 1024, 4, 0, 1>(d);
  run<32, 1024, 0, 12, 1>(d); run<64, 1024, 0, 12, 1>(d); run<128, 1024, 0, 12, 1>(d); run<32, 1024, 0, 4, 1>(d); run<128, 1024, 0, 4, 1>(d);
  run<32, 1024, 0, 36, 1>(d); run<128, 1024, 0, 36, 1>(d);
  run<32, 1024, 0, 0, 2>(d); run<32, 1024, 0, 0, 4>(d); run<64, 1024, 0, 0, 2>(d); run<64, 1024, 0, 0, 4>(d); run<128, 1024, 0, 0, 2>(d);
  run<32, 1280, 4, 12, 1>(d); run<64, 1280, 4, 12, 1>(d); run<128, 1280, 4, 12, 1>(d);
  run<32, 1280, 4, 0, 4>(d); run<64, 1280, 4, 0, 4>(d);
  return 0;
}
