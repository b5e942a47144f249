// Probe: tcgen05.mma.cta_group::2 (CTA pair, M = 256) with kind::tf32 — where do the operands and the result live?
//   * each CTA writes its own 128 x 32 A tile (K-major SWIZZLE_128B) and HALF of the B tile into its own shared memory
//     at identical offsets; which half (rows [0,N/2) or [N/2,N) of B) each CTA must hold is what the probe reports;
//   * the leader (cluster rank 0) issues 4 MMAs (K = 32) and commits with multicast to both CTAs' mbarriers;
//   * each CTA reads D rows from ITS tensor memory and compares against a host reference.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o umma_2cta_probe umma_2cta_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "../dahitra_b200/csrc/tc_common.cuh"
using namespace dhtc;

__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
probe(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D) {
  // A: [256][32] (rows 0..127 -> CTA 0, 128..255 -> CTA 1), B: [N][32], D: [256][N]
  extern __shared__ uint8_t raw[];
  __shared__ __align__(8) uint64_t done_bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = cluster_rank();
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  float* a_s = reinterpret_cast<float*>(raw + (base - smem_u32(raw)));     // 16 KB
  float* b_s = a_s + 128 * 32;                                              // N/2 rows x 32 floats
  for (int i = tid; i < 128 * 32; i += 128) { const int r = i >> 5, k = i & 31; a_s[sw128_idx(r, k)] = A[(rank * 128 + r) * 32 + k]; }
  for (int i = tid; i < (N / 2) * 32; i += 128) { const int r = i >> 5, k = i & 31; b_s[sw128_idx(r, k)] = B[(rank * (N / 2) + r) * 32 + k]; }
  if (tid == 0) { mbar_init(smem_u32(&done_bar), 1); mbar_fence_init(); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                       // both CTAs' operands are in place
  tc_fence_after();
  const uint32_t tmem = slot;
  if (rank == 0 && tid == 0) {
    const uint64_t ad = umma_desc_sw128(base), bd = umma_desc_sw128(base + 128 * 128);
    constexpr uint32_t idesc = umma_idesc_tf32(256, N);
    for (int k = 0; k < 4; ++k)
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                   "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                   ::"r"(tmem), "l"(ad + (uint64_t)(2 * k)), "l"(bd + (uint64_t)(2 * k)), "r"(idesc), "r"(k ? 1u : 0u) : "memory");
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(&done_bar)), "h"((uint16_t)3) : "memory");
  }
  mbar_wait(smem_u32(&done_bar), 0);
  tc_fence_after();
  for (int j = 0; j < N / 32; ++j) {
    uint32_t v[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(j * 32), v);
    for (int c = 0; c < 32; ++c) D[(size_t)(rank * 128 + tid) * N + j * 32 + c] = __uint_as_float(v[c]);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
  }
}

template <int N>
int run() {
  std::vector<float> A(256 * 32), B(N * 32), D(256 * N), R(256 * N);
  srand(1);
  auto q = [](float v) { uint32_t u; memcpy(&u, &v, 4); u &= 0xFFFFE000u; memcpy(&v, &u, 4); return v; };   // exact in TF32
  for (auto& v : A) v = q((rand() % 2001 - 1000) / 1000.f);
  for (auto& v : B) v = q((rand() % 2001 - 1000) / 1000.f);
  for (int m = 0; m < 256; ++m)
    for (int n = 0; n < N; ++n) { double s = 0; for (int k = 0; k < 32; ++k) s += (double)A[m * 32 + k] * B[n * 32 + k]; R[m * N + n] = (float)s; }
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, D.size() * 4);
  const int smem = 16384 + (N / 2) * 128 + 1024;
  cudaFuncSetAttribute(probe<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe<N><<<2, 128, smem>>>(dA, dB, dD);
  cudaError_t e = cudaDeviceSynchronize();
  printf("N=%d: %s\n", N, cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  // straight mapping, and the alternatives in case rows/columns land elsewhere
  double err = 0; int bad = 0;
  for (int i = 0; i < 256 * N; ++i) { const double d = fabs((double)D[i] - R[i]); if (d > err) err = d; if (d > 1e-3) ++bad; }
  printf("  straight mapping (CTA r: rows 128r.., B half r = rows r*N/2..): max|d| = %.3e, mismatches %d / %d\n", err, bad, 256 * N);
  if (bad) {
    for (int m : {0, 1, 127, 128, 255}) {
      printf("  row %3d got:", m); for (int n = 0; n < 6; ++n) printf(" %8.4f", D[m * N + n]); printf(" ... %8.4f %8.4f\n", D[m * N + N / 2], D[m * N + N - 1]);
      printf("      want:"); for (int n = 0; n < 6; ++n) printf(" %8.4f", R[m * N + n]); printf(" ... %8.4f %8.4f\n", R[m * N + N / 2], R[m * N + N - 1]);
    }
  }
  return bad ? 2 : 0;
}

int main() {
  int rc = 0;
  rc |= run<128>();
  rc |= run<64>();
  rc |= run<32>();
  rc |= run<256>();
  return rc;
}
