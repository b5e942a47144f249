// dahitra_b200 — auxiliary kernels next to the hot path: device-side confusion matrix.
//
// Replaces the per-batch  argmax -> .cpu().numpy() -> np.bincount  of the reference evaluator
// (models/evaluator.py:95-104, misc/metric_tool.py:141-158):
//     mask = (gt >= 0) & (gt < nc);  cm[gt[mask]][pred[mask]] += 1
// pred / gt are uint8 class maps (gt values >= nc, e.g. the 255 "ignore" label, are masked out).  The counts
// are ACCUMULATED into a caller-owned int64 [nc][nc] matrix, so one matrix can collect a whole evaluation run.
#include "common.cuh"

namespace {
__global__ void __launch_bounds__(256)
confusion_kernel(const uint8_t* __restrict__ pred, const uint8_t* __restrict__ gt, long long n, int nc,
                 unsigned long long* __restrict__ cm) {
  __shared__ unsigned int h[64];
  if (threadIdx.x < 64) h[threadIdx.x] = 0u;
  __syncthreads();
  const long long nvec = n / 16;
  const uint4* p4 = reinterpret_cast<const uint4*>(pred);
  const uint4* g4 = reinterpret_cast<const uint4*>(gt);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    const uint4 pv = __ldg(p4 + i), gv = __ldg(g4 + i);
    const unsigned int pw[4] = {pv.x, pv.y, pv.z, pv.w}, gw[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
    for (int w = 0; w < 4; ++w)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const unsigned int g = (gw[w] >> (8 * b)) & 0xffu, p = (pw[w] >> (8 * b)) & 0xffu;
        if (g < (unsigned)nc && p < (unsigned)nc) atomicAdd(&h[g * nc + p], 1u);
      }
  }
  if (blockIdx.x == 0) {   // tail
    for (long long i = nvec * 16 + threadIdx.x; i < n; i += blockDim.x) {
      const unsigned int g = gt[i], p = pred[i];
      if (g < (unsigned)nc && p < (unsigned)nc) atomicAdd(&h[g * nc + p], 1u);
    }
  }
  __syncthreads();
  if (threadIdx.x < nc * nc && h[threadIdx.x]) atomicAdd(&cm[threadIdx.x], (unsigned long long)h[threadIdx.x]);
}
}  // namespace

extern "C" int dahitra_confusion_matrix(const unsigned char* pred, const unsigned char* gt, long long n, int nc,
                                        long long* cm, void* stream) {
  DH_REQUIRE(pred && gt && cm, DH_E_NULL);
  DH_REQUIRE(n >= 0 && nc >= 1 && nc <= 8, DH_E_SHAPE);
  DH_REQUIRE(dh_aligned16(pred) && dh_aligned16(gt), DH_E_ALIGN);
  if (n == 0) return 0;
  long long blocks = (n / 16 + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 8) blocks = 148 * 8;
  confusion_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(pred, gt, n, nc, reinterpret_cast<unsigned long long*>(cm));
  DH_CHECK_LAUNCH();
  return 0;
}
