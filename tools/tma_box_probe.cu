// Probe: which un-swizzled TMA boxes over a planar fp32 (W, H, C, N) tensor are legal?  (stem halo fetch)
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tma_box_probe tma_box_probe.cu
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void probe(const __grid_constant__ CUtensorMap tm, int cx, int cy, int bytes, int dst_off, float* out, int nfloat) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
  const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem) + dst_off;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dst), "l"(&tm), "r"(b), "r"(cx), "r"(cy), "r"(0), "r"(0) : "memory");
  }
  uint32_t done = 0; int spins = 0;
  while (!done && spins++ < 1000000)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(b) : "memory");
  for (int i = threadIdx.x; i < nfloat; i += blockDim.x) out[i] = done ? reinterpret_cast<float*>(smem + dst_off)[i] : -777.f;
}

int main() {
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fp;
  const int W = 64, H = 64, C = 3, N = 2;
  std::vector<float> h((size_t)W * H * C * N);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 100000);
  float *d, *o; cudaMalloc(&d, h.size() * 4); cudaMalloc(&o, 65536 * 4);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  struct Case { int bx, by, bc, cx, cy, off; } cases[] = {
      {32, 21, 3, 0, 0, 0}, {32, 21, 3, -4, -3, 0}, {40, 21, 3, 0, 0, 0}, {40, 21, 3, -4, -3, 0}, {64, 21, 3, -4, -3, 0},
      {40, 21, 1, -4, -3, 0}, {40, 8, 3, -4, -3, 0}, {40, 21, 3, -4, 45, 0}, {40, 21, 3, 28, 45, 0}, {40, 21, 3, -4, -3, 10112}, {48, 21, 3, -4, -3, 0}, {40, 21, 3, -3, -3, 0}};
  for (auto& c : cases) {
    CUtensorMap tm;
    cuuint64_t dims[4] = {W, H, C, N}, st[3] = {W * 4, (cuuint64_t)W * H * 4, (cuuint64_t)W * H * C * 4};
    cuuint32_t box[4] = {(cuuint32_t)c.bx, (cuuint32_t)c.by, (cuuint32_t)c.bc, 1}, es[4] = {1, 1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    const int nfl = c.bx * c.by * c.bc;
    printf("box {%d,%d,%d,1} at (%d,%d) dst+%d: encode %d; ", c.bx, c.by, c.bc, c.cx, c.cy, c.off, (int)r);
    if (r != CUDA_SUCCESS) { printf("\n"); continue; }
    probe<<<1, 128, 64 * 1024>>>(tm, c.cx, c.cy, nfl * 4, c.off, o, nfl);
    cudaError_t e = cudaDeviceSynchronize();
    printf("run: %s", cudaGetErrorString(e));
    if (e != cudaSuccess) { printf("\n"); return 1; }
    std::vector<float> got(nfl);
    cudaMemcpy(got.data(), o, nfl * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int ci = 0; ci < c.bc; ++ci) for (int y = 0; y < c.by; ++y) for (int x = 0; x < c.bx; ++x) {
      const int gx = c.cx + x, gy = c.cy + y;
      const float want = (gx >= 0 && gx < W && gy >= 0 && gy < H) ? h[((size_t)ci * H + gy) * W + gx] : 0.f;
      if (got[(ci * c.by + y) * c.bx + x] != want) ++bad;
    }
    printf("; mismatches vs dense [c][y][x] layout: %d / %d (first value %.0f)\n", bad, nfl, got[0]);
  }
  return 0;
}
