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

// =====================================================================================================
// Input path: uint8 HWC images -> normalised fp32 NCHW tensors on the device, optionally
// cut into square tiles, so that only the decoded bytes cross PCIe (4x less than the fp32 tensors).
//   kind 0 (LEVIR loaders, datasets/data_utils.py:104-111: TF.to_tensor + TF.normalize(0.5, 0.5)):  (x / 255 - 0.5) / 0.5
//   kind 1 (xBD, xBD_code/utils.py:112-116 preprocess_inputs):                                      x / 127 - 1
// evaluated in fp32 with IEEE division in the reference's operation order => bit-identical tensors.
// tile > 0: image n yields T = (H/tile)*(W/tile) tiles; tile p starts at x0 = tile*(p / (H/tile)), y0 = tile*(p % (H/tile))
// (datasets/data_utils.py:65-66 with patch 0 at (0,0)); output [N*T][3][tile][tile].  tile = 0: whole images.
// =====================================================================================================
namespace {
__global__ void __launch_bounds__(256)
prepare_input_kernel(const uint8_t* __restrict__ src, int H, int W, int kind, int th, int tw, int tiles_y, int T,
                     long long total4, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // one thread = 4 consecutive x of one output row
  if (i >= total4) return;
  const int w4 = tw / 4;
  const int x4 = (int)(i % w4);
  long long r = i / w4;
  const int y = (int)(r % th); r /= th;
  const int t = (int)r;                                                        // output image (tile) index
  const int n = t / T, p = t % T;
  const int x0 = tw * (p / tiles_y), y0 = th * (p % tiles_y);
  const uint8_t* s = src + (((size_t)n * H + (y0 + y)) * W + (x0 + x4 * 4)) * 3;
  uint8_t b[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) b[k] = __ldg(s + k);
  float* o = out + ((size_t)t * 3 * th + y) * tw + x4 * 4;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float f = (float)b[k * 3 + c];
      v[k] = kind == 0 ? __fdiv_rn(__fsub_rn(__fdiv_rn(f, 255.f), 0.5f), 0.5f) : __fsub_rn(__fdiv_rn(f, 127.f), 1.f);
    }
    *reinterpret_cast<float4*>(o + (size_t)c * th * tw) = make_float4(v[0], v[1], v[2], v[3]);
  }
}
}  // namespace

extern "C" int dahitra_prepare_input_u8(const unsigned char* hwc, int N, int H, int W, int kind, int tile, float* nchw,
                                        void* stream) {
  DH_REQUIRE(hwc && nchw, DH_E_NULL);
  DH_REQUIRE(N > 0 && H > 0 && W > 0 && W % 4 == 0 && (kind == 0 || kind == 1), DH_E_SHAPE);
  DH_REQUIRE(tile == 0 || (tile % 4 == 0 && H % tile == 0 && W % tile == 0), DH_E_SHAPE);
  DH_REQUIRE(dh_aligned16(nchw), DH_E_ALIGN);
  const int th = tile ? tile : H, tw = tile ? tile : W;
  const int tiles_y = H / th, T = tiles_y * (W / tw);
  const long long total4 = (long long)N * T * th * (tw / 4);
  prepare_input_kernel<<<(unsigned)((total4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(hwc, H, W, kind, th, tw, tiles_y, T,
                                                                                          total4, nchw);
  DH_CHECK_LAUNCH();
  return 0;
}
