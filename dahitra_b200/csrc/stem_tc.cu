// dahitra_b200 — stem (7x7 stride-2 pad-3 conv 3 -> 64 + folded BN + ReLU) on the tensor cores.
//
// Replaces resnet.conv1 + bn1 + relu (reference models/networks.py:1120-1122, models/resnet.py:150-153).
// GEMM view: D[128 pixels][64] = A[128][K=147 -> 160] . Wt[64][160]^T, K ordered (r, s, ci) like DH_W_STEM_W.
// C_in = 3 cannot feed a TMA/UMMA row directly, so the im2col rows are assembled ON CHIP: the CTA stages the
// 21x37x3 input halo of its 8x16 output patch in shared memory once (NCHW planes in, the only HBM read), and
// each thread gathers the 160 taps of ITS pixel into the K-major SWIZZLE_128B A tile, 32 taps (one 128-byte
// row) per K step, double-buffered against the MMAs.  The filter arrives pre-swizzled (5 tiles of 64x32)
// through one bulk copy.  Accumulator in TMEM (64 columns); epilogue = bias + ReLU, NHWC stores.
#include "tc_common.cuh"

using namespace dhtc;

namespace {
constexpr int SK_TH = 8, SK_TW = 16;                 // output patch
constexpr int SK_HR = 2 * SK_TH + 5;                 // 21 halo rows
constexpr int SK_HC = 2 * SK_TW + 5;                 // 37 halo cols
constexpr int SK_HCP = 40;                           // padded row stride
constexpr int SK_PLANE = SK_HR * SK_HCP;             // 840 floats per channel
constexpr int SK_KSTEPS = 5;                         // 160 / 32
constexpr uint32_t SK_A_BYTES = 128 * 128;           // one A tile
constexpr uint32_t SK_B_BYTES = SK_KSTEPS * 64 * 128;    // 40 KB filter image
constexpr uint32_t SK_HALO_BYTES = (3 * SK_PLANE + 8) * 4;   // + a zero slot for the K padding
constexpr uint32_t SK_IDESC = umma_idesc_tf32(128, 64);
template <bool X3> struct SkCfg {       // X3: A tiles and the filter image come as TF32 hi + lo pairs (3 MMAs per product)
  static constexpr uint32_t NA = X3 ? 4 : 2;                      // A tiles: [buf0 hi, buf1 hi, buf0 lo, buf1 lo]
  static constexpr uint32_t NB = X3 ? 2 : 1;
  static constexpr uint32_t SMEM = NA * SK_A_BYTES + NB * SK_B_BYTES + 1024 + ((SK_HALO_BYTES + 15) & ~15u) + 160 * 4;
};

__device__ __forceinline__ float sk_tf32(float v) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
  return __uint_as_float(u);
}

template <bool X3>
__global__ void __launch_bounds__(128, X3 ? 1 : 2)
stem_tc_kernel(const float* __restrict__ x, long long xbs, int H, int W, int OH, int OW, int tilesX,
               const float* __restrict__ wtc, const float* __restrict__ bias, float* __restrict__ out) {
  extern __shared__ uint8_t sk_raw[];
  __shared__ __align__(8) uint64_t w_bar, free_bar[2], acc_bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int n = blockIdx.z;
  const int oy0 = (blockIdx.x / tilesX) * SK_TH, ox0 = (blockIdx.x % tilesX) * SK_TW;
  const uint32_t base = (smem_u32(sk_raw) + 1023u) & ~1023u;
  uint8_t* bp = sk_raw + (base - smem_u32(sk_raw));
  using Cfg = SkCfg<X3>;
  float* a_tile[2] = {reinterpret_cast<float*>(bp), reinterpret_cast<float*>(bp + SK_A_BYTES)};
  const uint32_t a_addr[2] = {base, base + SK_A_BYTES};
  constexpr uint32_t A_LO = 2 * SK_A_BYTES;                            // lo twins of the two A tiles (X3)
  const uint32_t b_addr = base + Cfg::NA * SK_A_BYTES;
  constexpr uint32_t OFF_H = Cfg::NA * SK_A_BYTES + Cfg::NB * SK_B_BYTES;
  float* halo = reinterpret_cast<float*>(bp + OFF_H);
  int* koff = reinterpret_cast<int*>(bp + OFF_H + ((SK_HALO_BYTES + 15) & ~15u));

  if (tid == 0) {
    mbar_init(smem_u32(&w_bar), 1); mbar_init(smem_u32(&free_bar[0]), 1); mbar_init(smem_u32(&free_bar[1]), 1);
    mbar_init(smem_u32(&acc_bar), 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {
    mbar_expect_tx(smem_u32(&w_bar), Cfg::NB * SK_B_BYTES);
    bulk_load_1d(b_addr, wtc, SK_B_BYTES, smem_u32(&w_bar));
    if (X3) bulk_load_1d(b_addr + SK_B_BYTES, wtc + SK_B_BYTES / 4, SK_B_BYTES, smem_u32(&w_bar));
  }
  // halo: input rows 2*oy0-3 .. +20, cols 2*ox0-3 .. +36, zero outside the image
  const float* xn = x + (size_t)n * xbs;
  const int iy0 = 2 * oy0 - 3, ix0 = 2 * ox0 - 3;
  for (int i = tid; i < 3 * SK_HR * SK_HC; i += 128) {
    const int ci = i / (SK_HR * SK_HC), rem = i - ci * (SK_HR * SK_HC);
    const int yy = rem / SK_HC, xx = rem - yy * SK_HC;
    const int iy = iy0 + yy, ix = ix0 + xx;
    float v = 0.f;
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = __ldg(xn + ((size_t)ci * H + iy) * W + ix);
    halo[ci * SK_PLANE + yy * SK_HCP + xx] = v;
  }
  if (tid < 8) halo[3 * SK_PLANE + tid] = 0.f;                       // zero slot (K padding 147..159)
  for (int k = tid; k < 160; k += 128) {
    int off = 3 * SK_PLANE;                                          // -> zero slot (independent of the pixel)
    if (k < 147) { const int tap = k / 3, ci = k - tap * 3, r = tap / 7, s = tap - r * 7; off = ci * SK_PLANE + r * SK_HCP + s; }
    koff[k] = off;
  }
  __syncthreads();

  const int py = tid / SK_TW, px = tid % SK_TW;
  const int pbase = (2 * py) * SK_HCP + 2 * px;
  for (int kt = 0; kt < SK_KSTEPS; ++kt) {
    const int buf = kt & 1;
    if (kt >= 2) mbar_wait(smem_u32(&free_bar[buf]), (uint32_t)(((kt >> 1) - 1) & 1));   // MMAs of step kt-2 have read it
    float v[32];
#pragma unroll
    for (int kk = 0; kk < 32; ++kk) {
      const int off = koff[kt * 32 + kk];
      v[kk] = halo[off + (off < 3 * SK_PLANE ? pbase : 0)];
    }
    float* at = a_tile[buf];
#pragma unroll
    for (int ch = 0; ch < 8; ++ch) {
      const int o = tid * 32 + ((ch ^ (tid & 7)) << 2);
      const float h0 = sk_tf32(v[ch * 4]), h1 = sk_tf32(v[ch * 4 + 1]), h2 = sk_tf32(v[ch * 4 + 2]), h3 = sk_tf32(v[ch * 4 + 3]);
      *reinterpret_cast<float4*>(at + o) = make_float4(h0, h1, h2, h3);
      if (X3) *reinterpret_cast<float4*>(at + A_LO / 4 + o) = make_float4(sk_tf32(v[ch * 4] - h0), sk_tf32(v[ch * 4 + 1] - h1),
                                                                        sk_tf32(v[ch * 4 + 2] - h2), sk_tf32(v[ch * 4 + 3] - h3));
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      if (kt == 0) mbar_wait(smem_u32(&w_bar), 0);
      tc_fence_after();
      const uint64_t ad = umma_desc_sw128(a_addr[buf]), bd = umma_desc_sw128(b_addr + (uint32_t)kt * 64 * 128);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_tf32(tmem_slot, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), SK_IDESC, (kt | k) ? 1u : 0u);
      if (X3) {
        const uint64_t al = umma_desc_sw128(a_addr[buf] + A_LO), bl = umma_desc_sw128(b_addr + SK_B_BYTES + (uint32_t)kt * 64 * 128);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_tf32(tmem_slot, al + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), SK_IDESC, 1u);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_tf32(tmem_slot, ad + (uint64_t)(2 * k), bl + (uint64_t)(2 * k), SK_IDESC, 1u);
      }
      umma_commit(smem_u32(&free_bar[buf]));
      if (kt == SK_KSTEPS - 1) umma_commit(smem_u32(&acc_bar));
    }
  }
  mbar_wait(smem_u32(&acc_bar), 0);
  tc_fence_after();
  const int oy = oy0 + py, ox = ox0 + px;
  const bool valid = oy < OH && ox < OW;
  const uint32_t tm = tmem_slot + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
  for (int j = 0; j < 2; ++j) {
    uint32_t u[32];
    tmem_ld32(tm + (uint32_t)(j * 32), u);
    if (valid) {
      float* op = out + ((size_t)(n * OH + oy) * OW + ox) * 64 + j * 32;
#pragma unroll
      for (int c4 = 0; c4 < 8; ++c4) {
        const float4 b = ldg4(bias + j * 32 + c4 * 4);
        st4(op + c4 * 4, make_float4(fmaxf(__uint_as_float(u[c4 * 4]) + b.x, 0.f), fmaxf(__uint_as_float(u[c4 * 4 + 1]) + b.y, 0.f),
                                     fmaxf(__uint_as_float(u[c4 * 4 + 2]) + b.z, 0.f), fmaxf(__uint_as_float(u[c4 * 4 + 3]) + b.w, 0.f)));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_slot, 64);
  }
}
}  // namespace

int dh_launch_stem_tc(const float* x, long long xbs, int N, int H, int W, const float* wtc, const float* b, float* out,
                      int x3, cudaStream_t s) {
  DH_REQUIRE(x && wtc && b && out, DH_E_NULL);
  DH_REQUIRE(N > 0 && H >= 8 && W >= 8 && H % 2 == 0 && W % 2 == 0, DH_E_SHAPE);
  DH_REQUIRE(dh_aligned16(wtc) && dh_aligned16(b) && dh_aligned16(out), DH_E_ALIGN);
  const int OH = H / 2, OW = W / 2;
  const int tx = dh_cdiv(OW, SK_TW), ty = dh_cdiv(OH, SK_TH);
  dim3 grid(tx * ty, 1, N);
  if (x3) {
    cudaError_t e = cudaFuncSetAttribute(stem_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SkCfg<true>::SMEM);
    if (e != cudaSuccess) return (int)e;
    stem_tc_kernel<true><<<grid, 128, SkCfg<true>::SMEM, s>>>(x, xbs, H, W, OH, OW, tx, wtc, b, out);
  } else {
    cudaError_t e = cudaFuncSetAttribute(stem_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SkCfg<false>::SMEM);
    if (e != cudaSuccess) return (int)e;
    stem_tc_kernel<false><<<grid, 128, SkCfg<false>::SMEM, s>>>(x, xbs, H, W, OH, OW, tx, wtc, b, out);
  }
  DH_CHECK_LAUNCH();
  return 0;
}
