// dahitra_b200 — implicit-GEMM convolution on the 5th-gen tensor cores (sm_100a):
// TMA (cp.async.bulk.tensor) -> 128B-swizzled shared memory -> tcgen05.mma kind::tf32 -> TMEM accumulator
// -> tcgen05.ld epilogue with folded-BN bias / residual / ReLU fused, NHWC fp32 in and out.
//
// GEMM view (stride-1 convolution, pad = K/2):  D[m][co] = sum_{tap, ci} A_tap[m][ci] * Wt[co][tap*Cin + ci]
//   M tile  = 128 output pixels = an 8 x 16 spatial patch of one image
//   N tile  = NT output channels (32 / 64 / 128)
//   K step  = 32 input channels of one filter tap (32 fp32 = 128 B = one swizzle row)
// A_tap is never materialised: for tap (r, s) the A tile is the 8x16 input patch shifted by (r-pad, s-pad),
// fetched by ONE 4-D TMA box {32 ch, 16, 8, 1} over the NHWC tensor.  Out-of-image coordinates (the zero
// padding, and ragged tile edges) are zero-filled by the TMA unit, so there is no im2col buffer, no halo
// logic and no predication in the main loop.  The virtual channel concat of two tensors (conv_decode,
// conv_layer2_0) is two tensor maps.  The filter is a 2-D K-major tensor [Cout][K] (box {32, NT}).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (one lane),
// warps 2..5 = epilogue (each owns the 32 TMEM lanes matching warp_id % 4).
// Pipeline: STAGES-deep smem ring with full/empty mbarriers; tcgen05.commit releases a stage when the MMAs
// that read it have retired, and signals the epilogue after the last K step.
#include "tc_common.cuh"   // mbarrier / TMA / UMMA / TMEM helpers shared with conv_tc2.cu
#include <mutex>
#include <unordered_map>

using namespace dhtc;

namespace {

constexpr int TC_TH = 8, TC_TW = 16;          // output patch
constexpr int TC_M = TC_TH * TC_TW;            // 128 rows
constexpr uint32_t TC_A_BYTES = TC_M * 128;    // 16 KB per stage

struct TcEpilogue {
  const float* bias; const float* res; float* out;
  int OH, OW, Cout, relu, tilesX, nsteps, cchunks0, cchunks, KW, pad, Cin;
  int stride;   // 1 or 2: the A tensor maps then carry elementStrides {1,s,s,1} and a box of 16s x 8s input pixels
  int ps;       // pixel-shuffle store (ConvArgs::ps)
};

template <int NT> struct TcCfg {
  static constexpr int STAGES = (NT == 128) ? 3 : (NT == 64 ? 4 : 5);      // ~96-100 KB -> 2 CTAs per SM
  static constexpr uint32_t B_BYTES = NT * 128;
  static constexpr uint32_t STAGE_BYTES = TC_A_BYTES + B_BYTES;
  static constexpr uint32_t SMEM = STAGES * STAGE_BYTES + 1024;             // + slack for the 1024 B alignment
  // instruction descriptor: D=F32 (bit 4), A=B=TF32 (2 at bits 7 and 10), K-major both, N>>3 at 17, M>>4 at 24
  static constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
};

template <int NT>
__global__ void __launch_bounds__(192, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
               const __grid_constant__ CUtensorMap tmB, const TcEpilogue e) {
  using Cfg = TcCfg<NT>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t tc_smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], accum_bar;
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tiles = (smem_u32(tc_smem_raw) + 1023u) & ~1023u;
  const int n = blockIdx.z;
  const int n0 = blockIdx.y * NT;
  const int oy0 = (blockIdx.x / e.tilesX) * TC_TH, ox0 = (blockIdx.x % e.tilesX) * TC_TW;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
    mbar_init(smem_u32(&accum_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA0) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  }
  if (warp == 1) {   // TMEM: NT fp32 accumulator columns x 128 lanes
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"((uint32_t)NT) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    if (lane == 0) {                                          // ---------------- TMA producer
      for (int ks = 0; ks < e.nsteps; ++ks) {
        const int st = ks % STAGES, round = ks / STAGES;
        mbar_wait(smem_u32(&empty_bar[st]), (round & 1) ^ 1);
        const int tap = ks / e.cchunks, cc = ks - tap * e.cchunks;
        const int r = tap / e.KW, s = tap - r * e.KW;
        const uint32_t a_dst = tiles + st * Cfg::STAGE_BYTES, b_dst = a_dst + TC_A_BYTES;
        const uint32_t bar = smem_u32(&full_bar[st]);
        mbar_expect_tx(bar, Cfg::STAGE_BYTES);
        const int ix = ox0 * e.stride + s - e.pad, iy = oy0 * e.stride + r - e.pad;
        if (cc < e.cchunks0) tma_load_4d(a_dst, &tmA0, bar, cc * 32, ix, iy, n);
        else                 tma_load_4d(a_dst, &tmA1, bar, (cc - e.cchunks0) * 32, ix, iy, n);
        tma_load_2d(b_dst, &tmB, bar, tap * e.Cin + cc * 32, n0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {                                          // ---------------- MMA issuer
      for (int ks = 0; ks < e.nsteps; ++ks) {
        const int st = ks % STAGES, round = ks / STAGES;
        mbar_wait(smem_u32(&full_bar[st]), round & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_addr = tiles + st * Cfg::STAGE_BYTES;
        const uint64_t adesc = umma_desc_sw128(a_addr), bdesc = umma_desc_sw128(a_addr + TC_A_BYTES);
#pragma unroll
        for (int k = 0; k < 4; ++k)   // UMMA_K = 8 tf32 = 32 B: advance the start address inside the swizzle row
          umma_tf32(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), Cfg::IDESC, (ks | k) ? 1u : 0u);
        umma_commit(smem_u32(&empty_bar[st]));                // frees the stage once these MMAs have read it
      }
      umma_commit(smem_u32(&accum_bar));                      // accumulator complete
    }
  } else {                                                    // ---------------- epilogue (warps 2..5)
    const int q = warp & 3;                                   // TMEM lane quarter this warp may read
    const int m = q * 32 + lane;
    const int oy = oy0 + m / TC_TW, ox = ox0 + m % TC_TW;
    const bool valid = (oy < e.OH) && (ox < e.OW);
    mbar_wait(smem_u32(&accum_bar), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const size_t row = ((size_t)(n * e.OH + oy) * e.OW + ox) * e.Cout + n0;
#pragma unroll 1
    for (int j = 0; j < NT / 32; ++j) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(j * 32), v);
      if (valid) {
        float* op = e.out + row + j * 32;
        if (e.ps)   // channel block j = output-pixel phase (j/2, j%2) of the x2-upsampled result, 32 channels each
          op = e.out + ((size_t)(n * 2 * e.OH + 2 * oy + (j >> 1)) * (2 * e.OW) + 2 * ox + (j & 1)) * 32;
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
          float4 o = make_float4(__uint_as_float(v[c4 * 4]), __uint_as_float(v[c4 * 4 + 1]),
                                 __uint_as_float(v[c4 * 4 + 2]), __uint_as_float(v[c4 * 4 + 3]));
          if (e.bias) { const float4 b = ldg4(e.bias + n0 + j * 32 + c4 * 4); o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w; }
          if (e.res) { const float4 rr = ldg4(e.res + row + j * 32 + c4 * 4); o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w; }
          if (e.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
          st4(op + c4 * 4, o);
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)NT) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  });
  return fn;
}

struct MapKey {
  const void* ptr; int d0, d1, d2, d3, b1, b2, es;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && d0 == o.d0 && d1 == o.d1 && d2 == o.d2 && d3 == o.d3 && b1 == o.b1 && b2 == o.b2 && es == o.es;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = (size_t)k.ptr;
    for (int v : {k.d0, k.d1, k.d2, k.d3, k.b1, k.b2, k.es}) h = h * 1000003u ^ (size_t)v;
    return h;
  }
};
std::mutex g_map_mu;
std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;

// NHWC activation map: dims (C, W, H, N), box (32, 16*es, 8*es, 1) traversed with element strides (1, es, es, 1)
// (es = conv stride: the box then lands as 16 x 8 pixels).  rank-2 filter map: dims (K, Cout), box (32, NT).
int get_map(CUtensorMap* out, const float* ptr, int rank, int d0, int d1, int d2, int d3, int b1, int b2, int es) {
  MapKey key{ptr, d0, d1, d2, d3, b1, b2, es};
  {
    std::lock_guard<std::mutex> lk(g_map_mu);
    auto it = g_maps.find(key);
    if (it != g_maps.end()) { *out = it->second; return 0; }
  }
  EncodeTiledFn enc = get_encode();
  if (!enc) return DH_E_VARIANT;
  cuuint64_t dims[4] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)d2, (cuuint64_t)d3};
  cuuint64_t strides[3] = {(cuuint64_t)d0 * 4, (cuuint64_t)d0 * d1 * 4, (cuuint64_t)d0 * d1 * d2 * 4};
  cuuint32_t box[4] = {32, (cuuint32_t)b1, (cuuint32_t)b2, 1};
  cuuint32_t estr[4] = {1, (cuuint32_t)es, (cuuint32_t)(rank == 4 ? es : 1), 1};
  CUtensorMap m;
  const CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, (void*)ptr, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return DH_E_SHAPE;
  {
    std::lock_guard<std::mutex> lk(g_map_mu);
    if (g_maps.size() > 4096) g_maps.clear();
    g_maps[key] = m;
  }
  *out = m;
  return 0;
}

template <int NT>
int launch(const CUtensorMap& A0, const CUtensorMap& A1, const CUtensorMap& Bm, const TcEpilogue& e, dim3 grid, cudaStream_t s) {
  cudaError_t err = cudaFuncSetAttribute(conv_tc_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TcCfg<NT>::SMEM);
  if (err != cudaSuccess) return (int)err;
  conv_tc_kernel<NT><<<grid, 192, TcCfg<NT>::SMEM, s>>>(A0, A1, Bm, e);
  DH_CHECK_LAUNCH();
  return 0;
}

}  // namespace

bool dh_conv_tc_eligible(const ConvArgs& a) {
  const bool base = a.wt != nullptr && (a.stride == 1 || a.stride == 2) && a.up == 1 && a.KH == a.KW &&
                    (a.KH == 1 || a.KH == 3) && a.pad == a.KH / 2 && a.C0 > 0 && a.C0 % 32 == 0 && a.C1 % 32 == 0 &&
                    (a.Cout == 32 || a.Cout == 64 || a.Cout == 128 || a.Cout == 256) && a.inH >= 1 && a.inW >= 1;
  if (!base) return false;
  if (a.ps) return a.Cout == 128 && a.stride == 1 && a.res == nullptr;
  return true;
}

int dh_launch_conv_tc(const ConvArgs& a, cudaStream_t s) {
  DH_REQUIRE(a.in0 && a.wt && a.out, DH_E_NULL);
  DH_REQUIRE(a.C1 == 0 || a.in1, DH_E_NULL);
  DH_REQUIRE(dh_conv_tc_eligible(a), DH_E_SHAPE);
  DH_REQUIRE(dh_aligned16(a.in0) && dh_aligned16(a.in1) && dh_aligned16(a.wt) && dh_aligned16(a.out) &&
             dh_aligned16(a.bias) && dh_aligned16(a.res), DH_E_ALIGN);
  const int Cin = a.C0 + a.C1, K = a.KH * a.KW * Cin;
  const int NT = a.Cout >= 128 ? 128 : a.Cout;
  const int OH = (a.inH + 2 * a.pad - a.KH) / a.stride + 1, OW = (a.inW + 2 * a.pad - a.KW) / a.stride + 1;
  CUtensorMap A0, A1, Bm;
  int rc = get_map(&A0, a.in0, 4, a.C0, a.inW, a.inH, a.N, TC_TW * a.stride, TC_TH * a.stride, a.stride);
  if (rc) return rc;
  if (a.C1) {
    rc = get_map(&A1, a.in1, 4, a.C1, a.inW, a.inH, a.N, TC_TW * a.stride, TC_TH * a.stride, a.stride);
    if (rc) return rc;
  } else {
    A1 = A0;
  }
  rc = get_map(&Bm, a.wt, 2, K, a.Cout, 1, 1, NT, 1, 1);
  if (rc) return rc;
  TcEpilogue e;
  e.bias = a.bias; e.res = a.res; e.out = a.out;
  e.OH = OH; e.OW = OW; e.Cout = a.Cout; e.relu = a.relu;
  e.tilesX = dh_cdiv(OW, TC_TW);
  e.cchunks0 = a.C0 / 32; e.cchunks = Cin / 32; e.nsteps = a.KH * a.KW * e.cchunks;
  e.KW = a.KW; e.pad = a.pad; e.Cin = Cin; e.stride = a.stride; e.ps = a.ps;
  dim3 grid(e.tilesX * dh_cdiv(OH, TC_TH), a.Cout / NT, a.N);
  switch (NT) {
    case 128: return launch<128>(A0, A1, Bm, e, grid, s);
    case 64: return launch<64>(A0, A1, Bm, e, grid, s);
    default: return launch<32>(A0, A1, Bm, e, grid, s);
  }
}
