// dahitra_b200 — streaming pixel-to-token cross-attention decoder (collapsed form).
//
// Replaces _forward_transformer_decoder + TransformerDecoder (reference models/networks.py:1288-1295,
// models/help_funcs.py:66-114,170-186).  Keys/values are 4 tokens per image and do not change with depth, so
// q.K^T and attn.V.Wo collapse to per-(image, layer) tables (built by decoder_tables_kernel, tokens.cu):
//     xhat = (x - mean(x)) / sqrt(var(x) + eps)
//     s[h,j] = xhat . A[:, h*4+j] + cA[h*4+j] ;  p = softmax_j(s[h,:])
//     x += sum_{h,j} p[h,j] Bv[h*4+j,:] + b_out
//     x += W2 gelu(W1f xhat' + b1f) + b2         (second LayerNorm folded into W1f/b1f)
// One thread owns one pixel (its 32 channels stay in registers through all layers); the per-layer tables
// are staged in shared memory and read as warp-wide broadcasts.  Activations are read once and written
// once: 128 B in (+128 B positional embedding) and 128 B out per pixel per call.
#include "common.cuh"

namespace {
constexpr int PD_T = 256;   // pixels (= threads) per CTA

template <int HEADS>
__global__ void __launch_bounds__(PD_T, 2)
pixel_decoder_kernel(const float* __restrict__ x, const float* __restrict__ pos, const float* __restrict__ tables,
                     const float* __restrict__ dec, int npix, int w, int depth, const float* __restrict__ skip,
                     int skip_up, float* __restrict__ out) {
  constexpr int H4 = HEADS * 4;
  constexpr int TAB = DH_TAB_FLOATS(HEADS);
  constexpr int MLPF = 1024 + 32 + 1024 + 32;
  constexpr size_t LSTRIDE = DH_DEC_LAYER_FLOATS(HEADS);
  __shared__ __align__(16) float tab_s[TAB];
  __shared__ __align__(16) float mlp_s[MLPF];
  const int tid = threadIdx.x, img = blockIdx.y;
  const int p = blockIdx.x * PD_T + tid;
  const bool valid = p < npix;

  float xr[32];
  if (valid) {
    const float* xp = x + ((size_t)img * npix + p) * 32;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 v = ldg4(xp + q * 4);
      xr[q * 4] = v.x; xr[q * 4 + 1] = v.y; xr[q * 4 + 2] = v.z; xr[q * 4 + 3] = v.w;
    }
    if (pos) {
      const float* pp = pos + (size_t)p * 32;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 v = ldg4(pp + q * 4);
        xr[q * 4] += v.x; xr[q * 4 + 1] += v.y; xr[q * 4 + 2] += v.z; xr[q * 4 + 3] += v.w;
      }
    }
  } else {
#pragma unroll
    for (int c = 0; c < 32; ++c) xr[c] = 0.f;
  }

  const float* A = tab_s; const float* cA = tab_s + 32 * H4; const float* Bv = cA + H4; const float* bout = Bv + H4 * 32;
  const float* W1f = mlp_s; const float* b1f = W1f + 1024; const float* W2t = b1f + 32; const float* b2 = W2t + 1024;

  for (int layer = 0; layer < depth; ++layer) {
    __syncthreads();   // previous layer's readers are done
    {
      const float4* tg = reinterpret_cast<const float4*>(tables + ((size_t)img * depth + layer) * TAB);
      for (int i = tid; i < TAB / 4; i += PD_T) reinterpret_cast<float4*>(tab_s)[i] = __ldg(tg + i);
      const float4* mg = reinterpret_cast<const float4*>(dec + (size_t)layer * LSTRIDE + (LSTRIDE - MLPF));
      for (int i = tid; i < MLPF / 4; i += PD_T) reinterpret_cast<float4*>(mlp_s)[i] = __ldg(mg + i);
    }
    __syncthreads();
    // ---- cross attention -------------------------------------------------------------------------
    float mu = 0.f;
#pragma unroll
    for (int c = 0; c < 32; ++c) mu += xr[c];
    mu *= (1.f / 32.f);
    float var = 0.f;
#pragma unroll
    for (int c = 0; c < 32; ++c) { const float d = xr[c] - mu; var = fmaf(d, d, var); }
    float rstd = 1.0f / sqrtf(var * (1.f / 32.f) + 1e-5f);
    float s[H4];
#pragma unroll
    for (int q = 0; q < H4 / 4; ++q) {
      const float4 v = *reinterpret_cast<const float4*>(cA + q * 4);
      s[q * 4] = v.x; s[q * 4 + 1] = v.y; s[q * 4 + 2] = v.z; s[q * 4 + 3] = v.w;
    }
#pragma unroll
    for (int c = 0; c < 32; ++c) {
      const float xh = (xr[c] - mu) * rstd;
#pragma unroll
      for (int q = 0; q < H4 / 4; ++q) {
        const float4 v = *reinterpret_cast<const float4*>(A + c * H4 + q * 4);
        s[q * 4] = fmaf(xh, v.x, s[q * 4]);
        s[q * 4 + 1] = fmaf(xh, v.y, s[q * 4 + 1]);
        s[q * 4 + 2] = fmaf(xh, v.z, s[q * 4 + 2]);
        s[q * 4 + 3] = fmaf(xh, v.w, s[q * 4 + 3]);
      }
    }
#pragma unroll
    for (int h = 0; h < HEADS; ++h) {
      const float mx = fmaxf(fmaxf(s[h * 4], s[h * 4 + 1]), fmaxf(s[h * 4 + 2], s[h * 4 + 3]));
      const float e0 = expf(s[h * 4] - mx), e1 = expf(s[h * 4 + 1] - mx), e2 = expf(s[h * 4 + 2] - mx), e3 = expf(s[h * 4 + 3] - mx);
      const float inv = 1.0f / (e0 + e1 + e2 + e3);
      s[h * 4] = e0 * inv; s[h * 4 + 1] = e1 * inv; s[h * 4 + 2] = e2 * inv; s[h * 4 + 3] = e3 * inv;
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 v = *reinterpret_cast<const float4*>(bout + q * 4);
      xr[q * 4] += v.x; xr[q * 4 + 1] += v.y; xr[q * 4 + 2] += v.z; xr[q * 4 + 3] += v.w;
    }
#pragma unroll
    for (int k = 0; k < H4; ++k) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 v = *reinterpret_cast<const float4*>(Bv + k * 32 + q * 4);
        xr[q * 4] = fmaf(s[k], v.x, xr[q * 4]);
        xr[q * 4 + 1] = fmaf(s[k], v.y, xr[q * 4 + 1]);
        xr[q * 4 + 2] = fmaf(s[k], v.z, xr[q * 4 + 2]);
        xr[q * 4 + 3] = fmaf(s[k], v.w, xr[q * 4 + 3]);
      }
    }
    // ---- MLP -------------------------------------------------------------------------------------
    mu = 0.f;
#pragma unroll
    for (int c = 0; c < 32; ++c) mu += xr[c];
    mu *= (1.f / 32.f);
    var = 0.f;
#pragma unroll
    for (int c = 0; c < 32; ++c) { const float d = xr[c] - mu; var = fmaf(d, d, var); }
    rstd = 1.0f / sqrtf(var * (1.f / 32.f) + 1e-5f);
    float hid[32];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 v = *reinterpret_cast<const float4*>(b1f + q * 4);
      hid[q * 4] = v.x; hid[q * 4 + 1] = v.y; hid[q * 4 + 2] = v.z; hid[q * 4 + 3] = v.w;
    }
#pragma unroll
    for (int c = 0; c < 32; ++c) {
      const float xh = (xr[c] - mu) * rstd;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 v = *reinterpret_cast<const float4*>(W1f + c * 32 + q * 4);
        hid[q * 4] = fmaf(xh, v.x, hid[q * 4]);
        hid[q * 4 + 1] = fmaf(xh, v.y, hid[q * 4 + 1]);
        hid[q * 4 + 2] = fmaf(xh, v.z, hid[q * 4 + 2]);
        hid[q * 4 + 3] = fmaf(xh, v.w, hid[q * 4 + 3]);
      }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 v = *reinterpret_cast<const float4*>(b2 + q * 4);
      xr[q * 4] += v.x; xr[q * 4 + 1] += v.y; xr[q * 4 + 2] += v.z; xr[q * 4 + 3] += v.w;
    }
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const float g = gelu_erf(hid[k]);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 v = *reinterpret_cast<const float4*>(W2t + k * 32 + q * 4);
        xr[q * 4] = fmaf(g, v.x, xr[q * 4]);
        xr[q * 4 + 1] = fmaf(g, v.y, xr[q * 4 + 1]);
        xr[q * 4 + 2] = fmaf(g, v.z, xr[q * 4 + 2]);
        xr[q * 4 + 3] = fmaf(g, v.w, xr[q * 4 + 3]);
      }
    }
  }

  if (valid) {
    if (skip) {   // + coarser-level result: nearest-x2-upsampled (networks.py:1329,1333) or same size (:1340)
      const int py = p / w, px = p - py * w;
      const float* sp = (skip_up == 2)
          ? skip + (((size_t)img * (npix / w / 2) + (py >> 1)) * (w >> 1) + (px >> 1)) * 32
          : skip + ((size_t)img * npix + p) * 32;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 v = ldg4(sp + q * 4);
        xr[q * 4] += v.x; xr[q * 4 + 1] += v.y; xr[q * 4 + 2] += v.z; xr[q * 4 + 3] += v.w;
      }
    }
    float* op = out + ((size_t)img * npix + p) * 32;
#pragma unroll
    for (int q = 0; q < 8; ++q) st4(op + q * 4, make_float4(xr[q * 4], xr[q * 4 + 1], xr[q * 4 + 2], xr[q * 4 + 3]));
  }
}
}  // namespace

int dh_launch_pixel_decoder(const float* x, const float* pos, const float* tables, const float* dec,
                            int nimg, int h, int w, int heads, int depth, const float* skip, int skip_up, float* out,
                            cudaStream_t s) {
  DH_REQUIRE(x && tables && dec && out, DH_E_NULL);
  DH_REQUIRE(nimg > 0 && h > 0 && w > 0 && depth >= 1 && (heads == 4 || heads == 8), DH_E_SHAPE);
  DH_REQUIRE(!skip || skip_up == 1 || (skip_up == 2 && h % 2 == 0 && w % 2 == 0), DH_E_SHAPE);
  DH_REQUIRE(dh_aligned16(x) && dh_aligned16(pos) && dh_aligned16(tables) && dh_aligned16(dec) && dh_aligned16(skip) &&
             dh_aligned16(out), DH_E_ALIGN);
  const int npix = h * w;
  dim3 grid(dh_cdiv(npix, PD_T), nimg);
  if (heads == 4)
    pixel_decoder_kernel<4><<<grid, PD_T, 0, s>>>(x, pos, tables, dec, npix, w, depth, skip, skip_up, out);
  else
    pixel_decoder_kernel<8><<<grid, PD_T, 0, s>>>(x, pos, tables, dec, npix, w, depth, skip, skip_up, out);
  DH_CHECK_LAUNCH();
  return 0;
}
