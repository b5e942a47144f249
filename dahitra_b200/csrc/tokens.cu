// dahitra_b200 — token path: squeeze(1x1 conv+ReLU) fused with the tokenizer's spatial softmax partials,
// the 1-layer token encoder, and the per-(pair,call,layer) attention tables of the collapsed pixel decoder.
#include "common.cuh"

// =====================================================================================================
// squeeze + tokenizer partials.
//   xs[p][:]  = relu(Wsq^T feat[p][:])                      (reference models/networks.py:1177-1184)
//   a[p][l]   = sum_c Wtok[c][l] xs[p][c]                   (conv_token, :1187-1189)
//   per CTA (128 pixels): m_l = max_p a, s_l = sum_p exp(a-m_l), t_l[c] = sum_p exp(a-m_l) xs[p][c]
// The spatial softmax over all N pixels (:1273-1280) is finished by the token-encoder kernel, which merges
// the per-CTA (m, s, t) triples with the usual log-sum-exp rescaling.
// =====================================================================================================
namespace {
constexpr int SQ_T = 128;   // threads per CTA
constexpr int SQ_P = 2;     // pixels per thread: every filter read from shared memory feeds two pixels
constexpr int SQ_PX = SQ_T * SQ_P;

__global__ void __launch_bounds__(SQ_T)
squeeze_tokens_kernel(const float* __restrict__ feat, int npix, int Cin, int nchunk, const float* __restrict__ wsq,
                      const float* __restrict__ wtok, float* __restrict__ xs, float* __restrict__ partials) {
  extern __shared__ __align__(16) float sm[];
  float* w_s = sm;                         // [Cin][32]
  float* wt_s = w_s + Cin * 32;            // [32][4]
  float* tile = wt_s + 128;                // [256][33]
  float* e_s = tile + SQ_PX * 33;          // [256][4]
  float* red = e_s + SQ_PX * 4;            // [4 warps][4]
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int img = blockIdx.y, chunk = blockIdx.x;
  for (int i = tid; i < Cin * 8; i += SQ_T) reinterpret_cast<float4*>(w_s)[i] = __ldg(reinterpret_cast<const float4*>(wsq) + i);
  if (tid < 32) reinterpret_cast<float4*>(wt_s)[tid] = __ldg(reinterpret_cast<const float4*>(wtok) + tid);
  __syncthreads();

  // thread t owns pixels chunk*256 + t and + t + 128 (both coalesce the same way across the warp)
  int p[SQ_P];
  bool valid[SQ_P];
  float acc[SQ_P][32];
#pragma unroll
  for (int u = 0; u < SQ_P; ++u) {
    p[u] = chunk * SQ_PX + u * SQ_T + tid;
    valid[u] = p[u] < npix;
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[u][j] = 0.f;
  }
  const float* fp0 = feat + ((size_t)img * npix + (valid[0] ? p[0] : 0)) * Cin;
  const float* fp1 = feat + ((size_t)img * npix + (valid[1] ? p[1] : 0)) * Cin;
  for (int c = 0; c < Cin; c += 4) {
    const float4 f0 = ldg4(fp0 + c), f1 = ldg4(fp1 + c);
    const float fv[SQ_P][4] = {{f0.x, f0.y, f0.z, f0.w}, {f1.x, f1.y, f1.z, f1.w}};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float* wr = w_s + (c + e) * 32;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 ww = *reinterpret_cast<const float4*>(wr + q * 4);
#pragma unroll
        for (int u = 0; u < SQ_P; ++u) {
          acc[u][q * 4 + 0] = fmaf(fv[u][e], ww.x, acc[u][q * 4 + 0]);
          acc[u][q * 4 + 1] = fmaf(fv[u][e], ww.y, acc[u][q * 4 + 1]);
          acc[u][q * 4 + 2] = fmaf(fv[u][e], ww.z, acc[u][q * 4 + 2]);
          acc[u][q * 4 + 3] = fmaf(fv[u][e], ww.w, acc[u][q * 4 + 3]);
        }
      }
    }
  }
  float lg[SQ_P][4];
#pragma unroll
  for (int u = 0; u < SQ_P; ++u) {
    lg[u][0] = lg[u][1] = lg[u][2] = lg[u][3] = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      acc[u][j] = fmaxf(acc[u][j], 0.f);
      const float4 wt = *reinterpret_cast<const float4*>(wt_s + j * 4);
      lg[u][0] = fmaf(acc[u][j], wt.x, lg[u][0]);
      lg[u][1] = fmaf(acc[u][j], wt.y, lg[u][1]);
      lg[u][2] = fmaf(acc[u][j], wt.z, lg[u][2]);
      lg[u][3] = fmaf(acc[u][j], wt.w, lg[u][3]);
    }
    if (valid[u]) {
      float* op = xs + ((size_t)img * npix + p[u]) * 32;
#pragma unroll
      for (int q = 0; q < 8; ++q) st4(op + q * 4, make_float4(acc[u][q * 4], acc[u][q * 4 + 1], acc[u][q * 4 + 2], acc[u][q * 4 + 3]));
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) tile[(u * SQ_T + tid) * 33 + j] = acc[u][j];
  }
  // block max of each token's logits
#pragma unroll
  for (int l = 0; l < 4; ++l) {
    const float m = warp_max(fmaxf(valid[0] ? lg[0][l] : -INFINITY, valid[1] ? lg[1][l] : -INFINITY));
    if (lane == 0) red[wid * 4 + l] = m;
  }
  __syncthreads();
#pragma unroll
  for (int l = 0; l < 4; ++l) {
    const float mx = fmaxf(fmaxf(red[l], red[4 + l]), fmaxf(red[8 + l], red[12 + l]));
#pragma unroll
    for (int u = 0; u < SQ_P; ++u) e_s[(u * SQ_T + tid) * 4 + l] = valid[u] ? expf(lg[u][l] - mx) : 0.f;
  }
  __syncthreads();
  // thread (l, c): weighted sum over the CTA's pixels
  const int l = tid >> 5, c = tid & 31;
  float t = 0.f, ssum = 0.f;
#pragma unroll 8
  for (int q = 0; q < SQ_PX; ++q) {
    const float e = e_s[q * 4 + l];
    ssum += e;
    t = fmaf(e, tile[q * 33 + c], t);
  }
  float* pp = partials + (((size_t)img * nchunk + chunk) * 4 + l) * 34;
  if (c == 0) {
    pp[0] = fmaxf(fmaxf(red[l], red[4 + l]), fmaxf(red[8 + l], red[12 + l]));
    pp[1] = ssum;
  }
  pp[2 + c] = t;
}
}  // namespace

int dh_launch_squeeze_tokens(const float* feat, int N, int npix, int Cin, const float* wsq, const float* wtok,
                             float* xs, float* partials, cudaStream_t s) {
  DH_REQUIRE(feat && wsq && wtok && xs && partials, DH_E_NULL);
  DH_REQUIRE(N > 0 && npix > 0 && Cin % 4 == 0 && Cin >= 4 && Cin <= 256, DH_E_SHAPE);
  DH_REQUIRE(dh_aligned16(feat) && dh_aligned16(wsq) && dh_aligned16(wtok) && dh_aligned16(xs), DH_E_ALIGN);
  const int nchunk = dh_cdiv(npix, SQ_PX);
  const int smem = (Cin * 32 + 128 + SQ_PX * 33 + SQ_PX * 4 + 16) * (int)sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(squeeze_tokens_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return (int)e;
  dim3 grid(nchunk, N);
  squeeze_tokens_kernel<<<grid, SQ_T, smem, s>>>(feat, npix, Cin, nchunk, wsq, wtok, xs, partials);
  DH_CHECK_LAUNCH();
  return 0;
}

// =====================================================================================================
// Token encoder: one CTA (8 warps = 8 tokens) per image pair.
//   tokens = softmax-merge(partials) ; tokens += pos ; 1 pre-norm self-attention layer + MLP
//   (reference models/networks.py:1282-1286 and the Transformer at :434-512).
// The q.k and v.out products are evaluated through the host-collapsed per-head 32x32 matrices
//   Mqk[h] = dim^-0.5 Wq[h]^T Wk[h],  MvoT[h][c'][c] = sum_d Wo[c][hd] Wv[hd][c']      (include/dahitra_b200.h)
// Output: mem[pair][0] = token1', mem[pair][1] = token2', mem[pair][2] = |token2' - token1'| (:1304,1315).
// =====================================================================================================
namespace {

__device__ __forceinline__ float ln_lane(float v, float g, float b) {   // LayerNorm(32) across a warp
  const float mu = warp_sum(v) * (1.f / 32.f);
  const float d = v - mu;
  const float var = warp_sum(d * d) * (1.f / 32.f);
  return d * (1.0f / sqrtf(var + 1e-5f)) * g + b;
}

__global__ void __launch_bounds__(256)
token_encoder_kernel(const float* __restrict__ partials, int B, int nchunk, const float* __restrict__ enc, int heads,
                     int add_pos, float* __restrict__ mem) {
  __shared__ float X[8][32], XN[8][32], Y[8][32], Hh[8][32];
  extern __shared__ __align__(16) float te_w[];                 // the whole encoder pack (DH_ENC_FLOATS(heads) floats)
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;   // warp w <-> token w; lane <-> channel
  const int pair = blockIdx.x;
  // The pack is cold in L2 by the time this kernel runs (GBs of activations have streamed through since the last
  // forward), and the attention loop below would otherwise walk it in 32 dependent batches of loads — 60 of the 70 us
  // this kernel took.  Fetch it once, asynchronously, while the partial softmax sums are merged.
  {
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(te_w);
    const int nvec = DH_ENC_FLOATS(heads) / 4;
    for (int i = tid; i < nvec; i += 256)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (uint32_t)i * 16u), "l"(enc + (size_t)i * 4) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  dh_pdl_wait();                       // the partials come from the previous launch; the weight fetch above does not
  dh_pdl_launch_dependents();
  const float* pos = te_w;
  const float* ln1g = te_w + 256; const float* ln1b = ln1g + 32;
  const float* Mqk = ln1b + 32;
  const float* MvoT = Mqk + (size_t)heads * 1024;
  const float* bo = MvoT + (size_t)heads * 1024;
  const float* ln2g = bo + 32; const float* ln2b = ln2g + 32;
  const float* W1t = ln2b + 32; const float* b1 = W1t + 1024;
  const float* W2t = b1 + 32; const float* b2 = W2t + 1024;

  // 1. finish the spatial softmax for token (img = w/4, l = w%4), channel = lane
  {
    const int img = (w < 4) ? pair : (B + pair), l = w & 3;
    const float* pp = partials + ((size_t)img * nchunk * 4 + l) * 34;
    float M = -INFINITY;
    for (int k = lane; k < nchunk; k += 32) M = fmaxf(M, __ldg(pp + (size_t)k * 4 * 34));
    M = warp_max(M);
    // merge in groups of 32 chunks: lane j evaluates the rescaling factor of chunk k0 + j once (one exp per chunk instead of
    // one per chunk AND lane), then every lane (= channel) accumulates its t over the group with the factors broadcast by
    // shuffle; 16 loads per lane are in flight at a time (the 1024^2 inputs have 512 chunks)
    float S = 0.f, T0 = 0.f, T1 = 0.f, T2 = 0.f, T3 = 0.f;
    for (int k0 = 0; k0 < nchunk; k0 += 32) {
      const int kk = k0 + lane;
      float sc = 0.f;
      if (kk < nchunk) {
        const float* q = pp + (size_t)kk * 4 * 34;
        sc = expf(__ldg(q) - M);
        S = fmaf(__ldg(q + 1), sc, S);
      }
      const int n = min(32, nchunk - k0);
      const float* tq = pp + (size_t)k0 * 4 * 34 + 2 + lane;
      int j = 0;
      for (; j + 16 <= n; j += 16) {                            // 16 loads in flight per lane (each an L2 round trip)
        float a[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) a[u] = __ldg(tq + (size_t)(j + u) * 136);
#pragma unroll
        for (int u = 0; u < 16; u += 4) {
          T0 = fmaf(a[u + 0], __shfl_sync(0xffffffffu, sc, j + u + 0), T0);
          T1 = fmaf(a[u + 1], __shfl_sync(0xffffffffu, sc, j + u + 1), T1);
          T2 = fmaf(a[u + 2], __shfl_sync(0xffffffffu, sc, j + u + 2), T2);
          T3 = fmaf(a[u + 3], __shfl_sync(0xffffffffu, sc, j + u + 3), T3);
        }
      }
      for (; j < n; ++j) T0 = fmaf(__ldg(tq + (size_t)j * 136), __shfl_sync(0xffffffffu, sc, j), T0);
    }
    S = warp_sum(S);
    const float T = (T0 + T1) + (T2 + T3);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                                            // the encoder pack is in shared memory
    float v = T / S;
    if (add_pos) v += pos[w * 32 + lane];
    X[w][lane] = v;
    XN[w][lane] = ln_lane(v, ln1g[lane], ln1b[lane]);
  }
  __syncthreads();
  // 2. attention, head by head
  float att_out = 0.f;
  for (int h = 0; h < heads; ++h) {
    const float* mq = Mqk + (size_t)h * 1024;
    const float* mv = MvoT + (size_t)h * 1024;
    float u = 0.f, y = 0.f;     // u[c'] = sum_c xn_w[c] Mqk[c][c'] ;  y[c] = sum_c' MvoT[c'][c] xn_w[c']
#pragma unroll 8
    for (int c = 0; c < 32; ++c) {
      const float xv = XN[w][c];
      u = fmaf(xv, mq[c * 32 + lane], u);
      y = fmaf(xv, mv[c * 32 + lane], y);
    }
    __syncthreads();            // previous head's readers of Y are done
    Y[w][lane] = y;
    float d[8], mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      d[j] = warp_sum(u * XN[j][lane]);
      mx = fmaxf(mx, d[j]);
    }
    float den = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) { d[j] = expf(d[j] - mx); den += d[j]; }
    const float inv = 1.f / den;
    __syncthreads();            // Y complete
#pragma unroll
    for (int j = 0; j < 8; ++j) att_out = fmaf(d[j] * inv, Y[j][lane], att_out);
  }
  float x = X[w][lane] + att_out + bo[lane];
  // 3. MLP
  const float xn2 = ln_lane(x, ln2g[lane], ln2b[lane]);
  __syncthreads();
  XN[w][lane] = xn2;
  __syncwarp();
  float hsum = b1[lane];
#pragma unroll 8
  for (int c = 0; c < 32; ++c) hsum = fmaf(XN[w][c], W1t[c * 32 + lane], hsum);
  Hh[w][lane] = gelu_erf(hsum);
  __syncwarp();
  float o = b2[lane];
#pragma unroll 8
  for (int k = 0; k < 32; ++k) o = fmaf(Hh[w][k], W2t[k * 32 + lane], o);
  x += o;
  __syncthreads();
  X[w][lane] = x;
  __syncthreads();
  float* mp = mem + (size_t)pair * 3 * 128;
  mp[w * 32 + lane] = x;                                    // calls 0 (tokens 0-3) and 1 (tokens 4-7)
  if (w < 4) mp[256 + w * 32 + lane] = fabsf(X[w + 4][lane] - X[w][lane]);
}
}  // namespace

int dh_launch_token_encoder(const float* partials, int B, int nchunk, const float* enc, int heads, int add_pos,
                            float* mem, cudaStream_t s) {
  DH_REQUIRE(partials && enc && mem, DH_E_NULL);
  DH_REQUIRE(B > 0 && nchunk > 0 && heads >= 1 && heads <= 16, DH_E_SHAPE);
  const int smem = DH_ENC_FLOATS(heads) * (int)sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(token_encoder_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return (int)e;
  return dh_launch(token_encoder_kernel, dim3(B), dim3(256), (size_t)smem, s, partials, B, nchunk, enc, heads, add_pos, mem);
}

// =====================================================================================================
// Decoder attention tables.  For memory tokens m (4x32) of one (pair, call) and decoder layer L:
//   mn_j = LN_L(m_j)                           (PreNorm2 shares the layer's LayerNorm, help_funcs.py:43-49)
//   A [c][h*4+j] = g_c * sum_c' Mqk_L[h][c][c'] mn_j[c']       (LN gamma of the query side folded in)
//   cA[h*4+j]    = sum_c b_c * (A/g)[c][h*4+j]                  (LN beta of the query side)
//   Bv[h*4+j][c] = sum_c' Mov_L[h][c][c'] mn_j[c']
// so that per pixel  s = xhat . A + cA ; p = softmax_j(s) ; x += sum p Bv + b_out   (xhat = (x-mu)/sigma).
// Table layout: A[32][4H] | cA[4H] | Bv[4H][32] | b_out[32].
// grid = (depth, ncalls*B); 128 threads (warp j <-> memory token j, lane <-> channel).
// =====================================================================================================
namespace {
__global__ void __launch_bounds__(128)
decoder_tables_kernel(const float* __restrict__ mem, int B, int first_call, const float* __restrict__ dec, int heads,
                      float* __restrict__ tables, int depth) {
  __shared__ float MN[4][32];
  const int lane = threadIdx.x & 31, j = threadIdx.x >> 5;
  const int layer = blockIdx.x;
  const int ci = blockIdx.y / B, pair = blockIdx.y % B;       // call index relative to first_call
  const int call = first_call + ci;
  const size_t lstride = DH_DEC_LAYER_FLOATS(heads);
  const float* L = dec + (size_t)layer * lstride;
  const float* ln1g = L; const float* ln1b = L + 32;
  const float* MqkT = L + 64;
  const float* MovT = MqkT + (size_t)heads * 1024;
  const float* bo = MovT + (size_t)heads * 1024;
  const int H4 = heads * 4;
  float* T = tables + ((size_t)blockIdx.y * depth + layer) * DH_TAB_FLOATS(heads);
  float* A = T; float* cA = T + 32 * H4; float* Bv = cA + H4; float* bout = Bv + H4 * 32;

  const float g = __ldg(ln1g + lane), b = __ldg(ln1b + lane);
  const float m = __ldg(mem + ((size_t)pair * 3 + call) * 128 + j * 32 + lane);
  MN[j][lane] = ln_lane(m, g, b);
  __syncwarp();
  for (int h = 0; h < heads; ++h) {
    const float* mq = MqkT + (size_t)h * 1024;
    const float* mv = MovT + (size_t)h * 1024;
    float a = 0.f, v = 0.f;
#pragma unroll 8
    for (int c2 = 0; c2 < 32; ++c2) {
      const float mn = MN[j][c2];
      a = fmaf(__ldg(mq + c2 * 32 + lane), mn, a);     // MqkT[h][c'][c]
      v = fmaf(__ldg(mv + c2 * 32 + lane), mn, v);     // MovT[h][c'][c]
    }
    const int col = h * 4 + j;
    A[lane * H4 + col] = g * a;
    const float ca = warp_sum(b * a);
    if (lane == 0) cA[col] = ca;
    Bv[col * 32 + lane] = v;
  }
  if (j == 0) bout[lane] = __ldg(bo + lane);
}
}  // namespace

int dh_launch_decoder_tables(const float* mem, int B, int first_call, int ncalls, const float* dec, int heads, int depth,
                             float* tables, cudaStream_t s) {
  DH_REQUIRE(mem && dec && tables, DH_E_NULL);
  DH_REQUIRE(B > 0 && first_call >= 0 && ncalls >= 1 && first_call + ncalls <= 3 && (heads == 4 || heads == 8) &&
             depth >= 1, DH_E_SHAPE);
  dim3 grid(depth, ncalls * B);
  decoder_tables_kernel<<<grid, 128, 0, s>>>(mem, B, first_call, dec, heads, tables, depth);
  DH_CHECK_LAUNCH();
  return 0;
}
