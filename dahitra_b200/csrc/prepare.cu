// dahitra_b200 — host-side weight preparation behind the C ABI (SURVEY.md §8b `dahitra_prepare_weights`): reference-layout
// state_dict tensors (host pointers) -> the prepared slot table `dahitra_forward` consumes.  No PyTorch, no CUDA calls: a
// C / C++ / ctypes consumer can go from a checkpoint's tensors to a forward without dahitra_b200/engine.py.
//
// What it does (the same algebra as engine.prepare_weights, in fp64, rounded to fp32 once at the end):
//   * eval-mode BatchNorm folded into the preceding convolution (reference models/resnet.py:57-73, help_funcs.py:7-15):
//       W' = W * g / sqrt(var + 1e-5),  b' = beta - mean * g / sqrt(var + 1e-5)
//   * filters re-laid-out OIHW -> [KH*KW*Cin][Cout] and, for the tensor-core kernels, K-major [Cout][KH*KW*Cin] in five
//     operand planes (TF32 hi / lo, bf16 pair, f16 + bf16 remainder, scaled f16 remainder — include/dahitra_b200.h)
//   * nearest-x2 upsample + 3x3 conv (conv_layer4/3/2, reference networks.py:1335-1351) as one 3x3 conv 32 -> 4x32 whose
//     taps are summed per output-pixel phase
//   * token encoder / pixel decoder projections collapsed to per-head 32x32 matrices (reference networks.py:434-512,
//     help_funcs.py:66-114): Mqk = dim^-0.5 Wq^T Wk, Mov = Wo Wv; second LayerNorm of each decoder layer folded into W1 / b1;
//     pre-swizzled TF32 hi / lo tiles and cumulative biases for the tensor-core decoder
//   * the stem's filter images for the tcgen05 stem (K ordered (ci, r, s8))
//   * positional embeddings NCHW -> [h*w][32]
#include "common.cuh"
#include <cmath>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace {

typedef std::vector<double> V;

struct Src {
  std::map<std::string, const dh_tensor*> m;
  bool ok = true;
  std::string missing;
  const dh_tensor* find(const std::string& k) const { auto it = m.find(k); return it == m.end() ? nullptr : it->second; }
  bool has(const std::string& k) const { return m.count(k) != 0; }
  // tensor as doubles; numel checked against `expect` (0 = any)
  V get(const std::string& k, size_t expect) {
    const dh_tensor* t = find(k);
    if (!t || !t->data) { if (ok) missing = k; ok = false; return V(expect, 0.0); }
    size_t n = 1;
    for (int i = 0; i < t->ndim; ++i) n *= (size_t)t->shape[i];
    if (expect && n != expect) { if (ok) missing = k + " (shape)"; ok = false; return V(expect, 0.0); }
    V v(n);
    if (t->dtype == DH_DTYPE_F32) { const float* p = (const float*)t->data; for (size_t i = 0; i < n; ++i) v[i] = p[i]; }
    else if (t->dtype == DH_DTYPE_F64) { const double* p = (const double*)t->data; for (size_t i = 0; i < n; ++i) v[i] = p[i]; }
    else { if (ok) missing = k + " (dtype)"; ok = false; }
    return v;
  }
};

// ---- scalar format helpers (round to nearest even, like torch's .to(float16 / bfloat16))
inline uint32_t f32_bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
inline float bits_f32(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
inline float tf32_round_f(float x) { return bits_f32((f32_bits(x) + 0x1000u) & ~0x1FFFu); }   // ties away from zero (cvt.rna.tf32.f32)
inline uint16_t bf16_bits(float f) {
  uint32_t u = f32_bits(f);
  if ((u & 0x7FFFFFFFu) > 0x7F800000u) return (uint16_t)((u >> 16) | 0x40);   // NaN
  u += 0x7FFFu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
inline uint16_t f16_bits(float f) {                       // IEEE binary16, RNE, overflow -> inf (callers clamp first)
  const uint32_t u = f32_bits(f), sign = (u >> 16) & 0x8000u;
  const uint32_t a = u & 0x7FFFFFFFu;
  if (a >= 0x7F800000u) return (uint16_t)(sign | (a > 0x7F800000u ? 0x7E00u : 0x7C00u));
  if (a >= 0x477FF000u) return (uint16_t)(sign | 0x7C00u);                    // rounds to >= 65520 -> inf
  if (a < 0x33000001u) return (uint16_t)sign;                                   // < 2^-25 (and the tie at 2^-25) -> 0
  int e = (int)(a >> 23) - 127;
  uint32_t m = (a & 0x7FFFFFu) | 0x800000u;                                     // 24-bit significand
  int shift;                                                                    // bits to drop
  uint32_t hexp;
  if (e >= -14) { shift = 13; hexp = (uint32_t)(e + 15); }
  else { shift = 13 + (-14 - e); hexp = 0; }
  uint32_t q = m >> shift;
  const uint32_t rem = m & ((1u << shift) - 1u), half = 1u << (shift - 1);
  if (rem > half || (rem == half && (q & 1u))) ++q;
  uint32_t h;
  if (hexp == 0) h = q;                                                         // subnormal (q may carry into the exponent: correct)
  else h = ((hexp - 1) << 10) + q;                                              // q includes the hidden bit: (hexp << 10) + (q - 0x400)
  return (uint16_t)(sign | h);
}
inline float f16_val(uint16_t h) {
  const uint32_t sign = (uint32_t)(h & 0x8000u) << 16, e = (h >> 10) & 0x1Fu, m = h & 0x3FFu;
  if (e == 0) { const float v = std::ldexp((float)m, -24); return sign ? -v : v; }
  if (e == 31) return bits_f32(sign | 0x7F800000u | (m << 13));
  return bits_f32(sign | ((e + 112u) << 23) | (m << 13));
}
inline double clamp16(double x) { return x > 65504.0 ? 65504.0 : (x < -65504.0 ? -65504.0 : x); }

// OIHW (folded) -> [KH*KW*Cin][Cout]
V khwc(const V& w, int O, int I, int K) {
  V r((size_t)K * K * I * O);
  for (int o = 0; o < O; ++o) for (int i = 0; i < I; ++i) for (int a = 0; a < K; ++a) for (int b = 0; b < K; ++b)
    r[((size_t)(a * K + b) * I + i) * O + o] = w[(((size_t)o * I + i) * K + a) * K + b];
  return r;
}
// [K][Cout] -> K-major [Cout][K]
V transpose(const V& m, int rows, int cols) {
  V r(m.size());
  for (int i = 0; i < rows; ++i) for (int j = 0; j < cols; ++j) r[(size_t)j * rows + i] = m[(size_t)i * cols + j];
  return r;
}
// K-major filter [Cout][K] (fp64) -> float32 [5][Cout][K] (engine.kmajor_split)
void kmajor_split(float* dst, const V& wt) {
  if (!dst) return;
  const size_t n = wt.size();
  float* p0 = dst; float* p1 = dst + n;
  uint16_t* p2 = reinterpret_cast<uint16_t*>(dst + 2 * n);
  uint16_t* p3 = reinterpret_cast<uint16_t*>(dst + 3 * n);
  uint16_t* p4 = reinterpret_cast<uint16_t*>(dst + 4 * n);
  for (size_t i = 0; i < n; ++i) {
    const double w = wt[i];
    const float hi = tf32_round_f((float)w);
    p0[i] = hi;
    p1[i] = tf32_round_f((float)(w - (double)hi));
    p2[i] = bf16_bits((float)w);
    p2[n + i] = bf16_bits((float)(w - (double)hi));
    const uint16_t h16 = f16_bits((float)clamp16(w));
    const double r = w - (double)f16_val(h16);
    p3[i] = h16;
    p3[n + i] = bf16_bits((float)r);
    p4[i] = f16_bits((float)clamp16(r * 2048.0));
    p4[n + i] = 0;
  }
}
inline size_t sw128(int r, int k) { return (size_t)r * 32 + ((((k >> 2) ^ (r & 7)) << 2) | (k & 3)); }
inline size_t sw128_16(int n, int k) { return (size_t)n * 64 + ((((k >> 3) ^ (n & 7)) << 3) | (k & 7)); }

struct Conv { V w, b; int O, I, K; };
// conv weight (+ optional bias) with an optional BatchNorm folded in
Conv fold(Src& s, const std::string& conv, const std::string& bn, int O, int I, int K) {
  Conv c; c.O = O; c.I = I; c.K = K;
  c.w = s.get(conv + ".weight", (size_t)O * I * K * K);
  if (bn.empty()) {
    if (s.has(conv + ".bias")) c.b = s.get(conv + ".bias", O);
    return c;
  }
  const V g = s.get(bn + ".weight", O), be = s.get(bn + ".bias", O), mu = s.get(bn + ".running_mean", O), var = s.get(bn + ".running_var", O);
  c.b.resize(O);
  const size_t per = (size_t)I * K * K;
  for (int o = 0; o < O; ++o) {
    const double k = g[o] / std::sqrt(var[o] + 1e-5);
    for (size_t i = 0; i < per; ++i) c.w[o * per + i] *= k;
    c.b[o] = be[o] - mu[o] * k;
  }
  return c;
}
// nn.Upsample(x2, nearest) + 3x3 pad-1 conv == four 2x2 convs on the low-res map (engine.upsample_phase_filter):
// returns the K-major filter [4*Cout][9*Cin] (row = phase*Cout + co, col = (u+1)*3*Cin + (v+1)*Cin + ci)
V phase_filter(const V& w, int cout, int cin) {
  V W3((size_t)4 * cout * 9 * cin, 0.0);
  for (int py = 0; py < 2; ++py) for (int px = 0; px < 2; ++px)
    for (int r = 0; r < 3; ++r) for (int s2 = 0; s2 < 3; ++s2) {
      // output phase py: filter row r lands on low-res offset u: py = 0: r0 -> -1, r1,r2 -> 0;  py = 1: r0,r1 -> 0, r2 -> +1
      const int u = py == 0 ? (r == 0 ? -1 : 0) : (r == 2 ? 1 : 0);
      const int v = px == 0 ? (s2 == 0 ? -1 : 0) : (s2 == 2 ? 1 : 0);
      for (int co = 0; co < cout; ++co) for (int ci = 0; ci < cin; ++ci)
        W3[((size_t)((py * 2 + px) * cout + co) * 9 + (u + 1) * 3 + (v + 1)) * cin + ci] += w[(((size_t)co * cin + ci) * 3 + r) * 3 + s2];
    }
  return W3;
}

constexpr double SCALE = 0.17677669529663687;   // 32^-0.5

}  // namespace

extern "C" long long dahitra_prepare_weights(const dh_tensor* tensors, int n_tensors, int variant, int output_nc, float* out_buf,
                                             long long out_floats, long long* slot_offsets) {
  if (!tensors || n_tensors <= 0) return DH_E_NULL;
  if (variant != DH_VARIANT_LEVIR && variant != DH_VARIANT_XBD) return DH_E_VARIANT;
  if (output_nc < 1 || output_nc > 8) return DH_E_SHAPE;
  Src s;
  for (int i = 0; i < n_tensors; ++i) if (tensors[i].name) s.m[tensors[i].name] = &tensors[i];
  struct TrunkConv { int w, b, wt; const char* conv; const char* bn; int O, I, K; };
  const TrunkConv trunk[] = {
      {DH_W_L1_0_C1_W, DH_W_L1_0_C1_B, DH_W_L1_0_C1_WT, "resnet.layer1.0.conv1", "resnet.layer1.0.bn1", 64, 64, 3},
      {DH_W_L1_0_C2_W, DH_W_L1_0_C2_B, DH_W_L1_0_C2_WT, "resnet.layer1.0.conv2", "resnet.layer1.0.bn2", 64, 64, 3},
      {DH_W_L1_1_C1_W, DH_W_L1_1_C1_B, DH_W_L1_1_C1_WT, "resnet.layer1.1.conv1", "resnet.layer1.1.bn1", 64, 64, 3},
      {DH_W_L1_1_C2_W, DH_W_L1_1_C2_B, DH_W_L1_1_C2_WT, "resnet.layer1.1.conv2", "resnet.layer1.1.bn2", 64, 64, 3},
      {DH_W_L2_0_C1_W, DH_W_L2_0_C1_B, DH_W_L2_0_C1_WT, "resnet.layer2.0.conv1", "resnet.layer2.0.bn1", 128, 64, 3},
      {DH_W_L2_0_C2_W, DH_W_L2_0_C2_B, DH_W_L2_0_C2_WT, "resnet.layer2.0.conv2", "resnet.layer2.0.bn2", 128, 128, 3},
      {DH_W_L2_0_DS_W, DH_W_L2_0_DS_B, DH_W_L2_0_DS_WT, "resnet.layer2.0.downsample.0", "resnet.layer2.0.downsample.1", 128, 64, 1},
      {DH_W_L2_1_C1_W, DH_W_L2_1_C1_B, DH_W_L2_1_C1_WT, "resnet.layer2.1.conv1", "resnet.layer2.1.bn1", 128, 128, 3},
      {DH_W_L2_1_C2_W, DH_W_L2_1_C2_B, DH_W_L2_1_C2_WT, "resnet.layer2.1.conv2", "resnet.layer2.1.bn2", 128, 128, 3},
      {DH_W_L3_0_C1_W, DH_W_L3_0_C1_B, DH_W_L3_0_C1_WT, "resnet.layer3.0.conv1", "resnet.layer3.0.bn1", 256, 128, 3},
      {DH_W_L3_0_C2_W, DH_W_L3_0_C2_B, DH_W_L3_0_C2_WT, "resnet.layer3.0.conv2", "resnet.layer3.0.bn2", 256, 256, 3},
      {DH_W_L3_0_DS_W, DH_W_L3_0_DS_B, DH_W_L3_0_DS_WT, "resnet.layer3.0.downsample.0", "resnet.layer3.0.downsample.1", 256, 128, 1},
      {DH_W_L3_1_C1_W, DH_W_L3_1_C1_B, DH_W_L3_1_C1_WT, "resnet.layer3.1.conv1", "resnet.layer3.1.bn1", 256, 256, 3},
      {DH_W_L3_1_C2_W, DH_W_L3_1_C2_B, DH_W_L3_1_C2_WT, "resnet.layer3.1.conv2", "resnet.layer3.1.bn2", 256, 256, 3}};
  // every slot's payload is collected first and then laid out in ENUM order at 256-byte granules, like engine.PreparedWeights
  std::vector<std::vector<float>> payload(DH_W_COUNT);
  std::vector<char> present(DH_W_COUNT, 0);
  auto emit = [&](int slot, const V& v) { payload[slot].assign(v.size(), 0.f); for (size_t i = 0; i < v.size(); ++i) payload[slot][i] = (float)v[i]; present[slot] = 1; };
  auto emit_split = [&](int slot, const V& wt) { payload[slot].assign(5 * wt.size(), 0.f); kmajor_split(payload[slot].data(), wt); present[slot] = 1; };
  auto emit_conv = [&](int wslot, int bslot, int wtslot, const Conv& c) {
    const V kh = khwc(c.w, c.O, c.I, c.K);
    emit(wslot, kh);
    if (bslot >= 0) emit(bslot, c.b);
    if (wtslot >= 0) emit_split(wtslot, transpose(kh, c.K * c.K * c.I, c.O));
  };
  // ---- trunk
  {
    const Conv c = fold(s, "resnet.conv1", "resnet.bn1", 64, 3, 7);
    emit(DH_W_STEM_W, khwc(c.w, 64, 3, 7));               // [147][64], K = (r, s, ci)
    emit(DH_W_STEM_B, c.b);
  }
  for (const TrunkConv& t : trunk) emit_conv(t.w, t.b, t.wt, fold(s, t.conv, t.bn, t.O, t.I, t.K));

  // ---- transformer levels
  const int lv_k[3] = {5, 4, 3}, lv_cin[3] = {256, 128, 64}, lv_heads[3] = {4, 4, 8}, lv_depth[3] = {4, 4, 8};
  for (int li = 0; li < 3; ++li) {
    const int k = lv_k[li], cin = lv_cin[li], H = lv_heads[li], depth = lv_depth[li], inner = H * 64;
    const std::string ks = std::to_string(k);
    const int base = DH_W_LV5_SQ + 6 * li;
    // squeeze [Cin][32] (+ K-major split [32][Cin]), token conv [32][4], conv_decode
    const V sqw = s.get("conv_squeeze_" + ks + ".0.weight", (size_t)32 * cin);          // (32, cin, 1, 1)
    emit(base + 0, transpose(sqw, 32, cin));
    emit_split(DH_W_LV5_SQ_WT + li, sqw);
    emit(base + 1, transpose(s.get("conv_token_" + ks + ".weight", 4 * 32), 4, 32));
    {
      Conv c; c.O = 32; c.I = 64; c.K = 3; c.w = s.get("conv_decode_" + ks + ".weight", (size_t)32 * 64 * 9);
      const V kh = khwc(c.w, 32, 64, 3);
      emit(base + 5, kh);
      emit_split(DH_W_LV5_DECODE_WT + li, transpose(kh, 9 * 64, 32));
    }
    // ---- token encoder pack
    {
      const std::string t = "transformer_" + ks + ".layers.0";
      V pack;
      V pos(256, 0.0);
      if (variant == DH_VARIANT_LEVIR) { if (s.has("pos_embedding_" + ks)) pos = s.get("pos_embedding_" + ks, 256); }
      else if (k == 5 && s.has("pos_embedding_3")) pos = s.get("pos_embedding_3", 256);
      const V g1 = s.get(t + ".0.fn.norm.weight", 32), b1n = s.get(t + ".0.fn.norm.bias", 32);
      const V wqkv = s.get(t + ".0.fn.fn.to_qkv.weight", (size_t)3 * inner * 32);
      const V wo = s.get(t + ".0.fn.fn.to_out.0.weight", (size_t)32 * inner);
      V mqk((size_t)H * 1024), mvoT((size_t)H * 1024);
      for (int h = 0; h < H; ++h) for (int c = 0; c < 32; ++c) for (int e = 0; e < 32; ++e) {
        double a = 0.0, v = 0.0;
        for (int d = 0; d < 64; ++d) {
          a += wqkv[((size_t)(h * 64 + d)) * 32 + c] * wqkv[((size_t)(inner + h * 64 + d)) * 32 + e];        // Wq[hd][c] Wk[hd][e]
          v += wo[(size_t)c * inner + h * 64 + d] * wqkv[((size_t)(2 * inner + h * 64 + d)) * 32 + e];      // Wo[c][hd] Wv[hd][e]
        }
        mqk[((size_t)h * 32 + c) * 32 + e] = SCALE * a;
        mvoT[((size_t)h * 32 + e) * 32 + c] = v;
      }
      auto app = [&](const V& v) { pack.insert(pack.end(), v.begin(), v.end()); };
      app(pos); app(g1); app(b1n); app(mqk); app(mvoT); app(s.get(t + ".0.fn.fn.to_out.0.bias", 32));
      app(s.get(t + ".1.fn.norm.weight", 32)); app(s.get(t + ".1.fn.norm.bias", 32));
      app(transpose(s.get(t + ".1.fn.fn.net.0.weight", 1024), 32, 32)); app(s.get(t + ".1.fn.fn.net.0.bias", 32));
      app(transpose(s.get(t + ".1.fn.fn.net.3.weight", 1024), 32, 32)); app(s.get(t + ".1.fn.fn.net.3.bias", 32));
      emit(base + 2, pack);
    }
    // ---- pixel decoder packs (CUDA-core layout and tensor-core layout)
    {
      V dec, dectc, cum(32, 0.0);
      for (int l = 0; l < depth; ++l) {
        const std::string d = "transformer_decoder_" + ks + ".layers." + std::to_string(l);
        const V wq = s.get(d + ".0.fn.fn.to_q.weight", (size_t)inner * 32), wk = s.get(d + ".0.fn.fn.to_k.weight", (size_t)inner * 32),
                wv = s.get(d + ".0.fn.fn.to_v.weight", (size_t)inner * 32), wo = s.get(d + ".0.fn.fn.to_out.0.weight", (size_t)32 * inner);
        const V bo = s.get(d + ".0.fn.fn.to_out.0.bias", 32);
        const V g2 = s.get(d + ".1.fn.norm.weight", 32), b2n = s.get(d + ".1.fn.norm.bias", 32);
        const V w1 = s.get(d + ".1.fn.fn.net.0.weight", 1024), b1 = s.get(d + ".1.fn.fn.net.0.bias", 32);
        const V w2 = s.get(d + ".1.fn.fn.net.3.weight", 1024), b2 = s.get(d + ".1.fn.fn.net.3.bias", 32);
        V mqkT((size_t)H * 1024), movT((size_t)H * 1024);
        for (int h = 0; h < H; ++h) for (int c = 0; c < 32; ++c) for (int e = 0; e < 32; ++e) {
          double a = 0.0, v = 0.0;
          for (int dd = 0; dd < 64; ++dd) {
            a += wq[((size_t)(h * 64 + dd)) * 32 + c] * wk[((size_t)(h * 64 + dd)) * 32 + e];
            v += wo[(size_t)c * inner + h * 64 + dd] * wv[((size_t)(h * 64 + dd)) * 32 + e];
          }
          mqkT[((size_t)h * 32 + e) * 32 + c] = SCALE * a;
          movT[((size_t)h * 32 + e) * 32 + c] = v;
        }
        V w1g(1024), b1f(32);                                   // second LayerNorm folded: W1f[o][c] = W1[o][c] g2[c]; b1f = b1 + W1 b2n
        for (int o = 0; o < 32; ++o) {
          double acc = 0.0;
          for (int c = 0; c < 32; ++c) { w1g[o * 32 + c] = w1[o * 32 + c] * g2[c]; acc += w1[o * 32 + c] * b2n[c]; }
          b1f[o] = b1[o] + acc;
        }
        auto app = [&](V& dst, const V& v) { dst.insert(dst.end(), v.begin(), v.end()); };
        app(dec, s.get(d + ".0.fn.norm.weight", 32)); app(dec, s.get(d + ".0.fn.norm.bias", 32));
        app(dec, mqkT); app(dec, movT); app(dec, bo);
        app(dec, transpose(w1g, 32, 32)); app(dec, b1f); app(dec, transpose(w2, 32, 32)); app(dec, b2);
        // tensor-core layer: [W1f_hi swz][W2_hi swz][b1f][cbA][cbM][W1f_lo swz][W2_lo swz]; B[n][k]: W1f n = hidden o, k = channel c; W2 n = channel, k = hidden
        V cbA(32), w1h(1024), w1l(1024), w2h(1024), w2l(1024);
        for (int c = 0; c < 32; ++c) { cbA[c] = cum[c] + bo[c]; cum[c] = cbA[c] + b2[c]; }
        for (int n = 0; n < 32; ++n) for (int kk = 0; kk < 32; ++kk) {
          const double a = w1g[n * 32 + kk], b = w2[n * 32 + kk];
          const float ah = tf32_round_f((float)a), bh = tf32_round_f((float)b);
          w1h[sw128(n, kk)] = ah; w1l[sw128(n, kk)] = tf32_round_f((float)(a - (double)ah));
          w2h[sw128(n, kk)] = bh; w2l[sw128(n, kk)] = tf32_round_f((float)(b - (double)bh));
        }
        app(dectc, w1h); app(dectc, w2h); app(dectc, b1f); app(dectc, cbA); app(dectc, cum); app(dectc, w1l); app(dectc, w2l);
      }
      emit(base + 3, dec);
      emit(DH_W_LV5_DECTC + li, dectc);
    }
    // ---- decoder positional embedding (1, 32, h, w) -> [h*w][32]
    {
      std::string key;
      if (variant == DH_VARIANT_LEVIR) key = "pos_embedding_decoder_" + ks;
      else if (k == 5) key = "pos_embedding_decoder_3";
      const dh_tensor* t = key.empty() ? nullptr : s.find(key);
      if (t && t->ndim == 4 && t->shape[1] == 32) {
        const int hw = (int)(t->shape[2] * t->shape[3]);
        const V pe = s.get(key, (size_t)32 * hw);
        V r((size_t)hw * 32);
        for (int c = 0; c < 32; ++c) for (int p = 0; p < hw; ++p) r[(size_t)p * 32 + c] = pe[(size_t)c * hw + p];
        emit(base + 4, r);
      }
    }
  }
  // ---- UNet head
  struct HeadConv { int w, b, psw, psb; const char* key; };
  const HeadConv up[3] = {{DH_W_CL4_W, DH_W_CL4_B, DH_W_CL4_PSWT, DH_W_CL4_PSB, "conv_layer4.0"},
                          {DH_W_CL3_W, DH_W_CL3_B, DH_W_CL3_PSWT, DH_W_CL3_PSB, "conv_layer3.0"},
                          {DH_W_CL2_W, DH_W_CL2_B, DH_W_CL2_PSWT, DH_W_CL2_PSB, "conv_layer2.0"}};
  for (const HeadConv& h : up) {
    const Conv c = fold(s, h.key, "", 32, 32, 3);
    emit_conv(h.w, h.b, -1, c);
    emit_split(h.psw, phase_filter(c.w, 32, 32));
    V pb(128);
    for (int i = 0; i < 128; ++i) pb[i] = c.b.empty() ? 0.0 : c.b[i % 32];
    emit(h.psb, pb);
  }
  emit_conv(DH_W_CL20A_W, DH_W_CL20A_B, DH_W_CL20A_WT, fold(s, "conv_layer2_0.0", "conv_layer2_0.1", 128, 128, 3));
  emit_conv(DH_W_CL20B_W, DH_W_CL20B_B, DH_W_CL20B_WT, fold(s, "conv_layer2_0.3", "", 32, 128, 3));
  {
    const V wc = s.get("classifier.weight", (size_t)output_nc * 32 * 9);          // (nc, 32, 3, 3) -> [9][nc][32]
    V r(wc.size());
    for (int o = 0; o < output_nc; ++o) for (int c = 0; c < 32; ++c) for (int t = 0; t < 9; ++t)
      r[((size_t)t * output_nc + o) * 32 + c] = wc[((size_t)o * 32 + c) * 9 + t];
    emit(DH_W_CLS_W, r);
    emit(DH_W_CLS_B, s.get("classifier.bias", output_nc));
  }
  {   // stem filter images for the tcgen05 stem (engine.stem_tc_image): K re-ordered to (ci, r, s8) = 21 groups of one zero + 7 taps,
      // padded to 192; TF32 hi / lo tiles, the folded FP16 tiles [f16(w) ; f16(2^11 r_w)] and the bf16 tiles, all pre-swizzled
    const Conv c = fold(s, "resnet.conv1", "resnet.bn1", 64, 3, 7);
    const V kh = khwc(c.w, 64, 3, 7);
    V wk((size_t)192 * 64, 0.0);
    for (int ci = 0; ci < 3; ++ci) for (int r = 0; r < 7; ++r) for (int s2 = 0; s2 < 7; ++s2) for (int co = 0; co < 64; ++co)
      wk[((size_t)((ci * 7 + r) * 8 + 1 + s2)) * 64 + co] = kh[((size_t)(r * 7 + s2) * 3 + ci) * 64 + co];
    std::vector<float>& img = payload[DH_W_STEM_WTC];
    img.assign(43008, 0.f);
    float* hi = img.data(); float* lo = hi + 12288;
    uint16_t* m16 = reinterpret_cast<uint16_t*>(hi + 24576); uint16_t* c16 = reinterpret_cast<uint16_t*>(hi + 36864);
    for (int kt = 0; kt < 6; ++kt) for (int n = 0; n < 64; ++n) for (int kk = 0; kk < 32; ++kk) {
      const double w = wk[(size_t)(kt * 32 + kk) * 64 + n];
      const float h = tf32_round_f((float)w);
      hi[kt * 2048 + sw128(n, kk)] = h;
      lo[kt * 2048 + sw128(n, kk)] = tf32_round_f((float)(w - (double)h));
    }
    for (int kt = 0; kt < 3; ++kt) for (int kk = 0; kk < 64; ++kk) for (int n = 0; n < 64; ++n) {
      const double w = wk[(size_t)(kt * 64 + kk) * 64 + n];
      const uint16_t h16 = f16_bits((float)clamp16(w));
      const double r = w - (double)f16_val(h16);
      m16[kt * 128 * 64 + sw128_16(n, kk)] = h16;
      m16[kt * 128 * 64 + sw128_16(64 + n, kk)] = f16_bits((float)clamp16(r * 2048.0));
      c16[kt * 64 * 64 + sw128_16(n, kk)] = bf16_bits((float)w);
    }
    present[DH_W_STEM_WTC] = 1;
  }
  if (!s.ok) return DH_E_WEIGHTS;
  // ---- lay the slots out in enum order
  long long total = 0;
  for (int i = 0; i < DH_W_COUNT; ++i) if (present[i]) total += (long long)((payload[i].size() + 63) / 64 * 64);
  if (!out_buf) return total;
  if (out_floats < total) return DH_E_WORKSPACE;
  long long off = 0;
  for (int i = 0; i < DH_W_COUNT; ++i) {
    if (!present[i]) { if (slot_offsets) slot_offsets[i] = -1; continue; }
    memcpy(out_buf + off, payload[i].data(), payload[i].size() * sizeof(float));
    const long long padded = (long long)((payload[i].size() + 63) / 64 * 64);
    for (long long j = (long long)payload[i].size(); j < padded; ++j) out_buf[off + j] = 0.f;
    if (slot_offsets) slot_offsets[i] = off;
    off += padded;
  }
  return total;
}
