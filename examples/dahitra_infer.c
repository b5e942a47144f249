/*
 * dahitra_infer — a host for the C ABI (include/dahitra_b200.h) written in plain C: no Python, no PyTorch.
 *
 * It does what the reference's evaluator does around `net_G(img_in1, img_in2)` (reference models/evaluator.py:156-180,
 * 89-92) for a consumer that links libdahitra_b200.so directly:
 *
 *   checkpoint tensors (by their reference state_dict keys)  --dahitra_prepare_weights (host)-->  prepared slots
 *   --one cudaMemcpy-->  device slot table  --dahitra_workspace_bytes / dahitra_forward-->  logits + uint8 class map
 *
 * Files (little endian; written / read by dahitra_b200/checkpoints.py: export_state_dict_bin, write_pairs_bin, read_result_bin):
 *   weights  "DHSD0001" | int32 n | n x { int32 name_len | name | int32 dtype (0 f32, 1 f64) | int32 ndim | int64 shape[4] |
 *                                        int64 nbytes | data }
 *   input    "DHIN0001" | int32 B, C, H, W | C = 3: x1 then x2, each (B,3,H,W) fp32  (LEVIR: two images per pair)
 *                                          | C = 6: one (B,6,H,W) fp32 tensor         (xBD: pre | post stacked on the channels)
 *   output   "DHOUT001" | int32 B, nc, H, W | logits (B,nc,H,W) fp32 | class map (B,H,W) uint8
 *
 * Usage
 *   dahitra_infer --weights w.bin --prepare-only [--variant V] [--nc N]             host only: size + FNV-1a of the slots
 *   dahitra_infer --weights w.bin --input x.bin --output y.bin [--variant V] [--nc N] [--flags F] [--repeat R]
 *   dahitra_infer --weights w.bin --synthetic BxHxW [--output y.bin] ...      uniform [-1, 1) images made on the host instead of a file
 * --repeat R times R further steps, each = H2D of the step's inputs from pinned memory + forward + D2H of the class map
 * (CUDA events on the launching stream), and prints pairs/s: the end-to-end rate of the C ABI without any Python
 * (one stream, so the copies do not overlap the forward; dahitra_b200/pipeline.py shows the overlapped form).
 * --resident leaves the copies out of the timed steps (inputs stay in HBM): the forward alone.
 *
 * Build: __graft_entry__.build() (or:  gcc -std=c99 -O2 -Iinclude -I/usr/local/cuda/include examples/dahitra_infer.c
 *        -Ldahitra_b200 -ldahitra_b200 -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/dahitra_b200 -o examples/bin/dahitra_infer)
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cuda_runtime_api.h>

#include "dahitra_b200.h"

#define DIE(...) do { fprintf(stderr, "dahitra_infer: " __VA_ARGS__); fputc('\n', stderr); exit(1); } while (0)
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) DIE("%s -> %s", #call, cudaGetErrorString(e_)); } while (0)
#define DH(call) do { int r_ = (call); if (r_ != 0) DIE("%s -> %d (%s)", #call, r_, r_ < 0 ? dahitra_error_string(r_) : cudaGetErrorString((cudaError_t)r_)); } while (0)

static void read_exact(FILE* f, void* dst, size_t n, const char* what) {
  if (n && fread(dst, 1, n, f) != n) DIE("short read (%s)", what);
}

static uint64_t fnv1a(const void* p, size_t n, uint64_t h) {
  const unsigned char* b = (const unsigned char*)p;
  for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
  return h;
}

/* the checkpoint: names and host copies of every tensor, as dahitra_prepare_weights takes them */
static dh_tensor* read_weights(const char* path, int* n_out) {
  FILE* f = fopen(path, "rb");
  if (!f) DIE("cannot open %s", path);
  char magic[8];
  int32_t n;
  read_exact(f, magic, 8, "magic");
  if (memcmp(magic, "DHSD0001", 8)) DIE("%s is not a DHSD0001 file", path);
  read_exact(f, &n, 4, "count");
  if (n < 1 || n > 100000) DIE("implausible tensor count %d", n);
  dh_tensor* t = (dh_tensor*)calloc((size_t)n, sizeof(dh_tensor));
  for (int i = 0; i < n; ++i) {
    int32_t len, dtype, ndim;
    int64_t shape[4], nbytes;
    read_exact(f, &len, 4, "name length");
    if (len < 1 || len > 4096) DIE("implausible name length %d", len);
    char* name = (char*)malloc((size_t)len + 1);
    read_exact(f, name, (size_t)len, "name");
    name[len] = 0;
    read_exact(f, &dtype, 4, "dtype");
    read_exact(f, &ndim, 4, "ndim");
    read_exact(f, shape, 32, "shape");
    read_exact(f, &nbytes, 8, "nbytes");
    if ((dtype != DH_DTYPE_F32 && dtype != DH_DTYPE_F64) || ndim < 0 || ndim > 4) DIE("tensor %s: bad dtype / ndim", name);
    int64_t numel = 1;
    for (int j = 0; j < ndim; ++j) numel *= shape[j];
    if (nbytes != numel * (dtype == DH_DTYPE_F64 ? 8 : 4)) DIE("tensor %s: %lld bytes for %lld elements", name, (long long)nbytes, (long long)numel);
    void* data = malloc(nbytes ? (size_t)nbytes : 8);
    read_exact(f, data, (size_t)nbytes, name);
    t[i].name = name;
    t[i].data = data;
    t[i].dtype = dtype;
    t[i].ndim = ndim;
    for (int j = 0; j < 4; ++j) t[i].shape[j] = j < ndim ? shape[j] : 1;
  }
  fclose(f);
  *n_out = n;
  return t;
}

int main(int argc, char** argv) {
  const char *wpath = NULL, *ipath = NULL, *opath = NULL, *synth = NULL;
  int variant = DH_VARIANT_LEVIR, nc = 2, flags = DH_FLAGS_TF32X3, repeat = 0, prepare_only = 0, resident = 0;
  for (int i = 1; i < argc; ++i) {
    const char* a = argv[i];
    const char* v = i + 1 < argc ? argv[i + 1] : NULL;
    if (!strcmp(a, "--prepare-only")) prepare_only = 1;
    else if (!strcmp(a, "--resident")) resident = 1;
    else if (!v) DIE("missing value after %s", a);
    else if (!strcmp(a, "--weights")) { wpath = v; ++i; }
    else if (!strcmp(a, "--input")) { ipath = v; ++i; }
    else if (!strcmp(a, "--output")) { opath = v; ++i; }
    else if (!strcmp(a, "--synthetic")) { synth = v; ++i; }
    else if (!strcmp(a, "--variant")) { variant = atoi(v); ++i; }
    else if (!strcmp(a, "--nc")) { nc = atoi(v); ++i; }
    else if (!strcmp(a, "--flags")) { flags = (int)strtol(v, NULL, 0); ++i; }
    else if (!strcmp(a, "--repeat")) { repeat = atoi(v); ++i; }
    else DIE("unknown option %s", a);
  }
  if (!wpath) DIE("--weights is required");
  if (dahitra_version() != DAHITRA_ABI_VERSION) DIE("library ABI %d, header ABI %d", dahitra_version(), DAHITRA_ABI_VERSION);

  /* ---- 1. checkpoint -> prepared slots (host code only) ---- */
  int n_tensors = 0;
  dh_tensor* tensors = read_weights(wpath, &n_tensors);
  long long n_floats = dahitra_prepare_weights(tensors, n_tensors, variant, nc, NULL, 0, NULL);
  if (n_floats <= 0) DIE("dahitra_prepare_weights (size query) -> %lld (%s)", n_floats, dahitra_error_string((int)n_floats));
  long long offs[DH_W_COUNT];
  float* prepared = NULL;
  if (prepare_only) {
    prepared = (float*)malloc((size_t)n_floats * 4);
  } else {
    CU(cudaMallocHost((void**)&prepared, (size_t)n_floats * 4));
  }
  long long wrote = dahitra_prepare_weights(tensors, n_tensors, variant, nc, prepared, n_floats, offs);
  if (wrote != n_floats) DIE("dahitra_prepare_weights -> %lld, expected %lld", wrote, n_floats);
  if (prepare_only) {
    int present = 0;
    for (int s = 0; s < DH_W_COUNT; ++s) present += offs[s] >= 0;
    printf("tensors=%d floats=%lld slots=%d present=%d data_fnv1a=%016llx offsets_fnv1a=%016llx\n", n_tensors, n_floats, (int)DH_W_COUNT, present,
           (unsigned long long)fnv1a(prepared, (size_t)n_floats * 4, 14695981039346656037ull),
           (unsigned long long)fnv1a(offs, sizeof(offs), 14695981039346656037ull));
    return 0;
  }
  if (!synth && (!ipath || !opath)) DIE("--input and --output are required (or --synthetic BxHxW, or --prepare-only)");

  /* ---- 2. inputs ---- */
  int32_t dims[4];
  if (synth) {
    int b = 0, h = 0, w = 0;
    if (sscanf(synth, "%dx%dx%d", &b, &h, &w) != 3) DIE("--synthetic expects BxHxW, got %s", synth);
    dims[0] = b; dims[1] = variant == DH_VARIANT_XBD ? 6 : 3; dims[2] = h; dims[3] = w;
  } else {
    FILE* f = fopen(ipath, "rb");
    if (!f) DIE("cannot open %s", ipath);
    char magic[8];
    read_exact(f, magic, 8, "magic");
    if (memcmp(magic, "DHIN0001", 8)) DIE("%s is not a DHIN0001 file", ipath);
    read_exact(f, dims, 16, "dims");
    fclose(f);
  }
  const int B = dims[0], C = dims[1], H = dims[2], W = dims[3];
  if (B < 1 || B > 4096 || (C != 3 && C != 6) || H < 32 || W < 32 || H > 8192 || W > 8192) DIE("bad input dims %d %d %d %d", B, C, H, W);
  const size_t in_floats = (size_t)B * 6 * H * W;              /* both images of every pair */
  float* h_in = NULL;
  CU(cudaMallocHost((void**)&h_in, in_floats * 4));
  if (synth) {
    uint32_t st = 12345u;                                      /* LCG; top 24 bits -> [-1, 1) */
    for (size_t i = 0; i < in_floats; ++i) { st = st * 1664525u + 1013904223u; h_in[i] = (float)(st >> 8) * (2.0f / 16777216.0f) - 1.0f; }
  } else {
    FILE* f = fopen(ipath, "rb");
    if (!f || fseek(f, 24, SEEK_SET)) DIE("cannot re-open %s", ipath);
    read_exact(f, h_in, in_floats * 4, "images");
    fclose(f);
  }
  /* C = 3: [x1 (B,3,H,W) | x2 (B,3,H,W)], images 3HW apart; C = 6: (B,6,H,W), pre / post halves of one image 6HW apart */
  const long long x_batch_stride = (C == 3 ? 3ll : 6ll) * H * W;
  const size_t x2_off = C == 3 ? (size_t)B * 3 * H * W : (size_t)3 * H * W;

  /* ---- 3. device memory: everything is the caller's ---- */
  cudaStream_t stream;
  CU(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  float *d_w = NULL, *d_in = NULL, *d_logits = NULL;
  unsigned char *d_map = NULL, *d_ws = NULL;
  const size_t ws_bytes = dahitra_workspace_bytes(variant, B, H, W, nc, flags);
  if (!ws_bytes) DIE("dahitra_workspace_bytes -> 0 (unsupported shape / variant / flags)");
  const size_t logit_floats = (size_t)B * nc * H * W, map_bytes = (size_t)B * H * W;
  CU(cudaMalloc((void**)&d_w, (size_t)n_floats * 4));
  CU(cudaMalloc((void**)&d_in, in_floats * 4));
  CU(cudaMalloc((void**)&d_logits, logit_floats * 4));
  CU(cudaMalloc((void**)&d_map, map_bytes));
  CU(cudaMalloc((void**)&d_ws, ws_bytes));
  CU(cudaMemcpyAsync(d_w, prepared, (size_t)n_floats * 4, cudaMemcpyHostToDevice, stream));
  const void* table[DH_W_COUNT];
  for (int s = 0; s < DH_W_COUNT; ++s) table[s] = offs[s] >= 0 ? (const void*)(d_w + offs[s]) : NULL;

  /* ---- 4. forward ---- */
  float* h_logits = NULL;
  unsigned char* h_map = NULL;
  CU(cudaMallocHost((void**)&h_logits, logit_floats * 4));
  CU(cudaMallocHost((void**)&h_map, map_bytes));
  CU(cudaMemcpyAsync(d_in, h_in, in_floats * 4, cudaMemcpyHostToDevice, stream));
  DH(dahitra_forward(table, DH_W_COUNT, d_in, d_in + x2_off, x_batch_stride, d_logits, d_map, d_ws, ws_bytes,
                     variant, B, H, W, nc, flags, stream));
  CU(cudaMemcpyAsync(h_logits, d_logits, logit_floats * 4, cudaMemcpyDeviceToHost, stream));
  CU(cudaMemcpyAsync(h_map, d_map, map_bytes, cudaMemcpyDeviceToHost, stream));
  CU(cudaStreamSynchronize(stream));

  if (opath) {
    FILE* f = fopen(opath, "wb");
    if (!f) DIE("cannot write %s", opath);
    const int32_t odims[4] = {B, nc, H, W};
    fwrite("DHOUT001", 1, 8, f);
    fwrite(odims, 4, 4, f);
    fwrite(h_logits, 4, logit_floats, f);
    fwrite(h_map, 1, map_bytes, f);
    if (fclose(f)) DIE("write to %s failed", opath);
  }
  {                                                            /* a fingerprint of the result for runs without an output file */
    double sum = 0.0;
    size_t cls1 = 0;
    for (size_t i = 0; i < logit_floats; ++i) sum += h_logits[i];
    for (size_t i = 0; i < map_bytes; ++i) cls1 += h_map[i] != 0;
    printf("logits: sum=%.6e fnv1a=%016llx; class map: %zu of %zu pixels not class 0\n", sum,
           (unsigned long long)fnv1a(h_logits, logit_floats * 4, 14695981039346656037ull), cls1, map_bytes);
  }
  printf("forward ok: B=%d nc=%d %dx%d variant=%d flags=%d workspace=%.1f MB prepared=%.1f MB\n", B, nc, H, W, variant, flags,
         ws_bytes / 1e6, n_floats * 4 / 1e6);

  /* ---- 5. optional: steady-state rate of the whole call sequence, host buffers in, class map out ---- */
  if (repeat > 0) {
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    for (int it = -3; it < repeat; ++it) {                     /* 3 untimed steps first */
      if (it == 0) CU(cudaEventRecord(e0, stream));
      if (!resident) CU(cudaMemcpyAsync(d_in, h_in, in_floats * 4, cudaMemcpyHostToDevice, stream));
      DH(dahitra_forward(table, DH_W_COUNT, d_in, d_in + x2_off, x_batch_stride, d_logits, d_map, d_ws, ws_bytes,
                         variant, B, H, W, nc, flags, stream));
      if (!resident) CU(cudaMemcpyAsync(h_map, d_map, map_bytes, cudaMemcpyDeviceToHost, stream));
    }
    CU(cudaEventRecord(e1, stream));
    CU(cudaEventSynchronize(e1));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, e0, e1));
    printf("repeat=%d ms_per_step=%.4f pairs_per_s=%.1f h2d_bytes_per_step=%zu d2h_bytes_per_step=%zu (%s)\n",
           repeat, ms / repeat, 1e3 * B * repeat / ms, resident ? (size_t)0 : in_floats * 4, resident ? (size_t)0 : map_bytes,
           resident ? "inputs resident in HBM, forward only" : "copies and forward serialised on one stream");
    CU(cudaEventDestroy(e0));
    CU(cudaEventDestroy(e1));
  }

  CU(cudaFree(d_ws)); CU(cudaFree(d_map)); CU(cudaFree(d_logits)); CU(cudaFree(d_in)); CU(cudaFree(d_w));
  CU(cudaFreeHost(h_map)); CU(cudaFreeHost(h_logits)); CU(cudaFreeHost(h_in)); CU(cudaFreeHost(prepared));
  CU(cudaStreamDestroy(stream));
  return 0;
}
