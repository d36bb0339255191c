// Handle management, chunk gather / merge, and the encoder / decoder drivers.
#include <stdarg.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

void lc_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* lc_last_error(void) { return g_err; }
extern "C" int lc_version(void) { return 100; }

extern "C" int lc_chunk_grid(int T, int H, int W, int grid[3]) {
  LC_REQUIRE(T >= CT && H >= CH && W >= CW, "Chunks size cannot be larger than the data size.");
  ChunkGeom g = make_geom(T, H, W);
  grid[0] = g.gt; grid[1] = g.gh; grid[2] = g.gw;
  return LC_OK;
}

// Host-only: the chunk index map the kernels use (dataLoader.py:457-481 origins, :503-532 kept part).
extern "C" int lc_chunk_window(int T, int H, int W, int idx, int out[6]) {
  LC_REQUIRE(T >= CT && H >= CH && W >= CW, "Chunks size cannot be larger than the data size.");
  ChunkGeom g = make_geom(T, H, W);
  LC_REQUIRE(idx >= 0 && idx < g.gt * g.gh * g.gw && out, "chunk index out of range");
  g.origin(idx, out[0], out[1], out[2]);
  g.keep_from(idx, out[3], out[4], out[5]);
  return LC_OK;
}

static int same_pad_lo(int n, int k, int s) {
  int out = (n + s - 1) / s;
  int total = (out - 1) * s + k - n;
  if (total < 0) total = 0;
  return total / 2;
}

extern "C" int lc_create(const lc_layer_t* layers, int n_enc, int n_dec, int in_channels,
                         int precision, int device, lc_handle_t* out) {
  LC_REQUIRE(layers && out, "null argument");
  LC_REQUIRE(n_enc > 0 && n_dec > 0 && n_enc + n_dec <= kMaxLayers, "bad layer count");
  LC_REQUIRE(precision == LC_PREC_FP32 || precision == LC_PREC_BF16 || precision == LC_PREC_FP16,
             "unknown precision");
  // the codec lives on `device`; the calling thread's current device is restored before returning
  struct DeviceGuard {
    int prev = -1;
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
  } guard;
  LC_CUDA(cudaGetDevice(&guard.prev));
  LC_CUDA(cudaSetDevice(device));
  lc_codec* h = new lc_codec();
  memset(h, 0, sizeof(*h));
  h->spans = new std::vector<ProfSpan>();
  h->pool = new std::vector<cudaEvent_t>();
  h->ws_ready = new std::vector<std::pair<void*, size_t>>();
  h->device = device;
  h->precision = precision;
  h->in_channels = in_channels;
  h->n_enc = n_enc;
  h->n_dec = n_dec;
  cudaDeviceProp prop;
  LC_CUDA(cudaGetDeviceProperties(&prop, device));
  h->num_sms = prop.multiProcessorCount;
  h->max_c = in_channels;

  int t = CT, hh = CH, w = CW, c = in_channels;
  for (int i = 0; i < n_enc + n_dec; ++i) {
    const lc_layer_t& S = layers[i];
    DevLayer& L = h->layers[i];
    if (S.cin != c) {
      lc_set_error("layer %d: cin %d does not match the previous layer's %d channels", i, S.cin, c);
      delete h;
      return LC_EINVAL;
    }
    L.kind = S.kind; L.k = S.k; L.stride = S.stride; L.cin = S.cin; L.cout = S.cout;
    L.relu = S.relu; L.res_save = S.res_save; L.res_add = S.res_add;
    L.in_t = t; L.in_h = hh; L.in_w = w;
    if (S.kind == 0) {
      L.out_t = (t + S.stride - 1) / S.stride;
      L.out_h = (hh + S.stride - 1) / S.stride;
      L.out_w = (w + S.stride - 1) / S.stride;
      L.pad_lo = same_pad_lo(t, S.k, S.stride);
      if (same_pad_lo(hh, S.k, S.stride) != L.pad_lo || same_pad_lo(w, S.k, S.stride) != L.pad_lo) {
        lc_set_error("layer %d: anisotropic 'same' padding is not supported", i);
        delete h;
        return LC_EUNSUPPORTED;
      }
    } else {
      L.out_t = t * S.stride; L.out_h = hh * S.stride; L.out_w = w * S.stride;
      L.pad_lo = (S.k - S.stride) / 2;
    }
    const int taps = S.k * S.k * S.k;
    std::vector<float> wr((size_t)taps * S.cin * S.cout);
    for (int tp = 0; tp < taps; ++tp)
      for (int ci = 0; ci < S.cin; ++ci)
        for (int co = 0; co < S.cout; ++co)
          wr[((size_t)tp * S.cin + ci) * S.cout + co] =
              (S.kind == 0) ? S.h_kernel[((size_t)tp * S.cin + ci) * S.cout + co]
                            : S.h_kernel[((size_t)tp * S.cout + co) * S.cin + ci];
    LC_CUDA(cudaMalloc(&L.d_w, wr.size() * sizeof(float)));
    LC_CUDA(cudaMemcpy(L.d_w, wr.data(), wr.size() * sizeof(float), cudaMemcpyHostToDevice));
    LC_CUDA(cudaMalloc(&L.d_b, S.cout * sizeof(float)));
    LC_CUDA(cudaMemcpy(L.d_b, S.h_bias, S.cout * sizeof(float), cudaMemcpyHostToDevice));
    t = L.out_t; hh = L.out_h; w = L.out_w; c = S.cout;
    if (c > h->max_c) h->max_c = c;
    if (i == n_enc - 1) h->latent_floats = t * hh * w * c;
  }
  if (t != CT || hh != CH || w != CW || c != 1) {
    lc_set_error("decoder output is (%d,%d,%d,%d), expected (16,48,48,1)", t, hh, w, c);
    delete h;
    return LC_EINVAL;
  }
  int rc = tc_prepare(h);
  if (rc != LC_OK) { delete h; return rc; }
  *out = h;
  return LC_OK;
}

extern "C" int lc_destroy(lc_handle_t h) {
  if (!h) return LC_OK;
  int prev = -1;
  cudaGetDevice(&prev);
  cudaSetDevice(h->device);
  for (int i = 0; i < h->n_enc + h->n_dec; ++i) {
    cudaFree(h->layers[i].d_w);
    cudaFree(h->layers[i].d_b);
    if (h->layers[i].d_wtc) cudaFree(h->layers[i].d_wtc);
  }
  tc_release(h);
  for (auto& sp : *h->spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
  for (auto& e : *h->pool) cudaEventDestroy(e);
  delete h->spans;
  delete h->pool;
  delete h->ws_ready;
  delete h;
  if (prev >= 0) cudaSetDevice(prev);
  return LC_OK;
}

// ---- per-layer kernel timing (CUDA events on the launching stream) -------------------------
static cudaEvent_t prof_event(lc_codec* h) {
  if (!h->pool->empty()) { cudaEvent_t e = h->pool->back(); h->pool->pop_back(); return e; }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}
void prof_begin(lc_codec* h, int layer, cudaStream_t s) {
  if (!h->profile) return;
  ProfSpan sp{layer, prof_event(h), prof_event(h)};
  cudaEventRecord(sp.a, s);
  h->spans->push_back(sp);
}
void prof_end(lc_codec* h, cudaStream_t s) {
  if (!h->profile) return;
  cudaEventRecord(h->spans->back().b, s);
}
extern "C" int lc_profile_enable(lc_handle_t h, int on) {
  LC_REQUIRE(h, "null handle");
  h->profile = on;
  return LC_OK;
}
extern "C" int lc_profile_read(lc_handle_t h, double* h_ms, long long* h_launches, int n) {
  LC_REQUIRE(h && h_ms && h_launches && n >= h->n_enc + h->n_dec, "bad argument");
  for (int i = 0; i < n; ++i) { h_ms[i] = 0; h_launches[i] = 0; }
  for (auto& sp : *h->spans) {
    LC_CUDA(cudaEventSynchronize(sp.b));
    float ms = 0;
    LC_CUDA(cudaEventElapsedTime(&ms, sp.a, sp.b));
    h_ms[sp.layer] += ms;
    h_launches[sp.layer] += 1;
    h->pool->push_back(sp.a);
    h->pool->push_back(sp.b);
  }
  h->spans->clear();
  return LC_OK;
}

extern "C" int lc_layer_backend(lc_handle_t h, int layer) {
  if (!h || layer < 0 || layer >= h->n_enc + h->n_dec) return LC_EINVAL;
  return tc_backend_of(h, layer);
}

extern "C" int lc_latent_floats(lc_handle_t h) { return h ? h->latent_floats : LC_EINVAL; }

// ---------------------------------------------------------------------------------------
// K2: standardise channel 0 and gather chunks (compress.py:66-76, dataLoader.py:437-481)
// one thread per (voxel, all C channels); channel 0: (x - mean) / std in IEEE fp32, no FMA
// ---------------------------------------------------------------------------------------
__global__ void chunk_gather_kernel(const float* __restrict__ x, ChunkGeom g, int C, float mean,
                                    float stdv, int standardise, int chunk_begin, int n_chunks,
                                    float* __restrict__ out) {
  const long long total = (long long)n_chunks * CT * CH * CW;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < total;
       v += (long long)gridDim.x * blockDim.x) {
    int lw = (int)(v % CW);
    long long r = v / CW;
    int lh = (int)(r % CH);
    r /= CH;
    int lt = (int)(r % CT);
    int chunk = (int)(r / CT);
    int t0, h0, w0;
    g.origin(chunk_begin + chunk, t0, h0, w0);
    const float* src = x + (((size_t)(t0 + lt) * g.H + (h0 + lh)) * g.W + (w0 + lw)) * C;
    float* dst = out + (size_t)v * C;
    dst[0] = standardise ? __fdiv_rn(__fsub_rn(src[0], mean), stdv) : src[0];
    for (int c = 1; c < C; ++c) dst[c] = src[c];
  }
}

// merge (dataLoader.py:484-532): chunk voxel (lt,lh,lw) is kept iff it is at or after the
// chunk's keep_from offsets; kept voxels of all chunks tile the field exactly once.
__global__ void chunk_merge_kernel(const float* __restrict__ dec, ChunkGeom g, int C, int chunk_begin,
                                   int n_chunks, float* __restrict__ recon) {
  const long long total = (long long)n_chunks * CT * CH * CW;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < total;
       v += (long long)gridDim.x * blockDim.x) {
    int lw = (int)(v % CW);
    long long r = v / CW;
    int lh = (int)(r % CH);
    r /= CH;
    int lt = (int)(r % CT);
    int chunk = (int)(r / CT);
    int t0, h0, w0, kt, kh, kw;
    g.origin(chunk_begin + chunk, t0, h0, w0);
    g.keep_from(chunk_begin + chunk, kt, kh, kw);
    if (lt < kt || lh < kh || lw < kw) continue;
    float* dst = recon + (((size_t)(t0 + lt) * g.H + (h0 + lh)) * g.W + (w0 + lw)) * C;
    for (int c = 0; c < C; ++c) dst[c] = dec[(size_t)v * C + c];
  }
}

static size_t act_floats_per_chunk(const lc_codec* h) {
  return (size_t)CT * CH * CW * (size_t)h->max_c;
}

extern "C" size_t lc_ae_workspace_bytes(lc_handle_t h, int max_chunks) {
  if (!h || max_chunks <= 0) return 0;
  if (h->precision != LC_PREC_FP32) return tc_workspace_bytes(h, max_chunks);
  return 3 * act_floats_per_chunk(h) * sizeof(float) * (size_t)max_chunks + 1024;
}

// The tensor-core kernels read the halo and guard slots of the 16-bit activation planes as zeros and never write
// them (act.cuh), so a workspace must be zeroed once before its first use and must not be used for anything else
// afterwards.  lc_ae_workspace_init does the zeroing and registers the buffer with the handle; lc_encode / lc_decode
// refuse a buffer that was not registered (fresh cudaMalloc memory would silently give different convolutions on the
// compressor and the decompressor, i.e. a broken error bound).
extern "C" int lc_ae_workspace_init(lc_handle_t h, void* d_ws, size_t ws_bytes, void* stream) {
  LC_REQUIRE(h && d_ws && ws_bytes > 0, "bad argument");
  LC_CUDA(cudaMemsetAsync(d_ws, 0, ws_bytes, as_stream(stream)));
  for (auto& e : *h->ws_ready)
    if (e.first == d_ws) { e.second = ws_bytes; return LC_OK; }
  if (h->ws_ready->size() >= 16) h->ws_ready->erase(h->ws_ready->begin());
  h->ws_ready->push_back({d_ws, ws_bytes});
  return LC_OK;
}

extern "C" int lc_ae_workspace_release(lc_handle_t h, void* d_ws) {
  LC_REQUIRE(h, "null handle");
  for (size_t i = 0; i < h->ws_ready->size(); ++i)
    if ((*h->ws_ready)[i].first == d_ws) { h->ws_ready->erase(h->ws_ready->begin() + i); break; }
  return LC_OK;
}

static bool ws_is_ready(const lc_codec* h, const void* d_ws, size_t need) {
  if (h->precision == LC_PREC_FP32) return true;      // the fp32 back end writes every value it reads
  for (auto& e : *h->ws_ready)
    if (e.first == d_ws && e.second >= need) return true;
  return false;
}

// Runs layers [first, first+count) with the fp32 direct back end.  bufs: 3 activation buffers.
static int run_layers_fp32(lc_codec* h, int first, int count, float* cur, float* bufs[3],
                           int n_chunks, float* final_out, float** last, cudaStream_t s) {
  // cur is one of bufs (or external input).  Skip tensor is kept in place.
  const float* skip = nullptr;
  for (int i = first; i < first + count; ++i) {
    DevLayer& L = h->layers[i];
    float* dst = nullptr;
    if (i == first + count - 1 && final_out) {
      dst = final_out;
    } else {
      for (int b = 0; b < 3; ++b)
        if (bufs[b] != cur && bufs[b] != skip) { dst = bufs[b]; break; }
    }
    prof_begin(h, i, s);
    int rc = conv_direct_launch(h, L, cur, dst, L.res_add ? skip : nullptr, n_chunks, s);
    prof_end(h, s);
    if (rc != LC_OK) return rc;
    if (L.res_add) skip = nullptr;
    if (L.res_save) skip = dst;
    cur = dst;
  }
  if (last) *last = cur;
  return LC_OK;
}

extern "C" int lc_chunk_gather(const float* d_x, int T, int H, int W, int C, float* d_chunks, void* stream) {
  LC_REQUIRE(d_x && d_chunks && C >= 1, "bad argument");
  LC_REQUIRE(T >= CT && H >= CH && W >= CW, "Chunks size cannot be larger than the data size.");
  ChunkGeom g = make_geom(T, H, W);
  const int n = g.gt * g.gh * g.gw;
  chunk_gather_kernel<<<kNumSM * 8, 256, 0, as_stream(stream)>>>(d_x, g, C, 0.f, 1.f, 0, 0, n, d_chunks);
  LC_LAUNCH_CHECK();
  return LC_OK;
}

extern "C" int lc_chunk_merge(const float* d_chunks, int T, int H, int W, int C, float* d_out, void* stream) {
  LC_REQUIRE(d_out && d_chunks && C >= 1, "bad argument");
  LC_REQUIRE(T >= CT && H >= CH && W >= CW, "Chunks size cannot be larger than the data size.");
  ChunkGeom g = make_geom(T, H, W);
  const int n = g.gt * g.gh * g.gw;
  chunk_merge_kernel<<<kNumSM * 8, 256, 0, as_stream(stream)>>>(d_chunks, g, C, 0, n, d_out);
  LC_LAUNCH_CHECK();
  return LC_OK;
}

extern "C" int lc_encode(lc_handle_t h, const float* d_x, int T, int H, int W, int C, float mean,
                         float stdv, int chunk_begin, int chunk_count, float* d_latent, void* d_ws,
                         size_t ws_bytes, void* stream) {
  LC_REQUIRE(h && d_x && d_latent && d_ws, "null argument");
  LC_REQUIRE(C == h->in_channels, "channel count does not match the model");
  LC_REQUIRE(T >= CT && H >= CH && W >= CW, "Chunks size cannot be larger than the data size.");
  ChunkGeom g = make_geom(T, H, W);
  LC_REQUIRE(chunk_begin >= 0 && chunk_count > 0 && chunk_begin + chunk_count <= g.gt * g.gh * g.gw,
             "chunk range out of bounds");
  LC_REQUIRE(ws_bytes >= lc_ae_workspace_bytes(h, chunk_count), "workspace too small");
  LC_REQUIRE(ws_is_ready(h, d_ws, lc_ae_workspace_bytes(h, chunk_count)),
             "workspace was not prepared with lc_ae_workspace_init (the 16-bit kernels need zeroed halos)");
  cudaStream_t s = as_stream(stream);
  if (h->precision != LC_PREC_FP32)
    return tc_encode(h, d_x, g, C, mean, stdv, chunk_begin, chunk_count, d_latent, d_ws, s);
  const size_t per = act_floats_per_chunk(h) * (size_t)chunk_count;
  float* base = reinterpret_cast<float*>(d_ws);
  float* bufs[3] = {base, base + per, base + 2 * per};
  const long long total = (long long)chunk_count * CT * CH * CW;
  int grid = (int)((total + 255) / 256 < (long long)h->num_sms * 8 ? (total + 255) / 256
                                                                    : (long long)h->num_sms * 8);
  chunk_gather_kernel<<<grid, 256, 0, s>>>(d_x, g, C, mean, stdv, 1, chunk_begin, chunk_count, bufs[0]);
  LC_LAUNCH_CHECK();
  return run_layers_fp32(h, 0, h->n_enc, bufs[0], bufs, chunk_count, d_latent, nullptr, s);
}

extern "C" int lc_decode(lc_handle_t h, const float* d_latent, int T, int H, int W, int chunk_begin,
                         int chunk_count, float* d_recon, void* d_ws, size_t ws_bytes,
                         void* stream) {
  LC_REQUIRE(h && d_latent && d_recon && d_ws, "null argument");
  LC_REQUIRE(T >= CT && H >= CH && W >= CW, "Chunks size cannot be larger than the data size.");
  ChunkGeom g = make_geom(T, H, W);
  LC_REQUIRE(chunk_begin >= 0 && chunk_count > 0 && chunk_begin + chunk_count <= g.gt * g.gh * g.gw,
             "chunk range out of bounds");
  LC_REQUIRE(ws_bytes >= lc_ae_workspace_bytes(h, chunk_count), "workspace too small");
  LC_REQUIRE(ws_is_ready(h, d_ws, lc_ae_workspace_bytes(h, chunk_count)),
             "workspace was not prepared with lc_ae_workspace_init (the 16-bit kernels need zeroed halos)");
  cudaStream_t s = as_stream(stream);
  if (h->precision != LC_PREC_FP32)
    return tc_decode(h, d_latent, g, chunk_begin, chunk_count, d_recon, d_ws, s);
  const size_t per = act_floats_per_chunk(h) * (size_t)chunk_count;
  float* base = reinterpret_cast<float*>(d_ws);
  float* bufs[3] = {base, base + per, base + 2 * per};
  // the latent is read in place as the first layer's input
  float* cur = nullptr;
  int rc = run_layers_fp32(h, h->n_enc, h->n_dec, const_cast<float*>(d_latent), bufs, chunk_count,
                           nullptr, &cur, s);
  if (rc != LC_OK) return rc;
  const long long total = (long long)chunk_count * CT * CH * CW;
  int grid = (int)((total + 255) / 256 < (long long)h->num_sms * 8 ? (total + 255) / 256
                                                                    : (long long)h->num_sms * 8);
  chunk_merge_kernel<<<grid, 256, 0, s>>>(cur, g, 1, chunk_begin, chunk_count, d_recon);
  LC_LAUNCH_CHECK();
  return LC_OK;
}
