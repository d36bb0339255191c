// Shared declarations of the lossycomp_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <utility>
#include <vector>

#include "../../include/lossycomp_b200.h"

void lc_set_error(const char* fmt, ...);

#define LC_CUDA(call)                                                                      \
  do {                                                                                     \
    cudaError_t e_ = (call);                                                               \
    if (e_ != cudaSuccess) {                                                               \
      lc_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));   \
      return LC_ECUDA;                                                                     \
    }                                                                                      \
  } while (0)

#define LC_REQUIRE(cond, msg)                                   \
  do {                                                          \
    if (!(cond)) {                                              \
      lc_set_error("%s:%d %s", __FILE__, __LINE__, msg);        \
      return LC_EINVAL;                                         \
    }                                                           \
  } while (0)

#define LC_LAUNCH_CHECK() LC_CUDA(cudaGetLastError())

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

constexpr int CT = LC_CHUNK_T, CH = LC_CHUNK_H, CW = LC_CHUNK_W;
constexpr int kMaxLayers = 32;
constexpr int kNumSM = 148;

// One layer as the device code sees it.
struct DevLayer {
  int kind, k, stride, cin, cout, relu, res_save, res_add;
  int in_t, in_h, in_w;      // input spatial extent
  int out_t, out_h, out_w;   // output spatial extent
  int pad_lo;                // Conv3D: "same" padding before; Conv3DTranspose: crop at the front
  float* d_w;                // fp32 weights re-laid as [tap][cin][cout] (both kinds)
  float* d_b;                // fp32 bias [cout]
  void* d_wtc;               // tensor-core weights (16-bit, UMMA core-matrix order), or null
  size_t wtc_bytes;
};

struct ProfSpan { int layer; cudaEvent_t a, b; };

struct lc_codec {
  int profile;                       // record a CUDA event pair around every conv launch
  std::vector<ProfSpan>* spans;      // pending (unread) spans
  std::vector<cudaEvent_t>* pool;    // recycled events
  int device;
  int precision;
  int in_channels;
  int n_enc, n_dec;
  int latent_floats;
  int max_c;                 // widest activation in channels
  DevLayer layers[kMaxLayers];
  int num_sms;
  void* tc_plan;             // TcPlan* (conv_tc.cu), null in LC_PREC_FP32 mode
  std::vector<std::pair<void*, size_t>>* ws_ready;   // workspaces zeroed by lc_ae_workspace_init
};

// chunk index -> origin in the field (dataLoader.py:457-481): lon fastest, tail re-anchored.
struct ChunkGeom {
  int T, H, W;
  int gt, gh, gw;
  __host__ __device__ inline void origin(int idx, int& t0, int& h0, int& w0) const {
    int iw = idx % gw;
    int ih = (idx / gw) % gh;
    int it = idx / (gw * gh);
    t0 = min(it * CT, T - CT);
    h0 = min(ih * CH, H - CH);
    w0 = min(iw * CW, W - CW);
  }
  // first local index (per axis) of the part merge_data keeps (dataLoader.py:503-532)
  __host__ __device__ inline void keep_from(int idx, int& kt, int& kh, int& kw) const {
    int iw = idx % gw;
    int ih = (idx / gw) % gh;
    int it = idx / (gw * gh);
    kt = it * CT - min(it * CT, T - CT);
    kh = ih * CH - min(ih * CH, H - CH);
    kw = iw * CW - min(iw * CW, W - CW);
  }
};

static inline ChunkGeom make_geom(int T, int H, int W) {
  ChunkGeom g;
  g.T = T; g.H = H; g.W = W;
  g.gt = (T + CT - 1) / CT; g.gh = (H + CH - 1) / CH; g.gw = (W + CW - 1) / CW;
  return g;
}

// profiling helpers (api.cu)
void prof_begin(lc_codec* h, int layer, cudaStream_t s);
void prof_end(lc_codec* h, cudaStream_t s);

// ---- conv back ends ---------------------------------------------------------------------
// fp32 CUDA-core direct convolution over NDHWC fp32 activations (conv_direct.cu)
int conv_direct_launch(const lc_codec* h, const DevLayer& L, const float* d_in, float* d_out,
                       const float* d_skip, int n_chunks, cudaStream_t s);

// tcgen05 implicit-GEMM back end (conv_tc.cu)
int tc_prepare(lc_codec* h);
void tc_release(lc_codec* h);
int tc_backend_of(const lc_codec* h, int layer);
size_t tc_workspace_bytes(const lc_codec* h, int max_chunks);
int tc_encode(lc_codec* h, const float* d_x, const ChunkGeom& g, int C, float mean, float stdv,
              int chunk_begin, int chunk_count, float* d_latent, void* d_ws, cudaStream_t s);
int tc_decode(lc_codec* h, const float* d_latent, const ChunkGeom& g, int chunk_begin, int chunk_count,
              float* d_recon, void* d_ws, cudaStream_t s);

struct ActView;
ActView f32_view(const DevLayer& L, bool output, const float* base);
int conv_direct_views(const lc_codec* h, const DevLayer& L, const float* d_w, const ActView& in,
                      const ActView& out, const ActView* skip, int n_chunks, cudaStream_t s);
