// Tensor-core (tcgen05 / TMEM / bulk-async-copy) back end of the autoencoder, LC_PREC_BF16 and
// LC_PREC_FP16 modes.
//
// Plane-sweep implicit GEMM (stride-1 k=3 convolutions, 80 % of the model's FLOPs):
//   * activations are 16-bit group-plane tensors (act.cuh): a 128-voxel tile of one time plane is a
//     128-row K-major UMMA operand, and the in-plane tap (dh,dw) is the same operand started
//     (dh*PW+dw) vectors later -- no im2col, no swizzle, one 1-D bulk copy per channel group;
//   * the time taps are folded into N: D[voxel, (dt,co)] = sum_{dh,dw,ci} X[t_in][voxel+(dh,dw)][ci] *
//     W[dt][dh][dw][ci][co], so one A tile read from shared memory feeds 3x23 output columns (N=80)
//     instead of 23 -- with N=32 the MMA would be bound by the 128 B/clk shared-memory read of A;
//   * a CTA sweeps t_in = 0..15 for a fixed 128-voxel tile: the dt-block of D belongs to output
//     plane t_in-dt+1, which is the same TMEM lane = the same epilogue thread for every t_in, so
//     the fold is undone by thread-local register accumulation (no cross-lane traffic);
//   * K steps pair channel groups across taps through the descriptor's leading-byte-offset
//     (27 groups of 8 channels -> 14 MMAs per plane instead of 18 with channels padded to 32);
//   * warp-specialised: warp 0 = bulk-copy producer, warp 1 = MMA issuer (one elected thread),
//     warps 2-5 = epilogue (TMEM -> registers -> bias/skip/ReLU -> 16-bit stores); 4-stage smem ring,
//     double-buffered TMEM accumulators.
// Everything is deterministic: fixed item -> CTA mapping is irrelevant to the result because each
// output voxel is produced by exactly one thread in a fixed accumulation order.
#include "act.cuh"
#include "umma.cuh"

namespace {

constexpr int kStages = 4;
constexpr int kTcThreads = 192;
constexpr int kMaxKs = 40;
constexpr int kTmemCols = 256;
constexpr int kAccStride = 128;   // TMEM columns between the two accumulator buffers

struct TcParams {
  const uint4* in;
  uint4* out;
  const uint4* skip;
  long long in_cs, out_cs, skip_cs;
  int in_S, in_GL, in_PW;
  int out_S, out_GL, out_PW;
  int skip_S, skip_GL, skip_PW;
  int G, T, H, W;
  const uint4* wimg;
  int wimg_bytes;
  const float* bias;
  int cout, relu, half;
  int n_items, tiles_per_chunk;
  int nks, N;
  int slab_slots;
  uint32_t a_off[kMaxKs], a_lbo[kMaxKs];
};

__device__ __forceinline__ void ld_block24(uint32_t taddr, float d[24]) {
  uint32_t v[8];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    umma::tmem_ld8(taddr + j * 8, v);
    umma::tmem_ld_wait();
#pragma unroll
    for (int k = 0; k < 8; ++k) d[j * 8 + k] = __uint_as_float(v[k]);
  }
}

// k=3, stride 1, dt folded (3 blocks of 24 columns)
__global__ void __launch_bounds__(kTcThreads, 2) conv3_tc_kernel(const __grid_constant__ TcParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t full[kStages], empty[kStages], tfull[2], tempty[2], wbar;
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_bias[24];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slab_bytes = p.slab_slots * 16;
  const int stage_bytes = p.G * slab_bytes;
  uint8_t* sB = smem;
  uint8_t* sA = smem + ((p.wimg_bytes + 127) & ~127);

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) { umma::mbar_init(&full[i], 1); umma::mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { umma::mbar_init(&tfull[i], 1); umma::mbar_init(&tempty[i], 4); }
    umma::mbar_init(&wbar, 1);
    umma::fence_barrier_init();
  }
  if (threadIdx.x < 24) s_bias[threadIdx.x] = (threadIdx.x < p.cout) ? p.bias[threadIdx.x] : 0.f;
  if (warp == 1) umma::tmem_alloc(&tmem_base_s, kTmemCols);
  umma::tcgen05_fence_before();
  __syncthreads();
  umma::tcgen05_fence_after();
  const uint32_t tmem = tmem_base_s;

  if (warp == 0) {
    // ===================== producer: bulk async copies (TMA engine) =====================
    if (lane == 0) {
      umma::mbar_arrive_expect_tx(&wbar, (uint32_t)p.wimg_bytes);
      umma::bulk_g2s(sB, p.wimg, (uint32_t)p.wimg_bytes, &wbar);
      int stage = 0, phase = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const int n = item / p.tiles_per_chunk, tile = item % p.tiles_per_chunk;
        // tile rows are slots [PW + 128*tile, +128); the slab starts PW+1 slots earlier
        const long long slab0 = (long long)n * p.in_cs + p.in_GL + tile * 128 - 1;
        for (int t = 0; t < p.T; ++t) {
          umma::mbar_wait(&empty[stage], phase ^ 1);
          umma::mbar_arrive_expect_tx(&full[stage], (uint32_t)stage_bytes);
          for (int g = 0; g < p.G; ++g)
            umma::bulk_g2s(sA + stage * stage_bytes + g * slab_bytes,
                           p.in + slab0 + (long long)(t * p.G + g) * p.in_S, (uint32_t)slab_bytes, &full[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: one elected thread =====================
    if (lane == 0) {
      umma::mbar_wait(&wbar, 0);
      const uint32_t idesc = umma::make_idesc(128, p.N, p.half ? 0 : 1);
      const uint32_t sA_u = umma::smem_u32(sA), sB_u = umma::smem_u32(sB);
      const uint32_t kstep_bytes = 2u * p.N * 16u;
      int stage = 0, phase = 0, it = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        for (int t = 0; t < p.T; ++t, ++it) {
          const int buf = it & 1;
          umma::mbar_wait(&tempty[buf], ((it >> 1) & 1) ^ 1);
          umma::mbar_wait(&full[stage], phase);
          umma::tcgen05_fence_after();
          const uint32_t a_base = sA_u + stage * stage_bytes;
          const uint32_t d_addr = tmem + buf * kAccStride;
          for (int ks = 0; ks < p.nks; ++ks) {
            const uint64_t da = umma::make_desc(a_base + p.a_off[ks], p.a_lbo[ks], 128);
            const uint64_t db = umma::make_desc(sB_u + ks * kstep_bytes, p.N * 16u, 128);
            umma::mma_f16(d_addr, da, db, idesc, ks > 0);
          }
          umma::commit(&empty[stage]);
          umma::commit(&tfull[buf]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ===================== epilogue: TMEM -> registers -> global =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
    int it = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      const int n = item / p.tiles_per_chunk, tile = item % p.tiles_per_chunk;
      const int slot = p.in_PW + tile * 128 + row;
      const int hp = slot / p.in_PW, wp = slot - hp * p.in_PW;
      const bool valid = (hp >= 1) && (hp <= p.H) && (wp >= 1) && (wp <= p.W);
      const long long o_base = (long long)n * p.out_cs + p.out_GL + hp * p.out_PW + wp;
      const long long s_base = (long long)n * p.skip_cs + p.skip_GL + hp * p.skip_PW + wp;
      float accP[24], accC[24], accN[24];
#pragma unroll
      for (int c = 0; c < 24; ++c) { accP[c] = 0.f; accC[c] = 0.f; }

      auto finalize = [&](int t, const float (&acc)[24]) {
        if (!valid) return;
        float y[24];
#pragma unroll
        for (int c = 0; c < 24; ++c) y[c] = acc[c] + s_bias[c];
        if (p.skip) {
#pragma unroll
          for (int g = 0; g < 3; ++g) {
            float f[8];
            unpack8(__ldg(p.skip + s_base + (long long)(t * p.G + g) * p.skip_S), p.half, f);
#pragma unroll
            for (int k = 0; k < 8; ++k) y[g * 8 + k] = __fadd_rn(y[g * 8 + k], f[k]);
          }
        }
#pragma unroll
        for (int g = 0; g < 3; ++g) {
          float f[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            float v = y[g * 8 + k];
            if (p.relu) v = fmaxf(v, 0.f);
            f[k] = (g * 8 + k < p.cout) ? v : 0.f;
          }
          p.out[o_base + (long long)(t * p.G + g) * p.out_S] = pack8(f, p.half);
        }
      };

      for (int t = 0; t < p.T; ++t, ++it) {
        const int buf = it & 1;
        umma::mbar_wait(&tfull[buf], (it >> 1) & 1);
        umma::tcgen05_fence_after();
        const uint32_t ta = lane_addr + buf * kAccStride;
        float d[24];
        ld_block24(ta, d);          // dt = 0 -> output plane t+1
#pragma unroll
        for (int c = 0; c < 24; ++c) accN[c] = d[c];
        ld_block24(ta + 24, d);     // dt = 1 -> output plane t
#pragma unroll
        for (int c = 0; c < 24; ++c) accC[c] += d[c];
        ld_block24(ta + 48, d);     // dt = 2 -> output plane t-1
#pragma unroll
        for (int c = 0; c < 24; ++c) accP[c] += d[c];
        umma::tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) umma::mbar_arrive(&tempty[buf]);
        if (t >= 1) finalize(t - 1, accP);
#pragma unroll
        for (int c = 0; c < 24; ++c) { accP[c] = accC[c]; accC[c] = accN[c]; }
      }
      finalize(p.T - 1, accP);
    }
  }
  umma::tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) umma::tmem_dealloc(tmem, kTmemCols);
}

// ------------------------------------------------------------------------------------------
// plan: tensors of one chunk block
// ------------------------------------------------------------------------------------------
struct TensorPlan {
  int fmt;                 // FMT_F32 / FMT_GP16
  int T, H, W, C, G, NP, PW, PH, S, GL;
  long long offset_vecs;   // offset inside the chunk block, in 16-byte vectors
  long long size_vecs;
};

struct LayerPlan {
  int backend;             // 0 = direct, 1 = conv3_tc
  float* d_w16;            // fp32 weights rounded through the 16-bit operand format [tap][cin][cout]
  TcParams tc;             // template parameters (pointers filled per call)
};

struct TcPlan {
  int n_layers;
  TensorPlan in_enc;                    // fp32 chunk input
  TensorPlan out[kMaxLayers];           // output tensor of every layer
  LayerPlan layer[kMaxLayers];
  long long block_vecs;                 // 16-byte vectors per chunk block
};

TensorPlan make_gp(int T, int H, int W, int C, int pad_hi) {
  TensorPlan t;
  memset(&t, 0, sizeof(t));
  t.fmt = FMT_GP16;
  t.T = T; t.H = H; t.W = W; t.C = C;
  t.G = (C + 7) / 8;
  t.NP = 1;
  t.PW = W + 1 + pad_hi;
  t.PH = H + 1 + pad_hi;
  t.GL = 8;
  // guard: a tile slab may over-read up to 128 + 3*(PW+1) vectors past the last interior line
  t.S = t.GL + t.PH * t.PW + 128 + 3 * (t.PW + 1);
  t.S = (t.S + 7) & ~7;
  t.size_vecs = (long long)T * t.G * t.S;
  return t;
}

TensorPlan make_f32(int T, int H, int W, int C) {
  TensorPlan t;
  memset(&t, 0, sizeof(t));
  t.fmt = FMT_F32;
  t.T = T; t.H = H; t.W = W; t.C = C;
  t.size_vecs = ((long long)T * H * W * C * 4 + 15) / 16;
  return t;
}

ActView view_of(const TensorPlan& t, void* block_base, long long block_vecs, int half) {
  ActView v;
  memset(&v, 0, sizeof(v));
  v.base = reinterpret_cast<uint4*>(block_base) + t.offset_vecs;
  v.fmt = t.fmt;
  v.half = half;
  v.T = t.T; v.H = t.H; v.W = t.W; v.C = t.C;
  v.G = t.G; v.NP = t.NP; v.PW = t.PW; v.S = t.S; v.GL = t.GL;
  v.chunk_stride = (t.fmt == FMT_F32) ? block_vecs * 4 : block_vecs;   // floats or vectors
  return v;
}

TcPlan* plan_of(const lc_codec* h) { return reinterpret_cast<TcPlan*>(h->tc_plan); }

bool tc_conv3_eligible(const DevLayer& L) {
  return L.kind == 0 && L.k == 3 && L.stride == 1 && L.cin == L.cout && L.cin > 16 && L.cin <= 24 &&
         L.in_t == CT && L.in_h == CH && L.in_w == CW;
}

// merge of the decoder's fp32 chunk output (dataLoader.py:484-532), strided chunk blocks
__global__ void merge_blocks_kernel(const float* __restrict__ dec, long long chunk_stride, ChunkGeom g,
                                    int chunk_begin, int n_chunks, float* __restrict__ recon) {
  const long long total = (long long)n_chunks * CT * CH * CW;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < total;
       v += (long long)gridDim.x * blockDim.x) {
    const int vox = (int)(v % (CT * CH * CW));
    const int chunk = (int)(v / (CT * CH * CW));
    const int lw = vox % CW, lh = (vox / CW) % CH, lt = vox / (CW * CH);
    int t0, h0, w0, kt, kh, kw;
    g.origin(chunk_begin + chunk, t0, h0, w0);
    g.keep_from(chunk_begin + chunk, kt, kh, kw);
    if (lt < kt || lh < kh || lw < kw) continue;
    recon[((size_t)(t0 + lt) * g.H + (h0 + lh)) * g.W + (w0 + lw)] = dec[(size_t)chunk * chunk_stride + vox];
  }
}

__global__ void gather_blocks_kernel(const float* __restrict__ x, ChunkGeom g, int C, float mean, float stdv,
                                     int chunk_begin, int n_chunks, float* __restrict__ out,
                                     long long chunk_stride) {
  const long long total = (long long)n_chunks * CT * CH * CW;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < total;
       v += (long long)gridDim.x * blockDim.x) {
    const int vox = (int)(v % (CT * CH * CW));
    const int chunk = (int)(v / (CT * CH * CW));
    const int lw = vox % CW, lh = (vox / CW) % CH, lt = vox / (CW * CH);
    int t0, h0, w0;
    g.origin(chunk_begin + chunk, t0, h0, w0);
    const float* src = x + (((size_t)(t0 + lt) * g.H + (h0 + lh)) * g.W + (w0 + lw)) * C;
    float* dst = out + (size_t)chunk * chunk_stride + (size_t)vox * C;
    dst[0] = __fdiv_rn(__fsub_rn(src[0], mean), stdv);
    for (int c = 1; c < C; ++c) dst[c] = src[c];
  }
}

int launch_conv3_tc(const lc_codec* h, const LayerPlan& lp, const TensorPlan& tin, const TensorPlan& tout,
                    const TensorPlan* tskip, void* block_base, long long block_vecs, int n_chunks,
                    cudaStream_t s) {
  TcParams p = lp.tc;
  uint4* base = reinterpret_cast<uint4*>(block_base);
  p.in = base + tin.offset_vecs;
  p.out = base + tout.offset_vecs;
  p.skip = tskip ? base + tskip->offset_vecs : nullptr;
  if (tskip) { p.skip_S = tskip->S; p.skip_GL = tskip->GL; p.skip_PW = tskip->PW; }
  p.in_cs = p.out_cs = p.skip_cs = block_vecs;
  p.n_items = n_chunks * p.tiles_per_chunk;
  const int stage_bytes = p.G * p.slab_slots * 16;
  const int smem = ((p.wimg_bytes + 127) & ~127) + kStages * stage_bytes + 128;
  static bool attr_done = false;
  if (!attr_done) {
    LC_CUDA(cudaFuncSetAttribute(conv3_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    attr_done = true;
  }
  if (smem > 100 * 1024) {
    lc_set_error("conv3_tc: %d bytes of shared memory needed", smem);
    return LC_EUNSUPPORTED;
  }
  int grid = 2 * h->num_sms;
  if (grid > p.n_items) grid = p.n_items;
  conv3_tc_kernel<<<grid, kTcThreads, smem, s>>>(p);
  LC_LAUNCH_CHECK();
  return LC_OK;
}

}  // namespace

int tc_prepare(lc_codec* h) {
  h->tc_plan = nullptr;
  if (h->precision == LC_PREC_FP32) return LC_OK;
  const int half = (h->precision == LC_PREC_FP16) ? 1 : 0;
  TcPlan* P = new TcPlan();
  memset(P, 0, sizeof(*P));
  const int nl = h->n_enc + h->n_dec;
  P->n_layers = nl;
  long long off = 0;
  auto place = [&](TensorPlan& t) {
    t.offset_vecs = off;
    off += (t.size_vecs + 7) & ~7LL;
  };
  P->in_enc = make_f32(CT, CH, CW, h->in_channels);
  place(P->in_enc);
  for (int i = 0; i < nl; ++i) {
    const DevLayer& L = h->layers[i];
    const bool last_enc = (i == h->n_enc - 1), last_dec = (i == nl - 1);
    if (last_enc || last_dec) {
      P->out[i] = make_f32(L.out_t, L.out_h, L.out_w, L.cout);   // latent / decoder output stay fp32
    } else {
      const DevLayer& Nx = h->layers[i + 1];
      const int pad_hi = (Nx.kind == 0 && Nx.stride == 1) ? (Nx.k - 1 - Nx.pad_lo) : 1;
      P->out[i] = make_gp(L.out_t, L.out_h, L.out_w, L.cout, pad_hi < 1 ? 1 : pad_hi);
    }
    if (!last_enc) place(P->out[i]);   // the latent goes straight to the caller's buffer
  }
  P->block_vecs = (off + 63) & ~63LL;

  for (int i = 0; i < nl; ++i) {
    DevLayer& L = h->layers[i];
    LayerPlan& lp = P->layer[i];
    const int taps = L.k * L.k * L.k;
    std::vector<float> w((size_t)taps * L.cin * L.cout), w16(w.size());
    LC_CUDA(cudaMemcpy(w.data(), L.d_w, w.size() * sizeof(float), cudaMemcpyDeviceToHost));
    for (size_t j = 0; j < w.size(); ++j) w16[j] = round_h16_host(w[j], half);
    LC_CUDA(cudaMalloc(&lp.d_w16, w16.size() * sizeof(float)));
    LC_CUDA(cudaMemcpy(lp.d_w16, w16.data(), w16.size() * sizeof(float), cudaMemcpyHostToDevice));
    lp.backend = 0;
    const bool in_is_gp = (i != 0) && (i != h->n_enc);
    const bool out_is_gp = P->out[i].fmt == FMT_GP16;
    if (tc_conv3_eligible(L) && in_is_gp && out_is_gp && !getenv("LOSSYCOMP_NO_TC")) {
      const TensorPlan& tin = P->out[i - 1];
      const TensorPlan& tout = P->out[i];
      TcParams& p = lp.tc;
      p.in_S = tin.S; p.in_GL = tin.GL; p.in_PW = tin.PW;
      p.out_S = tout.S; p.out_GL = tout.GL; p.out_PW = tout.PW;
      p.G = tin.G; p.T = L.in_t; p.H = L.in_h; p.W = L.in_w;
      p.cout = L.cout; p.relu = L.relu; p.half = half;
      p.bias = L.d_b;
      p.N = 80;
      p.tiles_per_chunk = (L.in_h * tin.PW + 127) / 128;
      p.slab_slots = 128 + 2 * (tin.PW + 1);
      // K groups, group-major so that operand addresses increase: (g, dh, dw)
      const int ngroups = p.G * 9;
      p.nks = (ngroups + 1) / 2;
      auto addr = [&](int gi) {
        const int g = gi / 9, tap = gi % 9, dh = tap / 3, dw = tap % 3;
        return (uint32_t)(g * p.slab_slots * 16 + (dh * tin.PW + dw) * 16);
      };
      for (int ks = 0; ks < p.nks; ++ks) {
        const int g0 = 2 * ks, g1 = (2 * ks + 1 < ngroups) ? 2 * ks + 1 : 2 * ks;
        p.a_off[ks] = addr(g0);
        p.a_lbo[ks] = addr(g1) - addr(g0);
      }
      // weight image [ks][khalf][n][8]; n = dt*24 + co
      std::vector<unsigned short> img((size_t)p.nks * 2 * p.N * 8, 0);
      for (int gi = 0; gi < ngroups; ++gi) {
        const int g = gi / 9, tap = gi % 9, dh = tap / 3, dw = tap % 3;
        for (int n = 0; n < p.N; ++n) {
          const int dt = n / 24, co = n % 24;
          if (dt >= 3 || co >= L.cout) continue;
          for (int k8 = 0; k8 < 8; ++k8) {
            const int ci = g * 8 + k8;
            if (ci >= L.cin) continue;
            const float v = w16[((size_t)((dt * 3 + dh) * 3 + dw) * L.cin + ci) * L.cout + co];
            unsigned short u;
            if (half) { __half hv = __float2half_rn(v); memcpy(&u, &hv, 2); }
            else { __nv_bfloat16 bv = __float2bfloat16_rn(v); memcpy(&u, &bv, 2); }
            img[((size_t)gi * p.N + n) * 8 + k8] = u;
          }
        }
      }
      p.wimg_bytes = (int)(img.size() * 2);
      LC_CUDA(cudaMalloc(&L.d_wtc, img.size() * 2));
      LC_CUDA(cudaMemcpy(L.d_wtc, img.data(), img.size() * 2, cudaMemcpyHostToDevice));
      L.wtc_bytes = img.size() * 2;
      p.wimg = reinterpret_cast<const uint4*>(L.d_wtc);
      lp.backend = 1;
    }
  }
  h->tc_plan = P;
  return LC_OK;
}

void tc_release(lc_codec* h) {
  TcPlan* P = plan_of(h);
  if (!P) return;
  for (int i = 0; i < P->n_layers; ++i)
    if (P->layer[i].d_w16) cudaFree(P->layer[i].d_w16);
  delete P;
  h->tc_plan = nullptr;
}

size_t tc_workspace_bytes(const lc_codec* h, int max_chunks) {
  const TcPlan* P = plan_of(h);
  if (!P) return 0;
  return (size_t)P->block_vecs * 16 * (size_t)max_chunks + 4096;
}

int tc_backend_of(const lc_codec* h, int layer) {
  const TcPlan* P = plan_of(h);
  return P ? P->layer[layer].backend : 0;
}

static int tc_run(lc_codec* h, int first, int count, const ActView& first_in, void* block_base, int n_chunks,
                  float* latent_out, cudaStream_t s) {
  TcPlan* P = plan_of(h);
  const int half = (h->precision == LC_PREC_FP16) ? 1 : 0;
  ActView cur = first_in;
  const TensorPlan* cur_t = nullptr;
  const TensorPlan* skip_t = nullptr;
  ActView skip_v;
  for (int i = first; i < first + count; ++i) {
    DevLayer& L = h->layers[i];
    LayerPlan& lp = P->layer[i];
    ActView out;
    if (i == h->n_enc - 1) {   // latent: contiguous fp32 in the caller's buffer
      out = f32_view(L, true, latent_out);
    } else {
      out = view_of(P->out[i], block_base, P->block_vecs, half);
    }
    prof_begin(h, i, s);
    int rc;
    if (lp.backend == 1) {
      rc = launch_conv3_tc(h, lp, *cur_t, P->out[i], L.res_add ? skip_t : nullptr, block_base, P->block_vecs,
                           n_chunks, s);
    } else {
      rc = conv_direct_views(h, L, lp.d_w16, cur, out, L.res_add ? &skip_v : nullptr, n_chunks, s);
    }
    prof_end(h, s);
    if (rc != LC_OK) return rc;
    if (L.res_save) { skip_v = out; skip_t = &P->out[i]; }
    cur = out;
    cur_t = &P->out[i];
  }
  return LC_OK;
}

int tc_encode(lc_codec* h, const float* d_x, const ChunkGeom& g, int C, float mean, float stdv, int chunk_begin,
              int chunk_count, float* d_latent, void* d_ws, cudaStream_t s) {
  TcPlan* P = plan_of(h);
  const int half = (h->precision == LC_PREC_FP16) ? 1 : 0;
  ActView in = view_of(P->in_enc, d_ws, P->block_vecs, half);
  gather_blocks_kernel<<<h->num_sms * 8, 256, 0, s>>>(d_x, g, C, mean, stdv, chunk_begin, chunk_count,
                                                      reinterpret_cast<float*>(in.base), in.chunk_stride);
  LC_LAUNCH_CHECK();
  return tc_run(h, 0, h->n_enc, in, d_ws, chunk_count, d_latent, s);
}

int tc_decode(lc_codec* h, const float* d_latent, const ChunkGeom& g, int chunk_begin, int chunk_count,
              float* d_recon, void* d_ws, cudaStream_t s) {
  TcPlan* P = plan_of(h);
  const int half = (h->precision == LC_PREC_FP16) ? 1 : 0;
  ActView in = f32_view(h->layers[h->n_enc], false, d_latent);
  int rc = tc_run(h, h->n_enc, h->n_dec, in, d_ws, chunk_count, nullptr, s);
  if (rc != LC_OK) return rc;
  ActView out = view_of(P->out[h->n_enc + h->n_dec - 1], d_ws, P->block_vecs, half);
  merge_blocks_kernel<<<h->num_sms * 8, 256, 0, s>>>(reinterpret_cast<const float*>(out.base), out.chunk_stride,
                                                     g, chunk_begin, chunk_count, d_recon);
  LC_LAUNCH_CHECK();
  return LC_OK;
}
