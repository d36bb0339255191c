// fp32 CUDA-core direct Conv3D / Conv3DTranspose.
//
// This is the LC_PREC_FP32 back end (reference-grade numerics: fp32 FMA, fixed tap order) and,
// in the tensor-core modes, the back end of the layers that have no tcgen05 kernel (the tiny
// inner levels of the autoencoder); there it reads and writes the same 16-bit group-plane
// activations as the tensor-core kernels (act.cuh) and uses weights rounded through the same
// 16-bit format, so a layer computes the same function whichever back end runs it.
//
// Semantics follow Keras (models.py:47-97; SURVEY.md §8c a-M): cross-correlation, "same" padding
// with before = total/2, Conv3DTranspose = full transposed convolution cropped by (k-s)/2.
// One thread owns one output voxel and all output channels; weights live in shared memory as
// [tap][cin][cout padded to 4] so that every FMA quad is fed by one broadcast LDS.128.
#include "act.cuh"

namespace {

struct ConvParams {
  ActView in, out, skip;   // skip.base == nullptr when there is no residual add
  const float* w;          // [taps][cin][cout]
  const float* b;
  int kind, k, stride, cin, cout, relu, pad_lo;
  long long total_vox;     // n_chunks * out voxels per chunk
};

template <int CO4>
__device__ __forceinline__ void fma_row(float (&acc)[CO4 * 4], const float x, const float4* __restrict__ wrow) {
#pragma unroll
  for (int c4 = 0; c4 < CO4; ++c4) {
    const float4 w = wrow[c4];
    acc[c4 * 4 + 0] = fmaf(x, w.x, acc[c4 * 4 + 0]);
    acc[c4 * 4 + 1] = fmaf(x, w.y, acc[c4 * 4 + 1]);
    acc[c4 * 4 + 2] = fmaf(x, w.z, acc[c4 * 4 + 2]);
    acc[c4 * 4 + 3] = fmaf(x, w.w, acc[c4 * 4 + 3]);
  }
}

template <int CO4>
__global__ void __launch_bounds__(512, 1) conv_direct_kernel(ConvParams p) {
  extern __shared__ float4 smem_w4[];
  float* smem_w = reinterpret_cast<float*>(smem_w4);
  const int taps = p.k * p.k * p.k;
  constexpr int cp = CO4 * 4;
  for (int i = threadIdx.x; i < taps * p.cin * cp; i += blockDim.x) {
    const int co = i % cp;
    const int r = i / cp;
    smem_w[i] = (co < p.cout) ? p.w[(size_t)r * p.cout + co] : 0.f;
  }
  __syncthreads();

  const int OT = p.out.T, OH = p.out.H, OW = p.out.W;
  const int IT = p.in.T, IH = p.in.H, IW = p.in.W;
  const int ovox = OT * OH * OW;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < p.total_vox;
       v += (long long)gridDim.x * blockDim.x) {
    const int chunk = (int)(v / ovox);
    int r = (int)(v % ovox);
    const int ow = r % OW;
    r /= OW;
    const int oh = r % OH;
    const int ot = r / OH;

    float acc[cp];
#pragma unroll
    for (int c = 0; c < cp; ++c) acc[c] = (c < p.cout) ? p.b[c] : 0.f;

    for (int kt = 0; kt < p.k; ++kt) {
      int it;
      if (p.kind == 0) {
        it = ot * p.stride + kt - p.pad_lo;
      } else {
        const int num = ot + p.pad_lo - kt;
        if (num < 0 || (num % p.stride) != 0) continue;
        it = num / p.stride;
      }
      if (it < 0 || it >= IT) continue;
      for (int kh = 0; kh < p.k; ++kh) {
        int ih;
        if (p.kind == 0) {
          ih = oh * p.stride + kh - p.pad_lo;
        } else {
          const int num = oh + p.pad_lo - kh;
          if (num < 0 || (num % p.stride) != 0) continue;
          ih = num / p.stride;
        }
        if (ih < 0 || ih >= IH) continue;
        for (int kw = 0; kw < p.k; ++kw) {
          int iw;
          if (p.kind == 0) {
            iw = ow * p.stride + kw - p.pad_lo;
          } else {
            const int num = ow + p.pad_lo - kw;
            if (num < 0 || (num % p.stride) != 0) continue;
            iw = num / p.stride;
          }
          if (iw < 0 || iw >= IW) continue;
          const float4* wt = smem_w4 + (size_t)((kt * p.k + kh) * p.k + kw) * p.cin * CO4;
          if (p.in.fmt == FMT_F32) {
            const float* xin = reinterpret_cast<const float*>(p.in.base) + (size_t)chunk * p.in.chunk_stride +
                               ((size_t)(it * IH + ih) * IW + iw) * p.cin;
            for (int ci = 0; ci < p.cin; ++ci) fma_row<CO4>(acc, __ldg(xin + ci), wt + ci * CO4);
          } else {
            const uint4* xin = reinterpret_cast<const uint4*>(p.in.base);
            for (int g = 0; g < p.in.G; ++g) {
              float f[8];
              unpack8(__ldg(xin + gp_vec_index(p.in, chunk, it, ih, iw, g)), p.in.half, f);
#pragma unroll
              for (int k8 = 0; k8 < 8; ++k8) {
                const int ci = g * 8 + k8;
                if (ci < p.cin) fma_row<CO4>(acc, f[k8], wt + ci * CO4);
              }
            }
          }
        }
      }
    }
    // ---- epilogue: (+ skip) -> ReLU -> store
    if (p.skip.base) {
      if (p.skip.fmt == FMT_F32) {
        const float* sk = reinterpret_cast<const float*>(p.skip.base) + (size_t)chunk * p.skip.chunk_stride +
                          (size_t)(v % ovox) * p.cout;
#pragma unroll
        for (int c = 0; c < cp; ++c)
          if (c < p.cout) acc[c] = __fadd_rn(acc[c], sk[c]);
      } else {
        const uint4* sk = reinterpret_cast<const uint4*>(p.skip.base);
#pragma unroll
        for (int g = 0; g < (cp + 7) / 8; ++g) {
          if (g < p.skip.G) {
            float f[8];
            unpack8(__ldg(sk + gp_vec_index(p.skip, chunk, ot, oh, ow, g)), p.skip.half, f);
#pragma unroll
            for (int k8 = 0; k8 < 8; ++k8) {
              const int c = g * 8 + k8;
              if (c < cp) {
                if (c < p.cout) acc[c] = __fadd_rn(acc[c], f[k8]);
              }
            }
          }
        }
      }
    }
    if (p.relu) {
#pragma unroll
      for (int c = 0; c < cp; ++c) acc[c] = fmaxf(acc[c], 0.f);
    }
    if (p.out.fmt == FMT_F32) {
      float* o = reinterpret_cast<float*>(p.out.base) + (size_t)chunk * p.out.chunk_stride +
                 (size_t)(v % ovox) * p.cout;
#pragma unroll
      for (int c = 0; c < cp; ++c)
        if (c < p.cout) o[c] = acc[c];
    } else {
      uint4* o = reinterpret_cast<uint4*>(p.out.base);
#pragma unroll
      for (int g = 0; g < (cp + 7) / 8; ++g) {
        if (g < p.out.G) {
          float f[8];
#pragma unroll
          for (int k8 = 0; k8 < 8; ++k8) {
            const int c = g * 8 + k8;
            f[k8] = (c < cp && c < p.cout) ? acc[c < cp ? c : 0] : 0.f;
          }
          o[gp_vec_index(p.out, chunk, ot, oh, ow, g)] = pack8(f, p.out.half);
        }
      }
    }
  }
}

// Small layers (the innermost levels: <= 128 output voxels per chunk): one warp per output voxel,
// lane = INPUT channel.  Every lane accumulates its channel's contribution to all output channels over
// all taps (independent weight loads -> plenty of memory-level parallelism, unlike a serial k loop),
// then the 32 partial sums of each output channel are added by a fixed butterfly.  Deterministic;
// the summation order differs from conv_direct_kernel, but a layer always runs on the same kernel.
template <int CO>
__global__ void __launch_bounds__(256) conv_small_kernel(ConvParams p) {
  const int lane = threadIdx.x & 31;
  const int OT = p.out.T, OH = p.out.H, OW = p.out.W;
  const int IT = p.in.T, IH = p.in.H, IW = p.in.W;
  const int ovox = OT * OH * OW;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int ci = lane;
  for (long long v = warp0; v < p.total_vox; v += nwarps) {
    const int chunk = (int)(v / ovox);
    int r = (int)(v % ovox);
    const int ow = r % OW;
    r /= OW;
    const int oh = r % OH;
    const int ot = r / OH;
    float acc[CO];
#pragma unroll
    for (int c = 0; c < CO; ++c) acc[c] = 0.f;
    for (int kt = 0; kt < p.k; ++kt) {
      int it;
      if (p.kind == 0) {
        it = ot * p.stride + kt - p.pad_lo;
      } else {
        const int num = ot + p.pad_lo - kt;
        if (num < 0 || (num % p.stride) != 0) continue;
        it = num / p.stride;
      }
      if (it < 0 || it >= IT) continue;
      for (int kh = 0; kh < p.k; ++kh) {
        int ih;
        if (p.kind == 0) {
          ih = oh * p.stride + kh - p.pad_lo;
        } else {
          const int num = oh + p.pad_lo - kh;
          if (num < 0 || (num % p.stride) != 0) continue;
          ih = num / p.stride;
        }
        if (ih < 0 || ih >= IH) continue;
        for (int kw = 0; kw < p.k; ++kw) {
          int iw;
          if (p.kind == 0) {
            iw = ow * p.stride + kw - p.pad_lo;
          } else {
            const int num = ow + p.pad_lo - kw;
            if (num < 0 || (num % p.stride) != 0) continue;
            iw = num / p.stride;
          }
          if (iw < 0 || iw >= IW) continue;
          float x = 0.f;
          if (ci < p.cin) {
            if (p.in.fmt == FMT_F32) {
              x = __ldg(reinterpret_cast<const float*>(p.in.base) + (size_t)chunk * p.in.chunk_stride +
                        ((size_t)(it * IH + ih) * IW + iw) * p.cin + ci);
            } else {
              const unsigned short* xp = reinterpret_cast<const unsigned short*>(
                  reinterpret_cast<const uint4*>(p.in.base) + gp_vec_index(p.in, chunk, it, ih, iw, ci >> 3));
              x = h16_to_float(xp[ci & 7], p.in.half);
            }
            const float* wt = p.w + ((size_t)((kt * p.k + kh) * p.k + kw) * p.cin + ci) * p.cout;
#pragma unroll
            for (int c = 0; c < CO; ++c)
              if (c < p.cout) acc[c] = fmaf(x, __ldg(wt + c), acc[c]);
          }
        }
      }
    }
    // butterfly over the input channels; afterwards every lane holds all sums
#pragma unroll
    for (int c = 0; c < CO; ++c) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
    }
    // epilogue: lane = output channel
    float y = 0.f;
#pragma unroll
    for (int c = 0; c < CO; ++c)
      if (c == lane) y = acc[c];
    const int co = lane;
    if (co < p.cout) y += p.b[co];
    if (p.skip.base) {
      float sk = 0.f;
      if (p.skip.fmt == FMT_F32) {
        if (co < p.cout)
          sk = reinterpret_cast<const float*>(p.skip.base)[(size_t)chunk * p.skip.chunk_stride +
                                                           (size_t)(v % ovox) * p.cout + co];
      } else if (co < p.skip.G * 8) {
        const unsigned short* sp = reinterpret_cast<const unsigned short*>(
            reinterpret_cast<const uint4*>(p.skip.base) + gp_vec_index(p.skip, chunk, ot, oh, ow, co >> 3));
        sk = h16_to_float(sp[co & 7], p.skip.half);
      }
      y = __fadd_rn(y, sk);
    }
    if (p.relu) y = fmaxf(y, 0.f);
    if (p.out.fmt == FMT_F32) {
      if (co < p.cout)
        reinterpret_cast<float*>(p.out.base)[(size_t)chunk * p.out.chunk_stride + (size_t)(v % ovox) * p.cout + co] = y;
    } else if (co < p.out.G * 8) {
      unsigned short* op = reinterpret_cast<unsigned short*>(
          reinterpret_cast<uint4*>(p.out.base) + gp_vec_index(p.out, chunk, ot, oh, ow, co >> 3));
      op[co & 7] = float_to_h16(co < p.cout ? y : 0.f, p.out.half);
    }
  }
}

template <int CO4>
int launch(const ConvParams& p, int num_sms, cudaStream_t s) {
  const size_t smem = (size_t)p.k * p.k * p.k * p.cin * CO4 * 4 * sizeof(float);
  if (smem > 227 * 1024) {
    lc_set_error("conv_direct: weights (%zu B) exceed shared memory", smem);
    return LC_EUNSUPPORTED;
  }
  LC_CUDA(cudaFuncSetAttribute(conv_direct_kernel<CO4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)smem));
  long long blocks_needed = (p.total_vox + 511) / 512;
  int grid = (int)(blocks_needed < num_sms ? blocks_needed : num_sms);
  if (grid < 1) grid = 1;
  conv_direct_kernel<CO4><<<grid, 512, smem, s>>>(p);
  LC_LAUNCH_CHECK();
  return LC_OK;
}

}  // namespace

ActView f32_view(const DevLayer& L, bool output, const float* base) {
  ActView v;
  memset(&v, 0, sizeof(v));
  v.base = const_cast<float*>(base);
  v.fmt = FMT_F32;
  v.T = output ? L.out_t : L.in_t;
  v.H = output ? L.out_h : L.in_h;
  v.W = output ? L.out_w : L.in_w;
  v.C = output ? L.cout : L.cin;
  v.chunk_stride = (long long)v.T * v.H * v.W * v.C;
  return v;
}

int conv_direct_views(const lc_codec* h, const DevLayer& L, const float* d_w, const ActView& in,
                      const ActView& out, const ActView* skip, int n_chunks, cudaStream_t s) {
  ConvParams p;
  p.in = in; p.out = out;
  if (skip) p.skip = *skip; else memset(&p.skip, 0, sizeof(p.skip));
  p.w = d_w; p.b = L.d_b;
  p.kind = L.kind; p.k = L.k; p.stride = L.stride; p.cin = L.cin; p.cout = L.cout;
  p.relu = L.relu; p.pad_lo = L.pad_lo;
  p.total_vox = (long long)n_chunks * L.out_t * L.out_h * L.out_w;
  if (L.out_t * L.out_h * L.out_w <= 128 && L.cout <= 24 && L.cin <= 32 && (out.fmt == FMT_F32 || out.G * 8 <= 32) &&
      !getenv("LOSSYCOMP_NO_SMALL")) {
    long long blocks = (p.total_vox + 7) / 8;
    if (blocks > (long long)h->num_sms * 16) blocks = (long long)h->num_sms * 16;
    if (blocks < 1) blocks = 1;
    if (L.cout <= 8) conv_small_kernel<8><<<(unsigned)blocks, 256, 0, s>>>(p);
    else conv_small_kernel<24><<<(unsigned)blocks, 256, 0, s>>>(p);
    LC_LAUNCH_CHECK();
    return LC_OK;
  }
  const int co4 = (L.cout + 3) / 4;
  switch (co4) {
    case 1: return launch<1>(p, h->num_sms, s);
    case 2: return launch<2>(p, h->num_sms, s);
    case 3: return launch<3>(p, h->num_sms, s);
    case 4: return launch<4>(p, h->num_sms, s);
    case 5: return launch<5>(p, h->num_sms, s);
    case 6: return launch<6>(p, h->num_sms, s);
    case 7: return launch<7>(p, h->num_sms, s);
    case 8: return launch<8>(p, h->num_sms, s);
    default:
      lc_set_error("conv_direct: cout %d not supported (max 32)", L.cout);
      return LC_EUNSUPPORTED;
  }
}

int conv_direct_launch(const lc_codec* h, const DevLayer& L, const float* d_in, float* d_out,
                       const float* d_skip, int n_chunks, cudaStream_t s) {
  ActView in = f32_view(L, false, d_in), out = f32_view(L, true, d_out), sk;
  if (d_skip) sk = f32_view(L, true, d_skip);
  return conv_direct_views(h, L, L.d_w, in, out, d_skip ? &sk : nullptr, n_chunks, s);
}
