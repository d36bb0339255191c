// fp32 CUDA-core direct Conv3D / Conv3DTranspose over channels-last fp32 activations.
//
// This is the LC_PREC_FP32 back end (reference-grade numerics: fp32 FMA, fixed tap order) and
// the back end of the small inner layers in the tensor-core modes.  Semantics follow Keras
// (models.py:47-97; SURVEY.md §8c a-M): cross-correlation, "same" padding with
// before = total/2, Conv3DTranspose = full transposed convolution cropped by (k-s)/2.
//
// One thread owns one output voxel and all output channels; weights live in shared memory as
// [tap][cin][cout padded to 4] so that every FMA quad is fed by one broadcast LDS.128.
#include "common.cuh"

namespace {

struct ConvParams {
  const float* in;
  float* out;
  const float* skip;
  const float* w;     // [taps][cin][cout]
  const float* b;
  int kind, k, stride, cin, cout, cout_pad, relu, pad_lo;
  int in_t, in_h, in_w, out_t, out_h, out_w;
  long long total_vox;  // n_chunks * out_t*out_h*out_w
};

template <int CO4>  // cout_pad / 4
__global__ void __launch_bounds__(512, 1) conv_direct_kernel(ConvParams p) {
  extern __shared__ float4 smem_w4[];
  float* smem_w = reinterpret_cast<float*>(smem_w4);
  const int taps = p.k * p.k * p.k;
  const int cp = CO4 * 4;
  // stage weights: [tap*cin + ci][cp], zero padded
  for (int i = threadIdx.x; i < taps * p.cin * cp; i += blockDim.x) {
    int co = i % cp;
    int r = i / cp;
    smem_w[i] = (co < p.cout) ? p.w[(size_t)r * p.cout + co] : 0.f;
  }
  __syncthreads();

  const int ovox = p.out_t * p.out_h * p.out_w;
  const int ivox = p.in_t * p.in_h * p.in_w;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < p.total_vox;
       v += (long long)gridDim.x * blockDim.x) {
    const int chunk = (int)(v / ovox);
    int r = (int)(v % ovox);
    const int ow = r % p.out_w;
    r /= p.out_w;
    const int oh = r % p.out_h;
    const int ot = r / p.out_h;
    const float* inb = p.in + (size_t)chunk * ivox * p.cin;

    float acc[CO4 * 4];
#pragma unroll
    for (int c = 0; c < CO4 * 4; ++c) acc[c] = (c < p.cout) ? p.b[c] : 0.f;

    for (int kt = 0; kt < p.k; ++kt) {
      int it;
      if (p.kind == 0) {
        it = ot * p.stride + kt - p.pad_lo;
      } else {
        int num = ot + p.pad_lo - kt;
        if (num < 0 || (num % p.stride) != 0) continue;
        it = num / p.stride;
      }
      if (it < 0 || it >= p.in_t) continue;
      for (int kh = 0; kh < p.k; ++kh) {
        int ih;
        if (p.kind == 0) {
          ih = oh * p.stride + kh - p.pad_lo;
        } else {
          int num = oh + p.pad_lo - kh;
          if (num < 0 || (num % p.stride) != 0) continue;
          ih = num / p.stride;
        }
        if (ih < 0 || ih >= p.in_h) continue;
        for (int kw = 0; kw < p.k; ++kw) {
          int iw;
          if (p.kind == 0) {
            iw = ow * p.stride + kw - p.pad_lo;
          } else {
            int num = ow + p.pad_lo - kw;
            if (num < 0 || (num % p.stride) != 0) continue;
            iw = num / p.stride;
          }
          if (iw < 0 || iw >= p.in_w) continue;
          const float* xin = inb + ((size_t)(it * p.in_h + ih) * p.in_w + iw) * p.cin;
          const float4* wt = smem_w4 + (size_t)((kt * p.k + kh) * p.k + kw) * p.cin * CO4;
          for (int ci = 0; ci < p.cin; ++ci) {
            const float x = __ldg(xin + ci);
#pragma unroll
            for (int c4 = 0; c4 < CO4; ++c4) {
              const float4 w = wt[ci * CO4 + c4];
              acc[c4 * 4 + 0] = fmaf(x, w.x, acc[c4 * 4 + 0]);
              acc[c4 * 4 + 1] = fmaf(x, w.y, acc[c4 * 4 + 1]);
              acc[c4 * 4 + 2] = fmaf(x, w.z, acc[c4 * 4 + 2]);
              acc[c4 * 4 + 3] = fmaf(x, w.w, acc[c4 * 4 + 3]);
            }
          }
        }
      }
    }
    float* o = p.out + (size_t)v * p.cout;
    const float* sk = p.skip ? p.skip + (size_t)v * p.cout : nullptr;
#pragma unroll
    for (int c = 0; c < CO4 * 4; ++c) {
      if (c < p.cout) {
        float y = acc[c];
        if (sk) y = __fadd_rn(y, sk[c]);
        if (p.relu) y = fmaxf(y, 0.f);
        o[c] = y;
      }
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

int conv_direct_launch(const lc_codec* h, const DevLayer& L, const float* d_in, float* d_out,
                       const float* d_skip, int n_chunks, cudaStream_t s) {
  ConvParams p;
  p.in = d_in; p.out = d_out; p.skip = d_skip; p.w = L.d_w; p.b = L.d_b;
  p.kind = L.kind; p.k = L.k; p.stride = L.stride; p.cin = L.cin; p.cout = L.cout;
  p.relu = L.relu; p.pad_lo = L.pad_lo;
  p.in_t = L.in_t; p.in_h = L.in_h; p.in_w = L.in_w;
  p.out_t = L.out_t; p.out_h = L.out_h; p.out_w = L.out_w;
  p.total_vox = (long long)n_chunks * L.out_t * L.out_h * L.out_w;
  const int co4 = (L.cout + 3) / 4;
  p.cout_pad = co4 * 4;
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
