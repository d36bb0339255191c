// Activation tensor views.
//
// Two storage formats exist on the convolution path:
//   FMT_F32   channels-last fp32 (N,T,H,W,C)  -- the Keras layout; used for the chunk input,
//             the latent and the decoder output, and for everything in LC_PREC_FP32 mode;
//   FMT_GP16  "group-plane" 16-bit (bf16 or fp16): channels are cut into groups of 8 (one
//             16-byte vector = the K extent of one UMMA core-matrix row), and every
//             (chunk, t, parity, group) owns a plane of S slots, one 16-byte vector per voxel of
//             the zero-haloed (H+pad)x(W+pad) image in row-major order with row pitch PW:
//                 vec index = (((n*T + t)*NP + p)*G + g)*S + GL + slot
//             plain (NP=1):   slot = (h+1)*PW + (w+1)
//             parity (NP=4):  hp=h+1, wp=w+1, p = (hp&1)*2 + (wp&1), slot = (hp>>1)*PW + (wp>>1)
//             Consecutive voxels of a line are consecutive vectors, so a conv tap (dh,dw) of a
//             128-voxel tile is the same 128-row UMMA operand started (dh*PW+dw)*16 bytes later,
//             and halo slots are real zeros in memory (never written after allocation).
#pragma once
#include "common.cuh"

enum { FMT_F32 = 0, FMT_GP16 = 1 };

struct ActView {
  void* base;
  int fmt;           // FMT_*
  int half;          // GP16 only: 0 = bf16, 1 = fp16
  int T, H, W, C;    // logical extent of one chunk
  int G, NP, PW, S, GL;
  long long chunk_stride;   // F32: floats per chunk; GP16: 16-byte vectors per chunk block
};

__host__ __device__ inline long long gp_vec_index(const ActView& v, int n, int t, int h, int w, int g) {
  int p = 0, slot;
  if (v.NP == 1) {
    slot = (h + 1) * v.PW + (w + 1);
  } else {
    const int hp = h + 1, wp = w + 1;
    p = (hp & 1) * 2 + (wp & 1);
    slot = (hp >> 1) * v.PW + (wp >> 1);
  }
  return (long long)n * v.chunk_stride + ((long long)(t * v.NP + p) * v.G + g) * v.S + v.GL + slot;
}

__device__ __forceinline__ float h16_to_float(unsigned short u, int half) {
  if (half) return __half2float(__ushort_as_half(u));
  return __uint_as_float((unsigned)u << 16);
}
__device__ __forceinline__ unsigned short float_to_h16(float f, int half) {
  if (half) return __half_as_ushort(__float2half_rn(fminf(fmaxf(f, -65504.f), 65504.f)));   // saturate
  return __bfloat16_as_ushort(__float2bfloat16_rn(f));
}
__device__ __forceinline__ void unpack8(const uint4 v, int half, float f[8]) {
  const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = h16_to_float((unsigned short)(w[i] & 0xffff), half);
    f[2 * i + 1] = h16_to_float((unsigned short)(w[i] >> 16), half);
  }
}
// two floats -> one packed 16-bit pair with a single F2FP instruction (same round-to-nearest-even
// results as two scalar conversions; fp16 saturates at +-65504 like float_to_h16)
__device__ __forceinline__ unsigned pack2(float a, float b, int half) {
  if (half) {
    const __half2 h = __floats2half2_rn(fminf(fmaxf(a, -65504.f), 65504.f), fminf(fmaxf(b, -65504.f), 65504.f));
    return *reinterpret_cast<const unsigned*>(&h);
  }
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const unsigned*>(&h);
}
// Fused (ReLU +) saturate + round-to-nearest-even + pack: one F2FP per pair.  Identical results to
// relu -> clamp -> convert done separately (.relu maps negatives to +0, .satfinite clamps to +-65504).
__device__ __forceinline__ unsigned pack2_act(float a, float b, int half, int relu) {
  unsigned r;
  if (half) {
    if (relu) asm("cvt.rn.satfinite.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    else asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  } else {
    if (relu) asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  }
  return r;
}
__device__ __forceinline__ uint4 pack8_act(const float f[8], int half, int relu) {
  return make_uint4(pack2_act(f[0], f[1], half, relu), pack2_act(f[2], f[3], half, relu),
                    pack2_act(f[4], f[5], half, relu), pack2_act(f[6], f[7], half, relu));
}
__device__ __forceinline__ uint4 pack8(const float f[8], int half) {
  return make_uint4(pack2(f[0], f[1], half), pack2(f[2], f[3], half), pack2(f[4], f[5], half),
                    pack2(f[6], f[7], half));
}

// round a float through the 16-bit operand format (host side, for weights)
static inline float round_h16_host(float f, int half) {
  if (half) return __half2float(__float2half_rn(f));
  return __bfloat162float(__float2bfloat16_rn(f));
}
