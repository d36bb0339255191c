// Tensor-core (tcgen05 / TMEM / bulk-async-copy) back end of the autoencoder, LC_PREC_BF16 and
// LC_PREC_FP16 modes.
//
// Plane-sweep implicit GEMM.  Common idea of all five kernel modes:
//   * activations are 16-bit group-plane tensors (act.cuh): a 128-voxel tile of one time plane is a
//     128-row K-major UMMA operand, and an in-plane tap (dh,dw) is the same operand started
//     (dh*PW+dw) vectors later -- no im2col, no swizzle, one 1-D bulk copy per channel-group plane;
//   * the time taps are folded into N: D[voxel, (dt,co)] = sum_{taps,ci} X[t_in][voxel+tap][ci] *
//     W[dt][tap][ci][co], so one A tile read from shared memory feeds up to 4x23 output columns;
//     with N=32 the MMA would be bound by the 128 B/clk shared-memory read of the A operand;
//   * a CTA sweeps the input planes t_in for a fixed 128-voxel tile: the dt-block of D belongs to an
//     output plane that is the same TMEM lane = the same epilogue thread for every t_in, so the fold
//     is undone by thread-local register accumulation (no cross-lane traffic, fixed order);
//   * K steps pair 8-channel groups across taps through the descriptor's leading-byte-offset
//     (k=3: 27 groups -> 14 MMAs per plane instead of 18 with channels padded to 32);
//   * warp-specialised (352 threads): warp 0 = bulk-copy producer, warps 1 and 10 = MMA issuers taking alternate
//     accumulators (the whole warp runs the loop so that every tcgen05.mma operand is uniform, one elected lane
//     issues), warps 2-9 = epilogue (two per TMEM lane quarter, 12 channels each: TMEM -> registers -> ReLU /
//     16-bit packing -> stores); smem ring of input planes, ring of 2-4 TMEM accumulators; two CTAs per SM
//     except DOWN and UP.
// Modes:
//   K3       Conv3D k=3 s=1 (the ResBlock convs, 80 % of the FLOPs): 9 taps x 3 groups, N = 3x24(+8); with a
//            residual input, three more K groups multiply the skip tile by identity weights (15 MMAs per plane)
//   K4       Conv3D k=4 s=1, first layer: input pre-folded along w by the chunk gather (4 taps x ceil(4*Cin/8)
//            groups, N = 4x24)
//   K4_OUT1  Conv3D k=4 s=1, last layer (cout = 1): time and w taps folded into N = 16, 4 h-taps x 3 groups,
//            fp32 output scattered straight into the merged field (= dataLoader.merge_data fused)
//   DOWN     Conv3D k=4 s=2 on a parity-split input (stride-2 taps become stride-1 taps of the four (h,w)-parity
//            planes), N = 2x24, weights selected by the parity of t_in
//   UP       Conv3DTranspose k=4 s=2: item = (chunk, output-line parity, tile); both output-column parities in
//            the same CTA (two accumulators per input plane, 4 taps x 3 groups, N = 4x24 = the four kt each),
//            packed results transposed across lanes so that every store covers 32 consecutive output columns
// Everything is deterministic: each output voxel is produced by exactly one thread in a fixed
// accumulation order, independent of the item -> CTA mapping.
#include "act.cuh"
#include "umma.cuh"

namespace {

constexpr int kTcThreads = 352;   // warp 0 producer, warps 1 and 10 MMA issuers (alternate planes), warps 2-9 epilogue
constexpr int kHalf = 12;         // output channels per epilogue thread (two threads share a TMEM lane)
constexpr int kMaxKs = 24;
constexpr int kMaxVar = 4;
constexpr int kMaxBufs = 4;       // TMEM accumulator ring (p.nbuf buffers, p.acc_stride columns apart)
constexpr int kMaxStages = 4;

enum { MODE_K3 = 1, MODE_K4 = 2, MODE_K4_OUT1 = 3, MODE_DOWN = 4, MODE_UP = 5 };

struct SweepParams {
  // input (GP16): planes per time step = in_NPG (parities x groups)
  const uint4* in;
  long long in_cs;
  int in_S, in_GL, in_NPG, in_T;
  // output (GP16) and skip (GP16, same logical extent as the output)
  ActView out, skip;
  // MODE_K4_OUT1: merged fp32 field
  float* out_f32;
  ChunkGeom geom;
  int chunk_begin;
  // row numbering of a tile: row r of tile j is slot row0 + 128*j + r of a pitch-row_PW image;
  // (hp, wp) = divmod(slot, row_PW); voxel = (hp - row_off, wp - row_off), valid inside [0,H)x[0,W)
  int row_PW, row0, row_off_h, row_off_w, rows_H, rows_W;
  int row_step, row_shift;       // tile j, row r -> slot row0 + row_step*j + r - row_shift (overlapping tiles)
  int tiles_per_chunk, n_items, n_classes;
  int slab_slots, slab_start;    // the stage holds slots [row0 + 128*j + slab_start, +slab_slots) of every plane
  int stages;
  int nbuf, acc_stride, tmem_cols;   // TMEM accumulator ring
  // MODE_K3 residual add done by the tensor core: the skip tile of plane t rides in stage t as skip.G extra K groups
  // multiplied by identity weights in the centre time block (skip_slots vectors per group, the first skip_pad of
  // them alignment padding); 0 = the epilogue adds the skip from global memory
  int skip_mma, skip_slots, skip_pad;
  // MMA
  int N, nks, nvar;
  const uint4* wimg;
  int wimg_bytes;
  uint32_t b_base[kMaxVar];
  uint32_t a_off[kMaxVar][kMaxKs], a_lbo[kMaxVar][kMaxKs];
  uint64_t a_desc[kMaxVar][kMaxKs], b_desc[kMaxVar][kMaxKs];   // descriptors for smem offsets relative to sA / sB
  const float* bias;
  int cout, relu, half, out_T;
  int debug;   // ablation: bit0 no MMA issue, bit2 no global stores (sweep_acc_kernel), bit3 no bulk copies, bit4 single issuer warp
  // sweep_acc_kernel (time taps accumulated in TMEM): weight image with the time blocks stored twice in decreasing
  // order, 8-row-group pitch b2_rows, descriptors relative to its start
  const uint4* wimg2;
  int wimg2_bytes, b2_rows, stages2;
  uint64_t b2_desc[kMaxKs];
  // CTA-pair form (cta_group::2): per-rank weight images holding the rows each CTA's half of N reads
  const uint4* wimg2c0;
  const uint4* wimg2c1;
  int wimg2c_bytes, b2c_rows, stages2c;
  uint64_t b2c_desc[kMaxKs];
  // sweep_acc_kernel: 1 = the epilogue stages every output plane in shared memory and writes it with bulk copies
  // (needs the output planes in the input's row numbering: a tile's 128 rows are 128 consecutive output slots)
  int tma_store;
};

// 12 consecutive accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void ld_cols12(uint32_t taddr, float d[kHalf]) {
  uint32_t v0[8], v1[4];
  umma::tmem_ld8(taddr, v0);
  umma::tmem_ld4(taddr + 8, v1);
  umma::tmem_ld_wait();
#pragma unroll
  for (int k = 0; k < 8; ++k) d[k] = __uint_as_float(v0[k]);
#pragma unroll
  for (int k = 0; k < 4; ++k) d[8 + k] = __uint_as_float(v1[k]);
}

// NB dt-blocks (24 columns apart) of 12 columns each: all TMEM loads are issued before the single wait
template <int NB>
__device__ __forceinline__ void ld_blocks(uint32_t taddr, float (&d)[NB][kHalf]) {
  uint32_t v0[NB][8], v1[NB][4];
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    umma::tmem_ld8(taddr + b * 24, v0[b]);
    umma::tmem_ld4(taddr + b * 24 + 8, v1[b]);
  }
  umma::tmem_ld_wait();
#pragma unroll
  for (int b = 0; b < NB; ++b) {
#pragma unroll
    for (int k = 0; k < 8; ++k) d[b][k] = __uint_as_float(v0[b][k]);
#pragma unroll
    for (int k = 0; k < 4; ++k) d[b][8 + k] = __uint_as_float(v1[b][k]);
  }
}

struct RowInfo {
  int n, cls, h, w;
  bool valid;
};

__device__ __forceinline__ RowInfo row_info(const SweepParams& p, int item, int row) {
  RowInfo r;
  const int per_chunk = p.tiles_per_chunk * p.n_classes;
  r.n = item / per_chunk;
  const int rem = item - r.n * per_chunk;
  r.cls = rem / p.tiles_per_chunk;
  const int tile = rem - r.cls * p.tiles_per_chunk;
  const int slot = p.row0 + tile * p.row_step + row - p.row_shift;
  const int hp = slot / p.row_PW, wp = slot - hp * p.row_PW;
  r.h = hp - p.row_off_h;
  r.w = wp - p.row_off_w;
  r.valid = (r.h >= 0) && (r.h < p.rows_H) && (r.w >= 0) && (r.w < p.rows_W);
  return r;
}

// Output / skip addressing of one tile row: vector pointers of (n, t = 0, h, w, group 0)
struct VoxPtr {
  uint4* o;
  const uint4* s;          // nullptr when the layer has no residual input
  long long o_t, s_t;      // vectors per time plane
  long long o_g, s_g;      // vectors per group plane
};

__device__ __forceinline__ VoxPtr vox_ptr(const SweepParams& p, int n, int h, int w) {
  VoxPtr a;
  a.o = reinterpret_cast<uint4*>(p.out.base) + gp_vec_index(p.out, n, 0, h, w, 0);
  a.o_t = (long long)p.out.NP * p.out.G * p.out.S;
  a.o_g = p.out.S;
  a.s = p.skip.base ? reinterpret_cast<const uint4*>(p.skip.base) + gp_vec_index(p.skip, n, 0, h, w, 0) : nullptr;
  a.s_t = (long long)p.skip.NP * p.skip.G * p.skip.S;
  a.s_g = p.skip.S;
  return a;
}

// The 12 channels [12*half, 12*half+12) of a voxel as stored: half 0 owns vector g0 and the low 8 bytes of
// g1, half 1 the high 8 bytes of g1 and vector g2.
struct SkipRegs {
  uint4 a;
  uint2 b;
};

__device__ __forceinline__ SkipRegs load_skip(const uint4* sk, long long s_g, int half) {
  SkipRegs r;
  if (half == 0) {
    r.a = __ldg(sk);
    r.b = __ldg(reinterpret_cast<const uint2*>(sk + s_g));
  } else {
    r.b = __ldg(reinterpret_cast<const uint2*>(sk + s_g) + 1);
    r.a = __ldg(sk + 2 * s_g);
  }
  return r;
}

template <int H16>
__device__ __forceinline__ void unpack2_t(unsigned w, float& lo, float& hi) {
  if (H16) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w));
    lo = f.x; hi = f.y;
  } else {
    lo = __uint_as_float(w << 16);
    hi = __uint_as_float(w & 0xffff0000u);
  }
}

// (ReLU +) saturate + round-to-nearest-even + pack of two floats: one F2FP
template <int H16, int RELU>
__device__ __forceinline__ unsigned pack2_t(float a, float b) {
  unsigned r;
  if (H16) {
    if (RELU) asm("cvt.rn.satfinite.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    else asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  } else {
    if (RELU) asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  }
  return r;
}

// y (bias already inside) (+ skip) -> ReLU -> 16-bit -> the voxel's 12 channels of this half at plane pointer o.
// Channels beyond cout need no masking: their weight columns, bias and skip padding are zero, so y is +0.
template <int H16, int RELU, bool SKIP>
__device__ __forceinline__ void store12(uint4* o, long long o_g, int G, int half, const float (&acc)[kHalf],
                                        const SkipRegs& sk) {
  float y[kHalf];
#pragma unroll
  for (int c = 0; c < kHalf; ++c) y[c] = acc[c];
  if (SKIP) {
    float f[kHalf];   // the skip channels in the order of y
    if (half == 0) {
      unpack2_t<H16>(sk.a.x, f[0], f[1]); unpack2_t<H16>(sk.a.y, f[2], f[3]);
      unpack2_t<H16>(sk.a.z, f[4], f[5]); unpack2_t<H16>(sk.a.w, f[6], f[7]);
      unpack2_t<H16>(sk.b.x, f[8], f[9]); unpack2_t<H16>(sk.b.y, f[10], f[11]);
    } else {
      unpack2_t<H16>(sk.b.x, f[0], f[1]); unpack2_t<H16>(sk.b.y, f[2], f[3]);
      unpack2_t<H16>(sk.a.x, f[4], f[5]); unpack2_t<H16>(sk.a.y, f[6], f[7]);
      unpack2_t<H16>(sk.a.z, f[8], f[9]); unpack2_t<H16>(sk.a.w, f[10], f[11]);
    }
#pragma unroll
    for (int c = 0; c < kHalf; ++c) y[c] = __fadd_rn(y[c], f[c]);
  }
  if (half == 0) {
    o[0] = make_uint4(pack2_t<H16, RELU>(y[0], y[1]), pack2_t<H16, RELU>(y[2], y[3]),
                      pack2_t<H16, RELU>(y[4], y[5]), pack2_t<H16, RELU>(y[6], y[7]));
    if (G > 1)
      *reinterpret_cast<uint2*>(o + o_g) = make_uint2(pack2_t<H16, RELU>(y[8], y[9]), pack2_t<H16, RELU>(y[10], y[11]));
  } else {
    if (G > 1)
      *(reinterpret_cast<uint2*>(o + o_g) + 1) = make_uint2(pack2_t<H16, RELU>(y[0], y[1]), pack2_t<H16, RELU>(y[2], y[3]));
    if (G > 2)
      o[2 * o_g] = make_uint4(pack2_t<H16, RELU>(y[4], y[5]), pack2_t<H16, RELU>(y[6], y[7]),
                              pack2_t<H16, RELU>(y[8], y[9]), pack2_t<H16, RELU>(y[10], y[11]));
  }
}

template <int MODE, int H16, int RELU>
__global__ void __launch_bounds__(kTcThreads, (MODE == MODE_DOWN || MODE == MODE_UP) ? 1 : 2)
sweep_tc_kernel(const __grid_constant__ SweepParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t full[kMaxStages], empty[kMaxStages], tfull[kMaxBufs], tempty[kMaxBufs], wbar;
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_bias[24];

  // warp index made provably warp-uniform, so that role branches are uniform control flow and the
  // issuer's descriptor arithmetic stays in uniform registers
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const int slab_bytes = p.slab_slots * 16;
  const int skip_bytes = p.skip_mma ? p.skip.G * p.skip_slots * 16 : 0;
  const int stage_bytes = p.in_NPG * slab_bytes + skip_bytes;
  uint8_t* sB = smem;
  uint8_t* sA = smem + ((p.wimg_bytes + 127) & ~127);
  if (threadIdx.x == 0) {
    for (int i = 0; i < p.stages; ++i) { umma::mbar_init(&full[i], 1); umma::mbar_init(&empty[i], 1); }
    for (int i = 0; i < p.nbuf; ++i) { umma::mbar_init(&tfull[i], 1); umma::mbar_init(&tempty[i], (MODE == MODE_K4_OUT1) ? 4 : 8); }
    umma::mbar_init(&wbar, 1);
    umma::fence_barrier_init();
  }
  if (threadIdx.x < 24) s_bias[threadIdx.x] = (threadIdx.x < p.cout) ? p.bias[threadIdx.x] : 0.f;
  if (warp == 1) umma::tmem_alloc(&tmem_base_s, (uint32_t)p.tmem_cols);
  umma::tcgen05_fence_before();
  __syncthreads();
  umma::tcgen05_fence_after();
  const uint32_t tmem = tmem_base_s;
  const int per_chunk = p.tiles_per_chunk * p.n_classes;

  if (warp == 0) {
    // ===================== producer: bulk async copies (TMA engine) =====================
    // The whole warp works: lane c issues copy c of a stage (input slabs first, then the skip tiles), lane 0 also
    // waits for the stage and posts the byte count.  One thread issuing every copy of every stage was the slowest
    // stage of the pipeline (its ~150 dependent scalar instructions per stage outlasted the MMAs of the stage).
    if (lane == 0) {
      umma::mbar_arrive_expect_tx(&wbar, (uint32_t)p.wimg_bytes);
      umma::bulk_g2s(sB, p.wimg, (uint32_t)p.wimg_bytes, &wbar);
    }
    const int ncop = p.in_NPG + (p.skip_mma ? p.skip.G : 0);      // <= 32 (checked by the host)
    const bool is_in = lane < p.in_NPG, act = lane < ncop;
    const int cs = lane - p.in_NPG;                               // skip copy index
    const uint32_t my_bytes = is_in ? (uint32_t)slab_bytes : (uint32_t)(p.skip_slots * 16);
    const int my_dst = is_in ? lane * slab_bytes : p.in_NPG * slab_bytes + cs * p.skip_slots * 16;
    const long long my_tstride = is_in ? (long long)p.in_NPG * p.in_S : (long long)p.skip.G * p.skip.S;
    int stage = 0, phase = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      const int n = item / per_chunk;
      const int tile = (item - n * per_chunk) % p.tiles_per_chunk;
      const uint4* src;
      if (is_in)
        src = p.in + (long long)n * p.in_cs + p.in_GL + p.row0 + tile * p.row_step - p.row_shift + p.slab_start +
              (long long)lane * p.in_S;
      else   // rows of this tile of the skip planes (same row numbering as the input planes)
        src = reinterpret_cast<const uint4*>(p.skip.base) + (long long)n * p.skip.chunk_stride + p.skip.GL + p.row0 +
              tile * 128 - p.skip_pad + (long long)cs * p.skip.S;
      for (int t = 0; t < p.in_T; ++t) {
        if (lane == 0) {
          umma::mbar_wait(&empty[stage], phase ^ 1);
          if (p.debug & 8) umma::mbar_arrive(&full[stage]);
          else umma::mbar_arrive_expect_tx(&full[stage], (uint32_t)stage_bytes);
        }
        __syncwarp();
        if (act && !(p.debug & 8))
          umma::bulk_g2s(sA + stage * stage_bytes + my_dst, src + (long long)t * my_tstride, my_bytes, &full[stage]);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 || warp == 10) {
    // ===================== MMA issuers =====================
    // Two issuer warps take alternate planes (different TMEM accumulators, so their MMAs may interleave
    // freely on the tensor pipe): the waits + commit one issuer spends per plane overlap the other's
    // MMAs instead of draining the pipe.  The whole warp runs the loop so that every operand of
    // tcgen05.mma is warp-uniform (descriptors come from the constant bank -> uniform registers, no
    // per-MMA R2UR traffic); one elected lane issues.
    {
      const int me = (warp == 10) ? 1 : 0;
      const int n_issuers = (p.debug & 16) ? 1 : 2;
      umma::mbar_wait(&wbar, 0);
      const uint32_t idesc = umma::make_idesc(128, p.N, H16 ? 0 : 1);
      const uint64_t sA_add = (uint64_t)(umma::smem_u32(sA) >> 4), sB_add = (uint64_t)(umma::smem_u32(sB) >> 4);
      int stage = 0, phase = 0, buf = 0, bphase = 0, step = 0;
      // MODE_UP runs two accumulator sub-steps per input plane (output-column parity 0 / 1, weight variants
      // 2*ph and 2*ph + 1) on the same stage; every other mode one
      constexpr int kSub = (MODE == MODE_UP) ? 2 : 1;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        int var = 0;
        if (MODE == MODE_UP) var = 2 * ((item % per_chunk) / p.tiles_per_chunk);
        for (int t = 0; t < p.in_T; ++t) {
          if (MODE == MODE_DOWN) var = t & 1;
#pragma unroll
          for (int sub = 0; sub < kSub; ++sub, ++step) {
            const bool mine = (n_issuers == 1) ? (me == 0) : ((step & 1) == me);
            if (mine) {
              umma::mbar_wait(&tempty[buf], bphase ^ 1);
              umma::mbar_wait(&full[stage], phase);
              umma::tcgen05_fence_after();
              const uint64_t a_add = sA_add + (uint64_t)((uint32_t)(stage * stage_bytes) >> 4);
              const uint32_t d_addr = tmem + buf * p.acc_stride;
              const int nks = (p.debug & 1) ? 0 : p.nks;
#pragma unroll 2
              for (int ks = 0; ks < nks; ++ks) {
                const uint64_t a = p.a_desc[var + sub][ks] + a_add, b = p.b_desc[var + sub][ks] + sB_add;
                if (umma::elect_one()) umma::mma_f16(d_addr, a, b, idesc, ks > 0);
              }
              __syncwarp();
              // one commit per accumulator: it signals the epilogue, which releases the smem stage
              if (umma::elect_one()) umma::commit(&tfull[buf]);
              __syncwarp();
            }
            if (++buf == p.nbuf) { buf = 0; bphase ^= 1; }
          }
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp >= 2 && warp <= 9) {
    // ===================== epilogue: TMEM -> registers -> global =====================
    // 8 warps: warp w reads TMEM lanes 32*(w&3)..+31 (hardware rule); warps 2-5 take output channels
    // 0..11 of every dt block, warps 6-9 channels 12..23.  The warps are issue-bound, so the loops below
    // are written to cost few instructions per plane: accumulators rotate through unrolled steps instead
    // of register moves, the bias is added where a plane's first block arrives (replacing a move), ReLU
    // and the 16-bit rounding are one F2FP per channel pair, nothing is masked.
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16) + half * kHalf;
    float bias_r[kHalf];
#pragma unroll
    for (int c = 0; c < kHalf; ++c) bias_r[c] = s_bias[half * kHalf + c];
    int ebuf = 0, ephase = 0, estage = 0;
    auto wait_d = [&](bool last_of_stage = true) -> uint32_t {
      umma::mbar_wait(&tfull[ebuf], ephase);
      umma::tcgen05_fence_after();
      // the MMAs that read this step's stage are complete: hand the stage back to the producer
      if (last_of_stage) {
        if (warp == 2 && lane == 0) umma::mbar_arrive_relaxed(&empty[estage]);
        if (++estage == p.stages) estage = 0;
      }
      return lane_addr + ebuf * p.acc_stride;
    };
    auto release_d = [&]() {
      umma::tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) umma::mbar_arrive_relaxed(&tempty[ebuf]);
      if (++ebuf == p.nbuf) { ebuf = 0; ephase ^= 1; }
    };
    const int G = p.out.G;
    const int T = p.in_T;
    if (MODE == MODE_K4_OUT1 && half == 1) {
      // single output channel: the second half of the epilogue warps has nothing to do
    } else
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      const RowInfo ri = row_info(p, item, row);
      const bool valid = ri.valid;
      VoxPtr vp;
      if (MODE != MODE_K4_OUT1 && MODE != MODE_UP) vp = vox_ptr(p, ri.n, ri.h, ri.w);
      const SkipRegs no_skip = {};
      if (MODE == MODE_K3) {
        // block dt -> output plane t_in - dt + 1.  Roles at step t: P = plane t-1 (completes), C = plane t,
        // N = plane t+1 (starts with this step's block 0 + bias).
        auto run = [&](auto skip_tag) {
          constexpr bool SKIP = decltype(skip_tag)::value;
          float a0[kHalf], a1[kHalf], a2[kHalf];
#pragma unroll
          for (int c = 0; c < kHalf; ++c) { a0[c] = 0.f; a1[c] = bias_r[c]; }
          uint4* op = vp.o - vp.o_t;            // plane t-1 of the output / skip tensors, advanced every step
          const uint4* sp = vp.s - vp.s_t;
          auto step = [&](float (&aP)[kHalf], float (&aC)[kHalf], float (&aN)[kHalf], int t) {
            SkipRegs sk = no_skip;
            // the skip values of the plane finished in this step are fetched before waiting for the MMA
            if (SKIP && t >= 1 && valid) sk = load_skip(sp, vp.s_g, half);
            const uint32_t ta = wait_d();
            if (SKIP) {
              // blocks 1,2 first: plane t-1 completes and is stored, freeing registers before block 0 arrives
              // (with the six skip registers the three-block form spills)
              {
                float db[2][kHalf];
                ld_blocks<2>(ta + 24, db);
#pragma unroll
                for (int c = 0; c < kHalf; ++c) { aC[c] += db[0][c]; aP[c] += db[1][c]; }
              }
              if (t >= 1 && valid) store12<H16, RELU, SKIP>(op, vp.o_g, G, half, aP, sk);
              float d0[kHalf];
              ld_cols12(ta, d0);
              release_d();
#pragma unroll
              for (int c = 0; c < kHalf; ++c) aN[c] = d0[c] + bias_r[c];
            } else {
              float db[3][kHalf];
              ld_blocks<3>(ta, db);
              release_d();
#pragma unroll
              for (int c = 0; c < kHalf; ++c) {
                aN[c] = db[0][c] + bias_r[c];
                aC[c] += db[1][c];
                aP[c] += db[2][c];
              }
              if (t >= 1 && valid) store12<H16, RELU, SKIP>(op, vp.o_g, G, half, aP, sk);
            }
            op += vp.o_t;
            if (SKIP) sp += vp.s_t;
          };
          auto finish = [&](float (&aP)[kHalf]) {
            if (!valid) return;
            SkipRegs sk = no_skip;
            if (SKIP) sk = load_skip(sp, vp.s_g, half);
            store12<H16, RELU, SKIP>(op, vp.o_g, G, half, aP, sk);
          };
          int t = 0;
          for (; t + 3 <= T; t += 3) { step(a0, a1, a2, t); step(a1, a2, a0, t + 1); step(a2, a0, a1, t + 2); }
          if (T - t == 0) {
            finish(a0);
          } else if (T - t == 1) {
            step(a0, a1, a2, t);
            finish(a1);
          } else {
            step(a0, a1, a2, t);
            step(a1, a2, a0, t + 1);
            finish(a2);
          }
        };
        if (vp.s != nullptr && !p.skip_mma) run(std::true_type{}); else run(std::false_type{});
      } else if (MODE == MODE_K4) {
        // k=4, pad 1/2: block dt -> output plane t_in - dt + 1.  Roles at step t: A3 = plane t-2 (completes),
        // A2 = t-1, A1 = t, A0 = t+1 (starts).
        float b0[kHalf], b1[kHalf], b2[kHalf], b3[kHalf];
#pragma unroll
        for (int c = 0; c < kHalf; ++c) { b0[c] = 0.f; b1[c] = 0.f; b2[c] = bias_r[c]; }
        uint4* op = vp.o - 2 * vp.o_t;          // plane t-2
        auto step = [&](float (&A3)[kHalf], float (&A2)[kHalf], float (&A1)[kHalf], float (&A0)[kHalf], int t) {
          // blocks 2,3 first: plane t-2 completes and is stored, which frees its registers for plane t+1
          // (four accumulator sets plus four blocks in flight would not fit the 80-register budget)
          const uint32_t ta = wait_d();
          {
            float db[2][kHalf];
            ld_blocks<2>(ta + 48, db);
#pragma unroll
            for (int c = 0; c < kHalf; ++c) { A2[c] += db[0][c]; A3[c] += db[1][c]; }
          }
          if (t >= 2 && valid) store12<H16, RELU, false>(op, vp.o_g, G, half, A3, no_skip);
          op += vp.o_t;
          {
            float db[2][kHalf];
            ld_blocks<2>(ta, db);
            release_d();
#pragma unroll
            for (int c = 0; c < kHalf; ++c) { A0[c] = db[0][c] + bias_r[c]; A1[c] += db[1][c]; }
          }
        };
        auto finish = [&](float (&A3)[kHalf], float (&A2)[kHalf]) {
          if (!valid) return;
          if (T >= 2) store12<H16, RELU, false>(op, vp.o_g, G, half, A3, no_skip);
          store12<H16, RELU, false>(op + vp.o_t, vp.o_g, G, half, A2, no_skip);
        };
        int t = 0;
        for (; t + 4 <= T; t += 4) {
          step(b0, b1, b2, b3, t); step(b1, b2, b3, b0, t + 1); step(b2, b3, b0, b1, t + 2); step(b3, b0, b1, b2, t + 3);
        }
        const int r = T - t;
        if (r == 0) {
          finish(b0, b1);
        } else if (r == 1) {
          step(b0, b1, b2, b3, t);
          finish(b1, b2);
        } else if (r == 2) {
          step(b0, b1, b2, b3, t); step(b1, b2, b3, b0, t + 1);
          finish(b2, b3);
        } else {
          step(b0, b1, b2, b3, t); step(b1, b2, b3, b0, t + 1); step(b2, b3, b0, b1, t + 2);
          finish(b3, b0);
        }
      } else if constexpr (MODE == MODE_K4_OUT1) {
        // cout = 1, taps dt and dw folded into N: column dt*4+dw of row r is the contribution of input
        // column w_r to output plane t_in-dt+1 at w_r-(dw-1).  The w shift crosses lanes, so the 16
        // values of every row go through a padded shared-memory tile; rows 1..125 of a tile are outputs
        // (tiles overlap by 3 rows).  fp32 output is scattered straight into the merged (T,H,W) field,
        // keeping only what merge_data keeps of this chunk.
        __shared__ float xch[2][128][17];
        float a0, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        int t0, h0, w0, kt, kh, kw;
        p.geom.origin(p.chunk_begin + ri.n, t0, h0, w0);
        p.geom.keep_from(p.chunk_begin + ri.n, kt, kh, kw);
        const bool keep_hw = valid && row >= 1 && row <= p.row_step && ri.h >= kh && ri.w >= kw;
        float* dst = p.out_f32 + ((size_t)t0 * p.geom.H + (h0 + ri.h)) * p.geom.W + (w0 + ri.w);
        const size_t plane = (size_t)p.geom.H * p.geom.W;
        const int rm = max(row - 1, 0), rp1 = min(row + 1, 127), rp2 = min(row + 2, 127);
        auto emit = [&](int t, float acc) {
          if (keep_hw && t >= kt) {
            float y = acc + s_bias[0];
            if (RELU) y = fmaxf(y, 0.f);
            dst[(size_t)t * plane] = y;
          }
        };
        for (int t = 0; t < T; ++t) {
          const uint32_t ta = wait_d() - half * kHalf;
          uint32_t v0[8], v1[8];
          umma::tmem_ld8(ta, v0);
          umma::tmem_ld8(ta + 8, v1);
          umma::tmem_ld_wait();
          release_d();
          float (*x)[17] = xch[t & 1];
#pragma unroll
          for (int k = 0; k < 8; ++k) { x[row][k] = __uint_as_float(v0[k]); x[row][8 + k] = __uint_as_float(v1[k]); }
          asm volatile("bar.sync 1, 128;" ::: "memory");
          float s[4];
#pragma unroll
          for (int dt = 0; dt < 4; ++dt)
            s[dt] = ((x[rm][dt * 4 + 0] + x[row][dt * 4 + 1]) + x[rp1][dt * 4 + 2]) + x[rp2][dt * 4 + 3];
          a0 = s[0];
          a1 += s[1];
          a2 += s[2];
          a3 += s[3];
          if (t >= 2) emit(t - 2, a3);
          a3 = a2; a2 = a1; a1 = a0;
        }
        emit(T - 2, a3);
        emit(T - 1, a2);
      } else if (MODE == MODE_DOWN) {
        // t_in = 2*to + dt - 1.  Even t_in = 2m: block0 = dt 1 -> plane m, block1 = dt 3 -> plane m-1 (completes).
        // Odd t_in = 2m+1: block0 = dt 0 -> plane m+1 (starts), block1 = dt 2 -> plane m.
        // pair(X, Y): X holds plane m-1, Y plane m on entry; on exit Y holds plane m and X plane m+1.
        float u[kHalf], v[kHalf];
#pragma unroll
        for (int c = 0; c < kHalf; ++c) { u[c] = 0.f; v[c] = bias_r[c]; }
        auto pair = [&](float (&X)[kHalf], float (&Y)[kHalf], int t) {
          {
            const uint32_t ta = wait_d();
            float db[2][kHalf];
            ld_blocks<2>(ta, db);
            release_d();
#pragma unroll
            for (int c = 0; c < kHalf; ++c) { Y[c] += db[0][c]; X[c] += db[1][c]; }
            const int m = t >> 1;
            if (m >= 1 && valid) store12<H16, RELU, false>(vp.o + (long long)(m - 1) * vp.o_t, vp.o_g, G, half, X, no_skip);
          }
          if (t + 1 < T) {
            const uint32_t ta = wait_d();
            float db[2][kHalf];
            ld_blocks<2>(ta, db);
            release_d();
#pragma unroll
            for (int c = 0; c < kHalf; ++c) { X[c] = db[0][c] + bias_r[c]; Y[c] += db[1][c]; }
          }
        };
        int t = 0;
        bool last_in_v = false;     // the last plane is in the Y array of the last pair() call
        for (; t + 4 <= T; t += 4) { pair(u, v, t); pair(v, u, t + 2); }
        if (t < T) {
          pair(u, v, t);
          last_in_v = true;
          if (t + 2 < T) { pair(v, u, t + 2); last_in_v = false; }
        }
        if (valid) {
          if (last_in_v) store12<H16, RELU, false>(vp.o + (long long)(p.out_T - 1) * vp.o_t, vp.o_g, G, half, v, no_skip);
          else store12<H16, RELU, false>(vp.o + (long long)(p.out_T - 1) * vp.o_t, vp.o_g, G, half, u, no_skip);
        }
      } else if (MODE == MODE_UP) {
        // Row r is input position (h, w); this item's class is the output-line parity ph, and both output-column
        // parities pw are computed here (two accumulators per plane), so that a warp can write runs of
        // consecutive output columns: a thread writing every other 16-byte vector leaves half-filled 32-byte
        // sectors behind, which cost the old one-parity-per-CTA form 1.0 of its 1.3 ms.
        // Block kt of sub-step (t, pw) belongs to output plane 2t + kt - 1: planes 2t-1 and 2t complete at step t.
        // Lane L stores what lane 16k + L/2 computed for pw = L & 1 (k = 0, 1): 32 consecutive columns per store.
        const int odd = lane & 1;
        uint4* ob[2];
        bool vs[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const RowInfo rs = row_info(p, item, q * 32 + 16 * k + (lane >> 1));
          vs[k] = rs.valid;
          const VoxPtr v = vox_ptr(p, rs.n, 2 * rs.h + rs.cls, 2 * rs.w + odd);
          ob[k] = v.o - v.o_t;                  // plane 2t-1
          vp = v;
        }
        float cX[2][kHalf], cY[2][kHalf];
#pragma unroll
        for (int c = 0; c < kHalf; ++c) { cX[0][c] = cX[1][c] = 0.f; cY[0][c] = cY[1][c] = bias_r[c]; }
        auto pack6 = [&](const float (&y)[kHalf], unsigned (&o)[6]) {
#pragma unroll
          for (int i = 0; i < 6; ++i) o[i] = pack2_t<H16, RELU>(y[2 * i], y[2 * i + 1]);
        };
        // the six packed registers of 12 channels: half 0 = vector g0 + low 8 bytes of g1, half 1 = high 8 bytes
        // of g1 + vector g2
        auto store6 = [&](uint4* o, const unsigned (&v)[6]) {
          if (half == 0) {
            o[0] = make_uint4(v[0], v[1], v[2], v[3]);
            if (G > 1) *reinterpret_cast<uint2*>(o + vp.o_g) = make_uint2(v[4], v[5]);
          } else {
            if (G > 1) *(reinterpret_cast<uint2*>(o + vp.o_g) + 1) = make_uint2(v[0], v[1]);
            if (G > 2) o[2 * vp.o_g] = make_uint4(v[2], v[3], v[4], v[5]);
          }
        };
        auto store_transposed = [&](const unsigned (&P0)[6], const unsigned (&P1)[6], long long plane_off, bool on) {
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const int src = 16 * k + (lane >> 1);
            unsigned v[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) {
              const unsigned x0 = __shfl_sync(0xffffffffu, P0[i], src);
              const unsigned x1 = __shfl_sync(0xffffffffu, P1[i], src);
              v[i] = odd ? x1 : x0;
            }
            if (on && vs[k]) store6(ob[k] + plane_off, v);
          }
        };
        for (int t = 0; t < T; ++t) {
          unsigned PX[2][6], PY[2][6];
#pragma unroll
          for (int pw = 0; pw < 2; ++pw) {
            const uint32_t ta = wait_d(pw == 1);
            {
              float db[2][kHalf];
              ld_blocks<2>(ta, db);
#pragma unroll
              for (int c = 0; c < kHalf; ++c) { db[0][c] += cX[pw][c]; db[1][c] += cY[pw][c]; }
              pack6(db[0], PX[pw]);
              pack6(db[1], PY[pw]);
            }
            {
              float db[2][kHalf];
              ld_blocks<2>(ta + 48, db);
              release_d();
#pragma unroll
              for (int c = 0; c < kHalf; ++c) { cX[pw][c] = db[0][c] + bias_r[c]; cY[pw][c] = db[1][c] + bias_r[c]; }
            }
          }
          store_transposed(PX[0], PX[1], 0, t >= 1);
          store_transposed(PY[0], PY[1], vp.o_t, true);
          ob[0] += 2 * vp.o_t;
          ob[1] += 2 * vp.o_t;
        }
        {
          unsigned PX[2][6];
          pack6(cX[0], PX[0]);
          pack6(cX[1], PX[1]);
          store_transposed(PX[0], PX[1], 0, true);
        }
      }
    }
  }
  umma::tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) umma::tmem_dealloc(tmem, (uint32_t)p.tmem_cols);
}

// ------------------------------------------------------------------------------------------
// sweep_acc_kernel: the stride-1 layers with the time taps ACCUMULATED IN TMEM (second-generation form of the
// K3 / K4 modes above).
//
// The plane-sweep MMA is unchanged -- D[voxel, (dt, co)] += X[t_in][voxel + tap][ci] * W[dt][tap][ci][co], N = NT
// time blocks of 24 columns -- but the NT blocks no longer land in a fresh accumulator that the epilogue folds in
// registers.  A bank of NT*24 TMEM columns holds the NT output planes currently under construction, plane p in
// slot p mod NT, and the MMA of input plane t adds block dt straight onto output plane t - dt + 1: the weight
// image stores the time blocks twice in decreasing order ([dt=NT-1 .. 0, NT-1 .. 1]) and the B descriptor of
// input plane t starts ((NT-2-t) mod NT) blocks in, which rotates the blocks onto the right slots.  The epilogue
// thread of a finished plane reads its 24 columns, writes the bias back into them (the slot's next plane starts
// from the bias, so the bias add is free and every MMA accumulates) and hands the bank back; conversion, packing
// and the global stores happen after that.  What this buys over the register fold:
//   * a third of the TMEM reads, no accumulator adds, no register rotation: ~60 instead of ~180 instructions per
//     voxel plane, and one thread owns all 24 channels of a voxel, so every store is a full 16-byte vector per lane
//     (512 contiguous bytes per warp) -- the old 12-channel halves wrote half-filled 32-byte sectors;
//   * one persistent CTA per SM with kAccBanks tiles in flight (bank b = tile 4r+b of the CTA, input planes
//     interleaved plane by plane): while the epilogue of one bank drains, the tensor pipe works on the other three,
//     the weights sit in shared memory once per SM, and 8-12 stages of input planes are in flight.
// Warps: 0 producer (the whole warp: lane 8*b + c issues bulk copy c of bank b's stage), 1-2 MMA issuers (banks 0/2 and
// 1/3), 3 TMEM allocation, 4.. epilogue (bank = (warp-4)/4, TMEM lane quarter = warp % 4).  Every bank owns its own
// ring of input stages.  Deterministic: fixed accumulation order per output voxel.
#ifndef LC_ACC_BANKS
#define LC_ACC_BANKS 4
#endif
constexpr int kAccBanks = LC_ACC_BANKS;
constexpr int kAccThreads = 32 * (4 + 4 * kAccBanks);
constexpr int kAccMaxStages = 12;

// NCTA = 2: the two CTAs of an SM pair (cluster of 2) run in lockstep on their own four tiles each, and ONE
// tcgen05.mma.cta_group::2 of the leader CTA multiplies both A tiles (M = 256: rows 0-127 from the leader's shared
// memory into the leader's TMEM lanes, rows 128-255 the peer's) with a B operand of which each CTA holds HALF the
// columns.  The kernel is bound by the shared-memory operand fetch (4 KB of A + 2.5 KB of B per M128 x N80 x K16
// step); halving B takes 19 % off it.  Everything else stays per CTA: producer, stages, epilogue, stores.  Barriers:
//   full[s]    leader: own bulk copies (expect_tx) + one arrival relayed from the peer (its issuer warp 1 waits for
//              its own full[s] and arrives remotely);  peer: own copies only
//   tfull[b]   committed by the leader into BOTH CTAs (multicast commit)
//   tempty[b]  leader's: 4 local + 4 remote epilogue warps
//   empty[s]   local (released by the local epilogue once tfull of the step that read the stage has fired)
// The peer's weight image is the leader's shifted by N/2 rows, so the same descriptor offsets address each half.
template <int NT, int H16, int RELU, int NCTA>
__global__ void __launch_bounds__(kAccThreads, 1)
sweep_acc_kernel(const __grid_constant__ SweepParams p) {
  constexpr int kBankCols = (NT * 24 + 15) & ~15;       // = the MMA's N (80 for NT = 3: 8 scratch columns)
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t full[kAccMaxStages], empty[kAccMaxStages], tfull[kAccBanks], tempty[kAccBanks], wbar;
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_bias[24];
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const int rank = (NCTA == 2) ? (int)umma::cluster_ctarank() : 0;
  const bool leader = rank == 0;
  const int slab_bytes = p.slab_slots * 16;
  const int skip_bytes = p.skip_mma ? p.skip.G * p.skip_slots * 16 : 0;
  const int stage_bytes = p.in_NPG * slab_bytes + skip_bytes;
  const int stages = (NCTA == 2) ? p.stages2c : p.stages2;
  // Every bank owns its own ring of spb stages (stage = bank + kAccBanks * (kb % spb) for the bank's step kb, parity
  // (kb / spb) & 1).  A ring shared by the banks is unsafe with parity waits: the two issuer warps each wait only for
  // the stages of their own banks, so a stage's previous use may belong to the other issuer's bank, and when that load
  // completes late (memory contention from a kernel of another stream) the barrier is still a whole phase behind, its
  // parity equals that of the awaited phase's predecessor and the wait passes on stale data.
  const int spb = stages / kAccBanks;
  const int wbytes = (NCTA == 2) ? p.wimg2c_bytes : p.wimg2_bytes;
  uint8_t* sB = smem;
  uint8_t* sA = smem + ((wbytes + 127) & ~127);
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) { umma::mbar_init(&full[i], (NCTA == 2 && leader) ? 2 : 1); umma::mbar_init(&empty[i], 1); }
    for (int i = 0; i < kAccBanks; ++i) { umma::mbar_init(&tfull[i], 1); umma::mbar_init(&tempty[i], 4 * NCTA); }
    umma::mbar_init(&wbar, 1);
    umma::fence_barrier_init();
  }
  if (threadIdx.x < 24) s_bias[threadIdx.x] = (threadIdx.x < p.cout) ? p.bias[threadIdx.x] : 0.f;
  if (warp == 3) {
    if (NCTA == 2) umma::tmem_alloc2(&tmem_base_s, 512u);
    else umma::tmem_alloc(&tmem_base_s, 512u);
  }
  umma::tcgen05_fence_before();
  if (NCTA == 2) umma::cluster_sync_all();     // the peer's barriers exist before anything arrives on them
  else __syncthreads();
  umma::tcgen05_fence_after();
  const uint32_t tmem = tmem_base_s;
  const int T = p.in_T;
  // Work items: unit u = blockIdx.x / NCTA of gridDim.x / NCTA units takes item groups u, u + units, ...; group j is
  // items NCTA*j + rank (the last group of an odd item count has no item for the peer: it re-runs the last item with
  // its stores off).  Round r works on groups 4r .. 4r+3 (one per bank).
  const int unit = (int)blockIdx.x / NCTA, units = (int)gridDim.x / NCTA;
  const int n_groups = (p.n_items + NCTA - 1) / NCTA;
  const int cnt = (unit < n_groups) ? (n_groups - unit + units - 1) / units : 0;
  const int rounds = (cnt + kAccBanks - 1) / kAccBanks;
  auto item_of = [&](int k) { return NCTA * (unit + k * units) + rank; };

  if (warp == 0) {
    // ===================== producer =====================
    // Whole warp: lane 8*b + c issues copy c of bank b's stage (input slabs, then skip tiles; at most 8 copies per
    // stage, checked by the host), lane 8*b also waits for the stage and posts the byte count -- the four banks'
    // stages of one input plane are filled by one pass of the loop instead of four passes of a single thread.
    static_assert(kAccBanks <= 4, "producer lanes are laid out for at most four banks");
    if (lane == 0) {
      umma::mbar_arrive_expect_tx(&wbar, (uint32_t)wbytes);
      umma::bulk_g2s(sB, (NCTA == 2) ? (leader ? p.wimg2c0 : p.wimg2c1) : p.wimg2, (uint32_t)wbytes, &wbar);
    }
    const int ncop = p.in_NPG + (p.skip_mma ? p.skip.G : 0);
    const int pb = lane >> 3, pc = lane & 7;
    const bool is_in = pc < p.in_NPG;
    const int cs = pc - p.in_NPG;
    const uint32_t my_bytes = is_in ? (uint32_t)slab_bytes : (uint32_t)(p.skip_slots * 16);
    const int my_dst = is_in ? pc * slab_bytes : p.in_NPG * slab_bytes + cs * p.skip_slots * 16;
    const long long my_tstride = is_in ? (long long)p.in_NPG * p.in_S : (long long)p.skip.G * p.skip.S;
    for (int r = 0; r < rounds; ++r) {
      const int nb = min(kAccBanks, cnt - kAccBanks * r);
      const bool act = pb < nb && pc < ncop;
      const int item = min(item_of(kAccBanks * r + min(pb, nb - 1)), p.n_items - 1);
      const int n = item / p.tiles_per_chunk;
      const int tile = item - n * p.tiles_per_chunk;
      const uint4* src;
      if (is_in)
        src = p.in + (long long)n * p.in_cs + p.in_GL + p.row0 + tile * p.row_step - p.row_shift + p.slab_start +
              (long long)pc * p.in_S;
      else
        src = reinterpret_cast<const uint4*>(p.skip.base) + (long long)n * p.skip.chunk_stride + p.skip.GL + p.row0 +
              tile * 128 - p.skip_pad + (long long)cs * p.skip.S;
      for (int t = 0; t < T; ++t) {
        const int kb = r * T + t;
        const int stage = pb + kAccBanks * (kb % spb), phase = (kb / spb) & 1;
        if (act && pc == 0) {
          umma::mbar_wait(&empty[stage], phase ^ 1);
          if (p.debug & 8) umma::mbar_arrive(&full[stage]);
          else umma::mbar_arrive_expect_tx(&full[stage], (uint32_t)stage_bytes);
        }
        __syncwarp();
        if (act && !(p.debug & 8))
          umma::bulk_g2s(sA + stage * stage_bytes + my_dst, src + (long long)t * my_tstride, my_bytes, &full[stage]);
      }
    }
  } else if (NCTA == 2 && !leader && warp == 1) {
    // ===================== peer CTA: relay "my stage is loaded" to the leader's full barrier =====================
    if (lane == 0) {
      umma::mbar_wait(&wbar, 0);                 // the first relayed arrival also vouches for this CTA's weights
      for (int r = 0; r < rounds; ++r) {
        const int nb = min(kAccBanks, cnt - kAccBanks * r);
        for (int t = 0; t < T; ++t) {
          const int kb = r * T + t;
          for (int b = 0; b < nb; ++b) {
            const int stage = b + kAccBanks * (kb % spb);
            umma::mbar_wait(&full[stage], (kb / spb) & 1);
            umma::mbar_arrive_cluster(umma::mapa_rank(&full[stage], 0));
          }
        }
      }
    }
  } else if ((warp == 1 || warp == 2) && leader) {
    // ===================== MMA issuers (alternate steps) =====================
    const int me = warp - 1;
    umma::mbar_wait(&wbar, 0);
    const uint32_t idesc = umma::make_idesc(128 * NCTA, kBankCols, H16 ? 0 : 1);
    const uint64_t sA_add = (uint64_t)(umma::smem_u32(sA) >> 4), sB_add = (uint64_t)(umma::smem_u32(sB) >> 4);
    for (int r = 0; r < rounds; ++r) {
      const int nb = min(kAccBanks, cnt - kAccBanks * r);
      for (int t = 0; t < T; ++t) {
        // start block of the rotated weight image: block i holds dt = (NT-1-i) mod NT; slot s must receive
        // dt = (t+1-s) mod NT  =>  i0 = (NT-2-t) mod NT
        const int i0 = (((NT - 2 - t) % NT) + NT) % NT;
        const uint64_t b_add = sB_add + (uint64_t)(i0 * 24);      // 24 rows of 16 bytes, in 16-byte units
        const int kb = r * T + t;                                 // this bank-step's index (all banks alike)
        for (int b = 0; b < nb; ++b) {
          const int stage = b + kAccBanks * (kb % spb), phase = (kb / spb) & 1;
          // a bank always belongs to the same issuer: its steps must be issued in order, each after the epilogue
          // of the one before (alternating issuers by step would let one run a whole phase ahead of the other on
          // the same bank whenever the number of active banks is odd)
          if ((b & 1) == me) {
            if (NCTA == 2) {
              umma::mbar_wait_cluster(&tempty[b], kb & 1);
              umma::mbar_wait_cluster(&full[stage], phase);
            } else {
              umma::mbar_wait(&tempty[b], kb & 1);
              umma::mbar_wait(&full[stage], phase);
            }
            umma::tcgen05_fence_after();
            const uint64_t a_add = sA_add + (uint64_t)((uint32_t)(stage * stage_bytes) >> 4);
            const uint32_t d_addr = tmem + b * kBankCols;
            const int nks = (p.debug & 1) ? 0 : p.nks;
#pragma unroll 2
            for (int ks = 0; ks < nks; ++ks) {
              const uint64_t a = p.a_desc[0][ks] + a_add, bd = ((NCTA == 2) ? p.b2c_desc[ks] : p.b2_desc[ks]) + b_add;
              if (umma::elect_one()) {
                if (NCTA == 2) umma::mma_f16_2cta(d_addr, a, bd, idesc, true);
                else umma::mma_f16(d_addr, a, bd, idesc, true);
              }
            }
            __syncwarp();
            if (umma::elect_one()) {
              if (NCTA == 2) umma::commit_2cta(&tfull[b]);
              else umma::commit(&tfull[b]);
            }
            __syncwarp();
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int b = (warp - 4) >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t bank_addr = tmem + ((uint32_t)(q * 32) << 16) + b * kBankCols;
    const uint32_t tempty_leader = (NCTA == 2) ? umma::mapa_rank(&tempty[b], 0) : 0u;
    uint32_t bias_u[24];
#pragma unroll
    for (int c = 0; c < 24; ++c) bias_u[c] = __float_as_uint(s_bias[c]);
    auto init_slot = [&](int slot) {
      umma::tmem_st8(bank_addr + slot * 24, bias_u);
      umma::tmem_st8(bank_addr + slot * 24 + 8, bias_u + 8);
      umma::tmem_st8(bank_addr + slot * 24 + 16, bias_u + 16);
    };
    auto hand_back = [&]() {
      umma::tmem_st_wait();
      umma::tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (NCTA == 2) umma::mbar_arrive_cluster(tempty_leader);
        else umma::mbar_arrive_relaxed(&tempty[b]);
      }
    };
    // 24 accumulator columns of one slot -> 12 packed 16-bit pairs ((ReLU,) saturate, round to nearest even)
    auto load_pack = [&](int slot, unsigned (&pk)[12]) {
      uint32_t v[24];
      umma::tmem_ld8(bank_addr + slot * 24, v);
      umma::tmem_ld8(bank_addr + slot * 24 + 8, v + 8);
      umma::tmem_ld8(bank_addr + slot * 24 + 16, v + 16);
      umma::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 12; ++i) pk[i] = pack2_t<H16, RELU>(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
    };
    const int G = p.out.G;
    // Output staging (p.tma_store): [bank][2][3 groups][128 rows] 16-byte vectors behind the input stages.  A warp
    // only touches its own 32 rows, so the hand-over to its bulk copies needs no block barrier.
    uint4* const obuf = reinterpret_cast<uint4*>(sA + spb * kAccBanks * stage_bytes) + (b * 2) * (3 * 128);
    int opar = 0;
    // every slot starts from the bias
#pragma unroll
    for (int s = 0; s < NT; ++s) init_slot(s);
    hand_back();
    for (int r = 0; kAccBanks * r + b < cnt; ++r) {
      const int nb = min(kAccBanks, cnt - kAccBanks * r);
      const int item_raw = item_of(kAccBanks * r + b);
      const int item = min(item_raw, p.n_items - 1);
      const RowInfo ri = row_info(p, item, row);
      const bool valid = ri.valid && item_raw < p.n_items && !(p.debug & 4);   // debug bit 2: no global stores
      const VoxPtr vp = vox_ptr(p, ri.n, ri.h, ri.w);
      // first output vector of this warp's 32 rows in plane 0, group 0 (tma_store: output slot == row slot)
      uint4* const warp_o = reinterpret_cast<uint4*>(p.out.base) + (long long)(item / p.tiles_per_chunk) * p.out.chunk_stride +
                            p.out.GL + p.row0 + (item % p.tiles_per_chunk) * 128 + q * 32;
      const bool item_ok = item_raw < p.n_items && !(p.debug & 4);
      auto store_plane = [&](int plane, const unsigned (&pk)[12]) {
        if (p.tma_store) {
          // rows outside the image are halo / guard slots of the output plane: they hold zeros and get zeros
          if (lane == 0) umma::bulk_wait_read<1>();      // the copies that read this staging buffer two planes ago
          __syncwarp();
          uint4* buf = obuf + opar * (3 * 128) + row;
          const uint4 z = make_uint4(0u, 0u, 0u, 0u);
          buf[0] = ri.valid ? make_uint4(pk[0], pk[1], pk[2], pk[3]) : z;
          if (G > 1) buf[128] = ri.valid ? make_uint4(pk[4], pk[5], pk[6], pk[7]) : z;
          if (G > 2) buf[256] = ri.valid ? make_uint4(pk[8], pk[9], pk[10], pk[11]) : z;
          umma::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (item_ok) {
              const uint4* src = obuf + opar * (3 * 128) + q * 32;
              uint4* dst = warp_o + (long long)plane * G * p.out.S;
              for (int g = 0; g < G; ++g) umma::bulk_s2g(dst + (long long)g * p.out.S, src + g * 128, 512u);
            }
            umma::bulk_commit();
          }
          opar ^= 1;
          return;
        }
        if (!valid) return;
        uint4* o = vp.o + (long long)plane * vp.o_t;
        o[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        if (G > 1) o[vp.o_g] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        if (G > 2) o[2 * vp.o_g] = make_uint4(pk[8], pk[9], pk[10], pk[11]);
      };
      for (int t = 0; t < T; ++t) {
        const int kb = r * T + t;
        if (NCTA == 2) umma::mbar_wait_cluster(&tfull[b], kb & 1);
        else umma::mbar_wait(&tfull[b], kb & 1);
        umma::tcgen05_fence_after();
        // the MMAs that read this step's stage are complete: hand the stage back to the producer
        if (q == 0 && lane == 0) umma::mbar_arrive_relaxed(&empty[b + kAccBanks * (kb % spb)]);
        const int pc = t - (NT - 2);                              // the plane that completed with this step
        if (t + 1 < T) {
          unsigned pk[12];
          const int slot = (pc + NT) % NT;
          if (pc >= 0) load_pack(slot, pk);
          init_slot(slot);
          hand_back();
          if (pc >= 0) store_plane(pc, pk);
        } else {
          // last input plane: the remaining NT-1 planes are complete; all slots restart from the bias
          unsigned pk[NT - 1][12];
#pragma unroll
          for (int j = 0; j < NT - 1; ++j)
            if (pc + j >= 0) load_pack((pc + j + NT) % NT, pk[j]);
#pragma unroll
          for (int s = 0; s < NT; ++s) init_slot(s);
          hand_back();
#pragma unroll
          for (int j = 0; j < NT - 1; ++j)
            if (pc + j >= 0) store_plane(pc + j, pk[j]);
        }
      }
    }
  }
  if (p.tma_store && warp >= 4 && lane == 0) umma::bulk_wait_all();   // staging is read, outputs are written
  umma::tcgen05_fence_before();
  if (NCTA == 2) umma::cluster_sync_all();      // no CTA frees TMEM or exits while its partner still uses the pair
  else __syncthreads();
  if (warp == 3) {
    if (NCTA == 2) umma::tmem_dealloc2(tmem, 512u);
    else umma::tmem_dealloc(tmem, 512u);
  }
}

// ------------------------------------------------------------------------------------------
// plan: tensors of one chunk block
// ------------------------------------------------------------------------------------------
enum { LAYOUT_PLAIN = 0, LAYOUT_PARITY = 1, LAYOUT_WFOLD = 2 };

struct TensorPlan {
  int fmt;                 // FMT_F32 / FMT_GP16
  int layout;
  int T, H, W, C, G, NP, PW, PH, S, GL;
  long long offset_vecs;   // offset inside the chunk block, in 16-byte vectors
  long long size_vecs;
};

struct LayerPlan {
  int backend;             // 0 = direct, else MODE_*
  float* d_w16;            // fp32 weights rounded through the 16-bit operand format [tap][cin][cout]
  SweepParams sp;          // template (pointers filled per call)
  int smem_bytes;
  int acc_smem_bytes;      // > 0: the layer runs sweep_acc_kernel (time taps accumulated in TMEM)
  void* d_wimg2;
  int acc2_smem_bytes;     // > 0: the CTA-pair form of sweep_acc_kernel can run the layer
  void* d_wimg2c[2];
};

struct TcPlan {
  int n_layers;
  TensorPlan in_enc;                    // chunk input: fp32 (direct first layer) or w-folded GP16
  TensorPlan out[kMaxLayers];           // output tensor of every layer
  LayerPlan layer[kMaxLayers];
  long long block_vecs;                 // 16-byte vectors per chunk block
  int wfold_k;
};

TensorPlan make_gp(int T, int H, int W, int C, int pad_hi, int layout) {
  TensorPlan t;
  memset(&t, 0, sizeof(t));
  t.fmt = FMT_GP16;
  t.layout = layout;
  t.T = T; t.H = H; t.W = W; t.C = C;
  t.G = (C + 7) / 8;
  t.GL = 8;
  if (layout == LAYOUT_PARITY) {
    t.NP = 4;
    t.PW = (W + 2) / 2;
    t.PH = (H + 2) / 2;
  } else {
    t.NP = 1;
    t.PW = W + 1 + pad_hi;
    t.PH = H + 1 + pad_hi;
  }
  // guard: a tile slab may over-read up to 128 + 3*(PW+1) vectors past the last interior line
  t.S = t.GL + t.PH * t.PW + 128 + 3 * (t.PW + 1);
  t.S = (t.S + 7) & ~7;
  t.size_vecs = (long long)T * t.NP * t.G * t.S;
  return t;
}

// chunk input folded along w for a k-tap first layer: vector (h,w) = { x[h][w+dw-pad][ci] }, pitch W
TensorPlan make_wfold(int T, int H, int W, int C, int k, int pad_lo) {
  TensorPlan t;
  memset(&t, 0, sizeof(t));
  t.fmt = FMT_GP16;
  t.layout = LAYOUT_WFOLD;
  t.T = T; t.H = H; t.W = W; t.C = k * C;
  t.G = (k * C + 7) / 8;
  t.NP = 1;
  t.GL = 8;
  t.PW = W;
  t.PH = H + k - 1;                 // pad_lo lines above, k-1-pad_lo below
  t.S = t.GL + t.PH * t.PW + 128 + 16;
  t.S = (t.S + 7) & ~7;
  t.size_vecs = (long long)T * t.G * t.S;
  (void)pad_lo;
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

unsigned short to_h16_host(float v, int half) {
  unsigned short u;
  if (half) { __half hv = __float2half_rn(v); memcpy(&u, &hv, 2); }
  else { __nv_bfloat16 bv = __float2bfloat16_rn(v); memcpy(&u, &bv, 2); }
  return u;
}

// which tensor-core kernel (if any) can run layer i
int tc_mode_of(const lc_codec* h, int i) {
  if (getenv("LOSSYCOMP_NO_TC")) return 0;
  const DevLayer& L = h->layers[i];
  const int nl = h->n_enc + h->n_dec;
  const bool full = (L.in_t == CT && L.in_h == CH && L.in_w == CW);
  const char* only = getenv("LOSSYCOMP_TC_MODES");   // debugging: comma list of mode numbers
  auto allowed = [&](int m) { return !only || strchr(only, '0' + m) != nullptr; };
  if (L.kind == 0 && L.k == 3 && L.stride == 1 && L.cin > 16 && L.cin <= 24 && L.cout > 16 && L.cout <= 24 &&
      full && i != 0 && i != h->n_enc && i != h->n_enc - 1 && i != nl - 1)
    return allowed(MODE_K3) ? MODE_K3 : 0;
  if (L.kind == 0 && L.k == 4 && L.stride == 1 && full) {
    if (i == 0 && 4 * L.cin <= 24 && L.cout > 16 && L.cout <= 24 && h->n_enc > 1)
      return allowed(MODE_K4) ? MODE_K4 : 0;
    if (i == nl - 1 && L.cout == 1 && L.cin > 16 && L.cin <= 24 && i != h->n_enc)
      return allowed(MODE_K4_OUT1) ? MODE_K4_OUT1 : 0;
  }
  // stride-2 conv: inner layers (24-channel group planes) and the FIRST layer of the models without a ResBlock
  // (results/models/{basic,coords,landsea}_model: 1 / 5 / 2 input channels in one 8-channel group, written in the
  // parity-split layout by the chunk gather)
  if (L.kind == 0 && L.k == 4 && L.stride == 2 && L.cout > 16 && L.cout <= 24 &&
      ((i != 0 && L.cin > 16 && L.cin <= 24) || (i == 0 && L.cin <= 8 && full)) &&
      L.in_t % 2 == 0 && L.in_h % 2 == 0 && L.in_w % 2 == 0 && L.in_h >= 12 && i != h->n_enc - 1)
    return allowed(MODE_DOWN) ? MODE_DOWN : 0;
  if (L.kind == 1 && L.k == 4 && L.stride == 2 && L.cin > 16 && L.cin <= 24 && L.cout > 16 && L.cout <= 24 &&
      L.in_h >= 6 && i != h->n_enc && i != nl - 1)
    return allowed(MODE_UP) ? MODE_UP : 0;
  return 0;
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

// K2: standardise channel 0 and gather chunks into fp32 chunk blocks (direct first layer)
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

// K2 (stride-2 tensor-core first layer): standardise and gather into the 16-bit group-plane layout the layer reads
// (parity-split for MODE_DOWN); one thread per voxel and channel group, channels beyond C are zero
__global__ void gather_gp_kernel(const float* __restrict__ x, ChunkGeom g, int C, float mean, float stdv,
                                 int chunk_begin, int n_chunks, ActView out) {
  const long long total = (long long)n_chunks * CT * CH * CW * out.G;
  uint4* o = reinterpret_cast<uint4*>(out.base);
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < total;
       v += (long long)gridDim.x * blockDim.x) {
    const int gi = (int)(v % out.G);
    long long r = v / out.G;
    const int lw = (int)(r % CW);
    r /= CW;
    const int lh = (int)(r % CH);
    r /= CH;
    const int lt = (int)(r % CT);
    const int chunk = (int)(r / CT);
    int t0, h0, w0;
    g.origin(chunk_begin + chunk, t0, h0, w0);
    const float* src = x + (((size_t)(t0 + lt) * g.H + (h0 + lh)) * g.W + (w0 + lw)) * C;
    float f[8];
#pragma unroll
    for (int k8 = 0; k8 < 8; ++k8) {
      const int c = gi * 8 + k8;
      float val = 0.f;
      if (c < C) {
        val = src[c];
        if (c == 0) val = __fdiv_rn(__fsub_rn(val, mean), stdv);
      }
      f[k8] = val;
    }
    o[gp_vec_index(out, chunk, lt, lh, lw, gi)] = pack8(f, out.half);
  }
}

// K2 (tensor-core first layer): standardise, gather and fold the k w-taps into pseudo-channels:
// vector g of voxel (h,w) holds pseudo-channels pc = dw*C + ci = x[h][w + dw - pad][ci] (0 outside)
__global__ void gather_wfold_kernel(const float* __restrict__ x, ChunkGeom g, int C, float mean, float stdv,
                                    int chunk_begin, int n_chunks, ActView out, int k, int pad_lo) {
  const long long total = (long long)n_chunks * CT * CH * CW * out.G;
  uint4* o = reinterpret_cast<uint4*>(out.base);
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < total;
       v += (long long)gridDim.x * blockDim.x) {
    const int lw = (int)(v % CW);
    long long r = v / CW;
    const int lh = (int)(r % CH);
    r /= CH;
    const int gi = (int)(r % out.G);
    r /= out.G;
    const int lt = (int)(r % CT);
    const int chunk = (int)(r / CT);
    int t0, h0, w0;
    g.origin(chunk_begin + chunk, t0, h0, w0);
    const float* line = x + (((size_t)(t0 + lt) * g.H + (h0 + lh)) * g.W + w0) * C;
    float f[8];
#pragma unroll
    for (int k8 = 0; k8 < 8; ++k8) {
      const int pc = gi * 8 + k8;
      const int dw = pc / C, ci = pc - dw * C;
      const int w = lw + dw - pad_lo;
      float val = 0.f;
      if (dw < k && w >= 0 && w < CW) {
        val = line[(size_t)w * C + ci];
        if (ci == 0) val = __fdiv_rn(__fsub_rn(val, mean), stdv);
      }
      f[k8] = val;
    }
    // slot = (h + pad_lo) * PW + w   (no halo columns: the w taps are inside the vector)
    const long long idx = (long long)chunk * out.chunk_stride + ((long long)lt * out.G + gi) * out.S + out.GL +
                          (lh + pad_lo) * out.PW + lw;
    o[idx] = pack8(f, out.half);
  }
}

// fast path of gather_wfold_kernel for the shipped models' input (C = 2, k = 4, one group): one thread per
// run of four voxels along w -- seven 8-byte loads and seven IEEE divisions feed four 16-byte stores (one
// thread per voxel standardised every input four times and was bound by the divisions)
__global__ void __launch_bounds__(256) gather_wfold_c2k4_kernel(const float* __restrict__ x, ChunkGeom g, float mean,
                                                                float stdv, int chunk_begin, int n_chunks, ActView out,
                                                                int pad_lo) {
  constexpr int kRun = 4;
  static_assert(CW % kRun == 0, "chunk width must be a multiple of the run length");
  const long long total = (long long)n_chunks * CT * CH * (CW / kRun);
  uint4* o = reinterpret_cast<uint4*>(out.base);
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < total;
       v += (long long)gridDim.x * blockDim.x) {
    const int lw0 = (int)(v % (CW / kRun)) * kRun;
    const int r1 = (int)(v / (CW / kRun));
    const int lh = r1 % CH;
    const int r2 = r1 / CH;
    const int lt = r2 % CT;
    const int chunk = r2 / CT;
    int t0, h0, w0;
    g.origin(chunk_begin + chunk, t0, h0, w0);
    const float2* line = reinterpret_cast<const float2*>(x) + ((size_t)(t0 + lt) * g.H + (h0 + lh)) * g.W + w0;
    float2 in[kRun + 3];
#pragma unroll
    for (int j = 0; j < kRun + 3; ++j) {
      const int w = lw0 + j - pad_lo;
      float2 val = make_float2(0.f, 0.f);
      if (w >= 0 && w < CW) {
        val = __ldg(line + w);
        val.x = __fdiv_rn(__fsub_rn(val.x, mean), stdv);
      }
      in[j] = val;
    }
    const long long idx = (long long)chunk * out.chunk_stride + (long long)lt * out.S + out.GL +
                          (lh + pad_lo) * out.PW + lw0;
#pragma unroll
    for (int r = 0; r < kRun; ++r) {
      float f[8];
#pragma unroll
      for (int dw = 0; dw < 4; ++dw) { f[2 * dw] = in[r + dw].x; f[2 * dw + 1] = in[r + dw].y; }
      o[idx + r] = pack8(f, out.half);
    }
  }
}

// Resident CTAs per SM: two (the epilogue warps of one CTA cover the MMA / barrier latency of the other; 80
// registers, half the TMEM columns, 110 KB of stages each) except for the stride-2 mode, whose 74 KB weight
// image plus two 31 KB parity-split stages need a whole SM, and the transposed mode, whose epilogue keeps
// the accumulators of both output-column parities in registers.
// LOSSYCOMP_TC_ONE_PER_SM=<comma list of mode numbers> forces one for more modes (experiments).
int ctas_per_sm(int mode) {
  static int one_mask = -1;
  if (one_mask < 0) {
    one_mask = (1 << MODE_DOWN) | (1 << MODE_UP);
    if (const char* e = getenv("LOSSYCOMP_TC_ONE_PER_SM"))
      for (const char* c = e; *c; ++c)
        if (*c >= '1' && *c <= '5') one_mask |= 1 << (*c - '0');
  }
  return (one_mask >> mode) & 1 ? 1 : 2;
}

template <int MODE, int H16, int RELU>
int launch_variant(const SweepParams& p, int grid, int smem, cudaStream_t s) {
  // the attribute belongs to the (function, device) pair: one flag per device, not per process
  static bool attr_done[64] = {};
  int dev = 0;
  LC_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr_done[dev]) {
    LC_CUDA(cudaFuncSetAttribute(sweep_tc_kernel<MODE, H16, RELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    if (dev >= 0 && dev < 64) attr_done[dev] = true;
  }
  sweep_tc_kernel<MODE, H16, RELU><<<grid, kTcThreads, smem, s>>>(p);
  return LC_OK;
}

template <int MODE>
int launch_mode(const lc_codec* h, const SweepParams& p, int grid, int smem, cudaStream_t s) {
  (void)h;
  if (p.half) return p.relu ? launch_variant<MODE, 1, 1>(p, grid, smem, s) : launch_variant<MODE, 1, 0>(p, grid, smem, s);
  return p.relu ? launch_variant<MODE, 0, 1>(p, grid, smem, s) : launch_variant<MODE, 0, 0>(p, grid, smem, s);
}

template <int NT, int H16, int RELU>
int launch_acc_variant(const SweepParams& p, int grid, int smem, cudaStream_t s) {
  static bool attr_done[64] = {};
  int dev = 0;
  LC_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr_done[dev]) {
    LC_CUDA(cudaFuncSetAttribute(sweep_acc_kernel<NT, H16, RELU, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    if (dev >= 0 && dev < 64) attr_done[dev] = true;
  }
  sweep_acc_kernel<NT, H16, RELU, 1><<<grid, kAccThreads, smem, s>>>(p);
  return LC_OK;
}

// CTA-pair form: clusters of two CTAs (one SM pair each); the grid is twice the number of pairs that can be resident
template <int NT, int H16, int RELU>
int launch_acc2_variant(const SweepParams& p, int n_items, int smem, cudaStream_t s) {
  static int max_pairs[64] = {};
  int dev = 0;
  LC_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) {
    lc_set_error("device index %d out of range", dev);
    return LC_EINVAL;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3(kAccThreads);
  cfg.dynamicSmemBytes = (size_t)smem;
  cfg.stream = s;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (!max_pairs[dev]) {
    LC_CUDA(cudaFuncSetAttribute(sweep_acc_kernel<NT, H16, RELU, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    cfg.gridDim = dim3(2);
    int n = 0;
    LC_CUDA(cudaOccupancyMaxActiveClusters(&n, sweep_acc_kernel<NT, H16, RELU, 2>, &cfg));
    if (n < 1) {
      lc_set_error("no CTA pair of sweep_acc_kernel fits the device");
      return LC_EUNSUPPORTED;
    }
    max_pairs[dev] = n;
    if (getenv("LOSSYCOMP_ACC2_PAIRS")) max_pairs[dev] = atoi(getenv("LOSSYCOMP_ACC2_PAIRS"));
    if (getenv("LOSSYCOMP_TC_VERBOSE")) fprintf(stderr, "[lossycomp] sweep_acc_kernel<%d> CTA pairs: occupancy %d, using %d\n", NT, n, max_pairs[dev]);
  }
  const int groups = (n_items + 1) / 2;
  const int pairs = max_pairs[dev] < groups ? max_pairs[dev] : groups;
  cfg.gridDim = dim3(2 * pairs);
  LC_CUDA(cudaLaunchKernelEx(&cfg, sweep_acc_kernel<NT, H16, RELU, 2>, p));
  return LC_OK;
}

// LOSSYCOMP_ACC_CTA2=0 keeps these layers on the single-CTA form.  Measured on the benchmark field (with the warp-wide
// producer): 3x3x3 layers 0.897 -> 0.852 ms, with a residual input 1.18 -> 0.98 ms, first layer 0.53 -> 0.54 ms.
bool acc_cta2_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("LOSSYCOMP_ACC_CTA2");
    on = e ? (atoi(e) != 0) : 1;
  }
  return on != 0;
}

int launch_sweep(const lc_codec* h, const LayerPlan& lp, SweepParams& p, cudaStream_t s) {
  if (lp.acc_smem_bytes > 0 && (lp.backend == MODE_K3 || lp.backend == MODE_K4)) {
    int grid = h->num_sms < p.n_items ? h->num_sms : p.n_items;
    int rc;
    const bool pair = lp.acc2_smem_bytes > 0 && acc_cta2_enabled();
    const int sm = pair ? lp.acc2_smem_bytes : lp.acc_smem_bytes;
#define LC_ACC_LAUNCH(NT, H, R) \
    (pair ? launch_acc2_variant<NT, H, R>(p, p.n_items, sm, s) : launch_acc_variant<NT, H, R>(p, grid, sm, s))
    if (lp.backend == MODE_K4) {
      if (p.half) rc = p.relu ? LC_ACC_LAUNCH(4, 1, 1) : LC_ACC_LAUNCH(4, 1, 0);
      else rc = p.relu ? LC_ACC_LAUNCH(4, 0, 1) : LC_ACC_LAUNCH(4, 0, 0);
    } else if (p.half) rc = p.relu ? LC_ACC_LAUNCH(3, 1, 1) : LC_ACC_LAUNCH(3, 1, 0);
    else rc = p.relu ? LC_ACC_LAUNCH(3, 0, 1) : LC_ACC_LAUNCH(3, 0, 0);
#undef LC_ACC_LAUNCH
    if (rc != LC_OK) return rc;
    LC_LAUNCH_CHECK();
    return LC_OK;
  }
  const int smem = lp.smem_bytes;
  const int per_sm = ctas_per_sm(lp.backend);
  int grid = per_sm * h->num_sms;
  if (grid > p.n_items) grid = p.n_items;
  int rc = LC_OK;
  switch (lp.backend) {
    case MODE_K3:
      rc = launch_mode<MODE_K3>(h, p, grid, smem, s);
      break;
    case MODE_K4:
      rc = launch_mode<MODE_K4>(h, p, grid, smem, s);
      break;
    case MODE_K4_OUT1:
      rc = launch_mode<MODE_K4_OUT1>(h, p, grid, smem, s);
      break;
    case MODE_DOWN:
      rc = launch_mode<MODE_DOWN>(h, p, grid, smem, s);
      break;
    case MODE_UP:
      rc = launch_mode<MODE_UP>(h, p, grid, smem, s);
      break;
    default:
      lc_set_error("unknown tensor-core mode %d", lp.backend);
      return LC_EINVAL;
  }
  if (rc != LC_OK) return rc;
  LC_LAUNCH_CHECK();
  return LC_OK;
}

// One K group = 8 (pseudo-)channels of one shifted view of one plane of the stage.
struct KGroup {
  uint32_t a_addr;        // byte offset inside the stage (filled once the slab size is final)
  int plane, off_slots;   // plane of the stage, slot offset of the shifted view inside the slab
  int tap_h, tap_w;       // kernel indices (kh, kw) this group multiplies with
  int ci0;                // first input channel of the group (or pseudo-channel for WFOLD)
};

}  // namespace

// Build the per-layer SweepParams template: geometry, K-step tables and the weight image.
static int build_sweep(lc_codec* h, TcPlan* P, int i, int mode, const TensorPlan& tin, const TensorPlan& tout,
                       const TensorPlan* tskip, const std::vector<float>& w16, int half) {
  DevLayer& L = h->layers[i];
  LayerPlan& lp = P->layer[i];
  SweepParams& p = lp.sp;
  memset(&p, 0, sizeof(p));
  p.in_S = tin.S; p.in_GL = tin.GL; p.in_NPG = tin.NP * tin.G; p.in_T = L.in_t;
  p.bias = L.d_b; p.cout = L.cout; p.relu = L.relu; p.half = half; p.out_T = L.out_t;
  p.n_classes = 1;
  const int G = tin.G;
  const int cin = L.cin, cout = L.cout, k = L.k;
  // variants: list of K groups (ordered by increasing operand address) and, per N block, the kt tap
  std::vector<std::vector<KGroup>> groups;       // [variant][group]
  std::vector<std::vector<int>> block_kt;        // [variant][block] -> kt (or -1)
  int block_cols = 24;
  p.row_step = 128; p.row_shift = 0;
  if (mode == MODE_K3) {
    // plain input; rows follow the input's padded numbering; taps (dh,dw) -> slab offset dh*PW+dw
    p.row_PW = tin.PW; p.row0 = tin.PW; p.row_off_h = 1; p.row_off_w = 1;
    p.rows_H = L.in_h; p.rows_W = L.in_w;
    p.tiles_per_chunk = (L.in_h * tin.PW + 127) / 128;
    p.slab_start = -(tin.PW + 1);
    p.slab_slots = 128 + (k - 1) * (tin.PW + 1);
    groups.resize(1);
    for (int g = 0; g < G; ++g)
      for (int dh = 0; dh < k; ++dh)
        for (int dw = 0; dw < k; ++dw)
          groups[0].push_back({0, g, dh * tin.PW + dw, dh, dw, g * 8});
    block_kt.resize(1);
    for (int dt = 0; dt < k; ++dt) block_kt[0].push_back(dt);
    p.N = 80;
    // Residual add on the tensor core: three more K groups (the 27 tap groups leave half a K step unused anyway:
    // 30 groups = 15 MMAs instead of 14) read the skip tile with identity weights, which removes the skip loads,
    // 12 conversions and 12 adds per voxel from the issue-bound epilogue.  Needs the skip tensor in the input's
    // row numbering.
    if (L.res_add && tskip && tskip->fmt == FMT_GP16 && tskip->layout == LAYOUT_PLAIN && tskip->PW == tin.PW &&
        tskip->G == G && !getenv("LOSSYCOMP_NO_SKIP_MMA")) {
      p.skip_mma = 1;
      p.skip_pad = (tskip->GL + p.row0) % 8;
      p.skip_slots = (128 + p.skip_pad + 7) & ~7;
      for (int g = 0; g < G; ++g) groups[0].push_back({0, -1 - g, 0, -1, -1, g * 8});
    }
  } else if (mode == MODE_K4_OUT1) {
    // plain input; taps dh only (dt and dw are folded into N = 4x4); tiles advance by 125 rows
    p.row_PW = tin.PW; p.row0 = tin.PW; p.row_off_h = 1; p.row_off_w = 1;
    p.rows_H = L.in_h; p.rows_W = L.in_w;
    p.row_step = 125; p.row_shift = 1;
    p.tiles_per_chunk = (L.in_h * tin.PW + p.row_step - 1) / p.row_step;
    p.slab_start = -tin.PW;
    p.slab_slots = 128 + (k - 1) * tin.PW;
    groups.resize(1);
    for (int g = 0; g < G; ++g)
      for (int dh = 0; dh < k; ++dh)
        groups[0].push_back({0, g, dh * tin.PW, dh, -2, g * 8});
    block_kt.resize(1);
    for (int dt = 0; dt < k; ++dt) block_kt[0].push_back(dt);
    block_cols = 4; p.N = 16;
  } else if (mode == MODE_K4) {
    // w-folded input (pitch W, no halo columns): taps dh only, pseudo-channel pc = dw*cin + ci
    p.row_PW = tin.PW; p.row0 = L.pad_lo * tin.PW; p.row_off_h = L.pad_lo; p.row_off_w = 0;
    p.rows_H = L.in_h; p.rows_W = L.in_w;
    p.tiles_per_chunk = (L.in_h * tin.PW + 127) / 128;
    p.slab_start = -L.pad_lo * tin.PW;
    p.slab_slots = 128 + (k - 1) * tin.PW;
    groups.resize(1);
    for (int g = 0; g < G; ++g)
      for (int dh = 0; dh < k; ++dh)
        groups[0].push_back({0, g, dh * tin.PW, dh, -1, g * 8});
    block_kt.resize(1);
    for (int dt = 0; dt < k; ++dt) block_kt[0].push_back(dt);
    p.N = 96;
  } else if (mode == MODE_DOWN) {
    // parity input: hp = 2*ho + kh -> plane (kh&1, kw&1), slot (ho + kh/2)*PW' + wo + kw/2
    p.row_PW = tin.PW; p.row0 = 0; p.row_off_h = 0; p.row_off_w = 0;
    p.rows_H = L.out_h; p.rows_W = L.out_w;
    p.tiles_per_chunk = (L.out_h * tin.PW + 127) / 128;
    p.slab_start = 0;
    p.slab_slots = 128 + tin.PW + 2;
    groups.resize(2);
    block_kt.resize(2);
    for (int var = 0; var < 2; ++var) {
      for (int par = 0; par < 4; ++par)
        for (int g = 0; g < G; ++g)
          for (int dh2 = 0; dh2 < 2; ++dh2)
            for (int dw2 = 0; dw2 < 2; ++dw2)
              groups[var].push_back({0, par * G + g, dh2 * tin.PW + dw2, 2 * dh2 + (par >> 1), 2 * dw2 + (par & 1),
                                     g * 8});
      // var = t_in & 1: odd -> kt {0, 2}; even -> kt {1, 3}
      block_kt[var] = var ? std::vector<int>{0, 2} : std::vector<int>{1, 3};
    }
    p.N = 48;
  } else if (mode == MODE_UP) {
    // plain input at input resolution; one weight variant per output parity class cls = 2*ph + pw:
    // p=0: (k=1, i=j), (k=3, i=j-1);  p=1: (k=2, i=j), (k=0, i=j+1)
    p.row_PW = tin.PW; p.row0 = tin.PW; p.row_off_h = 1; p.row_off_w = 1;
    p.rows_H = L.in_h; p.rows_W = L.in_w;
    p.tiles_per_chunk = (L.in_h * tin.PW + 127) / 128;
    p.n_classes = 2;                      // output-line parity; both column parities run in the same CTA
    p.slab_start = -(tin.PW + 1);
    p.slab_slots = 128 + 2 * (tin.PW + 1);
    groups.resize(4);
    block_kt.resize(4);
    for (int cls = 0; cls < 4; ++cls) {
      const int ph = cls >> 1, pw = cls & 1;
      // (offset, k) pairs per axis, in increasing offset order
      int offh[2], kh[2], offw[2], kw[2];
      if (ph == 0) { offh[0] = -1; kh[0] = 3; offh[1] = 0; kh[1] = 1; } else { offh[0] = 0; kh[0] = 2; offh[1] = 1; kh[1] = 0; }
      if (pw == 0) { offw[0] = -1; kw[0] = 3; offw[1] = 0; kw[1] = 1; } else { offw[0] = 0; kw[0] = 2; offw[1] = 1; kw[1] = 0; }
      for (int g = 0; g < G; ++g)
        for (int a = 0; a < 2; ++a)
          for (int b = 0; b < 2; ++b)
            groups[cls].push_back({0, g, (tin.PW + 1) + offh[a] * tin.PW + offw[b], kh[a], kw[b], g * 8});
      block_kt[cls] = {0, 1, 2, 3};
    }
    p.N = 96;
  }
  // Align the bulk copies: start every slab on a 128-byte boundary of the tensor (the copy engine is
  // several times slower on 16-byte-aligned sources) and make slabs a multiple of 128 bytes.
  {
    const int base = tin.GL + p.row0 - p.row_shift + p.slab_start;
    const int pad = (p.row_step % 8 == 0 && !getenv("LOSSYCOMP_NO_ALIGN")) ? (((base % 8) + 8) % 8) : 0;
    p.slab_start -= pad;
    p.slab_slots = (p.slab_slots + pad + 7) & ~7;
    for (auto& gv : groups)
      for (auto& kg : gv) {
        if (kg.plane >= 0) kg.a_addr = (uint32_t)((kg.plane * p.slab_slots + kg.off_slots + pad) * 16);
        else kg.a_addr = (uint32_t)((p.in_NPG * p.slab_slots + (-1 - kg.plane) * p.skip_slots + p.skip_pad) * 16);
      }
  }
  {   // every bulk copy must stay inside its own plane (guards included)
    const int first = tin.GL + p.row0 - p.row_shift + p.slab_start;
    const int last = first + (p.tiles_per_chunk - 1) * p.row_step + p.slab_slots;
    if (first < 0 || last > tin.S) {
      lc_set_error("layer %d: tile slabs [%d, %d) leave the input plane of %d vectors", i, first, last, tin.S);
      return LC_EINVAL;
    }
    for (auto& gv : groups)
      for (auto& kg : gv)
        if (kg.plane >= 0 && (kg.off_slots < 0 || (int)(kg.a_addr / 16) - kg.plane * p.slab_slots + 128 > p.slab_slots)) {
          lc_set_error("layer %d: operand view leaves its slab", i);
          return LC_EINVAL;
        }
  }
  if (p.skip_mma) {
    const int first = tskip->GL + p.row0 - p.skip_pad;
    const int last = first + (p.tiles_per_chunk - 1) * 128 + p.skip_slots;
    if (first < 0 || last > tskip->S) {
      lc_set_error("layer %d: skip tiles [%d, %d) leave the skip plane of %d vectors", i, first, last, tskip->S);
      return LC_EINVAL;
    }
  }
  p.nvar = (int)groups.size();
  const int ngroups = (int)groups[0].size();
  p.nks = (ngroups + 1) / 2;
  if (p.in_NPG + (p.skip_mma ? G : 0) > 32) {
    lc_set_error("layer %d: %d bulk copies per stage exceed the producer warp", i, p.in_NPG + (p.skip_mma ? G : 0));
    return LC_EUNSUPPORTED;
  }
  if (p.nks > kMaxKs || p.nvar > kMaxVar) {
    lc_set_error("layer %d: %d K steps / %d variants exceed the kernel tables", i, p.nks, p.nvar);
    return LC_EUNSUPPORTED;
  }
  const size_t var_elems = (size_t)p.nks * 2 * p.N * 8;
  std::vector<unsigned short> img(var_elems * p.nvar, 0);
  for (int var = 0; var < p.nvar; ++var) {
    p.b_base[var] = (uint32_t)(var * var_elems * 2);
    for (int ks = 0; ks < p.nks; ++ks) {
      const int g0 = 2 * ks, g1 = (2 * ks + 1 < ngroups) ? 2 * ks + 1 : 2 * ks;
      if (groups[var][g1].a_addr < groups[var][g0].a_addr) {
        lc_set_error("layer %d: K groups are not address ordered", i);
        return LC_EINVAL;
      }
      p.a_off[var][ks] = groups[var][g0].a_addr;
      p.a_lbo[var][ks] = groups[var][g1].a_addr - groups[var][g0].a_addr;
      auto host_desc = [](uint32_t addr, uint32_t lbo, uint32_t sbo) {
        uint64_t d = 0;
        d |= (uint64_t)((addr >> 4) & 0x3FFF);
        d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
        d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
        d |= (uint64_t)1 << 46;
        return d;
      };
      p.a_desc[var][ks] = host_desc(p.a_off[var][ks], p.a_lbo[var][ks], 128);
      p.b_desc[var][ks] = host_desc((uint32_t)(var * var_elems * 2 + (size_t)ks * 2 * p.N * 16), (uint32_t)p.N * 16, 128);
    }
    for (int gi = 0; gi < ngroups; ++gi) {
      const KGroup& kg = groups[var][gi];
      for (int n = 0; n < p.N; ++n) {
        const int blk = n / block_cols;
        int co = n % block_cols, kw_fold = -1;
        if (mode == MODE_K4_OUT1) { kw_fold = co; co = 0; }      // column = dt*4 + dw, single output channel
        if (blk >= (int)block_kt[var].size() || co >= cout) continue;
        const int kt = block_kt[var][blk];
        if (kg.plane < 0) {   // skip group: identity into the centre time block (output plane = input plane)
          if (kt == L.pad_lo && co >= kg.ci0 && co < kg.ci0 + 8)
            img[var * var_elems + ((size_t)gi * p.N + n) * 8 + (co - kg.ci0)] = to_h16_host(1.0f, half);
          continue;
        }
        for (int k8 = 0; k8 < 8; ++k8) {
          int kh = kg.tap_h, kw = kg.tap_w, ci = kg.ci0 + k8;
          if (mode == MODE_K4) {   // pseudo-channel -> (dw, ci)
            kw = ci / cin;
            ci = ci - kw * cin;
            if (kw >= k) continue;
          }
          if (kw_fold >= 0) kw = kw_fold;
          if (ci >= cin) continue;
          const float v = w16[((size_t)((kt * k + kh) * k + kw) * cin + ci) * cout + co];
          img[var * var_elems + ((size_t)gi * p.N + n) * 8 + k8] = to_h16_host(v, half);
        }
      }
    }
  }
  p.wimg_bytes = (int)(img.size() * 2);
  LC_CUDA(cudaMalloc(&L.d_wtc, img.size() * 2));
  LC_CUDA(cudaMemcpy(L.d_wtc, img.data(), img.size() * 2, cudaMemcpyHostToDevice));
  L.wtc_bytes = img.size() * 2;
  p.wimg = reinterpret_cast<const uint4*>(L.d_wtc);
  // a stage = the input slabs of one plane (+ the skip tile when the residual add runs on the tensor core)
  const int stage_bytes = p.in_NPG * p.slab_slots * 16 + (p.skip_mma ? G * p.skip_slots * 16 : 0);
  lp.acc_smem_bytes = 0;
  lp.d_wimg2 = nullptr;
  lp.acc2_smem_bytes = 0;
  lp.d_wimg2c[0] = lp.d_wimg2c[1] = nullptr;
  // Layers with a residual input run on sweep_acc_kernel too (LOSSYCOMP_ACC_SKIP=0: back to sweep_tc_kernel's register
  // fold).  Measured on the benchmark field: register fold 1.12 ms, accumulate form 1.17 ms with a single-thread producer
  // and one CTA per tile set, 0.98 ms with the warp-wide producer and CTA pairs.
  const bool acc_ok = !L.res_add || (p.skip_mma && !(getenv("LOSSYCOMP_ACC_SKIP") && !atoi(getenv("LOSSYCOMP_ACC_SKIP"))));
  const bool acc_mode = mode == MODE_K3 || (mode == MODE_K4 && L.pad_lo == 1 && !getenv("LOSSYCOMP_NO_ACC_K4"));
  if (acc_mode && acc_ok && cout <= 24 && p.in_NPG + (p.skip_mma ? G : 0) <= 8 && !getenv("LOSSYCOMP_NO_ACC")) {
    // weight image of sweep_acc_kernel: per 8-channel K group, the NT time blocks twice in decreasing order
    // ([dt = NT-1 .. 0, NT-1 .. 1], 24 rows each) so that a start offset of i0 blocks rotates them onto the TMEM
    // slots; padded to the rows the widest start (NT-1 blocks in) reads with N = bank columns
    const int NT = k;
    const int bank_cols = (NT * 24 + 15) & ~15;
    const int rows = ((NT - 1) * 24 + bank_cols + 7) & ~7;
    std::vector<unsigned short> img2((size_t)p.nks * 2 * rows * 8, 0);
    for (int gi = 0; gi < ngroups; ++gi) {
      const KGroup& kg = groups[0][gi];
      for (int i = 0; i < 2 * NT - 1; ++i) {
        const int kt = (NT - 1 - i % NT + NT) % NT;
        for (int co = 0; co < cout; ++co) {
          unsigned short* dst = &img2[((size_t)gi * rows + (size_t)i * 24 + co) * 8];
          if (kg.plane < 0) {   // skip group: identity into the centre time block
            if (kt == L.pad_lo && co >= kg.ci0 && co < kg.ci0 + 8) dst[co - kg.ci0] = to_h16_host(1.0f, half);
            continue;
          }
          for (int k8 = 0; k8 < 8; ++k8) {
            int ci = kg.ci0 + k8, kw = kg.tap_w;
            if (mode == MODE_K4) {   // w-folded input: pseudo-channel -> (dw, ci)
              kw = ci / cin;
              ci -= kw * cin;
              if (kw >= k) continue;
            }
            if (ci >= cin) continue;
            dst[k8] = to_h16_host(w16[((size_t)((kt * k + kg.tap_h) * k + kw) * cin + ci) * cout + co], half);
          }
        }
      }
    }
    auto host_desc2 = [](uint32_t addr, uint32_t lbo, uint32_t sbo) {
      uint64_t d = 0;
      d |= (uint64_t)((addr >> 4) & 0x3FFF);
      d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
      d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
      d |= (uint64_t)1 << 46;
      return d;
    };
    for (int ks = 0; ks < p.nks; ++ks)
      p.b2_desc[ks] = host_desc2((uint32_t)((size_t)ks * 2 * rows * 16), (uint32_t)rows * 16, 128);
    p.wimg2_bytes = (int)(img2.size() * 2);
    p.b2_rows = rows;
    LC_CUDA(cudaMalloc(&lp.d_wimg2, img2.size() * 2));
    LC_CUDA(cudaMemcpy(lp.d_wimg2, img2.data(), img2.size() * 2, cudaMemcpyHostToDevice));
    p.wimg2 = reinterpret_cast<const uint4*>(lp.d_wimg2);
    // Output staging for the bulk-copy epilogue (LOSSYCOMP_TMA_STORE=1; needs the output planes in the rows' own slot
    // numbering).  Off by default: measured equal to the per-thread 16-byte stores (0.868 vs 0.849 ms on the 3x3x3
    // layers in pair form) while its 48 KB of staging cost input stages.
    p.tma_store = (mode == MODE_K3 && tout.fmt == FMT_GP16 && tout.layout == LAYOUT_PLAIN && tout.PW == tin.PW &&
                   tout.G <= 3 && getenv("LOSSYCOMP_TMA_STORE") && atoi(getenv("LOSSYCOMP_TMA_STORE"))) ? 1 : 0;
    const int ostage = p.tma_store ? kAccBanks * 2 * 3 * 128 * 16 : 0;
    const int kAccSmem = 225 * 1024;
    const int fixed2 = ((p.wimg2_bytes + 127) & ~127) + 256 + ostage;
    int st2 = (kAccSmem - fixed2) / stage_bytes;
    if (st2 > kAccMaxStages) st2 = kAccMaxStages;
    if (getenv("LOSSYCOMP_ACC_STAGES")) st2 = atoi(getenv("LOSSYCOMP_ACC_STAGES"));
    st2 -= st2 % kAccBanks;                  // every bank owns stages / kAccBanks of them
    if (st2 >= 4) {
      p.stages2 = st2;
      lp.acc_smem_bytes = fixed2 + st2 * stage_bytes;
    }
    // CTA-pair form: rank r of the pair multiplies columns [r*N/2, (r+1)*N/2) of the rotated window, i.e. rows
    // [i0*24 + r*N/2, +N/2) of the image above; its own image holds rows [r*N/2, r*N/2 + rows_c) at offset 0, so
    // that one descriptor addresses the right half in both CTAs
    {
      const int half_cols = bank_cols / 2;
      const int rows_c = ((NT - 1) * 24 + half_cols + 7) & ~7;
      const bool swap = getenv("LOSSYCOMP_ACC2_SWAP") != nullptr;     // debugging: which CTA owns which half
      std::vector<unsigned short> imgc[2];
      for (int r = 0; r < 2; ++r) {
        imgc[r].assign((size_t)p.nks * 2 * rows_c * 8, 0);
        const int shift = (swap ? 1 - r : r) * half_cols;
        for (int gi = 0; gi < ngroups; ++gi)
          for (int j = 0; j < rows_c; ++j) {
            if (j + shift >= rows) continue;
            memcpy(&imgc[r][((size_t)gi * rows_c + j) * 8], &img2[((size_t)gi * rows + j + shift) * 8], 16);
          }
      }
      for (int ks = 0; ks < p.nks; ++ks)
        p.b2c_desc[ks] = host_desc2((uint32_t)((size_t)ks * 2 * rows_c * 16), (uint32_t)rows_c * 16, 128);
      p.wimg2c_bytes = (int)(imgc[0].size() * 2);
      p.b2c_rows = rows_c;
      for (int r = 0; r < 2; ++r) {
        LC_CUDA(cudaMalloc(&lp.d_wimg2c[r], imgc[r].size() * 2));
        LC_CUDA(cudaMemcpy(lp.d_wimg2c[r], imgc[r].data(), imgc[r].size() * 2, cudaMemcpyHostToDevice));
      }
      p.wimg2c0 = reinterpret_cast<const uint4*>(lp.d_wimg2c[0]);
      p.wimg2c1 = reinterpret_cast<const uint4*>(lp.d_wimg2c[1]);
      const int fixedc = ((p.wimg2c_bytes + 127) & ~127) + 256 + ostage;
      int stc = (kAccSmem - fixedc) / stage_bytes;
      if (stc > kAccMaxStages) stc = kAccMaxStages;
      if (getenv("LOSSYCOMP_ACC_STAGES")) stc = atoi(getenv("LOSSYCOMP_ACC_STAGES"));
      stc -= stc % kAccBanks;
      p.stages2c = stc;
      lp.acc2_smem_bytes = (stc >= 4 && lp.acc_smem_bytes > 0 && bank_cols % 16 == 0) ? fixedc + stc * stage_bytes : 0;
    }
  }
  // dynamic shared memory per CTA; the single-channel mode also holds a 17.6 KB static exchange tile
  const int budget = (ctas_per_sm(mode) == 2 ? (mode == MODE_K4_OUT1 ? 90 : 110) : 200) * 1024;
  const int fixed = ((p.wimg_bytes + 127) & ~127) + 256;
  int stages = (budget - fixed) / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  // an even ring: the two issuer warps take alternate steps, so with an odd ring a stage's previous use belongs to the
  // other issuer, whose wait this issuer never saw -- the same parity aliasing sweep_acc_kernel's private rings avoid
  if (stages > 2) stages -= stages % 2;
  if (stages < 2) {
    lc_set_error("layer %d: tensor-core kernel does not fit shared memory (%d + n*%d)", i, fixed, stage_bytes);
    return LC_EUNSUPPORTED;
  }
  p.stages = stages;
  {
    const bool two_per_sm = ctas_per_sm(mode) == 2;
    p.tmem_cols = two_per_sm ? 256 : 512;
    p.acc_stride = (p.N + 7) & ~7;
    p.nbuf = p.tmem_cols / p.acc_stride;
    if (p.nbuf > kMaxBufs) p.nbuf = kMaxBufs;
    if (getenv("LOSSYCOMP_TC_NBUF")) p.nbuf = atoi(getenv("LOSSYCOMP_TC_NBUF"));
  }
  lp.smem_bytes = fixed + stages * stage_bytes;
  lp.backend = mode;
  p.debug = getenv("LOSSYCOMP_TC_DEBUG") ? atoi(getenv("LOSSYCOMP_TC_DEBUG")) : 0;
  (void)tout;
  return LC_OK;
}

int tc_prepare(lc_codec* h) {
  h->tc_plan = nullptr;
  if (h->precision == LC_PREC_FP32) return LC_OK;
  const int half = (h->precision == LC_PREC_FP16) ? 1 : 0;
  TcPlan* P = new TcPlan();
  memset(P, 0, sizeof(*P));
  h->tc_plan = P;
  const int nl = h->n_enc + h->n_dec;
  P->n_layers = nl;
  std::vector<int> mode(nl);
  for (int i = 0; i < nl; ++i) mode[i] = tc_mode_of(h, i);
  long long off = 0;
  auto place = [&](TensorPlan& t) {
    t.offset_vecs = off;
    off += (t.size_vecs + 7) & ~7LL;
  };
  if (mode[0] == MODE_K4) {
    P->in_enc = make_wfold(CT, CH, CW, h->in_channels, h->layers[0].k, h->layers[0].pad_lo);
    P->wfold_k = h->layers[0].k;
  } else if (mode[0] == MODE_DOWN) {
    P->in_enc = make_gp(CT, CH, CW, h->in_channels, 1, LAYOUT_PARITY);
  } else {
    P->in_enc = make_f32(CT, CH, CW, h->in_channels);
  }
  place(P->in_enc);
  for (int i = 0; i < nl; ++i) {
    const DevLayer& L = h->layers[i];
    const bool last_enc = (i == h->n_enc - 1), last_dec = (i == nl - 1);
    if (last_enc || last_dec) {
      P->out[i] = make_f32(L.out_t, L.out_h, L.out_w, L.cout);   // latent / decoder output stay fp32
    } else {
      const DevLayer& Nx = h->layers[i + 1];
      const int pad_hi = (Nx.kind == 0 && Nx.stride == 1) ? (Nx.k - 1 - Nx.pad_lo) : 1;
      const int layout = (mode[i + 1] == MODE_DOWN) ? LAYOUT_PARITY : LAYOUT_PLAIN;
      P->out[i] = make_gp(L.out_t, L.out_h, L.out_w, L.cout, pad_hi < 1 ? 1 : pad_hi, layout);
    }
    if (!last_enc && !(last_dec && mode[i] == MODE_K4_OUT1)) place(P->out[i]);
  }
  P->block_vecs = (off + 63) & ~63LL;

  int skip_src = -1;      // layer whose output the next residual add reads
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
    if (mode[i]) {
      const TensorPlan& tin = (i == 0) ? P->in_enc : P->out[i - 1];
      int rc = build_sweep(h, P, i, mode[i], tin, P->out[i], (L.res_add && skip_src >= 0) ? &P->out[skip_src] : nullptr,
                           w16, half);
      if (rc != LC_OK) return rc;
    }
    if (L.res_save) skip_src = i;
  }
  return LC_OK;
}

void tc_release(lc_codec* h) {
  TcPlan* P = plan_of(h);
  if (!P) return;
  for (int i = 0; i < P->n_layers; ++i) {
    if (P->layer[i].d_w16) cudaFree(P->layer[i].d_w16);
    if (P->layer[i].d_wimg2) cudaFree(P->layer[i].d_wimg2);
    for (int r = 0; r < 2; ++r)
      if (P->layer[i].d_wimg2c[r]) cudaFree(P->layer[i].d_wimg2c[r]);
  }
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
                  float* latent_out, float* recon_out, const ChunkGeom* geom, int chunk_begin, cudaStream_t s) {
  TcPlan* P = plan_of(h);
  const int half = (h->precision == LC_PREC_FP16) ? 1 : 0;
  ActView cur = first_in;
  ActView skip_v;
  memset(&skip_v, 0, sizeof(skip_v));
  for (int i = first; i < first + count; ++i) {
    DevLayer& L = h->layers[i];
    LayerPlan& lp = P->layer[i];
    ActView out;
    memset(&out, 0, sizeof(out));
    if (i == h->n_enc - 1) {   // latent: contiguous fp32 in the caller's buffer
      out = f32_view(L, true, latent_out);
    } else if (lp.backend != MODE_K4_OUT1) {
      out = view_of(P->out[i], block_base, P->block_vecs, half);
    }
    prof_begin(h, i, s);
    int rc;
    if (lp.backend) {
      SweepParams p = lp.sp;
      p.in = reinterpret_cast<const uint4*>(cur.base);
      p.in_cs = cur.chunk_stride;
      p.out = out;
      if (L.res_add) p.skip = skip_v;
      p.n_items = n_chunks * p.tiles_per_chunk * p.n_classes;
      if (lp.backend == MODE_K4_OUT1) {
        p.out_f32 = recon_out;
        p.geom = *geom;
        p.chunk_begin = chunk_begin;
      }
      rc = launch_sweep(h, lp, p, s);
    } else {
      rc = conv_direct_views(h, L, lp.d_w16, cur, out, L.res_add ? &skip_v : nullptr, n_chunks, s);
    }
    prof_end(h, s);
    if (rc != LC_OK) return rc;
    if (L.res_save) skip_v = out;
    cur = out;
  }
  return LC_OK;
}

int tc_encode(lc_codec* h, const float* d_x, const ChunkGeom& g, int C, float mean, float stdv, int chunk_begin,
              int chunk_count, float* d_latent, void* d_ws, cudaStream_t s) {
  TcPlan* P = plan_of(h);
  const int half = (h->precision == LC_PREC_FP16) ? 1 : 0;
  ActView in = view_of(P->in_enc, d_ws, P->block_vecs, half);
  if (P->in_enc.fmt == FMT_F32) {
    gather_blocks_kernel<<<h->num_sms * 8, 256, 0, s>>>(d_x, g, C, mean, stdv, chunk_begin, chunk_count,
                                                        reinterpret_cast<float*>(in.base), in.chunk_stride);
  } else if (P->in_enc.layout != LAYOUT_WFOLD) {
    gather_gp_kernel<<<h->num_sms * 16, 256, 0, s>>>(d_x, g, C, mean, stdv, chunk_begin, chunk_count, in);
  } else {
    if (C == 2 && P->wfold_k == 4 && in.G == 1 && (reinterpret_cast<uintptr_t>(d_x) & 7) == 0)
      gather_wfold_c2k4_kernel<<<h->num_sms * 16, 256, 0, s>>>(d_x, g, mean, stdv, chunk_begin, chunk_count, in,
                                                                h->layers[0].pad_lo);
    else
      gather_wfold_kernel<<<h->num_sms * 8, 256, 0, s>>>(d_x, g, C, mean, stdv, chunk_begin, chunk_count, in,
                                                         P->wfold_k, h->layers[0].pad_lo);
  }
  LC_LAUNCH_CHECK();
  return tc_run(h, 0, h->n_enc, in, d_ws, chunk_count, d_latent, nullptr, nullptr, 0, s);
}

int tc_decode(lc_codec* h, const float* d_latent, const ChunkGeom& g, int chunk_begin, int chunk_count,
              float* d_recon, void* d_ws, cudaStream_t s) {
  TcPlan* P = plan_of(h);
  const int half = (h->precision == LC_PREC_FP16) ? 1 : 0;
  const int last = h->n_enc + h->n_dec - 1;
  ActView in = f32_view(h->layers[h->n_enc], false, d_latent);
  int rc = tc_run(h, h->n_enc, h->n_dec, in, d_ws, chunk_count, nullptr, d_recon, &g, chunk_begin, s);
  if (rc != LC_OK) return rc;
  if (P->layer[last].backend != MODE_K4_OUT1) {
    ActView out = view_of(P->out[last], d_ws, P->block_vecs, half);
    merge_blocks_kernel<<<h->num_sms * 8, 256, 0, s>>>(reinterpret_cast<const float*>(out.base), out.chunk_stride,
                                                       g, chunk_begin, chunk_count, d_recon);
    LC_LAUNCH_CHECK();
  }
  return LC_OK;
}
