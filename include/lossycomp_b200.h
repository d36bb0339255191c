/* lossycomp_b200.h -- C-ABI of the B200-native lossycomp codec hot path.
 *
 * The reference (SilkeDH/lossy-ml) is pure Python with no FFI of its own; its boundary for
 * this path is the two callables lossycomp/compress.py:21 `compress(array, err_threshold)` and
 * lossycomp/decompress.py:21 `decompress(blob)`.  Each entry point below replaces the span of
 * reference Python named in its comment; lossy-ml_b200/lossycomp/{compress,decompress}.py call
 * them through ctypes (see INTEGRATION.md for the stub a reference maintainer would add).
 *
 * Conventions
 *   - every function returns 0 on success, a negative LC_E* code otherwise; lc_last_error()
 *     returns a thread-local message for the last failure;
 *   - pointers named d_* are DEVICE pointers owned by the caller, h_* are host pointers;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing
 *     synchronises unless stated ("sync");
 *   - no hidden allocation after lc_create(): scratch comes from caller workspaces whose sizes
 *     the *_workspace_bytes queries return;
 *   - all kernels are deterministic (fixed reduction / accumulation order, no float atomics):
 *     the decoder output is bit-identical for any batch split, stream, GPU count or run, which
 *     is what makes the error bound hold across compress/decompress.
 */
#ifndef LOSSYCOMP_B200_H
#define LOSSYCOMP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LC_OK 0
#define LC_EINVAL (-1)
#define LC_ECUDA (-2)
#define LC_ENOMEM (-3)
#define LC_EUNSUPPORTED (-4)

#define LC_CHUNK_T 16
#define LC_CHUNK_H 48
#define LC_CHUNK_W 48

/* arithmetic used by the convolution kernels */
#define LC_PREC_FP32 0 /* CUDA-core fp32 FMA direct convolution (reference-grade numerics) */
#define LC_PREC_BF16 1 /* tcgen05 implicit GEMM, bf16 operands, fp32 TMEM accumulators */
#define LC_PREC_FP16 2 /* tcgen05 implicit GEMM, fp16 operands, fp32 TMEM accumulators */

typedef struct lc_codec* lc_handle_t;

/* one convolution layer of the autoencoder, Keras layouts (models.py:47-97) */
typedef struct {
  int32_t kind;     /* 0 = Conv3D, 1 = Conv3DTranspose */
  int32_t k;        /* cubic kernel size */
  int32_t stride;   /* 1 or 2 */
  int32_t cin, cout;
  int32_t relu;     /* fused ReLU on the output */
  int32_t res_save; /* output is the skip input of the following ResBlock */
  int32_t res_add;  /* add the saved skip tensor before the ReLU */
  const float* h_kernel; /* Conv3D: (k,k,k,cin,cout); Conv3DTranspose: (k,k,k,cout,cin) */
  const float* h_bias;   /* (cout) */
} lc_layer_t;

const char* lc_last_error(void);
int lc_version(void);

/* Replaces compress.py:41-57 / decompress.py:33-48 (parameter pickle, Keras graph build,
 * load_weights): uploads and re-lays-out the weights once.  n_enc encoder layers followed by
 * n_dec decoder layers, in execution order. */
int lc_create(const lc_layer_t* layers, int n_enc, int n_dec, int in_channels, int precision,
              int device, lc_handle_t* out);
/* Device rule for every other entry point: kernels are launched on the calling thread's CURRENT device, which must
 * be the device the handle, the pointers and the stream belong to (lc_create itself restores the caller's current
 * device before it returns). */
int lc_destroy(lc_handle_t h);
int lc_latent_floats(lc_handle_t h);
/* Which kernel runs layer i: 0 = CUDA-core direct conv, 1 = tcgen05 k3, 2 = k4 first layer, 3 = k4 last
 * layer (+ merge), 4 = stride-2 down, 5 = transposed up. */
int lc_layer_backend(lc_handle_t h, int layer); /* floats per chunk in the latent, e.g. 1*3*3*23 */

/* Per-layer kernel timing for bench.py's roofline: when enabled, every convolution launch is
 * bracketed by a CUDA event pair on the launching stream; lc_profile_read (sync) returns and
 * resets the accumulated milliseconds and launch counts per layer (n >= n_enc + n_dec). */
int lc_profile_enable(lc_handle_t h, int on);
int lc_profile_read(lc_handle_t h, double* h_ms, long long* h_launches, int n);

/* chunk grid of dataLoader.py:449: ceil(T/16), ceil(H/48), ceil(W/48) */
int lc_chunk_grid(int T, int H, int W, int grid[3]);

/* Host-only view of the same index map: out = {t0, h0, w0 (window origin, dataLoader.py:457-481),
 * kt, kh, kw (first local index merge_data keeps, dataLoader.py:503-532)} of chunk `idx`. */
int lc_chunk_window(int T, int H, int W, int idx, int out[6]);

/* dataLoader.py:437-481 chunk_data: (T,H,W,C) -> (N,16,48,48,C), N = product of lc_chunk_grid,
 * chunk index lon-fastest, tail windows re-anchored at axis_len-size (overlap, no padding). */
int lc_chunk_gather(const float* d_x, int T, int H, int W, int C, float* d_chunks, void* stream);
/* dataLoader.py:484-532 merge_data: (N,16,48,48,C) -> (T,H,W,C); a tail chunk contributes only
 * the voxels past its overlap with its predecessor. */
int lc_chunk_merge(const float* d_chunks, int T, int H, int W, int C, float* d_out, void* stream);

/* compress.py:64-65 -- statistics of channel 0 of d_x (T,H,W,C) float32.  Writes
 * d_out[0..3] = { sum x, sum x^2, max|x|, n } as doubles (fixed-order two-level reduction).
 * d_ws needs lc_stats_workspace_bytes(). */
size_t lc_stats_workspace_bytes(void);
int lc_stats(const float* d_x, int64_t n_vox, int C, double* d_out, void* d_ws, void* stream);
/* Shard-invariant form of the same statistics: the voxels are cut into consecutive blocks of block_vox (the codec
 * uses one 16-step time chunk = 16*H*W; the last block may be shorter) and each block is reduced on its own in a
 * fixed order; d_out[4*b .. 4*b+3] = { sum, sum^2, max|x|, n } of block b.  The caller adds the blocks in index
 * order, so the ranks of a time-sharded field (which exchange their blocks' partials) obtain bit for bit the mean
 * and std of the single-GPU call.  Same workspace as lc_stats. */
int lc_stats_blocks(const float* d_x, int64_t n_vox, int64_t block_vox, int C, double* d_out, void* d_ws,
                    void* stream);

/* compress.py:66-86 + dataLoader.py:437-481 + models.py:52-67 -- standardise channel 0,
 * cut chunks [chunk_begin, chunk_begin+chunk_count) of the (T,H,W,C) field (tail windows are
 * re-anchored at axis_len-size) and run the encoder.  d_latent receives
 * (chunk_count, latent_floats) float32 in Keras order (t,h,w,f). */
size_t lc_ae_workspace_bytes(lc_handle_t h, int max_chunks);
/* Workspace contract (16-bit precisions): the tensor-core kernels read the halo / guard slots of their activation
 * planes as zeros and never write them, so a workspace has to be zeroed ONCE before its first use and must then
 * belong to the codec alone (not shared with other work, not used as scratch between calls; lc_encode and lc_decode
 * calls on one workspace must be stream-ordered).  lc_ae_workspace_init enqueues the zeroing on `stream` and
 * registers the buffer with the handle; lc_encode / lc_decode return LC_EINVAL for a buffer that was not registered
 * (up to 16 buffers per handle; lc_ae_workspace_release forgets one before it is freed).  The fp32 precision has
 * no such requirement. */
int lc_ae_workspace_init(lc_handle_t h, void* d_ws, size_t ws_bytes, void* stream);
int lc_ae_workspace_release(lc_handle_t h, void* d_ws);
int lc_encode(lc_handle_t h, const float* d_x, int T, int H, int W, int C, float mean, float std,
              int chunk_begin, int chunk_count, float* d_latent, void* d_ws, size_t ws_bytes,
              void* stream);

/* compress.py:99-111 / decompress.py:66-72 + dataLoader.py:484-532 + models.py:70-85 -- run
 * the decoder on chunks [chunk_begin, +chunk_count) and merge: each chunk writes only the
 * voxels merge_data keeps from it into d_recon (T,H,W) float32 (standardised units).
 * d_latent points at the latent of chunk `chunk_begin`. */
int lc_decode(lc_handle_t h, const float* d_latent, int T, int H, int W, int chunk_begin,
              int chunk_count, float* d_recon, void* d_ws, size_t ws_bytes, void* stream);

/* compress.py:114-135 -- de-standardise, residual, 3-state mask, outlier compaction (C order)
 * and quantisation, one pass over HBM:
 *   r' = fl32(fl32(r*std)+mean); e = x0 - r'; m = |e|<=thr ? 0 : (e>0 ? 1 : 2)
 *   q  = int32(rint(|e| / thr2)) for outliers, appended in C order.
 * thr = fl32(abs_err), thr2 = fl32(2*abs_err) are passed by the host exactly as NumPy forms
 * them.  d_count receives the number of outliers (uint64).  d_ws: lc_scan_workspace_bytes(). */
size_t lc_scan_workspace_bytes(int64_t n);
int lc_quantize(const float* d_x, int C, const float* d_recon, int64_t n_vox, float mean, float std,
                float thr, float thr2, int8_t* d_mask, int32_t* d_q, unsigned long long* d_count,
                void* d_ws, void* stream);
/* lc_quantize plus the strict mode's bound check in the same pass: every element is also decoded exactly
 * as lc_reconstruct will decode it (decompress.py:86-104, thr2_dec = the double stored in blob slot 8) and
 * compared with x in fp64.  d_out3[0] = outliers, d_out3[1] = elements with |x - xhat| > abs_err (NaN counts),
 * d_out3[2] = bits of the largest |x - xhat| (a non-negative double). */
int lc_quantize_checked(const float* d_x, int C, const float* d_recon, int64_t n_vox, float mean, float std,
                        float thr, float thr2, double thr2_dec, double abs_err, int8_t* d_mask, int32_t* d_q,
                        unsigned long long* d_out3, void* d_ws, void* stream);

/* Latent quantisation (the reference stores the bottleneck tensor losslessly through fpzip, compress.py:142-144;
 * here it is rounded to IEEE half precision, which the decoder then consumes on both sides, so the error bound
 * is unaffected -- the residual stage absorbs the rounding).  lc_latent_quantize overwrites d_latent with the
 * dequantised values and writes the low / high byte planes of the half-precision codes; lc_latent_dequantize
 * rebuilds the fp32 latent from the planes. */
int lc_latent_quantize(float* d_latent, int64_t n, uint8_t* d_lo, uint8_t* d_hi, void* stream);
int lc_latent_dequantize(const uint8_t* d_lo, const uint8_t* d_hi, int64_t n, float* d_latent, void* stream);

/* compress.py:147-151,159-163 -- first difference [v0, v1-v0, ...] (int8 wraps like NumPy). */
int lc_diff_i32(const int32_t* d_in, int64_t n, int32_t* d_out, void* stream);
int lc_diff_i8(const int8_t* d_in, int64_t n, int8_t* d_out, void* stream);
/* decompress.py:78,83 -- inclusive prefix sum undoing the first difference. */
int lc_cumsum_i32(const int32_t* d_in, int64_t n, int32_t* d_out, void* d_ws, void* stream);
int lc_cumsum_i8(const int8_t* d_in, int64_t n, int8_t* d_out, void* d_ws, void* stream);

/* decompress.py:86-104 -- scatter de-quantised residuals by mask and add:
 *   val = m>0 ? fl32(double(q_j)*thr2) : 0, negated where m==2 (j = rank of the outlier);
 *   out = fl32(fl32(fl32(r*std)+mean) + val).   thr2 is the double stored in blob slot 8.
 * d_q holds n_q codes; a mask with more outliers than that (a damaged blob) reads zeros instead of memory past
 * the list, and the first 8 bytes of d_ws receive the outlier count of the mask (uint64) so that the caller can
 * compare it with n_q after its next synchronisation. */
int lc_reconstruct(const float* d_recon, const int8_t* d_mask, const int32_t* d_q, int64_t n_q, int64_t n_vox,
                   float mean, float std, double thr2, float* d_out, void* d_ws, void* stream);

/* exhaustive |x - xhat| <= abs_err check in fp64.  d_x has C interleaved channels (channel 0
 * is compared), d_xhat is (n_vox).  d_out[0] = violations (as double), d_out[1] = max |x-xhat|. */
int lc_verify_bound(const float* d_x, int C, const float* d_xhat, int64_t n_vox, double abs_err,
                    double* d_out, void* d_ws, void* stream);

/* ---- entropy stage (north_star stage 4; host table builder mirrors huffman.py:20-45) ----
 * Symbols are bytes (mask stream: packed trits; code stream: clipped first differences with
 * escapes).  Histogram is warp-aggregated in shared memory; d_hist is 256 x uint32. */
int lc_hist_u8(const uint8_t* d_sym, int64_t n, uint32_t* d_hist, void* stream);
/* Histograms of both candidate symbolisations of the int32 code stream in one pass:
 * d_hist[0..255] of (q + bias0), d_hist[256..511] of (q[i]-q[i-1] + bias1); values outside
 * [0, esc) are counted under `esc`. */
int lc_hist_codes(const int32_t* d_q, int64_t n, int esc, int bias0, int bias1, uint32_t* d_hist,
                  void* stream);
/* Map the int32 code stream (optionally first-differenced on the fly) to byte symbols:
 * sym = v + bias if 0 <= v+bias < esc else esc; escaped values go to d_esc in order. */
int lc_symbolize_i32(const int32_t* d_q, int64_t n, int diff, int bias, int esc, uint8_t* d_sym,
                     int32_t* d_esc, unsigned long long* d_nesc, void* d_ws, void* stream);
/* d_esc == NULL: the stream holds no escape symbol (the container header says so): plain element-wise inverse. */
int lc_unsymbolize_i32(const uint8_t* d_sym, int64_t n, int diff, int bias, int esc,
                       const int32_t* d_esc, int32_t* d_q, void* d_ws, void* stream);
/* lc_symbolize_i32 for a stream whose histogram shows that no value leaves [0, esc) after the bias: no escape
 * list, no scan, one pass at memory speed. */
int lc_symbolize_plain_i32(const int32_t* d_q, int64_t n, int diff, int bias, uint8_t* d_sym, void* stream);
/* Pack five 3-state mask values per byte (m0 + 3 m1 + 9 m2 + 27 m3 + 81 m4), tail zero padded. */
int lc_pack_trits(const int8_t* d_mask, int64_t n, uint8_t* d_sym, void* stream);
/* The same packing plus the 256-bin histogram of the packed bytes in one pass (d_hist: 256 x uint32, zeroed by the
 * call): what the entropy stage needs to build the mask's Huffman table and to decide whether run tokens pay. */
int lc_pack_trits_hist(const int8_t* d_mask, int64_t n, uint8_t* d_sym, uint32_t* d_hist, void* stream);
int lc_unpack_trits(const uint8_t* d_sym, int64_t n, int8_t* d_mask, void* stream);

/* Run tokens for sparse streams (large abs_error): a Huffman code cannot spend less than one bit per symbol,
 * so runs of the dominant byte `run_sym` are folded into the alphabet: a run of 2^(k+1) of them (k = 0..4) becomes
 * the token run_first + k, every other byte (a single run_sym included) stays a literal.  Runs do not cross
 * 32-byte segments and are decomposed largest power of two first.  The stream must not contain the bytes
 * run_first .. run_first+4 (the caller checks its histogram).  Mask bytes: run_sym 0, run_first 243; code symbols:
 * run_sym 1, run_first 250.  lc_rle_tokens writes the tokens (at most n), their count, their 256-bin histogram
 * and, when d_hist_in is not NULL, the histogram of the input bytes (same pass); lc_rle_expand is the inverse (d_sym: n bytes).  d_ws: lc_rle_workspace_bytes(n). */
size_t lc_rle_workspace_bytes(int64_t n);
int lc_rle_tokens(const uint8_t* d_sym, int64_t n, int run_sym, int run_first, uint8_t* d_tok,
                  unsigned long long* d_n_tok, uint32_t* d_hist, uint32_t* d_hist_in, void* d_ws, void* stream);
int lc_rle_expand(const uint8_t* d_tok, int64_t n_tok, int run_sym, int run_first, uint8_t* d_sym, int64_t n,
                  void* d_ws, void* stream);

/* Huffman bit packing from a host-built table: code[s] (right-aligned, MSB of the code is
 * emitted first) and len[s] (0..32).  The stream is cut into segments of `seg` symbols (a power of two, 64..4096); each
 * segment starts at the bit offset d_seg_bits[i] (uint64, exclusive scan written by the call),
 * bits are big-endian within bytes exactly like the '0101..' strings of huffman.py.
 * d_total_bits receives the stream length in bits.  d_out must be zero-initialised and hold
 * lc_huff_max_bytes(n) bytes. */
size_t lc_huff_max_bytes(int64_t n);
/* Host only: the prefix code of a 256-bin histogram with the reference's tree construction and tie-breaking
 * (lossycomp/huffman.py:20-45,71: leaves by (count, symbol), stable re-sort before each merge, first popped
 * = left = '0').  When that tree is deeper than max_bits the counts are flattened (>> k, floor 1) just enough
 * and *reference_tree = 0.  code/len/used: 256 entries each. */
int lc_huff_build_table(const int64_t* hist, int max_bits, uint32_t* code, uint8_t* len, uint8_t* used,
                        int* reference_tree);
size_t lc_huff_workspace_bytes(int64_t n, int seg);
int lc_huff_encode(const uint8_t* d_sym, int64_t n, const uint32_t* d_code, const uint8_t* d_len,
                   int seg, uint8_t* d_out, unsigned long long* d_seg_bits,
                   unsigned long long* d_total_bits, void* d_ws, void* stream);
/* Host only: the decode-side arrays lc_huff_decode consumes, from a table as lc_huff_build_table (or a blob header)
 * gives it.  lut: 2^lut_bits entries; tree: up to 1022 int32; *n_tree = entries written. */
int lc_huff_build_decode(const uint32_t* code, const uint8_t* len, const uint8_t* used, int lut_bits, uint16_t* lut,
                         int32_t* tree, int* n_tree);
/* Decode: d_lut is a 2^lut_bits table of (len << 8 | sym) for codes no longer than lut_bits;
 * longer codes walk d_tree (node = {child0, child1}, leaves are ~sym).  d_bits is n_bytes long (a multiple of 4,
 * zero padded by at least 16 bytes past the last code bit); reads are clamped to it. */
int lc_huff_decode(const uint8_t* d_bits, int64_t n_bytes, int64_t n, int seg, const unsigned long long* d_seg_bits,
                   const uint16_t* d_lut, int lut_bits, const int32_t* d_tree, uint8_t* d_sym,
                   void* stream);

/* synthetic ERA5-like field generated on the device for benchmarks:
 * (T,H,W,2) float32 window starting at global time t0 of level `level`. */
int lc_synth_field(float* d_x, int T, int H, int W, int t0, int level, uint64_t seed, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LOSSYCOMP_B200_H */
