"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.

A NumPy (integer/byte/elementwise stages) + torch-CPU fp32 (convolutions)
restatement of the reference codec path of SilkeDH/lossy-ml:

    lossycomp/compress.py:21-179      -> compress()
    lossycomp/decompress.py:21-108    -> decompress()
    lossycomp/dataLoader.py:437-532   -> chunk_data(), merge_data()
    lossycomp/models.py:47-97         -> encoder_forward(), decoder_forward()
    lossycomp/huffman.py:20-85        -> huffman_code(), huffman_decode()

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module, and only as the checker / the timed CPU baseline.
The product (lossy-ml_b200/) never imports it and has no CPU fallback.

Pinning: every non-NN stage is pinned bit-exactly against the
reference's own compress.py/decompress.py executed unmodified under import stubs
(tools/make_goldens.py -> tests/golden/*.npz) and against the pure-function KATs
the reference offers (chunk counts, merge(chunk(x)) == x, getHuffmanCode strings).
The convolution numerics (TensorFlow/Keras in the reference) cannot be run here.
The CONVENTIONS of this torch-CPU fp32 restatement (padding split, transposed-conv
crop, kernel layouts, residual adds) are pinned independently of torch by
tests/test_conv_definition.py (a tap-by-tap NumPy fp64 statement of the Keras
definition, all layers of model_2 and of the k=5 architecture, <= 1e-5) and on the
GPU box against cuDNN; what stays "parity unpinned" is TensorFlow's own summation
order (last bits of the conv outputs) and the fpzip byte stream of blob slot 0
(fpzip is absent; bz2 stands in, lossless).
"""
from __future__ import annotations

import bz2
import pickle

import numpy as np

CHUNK = (16, 48, 48)


# --------------------------------------------------------------------------
# dataLoader.py:437-532 -- chunk / merge with re-anchored (overlapping) tail windows
# --------------------------------------------------------------------------
def chunk_grid(shape, size=CHUNK):
    """Number of chunks per axis (dataLoader.py:449)."""
    return tuple(-(-int(shape[a]) // int(size[a])) for a in range(3))


def chunk_origin(idx, shape, size=CHUNK):
    """Start of chunk `idx` (lon fastest, dataLoader.py:458) on every axis; a window that would
    overrun the axis is anchored at axis_len - size instead (dataLoader.py:460-479)."""
    grid = chunk_grid(shape, size)
    ix = np.unravel_index(idx, grid)
    return tuple(int(min(ix[a] * size[a], shape[a] - size[a])) for a in range(3))


def chunk_data(data, size=CHUNK):
    assert all(data.shape[a] >= size[a] for a in range(3)), "Chunks size cannot be larger than the data size."
    grid = chunk_grid(data.shape, size)
    out = np.empty((int(np.prod(grid)),) + tuple(size[:3]) + (data.shape[3],), dtype=data.dtype)
    for i in range(out.shape[0]):
        t0, h0, w0 = chunk_origin(i, data.shape, size)
        out[i] = data[t0:t0 + size[0], h0:h0 + size[1], w0:w0 + size[2]]
    return out


def merge_data(chunks, size):
    """Inverse of chunk_data (dataLoader.py:484-532): a tail chunk contributes only its last
    `len - idx*cs` entries along the overrunning axis; all other voxels come from the chunk
    that owns them at its natural position."""
    cs = chunks.shape[1:4]
    grid = chunk_grid(size, cs)
    out = np.empty(tuple(size[:3]) + (chunks.shape[4],), dtype=chunks.dtype)
    for k in range(grid[0]):
        for j in range(grid[1]):
            for i in range(grid[2]):
                c = chunks[(k * grid[1] + j) * grid[2] + i]
                lo = (k * cs[0], j * cs[1], i * cs[2])
                n = tuple(min(cs[a], size[a] - lo[a]) for a in range(3))   # 'mn' in select_chunk
                out[lo[0]:lo[0] + n[0], lo[1]:lo[1] + n[1], lo[2]:lo[2] + n[2]] = \
                    c[cs[0] - n[0]:, cs[1] - n[1]:, cs[2] - n[2]:]
    return out


# --------------------------------------------------------------------------
# models.py:47-97 -- Autoencoder2 forward, torch CPU fp32
# --------------------------------------------------------------------------
def _same_pad(n, k, s):
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return total // 2, total - total // 2


def _conv_layer(x, L):
    """x: torch (N,C,T,H,W) float32. L: lossycomp.models.Layer (Keras-layout weights)."""
    import torch
    import torch.nn.functional as F
    w = torch.from_numpy(np.ascontiguousarray(L.kernel.transpose(4, 3, 0, 1, 2))).to(x.device)
    b = torch.from_numpy(L.bias).to(x.device)
    if L.kind == "conv":
        pads = []
        for n in reversed(x.shape[2:]):           # F.pad wants last axis first
            lo, hi = _same_pad(int(n), L.k, L.stride)
            pads += [lo, hi]
        return F.conv3d(F.pad(x, pads), w, b, stride=L.stride)
    # Conv3DTranspose, padding="same": full transposed conv cropped to 2*in from (k-2)//2.
    # Keras kernel (kt,kh,kw,C_out,C_in) -> torch conv_transpose weight (C_in,C_out,kt,kh,kw).
    y = F.conv_transpose3d(x, w, None, stride=L.stride)
    c = (L.k - L.stride) // 2
    t, h, ww = (int(n) * L.stride for n in x.shape[2:])
    y = y[:, :, c:c + t, c:c + h, c:c + ww]
    return y + b.view(1, -1, 1, 1, 1)


def run_layers(x_ndhwc, layers, batch=10, collect=None, device="cpu"):
    """Run a list of layers on channels-last numpy input (N,T,H,W,C) -> numpy channels-last.
    device="cuda" runs the same layer statements through cuDNN (tests only: an independent cross-check of the
    CUDA kernels, never a product path)."""
    import torch
    outs = []
    with torch.no_grad():
        for s in range(0, x_ndhwc.shape[0], batch):      # compress.py:84-86 (batch 10)
            x = torch.from_numpy(np.array(x_ndhwc[s:s + batch], dtype=np.float32, order="C")).permute(0, 4, 1, 2, 3).contiguous().to(device)
            skip = None
            for L in layers:
                y = _conv_layer(x, L)
                if L.res_add:
                    y = y + skip
                if L.relu:
                    y = torch.relu(y)
                if L.res_save:
                    skip = y
                x = y
                if collect is not None:
                    collect.setdefault(L.name, []).append(x.permute(0, 2, 3, 4, 1).cpu().numpy().copy())
            outs.append(x.permute(0, 2, 3, 4, 1).contiguous().cpu().numpy())
    return np.concatenate(outs, axis=0)


def encoder_forward(model, chunks):
    return run_layers(chunks, model.encoder)


def decoder_forward(model, latent):
    return run_layers(latent, model.decoder)


# --------------------------------------------------------------------------
# compress.py:64-168
# --------------------------------------------------------------------------
def standardise(array):
    """compress.py:64-67: float32 mean/std of channel 0 over the whole array, channel 0 only."""
    mean = np.mean(array[:, :, :, 0])
    std = np.std(array[:, :, :, 0])
    a = array.copy()
    a[:, :, :, 0] = (array[:, :, :, 0] - mean) / std
    return a, mean, std


def residual_codes(array, recon_std, mean, std, err_threshold):
    """compress.py:114-135.  recon_std: merged decoder output (T,H,W,1) in standardised units.
    Returns (mask int8 (T,H,W), q int32 (n_out,), recon float32 (T,H,W,1))."""
    recon = ((recon_std * std) + mean).astype(np.float32)
    error = (array[:, :, :, 0] - recon[:, :, :, 0]).astype(np.float32)
    thr = np.float32(err_threshold)       # python float vs float32 array compares in float32
    mask = np.zeros(error.shape, dtype=np.int8)
    mask[error > thr] = 1
    mask[error < -thr] = 2
    a = np.abs(error)
    a = a[a > thr]                         # C-order compaction of outliers
    thr2 = np.float32(err_threshold * 2.0)
    q = np.round(a / thr2).astype(np.int32)   # half-to-even, float32 divide
    return mask, q, recon


def first_difference(v):
    """compress.py:147-151 / :159-163: [v0, v1-v0, ...] in the array's own integer width
    (int8 arithmetic wraps, which cumsum in int64 + the small value range undoes)."""
    v = v.reshape(-1)
    if v.size == 0:
        return v.copy()
    return np.concatenate(([v[0]], np.diff(v))).astype(v.dtype)


def compress(array, err_threshold, model, latent_codec=None, inject=None, return_parts=False):
    """Restatement of compress.py:21-179 with `model` (lossycomp.models.Model) in place of the
    Keras model.  `inject` may carry {"mean","std"} and/or {"recon_std"} to pin stages."""
    assert err_threshold != 0, "The absolute error can't be 0."
    if array.dtype != np.float32:
        array = np.array(array, dtype=np.float32)
    assert array.shape[3] == 2, "Input should have 2 channels."
    inject = inject or {}
    array_std, mean, std = standardise(array)
    if "mean" in inject:
        mean, std = np.float32(inject["mean"]), np.float32(inject["std"])
        array_std = array.copy()
        array_std[:, :, :, 0] = (array[:, :, :, 0] - mean) / std
    chunks = chunk_data(array_std, CHUNK)
    latent = encoder_forward(model, chunks)                  # (N,1,3,3,F)
    if "recon_std" in inject:
        recon_std = inject["recon_std"]
    else:
        recon_std = merge_data(decoder_forward(model, latent), array.shape)
    mask, q, recon = residual_codes(array, recon_std, mean, std, err_threshold)
    thr2 = err_threshold * 2.0
    latent_s = np.squeeze(latent, axis=1)
    comp_data = (latent_codec or _latent_pack)(latent_s)
    dq = first_difference(q.astype(np.int32))
    dm = first_difference(mask.reshape(-1))
    comp_error = bz2.compress(dq.tobytes(), 9)
    comp_mask = bz2.compress(dm.tobytes(), 9)
    blob = pickle.dumps([comp_data, latent_s.shape, array.shape, dq.shape, mean, std,
                         comp_error, comp_mask, thr2])
    if return_parts:
        return blob, dict(mean=mean, std=std, latent=latent, recon_std=recon_std, recon=recon,
                          mask=mask, q=q, dq=dq, dm=dm)
    return blob


def _latent_pack(latent_s):
    # fpzip (compress.py:144) is not installable here; a lossless stand-in keeps the slot typed bytes.
    return bz2.compress(np.ascontiguousarray(latent_s, dtype="<f4").tobytes(), 9)


def _latent_unpack(b):
    return bz2.decompress(b)


# --------------------------------------------------------------------------
# decompress.py:54-108
# --------------------------------------------------------------------------
def reconstruct(recon_std, mean, std, mask, q, thr2):
    """decompress.py:86-104 given the un-diffed mask (T,H,W,1) and q (n_out,)."""
    error = q.astype(np.int32) * thr2                 # float64
    val = np.array(mask, dtype=np.float32)
    neg = np.where(mask == 2)
    val[val > 0] = error                              # float64 -> float32 on store
    val[neg] = val[neg] * -1.0
    val = val.astype(np.float32)
    out = (recon_std * std) + mean
    return (out + val).astype(np.float32)


def decompress(blob, model, latent_codec=None, inject=None):
    comp_data, lshape, array_size, error_shape, mean, std, comp_error, comp_mask, thr2 = pickle.loads(blob)
    latent = np.frombuffer((latent_codec or _latent_unpack)(comp_data), dtype=np.float32).reshape(lshape)
    latent = np.expand_dims(latent, axis=1)
    if inject and "recon_std" in inject:
        recon_std = inject["recon_std"]
    else:
        recon_std = merge_data(decoder_forward(model, latent), array_size)
    q = np.frombuffer(bz2.decompress(comp_error), dtype=np.int32).reshape(error_shape).cumsum()
    n = int(np.prod(recon_std.shape))
    mask = np.frombuffer(bz2.decompress(comp_mask), dtype=np.byte).reshape((n,)).cumsum().reshape(recon_std.shape)
    return reconstruct(recon_std, mean, std, mask, q, thr2)


# --------------------------------------------------------------------------
# huffman.py:20-85 -- the reference's (non-canonical) Huffman table construction
# --------------------------------------------------------------------------
def huffman_table(freqs):
    """freqs: dict symbol -> count.  Leaves are ordered by (count, symbol) (huffman.py:71); the
    queue is stably re-sorted by count before every merge, the two smallest are popped
    (first popped = left = '0') and the parent is appended at the end (huffman.py:20-33).
    Returns dict symbol -> bit string; a single-symbol alphabet gets the empty code."""
    leaves = sorted(freqs.items(), key=lambda kv: (kv[1], kv[0]))
    n = len(leaves)
    parent = {}
    queue = [(leaves[i][1], i) for i in range(n)]     # (count, node id); ids >= n are internal
    nxt = n
    while len(queue) > 1:
        queue.sort(key=lambda it: it[0])              # stable, like list.sort(key=freq)
        (fa, a), (fb, b) = queue.pop(0), queue.pop(0)
        parent[a] = (nxt, "0")
        parent[b] = (nxt, "1")
        queue.append((fa + fb, nxt))
        nxt += 1
    codes = {}
    for i, (sym, _) in enumerate(leaves):
        bits, node = "", i
        while node in parent:
            node, bit = parent[node][0], parent[node][1]
            bits = bit + bits
        codes[sym] = bits
    return codes


def huffman_code(string):
    """getHuffmanCode (huffman.py:63-85): returns [bit string, {char: code}]."""
    freqs = {}
    for ch in string:
        freqs[ch] = freqs.get(ch, 0) + 1
    table = huffman_table(freqs)
    return ["".join(table[ch] for ch in string), table]


def huffman_decode(bits, table):
    """decode_huffman (huffman.py:48-60): greedy prefix match."""
    inv = {v: k for k, v in table.items()}
    out, cur = [], ""
    for b in bits:
        cur += b
        if cur in inv:
            out.append(inv[cur])
            cur = ""
    return out


# --------------------------------------------------------------------------
# synthetic ERA5-like field -- CPU twin of the device generator
# --------------------------------------------------------------------------
def synthetic_field(T, H, W, seed=1234, level=0, t0=0, h0=0, w0=0, grid=(720, 1440)):
    """(T,H,W,2) float32 window of a synthetic global field on the `grid` lat/lon mesh:
    channel 0 temperature-like, channel 1 a land-sea mask in {0,1}."""
    GH, GW = grid
    lat = (np.pi / 2 - (np.arange(H, dtype=np.float64) + h0) * (np.pi / (GH - 1)))[None, :, None]
    lon = ((np.arange(W, dtype=np.float64) + w0) * (2 * np.pi / GW) - np.pi)[None, None, :]
    t = (np.arange(T, dtype=np.float64) + t0)[:, None, None]
    x = (288.0 - 45.0 * np.sin(lat) ** 2 + 6.0 * np.sin(3 * lon + 0.4 * lat) * np.cos(lat)
         + 3.0 * np.cos(2 * np.pi * t / 24.0 - lon) * np.cos(lat) + 2.0 * np.sin(11 * lon + 7 * lat)
         + 0.8 * np.sin(37 * lon - 29 * lat + 0.3 * t) - 6.5 * level)
    rng = np.random.default_rng(seed + 7919 * level)
    x = x + rng.normal(0.0, 0.15, size=(T, H, W))
    lsm = (np.sin(2 * lon + 1) + np.cos(3 * lat) > 0.3).astype(np.float32)
    out = np.empty((T, H, W, 2), dtype=np.float32)
    out[..., 0] = x.astype(np.float32)
    out[..., 1] = np.broadcast_to(lsm, (T, H, W))
    return out


# ------------------------------------------------------------------------------------------
# Restatements of this package's own container extensions (not part of the reference): the run
# tokens of sparse streams (include/lossycomp_b200.h: lc_rle_tokens / lc_rle_expand).
# ------------------------------------------------------------------------------------------
def rle_tokens(b, run_sym=0, run_first=243):
    """Bytes -> run tokens: runs of `run_sym` inside 32-byte segments, decomposed largest power of two
    first; a run of 2^(k+1) becomes token run_first + k (k = 0..4), everything else stays literal."""
    b = np.asarray(b, dtype=np.uint8)
    out = []
    for s0 in range(0, len(b), 32):
        seg = b[s0:s0 + 32]
        i = 0
        while i < len(seg):
            if seg[i] != run_sym:
                out.append(int(seg[i]))
                i += 1
                continue
            r = 1
            while i + r < len(seg) and seg[i + r] == run_sym:
                r += 1
            for k in range(5, -1, -1):
                if r & (1 << k):
                    out.append(run_sym if k == 0 else run_first - 1 + k)
            i += r
    return np.array(out, dtype=np.uint8)


def rle_expand(tok, n, run_sym=0, run_first=243):
    """Inverse of rle_tokens (n = number of bytes to produce)."""
    out = np.full(n, run_sym, dtype=np.uint8)
    pos = 0
    for t in np.asarray(tok, dtype=np.uint8):
        r = int(t) - run_first
        if 0 <= r <= 4:
            pos += 2 << r
        else:
            out[pos] = t
            pos += 1
    assert pos == n, (pos, n)
    return out
