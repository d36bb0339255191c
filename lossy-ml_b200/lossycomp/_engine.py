"""Host orchestration of the B200 codec: marshals buffers and launches the C-ABI kernels.

PyTorch is used for device memory, streams and (in dist.py) torch.distributed only; all the
arithmetic is in liblossycomp_b200.so.  The stage order follows the reference
(compress.py:64-168, decompress.py:54-104).
"""
from __future__ import annotations

import bz2
import ctypes as C
import contextlib
import hashlib
import json
import math
import os
import struct
from typing import Dict, Optional

import numpy as np

from . import _native as N
from . import huffman as HF
from .models import CHUNK, Model, load_model

# fp16 operands / fp32 accumulation on the tensor cores: compression factor within 0.01 % of the fp32
# autoencoder on the benchmark field (bf16: -1.3 %); "fp32" selects the CUDA-core reference-grade path
DEFAULT_PRECISION = "fp16"
LATENT_MAGIC = b"LCL1"
HUFF_MAGIC = b"LCH1"
SEG = 1024                      # symbols per independently decodable Huffman segment
LATENT_SEG = 256                # ... of the small latent plane, decoded on decompress()'s critical path
ESC = 255                       # escape symbol of the code stream
MASK_RUN = (0, 243)             # run tokens of the mask stream: runs of byte 0, tokens 243..247
CODE_RUN = (1, 250)             # run tokens of the code stream: runs of code 1 (undifferenced), tokens 250..254
RLE_GATE = 0.2                  # zero-run tokens of the mask are only tried when more than this share of its bytes is zero
# Revision of the autoencoder kernels' arithmetic (summation order, rounding points).  The error bound only
# holds when the decoder reproduces the compressor's reconstruction bit for bit, so blobs record the
# revision they were written with and a build with different arithmetic refuses them.
NUMERICS_REV = 5      # 2: bias folded into the first block of a plane; 3: residual add inside the MMA accumulation;
# 4: the 3x3x3 layers accumulate their time taps in TMEM starting from the bias (sweep_acc_kernel); 5: so do the
# 3x3x3 layers with a residual input
# chunks per convolution launch: large batches amortise kernel prologues and tails (900 chunks x 19 tiles =
# 57.8 waves of the 296 resident CTAs); the activation workspace is 19.4 MB per chunk
DEFAULT_BATCH = int(os.environ.get("LOSSYCOMP_BATCH", "1024"))


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


_PyBytes_New = C.pythonapi.PyBytes_FromStringAndSize
_PyBytes_New.restype = C.py_object
_PyBytes_New.argtypes = [C.c_void_p, C.c_ssize_t]
_PyBytes_Data = C.pythonapi.PyBytes_AsString
_PyBytes_Data.restype = C.c_void_p
_PyBytes_Data.argtypes = [C.py_object]


def _new_bytes(n: int):
    """An uninitialised bytes object of n bytes plus a writable uint8 view of it.  The C API allows filling a
    bytes object created with a NULL source before anything else sees it; the containers are built that way so
    that multi-megabyte payloads are written once, by the device->host copy itself."""
    obj = _PyBytes_New(None, n)
    if n == 0:
        return obj, np.zeros(0, dtype=np.uint8)
    view = np.ctypeslib.as_array((C.c_ubyte * n).from_address(_PyBytes_Data(obj)))
    return obj, view


class DevPieces:
    """A byte string whose large parts are still on the device: an ordered list of host pieces (bytes-like) and
    uint8 CUDA tensors.  The multi-GPU path sends the device parts over NCCL straight from the buffers the
    Huffman packer wrote (dist.compress_sharded); tobytes() is the single-GPU materialisation."""

    __slots__ = ("pieces",)

    def __init__(self, pieces):
        self.pieces = [p for p in pieces if _plen(p)]

    def __len__(self):
        return sum(_plen(p) for p in self.pieces)

    def tobytes(self) -> bytes:
        return b"".join(bytes(p) if not hasattr(p, "is_cuda") else p.cpu().numpy().tobytes() for p in self.pieces)


def _plen(p):
    return int(p.numel()) if hasattr(p, "numel") else len(p)


class Engine:
    """One codec instance per (device, model, precision); replaces the per-call Keras build +
    load_weights of compress.py:41-57 / decompress.py:33-48 by a one-time weight upload."""

    _cache: Dict[tuple, "Engine"] = {}

    @classmethod
    def get(cls, model_path: Optional[str] = None, precision: Optional[str] = None,
            device: Optional[int] = None, lane: int = 0) -> "Engine":
        """lane > 0: a further, independent instance for the same (model, precision, device) -- lossycomp.lanes
        keeps one field in flight per instance."""
        torch = N.require_cuda()
        precision = precision or os.environ.get("LOSSYCOMP_PRECISION", DEFAULT_PRECISION)
        device = torch.cuda.current_device() if device is None else device
        key = (model_path or "default", precision, device) + ((lane,) if lane else ())
        if key not in cls._cache:
            cls._cache[key] = Engine(load_model(model_path), precision, device)
        return cls._cache[key]

    def __init__(self, model: Model, precision: str = "fp32", device: int = 0):
        self.torch = N.require_cuda()
        self.lib = N.lib()
        self.model = model
        self.precision = precision
        self.device = device
        self.dev = self.torch.device("cuda", device)
        layers = model.layers
        arr = (N.LcLayer * len(layers))()
        self._keep = []
        wh = hashlib.sha1()
        for i, L in enumerate(layers):
            k = np.ascontiguousarray(L.kernel, dtype=np.float32)
            b = np.ascontiguousarray(L.bias, dtype=np.float32)
            wh.update(k.tobytes())
            wh.update(b.tobytes())
            self._keep += [k, b]
            arr[i] = N.LcLayer(0 if L.kind == "conv" else 1, L.k, L.stride, L.cin, L.cout, int(L.relu),
                               int(L.res_save), int(L.res_add), k.ctypes.data, b.ctypes.data)
        h = C.c_void_p()
        N.check(self.lib.lc_create(arr, len(model.encoder), len(model.decoder), model.arch.in_channels,
                                   N.PRECISIONS[precision], device, C.byref(h)), "lc_create")
        self.h = h
        self.weights_hash = wh.hexdigest()[:16]       # recorded in every blob: the decoder must be the compressor's
        self.latent_floats = self.lib.lc_latent_floats(h)
        self._ws = None
        self._ws_chunks = 0
        self._small = {}
        self.launches = 0          # kernels launched by this engine (bench.py gpu_launches)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.lc_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # -- per-layer kernel timing (bench.py roofline) --------------------------------
    def profile(self, on: bool):
        N.check(self.lib.lc_profile_enable(self.h, int(on)), "lc_profile_enable")

    BACKENDS = {0: "conv_direct/conv_small (CUDA core)", 1: "sweep_tc_kernel<K3>", 2: "sweep_tc_kernel<K4>",
                3: "sweep_tc_kernel<K4_OUT1>", 4: "sweep_tc_kernel<DOWN>", 5: "sweep_tc_kernel<UP>"}

    def layer_backends(self):
        return [self.lib.lc_layer_backend(self.h, i) for i in range(len(self.model.layers))]

    def profile_read(self):
        n = len(self.model.layers)
        ms = (C.c_double * n)()
        cnt = (C.c_longlong * n)()
        N.check(self.lib.lc_profile_read(self.h, ms, cnt, n), "lc_profile_read")
        return list(ms), list(cnt)

    # -- buffers ---------------------------------------------------------------
    def _stream(self):
        # kernels launch on the calling thread's current device: make it this engine's (several engines on
        # different GPUs may live in one process)
        if self.torch.cuda.current_device() != self.device:
            self.torch.cuda.set_device(self.device)
        return C.c_void_p(self.torch.cuda.current_stream(self.dev).cuda_stream)

    def side_stream(self):
        """Second stream of this engine (latent packing behind the decoder, entropy decoding beside it)."""
        side = self._small.get("side_stream")
        if side is None:
            side = self._small["side_stream"] = self.torch.cuda.Stream(device=self.dev)
        return side

    def _turn(self):
        """Context of one autoencoder pass.  Engines that share a device through lossycomp.lanes take turns on the
        GPU for these passes (self.turn, a lanes.GpuTurn): the persistent convolution kernels want every SM (and
        whole SM pairs), so two passes running against each other only fragment the machine; what the lanes
        overlap is one field's host round trips and small kernels with the other field's convolutions."""
        turn = getattr(self, "turn", None)
        return turn.section(self) if turn is not None else contextlib.nullcontext()

    def _ae_ws(self, n_chunks):
        if self._ws is None or self._ws_chunks < n_chunks:
            nbytes = self.lib.lc_ae_workspace_bytes(self.h, n_chunks)
            if self._ws is not None:
                self.lib.lc_ae_workspace_release(self.h, _ptr(self._ws))
            self._ws = None
            self._ws = self.torch.empty(nbytes, dtype=self.torch.uint8, device=self.dev)
            # zeroes the halo / guard slots the 16-bit kernels rely on and registers the buffer with the handle
            N.check(self.lib.lc_ae_workspace_init(self.h, _ptr(self._ws), nbytes, self._stream()), "lc_ae_workspace_init")
            self._ws_chunks = n_chunks
        return self._ws

    def _buf(self, name, nbytes):
        b = self._small.get(name)
        if b is None or b.numel() < nbytes:
            b = self.torch.empty(max(int(nbytes), 256), dtype=self.torch.uint8, device=self.dev)
            self._small[name] = b
        return b

    # -- stages ------------------------------------------------------------------
    def stats_blocks(self, x):
        """Per-time-chunk partial statistics of channel 0: float64 device tensor (ceil(T/16), 4) of
        {sum x, sum x^2, max|x|, n}.  No sync.  Blocks are reduced independently in a fixed order and added in
        index order by combine_blocks(), so a time-sharded field reproduces the single-GPU mean/std bit for bit."""
        T, H, W, Cc = x.shape
        nblk = -(-T // CHUNK[0])
        out = self.torch.empty((nblk, 4), dtype=self.torch.float64, device=self.dev)
        ws = self._buf("stats", self.lib.lc_stats_workspace_bytes())
        N.check(self.lib.lc_stats_blocks(_ptr(x), T * H * W, CHUNK[0] * H * W, Cc, _ptr(out), _ptr(ws), self._stream()),
                "lc_stats_blocks")
        self.launches += 2 * nblk
        return out

    @staticmethod
    def combine_blocks(blocks):
        """[(sum, sum^2, max|x|, n)] per 16-step block, in time order -> (mean float32, std float32, max|x|):
        compress.py:64-65 with fp64 accumulation (blocks added sequentially in index order)."""
        s = s2 = n = 0.0
        mx = 0.0
        for b in blocks:
            s += b[0]
            s2 += b[1]
            mx = max(mx, b[2])
            n += b[3]
        mean64 = s / n
        var = max(s2 / n - mean64 * mean64, 0.0)
        return np.float32(mean64), np.float32(math.sqrt(var)), mx

    def stats(self, x):
        """compress.py:64-65 -> (mean float32, std float32, max|x|).  Sync (reads 32 bytes per time chunk)."""
        return self.combine_blocks(self.stats_blocks(x).cpu().tolist())

    def n_chunks(self, T, H, W):
        g = (C.c_int * 3)()
        N.check(self.lib.lc_chunk_grid(T, H, W, C.byref(g)), "lc_chunk_grid")
        return g[0] * g[1] * g[2]

    def encode(self, x, mean, std, batch=None):
        """compress.py:66-86: standardise + chunk + encoder -> latent (N, latent_floats) float32."""
        T, H, W, Cc = x.shape
        n = self.n_chunks(T, H, W)
        batch = min(batch or DEFAULT_BATCH, n)
        latent = self.torch.empty((n, self.latent_floats), dtype=self.torch.float32, device=self.dev)
        ws = self._ae_ws(batch)
        with self._turn():
            for b in range(0, n, batch):
                c = min(batch, n - b)
                N.check(self.lib.lc_encode(self.h, _ptr(x), T, H, W, Cc, float(mean), float(std), b, c,
                                           C.c_void_p(latent.data_ptr() + b * self.latent_floats * 4),
                                           _ptr(ws), ws.numel(), self._stream()), "lc_encode")
                self.launches += 1 + len(self.model.encoder)
        return latent

    def decode(self, latent, T, H, W, batch=None):
        """compress.py:99-111 / decompress.py:66-72: decoder + merge -> recon (T,H,W) float32."""
        n = self.n_chunks(T, H, W)
        assert latent.shape[0] == n, "latent does not match the array size"
        batch = min(batch or DEFAULT_BATCH, n)
        recon = self.torch.empty((T, H, W), dtype=self.torch.float32, device=self.dev)
        ws = self._ae_ws(batch)
        with self._turn():
            for b in range(0, n, batch):
                c = min(batch, n - b)
                N.check(self.lib.lc_decode(self.h, C.c_void_p(latent.data_ptr() + b * self.latent_floats * 4),
                                           T, H, W, b, c, _ptr(recon), _ptr(ws), ws.numel(), self._stream()),
                        "lc_decode")
                self.launches += 1 + len(self.model.decoder)
        return recon

    def quantize(self, x, recon, mean, std, thr, check_abs_err=None):
        """compress.py:114-135.  thr is the (possibly guard-banded) absolute bound as a Python float;
        the float32 forms are made here exactly as NumPy makes them.  Sync (reads the count).
        With check_abs_err the same pass also decodes every element the way reconstruct() will and
        returns (mask, q, violations of |x - xhat| <= check_abs_err, max |x - xhat|)."""
        return self.quantize_result(self.quantize_launch(x, recon, mean, std, thr, check_abs_err))

    def quantize_launch(self, x, recon, mean, std, thr, check_abs_err=None):
        """The kernel of quantize() without its host sync: -> pending tuple for quantize_result()."""
        T, H, W, Cc = x.shape
        n = T * H * W
        mask = self.torch.empty(n, dtype=self.torch.int8, device=self.dev)
        q = self.torch.empty(n, dtype=self.torch.int32, device=self.dev)
        ws = self._buf("scan", self.lib.lc_scan_workspace_bytes(n))
        thr32 = np.float32(thr)
        thr2_32 = np.float32(thr * 2.0)
        self.launches += 1
        if check_abs_err is None:
            cnt = self.torch.zeros(1, dtype=self.torch.int64, device=self.dev)
            N.check(self.lib.lc_quantize(_ptr(x), Cc, _ptr(recon), n, float(mean), float(std), float(thr32),
                                         float(thr2_32), _ptr(mask), _ptr(q), _ptr(cnt), _ptr(ws), self._stream()),
                    "lc_quantize")
            return (mask, q, cnt, False)
        out3 = self.torch.empty(3, dtype=self.torch.int64, device=self.dev)
        N.check(self.lib.lc_quantize_checked(_ptr(x), Cc, _ptr(recon), n, float(mean), float(std), float(thr32),
                                             float(thr2_32), float(thr * 2.0), float(check_abs_err), _ptr(mask),
                                             _ptr(q), _ptr(out3), _ptr(ws), self._stream()), "lc_quantize_checked")
        return (mask, q, out3, True)

    def quantize_result(self, pending):
        """Sync: (mask, q[:n_out]) or, for the checked form, (mask, q[:n_out], violations, max |x - xhat|)."""
        mask, q, out, checked = pending
        if not checked:
            return mask, q[:int(out.item())]
        n_out, viol, bits = out.cpu().tolist()
        return mask, q[:n_out], int(viol), struct.unpack("<d", struct.pack("<q", bits))[0]

    def reconstruct(self, recon, mask, q, mean, std, thr2):
        """decompress.py:86-104 -> (T,H,W) float32."""
        n = recon.numel()
        out = self.torch.empty_like(recon)
        ws = self._buf("scan_rec", self.lib.lc_scan_workspace_bytes(n))
        N.check(self.lib.lc_reconstruct(_ptr(recon), _ptr(mask), _ptr(q), q.numel(), n, float(mean), float(std),
                                        float(thr2), _ptr(out), _ptr(ws), self._stream()), "lc_reconstruct")
        self.launches += 1
        self._rec_check = (ws, int(q.numel()))
        return out

    def check_reconstruct(self):
        """After the stream has been synchronised: the mask of the last reconstruct() must hold exactly as many
        outliers as there were codes (a damaged blob otherwise decodes silently wrong)."""
        chk = getattr(self, "_rec_check", None)
        if chk is None:
            return
        self._rec_check = None
        got = int(chk[0][:8].view(self.torch.int64).item())
        if got != chk[1]:
            raise N.NativeError(f"corrupt blob: the mask marks {got} outliers, the code stream holds {chk[1]}")

    def verify(self, x, xhat, abs_err):
        """Exhaustive fp64 bound check -> (violations, max |x - xhat|).  Sync."""
        T, H, W, Cc = x.shape
        out = self.torch.empty(2, dtype=self.torch.float64, device=self.dev)
        ws = self._buf("stats", self.lib.lc_stats_workspace_bytes())
        N.check(self.lib.lc_verify_bound(_ptr(x), Cc, _ptr(xhat), T * H * W, float(abs_err), _ptr(out),
                                         _ptr(ws), self._stream()), "lc_verify_bound")
        self.launches += 2
        v, mx = out.cpu().tolist()
        return int(v), mx

    # -- first difference / cumsum (compress.py:147-163, decompress.py:78,83) ------
    def diff(self, v):
        out = self.torch.empty_like(v)
        if v.numel():
            fn = self.lib.lc_diff_i32 if v.dtype == self.torch.int32 else self.lib.lc_diff_i8
            N.check(fn(_ptr(v), v.numel(), _ptr(out), self._stream()), "lc_diff")
            self.launches += 1
        return out

    def cumsum(self, v):
        out = self.torch.empty_like(v)
        if v.numel():
            ws = self._buf("scan", self.lib.lc_scan_workspace_bytes(v.numel()))
            fn = self.lib.lc_cumsum_i32 if v.dtype == self.torch.int32 else self.lib.lc_cumsum_i8
            N.check(fn(_ptr(v), v.numel(), _ptr(out), _ptr(ws), self._stream()), "lc_cumsum")
            self.launches += 1
        return out

    # -- entropy stage ---------------------------------------------------------------
    def hist(self, sym):
        h = self.torch.empty(256, dtype=self.torch.int32, device=self.dev)
        N.check(self.lib.lc_hist_u8(_ptr(sym), sym.numel(), _ptr(h), self._stream()), "lc_hist_u8")
        self.launches += 1
        return h.cpu().numpy().astype(np.int64) & 0xFFFFFFFF

    def huff_pack(self, sym, table: HF.Table, total_bits=None):
        """Symbols (uint8 device tensor) -> (payload bytes, per-segment bit lengths uint32, total bits).
        `total_bits` (known from the histogram) sizes the zero-initialised output exactly."""
        t = self.torch
        n = sym.numel()
        nseg = (n + SEG - 1) // SEG
        cap = self.lib.lc_huff_max_bytes(n) if total_bits is None else ((total_bits + 31) // 32) * 4 + 16
        out = t.zeros(cap, dtype=t.uint8, device=self.dev)
        segb = t.empty(max(nseg, 1) + 1, dtype=t.int64, device=self.dev)
        code = t.from_numpy(table.code.view(np.int32).copy()).to(self.dev)
        ln = t.from_numpy(table.len.copy()).to(self.dev)
        ws = self._buf("scan", self.lib.lc_huff_workspace_bytes(n, SEG))
        N.check(self.lib.lc_huff_encode(_ptr(sym), n, _ptr(code), _ptr(ln), SEG, _ptr(out), _ptr(segb),
                                        C.c_void_p(segb.data_ptr() + 8 * max(nseg, 1)), _ptr(ws), self._stream()),
                "lc_huff_encode")
        self.launches += 1
        offs = segb.cpu().numpy()                      # [segment offsets..., total bits]
        got_bits = int(offs[-1])
        assert total_bits is None or got_bits == total_bits
        nbytes = (got_bits + 7) // 8
        payload = out[:nbytes].cpu().numpy().tobytes()
        lens = np.diff(np.concatenate((offs[:nseg], [got_bits]))).astype(np.uint16) if nseg else \
            np.zeros(0, np.uint16)
        return payload, lens, got_bits

    def _stage_ring(self, nbytes):
        """A pinned staging buffer that is safe to overwrite: a ring of four, each guarded by the event recorded
        after its last host->device copy."""
        t = self.torch
        ring = self._small.setdefault("stage_ring", [])
        i = self._small["stage_next"] = (self._small.get("stage_next", -1) + 1) % 4
        while len(ring) <= i:
            ring.append([None, None])
        buf, ev = ring[i]
        if ev is not None:
            ev.synchronize()
        # every slot is sized for the largest payload seen so far: slots rotate among streams of very different
        # sizes (latent plane, codes, mask), and a pinned reallocation in the steady state costs milliseconds
        cap = self._small["stage_cap"] = max(self._small.get("stage_cap", 1 << 16), int(nbytes * 1.25))
        if buf is None or buf.numel() < nbytes:
            buf = t.empty(cap, dtype=t.uint8, pin_memory=True)
        ring[i] = [buf, None]
        return buf, ring[i]

    def huff_unpack(self, payload, n: int, seg_lens: np.ndarray, table: HF.Table, seg: int = SEG):
        """Huffman bits (bytes-like, or a uint8 CUDA tensor that is already on this device) -> n uint8 symbols on
        the device.  Host payloads, segment offsets and the decode tables travel in ONE pinned staging buffer and
        one asynchronous copy; a device payload is copied device-to-device into a word-aligned buffer with the
        16 spare bytes the decoder's look-ahead reads."""
        t = self.torch
        sym = t.empty(n, dtype=t.uint8, device=self.dev)
        if n == 0:
            return sym
        on_dev = hasattr(payload, "is_cuda")
        npay = int(payload.numel()) if on_dev else len(payload)
        pay_cap = ((npay + 3) // 4) * 4 + 16                      # the decoder reads whole words past the end
        offs = np.zeros(max(seg_lens.size, 1), dtype=np.int64)
        if seg_lens.size > 1:
            np.cumsum(seg_lens[:-1], dtype=np.int64, out=offs[1:])
        if int(offs[-1]) + int(seg_lens[-1] if seg_lens.size else 0) > 8 * npay:
            raise N.NativeError("corrupt 'LCH1' container: the segment lengths exceed the payload")
        lut, tree = table.decode_arrays_native()
        o_seg = 0 if on_dev else (pay_cap + 7) & ~7
        o_lut = o_seg + 8 * offs.size
        o_tree = (o_lut + 2 * lut.size + 3) & ~3
        total = o_tree + 4 * tree.size
        host, slot = self._stage_ring(total)
        hv = host.numpy()
        if not on_dev:
            if npay:
                hv[:npay] = np.frombuffer(payload, dtype=np.uint8)
            hv[npay:pay_cap] = 0
        hv[o_seg:o_lut] = offs.view(np.uint8)
        hv[o_lut:o_lut + 2 * lut.size] = lut.view(np.uint8)
        hv[o_tree:total] = tree.view(np.uint8)
        dbuf = t.empty(total, dtype=t.uint8, device=self.dev)
        dbuf.copy_(host[:total], non_blocking=True)
        ev = t.cuda.Event()
        ev.record(t.cuda.current_stream(self.dev))
        slot[1] = ev
        base = dbuf.data_ptr()
        if on_dev:
            dpay = t.empty(pay_cap, dtype=t.uint8, device=self.dev)
            dpay[:npay].copy_(payload, non_blocking=True)
            dpay[npay:].zero_()
            bits_ptr = dpay.data_ptr()
        else:
            bits_ptr = base
        N.check(self.lib.lc_huff_decode(C.c_void_p(bits_ptr), pay_cap, n, seg, C.c_void_p(base + o_seg), C.c_void_p(base + o_lut),
                                        HF.LUT_BITS, C.c_void_p(base + o_tree), _ptr(sym), self._stream()), "lc_huff_decode")
        self.launches += 1
        return sym

    def pack_stream(self, sym, extra: dict, hist=None) -> bytes:
        """uint8 symbols -> self-describing 'LCH1' container."""
        n = sym.numel()
        if hist is None:
            hist = self.hist(sym) if n else np.zeros(256, np.int64)
        table = HF.build_table_native(hist)
        bits = table.encoded_bits(hist)
        payload, lens, total_bits = self.huff_pack(sym, table, bits) if n else (b"", np.zeros(0, np.uint16), 0)
        meta = dict(extra, n=int(n), seg=SEG, bits=int(total_bits), ref_tree=bool(table.reference_tree))
        mj = json.dumps(meta).encode()
        tb = table.to_bytes()
        return b"".join([HUFF_MAGIC, struct.pack("<III", len(mj), len(tb), lens.size), mj, tb,
                         lens.tobytes(), payload])

    @staticmethod
    def stream_head_len(blob) -> int:
        """Bytes of an 'LCH1' container in front of its Huffman payload (magic, lengths, JSON, table, segment
        lengths)."""
        mv = memoryview(blob)
        if bytes(mv[:4]) != HUFF_MAGIC:
            raise N.NativeError("not an 'LCH1' container")
        lm, ltb, nl = struct.unpack("<III", mv[4:16])
        return 16 + lm + ltb + 2 * nl

    def unpack_stream(self, blob, dev_payload=None):
        """'LCH1' container -> (symbols on the device, header dict).  With dev_payload (uint8 CUDA tensor) `blob`
        holds only the part in front of the payload and the payload is taken from the device."""
        mv = memoryview(blob)
        if bytes(mv[:4]) != HUFF_MAGIC:
            raise N.NativeError("not an 'LCH1' container")
        lm, ltb, nl = struct.unpack("<III", mv[4:16])
        p = 16
        meta = json.loads(bytes(mv[p:p + lm])); p += lm
        table, _ = HF.Table.from_bytes(mv[p:p + ltb]); p += ltb
        lens = np.frombuffer(mv, dtype=np.uint16, count=nl, offset=p); p += 2 * nl
        n, seg = int(meta["n"]), int(meta.get("seg", SEG))
        if n < 0 or seg < 64 or seg > 4096 or seg & (seg - 1) or nl != (n + seg - 1) // seg:
            raise N.NativeError("corrupt 'LCH1' container: symbol count, segment size and segment table disagree")
        payload = dev_payload if dev_payload is not None else mv[p:]
        sym = self.huff_unpack(payload, n, lens, table, seg)   # host payload: copied once, into pinned memory
        return sym, meta

    def pack_mask(self, mask) -> bytes:
        """3-state mask -> five trits per byte -> Huffman ('LCH1', kind=mask)."""
        t = self.torch
        n = mask.numel()
        sym = t.empty((n + 4) // 5, dtype=t.uint8, device=self.dev)
        N.check(self.lib.lc_pack_trits(_ptr(mask), n, _ptr(sym), self._stream()), "lc_pack_trits")
        self.launches += 1
        return self.pack_stream(sym, {"kind": "mask5", "n_mask": int(n)})

    def unpack_mask(self, blob: bytes, dev_payload=None):
        """Slot 7 -> int8 mask on the device.  With dev_payload, `blob` is the container without its Huffman
        payload, which is taken from the device tensor (multi-GPU decode)."""
        t = self.torch
        from .dist import unpack_multipart
        parts = unpack_multipart(blob)
        if len(parts) > 1:
            return t.cat([self.unpack_mask(p) for p in parts])
        sym, meta = self.unpack_stream(blob, dev_payload)
        n = meta["n_mask"]
        if meta.get("rle"):                  # zero-run tokens -> 5-trit bytes
            tok, ns = sym, int(meta["n_sym"])
            sym = t.empty(ns, dtype=t.uint8, device=self.dev)
            ws = self._buf("scan_rle", self.lib.lc_rle_workspace_bytes(max(ns, tok.numel())))
            N.check(self.lib.lc_rle_expand(_ptr(tok), tok.numel(), int(meta.get("run_sym", 0)),
                                           int(meta.get("run_first", 243)), _ptr(sym), ns, _ptr(ws), self._stream()),
                    "lc_rle_expand")
            self.launches += 1
        mask = t.empty(n, dtype=t.int8, device=self.dev)
        N.check(self.lib.lc_unpack_trits(_ptr(sym), n, _ptr(mask), self._stream()), "lc_unpack_trits")
        self.launches += 1
        return mask

    def pack_codes(self, q) -> bytes:
        """int32 quantisation codes -> byte symbols (+ escapes) -> Huffman.  The stream is coded
        either as is or first-differenced (the reference's transform, compress.py:147-151),
        whichever the histograms say is shorter; the choice is recorded in the header."""
        t = self.torch
        n = q.numel()
        cand = ((0, 0), (1, 127))                      # (diff, bias)
        h2 = t.empty(512, dtype=t.int32, device=self.dev)
        N.check(self.lib.lc_hist_codes(_ptr(q), n, ESC, cand[0][1], cand[1][1], _ptr(h2), self._stream()),
                "lc_hist_codes")
        self.launches += 1
        h2 = h2.cpu().numpy().astype(np.int64) & 0xFFFFFFFF
        hists = (h2[:256], h2[256:])
        costs = [HF.build_table_native(h).encoded_bits(h) + 32 * int(h[ESC]) for h in hists]
        pick = 0 if (costs[0] <= costs[1] or n == 0) else 1
        diff, bias = cand[pick]
        ne = int(hists[pick][ESC])
        sym = t.empty(n, dtype=t.uint8, device=self.dev)
        esc = t.empty(max(ne, 1), dtype=t.int32, device=self.dev)
        nesc = t.zeros(1, dtype=t.int64, device=self.dev)
        ws = self._buf("scan", self.lib.lc_scan_workspace_bytes(max(n, 1)))
        N.check(self.lib.lc_symbolize_i32(_ptr(q), n, diff, bias, ESC, _ptr(sym), _ptr(esc), _ptr(nesc),
                                          _ptr(ws), self._stream()), "lc_symbolize_i32")
        self.launches += 1
        body = self.pack_stream(sym, {"kind": "codes", "diff": diff, "bias": bias, "esc": ESC, "n_esc": ne},
                                hist=hists[pick])
        escb = esc[:ne].cpu().numpy().astype("<i4").tobytes() if ne else b""
        return body + escb

    def unpack_codes(self, blob: bytes, dev_payload=None):
        """Slot 6 -> int32 codes on the device (dev_payload: see unpack_mask; the escape list stays in `blob`)."""
        t = self.torch
        from .dist import unpack_multipart
        parts = unpack_multipart(blob)
        if len(parts) > 1:
            return t.cat([self.unpack_codes(p) for p in parts])
        # the escape list trails the container; its length is in the header
        lm = struct.unpack("<I", blob[4:8])[0]
        meta = json.loads(bytes(blob[16:16 + lm]))
        ne = int(meta["n_esc"])
        mv = memoryview(blob)
        if ne < 0 or 4 * ne > len(mv):
            raise N.NativeError("corrupt code container: escape count exceeds the container")
        body, escb = (mv[:len(mv) - 4 * ne], mv[len(mv) - 4 * ne:]) if ne else (mv, b"")
        sym, meta = self.unpack_stream(body, dev_payload)
        n = meta["n"]
        if meta.get("rle"):                  # run tokens -> code symbols
            tok, n = sym, int(meta["n_sym"])
            sym = t.empty(n, dtype=t.uint8, device=self.dev)
            ws = self._buf("scan_rle", self.lib.lc_rle_workspace_bytes(max(n, tok.numel())))
            N.check(self.lib.lc_rle_expand(_ptr(tok), tok.numel(), int(meta["run_sym"]), int(meta["run_first"]), _ptr(sym),
                                           n, _ptr(ws), self._stream()), "lc_rle_expand")
            self.launches += 1
        q = t.empty(n, dtype=t.int32, device=self.dev)
        if n == 0:
            return q
        esc = t.from_numpy(np.frombuffer(escb, dtype="<i4").copy()).to(self.dev) if ne else None
        ws = self._buf("scan", self.lib.lc_scan_workspace_bytes(n))
        N.check(self.lib.lc_unsymbolize_i32(_ptr(sym), n, meta["diff"], meta["bias"], meta["esc"], _ptr(esc),
                                            _ptr(q), _ptr(ws), self._stream()), "lc_unsymbolize_i32")
        self.launches += 2 if meta["diff"] else 1
        return q

    # -- fused entropy stage: both streams, two host syncs in total --------------------------
    def _pinned(self, name, nbytes):
        b = self._small.get("pin_" + name)
        if b is None or b.numel() < nbytes:
            b = self.torch.empty(max(int(nbytes * 1.25), 1 << 16), dtype=self.torch.uint8, pin_memory=True)
            self._small["pin_" + name] = b
        return b

    def _launch_pack(self, sym, table: HF.Table, bits: int, tag: str, extra: dict, tail_len: int = 0, seg: int = SEG,
                     device_out: bool = False, defer: bool = False):
        """Enqueue the Huffman packing of `sym` (no sync) and allocate the finished 'LCH1' container as one
        uninitialised bytes object: everything but the segment lengths, the payload and the tail is known now.
        With device_out the payload stays on the device (no bytes object; _finish_pack returns DevPieces).  With
        defer the caller places the container inside a larger buffer (_place_pack) before collecting it."""
        t = self.torch
        n = sym.numel()
        nseg = (n + seg - 1) // seg
        cap = ((bits + 31) // 32) * 4 + 16
        out = t.zeros(cap, dtype=t.uint8, device=self.dev)
        segb = t.empty(max(nseg, 1) + 1, dtype=t.int64, device=self.dev)
        tab = np.concatenate((table.code.view(np.uint8), table.len))            # 1024 + 256 bytes, one H2D
        d_tab = t.from_numpy(tab.copy()).to(self.dev, non_blocking=True)
        ws = self._buf("scan_" + tag, self.lib.lc_huff_workspace_bytes(n, seg))
        N.check(self.lib.lc_huff_encode(_ptr(sym), n, C.c_void_p(d_tab.data_ptr()),
                                        C.c_void_p(d_tab.data_ptr() + 1024), seg, _ptr(out), _ptr(segb),
                                        C.c_void_p(segb.data_ptr() + 8 * max(nseg, 1)), _ptr(ws), self._stream()),
                "lc_huff_encode")
        self.launches += 1
        nbytes = (bits + 7) // 8
        h_seg = self._pinned("seg_" + tag, 8 * (nseg + 1))
        h_seg[:8 * (nseg + 1)].copy_(segb[:nseg + 1].view(t.uint8), non_blocking=True)
        meta = dict(extra, n=int(n), seg=seg, bits=int(bits), ref_tree=bool(table.reference_tree))
        mj = json.dumps(meta).encode()
        tb = table.to_bytes()
        head = b"".join([HUFF_MAGIC, struct.pack("<III", len(mj), len(tb), nseg), mj, tb])
        pay_off = len(head) + 2 * nseg
        if device_out:
            return dict(nseg=nseg, bits=bits, nbytes=nbytes, h_seg=h_seg, head=head, out=out, keep=(segb, d_tab),
                        device_out=True)
        job = dict(nseg=nseg, bits=bits, nbytes=nbytes, h_seg=h_seg, head_len=len(head), pay_off=pay_off, blob=None,
                   view=None, out=out, keep=(segb, d_tab), head=head, size=pay_off + nbytes + tail_len)
        if not defer:
            blob, view = _new_bytes(job["size"])
            self._place_pack(job, view, blob)
        return job

    @staticmethod
    def _place_pack(job, view, blob=None):
        """Give a deferred container its place: `view` is the uint8 window (job["size"] bytes) it is written to."""
        job["view"], job["blob"] = view, blob
        view[:job["head_len"]] = np.frombuffer(job.pop("head"), dtype=np.uint8)

    def _collect_pack(self, job):
        """Copy the packed payload from the device straight into its place in the container (blocking)."""
        n = job["nbytes"]
        if n and not job.get("device_out"):
            dst = self.torch.from_numpy(job["view"][job["pay_off"]:job["pay_off"] + n])
            dst.copy_(job["out"][:n])

    def _finish_pack(self, job, tail: bytes = b"") -> bytes:
        """After the stream sync: fill in the segment lengths and the tail; returns the container."""
        nseg, bits = job["nseg"], job["bits"]
        offs = np.frombuffer(job["h_seg"].numpy(), dtype=np.int64, count=nseg + 1)
        assert int(offs[nseg]) == bits
        if job.get("device_out"):
            lens = np.diff(offs[:nseg + 1]).astype("<u2").tobytes() if nseg else b""
            return DevPieces([job["head"] + lens, job["out"][:job["nbytes"]], tail])
        if nseg:
            lens = np.diff(offs[:nseg + 1]).astype("<u2")                                           # <= 1024*32 bits
            job["view"][job["head_len"]:job["pay_off"]] = lens.view(np.uint8)
        if tail:
            job["view"][job["pay_off"] + job["nbytes"]:] = np.frombuffer(tail, dtype=np.uint8)
        blob = job["blob"]
        job.clear()
        return blob

    def pack_streams(self, mask, q, device_out: bool = False, frame=None):
        """Entropy-code the mask and the code stream together -> (comp_error, comp_mask).
        Same containers as pack_codes()/pack_mask(); the two pipelines share their host syncs.
        With device_out the results are DevPieces whose Huffman payloads stay in device memory.
        frame = (bytes before, bytes after): the two containers are written as consecutive pickle byte strings
        (BINBYTES8) between them, straight into ONE bytes object -- the finished blob of compress() -- so the payloads
        cross PCIe into their final place and nothing is copied afterwards; the return value is then that blob."""
        t = self.torch
        n_m, n_q = mask.numel(), q.numel()
        if n_q == 0 or n_m == 0:
            ce, cm = self.pack_codes(q), self.pack_mask(mask)
            return (DevPieces([ce]), DevPieces([cm])) if device_out else (ce, cm)
        # ---- pass 1 (async): symbols of the mask + their histogram in one kernel, both code histograms in another
        n_sym = (n_m + 4) // 5
        sym_m = t.empty(n_sym, dtype=t.uint8, device=self.dev)
        # histograms: [0:256] mask bytes, [256:768] the two code candidates
        h3 = t.empty(768, dtype=t.int32, device=self.dev)
        N.check(self.lib.lc_pack_trits_hist(_ptr(mask), n_m, _ptr(sym_m), _ptr(h3), self._stream()), "lc_pack_trits_hist")
        cand = ((0, 0), (1, 127))
        N.check(self.lib.lc_hist_codes(_ptr(q), n_q, ESC, cand[0][1], cand[1][1],
                                       C.c_void_p(h3.data_ptr() + 1024), self._stream()), "lc_hist_codes")
        self.launches += 2
        hh = h3.cpu().numpy().astype(np.int64) & 0xFFFFFFFF                        # sync 1
        hist_m, hists_q = hh[:256], (hh[256:512], hh[512:768])
        # ---- host: tables.  Sparse masks (large abs_error: most 5-trit bytes are zero) are tried as zero-run
        # tokens too and go out that way when it is shorter; a mask with few zero bytes cannot gain from run tokens,
        # so the tokeniser (the slowest kernel of the stage) and its host sync only run when zero bytes are frequent
        tab_m = HF.build_table_native(hist_m)
        mask_extra = {"kind": "mask5", "n_mask": int(n_m)}
        if hist_m[MASK_RUN[0]] > RLE_GATE * n_sym and not hist_m[MASK_RUN[1]:MASK_RUN[1] + 5].any():
            tok_m = t.empty(n_sym, dtype=t.uint8, device=self.dev)
            ht = t.empty(258, dtype=t.int32, device=self.dev)
            ws_r = self._buf("scan_rle", self.lib.lc_rle_workspace_bytes(n_sym))
            N.check(self.lib.lc_rle_tokens(_ptr(sym_m), n_sym, *MASK_RUN, _ptr(tok_m), C.c_void_p(ht.data_ptr() + 1024),
                                           _ptr(ht), None, _ptr(ws_r), self._stream()), "lc_rle_tokens")
            self.launches += 1
            hv = ht.cpu().numpy()                                                  # sync (sparse regime only)
            n_tok = int(hv[256:258].view(np.int64)[0])
            hist_t = hv[:256].astype(np.int64) & 0xFFFFFFFF
            tab_t = HF.build_table_native(hist_t)
            if tab_t.encoded_bits(hist_t) + 16 * (n_tok // SEG) < tab_m.encoded_bits(hist_m) + 16 * (n_sym // SEG):
                sym_m, tab_m, hist_m = tok_m[:n_tok], tab_t, hist_t
                mask_extra.update(rle=1, n_sym=int(n_sym), run_sym=MASK_RUN[0], run_first=MASK_RUN[1])
        tabs_q = [HF.build_table_native(h) for h in hists_q]
        costs = [tb.encoded_bits(h) + 32 * int(h[ESC]) for tb, h in zip(tabs_q, hists_q)]
        pick = 0 if costs[0] <= costs[1] else 1
        diff, bias = cand[pick]
        ne = int(hists_q[pick][ESC])
        # ---- pass 2 (async): symbolise the codes, pack both streams, copy out
        sym_q = t.empty(n_q, dtype=t.uint8, device=self.dev)
        esc = None
        if ne == 0:           # the histogram shows no escapes: element-wise map at memory speed
            N.check(self.lib.lc_symbolize_plain_i32(_ptr(q), n_q, diff, bias, _ptr(sym_q), self._stream()),
                    "lc_symbolize_plain_i32")
        else:
            esc = t.empty(ne, dtype=t.int32, device=self.dev)
            nesc = t.zeros(1, dtype=t.int64, device=self.dev)
            ws = self._buf("scan", self.lib.lc_scan_workspace_bytes(n_q))
            N.check(self.lib.lc_symbolize_i32(_ptr(q), n_q, diff, bias, ESC, _ptr(sym_q), _ptr(esc), _ptr(nesc),
                                              _ptr(ws), self._stream()), "lc_symbolize_i32")
        self.launches += 1
        framed = frame is not None and not device_out
        job_m = self._launch_pack(sym_m, tab_m, tab_m.encoded_bits(hist_m), "m", mask_extra, device_out=device_out,
                                  defer=framed)
        tab_q, hist_q = tabs_q[pick], hists_q[pick]
        code_extra = {"kind": "codes", "diff": diff, "bias": bias, "esc": ESC, "n_esc": ne}
        # Large abs_error: nearly every code is 1 and the one-bit floor of the Huffman code dominates the slot.
        # Fold runs of 1 into run tokens (one more host sync, taken only when three codes in four are 1: below that
        # the tokens gain ~2 % of the slot at best) when that is shorter.
        if pick == 0 and hist_q[CODE_RUN[0]] > 0.75 * n_q and not hist_q[CODE_RUN[1]:CODE_RUN[1] + 5].any():
            tok_q = t.empty(n_q, dtype=t.uint8, device=self.dev)
            ht = t.empty(258, dtype=t.int32, device=self.dev)
            ws_q = self._buf("scan_rle", self.lib.lc_rle_workspace_bytes(n_q))
            N.check(self.lib.lc_rle_tokens(_ptr(sym_q), n_q, *CODE_RUN, _ptr(tok_q), C.c_void_p(ht.data_ptr() + 1024),
                                           _ptr(ht), None, _ptr(ws_q), self._stream()), "lc_rle_tokens")
            self.launches += 1
            hv = ht.cpu().numpy()                                                 # sync (sparse regime only)
            n_tok = int(hv[256:258].view(np.int64)[0])
            hist_t = hv[:256].astype(np.int64) & 0xFFFFFFFF
            tab_t = HF.build_table_native(hist_t)
            if tab_t.encoded_bits(hist_t) + 16 * (n_tok // SEG) < tab_q.encoded_bits(hist_q) + 16 * (n_q // SEG):
                sym_q, tab_q, hist_q = tok_q[:n_tok], tab_t, hist_t
                code_extra.update(rle=1, n_sym=int(n_q), run_sym=CODE_RUN[0], run_first=CODE_RUN[1])
        job_q = self._launch_pack(sym_q, tab_q, tab_q.encoded_bits(hist_q), "q", code_extra, tail_len=4 * ne,
                                  device_out=device_out, defer=framed)
        final = None
        if framed:
            before, after = frame
            hq = b"\x8e" + struct.pack("<Q", job_q["size"])            # pickle BINBYTES8: slot 6, then slot 7
            hm = b"\x8e" + struct.pack("<Q", job_m["size"])
            final, fv = _new_bytes(len(before) + len(hq) + job_q["size"] + len(hm) + job_m["size"] + len(after))
            o = 0
            for piece in (before, hq):
                fv[o:o + len(piece)] = np.frombuffer(piece, dtype=np.uint8)
                o += len(piece)
            self._place_pack(job_q, fv[o:o + job_q["size"]])
            o += job_q["size"]
            fv[o:o + len(hm)] = np.frombuffer(hm, dtype=np.uint8)
            o += len(hm)
            self._place_pack(job_m, fv[o:o + job_m["size"]])
            o += job_m["size"]
            fv[o:] = np.frombuffer(after, dtype=np.uint8)
        h_esc = None
        if ne:
            h_esc = self._pinned("esc", 4 * ne)
            h_esc[:4 * ne].copy_(esc[:ne].view(t.uint8), non_blocking=True)
        # the payloads land directly inside the returned bytes objects: no staging copy, no join
        self._collect_pack(job_m)
        self._collect_pack(job_q)
        self.torch.cuda.current_stream(self.dev).synchronize()                    # sync 2
        escb = bytes(memoryview(h_esc.numpy())[:4 * ne]) if ne else b""
        sizes = (job_q["size"], job_m["size"]) if framed else None
        comp_mask = self._finish_pack(job_m)
        comp_error = self._finish_pack(job_q, tail=escb)
        if framed:
            return final, sizes
        return comp_error, comp_mask

    # -- latent slot (compress.py:142-144; fpzip is replaced by a tagged container) -----
    def stage_latent(self, latent):
        """Start the device->pinned-host copy of the latent (no sync); pass the result to pack_latent()
        after any later sync of the stream so that the copy hides behind the decoder."""
        t = self.torch
        nb = latent.numel() * 4
        pin = self._pinned("latent", nb)
        pin[:nb].copy_(latent.reshape(-1).view(t.uint8), non_blocking=True)
        return pin, nb

    def _latent_meta(self, codec, **extra):
        return dict({"precision": self.precision, "model": self.model.arch.name, "filters": self.model.arch.filters,
                     "kernel": self.model.arch.kernel, "codec": codec, "rev": NUMERICS_REV,
                     "whash": self.weights_hash}, **extra)

    def check_latent_meta(self, meta, lat_cols=None):
        """The error bound only holds when the decoder that runs now is the one the compressor ran: same
        arithmetic revision, precision, architecture and weights.  Raises NativeError otherwise."""
        if meta.get("rev", 1) != NUMERICS_REV:
            raise N.NativeError(f"blob was written by kernel arithmetic revision {meta.get('rev', 1)}, this build is "
                                f"revision {NUMERICS_REV}: its reconstruction would not be reproduced bit for bit")
        if meta.get("precision") != self.precision:
            raise N.NativeError(f"blob was written with decoder precision {meta.get('precision')!r}, engine is "
                                f"{self.precision!r}: the error bound only holds with identical decoder numerics")
        a = self.model.arch
        if (meta.get("model"), meta.get("filters"), meta.get("kernel")) != (a.name, a.filters, a.kernel):
            raise N.NativeError(f"blob was written with model {meta.get('model')!r} (filters {meta.get('filters')}, kernel "
                                f"{meta.get('kernel')}), engine holds {a.name!r} (filters {a.filters}, kernel {a.kernel})")
        if meta.get("whash", self.weights_hash) != self.weights_hash:
            raise N.NativeError(f"blob was written with other weights of model {a.name!r} (hash {meta['whash']}, engine "
                                f"{self.weights_hash}): its reconstruction would not be reproduced")
        if lat_cols is not None and lat_cols != self.latent_floats:
            raise N.NativeError(f"latent has {lat_cols} values per chunk, the engine's decoder expects {self.latent_floats}")

    def pack_latent(self, latent, codec="raw", staged=None) -> bytes:
        mj = json.dumps(self._latent_meta(codec)).encode()
        if staged is not None:                      # float32 little-endian, already on the host
            raw = memoryview(staged[0].numpy())[:staged[1]]
        else:
            raw = latent.cpu().numpy().astype("<f4").tobytes()
        if codec == "bz2":
            raw = bz2.compress(raw, 9)
        return b"".join([LATENT_MAGIC, struct.pack("<I", len(mj)), mj, raw])

    # -- latent quantisation: half-precision codes, high byte plane Huffman-coded ('f16' codec) ---------------
    def quantize_latent(self, latent):
        """Round the latent to half precision IN PLACE (the decoder must consume the rounded values) and return
        (low byte plane, high byte plane, event recorded after the kernel)."""
        t = self.torch
        n = latent.numel()
        lo = t.empty(n, dtype=t.uint8, device=self.dev)
        hi = t.empty(n, dtype=t.uint8, device=self.dev)
        N.check(self.lib.lc_latent_quantize(_ptr(latent), n, _ptr(lo), _ptr(hi), self._stream()), "lc_latent_quantize")
        self.launches += 1
        ev = t.cuda.Event()
        ev.record(t.cuda.current_stream(self.dev))
        return lo, hi, ev

    def _f16_container(self, n, lo_bytes, hi_container) -> bytes:
        mj = json.dumps(self._latent_meta("f16", n=int(n))).encode()
        return b"".join([LATENT_MAGIC, struct.pack("<I", len(mj)), mj, lo_bytes, hi_container])

    def pack_latent_f16(self, lo, hi, ev, device_out: bool = False):
        """Entropy-code the planes of quantize_latent() on a side stream.  Call it after the decoder has been
        enqueued: only the side stream is synchronised, so the two host syncs of the Huffman stage hide behind
        the decoder's kernels."""
        t = self.torch
        side = self.side_stream()
        side.wait_event(ev)
        lo.record_stream(side)
        hi.record_stream(side)
        n = hi.numel()
        with t.cuda.stream(side):
            h = t.empty(256, dtype=t.int32, device=self.dev)
            N.check(self.lib.lc_hist_u8(_ptr(hi), n, _ptr(h), self._stream()), "lc_hist_u8")
            self.launches += 1
            if not device_out:
                h_lo = self._pinned("lat_lo", n)
                h_lo[:n].copy_(lo, non_blocking=True)
            hist = h.cpu().numpy().astype(np.int64) & 0xFFFFFFFF                  # syncs the side stream only
            table = HF.build_table_native(hist)
            job = self._launch_pack(hi, table, table.encoded_bits(hist), "lat", {"kind": "latent_hi"}, seg=LATENT_SEG,
                                    device_out=device_out)
            self._collect_pack(job)
            side.synchronize()
            hi_container = self._finish_pack(job)
        if device_out:          # header | low plane (device) | 'LCH1' of the high plane (payload on the device)
            mj = json.dumps(self._latent_meta("f16", n=int(n))).encode()
            return DevPieces([b"".join([LATENT_MAGIC, struct.pack("<I", len(mj)), mj]), lo] + hi_container.pieces)
        return self._f16_container(n, memoryview(h_lo.numpy())[:n], hi_container)

    def _split_f16(self, blob, meta, lm):
        n = int(meta["n"])
        off = 8 + lm
        return n, blob[off:off + n], blob[off + n:]

    def unpack_latent(self, blob: bytes, shape, dev=None):
        """Slot 0 -> (latent (N, latent_floats) float32 on the device, header dict).  A multi-part container
        ('LCM1', written by the time-sharded compressor) is the concatenation of its parts' chunks.  With
        dev = (low plane, Huffman payload of the high plane) as uint8 CUDA tensors, `blob` is the 'f16' container
        without those two parts (multi-GPU decode)."""
        from .dist import unpack_multipart
        t = self.torch
        parts = unpack_multipart(blob)
        if len(parts) > 1:
            outs, meta = [], None
            for p in parts:
                lat, meta = self.unpack_latent(p, None)
                outs.append(lat)
            lat = t.cat(outs)
            if shape is not None and lat.shape[0] != int(shape[0]):
                raise N.NativeError("latent parts do not add up to the chunk count of blob slot 1")
            return lat, meta
        if blob[:4] != LATENT_MAGIC:
            raise N.NativeError("blob slot 0 is not an 'LCL1' latent container (fpzip streams written by the "
                                "TensorFlow reference cannot be decoded: fpzip is not available and the decoder "
                                "numerics must match the compressor's, SURVEY.md H2)")
        (lm,) = struct.unpack("<I", blob[4:8])
        meta = json.loads(bytes(blob[8:8 + lm]))
        self.check_latent_meta(meta)
        F = self.latent_floats
        if meta.get("codec") == "f16":
            n = int(meta["n"])
            if n % F or (shape is not None and n != int(np.prod(shape))):
                raise N.NativeError("latent container does not match the chunk count / the engine's latent size")
            if dev is not None:
                lo, hi = dev[0], self.unpack_stream(blob[8 + lm:], dev[1])[0]
            else:
                _, lo_b, hic = self._split_f16(blob, meta, lm)
                lo = t.from_numpy(np.frombuffer(lo_b, dtype=np.uint8).copy()).to(self.dev, non_blocking=True)
                hi = self.unpack_stream(hic)[0]
            if lo.numel() != n or hi.numel() != n:
                raise N.NativeError("corrupt latent container: byte planes do not match the value count")
            lat = t.empty((n // F, F), dtype=t.float32, device=self.dev)
            N.check(self.lib.lc_latent_dequantize(_ptr(lo), _ptr(hi), n, _ptr(lat), self._stream()), "lc_latent_dequantize")
            self.launches += 1
            return lat, meta
        raw = blob[8 + lm:]
        if meta.get("codec") == "bz2":
            raw = bz2.decompress(raw)
        n = len(raw) // 4 if shape is None else int(np.prod(shape))
        if n % F or len(raw) < 4 * n:
            raise N.NativeError("latent container does not match the chunk count / the engine's latent size")
        lat = np.frombuffer(raw, dtype="<f4", count=n).reshape(n // F, F)
        return self.torch.from_numpy(lat.copy()).to(self.dev), meta
