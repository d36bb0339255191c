"""Failure behaviour of the CUDA path: workspace contract of the C-ABI, blobs decoded with the wrong engine,
damaged containers, repeated calls.  Everything must fail loudly (NativeError / ValueError), never decode garbage."""
import ctypes as C
import pickle

import numpy as np
import pytest

from conftest import golden_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch():
    import torch
    return torch


def _ptr(t):
    return C.c_void_p(t.data_ptr())


def test_workspace_contract_through_the_c_abi(torch):
    """lc_encode / lc_decode with a caller-owned workspace: a 0xFF-poisoned buffer passed through
    lc_ae_workspace_init gives bit-identical results to the engine's own workspace; a buffer that was never
    initialised is refused instead of silently producing other convolutions (include/lossycomp_b200.h)."""
    from lossycomp import _native as N
    from lossycomp._engine import Engine
    g = golden_case("ragged_tails")
    eng = Engine.get(None, "fp16", 0)
    lib = eng.lib
    x = torch.from_numpy(g["x"]).cuda()
    T, H, W, Cc = x.shape
    n = eng.n_chunks(T, H, W)
    want_lat = eng.encode(x, g["mean"], g["std"])
    want_rec = eng.decode(want_lat, T, H, W)
    nbytes = lib.lc_ae_workspace_bytes(eng.h, n)
    ws = torch.full((nbytes,), 0xFF, dtype=torch.uint8, device="cuda")
    lat = torch.empty_like(want_lat)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    args = (eng.h, _ptr(x), T, H, W, Cc, float(g["mean"]), float(g["std"]), 0, n, _ptr(lat), _ptr(ws), nbytes, stream)
    assert lib.lc_encode(*args) != 0                                      # refused: not initialised
    assert b"lc_ae_workspace_init" in lib.lc_last_error()
    N.check(lib.lc_ae_workspace_init(eng.h, _ptr(ws), nbytes, stream), "lc_ae_workspace_init")
    N.check(lib.lc_encode(*args), "lc_encode")
    assert torch.equal(lat, want_lat)
    rec = torch.empty_like(want_rec)
    N.check(lib.lc_decode(eng.h, _ptr(lat), T, H, W, 0, n, _ptr(rec), _ptr(ws), nbytes, stream), "lc_decode")
    assert torch.equal(rec, want_rec)
    # second use of the same workspace (halos must still be zero: the kernels never write them)
    N.check(lib.lc_encode(*args), "lc_encode")
    assert torch.equal(lat, want_lat)
    N.check(lib.lc_ae_workspace_release(eng.h, _ptr(ws)), "release")
    assert lib.lc_encode(*args) != 0


def test_blob_names_its_decoder(torch):
    """A blob decodes only with the architecture AND the weights it was written with (a landsea_model
    blob on the model_2 engine used to read out of bounds and return garbage)."""
    import os
    from lossycomp import _native as N
    from lossycomp._engine import Engine
    from lossycomp.compress import compress, compress_device
    from lossycomp.decompress import decompress, decompress_device
    from lossycomp.models import Model, load_model
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    landsea = os.path.join(root, "lossy-ml_b200", "weights", "landsea_model.lcw")
    g = golden_case("two_by_three")
    blob = compress(g["x"], 0.1, model=landsea)
    with pytest.raises(N.NativeError, match="model"):
        decompress(blob, model=os.path.join(root, "lossy-ml_b200", "weights", "model_2.lcw"))
    xhat = decompress(blob)                                   # picks weights/landsea_model.lcw from the blob header
    assert np.abs(g["x"][..., 0].astype(np.float64) - xhat[..., 0]).max() <= 0.1
    # same architecture, other weights: refused by the weight hash
    m2 = load_model()
    other = Model(m2.arch, [type(L)(**{**L.__dict__, "bias": L.bias + np.float32(1e-3)}) for L in m2.layers])
    e_other = Engine(other, "fp16", 0)
    slots, _ = compress_device(Engine.get(None, "fp16", 0), torch.from_numpy(g["x"]).cuda(), 0.1)
    with pytest.raises(N.NativeError, match="weights"):
        decompress_device(e_other, slots)


def test_damaged_blobs_raise(torch):
    from lossycomp import _native as N
    from lossycomp.compress import compress
    from lossycomp.decompress import decompress
    g = golden_case("two_by_three")
    blob = compress(g["x"], 0.1)
    slots = pickle.loads(blob)

    def broken(i, f):
        s = list(slots)
        s[i] = f(s[i])
        return pickle.dumps(s)

    bad = [
        broken(7, lambda b: b[:len(b) // 2]),                                   # truncated mask stream
        broken(6, lambda b: b[:40]),                                            # truncated code container
        broken(3, lambda v: (v[0] + 5,)),                                       # outlier count disagrees
        broken(0, lambda b: b[:-20]),                                           # truncated latent
        broken(1, lambda v: (v[0] + 1,) + tuple(v[1:])),                        # chunk count disagrees
        broken(7, lambda b: b[:16] + bytes(len(b) - 16)),                       # header / table zeroed
    ]
    for i, bb in enumerate(bad):
        with pytest.raises((N.NativeError, ValueError, AssertionError, KeyError, Exception)) as ei:
            decompress(bb)
        assert not isinstance(ei.value, (SystemExit, KeyboardInterrupt)), i
    # a mask whose outlier count differs from the code count (swap in the mask of another threshold)
    other = pickle.loads(compress(g["x"], 0.3))
    s = list(slots)
    s[7] = other[7]
    with pytest.raises(N.NativeError, match="outliers"):
        decompress(pickle.dumps(s))
    # the GPU is still healthy afterwards
    xhat = decompress(blob)
    assert np.abs(g["x"][..., 0].astype(np.float64) - xhat[..., 0]).max() <= 0.1


def test_prefix_code_validation():
    """lc_huff_build_decode rejects tables that are not complete prefix codes (host only)."""
    from lossycomp import _native as N
    lib = N.lib()

    def build(codes):   # {sym: (code, len)}
        code = np.zeros(256, np.uint32); ln = np.zeros(256, np.uint8); used = np.zeros(256, np.uint8)
        for s_, (c, l) in codes.items():
            code[s_], ln[s_], used[s_] = c, l, 1
        lut = np.zeros(1 << 11, np.uint16); tree = np.zeros(1022, np.int32); n = C.c_int(0)
        return lib.lc_huff_build_decode(code.ctypes.data, ln.ctypes.data, used.ctypes.data, 11, lut.ctypes.data,
                                        tree.ctypes.data, C.byref(n))
    assert build({1: (0, 1), 2: (2, 2), 3: (3, 2)}) == 0
    assert build({1: (0, 1), 2: (2, 2)}) != 0                      # incomplete
    assert build({1: (0, 1), 2: (1, 2), 3: (3, 2)}) != 0           # 01 extends 0
    assert build({1: (4, 2), 2: (1, 1)}) != 0                      # code wider than its length
    assert build({1: (0, 1), 2: (0, 1)}) != 0                      # duplicate
    assert build({7: (0, 0)}) == 0                                 # one-symbol alphabet


def test_soak_repeated_compress_is_bit_stable(torch):
    """200 compress() calls of the same field give memcmp-identical blobs (two MMA issuer warps, look-back scans,
    Huffman packing with atomics on disjoint words: none of it may depend on scheduling), on both engines."""
    import hashlib
    from lossycomp.compress import compress
    g = golden_case("odd_all_axes")
    for prec, reps in (("fp16", 200), ("fp32", 20)):
        ref = hashlib.sha256(compress(g["x"], 0.05, precision=prec)).hexdigest()
        for _ in range(reps):
            assert hashlib.sha256(compress(g["x"], 0.05, precision=prec)).hexdigest() == ref


def test_two_devices_in_one_process(torch):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from lossycomp.compress import compress
    from lossycomp.decompress import decompress
    g = golden_case("one_chunk")
    b0 = compress(g["x"], 0.1, device=0)
    b1 = compress(g["x"], 0.1, device=1)
    assert b0 == b1
    assert np.array_equal(decompress(b0, device=1), decompress(b1, device=0))


def test_no_writes_outside_the_buffers(torch):
    """compute-sanitizer is closed on this GPU pool (profiles/r02_sanitizer_closed.txt), so out-of-bounds WRITES are
    hunted with guard zones instead: every output of the C-ABI path (latent, reconstruction, mask, codes, the
    autoencoder workspace) sits between two 1 MB zones of a known pattern, on ragged shapes that exercise the tile
    tails of every kernel; the zones must be intact afterwards and the results must equal the unguarded run."""
    from lossycomp import _native as N
    from lossycomp._engine import Engine
    from oracle import lossycomp_oracle as O
    G = 1 << 20

    def guarded(nbytes, dtype):
        raw = torch.full((nbytes + 2 * G,), 0xA5, dtype=torch.uint8, device="cuda")
        return raw, raw[G:G + nbytes].view(dtype)

    def intact(raw, nbytes):
        return bool((raw[:G] == 0xA5).all()) and bool((raw[G + nbytes:] == 0xA5).all())

    for prec in ("fp16", "fp32"):
        eng = Engine.get(None, prec, 0)
        lib = eng.lib
        for shape in ((17, 49, 50), (33, 100, 97)):
            x = torch.from_numpy(O.synthetic_field(*shape, seed=9, h0=50, w0=900)).cuda()
            T, H, W, Cc = x.shape
            n = eng.n_chunks(T, H, W)
            mean, std, _ = eng.stats(x)
            want_lat = eng.encode(x, mean, std)
            want_rec = eng.decode(want_lat, T, H, W)
            want_mask, want_q = eng.quantize(x, want_rec, mean, std, 0.1)
            stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            ws_bytes = lib.lc_ae_workspace_bytes(eng.h, n)
            ws_raw, ws = guarded(ws_bytes, torch.uint8)
            N.check(lib.lc_ae_workspace_init(eng.h, _ptr(ws), ws_bytes, stream), "init")
            lat_raw, lat = guarded(want_lat.numel() * 4, torch.float32)
            N.check(lib.lc_encode(eng.h, _ptr(x), T, H, W, Cc, float(mean), float(std), 0, n, _ptr(lat), _ptr(ws), ws_bytes,
                                  stream), "lc_encode")
            rec_raw, rec = guarded(T * H * W * 4, torch.float32)
            N.check(lib.lc_decode(eng.h, _ptr(lat), T, H, W, 0, n, _ptr(rec), _ptr(ws), ws_bytes, stream), "lc_decode")
            m_raw, mask = guarded(T * H * W, torch.int8)
            q_raw, q = guarded(T * H * W * 4, torch.int32)
            cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
            sw = torch.empty(lib.lc_scan_workspace_bytes(T * H * W), dtype=torch.uint8, device="cuda")
            N.check(lib.lc_quantize(_ptr(x), Cc, _ptr(rec), T * H * W, float(mean), float(std), float(np.float32(0.1)),
                                    float(np.float32(0.2)), _ptr(mask), _ptr(q), _ptr(cnt), _ptr(sw), stream), "lc_quantize")
            torch.cuda.synchronize()
            assert intact(ws_raw, ws_bytes) and intact(lat_raw, want_lat.numel() * 4) and intact(rec_raw, T * H * W * 4)
            assert intact(m_raw, T * H * W) and intact(q_raw, T * H * W * 4)
            assert torch.equal(lat.view(want_lat.shape), want_lat) and torch.equal(rec.view(want_rec.shape), want_rec)
            k = int(cnt.item())
            assert k == want_q.numel() and torch.equal(q[:k], want_q) and torch.equal(mask, want_mask)
            N.check(lib.lc_ae_workspace_release(eng.h, _ptr(ws)), "release")
