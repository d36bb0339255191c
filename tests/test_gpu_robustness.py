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
