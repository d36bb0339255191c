"""The C-ABI shared library loads without a GPU and exports every symbol include/lossycomp_b200.h
declares; the ctypes table mirrors the header; the product path refuses to run without CUDA."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "lossycomp_b200.h")


def declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lc_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from lossycomp import _native
    if not os.path.exists(_native.LIB_PATH):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "lossy-ml_b200", "csrc"), "-j8"])
    return _native.lib()


def test_every_declared_symbol_is_exported_and_bound(lib):
    from lossycomp import _native
    names = declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert set(names) == set(_native.SIGNATURES), set(names) ^ set(_native.SIGNATURES)


def test_host_only_entry_points(lib):
    assert lib.lc_version() >= 100
    g = (ctypes.c_int * 3)()
    assert lib.lc_chunk_grid(32, 721, 1440, ctypes.byref(g)) == 0 and list(g) == [2, 16, 30]   # Chunking_data.ipynb cell 8
    assert lib.lc_chunk_grid(15, 48, 48, ctypes.byref(g)) != 0
    assert b"Chunks size cannot be larger" in lib.lc_last_error()
    assert lib.lc_scan_workspace_bytes(1 << 20) > 0 and lib.lc_stats_workspace_bytes() > 0


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from lossycomp import _native
    from lossycomp.compress import compress
    from lossycomp.decompress import decompress
    with pytest.raises(_native.NativeError):
        compress(np.zeros((16, 48, 48, 2), np.float32), 0.1)
    with pytest.raises(_native.NativeError):
        decompress(b"")


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "lossy-ml_b200", "lossycomp")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            assert "oracle" not in open(os.path.join(pkg, fn)).read().replace("the oracle", ""), fn


def test_model_spec_matches_survey():
    from lossycomp.models import MODEL_1_ARCH, load_model
    m = load_model()
    macs = m.arch.macs_per_chunk()
    assert macs == {"Encoder": 1339836480, "Decoder": 1285572672}          # 2625.4 MMAC
    assert sum(MODEL_1_ARCH.macs_per_chunk().values()) == 2395504800          # SURVEY: 2395.5 MMAC
    assert [L.name for L in m.layers][:3] == ["conv3d", "conv3d_1", "conv3d_2"]
    assert sum(L.kernel.size + L.bias.size for L in m.layers) == 332696
    assert abs(float(m.layers[-1].bias[0]) - 0.10914003) < 1e-7            # Decoder/conv3d_9/bias


def test_chunk_index_map_matches_oracle_property(lib):
    """The kernels' chunk window / merge map (host view: lc_chunk_window) against the oracle's restatement of
    dataLoader.py:437-532 on random shapes: every voxel is kept by exactly one chunk, at the right place."""
    from hypothesis import given, settings, strategies as st
    from oracle import lossycomp_oracle as O

    @settings(max_examples=60, deadline=None)
    @given(st.integers(16, 70), st.integers(48, 150), st.integers(48, 150))
    def check(T, H, W):
        g = (ctypes.c_int * 3)()
        assert lib.lc_chunk_grid(T, H, W, ctypes.byref(g)) == 0
        assert tuple(g) == O.chunk_grid((T, H, W))
        cover = np.zeros((T, H, W), dtype=np.int32)
        out = (ctypes.c_int * 6)()
        for idx in range(g[0] * g[1] * g[2]):
            assert lib.lc_chunk_window(T, H, W, idx, ctypes.byref(out)) == 0
            t0, h0, w0, kt, kh, kw = out
            assert (t0, h0, w0) == O.chunk_origin(idx, (T, H, W))
            cover[t0 + kt:t0 + 16, h0 + kh:h0 + 48, w0 + kw:w0 + 48] += 1
        assert cover.min() == 1 and cover.max() == 1

    check()
    # and the merge really is the inverse of chunk on the kept parts
    a = np.arange(33 * 100 * 97, dtype=np.float32).reshape(33, 100, 97, 1)
    assert np.array_equal(O.merge_data(O.chunk_data(a), a.shape), a)


def test_native_table_builder_equals_reference_mirror():
    """lc_huff_build_table (host C) against huffman.build_table (the Python mirror of the reference's tree,
    itself pinned to the reference's code strings in test_oracle_golden): same codes, lengths and flags,
    including heavy ties, single symbols, empty histograms and trees that need depth capping."""
    from lossycomp import huffman as HF
    rng = np.random.default_rng(3)
    hists = [np.zeros(256, np.int64)]
    one = np.zeros(256, np.int64); one[200] = 9; hists.append(one)
    fib = [1, 1]
    while len(fib) < 70:
        fib.append(fib[-1] + fib[-2])
    deep = np.zeros(256, np.int64); deep[:70] = fib; hists.append(deep)
    for trial in range(400):
        k = int(rng.integers(1, 257))
        h = np.zeros(256, np.int64)
        idx = rng.choice(256, k, replace=False)
        h[idx] = [rng.integers(1, 4, k), rng.integers(1, 1 << 30, k), np.ones(k, np.int64),
                  (2.0 ** rng.uniform(0, 40, k)).astype(np.int64) + 1][trial % 4]
        hists.append(h)
    for h in hists:
        a, b = HF.build_table(h), HF.build_table_native(h)
        assert np.array_equal(a.code, b.code) and np.array_equal(a.len, b.len)
        assert np.array_equal(a.used, b.used) and a.reference_tree == b.reference_tree
        assert a.to_bytes() == b.to_bytes()
        back, _ = HF.Table.from_bytes(b.to_bytes())
        assert back.codes == b.codes
    assert not HF.build_table_native(deep).reference_tree and HF.build_table_native(deep).len.max() <= 32


def test_native_decode_arrays_equal_python_builder():
    """lc_huff_build_decode (host C) against Table.decode_arrays: same direct table and tree for random codes,
    one-symbol alphabets and codes longer than the direct table."""
    from lossycomp import huffman as HF
    rng = np.random.default_rng(11)
    hists = []
    one = np.zeros(256, np.int64); one[17] = 3; hists.append(one)
    fib = [1, 1]
    while len(fib) < 40:
        fib.append(fib[-1] + fib[-2])
    deep = np.zeros(256, np.int64); deep[:40] = fib; hists.append(deep)          # code lengths up to 32 > LUT_BITS
    for trial in range(200):
        k = int(rng.integers(2, 257))
        h = np.zeros(256, np.int64)
        h[rng.choice(256, k, replace=False)] = [rng.integers(1, 5, k), rng.integers(1, 1 << 28, k),
                                               (2.0 ** rng.uniform(0, 30, k)).astype(np.int64) + 1][trial % 3]
        hists.append(h)
    for h in hists:
        t = HF.build_table_native(h)
        a, b = t.decode_arrays(), t.decode_arrays_native()
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
        back, used = HF.Table.from_bytes(memoryview(t.to_bytes()))
        assert used == len(t.to_bytes()) and np.array_equal(back.code, t.code) and np.array_equal(back.len, t.len)


def test_dumps_slots_loads_like_pickle():
    """The hand-assembled blob (one copy of the payloads) unpickles to exactly the list pickle.dumps would give,
    and is a well-formed pickle stream (pickletools is stricter than the unpickler about memo keys)."""
    import pickle
    import pickletools
    from lossycomp.compress import dumps_slots
    rng = np.random.default_rng(0)
    slots = [b"LCL1" + bytes(rng.integers(0, 255, 5000, dtype=np.uint8)), (900, 3, 3, 23), (24, 720, 1440, 2), (12345,),
             np.float32(265.5), np.float32(16.25), bytes(rng.integers(0, 255, 70000, dtype=np.uint8)), b"", 0.19998]
    blob = dumps_slots(slots)
    back = pickle.loads(blob)
    assert type(back) is list and len(back) == 9
    for a, b in zip(slots, back):
        assert type(a) is type(b) and a == b
    assert back == pickle.loads(pickle.dumps(slots))
    pickletools.dis(blob, out=open(os.devnull, "w"))


def test_coordinate_encodings_match_reference_values():
    """lossycomp/encodings.py:4-23 (encode_lon / encode_lat): values produced by the reference's own functions, and
    the channel order of dataLoader.py:284-301 for the 5-channel model's input."""
    from lossycomp.encodings import add_coords, encode_lat, encode_lon
    ref_lon = {-180: (0.0, 1.0), -100.5: (0.9832549075639545, 0.18223552549214767), 0: (1.2246467991473532e-16, -1.0),
               33.25: (-0.5482932295199135, -0.8362861558477597), 180: (-2.4492935982947064e-16, 1.0)}
    ref_lat = {-90: (0.0, 1.0), -30.25: (0.8638355052043957, 0.5037739770455263), 0: (1.0, 6.123233995736766e-17),
               45: (0.7071067811865476, -0.7071067811865475), 90: (1.2246467991473532e-16, -1.0)}
    for v, want in ref_lon.items():
        assert encode_lon(v) == want
    for v, want in ref_lat.items():
        assert encode_lat(v) == want
    with pytest.raises(AssertionError):
        encode_lon(181)
    with pytest.raises(AssertionError):
        encode_lat(-91)
    d = np.arange(2 * 3 * 4, dtype=np.float32).reshape(2, 3, 4)
    x = add_coords(d, [45, 0, -30.25], [-180, -100.5, 0, 33.25])
    assert x.shape == (2, 3, 4, 5) and x.dtype == np.float32 and np.array_equal(x[..., 0], d)
    assert np.allclose(x[1, 2, :, 1], 0.8638355052043957) and np.allclose(x[0, 2, :, 2], 0.5037739770455263)   # sin/cos(lat)
    assert np.allclose(x[0, :, 1, 3], 0.9832549075639545) and np.allclose(x[1, :, 1, 4], 0.18223552549214767)   # sin/cos(lon)


def test_stream_and_pipeline_entry_points_exist():
    """The bulk-job API of this package (no counterpart in the reference): stream generators, the begin / finish
    halves the pipelines are built from, and the lane set -- importable without a GPU, with the keyword surface the
    docs name."""
    import inspect
    from lossycomp import compress as Cm, decompress as Dm, dist as Di, lanes as La
    assert inspect.isgeneratorfunction(Cm.compress_stream) and inspect.isgeneratorfunction(Dm.decompress_stream)
    assert inspect.isgeneratorfunction(Di.compress_sharded_stream)
    assert "lanes" in inspect.signature(Cm.compress_stream).parameters
    assert {"lanes", "outs"} <= set(inspect.signature(Dm.decompress_stream).parameters)
    assert {"t0", "T", "dst", "lanes", "group"} <= set(inspect.signature(Di.compress_sharded_stream).parameters)
    for fn in (Cm.compress_begin, Cm.compress_finish, Di.compress_sharded_begin, Di.compress_sharded_finish):
        assert callable(fn)
    # compress_device(...) == compress_finish(compress_begin(...)): the halves take the whole's keywords between them
    whole = set(inspect.signature(Cm.compress_device).parameters) - {"eng", "x", "err_threshold"}
    halves = (set(inspect.signature(Cm.compress_begin).parameters) | set(inspect.signature(Cm.compress_finish).parameters)) \
        - {"eng", "x", "err_threshold", "st"}
    assert whole == halves, (whole ^ halves)
    assert La.n_lanes(3) == 3 and La.n_lanes(0) == 1
