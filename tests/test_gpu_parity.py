"""GPU parity tests: the CUDA path (through the C-ABI / the public API) against the CPU oracle
and the goldens made by the reference's own code.  Integer / byte / index stages: bit-exact.
Autoencoder (fp32 CUDA-core mode): |diff| <= 2e-4 on standardised outputs (fp32 vs fp32 with a
different summation order; typical values are O(1))."""
import bz2
import hashlib
import pickle

import numpy as np
import pytest

from conftest import GOLDEN_CASES, golden_case

pytestmark = pytest.mark.gpu

AE_FP32_TOL = 2e-4


@pytest.fixture(scope="module")
def eng():
    from lossycomp._engine import Engine
    return Engine.get(None, "fp32", 0)


@pytest.fixture(scope="module")
def torch():
    import torch
    return torch


def dev(torch, a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("shape", [(16, 48, 48, 2), (17, 49, 50, 2), (33, 100, 97, 1), (40, 96, 144, 3)])
def test_chunk_gather_merge_match_oracle(eng, torch, shape):
    from lossycomp import dataLoader as DL
    from oracle import lossycomp_oracle as O
    a = np.arange(np.prod(shape), dtype=np.float32).reshape(shape)
    c = DL.chunk_data(a, (16, 48, 48))
    np.testing.assert_array_equal(c, O.chunk_data(a))
    np.testing.assert_array_equal(DL.merge_data(c, shape), a)          # Chunking_data.ipynb cells 15, 19


def test_chunk_rejects_small(eng):
    from lossycomp import dataLoader as DL
    with pytest.raises(AssertionError):
        DL.chunk_data(np.zeros((15, 48, 48, 2), np.float32), (16, 48, 48))


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_stats(eng, torch, name):
    g = golden_case(name)
    mean, std, mx = eng.stats(dev(torch, g["x"]))
    assert mean.dtype == np.float32 and std.dtype == np.float32
    assert abs(float(mean) - float(g["mean"])) <= 2 * np.spacing(np.float32(g["mean"]))
    assert abs(float(std) - float(g["std"])) <= 4 * np.spacing(np.float32(g["std"]))
    assert mx == float(np.abs(g["x"][..., 0]).max())


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_autoencoder_fp32_matches_oracle(eng, torch, name):
    g = golden_case(name)
    x = dev(torch, g["x"])
    T, H, W, _ = g["x"].shape
    latent = eng.encode(x, g["mean"], g["std"])
    ref_lat = g["latent"].reshape(latent.shape[0], -1)
    assert np.abs(latent.cpu().numpy() - ref_lat).max() <= AE_FP32_TOL
    # decoder on the oracle's latent, so that only the decoder is compared
    recon = eng.decode(dev(torch, ref_lat), T, H, W)
    assert np.abs(recon.cpu().numpy() - g["recon_std"][..., 0]).max() <= AE_FP32_TOL


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_quantise_codes_bit_exact_on_reference_reconstruction(eng, torch, name):
    """compress.py:114-166 fed the reference's own reconstruction: mask, q, both first differences
    and the bz2 streams must equal what the reference produced."""
    g = golden_case(name)
    x = dev(torch, g["x"])
    recon = dev(torch, g["recon_std"][..., 0])
    thr = float(g["abs_err"])
    mask, q = eng.quantize(x, recon, g["mean"], g["std"], thr)
    ref_q = np.cumsum(g["dq"].astype(np.int64)).astype(np.int32)
    ref_m = np.cumsum(g["dm"].astype(np.int64)).astype(np.int8)
    np.testing.assert_array_equal(q.cpu().numpy(), ref_q)
    np.testing.assert_array_equal(mask.cpu().numpy(), ref_m)
    dq = eng.diff(q).cpu().numpy()
    dm = eng.diff(mask).cpu().numpy()
    np.testing.assert_array_equal(dq, g["dq"])
    np.testing.assert_array_equal(dm, g["dm"])
    assert hashlib.sha256(bz2.compress(dq.tobytes(), 9)).hexdigest() == str(g["comp_error_sha"])
    assert hashlib.sha256(bz2.compress(dm.tobytes(), 9)).hexdigest() == str(g["comp_mask_sha"])
    # cumsum is the exact inverse (decompress.py:78,83)
    np.testing.assert_array_equal(eng.cumsum(dev(torch, g["dq"])).cpu().numpy(), ref_q)
    np.testing.assert_array_equal(eng.cumsum(dev(torch, g["dm"])).cpu().numpy(), ref_m)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_reconstruct_bit_exact_and_violation_set(eng, torch, name):
    """decompress.py:86-104 on the reference's reconstruction and codes: output bit-identical to
    the reference's decompress(), including the elements where the reference exceeds the bound."""
    g = golden_case(name)
    recon = dev(torch, g["recon_std"][..., 0])
    q = dev(torch, np.cumsum(g["dq"].astype(np.int64)).astype(np.int32))
    m = dev(torch, np.cumsum(g["dm"].astype(np.int64)).astype(np.int8))
    xhat = eng.reconstruct(recon, m, q, g["mean"], g["std"], float(g["thr2"]))
    np.testing.assert_array_equal(xhat.cpu().numpy(), g["xhat"][..., 0])
    viol, mx = eng.verify(dev(torch, g["x"]), xhat, float(g["abs_err"]))
    assert viol == int(g["violations"])
    ref_max = np.abs(g["x"][..., 0].astype(np.float64) - g["xhat"][..., 0].astype(np.float64)).max()
    assert mx == ref_max


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_parity_mode_blob_slots_match_reference(eng, torch, name):
    """Whole compress_device in parity mode with the reference's mean/std/reconstruction injected:
    slots 1-8 equal the reference blob's (slot 0 is our latent container by design)."""
    from lossycomp.compress import compress_device
    from lossycomp.decompress import decompress_device
    g = golden_case(name)
    x = dev(torch, g["x"])
    inj = {"mean": g["mean"], "std": g["std"], "recon_std": dev(torch, g["recon_std"][..., 0])}
    slots, info = compress_device(eng, x, float(g["abs_err"]), mode="parity", codec="bz2", inject=inj)
    assert tuple(slots[1]) == tuple(g["latent_shape"]) and tuple(slots[2]) == tuple(g["array_shape"])
    assert tuple(slots[3]) == tuple(g["error_shape"])
    assert slots[4] == g["mean"] and slots[5] == g["std"] and slots[8] == float(g["thr2"])
    assert [type(s).__name__ for s in slots] == list(g["slot_types"])
    assert hashlib.sha256(slots[6]).hexdigest() == str(g["comp_error_sha"])
    assert hashlib.sha256(slots[7]).hexdigest() == str(g["comp_mask_sha"])
    out = decompress_device(eng, pickle.loads(pickle.dumps(slots)), inject=inj)
    np.testing.assert_array_equal(out.cpu().numpy(), g["xhat"][..., 0])


@pytest.mark.parametrize("codec", ["huff", "bz2"])
@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_public_api_roundtrip_strict_bound(name, codec):
    """compress() -> decompress() through the drop-in API: hard bound on every element."""
    from lossycomp.compress import compress
    from lossycomp.decompress import decompress
    g = golden_case(name)
    x, thr = g["x"], float(g["abs_err"])
    blob = compress(x, thr, codec=codec, precision="fp32")
    assert isinstance(blob, bytes)
    slots = pickle.loads(blob)
    assert len(slots) == 9 and slots[8] < 2 * thr
    xhat = decompress(blob)
    assert xhat.dtype == np.float32 and xhat.shape == x.shape[:3] + (1,)
    err = np.abs(x[..., 0].astype(np.float64) - xhat[..., 0].astype(np.float64))
    assert err.max() <= thr, f"{(err > thr).sum()} violations"
    # compression factor within 1% of the reference's for the same input (bz2 codec: same coder)
    if codec == "bz2":
        ref_len = int(g["comp_error_len"]) + int(g["comp_mask_len"])
        got = len(slots[6]) + len(slots[7])
        assert abs(got - ref_len) / ref_len < 0.01


def test_decoder_is_batch_invariant(eng, torch):
    """H2: the decoder output must be bit-identical for any batch split."""
    g = golden_case("odd_all_axes")
    T, H, W, _ = g["x"].shape
    lat = dev(torch, g["latent"].reshape(g["latent"].shape[0], -1))
    a = eng.decode(lat, T, H, W, batch=27).cpu().numpy()
    b = eng.decode(lat, T, H, W, batch=4).cpu().numpy()
    c = eng.decode(lat, T, H, W, batch=1).cpu().numpy()
    assert a.tobytes() == b.tobytes() == c.tobytes()


@pytest.mark.parametrize("n,alpha", [(1, 1), (5, 2), (4096, 3), (4097, 7), (100003, 40), (300000, 200)])
def test_huffman_bits_equal_reference_builder(eng, torch, n, alpha):
    """GPU bit packing from the host-built table == the '0101' string of the reference's
    getHuffmanCode for the same symbol sequence (symbols mapped to single characters)."""
    from lossycomp import huffman as HF
    from oracle import lossycomp_oracle as O
    rng = np.random.default_rng(n)
    p = rng.dirichlet(np.ones(alpha) * 0.3)
    sym = rng.choice(alpha, size=n, p=p).astype(np.uint8)
    table = HF.build_table(np.bincount(sym, minlength=256))
    assert table.reference_tree
    payload, lens, total_bits = eng.huff_pack(dev(torch, sym), table)
    bits, ref_table = O.huffman_code("".join(chr(0x100 + int(s)) for s in sym))
    assert {ord(k) - 0x100: v for k, v in ref_table.items()} == table.codes
    assert total_bits == len(bits)
    got = "".join(format(b, "08b") for b in payload)[:total_bits]
    assert got == bits
    back = eng.huff_unpack(payload, n, lens, table).cpu().numpy()
    np.testing.assert_array_equal(back, sym)


def test_mask_and_code_streams_roundtrip(eng, torch):
    rng = np.random.default_rng(5)
    mask = rng.choice(3, size=123457, p=[0.5, 0.3, 0.2]).astype(np.int8)
    blob = eng.pack_mask(dev(torch, mask))
    np.testing.assert_array_equal(eng.unpack_mask(blob).cpu().numpy(), mask)
    q = np.rint(np.abs(rng.normal(0, 6, size=77777))).astype(np.int32) + 1
    q[::1000] = 100000 + np.arange(q[::1000].size)      # escapes
    blob = eng.pack_codes(dev(torch, q))
    np.testing.assert_array_equal(eng.unpack_codes(blob).cpu().numpy(), q)
    empty = eng.pack_codes(dev(torch, np.zeros(0, np.int32)))
    assert eng.unpack_codes(empty).numel() == 0


def test_no_outliers_is_handled(torch):
    """The reference raises IndexError when nothing exceeds the bound (compress.py:148); we return a
    valid blob with an empty code stream."""
    from lossycomp.compress import compress
    from lossycomp.decompress import decompress
    g = golden_case("one_chunk")
    blob = compress(g["x"], 1000.0, precision="fp32")
    slots = pickle.loads(blob)
    assert slots[3] == (0,)
    xhat = decompress(blob)
    assert np.abs(g["x"][..., 0] - xhat[..., 0]).max() <= 1000.0


def test_api_errors(torch):
    from lossycomp.compress import compress
    with pytest.raises(AssertionError):
        compress(np.zeros((16, 48, 48, 2), np.float32), 0)
    with pytest.raises(AssertionError):
        compress(np.zeros((16, 48, 48, 3), np.float32), 0.1)
    with pytest.raises(AssertionError):
        compress(np.zeros((15, 48, 48, 2), np.float32), 0.1)


def test_time_sharded_compress_equals_single_call(eng, torch):
    """shards made of whole 16-step chunks + shared statistics reproduce the single-call
    mask / codes / reconstruction bit for bit (ranks emulated one after the other on one GPU)."""
    from lossycomp import dist as D
    from lossycomp.compress import compress_device
    from lossycomp.decompress import decompress_device
    g = golden_case("odd_all_axes")
    x = dev(torch, g["x"])
    T = x.shape[0]
    thr = float(g["abs_err"])
    ref_slots, ref_info = compress_device(eng, x, thr, mode="strict", codec="huff", return_parts=True)
    shards = D.shard_time(T, 2)
    assert shards == [(0, 16), (16, 33)]
    raw = [blk for a, b in shards for blk in eng.stats_blocks(x[a:b].contiguous()).cpu().tolist()]
    mean, std, mx = D.combine_stats(raw)
    # per-time-block partials added in time order: bit-identical to the single-GPU statistics, not merely close
    assert mean == ref_slots[4] and std == ref_slots[5]
    assert raw == eng.stats_blocks(x).cpu().tolist()
    rank_slots = []
    for a, b in shards:
        s, _ = compress_device(eng, x[a:b].contiguous(), thr, mode="strict", codec="huff",
                               inject={"mean": mean, "std": std, "max_abs": mx})
        rank_slots.append(pickle.loads(pickle.dumps(s)))
    slots = D.assemble_blob(rank_slots)
    assert slots[1] == ref_slots[1] and slots[2] == ref_slots[2] and slots[3] == ref_slots[3]
    assert slots[8] == ref_slots[8]
    np.testing.assert_array_equal(eng.unpack_mask(slots[7]).cpu().numpy(), ref_info["mask"].cpu().numpy())
    np.testing.assert_array_equal(eng.unpack_codes(slots[6]).cpu().numpy(), ref_info["q"].cpu().numpy())
    whole = decompress_device(eng, slots).cpu().numpy()
    np.testing.assert_array_equal(whole, decompress_device(eng, ref_slots).cpu().numpy())
    # sharded decode: every rank reconstructs only its rows
    for a, b in shards:
        part = decompress_device(eng, slots, rows=(a, b)).cpu().numpy()
        np.testing.assert_array_equal(part, whole[a:b])
    assert np.abs(g["x"][..., 0].astype(np.float64) - whole).max() <= thr


@pytest.mark.parametrize("abs_err", [1e-3, 1e-2, 0.3, 1.0])
def test_abs_error_sweep_strict_bound(abs_err):
    """BASELINE.json configs[3]: abs_err sweep 1e-3 ... 1; the hard bound holds at every setting."""
    from lossycomp.compress import compress
    from lossycomp.decompress import decompress
    g = golden_case("two_by_three")
    x = g["x"]
    blob = compress(x, abs_err)
    xhat = decompress(blob)
    err = np.abs(x[..., 0].astype(np.float64) - xhat[..., 0].astype(np.float64))
    assert err.max() <= abs_err, f"{(err > abs_err).sum()} violations at {abs_err}"
    assert len(blob) < x[..., 0].nbytes


def test_multi_level_field_container():
    """(time, level, lat, lon) variable = independent per-level problems with per-level bounds."""
    from lossycomp.field import compress_field, decompress_field
    from oracle import lossycomp_oracle as O
    levels = [O.synthetic_field(16, 96, 96, seed=7, level=l, h0=250, w0=600) for l in range(3)]
    var = np.stack([f[..., 0] for f in levels], axis=1)          # (T, L, H, W)
    lsm = levels[0][0, :, :, 1]
    errs = [0.05, 0.1, 0.5]
    c = compress_field(var, lsm, errs)
    back = decompress_field(c)
    assert back.shape == var.shape
    for l, e in enumerate(errs):
        assert np.abs(var[:, l].astype(np.float64) - back[:, l]).max() <= e


def test_host_io_paths_agree(torch):
    """Pinned and pageable inputs give the same blob; decompress(out=...) fills the caller's array and
    equals the freshly allocated result (the staged multi-threaded copies lose or reorder nothing)."""
    from lossycomp.compress import compress
    from lossycomp.decompress import decompress
    from oracle import lossycomp_oracle as O
    x = O.synthetic_field(35, 130, 151, seed=11)                 # odd sizes: slices end unaligned
    pinned = torch.empty(x.shape, dtype=torch.float32, pin_memory=True)
    pinned.copy_(torch.from_numpy(x))
    b_page = compress(x, 0.05)
    b_pin = compress(pinned.numpy(), 0.05)
    assert b_page == b_pin
    fresh = decompress(b_page)
    mine = np.full(x.shape[:3] + (1,), np.nan, dtype=np.float32)
    got = decompress(b_page, out=mine)
    assert got.base is mine or got is mine
    assert np.array_equal(fresh, mine)
    assert np.abs(x[..., 0].astype(np.float64) - fresh[..., 0]).max() <= 0.05
    with pytest.raises(AssertionError):
        decompress(b_page, out=np.empty(7, dtype=np.float32))


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_fused_bound_check_equals_reference_violation_set(eng, torch, name):
    """lc_quantize_checked on the reference's reconstruction, in the reference's arithmetic (thr = abs_err):
    same mask / codes as lc_quantize, and its in-pass violation count / maximum error equal those of the
    reference's own decompress() output -- i.e. it reproduces the reference's 1/2-ulp violations exactly."""
    g = golden_case(name)
    x = dev(torch, g["x"])
    recon = dev(torch, g["recon_std"][..., 0])
    thr = float(g["abs_err"])
    m0, q0 = eng.quantize(x, recon, g["mean"], g["std"], thr)
    m1, q1, viol, mx = eng.quantize(x, recon, g["mean"], g["std"], thr, check_abs_err=thr)
    assert torch.equal(m0, m1) and torch.equal(q0, q1)
    assert viol == int(g["violations"])
    ref_max = np.abs(g["x"][..., 0].astype(np.float64) - g["xhat"][..., 0].astype(np.float64)).max()
    assert mx == ref_max
    # and the independent two-kernel path agrees
    xhat = eng.reconstruct(recon, m1, q1, g["mean"], g["std"], thr * 2.0)
    assert eng.verify(x, xhat, thr) == (viol, mx)


def test_verify_modes_give_identical_blobs(eng, torch):
    from lossycomp.compress import compress_device
    from oracle import lossycomp_oracle as O
    x = dev(torch, O.synthetic_field(16, 100, 120, seed=5))
    a, ia = compress_device(eng, x, 0.02, verify="full")
    b, ib = compress_device(eng, x, 0.02)
    c, _ = compress_device(eng, x, 0.02, verify=False)
    assert pickle.dumps(a) == pickle.dumps(b) == pickle.dumps(c)
    assert ia["violations"] == ib["violations"] == 0 and ia["max_err"] == ib["max_err"] <= 0.02


def test_blob_from_other_kernel_revision_is_refused(eng, torch):
    """The decoder must reproduce the compressor's reconstruction bit for bit, so a blob written by kernels
    with different arithmetic (recorded as 'rev' in the latent container) is an error, not a silent bound loss."""
    import json
    import struct
    from lossycomp import _native as N
    from lossycomp.compress import compress_device
    from lossycomp.decompress import decompress_device
    from oracle import lossycomp_oracle as O
    x = dev(torch, O.synthetic_field(16, 48, 64, seed=2))
    slots, _ = compress_device(eng, x, 0.05)
    blob = slots[0]
    (lm,) = struct.unpack("<I", blob[4:8])
    meta = json.loads(blob[8:8 + lm])
    assert "rev" in meta
    meta["rev"] += 1
    mj = json.dumps(meta).encode()
    slots[0] = blob[:4] + struct.pack("<I", len(mj)) + mj + blob[8 + lm:]
    with pytest.raises(N.NativeError, match="revision"):
        decompress_device(eng, slots)


def test_full_size_field_properties(torch):
    """BASELINE.json configs[1] at full size (24 x 720 x 1440, 900 chunks, the shape bench.py times), through
    size-independent properties: every element within the bound, compress() deterministic (identical blobs from
    two calls), decoder idempotent and batch-split invariant, blob self-describing (9 slots, counts consistent),
    and the decoded mask / code streams reproduce the compressor's exactly."""
    from lossycomp import _native as N
    from lossycomp._engine import Engine
    from lossycomp.compress import compress_device
    from lossycomp.decompress import decompress_device
    eng = Engine.get(None, "fp16", 0)
    T, H, W = 24, 720, 1440
    x = torch.empty((T, H, W, 2), dtype=torch.float32, device="cuda")
    N.check(eng.lib.lc_synth_field(x.data_ptr(), T, H, W, 0, 0, 1234, None), "lc_synth_field")
    abs_err = 0.1
    slots, info = compress_device(eng, x, abs_err, return_parts=True)
    slots2, _ = compress_device(eng, x, abs_err)
    assert pickle.dumps(slots) == pickle.dumps(slots2)
    assert len(slots) == 9 and slots[1][0] == 900 and tuple(slots[2]) == (T, H, W, 2)
    assert slots[3][0] == info["q"].numel() == int((info["mask"] > 0).sum().item())
    assert torch.equal(eng.unpack_mask(slots[7]), info["mask"]) and torch.equal(eng.unpack_codes(slots[6]), info["q"])
    xhat = decompress_device(eng, slots)
    viol, mx = eng.verify(x, xhat, abs_err)
    assert viol == 0 and mx <= abs_err
    assert torch.equal(xhat, decompress_device(eng, slots))                 # idempotent
    assert torch.equal(xhat, decompress_device(eng, slots, batch=77))       # ragged batch split of the 900 chunks
    cf = T * H * W * 4 / len(pickle.dumps(slots))
    assert cf > 8.0


def _rle_tokens_reference(b):
    """The oracle's restatement of lc_rle_tokens (mask parameters)."""
    from oracle import lossycomp_oracle as O
    return O.rle_tokens(b, 0, 243)


@pytest.mark.parametrize("n,density", [(1, 0.0), (31, 0.5), (32, 0.0), (33, 0.02), (8192, 0.001), (8193, 0.3),
                                       (100003, 0.01), (100003, 0.9), (250000, 0.0)])
def test_zero_run_tokens_match_reference_and_invert(eng, torch, n, density):
    import ctypes as C
    rng = np.random.default_rng(n)
    b = np.where(rng.random(n) < density, rng.integers(1, 243, n), 0).astype(np.uint8)
    d_b = dev(torch, b)
    tok = torch.empty(n, dtype=torch.uint8, device="cuda")
    meta = torch.zeros(258, dtype=torch.int32, device="cuda")           # [0:256] histogram, [256:258] count
    ws = torch.empty(eng.lib.lc_rle_workspace_bytes(n), dtype=torch.uint8, device="cuda")
    from lossycomp import _native as N
    hin = torch.zeros(256, dtype=torch.int32, device="cuda")
    N.check(eng.lib.lc_rle_tokens(d_b.data_ptr(), n, 0, 243, tok.data_ptr(), C.c_void_p(meta.data_ptr() + 1024),
                                  meta.data_ptr(), hin.data_ptr(), ws.data_ptr(), None), "lc_rle_tokens")
    np.testing.assert_array_equal(hin.cpu().numpy().astype(np.int64), np.bincount(b, minlength=256))
    m = meta.cpu().numpy()
    n_tok = int(m[256:258].view(np.int64)[0])
    ref = _rle_tokens_reference(b)
    got = tok[:n_tok].cpu().numpy()
    np.testing.assert_array_equal(got, ref)
    np.testing.assert_array_equal(m[:256].astype(np.int64), np.bincount(ref, minlength=256))
    back = torch.full((n,), 7, dtype=torch.uint8, device="cuda")
    N.check(eng.lib.lc_rle_expand(tok.data_ptr(), n_tok, 0, 243, back.data_ptr(), n, ws.data_ptr(), None), "lc_rle_expand")
    np.testing.assert_array_equal(back.cpu().numpy(), b)
    # the code-stream parameters: runs of symbol 1, tokens 250..254 (same kernels)
    c = np.where(b == 0, 1, np.where(b == 1, 2, np.minimum(b, 200))).astype(np.uint8)
    N.check(eng.lib.lc_rle_tokens(dev(torch, c).data_ptr(), n, 1, 250, tok.data_ptr(), C.c_void_p(meta.data_ptr() + 1024),
                                  meta.data_ptr(), None, ws.data_ptr(), None), "lc_rle_tokens")
    n_tok2 = int(meta.cpu().numpy()[256:258].view(np.int64)[0])
    assert n_tok2 == n_tok
    N.check(eng.lib.lc_rle_expand(tok.data_ptr(), n_tok2, 1, 250, back.data_ptr(), n, ws.data_ptr(), None), "lc_rle_expand")
    np.testing.assert_array_equal(back.cpu().numpy(), c)


def test_sparse_mask_uses_zero_run_tokens_and_roundtrips(eng, torch):
    """At a large abs_error almost every 5-trit byte is zero: the container must switch to zero-run tokens (a
    plain Huffman stream cannot go below one bit per byte) and decode to the identical mask."""
    import json
    import struct
    rng = np.random.default_rng(5)
    n = 1_000_003
    mask = np.where(rng.random(n) < 0.0005, rng.integers(1, 3, n), 0).astype(np.int8)
    q = np.ones(int((mask > 0).sum()), dtype=np.int32)
    comp_error, comp_mask = eng.pack_streams(dev(torch, mask), dev(torch, q))
    lm = struct.unpack("<I", comp_mask[4:8])[0]
    meta = json.loads(comp_mask[16:16 + lm])
    assert meta.get("rle") == 1
    assert len(comp_mask) < n // 40 // 4            # far below the one-bit-per-byte floor of the plain stream
    assert np.array_equal(eng.unpack_mask(comp_mask).cpu().numpy(), mask)
    assert np.array_equal(eng.unpack_codes(comp_error).cpu().numpy(), q)
    # codes that are almost all 1 (with a few larger ones and an escape) also switch to run tokens
    q2 = np.ones(300_000, dtype=np.int32)
    q2[rng.integers(0, q2.size, 900)] = rng.integers(2, 6, 900)
    q2[12345] = 100000
    m2 = np.zeros(n, dtype=np.int8); m2[:q2.size] = 1
    ce2, cm2 = eng.pack_streams(dev(torch, m2), dev(torch, q2))
    meta2 = json.loads(ce2[16:16 + struct.unpack("<I", ce2[4:8])[0]])
    assert meta2.get("rle") == 1 and len(ce2) < q2.size // 8 // 3
    assert np.array_equal(eng.unpack_codes(ce2).cpu().numpy(), q2)
    assert np.array_equal(eng.unpack_mask(cm2).cpu().numpy(), m2)
    dense = np.where(rng.random(n) < 0.3, rng.integers(1, 3, n), 0).astype(np.int8)
    qd = rng.integers(1, 9, int((dense > 0).sum())).astype(np.int32)
    ce, cm = eng.pack_streams(dev(torch, dense), dev(torch, qd))
    assert "rle" not in json.loads(cm[16:16 + struct.unpack("<I", cm[4:8])[0]])
    assert np.array_equal(eng.unpack_mask(cm).cpu().numpy(), dense)


def test_latent_f16_codec_roundtrip_and_join(torch):
    """Default codec of the 16-bit engines: the latent is rounded to half precision in place (so compressor and
    decompressor decode the same values), stored as a raw low byte plane + a Huffman-coded high byte plane; shard
    containers join into one."""
    import json
    import struct
    from lossycomp._engine import Engine
    from lossycomp.compress import compress_device
    from lossycomp.decompress import decompress_device
    from oracle import lossycomp_oracle as O
    e16 = Engine.get(None, "fp16", 0)
    x = dev(torch, O.synthetic_field(32, 100, 150, seed=3))
    slots, info = compress_device(e16, x, 0.05, return_parts=True)
    blob = slots[0]
    meta = json.loads(blob[8:8 + struct.unpack("<I", blob[4:8])[0]])
    assert meta["codec"] == "f16" and meta["n"] == info["latent"].numel()
    assert len(blob) < 0.45 * 4 * info["latent"].numel()                 # < half of fp32 (2 bytes, high plane coded)
    lat, _ = e16.unpack_latent(blob, slots[1])
    assert torch.equal(lat, info["latent"])                              # exactly the values the compressor decoded
    assert torch.equal(lat, lat.half().float())                          # i.e. half-precision numbers
    xhat = decompress_device(e16, slots)
    assert e16.verify(x, xhat, 0.05)[0] == 0
    # raw codec on request; join of two shard containers == container of the concatenation
    raw_slots, _ = compress_device(e16, x, 0.05, latent_codec="raw")
    assert len(raw_slots[0]) > 4 * info["latent"].numel()
    half = lat.shape[0] // 2
    parts = []
    for part in (lat[:half], lat[half:]):
        p = part.contiguous().clone()
        parts.append(e16.pack_latent_f16(*e16.quantize_latent(p)))
    from lossycomp.dist import pack_multipart
    joined = pack_multipart(parts)
    back, _ = e16.unpack_latent(joined, slots[1])
    assert torch.equal(back, lat)
    # device-resident containers (what the multi-GPU path sends over NCCL) hold the same bytes
    dp = e16.pack_latent_f16(*e16.quantize_latent(lat.clone()), device_out=True)
    assert dp.tobytes() == e16.pack_latent_f16(*e16.quantize_latent(lat.clone()))
    ce, cm = e16.pack_streams(info["mask"], info["q"])
    dce, dcm = e16.pack_streams(info["mask"], info["q"], device_out=True)
    assert dce.tobytes() == ce and dcm.tobytes() == cm
    # and decode from device payloads
    hl = e16.stream_head_len(cm)
    m2 = e16.unpack_mask(cm[:hl], dev(torch, np.frombuffer(cm[hl:], dtype=np.uint8).copy()))
    assert torch.equal(m2, info["mask"])


def test_compress_stream_equals_compress(torch):
    """compress_stream() (H2D of field k+1 overlapped with the kernels of field k, double-buffered) yields exactly
    the blobs of separate compress() calls, for pinned and pageable inputs and changing shapes."""
    from lossycomp.compress import compress, compress_stream
    from oracle import lossycomp_oracle as O
    fields = [O.synthetic_field(16, 64, 80, seed=s) for s in (1, 2, 3)] + [O.synthetic_field(33, 50, 70, seed=4)]
    pinned = []
    for f in fields:
        p = torch.empty(f.shape, dtype=torch.float32, pin_memory=True)
        p.copy_(torch.from_numpy(f))
        pinned.append(p.numpy())
    want = [compress(f, 0.05) for f in fields]
    assert list(compress_stream(pinned, 0.05)) == want
    assert list(compress_stream(fields, 0.05)) == want              # pageable
    assert list(compress_stream([], 0.05)) == []
    assert list(compress_stream(pinned[:1], 0.05)) == want[:1]


@pytest.mark.parametrize("T,H,W,thr,seed", [(24, 336, 720, 0.1, 1234), (50, 241, 241, 0.6, 15)])
def test_benchmark_scale_window_against_oracle(torch, T, H, W, thr, seed):
    """The oracle on a 24x336x720 window of the benchmark field (5.8 M voxels, 2835 scan tiles, 210 chunks with a
    re-anchored time tail; a few seconds of CPU) and on the notebook's shape and bound ((50,241,241,2) at 0.6,
    notebooks/Compression_example.ipynb:128-135; tails on every axis, 3 % outliers; the oracle is pinned at this
    shape by tests/golden/digests.json): with the oracle's statistics and reconstruction injected the CUDA path's
    mask, codes, first differences, bz2 slots, reconstruction and violation set are bit-identical; the fp16
    tensor-core autoencoder stays within its stated tolerance of the oracle's fp32 one; and the Huffman codec
    decodes back to exactly the oracle's mask and codes."""
    from lossycomp._engine import Engine
    from lossycomp.compress import compress_device
    from lossycomp.decompress import decompress_device
    from lossycomp.models import load_model
    from oracle import lossycomp_oracle as O
    if seed == 1234:
        x = O.synthetic_field(T, H, W, seed=seed)
    else:
        x = O.synthetic_field(T, H, W, seed=seed, h0=100 + 7 * seed, w0=200 + 31 * seed)
    blob, p = O.compress(x, thr, load_model(), return_parts=True)
    ref = pickle.loads(blob)
    e32 = Engine.get(None, "fp32", 0)
    xd = dev(torch, x)
    recon = dev(torch, p["recon_std"][..., 0])
    mask, q = e32.quantize(xd, recon, p["mean"], p["std"], thr)
    np.testing.assert_array_equal(mask.cpu().numpy().reshape(p["mask"].shape), p["mask"])
    np.testing.assert_array_equal(q.cpu().numpy(), p["q"])
    np.testing.assert_array_equal(e32.diff(q).cpu().numpy(), p["dq"])
    np.testing.assert_array_equal(e32.diff(mask).cpu().numpy(), p["dm"])
    inj = {"mean": p["mean"], "std": p["std"], "recon_std": recon}
    slots, _ = compress_device(e32, xd, thr, mode="parity", codec="bz2", inject=inj)
    assert slots[6] == ref[6] and slots[7] == ref[7] and slots[3] == ref[3] and slots[8] == ref[8]
    xhat = decompress_device(e32, slots, inject=inj)
    want = O.decompress(blob, load_model(), inject={"recon_std": p["recon_std"]})[..., 0]
    np.testing.assert_array_equal(xhat.cpu().numpy(), want)
    v, _ = e32.verify(xd, xhat, thr)
    assert v == int((np.abs(x[..., 0].astype(np.float64) - want.astype(np.float64)) > thr).sum())
    # Huffman codec: containers decode to the oracle's streams
    hs, _ = compress_device(e32, xd, thr, mode="parity", codec="huff", inject=inj)
    np.testing.assert_array_equal(e32.unpack_mask(hs[7]).cpu().numpy().reshape(p["mask"].shape), p["mask"])
    np.testing.assert_array_equal(e32.unpack_codes(hs[6]).cpu().numpy(), p["q"])
    # statistics: fp64 GPU reduction vs NumPy's fp32 pairwise sums (last-bit differences allowed)
    mean, std, _ = e32.stats(xd)
    assert abs(float(mean) - float(p["mean"])) <= 4 * np.spacing(np.float32(p["mean"]))
    assert abs(float(std) - float(p["std"])) <= 8 * np.spacing(np.float32(p["std"]))
    # autoencoders: fp32 engine within 2e-4, fp16 tensor-core engine within 1e-2 * max(1, |latent|max) (latent) and
    # 2e-2 (reconstruction, standardised units) of the oracle
    lat_ref = p["latent"].reshape(p["latent"].shape[0], -1)
    lat32 = e32.encode(xd, p["mean"], p["std"]).cpu().numpy()
    assert np.abs(lat32 - lat_ref).max() <= AE_FP32_TOL * max(1.0, np.abs(lat_ref).max())
    e16 = Engine.get(None, "fp16", 0)
    lat16 = e16.encode(xd, p["mean"], p["std"]).cpu().numpy()
    assert np.abs(lat16 - lat_ref).max() <= 1e-2 * max(1.0, np.abs(lat_ref).max())
    rec16 = e16.decode(dev(torch, lat_ref), T, H, W).cpu().numpy()
    assert np.abs(rec16 - p["recon_std"][..., 0]).max() <= 2e-2


def test_fp32_engine_against_cudnn(eng, torch, model2):
    """Independent cross-check of the convolution kernels (SURVEY.md 7.3 allows cuDNN in tests only): the oracle's
    layer statements executed by torch on the GPU (cuDNN) against the fp32 CUDA-core engine, encoder and decoder,
    on chunks with tails on every axis."""
    from oracle import lossycomp_oracle as O
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = golden_case("odd_all_axes")
    x = g["x"]
    xs = x.copy()
    xs[..., 0] = (x[..., 0] - g["mean"]) / g["std"]
    chunks = O.chunk_data(xs)
    lat_cudnn = O.run_layers(chunks, model2.encoder, device="cuda")
    lat = eng.encode(dev(torch, x), g["mean"], g["std"]).cpu().numpy()
    scale = max(1.0, float(np.abs(lat_cudnn).max()))
    assert np.abs(lat - lat_cudnn.reshape(lat.shape)).max() <= AE_FP32_TOL * scale
    rec_cudnn = O.merge_data(O.run_layers(lat_cudnn, model2.decoder, device="cuda"), x.shape)[..., 0]
    rec = eng.decode(dev(torch, lat_cudnn.reshape(lat.shape)), *x.shape[:3]).cpu().numpy()
    assert np.abs(rec - rec_cudnn).max() <= AE_FP32_TOL * max(1.0, float(np.abs(rec_cudnn).max()))
    # and cuDNN agrees with the CPU oracle the goldens were made with
    assert np.abs(lat_cudnn - g["latent"]).max() <= AE_FP32_TOL * scale


def test_multi_variable_stack_abs_err_sweep():
    """BASELINE.json configs[3]: a multi-variable stack -- the reference concatenates variables on the level axis
    (dataLoader.py:416-424), i.e. V independent 2-channel problems -- with an abs_err sweep 1e-3 ... 1 (relative to
    each variable's scale): temperature (K), specific humidity (kg/kg, values near zero, where the bound check's
    fp64 branch is taken) and geopotential (m^2/s^2, large magnitudes, where the float32 guard band is widest).
    Every element of every variable meets its bound; the compression factor grows with the bound."""
    from lossycomp.field import compress_field, decompress_field
    from oracle import lossycomp_oracle as O
    base = O.synthetic_field(16, 96, 144, seed=21, h0=200, w0=400)
    t = base[..., 0].astype(np.float64)
    z = (t - t.mean()) / t.std()
    variables = {"temperature": (t, 1.0), "specific_humidity": (np.clip(8e-3 + 4e-3 * z, 0, None), 4e-3),
                 "geopotential": (5.4e4 + 2.5e3 * z, 2.5e3 / 16.0)}
    lsm = base[0, :, :, 1]
    for name, (v, scale) in variables.items():
        var = np.repeat(v[:, None].astype(np.float32), 5, axis=1)              # 5 "levels", one per bound
        errs = [scale * e for e in (1e-3, 1e-2, 0.1, 0.3, 1.0)]
        if name == "geopotential":
            errs[0] = max(errs[0], 0.05)                                       # float32 resolution at 5.4e4 is 0.004
        c = compress_field(var, lsm, errs)
        back = decompress_field(c)
        import pickle
        blobs = pickle.loads(c)["blobs"]
        sizes = []
        for l, e in enumerate(errs):
            d = np.abs(var[:, l].astype(np.float64) - back[:, l].astype(np.float64)).max()
            assert d <= e, (name, e, d)
            sizes.append(len(blobs[l]))
        assert sizes == sorted(sizes, reverse=True), (name, sizes)
        assert sizes[-1] < var[:, 0].nbytes / 8
