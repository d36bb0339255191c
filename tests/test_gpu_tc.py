"""Tensor-core (tcgen05) autoencoder path: against the CUDA-core back end on identical 16-bit
activations/weights, against the fp32 oracle within the 16-bit tolerance, and end to end."""
import os
import pickle

import numpy as np
import pytest

from conftest import GOLDEN_CASES, golden_case

pytestmark = pytest.mark.gpu

# 16-bit operands, fp32 accumulation, activations rounded to 16 bit after every layer:
# standardised outputs are O(1); bf16 has 8 significant bits, fp16 11.
TOL_VS_ORACLE = {"bf16": 6e-2, "fp16": 1e-2}
# same function, different fp32 summation order -> occasional 1-ulp flips of a 16-bit activation
TOL_TC_VS_DIRECT = {"bf16": 2e-2, "fp16": 3e-3}


def dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def make_engine(prec, no_tc=False):
    from lossycomp._engine import Engine
    from lossycomp.models import load_model
    old = os.environ.pop("LOSSYCOMP_NO_TC", None)
    if no_tc:
        os.environ["LOSSYCOMP_NO_TC"] = "1"
    try:
        return Engine(load_model(), prec, 0)
    finally:
        os.environ.pop("LOSSYCOMP_NO_TC", None)
        if old is not None:
            os.environ["LOSSYCOMP_NO_TC"] = old


@pytest.fixture(scope="module", params=["bf16", "fp16"])
def engines(request):
    prec = request.param
    return prec, make_engine(prec), make_engine(prec, no_tc=True)


@pytest.mark.parametrize("name", ["one_chunk", "two_by_three", "odd_all_axes"])
def test_tc_matches_cuda_core_backend(engines, name):
    prec, tc, direct = engines
    g = golden_case(name)
    x = dev(g["x"])
    T, H, W, _ = g["x"].shape
    la = tc.encode(x, g["mean"], g["std"]).cpu().numpy()
    lb = direct.encode(x, g["mean"], g["std"]).cpu().numpy()
    # one 16-bit ulp of the largest latent value (latents reach ~15)
    assert np.abs(la - lb).max() <= TOL_TC_VS_DIRECT[prec] * max(1.0, np.abs(lb).max() / 4), np.abs(la - lb).max()
    lat = dev(g["latent"].reshape(la.shape))
    ra = tc.decode(lat, T, H, W).cpu().numpy()
    rb = direct.decode(lat, T, H, W).cpu().numpy()
    d = np.abs(ra - rb)
    assert d.max() <= TOL_TC_VS_DIRECT[prec], d.max()
    assert d.mean() <= TOL_TC_VS_DIRECT[prec] / 20


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_tc_autoencoder_within_tolerance_of_oracle(engines, name):
    prec, tc, _ = engines
    g = golden_case(name)
    x = dev(g["x"])
    T, H, W, _ = g["x"].shape
    lat = tc.encode(x, g["mean"], g["std"]).cpu().numpy()
    ref = g["latent"].reshape(lat.shape)
    assert np.abs(lat - ref).max() <= TOL_VS_ORACLE[prec] * max(1.0, np.abs(ref).max())
    recon = tc.decode(dev(ref), T, H, W).cpu().numpy()
    assert np.abs(recon - g["recon_std"][..., 0]).max() <= TOL_VS_ORACLE[prec]


def test_tc_decoder_is_batch_invariant_and_repeatable(engines):
    prec, tc, _ = engines
    g = golden_case("odd_all_axes")
    T, H, W, _ = g["x"].shape
    lat = dev(g["latent"].reshape(g["latent"].shape[0], -1))
    a = tc.decode(lat, T, H, W, batch=27).cpu().numpy()
    b = tc.decode(lat, T, H, W, batch=5).cpu().numpy()
    c = tc.decode(lat, T, H, W, batch=1).cpu().numpy()
    d = tc.decode(lat, T, H, W, batch=27).cpu().numpy()
    assert a.tobytes() == b.tobytes() == c.tobytes() == d.tobytes()


@pytest.mark.parametrize("prec", ["bf16", "fp16"])
@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_tc_roundtrip_bound_and_cf(prec, name):
    from lossycomp.compress import compress
    from lossycomp.decompress import decompress
    g = golden_case(name)
    x, thr = g["x"], float(g["abs_err"])
    blob = compress(x, thr, precision=prec)
    xhat = decompress(blob)
    err = np.abs(x[..., 0].astype(np.float64) - xhat[..., 0].astype(np.float64))
    assert err.max() <= thr, f"{(err > thr).sum()} violations"
    ref = compress(x, thr, precision="fp32")
    # residual streams (codes + mask) of the 16-bit autoencoder within 1 % (fp16) / 3 % (bf16) of the fp32 one's;
    # the latent slot is not comparable (half-precision codes here, raw fp32 there) and only ever smaller
    import pickle
    sb, sr = pickle.loads(blob), pickle.loads(ref)
    got, want = len(sb[6]) + len(sb[7]), len(sr[6]) + len(sr[7])
    lim = 0.01 if prec == "fp16" else 0.03
    assert abs(got - want) / want < lim, (got, want)
    assert len(sb[0]) < len(sr[0])


def test_blob_records_decoder_precision():
    from lossycomp import _native as N
    from lossycomp.compress import compress
    from lossycomp.decompress import decompress
    g = golden_case("one_chunk")
    blob = compress(g["x"], 0.1, precision="bf16")
    with pytest.raises(N.NativeError):
        decompress(blob, precision="fp32")
    assert decompress(blob).shape == (16, 48, 48, 1)      # picks bf16 from the blob header


@pytest.mark.parametrize("prec,tol", [("fp32", 3e-4), ("fp16", 2e-2)])
def test_other_architectures_run_and_match_oracle(prec, tol):
    """results/models/landsea_model (2-ch, F=20, k=4, no ResBlock) with its shipped weights and the model_1
    architecture (F=20, k=5, res=1; weights not shipped -> seeded random): layers without a tensor-core
    kernel fall back to the CUDA-core back end inside the same precision mode."""
    import os
    from lossycomp._engine import Engine
    from lossycomp.models import MODEL_1_ARCH, Model
    from oracle import lossycomp_oracle as O
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    models = [Model.load(os.path.join(root, "lossy-ml_b200", "weights", "landsea_model.lcw")),
              Model.random(MODEL_1_ARCH, seed=3)]
    g = golden_case("ragged_tails")
    x = g["x"]
    for m in models:
        eng = Engine(m, prec, 0)
        xs, mean, std = O.standardise(x)
        chunks = O.chunk_data(xs)
        lat_ref = O.encoder_forward(m, chunks)
        rec_ref = O.merge_data(O.decoder_forward(m, lat_ref), x.shape)[..., 0]
        lat = eng.encode(dev(x), mean, std).cpu().numpy()
        scale = max(1.0, float(np.abs(lat_ref).max()))
        assert np.abs(lat - lat_ref.reshape(lat.shape)).max() <= tol * scale, m.arch.name
        rec = eng.decode(dev(lat_ref.reshape(lat.shape)), *x.shape[:3]).cpu().numpy()
        assert np.abs(rec - rec_ref).max() <= tol * max(1.0, float(np.abs(rec_ref).max())), m.arch.name


@pytest.mark.parametrize("name,cin", [("basic_model", 1), ("landsea_model", 2), ("coords_model", 5), ("3_convs", 1)])
def test_shipped_model_variants_on_tensor_cores(name, cin):
    """results/models/* (compress_test.py:90-115): the 1-channel, 2-channel (land-sea mask) and 5-channel (lat/lon
    sin/cos coordinates, encodings.py:4-23 + dataLoader.py:284-301) models without a ResBlock, with their shipped
    weights: every full-resolution layer runs on a tensor-core kernel, the autoencoder matches the oracle within the
    fp16 tolerance, and the public API round-trips with zero bound violations (the blob names its model, so
    decompress() finds the weights by itself)."""
    import os
    from lossycomp._engine import Engine
    from lossycomp.compress import compress
    from lossycomp.decompress import decompress
    from lossycomp.encodings import add_coords
    from lossycomp.models import Model
    from oracle import lossycomp_oracle as O
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    path = os.path.join(root, "lossy-ml_b200", "weights", f"{name}.lcw")
    m = Model.load(path)
    assert m.arch.in_channels == cin
    g = golden_case("two_by_three")
    x2 = g["x"]
    T, H, W, _ = x2.shape
    if cin == 1:
        x = np.ascontiguousarray(x2[..., :1])
    elif cin == 2:
        x = x2
    else:
        x = add_coords(x2[..., 0], np.linspace(60.0, 36.25, H), np.linspace(-20.0, 15.75, W))
        assert x.shape == (T, H, W, 5) and np.array_equal(x[..., 0], x2[..., 0])
    eng = Engine(m, "fp16", 0)
    shape = (16, 48, 48)
    for i, L in enumerate(m.layers):
        out = L.out_shape(shape)
        full_res = max(shape) == 48 or max(out) == 48
        if full_res:
            assert eng.layer_backends()[i] != 0, (name, L.name, "runs on CUDA cores")
        shape = out
    mean, std = np.mean(x[..., 0]), np.std(x[..., 0])
    xs = x.copy()
    xs[..., 0] = (x[..., 0] - mean) / std
    chunks = O.chunk_data(xs)
    lat_ref = O.encoder_forward(m, chunks)
    rec_ref = O.merge_data(O.decoder_forward(m, lat_ref), x.shape)[..., 0]
    lat = eng.encode(dev(x), mean, std).cpu().numpy()
    scale = max(1.0, float(np.abs(lat_ref).max()))
    assert np.abs(lat - lat_ref.reshape(lat.shape)).max() <= TOL_VS_ORACLE["fp16"] * scale
    rec = eng.decode(dev(lat_ref.reshape(lat.shape)), T, H, W).cpu().numpy()
    assert np.abs(rec - rec_ref).max() <= TOL_VS_ORACLE["fp16"] * max(1.0, float(np.abs(rec_ref).max()))
    blob = compress(x, 0.1, model=path)
    xhat = decompress(blob)
    assert xhat.shape == (T, H, W, 1)
    assert np.abs(x[..., 0].astype(np.float64) - xhat[..., 0]).max() <= 0.1
    assert len(blob) < x[..., 0].nbytes / 4
