"""An independent, loop-level NumPy statement of Keras Conv3D / Conv3DTranspose with padding="same", written from the
framework's definition (not from torch), against the oracle's torch-CPU layers -- the cross-check that pins the
oracle's convolution conventions (SURVEY.md 8c a-M) without TensorFlow.

Definitions used (channels-last, kernel (kt,kh,kw,Cin,Cout), cross-correlation):
  Conv3D, stride s:      out = ceil(in/s); total = max((out-1)*s + k - in, 0); before = total//2
                         y[o] = b + sum_k x[o*s + k - before] . w[k]            (x = 0 outside)
  Conv3DTranspose, s=2:  kernel (kt,kh,kw,Cout,Cin); it is the gradient of the stride-s "same" convolution that maps
                         the OUTPUT grid (2n) to the INPUT grid (n): that convolution pads before = (k-s)//2, so
                         y[j] = b + sum_i sum_k [i*s + k - before == j] x[i] . w[k]
"""
import numpy as np
import pytest

from oracle import lossycomp_oracle as O


def conv3d_same(x, w, b, s):
    """x (T,H,W,Cin) float64, w (k,k,k,Cin,Cout)."""
    k = w.shape[0]
    dims = x.shape[:3]
    out = [-(-n // s) for n in dims]
    before = [max((o - 1) * s + k - n, 0) // 2 for o, n in zip(out, dims)]
    y = np.zeros(tuple(out) + (w.shape[4],), dtype=np.float64) + b
    for a in range(k):
        for c in range(k):
            for d in range(k):
                # output positions o whose input position o*s + tap - before lies inside the array
                sl_o, sl_i = [], []
                for tap, n, o_n, bf in zip((a, c, d), dims, out, before):
                    o_lo = max(0, -(-(bf - tap) // s))
                    o_hi = min(o_n - 1, (n - 1 + bf - tap) // s)
                    sl_o.append(slice(o_lo, o_hi + 1))
                    sl_i.append(slice(o_lo * s + tap - bf, o_hi * s + tap - bf + 1, s))
                y[tuple(sl_o)] += x[tuple(sl_i)] @ w[a, c, d]
    return y


def conv3d_transpose_same(x, w, b, s=2):
    """x (T,H,W,Cin) float64, w (k,k,k,Cout,Cin) -> (2T,2H,2W,Cout)."""
    k = w.shape[0]
    dims = x.shape[:3]
    out = [n * s for n in dims]
    before = (k - s) // 2
    y = np.zeros(tuple(out) + (w.shape[3],), dtype=np.float64) + b
    for a in range(k):
        for c in range(k):
            for d in range(k):
                sl_o, sl_i = [], []
                for tap, n, o_n in zip((a, c, d), dims, out):
                    i_lo = max(0, -(-(before - tap) // s))
                    i_hi = min(n - 1, (o_n - 1 + before - tap) // s)
                    sl_i.append(slice(i_lo, i_hi + 1))
                    sl_o.append(slice(i_lo * s + tap - before, i_hi * s + tap - before + 1, s))
                y[tuple(sl_o)] += x[tuple(sl_i)] @ w[a, c, d].T
    return y


def run_definition(x, layers):
    """Every layer is fed the ORACLE's input of that layer (so differences do not accumulate); returns the
    largest relative difference per layer."""
    col = {}
    O.run_layers(x[None].astype(np.float32), layers, batch=1, collect=col)
    cur = x.astype(np.float64)
    worst = {}
    skip = None
    for L in layers:
        w, b = L.kernel.astype(np.float64), L.bias.astype(np.float64)
        y = conv3d_same(cur, w, b, L.stride) if L.kind == "conv" else conv3d_transpose_same(cur, w, b, L.stride)
        if L.res_add:
            y = y + skip
        if L.relu:
            y = np.maximum(y, 0.0)
        ref = col[L.name][0][0].astype(np.float64)
        assert ref.shape == y.shape, (L.name, ref.shape, y.shape)
        worst[L.name] = float(np.abs(y - ref).max() / max(1.0, np.abs(ref).max()))
        cur = ref                       # continue from the oracle's activations
        if L.res_save:
            skip = cur
    return worst


def _chunk(seed=5):
    x = O.synthetic_field(16, 48, 48, seed=seed, h0=250, w0=700)
    xs, _, _ = O.standardise(x)
    return xs[..., :]


def test_model_2_layers_match_the_keras_definition(model2):
    """All 14 layers of final_models/model_2 (k4 s1 pad 1/2, k3 s1, k4 s2, transposed k4 s2, ResBlock adds)."""
    x = _chunk()
    w_enc = run_definition(x, model2.encoder)
    lat = O.run_layers(x[None].astype(np.float32), model2.encoder, batch=1)[0]
    w_dec = run_definition(lat, model2.decoder)
    for name, d in {**w_enc, **w_dec}.items():
        assert d <= 1e-5, (name, d)        # fp32 torch layer vs fp64 definition
    assert len(w_enc) == 7 and len(w_dec) == 7


def test_model_1_architecture_k5_matches_the_keras_definition():
    """k = 5 (compress.py:41,57 hard-codes final_models/model_1): stride-1 pad 2/2, stride-2 pad 1/2, transposed
    crop 1/2 -- seeded random weights (the trained ones are not in the reference checkout)."""
    from lossycomp.models import MODEL_1_ARCH, Model
    m = Model.random(MODEL_1_ARCH, seed=3)
    x = _chunk(seed=6)
    w_enc = run_definition(x, m.encoder)
    lat = O.run_layers(x[None].astype(np.float32), m.encoder, batch=1)[0]
    w_dec = run_definition(lat, m.decoder)
    for name, d in {**w_enc, **w_dec}.items():
        assert d <= 1e-5, (name, d)


def test_a_shifted_convention_is_detected(model2):
    """The check has teeth: padding 2/1 instead of 1/2 for the k=4 stride-1 layer is far outside the tolerance."""
    x = _chunk().astype(np.float64)
    L = model2.encoder[0]
    col = {}
    O.run_layers(x[None].astype(np.float32), [L], batch=1, collect=col)
    ref = col[L.name][0][0]
    w, b = L.kernel.astype(np.float64), L.bias.astype(np.float64)
    good = np.maximum(conv3d_same(x, w, b, 1), 0)
    bad = np.maximum(conv3d_same(x[::-1, ::-1, ::-1], w[::-1, ::-1, ::-1], b, 1)[::-1, ::-1, ::-1], 0)   # pads 2/1
    assert np.abs(good - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max())
    assert np.abs(bad - ref).max() > 1e-2
