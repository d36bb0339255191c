"""Generate tests/golden/*.npz by running the REFERENCE's own lossycomp/compress.py and
lossycomp/decompress.py, unmodified, from /root/reference under import stubs.

Runs in the build container only (needs /root/reference).  What is stubbed, and why
(SURVEY.md §8c): tensorflow / xarray / fpzip are not installable here, and
`lossycomp.models` builds a Keras graph.  The stubs supply
  * tensorflow: random.set_seed, data.Dataset.from_tensor_slices(a).batch(n), keras.utils.Sequence
  * xarray: empty module (only the training loaders use it)
  * fpzip: compress/decompress stand-ins (lossless bz2 of the float32 bytes)
  * lossycomp.models.Autoencoder2.build: sub-models whose .predict runs the torch-CPU fp32
    convolution restatement (oracle/lossycomp_oracle.py::run_layers) with the shipped
    final_models/model_2 weights, batch 10 like the reference's predict loop.
Everything else -- statistics, chunk/merge, mask, quantisation, differencing, bz2, the blob and
the whole decompress arithmetic -- is executed by the reference's code, so the goldens pin the
oracle (and through it the CUDA path) bit-exactly for every non-NN stage.

    python tools/make_goldens.py
"""
import bz2 as _bz2
import hashlib
import importlib.util
import os
import pickle
import sys
import types

import numpy as np

sys.dont_write_bytecode = True   # /root/reference is read-only
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


b200_models = _load("b200_models", os.path.join(ROOT, "lossy-ml_b200", "lossycomp", "models.py"))
oracle = _load("lc_oracle", os.path.join(ROOT, "oracle", "lossycomp_oracle.py"))
MODEL = b200_models.Model.load(os.path.join(ROOT, "lossy-ml_b200", "weights", "model_2.lcw"))


# ---- stubs -----------------------------------------------------------------
class _DS:
    def __init__(self, a):
        self.a = a

    def batch(self, n):
        self.n = n
        return self


tf = types.ModuleType("tensorflow")
tf.random = types.SimpleNamespace(set_seed=lambda s: None)
tf.data = types.SimpleNamespace(Dataset=types.SimpleNamespace(from_tensor_slices=lambda a: _DS(a)))
tf.keras = types.ModuleType("tensorflow.keras")
tf.keras.utils = types.SimpleNamespace(Sequence=object)
sys.modules["tensorflow"] = tf
sys.modules["tensorflow.keras"] = tf.keras
sys.modules["xarray"] = types.ModuleType("xarray")

fpzip = types.ModuleType("fpzip")
fpzip.compress = lambda a, precision=0, order="C": _bz2.compress(np.ascontiguousarray(a, dtype="<f4").tobytes(), 9)
fpzip.decompress = lambda b, order="C": _bz2.decompress(b)
sys.modules["fpzip"] = fpzip

CAPTURE = {}


class _Sub:
    def __init__(self, layers, tag):
        self.layers_, self.tag = layers, tag

    def predict(self, ds, steps=None):
        out = oracle.run_layers(np.asarray(ds.a, dtype=np.float32), self.layers_, batch=ds.n)
        CAPTURE[self.tag] = out
        return out


class _Model:
    def __init__(self):
        self.layers = [None, _Sub(MODEL.encoder, "latent"), _Sub(MODEL.decoder, "decoded")]

    def load_weights(self, path):
        pass

    def summary(self):
        return ""


class Autoencoder2:
    @staticmethod
    def build(*a, **k):
        m = _Model()
        return (m.layers[1], m.layers[2], m)


sys.path.insert(0, REF)
import lossycomp  # noqa: E402  (the reference package, /root/reference/lossycomp)
assert lossycomp.__file__.startswith(REF), lossycomp.__file__
stub_models = types.ModuleType("lossycomp.models")
stub_models.Autoencoder2 = Autoencoder2
sys.modules["lossycomp.models"] = stub_models
os.chdir(os.path.join(REF, "lossycomp"))          # compress.py:41 opens '../results/...' relative to CWD

import lossycomp.compress as ref_c  # noqa: E402
import lossycomp.decompress as ref_d  # noqa: E402
import lossycomp.dataLoader as ref_dl  # noqa: E402
import lossycomp.huffman as ref_h  # noqa: E402

# capture the arrays the reference hands to bz2.compress (the differenced code / mask streams)
_real_bz2_compress = _bz2.compress
_bz2_log = []


def _spy(data, level=9):
    if isinstance(data, np.ndarray):
        _bz2_log.append(data.copy())
    return _real_bz2_compress(data, level)


ref_c.bz2 = types.SimpleNamespace(compress=_spy, decompress=_bz2.decompress)

CASES = [  # name, shape, abs_err, seed
    ("one_chunk", (16, 48, 48), 0.1, 11),
    ("ragged_tails", (17, 49, 50), 0.05, 12),
    ("two_by_three", (24, 96, 144), 0.1, 13),
    ("odd_all_axes", (33, 100, 97), 0.6, 14),
]


# digest-only goldens (arrays too large to commit: only hashes and scalars are kept, the input is regenerated from
# its seed): the notebook's shape and error bound (notebooks/Compression_example.ipynb:128-135: (50,241,241,2), 0.6)
DIGEST_CASES = [
    ("notebook_50x241x241", (50, 241, 241), 0.6, 15),
]


def sha(b):
    return hashlib.sha256(bytes(b)).hexdigest()


def digest_input(T, H, W, seed):
    return oracle.synthetic_field(T, H, W, seed=seed, h0=100 + 7 * seed, w0=200 + 31 * seed)


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, (T, H, W), thr, seed in CASES:
        x = oracle.synthetic_field(T, H, W, seed=seed, h0=200 + 7 * seed, w0=300 + 31 * seed)
        # make the land-sea channel fractional on part of the field, like real lsm
        x[..., 1] = np.clip(x[..., 1] * 0.75 + 0.1 * np.sin(np.arange(W, dtype=np.float32) * 0.37)[None, None, :], 0, 1)
        _bz2_log.clear()
        CAPTURE.clear()
        blob = ref_c.compress(x, thr)
        slots = pickle.loads(blob)
        dq, dm = _bz2_log[0], _bz2_log[1]
        latent = CAPTURE["latent"].copy()
        decoded = CAPTURE["decoded"].copy()
        recon_std = ref_dl.merge_data(decoded, x.shape)
        xhat = ref_d.decompress(blob)
        viol = int((np.abs(x[..., 0].astype(np.float64) - xhat[..., 0].astype(np.float64)) > thr).sum())
        np.savez_compressed(
            os.path.join(OUT, f"{name}.npz"),
            x=x, abs_err=np.float64(thr), mean=slots[4], std=slots[5],
            latent=latent, recon_std=recon_std.astype(np.float32),
            dq=dq.astype(np.int32), dm=dm.astype(np.int8), thr2=np.float64(slots[8]),
            latent_shape=np.array(slots[1]), array_shape=np.array(slots[2]), error_shape=np.array(slots[3]),
            comp_error_sha=sha(slots[6]), comp_mask_sha=sha(slots[7]),
            comp_error_len=len(slots[6]), comp_mask_len=len(slots[7]),
            xhat=xhat, violations=viol,
            slot_types=np.array([type(s).__name__ for s in slots]),
        )
        print(f"{name}: shape {x.shape} thr {thr} chunks {latent.shape[0]} outliers {dq.size} "
              f"({dq.size / (T * H * W):.3f}) blob {len(blob)} B  CF {x[..., 0].nbytes / len(blob):.2f} "
              f"violations {viol} max|e| {np.abs(x[..., 0] - xhat[..., 0]).max():.7f}")

    import json
    digests = {}
    for name, (T, H, W), thr, seed in DIGEST_CASES:
        x = digest_input(T, H, W, seed)
        _bz2_log.clear()
        CAPTURE.clear()
        blob = ref_c.compress(x, thr)
        slots = pickle.loads(blob)
        dq, dm = _bz2_log[0], _bz2_log[1]
        recon_std = ref_dl.merge_data(CAPTURE["decoded"], x.shape).astype(np.float32)
        xhat = ref_d.decompress(blob)
        viol = int((np.abs(x[..., 0].astype(np.float64) - xhat[..., 0].astype(np.float64)) > thr).sum())
        digests[name] = dict(
            shape=[T, H, W], abs_err=thr, seed=seed, x_sha=sha(x.tobytes()), mean=float(slots[4]), std=float(slots[5]),
            mean_hex=np.float32(slots[4]).tobytes().hex(), std_hex=np.float32(slots[5]).tobytes().hex(),
            latent_sha=sha(CAPTURE["latent"].astype("<f4").tobytes()), recon_std_sha=sha(recon_std.tobytes()),
            dq_sha=sha(dq.astype("<i4").tobytes()), dm_sha=sha(dm.astype("i1").tobytes()),
            comp_error_sha=sha(slots[6]), comp_mask_sha=sha(slots[7]), xhat_sha=sha(xhat.astype("<f4").tobytes()),
            n_chunks=int(slots[1][0]), n_out=int(slots[3][0]), thr2=float(slots[8]), violations=viol,
            blob_len=len(blob), cf=x[..., 0].nbytes / len(blob),
            max_abs_err=float(np.abs(x[..., 0].astype(np.float64) - xhat[..., 0]).max()))
        print(name, digests[name])
    with open(os.path.join(OUT, "digests.json"), "w") as fh:
        json.dump(digests, fh, indent=1, sort_keys=True)

    # chunk/merge known answers (notebooks/Chunking_data.ipynb cells 4, 8, 15, 19)
    facts = {}
    for shp in [(16, 720, 720, 1), (32, 721, 1440, 1), (50, 241, 241, 2), (17, 49, 50, 2)]:
        a = np.arange(np.prod(shp), dtype=np.float32).reshape(shp)
        c = ref_dl.chunk_data(a, (16, 48, 48))
        assert np.all(ref_dl.merge_data(c, shp) == a)
        facts[str(shp)] = (c.shape[0], [float(c[i, 0, 0, 0, 0]) for i in range(min(8, c.shape[0]))])
    hk = {s: ref_h.getHuffmanCode(s) for s in ["0010-1002-20001-10", "aabbccdd", "aaaaaaaa",
                                               "1,0,0,-1,2,0,0,0,1,-2,0,0,1,0,0,0,0,-1"]}
    with open(os.path.join(OUT, "kats.pkl"), "wb") as fh:
        pickle.dump({"chunk_facts": facts, "huffman": hk}, fh, protocol=4)
    print("chunk facts", facts)
    print("huffman", hk)


if __name__ == "__main__":
    main()
