"""Convert the Keras full-model HDF5 files shipped with the reference into the flat
`.lcw` weight files the B200 codec loads (lossycomp/models.py::Model.load).

Run in the build container only (the GPU box has no /root/reference):
    python tools/convert_weights.py
"""
import os
import pickle
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "lossy-ml_b200"))
sys.path.insert(0, HERE)

from hdf5_min import H5File  # noqa: E402
from lossycomp.models import Arch, Model  # noqa: E402

REF = "/root/reference/results"

# (output name, hdf5 path, history pickle, input channels)
SOURCES = [
    ("model_2", "final_models/model_2/weigths/weight.hdf5", "final_models/model_2/model-history.pkl", 2),
    ("landsea_model", "models/landsea_model/weigths/weight.hdf5", None, 2),
    # 1-channel and 5-channel (lat/lon sin/cos coordinates) variants of compress_test.py:90-115
    ("basic_model", "models/basic_model/weights/weight.hdf5", None, 1),
    ("coords_model", "models/coords_model/weigths/weight.hdf5", None, 5),
    ("3_convs", "models/3_convs/weights/weight.hdf5", None, 1),
]


def convert(name, h5, hist, cin):
    f = H5File(os.path.join(REF, h5))
    ds = f.walk_datasets("/model_weights")
    if hist:
        with open(os.path.join(REF, hist), "rb") as fh:
            p = pickle.load(fh)["parameters"]
        arch = Arch(filters=p["num_filters"], kernel=p["kernel_size"], res_blocks=p["res_blocks"],
                    num_convs=p["num_convs"], in_channels=cin, name=name)
    else:  # results/models/*: 4 reducing convs, no res block (compress_test.py:101-115)
        k = ds["Encoder/conv3d/kernel:0"].shape
        assert k[3] == cin, (name, k)
        n_convs = sum(1 for key in ds if key.startswith("Encoder/") and key.endswith("kernel:0"))
        arch = Arch(filters=k[4], kernel=k[0], res_blocks=0, num_convs=n_convs, in_channels=cin, name=name)
    layers = arch.layers()
    for L in layers:
        L.kernel = ds[f"{L.part}/{L.name}/kernel:0"]
        L.bias = ds[f"{L.part}/{L.name}/bias:0"]
    used = {f"{L.part}/{L.name}/{t}:0" for L in layers for t in ("kernel", "bias")}
    assert used == set(ds), (sorted(set(ds) - used), sorted(used - set(ds)))
    out = os.path.join(ROOT, "lossy-ml_b200", "weights", f"{name}.lcw")
    Model(arch, layers).save(out)
    m = Model.load(out)
    n = sum(L.kernel.size + L.bias.size for L in m.layers)
    print(f"{name}: {n} parameters -> {out} ({os.path.getsize(out)} bytes), MACs/chunk {arch.macs_per_chunk()}")


if __name__ == "__main__":
    for s in SOURCES:
        convert(*s)
