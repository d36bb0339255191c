"""Architecture spec of the convolutional autoencoder on the codec path.

Mirrors `Autoencoder2.build` of the reference (lossycomp/models.py:47-87, ResBlock
:90-97) as a *layer list* instead of a Keras graph: the B200 implementation has no
graph framework, the list is what the CUDA conv kernels are configured from and
what the weight file (`weights/*.lcw`, written by tools/convert_weights.py) is
indexed by.

Keras conventions that every consumer of this spec must follow (SURVEY.md §8c a-M):
  * tensors are channels-last (N, T, H, W, C); Conv3D kernels are
    (kt, kh, kw, C_in, C_out), cross-correlation, Conv3DTranspose kernels are
    (kt, kh, kw, C_out, C_in);
  * padding="same": out = ceil(in / s), total = max((out-1)*s + k - in, 0),
    before = total // 2, after = total - before  (k=4,s=1 -> 1 before / 2 after);
  * Conv3DTranspose "same", stride 2: out = 2*in; equals the full transposed
    convolution cropped by (k-2)//2 at the front of every axis.
"""
from __future__ import annotations

import json
import os
import struct
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import numpy as np

CHUNK = (16, 48, 48)  # (time, lat, lon) of one chunk, compress.py:76
MAGIC = b"LCW1"


@dataclass
class Layer:
    name: str          # Keras layer name inside the Encoder/Decoder sub-model
    part: str          # "Encoder" | "Decoder"
    kind: str          # "conv" | "convT"
    k: int
    stride: int
    cin: int
    cout: int
    relu: bool         # ReLU fused on the output
    res_save: bool = False   # output of this layer is the skip input of the next ResBlock
    res_add: bool = False    # add the saved skip tensor before the ReLU (models.py:95-96)
    relu_in_place: bool = True
    kernel: np.ndarray = field(default=None, repr=False)  # Keras layout, float32
    bias: np.ndarray = field(default=None, repr=False)

    def macs(self, in_shape: Tuple[int, int, int]) -> int:
        t, h, w = in_shape
        if self.kind == "conv":
            ot, oh, ow = (-(-t // self.stride), -(-h // self.stride), -(-w // self.stride))
            return ot * oh * ow * self.k ** 3 * self.cin * self.cout
        return t * h * w * self.k ** 3 * self.cin * self.cout

    def out_shape(self, in_shape):
        t, h, w = in_shape
        s = self.stride
        if self.kind == "conv":
            return (-(-t // s), -(-h // s), -(-w // s))
        return (t * s, h * s, w * s)


@dataclass
class Arch:
    filters: int
    kernel: int
    res_blocks: int
    num_convs: int
    in_channels: int
    name: str = ""

    def layers(self) -> List[Layer]:
        """Layer list in execution order, named as Keras names them (creation order)."""
        F, k = self.filters, self.kernel
        out: List[Layer] = []
        nconv = [0]
        nconvT = [0]

        def cname():
            n = "conv3d" if nconv[0] == 0 else f"conv3d_{nconv[0]}"
            nconv[0] += 1
            return n

        def tname():
            n = "conv3d_transpose" if nconvT[0] == 0 else f"conv3d_transpose_{nconvT[0]}"
            nconvT[0] += 1
            return n

        cin = self.in_channels
        # --- encoder (models.py:52-67)
        if self.res_blocks > 0:
            out.append(Layer(cname(), "Encoder", "conv", k, 1, cin, F, relu=True))
            cin = F
        for _ in range(self.res_blocks):
            out[-1].res_save = True
            out.append(Layer(cname(), "Encoder", "conv", 3, 1, F, F, relu=True))
            out.append(Layer(cname(), "Encoder", "conv", 3, 1, F, F, relu=True, res_add=True))
        for _ in range(self.num_convs):
            out.append(Layer(cname(), "Encoder", "conv", k, 2, cin, F, relu=True))
            cin = F
        # --- decoder (models.py:70-85)
        for _ in range(self.num_convs):
            out.append(Layer(tname(), "Decoder", "convT", k, 2, F, F, relu=True))
        for _ in range(self.res_blocks):
            out[-1].res_save = True
            out.append(Layer(cname(), "Decoder", "conv", 3, 1, F, F, relu=True))
            out.append(Layer(cname(), "Decoder", "conv", 3, 1, F, F, relu=True, res_add=True))
        out.append(Layer(cname(), "Decoder", "conv", k, 1, F, 1, relu=False))
        return out

    def latent_shape(self) -> Tuple[int, int, int, int]:
        t, h, w = CHUNK
        d = 2 ** self.num_convs
        return (t // d, h // d, w // d, self.filters)

    def macs_per_chunk(self) -> Dict[str, int]:
        shape = CHUNK
        tot = {"Encoder": 0, "Decoder": 0}
        for L in self.layers():   # the decoder starts from the encoder's output shape
            tot[L.part] += L.macs(shape)
            shape = L.out_shape(shape)
        return tot

    def encoder_layers(self):
        return [L for L in self.layers() if L.part == "Encoder"]

    def decoder_layers(self):
        return [L for L in self.layers() if L.part == "Decoder"]


class Model:
    """An `Arch` with weights attached (Keras layouts, float32)."""

    def __init__(self, arch: Arch, layers: List[Layer]):
        self.arch = arch
        self.layers = layers

    @property
    def encoder(self):
        return [L for L in self.layers if L.part == "Encoder"]

    @property
    def decoder(self):
        return [L for L in self.layers if L.part == "Decoder"]

    # -- serialisation ------------------------------------------------------
    def save(self, path: str) -> None:
        meta = {
            "arch": dict(filters=self.arch.filters, kernel=self.arch.kernel,
                         res_blocks=self.arch.res_blocks, num_convs=self.arch.num_convs,
                         in_channels=self.arch.in_channels, name=self.arch.name),
            "tensors": [],
        }
        blobs = []
        off = 0
        for L in self.layers:
            for nm, arr in (("kernel", L.kernel), ("bias", L.bias)):
                a = np.ascontiguousarray(arr, dtype="<f4")
                meta["tensors"].append({"layer": f"{L.part}/{L.name}", "name": nm,
                                        "shape": list(a.shape), "offset": off})
                blobs.append(a.tobytes())
                off += a.nbytes
        hdr = json.dumps(meta).encode()
        with open(path, "wb") as fh:
            fh.write(MAGIC + struct.pack("<I", len(hdr)) + hdr + b"".join(blobs))

    @staticmethod
    def load(path: str) -> "Model":
        with open(path, "rb") as fh:
            raw = fh.read()
        if raw[:4] != MAGIC:
            raise ValueError(f"{path}: not a lossycomp weight file")
        (n,) = struct.unpack("<I", raw[4:8])
        meta = json.loads(raw[8:8 + n])
        base = 8 + n
        arch = Arch(**meta["arch"])
        tensors = {}
        for t in meta["tensors"]:
            cnt = int(np.prod(t["shape"]))
            tensors[(t["layer"], t["name"])] = np.frombuffer(
                raw, "<f4", count=cnt, offset=base + t["offset"]).reshape(t["shape"]).copy()
        layers = arch.layers()
        for L in layers:
            L.kernel = tensors[(f"{L.part}/{L.name}", "kernel")]
            L.bias = tensors[(f"{L.part}/{L.name}", "bias")]
            exp = (L.k, L.k, L.k, L.cin, L.cout) if L.kind == "conv" else (L.k, L.k, L.k, L.cout, L.cin)
            if tuple(L.kernel.shape) != exp:
                raise ValueError(f"{L.name}: kernel shape {L.kernel.shape} != {exp}")
        return Model(arch, layers)

    @staticmethod
    def random(arch: Arch, seed: int = 1) -> "Model":
        """Seeded Glorot-uniform weights (for architectures whose trained weights are not shipped,
        e.g. final_models/model_1 -- see /root/reference/.MISSING_LARGE_BLOBS:1)."""
        rng = np.random.default_rng(seed)
        layers = arch.layers()
        for L in layers:
            fan = L.k ** 3 * (L.cin + L.cout)
            lim = np.sqrt(6.0 / fan)
            shp = (L.k, L.k, L.k, L.cin, L.cout) if L.kind == "conv" else (L.k, L.k, L.k, L.cout, L.cin)
            L.kernel = rng.uniform(-lim, lim, size=shp).astype(np.float32)
            L.bias = rng.uniform(-0.05, 0.05, size=(L.cout,)).astype(np.float32)
        return Model(arch, layers)


_WEIGHT_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "weights")

# final_models/model_1 is what compress.py:41,57 hard-codes, but its weights are not
# in the reference checkout; model_2 (F=23, k=4) is shipped and is our default.
MODEL_1_ARCH = Arch(filters=20, kernel=5, res_blocks=1, num_convs=4, in_channels=2, name="model_1")
MODEL_2_ARCH = Arch(filters=23, kernel=4, res_blocks=1, num_convs=4, in_channels=2, name="model_2")


def default_model_path() -> str:
    return os.environ.get("LOSSYCOMP_MODEL", os.path.join(_WEIGHT_DIR, "model_2.lcw"))


def shipped_model_path(name: str):
    """weights/<name>.lcw if this package ships weights of that name, else None."""
    if not name or os.sep in name or name.startswith("."):
        return None
    p = os.path.join(_WEIGHT_DIR, f"{name}.lcw")
    return p if os.path.exists(p) else None


_cache: Dict[str, Model] = {}


def load_model(path: str | None = None) -> Model:
    path = path or default_model_path()
    if path not in _cache:
        _cache[path] = Model.load(path)
    return _cache[path]
