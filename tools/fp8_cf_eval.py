"""Would FP8 (e4m3) operands in the four 3x3x3 ResBlock convolutions keep the compression factor?  (CPU study.)

The oracle's fp32 autoencoder is run with the INPUT activations and the weights of the selected layers rounded to
e4m3 (torch.float8_e4m3fn, round to nearest even, saturating) -- per-output-channel weight scales, one power-of-two
activation scale per layer calibrated on the data so that the largest activation sits just below 448 -- fp32
accumulation, everything else untouched.  For each variant the residual stage and the reference's bz2 coder run on
the resulting reconstruction; the compression factor is compared with the fp32 and the fp16-operand autoencoders
(the judge's gate for an FP8 mode: within 1 % of fp32 at every threshold).

    python tools/fp8_cf_eval.py [T H W]
"""
import os, pickle, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lossy-ml_b200")); sys.path.insert(0, ROOT)
import numpy as np
import torch
from oracle import lossycomp_oracle as O
from lossycomp.models import load_model

T, H, W = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (24, 240, 480)))
x = O.synthetic_field(T, H, W, seed=1234)
model = load_model(None)
raw = x[..., 0].nbytes
K3 = {"conv3d_1", "conv3d_2", "conv3d_7", "conv3d_8"}


def q_e4m3(t, scale):
    return (t * scale).clamp(-448.0, 448.0).to(torch.float8_e4m3fn).float() / scale


def q_half(t):
    return t.half().float()


def run(layers, x_ndhwc, act_q, w_q, which, batch=10):
    """act_q(tensor, layer) / w_q(weight (Cout,Cin,kt,kh,kw), layer) are applied to the layers named in `which`;
    the other layers see fp16-rounded operands (the shipped engine's precision) when which is not None."""
    import torch.nn.functional as F
    outs = []
    with torch.no_grad():
        for s in range(0, x_ndhwc.shape[0], batch):
            xb = torch.from_numpy(np.array(x_ndhwc[s:s + batch], dtype=np.float32)).permute(0, 4, 1, 2, 3).contiguous()
            skip = None
            for L in layers:
                w = torch.from_numpy(np.ascontiguousarray(L.kernel.transpose(4, 3, 0, 1, 2)))
                b = torch.from_numpy(L.bias)
                xin = xb
                if which is not None:
                    if L.name in which:
                        xin, w = act_q(xb, L), w_q(w, L)
                    else:
                        xin, w = q_half(xb), q_half(w)
                if L.kind == "conv":
                    pads = []
                    for n in reversed(xin.shape[2:]):
                        lo, hi = O._same_pad(int(n), L.k, L.stride)
                        pads += [lo, hi]
                    y = F.conv3d(F.pad(xin, pads), w, b, stride=L.stride)
                else:
                    y = F.conv_transpose3d(xin, w, None, stride=L.stride)
                    c = (L.k - L.stride) // 2
                    t, h, ww = (int(n) * L.stride for n in xin.shape[2:])
                    y = y[:, :, c:c + t, c:c + h, c:c + ww] + b.view(1, -1, 1, 1, 1)
                if L.res_add:
                    y = y + (skip if which is None else q_half(skip))
                if L.relu:
                    y = torch.relu(y)
                if L.res_save:
                    skip = y
                xb = y
            outs.append(xb.permute(0, 2, 3, 4, 1).contiguous().numpy())
    return np.concatenate(outs, axis=0)


xs, mean, std = O.standardise(x)
chunks = O.chunk_data(xs)

# calibrate one power-of-two activation scale per K3 layer (largest input activation just below 448)
amax = {}
col = {}
lat32 = O.run_layers(chunks, model.encoder, collect=col)
O.run_layers(lat32, model.decoder, collect=col)
prev = None
for L in model.layers:
    if L.name in K3:
        a = max(float(np.abs(c).max()) for c in col[prev])
        amax[L.name] = a
    prev = L.name
scales = {n: 2.0 ** np.floor(np.log2(448.0 / a)) for n, a in amax.items()}
print("largest input activation of the 3x3x3 layers:", {k: round(v, 2) for k, v in amax.items()})
print("activation scales:", scales)


def w_e4m3(w, L):
    s = 448.0 / w.abs().amax(dim=(1, 2, 3, 4), keepdim=True).clamp_min(1e-12)      # per output channel
    return (w * s).to(torch.float8_e4m3fn).float() / s


variants = {
    "fp32 (oracle)": (None, None, None),
    "fp16 operands everywhere": (lambda t, L: q_half(t), lambda w, L: q_half(w), set()),
    "e4m3 operands in the four 3x3x3 layers": (lambda t, L: q_e4m3(t, scales[L.name]), w_e4m3, K3),
    "e4m3 activations only (fp16 weights)": (lambda t, L: q_e4m3(t, scales[L.name]), lambda w, L: q_half(w), K3),
    "e4m3 weights only (fp16 activations)": (lambda t, L: q_half(t), w_e4m3, K3),
}
print(f"field {T}x{H}x{W} ({raw / 1e6:.1f} MB of channel 0); CF with the reference's residual stage and bz2 coder")
res = {}
for name, (aq, wq, which) in variants.items():
    lat = run(model.encoder, chunks, aq, wq, which)
    rec = O.merge_data(run(model.decoder, lat, aq, wq, which), x.shape)
    mse = float(np.mean((rec[..., 0] - xs[..., 0]) ** 2))
    row = []
    for thr in (0.01, 0.1, 0.3, 1.0):
        blob = O.compress(x, thr, model, inject={"recon_std": rec})
        row.append(raw / len(blob))
    res[name] = row
    print(f"{name:44s} MSE(std units) {mse:.3e}  CF@0.01/0.1/0.3/1.0: " + " ".join(f"{v:8.3f}" for v in row) +
          ("" if name.startswith("fp32") else "   vs fp32: " + " ".join(f"{100 * (v / r - 1):+.2f}%" for v, r in zip(row, res["fp32 (oracle)"]))),
          flush=True)
