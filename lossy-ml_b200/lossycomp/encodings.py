"""Coordinate encodings for the extra input channels of the 5-channel model (results/models/coords_model).

Host-side mirror of the reference's lossycomp/encodings.py:4-23 (`encode_lon`, `encode_lat`: same names, same
formulas, same assertions) plus the channel stacking its data generator does (dataLoader.py:284-301): the codec
input of the coords model is (time, lat, lon, 5) = [variable, sin(lat), cos(lat), sin(lon), cos(lon)] with the
latitude / longitude features broadcast over time.  Only channel 0 is standardised and reconstructed; the other
channels are side information for the autoencoder, like the land-sea mask of the 2-channel models.
"""
from __future__ import annotations

import math

import numpy as np


def encode_lon(lon):
    """Encode longitude across the globe from [-1, 1] using two features (encodings.py:4-12)."""
    assert lon <= 180
    assert lon >= -180
    lon = (lon + 180) / 360
    return math.sin(2 * math.pi * lon), math.cos(2 * math.pi * lon)


def encode_lat(lat):
    """Encode latitude across the globe from [-1, 1] using two features (encodings.py:15-23)."""
    assert lat <= 90
    assert lat >= -90
    lat = (lat + 90) / 180
    return math.sin(math.pi * lat), math.cos(math.pi * lat)


def add_coords(data, lat, lon):
    """(T, H, W) or (T, H, W, 1) variable + latitude (H,) / longitude (W,) in degrees -> (T, H, W, 5) float32 in the
    channel order of dataLoader.py:284-301: data, sin(lat), cos(lat), sin(lon), cos(lon)."""
    data = np.asarray(data, dtype=np.float32)
    if data.ndim == 3:
        data = data[..., None]
    T, H, W, _ = data.shape
    lat_st = np.array([encode_lat(float(v)) for v in lat], dtype=np.float64)      # (H, 2)
    lon_st = np.array([encode_lon(float(v)) for v in lon], dtype=np.float64)      # (W, 2)
    assert lat_st.shape[0] == H and lon_st.shape[0] == W
    out = np.empty((T, H, W, 5), dtype=np.float32)
    out[..., 0] = data[..., 0]
    out[..., 1] = lat_st[None, :, None, 0]
    out[..., 2] = lat_st[None, :, None, 1]
    out[..., 3] = lon_st[None, None, :, 0]
    out[..., 4] = lon_st[None, None, :, 1]
    return out
