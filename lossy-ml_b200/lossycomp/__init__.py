"""lossycomp -- B200-native drop-in for the SilkeDH/lossy-ml codec hot path.

Public API (same import paths and signatures as the reference):
    lossycomp.compress.compress(array, err_threshold, verbose=False) -> bytes
    lossycomp.decompress.decompress(compressed_data, verbose=False) -> np.ndarray
"""
__version__ = "0.1.0"
