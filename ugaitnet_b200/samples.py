"""The reference's on-disk sample format: one ``.h5`` file per 25-frame sub-sequence, written by
``dd.io.save(path, dict)`` in data/generateOFData.py:136-148 -- keys ``data`` (int16 [60,60,50] optical flow scaled by
compressFactor = 100, or uint8 [60,60,25] gray / depth / silhouette), ``label``, ``videoId``, ``gait``, ``frames``,
``bbs``, ``compressFactor`` [, ``cam``] -- and read by ``__load_dd`` (data/mj_dataGeneratorMMUWYHsingle.py:294-338).

deepdish stores a dict as the root group: numpy arrays as datasets, Python / numpy scalars as attributes of the
root group or 0-d datasets.  ``load_sample`` reads either through ugaitnet_b200.hdf5 (no h5py / pytables in the image);
``decode_sample`` restates ``__load_dd`` and the generator's channel-first layout.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np

from . import hdf5


# HostBatch(raw=...) specs of the stored sample types: (dtype, divisor, mul, sub) of __load_dd (:313-329), applied on the
# device by ugn_decode_samples -- optical flow int16 / compressFactor(100) * 0.1, gray / depth uint8 / 255 - 0.5,
# silhouette uint8 / 255.  Append (clip_min, clip_max) for the raw-magnitude clip of the optical flow (:318-321).
RAW_FLOW = ("int16", 100.0, 0.1, 0.0)
RAW_GRAY = ("uint8", 255.0, 1.0, 0.5)
RAW_SILHOUETTE = ("uint8", 255.0, 1.0, 0.0)


def load_sample(path) -> Dict[str, object]:
    f = hdf5.File(path)
    out: Dict[str, object] = {}
    for k, v in f.attrs.items():
        if not k.isupper() and k not in ("TITLE", "CLASS", "VERSION", "PYTABLES_FORMAT_VERSION", "DEEPDISH_IO_VERSION"):
            out[k] = v
    for k in f.keys():
        node = f[k]
        if not node.is_group:
            val = node.value
            out[k] = val[()] if getattr(val, "shape", None) == () else val
    return out


def save_sample(path, sample: Dict[str, object]):
    """Write a sample in the layout load_sample reads (arrays as datasets, scalars as 0-d datasets)."""
    w = hdf5.Writer()
    for k, v in sample.items():
        w.dataset(k, np.asarray(v))
    w.save(path)


def decode_sample(sample: Dict[str, object], silhouette: bool = False, ntype: int = 2, clip_max: float = 0,
                  clip_min: float = 0) -> Optional[np.ndarray]:
    """__load_dd (:294-338): compressFactor > 1 (optical flow): float32(data), optional magnitude clip on the RAW values
    (|x| > clip_max or < clip_min -> 1e-8), / compressFactor, * 0.1 when ntype == 2; otherwise uint8 / 255 (silhouette) or
    uint8 / 255 - 0.5 (gray, depth).  Returns the volume channels-first [C,H,W] (the generators' np.moveaxis(x, 2, 0)),
    float32; None for an empty sample."""
    data = np.asarray(sample["data"])
    if data.size == 0:
        return None
    cf = float(np.asarray(sample.get("compressFactor", 1)))
    if cf > 1:
        x = np.float32(data)
        if clip_max > 0:
            x[np.abs(x) > clip_max] = 1e-8
        if clip_min > 0:
            x[np.abs(x) < clip_min] = 1e-8
        x = x / np.float32(cf)
        if ntype == 2:
            x = x * np.float32(0.1)
    elif silhouette:
        x = np.float32(data) / np.float32(255.0)
    else:
        x = np.float32(data) / np.float32(255.0) - np.float32(0.5)
    return np.ascontiguousarray(np.moveaxis(x, 2, 0))
