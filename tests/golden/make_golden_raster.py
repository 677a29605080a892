#!/usr/bin/env python
"""Golden fixture of the input pipeline: runs the REFERENCE's own load_rgb / load_sar / load_dsm / RandomCrop
(pretraining/utils/multimodal_dfc2023.py, unmodified, real cv2) on synthetic raw rasters and stores raw inputs (seeds only)
plus the reference's outputs in tests/golden/raster.pt.

Authoring container only (reads /root/reference).  ``rasterio`` is not installed: the module's ``rasterio.open`` is served
by an in-memory stub that hands back the synthetic array registered under the "path" -- the decode is out of scope, every
line after it is the reference's.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/pretraining/utils/multimodal_dfc2023.py"

_RASTERS = {}


class _Handle:
    def __init__(self, arr):
        self.arr = arr

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def read(self, band=None):
        return self.arr.copy() if band is None else self.arr[band - 1].copy()


def load_reference():
    stub = types.ModuleType("rasterio")
    stub.open = lambda path: _Handle(_RASTERS[path])
    sys.modules["rasterio"] = stub
    spec = importlib.util.spec_from_file_location("ref_dfc2023", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def raw_rasters(seed, batch, factor, rgb_dtype, dsm_dtype):
    """synthetic decoded rasters; specials (NaN, inf, zero / negative backscatter) planted at fixed positions"""
    g = np.random.default_rng(seed)
    n = 256 * factor
    hi = 256 if rgb_dtype == np.uint8 else 4096
    rgb = g.integers(0, hi, (batch, 3, n, n)).astype(rgb_dtype) if rgb_dtype != np.float32 else \
        (g.random((batch, 3, n, n)) * 255).astype(np.float32)
    sar = (10.0 ** (g.normal(-0.8, 0.4, (batch, 1, n, n)))).astype(np.float32)
    sar[:, 0, 3, 5] = 0.0
    sar[:, 0, 7, 9] = -1.0
    sar[:, 0, 11, 2] = np.nan
    sar[:, 0, 13, 4] = np.inf
    if dsm_dtype == np.float32:
        dsm = (g.gamma(2.0, 4.0, (batch, 1, n, n))).astype(np.float32)
        dsm[:, 0, 5, 5] = np.nan
    else:
        dsm = g.integers(0, 60, (batch, 1, n, n)).astype(dsm_dtype)
    if rgb_dtype == np.float32:
        rgb[:, 1, 2, 2] = np.nan
    return rgb, sar, dsm


CASES = [
    dict(name="f2_u8", seed=11, batch=2, factor=2, rgb=np.uint8, dsm=np.float32, crop=80),
    dict(name="f1_u8", seed=12, batch=2, factor=1, rgb=np.uint8, dsm=np.float32, crop=None),
    dict(name="f4_u16", seed=13, batch=2, factor=4, rgb=np.uint16, dsm=np.uint16, crop=64),
    dict(name="f3_f32", seed=14, batch=2, factor=3, rgb=np.float32, dsm=np.uint8, crop=64),
    dict(name="f2_u16", seed=15, batch=2, factor=2, rgb=np.uint16, dsm=np.uint8, crop=None),
]


def main():
    ref = load_reference()
    out = {}
    for c in CASES:
        rgb, sar, dsm = raw_rasters(c["seed"], c["batch"], c["factor"], c["rgb"], c["dsm"])
        np.random.seed(c["seed"])
        tf = ref.RandomCrop(c["crop"]) if c["crop"] else None
        res = {"s1": [], "s2": [], "dem": []}
        for b in range(c["batch"]):
            _RASTERS.update(rgb=rgb[b], sar=sar[b], dsm=dsm[b])
            s = ref.load_rgb_sar_dsm({"rgb": "rgb", "sar": "sar", "dsm": "dsm", "id": b}, True, True, True, unlabeled=True)
            if tf:
                s = tf(s)
            for k in res:
                res[k].append(torch.from_numpy(np.ascontiguousarray(s[k])))
        # uncropped cases keep every 4th row / column of the 256 x 256 output plus its fp64 sum (fixture size)
        full = {k: torch.stack(v) for k, v in res.items()}
        out[c["name"]] = {k: (v if tf else v[..., ::4, ::4].contiguous()) for k, v in full.items()}
        out[c["name"]]["sum"] = {k: float(v.double().sum()) for k, v in full.items()}
        out[c["name"]]["dtypes"] = {k: str(v[0].dtype) for k, v in res.items()}
        print(c["name"], {k: (tuple(v.shape), v.dtype) for k, v in out[c["name"]].items() if torch.is_tensor(v)})
    torch.save(out, os.path.join(HERE, "raster.pt"))


if __name__ == "__main__":
    main()
