"""Input pipeline after the decode: oracle and CUDA kernel against the outputs of the reference's own
load_rgb / load_sar / load_dsm / RandomCrop (tests/golden/raster.pt, made by tests/golden/make_golden_raster.py)."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from oracle import raster_pipeline as O


def _gen(golden_dir):
    spec = importlib.util.spec_from_file_location("make_golden_raster", os.path.join(golden_dir, "make_golden_raster.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)     # does not touch /root/reference unless load_reference() is called
    return mod


def _cases(golden_dir):
    gen = _gen(golden_dir)
    fx = torch.load(os.path.join(golden_dir, "raster.pt"), weights_only=False)
    for c in gen.CASES:
        rgb, sar, dsm = gen.raw_rasters(c["seed"], c["batch"], c["factor"], c["rgb"], c["dsm"])
        np.random.seed(c["seed"])
        crop = None
        if c["crop"]:
            top, left = O.draw_crops(c["batch"], (c["crop"], c["crop"]))
            crop = (top, left, (c["crop"], c["crop"]))
        yield c, {"rgb": rgb, "sar": sar, "dsm": dsm}, crop, fx[c["name"]]


def _check(name, key, got, want_fx, cropped, exact, tol):
    got = torch.as_tensor(np.asarray(got))
    want = want_fx[key]
    if not cropped:   # the fixture keeps every 4th row / column and the fp64 sum of the full output
        s = float(got.double().sum())
        assert abs(s - want_fx["sum"][key]) <= 1e-6 * max(1.0, float(got.double().abs().sum())), (name, key, "sum")
        got = got[..., ::4, ::4]
    assert got.shape == want.shape and got.dtype == torch.float32, (name, key, got.shape, want.shape)
    if exact:
        assert torch.equal(got, want), (name, key, float((got - want).abs().max()))
    else:
        err = float((got - want).abs().max())
        assert err <= tol, (name, key, err)


def test_oracle_matches_reference_golden(golden_dir):
    """the numpy restatement (incl. cv2's INTER_AREA summation order) against the reference's own functions: the
    constant z-score modalities bit-exact, the per-image standardisation and the dB transform to 1 ulp-level noise"""
    for c, raw, crop, fx in _cases(golden_dir):
        outs = {"s2": [], "s1": [], "dem": []}
        for b in range(c["batch"]):
            s2, s1, dem = O.load_rgb(raw["rgb"][b]), O.load_sar(raw["sar"][b]), O.load_dsm(raw["dsm"][b])
            if crop:
                s2, s1, dem = (O.crop(x, crop[0][b], crop[1][b], crop[2]) for x in (s2, s1, dem))
            outs["s2"].append(s2); outs["s1"].append(s1); outs["dem"].append(dem)
        for k in outs:
            _check(c["name"], k, np.stack(outs[k]), fx, crop is not None, exact=True, tol=0)


def test_random_crop_draws_follow_the_reference(golden_dir):
    """RandomCrop.draw = the reference's RNG calls (top then left, sample by sample)"""
    pytest.importorskip("ctypes")
    from incomplete_multimodal_fusion_b200.utils import multimodal_dfc2023 as D
    np.random.seed(5)
    t1, l1 = O.draw_crops(7, (224, 224))
    np.random.seed(5)
    t2, l2 = D.RandomCrop(224).draw(7)
    assert (t1 == t2).all() and (l1 == l2).all()
    assert (D.rgb_MEAN == O.rgb_MEAN).all() and (D.sar_STD == O.sar_STD).all()


@pytest.mark.gpu
def test_device_pipeline_matches_reference_golden(golden_dir):
    """mmf_raster_prep through the drop-in module: integer rasters with constant z-score bit-exact; float paths (log10,
    per-image statistics) within 1e-5 absolute of the reference's normalised values (|values| = O(1))"""
    from incomplete_multimodal_fusion_b200.utils import multimodal_dfc2023 as D
    for c, raw, crop, fx in _cases(golden_dir):
        sample = {k: torch.from_numpy(v).cuda() for k, v in raw.items()}
        out = D.prepare_rgb_sar_dsm(sample, crop=crop)
        torch.cuda.synchronize()
        rgb_exact = c["rgb"] != np.float32
        _check(c["name"], "s2", out["s2"].cpu(), fx, crop is not None, exact=rgb_exact, tol=0 if rgb_exact else 1e-6)
        _check(c["name"], "s1", out["s1"].cpu(), fx, crop is not None, exact=False, tol=1e-5)
        _check(c["name"], "dem", out["dem"].cpu(), fx, crop is not None, exact=False, tol=1e-5)


@pytest.mark.gpu
def test_device_pipeline_matches_oracle_full_size():
    """BASELINE cfg-2 shape: 512 x 512 rasters, batch 16, 224 crops -- the kernel against the oracle on every pixel"""
    from incomplete_multimodal_fusion_b200.utils import multimodal_dfc2023 as D
    g = np.random.default_rng(3)
    B = 16
    rgb = g.integers(0, 256, (B, 3, 512, 512)).astype(np.uint8)
    sar = (10.0 ** g.normal(-0.8, 0.4, (B, 1, 512, 512))).astype(np.float32)
    dsm = g.gamma(2.0, 4.0, (B, 1, 512, 512)).astype(np.float32)
    np.random.seed(9)
    top, left = D.RandomCrop(224).draw(B)
    crop = (top, left, (224, 224))
    out = D.prepare_rgb_sar_dsm({"rgb": torch.from_numpy(rgb).cuda(), "sar": torch.from_numpy(sar).cuda(),
                                 "dsm": torch.from_numpy(dsm).cuda()}, crop=crop)
    for b in range(B):
        want = {"s2": O.load_rgb(rgb[b]), "s1": O.load_sar(sar[b]), "dem": O.load_dsm(dsm[b])}
        for k, w in want.items():
            w = torch.from_numpy(np.ascontiguousarray(O.crop(w, top[b], left[b], (224, 224))))
            got = out[k][b].cpu()
            if k == "s2":
                assert torch.equal(got, w)
            else:
                assert float((got - w).abs().max()) <= 1e-5, (k, b)


@pytest.mark.gpu
def test_device_pipeline_rejects_what_it_does_not_restate():
    from incomplete_multimodal_fusion_b200.utils import multimodal_dfc2023 as D
    with pytest.raises(NotImplementedError):
        D.prepare_rgb(torch.zeros(1, 3, 300, 300, dtype=torch.uint8, device="cuda"))     # non-integer INTER_AREA factor
    with pytest.raises(NotImplementedError):
        D.prepare_sar(torch.zeros(1, 1, 256, 256, dtype=torch.uint8, device="cuda"))
    with pytest.raises(RuntimeError):
        D.prepare_rgb(torch.zeros(1, 3, 256, 256, dtype=torch.uint8))                    # no CPU fallback
    with pytest.raises(ValueError):
        D.prepare_rgb(torch.zeros(2, 3, 256, 256, dtype=torch.uint8, device="cuda"),
                      crop=(np.array([40, 0]), np.array([0, 0]), (224, 224)))            # window leaves the raster


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["normal", "ties", "constant", "negative"])
def test_standardize_depth_matches_reference_expression(kind):
    """the sort-free truncated standardisation against the reference's expression (pretrain_mmae.py:452-459) in fp64-free
    torch on the CPU: continuous values, heavy ties (boundary values repeated inside and outside the kept slice), a
    constant image (variance 0 -> division by sqrt(1e-6)) and negative values (key order of the radix select)"""
    from incomplete_multimodal_fusion_b200.utils import multimodal_dfc2023 as D
    g = torch.Generator().manual_seed(4)
    B, H = 6, 224
    if kind == "normal":
        dem = torch.randn(B, 1, H, H, generator=g) * 7 + 5
    elif kind == "ties":
        dem = torch.randint(0, 12, (B, 1, H, H), generator=g).float()
    elif kind == "constant":
        dem = torch.full((B, 1, H, H), 3.25)
    else:
        dem = -torch.rand(B, 1, H, H, generator=g) * 100 + 20
    want = O.standardize_depth(dem.double()).float() if kind != "constant" else O.standardize_depth(dem)
    got = D.standardize_depth(dem.cuda()).cpu()
    scale = float(want.abs().max().clamp_min(1.0))
    assert float((got - want).abs().max()) <= 2e-5 * scale, (kind, float((got - want).abs().max()))
    ref32 = O.standardize_depth(dem)            # the reference's own fp32 arithmetic sits at the same distance
    assert float((got - ref32).abs().max()) <= 2e-4 * scale


# ------------------------------------------------------------------------------------------------
# host half: decode -> Dataset -> device loader (multimodal_dfc2023.py:180-238, pretrain_mmae.py:317-323)
# ------------------------------------------------------------------------------------------------
def _write_dataset(root, n, size, seed=0):
    import cv2
    rng = np.random.default_rng(seed)
    raws = []
    for sub in ("rgb", "sar", "dsm"):
        os.makedirs(os.path.join(root, sub), exist_ok=True)
    for i in range(n):
        rgb = rng.integers(0, 256, (3, size, size), dtype=np.uint8)
        sar = (rng.random((1, size, size), dtype=np.float32) * 2 + 1e-3).astype(np.float32)
        dsm = (rng.random((1, size, size), dtype=np.float32) * 40).astype(np.float32)
        name = "tile_%03d.tiff" % i
        cv2.imwrite(os.path.join(root, "rgb", name), np.ascontiguousarray(rgb.transpose(1, 2, 0)[:, :, ::-1]))   # cv2 writes BGR
        cv2.imwrite(os.path.join(root, "sar", name), sar[0])
        cv2.imwrite(os.path.join(root, "dsm", name), dsm[0])
        raws.append({"id": name, "rgb": rgb, "sar": sar, "dsm": dsm})
    return {r["id"]: r for r in raws}


def test_dataset_decodes_raw_rasters(tmp_path):
    """DFC2023.__getitem__ returns the files' own values (uint8 optical bands in file order, float32 SAR / DSM) and the crop
    origins the reference's RandomCrop would draw, in its order"""
    from incomplete_multimodal_fusion_b200.utils import multimodal_dfc2023 as D
    raws = _write_dataset(str(tmp_path), 3, 64)
    ds = D.DFC2023(str(tmp_path), transform=True, crop_size=224)
    assert len(ds) == 3
    np.random.seed(5)
    got = [ds[i] for i in range(3)]
    np.random.seed(5)
    for s in got:
        r = raws[s["id"]]
        assert s["rgb"].dtype == np.uint8 and np.array_equal(s["rgb"], r["rgb"])
        assert s["sar"].dtype == np.float32 and np.array_equal(s["sar"], r["sar"])
        assert s["dsm"].dtype == np.float32 and np.array_equal(s["dsm"], r["dsm"])
        top, left = np.random.randint(0, 256 - 224), np.random.randint(0, 256 - 224)     # RandomCrop.__call__ :66-73
        assert tuple(s["crop"]) == (top, left)


@pytest.mark.gpu
def test_device_loader_matches_reference_pipeline(tmp_path):
    """decode -> pinned batch -> copy stream -> mmf_raster_prep, against the (golden-pinned) restatement of the reference's
    load_rgb / load_sar / load_dsm + RandomCrop on the same files"""
    from incomplete_multimodal_fusion_b200.utils import multimodal_dfc2023 as D
    size, crop = 512, 224
    raws = _write_dataset(str(tmp_path), 5, size, seed=3)
    ds = D.DFC2023(str(tmp_path), transform=True, crop_size=crop)
    loader = D.DeviceBatchLoader(ds, batch_size=2, drop_last=True)
    np.random.seed(11)
    batches = list(loader)
    assert len(batches) == 2
    np.random.seed(11)
    for b in batches:
        assert set(b) == {"s1", "s2", "dem", "id"} and b["s2"].shape == (2, 3, crop, crop) and b["s2"].dtype == torch.float32
        for i, name in enumerate(b["id"]):
            r = raws[name]
            top, left = np.random.randint(0, 256 - crop), np.random.randint(0, 256 - crop)
            want = {"s2": O.crop(O.load_rgb(r["rgb"]), top, left, (crop, crop)), "s1": O.crop(O.load_sar(r["sar"]), top, left, (crop, crop)),
                    "dem": O.crop(O.load_dsm(r["dsm"]), top, left, (crop, crop))}
            assert torch.equal(b["s2"][i].cpu(), torch.as_tensor(want["s2"]))                       # integer raster: bit-exact
            assert float((b["s1"][i].cpu() - torch.as_tensor(want["s1"])).abs().max()) <= 2e-5
            assert float((b["dem"][i].cpu() - torch.as_tensor(want["dem"])).abs().max()) <= 2e-5
