"""TEST INFRASTRUCTURE (oracle): numpy restatement of the reference's per-sample raster transforms after the decode
(pretraining/utils/multimodal_dfc2023.py).  Only tests/ may import this.  No cv2 here: ``area_resize`` restates
cv2.resize(INTER_AREA) for integer factors (summation order probed against cv2 4.13 and pinned bit-exact by
tests/golden/raster.pt, which holds the outputs of the reference's own load_* functions).

    resiz_4pl        :10-16   per band cv2.resize into a float64 array
    normalize_rgb/sar:27-40   (float32 - float64) / float64, rounded back into the float32 array
    load_dsm         :99-112  nan_to_num, resize, float32, (x - mean) / sqrt(var + 1e-6)
    load_rgb         :115-125 nan_to_num, resize, float32, normalize_rgb
    load_sar         :128-139 10 log10(x + 1e-7), clip [-25, 0], nan_to_num, resize, float32, normalize_sar
    RandomCrop       :53-94   top = randint(0, h - new_h), left = randint(0, w - new_w), same window for all modalities
"""
import numpy as np

rgb_MEAN = np.array([81.29692, 87.93711, 72.041306])
rgb_STD = np.array([39.61512, 35.407978, 35.84708])
sar_MEAN = np.array([-7.9447875, ])
sar_STD = np.array([2.777256, ])


def area_resize(band, size):
    """cv2.resize(band, size, interpolation=cv2.INTER_AREA) for an integer shrink factor, same dtype as the input"""
    H, W = band.shape
    if H % size[1] or W % size[0] or H // size[1] != W // size[0]:
        raise NotImplementedError("integer factors only")
    f = H // size[1]
    if f == 1:
        return band.copy()
    blk = band.reshape(size[1], f, size[0], f).transpose(0, 2, 1, 3).reshape(size[1], size[0], f * f)   # row-major f x f
    if np.issubdtype(band.dtype, np.integer):
        s = blk.astype(np.int64).sum(-1)
        if f == 2:
            return ((s + 2) >> 2).astype(band.dtype)
        r = np.rint(s.astype(np.float32) * np.float32(1.0 / (f * f)))
        return np.clip(r, 0, np.iinfo(band.dtype).max).astype(band.dtype)
    assert band.dtype == np.float32
    if f == 2:
        return ((blk[..., 0] + blk[..., 1]) + (blk[..., 2] + blk[..., 3])) * np.float32(0.25)
    acc = np.zeros(blk.shape[:2], np.float32)
    k = 0
    while k + 4 <= f * f:
        acc = acc + (((blk[..., k] + blk[..., k + 1]) + blk[..., k + 2]) + blk[..., k + 3])
        k += 4
    while k < f * f:
        acc = acc + blk[..., k]
        k += 1
    return acc * np.float32(1.0 / (f * f))


def resiz_4pl(img, size):
    out = np.zeros((img.shape[0], size[0], size[1]))
    for i in range(img.shape[0]):
        out[i] = area_resize(img[i], size)
    return out


def _zscore(imgs, mean, std):
    for i in range(imgs.shape[0]):
        imgs[i] = (imgs[i] - mean[i]) / std[i]
    return imgs


def load_rgb(raw, size=(256, 256)):
    return _zscore(resiz_4pl(np.nan_to_num(raw), size).astype(np.float32), rgb_MEAN, rgb_STD)


def load_sar(raw, size=(256, 256)):
    with np.errstate(all="ignore"):
        sar = 10 * np.log10(raw + 0.0000001)
    sar = np.nan_to_num(np.clip(sar, -25, 0))
    return _zscore(resiz_4pl(sar, size).astype(np.float32), sar_MEAN, sar_STD)


def load_dsm(raw, size=(256, 256)):
    dsm = resiz_4pl(np.nan_to_num(raw), size).astype(np.float32)
    return (dsm - dsm.mean()) / np.sqrt(dsm.var() + 1e-6)


def crop(img, top, left, hw):
    return img[:, top: top + hw[0], left: left + hw[1]]


def draw_crops(batch, hw, h=256, w=256):
    """the RNG calls of RandomCrop.__call__, one sample after the other"""
    top, left = np.empty(batch, np.int32), np.empty(batch, np.int32)
    for b in range(batch):
        top[b] = np.random.randint(0, h - hw[0])
        left[b] = np.random.randint(0, w - hw[1])
    return top, left


def standardize_depth(dem):
    """pretrain_mmae.py:452-459 verbatim in meaning (torch; ``rearrange(.., 'b c h w -> b (c h w)')`` = flatten(1))"""
    import torch
    trunc = torch.sort(dem.flatten(1), dim=1)[0]
    trunc = trunc[:, int(0.1 * trunc.shape[1]): int(0.9 * trunc.shape[1])]
    return (dem - trunc.mean(dim=1)[:, None, None, None]) / torch.sqrt(trunc.var(dim=1)[:, None, None, None] + 1e-6)
