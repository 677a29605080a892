"""Device half of the DFC2023 input pipeline (reference: pretraining/utils/multimodal_dfc2023.py).

The reference does, per sample and on the host (``num_workers=0``): rasterio decode -> per-pixel transform ->
``cv2.resize(.., (256, 256), INTER_AREA)`` -> float32 -> z-score (``load_rgb`` :115-125, ``load_sar`` :128-139,
``load_dsm`` :99-112), optionally ``RandomCrop`` (:53-94), then the default collate and a pinned H2D copy of fp32
tensors (``pretrain_mmae.py:317-323, 447-450``).  Here the host only decodes and stacks the RAW rasters (uint8 / uint16 /
float32 at native size, 1-2 bytes per value for the optical bands); everything after the decode is one CUDA launch per
modality (``mmf_raster_prep``) producing the ``{'s1', 's2', 'dem'}`` fp32 tensors ``model(input_dict, ...)`` consumes.
The GeoTIFF decode itself (rasterio) stays out of scope.

Same names as the reference module where the meaning is the same (``rgb_MEAN`` .. ``dem_STD``, ``RandomCrop``,
``load_rgb_sar_dsm`` -> ``prepare_rgb_sar_dsm`` taking decoded arrays instead of paths).  No CPU path: tensors must
be CUDA tensors.
"""
import numpy as np
import torch

from .. import kernels as K

# per-band statistics of the reference (multimodal_dfc2023.py:27-50)
rgb_MEAN = np.array([81.29692, 87.93711, 72.041306])
rgb_STD = np.array([39.61512, 35.407978, 35.84708])
sar_MEAN = np.array([-7.9447875, ])
sar_STD = np.array([2.777256, ])
dem_MEAN = np.array([5.0160093, ])     # unused by load_dsm (per-image standardisation), kept for parity of the namespace
dem_STD = np.array([7.6128364, ])

RESIZE = (256, 256)   # the reference's fixed cv2.resize target


def _factor(raw, size):
    Hs, Ws = raw.shape[-2:]
    if Hs % size[0] or Ws % size[1] or Hs // size[0] != Ws // size[1]:
        raise NotImplementedError("raster %dx%d -> %s: only integer INTER_AREA factors are restated on the device" % (Hs, Ws, size))
    return Hs // size[0]


class RandomCrop(object):
    """The reference's RandomCrop (multimodal_dfc2023.py:53-94) as crop ORIGINS: the same two ``np.random.randint`` draws
    per sample, in the same order (top, then left); the window is applied by the device kernel to all modalities."""

    def __init__(self, output_size):
        assert isinstance(output_size, (int, tuple))
        self.output_size = (output_size, output_size) if isinstance(output_size, int) else output_size
        assert len(self.output_size) == 2

    def draw(self, batch, h=RESIZE[0], w=RESIZE[1]):
        new_h, new_w = self.output_size
        top, left = np.empty(batch, np.int32), np.empty(batch, np.int32)
        for b in range(batch):
            top[b] = np.random.randint(0, h - new_h)
            left[b] = np.random.randint(0, w - new_w)
        return top, left


def _crop_args(crop, batch, size, device):
    if crop is None:
        return None, None, size
    top, left, hw = crop
    if torch.is_tensor(top) and top.is_cuda:      # origins already uploaded (``upload_crop``): validated there, no copy here
        return top, left, tuple(hw)
    top, left = np.asarray(top, np.int32), np.asarray(left, np.int32)
    if top.shape != (batch,) or left.shape != (batch,):
        raise ValueError("crop origins must be [batch]")
    if top.min() < 0 or left.min() < 0 or top.max() + hw[0] > size[0] or left.max() + hw[1] > size[1]:
        raise ValueError("crop window leaves the resized raster")
    return torch.from_numpy(top).to(device), torch.from_numpy(left).to(device), tuple(hw)


def upload_crop(crop, batch, device, size=RESIZE):
    """validate the crop origins once and move them to the device: the same window then serves all modalities of the batch
    without a host -> device copy per launch"""
    return _crop_args(crop, batch, size, device)


def _prepare(raw, mode, mean, std, crop, size):
    if not raw.is_cuda:
        raise RuntimeError("prepare_*: CUDA tensors only (no CPU fallback)")
    raw = raw.contiguous()
    top, left, hw = _crop_args(crop, raw.shape[0], size, raw.device)
    return K.raster_prep(raw, mode, _factor(raw, size), mean, std, top, left, hw)


def prepare_rgb(raw, crop=None, size=RESIZE):
    """load_rgb (:115-125) after the decode: raw [B, 3, Hs, Ws] -> fp32 [B, 3, h, w]"""
    return _prepare(raw, K.RASTER_ZSCORE, rgb_MEAN, rgb_STD, crop, size)


def prepare_sar(raw, crop=None, size=RESIZE):
    """load_sar (:128-139) after the decode: raw float32 [B, 1, Hs, Ws] linear backscatter -> dB, clip, z-score"""
    if raw.dtype != torch.float32:
        raise NotImplementedError("prepare_sar: float32 rasters only")
    return _prepare(raw, K.RASTER_SAR_DB, sar_MEAN, sar_STD, crop, size)


def prepare_dsm(raw, crop=None, size=RESIZE):
    """load_dsm (:99-112) after the decode: raw [B, 1, Hs, Ws] -> per-image standardised fp32"""
    return _prepare(raw, K.RASTER_STANDARDIZE, None, None, crop, size)


def prepare_rgb_sar_dsm(sample, use_rgb=True, use_sar=True, use_dsm=True, crop=None):
    """load_rgb_sar_dsm (:151-177) + RandomCrop + collate on decoded batches: ``sample`` maps 'rgb' / 'sar' / 'dsm' to raw
    CUDA tensors; ``crop`` = (top[B], left[B], (h, w)) from ``RandomCrop.draw`` or None.  Returns the reference's keys."""
    if crop is not None:
        first = next(v for v in sample.values() if v is not None)
        crop = upload_crop(crop, first.shape[0], first.device)
    return {
        's1': prepare_sar(sample["sar"], crop) if use_sar else None,
        's2': prepare_rgb(sample["rgb"], crop) if use_rgb else None,
        'dem': prepare_dsm(sample["dsm"], crop) if use_dsm else None,
    }


def standardize_depth(dem):
    """The training loop's truncated depth standardisation (pretrain_mmae.py:452-459, ``--standardize_depth``): per sample,
    mean and variance of the values left after dropping the bottom and top 10 %, applied to the whole image.  One launch
    (radix select in shared memory) instead of the reference's full ``torch.sort`` of every sample."""
    if not dem.is_cuda:
        raise RuntimeError("standardize_depth: CUDA tensors only (no CPU fallback)")
    return K.trunc_standardize(dem.float().contiguous())
