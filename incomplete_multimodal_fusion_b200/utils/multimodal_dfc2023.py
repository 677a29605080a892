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


# ------------------------------------------------------------------------------------------------
# Host half: decode -> Dataset -> loader.  Reference: DFC2023 (multimodal_dfc2023.py:180-238) and the
# DataLoader of pretrain_mmae.py:317-323.  The reference's __getitem__ decodes AND transforms each sample on the host
# (5.4 ms of numpy / cv2 per sample); here __getitem__ only decodes and returns the RAW rasters (uint8 optical bands,
# float32 backscatter / heights, at native size) plus the crop origin drawn by the reference's RandomCrop calls; the
# loader stacks them into pinned host buffers, copies them to the GPU on a copy stream while the previous batch trains,
# and one mmf_raster_prep launch per modality produces the fp32 {'s1', 's2', 'dem'} tensors the model reads.
# ------------------------------------------------------------------------------------------------
import glob
import os


def decode_tiff(path, first_band_only=False):
    """(Geo)TIFF -> numpy [C, H, W] in the file's own dtype (what `rasterio.open(path).read()` returns).  rasterio is used
    when it is installed; otherwise OpenCV's libtiff reader (uint8 / uint16 / float32, strips or tiles, LZW / deflate),
    whose 3- and 4-channel results come back in BGR(A) order and are put back into file order."""
    try:
        import rasterio   # noqa: F401  (not in this image; the reference's decoder)
        with rasterio.open(path) as data:
            return data.read(1)[None] if first_band_only else data.read()
    except ImportError:
        pass
    import cv2
    img = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    if img is None:
        raise IOError("cannot decode %s" % path)
    if img.ndim == 2:
        img = img[:, :, None]
    elif img.shape[2] == 3:
        img = img[:, :, ::-1]
    elif img.shape[2] == 4:
        img = img[:, :, [2, 1, 0, 3]]
    img = np.ascontiguousarray(img.transpose(2, 0, 1))
    return img[:1] if first_band_only else img


class DFC2023(torch.utils.data.Dataset):
    """The reference's dataset class (same constructor, same directory layout <path>/{rgb,sar,dsm,lc}/*.tiff, same sample
    list order) returning RAW decoded rasters: {'rgb': uint8 [3, H, W], 'sar': float32 [1, H, W], 'dsm': float32
    [1, H, W], 'id': name[, 'label': [H, W]][, 'crop': int32 [2] = (top, left)]}.  With transform=True the crop origin is
    drawn here, per sample and in the reference's order (RandomCrop.__call__: top, then left, :66-73), so a single-process
    loader consumes numpy's global RNG exactly like the reference's; the window itself is cut on the device."""

    def __init__(self, path, use_rgb=True, use_sar=True, use_dsm=True, unlabeled=True, transform=False, crop_size=32):
        super().__init__()
        self.use_rgb, self.use_sar, self.use_dsm, self.unlabeled = use_rgb, use_sar, use_dsm, unlabeled
        self.transform = RandomCrop(crop_size) if transform else None
        assert os.path.exists(path)
        self.samples = []
        for rgb_loc in glob.glob(os.path.join(path, "rgb/*.tiff"), recursive=True):
            s = {"rgb": rgb_loc, "sar": rgb_loc.replace("rgb", "sar"), "dsm": rgb_loc.replace("rgb", "dsm"), "id": os.path.basename(rgb_loc)}
            if not unlabeled:
                s["lc"] = rgb_loc.replace("rgb", "lc")
            self.samples.append(s)

    def __len__(self):
        return len(self.samples)

    def __getitem__(self, index):
        s = self.samples[index]
        out = {"id": s["id"]}
        if self.use_rgb:
            out["rgb"] = decode_tiff(s["rgb"])
        if self.use_sar:
            out["sar"] = decode_tiff(s["sar"]).astype(np.float32, copy=False)
        if self.use_dsm:
            out["dsm"] = decode_tiff(s["dsm"], first_band_only=True)
        if not self.unlabeled:
            out["label"] = decode_tiff(s["lc"], first_band_only=True)[0]
        if self.transform is not None:
            top, left = self.transform.draw(1)
            out["crop"] = np.array([top[0], left[0]], np.int32)
        return out


class DeviceBatchLoader:
    """pretrain_mmae.py:317-323 + :447-450 for raw batches: wraps a torch DataLoader over `DFC2023` (default collate, pinned
    memory) and yields the reference's per-step dict {'s1', 's2', 'dem'[, 'label'], 'id'} of normalised fp32 CUDA tensors.
    Batch i + 1 is copied host -> device on a copy stream into the other of two device buffer sets while batch i is in
    use; the three mmf_raster_prep launches run on the consumer's stream."""

    def __init__(self, dataset, batch_size, device="cuda", sampler=None, shuffle=False, num_workers=0, drop_last=True,
                 use_rgb=True, use_sar=True, use_dsm=True):
        self.dataset, self.device = dataset, torch.device(device)
        self.use = dict(rgb=use_rgb, sar=use_sar, dsm=use_dsm)
        self.crop_hw = dataset.transform.output_size if getattr(dataset, "transform", None) is not None else None
        self.loader = torch.utils.data.DataLoader(dataset, batch_size=batch_size, sampler=sampler, shuffle=shuffle and sampler is None,
                                                  num_workers=num_workers, pin_memory=True, drop_last=drop_last)
        self.copy_stream = torch.cuda.Stream(device=self.device)

    def __len__(self):
        return len(self.loader)

    def _upload(self, batch):
        """pinned host batch -> device, asynchronously on the copy stream; returns (device tensors, ready event)"""
        dev = {}
        with torch.cuda.stream(self.copy_stream):
            for k in ("rgb", "sar", "dsm", "label", "crop"):
                if k in batch:
                    dev[k] = batch[k].to(self.device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        dev["id"] = batch["id"]
        return dev, ev

    def _prepare(self, dev):
        B = next(v for k, v in dev.items() if k in ("rgb", "sar", "dsm")).shape[0]
        crop = None
        if "crop" in dev and self.crop_hw is not None:
            c = dev["crop"].to(torch.int32)
            crop = (c[:, 0].contiguous(), c[:, 1].contiguous(), self.crop_hw)
        out = prepare_rgb_sar_dsm({"rgb": dev.get("rgb"), "sar": dev.get("sar"), "dsm": dev.get("dsm")},
                                  use_rgb="rgb" in dev, use_sar="sar" in dev, use_dsm="dsm" in dev, crop=crop)
        out = {k: v for k, v in out.items() if v is not None}
        if "label" in dev:
            out["label"] = dev["label"]
        out["id"] = dev["id"]
        assert all(v.shape[0] == B for k, v in out.items() if torch.is_tensor(v))
        return out

    def __iter__(self):
        it = iter(self.loader)
        try:
            nxt = self._upload(next(it))
        except StopIteration:
            return
        while nxt is not None:
            dev, ev = nxt
            try:
                nxt = self._upload(next(it))       # the next batch's copy overlaps this batch's use
            except StopIteration:
                nxt = None
            torch.cuda.current_stream(self.device).wait_event(ev)
            for v in dev.values():
                if torch.is_tensor(v):
                    v.record_stream(torch.cuda.current_stream(self.device))
            yield self._prepare(dev)
