"""Losses with the reference's names and signatures (reference: pretraining/multimae/criterion.py)."""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import functions as Fn


class _MaskedReconLoss(nn.Module):
    KIND = 0

    def __init__(self, patch_size: int = 16, stride: int = 1, norm_pix=False):
        super().__init__()
        self.patch_size = patch_size
        self.stride = stride
        self.scale_factor = patch_size // stride
        self.norm_pix = norm_pix
        if norm_pix:
            raise NotImplementedError("norm_pix targets are not built (disabled in every reference script)")

    def forward(self, input, target, mask=None):
        """input [B, C, H, W] (bf16 or fp32), target fp32, mask [B, n_patches] (1 = masked) or None -> scalar"""
        if mask is not None:
            mask = mask.to(torch.int64)
            if mask.dim() == 2 and mask.shape[0] == 1 and input.shape[0] > 1:
                mask = mask.expand(input.shape[0], -1)
            mask = mask.contiguous()
        return Fn.MaskedLossFn.apply(input, target.float(), mask, self.scale_factor, self.KIND)


class MaskedMSELoss(_MaskedReconLoss):
    """criterion.py:61-115: mean over channels, patch-mask weighted, per-sample normalised, batch nanmean"""
    KIND = 0


class MaskedL1Loss(_MaskedReconLoss):
    """criterion.py:118-172"""
    KIND = 1


class MaskedCrossEntropyLoss(nn.Module):
    """criterion.py:24-58: cross-entropy over class maps with the patch mask (the loss of the semantic modality `dnw`,
    pretrain_mmae_my.py:68-75).  input [B, C, H, W] logits (bf16 or fp32), target [B, H, W] int64, mask [B, n_patches]."""

    def __init__(self, patch_size: int = 16, stride: int = 1, label_smoothing: float = 0.0):
        super().__init__()
        self.patch_size = patch_size
        self.stride = stride
        self.scale_factor = patch_size // stride
        self.label_smoothing = label_smoothing
        if label_smoothing != 0.0:
            raise NotImplementedError("label smoothing is not built (0.0 in the reference's script, pretrain_mmae_my.py:74)")

    def forward(self, input, target, mask=None):
        if not input.is_cuda:
            raise RuntimeError("MaskedCrossEntropyLoss runs on CUDA only (no CPU fallback)")
        if mask is not None:
            mask = mask.to(torch.int64)
            if mask.dim() == 2 and mask.shape[0] == 1 and input.shape[0] > 1:
                mask = mask.expand(input.shape[0], -1)
            mask = mask.contiguous()
        return Fn.MaskedCEFn.apply(input, target.to(torch.int64), mask, self.scale_factor)


class HardNegtive_loss(nn.Module):
    """Debiased hard-negative contrastive loss (criterion.py:214-268).  The [2B, D] x [D, 2B] similarity runs on the
    tcgen05 GEMM with bf16 operands (torch.mm under the reference's autocast, Appendix A #19); normalisation, exp / log
    and the row reductions are fp32 device ops (no Python loop over the batch, no hard-coded .cuda())."""

    def __init__(self, tau_plus=0.1, beta=1.0, temperature=0.5, alpha=256, estimator='hard'):
        super().__init__()
        self.tau_plus, self.beta, self.temperature, self.alpha, self.estimator = tau_plus, beta, temperature, alpha, estimator

    def get_negative_mask(self, batch_size, device=None):
        """[2B, 2B] bool, False on the diagonal and on the positive pair (criterion.py:224-231 without the loop)"""
        eye = torch.eye(batch_size, dtype=torch.bool, device=device)
        return ~torch.cat([torch.cat([eye, eye], 1), torch.cat([eye, eye], 1)], 0)

    def _negative_index(self, batch_size, device):
        """[2B, 2B - 2] int64: column of the j-th True of get_negative_mask's row r (what masked_select walks)"""
        cache = self.__dict__.setdefault('_neg_index_cache', {})
        key = (batch_size, str(device))
        if key not in cache:
            j = torch.arange(2 * batch_size - 2, device=device)[None, :]
            b = (torch.arange(2 * batch_size, device=device) % batch_size)[:, None]
            cache[key] = j + (j >= b).long() + (j >= b + batch_size - 1).long()
        return cache[key]

    def forward(self, out_1, out_2):
        B = out_1.shape[0]
        if not out_1.is_cuda:
            raise RuntimeError("HardNegtive_loss runs on CUDA only (no CPU fallback)")
        o1 = F.normalize(out_1.float(), dim=1)
        o2 = F.normalize(out_2.float(), dim=1)
        out = torch.cat([o1, o2], dim=0)
        neg = torch.exp(Fn.MatmulNTFn.apply(out, out) / self.temperature)
        # the negatives of row r are all columns but r mod B and (r mod B) + B, in ascending order: a gather through a static
        # index instead of masked_select (whose output size is read back: a host synchronisation per loss call)
        neg = torch.gather(neg, 1, self._negative_index(B, out.device))
        pos = torch.exp((o1 * o2).sum(-1) / self.temperature)
        pos = torch.cat([pos, pos], 0)
        if self.estimator == 'hard':
            N = 2 * B - 2
            imp = (self.beta * neg.log()).exp()
            reweight = (imp * neg).sum(-1) / imp.mean(-1)
            Ng = (-self.tau_plus * N * pos + reweight) / (1 - self.tau_plus)
            Ng = torch.clamp(Ng, min=N * math.e ** (-1 / self.temperature))
        elif self.estimator == 'easy':
            Ng = neg.sum(-1)
        else:
            raise Exception('Invalid estimator selected. Please use any of [hard, easy]')
        return (-torch.log(pos / (pos + Ng))).mean()


def dino_loss_func(student_output, teacher_output, teacher_temp=0.04, student_temp=0.1):
    """criterion.py:328-335 -- one fused launch (forward + student gradient), fp32; the teacher is detached"""
    if not student_output.is_cuda:
        raise RuntimeError("dino_loss_func runs on CUDA only (no CPU fallback)")
    return Fn.DinoLossFn.apply(student_output, teacher_output, student_temp, teacher_temp)
