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

    def _norm_pix_target(self, target):
        """criterion.py:90-96 / :147-153: the target standardised per patch over its (p1 p2 c) values (unbiased variance,
        eps 1e-6).  A preprocessing of the detached target (no gradient flows into it), done as view reductions -- the
        patchify / unpatchify rearranges of the reference are pure index maps."""
        B, C, H, W = target.shape
        p = self.scale_factor
        t = target.float().view(B, C, H // p, p, W // p, p)
        mean = t.mean(dim=(1, 3, 5), keepdim=True)
        var = t.var(dim=(1, 3, 5), keepdim=True, unbiased=True)
        return ((t - mean) / torch.sqrt(var + 1e-6)).view(B, C, H, W)

    def forward(self, input, target, mask=None):
        """input [B, C, H, W] (bf16 or fp32), target fp32, mask [B, n_patches] (1 = masked) or None -> scalar"""
        if self.norm_pix:
            target = self._norm_pix_target(target.detach())
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
    """Debiased hard-negative contrastive loss (criterion.py:214-268), fused: normalisation, the [2B, 2B] similarity (bf16
    operands, fp32 accumulation: torch.mm under the reference's autocast, Appendix A #19), the negative mask of
    get_negative_mask (:224-231, a B-iteration Python loop + .cuda() in the reference), the debiased re-weighting, the
    loss AND both input gradients come out of one call of three small launches (mmf_hardneg_loss); no autograd graph of
    elementwise ops, no host synchronisation."""

    def __init__(self, tau_plus=0.1, beta=1.0, temperature=0.5, alpha=256, estimator='hard'):
        super().__init__()
        self.tau_plus, self.beta, self.temperature, self.alpha, self.estimator = tau_plus, beta, temperature, alpha, estimator

    def get_negative_mask(self, batch_size, device=None):
        """[2B, 2B] bool, False on the diagonal and on the positive pair (criterion.py:224-231 without the loop); kept for
        API parity -- the kernel applies the same exclusion by index"""
        eye = torch.eye(batch_size, dtype=torch.bool, device=device)
        return ~torch.cat([torch.cat([eye, eye], 1), torch.cat([eye, eye], 1)], 0)

    def forward(self, out_1, out_2):
        if not out_1.is_cuda:
            raise RuntimeError("HardNegtive_loss runs on CUDA only (no CPU fallback)")
        if self.estimator not in ('hard', 'easy'):
            raise Exception('Invalid estimator selected. Please use any of [hard, easy]')
        return Fn.HardNegLossFn.apply(out_1, out_2, self.tau_plus, self.beta, self.temperature, self.estimator == 'easy')


def dino_loss_func(student_output, teacher_output, teacher_temp=0.04, student_temp=0.1):
    """criterion.py:328-335 -- one fused launch (forward + student gradient), fp32; the teacher is detached"""
    if not student_output.is_cuda:
        raise RuntimeError("dino_loss_func runs on CUDA only (no CPU fallback)")
    return Fn.DinoLossFn.apply(student_output, teacher_output, student_temp, teacher_temp)
