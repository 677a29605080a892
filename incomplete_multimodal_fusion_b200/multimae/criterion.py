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


class HardNegtive_loss(nn.Module):
    """Debiased hard-negative contrastive loss (criterion.py:214-268).  Tiny [2B, 2B] problem: expressed with
    torch ops on the device (no Python loop over the batch, no hard-coded .cuda())."""

    def __init__(self, tau_plus=0.1, beta=1.0, temperature=0.5, alpha=256, estimator='hard'):
        super().__init__()
        self.tau_plus, self.beta, self.temperature, self.alpha, self.estimator = tau_plus, beta, temperature, alpha, estimator

    def forward(self, out_1, out_2):
        B = out_1.shape[0]
        o1 = F.normalize(out_1.float(), dim=1)
        o2 = F.normalize(out_2.float(), dim=1)
        out = torch.cat([o1, o2], dim=0)
        neg = torch.exp(out @ out.t() / self.temperature)
        eye = torch.eye(B, dtype=torch.bool, device=out.device)
        keep = ~torch.cat([torch.cat([eye, eye], 1), torch.cat([eye, eye], 1)], 0)
        neg = neg.masked_select(keep).view(2 * B, -1)
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
    """criterion.py:328-335 -- [B, D] problem, fp32"""
    s = F.log_softmax(F.normalize(student_output.float(), dim=1) / student_temp, dim=-1)
    t = F.softmax(F.normalize(teacher_output.float(), dim=1) / teacher_temp, dim=-1).detach()
    return (-t * s).sum(-1).mean()
