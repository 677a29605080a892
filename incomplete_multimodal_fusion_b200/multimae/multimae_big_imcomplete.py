"""Drop-in for the downstream incomplete-modality backbone: `ViTBaseline` of
downstream/instance_segmentation/modeling/multimae/multimae_big_imcomplete.py:418-454,534-680.

The encoder is the fusion-block MultiMAE of the pre-training package, run on whatever subset of {s1, s2, dem} is present:
an ABSENT modality contributes no tokens and -- unlike the pre-training model, which fills its slot with `mask_embedding`
-- no slot in the per-position modality attention (`all_idx` only holds the present modalities, :586-606, :633-647), so
the segment table and the slot map handed to the kernels are built over the present modalities only.  The fusion tokens
after blocks `flags` (every depth/4-th) are normalised with the final `norm`, reshaped to [B, D, H/P, W/P] and passed
through the four pyramid heads (:661-676).  The heads are plain torch modules (ConvTranspose2d / GroupNorm / MaxPool2d): they
are the FPN neck of the segmentation models, outside the encoder hot path.
"""
import os
import random
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence, Union

import torch
from torch import nn

from .. import functions as Fn
from ._core import MultiMAEBase, ZorroMask, block_params
from .zorro_utils import LayerNorm, TokenTypes


class MultiMAE(MultiMAEBase):
    """encoder of the downstream file (:43-118): the fusion-block model without the per-modality return tokens"""
    FUSION_BLOCKS = True

    def __init__(self, input_adapters: Dict[str, nn.Module], output_adapters: Optional[Dict[str, nn.Module]] = None,
                 in_domains: Sequence[str] = ('s1', 's2', 'dem'), num_global_tokens: int = 1, dim_tokens: int = 768, depth: int = 12,
                 dim_head: int = 64, heads: int = 8, ff_mult: int = 4, num_fusion_tokens: int = 16,
                 return_token_types=(TokenTypes.S1, TokenTypes.S2, TokenTypes.DEM, TokenTypes.FUSION),
                 drop_path_rate: float = 0.0, norm_layer: nn.Module = LayerNorm):
        self.MODALITIES = tuple(in_domains)
        super().__init__(input_adapters, output_adapters, num_global_tokens, dim_tokens, depth, dim_head, heads, ff_mult,
                         num_fusion_tokens, return_token_types, drop_path_rate, norm_layer)
        self.in_domains = list(in_domains)
        # the downstream encoder has no per-modality contrastive queries (they exist only in multimae_crossattn.py:105-109)
        del self.return_token_s1, self.return_token_s2, self.return_token_dem


class ViTBaseline(MultiMAE):
    def __init__(self, pretrained=None, pretrain_size=224, frozen_stages=12, freeze_attn=False, freeze_ffn=False, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.frozen_stages, self.freeze_attn, self.freeze_ffn = frozen_stages, freeze_attn, freeze_ffn
        self.cls_token = None
        self.num_block = len(self.blocks)
        self.pretrain_size = (pretrain_size, pretrain_size)
        self.flags = [i for i in range(-1, self.num_block, self.num_block // 4)][1:]
        D = self.dim_tokens
        self.up1 = nn.Sequential(nn.ConvTranspose2d(D, D, (2, 2), (2, 2)), nn.GroupNorm(32, D), nn.GELU(),
                                 nn.ConvTranspose2d(D, D, (2, 2), (2, 2)))
        self.up2 = nn.ConvTranspose2d(D, D, (2, 2), (2, 2))
        self.up3 = nn.Identity()
        self.up4 = nn.MaxPool2d(kernel_size=2, stride=2)
        self.incomplete_domains = list(self.in_domains)
        if isinstance(pretrained, str) and os.path.exists(pretrained):
            self.init_weights(pretrained)

    def init_weights(self, pretrained: str = None):
        if isinstance(pretrained, str):
            self.load_state_dict(torch.load(pretrained)['model'], strict=False)

    def forward_features(self, x: Union[Dict[str, torch.Tensor], torch.Tensor], mask_inputs: bool = False,
                         task_masks: Dict[str, torch.Tensor] = None, num_encoded_tokens: int = None, alphas=1.0,
                         sample_tasks_uniformly: bool = False):
        if self.training:   # a random non-empty subset of the modalities per step (:542-546)
            self.incomplete_domains = random.sample(self.in_domains, random.randint(1, len(self.in_domains)))
        else:
            self.incomplete_domains = list(self.in_domains)
        one = self.incomplete_domains[0]
        x = {one: x} if isinstance(x, torch.Tensor) else x
        present = [t for t in x if t in self.input_adapters and t in self.incomplete_domains]
        if not present:
            raise KeyError("none of the model's input modalities is present in the input")
        B, _, H, W = x[present[0]].shape
        device = x[present[0]].device
        if device.type != 'cuda':
            raise RuntimeError("incomplete_multimodal_fusion_b200.ViTBaseline runs on CUDA only (no CPU fallback)")
        D, Hh = self.dim_tokens, self.heads
        grids = {t: self.input_adapters[t].grid(H, W) for t in present}
        Fn_tok = self.fusion_tokens.shape[1]
        fus_ad = self.input_adapters['fusion']
        N_H, N_W = H // fus_ad.P_H, W // fus_ad.P_W
        carriers = OrderedDict((t, torch.empty(B, grids[t][0] * grids[t][1], 0, device=device)) for t in present)
        total = sum(c.shape[1] for c in carriers.values())
        if mask_inputs:
            nenc = num_encoded_tokens if num_encoded_tokens is not None else total
        else:
            nenc = int(total * 0.9) if self.training else total          # (:576-580)
        sizes = [c.shape[1] for c in carriers.values()]
        if task_masks is None:
            r = self._sample_masks(carriers, nenc, alphas=alphas, sample_tasks_uniformly=sample_tasks_uniformly)
        else:
            r = self._explicit_masks(task_masks, present, sizes, B, nenc, device)
        # eval encodes every token of the present modalities (nenc == total): the counts are the grid sizes and nothing is
        # read back; training's 0.9-of-the-tokens masking reads the counts once (counts_host) to size the patch-embed GEMMs
        idx = r.idx_list(present)
        slotmap = r["slotmap"]
        if slotmap is None:
            slotmap = torch.full((len(present), Fn_tok), -1, dtype=torch.int32, device=device)
            for m, ix in enumerate(idx):
                slotmap[m, ix.long()] = torch.arange(ix.numel(), dtype=torch.int32, device=device)
        counts = [int(i.numel()) for i in idx]
        if sum(counts) != nenc:
            raise ValueError(f"num_encoded_tokens={nenc} but the masks keep {sum(counts)} tokens")
        zmask = ZorroMask(counts, Fn_tok, device)          # segments = the PRESENT modalities, then the fusion tokens

        mod_args = []
        for t in present:
            ad = self.input_adapters[t]
            mod_args += [x[t].float(), ad.proj.weight, ad.proj.bias]
        meta_e = dict(B=B, D=D, P=self.input_adapters[present[0]].P_H, F=Fn_tok, nenc=nenc, idx=idx,
                      pos=[self.input_adapters[t].pos_table(*grids[t]) for t in present], pos_fusion=fus_ad.pos_table(N_H, N_W))
        X = Fn.EmbedFn.apply(meta_e, self.fusion_tokens, *mod_args)
        meta = dict(B=B, D=D, H=Hh, F=Fn_tok, nenc=nenc, fusion=True, depth=self.depth, I=int(D * self.ff_mult * 2 / 3),
                    seg=zmask.seg, nseg=zmask.nseg, slotmap=slotmap.contiguous(), grad_hook=getattr(self, 'grad_hook', None),
                    grad_hook_inplace=getattr(self, 'grad_hook_inplace', None), taps=list(self.flags))
        params = [self.mask_embedding]
        for fus, blk in zip(self.fus_blocks, self.blocks):
            params += block_params(fus) + block_params(blk)
        out = Fn.EncoderStackFn.apply(meta, X, *params)
        outs = [t.view(B, Fn_tok, D) for t in out[1:]]
        return outs, N_H, N_W

    def forward(self, input_dict) -> List[torch.Tensor]:
        outs, H, W = self.forward_features(input_dict)
        feats = []
        for f, up in zip(outs, (self.up1, self.up2, self.up3, self.up4)):
            bs, n, dim = f.shape
            t = Fn.layer_norm(f.reshape(bs * n, dim), self.norm.gamma, None, 1e-5, out_bf16=False).view(bs, n, dim)
            feats.append(up(t.transpose(1, 2).reshape(bs, dim, H, W)).contiguous())
        return feats
