"""Shared implementation of the two MultiMAE variants the reference trains:
`multimae.py` (plain zorro encoder) and `multimae_crossattn.py` (+ per-layer modality attention,
+ three per-modality return tokens).  Same constructor, `forward` signature, return tuples and
state_dict as the reference; the compute is the fused sm_100a path:

  masks (reference RNG call order kept verbatim) -> visible-patch embed GEMM -> EncoderStackFn
  -> final norm -> pooling head (one pool-attention launch for all return tokens) -> decoders.

Deliberate, documented differences from the reference:
  * always runs with CUDA-autocast(bf16) numerics (fp32 residual stream / norms / softmax stats,
    bf16 tensor-core operands), whatever the ambient autocast state; CPU tensors raise;
  * requires sum of visible tokens == num_encoded_tokens (the reference silently mis-slices
    otherwise, Appendix C) and raises instead;
  * explicit `task_masks` use a stable argsort (the reference's CUDA argsort tie order is undefined).
"""
import itertools
import math
from collections import OrderedDict
from typing import Dict, List, Optional, Tuple, Union

import torch
import torch.nn as nn
from torch.distributions.dirichlet import Dirichlet

from .. import functions as Fn
from .. import kernels as K
from .multimae_utils import trunc_normal_
from .zorro_utils import Attention, AttentionBiLSTM, Block, Block_Fusion, LayerNorm, Mlp, TokenTypes, ZorroMask, block_params

MODALITIES = ('s1', 's2', 'dem')   # hard-coded token order of the reference (multimae.py:378-407)
_MASK_STREAMS = {}                 # device index -> the mask sampler's side stream (see _sample_masks)
_VALIDATED_MASKS = {}              # explicit task_masks already checked against num_encoded_tokens (see _explicit_masks)


class MaskTables(dict):
    """What the mask builder leaves on the DEVICE for one step (one mask row for the whole batch): `mask`, `ids_restore`,
    `ids_keep`, `idx` (per-task ascending visible positions at the task offsets), `counts`, `seg` (zorro segment table),
    `slotmap`, `tok` (visible tokens in encoder order as global ids).  The fused path reads only these; nothing here
    needs the per-modality counts on the host.  `counts_host` / `idx_list()` are the legacy accessors (one host read-back)
    for callers that size tensors by the counts (the 4-modality semantic adapter, ViTBaseline)."""

    @property
    def counts_host(self):
        if "_counts_host" not in self:
            if self["nenc"] == sum(self["sizes"]):          # nothing masked: the counts are the sizes, no read-back
                self["_counts_host"] = list(self["sizes"])
            else:
                side = self.get("stream")            # read on the builder's own stream: no wait for the main stream's queue
                if side is not None:
                    with torch.cuda.stream(side):
                        self["_counts_host"] = self["counts"].tolist()
                else:
                    self["_counts_host"] = self["counts"].tolist()
        return self["_counts_host"]

    def idx_list(self, order=None):
        cnt = self.counts_host
        off = [0]
        for n in self["sizes"]:
            off.append(off[-1] + n)
        where = {t: i for i, t in enumerate(self["tasks"])}
        order = order or self["tasks"]
        return [self["idx"][off[where[t]]: off[where[t]] + cnt[where[t]]] for t in order]


class MultiMAEBase(nn.Module):
    FUSION_BLOCKS = False              # per-layer modality attention (multimae_crossattn.py)
    LSTM_FUSION = False                # one BiLSTM-initialised fusion token per visible token (multimae_lstm_s2dsm.py)
    MODALITIES = MODALITIES            # token order of the variant's forward
    TYPE_IDS = {'s1': TokenTypes.S1.value, 's2': TokenTypes.S2.value, 'dem': TokenTypes.DEM.value}
    FUSION_TYPE_ID = TokenTypes.FUSION.value   # (the 4-modality variant renumbers: multimae_quadruplet.py)

    def __init__(self, input_adapters: Dict[str, nn.Module], output_adapters: Optional[Dict[str, nn.Module]],
                 num_global_tokens: int = 1, dim_tokens: int = 768, depth: int = 12, dim_head: int = 64, heads: int = 8,
                 ff_mult: int = 4, num_fusion_tokens: int = 16,
                 return_token_types: Tuple[TokenTypes] = (TokenTypes.S1, TokenTypes.S2, TokenTypes.DEM, TokenTypes.FUSION),
                 drop_path_rate: float = 0.0, norm_layer: nn.Module = LayerNorm):
        super().__init__()
        if drop_path_rate != 0.0:
            raise NotImplementedError("stochastic depth is not built (rate 0 in every reference script)")
        for adapter in input_adapters.values():
            adapter.init(dim_tokens=dim_tokens)
        self.input_adapters = nn.ModuleDict(input_adapters)
        if output_adapters is not None:
            for adapter in output_adapters.values():
                adapter.init(dim_tokens_enc=dim_tokens)
            self.output_adapters = nn.ModuleDict(output_adapters)
        else:
            self.output_adapters = None
        assert num_fusion_tokens == input_adapters[self.MODALITIES[0]].num_patches

        self.dim_tokens, self.depth, self.heads, self.dim_head, self.ff_mult = dim_tokens, depth, heads, dim_head, ff_mult
        self.max_return_tokens = len(return_token_types)
        self.return_token_types = return_token_types
        self.register_buffer('return_token_types_tensor', torch.tensor([t.value for t in return_token_types]), persistent=False)

        self.return_tokens = nn.Parameter(torch.randn(1, self.max_return_tokens, dim_tokens))
        trunc_normal_(self.return_tokens, std=0.02)
        self.attn_pool = Attention(dim=dim_tokens, dim_head=dim_head, heads=heads)
        self.fusion_tokens = nn.Parameter(torch.randn(1, num_fusion_tokens, dim_tokens))
        trunc_normal_(self.fusion_tokens, std=0.02)
        if self.FUSION_BLOCKS:
            self.return_token_s1 = nn.Parameter(torch.randn(1, 1, dim_tokens))
            self.return_token_s2 = nn.Parameter(torch.randn(1, 1, dim_tokens))
            self.return_token_dem = nn.Parameter(torch.randn(1, 1, dim_tokens))
        self.mlp = Mlp(in_features=dim_tokens, hidden_features=int(dim_tokens * 4.0))
        if self.LSTM_FUSION:
            self.attn_lstm = AttentionBiLSTM(dim_tokens)
        if self.FUSION_BLOCKS:
            self.fus_blocks = nn.ModuleList([Block_Fusion(dim=dim_tokens, dim_head=dim_head, heads=heads, ff_mult=ff_mult,
                                                          norm_layer=norm_layer) for _ in range(depth)])
            self.mask_embedding = nn.Parameter(torch.zeros(1, num_fusion_tokens, dim_tokens))
        self.blocks = nn.ModuleList([Block(dim=dim_tokens, dim_head=dim_head, heads=heads, ff_mult=ff_mult, drop_path=0.,
                                           norm_layer=norm_layer) for _ in range(depth)])
        self.norm = LayerNorm(dim_tokens)
        self._init_parameters()
        # loading a checkpoint writes the parameters in place: drop the bf16 weight images cached from the old values
        self.register_load_state_dict_post_hook(lambda module, incompatible: Fn.invalidate_weight_cache())

    # ---- initialisation: multimae.py:116-142 ----
    def _init_parameters(self):
        for name, m in self.named_modules():
            if isinstance(m, nn.Linear):
                fan_out, fan_in = m.weight.shape
                if 'qkv' in name:
                    fan_out //= 3
                elif 'kv' in name:
                    fan_out //= 2
                bound = math.sqrt(6. / float(fan_out + fan_in))
                nn.init.uniform_(m.weight, -bound, bound)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.LayerNorm):
                nn.init.constant_(m.bias, 0)
                nn.init.constant_(m.weight, 1.0)
            elif isinstance(m, nn.Conv2d) and '.proj' in name:
                w = m.weight.data
                nn.init.xavier_uniform_(w.view([w.shape[0], -1]))

    def get_num_layers(self):
        return len(self.blocks)

    @torch.jit.ignore
    def no_weight_decay(self):
        skip = {'global_tokens'}
        for group, adapters in (('input_adapters', self.input_adapters), ('output_adapters', self.output_adapters or {})):
            for task, adapter in adapters.items():
                if hasattr(adapter, 'no_weight_decay'):
                    skip |= {f'{group}.{task}.{name}' for name in adapter.no_weight_decay()}
        return skip

    # ---- mask sampling: same torch RNG calls, in the same order, as multimae.py:165-255 ----
    def sample_alphas(self, B: int, n_tasks: int, alphas: float = 1.0, eps: float = 1e-5):
        choices = torch.Tensor([list(c) for c in itertools.product([0, 1], repeat=n_tasks)][1:])
        picked = torch.index_select(choices, 0, torch.randint(0, len(choices), (B,)))
        return picked * torch.tensor(alphas) + eps

    def _sample_masks(self, input_tokens: Dict[str, torch.Tensor], num_encoded_tokens: int,
                      alphas: Union[float, List[float]] = 1.0, sample_tasks_uniformly: bool = False):
        """The reference's torch RNG calls verbatim (same generator, shapes and order: CPU Dirichlet :208, one
        torch.rand(1, n) per task on the device :217, rand_like over all tokens :241), then ONE mask-builder launch
        (kernels.mask_build) for the argsorts / gathers / where / nonzero bookkeeping of :210-255 and :378-426."""
        first = next(iter(input_tokens.values()))
        B, device = first.shape[0], first.device
        if device.type != 'cuda':
            raise RuntimeError("mask sampling runs on CUDA only (no CPU fallback)")
        sizes = [t.shape[1] for t in input_tokens.values()]
        alphas = [alphas] * len(sizes) if isinstance(alphas, float) else alphas
        if sample_tasks_uniformly:
            share_cpu = Dirichlet(self.sample_alphas(1, len(sizes), alphas=alphas)).sample()
        else:
            share_cpu = Dirichlet(torch.Tensor(alphas)).sample((1,))
        Fn_tok = self.fusion_tokens.shape[1]
        want_slot = (self.FUSION_BLOCKS or self.LSTM_FUSION) and all(n == Fn_tok for n in sizes)
        # The draws and the mask builder run on their own (high-priority) stream: they depend on nothing of the step, so
        # they can run ahead of whatever the main stream still holds.  The mask depends on the random draws only (whose
        # Philox offsets are assigned in host call order, whatever the stream), so the values are unchanged.  Nothing is
        # read back: every consumer takes the device tables (MaskTables).
        main = torch.cuda.current_stream(device)
        side = _MASK_STREAMS.get(device.index)
        if side is None:
            side = _MASK_STREAMS[device.index] = torch.cuda.Stream(device=device, priority=-1)
        with torch.cuda.stream(side):
            share = share_cpu.to(device)
            noise1 = torch.cat([torch.rand(1, n, device=device) for n in sizes], dim=1)
            noise2 = torch.rand_like(noise1)
            mask, ids_restore, ids_keep, idx, counts, seg, slotmap, tok = K.mask_build(
                noise1.view(-1), noise2.view(-1), share.float().view(-1).contiguous(), sizes, num_encoded_tokens, Fn_tok, want_slot)
        main.wait_stream(side)
        for t in (mask, ids_restore, ids_keep, idx, counts, seg, slotmap, tok):
            if t is not None:
                t.record_stream(main)
        return MaskTables(tasks=list(input_tokens.keys()), sizes=sizes, B=B, nenc=num_encoded_tokens, mask=mask, ids_restore=ids_restore,
                          ids_keep=ids_keep, idx=idx, counts=counts, seg=seg, slotmap=slotmap, tok=tok, stream=side)

    def _explicit_masks(self, task_masks: Dict[str, torch.Tensor], tasks: List[str], sizes: List[int], B: int, nenc: int, device):
        """Caller-provided masks (multimae.py:372-376 + the selections of :378-383) through the same single-CTA builder: a
        stable partition of row 0 of the masks (the reference selects tokens from row 0 as well, :378-382) gives
        ids_shuffle / ids_restore / ids_keep and every table of the sampled branch, on the device.  The reference reads
        (mask_all == 0).sum() and three nonzero() results back; here the kept-token count is checked against
        num_encoded_tokens ONCE per mask object (keyed on storage, version and shape -- inference loops pass the same
        masks for every batch), so repeated calls have no host round-trip at all."""
        Fn_tok = self.fusion_tokens.shape[1]
        want_slot = (self.FUSION_BLOCKS or self.LSTM_FUSION) and all(n == Fn_tok for n in sizes)
        rows = [task_masks[t][0].to(device=device, dtype=torch.int64) for t in tasks]
        given = torch.cat(rows).contiguous()
        ids_restore, ids_keep, idx, counts, seg, slotmap, tok, err = K.mask_explicit(given, sizes, nenc, Fn_tok, want_slot)
        key = tuple((task_masks[t].data_ptr(), task_masks[t]._version, tuple(task_masks[t].shape)) for t in tasks) + (nenc,)
        if key not in _VALIDATED_MASKS:
            if int(err.item()):
                kept = int((given == 0).sum())
                raise ValueError(f"num_encoded_tokens={nenc} but the masks keep {kept} tokens; pass the true count "
                                 "(the reference mis-slices silently in this case)")
            if len(_VALIDATED_MASKS) > 256:
                _VALIDATED_MASKS.clear()
            _VALIDATED_MASKS[key] = True
        Bm = next(iter(task_masks.values())).shape[0]
        return MaskTables(tasks=tasks, sizes=sizes, B=B, nenc=nenc, mask=given, ids_restore=ids_restore.unsqueeze(0).expand(Bm, -1),
                          ids_keep=ids_keep.unsqueeze(0).expand(Bm, -1), idx=idx, counts=counts, seg=seg, slotmap=slotmap, tok=tok)

    def generate_random_masks(self, input_tokens: Dict[str, torch.Tensor], num_encoded_tokens: int,
                              alphas: Union[float, List[float]] = 1.0, sample_tasks_uniformly: bool = False):
        """One mask for the whole batch (the reference's edit of MultiMAE's per-sample masking).  Returns
        (task_masks {task: [B, n] int64, 0 = keep}, ids_keep [B, nenc], ids_restore [B, sum n])."""
        r = self._sample_masks(input_tokens, num_encoded_tokens, alphas, sample_tasks_uniformly)
        return self._public_masks(r)

    @staticmethod
    def _public_masks(r):
        B = r["B"]
        masks = {t: m.unsqueeze(0).repeat(B, 1) for t, m in zip(r["tasks"], torch.split(r["mask"], r["sizes"]))}
        return masks, r["ids_keep"].unsqueeze(0).repeat(B, 1), r["ids_restore"].unsqueeze(0).repeat(B, 1)

    @staticmethod
    def make_mask(N_H, N_W, xy_idxs, full_tasks=[], indicate_visible=True, flatten=True, device='cuda'):
        """masks from lists of visible (x, y) patch coordinates (multimae.py:257-285)"""
        masks = {}
        for k, pts in xy_idxs.items():
            m = torch.ones(N_H, N_W, device=device)
            pts = torch.LongTensor(pts)
            if len(pts) > 0:
                m[pts[:, 1], pts[:, 0]] = 0
            masks[k] = m
        for task in full_tasks:
            masks[task][:] = 0
        if not indicate_visible:
            masks = {k: 1 - v for k, v in masks.items()}
        if flatten:
            masks = {k: v.flatten().unsqueeze(0) for k, v in masks.items()}
        return masks

    def generate_input_info(self, input_task_tokens, image_size):
        info = OrderedDict()
        info['tasks'] = {}
        start = 0
        for domain, tensor in input_task_tokens.items():
            n = tensor.shape[1]
            info['tasks'][domain] = {'num_tokens': n, 'has_2d_posemb': True, 'start_idx': start, 'end_idx': start + n}
            start += n
        info['image_size'] = image_size
        info['num_task_tokens'] = start
        return info

    # ---- forward ----
    def _device_const(self, key, device, make):
        """small constant device tensors, built once per (key, device) instead of a pageable H2D copy per step"""
        cache = self.__dict__.setdefault('_const_cache', {})
        k = (key, str(device))
        if k not in cache:
            cache[k] = make().to(device)
        return cache[k]

    def forward(self, x: Union[Dict[str, torch.Tensor], torch.Tensor], mask_inputs: bool = True,
                task_masks: Dict[str, torch.Tensor] = None, num_encoded_tokens: int = 128,
                alphas: Union[float, List[float]] = 1.0, sample_tasks_uniformly: bool = False,
                fp32_output_adapters: List[str] = [], return_token_indices: Optional[Tuple[int]] = None):
        MODALITIES = self.MODALITIES
        first = MODALITIES[0]
        x = {first: x} if isinstance(x, torch.Tensor) else x
        B, _, H, W = x[first].shape
        device = x[first].device
        if device.type != 'cuda':
            raise RuntimeError("incomplete_multimodal_fusion_b200.MultiMAE runs on CUDA only (no CPU fallback)")
        for t in MODALITIES:
            if t not in x or t not in self.input_adapters:
                raise KeyError(f"input '{t}' is required (the reference hard-codes its modalities, multimae.py:378-383)")
        D, Hh = self.dim_tokens, self.heads
        tasks = [t for t in x if t in self.input_adapters]
        grids = {t: self.input_adapters[t].grid(H, W) for t in tasks}
        Fn_tok = self.fusion_tokens.shape[1]
        # shape carriers: the mask sampler only looks at .shape[0], .shape[1] and .device
        carriers = OrderedDict((t, torch.empty(B, grids[t][0] * grids[t][1], 0, device=device)) for t in tasks)
        input_info = self.generate_input_info(carriers, image_size=(H, W))
        nenc = num_encoded_tokens if mask_inputs else sum(c.shape[1] for c in carriers.values())

        if task_masks is None:
            r = self._sample_masks(carriers, nenc, alphas=alphas, sample_tasks_uniformly=sample_tasks_uniformly)
            task_masks, ids_keep, ids_restore = self._public_masks(r)
        else:
            r = self._explicit_masks(task_masks, tasks, [c.shape[1] for c in carriers.values()], B, nenc, device)
            ids_keep, ids_restore = r["ids_keep"], r["ids_restore"]
        in_order = r["tasks"][:len(MODALITIES)] == list(MODALITIES)       # the input dict lists the modalities in encoder order
        kinds = [getattr(self.input_adapters[t], 'KIND', 'patch') for t in MODALITIES]
        # device-table path: all modalities are plain patch adapters in encoder order (every reference script); otherwise
        # the legacy path sizes per-modality tensors by the counts (one host read-back)
        fused_tables = in_order and len(r["tasks"]) == len(MODALITIES) and all(k == 'patch' for k in kinds) and \
            all(sz == Fn_tok for sz in r["sizes"])
        n_tail = nenc if self.LSTM_FUSION else Fn_tok          # fusion tokens in the encoder sequence
        nseg = len(MODALITIES) + 1
        slotmap = r["slotmap"]
        if fused_tables:
            idx = counts = None
            seg = r["seg"]
            if n_tail != Fn_tok:                               # (LSTM variant: one fusion token per visible token)
                seg = torch.cat([seg[:-1], seg[-2:-1] + n_tail])
        else:
            idx = r.idx_list(MODALITIES)
            where = {t: i for i, t in enumerate(r["tasks"])}
            counts = [r.counts_host[where[t]] for t in MODALITIES]
            if sum(counts) != nenc:
                raise ValueError(f"num_encoded_tokens={nenc} but the masks keep {sum(counts)} tokens; pass the true count "
                                 "(the reference mis-slices silently in this case)")
            seg = ZorroMask(counts, n_tail, device).seg
            if slotmap is not None and not in_order:
                slotmap = torch.stack([slotmap[where[t]] for t in MODALITIES]).contiguous()

        # ---- tokens: visible-patch embedding + fusion tokens, planar layout ----
        mod_args, pads = [], {}
        for m, t in enumerate(MODALITIES):
            ad = self.input_adapters[t]
            if kinds[m] == 'semseg':          # class map [B, H, W] (SemSegInputAdapter): embedded through a one-hot GEMM
                mod_args += ad.embed_args(x[t])
                pads[m] = ad.emb_padding_idx
            else:
                mod_args += [x[t].float(), ad.proj.weight, ad.proj.bias]
        fus_ad = self.input_adapters['fusion']
        pos_fusion = fus_ad.pos_table(H // fus_ad.P_H, W // fus_ad.P_W)
        meta_e = dict(B=B, D=D, P=self.input_adapters[first].P_H, F=0 if self.LSTM_FUSION else Fn_tok, nenc=nenc, idx=idx,
                      tok=r["tok"] if fused_tables else None,
                      pos=[self.input_adapters[t].pos_table(*grids[t]) for t in MODALITIES], pos_fusion=pos_fusion, kinds=kinds,
                      padding_idx=pads)
        X = Fn.EmbedFn.apply(meta_e, self.fusion_tokens, *mod_args)
        # learnable positional embeddings (off in every reference script): their gradient rides a pass-through node
        pos_g = [self.input_adapters[t].pos_table_grad(*grids[t]) for t in MODALITIES]
        fus_g = None if self.LSTM_FUSION else fus_ad.pos_table_grad(H // fus_ad.P_H, W // fus_ad.P_W)
        if any(g is not None for g in pos_g) or fus_g is not None:
            if not fused_tables:
                raise NotImplementedError("learnable positional embeddings need the device-table path (patch adapters in encoder order)")
            tab = None
            if any(g is not None for g in pos_g):
                tab = torch.cat([g if g is not None else p_.detach() for g, p_ in zip(pos_g, meta_e["pos"])], 0)
            X = Fn.PosEmbGradFn.apply(X, r["tok"], dict(B=B, nenc=nenc, F=meta_e["F"]), tab, fus_g)
        complete_fusion = None
        if self.LSTM_FUSION:
            # multimae_lstm_s2dsm.py:384-434: the fusion token of every visible position, merged with that position's
            # modality token by the BiLSTM (cuDNN under bf16 autocast, like the reference's AMP step)
            pos_fusion_t = fus_ad.pos_table_grad(H // fus_ad.P_H, W // fus_ad.P_W)     # on the tape when learnable
            complete_fusion = self.fusion_tokens[0] + (pos_fusion if pos_fusion_t is None else pos_fusion_t)   # [F, D]
            sel = (r["tok"].long() % Fn_tok) if fused_tables else torch.cat([i.long() for i in idx])   # patch of every visible token
            pairs = torch.stack([X.view(B, nenc, D), complete_fusion[sel].unsqueeze(0).expand(B, -1, -1)], dim=2)
            with torch.autocast('cuda', dtype=torch.bfloat16):
                fus0 = self.attn_lstm(pairs.reshape(B * nenc, 2, D))
            X = torch.cat([X, fus0.float()], dim=0)                                   # planar: modality rows, then fusion rows

        # ---- encoder stack ----
        if self.FUSION_BLOCKS and slotmap is None:            # (legacy path with grids that differ from the fusion grid)
            slotmap = torch.full((len(MODALITIES), Fn_tok), -1, dtype=torch.int32, device=device)
            for m, ix in enumerate(idx):
                slotmap[m, ix.long()] = torch.arange(ix.numel(), dtype=torch.int32, device=device)
        meta = dict(B=B, D=D, H=Hh, F=n_tail, nenc=nenc, fusion=self.FUSION_BLOCKS, depth=self.depth,
                    I=int(D * self.ff_mult * 2 / 3), seg=seg, nseg=nseg, slotmap=slotmap,
                    grad_hook=getattr(self, 'grad_hook', None),
                    grad_hook_inplace=getattr(self, 'grad_hook_inplace', None))
        params = []
        if self.FUSION_BLOCKS:
            params.append(self.mask_embedding)
            for fus, blk in zip(self.fus_blocks, self.blocks):
                params += block_params(fus) + block_params(blk)
        else:
            for blk in self.blocks:
                params += block_params(blk)
        X = Fn.EncoderStackFn.apply(meta, X, *params)
        Mh = B * nenc
        T = Fn.layer_norm(X, self.norm.gamma, None, 1e-5, out_bf16=False)           # final norm, fp32 (multimae.py:431)
        ori_tokens = T[:Mh].view(B, nenc, D)
        enc_fusion = T[Mh:].view(B, n_tail, D)

        # ---- pooling head: all return tokens in one pool-attention launch ----
        rt = self.return_tokens
        rtypes = self.return_token_types_tensor
        if return_token_indices is not None:
            assert len(set(return_token_indices)) == len(return_token_indices), 'all indices must be unique'
            assert all(i < self.max_return_tokens for i in return_token_indices), \
                'indices must range from 0 to max_num_return_tokens - 1'
            sel = torch.tensor(return_token_indices, dtype=torch.long, device=device)
            rt, rtypes = rt[:, sel], rtypes[sel]
        R = rt.shape[1]
        queries = [rt[0]]
        if self.FUSION_BLOCKS:
            queries += [self.return_token_s1[0], self.return_token_s2[0], self.return_token_dem[0]]
        queries = torch.cat(queries, 0)
        Rt = queries.shape[0]
        N = nenc + n_tail
        # token type of every position from the device segment table (a repeat_interleave with device repeats reads its
        # output size back: a full host synchronisation in the middle of the step, after which the small decoder / loss
        # kernels ran host-bound)
        type_ids = self._device_const(('type_ids', tuple(MODALITIES)), device,
                                      lambda: torch.tensor([self.TYPE_IDS[t] for t in MODALITIES] + [self.FUSION_TYPE_ID]))
        pos = self._device_const(('arange', N), device, lambda: torch.arange(N, dtype=torch.int32))
        types = type_ids[(pos[:, None] >= seg[None, 1:-1]).sum(1)]
        pmask = torch.zeros(Rt, N, dtype=torch.uint8, device=device)
        pmask[:R] = ((rtypes[:, None] == types[None, :]) | (rtypes[:, None] == self.FUSION_TYPE_ID)).to(torch.uint8)
        mode = torch.zeros(Rt, dtype=torch.int32, device=device)
        if self.FUSION_BLOCKS:
            # per-modality pools over the fusion tokens at that modality's visible positions (multimae_crossattn.py:529-543):
            # straight from the mask rows (0 = visible), no index lists
            vis = torch.stack([(task_masks[t][0] == 0) for t in MODALITIES]).to(device=device, dtype=torch.uint8)    # [M, F]
            pmask[R:R + len(MODALITIES), nenc:nenc + vis.shape[1]] = vis
            mode[R:] = 1                            # empty context -> zeros, not the uniform fallback
        ap = self.attn_pool
        # the learned queries are batch-invariant rows: their normalised values and projection stay fp32 on the autograd
        # tape so that the batch-summed gradient is not rounded to bf16 once for the whole batch (LinearFn.precise_grad)
        qn = Fn.layer_norm(queries.float(), ap.norm.gamma, None, 1e-5, out_bf16=False)
        q = Fn.linear(qn, ap.to_q.weight, out_f32=True, precise_grad=True)
        kv = Fn.linear(T, ap.to_kv.weight)
        pooled = Fn.PoolAttnFn.apply(q, kv, pmask, mode, B, Hh, N, nenc, ap.scale)
        r = Fn.linear(pooled.view(B * Rt, Hh * 64), ap.to_out.weight)
        rn = Fn.layer_norm(r.float(), self.norm.gamma, None, 1e-5, out_bf16=True)
        r = (r + self.mlp(rn)).view(B, Rt, D)
        return_tokens = r[:, :R]

        if self.output_adapters is None:
            tokens = torch.cat([ori_tokens, enc_fusion], dim=1)
            return tokens, return_tokens, task_masks

        dec_in = enc_fusion
        if self.LSTM_FUSION:
            # decoders read the whole fusion grid with the encoded tokens scattered back (multimae_lstm_s2dsm.py:470-477;
            # the reference writes position by position in order, so a position visible in both modalities keeps the
            # later, dem, copy: one index_copy per modality in order reproduces that deterministically)
            if fused_tables:
                # slotmap[m][p] = rank of position p among modality m's visible tokens (or -1); the LAST modality that sees
                # p wins, as in the reference's in-order writes: one gather instead of per-modality index_copy
                src = torch.full((Fn_tok,), -1, dtype=torch.int64, device=device)
                for m in range(len(MODALITIES)):
                    sm = slotmap[m].long()
                    src = torch.where(sm >= 0, sm + seg[m].long(), src)
                picked = enc_fusion[:, src.clamp_min(0)]
                dec_in = torch.where((src >= 0)[None, :, None], picked, complete_fusion.unsqueeze(0).expand(B, -1, -1))
            else:
                dec_in = complete_fusion.unsqueeze(0).expand(B, -1, -1).clone()
                o = 0
                for ix in idx:
                    dec_in = dec_in.index_copy(1, ix.long(), enc_fusion[:, o:o + ix.numel()])
                    o += ix.numel()
        preds = {}
        for domain, adapter in self.output_adapters.items():
            p = adapter(encoder_tokens=dec_in, input_info=input_info, ids_keep=ids_keep, ids_restore=ids_restore)
            preds[domain] = p.float() if domain in fp32_output_adapters else p
        out = (preds, task_masks, return_tokens, ori_tokens, enc_fusion)
        if self.FUSION_BLOCKS:
            out = out + (r[:, R:R + 1], r[:, R + 1:R + 2], r[:, R + 2:R + 3])
        return out
