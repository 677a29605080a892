"""Oracle (TEST INFRASTRUCTURE) for the two-modality BiLSTM-fusion variant the reference trains in
``pretrain_mmae_s2dsm.py``: ``pretraining/multimae/multimae_lstm_s2dsm.py`` (BASELINE config 1).

Differences from ``multimae.py`` restated here (line numbers of multimae_lstm_s2dsm.py):
  * modalities are s2 and dem (:340-352); one fusion token per VISIBLE token: the fusion tokens (with pos-emb) at the
    visible positions cat(s2_idx, dem_idx) (:384-389);
  * each (token, fusion token) pair runs through ``AttentionBiLSTM`` (zorro_utils.py:261-299: bidirectional LSTM over
    the length-2 sequence, directions summed, tanh-attention pooling) to initialise the fusion token (:428-434);
  * plain zorro blocks over [s2 | dem | fusion] (:435-438), pooling with return types (S2, DEM, FUSION) (:441-467);
  * decoders read the full fusion-token grid with the encoded fusion tokens scattered back (later index wins, :474-477).
"""
from collections import OrderedDict
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from . import functional as O
from .functional import DEM, FUSION, S2, OracleConfig


def lstm_config(**kw) -> OracleConfig:
    """OracleConfig for the s2 (4 channels) + dem (1 channel) model of pretrain_mmae_s2dsm.py:45-65"""
    base = dict(variant="lstm_s2dsm", channels=OrderedDict([("s2", 4), ("dem", 1)]), return_token_types=(S2, DEM, FUSION),
                out_tasks=("s2", "dem"))
    base.update(kw)
    return OracleConfig(**base)


def add_lstm_params(sd, cfg: OracleConfig, seed: int = 0):
    """attn_lstm.* keys of multimae_lstm_s2dsm.py:107 (nn.LSTM default init U(-1/sqrt(D), 1/sqrt(D)); the attention
    Linear(D, 1) is xavier-initialised by `self.apply(_init_weights)`, :127-142)"""
    g = torch.Generator().manual_seed(seed + 1000)
    D = cfg.dim
    k = 1.0 / D ** 0.5
    for sfx in ("", "_reverse"):
        sd[f"attn_lstm.lstm.weight_ih_l0{sfx}"] = (torch.rand(4 * D, D, generator=g) * 2 - 1) * k
        sd[f"attn_lstm.lstm.weight_hh_l0{sfx}"] = (torch.rand(4 * D, D, generator=g) * 2 - 1) * k
        sd[f"attn_lstm.lstm.bias_ih_l0{sfx}"] = (torch.rand(4 * D, generator=g) * 2 - 1) * k
        sd[f"attn_lstm.lstm.bias_hh_l0{sfx}"] = (torch.rand(4 * D, generator=g) * 2 - 1) * k
    sd["attn_lstm.attention.attention.weight"] = O._xavier(1, D, g)
    sd["attn_lstm.attention.attention.bias"] = torch.zeros(1)
    return sd


def init_state_dict(cfg: OracleConfig, seed: int = 0):
    return add_lstm_params(O.init_state_dict(cfg, seed), cfg, seed)


def _lstm_dir(x, w_ih, w_hh, b_ih, b_hh, reverse: bool):
    """one direction of a single-layer LSTM over [n, T, D] (gate order i, f, g, o as torch.nn.LSTM)"""
    n, T, D = x.shape
    h = x.new_zeros(n, D)
    c = x.new_zeros(n, D)
    outs = [None] * T
    for t in (range(T - 1, -1, -1) if reverse else range(T)):
        gates = x[:, t] @ w_ih.t() + b_ih + h @ w_hh.t() + b_hh
        i, f, g, o = gates.chunk(4, dim=-1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
        h = torch.sigmoid(o) * torch.tanh(c)
        outs[t] = h
    return torch.stack(outs, dim=1)


def attention_bilstm(sd, pfx: str, x, cfg: OracleConfig):
    """AttentionBiLSTM.forward, zorro_utils.py:292-299 (mask=None): y = fwd + bwd hidden states,
    alpha = softmax_t(Linear(tanh(y))), r = sum_t alpha_t y_t"""
    x = O._f32(x, cfg)
    p = pfx + "lstm."
    yf = _lstm_dir(x, sd[p + "weight_ih_l0"], sd[p + "weight_hh_l0"], sd[p + "bias_ih_l0"], sd[p + "bias_hh_l0"], False)
    yb = _lstm_dir(x, sd[p + "weight_ih_l0_reverse"], sd[p + "weight_hh_l0_reverse"], sd[p + "bias_ih_l0_reverse"],
                   sd[p + "bias_hh_l0_reverse"], True)
    y = yf + yb
    m = torch.tanh(y) @ sd[pfx + "attention.attention.weight"].t() + sd[pfx + "attention.attention.bias"]
    alpha = F.softmax(m.squeeze(2), dim=1).unsqueeze(1)
    return alpha.bmm(y).squeeze(1)


def multimae_lstm_forward(sd, cfg: OracleConfig, x: Dict[str, torch.Tensor], mask_inputs: bool = True,
                          task_masks: Optional[Dict[str, torch.Tensor]] = None, num_encoded_tokens: int = 128,
                          alphas=1.0, sample_tasks_uniformly: bool = False, decode: bool = True,
                          return_token_indices=None):
    """multimae_lstm_s2dsm.py:312-502 -> (preds, task_masks, return_tokens, ori_tokens, encoder_fusion_tokens)"""
    dev = x["s2"].device
    B, _, H, W = x["s2"].shape
    tok = OrderedDict((t, O.patch_embed(sd, f"input_adapters.{t}.", img, cfg)) for t, img in x.items() if t in cfg.channels)
    complete = O.add_fusion_posemb(sd, sd["fusion_tokens"].expand(B, -1, -1))
    n_per_task = OrderedDict((t, v.shape[1]) for t, v in tok.items())
    input_info = O.build_input_info(n_per_task, (H, W))
    if not mask_inputs:
        num_encoded_tokens = sum(n_per_task.values())
    if task_masks is None:
        task_masks, ids_keep, ids_restore = O.generate_random_masks(
            n_per_task, B, num_encoded_tokens, dev, alphas=alphas, sample_tasks_uniformly=sample_tasks_uniformly)
    else:
        ids_keep, ids_restore = O.masks_from_task_masks(task_masks, list(tok.keys()))
    nenc = num_encoded_tokens
    s2_idx = (task_masks["s2"][0] == 0).nonzero(as_tuple=True)[0]
    dem_idx = (task_masks["dem"][0] == 0).nonzero(as_tuple=True)[0]
    sel = torch.cat([s2_idx, dem_idx], 0)
    tokens = torch.cat([tok["s2"][:, s2_idx], tok["dem"][:, dem_idx], complete[:, sel]], dim=1)
    types = torch.tensor([S2] * len(s2_idx) + [DEM] * len(dem_idx) + [FUSION] * len(sel), dtype=torch.long, device=dev)
    zmask = (types[:, None] == types[None, :]) | (types[:, None] == FUSION)

    pairs = torch.stack([tokens[:, :nenc], tokens[:, nenc:]], dim=2).reshape(B * nenc, 2, cfg.dim)
    fus = attention_bilstm(sd, "attn_lstm.", pairs, cfg).reshape(B, nenc, cfg.dim)
    tokens = torch.cat([tokens[:, :nenc], fus], dim=1)
    for i in range(cfg.depth):
        tokens = O.zorro_block(sd, f"blocks.{i}.", tokens, zmask, cfg)
    tokens = O.zorro_layer_norm(tokens, sd["norm.gamma"], cfg)

    rt = sd["return_tokens"]
    rtypes = list(cfg.return_token_types)
    if return_token_indices is not None:
        rt = rt[:, list(return_token_indices)]
        rtypes = [rtypes[i] for i in return_token_indices]
    rtt = torch.tensor(rtypes, dtype=torch.long, device=dev)
    pmask = (rtt[:, None] == types[None, :]) | (rtt[:, None] == FUSION)
    r = O.zorro_attention(sd, "attn_pool.", rt.expand(B, -1, -1), cfg, context=tokens, attn_mask=pmask)
    return_tokens = r + O.biased_mlp(sd, "mlp.", O.zorro_layer_norm(r, sd["norm.gamma"], cfg), cfg)
    if not decode:
        return tokens, return_tokens, task_masks

    ori_tokens = tokens[:, :nenc]
    enc_fusion = tokens[:, nenc:]
    full = complete.clone().to(enc_fusion.dtype)
    n1 = len(s2_idx)
    full[:, s2_idx] = enc_fusion[:, :n1]          # the reference assigns position by position in order (:474-477):
    full[:, dem_idx] = enc_fusion[:, n1:]         # a position visible in both modalities ends with the dem copy
    preds = {}
    for t in cfg.out_tasks:
        if cfg.decoder == "simple":
            preds[t] = O.simple_output_adapter(sd, t, full, (H, W), cfg)
        else:
            preds[t] = O.xattn_output_adapter(sd, t, full, input_info, ids_keep, ids_restore, cfg)
    return preds, task_masks, return_tokens, ori_tokens, enc_fusion


def pretrain_loss_s2dsm(out, targets, cfg: OracleConfig):
    """pretrain_mmae_s2dsm.py:470-492: masked MSE (s2) + masked L1 (dem) + HardNegtive_loss over the three pairs of
    pooled return tokens (weight 1)"""
    preds, masks = out[0], out[1]
    total = O.masked_mse_loss(preds["s2"].float(), targets["s2"], masks.get("s2"), cfg.patch) + \
        O.masked_l1_loss(preds["dem"].float(), targets["dem"], masks.get("dem"), cfg.patch)
    a, b, c = [t.squeeze(1).float() for t in torch.chunk(out[2], 3, dim=1)]
    total = total + O.hard_negative_loss(a, b) + O.hard_negative_loss(a, c) + O.hard_negative_loss(b, c)
    return total
