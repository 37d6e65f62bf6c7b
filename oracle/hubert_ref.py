"""CPU fp32 restatement of the HuBERT forward the reference calls.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The arithmetic lives in the
third-party `transformers.models.hubert` (pinned ==4.35.0 in the reference's
requirements.txt:8; 5.5.0 installed), reached from I_ea/model.py:82-88 and, for
I_da, through fairseq `extract_features` (I_da/src/hubert_feature_reader.py:60-65,
fairseq absent => HF arithmetic is the stand-in; fairseq-vs-HF parity unpinned).

All functions take a flat `params` dict with HF state-dict key names and a
`HubertCfg` (oracle/params.py).  `taps`, if given, collects per-stage activations
for bisecting the CUDA path.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def fold_pos_conv_weight(params: dict, prefix: str = "") -> torch.Tensor:
    """weight_norm(dim=2) of the positional conv (HF:59-78): w = g * v / ||v||_(0,1).
    Accepts old (`weight_g/_v`) and new (`parametrizations.weight.original0/1`) names."""
    base = prefix + "encoder.pos_conv_embed.conv."
    if base + "weight" in params:
        return params[base + "weight"]
    if base + "weight_g" in params:
        g, v = params[base + "weight_g"], params[base + "weight_v"]
    else:
        g = params[base + "parametrizations.weight.original0"]
        v = params[base + "parametrizations.weight.original1"]
    return g * v / v.pow(2).sum(dim=(0, 1), keepdim=True).sqrt()


def feature_encoder(params, cfg, input_values, prefix="", taps=None):
    """HF:203-213 `HubertFeatureEncoder.forward` -> [B, 512, T].
    group: conv0 -> GroupNorm(512,512) -> GELU, then conv+GELU (HF:154-175, 106-124);
    layer: every layer conv(+bias) -> LayerNorm(C) over channels -> GELU (HF:127-151)."""
    h = input_values[:, None, :]
    for i, s in enumerate(cfg.conv_stride):
        b = f"{prefix}feature_extractor.conv_layers.{i}."
        h = F.conv1d(h, params[b + "conv.weight"], params.get(b + "conv.bias"), stride=s)
        if cfg.feat_extract_norm == "group" and i == 0:
            c = h.shape[1]
            h = F.group_norm(h, c, params[b + "layer_norm.weight"], params[b + "layer_norm.bias"], eps=1e-5)
        elif cfg.feat_extract_norm == "layer":
            h = F.layer_norm(h.transpose(1, 2), (h.shape[1],), params[b + "layer_norm.weight"],
                             params[b + "layer_norm.bias"], eps=1e-5).transpose(1, 2)
        h = F.gelu(h)  # ACT2FN["gelu"] == exact erf GELU
        if taps is not None:
            taps[f"conv{i}"] = h.transpose(1, 2).contiguous()
    return h


def feature_vector_attention_mask(cfg, feat_len: int, attention_mask):
    """HF:690-700 `_get_feature_vector_attention_mask` -> bool [B, T]."""
    out_len = cfg.feat_lengths(attention_mask.sum(-1).to(torch.long))
    ar = torch.arange(feat_len)[None, :]
    return ar < out_len[:, None]


def attention(params, cfg, x, base, key_mask):
    """HF:296-345 + eager_attention_forward HF:234-259 (softmax in fp32, scale d^-1/2,
    additive -inf/finfo.min key-padding mask)."""
    B, T, H = x.shape
    nh = cfg.num_attention_heads
    d = H // nh
    q = F.linear(x, params[base + "q_proj.weight"], params[base + "q_proj.bias"]).view(B, T, nh, d).transpose(1, 2)
    k = F.linear(x, params[base + "k_proj.weight"], params[base + "k_proj.bias"]).view(B, T, nh, d).transpose(1, 2)
    v = F.linear(x, params[base + "v_proj.weight"], params[base + "v_proj.bias"]).view(B, T, nh, d).transpose(1, 2)
    s = torch.matmul(q, k.transpose(2, 3)) * (d ** -0.5)
    if key_mask is not None:
        s = s.masked_fill(~key_mask[:, None, None, :], torch.finfo(s.dtype).min)
    p = F.softmax(s, dim=-1)
    o = torch.matmul(p, v).transpose(1, 2).reshape(B, T, H)
    return F.linear(o, params[base + "out_proj.weight"], params[base + "out_proj.bias"])


def feed_forward(params, x, base):
    """HF:362-369: Linear H->4H, exact GELU, Linear 4H->H."""
    h = F.gelu(F.linear(x, params[base + "intermediate_dense.weight"], params[base + "intermediate_dense.bias"]))
    return F.linear(h, params[base + "output_dense.weight"], params[base + "output_dense.bias"])


def _ln(params, x, base, eps):
    return F.layer_norm(x, (x.shape[-1],), params[base + "weight"], params[base + "bias"], eps=eps)


def pos_conv_embed(params, cfg, x, prefix=""):
    """HF:83-92: grouped Conv1d k=128 pad=64, drop last step (HF:95-103), GELU."""
    w = fold_pos_conv_weight(params, prefix)
    k = cfg.num_conv_pos_embeddings
    y = F.conv1d(x.transpose(1, 2), w, params[prefix + "encoder.pos_conv_embed.conv.bias"],
                 padding=k // 2, groups=cfg.num_conv_pos_embedding_groups)
    if k % 2 == 0:
        y = y[:, :, :-1]
    return F.gelu(y).transpose(1, 2)


def hubert_forward(params, cfg, input_values, attention_mask=None, prefix="", taps=None,
                   output_layer=None):
    """HF:889-958 `HubertModel.forward` in eval mode -> last_hidden_state [B, T, H].

    `output_layer` (1-based, fairseq semantics used by I_da hubert_feature_reader.py:60-65)
    stops after that many transformer layers; None / -1 => all layers."""
    eps = cfg.layer_norm_eps
    feats = feature_encoder(params, cfg, input_values, prefix, taps).transpose(1, 2)  # [B,T,512]
    key_mask = None
    if attention_mask is not None:
        key_mask = feature_vector_attention_mask(cfg, feats.shape[1], attention_mask)
    # feature projection HF:225-231
    h = _ln(params, feats, prefix + "feature_projection.layer_norm.", eps)
    h = F.linear(h, params[prefix + "feature_projection.projection.weight"],
                 params[prefix + "feature_projection.projection.bias"])
    if taps is not None:
        taps["proj"] = h.clone()
    # _mask_hidden_states is a no-op in eval / with mask_time_prob=0 (I_ea/model.py:58-63)
    if key_mask is not None:
        h = h * key_mask[:, :, None]  # HF:429-432 zero padded frames
        if bool(key_mask.all()):
            key_mask_attn = None  # create_bidirectional_mask returns None when nothing is padded
        else:
            key_mask_attn = key_mask
    else:
        key_mask_attn = None
    h = h + pos_conv_embed(params, cfg, h, prefix)  # HF:440-441
    if not cfg.do_stable_layer_norm:
        h = _ln(params, h, prefix + "encoder.layer_norm.", eps)  # HF:442
    if taps is not None:
        taps["enc_in"] = h.clone()
    n_layers = cfg.num_hidden_layers
    if output_layer is not None and output_layer > 0:
        n_layers = min(n_layers, output_layer)
    for l in range(n_layers):
        b = f"{prefix}encoder.layers.{l}."
        if cfg.do_stable_layer_norm:  # HF:525-548 pre-LN
            a = attention(params, cfg, _ln(params, h, b + "layer_norm.", eps), b + "attention.", key_mask_attn)
            h = h + a
            h = h + feed_forward(params, _ln(params, h, b + "final_layer_norm.", eps), b + "feed_forward.")
        else:  # HF:388-405 post-LN
            a = attention(params, cfg, h, b + "attention.", key_mask_attn)
            h = _ln(params, h + a, b + "layer_norm.", eps)
            h = _ln(params, h + feed_forward(params, h, b + "feed_forward."), b + "final_layer_norm.", eps)
        if taps is not None:
            taps[f"layer{l}"] = h.clone()
    if cfg.do_stable_layer_norm and n_layers == cfg.num_hidden_layers:
        h = _ln(params, h, prefix + "encoder.layer_norm.", eps)  # HF:613
    return h


def custom_model_forward(params, cfg, input_values, attention_mask=None, taps=None):
    """I_ea/model.py:80-89 `CustomModel.forward`: HubertModel -> LayerNorm(H) -> Linear(H, 80).
    Keys: `base_model.*` and `final_layers.{0,1}.*`."""
    h = hubert_forward(params, cfg, input_values, attention_mask, prefix="base_model.", taps=taps)
    h = F.layer_norm(h, (h.shape[-1],), params["final_layers.0.weight"], params["final_layers.0.bias"], eps=1e-5)
    return F.linear(h, params["final_layers.1.weight"], params["final_layers.1.bias"])


def get_feats(params, cfg, signal, normalize: bool, layer: int, max_chunk: int = 1_600_000):
    """I_da/src/hubert_feature_reader.py:44-67 `HubertFeatureReader.get_feats`:
    optional whole-utterance F.layer_norm (eps 1e-5), 1.6 M-sample chunks,
    extract_features(padding_mask=None, mask=False, output_layer=layer) -> [T, H]."""
    x = torch.as_tensor(signal, dtype=torch.float32)
    if normalize:
        x = F.layer_norm(x, x.shape)
    x = x.view(1, -1)
    feat = []
    for start in range(0, x.size(1), max_chunk):
        feat.append(hubert_forward(params, cfg, x[:, start:start + max_chunk], None, output_layer=layer))
    return torch.cat(feat, 1).squeeze(0)
