"""CPU fp32 restatement of the HiFi-GAN generators on the hot path.

TEST INFRASTRUCTURE (see oracle/__init__.py).
  * I_ea  `Generator.forward`      I_ea/hifi_gan/models.py:107-123 (+ResBlock1 :36-43,
    ResBlock2 :66-71, get_padding utils.py:47-48)
  * I_da  `Generator.forward`      I_da/src/models.py:209-225 (same graph, conv_pre in-dim
    = model_in_dim)
  * I_da  `CodeGenerator.forward`  I_da/src/model.py:121-189 (`_upsample` :78-119)
`params` is a state dict with the reference key names, weight-norm either present
(`weight_g/_v`) or already removed (`weight`).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .params import fold_weight_norm

LRELU_SLOPE = 0.1  # models.py:9


def get_padding(kernel_size: int, dilation: int = 1) -> int:
    """I_ea/hifi_gan/utils.py:47-48."""
    return int((kernel_size * dilation - dilation) / 2)


def _resblock1(p, base, x, k, dils, taps=None):
    """models.py:36-43: 3 x (lrelu -> conv(k, d) -> lrelu -> conv(k, 1) -> + x)."""
    for m, d in enumerate(dils):
        xt = F.leaky_relu(x, LRELU_SLOPE)
        xt = F.conv1d(xt, p[f"{base}.convs1.{m}.weight"], p[f"{base}.convs1.{m}.bias"],
                      dilation=d, padding=get_padding(k, d))
        xt = F.leaky_relu(xt, LRELU_SLOPE)
        xt = F.conv1d(xt, p[f"{base}.convs2.{m}.weight"], p[f"{base}.convs2.{m}.bias"],
                      padding=get_padding(k, 1))
        x = xt + x
    return x


def _resblock2(p, base, x, k, dils):
    """models.py:66-71: 2 x (lrelu -> conv(k, d) -> + x)."""
    for m, d in enumerate(dils):
        xt = F.leaky_relu(x, LRELU_SLOPE)
        xt = F.conv1d(xt, p[f"{base}.convs.{m}.weight"], p[f"{base}.convs.{m}.bias"],
                      dilation=d, padding=get_padding(k, d))
        x = xt + x
    return x


def generator_forward(params, cfg, x, taps=None):
    """`Generator.forward`: x [B, in_dim, Tm] -> [B, 1, Tm * prod(upsample_rates)]."""
    p = fold_weight_norm(params)
    nk = len(cfg.resblock_kernel_sizes)
    x = F.conv1d(x, p["conv_pre.weight"], p["conv_pre.bias"], padding=3)
    if taps is not None:
        taps["conv_pre"] = x.transpose(1, 2).contiguous()
    for i, (u, k) in enumerate(zip(cfg.upsample_rates, cfg.upsample_kernel_sizes)):
        x = F.leaky_relu(x, LRELU_SLOPE)
        x = F.conv_transpose1d(x, p[f"ups.{i}.weight"], p[f"ups.{i}.bias"], stride=u, padding=(k - u) // 2)
        if taps is not None:
            taps[f"ups{i}"] = x.transpose(1, 2).contiguous()
        xs = None
        for j, (rk, dil) in enumerate(zip(cfg.resblock_kernel_sizes, cfg.resblock_dilation_sizes)):
            base = f"resblocks.{i * nk + j}"
            r = _resblock1(p, base, x, rk, dil) if cfg.resblock == "1" else _resblock2(p, base, x, rk, dil)
            xs = r if xs is None else xs + r
        x = xs / nk
        if taps is not None:
            taps[f"stage{i}"] = x.transpose(1, 2).contiguous()
    x = F.leaky_relu(x)  # NB default slope 0.01 (models.py:119)
    x = F.conv1d(x, p["conv_post.weight"], p["conv_post.bias"], padding=3)
    return torch.tanh(x)


def upsample_repeat(signal, max_frames: int):
    """I_da/src/model.py:78-119 `_upsample`: repeat each step max_frames // T times."""
    if signal.dim() == 2:
        signal = signal.unsqueeze(2)
    elif signal.dim() != 3:
        signal = signal.view(-1, 1, 1)
    b, c, t = signal.shape
    signal = signal.unsqueeze(3).repeat(1, 1, 1, max_frames // t)
    if (max_frames - signal.shape[2] * signal.shape[3]) // signal.shape[3] > 0:
        raise NotImplementedError("Padding condition signal - misalignment between condition features.")
    return signal.reshape(b, c, max_frames)


def code_generator_front(params, code, z_p, emb):
    """I_da/src/model.py:141-172 for the shipped config (f0_stats set, multispkr set,
    no code VQ): emb_c(code)^T (+) emb_p(z_p)^T repeat-upsampled (+) d-vector repeat-upsampled
    -> [B, 3*E, T].  `z_p` are the f0 VQ bin indices [B, T/4] (the frozen f0 VQ-VAE encoder
    that produces them, model.py:148-153, is SURVEY 8f "next" row 2, outside this path)."""
    emb_c = F.embedding(code, params["emb_c.weight"]).transpose(1, 2)
    emb_p = F.embedding(z_p, params["emb_p.weight"]).transpose(1, 2)
    if emb_c.shape[-1] < emb_p.shape[-1]:
        emb_c = upsample_repeat(emb_c, emb_p.shape[-1])
    else:
        emb_p = upsample_repeat(emb_p, emb_c.shape[-1])
    x = torch.cat([emb_c, emb_p], dim=1)
    emb_s = upsample_repeat(emb, x.shape[-1])
    return torch.cat([x, emb_s], dim=1)


def code_generator_forward(params, cfg, code, z_p, emb):
    """`CodeGenerator.forward(code=, f0=, emb=, spkr=)` with the f0 branch already quantised."""
    return generator_forward(params, cfg, code_generator_front(params, code, z_p, emb))


def to_int16(y):
    """`generate` I_da/src/dataset.py:241-243 / I_ea/predict.py:125-127:
    audio * 32768 -> numpy astype('int16') (C truncation toward zero)."""
    a = (y.squeeze() * 32768.0).cpu().numpy()
    return a.astype("int16")
