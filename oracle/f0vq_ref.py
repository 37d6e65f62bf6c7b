"""CPU restatement of the frozen f0 VQ-VAE encoder + quantiser inside I_da's CodeGenerator (test infrastructure only).

Follows I_da/src/model.py:148-152 (`fo_vqvae.encoder(fo)` then `fo_vqvae.vq(h_p)[0]`), with the encoder of
I_da/src/modules/jukebox.py:11-113 + resnet.py:30-96 and `BottleneckBlock.quantise` of I_da/src/modules/vq.py:118-128, for
the shipped configuration (I_da/configs/VCTK/hubert_lut.json:36-49).  Pinned by oracle/make_golden.py against the
reference's own `Encoder` module and `BottleneckBlock.quantise` (tests/golden/f0vq_golden.npz).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

F0_QUANTIZER = {   # I_da/configs/VCTK/hubert_lut.json:36-49
    "f0_vq_params": {"l_bins": 20, "emb_width": 128, "mu": 0.99, "levels": 1},
    "f0_encoder_params": {"input_emb_width": 1, "output_emb_width": 128, "levels": 1, "downs_t": [4], "strides_t": [2],
                          "width": 32, "depth": 4, "m_conv": 1.0, "dilation_growth_rate": 3},
}


def make_params(cfg=F0_QUANTIZER, seed: int = 1234, scale: float = 1.0):
    """Seeded state dict with the reference's key names (PyTorch-default-like uniform init, codebook ~ N(0, 1))."""
    g = torch.Generator().manual_seed(seed)
    e, v = cfg["f0_encoder_params"], cfg["f0_vq_params"]
    w, depth, down, s = e["width"], e["depth"], e["downs_t"][0], e["strides_t"][0]
    sd = {}

    def conv(name, cout, cin, k):
        bound = scale / (cin * k) ** 0.5
        sd[name + ".weight"] = (torch.rand(cout, cin, k, generator=g) * 2 - 1) * bound
        sd[name + ".bias"] = (torch.rand(cout, generator=g) * 2 - 1) * bound

    for i in range(down):
        conv(f"encoder.level_blocks.0.model.{i}.0", w, e["input_emb_width"] if i == 0 else w, 2 * s)
        for d in range(depth):
            conv(f"encoder.level_blocks.0.model.{i}.1.model.{d}.model.1", w, w, 3)
            conv(f"encoder.level_blocks.0.model.{i}.1.model.{d}.model.3", w, w, 1)
    conv(f"encoder.level_blocks.0.model.{down}", e["output_emb_width"], w, 3)
    sd["vq.level_blocks.0.k"] = torch.randn(v["l_bins"], v["emb_width"], generator=g) * 0.3
    return sd


def encoder_forward(sd, f0: torch.Tensor, cfg=F0_QUANTIZER) -> torch.Tensor:
    """f0 [B, 1, L] -> h [B, 128, L/16]."""
    e = cfg["f0_encoder_params"]
    s, growth = e["strides_t"][0], e["dilation_growth_rate"]
    x = f0
    for i in range(e["downs_t"][0]):
        n = f"encoder.level_blocks.0.model.{i}"
        x = F.conv1d(x, sd[n + ".0.weight"], sd[n + ".0.bias"], stride=s, padding=s // 2)
        for d in range(e["depth"]):
            r, dil = f"{n}.1.model.{d}.model", growth ** d
            h = F.conv1d(F.relu(x), sd[r + ".1.weight"], sd[r + ".1.bias"], padding=dil, dilation=dil)
            x = x + F.conv1d(F.relu(h), sd[r + ".3.weight"], sd[r + ".3.bias"])
    n = f"encoder.level_blocks.0.model.{e['downs_t'][0]}"
    return F.conv1d(x, sd[n + ".weight"], sd[n + ".bias"], padding=1)


def quantise(h: torch.Tensor, k: torch.Tensor) -> torch.Tensor:
    """vq.py:98-101 (NCT -> [N*T, C]) + :118-128: distance = |x|^2 - 2 x.k^T + |k|^2, argmin over bins -> [N, T]."""
    N, C, T = h.shape
    x = h.permute(0, 2, 1).reshape(-1, C)
    kw = k.t()
    dist = x.pow(2).sum(-1, keepdim=True) - 2 * (x @ kw) + kw.pow(2).sum(0, keepdim=True)
    return dist.argmin(dim=-1).view(N, T)


def f0_to_bins(sd, f0: torch.Tensor, cfg=F0_QUANTIZER) -> torch.Tensor:
    return quantise(encoder_forward(sd, f0, cfg), sd["vq.level_blocks.0.k"])
