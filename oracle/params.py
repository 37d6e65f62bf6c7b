"""Seeded synthetic weights with the reference's checkpoint key names.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference ships no weights
(HF-hub links only, README.md:22), and north_star prescribes random-init weights,
so both the oracle and the CUDA path are fed from these deterministic state
dicts (CPU torch.Generator => identical on every box with the same torch).

Key names follow SURVEY.md section 8b:
  HuBERT  - HF `HubertModel.state_dict()` (HF:178-213, 216-231, 45-92, 372-405)
  head    - `final_layers.{0,1}.*`           (I_ea/model.py:75-78)
  HiFi-GAN- `conv_pre|ups.i|resblocks.n.convs{1,2}.m|conv_post`.{bias,weight_g,weight_v}
            (I_ea/hifi_gan/models.py:82-105, checkpoints are saved *with*
            weight-norm, predict.py:119-122)
  I_da    - same + `emb_c.weight`, `emb_p.weight`, `emb_s.weight` (I_da/src/model.py:48-76)
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import torch


@dataclass
class HubertCfg:
    """Subset of HF HubertConfig that changes the arithmetic (HF configuration_hubert.py)."""

    hidden_size: int = 768
    num_hidden_layers: int = 12
    num_attention_heads: int = 12
    intermediate_size: int = 3072
    feat_extract_norm: str = "group"  # "group" (base) | "layer" (large)
    conv_bias: bool = False
    do_stable_layer_norm: bool = False
    conv_dim: tuple = (512,) * 7
    conv_kernel: tuple = (10, 3, 3, 3, 3, 2, 2)
    conv_stride: tuple = (5, 2, 2, 2, 2, 2, 2)
    num_conv_pos_embeddings: int = 128
    num_conv_pos_embedding_groups: int = 16
    layer_norm_eps: float = 1e-5

    @staticmethod
    def base() -> "HubertCfg":
        return HubertCfg()

    @staticmethod
    def large() -> "HubertCfg":
        # public facebook/hubert-large-ls960-ft shape
        return HubertCfg(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16,
                         intermediate_size=4096, feat_extract_norm="layer", conv_bias=True,
                         do_stable_layer_norm=True)

    @staticmethod
    def tiny(stable: bool = False) -> "HubertCfg":
        """Small shape for fast CPU tests; same code paths as base/large."""
        return HubertCfg(hidden_size=128, num_hidden_layers=2, num_attention_heads=2,
                         intermediate_size=256, feat_extract_norm="layer" if stable else "group",
                         conv_bias=stable, do_stable_layer_norm=stable,
                         conv_dim=(64,) * 7, num_conv_pos_embeddings=16,
                         num_conv_pos_embedding_groups=4)

    def feat_lengths(self, n):
        """HF:675-688 `_get_feat_extract_output_lengths`."""
        for k, s in zip(self.conv_kernel, self.conv_stride):
            n = (n - k) // s + 1
        return n


@dataclass
class HifiCfg:
    """HiFi-GAN generator hyper-parameters (I_ea/hifi_gan/config_v1.json:11-15,
    I_da/configs/VCTK/hubert_lut.json:13-20)."""

    upsample_rates: tuple = (8, 8, 2, 2)
    upsample_kernel_sizes: tuple = (16, 16, 4, 4)
    upsample_initial_channel: int = 512
    resblock_kernel_sizes: tuple = (3, 7, 11)
    resblock_dilation_sizes: tuple = ((1, 3, 5), (1, 3, 5), (1, 3, 5))
    model_in_dim: int = 80
    resblock: str = "1"
    # I_da CodeGenerator front (model.py:48-76)
    num_embeddings: int = 0
    embedding_dim: int = 128
    f0_bins: int = 20

    @staticmethod
    def v1() -> "HifiCfg":
        return HifiCfg()

    @staticmethod
    def v2() -> "HifiCfg":
        """I_ea/hifi_gan/config_v2.json:11-15 (ResBlock1, 128 initial channels: stages of 64 / 32 / 16 / 8)."""
        return HifiCfg(upsample_initial_channel=128)

    @staticmethod
    def v3() -> "HifiCfg":
        """I_ea/hifi_gan/config_v3.json:2,11-15 (ResBlock2: two convs per block, models.py:52-73)."""
        return HifiCfg(upsample_rates=(8, 8, 4), upsample_kernel_sizes=(16, 16, 8), upsample_initial_channel=256,
                       resblock_kernel_sizes=(3, 5, 7), resblock_dilation_sizes=((1, 2), (2, 6), (3, 12)), resblock="2")

    @staticmethod
    def ida() -> "HifiCfg":
        return HifiCfg(upsample_rates=(5, 4, 4, 2, 2), upsample_kernel_sizes=(11, 8, 8, 4, 4),
                       model_in_dim=384, num_embeddings=500)

    @staticmethod
    def tiny(ida: bool = False) -> "HifiCfg":
        if ida:
            return HifiCfg(upsample_rates=(5, 4, 2), upsample_kernel_sizes=(11, 8, 4),
                           upsample_initial_channel=64, model_in_dim=48, num_embeddings=50,
                           embedding_dim=16)
        return HifiCfg(upsample_rates=(8, 2), upsample_kernel_sizes=(16, 4),
                       upsample_initial_channel=64, model_in_dim=80)

    @property
    def total_upsample(self) -> int:
        return math.prod(self.upsample_rates)

    def as_attrdict(self) -> dict:
        return dict(upsample_rates=list(self.upsample_rates),
                    upsample_kernel_sizes=list(self.upsample_kernel_sizes),
                    upsample_initial_channel=self.upsample_initial_channel,
                    resblock_kernel_sizes=list(self.resblock_kernel_sizes),
                    resblock_dilation_sizes=[list(d) for d in self.resblock_dilation_sizes],
                    model_in_dim=self.model_in_dim, resblock=self.resblock,
                    num_embeddings=self.num_embeddings, embedding_dim=self.embedding_dim)


def _randn(g, *shape, std=1.0):
    return torch.randn(*shape, generator=g, dtype=torch.float32) * std


def make_hubert_params(cfg: HubertCfg, seed: int = 1234, prefix: str = "") -> dict:
    """State dict with HF key names; init scales follow HF:640-673 `_init_weights`
    (Linear N(0,0.02), conv kaiming-normal, norm affine perturbed so that the
    affine terms are actually exercised)."""
    g = torch.Generator().manual_seed(seed)
    p = {}
    cin = 1
    for i, (c, k) in enumerate(zip(cfg.conv_dim, cfg.conv_kernel)):
        p[f"feature_extractor.conv_layers.{i}.conv.weight"] = _randn(g, c, cin, k, std=math.sqrt(2.0 / (cin * k)))
        if cfg.conv_bias:
            p[f"feature_extractor.conv_layers.{i}.conv.bias"] = _randn(g, c, std=0.05)
        if cfg.feat_extract_norm == "layer" or i == 0:
            p[f"feature_extractor.conv_layers.{i}.layer_norm.weight"] = 1.0 + _randn(g, c, std=0.1)
            p[f"feature_extractor.conv_layers.{i}.layer_norm.bias"] = _randn(g, c, std=0.1)
        cin = c
    H, F = cfg.hidden_size, cfg.intermediate_size
    p["feature_projection.layer_norm.weight"] = 1.0 + _randn(g, cin, std=0.1)
    p["feature_projection.layer_norm.bias"] = _randn(g, cin, std=0.1)
    p["feature_projection.projection.weight"] = _randn(g, H, cin, std=0.02)
    p["feature_projection.projection.bias"] = _randn(g, H, std=0.02)
    K, G = cfg.num_conv_pos_embeddings, cfg.num_conv_pos_embedding_groups
    v = _randn(g, H, H // G, K, std=2.0 * math.sqrt(1.0 / (K * H)))
    p["encoder.pos_conv_embed.conv.parametrizations.weight.original1"] = v
    # weight-norm dim=2: g has shape [1,1,K] (HF:78); start at ||v|| then perturb
    p["encoder.pos_conv_embed.conv.parametrizations.weight.original0"] = (
        v.pow(2).sum(dim=(0, 1), keepdim=True).sqrt() * (1.0 + _randn(g, 1, 1, K, std=0.1)))
    p["encoder.pos_conv_embed.conv.bias"] = _randn(g, H, std=0.02)
    p["encoder.layer_norm.weight"] = 1.0 + _randn(g, H, std=0.1)
    p["encoder.layer_norm.bias"] = _randn(g, H, std=0.1)
    for l in range(cfg.num_hidden_layers):
        b = f"encoder.layers.{l}."
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            p[b + f"attention.{n}.weight"] = _randn(g, H, H, std=0.02 * 2)
            p[b + f"attention.{n}.bias"] = _randn(g, H, std=0.02)
        p[b + "layer_norm.weight"] = 1.0 + _randn(g, H, std=0.1)
        p[b + "layer_norm.bias"] = _randn(g, H, std=0.1)
        p[b + "feed_forward.intermediate_dense.weight"] = _randn(g, F, H, std=0.02 * 2)
        p[b + "feed_forward.intermediate_dense.bias"] = _randn(g, F, std=0.02)
        p[b + "feed_forward.output_dense.weight"] = _randn(g, H, F, std=0.02 * 2)
        p[b + "feed_forward.output_dense.bias"] = _randn(g, H, std=0.02)
        p[b + "final_layer_norm.weight"] = 1.0 + _randn(g, H, std=0.1)
        p[b + "final_layer_norm.bias"] = _randn(g, H, std=0.1)
    p["masked_spec_embed"] = torch.rand(H, generator=g)
    return {prefix + k: v for k, v in p.items()}


def make_head_params(hidden: int, out_dim: int = 80, seed: int = 4321) -> dict:
    """`final_layers` = LayerNorm(H) -> Linear(H, codebook_dim)  (I_ea/model.py:75-78)."""
    g = torch.Generator().manual_seed(seed)
    return {
        "final_layers.0.weight": 1.0 + _randn(g, hidden, std=0.1),
        "final_layers.0.bias": _randn(g, hidden, std=0.1),
        "final_layers.1.weight": _randn(g, out_dim, hidden, std=1.0 / math.sqrt(hidden)),
        "final_layers.1.bias": _randn(g, out_dim, std=0.1),
    }


def _wn_pair(g, shape, std):
    """(weight_g, weight_v) for torch weight_norm(dim=0): g=[C0,1,1] (SURVEY 7 hard parts)."""
    v = _randn(g, *shape, std=std)
    norm = v.pow(2).sum(dim=(1, 2), keepdim=True).sqrt()
    gg = norm * (1.0 + 0.1 * _randn(g, shape[0], 1, 1))
    return gg, v


def make_generator_params(cfg: HifiCfg, seed: int = 1234, init: str = "unit") -> dict:
    """HiFi-GAN generator state dict *with* weight-norm tensors.

    init="reference": N(0, 0.01) as `init_weights` (I_ea/hifi_gan/utils.py:24-32) - outputs are
    tiny (|y| << 1).  init="unit": variance-preserving scales so every stage is O(1) and SNR is
    meaningful (SURVEY 8d asks for both)."""
    g = torch.Generator().manual_seed(seed)
    p = {}

    def put(name, shape, fan_in):
        std = 0.01 if init == "reference" else math.sqrt(1.0 / fan_in)
        p[name + ".weight_g"], p[name + ".weight_v"] = _wn_pair(g, shape, std)
        cout = shape[1] if name.startswith("ups.") else shape[0]
        p[name + ".bias"] = _randn(g, cout, std=0.01 if init == "reference" else 0.05)

    c0 = cfg.upsample_initial_channel
    put("conv_pre", (c0, cfg.model_in_dim, 7), cfg.model_in_dim * 7)
    ch = c0
    for i, (u, k) in enumerate(zip(cfg.upsample_rates, cfg.upsample_kernel_sizes)):
        cin, ch = c0 // (2 ** i), c0 // (2 ** (i + 1))
        # ConvTranspose1d weight is [Cin, Cout, k]; each output sample sees k/u taps
        put(f"ups.{i}", (cin, ch, k), cin * k / u)
        for j, (rk, dil) in enumerate(zip(cfg.resblock_kernel_sizes, cfg.resblock_dilation_sizes)):
            n = i * len(cfg.resblock_kernel_sizes) + j
            for m in range(len(dil)):
                # residual branches are kept small so the stack stays O(1)
                if cfg.resblock == "1":
                    put(f"resblocks.{n}.convs1.{m}", (ch, ch, rk), ch * rk)
                    put(f"resblocks.{n}.convs2.{m}", (ch, ch, rk), ch * rk * 4)
                else:   # ResBlock2: one conv per dilation (models.py:55-61)
                    put(f"resblocks.{n}.convs.{m}", (ch, ch, rk), ch * rk * 4)
    put("conv_post", (1, ch, 7), ch * 7)
    if cfg.num_embeddings:
        p["emb_c.weight"] = _randn(g, cfg.num_embeddings, cfg.embedding_dim)
        p["emb_p.weight"] = _randn(g, cfg.f0_bins, cfg.embedding_dim)
        p["emb_s.weight"] = _randn(g, 200, cfg.embedding_dim)
    return p


def fold_weight_norm(p: dict) -> dict:
    """`remove_weight_norm()` (models.py:125-132): weight = g * v / ||v||, norm over all dims but 0."""
    out = {}
    for k, v in p.items():
        if k.endswith(".weight_g"):
            base = k[: -len(".weight_g")]
            vv = p[base + ".weight_v"]
            out[base + ".weight"] = v * vv / vv.pow(2).sum(dim=(1, 2), keepdim=True).sqrt()
        elif k.endswith(".weight_v"):
            continue
        else:
            out[k] = v
    return out


def make_codebook(dim: int = 80, k: int = 100, seed: int = 77) -> torch.Tensor:
    """Synthetic k-means codebook C[dim, K] (ApplyKmeans.C layout, I_ea/dataset/km_label.py:10-34)."""
    g = torch.Generator().manual_seed(seed)
    return _randn(g, dim, k)
