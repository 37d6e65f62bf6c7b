"""HuBERT encoder on the B200 kernels, behind the reference's module surface.

Drop-in for:
  * HF `HubertModel.forward(input_values, attention_mask) -> .last_hidden_state`  HF:889-958
  * `CustomModel.forward` (HubertModel -> LayerNorm -> Linear(H,80))               I_ea/model.py:80-89
  * fairseq-style `extract_features(source, padding_mask, mask, output_layer)`       I_da/src/hubert_feature_reader.py:60-65
Checkpoint key names are the reference's (HF state dict, old or new weight-norm names).
All arithmetic runs in libsib_b200.so; there is no PyTorch fallback.
"""
from __future__ import annotations

import os

from dataclasses import dataclass
from types import SimpleNamespace

import torch

from . import ops
from .module import SibModule
from .ops import ACT_GELU, ACT_NONE, Plan, SibError


@dataclass
class HubertConfig:
    """The HF `HubertConfig` fields that change the arithmetic (same names, same defaults)."""

    hidden_size: int = 768
    num_hidden_layers: int = 12
    num_attention_heads: int = 12
    intermediate_size: int = 3072
    feat_extract_norm: str = "group"
    conv_bias: bool = False
    do_stable_layer_norm: bool = False
    conv_dim: tuple = (512,) * 7
    conv_kernel: tuple = (10, 3, 3, 3, 3, 2, 2)
    conv_stride: tuple = (5, 2, 2, 2, 2, 2, 2)
    num_conv_pos_embeddings: int = 128
    num_conv_pos_embedding_groups: int = 16
    layer_norm_eps: float = 1e-5
    # SpecAugment knobs the reference zeroes (I_ea/model.py:58-63); eval-only path ignores them
    mask_time_prob: float = 0.0
    mask_feature_prob: float = 0.0
    mask_feature_length: int = 0
    mask_feature_min_masks: int = 0
    mask_time_length: int = 0
    mask_time_min_masks: int = 0

    @staticmethod
    def base():
        return HubertConfig()

    @staticmethod
    def large():
        return HubertConfig(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16,
                            intermediate_size=4096, feat_extract_norm="layer", conv_bias=True,
                            do_stable_layer_norm=True)

    @classmethod
    def from_any(cls, cfg):
        """Accepts this class, an HF HubertConfig or any object/dict with the same attribute names."""
        if isinstance(cls, type) and isinstance(cfg, cls):
            return cfg
        get = (lambda k, d: cfg.get(k, d)) if isinstance(cfg, dict) else (lambda k, d: getattr(cfg, k, d))
        base = cls()
        kw = {f: get(f, getattr(base, f)) for f in base.__dataclass_fields__}
        for f in ("conv_dim", "conv_kernel", "conv_stride"):
            kw[f] = tuple(kw[f])
        return cls(**kw)

    def feat_extract_output_length(self, n: int) -> int:
        """HF:675-688."""
        for k, s in zip(self.conv_kernel, self.conv_stride):
            n = (n - k) // s + 1
        return n


def expected_hubert_keys(cfg: HubertConfig, prefix: str = "", new_wn_names: bool = True):
    keys = []
    for i in range(len(cfg.conv_dim)):
        b = f"feature_extractor.conv_layers.{i}."
        keys.append(b + "conv.weight")
        if cfg.conv_bias:
            keys.append(b + "conv.bias")
        if cfg.feat_extract_norm == "layer" or i == 0:
            keys += [b + "layer_norm.weight", b + "layer_norm.bias"]
    keys += ["feature_projection.layer_norm.weight", "feature_projection.layer_norm.bias",
             "feature_projection.projection.weight", "feature_projection.projection.bias",
             "encoder.pos_conv_embed.conv.bias", "encoder.layer_norm.weight", "encoder.layer_norm.bias"]
    if new_wn_names:
        keys += ["encoder.pos_conv_embed.conv.parametrizations.weight.original0",
                 "encoder.pos_conv_embed.conv.parametrizations.weight.original1"]
    else:
        keys += ["encoder.pos_conv_embed.conv.weight_g", "encoder.pos_conv_embed.conv.weight_v"]
    for l in range(cfg.num_hidden_layers):
        b = f"encoder.layers.{l}."
        for n in ("attention.q_proj", "attention.k_proj", "attention.v_proj", "attention.out_proj", "layer_norm",
                  "feed_forward.intermediate_dense", "feed_forward.output_dense", "final_layer_norm"):
            keys += [b + n + ".weight", b + n + ".bias"]
    keys.append("masked_spec_embed")
    return [prefix + k for k in keys]


def hubert_param_shapes(cfg: HubertConfig, prefix: str = ""):
    """Shape of every tensor in an HF `HubertModel.state_dict()` for this config (HF:106-231, 45-92, 262-405)."""
    H, F = cfg.hidden_size, cfg.intermediate_size
    K, G = cfg.num_conv_pos_embeddings, cfg.num_conv_pos_embedding_groups
    shapes, cin = {}, 1
    for i, (c, k) in enumerate(zip(cfg.conv_dim, cfg.conv_kernel)):
        b = f"feature_extractor.conv_layers.{i}."
        shapes[b + "conv.weight"] = (c, cin, k)
        if cfg.conv_bias:
            shapes[b + "conv.bias"] = (c,)
        if cfg.feat_extract_norm == "layer" or i == 0:
            shapes[b + "layer_norm.weight"] = shapes[b + "layer_norm.bias"] = (c,)
        cin = c
    shapes["feature_projection.layer_norm.weight"] = shapes["feature_projection.layer_norm.bias"] = (cin,)
    shapes["feature_projection.projection.weight"], shapes["feature_projection.projection.bias"] = (H, cin), (H,)
    shapes["encoder.pos_conv_embed.conv.bias"] = (H,)
    shapes["encoder.pos_conv_embed.conv.parametrizations.weight.original0"] = (1, 1, K)
    shapes["encoder.pos_conv_embed.conv.parametrizations.weight.original1"] = (H, H // G, K)
    shapes["encoder.layer_norm.weight"] = shapes["encoder.layer_norm.bias"] = (H,)
    for l in range(cfg.num_hidden_layers):
        b = f"encoder.layers.{l}."
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            shapes[b + f"attention.{n}.weight"], shapes[b + f"attention.{n}.bias"] = (H, H), (H,)
        for n in ("layer_norm", "final_layer_norm"):
            shapes[b + n + ".weight"] = shapes[b + n + ".bias"] = (H,)
        shapes[b + "feed_forward.intermediate_dense.weight"], shapes[b + "feed_forward.intermediate_dense.bias"] = (F, H), (F,)
        shapes[b + "feed_forward.output_dense.weight"], shapes[b + "feed_forward.output_dense.bias"] = (H, F), (H,)
    shapes["masked_spec_embed"] = (H,)
    return {prefix + k: v for k, v in shapes.items()}


def _init_param(name: str, shape, gen=None) -> torch.Tensor:
    """Placeholder initialisation until a checkpoint is loaded (the reference scripts always load one): norm scales 1,
    biases 0, weights N(0, 0.02) - HF:640-673 in spirit; weight-norm gains are set by the caller."""
    leaf = name.rsplit(".", 1)[-1]
    if leaf == "bias":
        return torch.zeros(shape)
    if leaf == "weight" and len(shape) == 1:
        return torch.ones(shape)
    return torch.empty(shape).normal_(0.0, 0.02, generator=gen)


_OLD_WN = (("parametrizations.weight.original0", "weight_g"), ("parametrizations.weight.original1", "weight_v"))


def _to_new_wn_names(sd: dict) -> dict:
    """Old-style weight-norm names of the pos-conv (`weight_g` / `weight_v`, torch < 2.1 checkpoints) -> the
    parametrization names the module registers (SURVEY 8b accepts both)."""
    out = {}
    for k, v in sd.items():
        if "pos_conv_embed.conv." in k:
            for new, old in _OLD_WN:
                if k.endswith("conv." + old):
                    k = k[: -len(old)] + new
        out[k] = v
    return out


class HubertModel(SibModule):
    """HF `HubertModel` surface over the sm_100a kernels (eval mode, SpecAugment off)."""

    def __init__(self, config=None, precision: str = "fp32", key_prefix: str = ""):
        super().__init__()
        self.config = HubertConfig.from_any(config) if config is not None else HubertConfig()
        self.precision = precision
        self._prefix = key_prefix
        self.use_cuda_graph = False
        for name, shape in hubert_param_shapes(self.config).items():
            self._add_param(name, _init_param(name, shape))
        with torch.no_grad():   # weight-norm gain of the pos-conv starts at ||v|| (HF:59-78)
            b = "encoder.pos_conv_embed.conv.parametrizations.weight."
            self._sd[b + "original0"].copy_(self._sd[b + "original1"].pow(2).sum(dim=(0, 1), keepdim=True).sqrt())
        self._sd_cache = None
        # independent half-batch chains through the transformer stack (ops.Plan.chain).  Measured on B200 at 32 x 4 s: two chains
        # 11.15-11.33 ms per step against 10.97-11.04 for one (the persistent GEMMs own every SM, so the second chain only
        # adds smaller, less efficient tiles) - kept as an A/B switch, off by default
        self.transformer_chains = int(os.environ.get("SIB_HUBERT_CHAINS", "1")) if precision == "bf16" else 1
        # bf16 arm, OFF by default: the LayerNorms of the transformer loop folded into the neighbouring linear layers
        # (sib_linear_ln_bf16: no LayerNorm launch per layer, the normalised tensors are never materialised).  Measured on
        # B200 at 32 x 4 s, three same-box alternations: 10.55-10.65 ms per step with the fold against 10.31-10.36 without -
        # the GEMM epilogues are what paces the short HuBERT GEMMs, so moving the normalisation into them costs more than the
        # 24 small LayerNorm launches it removes; it also makes a row's statistics depend on the N tiling the cost model picks
        # for the batch size, which breaks the bit-exact shard invariance.  Kept as a tested option (SIB_HUBERT_FOLD_LN=1).
        self.fold_layernorm = precision == "bf16" and os.environ.get("SIB_HUBERT_FOLD_LN", "0") != "0"
        # bf16 arm, OFF by default: tile-level dataflow between the launches of the transformer loop (ops.FlowChain / sib_flow)
        # instead of grid-wide kernel boundaries; bit-identical results.  1 = the GEMM / LayerNorm hand-offs, 2 = the attention
        # kernel's two boundaries as well.  Measured on B200, same-box alternations of level 2 against 0: 32 x 4 s HuBERT-base
        # 10.37-10.42 against 10.40-10.51 ms per step (-0.7 %), 64 x 4 s I_da 22.70-22.74 against 22.45-22.62 (+0.8 %),
        # HuBERT-large 128 x 6 s 89.2-89.4 against 88.6-89.3 (neutral): the overlap it buys (6-8 us per half layer) is about
        # what its release / acquire pairs cost, because programmatic dependent launch already hides most of every boundary
        # and a late-starting CTA still owns a full static share of tiles.  SIB_HUBERT_FLOW=1|2 enables it.
        self.flow = int(os.environ.get("SIB_HUBERT_FLOW", "0")) if precision == "bf16" else 0

    # ---- state
    def _expected_keys(self):
        return expected_hubert_keys(self.config, "")

    def _adapt_state_dict(self, sd):
        return _to_new_wn_names(sd)

    def _w(self, name):
        return self._sd[name]

    def _pos_conv_weight(self):
        """weight_norm(dim=2) folded on the device: w = v * (g / ||v||_(0,1))  (HF:59-78)."""
        b = "encoder.pos_conv_embed.conv.parametrizations.weight."
        return ops.weight_norm_fold(self._sd[b + "original1"], self._sd[b + "original0"].reshape(-1), dim=2)

    def _pack(self):
        """One-time weight packing into kernel layouts (fold weight-norm, transpose, fuse QKV)."""
        if self._packed is not None:
            return self._packed
        self._require_cuda()
        cfg, P = self.config, {}
        w0 = self._w("feature_extractor.conv_layers.0.conv.weight")
        P["conv0.w"] = w0.reshape(w0.shape[0], -1).contiguous()
        for i in range(len(cfg.conv_dim)):
            b = f"feature_extractor.conv_layers.{i}."
            if i > 0:
                P[f"conv{i}.w"] = ops.pack_conv_weight(self._w(b + "conv.weight"))
            P[f"conv{i}.b"] = self._sd.get(b + "conv.bias")
            if b + "layer_norm.weight" in self._sd:
                P[f"conv{i}.g"], P[f"conv{i}.beta"] = self._w(b + "layer_norm.weight"), self._w(b + "layer_norm.bias")
        P["proj.w"] = ops.pack_linear_weight(self._w("feature_projection.projection.weight"))
        P["pos.w"] = ops.pack_conv_weight(self._pos_conv_weight(), cfg.num_conv_pos_embedding_groups)
        for l in range(cfg.num_hidden_layers):
            b = f"encoder.layers.{l}.attention."
            qkv_w = torch.cat([self._w(b + "q_proj.weight"), self._w(b + "k_proj.weight"), self._w(b + "v_proj.weight")], 0)
            P[f"l{l}.qkv.w"] = ops.pack_linear_weight(qkv_w)
            P[f"l{l}.qkv.b"] = torch.cat([self._w(b + "q_proj.bias"), self._w(b + "k_proj.bias"), self._w(b + "v_proj.bias")]).contiguous()
            P[f"l{l}.o.w"] = ops.pack_linear_weight(self._w(b + "out_proj.weight"))
            f = f"encoder.layers.{l}.feed_forward."
            P[f"l{l}.ff1.w"] = ops.pack_linear_weight(self._w(f + "intermediate_dense.weight"))
            P[f"l{l}.ff2.w"] = ops.pack_linear_weight(self._w(f + "output_dense.weight"))
        if self.precision == "bf16":
            if self.fold_layernorm:
                self._pack_folded(P)
            # tcgen05 operand layout (K-major bf16); conv0 / norms / biases stay fp32
            for k in [k for k in P if k.endswith(".w") and k != "conv0.w"]:
                P[k] = ops.to_kmajor_bf16(P[k])
        elif self.precision != "fp32":
            raise SibError(f"unknown precision {self.precision!r} (fp32 | bf16)")
        self._packed = P
        self._plans.clear()
        return P

    def _pack_folded(self, P):
        """Weights of the linear layers that consume a LayerNorm output, with that LayerNorm folded in (one-time):
        LN(t) W + b = r (t W' - mu s) + c,  W' = diag(gamma) W,  s = column sums of the bf16-rounded W',  c = beta W + b.
        post-LN (HF:388-405): QKV of layer l takes final_layer_norm of layer l-1, FFN-in takes layer_norm of layer l;
        pre-LN  (HF:525-548): QKV takes layer_norm, FFN-in final_layer_norm of the same layer."""
        cfg = self.config

        def fold(wkey, bias, ln_name):
            g, be = self._w(ln_name + ".weight"), self._w(ln_name + ".bias")
            w = P[wkey]                                        # [1][1][K][N] fp32
            wf = w * g.view(1, 1, -1, 1)
            P[wkey + "f"] = wf                                 # K-major bf16 below (key ends in ".wf" -> handled explicitly)
            P[wkey + "f.colsum"] = wf.to(torch.bfloat16).to(torch.float32).sum(dim=2).reshape(-1).contiguous()
            P[wkey + "f.c"] = (be.view(1, -1) @ w[0, 0]).reshape(-1).add(bias).contiguous()

        for l in range(cfg.num_hidden_layers):
            b = f"encoder.layers.{l}."
            f1b = self._w(b + "feed_forward.intermediate_dense.bias")
            if cfg.do_stable_layer_norm:
                fold(f"l{l}.qkv.w", P[f"l{l}.qkv.b"], b + "layer_norm")
                fold(f"l{l}.ff1.w", f1b, b + "final_layer_norm")
            else:
                if l > 0:
                    fold(f"l{l}.qkv.w", P[f"l{l}.qkv.b"], f"encoder.layers.{l - 1}.final_layer_norm")
                fold(f"l{l}.ff1.w", f1b, b + "layer_norm")
        for k in [k for k in P if k.endswith(".wf")]:
            P[k] = ops.to_kmajor_bf16(P[k])

    # ---- plan
    def _build_plan(self, B: int, N: int, padded: bool, n_layers: int):
        cfg, P, dev = self.config, self._pack(), self._device
        bf16 = self.precision == "bf16"
        f32 = dict(device=dev, dtype=torch.bfloat16 if bf16 else torch.float32)   # activation dtype of this plan
        real_f32 = dict(device=dev, dtype=torch.float32)

        def act_buf(b_, t_, c_):
            # + slack: the stride-s TMA view [B, ceil(T/s), s*C] may touch one row past the last frame
            flat = torch.zeros(b_ * t_ * c_ + 4 * c_, **f32)
            return flat[: b_ * t_ * c_].view(b_, t_, c_)
        lens = [N]
        for k, s in zip(cfg.conv_kernel, cfg.conv_stride):
            lens.append((lens[-1] - k) // s + 1)
        if lens[-1] < 1:
            raise SibError(f"input of {N} samples is shorter than the receptive field of the feature encoder")
        T, H, C = lens[-1], cfg.hidden_size, cfg.conv_dim[0]
        eps = cfg.layer_norm_eps
        io = SimpleNamespace(wave=torch.empty(B, N, **real_f32), key_len=torch.empty(B, dtype=torch.int32, device=dev) if padded else None)
        plan = Plan()
        fold = bf16 and self.fold_layernorm and self.transformer_chains <= 1 and n_layers > 0
        with plan.record():
            chain = None
            if bf16 and self.flow and not fold and self.transformer_chains <= 1 and n_layers > 0:
                chain = ops.FlowChain(B * T, 7 * n_layers, dev)
                chain.reset()                  # first launch of the plan: the counters are zero before anything signals
            # ---- feature encoder (HF:203-213)
            t0 = lens[1]
            a = act_buf(B, t0, C)
            k0, s0 = cfg.conv_kernel[0], cfg.conv_stride[0]
            if cfg.feat_extract_norm == "group":
                mean, rstd = torch.empty(B, C, **real_f32), torch.empty(B, C, **real_f32)
                if (k0, s0) == (10, 5):
                    # closed-form statistics from the waveform's lag sums: no [B, T0, C] evaluation pass
                    ops.conv0_gn_stats(io.wave, P["conv0.w"], P["conv0.b"], C, k0, s0, t0, 1e-5, mean, rstd)
                else:
                    nt = ops.conv0_num_tiles(t0)
                    part = torch.empty(B, nt, C, 2, **real_f32)
                    ops.conv0(0, io.wave, P["conv0.w"], P["conv0.b"], C, k0, s0, t0, partial=part)
                    ops.gn_finalize(part, B, nt, C, t0, 1e-5, mean, rstd)
                ops.conv0(1, io.wave, P["conv0.w"], P["conv0.b"], C, k0, s0, t0, mean=mean, rstd=rstd,
                          gamma=P["conv0.g"], beta=P["conv0.beta"], y=a)
            else:
                ops.conv0(2, io.wave, P["conv0.w"], P["conv0.b"], C, k0, s0, t0, y=a)
                ops.layernorm(a, P["conv0.g"], P["conv0.beta"], a, 1e-5, post_act=ACT_GELU)
            for i in range(1, len(cfg.conv_dim)):
                y = act_buf(B, lens[i + 1], cfg.conv_dim[i])
                k, s = cfg.conv_kernel[i], cfg.conv_stride[i]
                if cfg.feat_extract_norm == "layer":
                    ops.conv1d(a, P[f"conv{i}.w"], P[f"conv{i}.b"], y, list(range(k)), stride=s)
                    ops.layernorm(y, P[f"conv{i}.g"], P[f"conv{i}.beta"], y, 1e-5, post_act=ACT_GELU)
                else:
                    ops.conv1d(a, P[f"conv{i}.w"], P[f"conv{i}.b"], y, list(range(k)), stride=s, post_act=ACT_GELU)
                a = y
            # ---- feature projection (HF:225-231)
            ln = torch.empty(B, T, cfg.conv_dim[-1], **f32)
            ops.layernorm(a, self._w("feature_projection.layer_norm.weight"), self._w("feature_projection.layer_norm.bias"), ln, eps)
            h = torch.empty(B, T, H, **f32)
            ops.linear(ln.view(B * T, -1), P["proj.w"], self._w("feature_projection.projection.bias"), h.view(B * T, H))
            if padded:
                ops.zero_padded_frames(h, io.key_len)  # HF:429-432
            # ---- positional conv embedding (HF:83-92, 440-442)
            kp = cfg.num_conv_pos_embeddings
            taps = [j - kp // 2 for j in range(kp)]  # pad k//2; the dropped last step is simply not computed
            h2 = torch.empty(B, T, H, **f32)
            tmp = torch.empty(B, T, H, **f32)
            if cfg.do_stable_layer_norm:
                ops.conv1d(h, P["pos.w"], self._w("encoder.pos_conv_embed.conv.bias"), h2, taps,
                           groups=cfg.num_conv_pos_embedding_groups, post_act=ACT_GELU, residual=h, res_after_act=True)
            else:
                ops.conv1d(h, P["pos.w"], self._w("encoder.pos_conv_embed.conv.bias"), tmp, taps,
                           groups=cfg.num_conv_pos_embedding_groups, post_act=ACT_GELU)
                ops.layernorm(tmp, self._w("encoder.layer_norm.weight"), self._w("encoder.layer_norm.bias"), h2, eps, residual=h)
            h = h2
            # ---- transformer layers
            qkv = torch.empty(B, T, 3 * H, **f32)
            # with the attention kernel inside the dataflow chain, the projection of layer l + 1 may start on early row blocks
            # while a late utterance of layer l still reads its keys / values: the two layers' projections alternate buffers
            qkv_alt = torch.empty(B, T, 3 * H, **f32) if chain is not None else qkv
            att = torch.empty(B, T, H, **f32)
            ff = torch.empty(B, T, cfg.intermediate_size, **f32)
            nrm = torch.empty(B, T, H, **f32)
            # Two independent half-batch chains through the transformer stack (every op is row- or utterance-wise): the
            # GEMMs of 32 x 199 frames are 1.0 - 4.1 waves of tiles, so a single chain leaves most SMs idle during each
            # kernel's last wave; the other chain's kernel fills them (ops.Plan.chain).  Same buffers, disjoint row ranges.
            if fold:
                h = self._record_layers_folded(P, h, qkv, att, ff, tmp, nrm, io.key_len, eps, n_layers)
            halves = [(0, B)]
            if fold:
                halves = []
            elif self.transformer_chains > 1 and B >= 2 * self.transformer_chains and not padded:
                nc = self.transformer_chains
                halves = [(B * c // nc, B * (c + 1) // nc) for c in range(nc)]
                plan.fork()
            # Tile-level dataflow through the loop (ops.FlowChain / sib_flow): out-proj -> LN -> FFN-in -> FFN-out -> LN -> QKV
            # hand their rows over per 128-row block, so the tail of one kernel overlaps the head of the next.
            edge = None
            for l in range(n_layers if halves else 0):   # layer-major order: the host feeds both streams alternately
                for c, (b0, b1) in enumerate(halves):
                    with plan.chain(c):
                        edge = self._record_layer(l, P, h[b0:b1], (qkv if l % 2 == 0 else qkv_alt)[b0:b1], att[b0:b1], ff[b0:b1], nrm[b0:b1], tmp[b0:b1],
                                                  None if io.key_len is None else io.key_len[b0:b1], eps, chain=chain, edge_in=edge,
                                                  last=l == n_layers - 1)
            if len(halves) > 1:
                plan.join()
            if cfg.do_stable_layer_norm and n_layers == cfg.num_hidden_layers:
                ops.layernorm(h, self._w("encoder.layer_norm.weight"), self._w("encoder.layer_norm.bias"), h, eps)  # HF:613
            if bf16:
                out32 = torch.empty(B, T, H, **real_f32)
                ops.cast_to_f32(h, out32)
                h = out32
        io.out = h
        io.T = T
        io.plan = plan
        if self.use_cuda_graph:
            plan.capture()
        return io

    def _record_layers_folded(self, P, h, qkv, att, ff, t1, t2, key_len, eps, n_layers):
        """The transformer loop with every LayerNorm folded into its neighbours (`ops.linear_ln`): the normalised tensors are
        never written.  post-LN: the stream alternates between two RAW tensors t1 (attention block output + residual) and t2
        (FFN output + residual), each with the partial row statistics its producer leaves; LN(t) is applied on the fly -
        as two per-row scalars in the epilogue of the linear layer that consumes it, and rebuilt from the residual tile where
        it is the residual.  pre-LN: the stream h stays raw anyway; the producers only add the statistics.
        Returns the tensor holding the loop's result (materialised by one last LayerNorm launch for post-LN)."""
        cfg = self.config
        B, T, H = h.shape
        M = B * T
        dev = h.device
        st_a = torch.zeros(M, ops.LN_SLOTS, 2, device=dev, dtype=torch.float32)   # rows of t1 / of h after the attention block
        st_b = torch.zeros(M, ops.LN_SLOTS, 2, device=dev, dtype=torch.float32)   # rows of t2 / of h after the FFN block
        v2 = lambda t: t.view(M, -1)   # noqa: E731
        prev = None                    # post-LN: (raw tensor, its statistics, gamma, beta) of the layer input
        for l in range(n_layers):
            b = f"encoder.layers.{l}."
            ob, f2b = self._w(b + "attention.out_proj.bias"), self._w(b + "feed_forward.output_dense.bias")
            ln1 = (self._w(b + "layer_norm.weight"), self._w(b + "layer_norm.bias"))
            ln2 = (self._w(b + "final_layer_norm.weight"), self._w(b + "final_layer_norm.bias"))
            if cfg.do_stable_layer_norm:   # HF:525-548
                if l == 0:                 # nobody has produced the statistics of h yet: one real LayerNorm
                    ops.layernorm(h, ln1[0], ln1[1], t1, eps)
                    ops.linear(v2(t1), P[f"l{l}.qkv.w"], P[f"l{l}.qkv.b"], v2(qkv))
                else:
                    ops.linear_ln(v2(h), P[f"l{l}.qkv.wf"], P[f"l{l}.qkv.wf.c"], v2(qkv), mode="apply", n_norm=H, eps=eps,
                                  stats_in=st_b, colsum=P[f"l{l}.qkv.wf.colsum"])
                ops.attention(qkv, key_len, att, cfg.num_attention_heads)
                ops.linear_ln(v2(att), P[f"l{l}.o.w"], ob, v2(h), mode="residual", n_norm=H, eps=eps, residual=v2(h), stats_out=st_a)
                ops.linear_ln(v2(h), P[f"l{l}.ff1.wf"], P[f"l{l}.ff1.wf.c"], v2(ff), mode="apply", n_norm=H, eps=eps,
                              stats_in=st_a, colsum=P[f"l{l}.ff1.wf.colsum"], post_act=ACT_GELU)
                ops.linear_ln(v2(ff), P[f"l{l}.ff2.w"], f2b, v2(h), mode="residual", n_norm=H, eps=eps, residual=v2(h), stats_out=st_b)
                continue
            # post-LN, HF:388-405
            if prev is None:
                ops.linear(v2(h), P[f"l{l}.qkv.w"], P[f"l{l}.qkv.b"], v2(qkv))
                res_kw = dict(residual=v2(h))
            else:
                pt, pst, pg, pb = prev
                ops.linear_ln(v2(pt), P[f"l{l}.qkv.wf"], P[f"l{l}.qkv.wf.c"], v2(qkv), mode="apply", n_norm=H, eps=eps,
                              stats_in=pst, colsum=P[f"l{l}.qkv.wf.colsum"])
                res_kw = dict(residual=v2(pt), stats_in=pst, gamma=pg, beta=pb)
            ops.attention(qkv, key_len, att, cfg.num_attention_heads)
            ops.linear_ln(v2(att), P[f"l{l}.o.w"], ob, v2(t1), mode="residual", n_norm=H, eps=eps, stats_out=st_a, **res_kw)
            ops.linear_ln(v2(t1), P[f"l{l}.ff1.wf"], P[f"l{l}.ff1.wf.c"], v2(ff), mode="apply", n_norm=H, eps=eps,
                          stats_in=st_a, colsum=P[f"l{l}.ff1.wf.colsum"], post_act=ACT_GELU)
            ops.linear_ln(v2(ff), P[f"l{l}.ff2.w"], f2b, v2(t2), mode="residual", n_norm=H, eps=eps, residual=v2(t1),
                          stats_in=st_a, gamma=ln1[0], beta=ln1[1], stats_out=st_b)
            prev = (t2, st_b, ln2[0], ln2[1])
        if cfg.do_stable_layer_norm:
            return h
        pt, _, pg, pb = prev
        ops.layernorm(pt, pg, pb, h, eps)          # the layer stack's output is the one LayerNorm that must exist in memory
        return h

    def _record_layer(self, l, P, h, qkv, att, ff, nrm, tmp, key_len, eps, chain=None, edge_in=None, last=False):
        """Transformer layer l (HF:388-405 post-LN / HF:525-548 pre-LN) on a contiguous batch range of the plan's buffers.
        With a `chain` (ops.FlowChain) every hand-off except the two around the attention kernel is per 128-row block;
        `edge_in` is the edge the previous layer's last launch signals, the return value this layer's."""
        cfg = self.config
        Bc, T, H = h.shape
        M = Bc * T
        b = f"encoder.layers.{l}."
        ob, f2b = self._w(b + "attention.out_proj.bias"), self._w(b + "feed_forward.output_dense.bias")
        f1b = self._w(b + "feed_forward.intermediate_dense.bias")
        ln1 = (self._w(b + "layer_norm.weight"), self._w(b + "layer_norm.bias"))
        ln2 = (self._w(b + "final_layer_norm.weight"), self._w(b + "final_layer_norm.bias"))
        if chain is not None:
            # Buffer reuse under dataflow (no grid-wide barrier any more), checked per 128-row block m:
            #  * every linear layer / LayerNorm of the loop reads and writes block m only, and its wait on block m of its producer
            #    implies (transitively, release / acquire is cumulative) that all earlier launches have finished block m - so
            #    tmp / nrm / ff / h are overwritten only after their last reader of that block is done; the in-place residual
            #    updates of the pre-LN stack read and write the same tile;
            #  * the attention kernel reads whole utterances: it waits for every block an utterance touches, and the out-projection
            #    waits for the attention rows of its block, i.e. for all launches of the previous layers on the blocks of the
            #    utterances in it.  The one buffer that is NOT safe is qkv: a late utterance of layer l may still read keys in a
            #    block that layer l + 1's projection (gated only on its own block) may overwrite - the caller alternates two buffers.
            I = ff.shape[-1]
            v2 = lambda t: t.view(M, -1)   # noqa: E731
            e1, e3 = chain.edge("linear", H), chain.edge("linear", I)
            e2 = chain.edge("layernorm", H)
            eq, eo = (chain.edge("linear", 3 * H), chain.edge("attention", H)) if int(self.flow) >= 2 else (None, None)
            heads = cfg.num_attention_heads
            if cfg.do_stable_layer_norm:  # HF:525-548; h is updated in place by the two residual GEMMs
                ea = chain.edge("layernorm", H)
                e4 = chain.edge("linear", H)
                ops.layernorm(h, ln1[0], ln1[1], nrm, eps, wait=edge_in, signal=ea)
                ops.linear(v2(nrm), P[f"l{l}.qkv.w"], P[f"l{l}.qkv.b"], v2(qkv), wait=ea, signal=eq)
                ops.attention(qkv, key_len, att, heads, wait=eq, signal=eo)
                ops.linear(v2(att), P[f"l{l}.o.w"], ob, v2(h), residual=v2(h), wait=eo, signal=e1)
                ops.layernorm(h, ln2[0], ln2[1], nrm, eps, wait=e1, signal=e2)
                ops.linear(v2(nrm), P[f"l{l}.ff1.w"], f1b, v2(ff), post_act=ACT_GELU, wait=e2, signal=e3)
                # the loop's last launch hands over to a plain launch (grid-wide wait): no counter needed
                ops.linear(v2(ff), P[f"l{l}.ff2.w"], f2b, v2(h), residual=v2(h), wait=e3, signal=None if last else e4)
                return e4
            e4, e5 = chain.edge("linear", H), chain.edge("layernorm", H)
            ops.linear(v2(h), P[f"l{l}.qkv.w"], P[f"l{l}.qkv.b"], v2(qkv), wait=edge_in, signal=eq)
            ops.attention(qkv, key_len, att, heads, wait=eq, signal=eo)
            ops.linear(v2(att), P[f"l{l}.o.w"], ob, v2(tmp), wait=eo, signal=e1)
            ops.layernorm(tmp, ln1[0], ln1[1], nrm, eps, residual=h, wait=e1, signal=e2)
            ops.linear(v2(nrm), P[f"l{l}.ff1.w"], f1b, v2(ff), post_act=ACT_GELU, wait=e2, signal=e3)
            ops.linear(v2(ff), P[f"l{l}.ff2.w"], f2b, v2(tmp), wait=e3, signal=e4)
            ops.layernorm(tmp, ln2[0], ln2[1], h, eps, residual=nrm, wait=e4, signal=None if last else e5)
            return e5
        if cfg.do_stable_layer_norm:  # HF:525-548
            ops.layernorm(h, ln1[0], ln1[1], nrm, eps)
            ops.linear(nrm.view(M, H), P[f"l{l}.qkv.w"], P[f"l{l}.qkv.b"], qkv.view(M, 3 * H))
            ops.attention(qkv, key_len, att, cfg.num_attention_heads)
            ops.linear(att.view(M, H), P[f"l{l}.o.w"], ob, h.view(M, H), residual=h.view(M, H))
            ops.layernorm(h, ln2[0], ln2[1], nrm, eps)
            ops.linear(nrm.view(M, H), P[f"l{l}.ff1.w"], f1b, ff.view(M, -1), post_act=ACT_GELU)
            ops.linear(ff.view(M, -1), P[f"l{l}.ff2.w"], f2b, h.view(M, H), residual=h.view(M, H))
        else:  # HF:388-405
            ops.linear(h.view(M, H), P[f"l{l}.qkv.w"], P[f"l{l}.qkv.b"], qkv.view(M, 3 * H))
            ops.attention(qkv, key_len, att, cfg.num_attention_heads)
            ops.linear(att.view(M, H), P[f"l{l}.o.w"], ob, tmp.view(M, H))
            ops.layernorm(tmp, ln1[0], ln1[1], nrm, eps, residual=h)
            ops.linear(nrm.view(M, H), P[f"l{l}.ff1.w"], f1b, ff.view(M, -1), post_act=ACT_GELU)
            ops.linear(ff.view(M, -1), P[f"l{l}.ff2.w"], f2b, tmp.view(M, H))
            ops.layernorm(tmp, ln2[0], ln2[1], h, eps, residual=nrm)

    def _key_len(self, attention_mask, N):
        """HF:690-700: valid frames per utterance from the sample-level attention mask."""
        lens = attention_mask.sum(-1).to(torch.int64)
        out = lens.clone()
        for k, s in zip(self.config.conv_kernel, self.config.conv_stride):
            out = torch.div(out - k, s, rounding_mode="floor") + 1
        return out.to(torch.int32)

    def _run(self, input_values, attention_mask=None, n_layers=None):
        self._require_cuda()
        if input_values.dim() != 2:
            raise SibError(f"input_values must be [batch, samples], got {tuple(input_values.shape)}")
        B, N = input_values.shape
        L = self.config.num_hidden_layers if n_layers is None else n_layers
        padded = attention_mask is not None and not bool(attention_mask.to(torch.bool).all())
        key = (B, N, padded, L)
        with torch.cuda.device(self._device):   # the C side launches on the current device: make it this module's
            self._pack()
            io = self._cached_plan(key, lambda: self._build_plan(B, N, padded, L))
            io.wave.copy_(input_values.to(self._device, torch.float32), non_blocking=True)
            if padded:
                io.key_len.copy_(self._key_len(attention_mask.to(self._device), N))
            io.plan.run()
        return io

    # ---- reference surface
    def forward(self, input_values, attention_mask=None, **_unused):
        io = self._run(input_values, attention_mask)
        return SimpleNamespace(last_hidden_state=io.out.clone())   # fresh tensor, as the reference returns

    def extract_features(self, source, padding_mask=None, mask=False, output_layer=None):
        """fairseq HuBERT surface used by I_da (hubert_feature_reader.py:60-65); output_layer is 1-based,
        None / -1 / 0 => all layers.  fairseq-vs-HF equivalence: parity unpinned (fairseq absent)."""
        if mask:
            raise SibError("mask=True (SpecAugment) is a training feature; the reference calls mask=False")
        n_layers = None if (output_layer is None or output_layer <= 0) else min(output_layer, self.config.num_hidden_layers)
        am = None if padding_mask is None else (~padding_mask.to(torch.bool)).to(torch.int64)
        io = self._run(source, am, n_layers)
        return io.out.clone(), padding_mask


class CustomModel(SibModule):
    """I_ea/model.py:22-89 `CustomModel`: HubertModel + final_layers = LayerNorm(H) -> Linear(H, codebook_dim).
    Constructed from a config instead of `from_pretrained` (no network on the path); state-dict keys are
    `base_model.*` and `final_layers.{0,1}.*` (211 tensors for HuBERT-base)."""

    def __init__(self, codebook_dim=80, type="large", load_pretrained=False, train_encoder=False, loss_function="",
                 config=None, precision: str = "fp32"):
        super().__init__()
        if config is None:
            config = HubertConfig.base() if type == "base" else HubertConfig.large()
        self.base_model = HubertModel(config, precision=precision)
        self.last_hidden_dim = self.base_model.config.hidden_size
        self.codebook_dim = 100 if loss_function == "softmax" else codebook_dim
        H = self.last_hidden_dim
        self._add_param("final_layers.0.weight", torch.ones(H))
        self._add_param("final_layers.0.bias", torch.zeros(H))
        self._add_param("final_layers.1.weight", torch.empty(self.codebook_dim, H).normal_(0.0, 0.02))
        self._add_param("final_layers.1.bias", torch.zeros(self.codebook_dim))
        self._head = None

    def _expected_keys(self):
        return (expected_hubert_keys(self.base_model.config, "base_model.") +
                ["final_layers.0.weight", "final_layers.0.bias", "final_layers.1.weight", "final_layers.1.bias"])

    def _adapt_state_dict(self, sd):
        return _to_new_wn_names(sd)

    def _head_weight(self):
        if self._head is None:
            self._head = ops.pack_linear_weight(self._sd["final_layers.1.weight"])
        return self._head

    def forward_frames(self, input_values, attention_mask, pos_t, len_t, off_t, n_rows: int):
        """final_layers applied ONLY to the frames [pos_b, pos_b + len_b) of every utterance: what predict.py:164-168
        gathers out of the full `outputs` anyway.  LayerNorm and Linear are row-wise, so the gathered rows are
        bit-identical to gather(forward(...)); the head runs on sum(len) rows instead of B*T.  Returns ([n_rows, D], T)."""
        io = self.base_model._run(input_values, attention_mask)
        h = io.out
        B, T, H = h.shape
        sd = self._sd
        with torch.cuda.device(h.device):
            rows = torch.empty(max(n_rows, 1), H, device=h.device, dtype=torch.float32)[:n_rows]
            out = torch.empty(max(n_rows, 1), self.codebook_dim, device=h.device, dtype=torch.float32)[:n_rows]
            if n_rows > 0:
                ops.gather_frames(h, pos_t, len_t, off_t, rows)
                nrm = torch.empty_like(rows)
                ops.layernorm(rows, sd["final_layers.0.weight"], sd["final_layers.0.bias"], nrm, 1e-5)
                if self.codebook_dim <= 128 and H <= 8192:   # a few hundred rows x 80 outputs: one CTA per row
                    ops.linear_skinny(nrm, self._head_weight(), sd["final_layers.1.bias"], out)
                else:
                    ops.linear(nrm, self._head_weight(), sd["final_layers.1.bias"], out)
        return out, T

    def forward(self, input_values, attention_mask=None):
        io = self.base_model._run(input_values, attention_mask)
        h = io.out
        B, T, H = h.shape
        sd = self._sd
        with torch.cuda.device(h.device):
            nrm = torch.empty_like(h)
            ops.layernorm(h, sd["final_layers.0.weight"], sd["final_layers.0.bias"], nrm, 1e-5)
            out = torch.empty(B, T, self.codebook_dim, device=h.device, dtype=torch.float32)
            ops.linear(nrm.view(B * T, H), self._head_weight(), sd["final_layers.1.bias"], out.view(B * T, -1))
        return out
