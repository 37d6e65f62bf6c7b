"""Frozen f0 VQ-VAE encoder + nearest-bin quantiser of the I_da decoder.

`CodeGenerator.forward` runs `self.fo_vqvae.encoder(fo)` and `self.fo_vqvae.vq(h_p)[0]` to turn the continuous
fundamental-frequency track [B, 1, L] into pitch bins z_p [B, L/16] (I_da/src/model.py:148-153):
  encoder = Jukebox `EncoderConvBlock` (I_da/src/modules/jukebox.py:11-113), for the shipped config
            (configs/VCTK/hubert_lut.json:36-49: width 32, depth 4, downs_t [4], strides_t [2], growth 3):
            4 x [Conv1d(k 4, s 2, p 1) -> Resnet1D: 4 x (x + Conv1d_1x1(relu(Conv1d_k3,dil 3^i(relu(x)))))],
            then Conv1d(32 -> 128, k 3, p 1)          (I_da/src/modules/resnet.py:30-96)
  vq      = `BottleneckBlock.quantise`: argmin_k ||h - k_k||^2 over l_bins = 20 rows (I_da/src/modules/vq.py:118-128)
All convolutions run on `sib_conv1d_f32` (fp32: the output is an integer index and must match exactly), the assignment
on `sib_l2_argmin_f32`.  State-dict keys are the reference's (`encoder.level_blocks.0.model.*`, `vq.level_blocks.0.k`,
as found under `fo_vqvae.` in a CodeGenerator checkpoint or under `ckpt['generator']` of the f0 VQ-VAE itself;
`decoder.*` is training-only and ignored).
"""
from __future__ import annotations

import torch

from . import ops
from .module import SibModule
from .ops import SibError


class F0Quantizer(SibModule):
    def __init__(self, f0_quantizer: dict):
        super().__init__()
        enc = dict(f0_quantizer["f0_encoder_params"])
        vq = dict(f0_quantizer["f0_vq_params"])
        if int(enc.get("levels", 1)) != 1 or int(vq.get("levels", 1)) != 1:
            raise SibError("F0Quantizer: only the single-level f0 VQ-VAE of the shipped configs is supported")
        down, stride = enc["downs_t"], enc["strides_t"]
        down = down[0] if isinstance(down, (list, tuple)) and len(down) == 1 else down
        stride = stride[0] if isinstance(stride, (list, tuple)) and len(stride) == 1 else stride
        if isinstance(down, (list, tuple)) or isinstance(stride, (list, tuple)):
            raise SibError("F0Quantizer: per-block stride lists are not used by the shipped configs")
        self.down_t, self.stride_t = int(down), int(stride)
        self.in_width, self.out_width = int(enc["input_emb_width"]), int(enc["output_emb_width"])
        self.width, self.depth = int(enc["width"]), int(enc["depth"])
        self.state = int(float(enc.get("m_conv", 1.0)) * self.width)
        self.growth = int(enc.get("dilation_growth_rate", 1))
        self.cycle = enc.get("dilation_cycle", None)
        self.res_scale = 1.0 if not enc.get("res_scale", False) else 1.0 / (self.depth ** 0.5)
        self.l_bins, self.emb_width = int(vq["l_bins"]), int(vq["emb_width"])
        if self.emb_width != self.out_width:
            raise SibError("F0Quantizer: encoder output width must equal the codebook width")
        self.hop = self.stride_t ** self.down_t
        for name, shape in self._conv_shapes().items():
            self._add_param(name + ".weight", torch.empty(shape).normal_(0.0, 0.02))
            self._add_param(name + ".bias", torch.zeros(shape[0]))
        self._add_param("vq.level_blocks.0.k", torch.randn(self.l_bins, self.emb_width))

    def _conv_shapes(self):
        """Conv1d weight shapes of the Jukebox encoder (jukebox.py:89-110, resnet.py:36-52)."""
        shapes = {}
        for i in range(self.down_t):
            n = f"encoder.level_blocks.0.model.{i}"
            shapes[n + ".0"] = (self.width, self.in_width if i == 0 else self.width, 2 * self.stride_t)
            for d in range(self.depth):
                shapes[f"{n}.1.model.{d}.model.1"] = (self.state, self.width, 3)
                shapes[f"{n}.1.model.{d}.model.3"] = (self.width, self.state, 1)
        shapes[f"encoder.level_blocks.0.model.{self.down_t}"] = (self.out_width, self.width, 3)
        return shapes

    def _dilation(self, d):
        return self.growth ** (d if self.cycle is None else d % int(self.cycle))

    def _conv_names(self):
        names = []
        for i in range(self.down_t):
            names.append(f"encoder.level_blocks.0.model.{i}.0")
            for d in range(self.depth):
                names += [f"encoder.level_blocks.0.model.{i}.1.model.{d}.model.1",
                          f"encoder.level_blocks.0.model.{i}.1.model.{d}.model.3"]
        return names + [f"encoder.level_blocks.0.model.{self.down_t}"]

    def _expected_keys(self):
        return [n + s for n in self._conv_names() for s in (".weight", ".bias")] + ["vq.level_blocks.0.k"]

    def _adapt_state_dict(self, sd):
        return {k: v for k, v in sd.items() if not k.startswith("decoder.")}   # training-only half of the VQ-VAE

    def _pack(self):
        if self._packed is None:
            self._require_cuda()
            self._packed = {n: ops.pack_conv_weight(self._sd[n + ".weight"]) for n in self._conv_names()}
        return self._packed

    def encode_features(self, f0):
        """f0 [B, in_width, L] (reference layout) -> encoder output, frame-major [B, L / hop, out_width]."""
        self._require_cuda()
        if f0.dim() != 3 or f0.shape[1] != self.in_width:
            raise SibError(f"F0Quantizer: f0 must be [B, {self.in_width}, L], got {tuple(f0.shape)}")
        dev = self._device
        with torch.cuda.device(dev):
            return self._encode_features(f0, dev)

    def _encode_features(self, f0, dev):
        P, sd = self._pack(), self._sd
        x = f0.to(dev, torch.float32)
        x = x.transpose(1, 2).contiguous() if self.in_width > 1 else x.reshape(x.shape[0], x.shape[2], 1).contiguous()
        B, L, _ = x.shape
        k, s, p = 2 * self.stride_t, self.stride_t, self.stride_t // 2
        for i in range(self.down_t):
            n = f"encoder.level_blocks.0.model.{i}"
            t_out = (L + 2 * p - k) // s + 1
            if t_out <= 0:
                raise SibError(f"F0Quantizer: f0 series of {f0.shape[2]} frames is too short for {self.down_t} stride-{s} stages")
            y = torch.empty(B, t_out, self.width, device=dev, dtype=torch.float32)
            ops.conv1d(x, P[n + ".0"], sd[n + ".0.bias"], y, [j - p for j in range(k)], stride=s)   # jukebox.py:89-96
            x, L = y, t_out
            for d in range(self.depth):                                                            # resnet.py:36-52
                r, dl = f"{n}.1.model.{d}.model", self._dilation(d)
                h = torch.empty(B, L, self.state, device=dev, dtype=torch.float32)
                ops.conv1d(x, P[r + ".1"], sd[r + ".1.bias"], h, ops.conv_taps(3, dl, dl), pre_slope=0.0)   # ReLU -> k3
                if self.res_scale != 1.0:
                    raise SibError("F0Quantizer: res_scale is not used by the shipped configs")
                y = torch.empty_like(x)
                ops.conv1d(h, P[r + ".3"], sd[r + ".3.bias"], y, [0], pre_slope=0.0, residual=x)            # ReLU -> 1x1, + x
                x = y
        n = f"encoder.level_blocks.0.model.{self.down_t}"
        out = torch.empty(B, L, self.out_width, device=dev, dtype=torch.float32)
        ops.conv1d(x, P[n], sd[n + ".bias"], out, ops.conv_taps(3, 1, 1))                                   # jukebox.py:110
        return out

    def encode(self, f0):
        """f0 [B, 1, L] -> pitch bins z_p int64 [B, L / hop] (model.py:148-152)."""
        h = self.encode_features(f0)
        B, T, W = h.shape
        with torch.cuda.device(h.device):
            z = torch.empty(B * T, device=h.device, dtype=torch.int64)
            ops.l2_argmin(h.view(B * T, W), self._sd["vq.level_blocks.0.k"], z)
        return z.view(B, T)

    forward = encode
