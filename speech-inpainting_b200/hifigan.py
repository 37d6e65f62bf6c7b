"""HiFi-GAN generators on the B200 kernels, behind the reference's module surface.

Drop-in for:
  * `Generator(h).forward(x[B, in_dim, Tm]) -> [B, 1, Tm*prod(rates)]`   I_ea/hifi_gan/models.py:76-132,
                                                                          I_da/src/models.py:156-233
  * `CodeGenerator(h).forward(code=, f0=, emb=, spkr=)`                  I_da/src/model.py:42-189
`load_state_dict` takes the reference checkpoints (`ckpt['generator']`) with weight-norm tensors
(`weight_g/weight_v`) or after `remove_weight_norm()` (`weight`).  Weight-norm folding, the poly-phase
re-layout of ConvTranspose1d and the transposition into kernel layouts happen once, on first use.
"""
from __future__ import annotations

import math
import os
from types import SimpleNamespace

import torch

from . import ops
from .module import SibModule
from .ops import ACT_NONE, ACT_TANH, Plan, SibError

LRELU_SLOPE = 0.1  # models.py:9


def get_padding(kernel_size: int, dilation: int = 1) -> int:
    """I_ea/hifi_gan/utils.py:47-48."""
    return int((kernel_size * dilation - dilation) / 2)


class AttrDict(dict):
    """I_ea/hifi_gan/env.py:5-8."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.__dict__ = self


def _hget(h, name, default=None):
    if isinstance(h, dict):
        return h.get(name, default)
    return getattr(h, name, default)


class Generator(SibModule):
    def __init__(self, h, precision: str = "fp32"):
        super().__init__()
        self.h = h
        self.precision = precision
        self.upsample_rates = list(_hget(h, "upsample_rates"))
        self.upsample_kernel_sizes = list(_hget(h, "upsample_kernel_sizes"))
        self.c0 = int(_hget(h, "upsample_initial_channel"))
        self.rb_kernels = list(_hget(h, "resblock_kernel_sizes"))
        self.rb_dilations = [list(d) for d in _hget(h, "resblock_dilation_sizes")]
        self.resblock = str(_hget(h, "resblock", "1"))
        # I_ea hard-codes 80 mel bins (models.py:86); I_da reads model_in_dim (src/models.py:171)
        self.in_dim = int(_hget(h, "model_in_dim", 80) or 80)
        self.num_kernels = len(self.rb_kernels)
        self.num_upsamples = len(self.upsample_rates)
        self.total_upsample = math.prod(self.upsample_rates)
        self._weight_norm_removed = False
        self.use_cuda_graph = False
        self.use_fused_resunits = True   # bf16 arm: fused ResBlock1 units on the narrow stages (A/B switch)
        # bf16 arm: unfused convs apply the leaky-relu that precedes them to their A tile in shared memory, so producers
        # stop writing an activated copy of every tensor (A/B switch, SIB_SMEM_PREACT=0)
        self.use_smem_preact = os.environ.get("SIB_SMEM_PREACT", "1") != "0"
        # bf16 arm: utterances per chunk when a ResBlock's unit chain is run chunk by chunk (0 = whole batch per launch)
        self.rb_chunk = int(os.environ.get("SIB_RB_CHUNK", "0"))
        # bf16 arm: the ResBlocks of a stage as concurrent chains on separate streams (A/B switch SIB_RB_STREAMS=0/1)
        self.rb_streams = os.environ.get("SIB_RB_STREAMS", "0") != "0"
        # parameters under the reference's checkpoint names, WITH weight-norm (weight_g / weight_v) as `Generator.__init__`
        # creates them (models.py:82-105): 234 tensors for V1
        for name, shape in self._conv_shapes().items():
            v = torch.empty(shape).normal_(0.0, 0.01)                       # init_weights, utils.py:24-32
            self._add_param(name + ".weight_v", v)
            self._add_param(name + ".weight_g", v.pow(2).sum(dim=(1, 2), keepdim=True).sqrt())
            self._add_param(name + ".bias", torch.zeros(shape[1] if name.startswith("ups.") else shape[0]))

    # ---- state
    def _conv_names(self):
        names = ["conv_pre"] + [f"ups.{i}" for i in range(self.num_upsamples)]
        for i in range(self.num_upsamples):
            for j in range(self.num_kernels):
                n = i * self.num_kernels + j
                if self.resblock == "1":
                    for m in range(len(self.rb_dilations[j])):
                        names += [f"resblocks.{n}.convs1.{m}", f"resblocks.{n}.convs2.{m}"]
                else:
                    names += [f"resblocks.{n}.convs.{m}" for m in range(len(self.rb_dilations[j]))]
        return names + ["conv_post"]

    def _conv_shapes(self):
        """torch weight shape of every conv: Conv1d [Cout, Cin, k], ConvTranspose1d [Cin, Cout, k] (models.py:82-105)."""
        shapes = {"conv_pre": (self.c0, self.in_dim, 7)}
        ch = self.c0
        for i, k in enumerate(self.upsample_kernel_sizes):
            cin, ch = self.c0 // (2 ** i), self.c0 // (2 ** (i + 1))
            shapes[f"ups.{i}"] = (cin, ch, k)
            for j, rk in enumerate(self.rb_kernels):
                n = i * self.num_kernels + j
                for m in range(len(self.rb_dilations[j])):
                    if self.resblock == "1":
                        shapes[f"resblocks.{n}.convs1.{m}"] = shapes[f"resblocks.{n}.convs2.{m}"] = (ch, ch, rk)
                    else:
                        shapes[f"resblocks.{n}.convs.{m}"] = (ch, ch, rk)
        shapes["conv_post"] = (1, ch, 7)
        return shapes

    def _extra_keys(self):
        return []

    def _expected_keys(self):
        keys = []
        for n in self._conv_names():
            keys += [n + ".bias", n + ".weight_g", n + ".weight_v"]
        return keys + self._extra_keys()

    def _adapt_state_dict(self, sd):
        folded = any(k.endswith(".weight") and not k.startswith(("emb_", "fo_vqvae.")) for k in sd)
        if folded and not self._weight_norm_removed:   # checkpoint saved after remove_weight_norm(): same structure here
            self._drop_weight_norm(fold=False)
        return sd

    def _drop_weight_norm(self, fold: bool):
        """Replace (weight_g, weight_v) by `weight`, as torch's remove_weight_norm does (156 tensors left for V1)."""
        sd = self._sd
        for n in self._conv_names():
            if not self._has_param(n + ".weight_v"):
                continue
            v, g = sd[n + ".weight_v"], sd[n + ".weight_g"]
            w = ops.weight_norm_fold(v, g.reshape(-1), dim=0) if fold else torch.empty_like(v)
            # torch registers `weight` first, then bias stays: key order is irrelevant to load_state_dict
            self._del_param(n + ".weight_g")
            self._del_param(n + ".weight_v")
            self._add_param(n + ".weight", w)
        self._weight_norm_removed = True

    def remove_weight_norm(self):
        """models.py:125-132: fold weight = v * (g / ||v||) (one `sib_weight_norm_fold_f32` launch per conv) and replace
        the weight-norm pair by the folded tensor, so that `state_dict()` afterwards is the reference's 156-key one."""
        if self._weight_norm_removed:
            return
        self._require_cuda()
        with torch.cuda.device(self._device):
            self._drop_weight_norm(fold=True)

    def _weight(self, name):
        """Folded conv weight: `weight` after remove_weight_norm(), else v * (g / ||v||) folded on the device."""
        sd = self._sd
        if name + ".weight" in sd:
            return sd[name + ".weight"]
        return ops.weight_norm_fold(sd[name + ".weight_v"], sd[name + ".weight_g"].reshape(-1), dim=0)

    def _pack(self):
        if self._packed is not None:
            return self._packed
        self._require_cuda()
        P = {}
        P["conv_pre.w"] = ops.pack_conv_weight(self._weight("conv_pre"))
        for i, (u, k) in enumerate(zip(self.upsample_rates, self.upsample_kernel_sizes)):
            w, b, offs = ops.pack_conv_transpose(self._weight(f"ups.{i}"), self._sd[f"ups.{i}.bias"], u, (k - u) // 2)
            P[f"ups.{i}.w"], P[f"ups.{i}.b"], P[f"ups.{i}.taps"] = w, b, offs
        for n in self._conv_names():
            if n.startswith("resblocks."):
                P[n + ".w"] = ops.pack_conv_weight(self._weight(n))
        wp = self._weight("conv_post")  # [1, C, 7] -> [k][C]
        P["conv_post.w"] = wp[0].t().contiguous()
        if self.precision == "bf16":
            for k in [k for k in P if k.endswith(".w") and k != "conv_post.w"]:
                P[k] = ops.to_kmajor_bf16(P[k])  # tcgen05 operand layout
        elif self.precision != "fp32":
            raise SibError(f"unknown precision {self.precision!r} (fp32 | bf16)")
        self._packed = P
        self._plans.clear()
        return P

    # ---- plan (bf16 / tcgen05 arm)
    def _build_plan_bf16(self, B: int, Tm: int, frame_major_in: bool):
        """Same graph as `_build_plan`, bf16 activations.  The TMA-fed MMA cannot transform its A operand on the fly, so a
        leaky_relu that precedes a conv (models.py:37,39,109,119) is either applied by the consumer to its landed A tile
        in shared memory (fused units; `sib_conv1d_bf16` in halo mode with `use_smem_preact`) or produced by the PREVIOUS
        kernel's epilogue as a second output lrelu(y) (`y_act`); convs1 write only lrelu(t1)."""
        P, dev = self._pack(), self._device
        f32 = dict(device=dev, dtype=torch.float32)
        b16 = dict(device=dev, dtype=torch.bfloat16)
        io = SimpleNamespace()
        io.x_cf = None if frame_major_in else torch.empty(B, self.in_dim, Tm, **f32)
        io.x = torch.empty(B, Tm, self.in_dim, **f32)
        chans = [self.c0 // (2 ** (i + 1)) for i in range(self.num_upsamples)]
        lens, L = [], Tm
        for u in self.upsample_rates:
            L *= u
            lens.append(L)
        big = max(l * c for l, c in zip(lens, chans))
        # the MRF's ResBlocks of a stage are independent chains of units (models.py:113-118); with `rb_streams` they are
        # recorded on separate plan chains (own scratch tensors each), so the tail of one chain's kernel overlaps the head
        # of another's instead of leaving SMs idle at every launch boundary
        n_ch = self.num_kernels if (self.rb_streams and self.num_kernels > 1) else 1
        pool = [torch.empty(B * big, **b16) for _ in range(11 + 5 * (n_ch - 1))]
        UP, UPA, PA, PAA, PB, PBA, T1, XS0, XS0A, XS1, XS1A = range(11)

        def view(i, L_, C_, ch=0):
            if ch > 0 and i in (PA, PAA, PB, PBA, T1):          # chain-private scratch
                i = 11 + 5 * (ch - 1) + (PA, PAA, PB, PBA, T1).index(i)
            return pool[i][: B * L_ * C_].view(B, L_, C_)

        plan = Plan()
        with plan.record():
            if not frame_major_in:
                ops.transpose(io.x_cf, io.x)
            xb = torch.empty(B, Tm, self.in_dim, **b16)
            ops.cast_to_bf16(io.x, xb)
            cur_act = torch.empty(B, Tm, self.c0, **b16)
            ops.conv1d(xb, P["conv_pre.w"], self._sd["conv_pre.bias"], cur_act, ops.conv_taps(7, 1, 3),
                       post_act=ops.ACT_LRELU, post_slope=LRELU_SLOPE)
            t_in = Tm
            # Which consumers activate their own input?  Fused units do (raw x in); an unfused conv does when the tcgen05
            # kernel can apply the leaky-relu to its A tile in shared memory (halo mode); conv_post has a pre-slope of its
            # own.  Producers write an activated copy (y_act) only for the consumers that are left.
            pre = self.use_smem_preact
            c_prev = [self.c0] + chans[:-1]
            t_prev = [Tm] + lens[:-1]
            # (queried for the larger launch - with the activated second output - so that the answer holds either way)
            ups_pa = [pre and i > 0 and ops.conv_pre_act_supported(B, t_prev[i], c_prev[i], u * chans[i], P[f"ups.{i}.taps"],
                                                                   LRELU_SLOPE, has_y_act=True)
                      for i, u in enumerate(self.upsample_rates)]
            cur, cur_act = None, cur_act     # conv_pre wrote lrelu(y) only (nobody needs its raw output)
            for i, u in enumerate(self.upsample_rates):
                C_, L_ = chans[i], lens[i]
                last_stage = i == self.num_upsamples - 1
                next_needs_act = (not pre) if last_stage else (not ups_pa[i + 1])
                up, up_act = view(UP, L_, C_), view(UPA, L_, C_)
                xs, xs_act = (view(XS0, L_, C_), view(XS0A, L_, C_)) if i % 2 == 0 else (view(XS1, L_, C_), view(XS1A, L_, C_))
                # Narrow stages (C <= 64) run fused ResBlock1 units (sib_resunit_bf16: lrelu -> conv1 -> lrelu -> conv2 ->
                # + x in one kernel, raw x in, y out) wherever the unit's weights fit in shared memory.
                fused, self_act = [], []
                for j, (rk, dils) in enumerate(zip(self.rb_kernels, self.rb_dilations)):
                    row, srow = [False] * len(dils), [False] * len(dils)
                    for m in reversed(range(len(dils))):
                        last_m = m == len(dils) - 1
                        # a unit writes lrelu(y) only for a consumer that cannot activate its own input
                        need_act = (j == self.num_kernels - 1 and next_needs_act) if last_m else (not srow[m + 1])
                        row[m] = (self.resblock == "1" and self.use_fused_resunits and
                                  ops.resunit_supported(C_, rk, dils[m], last_m and j > 0, need_act))
                        # ResBlock1's first conv has neither residual nor second output; a ResBlock2 conv has both
                        rb2 = self.resblock != "1"
                        srow[m] = row[m] or (pre and ops.conv_pre_act_supported(
                            B, L_, C_, C_, ops.conv_taps(rk, dils[m], get_padding(rk, dils[m])), LRELU_SLOPE,
                            has_residual=rb2, has_y_act=rb2))
                    fused.append(row)
                    self_act.append(srow)
                need_up_act = not all(sa[0] for sa in self_act)
                if ups_pa[i]:   # lrelu(0.1) of models.py:110 on the raw output of the previous stage, in shared memory
                    ops.conv1d(cur, P[f"ups.{i}.w"], P[f"ups.{i}.b"], up.view(B, t_in, u * C_), P[f"ups.{i}.taps"],
                               pre_slope=LRELU_SLOPE, y_act=up_act.view(B, t_in, u * C_) if need_up_act else None,
                               act2_slope=LRELU_SLOPE)
                else:
                    ops.conv1d(cur_act, P[f"ups.{i}.w"], P[f"ups.{i}.b"], up.view(B, t_in, u * C_), P[f"ups.{i}.taps"],
                               y_act=up_act.view(B, t_in, u * C_) if need_up_act else None, act2_slope=LRELU_SLOPE)
                # A ResBlock's chain of units runs per CHUNK of utterances (`rb_chunk`): a chunk's tensors (45 MB for 8
                # utterances of the 32 x 4 s workload) then stay in the 126 MB L2 from one unit to the next - the conv1 -> conv2
                # intermediate of the unfused units and the unit -> unit hand-offs never make the HBM round trip.
                cb = self.rb_chunk if self.rb_chunk and self.rb_chunk > 0 else B
                chunks = [(b0, min(b0 + cb, B)) for b0 in range(0, B, cb)]
                if n_ch > 1:
                    plan.fork()
                for j, (rk, dils) in enumerate(zip(self.rb_kernels, self.rb_dilations)):
                    n = i * self.num_kernels + j
                    last_j = j == self.num_kernels - 1
                    chn = j if n_ch > 1 else 0
                    for b0, b1 in chunks:
                      with plan.chain(chn):
                        sl = lambda t: None if t is None else t[b0:b1]   # noqa: E731
                        xcur, xcur_act = sl(up), (sl(up_act) if need_up_act else None)
                        for m, dl in enumerate(dils):
                            last_m = m == len(dils) - 1
                            if last_m:
                                dst, dst_act = sl(xs), (sl(xs_act) if (last_j and next_needs_act) else None)
                                # the next consumer of xs is lrelu(0.1)->ups[i+1] or lrelu(0.01)->conv_post
                                slope2 = 0.01 if last_stage else LRELU_SLOPE
                            else:
                                dst, dst_act = ((sl(view(PA, L_, C_, chn)), sl(view(PAA, L_, C_, chn))) if m % 2 == 0
                                                else (sl(view(PB, L_, C_, chn)), sl(view(PBA, L_, C_, chn))))
                                slope2 = LRELU_SLOPE
                                if self_act[j][m + 1]:
                                    dst_act = None      # the consumer activates its own input tile
                            acc = last_m and j > 0
                            scale = (1.0 / self.num_kernels) if (last_m and last_j) else 1.0
                            if acc and n_ch > 1:          # the running MRF sum: after the previous ResBlock's last unit
                                plan.wait(("xs", i, j - 1, b0))
                            if fused[j][m]:
                                ops.resunit(xcur, P[f"resblocks.{n}.convs1.{m}.w"], self._sd[f"resblocks.{n}.convs1.{m}.bias"],
                                            P[f"resblocks.{n}.convs2.{m}.w"], self._sd[f"resblocks.{n}.convs2.{m}.bias"],
                                            dst, rk, dl, y_act=dst_act, accumulate=acc, out_scale=scale,
                                            slope_in=LRELU_SLOPE, slope_mid=LRELU_SLOPE, act2_slope=slope2)
                            else:
                                kw = dict(accumulate=acc, out_scale=scale, residual=xcur, y_act=dst_act, act2_slope=slope2)
                                # first conv of the unit: raw x + in-kernel leaky-relu, or the producer's activated copy
                                src, pkw = (xcur, dict(pre_slope=LRELU_SLOPE)) if self_act[j][m] else (xcur_act, {})
                                if self.resblock == "1":
                                    t1 = sl(view(T1, L_, C_, chn))
                                    ops.conv1d(src, P[f"resblocks.{n}.convs1.{m}.w"], self._sd[f"resblocks.{n}.convs1.{m}.bias"],
                                               t1, ops.conv_taps(rk, dl, get_padding(rk, dl)), post_act=ops.ACT_LRELU,
                                               post_slope=LRELU_SLOPE, **pkw)
                                    ops.conv1d(t1, P[f"resblocks.{n}.convs2.{m}.w"], self._sd[f"resblocks.{n}.convs2.{m}.bias"],
                                               dst, ops.conv_taps(rk, 1, get_padding(rk, 1)), **kw)
                                else:
                                    ops.conv1d(src, P[f"resblocks.{n}.convs.{m}.w"], self._sd[f"resblocks.{n}.convs.{m}.bias"],
                                               dst, ops.conv_taps(rk, dl, get_padding(rk, dl)), **kw, **pkw)
                            if last_m and n_ch > 1 and not last_j:
                                plan.signal(("xs", i, j, b0))
                            xcur, xcur_act = dst, dst_act
                if n_ch > 1:
                    plan.join()
                cur, cur_act = xs, (xs_act if next_needs_act else None)
                t_in = L_
            io.y = torch.empty(B, 1, lens[-1], **f32)
            if pre:   # lrelu(0.01) of models.py:119 as conv_post's own pre-activation on the raw stage output
                ops.conv1d_cout1(cur, P["conv_post.w"], self._sd["conv_post.bias"], io.y.view(B, lens[-1]), 7, 3, 0.01, ACT_TANH)
            else:     # cur_act already holds lrelu(x, 0.01): identity pre-activation (slope 1)
                ops.conv1d_cout1(cur_act, P["conv_post.w"], self._sd["conv_post.bias"], io.y.view(B, lens[-1]), 7, 3, 1.0, ACT_TANH)
        io.plan = plan
        if self.use_cuda_graph:
            plan.capture()
        return io

    # ---- plan (fp32 arm)
    def _build_plan(self, B: int, Tm: int, frame_major_in: bool):
        if self.precision == "bf16":
            return self._build_plan_bf16(B, Tm, frame_major_in)
        P, dev = self._pack(), self._device
        f32 = dict(device=dev, dtype=torch.float32)
        io = SimpleNamespace()
        io.x_cf = None if frame_major_in else torch.empty(B, self.in_dim, Tm, **f32)
        io.x = torch.empty(B, Tm, self.in_dim, **f32)
        chans = [self.c0 // (2 ** (i + 1)) for i in range(self.num_upsamples)]
        lens, L = [], Tm
        for u in self.upsample_rates:
            L *= u
            lens.append(L)
        big = max(l * c for l, c in zip(lens, chans))
        pool = [torch.empty(B * big, **f32) for _ in range(6)]

        def view(i, L_, C_):
            return pool[i][: B * L_ * C_].view(B, L_, C_)

        plan = Plan()
        with plan.record():
            if not frame_major_in:
                ops.transpose(io.x_cf, io.x)  # [B,C,T] -> [B,T,C]
            cur = torch.empty(B, Tm, self.c0, **f32)
            ops.conv1d(io.x, P["conv_pre.w"], self._sd["conv_pre.bias"], cur, ops.conv_taps(7, 1, 3))  # models.py:108
            t_in = Tm
            for i, u in enumerate(self.upsample_rates):
                C_, L_ = chans[i], lens[i]
                # xs alternates between two slots so that the next stage can read it while writing its own
                up, xs = view(0, L_, C_), view(1 if i % 2 == 0 else 5, L_, C_)
                # lrelu(0.1) -> ConvTranspose1d as poly-phase conv: y[B, t_in, u*C] == [B, t_in*u, C]  (models.py:110-111)
                ops.conv1d(cur, P[f"ups.{i}.w"], P[f"ups.{i}.b"], up.view(B, t_in, u * C_), P[f"ups.{i}.taps"],
                           pre_slope=LRELU_SLOPE)
                for j, (rk, dils) in enumerate(zip(self.rb_kernels, self.rb_dilations)):
                    n = i * self.num_kernels + j
                    last_j = j == self.num_kernels - 1
                    xcur = up
                    for m, dl in enumerate(dils):
                        last_m = m == len(dils) - 1
                        # MRF sum / num_kernels folded into the last residual epilogue of each block (models.py:113-118)
                        dst = xs if last_m else view(2 + (m & 1), L_, C_)
                        kw = dict(accumulate=last_m and j > 0, out_scale=(1.0 / self.num_kernels) if (last_m and last_j) else 1.0)
                        if self.resblock == "1":  # models.py:36-43
                            t1 = view(4, L_, C_)
                            ops.conv1d(xcur, P[f"resblocks.{n}.convs1.{m}.w"], self._sd[f"resblocks.{n}.convs1.{m}.bias"],
                                       t1, ops.conv_taps(rk, dl, get_padding(rk, dl)), pre_slope=LRELU_SLOPE)
                            ops.conv1d(t1, P[f"resblocks.{n}.convs2.{m}.w"], self._sd[f"resblocks.{n}.convs2.{m}.bias"],
                                       dst, ops.conv_taps(rk, 1, get_padding(rk, 1)), pre_slope=LRELU_SLOPE,
                                       residual=xcur, **kw)
                        else:  # models.py:66-71
                            ops.conv1d(xcur, P[f"resblocks.{n}.convs.{m}.w"], self._sd[f"resblocks.{n}.convs.{m}.bias"],
                                       dst, ops.conv_taps(rk, dl, get_padding(rk, dl)), pre_slope=LRELU_SLOPE,
                                       residual=xcur, **kw)
                        xcur = dst
                cur = xs
                t_in = L_
            io.y = torch.empty(B, 1, lens[-1], **f32)
            # lrelu(default 0.01) -> conv_post -> tanh (models.py:119-121)
            ops.conv1d_cout1(cur, P["conv_post.w"], self._sd["conv_post.bias"], io.y.view(B, lens[-1]), 7, 3, 0.01, ACT_TANH)
        io.plan = plan
        if self.use_cuda_graph:
            plan.capture()
        return io

    def _run(self, x, frame_major_in=False):
        self._require_cuda()
        if x.dim() != 3:
            raise SibError(f"Generator input must be 3-D, got {tuple(x.shape)}")
        if frame_major_in:
            B, Tm, Cin = x.shape
        else:
            B, Cin, Tm = x.shape
        if Cin != self.in_dim:
            raise RuntimeError(f"expected input with {self.in_dim} channels, got {Cin}")  # torch conv1d raises RuntimeError too
        key = (B, Tm, frame_major_in)
        with torch.cuda.device(self._device):   # the C side launches on the current device: make it this module's
            self._pack()
            io = self._cached_plan(key, lambda: self._build_plan(B, Tm, frame_major_in))
            (io.x if frame_major_in else io.x_cf).copy_(x.to(self._device, torch.float32), non_blocking=True)
            io.plan.run()
        return io

    def forward(self, x):
        """x [B, in_dim, Tm] (channels-first, as the reference) -> a fresh [B, 1, Tm*prod(rates)] tensor."""
        return self._run(x).y.clone()

    def forward_frame_major(self, x, clone: bool = True):
        """x [B, Tm, in_dim] frame-major (what extend_mel / embed_concat kernels emit) -> same output.  clone=False
        returns the plan's own output buffer (valid until the next call with this shape)."""
        y = self._run(x, frame_major_in=True).y
        return y.clone() if clone else y


class CodeGenerator(Generator):
    """I_da/src/model.py:42-189 for the shipped config (no code VQ; f0_stats and multispkr set).

    The frozen f0 VQ-VAE encoder + quantiser (model.py:148-153) is `self.fo_vqvae`, an `F0Quantizer`
    running on the same fp32 conv / argmin kernels: `forward(code=, f0=, emb=, spkr=)` is the reference call.  Its weights
    arrive under `fo_vqvae.*` in a CodeGenerator checkpoint, or through `load_f0_quantizer(ckpt['generator'])` (the file
    `h.f0_quantizer_path` points at, model.py:63-71).  Callers that already hold the bins may pass `f0_code=` instead."""

    def __init__(self, h, precision: str = "fp32"):
        super().__init__(h, precision)
        self.num_embeddings = int(_hget(h, "num_embeddings"))
        self.embedding_dim = int(_hget(h, "embedding_dim"))
        fq = _hget(h, "f0_quantizer") or {}
        self.f0_bins = int(fq.get("f0_vq_params", {}).get("l_bins", 20)) if isinstance(fq, dict) else 20
        self.n_speakers = int(_hget(h, "num_speakers", 200) or 200)
        E = self.embedding_dim
        self._add_param("emb_c.weight", torch.randn(self.num_embeddings, E))      # nn.Embedding default init N(0, 1)
        self._add_param("emb_p.weight", torch.randn(self.f0_bins, E))
        self._add_param("emb_s.weight", torch.randn(self.n_speakers, E))          # unused on the d-vector path (model.py:139)
        self.fo_vqvae = None
        self._f0_loaded = False
        if isinstance(fq, dict) and "f0_encoder_params" in fq:
            from .f0vq import F0Quantizer
            self.fo_vqvae = F0Quantizer(fq)

    def _extra_keys(self):
        return ["emb_c.weight", "emb_p.weight", "emb_s.weight"]

    def _adapt_state_dict(self, sd):
        sd = super()._adapt_state_dict(sd)
        if "emb_s.weight" in sd and tuple(sd["emb_s.weight"].shape) != tuple(self._sd["emb_s.weight"].shape):
            self._add_param("emb_s.weight", torch.empty_like(sd["emb_s.weight"], device=self._device))
        fo = [k for k in sd if k.startswith("fo_vqvae.")]
        if self.fo_vqvae is None:
            for k in fo:            # a checkpoint with the frozen f0 VQ-VAE loaded into a generator configured without it
                sd.pop(k)
        else:
            for k in [k for k in fo if k.startswith("fo_vqvae.decoder.")]:   # training-only half of the VQ-VAE
                sd.pop(k)
            if fo:
                self._f0_loaded = True
            else:                   # generator-only checkpoint: keep whatever the quantiser holds
                sd.update({"fo_vqvae." + k: v for k, v in self.fo_vqvae.state_dict().items()})
        return sd

    def load_f0_quantizer(self, sd, strict: bool = True):
        """`torch.load(h.f0_quantizer_path)["generator"]` (model.py:66-69)."""
        if self.fo_vqvae is None:
            raise SibError("this CodeGenerator was configured without an f0_quantizer")
        self._f0_loaded = True
        return self.fo_vqvae.load_state_dict(sd, strict)

    def forward(self, **kwargs):
        self._require_cuda()
        code = kwargs["code"]
        if "f0_code" not in kwargs:
            if kwargs.get("f0") is None or self.fo_vqvae is None or not self._f0_loaded:
                raise SibError("CodeGenerator needs f0= together with loaded fo_vqvae weights (load_state_dict with "
                               "fo_vqvae.* keys or load_f0_quantizer), or precomputed bins through f0_code=")
            kwargs = dict(kwargs, f0_code=self.fo_vqvae.encode(kwargs["f0"]))   # model.py:148-152
        zp, emb = kwargs["f0_code"], kwargs["emb"]
        dev = self._device
        code = code.to(dev, torch.int64).contiguous()
        zp = zp.to(dev, torch.int64).contiguous()
        emb = emb.to(dev, torch.float32).contiguous()
        B, T = code.shape
        E = self.embedding_dim
        Tmax = max(T, zp.shape[1])
        if Tmax != T:
            raise SibError("pitch series longer than the code series is not produced by the reference pipeline")
        with torch.cuda.device(dev):
            x = torch.empty(B, T, 2 * E + emb.shape[1], device=dev, dtype=torch.float32)
            ops.embed_concat(code, zp, emb, self._sd["emb_c.weight"], self._sd["emb_p.weight"], x)
        return self.forward_frame_major(x)
