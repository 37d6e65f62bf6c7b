"""Tensor-level wrappers over the C-ABI kernels + launch plans.

PyTorch is plumbing here (device memory, streams); every op below is a call into
`libsib_b200.so` on the current CUDA stream.  A `Plan` records the launches of one forward pass
over static buffers so that it can be replayed with ~2 us of host time per kernel or captured
into a CUDA graph (no tracing compiler involved).
"""
from __future__ import annotations

import contextlib
import ctypes as C

import torch

from . import _lib
from ._lib import ACT_GELU, ACT_LRELU, ACT_NONE, ACT_TANH, ConvDesc, ResUnitDesc, SibError  # noqa: F401

_active_plan = None


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _chk(t, dtype=None, name="tensor"):
    if t is None:
        return None
    if not t.is_cuda:
        raise SibError(f"{name} must be a CUDA tensor (no CPU fallback on this path)")
    if t.device.index != torch.cuda.current_device():
        # the C side launches on the CURRENT device (stream, SM count, attribute cache): a tensor of another device would
        # be an illegal address or silent peer traffic.  Modules / pipelines enter `torch.cuda.device(their device)`.
        raise SibError(f"{name} lives on cuda:{t.device.index} but the current device is cuda:{torch.cuda.current_device()}; "
                       "wrap the call in `with torch.cuda.device(...)`")
    if dtype is not None and t.dtype != dtype:
        raise SibError(f"{name} must be {dtype}, got {t.dtype}")
    return t


def _p(t):
    return None if t is None else t.data_ptr()


def _dt(t):
    """sib_dtype code of a tensor (SIB_F32 = 0, SIB_BF16 = 1)."""
    if t is None or t.dtype == torch.float32:
        return 0
    if t.dtype == torch.bfloat16:
        return 1
    raise SibError(f"unsupported dtype {t.dtype}")


class Plan:
    """Recorded launch list of one forward pass over static buffers.

    Steps may be tagged with a chain id (`with plan.chain(c):`).  Chain 0 runs on the caller's stream, chain c > 0 on a side
    stream of the plan; `fork()` / `join()` mark where the side chains branch off and come back (events, no host sync).
    Independent chains - e.g. the two half-batches of the transformer stack - let the tail of one kernel (a partial last
    wave of tiles) overlap the head of the other chain's kernel instead of leaving SMs idle."""

    def __init__(self):
        self.steps = []   # (cfunc, args-without-stream, name)
        self.chains = []  # chain id per step
        self.marks = {}   # step index -> list of "fork" / "join" executed before that step
        self.keep = []    # tensors / descs that must outlive the plan
        self.graph = None
        self._chain = 0
        self._side = {}   # chain id -> torch.cuda.Stream
        self._n_chains = 1
        self.device = torch.cuda.current_device() if torch.cuda.is_available() else None

    @contextlib.contextmanager
    def record(self):
        global _active_plan
        prev, _active_plan = _active_plan, self
        try:
            yield self
        finally:
            _active_plan = prev

    @contextlib.contextmanager
    def chain(self, c: int):
        prev, self._chain = self._chain, c
        self._n_chains = max(self._n_chains, c + 1)
        try:
            yield self
        finally:
            self._chain = prev

    def fork(self):
        self.marks.setdefault(len(self.steps), []).append("fork")

    def join(self):
        self.marks.setdefault(len(self.steps), []).append("join")

    def signal(self, tag):
        """Record an event on the CURRENT chain's stream after the steps recorded so far (see `wait`)."""
        self.marks.setdefault(len(self.steps), []).append(("signal", tag, self._chain))

    def wait(self, tag):
        """The current chain's stream waits for the event `tag` of another chain before its next step."""
        self.marks.setdefault(len(self.steps), []).append(("wait", tag, self._chain))

    def _launch_all(self, multi_stream: bool):
        main = torch.cuda.current_stream()
        s_main = main.cuda_stream
        if not multi_stream or self._n_chains == 1:
            for fn, args, name in self.steps:
                rc = fn(*args, s_main)
                if rc:
                    _lib.check(rc, name)
            return
        for c in range(1, self._n_chains):
            if c not in self._side:
                self._side[c] = torch.cuda.Stream(main.device)
        handles = {0: s_main, **{c: st.cuda_stream for c, st in self._side.items()}}
        streams = {0: main, **self._side}
        tagged = {}
        for i, ((fn, args, name), c) in enumerate(zip(self.steps, self.chains)):
            for m in self.marks.get(i, ()):
                if m == "fork":
                    ev = torch.cuda.Event()
                    ev.record(main)
                    for st in self._side.values():
                        st.wait_event(ev)
                elif m == "join":
                    for st in self._side.values():
                        ev = torch.cuda.Event()
                        ev.record(st)
                        main.wait_event(ev)
                elif m[0] == "signal":
                    ev = torch.cuda.Event()
                    ev.record(streams[m[2]])
                    tagged[m[1]] = ev
                else:   # ("wait", tag, chain)
                    streams[m[2]].wait_event(tagged[m[1]])
            rc = fn(*args, handles[c])
            if rc:
                _lib.check(rc, name)
        for m in self.marks.get(len(self.steps), ()):   # a join after the last step
            if m == "join":
                for st in self._side.values():
                    ev = torch.cuda.Event()
                    ev.record(st)
                    main.wait_event(ev)

    def run(self):
        if self.device is not None and self.device != torch.cuda.current_device():
            raise SibError(f"plan recorded on cuda:{self.device} replayed with cuda:{torch.cuda.current_device()} current")
        if self.graph is not None:
            self.graph.replay()
            return
        self._launch_all(multi_stream=_multi_chain_enabled())

    def capture(self):
        """Capture the launch list into a CUDA graph (launch-bound inner loops, one replay per forward); chains are
        serialised into the capturing stream."""
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            self._launch_all(multi_stream=False)  # warm-up outside capture (sets func attributes)
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._launch_all(multi_stream=False)
        self.graph = g

    def __len__(self):
        return len(self.steps)


_chains_on = None


def _multi_chain_enabled() -> bool:
    global _chains_on
    if _chains_on is None:
        import os
        _chains_on = os.environ.get("SIB_CHAINS", "1") != "0"
    return _chains_on


def set_multi_chain(enabled: bool) -> bool:
    """Run independent plan chains on separate streams (default) or serialise them on the caller's stream (A/B switch,
    also SIB_CHAINS=0).  Returns the previous setting."""
    global _chains_on
    prev = _multi_chain_enabled()
    _chains_on = bool(enabled)
    return prev


def _emit(name, args, keep=()):
    fn = getattr(_lib.lib(), name)
    if _active_plan is not None:
        _active_plan.steps.append((fn, args, name))
        _active_plan.chains.append(_active_plan._chain)
        _active_plan.keep.extend(keep)
        return
    rc = fn(*args, _stream())
    if rc:
        _lib.check(rc, name)


def set_pdl(enabled: bool) -> bool:
    """Programmatic dependent launch on / off (returns the previous setting); off gives clean per-kernel timings."""
    return bool(_lib.lib().sib_set_pdl(int(enabled)))


def launch_count() -> int:
    return int(_lib.lib().sib_launch_count())


# ----------------------------------------------------------------------------- weight packing
def pack_conv_weight(w: torch.Tensor, groups: int = 1) -> torch.Tensor:
    """torch Conv1d weight [Cout, Cin/g, k] -> [g][k][Cin/g][Cout/g] (fp32 SIMT layout)."""
    cout, cin_g, k = w.shape
    return w.view(groups, cout // groups, cin_g, k).permute(0, 3, 2, 1).contiguous()


def pack_linear_weight(w: torch.Tensor) -> torch.Tensor:
    """torch Linear weight [out, in] -> [1][1][in][out]."""
    return w.t().contiguous().view(1, 1, w.shape[1], w.shape[0])


def kblock(c_in_per_group: int):
    """(cc, tb) K-block geometry of the tcgen05 kernel for this channel count (sib_conv1d_bf16_kblock)."""
    cc, tb = C.c_int(), C.c_int()
    _lib.check(_lib.lib().sib_conv1d_bf16_kblock(c_in_per_group, C.byref(cc), C.byref(tb)), "sib_conv1d_bf16_kblock")
    return cc.value, tb.value


def to_kmajor_bf16(w_packed: torch.Tensor) -> torch.Tensor:
    """fp32 SIMT layout [g][taps][cin_g][cout_g] -> tcgen05 layout [g][cin_g/cc][taps][cout_g][cc] bf16:
    every (channel-chunk, tap) slab is one K-major B operand [cout_g rows x cc]."""
    g, taps, cin_g, cout_g = w_packed.shape
    cc, _ = kblock(cin_g)
    w = w_packed.permute(0, 1, 3, 2).reshape(g, taps, cout_g, cin_g // cc, cc).permute(0, 3, 1, 2, 4)
    return w.to(torch.bfloat16).contiguous()


def weight_norm_fold(v: torch.Tensor, g: torch.Tensor, dim: int = 0) -> torch.Tensor:
    """torch weight_norm folded on the device: v * (g / ||v||), norm over every dim but `dim` (0, or the last one)."""
    _chk(v, torch.float32, "weight_v"); _chk(g, torch.float32, "weight_g")
    v, g = v.contiguous(), g.contiguous()
    w = torch.empty_like(v)
    if dim == 0:
        outer, inner, last = v.shape[0], v.numel() // v.shape[0], 0
    elif dim in (v.dim() - 1, -1):
        inner = v.shape[-1]
        outer, last = v.numel() // inner, 1
    else:
        raise SibError(f"weight_norm_fold: dim={dim} unsupported (0 or last)")
    if g.numel() != (inner if last else outer):
        raise SibError(f"weight_norm_fold: weight_g has {g.numel()} entries for a {tuple(v.shape)} weight_v (dim {dim})")
    _emit("sib_weight_norm_fold_f32", (_p(v), _p(g), _p(w), outer, inner, last), keep=(v, g, w))
    return w


def conv_taps(k: int, dilation: int, padding: int):
    return [j * dilation - padding for j in range(k)]


def pack_conv_transpose(w: torch.Tensor, bias, stride: int, padding: int):
    """ConvTranspose1d weight [Cin, Cout, k] -> poly-phase conv over the *input* grid:
    output row q holds the `stride` phases of samples q*stride+p, so y[B, T, stride*Cout] is
    bit-for-bit the frame-major [B, T*stride, Cout].  Returns (w_packed[1][taps][Cin][s*Cout],
    bias[s*Cout], tap_offsets).  For o = q*s+p: j = r + m*s, r = (p+pad) % s, i = q + (p+pad)//s - m."""
    cin, cout, k = w.shape
    entries = []
    for p in range(stride):
        r, delta = (p + padding) % stride, (p + padding) // stride
        m = 0
        while r + m * stride < k:
            entries.append((delta - m, p, r + m * stride))
            m += 1
    offs = sorted({e[0] for e in entries})
    packed = torch.zeros(len(offs), cin, stride * cout, dtype=w.dtype, device=w.device)
    for off, p, j in entries:
        packed[offs.index(off), :, p * cout:(p + 1) * cout] = w[:, :, j]
    b = None if bias is None else bias.repeat(stride)
    return packed.unsqueeze(0).contiguous(), b, offs


# ----------------------------------------------------------------------------- conv / linear
def make_desc(batch, t_in, t_out, c_in, c_out, taps, *, stride=1, groups=1, x_row=None, x_batch=None,
              y_row=None, y_batch=None, r_row=None, r_batch=None, pre_slope=None, post_act=ACT_NONE,
              post_slope=0.0, out_scale=1.0, accumulate=False, res_after_act=False,
              act2_slope=0.0) -> ConvDesc:
    if len(taps) > _lib.SIB_MAX_TAPS:
        raise SibError(f"{len(taps)} taps > {_lib.SIB_MAX_TAPS}")
    d = ConvDesc()
    d.batch, d.t_in, d.t_out, d.c_in, d.c_out, d.groups = batch, t_in, t_out, c_in, c_out, groups
    d.n_taps, d.stride = len(taps), stride
    for i, o in enumerate(taps):
        d.tap_offset[i] = o
    d.x_row_stride = c_in if x_row is None else x_row
    d.x_batch_stride = t_in * d.x_row_stride if x_batch is None else x_batch
    d.y_row_stride = c_out if y_row is None else y_row
    d.y_batch_stride = t_out * d.y_row_stride if y_batch is None else y_batch
    d.r_row_stride = c_out if r_row is None else r_row
    d.r_batch_stride = t_out * d.r_row_stride if r_batch is None else r_batch
    d.pre_act = ACT_LRELU if pre_slope is not None else ACT_NONE
    d.pre_slope = 0.0 if pre_slope is None else pre_slope
    d.post_act, d.post_slope, d.out_scale = post_act, post_slope, out_scale
    d.accumulate, d.res_after_act = int(accumulate), int(res_after_act)
    d.act2_slope = act2_slope
    return d


def conv1d(x, w, bias, y, taps, *, stride=1, groups=1, residual=None, y_act=None, **kw):
    """x [B,T_in,Cin], y [B,T_out,Cout] frame-major (fp32 or bf16, chosen by x.dtype)."""
    B, t_in, c_in = x.shape
    _, t_out, c_out = y.shape
    d = make_desc(B, t_in, t_out, c_in, c_out, taps, stride=stride, groups=groups,
                  x_row=x.stride(1), x_batch=x.stride(0), y_row=y.stride(1), y_batch=y.stride(0),
                  r_row=None if residual is None else residual.stride(1),
                  r_batch=None if residual is None else residual.stride(0), **kw)
    if x.dtype == torch.float32:
        _chk(w, torch.float32, "w"); _chk(y, torch.float32, "y"); _chk(bias, torch.float32, "bias")
        _chk(residual, torch.float32, "residual")
        if y_act is not None:
            raise SibError("y_act is a bf16-path feature")
        _emit("sib_conv1d_f32", (C.byref(d), _p(x), _p(w), _p(bias), _p(residual), _p(y)), keep=(d, x, w, bias, residual, y))
    elif x.dtype == torch.bfloat16:
        _chk(w, torch.bfloat16, "w"); _chk(y, torch.bfloat16, "y"); _chk(bias, torch.float32, "bias")
        _chk(residual, torch.bfloat16, "residual"); _chk(y_act, torch.bfloat16, "y_act")
        _emit("sib_conv1d_bf16", (C.byref(d), _p(x), _p(w), _p(bias), _p(residual), _p(y), _p(y_act)),
              keep=(d, x, w, bias, residual, y, y_act))
    else:
        raise SibError(f"unsupported dtype {x.dtype}")


def conv_pre_act_supported(batch, t, c_in, c_out, taps, pre_slope=0.1, has_residual=False, has_y_act=False) -> bool:
    """Whether the tcgen05 conv can apply leaky-relu to its A tile in shared memory for this (stride-1) layer when it is
    launched with / without a residual input and an activated second output (both cost epilogue staging memory)."""
    d = make_desc(batch, t, t, c_in, c_out, taps, pre_slope=pre_slope)
    return bool(_lib.lib().sib_conv1d_bf16_pre_act_supported(C.byref(d), int(has_residual), int(has_y_act)))


def resunit_supported(c: int, k: int, dilation: int, accumulate: bool = False, has_y_act: bool = False) -> bool:
    return bool(_lib.lib().sib_resunit_bf16_supported(c, k, dilation, int(accumulate), int(has_y_act)))


def resunit(x, w1, b1, w2, b2, y, k, dilation, *, y_act=None, accumulate=False, out_scale=1.0, slope_in=0.1,
            slope_mid=0.1, act2_slope=0.1):
    """Fused ResBlock1 unit: y = (conv2(lrelu(conv1(lrelu(x)) + b1)) + b2 + x [+ y]) * out_scale on bf16 [B,T,C]."""
    B, T, Cc = x.shape
    for t, nme in ((x, "x"), (w1, "w1"), (w2, "w2"), (y, "y"), (y_act, "y_act")):
        _chk(t, torch.bfloat16, nme)
    _chk(b1, torch.float32, "b1"); _chk(b2, torch.float32, "b2")
    d = ResUnitDesc()
    d.batch, d.t, d.c, d.k, d.dilation = B, T, Cc, k, dilation
    d.accumulate = int(accumulate)
    d.slope_in, d.slope_mid, d.out_scale, d.act2_slope = slope_in, slope_mid, out_scale, act2_slope
    d.x_batch_stride, d.y_batch_stride = x.stride(0), y.stride(0)
    d.x_row_stride, d.y_row_stride = x.stride(1), y.stride(1)
    _emit("sib_resunit_bf16", (C.byref(d), _p(x), _p(w1), _p(b1), _p(w2), _p(b2), _p(y), _p(y_act)),
          keep=(d, x, w1, b1, w2, b2, y, y_act))


def linear(x2d, w, bias, y2d, *, residual=None, wait=None, signal=None, **kw):
    """x2d [M,K] @ packed w [1][1][K][N] (+bias) -> y2d [M,N].
    wait / signal: `FlowEdge`s of a `FlowChain` (bf16 only) - the launch then reads its rows block by block as their
    producer finishes them instead of waiting for the previous grid, and / or publishes its own row blocks (`sib_flow`)."""
    if wait is None and signal is None:
        conv1d(x2d.unsqueeze(0), w, bias, y2d.unsqueeze(0), [0],
               residual=None if residual is None else residual.unsqueeze(0), **kw)
        return
    M, K = x2d.shape
    N = y2d.shape[1]
    for t, nme in ((x2d, "x"), (w, "w"), (y2d, "y"), (residual, "residual")):
        _chk(t, torch.bfloat16, nme)
    _chk(bias, torch.float32, "bias")
    d = make_desc(1, M, M, K, N, [0], x_row=x2d.stride(0), x_batch=M * x2d.stride(0), y_row=y2d.stride(0), y_batch=M * y2d.stride(0),
                  r_row=None if residual is None else residual.stride(0),
                  r_batch=None if residual is None else M * residual.stride(0), **kw)
    fl = _flow_struct(M, wait, signal, N)
    _emit("sib_linear_flow_bf16", (C.byref(d), C.byref(fl), _p(x2d), _p(w), _p(bias), _p(residual), _p(y2d)),
          keep=(d, fl, x2d, w, bias, residual, y2d, wait, signal))


class FlowEdge:
    """One counter row of a `FlowChain`: written by exactly one producer launch, polled by its consumers."""

    def __init__(self, counters, kind, n, rows):
        self.counters, self.kind, self.n, self.rows = counters, kind, n, rows

    def targets(self):
        """(value of a complete 128-row block, value of the last block) - a linear layer stores whole 32-row quarters
        (4 x n per block whatever the number of valid rows), a LayerNorm adds n / 32 per valid row."""
        full = 4 * self.n
        if self.kind == "linear":
            return full, full
        # "layernorm" / "attention": n / 32 per valid row
        valid_last = self.rows - 128 * ((self.rows - 1) // 128)
        return full, valid_last * (self.n // 32)


class FlowChain:
    """Dataflow counters for a chain of row-wise dependent launches over `rows` flat rows (`sib_flow`): one int32 per
    128-row block and edge.  `reset()` is recorded at the head of the plan that uses the chain."""

    def __init__(self, rows: int, n_edges: int, device):
        self.rows = rows
        self.n_rb = 2 * ((rows + 255) // 256)       # the odd CTA of a last pair tile signals its block even when it is empty
        self.counters = torch.zeros(max(n_edges, 1), self.n_rb, dtype=torch.int32, device=device)
        self._next = 0

    def reset(self):
        _emit("sib_fill_zero", (_p(self.counters), self.counters.numel() * 4), keep=(self.counters,))

    def edge(self, kind: str, n: int) -> FlowEdge:
        if kind not in ("linear", "layernorm", "attention"):
            raise SibError(f"FlowChain.edge: unknown producer kind {kind!r}")
        if kind != "linear" and n % 32:
            raise SibError("FlowChain.edge: a row-wise producer needs a row length that is a multiple of 32")
        if self._next >= self.counters.shape[0]:
            raise SibError("FlowChain: more edges requested than allocated")
        e = FlowEdge(self.counters[self._next], kind, n, self.rows)
        self._next += 1
        return e


def _flow_struct(rows, wait, signal, n_out):
    fl = _lib.Flow()
    for e in (wait, signal):
        if e is not None and e.rows != rows:
            raise SibError(f"flow edge over {e.rows} rows used by a launch over {rows} rows")
    if wait is not None:
        fl.wait = _p(wait.counters)
        fl.wait_target, fl.wait_target_last = wait.targets()
    if signal is not None:
        if signal.n != n_out:
            raise SibError(f"flow edge declared for rows of {signal.n} values, the launch writes rows of {n_out}")
        fl.signal = _p(signal.counters)
    return fl


LN_SLOTS = _lib.SIB_LN_SLOTS


def linear_ln(x2d, w, bias, y2d, *, mode, n_norm, eps=1e-5, stats_in=None, colsum=None, gamma=None, beta=None, stats_out=None,
              residual=None, post_act=ACT_NONE):
    """bf16 linear layer with a LayerNorm folded into it (`sib_linear_ln_bf16`).
    mode "apply":    y = act(LN(t) W + b) from the RAW t (= x2d), W' = diag(gamma) W (= w), colsum of W', c = beta W + b (= bias)
                     and the row statistics of t (`stats_in` [M, LN_SLOTS, 2]).
    mode "residual": y = x W + b + LN(residual) (`stats_in` / gamma / beta; plain residual when stats_in is None), and the
                     partial row statistics of y go to `stats_out`."""
    M, K = x2d.shape
    N = y2d.shape[1]
    for t, nme in ((x2d, "x"), (w, "w"), (y2d, "y"), (residual, "residual")):
        _chk(t, torch.bfloat16, nme)
    for t, nme in ((bias, "bias"), (stats_in, "stats_in"), (colsum, "colsum"), (gamma, "gamma"), (beta, "beta"), (stats_out, "stats_out")):
        _chk(t, torch.float32, nme)
    for t in (stats_in, stats_out):
        if t is not None and (tuple(t.shape) != (M, LN_SLOTS, 2) or not t.is_contiguous()):
            raise SibError(f"linear_ln: statistics buffers must be dense [M, {LN_SLOTS}, 2] fp32 tensors")
    d = make_desc(1, M, M, K, N, [0], x_row=x2d.stride(0), x_batch=M * x2d.stride(0), y_row=y2d.stride(0), y_batch=M * y2d.stride(0),
                  r_row=None if residual is None else residual.stride(0),
                  r_batch=None if residual is None else M * residual.stride(0), post_act=post_act)
    ln = _lib.LnFold()
    ln.stats_in, ln.colsum, ln.gamma, ln.beta, ln.stats_out = _p(stats_in), _p(colsum), _p(gamma), _p(beta), _p(stats_out)
    ln.mode = _lib.SIB_LN_APPLY if mode == "apply" else _lib.SIB_LN_RESIDUAL
    ln.n_norm, ln.eps = n_norm, eps
    _emit("sib_linear_ln_bf16", (C.byref(d), C.byref(ln), _p(x2d), _p(w), _p(bias), _p(residual), _p(y2d)),
          keep=(d, ln, x2d, w, bias, residual, y2d, stats_in, colsum, gamma, beta, stats_out))


def linear_skinny(x2d, w, bias, y2d):
    """x2d [M,K] fp32 @ packed w [1][1][K][N] (+bias) -> y2d [M,N] for N <= 128: one CTA per row (few-hundred-row heads)."""
    M, K = x2d.shape
    N = y2d.shape[1]
    for t, nme in ((x2d, "x"), (w, "w"), (bias, "bias"), (y2d, "y")):
        _chk(t, torch.float32, nme)
    if not (x2d.is_contiguous() and y2d.is_contiguous() and w.is_contiguous()) or w.numel() != K * N:
        raise SibError("linear_skinny: operands must be contiguous and w must hold K x N values")
    _emit("sib_linear_skinny_f32", (_p(x2d), _p(w), _p(bias), _p(y2d), M, K, N), keep=(x2d, w, bias, y2d))


def conv1d_cout1(x, w, bias, y, k, pad, pre_slope, post_act):
    B, T, Cc = x.shape
    _chk(x, None, "x"); _chk(y, torch.float32, "y")
    _emit("sib_conv1d_cout1", (_p(x), _dt(x), _p(w), _p(bias), _p(y), B, T, Cc, k, pad, pre_slope, post_act),
          keep=(x, w, bias, y))


# ----------------------------------------------------------------------------- norms / attention
def conv0(mode, wave, w, bias, c, k, stride, t0, *, partial=None, mean=None, rstd=None, gamma=None, beta=None, y=None):
    B, n = wave.shape
    _chk(wave, torch.float32, "wave")
    _emit("sib_conv0", (mode, _p(wave), B, n, wave.stride(0), _p(w), _p(bias), c, k, stride, t0, _p(partial),
                        _p(mean), _p(rstd), _p(gamma), _p(beta), _p(y), _dt(y)),
          keep=(wave, w, bias, partial, mean, rstd, gamma, beta, y))


def conv0_num_tiles(t0: int) -> int:
    return int(_lib.lib().sib_conv0_num_tiles(t0))


def gn_finalize(partial, batch, n_tiles, c, t0, eps, mean, rstd):
    _emit("sib_gn_finalize_f32", (_p(partial), batch, n_tiles, c, t0, eps, _p(mean), _p(rstd)), keep=(partial, mean, rstd))


def conv0_gn_stats(wave, w, bias, c, k, stride, t0, eps, mean, rstd):
    """GroupNorm statistics of conv0's output in closed form from the waveform (no conv0 evaluation pass)."""
    B, n = wave.shape
    _chk(wave, torch.float32, "wave"); _chk(mean, torch.float32, "mean"); _chk(rstd, torch.float32, "rstd")
    _emit("sib_conv0_gn_stats_f32", (_p(wave), B, n, wave.stride(0), _p(w), _p(bias), c, k, stride, t0, eps, _p(mean), _p(rstd)),
          keep=(wave, w, bias, mean, rstd))


def layernorm(x, gamma, beta, y, eps=1e-5, residual=None, post_act=ACT_NONE, wait=None, signal=None):
    c = x.shape[-1]
    rows = x.numel() // c
    _chk(x, None, "x"); _chk(y, None, "y"); _chk(gamma, torch.float32, "gamma")
    if not (x.is_contiguous() and y.is_contiguous() and (residual is None or residual.is_contiguous())):
        raise SibError("layernorm: tensors must be contiguous")
    if wait is not None or signal is not None:
        for t, nme in ((x, "x"), (y, "y"), (residual, "residual")):
            _chk(t, torch.bfloat16, nme)
        if post_act != ACT_NONE:
            raise SibError("layernorm: the dataflow variant has no activation")
        fl = _flow_struct(rows, wait, signal, c)
        _emit("sib_layernorm_flow_bf16", (_p(x), _p(residual), _p(gamma), _p(beta), _p(y), rows, c, eps, C.byref(fl)),
              keep=(x, residual, gamma, beta, y, fl, wait, signal))
        return
    _emit("sib_layernorm", (_p(x), _dt(x), _p(residual), _dt(residual), _p(gamma), _p(beta), _p(y), _dt(y), rows, c, eps,
                            post_act), keep=(x, residual, gamma, beta, y))


def attention(qkv, key_len, out, heads, wait=None, signal=None):
    B, T, H3 = qkv.shape
    _chk(qkv, None, "qkv"); _chk(out, qkv.dtype, "out"); _chk(key_len, torch.int32, "key_len")
    if wait is not None or signal is not None:
        _chk(qkv, torch.bfloat16, "qkv")
        if not (qkv.is_contiguous() and out.is_contiguous()):
            raise SibError("attention: the dataflow variant needs dense [B, T, 3H] / [B, T, H] tensors (flat rows)")
        fl = _flow_struct(B * T, wait, signal, H3 // 3)
        _emit("sib_attention_flow_bf16", (_p(qkv), _p(key_len), _p(out), B, T, heads, H3 // 3 // heads, C.byref(fl)),
              keep=(qkv, key_len, out, fl, wait, signal))
        return
    _emit("sib_attention", (_p(qkv), _dt(qkv), _p(key_len), _p(out), B, T, heads, H3 // 3 // heads), keep=(qkv, key_len, out))


def zero_padded_frames(h, key_len):
    B, T, Cc = h.shape
    _emit("sib_zero_padded_frames", (_p(h), _dt(h), _p(_chk(key_len, torch.int32, "key_len")), B, T, Cc), keep=(h, key_len))


# ----------------------------------------------------------------------------- glue
def zero_ranges(wave, lo, hi, add_eps=0.0):
    B, n = wave.shape
    _chk(wave, torch.float32, "wave"); _chk(lo, torch.int32, "lo"); _chk(hi, torch.int32, "hi")
    if not wave.is_contiguous():
        raise SibError("wave must be contiguous")
    _emit("sib_zero_ranges_f32", (_p(wave), B, n, _p(lo), _p(hi), add_eps), keep=(wave, lo, hi))


def mask_peak_normalize(x, y, lo=None, hi=None, scale=0.95):
    """y = librosa.util.normalize(x with [lo, hi) zeroed) * scale per utterance (predict.py:99-103)."""
    B, n = x.shape
    _chk(x, torch.float32, "x"); _chk(y, torch.float32, "y"); _chk(lo, torch.int32, "lo"); _chk(hi, torch.int32, "hi")
    _emit("sib_mask_peak_normalize_f32", (_p(x), _p(y), B, n, _p(lo), _p(hi), scale), keep=(x, y, lo, hi))


def znorm(x, y, lengths=None, eps=1e-7):
    B, n = x.shape
    _chk(x, torch.float32, "x"); _chk(y, torch.float32, "y"); _chk(lengths, torch.int32, "lengths")
    _emit("sib_znorm_f32", (_p(x), _p(y), B, n, _p(lengths), eps), keep=(x, y, lengths))


def gather_frames(src, pos, length, off, out):
    B, T, Dm = src.shape
    _chk(src, torch.float32, "src"); _chk(out, torch.float32, "out")
    for t, nme in ((pos, "pos"), (length, "len"), (off, "off")):
        _chk(t, torch.int32, nme)
    _emit("sib_gather_frames_f32", (_p(src), B, T, Dm, _p(pos), _p(length), _p(off), _p(out)), keep=(src, pos, length, off, out))


def cos_argmax(v, cc, labels):
    M, Dm = v.shape
    _chk(v, torch.float32, "v"); _chk(cc, torch.float32, "cc"); _chk(labels, torch.int64, "labels")
    _emit("sib_cos_argmax_f32", (_p(v), _p(cc), M, cc.shape[0], Dm, _p(labels)), keep=(v, cc, labels))


def l2_argmin(f, mu, labels):
    M, Dm = f.shape
    _chk(f, torch.float32, "f"); _chk(mu, torch.float32, "mu"); _chk(labels, torch.int64, "labels")
    _emit("sib_l2_argmin_f32", (_p(f), _p(mu), M, mu.shape[0], Dm, _p(labels)), keep=(f, mu, labels))


def kmeans_pack(mu):
    """Centres [K, D] -> (packed mu^T for the fp32 GEMM, bias -0.5 ||mu_k||^2) for `kmeans_assign`."""
    _chk(mu, torch.float32, "mu")
    mu = mu.contiguous()
    K, Dm = mu.shape
    bias = torch.empty(K, device=mu.device, dtype=torch.float32)
    _emit("sib_row_sqnorm_f32", (_p(mu), K, Dm, -0.5, _p(bias)), keep=(mu, bias))
    return pack_linear_weight(mu), bias


def kmeans_assign(f, mu_packed, mu_bias, labels, scores=None):
    """labels[m] = argmin_k ||f[m] - mu_k||^2 through one fp32 GEMM (scores = f mu^T - 0.5 ||mu||^2) and a row argmax:
    sklearn's own float32 formulation of KMeans.predict (I_da/scripts/inpainting.py:204-205)."""
    M, Dm = f.shape
    K = mu_bias.shape[0]
    _chk(f, torch.float32, "f"); _chk(labels, torch.int64, "labels")
    if scores is None:
        scores = torch.empty(M, K, device=f.device, dtype=torch.float32)
    linear(f.contiguous(), mu_packed, mu_bias, scores)
    _emit("sib_row_argmax_f32", (_p(scores), M, K, _p(labels)), keep=(scores, labels))
    return labels


def paste_centroids(mel, cc, center, labels, pos, length, off):
    B, Dm, T = mel.shape
    _chk(mel, torch.float32, "mel"); _chk(labels, torch.int64, "labels"); _chk(cc, torch.float32, "cc")
    if not (mel.is_contiguous() and cc.is_contiguous()) or cc.shape[1] != Dm:
        raise SibError("paste_centroids: mel must be a dense [B, D, T] tensor and cc a dense [K, D] codebook")
    _emit("sib_paste_centroids_f32", (_p(mel), B, Dm, T, _p(cc), _p(center), _p(labels), _p(pos), _p(length), _p(off),
                                      cc.shape[0]), keep=(mel, cc, center, labels, pos, length, off))


def extend_mel_len(t: int) -> int:
    """floor(T * 441/256) with the float arithmetic F.interpolate uses."""
    import math
    return int(math.floor(float(t) * (441 / 256)))


def extend_mel(mel, out, frame_major: bool):
    B, Dm, T = mel.shape
    tm = out.shape[1] if frame_major else out.shape[2]
    _chk(mel, torch.float32, "mel"); _chk(out, None, "out")
    if not (mel.is_contiguous() and out.is_contiguous()):
        raise SibError("extend_mel: mel [B, D, T] and out must be dense tensors")
    _emit("sib_extend_mel", (_p(mel), _p(out), _dt(out), B, Dm, T, tm, int(frame_major)), keep=(mel, out))


def transpose(x, out):
    """[B,R,C] -> [B,C,R]"""
    B, R, Cc = x.shape
    _chk(x, torch.float32, "x"); _chk(out, torch.float32, "out")
    _emit("sib_transpose_f32", (_p(x), _p(out), B, R, Cc), keep=(x, out))


def embed_concat(code, zp, spk, emb_c, emb_p, out):
    B, T = code.shape
    _chk(code, torch.int64, "code"); _chk(zp, torch.int64, "zp"); _chk(spk, torch.float32, "spk")
    _chk(emb_c, torch.float32, "emb_c"); _chk(emb_p, torch.float32, "emb_p")
    if not all(t.is_contiguous() for t in (code, zp, spk, emb_c, emb_p, out)):
        raise SibError("embed_concat: operands must be dense")
    _emit("sib_embed_concat_f32", (_p(code), _p(zp), _p(spk), _p(emb_c), _p(emb_p), _p(out), B, T, zp.shape[1],
                                   emb_c.shape[1], spk.shape[1], emb_c.shape[0], emb_p.shape[0]),
          keep=(code, zp, spk, emb_c, emb_p, out))


def pack_int16(y, out):
    _chk(y, torch.float32, "y"); _chk(out, torch.int16, "out")
    _emit("sib_pack_int16_f32", (_p(y), _p(out), y.numel()), keep=(y, out))


def mel_spectrogram(wave, basis, out, hop, pad):
    B, n = wave.shape
    _chk(wave, torch.float32, "wave"); _chk(basis, torch.float32, "basis"); _chk(out, torch.float32, "out")
    ws = torch.empty(int(_lib.lib().sib_mel_workspace_bytes(basis.shape[0])), dtype=torch.uint8, device=wave.device)
    _emit("sib_mel_spectrogram_f32", (_p(wave), B, n, hop, pad, _p(basis), basis.shape[0], _p(out), out.shape[2], _p(ws)),
          keep=(wave, basis, out, ws))


def cast_to_bf16(x, out):
    _emit("sib_cast_f32_to_bf16", (_p(x), _p(out), x.numel()), keep=(x, out))


def cast_to_f32(x, out):
    _emit("sib_cast_bf16_to_f32", (_p(x), _p(out), x.numel()), keep=(x, out))


# ----------------------------------------------------------------------------- audio formats / metrics
def resample(x, filt, up, down, first, y, len_in=None, len_out=None):
    """y[b, n] = sum_j filt[j][n % up] * x[b, (n*down)//up + first + j]; x int16 PCM (scaled by 1/32768) or float32."""
    B, n_in = x.shape
    _chk(x, None, "x"); _chk(filt, torch.float32, "filt"); _chk(y, torch.float32, "y"); _chk(len_in, torch.int32, "len_in")
    _chk(len_out, torch.int32, "len_out")
    if x.dtype == torch.int16:
        dt = 2
    elif x.dtype == torch.float32:
        dt = 0
    else:
        raise SibError(f"resample: input must be int16 or float32, got {x.dtype}")
    if x.stride(1) != 1 or y.stride(1) != 1 or not filt.is_contiguous() or filt.shape[1] != up:
        raise SibError("resample: rows must be dense and filt must be [taps][up]")
    _emit("sib_resample", (_p(x), dt, B, n_in, x.stride(0), _p(len_in), _p(filt), up, down, filt.shape[0], first, _p(y),
                           y.shape[1], y.stride(0), _p(len_out)), keep=(x, filt, y, len_in, len_out))


def si_sdr(est, ref, out, lengths=None, eps=1.1920929e-07):
    B, n = est.shape
    for t, nme in ((est, "est"), (ref, "ref"), (out, "out")):
        _chk(t, torch.float32, nme)
    _chk(lengths, torch.int32, "lengths")
    if not (est.is_contiguous() and ref.is_contiguous()) or ref.shape != est.shape:
        raise SibError("si_sdr: est / ref must be contiguous [B, n] tensors of the same shape")
    ws = torch.empty(int(_lib.lib().sib_si_sdr_workspace_bytes(B)), dtype=torch.uint8, device=est.device)
    _emit("sib_si_sdr_f32", (_p(est), _p(ref), B, n, _p(lengths), eps, _p(out), _p(ws)), keep=(est, ref, out, lengths, ws))


def abs_diff_mean(a, b, out):
    """out[i] = mean |a[i] - b[i]| over everything but the leading axis."""
    _chk(a, torch.float32, "a"); _chk(b, torch.float32, "b"); _chk(out, torch.float32, "out")
    if a.shape != b.shape or not (a.is_contiguous() and b.is_contiguous()):
        raise SibError("abs_diff_mean: operands must be contiguous and of the same shape")
    B = a.shape[0]
    ws = torch.empty(int(_lib.lib().sib_abs_diff_workspace_bytes(B)), dtype=torch.uint8, device=a.device)
    _emit("sib_abs_diff_mean_f32", (_p(a), _p(b), B, a.numel() // B, _p(out), _p(ws)), keep=(a, b, out, ws))
