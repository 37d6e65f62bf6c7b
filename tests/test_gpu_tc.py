"""-m gpu: the bf16 tcgen05/TMA implicit-GEMM kernel (sib_conv1d_bf16) against torch fp32 run on the same
bf16-rounded operands.  Accumulation is fp32 in TMEM, so the only differences are summation order and the
final bf16 rounding of the output: tolerance = 1 bf16 ulp of the result (2^-8 relative) + small absolute."""
import math

import pytest
import torch
import torch.nn.functional as F

from util import max_abs, to_frame_major

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sib():
    import speech_inpainting_b200 as m
    m._load_lib()
    return m


def _rand(*shape, seed=0, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale


def _bf(x):
    return x.to(torch.bfloat16).to(torch.float32)


def _close(ref, got, tag=""):
    err = (ref - got).abs()
    tol = 2.0 ** -7 * ref.abs() + 2e-2 * float(ref.abs().mean())
    bad = (err > tol).sum().item()
    assert bad == 0, f"{tag}: {bad} elements out of tolerance, max err {err.max():.4g}, ref max {ref.abs().max():.4g}"


def _halo_buf(x_fm, halo):
    """[B,T,C] -> zero-haloed buffer; returns (buffer, view of the valid rows)."""
    B, T, C = x_fm.shape
    buf = torch.zeros(B * (T + 2 * halo) + 256, C, dtype=torch.bfloat16, device="cuda")  # + slack rows at the end
    v = buf[: B * (T + 2 * halo)].view(B, T + 2 * halo, C)[:, halo:halo + T]
    v.copy_(x_fm)
    return buf, v


TC_CASES = [
    # B, T, Cin, Cout, k, stride, dil, pad, groups, halo
    (1, 333, 768, 768, 1, 1, 1, 0, 1, 0),      # linear
    (1, 6368, 768, 2304, 1, 1, 1, 0, 1, 0),    # QKV projection at config-2 size
    (2, 300, 128, 128, 7, 1, 3, 9, 1, 0),      # dilated ResBlock conv, OOB zero padding
    (2, 517, 256, 256, 11, 1, 5, 25, 1, 0),
    (3, 401, 512, 512, 3, 2, 1, 0, 1, 0),      # HuBERT conv1-4 (stride 2, odd length)
    (2, 400, 512, 512, 2, 2, 1, 0, 1, 0),      # HuBERT conv5-6
    (2, 700, 64, 64, 3, 1, 1, 1, 1, 0),
    (2, 900, 32, 32, 11, 1, 5, 25, 1, 32),     # tap-blocked (cc=32, tb=2), halo
    (2, 900, 32, 32, 3, 1, 3, 3, 1, 32),
    (2, 500, 16, 16, 7, 1, 1, 3, 1, 32),       # I_da last stage (cc=16, tb=4)
    (2, 99, 768, 768, 128, 1, 1, 64, 16, 64),  # pos-conv: 48 ch / group -> cc=16, tb=4
    (1, 120, 1024, 1024, 128, 1, 1, 64, 16, 64),  # large pos-conv: 64 ch / group -> cc=64
    (2, 50, 80, 512, 7, 1, 1, 3, 1, 8),        # conv_pre (80 = 5 x 16)
    (1, 200, 512, 80, 1, 1, 1, 0, 1, 0),       # head: N tail (80 = 64 + 16)
]


@pytest.mark.parametrize("case", TC_CASES)
def test_conv1d_bf16_tc(sib, case):
    B, T, Cin, Cout, k, s, d, p, g, halo = case
    x = _bf(_rand(B, Cin, T, seed=1))
    w = _bf(_rand(Cout, Cin // g, k, seed=2, scale=1.0 / math.sqrt(Cin // g * k)))
    b = _rand(Cout, seed=3, scale=0.1)
    ref = F.conv1d(x, w, b, stride=s, dilation=d, padding=p, groups=g)
    t_out = min(ref.shape[-1], T) if k == 128 else ref.shape[-1]
    ref = to_frame_major(ref[..., :t_out])
    wk = sib.ops.to_kmajor_bf16(sib.ops.pack_conv_weight(w.cuda(), g))
    xfm = to_frame_major(x).to(torch.bfloat16).cuda()
    if halo:
        keep, xd = _halo_buf(xfm, halo)
    else:
        keep = torch.zeros(xfm.numel() + 4096, dtype=torch.bfloat16, device="cuda")  # slack for the stride-s view
        xd = keep[: xfm.numel()].view_as(xfm)
        xd.copy_(xfm)
    y = torch.full((B, t_out, Cout), float("nan"), dtype=torch.bfloat16, device="cuda")
    sib.ops.conv1d(xd, wk, b.cuda(), y, sib.ops.conv_taps(k, d, p), stride=s, groups=g)
    torch.cuda.synchronize()
    _close(ref, y.float().cpu(), f"case {case}")


def test_conv1d_bf16_epilogue_variants(sib):
    B, T, Cc = 2, 300, 128
    x, w, b = _bf(_rand(B, Cc, T, seed=5)), _bf(_rand(Cc, Cc, 3, seed=6, scale=0.1)), _rand(Cc, seed=7)
    res, y0 = _bf(_rand(B, T, Cc, seed=8)), _bf(_rand(B, T, Cc, seed=9))
    conv = to_frame_major(F.conv1d(x, w, b, padding=1))
    xd = to_frame_major(x).to(torch.bfloat16).cuda()
    wk = sib.ops.to_kmajor_bf16(sib.ops.pack_conv_weight(w.cuda()))
    taps = sib.ops.conv_taps(3, 1, 1)
    # residual + accumulate + scale, second (activated) output
    y = y0.to(torch.bfloat16).cuda()
    y2 = torch.empty_like(y)
    sib.ops.conv1d(xd, wk, b.cuda(), y, taps, residual=res.to(torch.bfloat16).cuda(), accumulate=True, out_scale=1 / 3,
                   y_act=y2, act2_slope=0.1)
    ref = (conv + res + y0) / 3
    _close(ref, y.float().cpu(), "acc")
    _close(F.leaky_relu(_bf(ref), 0.1), y2.float().cpu(), "y_act")
    # gelu then residual (pos-conv wiring), lrelu post-act, tanh
    for act, fn in ((sib.ops.ACT_GELU, F.gelu), (sib.ops.ACT_TANH, torch.tanh)):
        y = torch.empty(B, T, Cc, dtype=torch.bfloat16, device="cuda")
        sib.ops.conv1d(xd, wk, b.cuda(), y, taps, post_act=act, residual=res.to(torch.bfloat16).cuda(), res_after_act=True)
        _close(fn(conv) + res, y.float().cpu(), f"act{act}")
    y = torch.empty(B, T, Cc, dtype=torch.bfloat16, device="cuda")
    sib.ops.conv1d(xd, wk, b.cuda(), y, taps, post_act=sib.ops.ACT_LRELU, post_slope=0.1)
    _close(F.leaky_relu(conv, 0.1), y.float().cpu(), "lrelu")


@pytest.mark.parametrize("ks,cin,cout", [((16, 8), 512, 256), ((4, 2), 128, 64), ((4, 2), 64, 32), ((11, 5), 512, 256),
                                         ((4, 2), 32, 16)])
def test_conv_transpose_bf16_tc(sib, ks, cin, cout):
    k, s = ks
    B, T = 2, 45
    x, w, b = _bf(_rand(B, cin, T, seed=1)), _bf(_rand(cin, cout, k, seed=2, scale=1 / math.sqrt(cin * k / s))), _rand(cout, seed=3)
    ref = to_frame_major(F.conv_transpose1d(x, w, b, stride=s, padding=(k - s) // 2))
    wp, bp, taps = sib.ops.pack_conv_transpose(w.cuda(), b.cuda(), s, (k - s) // 2)
    keep, xd = _halo_buf(to_frame_major(x).to(torch.bfloat16).cuda(), 8)
    y = torch.empty(B, T * s, cout, dtype=torch.bfloat16, device="cuda")
    sib.ops.conv1d(xd, sib.ops.to_kmajor_bf16(wp), bp, y.view(B, T, s * cout), taps)
    _close(ref, y.float().cpu(), f"convT {ks}")


@pytest.mark.parametrize("B,T,nh,padded", [(3, 99, 4, False), (2, 199, 12, True), (2, 128, 2, False), (3, 130, 4, True),
                                            (2, 499, 3, True), (1, 257, 2, False), (2, 16, 2, True)])
def test_attention_bf16_tc(sib, B, T, nh, padded):
    """tcgen05 attention (HF:234-259) vs torch fp32 on the same bf16-rounded q|k|v: S and O accumulate in fp32 (TMEM),
    P is rounded to bf16 before P.V, the output to bf16 - tolerance a few bf16 ulps of the output scale."""
    d = 64
    H = nh * d
    qkv = _bf(_rand(B, T, 3 * H, seed=T + nh))
    kl = None
    if padded:
        kl = torch.tensor([T, max(1, T - 17), 5][:B], dtype=torch.int32)
    q, k, v = [t.view(B, T, nh, d).transpose(1, 2) for t in qkv.split(H, dim=-1)]
    s = torch.matmul(q, k.transpose(2, 3)) * d ** -0.5
    if padded:
        km = torch.arange(T)[None, :] < kl[:, None]
        s = s.masked_fill(~km[:, None, None, :], torch.finfo(torch.float32).min)
    ref = torch.matmul(F.softmax(s, -1), v).transpose(1, 2).reshape(B, T, H)
    out = torch.full((B, T, H), float("nan"), device="cuda", dtype=torch.bfloat16)
    sib.ops.attention(qkv.cuda().to(torch.bfloat16), None if kl is None else kl.cuda(), out, nh)
    got = out.float().cpu()
    assert torch.isfinite(got).all()
    err = (ref - got).abs().max().item()
    assert err < 2.5e-2 * max(1.0, ref.abs().max().item()), f"max err {err:.4g} (ref max {ref.abs().max():.3g})"
    assert (ref - got).abs().mean().item() < 3e-3


RU_CASES = [
    # B, T, C, k, dil, accumulate, y_act, scale
    (2, 1000, 32, 3, 1, False, False, 1.0),
    (2, 777, 32, 7, 3, False, False, 1.0),
    (3, 500, 32, 11, 5, True, True, 1.0 / 3),
    (1, 118, 32, 11, 1, False, True, 1.0),       # exactly one tile
    (2, 119, 32, 11, 5, True, False, 1.0),       # one row into the second tile
    (2, 640, 64, 3, 5, False, False, 1.0),
    (2, 300, 64, 3, 1, True, True, 1.0 / 3),
    (2, 333, 64, 7, 3, False, True, 1.0),
    (2, 50, 64, 7, 1, False, False, 1.0),        # shorter than one tile
    (2, 300, 64, 11, 5, False, False, 1.0),      # 176 KB of weights: CTA-pair variant (output channels split over two SMs)
    (1, 300, 64, 11, 1, True, True, 1.0 / 3),    # odd tile count: the pair's second CTA runs a masked duplicate
    (3, 119, 64, 11, 3, True, False, 1.0),
    (8, 6000, 64, 11, 5, True, True, 1.0 / 3),   # 408 tiles: every persistent CTA pair loops over several tile pairs
    (2, 1000, 128, 3, 1, False, False, 1.0),     # C = 128 (stage 2): wide CTA-pair variant, two 64-channel chunks per row
    (1, 126, 128, 3, 5, False, False, 1.0),      # exactly one tile (R = 126), odd tile count -> masked duplicate in the pair
    (3, 127, 128, 3, 3, False, False, 1.0),      # one row into the second tile
    (8, 5000, 128, 3, 5, False, False, 1.0 / 3), # 320 tiles: every pair loops; output scale
    (2, 300, 128, 1, 1, False, False, 1.0),      # k = 1
    (2, 400, 16, 3, 1, False, False, 1.0),       # I_da last stage
    (2, 401, 16, 11, 5, True, True, 1.0 / 3),
]


@pytest.mark.parametrize("case", RU_CASES)
def test_resunit_bf16(sib, case):
    """Fused ResBlock1 unit (models.py:36-43) vs the two torch convs in fp32 on the same bf16-rounded operands; the
    intermediate is rounded to bf16 in the kernel (it is the A operand of conv2), so the reference rounds it too."""
    B, T, C, k, dil, accumulate, want_act, scale = case
    ops = sib.ops
    if not ops.resunit_supported(C, k, dil, accumulate, want_act):
        pytest.skip("configuration does not fit in shared memory")
    x = _bf(_rand(B, C, T, seed=1))
    w1 = _bf(_rand(C, C, k, seed=2, scale=1.0 / math.sqrt(C * k)))
    w2 = _bf(_rand(C, C, k, seed=3, scale=1.0 / math.sqrt(C * k)))
    b1, b2 = _rand(C, seed=4, scale=0.1), _rand(C, seed=5, scale=0.1)
    y_old = _bf(_rand(B, C, T, seed=6))
    t1 = _bf(F.leaky_relu(F.conv1d(F.leaky_relu(x, 0.1), w1, b1, padding=(k - 1) * dil // 2, dilation=dil), 0.1))
    ref = F.conv1d(t1, w2, b2, padding=(k - 1) // 2) + x
    if accumulate:
        ref = ref + y_old
    ref = ref * scale
    xd = to_frame_major(x).cuda().to(torch.bfloat16).contiguous()
    yd = to_frame_major(y_old).cuda().to(torch.bfloat16).contiguous() if accumulate else \
        torch.full((B, T, C), float("nan"), device="cuda", dtype=torch.bfloat16)
    ya = torch.full((B, T, C), float("nan"), device="cuda", dtype=torch.bfloat16) if want_act else None
    w1d = ops.to_kmajor_bf16(ops.pack_conv_weight(w1.cuda()))
    w2d = ops.to_kmajor_bf16(ops.pack_conv_weight(w2.cuda()))
    ops.resunit(xd, w1d, b1.cuda(), w2d, b2.cuda(), yd, k, dil, y_act=ya, accumulate=accumulate, out_scale=scale,
                act2_slope=0.01)
    torch.cuda.synchronize()
    got = yd.float().cpu()
    assert torch.isfinite(got).all()
    _close(to_frame_major(ref), got, f"resunit {case}")
    if want_act:
        _close(to_frame_major(F.leaky_relu(ref, 0.01)), ya.float().cpu(), f"resunit y_act {case}")


@pytest.mark.parametrize("c,res,act", [(768, True, False), (512, False, False), (1024, True, True), (80, True, False), (2048, False, False)])
def test_layernorm_bf16(sib, c, res, act):
    """bf16 LayerNorm (vectorised path for C % 8 == 0, scalar otherwise) vs torch fp32 on the same bf16 inputs."""
    x, r = _bf(_rand(41, 7, c, seed=1, scale=2.0)), _bf(_rand(41, 7, c, seed=2))
    g, b = 1 + 0.1 * _rand(c, seed=3), _rand(c, seed=4)
    y = torch.empty(41, 7, c, device="cuda", dtype=torch.bfloat16)
    sib.ops.layernorm(x.cuda().to(torch.bfloat16), g.cuda(), b.cuda(), y, 1e-5,
                      residual=r.cuda().to(torch.bfloat16) if res else None,
                      post_act=sib.ops.ACT_GELU if act else sib.ops.ACT_NONE)
    ref = F.layer_norm(x + r if res else x, (c,), g, b, 1e-5)
    if act:
        ref = F.gelu(ref)
    _close(ref, y.float().cpu(), f"layernorm bf16 c={c}")


@pytest.mark.parametrize("B,T,C,k", [(2, 1000, 32, 7), (3, 257, 16, 7), (1, 255, 64, 3), (2, 5, 32, 7)])
def test_conv_cout1_bf16(sib, B, T, C, k):
    """conv_post (models.py:119-121: lrelu(0.01) -> Conv1d(C, 1, 7) -> tanh) on bf16 frame-major activations."""
    x = _bf(_rand(B, C, T, seed=1))
    w, b = _rand(1, C, k, seed=2, scale=0.2), _rand(1, seed=3, scale=0.1)
    ref = torch.tanh(F.conv1d(F.leaky_relu(x, 0.01), w, b, padding=k // 2))
    y = torch.empty(B, T, device="cuda")
    sib.ops.conv1d_cout1(to_frame_major(x).cuda().to(torch.bfloat16).contiguous(), w[0].t().contiguous().cuda(), b.cuda(), y,
                         k, k // 2, 0.01, sib.ops.ACT_TANH)
    assert max_abs(ref[:, 0], y.cpu()) < 2e-5


def test_new_operators_fail_loudly(sib):
    """Unsupported configurations of the fused operators raise SibError - nothing falls back to another path."""
    ops = sib.ops
    assert ops.resunit_supported(32, 11, 5) and ops.resunit_supported(64, 11, 5) and ops.resunit_supported(64, 3, 1, True, True)
    assert ops.resunit_supported(128, 3, 5)           # C = 128, k = 3: wide CTA-pair units
    assert not ops.resunit_supported(128, 7, 1)       # ... but 448 KB of weights do not fit: the conv kernel runs k = 7 / 11
    assert not ops.resunit_supported(128, 3, 1, True, False) and not ops.resunit_supported(128, 3, 1, False, True)
    assert not ops.resunit_supported(256, 3, 1)       # wider stages use the conv kernel
    assert not ops.resunit_supported(64, 4, 1)        # even kernel sizes have no "same" padding
    x = torch.zeros(1, 64, 128, device="cuda", dtype=torch.bfloat16)
    w = torch.zeros(1, 2, 7, 128, 64, device="cuda", dtype=torch.bfloat16)
    b = torch.zeros(128, device="cuda")
    with pytest.raises(sib.SibError, match="k <= 3"):
        ops.resunit(x, w, b, w, b, torch.empty_like(x), 7, 1)
    x32 = torch.zeros(1, 64, 32, device="cuda", dtype=torch.bfloat16)
    w32 = torch.zeros(1, 1, 3, 32, 32, device="cuda", dtype=torch.bfloat16)
    b32 = torch.zeros(32, device="cuda")
    with pytest.raises(sib.SibError, match="in-place"):
        ops.resunit(x32, w32, b32, w32, b32, x32, 3, 1)
    with pytest.raises(sib.SibError, match="slopes"):
        ops.resunit(x32, w32, b32, w32, b32, torch.empty_like(x32), 3, 1, slope_in=1.5)
    with pytest.raises(sib.SibError):
        ops.conv0_gn_stats(torch.zeros(1, 4000, device="cuda"), torch.zeros(512, 3, device="cuda"), None, 512, 3, 2, 100, 1e-5,
                           torch.empty(1, 512, device="cuda"), torch.empty(1, 512, device="cuda"))
    # PDL switch round-trips
    prev = ops.set_pdl(False)
    assert ops.set_pdl(prev) is False


PRE_ACT_CASES = [
    # B, T, C_in, C_out, k, dil  (stride 1, "same" padding): single CTAs, CTA pairs, resident and streamed weights, N tiles
    (2, 300, 128, 128, 3, 1),
    (3, 1000, 128, 128, 7, 5),
    (2, 2000, 128, 128, 11, 3),
    (2, 700, 256, 256, 11, 5),
    (2, 517, 256, 256, 3, 3),
    (2, 640, 256, 1024, 3, 1),     # poly-phase upsampling conv: several N tiles per halo tile
    (2, 900, 64, 64, 7, 1),
    (2, 900, 32, 32, 3, 3),        # tap-blocked K rows (cc = 32)
]


@pytest.mark.parametrize("case", PRE_ACT_CASES)
def test_conv1d_bf16_pre_activation_in_shared_memory(sib, case):
    """leaky-relu applied to the landed A tile inside the kernel == conv of the activated tensor (models.py:37,109)."""
    B, T, Cin, Cout, k, d = case
    pad = (k * d - d) // 2
    x = _bf(_rand(B, Cin, T, seed=11))
    w = _bf(_rand(Cout, Cin, k, seed=12, scale=1.0 / math.sqrt(Cin * k)))
    b = _rand(Cout, seed=13, scale=0.1)
    taps = sib.ops.conv_taps(k, d, pad)
    assert sib.ops.conv_pre_act_supported(B, T, Cin, Cout, taps)
    xa = _bf(F.leaky_relu(x, 0.1))                      # what the producer's y_act output would hold
    ref = to_frame_major(F.conv1d(xa, w, b, dilation=d, padding=pad))
    wk = sib.ops.to_kmajor_bf16(sib.ops.pack_conv_weight(w.cuda()))
    xd = to_frame_major(x).to(torch.bfloat16).cuda()
    y = torch.full((B, T, Cout), float("nan"), dtype=torch.bfloat16, device="cuda")
    sib.ops.conv1d(xd, wk, b.cuda(), y, taps, pre_slope=0.1)
    # the in-kernel activation rounds slope * x twice (bf16 product + bf16 correction): within one bf16 ulp of xa
    _close(ref, y.float().cpu(), f"pre-act {case}")
    # and it matches the y_act route bit for bit wherever x >= 0 everywhere (identity on non-negative inputs)
    xp = x.abs()
    y1 = torch.empty_like(y)
    y2 = torch.empty_like(y)
    xpd = to_frame_major(xp).to(torch.bfloat16).cuda()
    sib.ops.conv1d(xpd, wk, b.cuda(), y1, taps, pre_slope=0.1)
    sib.ops.conv1d(xpd, wk, b.cuda(), y2, taps)
    assert torch.equal(y1, y2)


def test_conv1d_bf16_pre_activation_rejected_outside_halo_mode(sib):
    x = torch.zeros(1, 256, 768, dtype=torch.bfloat16, device="cuda")
    w = sib.ops.to_kmajor_bf16(sib.ops.pack_linear_weight(torch.zeros(768, 768, device="cuda")))
    y = torch.empty(1, 256, 768, dtype=torch.bfloat16, device="cuda")
    assert not sib.ops.conv_pre_act_supported(1, 256, 768, 768, [0])
    with pytest.raises(sib.SibError, match="halo"):
        sib.ops.conv1d(x, w, None, y, [0], pre_slope=0.1)
