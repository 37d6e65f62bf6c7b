"""-m gpu: every C-ABI kernel against a plain PyTorch fp32 statement of the same op (through ctypes)."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from util import max_abs, snr_db, to_frame_major

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sib():
    import speech_inpainting_b200 as m
    m._load_lib()
    return m


def _rand(*shape, seed=0, scale=1.0):
    return (torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale)


CONV_CASES = [
    # B, T, Cin, Cout, k, stride, dil, pad, groups
    (2, 300, 64, 96, 3, 1, 1, 1, 1),
    (1, 517, 128, 128, 7, 1, 3, 9, 1),
    (3, 130, 32, 32, 11, 1, 5, 25, 1),
    (2, 401, 512, 512, 3, 2, 1, 0, 1),   # HuBERT conv1-4 shape
    (2, 200, 512, 512, 2, 2, 1, 0, 1),   # HuBERT conv5-6 shape
    (2, 99, 768, 768, 128, 1, 1, 64, 16),  # pos-conv (k=128, groups=16) - output trimmed to T
    (1, 50, 80, 512, 7, 1, 1, 3, 1),     # conv_pre
    (2, 77, 20, 36, 5, 1, 2, 4, 2),      # odd sizes -> scalar path
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv1d_f32(sib, case):
    B, T, Cin, Cout, k, s, d, p, g = case
    x = _rand(B, Cin, T, seed=1)
    w = _rand(Cout, Cin // g, k, seed=2, scale=1.0 / math.sqrt(Cin // g * k))
    b = _rand(Cout, seed=3, scale=0.1)
    ref = F.conv1d(F.leaky_relu(x, 0.1), w, b, stride=s, dilation=d, padding=p, groups=g)
    t_out = min(ref.shape[-1], T) if (k == 128) else ref.shape[-1]
    ref = ref[..., :t_out]
    res = _rand(B, Cout, t_out, seed=4)
    ref = torch.tanh(ref + res) if Cout == 96 else F.gelu(ref + res)
    xd, wd = to_frame_major(x).cuda(), sib.ops.pack_conv_weight(w.cuda(), g)
    y = torch.empty(B, t_out, Cout, device="cuda")
    sib.ops.conv1d(xd, wd, b.cuda(), y, sib.ops.conv_taps(k, d, p), stride=s, groups=g, pre_slope=0.1,
                   residual=to_frame_major(res).cuda(), post_act=sib.ops.ACT_TANH if Cout == 96 else sib.ops.ACT_GELU)
    assert max_abs(to_frame_major(ref), y.cpu()) < 2e-4
    assert snr_db(to_frame_major(ref), y.cpu()) > 80


def test_conv1d_accumulate_scale_and_res_after_act(sib):
    B, T, Cc = 2, 150, 64
    x, w, b = _rand(B, Cc, T, seed=5), _rand(Cc, Cc, 3, seed=6, scale=0.1), _rand(Cc, seed=7)
    y0 = _rand(B, T, Cc, seed=8)
    y = y0.clone().cuda()
    sib.ops.conv1d(to_frame_major(x).cuda(), sib.ops.pack_conv_weight(w.cuda()), b.cuda(), y, sib.ops.conv_taps(3, 1, 1),
                   accumulate=True, out_scale=1.0 / 3)
    ref = (to_frame_major(F.conv1d(x, w, b, padding=1)) + y0) / 3
    assert max_abs(ref, y.cpu()) < 1e-5
    res = _rand(B, T, Cc, seed=9)
    y2 = torch.empty(B, T, Cc, device="cuda")
    sib.ops.conv1d(to_frame_major(x).cuda(), sib.ops.pack_conv_weight(w.cuda()), b.cuda(), y2, sib.ops.conv_taps(3, 1, 1),
                   post_act=sib.ops.ACT_GELU, residual=res.cuda(), res_after_act=True)
    ref2 = F.gelu(to_frame_major(F.conv1d(x, w, b, padding=1))) + res
    assert max_abs(ref2, y2.cpu()) < 1e-5


@pytest.mark.parametrize("ks", [(16, 8), (4, 2), (11, 5), (8, 4)])
def test_conv_transpose_polyphase(sib, ks):
    k, s = ks
    B, T, Cin, Cout = 2, 37, 64, 32
    x, w, b = _rand(B, Cin, T, seed=1), _rand(Cin, Cout, k, seed=2, scale=0.1), _rand(Cout, seed=3)
    ref = F.conv_transpose1d(F.leaky_relu(x, 0.1), w, b, stride=s, padding=(k - s) // 2)
    assert ref.shape[-1] == T * s
    wp, bp, taps = sib.ops.pack_conv_transpose(w.cuda(), b.cuda(), s, (k - s) // 2)
    y = torch.empty(B, T * s, Cout, device="cuda")
    sib.ops.conv1d(to_frame_major(x).cuda(), wp, bp, y.view(B, T, s * Cout), taps, pre_slope=0.1)
    assert max_abs(to_frame_major(ref), y.cpu()) < 1e-5


def test_linear_f32(sib):
    M, K, N = 333, 768, 80
    x, w, b = _rand(M, K, seed=1), _rand(N, K, seed=2, scale=0.05), _rand(N, seed=3)
    y = torch.empty(M, N, device="cuda")
    sib.ops.linear(x.cuda(), sib.ops.pack_linear_weight(w.cuda()), b.cuda(), y)
    assert max_abs(F.linear(x, w, b), y.cpu()) < 1e-4


def test_conv_cout1_tanh(sib):
    B, T, Cc = 2, 1000, 32
    x, w, b = _rand(B, Cc, T, seed=1), _rand(1, Cc, 7, seed=2, scale=0.1), _rand(1, seed=3)
    ref = torch.tanh(F.conv1d(F.leaky_relu(x), w, b, padding=3))
    y = torch.empty(B, T, device="cuda")
    sib.ops.conv1d_cout1(to_frame_major(x).cuda(), w[0].t().contiguous().cuda(), b.cuda(), y, 7, 3, 0.01, sib.ops.ACT_TANH)
    assert max_abs(ref[:, 0], y.cpu()) < 1e-5


@pytest.mark.parametrize("c", [80, 512, 768, 1024, 1500])
def test_layernorm(sib, c):
    x, r = _rand(37, 5, c, seed=1, scale=2.0), _rand(37, 5, c, seed=2)
    g, b = 1 + 0.1 * _rand(c, seed=3), _rand(c, seed=4)
    y = torch.empty(37, 5, c, device="cuda")
    sib.ops.layernorm(x.cuda(), g.cuda(), b.cuda(), y, 1e-5, residual=r.cuda(), post_act=sib.ops.ACT_GELU)
    ref = F.gelu(F.layer_norm(x + r, (c,), g, b, 1e-5))
    assert max_abs(ref, y.cpu()) < 2e-5


@pytest.mark.parametrize("T,padded", [(99, False), (199, True), (64, False), (130, True)])
def test_attention(sib, T, padded):
    B, nh, d = 3, 4, 64
    H = nh * d
    qkv = _rand(B, T, 3 * H, seed=T)
    kl = torch.tensor([T, T - 17, 5], dtype=torch.int32) if padded else None
    q, k, v = [t.view(B, T, nh, d).transpose(1, 2) for t in qkv.split(H, dim=-1)]
    s = torch.matmul(q, k.transpose(2, 3)) * d ** -0.5
    if padded:
        km = torch.arange(T)[None, :] < kl[:, None]
        s = s.masked_fill(~km[:, None, None, :], torch.finfo(torch.float32).min)
    ref = torch.matmul(F.softmax(s, -1), v).transpose(1, 2).reshape(B, T, H)
    out = torch.empty(B, T, H, device="cuda")
    sib.ops.attention(qkv.cuda(), None if kl is None else kl.cuda(), out, nh)
    assert max_abs(ref, out.cpu()) < 2e-5


@pytest.mark.parametrize("mode", ["group", "layer"])
def test_conv0(sib, mode):
    B, N, Cc = 2, 4000, 512
    x = _rand(B, N, seed=1)
    w, g, be = _rand(Cc, 1, 10, seed=2, scale=0.4), 1 + 0.1 * _rand(Cc, seed=3), 0.1 * _rand(Cc, seed=4)
    bias = 0.1 * _rand(Cc, seed=5) if mode == "layer" else None
    h = F.conv1d(x[:, None], w, bias, stride=5)
    t0 = h.shape[-1]
    y = torch.empty(B, t0, Cc, device="cuda")
    wd = w.reshape(Cc, 10).contiguous().cuda()
    if mode == "group":
        ref = F.gelu(F.group_norm(h, Cc, g, be, 1e-5))
        nt = sib.ops.conv0_num_tiles(t0)
        part = torch.empty(B, nt, Cc, 2, device="cuda")
        mean, rstd = torch.empty(B, Cc, device="cuda"), torch.empty(B, Cc, device="cuda")
        sib.ops.conv0(0, x.cuda(), wd, None, Cc, 10, 5, t0, partial=part)
        sib.ops.gn_finalize(part, B, nt, Cc, t0, 1e-5, mean, rstd)
        # closed-form statistics from the waveform's lag sums must agree with the evaluated ones
        mean2, rstd2 = torch.empty_like(mean), torch.empty_like(rstd)
        sib.ops.conv0_gn_stats(x.cuda(), wd, None, Cc, 10, 5, t0, 1e-5, mean2, rstd2)
        assert max_abs(h.mean(-1), mean2.cpu()) < 1e-5
        assert max_abs(1.0 / torch.sqrt(h.var(-1, unbiased=False) + 1e-5), rstd2.cpu()) < 1e-4
        assert max_abs(mean.cpu(), mean2.cpu()) < 1e-5 and max_abs(rstd.cpu(), rstd2.cpu()) < 1e-4
        sib.ops.conv0(1, x.cuda(), wd, None, Cc, 10, 5, t0, mean=mean2, rstd=rstd2, gamma=g.cuda(), beta=be.cuda(), y=y)
        yb = torch.empty(B, t0, Cc, device="cuda", dtype=torch.bfloat16)
        sib.ops.conv0(1, x.cuda(), wd, None, Cc, 10, 5, t0, mean=mean2, rstd=rstd2, gamma=g.cuda(), beta=be.cuda(), y=yb)
        assert max_abs(to_frame_major(ref), yb.float().cpu()) < 2e-2
    else:
        ref = F.gelu(F.layer_norm(h.transpose(1, 2), (Cc,), g, be, 1e-5)).transpose(1, 2)
        sib.ops.conv0(2, x.cuda(), wd, bias.cuda(), Cc, 10, 5, t0, y=y)
        sib.ops.layernorm(y, g.cuda(), be.cuda(), y, 1e-5, post_act=sib.ops.ACT_GELU)
    assert max_abs(to_frame_major(ref), y.cpu()) < 3e-5


def test_znorm_and_zero_ranges(sib):
    from oracle import glue_ref
    B, N = 3, 32000
    x = 0.1 * _rand(B, N, seed=1) + 0.01
    lengths = torch.tensor([N, N - 5000, 1234], dtype=torch.int32)
    y = torch.empty(B, N, device="cuda")
    sib.ops.znorm(x.cuda(), y, None, 1e-7)
    assert max_abs(glue_ref.processor_znorm(x), y.cpu()) < 2e-5
    sib.ops.znorm(x.cuda(), y, lengths.cuda(), 1e-7)
    assert max_abs(glue_ref.processor_znorm(x, lengths), y.cpu()) < 2e-5
    sib.ops.znorm(x.cuda(), y, None, 1e-5)
    assert max_abs(glue_ref.fairseq_layer_norm(x[0]), y[0].cpu()) < 2e-5
    # every kernel variant (8 / 16 / 32 samples per thread in the 8-CTA cluster kernel, one-CTA fallback), ragged lengths,
    # and batch invariance: an utterance normalises to the same bits alone or inside a batch
    for n in (64000, 96000, 160000, 300000, 777):
        xx = 0.1 * _rand(4, n, seed=n) + 0.02
        ll = torch.tensor([n, n // 2 + 3, 1, n - 1], dtype=torch.int32)
        yy = torch.empty(4, n, device="cuda")
        sib.ops.znorm(xx.cuda(), yy, ll.cuda(), 1e-7)
        assert max_abs(glue_ref.processor_znorm(xx, ll), yy.cpu()) < 3e-5, n
        y1 = torch.empty(1, n, device="cuda")
        sib.ops.znorm(xx[1:2].cuda().contiguous(), y1, ll[1:2].cuda(), 1e-7)
        assert torch.equal(y1[0], yy[1])
    # zero ranges: bit exact, numpy slice semantics incl. empty / clamped ranges
    lo, hi = [14480, 31000, 500], [17599, 40000, 400]
    xd = x.clone().cuda()
    sib.ops.zero_ranges(xd, torch.tensor(lo, dtype=torch.int32).cuda(), torch.tensor(hi, dtype=torch.int32).cuda())
    for b in range(B):
        assert np.array_equal(glue_ref.apply_zero_range(x[b].numpy(), lo[b], hi[b]), xd[b].cpu().numpy())
    xd = x.clone().cuda()
    sib.ops.zero_ranges(xd, torch.tensor([24000] * B, dtype=torch.int32).cuda(),
                        torch.tensor([24000 + 6400] * B, dtype=torch.int32).cuda(), add_eps=1e-6)
    ref, fs = glue_ref.ida_mask(x[0].numpy(), 6400)
    assert fs == 24000 and np.array_equal(ref.astype(np.float32), xd[0].cpu().numpy())


def test_glue_gather_assign_paste(sib, golden_dir):
    from oracle import glue_ref
    from oracle.params import make_codebook
    gold = np.load(f"{golden_dir}/glue_golden.npz")
    for K in (100, 500):
        C = make_codebook(80, K, seed=77)
        vals = _rand(3, 10, 80, seed=K)
        cc, center = glue_ref.codebook_center(C)
        labels = torch.empty(30, dtype=torch.int64, device="cuda")
        sib.ops.cos_argmax(vals.view(30, 80).cuda(), cc.contiguous().cuda(), labels)
        assert np.array_equal(labels.cpu().numpy().reshape(3, 10), gold[f"cos_sim_pred_{K}"])  # reference labels, exact
        mel = _rand(1, 80, 50, seed=5).cuda()
        i32 = lambda v: torch.tensor(v, dtype=torch.int32).cuda()
        sib.ops.paste_centroids(mel, cc.contiguous().cuda(), center.cuda(), labels[:10].contiguous(), i32([7]), i32([10]), i32([0]))
        assert max_abs(torch.from_numpy(gold[f"paste_{K}"]), mel.cpu()) < 1e-6
    # ragged gather
    out = _rand(4, 60, 80, seed=3)
    pos, ln = [0, 13, 59, 20], [5, 20, 1, 0]
    off = [0, 5, 25, 26]
    got = torch.empty(26, 80, device="cuda")
    i32 = lambda v: torch.tensor(v, dtype=torch.int32).cuda()
    sib.ops.gather_frames(out.cuda(), i32(pos), i32(ln), i32(off), got)
    ref = torch.cat(glue_ref.gather_mask_frames(out, pos, ln), 0)
    assert torch.equal(ref, got.cpu())
    # k-means assignment
    f, mu = _rand(199, 768, seed=1), _rand(500, 768, seed=2)
    lab = torch.empty(199, dtype=torch.int64, device="cuda")
    sib.ops.l2_argmin(f.cuda(), mu.cuda(), lab)
    assert torch.equal(glue_ref.kmeans_predict(f, mu), lab.cpu())
    # ... and at scale through the fp32 GEMM + row argmax (sklearn's float32 form): same labels as sklearn's golden vectors
    gk = np.load(f"{golden_dir}/kmeans_golden.npz")
    for K, H, M in ((100, 768, 400), (500, 768, 600), (500, 1024, 300)):
        g = torch.Generator().manual_seed(K + H)
        mu = torch.randn(K, H, generator=g) * 0.5
        idx = torch.randint(0, K, (M,), generator=g)
        f = torch.cat([mu[idx[: M // 2]] + 0.3 * torch.randn(M // 2, H, generator=g), torch.randn(M - M // 2, H, generator=g) * 0.5])
        packed, bias = sib.ops.kmeans_pack(mu.cuda())
        lab = torch.empty(M, dtype=torch.int64, device="cuda")
        sib.ops.kmeans_assign(f.cuda(), packed, bias, lab)
        assert np.array_equal(lab.cpu().numpy(), gk[f"labels_{K}_{H}"])          # sklearn.cluster.KMeans.predict
        lab2 = torch.empty(M, dtype=torch.int64, device="cuda")
        sib.ops.l2_argmin(f.cuda(), mu.cuda(), lab2)
        assert torch.equal(lab, lab2)
    # ties resolve to the lowest index
    sc = torch.zeros(3, 37, device="cuda"); sc[1, 5] = sc[1, 20] = 2.0; sc[2, 36] = 1.0
    lab = torch.empty(3, dtype=torch.int64, device="cuda")
    sib._load_lib().sib_row_argmax_f32(sc.data_ptr(), 3, 37, lab.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert lab.cpu().tolist() == [0, 5, 36]


@pytest.mark.parametrize("T", [37, 100, 200])
def test_extend_mel(sib, golden_dir, T):
    gold = torch.from_numpy(np.load(f"{golden_dir}/glue_golden.npz")[f"extend_mel_{T}"])
    spec = _rand(2, 80, T, seed=T)  # same seed as make_golden
    spec = torch.randn(2, 80, T, generator=torch.Generator().manual_seed(T))
    out = sib.extend_mel(spec.cuda())
    assert out.shape == gold.shape
    assert max_abs(gold, out.cpu()) < 5e-5
    fm = torch.empty(2, gold.shape[-1], 80, device="cuda")
    sib.ops.extend_mel(spec.cuda(), fm, frame_major=True)
    assert torch.equal(fm.transpose(1, 2), out)


def test_transpose_embed_pack(sib):
    from oracle import hifigan_ref
    x = _rand(3, 45, 70, seed=1)
    out = torch.empty(3, 70, 45, device="cuda")
    sib.ops.transpose(x.cuda(), out)
    assert torch.equal(x.transpose(1, 2), out.cpu())
    B, T, E = 2, 16, 128
    code = torch.randint(0, 500, (B, T), generator=torch.Generator().manual_seed(1))
    zp = torch.randint(0, 20, (B, T // 4), generator=torch.Generator().manual_seed(2))
    emb = _rand(B, E, seed=3)
    p = {"emb_c.weight": _rand(500, E, seed=4), "emb_p.weight": _rand(20, E, seed=5)}
    ref = hifigan_ref.code_generator_front(p, code, zp, emb)
    got = torch.empty(B, T, 3 * E, device="cuda")
    sib.ops.embed_concat(code.cuda(), zp.cuda(), emb.cuda(), p["emb_c.weight"].cuda(), p["emb_p.weight"].cuda(), got)
    assert torch.equal(to_frame_major(ref), got.cpu())
    y = torch.cat([torch.tanh(_rand(1, 1, 5000, seed=6) * 3), torch.tensor([[[1.0, -1.0, 0.99999, -0.99999, 0.0]]])], -1)
    out16 = torch.empty(y.shape, dtype=torch.int16, device="cuda")
    sib.ops.pack_int16(y.cuda(), out16)
    assert np.array_equal(hifigan_ref.to_int16(y), out16.cpu().numpy().squeeze())


@pytest.mark.parametrize("hop,pad,fmax", [(256, None, None), (441, 312, 8000)])
def test_mel_spectrogram(sib, golden_dir, hop, pad, fmax):
    gold = torch.from_numpy(np.load(f"{golden_dir}/mel_golden.npz")[f"mel_hop{hop}"])
    y = 0.3 * torch.randn(2, 22050, generator=torch.Generator().manual_seed(3)).clamp(-3, 3)
    out = sib.mel_spectrogram(y.cuda(), hop_size=hop, fmax=fmax, pad=pad)
    assert out.shape == gold.shape
    assert max_abs(gold, out.cpu()) < 2e-3   # log of fp32 magnitudes; bulk error is ~1e-5
    assert float((gold - out.cpu()).abs().mean()) < 2e-5


def test_errors_are_loud(sib):
    x = torch.zeros(1, 8, 16)
    with pytest.raises(sib.SibError):
        sib.extend_mel(x)  # CPU tensor: no fallback
    d = sib.ops.make_desc(1, 8, 8, 16, 16, [0], groups=3)
    xd = torch.zeros(1, 8, 16, device="cuda")
    with pytest.raises(sib.SibError, match="groups"):
        sib.ops.conv1d(xd, xd, None, xd.clone(), [0], groups=3)


def test_masked_feature_mel_front_end(sib):
    """predict.py:99-104 on the device (22 kHz zero-mask -> normalize * 0.95 -> get_mel) vs the oracle."""
    from oracle import glue_ref, mel_ref
    g = torch.Generator().manual_seed(11)
    B, S = 3, 44100
    wave = (0.2 * torch.randn(B, S, generator=g)).clamp(-1, 1)
    wave[2] = 0.0                                             # silent utterance: normalize leaves it untouched
    ranges = [glue_ref.iea_mask_indices(0.9, 1.1)["zero22"], (0, 0), (100, 5000)]
    xd = wave.cuda()
    yd = torch.empty_like(xd)
    lo = torch.tensor([r[0] for r in ranges], dtype=torch.int32).cuda()
    hi = torch.tensor([r[1] for r in ranges], dtype=torch.int32).cuda()
    sib.ops.mask_peak_normalize(xd, yd, lo, hi, 0.95)
    for b in range(B):
        w = wave[b].numpy().copy()
        w[ranges[b][0]:ranges[b][1]] = 0
        ref = (mel_ref.peak_normalize(w) * np.float32(0.95)).astype(np.float32)
        assert np.array_equal(ref, yd[b].cpu().numpy()), f"utterance {b}: masked / normalised samples differ"   # bit exact
    mel = sib.masked_feature_mel(wave, ranges)
    assert mel.shape == (B, 80, S // 441)
    for b in range(2):
        ref = mel_ref.masked_feature_mel(wave[b].numpy(), *ranges[b])
        assert max_abs(ref[0], mel[b].cpu()) < 2e-3 and float((ref[0] - mel[b].cpu()).abs().mean()) < 2e-5


def test_device_metrics(sib, golden_dir):
    """SI-SDR (I_ea/metrics.py:127-141) and mel-L1 reduced on the device vs their CPU restatements."""
    from oracle import mel_ref, metrics_ref
    g = torch.Generator().manual_seed(21)
    ref = 0.3 * torch.randn(3, 22050, generator=g)
    est = ref * torch.tensor([[1.0], [0.5], [2.0]]) + torch.tensor([[0.01], [0.1], [0.5]]) * torch.randn(3, 22050, generator=g)
    got = sib.si_sdr(est.cuda(), ref.cuda()).cpu()
    golden = np.load(f"{golden_dir}/metrics_golden.npz")["sisdr"]   # the reference's own method on these inputs
    for b in range(3):
        want = metrics_ref.sisdr(est[b].numpy().astype(np.float64), ref[b].numpy().astype(np.float64))
        assert abs(want - float(got[b])) < 1e-3 and abs(golden[b] - float(got[b])) < 1e-3, (b, want, float(got[b]))
    assert got[0] > got[1] > got[2]
    # ragged lengths == each utterance alone; batch composition does not change a bit (fixed reduction order)
    lens = [22050, 10000, 333]
    rag = sib.si_sdr(est.cuda(), ref.cuda(), lengths=lens).cpu()
    for b, n in enumerate(lens):
        alone = sib.si_sdr(est[b, :n].cuda(), ref[b, :n].cuda()).cpu()
        assert float(alone) == float(rag[b])
        assert abs(float(alone) - metrics_ref.sisdr(est[b, :n].numpy().astype(np.float64), ref[b, :n].numpy().astype(np.float64))) < 1e-3
    # a perfect estimate saturates at the eps floor like the reference (no inf / nan)
    assert torch.isfinite(sib.si_sdr(ref.cuda(), ref.cuda())).all()
    l1 = sib.mel_l1(est.cuda(), ref.cuda())
    assert abs(l1 - mel_ref.mel_l1(est, ref)) < 1e-4
    per = sib.mel_l1(est.cuda(), ref.cuda(), per_utterance=True)
    assert per.shape == (3,) and abs(float(per.double().mean()) - l1) < 1e-9
    one = sib.mel_l1(est[1:2].cuda(), ref[1:2].cuda(), per_utterance=True)
    assert float(one[0]) == float(per[1])
    with pytest.raises(sib.SibError):
        sib.si_sdr(est, ref)


RESAMPLE_CASES = [(16000, 22050), (22050, 16000), (48000, 16000), (24000, 22050), (16000, 16000)]


@pytest.mark.parametrize("rates", RESAMPLE_CASES)
def test_resample_vs_oracle_and_golden(sib, golden_dir, rates):
    """the poly-phase resampler vs the float64 oracle and the torchaudio-generated golden vectors."""
    from oracle import resample_ref as R
    o, n = rates
    g = np.load(f"{golden_dir}/resample_golden.npz")
    x = torch.from_numpy(g["x"])
    y = sib.resample(x.cuda(), o, n).cpu().numpy()
    want = R.resample(g["x"], o, n)
    assert y.shape == want.shape and np.abs(y - want).max() < 2e-6
    if f"y_{o}_{n}" in g:
        assert np.abs(y - g[f"y_{o}_{n}"]).max() < 2e-6
    # int16 PCM input (x / 32768 on load) and a ragged, zero-padded batch: every row == that utterance alone
    pcm = (x * 20000).round().clamp(-32768, 32767).to(torch.int16)
    lens = [12000, 7001, 1]
    batch = torch.zeros(3, 12000, dtype=torch.int16)
    for b, ln in enumerate(lens):
        batch[b, :ln] = pcm[:ln]
    yb = sib.resample(batch.cuda(), o, n, lengths=lens).cpu().numpy()
    for b, ln in enumerate(lens):
        alone = sib.resample(pcm[:ln].cuda(), o, n).cpu().numpy()
        m = R.out_length(ln, o, n)
        assert alone.shape == (m,) and np.array_equal(yb[b, :m], alone) and not yb[b, m:].any()
        ref = R.resample(R.pcm16_to_float(pcm[:ln].numpy()), o, n)
        assert np.abs(alone - ref).max() < 2e-6


def test_resample_long_batch_property(sib):
    """BASELINE-size property (32 x 4 s): 16 k -> 22.05 k -> 16 k returns the band-limited input; a constant stays put."""
    g = torch.Generator().manual_seed(4)
    t = torch.arange(64000) / 16000.0
    fade = torch.ones(64000)
    fade[:4000] = torch.hann_window(8000, periodic=False)[:4000]
    fade[-4000:] = torch.hann_window(8000, periodic=False)[4000:]
    x = torch.stack([0.3 * torch.sin(2 * math.pi * (200.0 + 150.0 * b) * t + b) for b in range(32)]) * fade   # <= 4.85 kHz
    up = sib.resample(x.cuda(), 16000, 22050)
    assert up.shape == (32, 88200)
    back = sib.resample(up, 22050, 16000).cpu()
    assert back.shape == (32, 64000)
    # pass-band gain error of two 64-crossing Kaiser filters: ~1e-3 relative (the float64 oracle shows the same)
    assert float((back - x).abs().max()) < 1e-3
    ones = sib.resample(torch.ones(1, 8000).cuda(), 16000, 22050).cpu()
    assert float((ones[0, 300:-300] - 1.0).abs().max()) < 1e-4


def test_load_wav_batch(sib, tmp_path):
    """The librosa.load(sr=22050) / librosa.load(sr=16000) pair (I_ea/predict.py:79-80) for a batch of files."""
    from oracle import resample_ref as R
    rng = np.random.default_rng(3)
    files, pcms = [], []
    for i, (n, sr) in enumerate([(9000, 22050), (7000, 22050), (5000, 16000)]):
        pcm = (rng.standard_normal(n) * 4000).astype(np.int16)
        sib.write_wav(tmp_path / f"u{i}.wav", pcm, sr)
        files.append(tmp_path / f"u{i}.wav")
        pcms.append((pcm, sr))
    out = sib.load_wav_batch(files[:2])                      # same source rate: one launch per target rate
    for sr in (16000, 22050):
        wave, lens = out[sr]
        assert wave.is_cuda and lens.tolist() == [R.out_length(p.shape[0], s, sr) for p, s in pcms[:2]]
        for b, (p, s) in enumerate(pcms[:2]):
            ref = R.resample(R.pcm16_to_float(p), s, sr)
            got = wave[b].cpu().numpy()
            assert np.abs(got[:len(ref)] - ref).max() < 2e-6 and not got[len(ref):].any()
    mixed = sib.load_wav_batch(files, target_srs=(16000,))   # mixed source rates
    wave, lens = mixed[16000]
    for b, (p, s) in enumerate(pcms):
        ref = R.resample(R.pcm16_to_float(p), s, 16000)
        assert np.abs(wave[b, :len(ref)].cpu().numpy() - ref).max() < 2e-6
    assert lens.tolist()[2] == 5000 and np.array_equal(wave[2, :5000].cpu().numpy(), pcms[2][0].astype(np.float32) / 32768)


def test_linear_skinny_head(sib):
    """CustomModel head (Linear(H, 80), I_ea/model.py:75-78) on gathered frames: row-wise kernel vs torch fp32; every row is
    computed the same way whatever the batch (bit-identical alone or among 320 rows)."""
    for M, K, N in ((320, 768, 80), (1, 1024, 80), (37, 768, 100), (5, 64, 128)):
        x, w, b = _rand(M, K, seed=31), _rand(N, K, seed=32, scale=K ** -0.5), _rand(N, seed=33)
        ref = x @ w.t() + b
        wd = sib.ops.pack_linear_weight(w.cuda())
        y = torch.empty(M, N, device="cuda")
        sib.ops.linear_skinny(x.cuda(), wd, b.cuda(), y)
        assert max_abs(ref, y.cpu()) < 2e-5
        y1 = torch.empty(1, N, device="cuda")
        sib.ops.linear_skinny(x[M // 2: M // 2 + 1].cuda().contiguous(), wd, b.cuda(), y1)
        assert torch.equal(y1[0], y[M // 2])
    with pytest.raises(sib.SibError):
        sib.ops.linear_skinny(torch.zeros(2, 8, device="cuda"), torch.zeros(8 * 200, device="cuda"), None, torch.zeros(2, 200, device="cuda"))
