#!/usr/bin/env python
"""Micro-benchmark of the bandwidth-bound kernels (STFT/mel, z-norm, LayerNorm, conv0 + GroupNorm + GELU, conv_post,
int16 pack) against their algorithmic bytes (CUDA events, L2 flushed).  Usage: python scripts/aux_microbench.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import speech_inpainting_b200 as sib  # noqa: E402

ops = sib.ops
flush = None


def timeit(fn, iters=5):
    global flush
    if flush is None:
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for it in range(iters + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if it >= 2:
            ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def report(name, ms, nbytes):
    print(f"{name:34s} {ms * 1e3:9.1f} us   {nbytes / ms / 1e6:8.1f} GB/s   (algorithmic bytes {nbytes / 1e6:.1f} MB)")


def main():
    dev = "cuda"
    # calibration of the method (memset flush before every run leaves dirty lines that the timed kernel evicts): what a
    # plain device-to-device copy of one stage-2 activation tensor (32 x 22016 x 128 bf16 = 180 MB) reaches here
    src = torch.randn(32, 22016, 128, device=dev).to(torch.bfloat16)
    dst = torch.empty_like(src)
    report("torch copy_ 180 MB (calibration)", timeit(lambda: dst.copy_(src)), 4.0 * src.numel())
    big = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
    big2 = torch.empty_like(big)
    report("torch copy_ 1 GiB (calibration)", timeit(lambda: big2.copy_(big)), 2.0 * big.numel())
    del big, big2
    # STFT / mel: 256 x 4 s at 22.05 kHz (the mel-L1 check of a 256-utterance shard), both hop sizes
    B, S = 256, 88064
    y = torch.randn(B, S, device=dev) * 0.1
    for hop, pad, tag in ((256, 384, "mel hop 256 (mel-L1 metric)"), (441, 312, "mel hop 441 (features)")):
        out = sib.mel_spectrogram(y, hop_size=hop, pad=pad)
        ms = timeit(lambda: sib.mel_spectrogram(y, hop_size=hop, pad=pad))
        report(tag, ms, 4.0 * B * S + 4.0 * out.numel())
    # z-norm of 256 x 4 s at 16 kHz
    x = torch.randn(256, 64000, device=dev)
    xn = torch.empty_like(x)
    report("z-norm 256x64000 f32", timeit(lambda: ops.znorm(x, xn, None, 1e-7)), 8.0 * x.numel())
    # LayerNorm + residual, bf16 6368 x 768 and a 16x larger instance
    for rows in (6368, 101888):
        a = torch.randn(rows, 768, device=dev).to(torch.bfloat16)
        r = torch.randn(rows, 768, device=dev).to(torch.bfloat16)
        o = torch.empty_like(a)
        g, b = torch.ones(768, device=dev), torch.zeros(768, device=dev)
        report(f"layernorm+res bf16 {rows}x768", timeit(lambda: ops.layernorm(a, g, b, o, 1e-5, residual=r)), 6.0 * a.numel())
    # conv0 + GroupNorm + GELU (bf16 out), 32 x 4 s
    wave = torch.randn(32, 64000, device=dev)
    w0, gam, bet = torch.randn(512, 10, device=dev) * 0.3, torch.ones(512, device=dev), torch.zeros(512, device=dev)
    t0 = (64000 - 10) // 5 + 1
    mean, rstd = torch.empty(32, 512, device=dev), torch.empty(32, 512, device=dev)
    yb = torch.empty(32, t0, 512, device=dev, dtype=torch.bfloat16)
    report("conv0 GN stats (closed form)", timeit(lambda: ops.conv0_gn_stats(wave, w0, None, 512, 10, 5, t0, 1e-5, mean, rstd)), 4.0 * wave.numel())
    report("conv0 + GN + GELU apply -> bf16", timeit(lambda: ops.conv0(1, wave, w0, None, 512, 10, 5, t0, mean=mean, rstd=rstd, gamma=gam, beta=bet, y=yb)),
           4.0 * wave.numel() + 2.0 * yb.numel())
    # conv_post: bf16 [32, 88064, 32] -> f32 [32, 88064]
    xa = torch.randn(32, 88064, 32, device=dev).to(torch.bfloat16)
    wp, bp = torch.randn(7, 32, device=dev) * 0.1, torch.zeros(1, device=dev)
    yo = torch.empty(32, 88064, device=dev)
    report("conv_post (C=32 -> 1, k7, tanh)", timeit(lambda: ops.conv1d_cout1(xa, wp, bp, yo, 7, 3, 1.0, ops.ACT_TANH)), 2.0 * xa.numel() + 4.0 * yo.numel())
    # int16 pack
    yi = torch.empty(32, 88064, device=dev, dtype=torch.int16)
    report("pack int16 32x88064", timeit(lambda: ops.pack_int16(yo, yi)), 6.0 * yo.numel())
    # 8f row 3: int16 PCM at 22.05 kHz -> float32 at 16 kHz and 16 k -> 22.05 k, 256 x 4 s
    pcm = (torch.randn(256, 88200, device=dev) * 3000).to(torch.int16)
    r16 = sib.resample(pcm, 22050, 16000)
    report("resample i16 22.05k -> f32 16k", timeit(lambda: sib.resample(pcm, 22050, 16000)), 2.0 * pcm.numel() + 4.0 * r16.numel())
    r22 = sib.resample(r16, 16000, 22050)
    report("resample f32 16k -> 22.05k", timeit(lambda: sib.resample(r16, 16000, 22050)), 4.0 * r16.numel() + 4.0 * r22.numel())
    # 8f row 4: SI-SDR (two passes over both signals) and the mel-L1 reduction, 256 x 4 s
    out = torch.empty(256, device=dev)
    report("si_sdr 256x88200 (2 passes)", timeit(lambda: ops.si_sdr(r22, r22, out)), 16.0 * r22.numel())


if __name__ == "__main__":
    main()
