#!/usr/bin/env python
"""Micro-benchmark of sib_conv1d_bf16 on the layer shapes of the headline workload (CUDA events, L2 flushed).
Usage: python scripts/tc_microbench.py [--only NAME] [--iters N]     (used for tuning and for ncu captures)"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import speech_inpainting_b200 as sib  # noqa: E402

ops = sib.ops
# name: (B, T, Cin, Cout, k, dil, stride, residual, y_act)
SHAPES = {
    "bigk": (1, 16384, 8192, 4096, 1, 1, 1, False, False),
    "bigconv": (32, 4096, 512, 512, 11, 1, 1, False, False),
    "qkv": (1, 6368, 768, 2304, 1, 1, 1, False, False),
    "ffn1": (1, 6368, 768, 3072, 1, 1, 1, False, False),
    "ffn1g": (1, 6368, 768, 3072, 1, 1, 1, False, False, "gelu"),
    "ffn2": (1, 6368, 3072, 768, 1, 1, 1, False, False),
    "oproj": (1, 6368, 768, 768, 1, 1, 1, False, False),
    "tiny1": (1, 256, 768, 256, 1, 1, 1, False, False),          # one pair tile: the fixed cost of a launch
    "tiny74": (1, 18944, 768, 256, 1, 1, 1, False, False),       # exactly one round of 74 pair tiles
    "tiny148": (1, 37888, 768, 256, 1, 1, 1, False, False),      # two rounds
    "hubconv1": (32, 12799, 512, 512, 3, 1, 2, False, False),
    "s1k11": (32, 2752, 256, 256, 11, 1, 1, True, True),
    "s2k1": (32, 22016, 128, 128, 1, 1, 1, False, False),          # data-movement floor of the stage-2 tiles (8 MMAs per tile)
    "s2k3c1": (32, 22016, 128, 128, 3, 1, 1, False, False),
    "s2k3c1p": (32, 22016, 128, 128, 3, 1, 1, False, False, "pre"),   # first conv of a unit: leaky-relu applied to the A tile in smem
    "s2k7c1p": (32, 22016, 128, 128, 7, 1, 1, False, False, "pre"),
    "s1k3c1p": (32, 2752, 256, 256, 3, 1, 1, False, False, "pre"),
    "s2k3c2": (32, 22016, 128, 128, 3, 1, 1, True, True),
    "s2k11c2": (32, 22016, 128, 128, 11, 5, 1, True, True),
    "s2k3c2n": (32, 22016, 128, 128, 3, 1, 1, True, False),     # second conv of a unit whose consumer activates its own input
    "s2k7c2n": (32, 22016, 128, 128, 7, 1, 1, True, False),
    "s2k11c2n": (32, 22016, 128, 128, 11, 1, 1, True, False),
    "s1k3c2n": (32, 2752, 256, 256, 3, 1, 1, True, False),
    "s1k7c2n": (32, 2752, 256, 256, 7, 1, 1, True, False),
    "s1k11c2n": (32, 2752, 256, 256, 11, 1, 1, True, False),
    "s3k3c1": (32, 44032, 64, 64, 3, 1, 1, False, False),
    "s3k7c2": (32, 44032, 64, 64, 7, 3, 1, True, True),
    "s3k11c1": (32, 44032, 64, 64, 11, 5, 1, False, False),
    "s3k11c2": (32, 44032, 64, 64, 11, 1, 1, True, True),
    "s4k3c1": (32, 88064, 32, 32, 3, 1, 1, False, False),
    "s4k11c2": (32, 88064, 32, 32, 11, 5, 1, True, True),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    ap.add_argument("--iters", type=int, default=5)
    a = ap.parse_args()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for name, spec in SHAPES.items():
        B, T, Cin, Cout, k, dil, stride, res, yact = spec[:9]
        act = ops.ACT_GELU if len(spec) > 9 and spec[9] == "gelu" else ops.ACT_NONE
        pre = dict(pre_slope=0.1, post_act=ops.ACT_LRELU, post_slope=0.1) if len(spec) > 9 and spec[9] == "pre" else {}
        if a.only and name not in a.only.split(","):
            continue
        t_out = (T - k) // stride + 1 if stride > 1 else T
        pad = 0 if stride > 1 else (k * dil - dil) // 2
        x = torch.randn(B * T * Cin + 4096, device="cuda").to(torch.bfloat16)[: B * T * Cin].view(B, T, Cin)
        w = ops.to_kmajor_bf16(ops.pack_conv_weight(torch.randn(Cout, Cin, k, device="cuda") * 0.05))
        bias = torch.randn(Cout, device="cuda")
        y = torch.empty(B, t_out, Cout, dtype=torch.bfloat16, device="cuda")
        r = torch.randn(B, t_out, Cout, device="cuda").to(torch.bfloat16) if res else None
        y2 = torch.empty_like(y) if yact else None
        taps = ops.conv_taps(k, dil, pad)
        ts = []
        for it in range(a.iters + 2):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.conv1d(x, w, bias, y, taps, stride=stride, residual=r, y_act=y2, act2_slope=0.1, **(pre or dict(post_act=act)))
            e1.record()
            torch.cuda.synchronize()
            if it >= 2:
                ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        fl = 2.0 * B * t_out * Cout * Cin * k
        byt = 2.0 * (B * T * Cin + B * t_out * Cout * (1 + int(res) + int(yact))) + w.numel() * 2
        print(f"{name:9s} {ms:8.3f} ms  {fl / ms / 1e9:7.1f} TFLOP/s  {byt / ms / 1e6:7.1f} GB/s (algorithmic bytes {byt / 1e6:.0f} MB)")


if __name__ == "__main__":
    main()
