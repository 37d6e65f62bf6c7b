#!/usr/bin/env python
"""Micro-benchmark of the fused ResBlock1 unit (sib_resunit_bf16) at the HiFi-GAN V1 stage-3 / stage-4 shapes of the
headline workload (CUDA events, L2 flushed).  Usage: python scripts/resunit_microbench.py [--only NAME] [--iters N]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import speech_inpainting_b200 as sib  # noqa: E402

ops = sib.ops
# name: (B, T, C, k, dil, accumulate, y_act)
SHAPES = {
    "s4k3d1": (32, 88064, 32, 3, 1, False, False),
    "s4k7d3": (32, 88064, 32, 7, 3, False, False),
    "s4k11d5": (32, 88064, 32, 11, 5, True, True),
    "s3k3d1": (32, 44032, 64, 3, 1, False, False),
    "s3k7d3": (32, 44032, 64, 7, 3, False, False),
    "s3k11d5": (32, 44032, 64, 11, 5, True, False),
    "s2k3d1": (32, 22016, 128, 3, 1, False, False),      # stage 2, C = 128: wide CTA-pair variant
    "s2k3d5": (32, 22016, 128, 3, 5, False, False),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    ap.add_argument("--iters", type=int, default=5)
    a = ap.parse_args()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for name, (B, T, C, k, dil, acc, yact) in SHAPES.items():
        if a.only and name not in a.only.split(","):
            continue
        x = torch.randn(B, T, C, device="cuda").to(torch.bfloat16)
        w1 = ops.to_kmajor_bf16(ops.pack_conv_weight(torch.randn(C, C, k, device="cuda") * 0.05))
        w2 = ops.to_kmajor_bf16(ops.pack_conv_weight(torch.randn(C, C, k, device="cuda") * 0.05))
        b1, b2 = torch.randn(C, device="cuda"), torch.randn(C, device="cuda")
        y = torch.zeros(B, T, C, dtype=torch.bfloat16, device="cuda")
        y2 = torch.empty_like(y) if yact else None
        ts = []
        for it in range(a.iters + 2):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.resunit(x, w1, b1, w2, b2, y, k, dil, y_act=y2, accumulate=acc)
            e1.record()
            torch.cuda.synchronize()
            if it >= 2:
                ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        fl = 2 * 2.0 * B * T * C * C * k
        byt = 2.0 * B * T * C * (2 + int(acc) + int(yact))
        print(f"{name:9s} {ms:8.3f} ms  {fl / ms / 1e9:7.1f} TFLOP/s  {byt / ms / 1e6:7.1f} GB/s (algorithmic bytes {byt / 1e6:.0f} MB)")


if __name__ == "__main__":
    main()
