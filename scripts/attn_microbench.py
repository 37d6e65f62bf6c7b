#!/usr/bin/env python
"""Micro-benchmark of sib_attention (bf16 tcgen05 arm) on the transformer shapes of the bench workloads
(CUDA events, L2 flushed).  Usage: python scripts/attn_microbench.py [--only NAME] [--iters N]   (also the ncu target)"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import speech_inpainting_b200 as sib  # noqa: E402

ops = sib.ops
SHAPES = {"base_4s": (32, 199, 12), "base_10s": (32, 499, 12), "large_6s": (32, 299, 16), "ida_4s": (64, 199, 12)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    ap.add_argument("--iters", type=int, default=9)
    a = ap.parse_args()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for name, (B, T, heads) in SHAPES.items():
        if a.only and name not in a.only.split(","):
            continue
        H = heads * 64
        qkv = torch.randn(B, T, 3 * H, device="cuda").to(torch.bfloat16)
        out = torch.empty(B, T, H, device="cuda", dtype=torch.bfloat16)
        ts = []
        for it in range(a.iters + 2):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.attention(qkv, None, out, heads)
            e1.record()
            torch.cuda.synchronize()
            if it >= 2:
                ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        fl = 4.0 * B * heads * T * T * 64
        print(f"{name:9s} {ms * 1e3:8.1f} us  {fl / ms / 1e9:7.1f} TFLOP/s  items {B * heads * ((T + 127) // 128)}")


if __name__ == "__main__":
    main()
