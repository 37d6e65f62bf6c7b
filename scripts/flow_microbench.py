#!/usr/bin/env python
"""Micro-benchmark of the dataflow launches (sib_flow) on the HuBERT-base transformer shapes at 32 x 199 rows:
(1) every linear layer / LayerNorm alone: plain, signal only, wait only (counters preset), both - the cost of the
counters themselves; (2) the chain out-proj -> LN -> FFN-in -> FFN-out -> LN -> QKV as plain PDL launches and as dataflow
launches - what the overlap buys.  CUDA events, L2 flushed, median of --iters."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import speech_inpainting_b200 as sib  # noqa: E402

ops = sib.ops
M, H, I = 6368, 768, 3072


def timed(fn, flush, iters):
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
    return sorted(ts)[len(ts) // 2] * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=9)
    ap.add_argument("--no-chain", action="store_true", help="single launches only (library variants without signalling code)")
    a = ap.parse_args()
    dev = "cuda"
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    bf = dict(device=dev, dtype=torch.bfloat16)
    g = torch.Generator(device=dev).manual_seed(1)
    rnd = lambda *s: torch.randn(*s, device=dev, generator=g)   # noqa: E731
    W = {n: ops.to_kmajor_bf16(ops.pack_linear_weight(0.03 * rnd(o, i))) for n, (i, o) in
         dict(qkv=(H, 3 * H), o=(H, H), ff1=(H, I), ff2=(I, H)).items()}
    Bv = {n: 0.1 * rnd(o) for n, o in dict(qkv=3 * H, o=H, ff1=I, ff2=H).items()}
    gam, bet = 1 + 0.1 * rnd(H), 0.1 * rnd(H)
    xh, xi = rnd(M, H).to(torch.bfloat16), rnd(M, I).to(torch.bfloat16)
    out = {n: torch.empty(M, o, **bf) for n, o in dict(qkv=3 * H, o=H, ff1=I, ff2=H, ln=H, ln2=H).items()}
    shapes = dict(qkv=(xh, "qkv", ops.ACT_NONE), oproj=(xh, "o", ops.ACT_NONE), ffn1g=(xh, "ff1", ops.ACT_GELU), ffn2=(xi, "ff2", ops.ACT_NONE))
    chain = ops.FlowChain(M, 8, dev)
    full = torch.full_like(chain.counters, 1 << 28)

    def preset():
        chain.counters.copy_(full)

    print("== single launches (us): plain / signal / wait (preset) / both")
    for name, (x, wn, act) in shapes.items():
        res = []
        for use_wait, use_sig in ((False, False), (False, True), (True, False), (True, True)):
            chain._next = 0
            ew = chain.edge("linear", x.shape[1])
            es = chain.edge("linear", out[wn].shape[1])

            def fn():
                ops.linear(x, W[wn], Bv[wn], out[wn], post_act=act, wait=ew if use_wait else None, signal=es if use_sig else None)
            preset()
            res.append(timed(fn, flush, a.iters))
        print(f"{name:6s} " + " ".join(f"{t:7.1f}" for t in res))
    res = []
    for use_wait, use_sig in ((False, False), (False, True), (True, False), (True, True)):
        chain._next = 0
        ew, es = chain.edge("linear", H), chain.edge("layernorm", H)

        def fn():
            ops.layernorm(out["o"], gam, bet, out["ln"], 1e-5, residual=xh, wait=ew if use_wait else None, signal=es if use_sig else None)
        preset()
        res.append(timed(fn, flush, a.iters))
    print("ln     " + " ".join(f"{t:7.1f}" for t in res))

    if a.no_chain:
        return
    print("== chain out-proj -> LN -> FFN-in -> FFN-out -> LN -> QKV (us)")
    for flow in (False, True, False, True):
        chain._next = 0
        e = [chain.edge(k, n) for k, n in (("linear", H), ("layernorm", H), ("linear", I), ("linear", H), ("layernorm", H))] if flow else [None] * 5

        def fn():
            if flow:
                chain.reset()
            ops.linear(xh, W["o"], Bv["o"], out["o"], signal=e[0])
            ops.layernorm(out["o"], gam, bet, out["ln"], 1e-5, residual=xh, wait=e[0], signal=e[1])
            ops.linear(out["ln"], W["ff1"], Bv["ff1"], out["ff1"], post_act=ops.ACT_GELU, wait=e[1], signal=e[2])
            ops.linear(out["ff1"], W["ff2"], Bv["ff2"], out["ff2"], wait=e[2], signal=e[3])
            ops.layernorm(out["ff2"], gam, bet, out["ln2"], 1e-5, residual=out["ln"], wait=e[3], signal=e[4])
            ops.linear(out["ln2"], W["qkv"], Bv["qkv"], out["qkv"], wait=e[4])
        print(f"flow={int(flow)}  {timed(fn, flush, a.iters):7.1f}")


if __name__ == "__main__":
    main()
