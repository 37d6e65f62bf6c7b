"""-m gpu: the "honest GPU baseline" of SURVEY 8d - the reference's PyTorch path (oracle port: cuDNN convs, cuBLAS GEMMs,
eager attention) on the SAME B200, fp32 and bf16 autocast, against the product pipeline on the headline workload
(BASELINE configs[1]: 32 x 4 s, HuBERT-base + head + HiFi-GAN V1).  The reference has no Blackwell kernels of its own, so
this is the number a user gets today by moving the reference to the GPU unchanged.  Printed, and asserted to be slower."""
import os
import sys
import time

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


def test_product_beats_pytorch_on_the_same_gpu(capsys):
    import bench
    from oracle.params import fold_weight_norm
    dev = torch.device("cuda", 0)
    wl = bench.WORKLOADS["cfg2"]
    st = bench.make_state(wl)
    sib, pipe, _ = bench.build_pipeline(wl, st, "bf16", dev)
    sd, ocfg, gp, gcfg, C = st["sd"], st["ocfg"], st["gp"], st["gcfg"], st["C"]
    wave, mel, pos, ln = bench.workload(wl["batch"], wl["seconds"])
    audio_s = wl["batch"] * wl["seconds"]

    def timed(fn, iters=3):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(iters):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / iters

    wave_d, mel_d = wave.to(dev), mel.to(dev)
    ours = timed(lambda: pipe(wave_d, mel_d, pos, ln), iters=10)
    state = ({k: v.to(dev) for k, v in sd.items()}, ocfg, {k: v.to(dev) for k, v in fold_weight_norm(gp).items()}, gcfg, C.to(dev))
    torch.backends.cudnn.benchmark = True
    res = {}

    def torch_step():
        """The oracle port's torch modules moved to the GPU unchanged (cuDNN convs, cuBLAS GEMMs, eager attention)."""
        from oracle import glue_ref, hifigan_ref, hubert_ref
        sd_, ocfg_, gp_, gcfg_, C_ = state
        with torch.no_grad():
            x = wave_d.clone()
            for b in range(x.shape[0]):
                lo, hi = glue_ref.iea_zero_range_from_frames(pos[b], ln[b])
                x[b, lo:hi] = 0
            out = hubert_ref.custom_model_forward(sd_, ocfg_, glue_ref.processor_znorm(x))
            labels = [glue_ref.cos_sim_argmax(v, C_) for v in glue_ref.gather_mask_frames(out, pos, ln)]
            feats = glue_ref.extend_mel(glue_ref.paste_centroids(mel_d, C_, labels, pos))
            return hifigan_ref.generator_forward(gp_, gcfg_, feats)

    try:
        res["fp32"] = timed(torch_step)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            res["bf16 autocast"] = timed(torch_step)
    except RuntimeError as e:   # the oracle port is written for the CPU; a device mismatch inside it is not a product failure
        pytest.skip(f"oracle port does not run on the GPU as is: {e}")
    with capsys.disabled():
        print(f"\n[torch-on-B200 baseline, 32x4 s] product {audio_s / ours:9.0f} audio-s/s ({ours * 1e3:.2f} ms/step)")
        for k, v in res.items():
            print(f"[torch-on-B200 baseline, 32x4 s] PyTorch {k:14s} {audio_s / v:9.0f} audio-s/s ({v * 1e3:.1f} ms/step) "
                  f"-> product is {v / ours:.1f}x faster")
    assert ours < min(res.values())
