#!/usr/bin/env python
"""Headline benchmark: inpainted audio-seconds per second, HuBERT -> HiFi-GAN end to end.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision fp32|bf16]

Workload (BASELINE.json configs[1]): I_ea informed inpainting, HuBERT-base + head + HiFi-GAN V1 with
random-init weights, 32 x 4 s 16 kHz utterances per GPU, 200 ms mask, synthetic data (SURVEY 8d cfg 2).
A step = one pass of `InformedInpainter` over one batch: zero-mask -> z-norm -> HuBERT -> head ->
gather -> cos-sim argmax -> centroid paste -> extend_mel -> HiFi-GAN -> waveform.

  value     device-resident inputs, CUDA-event timed, max over ranks, L2 flushed between steps
  e2e       same steps through the public streaming API (InformedInpainter.stream) from pinned HOST buffers: every step
            uploads its inputs and downloads its int16 result inside ONE timed region around all K steps (copies of
            neighbouring steps overlap the compute on copy streams); the serial one-call-per-step figure is kept beside it
  roofline  dominant kernel family (the implicit-GEMM conv/linear kernel): algorithmic FLOPs / event time
  cpu_baseline  the reference's CPU path (oracle port, torch fp32, all host threads) on a bounded sample

`--impl reference` times that CPU path alone (rank 0 only) and prints the same JSON line.
One rank per GPU under torchrun; utterances shard over ranks with no data-path collective (weak scaling).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "inpainted_audio_seconds_per_second"
UNIT = "audio-s/s"
BATCH, SECONDS, SR = 32, 4, 16000
MASK_START, MASK_END = 1.5, 1.7  # 200 ms informed mask (per-utterance positions are randomised below)


def workload(batch=BATCH, seconds=SECONDS, seed=1234):
    g = torch.Generator().manual_seed(seed)
    n = seconds * SR
    wave = 0.1 * torch.randn(batch, n, generator=g)
    t_mel = n * 22050 // SR // 441  # hop-441 frames of the 22.05 kHz rendition (mel_dump.py:16)
    mel = torch.randn(batch, 80, t_mel, generator=g)
    T = (n - 400) // 320 + 1
    L = 10  # 200 ms / 20 ms
    pos = torch.randint(0, T - L, (batch,), generator=g).tolist()
    return wave, mel, pos, [L] * batch


def flops_per_utterance(n_samples, t_mel):
    """SURVEY 8d algorithmic FLOPs: HuBERT-base F_hub(N) + head + HiFi-GAN V1 614.1 MFLOP x Tm."""
    lens, L = [], n_samples
    for k, s in zip((10, 3, 3, 3, 3, 2, 2), (5, 2, 2, 2, 2, 2, 2)):
        L = (L - k) // s + 1
        lens.append(L)
    T = lens[-1]
    hub = 2 * (5120 * lens[0] + 786432 * sum(lens[1:5]) + 524288 * sum(lens[5:7]))
    hub += T * (2 * 512 * 768 + 2 * 768 * 48 * 128 + 12 * (8 * 768 ** 2 + 4 * 768 * 3072)) + 12 * 4 * T * T * 768
    hub += 2 * T * 768 * 80
    tm = int(t_mel * 441 / 256)
    return hub + 614.1e6 * tm, tm


def build_models(precision, device):
    import speech_inpainting_b200 as sib
    from oracle.params import HifiCfg, HubertCfg, make_codebook, make_generator_params, make_head_params, make_hubert_params
    ocfg, gcfg = HubertCfg.base(), HifiCfg.v1()
    sd = make_hubert_params(ocfg, 1234, prefix="base_model.")
    sd.update(make_head_params(768, 80))
    gp = make_generator_params(gcfg, 1234, "unit")
    C = make_codebook(80, 100)
    model = sib.CustomModel(80, "base", False, config=sib.HubertConfig.base(), precision=precision).to(device)
    model.load_state_dict(sd)
    gen = sib.Generator(sib.AttrDict(gcfg.as_attrdict()), precision=precision).to(device)
    gen.load_state_dict(gp)
    gen.remove_weight_norm()
    return sib, sib.InformedInpainter(model.eval(), gen.eval(), C), (sd, ocfg, gp, gcfg, C)


def cpu_reference_step(state, wave, mel, pos, ln):
    """The reference's CPU path (SURVEY 8d 'CPU baseline'): oracle modules, torch fp32, eager attention."""
    from oracle import glue_ref, hifigan_ref, hubert_ref
    sd, ocfg, gp, gcfg, C = state
    with torch.no_grad():
        x = wave.clone()
        for b in range(x.shape[0]):
            lo, hi = glue_ref.iea_zero_range_from_frames(pos[b], ln[b])
            x[b, lo:hi] = 0
        out = hubert_ref.custom_model_forward(sd, ocfg, glue_ref.processor_znorm(x))
        labels = [glue_ref.cos_sim_argmax(v, C) for v in glue_ref.gather_mask_frames(out, pos, ln)]
        feats = glue_ref.extend_mel(glue_ref.paste_centroids(mel, C, labels, pos))
        return hifigan_ref.generator_forward(gp, gcfg, feats)


def time_cpu_reference(state, steps, warmup, sample_utts):
    from oracle.params import fold_weight_norm
    sd, ocfg, gp, gcfg, C = state
    state = (sd, ocfg, fold_weight_norm(gp), gcfg, C)  # remove_weight_norm() once, as predict.py:122
    torch.set_num_threads(os.cpu_count() or 1)
    wave, mel, pos, ln = workload(batch=sample_utts)
    for _ in range(warmup):
        cpu_reference_step(state, wave, mel, pos, ln)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        cpu_reference_step(state, wave, mel, pos, ln)
        ts.append(time.perf_counter() - t0)
    total = sum(ts)
    return sample_utts * SECONDS * steps / total, total / steps


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons during the timed region (B200_PROFILING.md recipe).  NVML in-process
    (20 ms period) so that short timed regions still get tens of samples; `nvidia-smi` as the fallback."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()
        self.nvml = self.handle = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES through the PCI bus id
            bus = torch.cuda.get_device_properties(index).pci_bus_id
            self.handle = None
            for i in range(pynvml.nvmlDeviceGetCount()):
                h = pynvml.nvmlDeviceGetHandleByIndex(i)
                if pynvml.nvmlDeviceGetPciInfo(h).bus == bus:
                    self.handle = h
            if self.handle is None:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n, h = self.nvml, self.handle
        sm = n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM)
        r = n.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") \
            else n.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        flag = lambda name: "Active" if r & getattr(n, name, 0) else "Not Active"  # noqa: E731
        try:
            pw = n.nvmlDeviceGetPowerUsage(h) / 1000.0
        except Exception:
            pw = 0.0
        return [str(sm), str(mx), f"{pw:.1f}", flag("nvmlClocksThrottleReasonHwSlowdown"),
                flag("nvmlClocksThrottleReasonHwThermalSlowdown"), flag("nvmlClocksThrottleReasonSwThermalSlowdown"),
                flag("nvmlClocksThrottleReasonSwPowerCap")]

    def run(self):
        while not self._halt.is_set():
            try:
                if self.nvml is not None:
                    self.rows.append(self._sample_nvml())
                else:
                    out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                    self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            self._halt.wait(0.02 if self.nvml is not None else 0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=5)
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        pw = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_mhz_min": sm[0] if sm else None,
                "sm_max_mhz": mx[0] if mx else None, "power_w_max": max(pw) if pw else None,
                "reasons": sorted(reasons), "samples": len(self.rows),
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def profile_plans(sib, plans, detail=None):
    """Per-launch CUDA-event timing of the recorded launch lists, aggregated per C-ABI kernel family.
    Run after the timed region; gives the dominant kernel's share and its achieved FLOP rate."""
    agg = {}
    stream = torch.cuda.current_stream().cuda_stream
    prev_pdl = sib.ops.set_pdl(False)   # stream-ordered launches: consecutive kernels must not overlap under the events
    for plan in plans:
        evs = []
        for fn, args, name in plan.steps:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = fn(*args, stream)
            e1.record()
            assert rc == 0, name
            evs.append((name, args, e0, e1))
        torch.cuda.synchronize()
        for name, args, e0, e1 in evs:
            a = agg.setdefault(name, {"ms": 0.0, "launches": 0, "flops": 0.0})
            a["ms"] += e0.elapsed_time(e1)
            a["launches"] += 1
            if name in ("sib_conv1d_f32", "sib_conv1d_bf16"):
                d = args[0]._obj
                fl = 2.0 * d.batch * d.t_out * d.c_out * (d.c_in // d.groups) * d.n_taps
                a["flops"] += fl
                if detail is not None:
                    ms_ = e0.elapsed_time(e1)
                    detail.append({"kernel": name, "B": d.batch, "t_out": d.t_out, "c_in": d.c_in, "c_out": d.c_out,
                                   "groups": d.groups, "taps": d.n_taps, "stride": d.stride, "ms": round(ms_, 4),
                                   "tflops": round(fl / ms_ / 1e9, 1) if ms_ > 0 else None})
            elif name == "sib_resunit_bf16":
                d = args[0]._obj
                ms_ = e0.elapsed_time(e1)
                fl = 2.0 * 2.0 * d.batch * d.t * d.c * d.c * d.k          # two k-tap convs per unit
                n_t = 2 + int(bool(d.accumulate)) + int(args[7] is not None)  # x in, y out (+ running sum in, + lrelu(y) out)
                by = 2.0 * d.batch * d.t * d.c * n_t
                a["flops"] += fl
                a["bytes"] = a.get("bytes", 0.0) + by
                if detail is not None:
                    detail.append({"kernel": name, "B": d.batch, "t": d.t, "c": d.c, "k": d.k, "dilation": d.dilation,
                                   "tensors": n_t, "ms": round(ms_, 4), "tflops": round(fl / ms_ / 1e9, 1),
                                   "gbs": round(by / ms_ / 1e6, 1)})
            elif detail is not None:
                detail.append({"kernel": name, "ms": round(e0.elapsed_time(e1), 4)})
    sib.ops.set_pdl(prev_pdl)
    return agg


def ncu_traffic(kernel_key):
    """Per-launch DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of a kernel family from the newest
    committed ncu capture of this same command (profiles/r*_ncu_step_*_summary.json); None if there is none."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_step_*_summary.json")))
    if not files:
        return None, None
    try:
        doc = json.load(open(files[-1]))
        for k in doc["kernels"]:
            if kernel_key in k["kernel"]:
                per_launch = (k["dram_read_bytes_per_step"] + k["dram_write_bytes_per_step"]) / k["launches_per_step"]
                return per_launch, os.path.basename(files[-1])
    except Exception:
        pass
    return None, None


_REAL_STDOUT = None


def _claim_stdout():
    """stdout must carry exactly ONE JSON line.  Libraries below us write to fd 1 on their own (NCCL prints its version
    banner there from C), so fd 1 is pointed at stderr for the whole run and the JSON line goes to the saved descriptor."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def _emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, line)


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("SIB_PRECISION", "bf16"), choices=["fp32", "bf16"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--cpu-sample-utts", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--breakdown", default=None, help="write the per-launch timing table (JSON) to this path")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_samples = SECONDS * SR
    fl_utt, tm = flops_per_utterance(n_samples, n_samples * 22050 // SR // 441)
    config = {"workload": "I_ea informed inpainting, HuBERT-base + head + HiFi-GAN V1 (random init), "
                          f"{args.batch}x{SECONDS} s 16 kHz utterances per GPU, 200 ms mask (BASELINE configs[1])",
              "batch_per_gpu": args.batch, "utterance_seconds": SECONDS, "mask_ms": 200, "parallelism": f"dp{world}",
              "l2": "256 MiB memset between timed steps (L2 flush); per-step working set >> 126 MB L2"}

    if args.impl == "reference":
        if rank != 0:
            return
        from oracle.params import HifiCfg, HubertCfg, make_codebook, make_generator_params, make_head_params, make_hubert_params
        ocfg, gcfg = HubertCfg.base(), HifiCfg.v1()
        sd = make_hubert_params(ocfg, 1234, prefix="base_model.")
        sd.update(make_head_params(768, 80))
        state = (sd, ocfg, make_generator_params(gcfg, 1234, "unit"), gcfg, make_codebook(80, 100))
        v, sec = time_cpu_reference(state, args.steps, args.warmup, args.cpu_sample_utts)
        sample = (f"{args.cpu_sample_utts}x{SECONDS} s utterances per step of the same workload "
                  f"(1/{args.batch // args.cpu_sample_utts} of the batch), {args.steps} steps after {args.warmup} warm-up")
        _emit(({"metric": METRIC, "value": v, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                           "sample": sample},
                          "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "gpu_launches": 0}))
        return

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback for the product path)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    sib, pipe, state = build_models(args.precision, dev)
    # rank r owns utterances [r*B, (r+1)*B) of the global batch (weak scaling; no data-path collective)
    wave, mel, pos, ln = workload(batch=args.batch, seed=1234 + rank)
    wave_d, mel_d = wave.to(dev), mel.to(dev)
    wave_h, mel_h = wave.pin_memory(), mel.pin_memory()
    out_h = torch.empty(args.batch, 1, tm * 256, dtype=torch.int16).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step_device():
        return pipe(wave_d, mel_d, pos, ln)

    def step_e2e():
        res = pipe(wave_h, mel_h, pos, ln, return_int16=True)  # H2D copies of wave/mel happen inside
        out_h.copy_(res.int16, non_blocking=True)
        return res

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        evs = []
        barrier()
        for _ in range(steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            evs.append((e0, e1))
        barrier()
        ms = sum(a.elapsed_time(b) for a, b in evs)
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for _ in range(max(args.warmup, 3)):
        step_device()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    n0 = sib.ops.launch_count()
    torch.cuda.nvtx.range_push("sib_timed")  # ncu --nvtx --nvtx-include "sib_timed/" profiles exactly these launches
    ms = timed(step_device, args.steps)
    torch.cuda.nvtx.range_pop()
    launches = sib.ops.launch_count() - n0
    for _ in range(2):
        step_e2e()
    ms_e2e_serial = timed(step_e2e, args.steps)

    # e2e through the public streaming API (InformedInpainter.stream): every step uploads its own inputs from pinned host
    # memory and downloads its int16 result; uploads / downloads of neighbouring steps overlap the compute on a copy
    # stream.  Timed as ONE region around all K steps (L2-flush memsets included), device events, max over ranks.
    def host_batches(n):
        for _ in range(n):
            flush.zero_()
            yield {"wave16": wave_h, "mel": mel_h, "mask_pos": pos, "mask_len": ln}

    for _ in pipe.stream(host_batches(2)):
        pass
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n_done = sum(1 for _ in pipe.stream(host_batches(args.steps)))   # each yield = that step's result is on the host
    e1.record()
    barrier()
    assert n_done == args.steps
    ms_e2e = e0.elapsed_time(e1)
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())
    clocks = sampler.stop()

    audio_s = args.batch * SECONDS * world
    value = audio_s * args.steps / (ms / 1e3)
    e2e = audio_s * args.steps / (ms_e2e / 1e3)

    roofline = cpu = None
    if rank == 0:
        plans = [io.plan for io in pipe.model.base_model._plans.values()] + [io.plan for io in pipe.generator._plans.values()]
        detail = [] if args.breakdown else None
        agg = profile_plans(sib, plans, detail)
        if args.breakdown:
            os.makedirs(os.path.dirname(os.path.abspath(args.breakdown)), exist_ok=True)
            with open(args.breakdown, "w") as f:
                json.dump(detail, f, indent=0)
        total_ms = sum(a["ms"] for a in agg.values())
        name, top = max(agg.items(), key=lambda kv: kv[1]["ms"])
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = peaks.get("bf16_tflops_sustained", 1400.0) if args.precision == "bf16" else None
        ach = top["flops"] / (top["ms"] / 1e3) / 1e12 if top["ms"] else 0.0
        # fp32 SIMT arm: no tensor pipe; the bound quoted is still the tensor roofline the bf16 arm is judged by
        peak = peaks.get("bf16_tflops_sustained", 1400.0)
        ncu_key = {"sib_conv1d_bf16": "conv1d_bf16_tc_kernel", "sib_resunit_bf16": "resunit_tc_kernel"}.get(name, name)
        traffic, traffic_src = ncu_traffic(ncu_key)
        roofline = {"bound": "tensor", "kernel": name, "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                    "frac": ach / peak, "traffic": traffic,
                    "traffic_note": (f"DRAM bytes per launch, averaged over the family's launches of one step ({traffic_src})"
                                     if traffic else "no ncu capture committed"),
                    "algorithmic_flops_per_launch": top["flops"] / max(top["launches"], 1),
                    "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)"
                    if peaks else "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)",
                    "kernel_share_of_step": top["ms"] / total_ms if total_ms else None,
                    "launches_per_step": top["launches"], "algorithmic_flops_per_step": top["flops"],
                    "per_kernel_ms": {k: round(v["ms"], 3) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])}}
        ru = agg.get("sib_resunit_bf16")
        if ru and ru.get("bytes") and ru["ms"] > 0:
            # second family: the fused ResBlock units of the narrow stages are HBM-bound (x in, y out)
            hbm_peak = peaks.get("hbm_gbs", 6650.0)
            gbs = ru["bytes"] / (ru["ms"] / 1e3) / 1e9
            t2, _ = ncu_traffic("resunit_tc_kernel")
            roofline["secondary"] = [{"bound": "hbm", "kernel": "sib_resunit_bf16", "achieved": gbs, "peak": hbm_peak,
                                      "unit": "GB/s", "frac": gbs / hbm_peak, "traffic": t2,
                                      "algorithmic_bytes_per_launch": ru["bytes"] / ru["launches"],
                                      "tflops": ru["flops"] / (ru["ms"] / 1e3) / 1e12,
                                      "kernel_share_of_step": ru["ms"] / total_ms}]
        if not args.no_cpu_baseline and world == 1:   # the CPU baseline is an N = 1 figure: other ranks would be spinning beside it
            v, sec = time_cpu_reference(state, 3, 1, args.cpu_sample_utts)
            cpu = {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                   "sample": f"{args.cpu_sample_utts}x{SECONDS} s utterances of the same workload, 3 steps after 1 warm-up, "
                             f"{sec:.2f} s per step; oracle port of the reference's PyTorch CPU path"}
        _emit(({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.precision == "fp32" else "bf16", "data": "synthetic", "config": config,
            "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "api": "InformedInpainter.stream (2-deep upload / compute / download pipeline; one timed region around "
                           "all steps, L2-flush memsets included)",
                    "serial_ms_per_step": ms_e2e_serial / args.steps,
                    "h2d_bytes_per_step": (wave_h.numel() + mel_h.numel()) * 4 * world,
                    "d2h_bytes_per_step": out_h.numel() * 2 * world},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "algorithmic_tflop_per_step": fl_utt * args.batch * world / 1e12,
        }))   # written straight to the descriptor before the teardown collectives: a rank dying there cannot take it along
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
