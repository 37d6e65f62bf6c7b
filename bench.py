#!/usr/bin/env python
"""Headline benchmark: inpainted audio-seconds per second, HuBERT -> HiFi-GAN end to end.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision fp32|bf16]
                    [--workload cfg2|cfg3|cfg4|cfg5]

Default workload cfg2 (BASELINE.json configs[1], the headline): I_ea informed inpainting, HuBERT-base + head + HiFi-GAN V1
with random-init weights, 32 x 4 s 16 kHz utterances per GPU, 200 ms mask, synthetic data (SURVEY 8d cfg 2), weak scaling.
A step = one pass of `InformedInpainter` over the rank's utterances: zero-mask -> z-norm -> HuBERT -> head ->
gather -> cos-sim argmax -> centroid paste -> extend_mel -> HiFi-GAN -> waveform -> int16.
--workload cfg5 = BASELINE configs[4]: ONE fixed job of 1024 x 10 s utterances sharded contiguously over the ranks
(`shard_batch`), micro-batches of 32 through the streaming API, "scaling": "strong" (time-to-finish shrinks with N);
cfg4 = HuBERT-large 128 x 6 s, ragged masks; cfg3 = I_da blind inpainting 64 x 4 s (BlindInpainter).

  value     device-resident inputs, CUDA-event timed, max over ranks, L2 flushed between steps
  e2e       same steps through the public streaming API (InformedInpainter.stream) from pinned HOST buffers: every step
            uploads its inputs and downloads its int16 result inside ONE timed region around all K steps (copies of
            neighbouring steps overlap the compute on copy streams); the serial one-call-per-step figure is kept beside it
  roofline  dominant kernel family (the implicit-GEMM conv/linear kernel): algorithmic FLOPs / event time
  cpu_baseline  the reference's CPU path (oracle port, torch fp32, all host threads) on a bounded sample

`--impl reference` times that CPU path alone (rank 0 only) and prints the same JSON line.
One rank per GPU under torchrun; utterances shard over ranks with no data-path collective.
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
SR = 16000

# BASELINE.json configs[1..4].  `batch` = utterances per GPU (weak) or per JOB (strong); every step runs
# them in micro-batches of `micro`; `cpu_utts` = utterances per step of the bounded CPU sample.
WORKLOADS = {
    "cfg2": dict(flavour="iea", size="base", batch=32, seconds=4, lens=[10], K=100, micro=32, scaling="weak", cpu_utts=8,
                 name="I_ea informed inpainting, HuBERT-base + head + HiFi-GAN V1 (random init), {b}x4 s 16 kHz utterances per GPU, "
                      "200 ms mask (BASELINE configs[1])"),
    "cfg3": dict(flavour="ida", size="base", batch=64, seconds=4, mask=6400, K=500, micro=64, scaling="weak", cpu_utts=2,
                 name="I_da blind inpainting, HuBERT-base (z-norm off) + k-means K=500 + hubert_lut CodeGenerator (random init), "
                      "{b}x4 s per GPU, 400 ms gap at 1.5 s; per utterance 2 HuBERT passes + 2 CodeGenerator passes as "
                      "scripts/inpainting.py:195-259, audio-s counted once (BASELINE configs[2])"),
    "cfg4": dict(flavour="iea", size="large", batch=128, seconds=6, lens=[1, 2, 3, 4, 5, 10, 15, 20], K=500, micro=32,
                 scaling="weak", cpu_utts=2,
                 name="I_ea informed inpainting, HuBERT-large + head + HiFi-GAN V1 (random init), {b}x6 s per GPU in micro-batches "
                      "of 32, mask lengths 1..20 frames, K=500 (BASELINE configs[3])"),
    "cfg5": dict(flavour="iea", size="base", batch=1024, seconds=10, lens=[10], K=100, micro=32, scaling="strong", cpu_utts=2,
                 name="I_ea batch-sharded sweep: ONE fixed job of {b}x10 s utterances (10 240 audio-s) split contiguously over the "
                      "ranks, micro-batches of 32, 200 ms mask (BASELINE configs[4])"),
}


def workload(batch, seconds, seed=1234, lens=(10,)):
    g = torch.Generator().manual_seed(seed)
    n = seconds * SR
    wave = 0.1 * torch.randn(batch, n, generator=g)
    t_mel = n * 22050 // SR // 441  # hop-441 frames of the 22.05 kHz rendition (mel_dump.py:16)
    mel = torch.randn(batch, 80, t_mel, generator=g)
    T = (n - 400) // 320 + 1
    ln = [lens[i % len(lens)] for i in range(batch)]
    pos = [int(torch.randint(0, T - l, (1,), generator=g)) for l in ln]
    return wave, mel, pos, ln


def hubert_flops(cfg, n_samples, head_dim=0):
    """SURVEY 8d F_hub(N): feature encoder + projection + pos-conv + layers (+ head)."""
    lens, L = [], n_samples
    for k, s in zip(cfg.conv_kernel, cfg.conv_stride):
        L = (L - k) // s + 1
        lens.append(L)
    T, H, F = lens[-1], cfg.hidden_size, cfg.intermediate_size
    fl, cin = 0.0, 1
    for t, c, k in zip(lens, cfg.conv_dim, cfg.conv_kernel):
        fl += 2.0 * t * c * cin * k
        cin = c
    fl += 2.0 * T * cin * H + 2.0 * T * H * (H // cfg.num_conv_pos_embedding_groups) * cfg.num_conv_pos_embeddings
    fl += cfg.num_hidden_layers * (T * (8.0 * H * H + 4.0 * H * F) + 4.0 * T * T * H)
    return fl + 2.0 * T * H * head_dim, T


def generator_flops(gcfg, t_in):
    """2 x MACs of every conv of `Generator.forward` (614.1 MFLOP per mel frame for V1, 322.6 per code frame for I_da)."""
    c0 = gcfg.upsample_initial_channel
    fl, L, ch = 2.0 * t_in * gcfg.model_in_dim * c0 * 7, t_in, c0
    for i, (u, k) in enumerate(zip(gcfg.upsample_rates, gcfg.upsample_kernel_sizes)):
        cin, ch = c0 // (2 ** i), c0 // (2 ** (i + 1))
        L *= u
        fl += 2.0 * L * cin * ch * k / u
        per = 2 if gcfg.resblock == "1" else 1
        for rk, dil in zip(gcfg.resblock_kernel_sizes, gcfg.resblock_dilation_sizes):
            fl += per * len(dil) * 2.0 * L * ch * ch * rk
    return fl + 2.0 * L * ch * 7


def make_state(wl):
    from oracle.params import HifiCfg, HubertCfg, make_codebook, make_generator_params, make_head_params, make_hubert_params
    ocfg = HubertCfg.base() if wl["size"] == "base" else HubertCfg.large()
    if wl["flavour"] == "iea":
        gcfg = HifiCfg.v1()
        sd = make_hubert_params(ocfg, 1234, prefix="base_model.")
        sd.update(make_head_params(ocfg.hidden_size, 80))
        return dict(sd=sd, ocfg=ocfg, gp=make_generator_params(gcfg, 1234, "unit"), gcfg=gcfg, C=make_codebook(80, wl["K"]))
    gcfg = HifiCfg.ida()
    g = torch.Generator().manual_seed(52)
    mu = torch.randn(gcfg.num_embeddings, ocfg.hidden_size, generator=g) * 0.5
    return dict(hp=make_hubert_params(ocfg, 52), ocfg=ocfg, gp=make_generator_params(gcfg, 52, "unit"), gcfg=gcfg, mu=mu)


def build_pipeline(wl, st, precision, device):
    import speech_inpainting_b200 as sib
    cfg = sib.HubertConfig.base() if wl["size"] == "base" else sib.HubertConfig.large()
    if wl["flavour"] == "iea":
        model = sib.CustomModel(80, wl["size"], False, config=cfg, precision=precision).to(device)
        model.load_state_dict(st["sd"])
        gen = sib.Generator(sib.AttrDict(st["gcfg"].as_attrdict()), precision=precision).to(device)
        gen.load_state_dict(st["gp"])
        gen.remove_weight_norm()
        return sib, sib.InformedInpainter(model.eval(), gen.eval(), st["C"]), [model.base_model, gen]
    hub = sib.HubertModel(cfg, precision=precision).to(device)
    hub.load_state_dict(st["hp"])
    gen = sib.CodeGenerator(sib.AttrDict(st["gcfg"].as_attrdict()), precision=precision).to(device)
    gen.load_state_dict(st["gp"])
    return sib, sib.BlindInpainter(hub.eval(), gen.eval(), st["mu"], layer=-1, normalize=False), [hub, gen]


def ida_inputs(batch, seconds, gcfg, T, seed=52):
    g = torch.Generator().manual_seed(seed)
    wave = 0.1 * torch.randn(batch, seconds * SR, generator=g)
    zp = torch.randint(0, 20, (batch, T // 4 + 1), generator=g)
    emb = torch.randn(batch, gcfg.embedding_dim, generator=g)
    return wave, zp, emb


def time_cpu_reference(wl, st, steps, warmup, sample_utts):
    """The reference's CPU path on a bounded sample of the workload, all host threads (oracle/ref_modules.py: the real
    transformers.HubertModel; the reference's own Generator when /root/reference exists, else its pinned port)."""
    from oracle.ref_modules import ReferenceIda, ReferenceIea
    torch.set_num_threads(os.cpu_count() or 1)
    if wl["flavour"] == "iea":
        ref = ReferenceIea(st["sd"], st["ocfg"], st["gp"], st["gcfg"], st["C"])
        wave, mel, pos, ln = workload(sample_utts, wl["seconds"], lens=wl["lens"])
        step = lambda: ref(wave, mel, pos, ln)   # noqa: E731
    else:
        ref = ReferenceIda(st["hp"], st["ocfg"], st["gp"], st["gcfg"], st["mu"])
        wave, zp, emb = ida_inputs(sample_utts, wl["seconds"], st["gcfg"], st["ocfg"].feat_lengths(wl["seconds"] * SR))
        step = lambda: ref(wave, wl["mask"], zp, emb)   # noqa: E731
    for _ in range(warmup):
        step()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    total = sum(ts)
    return sample_utts * wl["seconds"] * steps / total, total / steps, ref


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
            if name == "sib_linear_ln_bf16":     # same kernel family (conv1d_bf16_tc_kernel), LayerNorm-folding epilogues
                name = "sib_conv1d_bf16"
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


def ncu_traffic(kernel_key, workload_key):
    """Per-launch DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of a kernel family from the newest committed
    ncu capture of this same command and workload (profiles/r*_ncu_step_*_summary.json); (None, None) if there is none.
    The summary carries the git head of the tree it was captured on, which goes into `roofline.traffic_note`."""
    import glob
    import re

    def order(f):   # (round, capture version) from r02_ncu_step_cfg2_v4_summary.json; a checkout gives every file the same mtime
        m = re.search(r"r(\d+)_ncu_step_.*?v(\d+)_summary", os.path.basename(f))
        return (int(m.group(1)), int(m.group(2))) if m else (0, 0)
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_step_*_summary.json")), key=lambda f: (order(f), os.path.getmtime(f)))
    for f in reversed(files):
        try:
            doc = json.load(open(f))
            if doc.get("workload", "cfg2") != workload_key:
                continue
            for k in doc["kernels"]:
                if kernel_key in k["kernel"]:
                    per_launch = (k["dram_read_bytes_per_step"] + k["dram_write_bytes_per_step"]) / k["launches_per_step"]
                    return per_launch, (f"{os.path.basename(f)}, captured on git {doc.get('git_head', 'unknown (round 1)')}"
                                        f" / source digest {doc.get('source_digest', 'n/a')}")
        except Exception:
            continue
    return None, None


def git_head():
    """Git head when there is a work tree (the build container); the GPU boxes receive a snapshot without .git, so the line
    always carries `source_digest()` as well."""
    try:
        return subprocess.run(["git", "-C", ROOT, "rev-parse", "--short=12", "HEAD"], capture_output=True, text=True,
                              timeout=10).stdout.strip() or None
    except Exception:
        return None


def source_digest():
    """sha256 over the product sources (csrc, package .py, header, bench.py): identifies the tree a number or an ncu capture
    was taken on, with or without git.  `python bench.py --print-digest` prints it."""
    import glob
    import hashlib
    h = hashlib.sha256()
    pkg = os.path.join(ROOT, "speech-inpainting_b200")
    files = sorted(glob.glob(os.path.join(pkg, "csrc", "*.cu*")) + glob.glob(os.path.join(pkg, "*.py")) +
                   glob.glob(os.path.join(ROOT, "include", "*.h")) + [os.path.join(ROOT, "bench.py")])
    for f in files:
        h.update(os.path.relpath(f, ROOT).encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:12]


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
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("SIB_PRECISION", "bf16"), choices=["fp32", "bf16"])
    ap.add_argument("--workload", default=os.environ.get("SIB_WORKLOAD", "cfg2"), choices=sorted(WORKLOADS),
                    help="cfg2 = the headline (BASELINE configs[1]); cfg5 = the fixed 1024 x 10 s job, STRONG scaling")
    ap.add_argument("--batch", type=int, default=None, help="utterances per GPU (weak) / per job (strong); default per workload")
    ap.add_argument("--cpu-sample-utts", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--breakdown", default=None, help="write the per-launch timing table (JSON) to this path")
    ap.add_argument("--print-digest", action="store_true", help="print the source digest of this tree and exit")
    args = ap.parse_args()
    if args.print_digest:
        _emit({"source_digest": source_digest(), "git_head": git_head()})
        return

    wl = dict(WORKLOADS[args.workload])
    if args.batch:
        wl["batch"] = args.batch
    if args.steps is None:
        args.steps = 20 if args.workload in ("cfg2", "cfg3") else (6 if args.workload == "cfg4" else 3)
    cpu_utts = args.cpu_sample_utts or wl["cpu_utts"]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    seconds, n_samples = wl["seconds"], wl["seconds"] * SR
    config = {"workload": wl["name"].format(b=wl["batch"]), "workload_key": args.workload,
              "batch_per_gpu" if wl["scaling"] == "weak" else "batch_per_job": wl["batch"], "micro_batch": wl["micro"],
              "utterance_seconds": seconds, "parallelism": f"dp{world}",
              "l2": "256 MiB memset between timed steps (L2 flush); per-step working set >> 126 MB L2"}

    if args.impl == "reference":
        if rank != 0:
            return
        st = make_state(wl)
        v, sec, ref = time_cpu_reference(wl, st, args.steps, args.warmup, cpu_utts)
        sample = (f"{cpu_utts}x{seconds} s utterances per step of the same workload, {args.steps} steps after {args.warmup} "
                  f"warm-up, {sec:.2f} s per step; " + "; ".join(f"{k}: {v_}" for k, v_ in ref.parts.items()))
        _emit(({"metric": METRIC, "value": v, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
                "scaling": wl["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": ref.kind, "sample": sample},
                "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}))
        return

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback for the product path)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    st = make_state(wl)
    sib, pipe, modules = build_pipeline(wl, st, args.precision, dev)

    # ---- this rank's share.  weak: `batch` utterances per GPU; strong: a contiguous shard of the ONE job
    # (I_da/scripts/inference.py:244-248,320: a fixed file list split over the workers) - no data-path collective either way
    if wl["scaling"] == "strong":
        lo, hi = sib.shard_batch(wl["batch"], world, rank)
        my_utts, total_utts = hi - lo, wl["batch"]
    else:
        my_utts, total_utts = wl["batch"], wl["batch"] * world
    micro = [min(wl["micro"], my_utts - o) for o in range(0, my_utts, wl["micro"])]     # micro-batch sizes of one step
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    if wl["flavour"] == "iea":
        # one seeded micro-batch per distinct size (synthetic data: every micro-batch of a step re-uses it)
        data = {}
        for mb in sorted(set(micro)):
            wave, mel, pos, ln = workload(mb, seconds, seed=1234 + rank, lens=wl["lens"])
            data[mb] = dict(wave_d=wave.to(dev), mel_d=mel.to(dev), pos=pos, ln=ln,
                            host={"wave16": wave.pin_memory(), "mel": mel.pin_memory(), "mask_pos": pos, "mask_len": ln})
        tm = sib.ops.extend_mel_len(n_samples * 22050 // SR // 441)
        out_samples = tm * 256
        h2d = sum((data[mb]["host"]["wave16"].numel() + data[mb]["host"]["mel"].numel()) * 4 for mb in micro)
        d2h = sum(mb * out_samples * 2 for mb in micro)
        fl_h, _ = hubert_flops(st["ocfg"], n_samples, 0)
        fl_utt = fl_h + generator_flops(st["gcfg"], tm)

        def step_device():
            for mb in micro:
                d = data[mb]
                pipe(d["wave_d"], d["mel_d"], d["pos"], d["ln"], return_int16=True)

        def host_batches(steps):
            for _ in range(steps):
                flush.zero_()
                for mb in micro:
                    yield data[mb]["host"]

        def run_e2e(steps):
            return sum(1 for _ in pipe.stream(host_batches(steps))) == steps * len(micro)
        e2e_api = ("InformedInpainter.stream (2-deep upload / compute / download pipeline over pinned host micro-batches; one "
                   "timed region around all steps, L2-flush memsets included)")
    else:
        T = st["ocfg"].feat_lengths(n_samples)
        data = {}
        for mb in sorted(set(micro)):
            wave, zp, emb = ida_inputs(mb, seconds, st["gcfg"], T, seed=52 + rank)
            data[mb] = dict(wave_d=wave.to(dev), zp_d=zp.to(dev), emb_d=emb.to(dev), wave_h=wave.pin_memory(), zp_h=zp.pin_memory(),
                            emb_h=emb.pin_memory(), out_h=torch.empty(mb, 1, (T // 4 * 4) * 320, dtype=torch.int16).pin_memory())
        n_code = sib.ida_matched_frames(n_samples, T, 4 * T)
        h2d = sum(data[mb]["wave_h"].numel() * 4 + data[mb]["zp_h"].numel() * 8 + data[mb]["emb_h"].numel() * 4 for mb in micro)
        d2h = sum(mb * n_code * 320 * 2 for mb in micro)
        fl_h, _ = hubert_flops(st["ocfg"], n_samples, 0)
        fl_utt = 2 * fl_h + 2 * generator_flops(st["gcfg"], n_code) + 2 * 2.0 * T * st["gcfg"].num_embeddings * st["ocfg"].hidden_size

        def step_device():
            for mb in micro:
                d = data[mb]
                pipe(d["wave_d"], wl["mask"], d["zp_d"], d["emb_d"], informed=False, return_int16=True)

        def run_e2e(steps):
            for _ in range(steps):
                flush.zero_()
                for mb in micro:
                    d = data[mb]
                    res = pipe(d["wave_h"].to(dev, non_blocking=True), wl["mask"], d["zp_h"].to(dev, non_blocking=True),
                               d["emb_h"].to(dev, non_blocking=True), informed=False, return_int16=True)
                    d["out_h"][:, :, : res.int16.shape[-1]].copy_(res.int16, non_blocking=True)
            return True
        e2e_api = "BlindInpainter.__call__ on pinned host tensors + int16 download per micro-batch (serial)"

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

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
        return max_over_ranks(sum(a.elapsed_time(b) for a, b in evs))

    warm = max(args.warmup, 3)
    for _ in range(warm):
        step_device()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    n0 = sib.ops.launch_count()
    torch.cuda.nvtx.range_push("sib_timed")  # ncu --nvtx --nvtx-include "sib_timed/" profiles exactly these launches
    ms = timed(step_device, args.steps)
    torch.cuda.nvtx.range_pop()
    launches = sib.ops.launch_count() - n0

    # e2e through the public API from pinned HOST buffers: every micro-batch uploads its inputs and downloads its int16
    # result inside ONE timed region around all K steps (device events, max over ranks)
    # warm-up of the streaming path: the 2-deep pipeline owns two device slots and three pinned result buffers, which are
    # allocated on first use (cudaHostAlloc can take tens of milliseconds on a busy host) - one step would leave two of the
    # three to be allocated inside the timed region
    assert run_e2e(max(warm, 4))
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    assert run_e2e(args.steps)
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop()

    audio_s = total_utts * seconds
    value = audio_s * args.steps / (ms / 1e3)
    e2e = audio_s * args.steps / (ms_e2e / 1e3)
    utts_all = [my_utts]
    if world > 1:
        import torch.distributed as dist
        t = torch.zeros(world, device=dev)
        t[rank] = my_utts
        dist.all_reduce(t)
        utts_all = [int(v) for v in t.tolist()]

    roofline = cpu = None
    if rank == 0:
        plans = [io.plan for m in modules for io in m._plans.values()]
        if getattr(pipe, "gen", None) is not None and getattr(pipe.gen, "fo_vqvae", None) is not None:
            plans += [io.plan for io in pipe.gen.fo_vqvae._plans.values()]
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
        ach = top["flops"] / (top["ms"] / 1e3) / 1e12 if top["ms"] else 0.0
        # (fp32 SIMT arm: no tensor pipe; the bound quoted is still the tensor roofline the bf16 arm is judged by)
        peak = peaks.get("bf16_tflops_sustained", 1400.0)
        ncu_key = {"sib_conv1d_bf16": "conv1d_bf16_tc_kernel", "sib_resunit_bf16": "resunit_tc_kernel"}.get(name, name)
        traffic, traffic_src = ncu_traffic(ncu_key, args.workload)
        roofline = {"bound": "tensor", "kernel": name, "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                    "frac": ach / peak, "traffic": traffic,
                    "traffic_note": (f"DRAM bytes per launch, averaged over the family's launches of one step ({traffic_src}); "
                                     f"this run: git {git_head()} / source digest {source_digest()}" if traffic
                                     else "no ncu capture of this workload committed"),
                    "algorithmic_flops_per_launch": top["flops"] / max(top["launches"], 1),
                    "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)"
                    if peaks else "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)",
                    "kernel_share_of_step": top["ms"] / total_ms if total_ms else None,
                    "launches_per_step": top["launches"], "algorithmic_flops_per_step": top["flops"],
                    "whole_step": {"algorithmic_tflop": fl_utt * my_utts / 1e12,
                                   "achieved_tflops": fl_utt * my_utts / 1e12 / (ms / args.steps / 1e3),
                                   "frac_of_peak": fl_utt * my_utts / 1e12 / (ms / args.steps / 1e3) / peak},
                    "per_kernel_ms": {k: round(v["ms"], 3) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])},
                    "note": "per-kernel times: one instrumented replay of every recorded plan (PDL off, CUDA events per launch)"}
        ru = agg.get("sib_resunit_bf16")
        if ru and ru.get("bytes") and ru["ms"] > 0 and name != "sib_resunit_bf16":
            # second family: the fused ResBlock units of the narrow stages (x in, y out)
            hbm_peak = peaks.get("hbm_gbs", 6650.0)
            gbs = ru["bytes"] / (ru["ms"] / 1e3) / 1e9
            t2, _ = ncu_traffic("resunit_tc_kernel", args.workload)
            roofline["secondary"] = [{"bound": "hbm", "kernel": "sib_resunit_bf16", "achieved": gbs, "peak": hbm_peak,
                                      "unit": "GB/s", "frac": gbs / hbm_peak, "traffic": t2,
                                      "algorithmic_bytes_per_launch": ru["bytes"] / ru["launches"],
                                      "tflops": ru["flops"] / (ru["ms"] / 1e3) / 1e12,
                                      "kernel_share_of_step": ru["ms"] / total_ms}]
        if not args.no_cpu_baseline and world == 1:   # the CPU baseline is an N = 1 figure: other ranks would be spinning beside it
            v, sec, ref = time_cpu_reference(wl, st, 3, 1, cpu_utts)
            cpu = {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": ref.kind,
                   "sample": f"{cpu_utts}x{seconds} s utterances of the same workload, 3 steps after 1 warm-up, {sec:.2f} s per step; "
                             + "; ".join(f"{k}: {v_}" for k, v_ in ref.parts.items())}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None,
            "dtype": "f32" if args.precision == "fp32" else "bf16", "data": "synthetic", "config": config,
            "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e / args.steps, "api": e2e_api,
                    "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world},   # weak: every rank moves its own batch
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "algorithmic_tflop_per_step": fl_utt * total_utts / 1e12,
            "utterances_per_rank": utts_all, "micro_batches_per_rank_per_step": len(micro), "git_head": git_head(),
            "source_digest": source_digest(),
        }
        if wl["scaling"] == "strong":   # the whole job's bytes: every utterance goes up and comes down exactly once
            per_utt_in = h2d / max(my_utts, 1)
            line["e2e"]["h2d_bytes_per_step"] = int(per_utt_in * total_utts)
            line["e2e"]["d2h_bytes_per_step"] = int(d2h / max(my_utts, 1) * total_utts)
            line["time_to_finish_job_ms"] = ms / args.steps
            line["e2e"]["time_to_finish_job_ms"] = ms_e2e / args.steps
        _emit(line)   # written straight to the descriptor before the teardown collectives: a rank dying there cannot take it along
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
