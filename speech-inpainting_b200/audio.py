"""Wave-file batcher and 16 kHz <-> 22.05 kHz resampling on the device.

The reference reads every utterance twice with `librosa.load(path, sr=22050)` / `sr=16000` (I_ea/predict.py:79-80; librosa
0.9.1 -> resampy `kaiser_best`) and preprocesses corpora with `resampy.resample(data, sr, 16000)`
(I_da/scripts/preprocess.py:43-45), one file at a time on the host.  Here the int16 PCM goes to the GPU as it is on
disk (2 bytes per sample) and one kernel launch per (source rate, target rate) converts and resamples the whole batch.

Filter: a Kaiser-windowed sinc with resampy's published `kaiser_best` parameters (64 zero crossings, roll-off
0.9475937167399596, beta 14.769656459379492), evaluated EXACTLY at the 441 (or 320) fractional delays of the rational
ratio instead of resampy's 512-entries-per-crossing table with linear interpolation.  Equivalence to resampy itself is
parity-unpinned (resampy / librosa are not in the image); the test suite pins this formulation against
`torchaudio.functional.resample(..., resampling_method="sinc_interp_kaiser")` with the same parameters.
"""
from __future__ import annotations

import math
import struct

import numpy as np
import torch

from . import ops
from .ops import SibError

KAISER_BEST = dict(lowpass_filter_width=64, rolloff=0.9475937167399596, beta=14.769656459379492)
_filter_cache = {}


def resample_filter(orig_sr: int, new_sr: int, lowpass_filter_width: int = 64, rolloff: float = 0.9475937167399596,
                    beta: float = 14.769656459379492, window: str = "kaiser"):
    """Poly-phase prototype of the rational resampler `orig_sr -> new_sr`.

    Returns (filt float32 [taps][up], up, down, first): output n reads inputs (n*down)//up + first + j, j < taps, with
    weights filt[j][n % up]."""
    if orig_sr <= 0 or new_sr <= 0:
        raise SibError("sample rates must be positive")
    g = math.gcd(int(orig_sr), int(new_sr))
    down, up = int(orig_sr) // g, int(new_sr) // g
    if up == down:
        return np.ones((1, 1), np.float32), 1, 1, 0
    cutoff = min(up, down) * rolloff            # in units of the common grid rate / (up*down)
    half = int(math.ceil(lowpass_filter_width * down / cutoff))   # input samples either side of the output instant
    taps, first = 2 * half, -half + 1
    r = np.arange(up, dtype=np.int64)
    base_idx = (r * down) // up                  # floor of the output instant on the input grid
    frac_pos = (r * down).astype(np.float64) / up
    j = np.arange(taps, dtype=np.int64)
    d = (base_idx[None, :] + first + j[:, None]).astype(np.float64) - frac_pos[None, :]    # [taps][up] input offsets
    t = np.clip(d * cutoff / down, -lowpass_filter_width, lowpass_filter_width)
    if window == "kaiser":
        win = np.i0(beta * np.sqrt(np.maximum(0.0, 1.0 - (t / lowpass_filter_width) ** 2))) / np.i0(beta)
    elif window == "hann":
        win = np.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    else:
        raise SibError(f"unknown window {window!r}")
    filt = np.sinc(t) * win * (cutoff / down)
    return np.ascontiguousarray(filt.astype(np.float32)), up, down, first


def _device_filter(orig_sr, new_sr, device, **kw):
    key = (int(orig_sr), int(new_sr), str(device), tuple(sorted(kw.items())))
    if key not in _filter_cache:
        f, up, down, first = resample_filter(orig_sr, new_sr, **kw)
        _filter_cache[key] = (torch.from_numpy(f).to(device), up, down, first)
    return _filter_cache[key]


def resampled_length(n: int, orig_sr: int, new_sr: int) -> int:
    """ceil(n * new / orig) in integers (librosa.resample / torchaudio output length)."""
    g = math.gcd(int(orig_sr), int(new_sr))
    down, up = int(orig_sr) // g, int(new_sr) // g
    return -((-int(n) * up) // down)


def resample(x, orig_sr: int, new_sr: int, lengths=None, out=None, **filter_kw):
    """x [B, n] (or [n]) CUDA int16 PCM or float32 at `orig_sr` -> float32 [B, ceil(n*new/orig)] at `new_sr`.
    int16 input is scaled by 1/32768 (soundfile / librosa convention).  `lengths` (per-utterance valid samples) makes
    zero-padded batches behave as if every utterance had been resampled alone (output zero past ceil(len*new/orig));
    `out` is an optional preallocated float32 [B, >= n_out] destination."""
    if not x.is_cuda:
        raise SibError("resample: x must be a CUDA tensor (no CPU fallback)")
    squeeze = x.dim() == 1
    xb = x.reshape(1, -1) if squeeze else x
    if xb.dtype not in (torch.int16, torch.float32):
        xb = xb.to(torch.float32)
    if xb.stride(1) != 1:
        xb = xb.contiguous()
    filt, up, down, first = _device_filter(orig_sr, new_sr, xb.device, **filter_kw)
    n_out = resampled_length(xb.shape[1], orig_sr, new_sr)
    y = torch.empty(xb.shape[0], n_out, device=xb.device, dtype=torch.float32) if out is None else out
    li = lo = None
    if lengths is not None:
        lens = torch.as_tensor(lengths, dtype=torch.int64).cpu()
        li = lens.to(torch.int32).to(xb.device)
        lo = torch.tensor([resampled_length(int(v), orig_sr, new_sr) for v in lens], dtype=torch.int32).to(xb.device)
    ops.resample(xb, filt, up, down, first, y, li, lo)
    return y[0] if squeeze else y


# ----------------------------------------------------------------------------- RIFF/WAVE (16-bit PCM) on the host
def read_wav(path):
    """Minimal RIFF/WAVE reader for 16-bit PCM (what LJSpeech / VCTK / the reference's own prediction/*.wav hold).
    Returns (int16 ndarray [n, channels], sample_rate).  Anything else raises - no silent conversion."""
    with open(path, "rb") as f:
        blob = f.read()
    if len(blob) < 12 or blob[:4] != b"RIFF" or blob[8:12] != b"WAVE":
        raise SibError(f"{path}: not a RIFF/WAVE file")
    pos, fmt, data = 12, None, None
    while pos + 8 <= len(blob):
        cid, size = blob[pos:pos + 4], struct.unpack_from("<I", blob, pos + 4)[0]
        body = blob[pos + 8:pos + 8 + size]
        if cid == b"fmt ":
            fmt = body
        elif cid == b"data":
            data = body
            break
        pos += 8 + size + (size & 1)
    if fmt is None or data is None or len(fmt) < 16:
        raise SibError(f"{path}: missing fmt/data chunk")
    tag, channels, sr, _, block_align, bits = struct.unpack_from("<HHIIHH", fmt, 0)
    if tag == 0xFFFE and len(fmt) >= 26:   # WAVE_FORMAT_EXTENSIBLE: the sub-format GUID starts with the real tag
        tag = struct.unpack_from("<H", fmt, 24)[0]
    if tag != 1 or bits != 16 or channels < 1:
        raise SibError(f"{path}: only 16-bit PCM is supported (format tag {tag}, {bits} bits, {channels} channels)")
    n = len(data) // (2 * channels)
    pcm = np.frombuffer(data, dtype="<i2", count=n * channels).reshape(n, channels)
    return pcm, int(sr)


def write_wav(path, pcm, sample_rate: int):
    """16-bit PCM RIFF/WAVE writer; `pcm` int16 [n] or [n, channels] (numpy or tensor) - the int16 a19 produces."""
    if isinstance(pcm, torch.Tensor):
        pcm = pcm.detach().cpu().numpy()
    pcm = np.asarray(pcm)
    if pcm.dtype != np.int16:
        raise SibError(f"write_wav wants int16 samples (pack_int16 output), got {pcm.dtype}")
    if pcm.ndim == 1:
        pcm = pcm[:, None]
    n, channels = pcm.shape
    payload = np.ascontiguousarray(pcm.astype("<i2")).tobytes()
    hdr = b"RIFF" + struct.pack("<I", 36 + len(payload)) + b"WAVE" + b"fmt " + struct.pack(
        "<IHHIIHH", 16, 1, channels, int(sample_rate), int(sample_rate) * channels * 2, channels * 2, 16)
    with open(path, "wb") as f:
        f.write(hdr + b"data" + struct.pack("<I", len(payload)) + payload)


def load_wav_batch(paths, target_srs=(16000, 22050), device=None):
    """The `librosa.load(path, sr=22050)` + `librosa.load(path, sr=16000)` pair of I_ea/predict.py:79-80 for a batch of
    files: PCM is staged in one pinned int16 buffer, copied once, and converted / resampled on the device.

    Returns {sr: (wave float32 [B, Nmax_sr] zero padded, lengths int32 [B])} with B in the order of `paths`."""
    if device is None:
        if not torch.cuda.is_available():
            raise SibError("load_wav_batch needs a CUDA device (no CPU fallback)")
        device = torch.device("cuda", torch.cuda.current_device())
    files = [read_wav(p) for p in paths]
    if not files:
        raise SibError("load_wav_batch: no files")
    B = len(files)
    n_max = max(f[0].shape[0] for f in files)
    mono16 = all(f[0].shape[1] == 1 for f in files)
    stage = torch.zeros(B, n_max, dtype=torch.int16 if mono16 else torch.float32).pin_memory()
    for i, (pcm, _) in enumerate(files):
        if mono16:
            stage[i, :pcm.shape[0]] = torch.from_numpy(np.array(pcm[:, 0]))
        else:   # librosa.to_mono: mean over channels of the float signal
            stage[i, :pcm.shape[0]] = torch.from_numpy((pcm.astype(np.float32) / 32768.0).mean(axis=1))
    dev = stage.to(device, non_blocking=True)
    n_in = torch.tensor([f[0].shape[0] for f in files], dtype=torch.int32)
    src = [f[1] for f in files]
    out = {}
    for sr in target_srs:
        lens = torch.tensor([resampled_length(int(n), s, sr) for n, s in zip(n_in, src)], dtype=torch.int32)
        wave = torch.empty(B, int(lens.max()), device=device, dtype=torch.float32)
        if len(set(src)) == 1:    # the usual case: one launch for the whole batch, written in place
            resample(dev, src[0], sr, lengths=n_in, out=wave)
        else:                     # one launch per source rate
            for s in sorted(set(src)):
                idx = [i for i, v in enumerate(src) if v == s]
                part = torch.empty(len(idx), wave.shape[1], device=device, dtype=torch.float32)
                resample(dev[idx].contiguous(), s, sr, lengths=n_in[idx], out=part)
                wave[idx] = part
        out[int(sr)] = (wave, lens)
    return out
