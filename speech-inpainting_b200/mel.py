"""Log-mel spectrogram on the B200 STFT kernel.

`mel_spectrogram(y, n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax, center=False)` keeps
the reference signature (I_ea/hifi_gan/meldataset.py:49-79); `get_mel` is I_ea/dataset/mel_dump.py:96-98
(hop 441, pad 312).  The slaney filterbank (librosa.filters.mel defaults, librosa<0.10 positional call
at meldataset.py:62) is built on the host once per (sr, n_fft, n_mels, fmin, fmax) - equivalence to
librosa itself is parity-unpinned (librosa is not in the image), see DESIGN.md.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from .ops import SibError

_basis_cache = {}


def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp, min_log_hz = 200.0 / 3, 1000.0
    min_log_mel, logstep = min_log_hz / f_sp, np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, f / f_sp)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp, min_log_hz = 200.0 / 3, 1000.0
    min_log_mel, logstep = min_log_hz / f_sp, np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank(sr, n_fft, n_mels, fmin, fmax) -> np.ndarray:
    """Slaney-scale, area-normalised triangular filters -> float32 [n_mels, 1 + n_fft//2]."""
    fmax = sr / 2.0 if fmax is None else fmax
    fftfreqs = np.linspace(0, sr / 2.0, 1 + n_fft // 2)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    w = np.maximum(0, np.minimum(-ramps[:-2] / fdiff[:-1, None], ramps[2:] / fdiff[1:, None]))
    return (w * (2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels]))[:, None]).astype(np.float32)


def mel_spectrogram(y, n_fft=1024, num_mels=80, sampling_rate=22050, hop_size=256, win_size=1024, fmin=0,
                    fmax=8000, center=False, pad=None):
    """y [B, S] CUDA float32 -> log-mel [B, num_mels, frames] (meldataset.py:49-79)."""
    if n_fft != 1024 or win_size != 1024 or center:
        raise SibError("the STFT kernel is specialised for n_fft = win_size = 1024, center=False (the reference's settings)")
    if not y.is_cuda:
        raise SibError("mel_spectrogram: y must be a CUDA tensor (no CPU fallback)")
    y = y.to(torch.float32).contiguous()
    pad = int((n_fft - hop_size) / 2) if pad is None else pad
    key = (sampling_rate, n_fft, num_mels, fmin, fmax, y.device)
    if key not in _basis_cache:
        _basis_cache[key] = torch.from_numpy(mel_filterbank(sampling_rate, n_fft, num_mels, fmin, fmax)).to(y.device)
    B, S = y.shape
    frames = 1 + (S + 2 * pad - n_fft) // hop_size
    out = torch.empty(B, num_mels, frames, device=y.device, dtype=torch.float32)
    ops.mel_spectrogram(y, _basis_cache[key], out, hop_size, pad)
    return out


def get_mel(x, hop_size=441):
    """I_ea/dataset/mel_dump.py:96-98 (n_fft 1024, 80 mels, 22.05 kHz, pad 312, fmax 8000)."""
    return mel_spectrogram(x, 1024, 80, 22050, hop_size, 1024, 0, 8000, pad=312)


def masked_feature_mel(wave22, zero22=None, scale=0.95):
    """The feature front-end of I_ea/predict.py:99-104 on the device: zero [lo, hi) of every 22.05 kHz utterance
    (`zero22` = list of (lo, hi) or None), `normalize(.) * 0.95` (librosa.util.normalize), `get_mel`.
    wave22 [B, S] float32 (host or CUDA) -> log-mel [B, 80, S // 441]."""
    dev = wave22.device if wave22.is_cuda else torch.device("cuda", torch.cuda.current_device())
    x = wave22.to(dev, torch.float32, non_blocking=True).contiguous()
    lo = hi = None
    if zero22 is not None:
        if len(zero22) != x.shape[0]:
            raise SibError("zero22 must have one (lo, hi) pair per utterance")
        lo = torch.as_tensor([int(r[0]) for r in zero22], dtype=torch.int32).to(dev)
        hi = torch.as_tensor([int(r[1]) for r in zero22], dtype=torch.int32).to(dev)
    y = torch.empty_like(x)
    ops.mask_peak_normalize(x, y, lo, hi, scale)
    return get_mel(y)


def mel_l1(y_a, y_b, sampling_rate=22050, per_utterance=False):
    """mel-L1 acceptance metric (hop 256, fmax=None; I_ea/hifi_gan/train.py:224-227), reduced on the device
    (SURVEY 8f row 4).  Returns the batch mean as a float, or the per-utterance means [B] with `per_utterance`."""
    ma = mel_spectrogram(y_a, sampling_rate=sampling_rate, fmax=None)
    mb = mel_spectrogram(y_b, sampling_rate=sampling_rate, fmax=None)
    per = torch.empty(ma.shape[0], device=ma.device, dtype=torch.float32)
    ops.abs_diff_mean(ma, mb, per)
    return per if per_utterance else float(per.double().mean())


def si_sdr(est, ref, lengths=None):
    """Scale-invariant SDR in dB per utterance (I_ea/metrics.py:127-141) for CUDA float32 [B, n] (or [n]) waveforms."""
    if not (est.is_cuda and ref.is_cuda):
        raise SibError("si_sdr: operands must be CUDA tensors (no CPU fallback)")
    squeeze = est.dim() == 1
    e = est.reshape(1, -1) if squeeze else est
    r = ref.reshape(1, -1) if squeeze else ref
    e, r = e.to(torch.float32).contiguous(), r.to(torch.float32).contiguous()
    if lengths is not None:
        lengths = torch.as_tensor(lengths, dtype=torch.int32).to(e.device)
    out = torch.empty(e.shape[0], device=e.device, dtype=torch.float32)
    ops.si_sdr(e, r, out, lengths)
    return out[0] if squeeze else out
