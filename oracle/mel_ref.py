"""CPU restatement of the log-mel spectrograms on / next to the hot path.

TEST INFRASTRUCTURE (see oracle/__init__.py).
  * mel-L1 metric mel:  I_ea/hifi_gan/meldataset.py:49-79 (hop 256, pad (n_fft-hop)/2 = 384)
  * I_ea feature mel:   I_ea/dataset/mel_dump.py:40-98     (hop 441, pad 312, fmax 8000)
Both: reflect-pad, STFT n_fft = win = 1024 periodic hann, center=False,
sqrt(re^2+im^2+1e-9), slaney mel 80x513, log(clamp(., 1e-5)).

`librosa.filters.mel` (librosa absent; called positionally => librosa<0.10) is restated below
from its published algorithm (Slaney auditory-toolbox mel scale + area normalisation).
Equivalence to librosa 0.9.1 itself is PARITY UNPINNED; the restatement is pinned against
torchaudio.functional.melscale_fbanks(norm="slaney", mel_scale="slaney") in make_golden.py.
The same filterbank tensor feeds the oracle and the kernel, so it cancels in comparisons.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, mels)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    freqs = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), freqs)


def slaney_mel_filterbank(sr: int, n_fft: int, n_mels: int, fmin: float, fmax) -> np.ndarray:
    """librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax) defaults (htk=False, norm='slaney')
    -> float32 [n_mels, 1 + n_fft//2]."""
    if fmax is None:
        fmax = sr / 2.0
    fftfreqs = np.linspace(0, sr / 2.0, 1 + n_fft // 2)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    w = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2: n_mels + 2] - mel_f[:n_mels])
    return (w * enorm[:, None]).astype(np.float32)


def mel_spectrogram(y: torch.Tensor, n_fft=1024, num_mels=80, sampling_rate=22050, hop_size=256,
                    win_size=1024, fmin=0, fmax=8000, pad=None) -> torch.Tensor:
    """y [B, S] -> [B, num_mels, 1 + (S + 2*pad - n_fft)//hop].
    pad=None => int((n_fft-hop)/2) (meldataset.py:65); mel_dump.py:75 passes 312 with hop 441."""
    if pad is None:
        pad = int((n_fft - hop_size) / 2)
    basis = torch.from_numpy(slaney_mel_filterbank(sampling_rate, n_fft, num_mels, fmin, fmax))
    window = torch.hann_window(win_size)
    y = F.pad(y.unsqueeze(1), (pad, pad), mode="reflect").squeeze(1)
    spec = torch.stft(y, n_fft, hop_length=hop_size, win_length=win_size, window=window, center=False,
                      normalized=False, onesided=True, return_complex=True)
    spec = torch.sqrt(spec.real.pow(2) + spec.imag.pow(2) + 1e-9)
    spec = torch.matmul(basis, spec)
    return torch.log(torch.clamp(spec, min=1e-5))


def feature_mel(y22: torch.Tensor) -> torch.Tensor:
    """`get_mel` I_ea/dataset/mel_dump.py:96-98: hop 441, pad 312, fmax 8000."""
    return mel_spectrogram(y22, hop_size=441, fmax=8000, pad=312)


def peak_normalize(x: np.ndarray) -> np.ndarray:
    """librosa.util.normalize(x) for a 1-D float32 signal (norm=inf, threshold=tiny, fill=None): x / max|x|, unchanged
    when the peak is below the smallest normal float (librosa is absent: restated from its published algorithm;
    oracle/make_golden.py stubs it the same way)."""
    x = np.asarray(x, dtype=np.float32)
    peak = np.abs(x).max() if x.size else np.float32(0)
    return x if peak < np.finfo(np.float32).tiny else (x / peak).astype(np.float32)


def masked_feature_mel(wave22: np.ndarray, lo: int, hi: int) -> torch.Tensor:
    """I_ea/predict.py:99-104: wave_22_masked[start:end] = 0; normalize(.) * 0.95; get_mel -> [1, 80, T']."""
    w = np.array(wave22, dtype=np.float32, copy=True)
    w[lo:hi] = 0
    w = (peak_normalize(w) * np.float32(0.95)).astype(np.float32)
    return feature_mel(torch.from_numpy(w)[None])


def mel_l1(y_a: torch.Tensor, y_b: torch.Tensor, sampling_rate=22050) -> float:
    """mel-L1 acceptance metric: F.l1_loss of hop-256 log-mels with fmax=None
    (I_ea/hifi_gan/train.py:224-227 uses fmax_for_loss = null)."""
    ma = mel_spectrogram(y_a, sampling_rate=sampling_rate, fmax=None)
    mb = mel_spectrogram(y_b, sampling_rate=sampling_rate, fmax=None)
    return float(F.l1_loss(ma, mb))
