"""Shared helpers for the parity tests."""
import math

import torch


def snr_db(ref: torch.Tensor, x: torch.Tensor) -> float:
    ref, x = ref.double().flatten(), x.double().flatten()
    num = ref.pow(2).sum()
    den = (ref - x).pow(2).sum()
    if den == 0:
        return math.inf
    return float(10.0 * torch.log10(num / den))


def max_abs(ref, x) -> float:
    return float((ref.double() - x.double()).abs().max())


def to_frame_major(x):  # [B,C,T] -> [B,T,C]
    return x.transpose(1, 2).contiguous()
