"""CPU restatement of the waveform metrics the reference evaluates with (test infrastructure only).

SI-SDR follows I_ea/metrics.py:127-141: project the estimate on the reference with the regularised gain
a = (eps + <r, e>) / (<r, r> + eps), then 10 log10((eps + |a r|^2) / (eps + |e - a r|^2)), eps = machine epsilon of the
ESTIMATE's dtype.  Pinned by oracle/make_golden.py against that very method (its source is lifted out of the file with
`ast`, because `I_ea/metrics.py:7` does not import on any torch version - SURVEY 8b) -> tests/golden/metrics_golden.npz.
"""
from __future__ import annotations

import numpy as np


def sisdr(x_est: np.ndarray, x_ref: np.ndarray, eps=None) -> float:
    e = np.ravel(x_est)
    r = np.ravel(x_ref)
    tiny = np.finfo(e.dtype).eps if eps is None else eps
    gain = (tiny + np.dot(r, e)) / (np.dot(r, r) + tiny)
    target = gain * r
    noise = e - target
    return float(10.0 * np.log10((tiny + np.sum(target ** 2)) / (tiny + np.sum(noise ** 2))))
