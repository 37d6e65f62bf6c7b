"""CPU restatement of the integer / glue steps around the two models.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Integer functions are pure Python
so they can be pinned bit-exactly against the reference's golden wav pair.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------- a1 mask arithmetic
def iea_mask_indices(start_sec: float, end_sec: float, sr16: int = 16000, sr22: int = 22050):
    """I_ea/predict.py:85-90,99-100,133.

    Returns dict(mask_len=L frames, mask_pos=frame index, zero16=(lo,hi) half-open zero range of
    the 16 kHz wave, zero22=(lo,hi) zero range of the 22.05 kHz wave)."""
    mask_ms = int((end_sec - start_sec) * 1000)          # :85-86
    mask_len = mask_ms // 20                              # :87
    start_mask = int(start_sec * sr16)                    # :88
    end_mask = int(end_sec * sr16)                        # :89
    mask_pos = start_mask // 320                          # :90
    zero22 = (start_mask * sr22 // sr16, end_mask * sr22 // sr16)  # :99-100
    zero16 = (mask_pos * 320 + 80, (mask_pos + mask_len) * 320 + 79 - 80)  # :133
    return dict(mask_len=mask_len, mask_pos=mask_pos, zero16=zero16, zero22=zero22)


def iea_zero_range_from_frames(mask_pos: int, mask_len: int):
    """I_ea/predict.py:133 == I_ea/dataset/dataset.py:82 (training uses the same formula)."""
    return (mask_pos * 320 + 80, (mask_pos + mask_len) * 320 + 79 - 80)


def apply_zero_range(wave: np.ndarray, lo: int, hi: int) -> np.ndarray:
    """numpy slice-assign semantics (clamped, empty when hi<=lo)."""
    out = wave.copy()
    out[lo:hi] = 0
    return out


def ida_mask(y: np.ndarray, mask_size: int, sampling_rate: int = 16000):
    """I_da/scripts/inpainting.py:187-191: frame_start=int(sr*3/2); y_inp=(y+1e-6)*mask."""
    frame_start = int(sampling_rate * 3 / 2)
    mask = np.ones_like(y)
    mask[frame_start: frame_start + mask_size] = 0
    return (y + 1e-6) * mask, frame_start


def ida_splice_codes(code: np.ndarray, code_inp: np.ndarray, frame_start: int, mask_size: int,
                     hop: int = 320) -> np.ndarray:
    """I_da/scripts/inpainting.py:207-214: keep predicted units only inside the gap."""
    out = code_inp.copy()
    out[: frame_start // hop] = code[: frame_start // hop]
    out[(frame_start + mask_size) // hop:] = code[(frame_start + mask_size) // hop:]
    return out


def ida_trim(n_samples: int, hop: int = 320):
    """I_da/scripts/inpainting.py:243-256: drop the tail so that n % 1280 == 0.
    Returns (samples_to_remove, code_frames_to_remove, f0_frames_to_remove)."""
    to_remove = n_samples % (16 * 80)
    assert to_remove % hop == 0
    return to_remove, to_remove // hop, to_remove // 80


def ida_matched_frames(n_samples: int, n_code: int, n_f0: int, hop: int = 320) -> int:
    """`match_length([(audio,1),(mask,1),(code,320),(f0,80)])` (I_da/src/multiseries.py:5-73: unit =
    lcm(1,1,320,80) = 320 samples, n_unit = min over series) followed by the 1280-sample tail trim of
    I_da/scripts/inpainting.py:243-256 -> number of code frames fed to the generator."""
    lcm = int(np.lcm.reduce([1, 1, hop, 80]))
    fpu = [lcm // 1, lcm // 1, lcm // hop, lcm // 80]
    n_unit = min(n_samples // fpu[0], n_samples // fpu[1], n_code // fpu[2], n_f0 // fpu[3])
    to_remove, rm_code, _ = ida_trim(n_unit * lcm, hop)
    return n_unit * fpu[2] - rm_code


# ----------------------------------------------------------------------------- a2 z-norm
def processor_znorm(x: torch.Tensor, lengths=None, padding_value: float = 0.0) -> torch.Tensor:
    """HF wav2vec2/feature_extraction_wav2vec2.py:78-97 `zero_mean_unit_var_norm`:
    (x - mean) / sqrt(var + 1e-7) per utterance (population variance); with an attention mask
    the statistics use the first `length` samples and the tail is set to padding_value."""
    out = torch.empty_like(x)
    for i in range(x.shape[0]):
        n = x.shape[1] if lengths is None else int(lengths[i])
        v = x[i, :n].double()
        m, var = v.mean(), v.var(unbiased=False)
        out[i] = ((x[i].double() - m) / torch.sqrt(var + 1e-7)).float()
        if n < x.shape[1]:
            out[i, n:] = padding_value
    return out


def fairseq_layer_norm(x: torch.Tensor) -> torch.Tensor:
    """I_da/src/hubert_feature_reader.py:53-54 `F.layer_norm(x, x.shape)` (eps 1e-5)."""
    return F.layer_norm(x, x.shape)


# ----------------------------------------------------------------------------- a11-a13
def gather_mask_frames(outputs: torch.Tensor, mask_pos, mask_len) -> list:
    """I_ea/predict.py:164-168 per utterance (the reference allocates with mask_len[0]; ragged
    lengths are reproduced per utterance, SURVEY 7 'variable mask lengths')."""
    return [outputs[i, int(mask_pos[i]): int(mask_pos[i]) + int(mask_len[i]), :] for i in range(outputs.shape[0])]


def codebook_center(C: torch.Tensor):
    """I_ea/loss_fn.py:10-14: all_embeds_t=[1,K,D]; center_=mean over K; centred codebook."""
    all_t = C.T
    center = all_t.mean(dim=0)
    return all_t - center[None, :], center


def cos_sim_argmax(values: torch.Tensor, C: torch.Tensor) -> torch.Tensor:
    """I_ea/loss_fn.py:44-46: argmax_k cosine_similarity(v[:,None,:], Cc[None,:,:], dim=-1), eps=1e-8."""
    Cc, _ = codebook_center(C)
    v = values.reshape(-1, values.shape[-1])
    sim = F.cosine_similarity(v.unsqueeze(1), Cc.unsqueeze(0), dim=-1)
    return torch.argmax(sim, dim=1)


def paste_centroids(mel: torch.Tensor, C: torch.Tensor, labels_per_utt: list, mask_pos) -> torch.Tensor:
    """I_ea/predict.py:184-187: mel[b,:,pos:pos+L] = ((C-c)[pred] + c)^T."""
    Cc, center = codebook_center(C)
    out = mel.clone()
    for b, lab in enumerate(labels_per_utt):
        L = lab.numel()
        if L:
            out[b, :, int(mask_pos[b]): int(mask_pos[b]) + L] = (Cc[lab] + center).T
    return out


def extend_mel(spec: torch.Tensor) -> torch.Tensor:
    """I_ea/hifi_gan/inference_modified.py:16-19: bilinear, scale_factor=(1, 441/256),
    align_corners=False -> [B, 80, floor(T*441/256)]."""
    return F.interpolate(spec.unsqueeze(0), scale_factor=(1, 441 / 256), mode="bilinear",
                         align_corners=False).squeeze(0)


def extend_mel_explicit(spec: torch.Tensor) -> torch.Tensor:
    """Closed form of `extend_mel` (what the kernel implements): src = max(0, (dst+0.5)*256/441-0.5)
    with the *given* scale, linear blend of floor / floor+1 clamped to T-1."""
    T = spec.shape[-1]
    Tm = int(np.floor(T * (441 / 256)))
    scale = np.float32(1.0 / (441 / 256))
    dst = torch.arange(Tm, dtype=torch.float32)
    src = torch.clamp((dst + 0.5) * float(scale) - 0.5, min=0.0)
    i0 = src.floor().long().clamp(max=T - 1)
    i1 = (i0 + 1).clamp(max=T - 1)
    w1 = src - i0.float()
    return spec[..., i0] * (1.0 - w1) + spec[..., i1] * w1


# ----------------------------------------------------------------------------- a17
def kmeans_predict(feats: torch.Tensor, centers: torch.Tensor) -> torch.Tensor:
    """I_da/scripts/inpainting.py:204-205 sklearn `KMeans.predict` == argmin_k ||f - mu_k||^2
    (ties -> lowest k)."""
    d = (feats.double()[:, None, :] - centers.double()[None, :, :]).pow(2).sum(-1)
    return torch.argmin(d, dim=1)


def kmeans_predict_f32(feats: torch.Tensor, centers: torch.Tensor) -> torch.Tensor:
    """The same assignment in the arithmetic sklearn itself uses for float32 features (`_labels_inertia` ->
    argmin_k (||mu_k||^2 - 2 f.mu_k), float32): pinned against `sklearn.cluster.KMeans.predict` by
    oracle/make_golden.py:golden_kmeans.  Differs from `kmeans_predict` only on near-ties."""
    f, mu = feats.float(), centers.float()
    return torch.argmin((mu * mu).sum(1)[None, :] - 2.0 * (f @ mu.T), dim=1)


def ida_emb_longtensor(emb):
    """I_da/scripts/inpainting.py:233,239 (and src/dataset.py:437): the speaker d-vector reaches the generator as
    `torch.LongTensor(emb)` - every component truncated toward zero; CodeGenerator.forward then concatenates it with
    the float embeddings (type promotion back to float, src/model.py:139,170-172)."""
    return torch.as_tensor(emb).to(torch.int64)
