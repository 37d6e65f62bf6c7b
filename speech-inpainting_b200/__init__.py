"""B200-native (sm_100a) implementation of Speech-Inpainting's inference hot path:
masked 16 kHz waveform -> HuBERT encoder -> head / codebook assignment -> HiFi-GAN generator -> waveform.

Public surface mirrors the reference's modules: `HubertModel`, `CustomModel`,
`Generator`, `CodeGenerator`, `mel_spectrogram`, `extend_mel`, plus the batched pipelines
`InformedInpainter` (I_ea/predict.py) and `BlindInpainter` (I_da/scripts/inpainting.py).
Everything executes in `libsib_b200.so` (hand-written CUDA); importing fails if it is not built.
"""
from ._lib import SibError, lib as _load_lib, exported_symbols  # noqa: F401
from .hubert import HubertConfig, HubertModel, CustomModel  # noqa: F401
from .hifigan import Generator, CodeGenerator, AttrDict, get_padding  # noqa: F401
from .f0vq import F0Quantizer  # noqa: F401
from .mel import mel_spectrogram, get_mel, mel_l1, si_sdr, mel_filterbank, masked_feature_mel  # noqa: F401
from .audio import resample, resample_filter, read_wav, write_wav, load_wav_batch  # noqa: F401
from .inpaint import (InformedInpainter, BlindInpainter, iea_mask_indices, iea_zero_range, extend_mel,  # noqa: F401
                      shard_batch, ida_matched_frames, predict_files)
from . import ops  # noqa: F401

__version__ = "0.1.0"
