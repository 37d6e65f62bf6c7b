"""Generate tests/golden/* by running the REAL reference (and HF transformers) in this container.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Run from the repo root:

    python -m oracle.make_golden            # needs /root/reference (read-only) + transformers

Nothing at test / bench time reads /root/reference: the outputs below are committed.  Inputs and
weights are regenerated from seeds (oracle/params.py), so only outputs are stored.
The script also asserts that the oracle restatement agrees with the reference on every vector
it writes (the "pin").
"""
from __future__ import annotations

import hashlib
import importlib.machinery
import json
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("SIB_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
sys.dont_write_bytecode = True


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.use = lambda *a, **k: None
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def import_reference():
    """Import shims from SURVEY.md 8c / Appendix A (transformers must be imported first)."""
    import transformers  # noqa: F401
    from transformers import HubertModel, HubertConfig  # noqa: F401
    from oracle.mel_ref import slaney_mel_filterbank

    for n in ["matplotlib", "matplotlib.pylab", "matplotlib.pyplot", "soundfile", "kaldi_io", "fairseq",
              "amfm_decompy", "amfm_decompy.basic_tools", "amfm_decompy.pYAAPT"]:
        _stub(n)
    _stub("npy_append_array", NpyAppendArray=None)
    lib = _stub("librosa")
    lib.util = _stub("librosa.util", normalize=lambda x: x / np.abs(x).max())
    # meldataset.py:62 calls librosa_mel_fn(sr, n_fft, n_mels, fmin, fmax) positionally
    lib.filters = _stub("librosa.filters", mel=lambda sr, n_fft, n_mels, fmin, fmax: slaney_mel_filterbank(
        sr, n_fft, n_mels, fmin, fmax))
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(REF, "I_da"))  # I_da modules import `src.*` relative to I_da/
    import I_ea
    import I_ea.hifi_gan
    import I_ea.dataset
    sys.modules["Inpainting"] = I_ea
    sys.modules["Inpainting.hifi_gan"] = I_ea.hifi_gan
    sys.modules["Inpainting.dataset"] = I_ea.dataset
    return I_ea


class AttrDict(dict):
    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.__dict__ = self


def golden_mask():
    """a1: the only golden vector in the reference tree."""
    from scipy.io import wavfile
    from oracle.glue_ref import iea_zero_range_from_frames
    d = os.path.join(REF, "I_ea/prediction/LJ050-0271")
    sr_o, orig = wavfile.read(os.path.join(d, "orig.wav"))
    sr_m, masked = wavfile.read(os.path.join(d, "masked.wav"))
    assert sr_o == sr_m == 16000 and orig.shape == masked.shape
    from oracle.glue_ref import apply_zero_range
    diff = np.nonzero(orig != masked)[0]
    pos, L = 149, 20
    lo, hi = iea_zero_range_from_frames(pos, L)
    # THE PIN: replaying predict.py:133 on orig.wav reproduces masked.wav bit for bit
    assert np.array_equal(masked, apply_zero_range(orig, lo, hi))
    assert lo <= int(diff.min()) and int(diff.max()) < hi
    out = dict(n_samples=int(len(orig)), sample_rate=16000, mask_pos=pos, mask_len=L,
               first_diff=int(diff.min()), last_diff_plus1=int(diff.max()) + 1,
               zero_range=[int(lo), int(hi)], n_zeroed=int(hi - lo),
               orig_sha256=hashlib.sha256(orig.tobytes()).hexdigest(),
               masked_sha256=hashlib.sha256(masked.tobytes()).hexdigest(),
               # edge windows [lo-4, lo+4) and [hi-4, hi+4) of both files, to replay the edit in tests
               orig_lo_window=orig[lo - 4: lo + 4].tolist(), masked_lo_window=masked[lo - 4: lo + 4].tolist(),
               orig_hi_window=orig[hi - 4: hi + 4].tolist(), masked_hi_window=masked[hi - 4: hi + 4].tolist(),
               source="I_ea/prediction/LJ050-0271/{orig,masked}.wav; predict.py:133")
    with open(os.path.join(OUT, "mask_golden.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("mask golden:", out["zero_range"], "n_zeroed", out["n_zeroed"])


def golden_hubert():
    from transformers import HubertModel, HubertConfig
    from oracle.params import HubertCfg, make_hubert_params
    from oracle.hubert_ref import hubert_forward

    def hf_cfg(c: HubertCfg):
        return HubertConfig(hidden_size=c.hidden_size, num_hidden_layers=c.num_hidden_layers,
                            num_attention_heads=c.num_attention_heads, intermediate_size=c.intermediate_size,
                            feat_extract_norm=c.feat_extract_norm, conv_bias=c.conv_bias,
                            do_stable_layer_norm=c.do_stable_layer_norm, conv_dim=list(c.conv_dim),
                            conv_kernel=list(c.conv_kernel), conv_stride=list(c.conv_stride),
                            num_conv_pos_embeddings=c.num_conv_pos_embeddings,
                            num_conv_pos_embedding_groups=c.num_conv_pos_embedding_groups,
                            attn_implementation="eager")

    res = {}
    cases = [("tiny_group", HubertCfg.tiny(False), 2, 3000, True), ("tiny_layer", HubertCfg.tiny(True), 2, 3000, True),
             ("base", HubertCfg.base(), 1, 4000, False), ("large", HubertCfg.large(), 1, 2400, False)]
    for name, cfg, B, N, with_pad in cases:
        params = make_hubert_params(cfg, seed=1234)
        model = HubertModel(hf_cfg(cfg)).eval()
        missing = model.load_state_dict(params, strict=True)
        g = torch.Generator().manual_seed(99)
        x = 0.1 * torch.randn(B, N, generator=g)
        with torch.no_grad():
            y_hf = model(x).last_hidden_state
            y_or = hubert_forward(params, cfg, x)
        err = (y_hf - y_or).abs().max().item()
        print(f"hubert {name}: HF vs oracle max-abs {err:.3e}  out {tuple(y_hf.shape)} missing={missing}")
        assert err < 5e-5, err
        res[name + "_out"] = y_hf.numpy()
        if with_pad:
            am = torch.ones(B, N, dtype=torch.long)
            am[1, N - 900:] = 0
            xp = x.clone()
            xp[1, N - 900:] = 0
            with torch.no_grad():
                y_hf = model(xp, attention_mask=am).last_hidden_state
                y_or = hubert_forward(params, cfg, xp, am)
            err = (y_hf - y_or).abs().max().item()
            print(f"hubert {name} padded: max-abs {err:.3e}")
            assert err < 5e-5, err
            res[name + "_padded_out"] = y_hf.numpy()
        del model
    np.savez_compressed(os.path.join(OUT, "hubert_golden.npz"), **res)


def golden_hifigan():
    from I_ea.hifi_gan.models import Generator
    from oracle.params import HifiCfg, make_generator_params
    from oracle.hifigan_ref import generator_forward

    res = {}
    # V1 / V2 (ResBlock1) and V3 (ResBlock2, models.py:52-73): the three generator configs the reference ships
    for name, cfg, B, T, init in [("v1_unit", HifiCfg.v1(), 2, 6, "unit"), ("v1_ref", HifiCfg.v1(), 1, 5, "reference"),
                                  ("tiny_unit", HifiCfg.tiny(), 2, 9, "unit"), ("v2_unit", HifiCfg.v2(), 2, 7, "unit"),
                                  ("v3_unit", HifiCfg.v3(), 2, 11, "unit"), ("v3_ref", HifiCfg.v3(), 1, 6, "reference")]:
        params = make_generator_params(cfg, seed=1234, init=init)
        h = AttrDict(cfg.as_attrdict())
        gen = Generator(h)
        # torch>=2.1 legacy weight_norm still exposes weight_g / weight_v
        gen.load_state_dict({k: v for k, v in params.items() if not k.startswith("emb_")}, strict=True)
        gen.eval()
        gen.remove_weight_norm()
        x = torch.randn(B, cfg.model_in_dim, T, generator=torch.Generator().manual_seed(7))
        with torch.no_grad():
            y_ref = gen(x)
            y_or = generator_forward(params, cfg, x)
        err = (y_ref - y_or).abs().max().item()
        print(f"hifigan {name}: reference vs oracle max-abs {err:.3e} |y|max {y_ref.abs().max():.3e} out {tuple(y_ref.shape)}")
        assert err < 1e-5 * max(1.0, float(y_ref.abs().max())), err
        res[name + "_out"] = y_ref.numpy()

    # I_da generator (src.models.Generator); needs cwd-relative `src` package
    from src.models import Generator as IdaGenerator
    for name, cfg, B, T in [("ida_unit", HifiCfg.ida(), 1, 4), ("ida_tiny", HifiCfg.tiny(True), 2, 8)]:
        params = make_generator_params(cfg, seed=1234, init="unit")
        gen = IdaGenerator(AttrDict(cfg.as_attrdict()))
        gen.load_state_dict({k: v for k, v in params.items() if not k.startswith("emb_")}, strict=True)
        gen.eval()
        gen.remove_weight_norm()
        x = torch.randn(B, cfg.model_in_dim, T, generator=torch.Generator().manual_seed(7))
        with torch.no_grad():
            y_ref = gen(x)
            y_or = generator_forward(params, cfg, x)
        err = (y_ref - y_or).abs().max().item()
        print(f"hifigan {name}: reference vs oracle max-abs {err:.3e} out {tuple(y_ref.shape)}")
        assert err < 1e-5, err
        res[name + "_out"] = y_ref.numpy()
    np.savez_compressed(os.path.join(OUT, "hifigan_golden.npz"), **res)


def golden_glue():
    from oracle import glue_ref
    from oracle.params import make_codebook
    res = {}
    # extend_mel (inference_modified.py:16-19)
    from I_ea.hifi_gan.inference_modified import extend_mel as ref_extend
    for T in (37, 100, 200):
        spec = torch.randn(2, 80, T, generator=torch.Generator().manual_seed(T))
        y_ref = ref_extend(spec)
        assert torch.equal(y_ref, glue_ref.extend_mel(spec))
        err = (y_ref - glue_ref.extend_mel_explicit(spec)).abs().max().item()
        print(f"extend_mel T={T}: -> {tuple(y_ref.shape)} explicit-form max-abs {err:.2e}")
        assert err < 2e-5 and y_ref.shape[-1] == int(T * 441 / 256)
        res[f"extend_mel_{T}"] = y_ref.numpy()
    # cos_sim argmax (loss_fn.py:26-47) - build LossFunction without its kmeans file
    from I_ea.loss_fn import LossFunction
    for K in (100, 500):
        C = make_codebook(80, K, seed=77)
        lf = object.__new__(LossFunction)
        lf.all_embeds = C
        lf.all_embeds_t = C.T[None, :, :]
        lf.center_ = lf.all_embeds_t.squeeze(0).mean(dim=0)
        lf.all_embeds_t_c = lf.all_embeds_t - lf.center_.unsqueeze(0).unsqueeze(0)
        vals = torch.randn(3, 10, 80, generator=torch.Generator().manual_seed(K))
        labels = torch.zeros(3, 10, dtype=torch.long)
        _, pred = lf.cos_sim(vals, labels)
        mine = glue_ref.cos_sim_argmax(vals, C).view(3, 10)
        assert torch.equal(pred, mine)
        res[f"cos_sim_pred_{K}"] = pred.numpy()
        # paste (predict.py:184-187)
        mel = torch.randn(1, 80, 50, generator=torch.Generator().manual_seed(5))
        ref_mel = mel.clone()
        pred_mels = lf.all_embeds_t_c[0, pred[0, :], :] + lf.center_
        ref_mel[0, :, 7:7 + 10] = pred_mels.T
        assert torch.equal(ref_mel, glue_ref.paste_centroids(mel, C, [pred[0]], [7]))
        res[f"paste_{K}"] = ref_mel.numpy()
    print("cos_sim / paste pinned")
    # _upsample (I_da/src/model.py:78-119) - static method, importable with stubs
    try:
        for n in ["src.modules.dvector", "src.modules.ge2e", "src.modules.ge2e_dataset",
                  "src.modules.infinite_dataloader", "src.modules.wav2mel"]:
            _stub(n, AttentivePooledLSTMDvector=None, GE2ELoss=None, GE2EDataset=None, collate_batch=None,
                  InfiniteDataLoader=None, infinite_iterator=None, Wav2Mel=None)
        from src.model import CodeGenerator
        sig = torch.randn(2, 16, 5)
        assert torch.equal(CodeGenerator._upsample(sig, 20), __import__("oracle.hifigan_ref", fromlist=["x"]).upsample_repeat(sig, 20))
        emb = torch.randn(2, 16)
        assert torch.equal(CodeGenerator._upsample(emb, 20), __import__("oracle.hifigan_ref", fromlist=["x"]).upsample_repeat(emb, 20))
        print("CodeGenerator._upsample pinned")
        res["upsample_pinned"] = np.array([1])
    except Exception as e:  # pragma: no cover
        print("CodeGenerator import failed (parity unpinned for _upsample):", repr(e))
        res["upsample_pinned"] = np.array([0])
    np.savez_compressed(os.path.join(OUT, "glue_golden.npz"), **res)


def golden_kmeans():
    """a17: `kmeans_model.predict(feats)` (I_da/scripts/inpainting.py:204-205) with the REAL sklearn estimator
    (joblib-loaded `KMeans` in the reference; here one whose fitted attributes are set directly - predict only reads
    cluster_centers_).  Pins oracle.glue_ref.kmeans_predict (exact argmin in float64) and kmeans_predict_f32."""
    import sklearn
    from sklearn.cluster import KMeans
    from oracle import glue_ref
    res = {"sklearn_version": np.array(sklearn.__version__)}
    for K, H, M in ((100, 768, 400), (500, 768, 600), (500, 1024, 300)):
        g = torch.Generator().manual_seed(K + H)
        mu = torch.randn(K, H, generator=g) * 0.5
        # features near the centroids (as real HuBERT features are) plus plain noise rows: unambiguous and hard cases
        idx = torch.randint(0, K, (M,), generator=g)
        f = torch.cat([mu[idx[: M // 2]] + 0.3 * torch.randn(M // 2, H, generator=g), torch.randn(M - M // 2, H, generator=g) * 0.5])
        km = KMeans(n_clusters=K, n_init=1)
        km.cluster_centers_ = mu.numpy().astype(np.float32)
        km._n_threads = 1
        km.n_features_in_ = H
        km._n_features_out = K
        want = km.predict(f.numpy())
        mine = glue_ref.kmeans_predict(f, mu).numpy()
        mine32 = glue_ref.kmeans_predict_f32(f, mu).numpy()
        agree, agree32 = float((want == mine).mean()), float((want == mine32).mean())
        print(f"kmeans K={K} H={H}: sklearn {sklearn.__version__} predict vs oracle exact-argmin {agree:.4f}, vs f32 form {agree32:.4f}")
        assert agree == 1.0, "oracle kmeans_predict differs from sklearn KMeans.predict"
        res[f"labels_{K}_{H}"] = want.astype(np.int64)
    np.savez_compressed(os.path.join(OUT, "kmeans_golden.npz"), **res)


def golden_mel():
    import torchaudio
    from oracle import mel_ref
    res = {}
    for fmax in (8000, None):
        mine = mel_ref.slaney_mel_filterbank(22050, 1024, 80, 0, fmax)
        ta = torchaudio.functional.melscale_fbanks(513, 0.0, float(fmax or 22050 / 2), 80, 22050,
                                                   norm="slaney", mel_scale="slaney").T.numpy()
        err = np.abs(mine - ta).max()
        print(f"mel filterbank fmax={fmax}: vs torchaudio max-abs {err:.2e}")
        assert err < 1e-6
    from I_ea.hifi_gan.meldataset import mel_spectrogram as ref_mel
    from I_ea.dataset.mel_dump import get_mel as ref_get_mel
    y = 0.3 * torch.randn(2, 22050, generator=torch.Generator().manual_seed(3)).clamp(-3, 3)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m_ref = ref_mel(y, 1024, 80, 22050, 256, 1024, 0, None)
        m_feat = ref_get_mel(y)
    m_or = mel_ref.mel_spectrogram(y, fmax=None)
    f_or = mel_ref.feature_mel(y)
    e1, e2 = (m_ref - m_or).abs().max().item(), (m_feat - f_or).abs().max().item()
    print(f"mel_spectrogram hop256 {tuple(m_ref.shape)} max-abs {e1:.2e}; feature mel hop441 {tuple(m_feat.shape)} max-abs {e2:.2e}")
    assert e1 < 1e-4 and e2 < 1e-4
    res["mel_hop256"] = m_ref.numpy()
    res["mel_hop441"] = m_feat.numpy()
    np.savez_compressed(os.path.join(OUT, "mel_golden.npz"), **res)


def golden_resample():
    """8f row 3: the oracle resampler against torchaudio's Kaiser-windowed sinc with resampy's kaiser_best parameters
    (float64, so the comparison sees the formulation, not fp32 kernel rounding)."""
    import torchaudio
    from oracle import resample_ref as R
    res = {}
    x = 0.25 * torch.randn(12000, generator=torch.Generator().manual_seed(9), dtype=torch.float64)
    x[3000:3400] = 0.0
    res["x"] = x.numpy().astype(np.float32)
    for o, n in ((16000, 22050), (22050, 16000), (48000, 16000), (24000, 22050)):
        ta = torchaudio.functional.resample(torch.from_numpy(res["x"].astype(np.float64)), o, n, lowpass_filter_width=R.ZEROS,
                                            rolloff=R.ROLLOFF, resampling_method="sinc_interp_kaiser", beta=R.BETA).numpy()
        mine = R.resample(res["x"], o, n)
        err = np.abs(ta - mine).max()
        print(f"resample {o}->{n}: {mine.shape[0]} samples, oracle vs torchaudio(float64) max-abs {err:.2e}")
        assert ta.shape == mine.shape and err < 2e-7   # torchaudio holds beta as a float32 tensor
        res[f"y_{o}_{n}"] = ta
    np.savez_compressed(os.path.join(OUT, "resample_golden.npz"), **res)


def golden_f0vq():
    """8f row 2: the reference's own Jukebox Encoder (I_da/src/modules/jukebox.py) and BottleneckBlock.quantise
    (vq.py:118-128; the class itself cannot be built without a GPU - reset_k() calls .cuda(), vq.py:22)."""
    from oracle import f0vq_ref
    for n in ["src.modules.dvector", "src.modules.ge2e", "src.modules.ge2e_dataset",
              "src.modules.infinite_dataloader", "src.modules.wav2mel"]:
        if n not in sys.modules:
            _stub(n, AttentivePooledLSTMDvector=None, GE2ELoss=None, GE2EDataset=None, collate_batch=None,
                  InfiniteDataLoader=None, infinite_iterator=None, Wav2Mel=None)
    from src.modules.jukebox import Encoder
    from src.modules.vq import BottleneckBlock
    cfg = f0vq_ref.F0_QUANTIZER
    enc = Encoder(**cfg["f0_encoder_params"]).eval()
    sd = f0vq_ref.make_params(cfg, seed=1234)
    enc_sd = {k[len("encoder."):]: v for k, v in sd.items() if k.startswith("encoder.")}
    enc.load_state_dict(enc_sd)   # strict: key names and shapes of the restatement == the reference module's
    res = {}
    for B, L in ((2, 784), (1, 160)):
        f0 = torch.randn(B, 1, L, generator=torch.Generator().manual_seed(L))
        with torch.no_grad():
            h_ref = enc(f0)[0]
            z_ref, _ = BottleneckBlock.quantise(types.SimpleNamespace(k=sd["vq.level_blocks.0.k"]),
                                                h_ref.permute(0, 2, 1).reshape(-1, h_ref.shape[1]))
        h = f0vq_ref.encoder_forward(sd, f0)
        z = f0vq_ref.quantise(h, sd["vq.level_blocks.0.k"])
        err = (h - h_ref).abs().max().item()
        print(f"f0 encoder B={B} L={L}: -> {tuple(h_ref.shape)} max-abs {err:.2e}; bins equal: {torch.equal(z.view(-1), z_ref)}")
        assert err < 1e-5 and torch.equal(z.view(-1), z_ref)
        res[f"h_{L}"] = h_ref.numpy()
        res[f"z_{L}"] = z_ref.view(B, -1).numpy()
    np.savez_compressed(os.path.join(OUT, "f0vq_golden.npz"), **res)


def golden_metrics():
    """8f row 4: Metrics.sisdr lifted out of I_ea/metrics.py with ast (the module itself fails to import)."""
    import ast
    from oracle import metrics_ref
    src = open(os.path.join(REF, "I_ea", "metrics.py")).read()
    fn = next(n for n in ast.walk(ast.parse(src)) if isinstance(n, ast.FunctionDef) and n.name == "sisdr")
    ns = {"np": np}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "I_ea/metrics.py", "exec"), ns)
    g = torch.Generator().manual_seed(21)
    ref = 0.3 * torch.randn(3, 22050, generator=g)
    est = ref * torch.tensor([[1.0], [0.5], [2.0]]) + torch.tensor([[0.01], [0.1], [0.5]]) * torch.randn(3, 22050, generator=g)
    vals = []
    for b in range(3):
        e, r = est[b].numpy().astype(np.float64), ref[b].numpy().astype(np.float64)
        want = float(ns["sisdr"](None, e, r))
        got = metrics_ref.sisdr(e, r)
        assert abs(want - got) < 1e-9, (want, got)
        vals.append(want)
    print("sisdr pinned:", [round(v, 4) for v in vals])
    np.savez_compressed(os.path.join(OUT, "metrics_golden.npz"), sisdr=np.array(vals))


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    import_reference()
    golden_mask()
    golden_kmeans()
    golden_glue()
    golden_mel()
    golden_hifigan()
    golden_hubert()
    golden_resample()
    golden_f0vq()
    golden_metrics()
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
