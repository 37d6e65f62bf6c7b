"""CPU (-m "not gpu"): the oracle against the golden vectors produced by the REAL reference
(oracle/make_golden.py ran /root/reference + transformers in the build container; nothing here reads /root/reference)."""
import json

import numpy as np
import pytest
import torch

from util import max_abs


def test_mask_golden_from_reference_wav_pair(golden_dir):
    """a1 - the only golden vector in the reference tree: I_ea/prediction/LJ050-0271/{orig,masked}.wav."""
    from oracle import glue_ref
    g = json.load(open(f"{golden_dir}/mask_golden.json"))
    lo, hi = glue_ref.iea_zero_range_from_frames(g["mask_pos"], g["mask_len"])
    assert [lo, hi] == g["zero_range"] == [47760, 54079]
    assert hi - lo == g["n_zeroed"] == 20 * 320 - 81
    assert lo <= g["first_diff"] and g["last_diff_plus1"] <= hi
    # replay predict.py:133 on the edge windows of the real orig.wav -> must give masked.wav's samples
    w = np.array(g["orig_lo_window"], dtype=np.int16)
    assert glue_ref.apply_zero_range(w, 4, 8).tolist() == g["masked_lo_window"]
    w = np.array(g["orig_hi_window"], dtype=np.int16)
    assert glue_ref.apply_zero_range(w, 0, 4).tolist() == g["masked_hi_window"]


def test_mask_indices_config_values():
    from oracle import glue_ref
    # 200 ms at 0.9 s -> pos 45, L 10, 3119 zeroed samples; 400 ms -> L 20, 6319 samples
    m = glue_ref.iea_mask_indices(0.9, 1.1)
    assert (m["mask_pos"], m["mask_len"], m["zero16"]) == (45, 10, (14480, 17599)) and m["zero16"][1] - m["zero16"][0] == 3119
    m = glue_ref.iea_mask_indices(2.98, 3.38)
    assert m["mask_len"] == 19  # int((3.38-2.98)*1000) = 399 -> 19 frames: the float truncation of predict.py:85-87
    assert glue_ref.iea_mask_indices(1.0, 1.4)["mask_len"] == 19   # (1.4-1.0)*1000 = 399.99999999999994 as well
    m = glue_ref.iea_mask_indices(0.5, 0.9)
    assert m["mask_len"] == 20 and m["zero16"][1] - m["zero16"][0] == 6319
    assert m["zero22"] == (8000 * 22050 // 16000, int(0.9 * 16000) * 22050 // 16000)
    y = np.arange(64000, dtype=np.float32)
    y_inp, fs = glue_ref.ida_mask(y, 6400)
    assert fs == 24000 and np.all(y_inp[24000:30400] == 0) and y_inp[23999] == np.float32(23999 + 1e-6)
    code, code_inp = np.arange(199), -np.arange(199)
    out = glue_ref.ida_splice_codes(code, code_inp, fs, 6400)
    assert np.array_equal(out[:75], code[:75]) and np.array_equal(out[75:95], code_inp[75:95]) and np.array_equal(out[95:], code[95:])
    assert glue_ref.ida_trim(63680) == (960, 3, 12) and glue_ref.ida_matched_frames(64000, 199, 800) == 196


@pytest.mark.parametrize("name,B,N,pad", [("tiny_group", 2, 3000, True), ("tiny_layer", 2, 3000, True), ("base", 1, 4000, False)])
def test_hubert_oracle_vs_transformers_golden(golden_dir, name, B, N, pad):
    from oracle import hubert_ref
    from oracle.params import HubertCfg, make_hubert_params
    gold = np.load(f"{golden_dir}/hubert_golden.npz")
    cfg = {"tiny_group": HubertCfg.tiny(False), "tiny_layer": HubertCfg.tiny(True), "base": HubertCfg.base()}[name]
    params = make_hubert_params(cfg, 1234)
    x = 0.1 * torch.randn(B, N, generator=torch.Generator().manual_seed(99))
    with torch.no_grad():
        y = hubert_ref.hubert_forward(params, cfg, x)
    assert max_abs(torch.from_numpy(gold[name + "_out"]), y) < 5e-5
    if pad:
        am = torch.ones(B, N, dtype=torch.long)
        am[1, N - 900:] = 0
        xp = x.clone()
        xp[1, N - 900:] = 0
        with torch.no_grad():
            y = hubert_ref.hubert_forward(params, cfg, xp, am)
        assert max_abs(torch.from_numpy(gold[name + "_padded_out"]), y) < 5e-5


def test_hubert_oracle_live_pin_against_transformers():
    """transformers is an installed library (not /root/reference): pin the restatement live as well."""
    transformers = pytest.importorskip("transformers")
    from oracle import hubert_ref
    from oracle.params import HubertCfg, make_hubert_params
    cfg = HubertCfg.tiny(True)
    hf = transformers.HubertConfig(hidden_size=cfg.hidden_size, num_hidden_layers=cfg.num_hidden_layers,
                                   num_attention_heads=cfg.num_attention_heads, intermediate_size=cfg.intermediate_size,
                                   feat_extract_norm=cfg.feat_extract_norm, conv_bias=cfg.conv_bias,
                                   do_stable_layer_norm=cfg.do_stable_layer_norm, conv_dim=list(cfg.conv_dim),
                                   conv_kernel=list(cfg.conv_kernel), conv_stride=list(cfg.conv_stride),
                                   num_conv_pos_embeddings=cfg.num_conv_pos_embeddings,
                                   num_conv_pos_embedding_groups=cfg.num_conv_pos_embedding_groups, attn_implementation="eager")
    params = make_hubert_params(cfg, 7)
    model = transformers.HubertModel(hf).eval()
    model.load_state_dict(params, strict=True)
    x = 0.1 * torch.randn(2, 2500, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        assert max_abs(model(x).last_hidden_state, hubert_ref.hubert_forward(params, cfg, x)) < 5e-5


@pytest.mark.parametrize("name,kind,B,T,init", [("v1_unit", "v1", 2, 6, "unit"), ("v1_ref", "v1", 1, 5, "reference"),
                                                ("tiny_unit", "tiny", 2, 9, "unit"), ("ida_unit", "ida", 1, 4, "unit"),
                                                ("ida_tiny", "ida_tiny", 2, 8, "unit"), ("v2_unit", "v2", 2, 7, "unit"),
                                                ("v3_unit", "v3", 2, 11, "unit"), ("v3_ref", "v3", 1, 6, "reference")])
def test_generator_oracle_vs_reference_golden(golden_dir, name, kind, B, T, init):
    """V1 / V2 (ResBlock1), V3 (ResBlock2, models.py:52-73) and the I_da generator against the reference's own modules."""
    from oracle import hifigan_ref
    from oracle.params import HifiCfg, make_generator_params
    gold = torch.from_numpy(np.load(f"{golden_dir}/hifigan_golden.npz")[name + "_out"])
    cfg = {"v1": HifiCfg.v1(), "v2": HifiCfg.v2(), "v3": HifiCfg.v3(), "tiny": HifiCfg.tiny(), "ida": HifiCfg.ida(),
           "ida_tiny": HifiCfg.tiny(True)}[kind]
    params = make_generator_params(cfg, 1234, init)
    x = torch.randn(B, cfg.model_in_dim, T, generator=torch.Generator().manual_seed(7))
    with torch.no_grad():
        y = hifigan_ref.generator_forward(params, cfg, x)
    assert y.shape == gold.shape == (B, 1, T * cfg.total_upsample)
    assert max_abs(gold, y) < 1e-5 * max(1.0, float(gold.abs().max()))


def test_kmeans_oracle_vs_sklearn_golden(golden_dir):
    """a17: `kmeans_predict` against labels produced by the real `sklearn.cluster.KMeans.predict`
    (oracle/make_golden.py:golden_kmeans), and live against the installed sklearn when it is importable."""
    from oracle import glue_ref
    gold = np.load(f"{golden_dir}/kmeans_golden.npz")
    for K, H, M in ((100, 768, 400), (500, 768, 600), (500, 1024, 300)):
        g = torch.Generator().manual_seed(K + H)
        mu = torch.randn(K, H, generator=g) * 0.5
        idx = torch.randint(0, K, (M,), generator=g)
        f = torch.cat([mu[idx[: M // 2]] + 0.3 * torch.randn(M // 2, H, generator=g), torch.randn(M - M // 2, H, generator=g) * 0.5])
        want = gold[f"labels_{K}_{H}"]
        assert np.array_equal(glue_ref.kmeans_predict(f, mu).numpy(), want)
        assert np.array_equal(glue_ref.kmeans_predict_f32(f, mu).numpy(), want)
        assert np.array_equal(want[: M // 2], idx[: M // 2].numpy())          # the near-centroid rows are unambiguous
    cluster = pytest.importorskip("sklearn.cluster")
    km = cluster.KMeans(n_clusters=K, n_init=1)
    km.cluster_centers_, km._n_threads, km.n_features_in_, km._n_features_out = mu.numpy().astype(np.float32), 1, H, K
    assert np.array_equal(km.predict(f.numpy()), want)


def test_ida_emb_longtensor_quirk():
    """I_da/scripts/inpainting.py:233: the d-vector goes through torch.LongTensor -> truncation toward zero."""
    from oracle import glue_ref
    emb = np.array([0.9, -0.9, 1.5, -2.7, 0.0, 3.0], dtype=np.float32)
    assert glue_ref.ida_emb_longtensor(emb).tolist() == torch.LongTensor(emb).tolist() == [0, 0, 1, -2, 0, 3]


def test_glue_oracle_vs_reference_golden(golden_dir):
    from oracle import glue_ref
    from oracle.params import make_codebook
    gold = np.load(f"{golden_dir}/glue_golden.npz")
    for T in (37, 100, 200):
        spec = torch.randn(2, 80, T, generator=torch.Generator().manual_seed(T))
        out = glue_ref.extend_mel(spec)
        assert out.shape[-1] == int(T * 441 / 256) == gold[f"extend_mel_{T}"].shape[-1]   # 100 -> 172, 200 -> 344
        assert max_abs(torch.from_numpy(gold[f"extend_mel_{T}"]), out) < 1e-6
        assert max_abs(out, glue_ref.extend_mel_explicit(spec)) < 3e-5
    for K in (100, 500):
        C = make_codebook(80, K, seed=77)
        vals = torch.randn(3, 10, 80, generator=torch.Generator().manual_seed(K))
        pred = glue_ref.cos_sim_argmax(vals, C).view(3, 10)
        assert np.array_equal(pred.numpy(), gold[f"cos_sim_pred_{K}"])
        mel = torch.randn(1, 80, 50, generator=torch.Generator().manual_seed(5))
        assert max_abs(torch.from_numpy(gold[f"paste_{K}"]), glue_ref.paste_centroids(mel, C, [pred[0]], [7])) == 0
    assert int(gold["upsample_pinned"][0]) == 1   # CodeGenerator._upsample was importable when the fixtures were made


def test_mel_oracle_vs_reference_golden(golden_dir):
    from oracle import mel_ref
    gold = np.load(f"{golden_dir}/mel_golden.npz")
    y = 0.3 * torch.randn(2, 22050, generator=torch.Generator().manual_seed(3)).clamp(-3, 3)
    m256 = mel_ref.mel_spectrogram(y, fmax=None)
    m441 = mel_ref.feature_mel(y)
    assert m256.shape == (2, 80, 86) and m441.shape == (2, 80, 50)     # 22 050 samples @441 -> 50 frames
    assert max_abs(torch.from_numpy(gold["mel_hop256"]), m256) < 1e-4
    assert max_abs(torch.from_numpy(gold["mel_hop441"]), m441) < 1e-4
    torchaudio = pytest.importorskip("torchaudio")
    ta = torchaudio.functional.melscale_fbanks(513, 0.0, 8000.0, 80, 22050, norm="slaney", mel_scale="slaney").T.numpy()
    assert np.abs(mel_ref.slaney_mel_filterbank(22050, 1024, 80, 0, 8000) - ta).max() < 1e-6


def test_znorm_and_int16_oracle():
    from oracle import glue_ref, hifigan_ref
    x = torch.randn(2, 1000) * 0.1 + 0.3
    z = glue_ref.processor_znorm(x)
    assert abs(float(z[0].mean())) < 1e-5 and abs(float(z[0].var(unbiased=False)) - 1) < 1e-3
    zl = glue_ref.processor_znorm(x, torch.tensor([1000, 600]))
    assert torch.all(zl[1, 600:] == 0) and abs(float(zl[1, :600].mean())) < 1e-5
    assert max_abs(glue_ref.fairseq_layer_norm(x[0]), torch.nn.functional.layer_norm(x[0], x[0].shape)) == 0
    y = torch.tensor([[[0.5, -0.5, 0.99999, -1.0, 3.0e-5]]])
    assert hifigan_ref.to_int16(y).tolist() == [16384, -16384, 32767, -32768, 0]


def test_feature_front_end_oracle():
    """predict.py:99-104 restated: the mask is applied before the peak normalisation and silent input stays silent."""
    from oracle import glue_ref, mel_ref
    rng = np.random.default_rng(0)
    w = (0.3 * rng.standard_normal(22050)).astype(np.float32)
    lo, hi = glue_ref.iea_mask_indices(0.4, 0.6)["zero22"]
    assert (lo, hi) == (int(0.4 * 16000) * 22050 // 16000, int(0.6 * 16000) * 22050 // 16000)
    n = mel_ref.peak_normalize(w)
    assert np.isclose(np.abs(n).max(), 1.0) and np.array_equal(mel_ref.peak_normalize(np.zeros(8, np.float32)), np.zeros(8, np.float32))
    m = mel_ref.masked_feature_mel(w, lo, hi)
    assert m.shape == (1, 80, 22050 // 441)
    # frames whose 1024-sample window lies entirely inside the zeroed range are log(1e-5 clamp) silence
    f = (lo + 1024 + 312) // 441 + 1
    assert torch.allclose(m[0, :, f], torch.full((80,), float(np.log(1e-5))), atol=1e-3)
