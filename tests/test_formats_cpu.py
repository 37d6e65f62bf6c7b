"""CPU (-m "not gpu"): SURVEY 8f rows 2-4 - oracles against their golden vectors, and the host-side logic of the product
(poly-phase filter design, RIFF/WAVE reader / writer, F0Quantizer state-dict surface).  No kernel runs here."""
import numpy as np
import pytest
import torch


@pytest.fixture(scope="module")
def sib():
    import speech_inpainting_b200 as m
    return m


def test_resample_oracle_vs_torchaudio_golden(golden_dir):
    from oracle import resample_ref as R
    g = np.load(f"{golden_dir}/resample_golden.npz")
    for o, n in ((16000, 22050), (22050, 16000), (48000, 16000), (24000, 22050)):
        y = R.resample(g["x"], o, n)
        assert y.shape == g[f"y_{o}_{n}"].shape == (R.out_length(12000, o, n),)
        assert np.abs(y - g[f"y_{o}_{n}"]).max() < 2e-7
    assert np.array_equal(R.resample(g["x"], 16000, 16000), g["x"].astype(np.float64))
    assert R.out_length(64000, 16000, 22050) == 88200 and R.out_length(1, 16000, 22050) == 2


def _polyphase_on_cpu(sib, x, o, n):
    """What sib_resample computes, written with numpy from the product's own filter table."""
    f, up, down, first = sib.resample_filter(o, n)
    n_out = sib.audio.resampled_length(len(x), o, n)
    idx = np.arange(n_out)
    q, r = (idx * down) // up, idx % up
    pad = f.shape[0] + abs(first)
    xp = np.concatenate([np.zeros(pad), x.astype(np.float64), np.zeros(pad + down)])
    y = np.zeros(n_out)
    for j in range(f.shape[0]):
        y += f[j, r].astype(np.float64) * xp[q + first + j + pad]
    return y


def test_product_filter_design_matches_oracle(sib, golden_dir):
    from oracle import resample_ref as R
    x = np.load(f"{golden_dir}/resample_golden.npz")["x"]
    for o, n in ((16000, 22050), (22050, 16000), (48000, 16000), (44100, 22050)):
        f, up, down, first = sib.resample_filter(o, n)
        assert f.dtype == np.float32 and f.shape == (2 * (-first + 1), up) and up * o == down * n
        assert np.abs(_polyphase_on_cpu(sib, x, o, n) - R.resample(x, o, n)).max() < 1e-6
        # DC gain of every phase is 1 (to the stop-band ripple): a constant stays a constant
        assert np.abs(f.astype(np.float64).sum(0) - 1.0).max() < 2e-3 if n > o else True
    f, up, down, first = sib.resample_filter(16000, 16000)
    assert (f.tolist(), up, down, first) == ([[1.0]], 1, 1, 0)
    assert (up, down) == (1, 1) and sib.resample_filter(16000, 22050)[1:3] == (441, 320)
    assert [sib.audio.resampled_length(v, 16000, 22050) for v in (0, 1, 320, 64000, 63999)] == [0, 2, 441, 88200, 88199]
    with pytest.raises(sib.SibError):
        sib.resample(torch.zeros(4), 16000, 22050)   # CPU tensor: no fallback


def test_wav_reader_writer_roundtrip(sib, tmp_path):
    from scipy.io import wavfile
    rng = np.random.default_rng(0)
    mono = rng.integers(-32768, 32767, 4001, dtype=np.int16)
    stereo = rng.integers(-32768, 32767, (300, 2), dtype=np.int16)
    sib.write_wav(tmp_path / "m.wav", mono, 22050)
    sib.write_wav(tmp_path / "s.wav", torch.from_numpy(stereo), 16000)
    sr, back = wavfile.read(tmp_path / "m.wav")            # an independent reader accepts what we write
    assert sr == 22050 and np.array_equal(back, mono)
    pcm, sr = sib.read_wav(tmp_path / "m.wav")
    assert sr == 22050 and pcm.shape == (4001, 1) and np.array_equal(pcm[:, 0], mono)
    pcm, sr = sib.read_wav(tmp_path / "s.wav")
    assert sr == 16000 and np.array_equal(pcm, stereo)
    wavfile.write(tmp_path / "w.wav", 16000, mono)          # ... and we read what an independent writer produces
    assert np.array_equal(sib.read_wav(tmp_path / "w.wav")[0][:, 0], mono)
    wavfile.write(tmp_path / "f.wav", 16000, mono.astype(np.float32) / 32768)
    with pytest.raises(sib.SibError):
        sib.read_wav(tmp_path / "f.wav")                    # float WAV: refused loudly, not converted silently
    (tmp_path / "junk.wav").write_bytes(b"not a wave file at all")
    with pytest.raises(sib.SibError):
        sib.read_wav(tmp_path / "junk.wav")
    with pytest.raises(sib.SibError):
        sib.write_wav(tmp_path / "x.wav", mono.astype(np.float32), 16000)
    with pytest.raises(sib.SibError):
        sib.load_wav_batch([tmp_path / "m.wav"])            # needs a device


def test_f0vq_oracle_vs_reference_golden(golden_dir):
    from oracle import f0vq_ref
    g = np.load(f"{golden_dir}/f0vq_golden.npz")
    sd = f0vq_ref.make_params(seed=1234)
    for B, L in ((2, 784), (1, 160)):
        f0 = torch.randn(B, 1, L, generator=torch.Generator().manual_seed(L))
        h = f0vq_ref.encoder_forward(sd, f0)
        assert h.shape == (B, 128, L // 16) and float((h - torch.from_numpy(g[f"h_{L}"])).abs().max()) < 1e-5
        assert np.array_equal(f0vq_ref.f0_to_bins(sd, f0).numpy(), g[f"z_{L}"])
    assert len(set(g["z_784"].ravel().tolist())) > 5    # the fixture exercises more than a couple of bins


def test_f0_quantizer_state_dict_surface(sib):
    from oracle import f0vq_ref
    q = sib.F0Quantizer(f0vq_ref.F0_QUANTIZER)
    sd = f0vq_ref.make_params()
    assert set(q._expected_keys()) == set(sd) and q.hop == 16 and len(q._conv_names()) == 4 * 9 + 1
    q.load_state_dict(dict(sd, **{"decoder.level_blocks.0.model.0.weight": torch.zeros(1)}))   # decoder.* ignored
    with pytest.raises(RuntimeError):
        q.load_state_dict({k: v for k, v in sd.items() if not k.endswith(".0.bias")})
    with pytest.raises(sib.SibError):
        q.encode(torch.zeros(1, 1, 160))                    # not on a CUDA device: no fallback
    with pytest.raises(sib.SibError):
        sib.F0Quantizer(dict(f0vq_ref.F0_QUANTIZER, f0_encoder_params=dict(f0vq_ref.F0_QUANTIZER["f0_encoder_params"], levels=2)))
    h = sib.AttrDict(upsample_rates=[5, 4], upsample_kernel_sizes=[11, 8], upsample_initial_channel=32,
                     resblock_kernel_sizes=[3], resblock_dilation_sizes=[[1, 3, 5]], model_in_dim=384, resblock="1",
                     num_embeddings=10, embedding_dim=128, f0_quantizer=f0vq_ref.F0_QUANTIZER)
    gen = sib.CodeGenerator(h)
    assert isinstance(gen.fo_vqvae, sib.F0Quantizer)
    gen.load_f0_quantizer(sd)
    assert set(gen.fo_vqvae.state_dict()) == set(sd)


def test_metrics_oracle_vs_reference_golden(golden_dir):
    from oracle import metrics_ref
    want = np.load(f"{golden_dir}/metrics_golden.npz")["sisdr"]
    g = torch.Generator().manual_seed(21)
    ref = 0.3 * torch.randn(3, 22050, generator=g)
    est = ref * torch.tensor([[1.0], [0.5], [2.0]]) + torch.tensor([[0.01], [0.1], [0.5]]) * torch.randn(3, 22050, generator=g)
    for b in range(3):
        got = metrics_ref.sisdr(est[b].numpy().astype(np.float64), ref[b].numpy().astype(np.float64))
        assert abs(got - want[b]) < 1e-9
    # scale invariance, the property the metric is named after
    e, r = est[1].numpy().astype(np.float64), ref[1].numpy().astype(np.float64)
    assert abs(metrics_ref.sisdr(3.7 * e, r) - metrics_ref.sisdr(e, r)) < 1e-9
