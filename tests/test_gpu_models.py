"""-m gpu: module-level parity - CUDA path vs the CPU oracle on the same seeded inputs / weights, and vs the
golden vectors produced by the real reference (tests/golden, oracle/make_golden.py).

Tolerances (fp32 path): waveform SNR >= 40 dB is north_star's bar; measured margins are far larger, so the
asserts use 60 dB to catch regressions early.  Integer outputs (labels, mask indices, lengths) are exact."""
import numpy as np
import pytest
import torch

from util import max_abs, snr_db

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sib():
    import speech_inpainting_b200 as m
    m._load_lib()
    return m


def _hub_cfg(sib, ocfg):
    return sib.HubertConfig.from_any(ocfg)


def _hubert_pair(sib, name):
    from oracle.params import HubertCfg, make_hubert_params
    ocfg = {"tiny_group": HubertCfg.tiny(False), "tiny_layer": HubertCfg.tiny(True), "base": HubertCfg.base(),
            "large": HubertCfg.large()}[name]
    params = make_hubert_params(ocfg, seed=1234)
    model = sib.HubertModel(_hub_cfg(sib, ocfg)).to("cuda")
    model.load_state_dict(params)
    return ocfg, params, model.eval()


@pytest.mark.parametrize("name,B,N", [("tiny_group", 2, 3000), ("tiny_layer", 2, 3000), ("base", 1, 4000), ("large", 1, 2400)])
def test_hubert_vs_reference_golden(sib, golden_dir, name, B, N):
    gold = np.load(f"{golden_dir}/hubert_golden.npz")
    ocfg, params, model = _hubert_pair(sib, name)
    x = 0.1 * torch.randn(B, N, generator=torch.Generator().manual_seed(99))
    y = model(x.cuda()).last_hidden_state.cpu()
    ref = torch.from_numpy(gold[name + "_out"])   # transformers.HubertModel output
    assert y.shape == ref.shape
    assert max_abs(ref, y) < 2e-4 and snr_db(ref, y) > 80
    if name.startswith("tiny"):
        am = torch.ones(B, N, dtype=torch.long)
        am[1, N - 900:] = 0
        xp = x.clone()
        xp[1, N - 900:] = 0
        y = model(xp.cuda(), attention_mask=am.cuda()).last_hidden_state.cpu()
        ref = torch.from_numpy(gold[name + "_padded_out"])
        assert max_abs(ref, y) < 2e-4 and snr_db(ref, y) > 80


def test_hubert_per_layer_bisect_and_extract_features(sib):
    """Per-stage taps of the oracle vs the CUDA plan's intermediate result at a truncated depth."""
    from oracle import hubert_ref
    ocfg, params, model = _hubert_pair(sib, "tiny_group")
    x = 0.1 * torch.randn(2, 5000, generator=torch.Generator().manual_seed(5))
    for layer in (1, 2):
        ref = hubert_ref.hubert_forward(params, ocfg, x, output_layer=layer)
        feat, pm = model.extract_features(x.cuda(), padding_mask=None, mask=False, output_layer=layer)
        assert pm is None and max_abs(ref, feat.cpu()) < 1e-4
    ref = hubert_ref.get_feats(params, ocfg, x[0].numpy(), normalize=True, layer=-1)
    xn = torch.empty(1, 5000, device="cuda")
    sib.ops.znorm(x[:1].cuda(), xn, None, 1e-5)
    feat, _ = model.extract_features(xn, output_layer=-1)
    assert max_abs(ref, feat[0].cpu()) < 2e-4


def test_hubert_base_2s_vs_oracle(sib):
    """config #1 encoder shape: 1 x 2 s -> [1, 99, 768]."""
    from oracle import hubert_ref
    ocfg, params, model = _hubert_pair(sib, "base")
    x = 0.1 * torch.randn(1, 32000, generator=torch.Generator().manual_seed(1234))
    ref = hubert_ref.hubert_forward(params, ocfg, x)
    y = model(x.cuda()).last_hidden_state.cpu()
    assert y.shape == (1, 99, 768)
    assert max_abs(ref, y) < 3e-4 and snr_db(ref, y) > 80


def test_custom_model_state_dict_surface(sib):
    from oracle import hubert_ref
    from oracle.params import HubertCfg, make_head_params, make_hubert_params
    ocfg = HubertCfg.tiny(False)
    sd = make_hubert_params(ocfg, 1234, prefix="base_model.")
    sd.update(make_head_params(ocfg.hidden_size, 80))
    m = sib.CustomModel(codebook_dim=80, type="base", load_pretrained=False, config=_hub_cfg(sib, ocfg)).to("cuda")
    with pytest.raises(RuntimeError):
        m.load_state_dict({k: v for k, v in sd.items() if "q_proj" not in k})
    # old-style weight-norm names are accepted too
    old = {k.replace("parametrizations.weight.original0", "weight_g").replace("parametrizations.weight.original1", "weight_v"): v
           for k, v in sd.items()}
    m.load_state_dict(old)
    m.eval()
    x = 0.1 * torch.randn(2, 4000, generator=torch.Generator().manual_seed(3))
    ref = hubert_ref.custom_model_forward(sd, ocfg, x)
    y = m(x.cuda(), None).cpu()
    assert y.shape == ref.shape == (2, ocfg.feat_lengths(4000), 80)
    assert max_abs(ref, y) < 2e-4


GEN_CASES = [("v1_unit", "v1", 2, 6, "unit"), ("v1_ref", "v1", 1, 5, "reference"), ("tiny_unit", "tiny", 2, 9, "unit"),
             ("ida_unit", "ida", 1, 4, "unit"), ("ida_tiny", "ida_tiny", 2, 8, "unit"),
             # config_v2.json (ResBlock1, 128 channels) and config_v3.json (ResBlock2, models.py:52-73)
             ("v2_unit", "v2", 2, 7, "unit"), ("v3_unit", "v3", 2, 11, "unit"), ("v3_ref", "v3", 1, 6, "reference")]


def _gen_cfg(kind):
    from oracle.params import HifiCfg
    return {"v1": HifiCfg.v1(), "v2": HifiCfg.v2(), "v3": HifiCfg.v3(), "tiny": HifiCfg.tiny(), "ida": HifiCfg.ida(),
            "ida_tiny": HifiCfg.tiny(True)}[kind]


@pytest.mark.parametrize("name,kind,B,T,init", GEN_CASES)
def test_generator_vs_reference_golden(sib, golden_dir, name, kind, B, T, init):
    from oracle.params import fold_weight_norm, make_generator_params
    gold = torch.from_numpy(np.load(f"{golden_dir}/hifigan_golden.npz")[name + "_out"])
    cfg = _gen_cfg(kind)
    params = {k: v for k, v in make_generator_params(cfg, 1234, init).items() if not k.startswith("emb_")}
    gen = sib.Generator(sib.AttrDict(cfg.as_attrdict())).to("cuda")
    gen.load_state_dict(params)       # with weight_g / weight_v, as the reference checkpoints
    gen.eval()
    assert set(gen.state_dict()) == set(params)
    gen.remove_weight_norm()          # folds on the device and leaves the reference's post-removal key set
    assert set(gen.state_dict()) == set(fold_weight_norm(params)) and len(gen.state_dict()) == 2 * len(params) // 3
    x = torch.randn(B, cfg.model_in_dim, T, generator=torch.Generator().manual_seed(7))
    y = gen(x.cuda()).cpu()
    assert y.shape == gold.shape == (B, 1, T * cfg.total_upsample)
    assert snr_db(gold, y) > 60, snr_db(gold, y)
    assert max_abs(gold, y) < 1e-4 * max(1.0, float(gold.abs().max()))
    # folded ("remove_weight_norm"-ed) checkpoints load too and give the same answer
    gen2 = sib.Generator(sib.AttrDict(cfg.as_attrdict())).to("cuda")
    gen2.load_state_dict(fold_weight_norm(params))
    assert max_abs(y, gen2(x.cuda()).cpu()) < 1e-5


def test_generator_longer_input_vs_oracle(sib):
    from oracle import hifigan_ref
    from oracle.params import HifiCfg, make_generator_params
    cfg = HifiCfg.v1()
    params = make_generator_params(cfg, 1234, "unit")
    gen = sib.Generator(sib.AttrDict(cfg.as_attrdict())).to("cuda")
    gen.load_state_dict(params)
    x = torch.randn(2, 80, 43, generator=torch.Generator().manual_seed(11))
    ref = hifigan_ref.generator_forward(params, cfg, x)
    y = gen(x.cuda()).cpu()
    assert snr_db(ref, y) > 60 and max_abs(ref, y) < 1e-4


def test_code_generator_vs_oracle(sib):
    from oracle import hifigan_ref
    from oracle.params import HifiCfg, make_generator_params
    cfg = HifiCfg.tiny(True)
    params = make_generator_params(cfg, 1234, "unit")
    h = sib.AttrDict(cfg.as_attrdict())
    gen = sib.CodeGenerator(h).to("cuda")
    gen.load_state_dict(params)
    g = torch.Generator().manual_seed(2)
    code = torch.randint(0, cfg.num_embeddings, (2, 12), generator=g)
    zp = torch.randint(0, 20, (2, 3), generator=g)
    emb = torch.randn(2, cfg.embedding_dim, generator=g)
    ref = hifigan_ref.code_generator_forward(params, cfg, code, zp, emb)
    y = gen(code=code.cuda(), f0_code=zp.cuda(), emb=emb.cuda(), spkr=torch.zeros(2, 1, dtype=torch.long)).cpu()
    assert y.shape == ref.shape == (2, 1, 12 * cfg.total_upsample)
    assert snr_db(ref, y) > 60
    with pytest.raises(sib.SibError):
        gen(code=code.cuda(), f0=torch.zeros(2, 1, 48), emb=emb.cuda(), spkr=None)


def test_f0_quantizer_vs_oracle_and_reference_golden(sib, golden_dir):
    """frozen f0 VQ-VAE encoder + quantiser (I_da/src/model.py:148-152) - integer output, must be exact."""
    from oracle import f0vq_ref
    g = np.load(f"{golden_dir}/f0vq_golden.npz")
    sd = f0vq_ref.make_params(seed=1234)
    q = sib.F0Quantizer(f0vq_ref.F0_QUANTIZER).to("cuda")
    q.load_state_dict(sd)
    for B, L in ((2, 784), (1, 160)):
        f0 = torch.randn(B, 1, L, generator=torch.Generator().manual_seed(L))
        h = q.encode_features(f0.cuda()).cpu()                       # frame-major [B, L/16, 128]
        assert max_abs(torch.from_numpy(g[f"h_{L}"]).transpose(1, 2), h) < 2e-5
        z = q.encode(f0.cuda()).cpu()
        assert z.dtype == torch.int64 and np.array_equal(z.numpy(), g[f"z_{L}"])
    # a batch the oracle handles in seconds, odd batch size, other seed: bins bit-equal, and batch-invariant
    sd2 = f0vq_ref.make_params(seed=7, scale=1.5)
    q.load_state_dict(sd2)
    f0 = torch.randn(5, 1, 3200, generator=torch.Generator().manual_seed(1))
    z = q.encode(f0.cuda()).cpu()
    want = f0vq_ref.f0_to_bins(sd2, f0)
    assert z.shape == (5, 200) and float((z == want).float().mean()) >= 0.999
    assert torch.equal(q.encode(f0[3:4].cuda()).cpu(), z[3:4])
    with pytest.raises(sib.SibError):
        q.encode(torch.zeros(1, 1, 8).cuda())                         # shorter than one bin


def test_code_generator_quantises_f0_itself(sib):
    """CodeGenerator.forward(code=, f0=, emb=, spkr=) - the reference call (model.py:121-189) without precomputed bins."""
    from oracle import f0vq_ref, hifigan_ref
    from oracle.params import HifiCfg, make_generator_params
    cfg = HifiCfg.tiny(True)
    params = make_generator_params(cfg, 1234, "unit")
    fsd = f0vq_ref.make_params(seed=5)
    h = sib.AttrDict(dict(cfg.as_attrdict(), f0_quantizer=f0vq_ref.F0_QUANTIZER))
    gen = sib.CodeGenerator(h).to("cuda")
    gen.load_state_dict(dict(params, **{"fo_vqvae." + k: v for k, v in fsd.items()}))   # checkpoint layout of model.py:63-71
    g = torch.Generator().manual_seed(2)
    code = torch.randint(0, cfg.num_embeddings, (2, 12), generator=g)
    f0 = torch.randn(2, 1, 48, generator=g)
    emb = torch.randn(2, cfg.embedding_dim, generator=g)
    zp = f0vq_ref.f0_to_bins(fsd, f0)
    assert zp.shape == (2, 3)
    ref = hifigan_ref.code_generator_forward(params, cfg, code, zp, emb)
    y = gen(code=code.cuda(), f0=f0.cuda(), emb=emb.cuda(), spkr=torch.zeros(2, 1, dtype=torch.long)).cpu()
    assert y.shape == ref.shape and snr_db(ref, y) > 60
    y2 = gen(code=code.cuda(), f0_code=zp.cuda(), emb=emb.cuda()).cpu()
    assert torch.equal(y, y2)


def _iea_oracle(params_h, ocfg, gparams, gcfg, C, wave, mel, pos, ln):
    from oracle import glue_ref, hifigan_ref, hubert_ref
    x = wave.clone()
    for b in range(x.shape[0]):
        lo, hi = glue_ref.iea_zero_range_from_frames(pos[b], ln[b])
        x[b] = torch.from_numpy(glue_ref.apply_zero_range(x[b].numpy(), lo, hi))
    xn = glue_ref.processor_znorm(x)
    out = hubert_ref.custom_model_forward(params_h, ocfg, xn, None)
    vals = glue_ref.gather_mask_frames(out, pos, ln)
    labels = [glue_ref.cos_sim_argmax(v, C) for v in vals]
    mel2 = glue_ref.paste_centroids(mel, C, labels, pos)
    feats = glue_ref.extend_mel(mel2)
    return hifigan_ref.generator_forward(gparams, gcfg, feats), torch.cat(labels), mel2


def test_informed_inpainting_config1_end_to_end(sib):
    """BASELINE config #1: HuBERT-base + head + HiFi-GAN V1, 1 x 2 s, 200 ms mask at 0.9 s, fp32."""
    from oracle.params import HifiCfg, HubertCfg, make_codebook, make_generator_params, make_head_params, make_hubert_params
    ocfg, gcfg = HubertCfg.base(), HifiCfg.v1()
    sd = make_hubert_params(ocfg, 1234, prefix="base_model.")
    sd.update(make_head_params(768, 80))
    gparams = {k: v for k, v in make_generator_params(gcfg, 1234, "unit").items()}
    C = make_codebook(80, 100)
    g = torch.Generator().manual_seed(1234)
    wave = 0.1 * torch.randn(1, 32000, generator=g)
    mel = torch.randn(1, 80, 100, generator=g)
    mi = sib.iea_mask_indices(0.9, 1.1)
    assert (mi["mask_pos"], mi["mask_len"], mi["zero16"]) == (45, 10, (14480, 17599))   # SURVEY 8d cfg 1
    ref_wave, ref_labels, ref_mel = _iea_oracle(sd, ocfg, gparams, gcfg, C, wave, mel, [45], [10])
    model = sib.CustomModel(80, "base", False, config=_hub_cfg(sib, ocfg)).to("cuda")
    model.load_state_dict(sd)
    gen = sib.Generator(sib.AttrDict(gcfg.as_attrdict())).to("cuda")
    gen.load_state_dict(gparams)
    gen.remove_weight_norm()
    pipe = sib.InformedInpainter(model.eval(), gen.eval(), C)
    res = pipe(wave, mel, [45], [10], return_int16=True)
    assert res.wave.shape == ref_wave.shape == (1, 1, 172 * 256)
    assert torch.equal(ref_labels, res.labels.cpu())            # integer outputs exact
    assert max_abs(ref_mel, res.mel.cpu()) < 1e-6
    s = snr_db(ref_wave, res.wave.cpu())
    assert s > 60, s                                            # north_star bar: >= 40 dB
    from oracle import hifigan_ref
    i16 = hifigan_ref.to_int16(ref_wave)
    assert np.abs(i16.astype(np.int32) - res.int16.cpu().numpy().squeeze().astype(np.int32)).max() <= 1


def test_informed_inpainting_ragged_masks_tiny(sib):
    """config #4 style: variable mask lengths per utterance (incl. L=0 and a mask touching the last frame)."""
    from oracle.params import HifiCfg, HubertCfg, make_codebook, make_generator_params, make_head_params, make_hubert_params
    ocfg, gcfg = HubertCfg.tiny(True), HifiCfg.tiny()
    sd = make_hubert_params(ocfg, 1, prefix="base_model.")
    sd.update(make_head_params(ocfg.hidden_size, 80))
    gparams = make_generator_params(gcfg, 2, "unit")
    C = make_codebook(80, 500)
    g = torch.Generator().manual_seed(7)
    B, N = 4, 16000
    T = ocfg.feat_lengths(N)
    wave, mel = 0.1 * torch.randn(B, N, generator=g), torch.randn(B, 80, 50, generator=g)
    pos, ln = [3, 20, T - 5, 10], [1, 15, 5, 0]
    ref_wave, ref_labels, _ = _iea_oracle(sd, ocfg, gparams, gcfg, C, wave, mel, pos, ln)
    model = sib.CustomModel(80, "large", False, config=_hub_cfg(sib, ocfg)).to("cuda")
    model.load_state_dict(sd)
    gen = sib.Generator(sib.AttrDict(gcfg.as_attrdict())).to("cuda")
    gen.load_state_dict(gparams)
    res = sib.InformedInpainter(model, gen, C)(wave, mel, pos, ln)
    assert torch.equal(ref_labels, res.labels.cpu())
    assert snr_db(ref_wave, res.wave.cpu()) > 60
    with pytest.raises(sib.SibError):
        sib.InformedInpainter(model, gen, C)(wave, mel, [T - 2] * B, [5] * B)   # mask beyond the last frame


def test_blind_inpainting_tiny(sib):
    """I_da path (config #3 style) on tiny shapes: units exact, splice exact, waveform SNR."""
    from oracle import glue_ref, hifigan_ref, hubert_ref
    from oracle.params import HifiCfg, HubertCfg, make_generator_params, make_hubert_params
    ocfg, gcfg = HubertCfg.tiny(False), HifiCfg.tiny(True)
    hp = make_hubert_params(ocfg, 3)
    gp = make_generator_params(gcfg, 4, "unit")
    g = torch.Generator().manual_seed(52)
    B, N, mask = 2, 32000, 6400
    wave = 0.1 * torch.randn(B, N, generator=g)
    mu = torch.randn(gcfg.num_embeddings, ocfg.hidden_size, generator=g) * 0.5
    T = ocfg.feat_lengths(N)
    n_code = glue_ref.ida_matched_frames(N, T, 4 * T)
    zp = torch.randint(0, 20, (B, T // 4 + 1), generator=g)
    emb = torch.randn(B, gcfg.embedding_dim, generator=g)
    hub = sib.HubertModel(_hub_cfg(sib, ocfg)).to("cuda")
    hub.load_state_dict(hp)
    # tiny generator upsamples by 40, not 320 - only the code path is under test here
    gen = sib.CodeGenerator(sib.AttrDict(gcfg.as_attrdict())).to("cuda")
    gen.load_state_dict(gp)
    res = sib.BlindInpainter(hub, gen, mu, layer=-1, normalize=False)(wave, mask, zp, emb, informed=True)
    for b in range(B):
        y_inp, fs = glue_ref.ida_mask(wave[b].numpy(), mask)
        assert np.array_equal(y_inp.astype(np.float32), res.audio_mask[b].cpu().numpy())
        code = glue_ref.kmeans_predict(hubert_ref.get_feats(hp, ocfg, wave[b].numpy(), False, -1), mu).numpy()
        code_inp = glue_ref.kmeans_predict(hubert_ref.get_feats(hp, ocfg, y_inp.astype(np.float32), False, -1), mu).numpy()
        code_inp = glue_ref.ida_splice_codes(code, code_inp, fs, mask)
        assert np.array_equal(code[:n_code], res.code[b].cpu().numpy())
        assert np.array_equal(code_inp[:n_code], res.code_inpainting[b].cpu().numpy())
        # inpainting.py:233: the d-vector reaches the generator as torch.LongTensor(emb) (truncated toward zero)
        emb_l = glue_ref.ida_emb_longtensor(emb[b:b + 1])
        ref = hifigan_ref.code_generator_forward(gp, gcfg, torch.from_numpy(code_inp[:n_code])[None], zp[b:b + 1, : n_code // 4], emb_l)
        assert snr_db(ref, res.audio_inp[b:b + 1].cpu()) > 60
    raw = sib.BlindInpainter(hub, gen, mu, layer=-1, normalize=False, emb_as_long=False)(wave, mask, zp, emb, informed=True)
    assert not torch.equal(raw.audio_inp, res.audio_inp)        # the raw float d-vector is a different (non-reference) input


def test_mask_golden_on_device(sib, golden_dir):
    """a1: replay predict.py:133 on the device against the reference's own orig/masked wav pair."""
    import json
    gold = json.load(open(f"{golden_dir}/mask_golden.json"))
    lo, hi = sib.iea_zero_range(gold["mask_pos"], gold["mask_len"])
    assert [lo, hi] == gold["zero_range"] and hi - lo == gold["n_zeroed"] == 6319
    n = gold["n_samples"]
    wave = torch.arange(1, n + 1, dtype=torch.float32)[None].cuda()
    i32 = lambda v: torch.tensor(v, dtype=torch.int32).cuda()
    sib.ops.zero_ranges(wave, i32([lo]), i32([hi]))
    z = (wave[0] == 0).nonzero().flatten().cpu()
    assert int(z[0]) == gold["zero_range"][0] and int(z[-1]) + 1 == gold["zero_range"][1] and len(z) == 6319
    # edge windows of the real files: orig -> masked
    for side, idx in (("lo", lo), ("hi", hi)):
        o = torch.tensor(gold[f"orig_{side}_window"], dtype=torch.float32)[None].cuda()
        rel_lo, rel_hi = (4, 8 + 100) if side == "lo" else (-100, 4)
        sib.ops.zero_ranges(o, i32([max(rel_lo, 0)]), i32([min(rel_hi, 8)]))
        assert o[0].cpu().tolist() == [float(v) for v in gold[f"masked_{side}_window"]]


def test_plan_replay_and_cuda_graph(sib):
    from oracle.params import HifiCfg, make_generator_params
    cfg = HifiCfg.tiny()
    gen = sib.Generator(sib.AttrDict(cfg.as_attrdict())).to("cuda")
    gen.load_state_dict(make_generator_params(cfg, 1, "unit"))
    x = torch.randn(2, 80, 20, generator=torch.Generator().manual_seed(1)).cuda()
    y1 = gen(x)
    n0 = sib.ops.launch_count()
    y2 = gen(x)
    assert sib.ops.launch_count() - n0 == len(gen._plans[(2, 20, False)].plan)
    assert torch.equal(y1, y2)
    gen2 = sib.Generator(sib.AttrDict(cfg.as_attrdict())).to("cuda")
    gen2.use_cuda_graph = True
    gen2.load_state_dict(make_generator_params(cfg, 1, "unit"))
    assert torch.equal(y1, gen2(x)) and torch.equal(y1, gen2(x))


def test_informed_inpainting_from_wave22(sib):
    """The pipeline fed with the 22.05 kHz rendition instead of a precomputed mel equals the pipeline
    fed with the oracle's mel of the same masked, normalised signal."""
    from oracle import glue_ref, mel_ref
    from oracle.params import HifiCfg, HubertCfg, make_codebook, make_generator_params, make_head_params, make_hubert_params
    ocfg, gcfg = HubertCfg.tiny(False), HifiCfg.tiny()
    sd = make_hubert_params(ocfg, 1234, prefix="base_model.")
    sd.update(make_head_params(ocfg.hidden_size, 80))
    gp = make_generator_params(gcfg, 1234, "unit")
    C = make_codebook(80, 100)
    g = torch.Generator().manual_seed(5)
    B, N = 2, 32000
    wave16 = 0.1 * torch.randn(B, N, generator=g)
    wave22 = (0.1 * torch.randn(B, N * 22050 // 16000, generator=g)).clamp(-1, 1)
    idx = [glue_ref.iea_mask_indices(0.9, 1.1), glue_ref.iea_mask_indices(0.2, 0.6)]
    pos, ln = [i["mask_pos"] for i in idx], [i["mask_len"] for i in idx]
    model = sib.CustomModel(80, "base", False, config=_hub_cfg(sib, ocfg)).to("cuda")
    model.load_state_dict(sd)
    gen = sib.Generator(sib.AttrDict(gcfg.as_attrdict())).to("cuda")
    gen.load_state_dict(gp)
    pipe = sib.InformedInpainter(model, gen, C)
    res = pipe(wave16, None, pos, ln, wave22=wave22, zero22=[i["zero22"] for i in idx])
    mel_ref_ = torch.cat([mel_ref.masked_feature_mel(wave22[b].numpy(), *idx[b]["zero22"]) for b in range(B)])
    res2 = pipe(wave16, mel_ref_, pos, ln)
    assert torch.equal(res.labels, res2.labels)
    assert snr_db(res2.wave.cpu(), res.wave.cpu()) > 60
    with pytest.raises(sib.SibError, match="mel or wave22"):
        pipe(wave16, None, pos, ln)


def test_predict_files_wav_to_wav(sib, tmp_path):
    """I_ea/predict.py:66-207 from files to files (SURVEY 8f rows 1 + 3): wav -> device resampler -> feature mel ->
    informed pipeline -> inpainted.wav; equals the pipeline fed with the oracle's resampled signals."""
    from oracle import glue_ref, resample_ref as R
    from oracle.params import HifiCfg, HubertCfg, make_codebook, make_generator_params, make_head_params, make_hubert_params
    ocfg, gcfg = HubertCfg.tiny(False), HifiCfg.tiny()
    sd = make_hubert_params(ocfg, 1234, prefix="base_model.")
    sd.update(make_head_params(ocfg.hidden_size, 80))
    model = sib.CustomModel(80, "base", False, config=_hub_cfg(sib, ocfg)).to("cuda")
    model.load_state_dict(sd)
    gen = sib.Generator(sib.AttrDict(gcfg.as_attrdict())).to("cuda")
    gen.load_state_dict(make_generator_params(gcfg, 1234, "unit"))
    pipe = sib.InformedInpainter(model, gen, make_codebook(80, 100))
    rng = np.random.default_rng(11)
    t = np.arange(44100) / 22050.0
    files, pcms = [], []
    for i, n in enumerate((44100, 44100, 33075)):     # two utterances of 2 s and one of 1.5 s, 22.05 kHz PCM-16 on disk
        pcm = (3000 * np.sin(2 * np.pi * (180 + 40 * i) * t[:n]) + 800 * rng.standard_normal(n)).astype(np.int16)
        sib.write_wav(tmp_path / f"utt{i}.wav", pcm, 22050)
        files.append(tmp_path / f"utt{i}.wav")
        pcms.append(pcm)
    out = sib.predict_files(pipe, files, 0.5, 0.7, save_dir=tmp_path / "pred")
    idx = glue_ref.iea_mask_indices(0.5, 0.7)
    for i, pcm in enumerate(pcms):
        w22 = R.pcm16_to_float(pcm)
        w16 = R.resample(w22, 22050, 16000)
        ref = pipe(torch.from_numpy(w16.astype(np.float32))[None], None, idx["mask_pos"], idx["mask_len"],
                   wave22=torch.from_numpy(w22.astype(np.float32))[None], zero22=[idx["zero22"]], return_int16=True)
        assert torch.equal(out[i].labels, ref.labels.cpu())
        a, b = out[i].int16.float(), ref.int16.reshape(-1).cpu().float()
        assert a.shape == b.shape and float((a - b).abs().max()) <= 2          # resampler rounding (2e-6) -> <= 2 LSB
        back, sr = sib.read_wav(tmp_path / "pred" / f"utt{i}" / "inpainted.wav")
        assert sr == 22050 and np.array_equal(back[:, 0], out[i].int16.numpy())
        masked, sr = sib.read_wav(tmp_path / "pred" / f"utt{i}" / "masked.wav")
        lo, hi = idx["zero16"]
        assert sr == 16000 and not masked[lo:hi].any() and masked[:lo].any() and masked.shape[0] == len(w16)


def test_stream_pipeline_matches_direct_calls(sib):
    """InformedInpainter.stream (upload / compute / download overlapped over `depth` slots) returns, in order, exactly
    what one direct call per batch returns - including when slots and host buffers are recycled."""
    from oracle.params import HifiCfg, HubertCfg, make_codebook, make_generator_params, make_head_params, make_hubert_params
    ocfg, gcfg = HubertCfg.tiny(False), HifiCfg.tiny()
    sd = make_hubert_params(ocfg, 1234, prefix="base_model.")
    sd.update(make_head_params(ocfg.hidden_size, 80))
    model = sib.CustomModel(80, "base", False, config=_hub_cfg(sib, ocfg)).to("cuda")
    model.load_state_dict(sd)
    gen = sib.Generator(sib.AttrDict(gcfg.as_attrdict())).to("cuda")
    gen.load_state_dict(make_generator_params(gcfg, 1234, "unit"))
    pipe = sib.InformedInpainter(model, gen, make_codebook(80, 100))
    g = torch.Generator().manual_seed(8)
    batches = []
    for i in range(5):
        batches.append({"wave16": (0.1 * torch.randn(3, 16000, generator=g)).pin_memory(),
                        "mel": torch.randn(3, 80, 50, generator=g).pin_memory(),
                        "mask_pos": [5 + i, 20, 31], "mask_len": [10, 3 + i, 0]})
    direct = [pipe(b["wave16"], b["mel"], b["mask_pos"], b["mask_len"], return_int16=True) for b in batches]
    direct = [(r.int16.cpu().clone(), r.labels.cpu().clone()) for r in direct]
    for depth in (1, 2, 3):
        n = 0
        for out, (pcm, lab) in zip(pipe.stream(iter(batches), depth=depth), direct):
            assert torch.equal(out.int16, pcm) and torch.equal(out.labels, lab)   # checked before the buffers are recycled
            n += 1
        assert n == len(batches)


def test_blind_inpainting_with_continuous_f0(sib):
    """BlindInpainter(f0=...) - the reference passes the raw f0 track (inpainting.py:231) and CodeGenerator quantises it
    (model.py:148-152): same result as handing over the oracle's bins, including the 1280-sample trim of the f0 series."""
    from oracle import f0vq_ref, glue_ref
    from oracle.params import HifiCfg, HubertCfg, make_generator_params, make_hubert_params
    ocfg, gcfg = HubertCfg.tiny(False), HifiCfg.tiny(True)
    hp, gp, fsd = make_hubert_params(ocfg, 3), make_generator_params(gcfg, 4, "unit"), f0vq_ref.make_params(seed=9)
    g = torch.Generator().manual_seed(53)
    B, N, mask = 2, 32000, 6400
    wave = 0.1 * torch.randn(B, N, generator=g)
    mu = torch.randn(gcfg.num_embeddings, ocfg.hidden_size, generator=g) * 0.5
    T = ocfg.feat_lengths(N)
    f0 = torch.randn(B, 1, 4 * T, generator=g)                 # hop 80: four f0 frames per code frame
    emb = torch.randn(B, gcfg.embedding_dim, generator=g)
    hub = sib.HubertModel(_hub_cfg(sib, ocfg)).to("cuda")
    hub.load_state_dict(hp)
    gen = sib.CodeGenerator(sib.AttrDict(dict(gcfg.as_attrdict(), f0_quantizer=f0vq_ref.F0_QUANTIZER))).to("cuda")
    gen.load_state_dict(gp)
    gen.load_f0_quantizer(fsd)
    pipe = sib.BlindInpainter(hub, gen, mu, layer=-1, normalize=False)
    res = pipe(wave, mask, emb=emb, f0=f0, informed=True)
    n_code = glue_ref.ida_matched_frames(N, T, 4 * T)
    zp = f0vq_ref.f0_to_bins(fsd, f0[..., : 4 * n_code])        # inpainting.py:243-256: f0 trimmed with the codes
    assert zp.shape == (B, n_code // 4)
    ref = pipe(wave, mask, zp, emb, informed=True)
    assert torch.equal(res.code_inpainting, ref.code_inpainting)
    assert torch.equal(res.audio_inp, ref.audio_inp) and torch.equal(res.audio_gen, ref.audio_gen)
    with pytest.raises(sib.SibError):
        pipe(wave, mask, zp, emb, f0=f0)                         # both given
    with pytest.raises(sib.SibError):
        pipe(wave, mask, emb=emb)                                # neither given


def test_modules_are_real_nn_modules_used_as_the_reference_scripts_do(sib, tmp_path):
    """the shims are `nn.Module`s built, moved, loaded and saved exactly as I_ea/predict.py:117-122,145-150 do:
    Generator(h).to(device); load_state_dict(ckpt['generator']); eval(); remove_weight_norm();
    CustomModel(codebook_dim=80, type=..., load_pretrained=False); model.to(device); load_state_dict(torch.load(ckpt)); eval()."""
    from torch import nn
    from oracle import hifigan_ref, hubert_ref
    from oracle.params import HifiCfg, HubertCfg, fold_weight_norm, make_generator_params, make_head_params, make_hubert_params
    device = torch.device("cuda")
    gcfg = HifiCfg.v1()
    gp = make_generator_params(gcfg, 1234, "unit")
    torch.save({"generator": gp}, tmp_path / "g_ckpt")                    # checkpoint layout of hifi_gan/train.py
    h = sib.AttrDict(gcfg.as_attrdict())
    generator = sib.Generator(h).to(device)
    assert isinstance(generator, nn.Module) and len(generator.state_dict()) == 234 == len(list(generator.named_parameters()))
    state_dict_g = torch.load(tmp_path / "g_ckpt", map_location=device)
    generator.load_state_dict(state_dict_g["generator"])
    generator.eval()
    generator.remove_weight_norm()
    assert len(generator.state_dict()) == 156 and not generator.training
    x = torch.randn(1, 80, 12, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        y = generator(x.to(device))
    ref = hifigan_ref.generator_forward(gp, gcfg, x)
    assert snr_db(ref, y.cpu()) > 60
    for k, v in fold_weight_norm(gp).items():                              # the folded tensors are the reference's
        assert max_abs(v, generator.state_dict()[k].cpu()) < 1e-6 * max(1.0, float(v.abs().max())), k
    # whole-module pickling (torch.save(model)) and a hook
    torch.save(generator, tmp_path / "g_module")
    g2 = torch.load(tmp_path / "g_module", weights_only=False)
    seen = []
    g2.register_forward_hook(lambda m, i, o: seen.append(tuple(o.shape)))
    assert torch.equal(g2(x.to(device)), y) and seen == [(1, 1, 12 * 256)]
    # CustomModel: tiny config stands in for from_pretrained (no network); keys and semantics are the reference's
    ocfg = HubertCfg.tiny(False)
    sd = make_hubert_params(ocfg, 5, prefix="base_model.")
    sd.update(make_head_params(ocfg.hidden_size, 80))
    torch.save(sd, tmp_path / "m_ckpt")
    model = sib.CustomModel(codebook_dim=80, type="base", load_pretrained=False, config=_hub_cfg(sib, ocfg))
    model.to(device)
    model.load_state_dict(torch.load(tmp_path / "m_ckpt", map_location="cuda"))
    model.eval()
    assert isinstance(model, nn.Module) and isinstance(model.base_model, nn.Module)
    assert set(model.state_dict()) == set(sd) and all(not p.requires_grad for p in model.parameters())
    for p_ in model.base_model.encoder.parameters():                       # I_ea/model.py:53-55 touches these
        assert p_.is_cuda
    assert any(isinstance(m, nn.Module) for m in model.modules()) and model.base_model.config.hidden_size == 128
    xw = 0.1 * torch.randn(2, 4000, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        out = model(xw.to(device), None)
    assert max_abs(hubert_ref.custom_model_forward(sd, ocfg, xw), out.cpu()) < 2e-4
    # .half() / .float() round trip keeps the module usable (weights are re-packed from the parameters)
    model.half().float()
    out2 = model(xw.to(device), None)
    assert snr_db(out.cpu(), out2.cpu()) > 50
    model.load_state_dict(torch.load(tmp_path / "m_ckpt", map_location="cuda"))
    assert torch.equal(model(xw.to(device), None), out)


def test_plan_cache_is_bounded_over_many_distinct_lengths(sib):
    """one multi-GB plan per distinct input length must not accumulate - an LRU bounds the resident plans, and
    device memory stays flat over 60 distinct lengths."""
    from oracle.params import HifiCfg, HubertCfg, make_generator_params, make_hubert_params
    ocfg, gcfg = HubertCfg.tiny(False), HifiCfg.tiny()
    hub = sib.HubertModel(_hub_cfg(sib, ocfg)).to("cuda")
    hub.load_state_dict(make_hubert_params(ocfg, 1))
    gen = sib.Generator(sib.AttrDict(gcfg.as_attrdict())).to("cuda")
    gen.load_state_dict(make_generator_params(gcfg, 1, "unit"))
    first = {}
    shapes = [(4000 + 160 * i, 20 + i) for i in range(60)]

    def run(hub_, gen_, todo):
        out = None
        for n, tm in todo:
            y = hub_(torch.zeros(2, n, device="cuda") + 1e-4 * n).last_hidden_state
            w = gen_(torch.full((2, 80, tm), 1e-4 * n, device="cuda"))
            out = out or (y.clone(), w.clone())
        torch.cuda.synchronize()
        return out

    base = torch.cuda.memory_allocated()
    first = run(hub, gen, shapes)
    after_60 = torch.cuda.memory_allocated() - base
    assert len(hub._plans) <= hub._plans.max_plans and len(gen._plans) <= gen._plans.max_plans
    assert hub._plans.evictions >= 60 - hub._plans.max_plans
    # what the resident plans alone need: fresh modules that only ever saw the last `max_plans` shapes
    hub2 = sib.HubertModel(_hub_cfg(sib, ocfg)).to("cuda")
    hub2.load_state_dict(make_hubert_params(ocfg, 1))
    gen2 = sib.Generator(sib.AttrDict(gcfg.as_attrdict())).to("cuda")
    gen2.load_state_dict(make_generator_params(gcfg, 1, "unit"))
    base2 = torch.cuda.memory_allocated()
    run(hub2, gen2, shapes[-hub._plans.max_plans:])
    resident = torch.cuda.memory_allocated() - base2
    assert after_60 <= 1.25 * resident + (4 << 20), (after_60, resident)       # 60 shapes hold what 6 shapes hold
    # an evicted shape is simply re-planned and reproduces its first answer bit for bit
    y0 = hub(torch.zeros(2, 4000, device="cuda") + 1e-4 * 4000).last_hidden_state
    assert torch.equal(y0, first[0]) and torch.equal(gen(torch.full((2, 80, 20), 1e-4 * 4000, device="cuda")), first[1])


def test_wrong_current_device_is_refused_loudly(sib):
    """the C side launches on the CURRENT device; modules enter their own device, raw ops refuse a mismatch."""
    if torch.cuda.device_count() < 2:
        x = torch.zeros(4, 8, device="cuda")
        sib.ops.cast_to_bf16(x, torch.empty(4, 8, device="cuda", dtype=torch.bfloat16))   # same device: fine
        pytest.skip("needs two GPUs for the cross-device half")
    from oracle.params import HifiCfg, make_generator_params
    gcfg = HifiCfg.tiny()
    gp = make_generator_params(gcfg, 1, "unit")
    g0 = sib.Generator(sib.AttrDict(gcfg.as_attrdict())).to("cuda:0"); g0.load_state_dict(gp)
    g1 = sib.Generator(sib.AttrDict(gcfg.as_attrdict())).to("cuda:1"); g1.load_state_dict(gp)
    x = torch.randn(1, 80, 16)
    assert torch.cuda.current_device() == 0
    y0, y1 = g0(x.to("cuda:0")), g1(x.to("cuda:1"))       # g1 runs on cuda:1 although cuda:0 is current
    assert y1.device.index == 1 and torch.equal(y0.cpu(), y1.cpu())
    with pytest.raises(sib.SibError, match="current device"):
        sib.ops.znorm(torch.zeros(1, 64, device="cuda:1"), torch.zeros(1, 64, device="cuda:1"))


def test_out_of_range_codes_trap_on_the_device(sib):
    """a k-means unit / pitch bin / codebook label outside its table is a device-side assert (as
    nn.Embedding), not a silent out-of-bounds read.  Run in a child process: the trap poisons the CUDA context."""
    import subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    child = (
        "import sys, torch; sys.path.insert(0, %r)\n"
        "import speech_inpainting_b200 as sib\n"
        "code = torch.tensor([[1, 2, 99]], device='cuda'); zp = torch.zeros(1, 3, dtype=torch.int64, device='cuda')\n"
        "out = torch.empty(1, 3, 40, device='cuda')\n"
        "sib.ops.embed_concat(code, zp, torch.zeros(1, 8, device='cuda'), torch.zeros(50, 16, device='cuda'),"
        " torch.zeros(20, 16, device='cuda'), out)\n"
        "torch.cuda.synchronize()\n" % root)
    r = subprocess.run([sys.executable, "-c", child], capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and ("outside the 50" in r.stdout + r.stderr or "CUDA error" in r.stderr), r.stderr[-500:]
