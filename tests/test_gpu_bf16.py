"""-m gpu: the bf16 tensor-core plans (precision="bf16") against the fp32 CPU oracle.

Acceptance (BASELINE north_star): bf16 is judged by a mel-L1 bound on the waveform (hop-256 log-mel,
fmax=None, I_ea/hifi_gan/meldataset.py:49-79) and by the code-agreement rate of the integer outputs; SNR is
reported and bounded loosely.  Measured values on B200 are recorded in DESIGN.md."""
import numpy as np
import pytest
import torch

from util import max_abs, snr_db

pytestmark = pytest.mark.gpu

MEL_L1_BOUND = 0.02      # log-mel units; measured on B200: 0.0045 (V1 generator), 0.0050 (I_ea config 1)
SNR_BOUND_DB = 30.0      # vs fp32 oracle; measured on B200: generator 40.7-49.9 dB, HuBERT-base 37.2 dB


@pytest.fixture(scope="module")
def sib():
    import speech_inpainting_b200 as m
    m._load_lib()
    return m


@pytest.mark.parametrize("kind,B,T", [("tiny", 2, 40), ("v1", 2, 43), ("ida", 1, 12), ("v3", 2, 37)])
def test_generator_bf16(sib, kind, B, T):
    """v3 = config_v3.json: ResBlock2 (models.py:52-73) through the tcgen05 convs with residual + in-kernel leaky-relu."""
    from oracle import hifigan_ref, mel_ref
    from oracle.params import HifiCfg, make_generator_params
    cfg = {"v1": HifiCfg.v1(), "v3": HifiCfg.v3(), "tiny": HifiCfg.tiny(), "ida": HifiCfg.ida(), "ida_tiny": HifiCfg.tiny(True)}[kind]
    params = {k: v for k, v in make_generator_params(cfg, 1234, "unit").items() if not k.startswith("emb_")}
    gen = sib.Generator(sib.AttrDict(cfg.as_attrdict()), precision="bf16").to("cuda")
    gen.load_state_dict(params)
    x = torch.randn(B, cfg.model_in_dim, T, generator=torch.Generator().manual_seed(11))
    ref = hifigan_ref.generator_forward(params, cfg, x)
    y = gen(x.cuda()).cpu()
    assert y.shape == ref.shape
    s = snr_db(ref, y)
    print(f"\n[bf16 generator {kind}] SNR {s:.1f} dB, max-abs {max_abs(ref, y):.4f}, |ref|max {ref.abs().max():.3f}")
    assert s > SNR_BOUND_DB
    if ref.shape[-1] >= 4096:
        l1 = mel_ref.mel_l1(ref[:, 0], y[:, 0])
        print(f"[bf16 generator {kind}] mel-L1 {l1:.4f}")
        assert l1 < MEL_L1_BOUND


@pytest.mark.parametrize("name,B,N", [("tiny_group", 2, 8000), ("tiny_layer", 2, 8000), ("base", 2, 32000), ("large", 2, 24000)])
def test_hubert_bf16(sib, name, B, N):
    from oracle import hubert_ref
    from oracle.params import HubertCfg, make_hubert_params
    ocfg = {"tiny_group": HubertCfg.tiny(False), "tiny_layer": HubertCfg.tiny(True), "base": HubertCfg.base(),
            "large": HubertCfg.large()}[name]
    params = make_hubert_params(ocfg, 1234)
    model = sib.HubertModel(sib.HubertConfig.from_any(ocfg), precision="bf16").to("cuda")
    model.load_state_dict(params)
    x = 0.1 * torch.randn(B, N, generator=torch.Generator().manual_seed(3))
    am = torch.ones(B, N, dtype=torch.long)
    am[1, N - 2000:] = 0
    x[1, N - 2000:] = 0
    for mask in (None, am):
        ref = hubert_ref.hubert_forward(params, ocfg, x, mask)
        y = model(x.cuda(), None if mask is None else mask.cuda()).last_hidden_state.cpu()
        s = snr_db(ref, y)
        print(f"\n[bf16 hubert {name} padded={mask is not None}] SNR {s:.1f} dB, max-abs {max_abs(ref, y):.4f}")
        assert s > SNR_BOUND_DB


@pytest.mark.parametrize("name,B,N", [("tiny_group", 2, 8000), ("tiny_layer", 2, 8000), ("base", 2, 32000)])
def test_hubert_bf16_with_folded_layernorm(sib, name, B, N):
    """The optional LayerNorm folding (`sib_linear_ln_bf16`: LN(t) W + b = r (t W' - mu s) + c in the epilogue of the consuming
    linear layer, LN of the residual rebuilt from the raw residual tile, row statistics emitted by the producers) against the
    same CPU oracle; post-LN (group) and pre-LN (layer) stacks."""
    from oracle import hubert_ref
    from oracle.params import HubertCfg, make_hubert_params
    ocfg = {"tiny_group": HubertCfg.tiny(False), "tiny_layer": HubertCfg.tiny(True), "base": HubertCfg.base()}[name]
    params = make_hubert_params(ocfg, 1234)
    model = sib.HubertModel(sib.HubertConfig.from_any(ocfg), precision="bf16").to("cuda")
    model.fold_layernorm = True
    model.load_state_dict(params)
    x = 0.1 * torch.randn(B, N, generator=torch.Generator().manual_seed(3))
    ref = hubert_ref.hubert_forward(params, ocfg, x, None)
    n0 = sib.ops.launch_count()
    y = model(x.cuda()).last_hidden_state.cpu()
    plan = model._plans.values()[0].plan
    outside = 10 if ocfg.feat_extract_norm == "layer" else 3       # feature encoder / projection / encoder-level LayerNorms
    assert sum(1 for _, _, nme in plan.steps if nme == "sib_layernorm") <= outside    # none left inside the layer loop
    assert sum(1 for _, _, nme in plan.steps if nme == "sib_linear_ln_bf16") >= 4 * ocfg.num_hidden_layers - 1
    s = snr_db(ref, y)
    print(f"\n[bf16 hubert {name}, folded LayerNorm] SNR {s:.1f} dB, max-abs {max_abs(ref, y):.4f}")
    assert s > SNR_BOUND_DB


def test_informed_inpainting_bf16_config1(sib):
    """config #1 shapes through the bf16 arm: labels vs fp32 oracle (agreement rate), waveform mel-L1."""
    from oracle import mel_ref
    from oracle.params import HifiCfg, HubertCfg, make_codebook, make_generator_params, make_head_params, make_hubert_params
    from test_gpu_models import _iea_oracle
    ocfg, gcfg = HubertCfg.base(), HifiCfg.v1()
    sd = make_hubert_params(ocfg, 1234, prefix="base_model.")
    sd.update(make_head_params(768, 80))
    gparams = make_generator_params(gcfg, 1234, "unit")
    C = make_codebook(80, 100)
    g = torch.Generator().manual_seed(1234)
    B = 4
    wave, mel = 0.1 * torch.randn(B, 32000, generator=g), torch.randn(B, 80, 100, generator=g)
    pos, ln = [45, 10, 70, 0], [10, 20, 5, 10]
    ref_wave, ref_labels, _ = _iea_oracle(sd, ocfg, gparams, gcfg, C, wave, mel, pos, ln)
    model = sib.CustomModel(80, "base", False, config=sib.HubertConfig.base(), precision="bf16").to("cuda")
    model.load_state_dict(sd)
    gen = sib.Generator(sib.AttrDict(gcfg.as_attrdict()), precision="bf16").to("cuda")
    gen.load_state_dict(gparams)
    res = sib.InformedInpainter(model, gen, C)(wave, mel, pos, ln)
    agree = float((ref_labels == res.labels.cpu()).float().mean())
    # compare waveforms only where the pasted codes agree (a flipped code changes a whole 20 ms frame by design)
    l1 = mel_ref.mel_l1(ref_wave[:, 0], res.wave[:, 0].cpu())
    s = snr_db(ref_wave, res.wave.cpu())
    print(f"\n[bf16 I_ea cfg1] label agreement {agree:.3f}, waveform SNR {s:.1f} dB, mel-L1 {l1:.4f}")
    assert agree >= 0.98
    assert l1 < MEL_L1_BOUND


def test_bf16_rejects_unsupported_channel_counts_loudly(sib):
    """c_in/groups must be a multiple of 16 on the tcgen05 arm: no silent fallback to another path."""
    from oracle.params import HifiCfg, make_generator_params
    cfg = HifiCfg.tiny(True)   # last stage has 8 channels
    gen = sib.Generator(sib.AttrDict(cfg.as_attrdict()), precision="bf16").to("cuda")
    gen.load_state_dict({k: v for k, v in make_generator_params(cfg, 1, "unit").items() if not k.startswith("emb_")})
    with pytest.raises(sib.SibError, match="multiple of 16"):
        gen(torch.randn(1, cfg.model_in_dim, 8).cuda())
