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


@pytest.mark.parametrize("name,B,N", [("tiny_group", 3, 8000), ("tiny_layer", 3, 8000), ("base", 5, 32000), ("base", 32, 64000),
                                      ("large", 3, 24000)])
def test_hubert_bf16_dataflow_is_bit_identical(sib, name, B, N):
    """Tile-level dataflow through the transformer loop (`sib_flow`: per-128-row-block counters instead of grid-wide kernel
    boundaries between out-proj / LayerNorm / FFN-in / FFN-out / LayerNorm / QKV) changes WHEN rows are read, never the
    arithmetic: the output must equal the plain launch chain's bit for bit - post-LN and pre-LN stacks, row counts with a
    partial last block, padded batches, and a repeated replay of the same plan (counters reset at the head of every run)."""
    from oracle.params import HubertCfg, make_hubert_params
    ocfg = {"tiny_group": HubertCfg.tiny(False), "tiny_layer": HubertCfg.tiny(True), "base": HubertCfg.base(),
            "large": HubertCfg.large()}[name]
    params = make_hubert_params(ocfg, 1234)
    x = 0.1 * torch.randn(B, N, generator=torch.Generator().manual_seed(5))
    am = torch.ones(B, N, dtype=torch.long)
    am[1, N - 2000:] = 0
    outs = {}
    L = ocfg.num_hidden_layers
    want = {0: 0, 1: 6 * L - (0 if ocfg.do_stable_layer_norm else 1), 2: 7 * L}   # level 2: every launch of the loop
    for flow in (0, 1, 2):
        model = sib.HubertModel(sib.HubertConfig.from_any(ocfg), precision="bf16").to("cuda")
        model.flow = flow
        model.load_state_dict(params)
        res = []
        for mask in (None, am, None):
            res.append(model(x.cuda(), None if mask is None else mask.cuda()).last_hidden_state.cpu())
        plan = list(model._plans.values())[0].plan
        n_flow = sum(1 for _, _, nme in plan.steps if nme in ("sib_linear_flow_bf16", "sib_layernorm_flow_bf16", "sib_attention_flow_bf16"))
        assert n_flow == want[flow]
        assert torch.equal(res[0], res[2])
        outs[flow] = res
    for level in (1, 2):
        for a, b in zip(outs[0], outs[level]):
            assert torch.isfinite(a).all()
            assert torch.equal(a, b)


def test_linear_flow_chain_against_plain_launches(sib):
    """Three chained linear layers + a LayerNorm through `sib_linear_flow_bf16` / `sib_layernorm_flow_bf16` on 6368 rows
    (49 full blocks + 96 rows; CTA pairs) and on 300 rows against the same launches without counters."""
    ops = sib.ops
    g = torch.Generator().manual_seed(9)
    for M in (6368, 300, 128):
        H, I = 768, 3072
        x = torch.randn(M, H, generator=g).to("cuda", torch.bfloat16)
        w1 = ops.to_kmajor_bf16(ops.pack_linear_weight((0.03 * torch.randn(I, H, generator=g)).cuda()))
        w2 = ops.to_kmajor_bf16(ops.pack_linear_weight((0.03 * torch.randn(H, I, generator=g)).cuda()))
        b1, b2 = (0.1 * torch.randn(I, generator=g)).cuda(), (0.1 * torch.randn(H, generator=g)).cuda()
        gam, bet = (1 + 0.1 * torch.randn(H, generator=g)).cuda(), (0.1 * torch.randn(H, generator=g)).cuda()
        res = {}
        for flow in (False, True):
            h1 = torch.empty(M, I, device="cuda", dtype=torch.bfloat16)
            h2 = torch.empty(M, H, device="cuda", dtype=torch.bfloat16)
            n1 = torch.empty(M, H, device="cuda", dtype=torch.bfloat16)
            h3 = torch.empty(M, I, device="cuda", dtype=torch.bfloat16)
            if flow:
                ch = ops.FlowChain(M, 3, "cuda")
                for _ in range(3):      # replayed: the counters are reset at the head of every run
                    ch._next = 0
                    ch.reset()
                    e1, e2, e3 = ch.edge("linear", I), ch.edge("linear", H), ch.edge("layernorm", H)
                    ops.linear(x, w1, b1, h1, post_act=ops.ACT_GELU, signal=e1)
                    ops.linear(h1, w2, b2, h2, residual=x, wait=e1, signal=e2)
                    ops.layernorm(h2, gam, bet, n1, 1e-5, residual=x, wait=e2, signal=e3)
                    ops.linear(n1, w1, b1, h3, wait=e3)
            else:
                ops.linear(x, w1, b1, h1, post_act=ops.ACT_GELU)
                ops.linear(h1, w2, b2, h2, residual=x)
                ops.layernorm(h2, gam, bet, n1, 1e-5, residual=x)
                ops.linear(n1, w1, b1, h3)
            torch.cuda.synchronize()
            res[flow] = (h1.cpu(), h2.cpu(), n1.cpu(), h3.cpu())
        for a, b in zip(res[False], res[True]):
            assert torch.isfinite(a.float()).all()
            assert torch.equal(a, b)


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
