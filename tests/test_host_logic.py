"""CPU (-m "not gpu"): host-side logic of the product package, the C-ABI surface, and the N>1 sharding path (gloo)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def sib():
    import speech_inpainting_b200 as m
    return m


def test_host_integer_arithmetic_matches_oracle(sib):
    from oracle import glue_ref
    for s, e in [(0.9, 1.1), (1.0, 1.4), (2.98, 3.38), (0.0, 0.02), (3.5, 3.9)]:
        a, b = sib.iea_mask_indices(s, e), glue_ref.iea_mask_indices(s, e)
        assert a == b
    for pos, L in [(0, 1), (149, 20), (45, 10), (7, 0)]:
        assert sib.iea_zero_range(pos, L) == glue_ref.iea_zero_range_from_frames(pos, L)
    for n, t, f in [(64000, 199, 800), (63999, 199, 799), (160000, 499, 2000), (32000, 99, 400), (5000, 15, 62)]:
        assert sib.ida_matched_frames(n, t, f) == glue_ref.ida_matched_frames(n, t, f)
    cfg = sib.HubertConfig.base()
    assert [cfg.feat_extract_output_length(n) for n in (32000, 64000, 96000, 160000)] == [99, 199, 299, 499]
    assert sib.ops.extend_mel_len(100) == 172 and sib.ops.extend_mel_len(200) == 344
    assert sib.get_padding(11, 5) == 25 and sib.get_padding(3, 1) == 1


def test_shard_batch_partition(sib):
    for n, w in [(1024, 8), (10, 4), (3, 8), (128, 1)]:
        spans = [sib.shard_batch(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1


def test_conv_transpose_polyphase_packing_matches_torch(sib):
    """The poly-phase re-layout of ConvTranspose1d is pure host index work: check it on the CPU with F.conv1d."""
    import torch.nn.functional as F
    for k, s in [(16, 8), (4, 2), (11, 5), (8, 4)]:
        cin, cout, T = 6, 4, 13
        g = torch.Generator().manual_seed(k)
        x, w, b = torch.randn(2, cin, T, generator=g), torch.randn(cin, cout, k, generator=g), torch.randn(cout, generator=g)
        ref = F.conv_transpose1d(x, w, b, stride=s, padding=(k - s) // 2)
        wp, bp, taps = sib.ops.pack_conv_transpose(w, b, s, (k - s) // 2)   # [1][taps][cin][s*cout]
        lo, hi = -min(taps), max(taps)
        xp = F.pad(x, (lo, hi))
        y = torch.zeros(2, T, s * cout)
        for ti, off in enumerate(taps):
            y += torch.einsum("bct,cn->btn", xp[:, :, lo + off: lo + off + T], wp[0, ti])
        y = (y + bp).reshape(2, T * s, cout).transpose(1, 2)
        assert torch.allclose(ref, y, atol=1e-5)


def test_state_dict_surface_on_cpu(sib):
    from oracle.params import HifiCfg, HubertCfg, make_generator_params, make_head_params, make_hubert_params
    ocfg = HubertCfg.tiny(False)
    sd = make_hubert_params(ocfg, 1, prefix="base_model.")
    sd.update(make_head_params(ocfg.hidden_size, 80))
    m = sib.CustomModel(80, "base", False, config=sib.HubertConfig.from_any(ocfg))
    m.load_state_dict(sd)
    assert m.base_model.config.hidden_size == 128 and len(list(m.parameters())) == len(sd)
    m.base_model.config.mask_time_prob = 0           # the reference writes these fields (I_ea/model.py:58-63)
    with pytest.raises(RuntimeError):
        m.load_state_dict({k: v for k, v in sd.items() if "k_proj" not in k})
    with pytest.raises(sib.SibError, match="no CPU fallback"):
        m(torch.zeros(1, 4000))                      # never silently runs on the CPU
    gcfg = HifiCfg.tiny()
    gen = sib.Generator(sib.AttrDict(gcfg.as_attrdict()))
    gp = make_generator_params(gcfg, 1)
    gen.load_state_dict(gp)
    assert len(gp) == 3 * len(gen._conv_names())     # bias + weight_g + weight_v per conv (234 keys for V1)
    v1 = sib.Generator(sib.AttrDict(HifiCfg.v1().as_attrdict()))
    assert len(v1._expected_keys()) == 234 and v1.total_upsample == 256
    with pytest.raises(sib.SibError):
        gen(torch.zeros(1, 80, 8))
    with pytest.raises(sib.SibError):
        gen.train(True)


def test_c_abi_library_exports_every_declared_symbol(sib):
    """include/speech_inpainting_b200.h <-> libsib_b200.so <-> the ctypes binding (no compute calls)."""
    header = open(os.path.join(ROOT, "include", "speech_inpainting_b200.h")).read()
    declared = set(re.findall(r"^\s*(?:int|long long|size_t|const char\*)\s+(sib_\w+)\s*\(", header, flags=re.M))
    assert len(declared) >= 30
    lib_path = os.path.join(ROOT, "speech-inpainting_b200", "libsib_b200.so")
    if not os.path.exists(lib_path):
        pytest.skip("libsib_b200.so not built yet (run __graft_entry__.build())")
    out = subprocess.run(["nm", "-D", "--defined-only", lib_path], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert declared <= exported, f"declared but not exported: {sorted(declared - exported)}"
    assert set(sib.exported_symbols()) <= declared, sorted(set(sib.exported_symbols()) - declared)
    lib = ctypes.CDLL(lib_path)          # loads without a GPU
    lib.sib_abi_version.restype = ctypes.c_int
    assert lib.sib_abi_version() == 2


@pytest.mark.parametrize("cname,pyname", [("sib_conv_desc", "ConvDesc"), ("sib_resunit_desc", "ResUnitDesc"), ("sib_ln_fold", "LnFold"),
                                          ("sib_flow", "Flow")])
def test_struct_layouts_match_header(sib, tmp_path, cname, pyname):
    """struct layout agreement between include/*.h (compiled with gcc) and every ctypes mirror."""
    from speech_inpainting_b200 import _lib
    mirror = getattr(_lib, pyname)
    fields = [f[0] for f in mirror._fields_]
    src = tmp_path / "layout.c"
    lines = "\n".join(f'  printf("{f} %zu\\n", offsetof({cname}, {f}));' for f in fields)
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "speech_inpainting_b200.h"\nint main(void) {\n'
                   f'  printf("sizeof %zu\\n", sizeof({cname}));\n{lines}\n  return 0;\n}}\n')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    assert int(out["sizeof"]) == ctypes.sizeof(mirror)
    for f in fields:
        assert int(out[f]) == getattr(mirror, f).offset, f


def test_flow_edge_targets():
    """Dataflow counters (`sib_flow`): what a complete 128-row block is worth per producer kind.  A linear layer stores whole
    32-row quarters of every tile (4 x n per block, valid rows or not); LayerNorm and attention add n / 32 per VALID row,
    so only their last block's target depends on the row count."""
    from speech_inpainting_b200.ops import FlowEdge
    assert FlowEdge(None, "linear", 768, 6368).targets() == (3072, 3072)
    assert FlowEdge(None, "layernorm", 768, 6368).targets() == (3072, 96 * 24)          # 6368 = 49 x 128 + 96
    assert FlowEdge(None, "attention", 1024, 256).targets() == (4096, 128 * 32)        # a full last block
    assert FlowEdge(None, "layernorm", 128, 72).targets() == (512, 72 * 4)


def test_no_product_import_of_oracle():
    """The oracle is test infrastructure: nothing under the product package may import it."""
    pkg = os.path.join(ROOT, "speech-inpainting_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src, fn


_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
import speech_inpainting_b200 as sib
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
n = 37
lo, hi = sib.shard_batch(n, world, rank)
# every rank owns a disjoint utterance range; the only collective is the timing/bookkeeping reduction
owned = torch.zeros(n, dtype=torch.int64); owned[lo:hi] = 1
dist.all_reduce(owned)
assert torch.all(owned == 1), owned
t = torch.tensor([10.0 + rank])
dist.all_reduce(t, op=dist.ReduceOp.MAX)
assert float(t) == 10.0 + world - 1
audio_s = torch.tensor([float(hi - lo) * 4.0]); dist.all_reduce(audio_s)
assert float(audio_s) == n * 4.0
dist.barrier(); dist.destroy_process_group()
print("ok", rank)
'''


def test_two_rank_sharding_over_gloo(tmp_path):
    """N>1 path on the CPU: world_size 2, gloo, 127.0.0.1 - utterance shards are disjoint and complete; timing is max-reduced."""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29611", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script), ROOT], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs)


def test_bench_reference_arm_prints_exactly_one_json_line():
    """bench.py contract: stdout is ONE JSON line (the driver parses it).  The reference arm runs on the CPU, so it is
    checked here; fd 1 is re-pointed at stderr inside bench.py, so even C-level prints below Python cannot add lines."""
    import json
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-sample-utts", "1"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "inpainted_audio_seconds_per_second" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["e2e"]["h2d_bytes_per_step"] == 0 and d["higher_is_better"] is True
    assert "transformers.HubertModel" in d["cpu_baseline"]["sample"]      # the real dependency, not a port, runs HuBERT
