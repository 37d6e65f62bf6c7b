import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # keep libsib_b200.so in step with csrc/ (a no-op when the per-file digests match; needs nvcc, else the prebuilt .so is used)
    try:
        import importlib.util
        spec = importlib.util.spec_from_file_location("_sib_build", os.path.join(ROOT, "speech-inpainting_b200", "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        if os.path.exists(mod.NVCC):
            mod.build()
    except Exception as e:  # pragma: no cover - a stale library then fails the ABI / symbol tests loudly
        print(f"[conftest] could not rebuild libsib_b200.so: {e}", file=sys.stderr)


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
