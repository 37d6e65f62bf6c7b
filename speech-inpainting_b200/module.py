"""`nn.Module` plumbing shared by the reference-surface shims (HubertModel, CustomModel, Generator, CodeGenerator,
F0Quantizer).

The reference's model surface is `torch.nn.Module` (I_ea/model.py:21, I_ea/hifi_gan/models.py:76, I_da/src/model.py:42):
scripts call `.to(device)`, `.load_state_dict(torch.load(ckpt))`, `.eval()`, `.parameters()`, `.state_dict()`,
`torch.save`.  The shims therefore ARE `nn.Module`s whose parameters carry the reference's checkpoint key names
(built as a tree of plain container modules), while the arithmetic stays in libsib_b200.so: kernel-layout copies of the
weights (`_packed`) and the recorded launch plans (`_plans`) are derived, non-persistent state that is dropped whenever
the parameters move or change.
"""
from __future__ import annotations

import os
from collections import OrderedDict

import torch
from torch import nn

from .ops import SibError


class _Node(nn.Module):
    """Plain container: gives dotted checkpoint keys their module hierarchy (`encoder.layers.3.attention.q_proj.weight`)."""

    def forward(self, *a, **k):  # pragma: no cover - never called
        raise SibError("container node of a libsib_b200 module: call the owning model instead")


class PlanCache:
    """LRU of recorded launch plans keyed by input shape.  Every plan pins its own activation buffers (about 2 GB for
    HiFi-GAN V1 and 1.3 GB for HuBERT-base at 32 x 4 s), so a serving loop over ragged lengths must not keep one per shape
    for ever: at most `max_plans` plans and `max_bytes` of plan memory stay resident, the least recently used go first.
    Dropped buffers return to the caching allocator (stream-ordered: kernels already queued on them are safe)."""

    def __init__(self, max_plans: int | None = None, max_bytes: int | None = None):
        self.max_plans = int(os.environ.get("SIB_PLAN_CACHE", "6")) if max_plans is None else max_plans
        self.max_bytes = int(float(os.environ.get("SIB_PLAN_CACHE_GB", "48")) * (1 << 30)) if max_bytes is None else max_bytes
        self._d = OrderedDict()   # key -> (io, bytes)
        self.evictions = 0

    def get(self, key):
        hit = self._d.get(key)
        if hit is None:
            return None
        self._d.move_to_end(key)
        return hit[0]

    def put(self, key, io, nbytes: int):
        self._d[key] = (io, int(nbytes))
        self._d.move_to_end(key)
        while len(self._d) > 1 and (len(self._d) > self.max_plans or self.total_bytes() > self.max_bytes):
            self._d.popitem(last=False)
            self.evictions += 1

    def total_bytes(self) -> int:
        return sum(b for _, b in self._d.values())

    def clear(self):
        self._d.clear()

    def values(self):
        return [io for io, _ in self._d.values()]

    def __getitem__(self, key):
        return self._d[key][0]

    def __contains__(self, key):
        return key in self._d

    def __len__(self):
        return len(self._d)


class SibModule(nn.Module):
    """Base of the shims: reference-named parameters + derived kernel state."""

    _TRANSIENT = ("_packed", "_plans", "_sd_cache", "_head", "_stream_cache")

    def __init__(self):
        super().__init__()
        self._packed = None
        self._plans = PlanCache()
        self._sd_cache = None
        self.training = False          # inference-only path: born in eval mode

    # ---- parameter tree
    def _node_for(self, dotted: str, create: bool):
        mod, parts = self, dotted.split(".")
        for p in parts[:-1]:
            nxt = mod._modules.get(p)
            if nxt is None:
                if not create:
                    return None, parts[-1]
                nxt = _Node()
                nxt.training = False
                mod.add_module(p, nxt)
            mod = nxt
        return mod, parts[-1]

    def _add_param(self, dotted: str, value: torch.Tensor):
        mod, leaf = self._node_for(dotted, True)
        if leaf in mod._parameters:
            del mod._parameters[leaf]
        mod.register_parameter(leaf, nn.Parameter(value, requires_grad=False))
        self._invalidate()

    def _del_param(self, dotted: str):
        mod, leaf = self._node_for(dotted, False)
        if mod is not None and leaf in mod._parameters:
            del mod._parameters[leaf]
        self._invalidate()

    def _has_param(self, dotted: str) -> bool:
        mod, leaf = self._node_for(dotted, False)
        return mod is not None and leaf in mod._parameters

    # ---- derived state
    def _invalidate(self):
        self._packed = None
        self._sd_cache = None
        if getattr(self, "_plans", None) is not None:
            self._plans.clear()
        if hasattr(self, "_head"):
            self._head = None

    @property
    def _sd(self):
        """name -> fp32 contiguous tensor view of every parameter (what the kernels read; cached until the parameters
        move or are reloaded)."""
        if self._sd_cache is None:
            self._sd_cache = {k: (p.detach() if p.dtype == torch.float32 and p.is_contiguous()
                                  else p.detach().to(torch.float32).contiguous())
                              for k, p in self.named_parameters()}
        return self._sd_cache

    @property
    def _device(self) -> torch.device:
        for p in self.parameters():
            return p.device
        return torch.device("cpu")

    def _require_cuda(self):
        if self._device.type != "cuda":
            raise SibError("module must be moved to a CUDA device before forward (no CPU fallback)")

    def _apply(self, fn, *a, **k):
        r = super()._apply(fn, *a, **k)
        for m in self.modules():
            if isinstance(m, SibModule):
                m._invalidate()
        return r

    def refresh(self):
        """Re-derive the kernel-layout weights after parameters were modified in place."""
        for m in self.modules():
            if isinstance(m, SibModule):
                m._invalidate()
        return self

    # ---- nn.Module surface
    def _adapt_state_dict(self, sd: dict) -> dict:
        """Hook: translate alternative checkpoint key styles / restructure parameters before loading."""
        return sd

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        sd = self._adapt_state_dict(dict(state_dict))
        r = super().load_state_dict(sd, strict=strict, assign=assign)
        for p in self.parameters():
            p.requires_grad_(False)
        self.refresh()
        return r

    def train(self, mode: bool = True):
        if mode:
            raise SibError("this is an inference-only path (SURVEY 2: training is out of scope)")
        return super().train(False)

    def __getstate__(self):
        st = self.__dict__.copy()
        for k in self._TRANSIENT:       # ctypes descriptors / device plans are rebuilt on first use after torch.load
            st.pop(k, None)
        return st

    def __setstate__(self, st):
        super().__setstate__(st)
        self._packed = None
        self._plans = PlanCache()
        self._sd_cache = None
        if "_head" in self._TRANSIENT and not hasattr(self, "_head"):
            self._head = None

    def _cached_plan(self, key, builder):
        io = self._plans.get(key)
        if io is None:
            dev = self._device
            before = torch.cuda.memory_allocated(dev)
            io = builder()
            self._plans.put(key, io, max(torch.cuda.memory_allocated(dev) - before, 0))
        return io
