"""Import the UNMODIFIED reference modules from /root/reference on CPU.

TEST INFRASTRUCTURE ONLY — used by oracle/gen_golden.py (run in the build
container, where /root/reference is mounted) to pin the oracle restatement and
to generate tests/golden/*.  Never imported by the product path, the GPU tests,
smoke() or bench.py (the reference does not exist on the GPU box).

What it does:
  * puts the diffusers/mmgp stand-ins (this directory) and /root/reference on sys.path
  * registers namespace stubs for `wan`, `wan.modules`, `wan.utils`, `wan.distributed`
    so sub-modules import without running wan/__init__.py (which pulls CLIP/T5/decord)
  * makes torch.cuda.get_device_capability() answer (10, 0) when no GPU is present,
    because utils/attention.py:7 queries it at import time.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("LTX_REFERENCE_ROOT", "/root/reference")
_HERE = os.path.dirname(os.path.abspath(__file__))


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "ltx_video"))


def install():
    import torch

    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    for p in (_HERE, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    if not torch.cuda.is_available():
        torch.cuda.get_device_capability = lambda *a, **k: (10, 0)
    for name in ("wan", "wan.modules", "wan.utils", "wan.distributed", "wan.configs"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = [os.path.join(REFERENCE_ROOT, *name.split("."))]
            sys.modules[name] = m
    from mmgp import offload

    offload.shared_state["_attention"] = "sdpa"
    return True
