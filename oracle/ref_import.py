"""Import the UNMODIFIED reference from /root/reference (only where it exists: this container).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  Used by
``oracle/make_golden.py`` and by the CPU tests that compare the oracle with the
live reference; nothing that runs on the GPU box may depend on it.

Shims, test-harness only:
* ``mmedit`` is an un-vendored, unpinned dependency -> stub package re-exporting the in-repo
  ``models/spynet_arch`` twins so ``models/basicvsr_arch*.py`` import.
* ``sys.argv`` is parked while importing ``export_onnx`` (it has a CLI ``__main__`` guard only, but
  importing under pytest would otherwise see pytest's argv).
"""
from __future__ import annotations

import os
import sys
import types
import warnings

REF = os.environ.get("B200SR_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "models", "basic_wdsr_b.py"))


_done = False


def install() -> None:
    global _done
    if _done:
        return
    if not available():
        raise RuntimeError(f"reference tree not found at {REF}")
    sys.dont_write_bytecode = True          # the tree is read-only
    if REF not in sys.path:
        sys.path.insert(0, REF)
    warnings.filterwarnings("ignore", message=".*weight_norm.*")
    import models.spynet_arch as sa

    for n in ["mmedit", "mmedit.models", "mmedit.models.common", "mmedit.models.backbones",
              "mmedit.models.backbones.sr_backbones", "mmedit.models.backbones.sr_backbones.basicvsr_net",
              "mmedit.core", "mmedit.core.evaluation", "mmedit.core.evaluation.metrics"]:
        sys.modules.setdefault(n, types.ModuleType(n))
    common = sys.modules["mmedit.models.common"]
    common.flow_warp = sa.flow_warp
    common.PixelShufflePack = object

    class SPyNet(sa.SpyNet):
        def __init__(self, pretrained=None):
            super().__init__(None)

    bn = sys.modules["mmedit.models.backbones.sr_backbones.basicvsr_net"]
    bn.SPyNet = SPyNet
    bn.ResidualBlocksWithInputConv = object
    _done = True


def modules():
    """Returns a namespace with the reference classes on the hot path."""
    install()
    import importlib
    ns = types.SimpleNamespace()
    from models.basic_wdsr_b import BASIC_MODEL, Block as BasicBlock
    import models.wdsr_b as wdsr_b
    import models.ops as ops
    import models.spynet_arch as sa
    argv, sys.argv = sys.argv, ["export_onnx"]
    try:
        export_onnx = importlib.import_module("export_onnx")
    finally:
        sys.argv = argv
    import models.basicvsr_arch_origin as origin
    import models.basicvsr_arch as fork
    ns.BASIC_MODEL, ns.BasicBlock = BASIC_MODEL, BasicBlock
    ns.wdsr_b, ns.ops, ns.spynet_arch, ns.export_onnx = wdsr_b, ops, sa, export_onnx
    ns.basicvsr_origin, ns.basicvsr_fork = origin, fork
    return ns


def naive_model():
    """models/naive_multi_model_easy.Naive_model.  The module imports the whole training stack at its top (tensorboard, torchvision,
    skimage, mmedit's metrics, ...): everything that is missing is replaced by an empty stub module -- none of it is touched by the class."""
    install()
    import importlib

    class _Any(types.ModuleType):
        def __getattr__(self, k):
            if k.startswith("__"):
                raise AttributeError(k)
            return object

    for n in list(sys.modules):
        if n.startswith("mmedit.core"):
            sys.modules[n].psnr = sys.modules[n].ssim = object
    for _ in range(40):
        try:
            return importlib.import_module("models.naive_multi_model_easy").Naive_model
        except ModuleNotFoundError as e:
            sys.modules[e.name] = _Any(e.name)
    raise RuntimeError("could not import models.naive_multi_model_easy")
