"""CUDA-graph replay of a forward: the serving form of the hot path.

Every forward of this package is capture-safe (no host synchronisation, no allocation outside torch's capture pool, launches on
the current stream, the video path's second stream forked and joined by events), so a whole WDSR forward (18 launches) or a whole
BasicVSR clip (~2,100 launches on two streams) replays as ONE graph launch: the Python / ctypes cost per kernel disappears.

    g = Graphed(model, example_x)            # or Graphed(vsr, example_clip, 720, 1280)
    y = g(x)                                  # x is copied into the captured input buffer; y is the captured output buffer

The reference has no counterpart (it calls cuDNN eagerly); outputs are bit-identical to the eager call (tests/test_gpu_*).

Contract on weights: the captured graph holds DEVICE POINTERS into the folded-weight plans that were live at capture.  ``Graphed``
keeps those plans alive and records every sub-module's plan signature; ``__call__`` compares them first and re-captures when a
parameter was modified, re-assigned or loaded since (``load_state_dict``, in-place edits).  Writes through ``.data`` are invisible to
the signature (see ``_PlanCacheMixin``): call ``module.invalidate()`` after them, which also forces the re-capture here.
"""
from __future__ import annotations

import torch

from . import _lib

__all__ = ["Graphed"]


class Graphed:
    def __init__(self, module: torch.nn.Module, example: torch.Tensor, *args, warmup: int = 2, **kwargs):
        _lib.require_cuda_tensor(example, "example")
        self.module, self.args, self.kwargs = module, args, kwargs
        self.warmup = max(1, warmup)
        self.x = example.clone()
        self.stream = torch.cuda.Stream(device=example.device)
        self._capture()

    # modules of this package that cache a folded plan (wdsr._PlanCacheMixin, split.Split_Block, video conv handles)
    def _planned(self):
        return [m for m in self.module.modules() if hasattr(m, "_plan_sig") or hasattr(m, "_signature")]

    def _signatures(self):
        dev = self.x.device
        sigs = []
        for m in self._planned():
            if getattr(m, "_plan_sig", None) is None and hasattr(m, "_signature"):
                sigs.append(None)                      # invalidated since capture
            elif hasattr(m, "_signature"):
                sigs.append(m._signature(dev))
            else:
                sigs.append(tuple((p.data_ptr(), p._version) for p in m.parameters()))
        return sigs

    def _capture(self):
        dev = self.x.device
        self.stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.no_grad(), torch.cuda.stream(self.stream):
            for _ in range(self.warmup):              # plans, tensor maps and the caching allocator settle before the capture
                self.module(self.x, *self.args, **self.kwargs)
            self.stream.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=self.stream):
                self.y = self.module(self.x, *self.args, **self.kwargs)
        torch.cuda.current_stream(dev).wait_stream(self.stream)
        # strong references: the graph reads these plans' device memory on every replay
        self._held = [getattr(m, a) for m in self.module.modules() for a in ("_plan", "_plan_obj", "_handles") if getattr(m, a, None) is not None]
        self._sigs = self._signatures()

    def __call__(self, x: torch.Tensor, clone: bool = False):
        """Replay on the caller's current stream.  The result lives in the captured output buffer (overwritten by the next
        call) unless ``clone=True``."""
        if x.shape != self.x.shape or x.dtype != self.x.dtype:
            raise RuntimeError(f"Graphed: captured for {tuple(self.x.shape)} {self.x.dtype}, got {tuple(x.shape)} {x.dtype}")
        if self._signatures() != self._sigs:          # weights changed since capture: the old plans' pointers are stale
            self._capture()
        self.x.copy_(x, non_blocking=True)
        self.graph.replay()
        y = self.y
        if clone:
            y = tuple(t.clone() for t in y) if isinstance(y, tuple) else y.clone()
        return y
