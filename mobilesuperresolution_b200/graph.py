"""CUDA-graph replay of a forward: the serving form of the hot path.

Every forward of this package is capture-safe (no host synchronisation, no allocation outside torch's capture pool, launches on
the current stream, the video path's second stream forked and joined by events), so a whole WDSR forward (18 launches) or a whole
BasicVSR clip (~2,100 launches on two streams) replays as ONE graph launch: the Python / ctypes cost per kernel disappears.

    g = Graphed(model, example_x)            # or Graphed(vsr, example_clip, 720, 1280)
    y = g(x)                                  # x is copied into the captured input buffer; y is the captured output buffer

The reference has no counterpart (it calls cuDNN eagerly); outputs are bit-identical to the eager call (tests/test_gpu_*).
"""
from __future__ import annotations

import torch

from . import _lib

__all__ = ["Graphed"]


class Graphed:
    def __init__(self, module: torch.nn.Module, example: torch.Tensor, *args, warmup: int = 2, **kwargs):
        _lib.require_cuda_tensor(example, "example")
        self.module, self.args, self.kwargs = module, args, kwargs
        self.x = example.clone()
        self.stream = torch.cuda.Stream(device=example.device)
        self.stream.wait_stream(torch.cuda.current_stream(example.device))
        with torch.no_grad(), torch.cuda.stream(self.stream):
            for _ in range(max(1, warmup)):           # plans, tensor maps and the caching allocator settle before the capture
                module(self.x, *args, **kwargs)
            self.stream.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=self.stream):
                self.y = module(self.x, *args, **kwargs)
        torch.cuda.current_stream(example.device).wait_stream(self.stream)

    def __call__(self, x: torch.Tensor, clone: bool = False):
        """Replay on the caller's current stream.  The result lives in the captured output buffer (overwritten by the next
        call) unless ``clone=True``."""
        if x.shape != self.x.shape or x.dtype != self.x.dtype:
            raise RuntimeError(f"Graphed: captured for {tuple(self.x.shape)} {self.x.dtype}, got {tuple(x.shape)} {x.dtype}")
        self.x.copy_(x, non_blocking=True)
        self.graph.replay()
        y = self.y
        if clone:
            y = tuple(t.clone() for t in y) if isinstance(y, tuple) else y.clone()
        return y
