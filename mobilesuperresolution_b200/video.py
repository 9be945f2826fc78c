"""Video path over the B200 C ABI: ``flow_warp``, ``SpyNet``, ``BasicVSR_origin`` / ``BasicVSR`` (fork) / ``MotionVectorVSR``.

Mirrors models/spynet_arch.py (the in-repo twin of the un-vendored ``mmedit`` functions the BasicVSR files import),
models/basicvsr_arch_origin.py, models/basicvsr_arch.py and models/mvvsr_arch.py: same constructors, ``forward`` signatures and
``state_dict`` keys.  In bf16 precision the convolutions run on the tcgen05 kernels (csrc/conv_tc5.cuh, conv7_tc5.cuh) with the
tensors between consecutive convolutions kept planar-8; fp32 precision is the true-fp32 FFMA parity arm (1e-4 gate).
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib


def _ptr(t: torch.Tensor) -> ctypes.c_void_p:
    return ctypes.c_void_p(t.data_ptr())


def flow_warp(x: torch.Tensor, flow: torch.Tensor, interp_mode: str = "bilinear", padding_mode: str = "zeros",
              align_corners: bool = True) -> torch.Tensor:
    """Warp ``x`` (n,c,h,w) with ``flow`` (n,h,w,2).  Signature and assert of models/spynet_arch.py:98-129.

    ``flow`` may be any strided view (the callers pass ``flow.permute(0,2,3,1)``): it is consumed in place.
    Only the configuration the reference's callers use is accelerated (bilinear, zeros|border, align_corners=True).
    """
    assert x.size()[-2:] == flow.size()[1:3]
    if interp_mode != "bilinear" or not align_corners or padding_mode not in ("zeros", "border"):
        raise NotImplementedError("b200sr.flow_warp: bilinear, padding_mode zeros|border, align_corners=True only")
    _lib.require_cuda_tensor(x, "x")
    _lib.require_cuda_tensor(flow, "flow")
    if x.dtype != torch.float32 or flow.dtype != torch.float32:
        raise TypeError("b200sr.flow_warp: float32 tensors (use flow_warp_nhwc for the bf16 internal layout)")
    x = x.contiguous()
    n, c, h, w = x.shape
    y = torch.empty_like(x)
    if x.numel() == 0:
        return y
    sn, sh, sw, sc = flow.stride()
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().b200sr_flow_warp_nchw(
            _ptr(x), _ptr(flow), sn, sh, sw, sc, _ptr(y), n, c, h, w,
            _lib.PAD_BORDER if padding_mode == "border" else _lib.PAD_ZEROS, _lib.current_stream_ptr(x.device)))
    return y


def flow_warp_nhwc(x: torch.Tensor, flow_nchw: torch.Tensor, padding_mode: str = "zeros", out: Optional[torch.Tensor] = None,
                   out_coff: int = 0) -> torch.Tensor:
    """Internal-layout warp: ``x`` (n,h,w,c) float32|bfloat16, contiguous or a channel window ``wide[..., a:a+c]`` of a contiguous
    (n,h,w,C) tensor; ``flow_nchw`` (n,2,h,w) float32.  ``out``: write into channels [out_coff, out_coff + c) of this wider (n,h,w,C)
    tensor instead of a new one (a concatenation without the copy)."""
    _lib.require_cuda_tensor(x, "x")
    assert flow_nchw.is_contiguous() and flow_nchw.dtype == torch.float32
    n, h, w, c = x.shape
    x_cs = x.stride(2) if h * w > 1 else c
    assert x.stride(3) == 1 and x_cs >= c and (n * h * w <= 1 or x.stride() == (h * w * x_cs, w * x_cs, x_cs, 1)), "x: NHWC or a channel window of NHWC"
    assert tuple(flow_nchw.shape) == (n, 2, h, w)
    y = torch.empty((n, h, w, c), dtype=x.dtype, device=x.device) if out is None else out
    assert y.is_contiguous() and y.dtype == x.dtype and tuple(y.shape[:3]) == (n, h, w)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().b200sr_flow_warp_nhwc_windows(
            _ptr(x), x_cs, 0, _ptr(flow_nchw), _ptr(y), y.shape[-1], out_coff, n, c, h, w,
            _lib.PAD_BORDER if padding_mode == "border" else _lib.PAD_ZEROS, _lib.dtype_code(x.dtype),
            _lib.current_stream_ptr(x.device)))
    return y


def _cat_feats(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """``torch.cat([a, b], dim=-1)`` -- for free when the two are the adjacent channel windows of one tensor (propagate lays them out so)."""
    base = a._base
    if (base is not None and base is b._base and base.is_contiguous() and base.dim() == 4 and base.shape[-1] == a.shape[-1] + b.shape[-1]
            and a.stride() == base.stride() == b.stride() and a.storage_offset() == base.storage_offset()
            and b.storage_offset() == base.storage_offset() + a.shape[-1]):
        return base
    return torch.cat([a, b], dim=-1)


# ======================================================================================================
# SPyNet + BasicVSR over the C ABI
# ======================================================================================================
import math
import os
from typing import Dict, List, Optional, Tuple

import torch.nn as nn

ACT_NONE, ACT_RELU, ACT_LRELU = 0, 1, 2


class _ConvHandle:
    """One ``b200sr_conv_t``: an nn.Conv2d's filters packed for both arithmetic paths, resident on one device."""

    def __init__(self, conv: nn.Conv2d, device: torch.device):
        w = conv.weight.detach().float().cpu().contiguous()
        b = conv.bias.detach().float().cpu().contiguous() if conv.bias is not None else None
        self.cout, self.cin, self.k = int(w.shape[0]), int(w.shape[1]), int(w.shape[2])
        h = ctypes.c_void_p()
        with torch.cuda.device(device):
            _lib.check(_lib.lib().b200sr_conv_create(self.cin, self.cout, self.k, _ptr(w), _ptr(b) if b is not None else None,
                                                     ctypes.byref(h)))
        self._h, self.device = h, device

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                _lib.lib().b200sr_conv_destroy(h)
            except Exception:
                pass

    def set_max_ctas(self, n: int) -> None:
        """Grid cap of this conv's tcgen05 launches (0 = one CTA per SM)."""
        _lib.check(_lib.lib().b200sr_conv_set_max_ctas(self._h, int(n)))

    def tcgen05_ok(self) -> bool:
        """True when a tcgen05 kernel (3x3 (64..80) -> 64 k, SPyNet 7x7 layers) serves this conv in bf16 -- the only kernels that take
        planar-8 tensors."""
        return bool(_lib.lib().b200sr_conv_tcgen05_ok(self._h))

    def __call__(self, x: torch.Tensor, precision: str, act: int = ACT_NONE, x_coff: int = 0, out: Optional[torch.Tensor] = None,
                 y_coff: int = 0, residual: Optional[torch.Tensor] = None, shuffle: int = 1,
                 out_dtype: Optional[torch.dtype] = None, x_planar: bool = False, y_planar: bool = False) -> torch.Tensor:
        """NHWC tensors (n,h,w,c); ``x_planar`` / ``y_planar``: planar-8 tensors (n,c/8,h,w,8) instead (the residual follows x)."""
        assert x.is_contiguous()
        if x_planar:
            n, _, h, w, _ = x.shape
            xcs = self.cin
        else:
            n, h, w, xcs = x.shape
        if out is None:
            oc = self.cout // (shuffle * shuffle)
            shape = (n, oc // 8, h * shuffle, w * shuffle, 8) if y_planar else (n, h * shuffle, w * shuffle, oc)
            out = torch.empty(shape, dtype=out_dtype or x.dtype, device=x.device)
        ycs = self.cout if y_planar else out.shape[-1]
        rcs = 0 if residual is None else self.cout if x_planar else residual.shape[-1]
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().b200sr_conv_forward_layout(
                self._h, _ptr(x), int(x_planar), xcs, x_coff, _ptr(out), int(y_planar), ycs, y_coff,
                _ptr(residual) if residual is not None else None, rcs, 0,
                n, h, w, act, shuffle, _lib.dtype_code(x.dtype), _lib.dtype_code(out.dtype), _lib.precision_code(precision),
                _lib.current_stream_ptr(x.device)))
        return out


class _VideoPlanMixin:
    precision: str = "fp32"

    def set_precision(self, precision: str):
        _lib.precision_code(precision)
        self.precision = precision
        for m in self.children():
            if isinstance(m, _VideoPlanMixin):
                m.set_precision(precision)
        return self

    def _act_dtype(self) -> torch.dtype:
        return torch.float32 if self.precision == "fp32" else torch.bfloat16

    def _convs(self, device) -> Dict[str, _ConvHandle]:
        """Conv handles of this module's own nn.Conv2d leaves, rebuilt when a parameter changes (cf. wdsr._PlanCacheMixin)."""
        own = [(n, m) for n, m in self.named_modules() if isinstance(m, nn.Conv2d) and not n.startswith("spynet.")]
        sig = (str(device),) + tuple((p.data_ptr(), p._version) for _, m in own for p in m.parameters())
        if getattr(self, "_conv_sig", None) != sig:
            rot = self._feat_first()
            self._conv_cache = {n: _ConvHandle(_rotate_in_channels(m, 3) if rot and n.endswith("_trunk.main.0") else m, device)
                                for n, m in own}
            self._conv_sig = sig
        return self._conv_cache

    def _feat_first(self) -> bool:
        """BasicVSR trunks: keep the trunk input as [feat | x_i] (see propagate) when num_feat is a multiple of 8."""
        return getattr(self, "num_feat", 0) > 0 and self.num_feat % 8 == 0


def _rotate_in_channels(conv: nn.Conv2d, k: int) -> nn.Conv2d:
    """The same convolution for an input whose first ``k`` channels were moved to the end."""
    r = nn.Conv2d(conv.in_channels, conv.out_channels, conv.kernel_size, conv.stride, conv.padding, bias=conv.bias is not None)
    with torch.no_grad():
        r.weight.copy_(torch.cat([conv.weight[:, k:], conv.weight[:, :k]], 1))
        if conv.bias is not None:
            r.bias.copy_(conv.bias)
    return r


class BasicModule(nn.Module):
    """Parameter container of models/spynet_arch.py:10-25 (7x7 convs 8-32-64-32-16-2 with ReLU between)."""

    def __init__(self):
        super().__init__()
        self.basic_module = nn.Sequential(
            nn.Conv2d(8, 32, 7, 1, 3), nn.ReLU(inplace=False), nn.Conv2d(32, 64, 7, 1, 3), nn.ReLU(inplace=False),
            nn.Conv2d(64, 32, 7, 1, 3), nn.ReLU(inplace=False), nn.Conv2d(32, 16, 7, 1, 3), nn.ReLU(inplace=False),
            nn.Conv2d(16, 2, 7, 1, 3))


class SpyNet(nn.Module, _VideoPlanMixin):
    """SPyNet optical flow, constructor / forward / state_dict of models/spynet_arch.py:29-96.

    ``forward(ref, supp)`` -> flow (n,2,h,w) float32.  Flows, sampling positions and the pyramid are always fp32;
    with ``set_precision('bf16')`` only the 7x7 convolutions run on bf16 tensor-core operands.
    """

    def __init__(self, load_path=None):
        super().__init__()
        self.basic_module = nn.ModuleList([BasicModule() for _ in range(6)])
        if load_path:
            self.load_state_dict(torch.load(load_path, map_location=lambda storage, loc: storage)["params"])
        self.register_buffer("mean", torch.Tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1))
        self.register_buffer("std", torch.Tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1))

    @staticmethod
    def remap_mmedit_state_dict(sd):
        """``mmedit`` SPyNet checkpoints wrap every conv in a ConvModule (``...basic_module.I.conv.weight``, I in 0..4);
        the in-repo layout numbers the Sequential slots (``...basic_module.{0,2,4,6,8}.weight``).  SURVEY.md 8c."""
        out = {}
        for k, v in sd.items():
            parts = k.split(".")
            if len(parts) == 6 and parts[0] == "basic_module" and parts[2] == "basic_module" and parts[4] == "conv":
                k = ".".join([parts[0], parts[1], parts[2], str(2 * int(parts[3])), parts[5]])
            out[k] = v
        return out

    def forward(self, ref: torch.Tensor, supp: torch.Tensor) -> torch.Tensor:
        assert ref.size() == supp.size()
        _lib.require_cuda_tensor(ref, "ref")
        _lib.require_cuda_tensor(supp, "supp")
        L = _lib.lib()
        dev = ref.device
        n, _, h, w = ref.shape
        w_up = int(math.floor(math.ceil(w / 32.0) * 32.0))
        h_up = int(math.floor(math.ceil(h / 32.0) * 32.0))
        if h_up < 64 or w_up < 64:
            raise RuntimeError("Input and output sizes should be greater than 0 (SPyNet needs at least 33 pixels per side)")
        convs = self._convs(dev)
        st = _lib.current_stream_ptr(dev)
        # the normalisation constants travel as kernel arguments; read the (device) buffers back only when they change -- a
        # device->host copy per call is a stream sync and forbids CUDA-graph capture of the forward
        nsig = (self.mean.data_ptr(), self.mean._version, self.std.data_ptr(), self.std._version)
        if getattr(self, "_norm_sig", None) != nsig:
            mean = self.mean.detach().float().cpu().view(-1).tolist() + [0.0]
            inv_std = (1.0 / self.std.detach().float().cpu().view(-1)).tolist() + [1.0]
            self._norm_args = ((ctypes.c_float * 4)(*mean), (ctypes.c_float * 4)(*inv_std))
            self._norm_sig = nsig
        sub, mul = self._norm_args
        # ONE C-ABI call (b200sr_spynet_forward): the ~100 launches of the pyramid are sequenced in C on a cached workspace -- no
        # per-launch ctypes call, no torch.empty per intermediate
        hkey = "__spynet_handles__"
        if hkey not in convs:
            hs = [convs[f"basic_module.{level}.basic_module.{idx}"]._h for level in range(6) for idx in (0, 2, 4, 6, 8)]
            convs[hkey] = (ctypes.c_void_p * len(hs))(*[h.value for h in hs])
        prec = _lib.precision_code(self.precision)
        need = L.b200sr_spynet_workspace_bytes(n, h, w, prec)
        ws = self.__dict__.get("_ws")
        if ws is None or ws.numel() < need or ws.device != dev:
            ws = self.__dict__["_ws"] = torch.empty(need, dtype=torch.uint8, device=dev)
        ref, supp = ref.contiguous(), supp.contiguous()
        if supp.dtype != ref.dtype:
            supp = supp.to(ref.dtype)
        out = torch.empty((n, 2, h, w), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(L.b200sr_spynet_forward(convs[hkey], _ptr(ref), _ptr(supp), _lib.dtype_code(ref.dtype), _ptr(out), n, h, w, prec, sub, mul,
                                               _ptr(ws), ws.numel(), st))
        return out


class ResidualBlockNoBN(nn.Module):
    """Parameter container of models/basicvsr_arch_origin.py:115-137."""

    def __init__(self, num_feat=64, res_scale=1, pytorch_init=False):
        super().__init__()
        self.res_scale = res_scale
        self.conv1 = nn.Conv2d(num_feat, num_feat, 3, 1, 1, bias=True)
        self.conv2 = nn.Conv2d(num_feat, num_feat, 3, 1, 1, bias=True)
        self.relu = nn.ReLU(inplace=True)


class ConvResidualBlocks(nn.Module):
    """Parameter container of models/basicvsr_arch_origin.py:98-113 (conv + LeakyReLU(0.1) + num_block residual blocks)."""

    def __init__(self, num_in_ch=3, num_out_ch=64, num_block=15):
        super().__init__()
        self.main = nn.Sequential(nn.Conv2d(num_in_ch, num_out_ch, 3, 1, 1, bias=True), nn.LeakyReLU(negative_slope=0.1, inplace=True),
                                  nn.Sequential(*[ResidualBlockNoBN(num_feat=num_out_ch) for _ in range(num_block)]))


class _VsrBase(nn.Module, _VideoPlanMixin):
    num_feat: int

    def get_flow(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """models/basicvsr_arch_origin.py:42-51: both directions of all n-1 frame pairs, batched through SPyNet."""
        b, n, c, h, w = x.size()
        x = x.contiguous()
        self.spynet.set_precision(self.precision)
        # the reference's two SPyNet calls (x_1 -> x_2, x_2 -> x_1) as ONE batch of 2 b (n-1) pairs: every op is per sample, so the
        # flows are the same numbers, and the launch-bound coarse pyramid levels run once instead of twice
        # (torch.cat([x_1, x_2]) / torch.cat([x_2, x_1]) as per-clip contiguous copies: device-to-device memcpy nodes instead of cat kernels)
        m = b * (n - 1)
        ref, supp = x.new_empty((2 * m, c, h, w)), x.new_empty((2 * m, c, h, w))
        for k in range(b):
            lo, hi = k * (n - 1), (k + 1) * (n - 1)
            ref[lo:hi].copy_(x[k, :-1]), ref[m + lo:m + hi].copy_(x[k, 1:])
            supp[lo:hi].copy_(x[k, 1:]), supp[m + lo:m + hi].copy_(x[k, :-1])
        flows = self.spynet(ref, supp)
        flows_backward = flows[:m].view(b, n - 1, 2, h, w)
        flows_forward = flows[m:].view(b, n - 1, 2, h, w)
        return flows_forward, flows_backward

    def _trunk(self, convs, name: str, buf: torch.Tensor, num_block: int, out: Optional[torch.Tensor] = None, out_coff: int = 0) -> torch.Tensor:
        """``out``: a wider (n,h,w,C) tensor whose channels [out_coff, out_coff + num_feat) receive the features; returns that window."""
        first = convs[f"{name}.main.0"]
        # bf16 on the tcgen05 kernel: the trunk's private 64-channel tensors live in the planar-8 layout (TMA box rows of 512
        # contiguous bytes instead of one request per pixel and chunk); the last conv writes the NHWC features the callers read
        planar = (self.precision != "fp32" and num_block > 0 and buf.dtype == torch.bfloat16 and first.cout == 64 and first.tcgen05_ok()
                  and convs[f"{name}.main.2.0.conv1"].tcgen05_ok())
        if not planar:
            t = first(buf, self.precision, ACT_LRELU, out_dtype=self._act_dtype())
            for k in range(num_block):
                o = convs[f"{name}.main.2.{k}.conv1"](t, self.precision, ACT_RELU)
                t = convs[f"{name}.main.2.{k}.conv2"](o, self.precision, ACT_NONE, residual=t)
            if out is None:
                return t
            win = out[..., out_coff:out_coff + t.shape[-1]]
            win.copy_(t)
            return win
        # 2 * num_block + 1 dependent launches of a few microseconds each: the host side is ONE ABI call (b200sr_vsr_trunk_forward
        # sequences them in C), two planar-8 buffers (conv2 adds its residual in place: y = t + conv(o))
        n, h, w, cs = buf.shape
        dev = buf.device
        t = torch.empty((n, 8, h, w, 8), dtype=torch.bfloat16, device=dev)
        o = torch.empty_like(t)
        if out is None:
            out, out_coff = torch.empty((n, h, w, 64), dtype=torch.bfloat16, device=dev), 0
        hkey = "__handles__:" + name   # lives and dies with this set of conv handles
        if hkey not in convs:
            hs = [convs[f"{name}.main.2.{k}.{c}"]._h for k in range(num_block) for c in ("conv1", "conv2")]
            convs[hkey] = (ctypes.c_void_p * len(hs))(*[h.value for h in hs])
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().b200sr_vsr_trunk_forward_into(first._h, convs[hkey], num_block, buf.data_ptr(), cs, t.data_ptr(), o.data_ptr(),
                                                                out.data_ptr(), out.shape[-1], out_coff, n, h, w, _lib.current_stream_ptr(dev)))
        return out if out.shape[-1] == 64 else out[..., out_coff:out_coff + 64]

    def propagate(self, x: torch.Tensor, flows_forward: torch.Tensor, flows_backward: torch.Tensor):
        """The two recurrent loops (models/basicvsr_arch_origin.py:61-82) -> per-frame NHWC features (backward, forward).
        A clip's time axis is inherently sequential; parallelism comes from the clip batch ``b``."""
        _lib.require_cuda_tensor(x, "x")
        b, n, _, h, w = x.shape
        dev, adt, nf = x.device, self._act_dtype(), self.num_feat
        convs = self._convs(dev)
        nb = len(self.backward_trunk.main[2])
        cs = -(-(nf + 3) // 16) * 16
        x = x.contiguous()
        L = _lib.lib()

        # trunk input = cat([x_i, feat]) (:69,81).  When the feature count keeps the stores 16-byte aligned the buffer holds
        # [feat | x_i | 0] instead -- the first conv's handle was built with its input channels rotated the same way (_convs) -- so that
        # flow_warp writes the warped features in place
        feat_first = self._feat_first()
        xco = nf if feat_first else 0

        # per frame ONE tensor [backward features | forward features]: the two trunks write its halves in place, so the reconstruction's
        # torch.cat([out_l[i], feat_prop], dim=1) (:84) is free (_cat_feats)
        fused = [torch.empty((b, h, w, 2 * nf), dtype=adt, device=dev) for _ in range(n)]

        def run(trunk: str, order, flows, flow_index, coff):
            feats: List[Optional[torch.Tensor]] = [None] * n
            feat = None
            st = _lib.current_stream_ptr(dev)   # the stream this direction was forked onto
            # ONE trunk input per direction, zero-filled once: every frame rewrites its x_i channels and (from the second step on) its feature
            # channels; the first step needs zero features, the pad channels stay zero
            buf = torch.empty((b, h, w, cs), dtype=adt, device=dev)
            with torch.cuda.device(dev):
                _lib.check(L.b200sr_zero_async(_ptr(buf), buf.numel() * buf.element_size(), st))
            for step, i in enumerate(order):
                xi = x[:, i]
                with torch.cuda.device(dev):
                    _lib.check(L.b200sr_nchw3_to_nhwc(_ptr(xi), _lib.dtype_code(x.dtype), x.stride(0), _ptr(buf), _lib.dtype_code(adt),
                                                      b, h, w, cs, xco, st))
                if step > 0:
                    fl = flows[:, flow_index(i)].contiguous()
                    if feat_first:
                        flow_warp_nhwc(feat, fl, out=buf)          # straight into channels [0, nf) of the trunk input
                    elif nf % (4 if adt == torch.float32 else 8) == 0:
                        buf[..., 3:3 + nf] = flow_warp_nhwc(feat, fl)
                    else:   # odd feature counts (the fork's BasicVSR only runs for num_feat = 3): the reference-layout NCHW kernel, fp32
                        wv = flow_warp(feat.float().permute(0, 3, 1, 2).contiguous(), fl.permute(0, 2, 3, 1))
                        buf[..., 3:3 + nf] = wv.permute(0, 2, 3, 1).to(adt)
                feat = self._trunk(convs, trunk, buf, nb, out=fused[i], out_coff=coff)
                feats[i] = feat
            return feats

        # the two directions are independent recurrences of small launches (one 180x320 frame is 253 tiles on 148 SMs): the
        # forward one runs on a side stream (fork / join by events, CUDA-graph capturable) so that they fill each other's tails
        main = torch.cuda.current_stream(dev)
        side = main if os.environ.get("B200SR_ONE_STREAM") == "1" else self._side_stream(dev)   # developer A/B switch
        # ... each on half of the SMs (grids of the one-frame trunk launches capped): 14.7 -> 12.4 ms per 15-frame clip
        half = 0 if side is main or os.environ.get("B200SR_TRUNK_FULL_GRID") == "1" else torch.cuda.get_device_properties(dev).multi_processor_count // 2
        if half and os.environ.get("B200SR_TRUNK_CTAS"):     # developer experiment: another grid cap for the one-frame trunk launches
            half = int(os.environ["B200SR_TRUNK_CTAS"])
        for name, c in convs.items():
            if name.startswith(("backward_trunk.", "forward_trunk.")):
                c.set_max_ctas(half)
        side.wait_stream(main)
        back = run("backward_trunk", range(n - 1, -1, -1), flows_backward, lambda i: i, 0)
        with torch.cuda.stream(side):
            fwd = run("forward_trunk", range(0, n), flows_forward, lambda i: i - 1, nf)
        main.wait_stream(side)
        for f in fused:
            f.record_stream(side)
        return back, fwd

    # ---- tail of the fork's BasicVSR and of MotionVectorVSR: lrelu(fusion) -> conv_last = ConvTranspose2d(2nf, 3, 5, stride 4) -> bilinear
    #      resize to (height, weight) + bilinear base (models/basicvsr_arch.py:93-102, models/mvvsr_arch.py:95-104)
    def _deconv_handle(self, device):
        """(generic 3x3 handle of the transposed conv, [tcgen05 halves] or None).  bf16 with 2 nf = 128 input channels and 3 x 16 <= 64 phase
        outputs: the 3x3 conv runs as TWO tcgen05 3x3 convs over the two 64-channel windows of the fusion output (outputs padded to 64), the
        second one adding the first one's result as its residual -- 2 x ~8 us instead of 105 us per 180 x 320 frame on the generic mma.sync
        kernel (the tcgen05 3x3 kernel takes 64..80 input channels)."""
        sig = (str(device),) + tuple((p.data_ptr(), p._version) for p in self.conv_last.parameters())
        if getattr(self, "_tail_sig", None) != sig:
            conv = _transposed_s4k5_as_conv3x3(self.conv_last)
            halves = None
            if conv.in_channels == 128 and conv.out_channels <= 64:
                halves = []
                for k in range(2):
                    c = nn.Conv2d(64, 64, 3, 1, 1, bias=True)
                    with torch.no_grad():
                        c.weight.zero_(), c.bias.zero_()
                        c.weight[:conv.out_channels].copy_(conv.weight[:, 64 * k:64 * (k + 1)])
                        if k == 0:
                            c.bias[:conv.out_channels].copy_(conv.bias)
                    halves.append(_ConvHandle(c, device))
                if not all(hd.tcgen05_ok() for hd in halves):
                    halves = None
            self._tail_handle, self._tail_halves, self._tail_sig = _ConvHandle(conv, device), halves, sig
        return self._tail_handle, self._tail_halves

    def _deconv_tail(self, x: torch.Tensor, back, fwd, height: int, weight: int) -> torch.Tensor:
        b, n, _, h, w = x.shape
        dev, p = x.device, self.precision
        convs, (tail, halves) = self._convs(dev), self._deconv_handle(dev)
        L, st = _lib.lib(), _lib.current_stream_ptr(dev)
        out = torch.empty((b, n, 3, height, weight), dtype=torch.float32, device=dev)
        for i in range(n):
            o = convs["fusion"](_cat_feats(back[i], fwd[i]), p, ACT_LRELU)
            op = torch.empty((b, h + 1, w + 1, o.shape[-1]), dtype=o.dtype, device=dev)
            with torch.cuda.device(dev):                                     # F.pad(o, (0, 0, 0, 1, 0, 1)): one zero row / column, the fifth tap's outputs
                _lib.check(L.b200sr_pad_bottom_right_async(_ptr(o), _ptr(op), b, h, w, o.shape[-1] * o.element_size(), st))
            o = op
            if halves is not None and p != "fp32" and o.dtype == torch.bfloat16:
                ta = halves[0](o, p, ACT_NONE, x_coff=0)                     # (b, h+1, w+1, 64): channels 0..63 of the fusion output
                t = halves[1](o, p, ACT_NONE, x_coff=64, residual=ta)        # + channels 64..127
                tdt = _lib.BF16
            else:
                t = tail(o, p, ACT_NONE, out_dtype=torch.float32)           # (b, h+1, w+1, 3*16)
                tdt = _lib.F32
            xi = x[:, i]
            with torch.cuda.device(dev):                                     # shuffle(4) + crop + resize + base + add: one kernel
                _lib.check(L.b200sr_vsr_deconv_tail(_ptr(t), tdt, t.shape[-1], _ptr(xi), _lib.dtype_code(x.dtype), x.stride(0), _ptr(out[:, i]),
                                                    out.stride(0), b, h, w, height, weight, st))
        return out

    def _side_stream(self, dev) -> "torch.cuda.Stream":
        key = str(dev)
        cache = self.__dict__.setdefault("_side_streams", {})
        if key not in cache:
            cache[key] = torch.cuda.Stream(device=dev)
        return cache[key]


class BasicVSR_origin(_VsrBase):
    """Canonical BasicVSR x4: constructor / forward / state_dict of models/basicvsr_arch_origin.py:10-96.

    ``forward(x, height, weight)``: x (b,n,3,h,w) -> (b,n,3,height,weight) float32.
    """

    def __init__(self, num_feat=64, num_block=15, spynet_path=None):
        super().__init__()
        self.num_feat = num_feat
        self.spynet = SpyNet(spynet_path)
        self.scale = 4
        self.backward_trunk = ConvResidualBlocks(num_feat + 3, num_feat, num_block)
        self.forward_trunk = ConvResidualBlocks(num_feat + 3, num_feat, num_block)
        self.fusion = nn.Conv2d(num_feat * 2, num_feat, 1, 1, 0, bias=True)
        self.upconv1 = nn.Conv2d(num_feat, num_feat * 4, 3, 1, 1, bias=True)
        self.upconv2 = nn.Conv2d(num_feat, 64 * 4, 3, 1, 1, bias=True)
        self.conv_hr = nn.Conv2d(64, 64, 3, 1, 1)
        self.conv_last = nn.Conv2d(64, 3, 3, 1, 1)
        self.pixel_shuffle = nn.PixelShuffle(2)
        self.lrelu = nn.LeakyReLU(negative_slope=0.1, inplace=True)

    def forward(self, x: torch.Tensor, height: int, weight: int) -> torch.Tensor:
        _lib.require_cuda_tensor(x, "x")
        flows_forward, flows_backward = self.get_flow(x)
        back, fwd = self.propagate(x, flows_forward, flows_backward)
        b, n, _, h, w = x.shape
        dev = x.device
        convs = self._convs(dev)
        L, st, p = _lib.lib(), _lib.current_stream_ptr(dev), self.precision
        x = x.contiguous()
        out = torch.empty((b, n, 3, height, weight), dtype=torch.float32, device=dev)
        for i in range(n):
            o = convs["fusion"](_cat_feats(back[i], fwd[i]), p, ACT_LRELU)
            o = convs["upconv1"](o, p, ACT_LRELU, shuffle=2)          # lrelu(pixel_shuffle(conv)) == shuffle(lrelu(conv))
            hr_planar = p != "fp32" and convs["upconv2"].tcgen05_ok() and convs["conv_hr"].tcgen05_ok() and convs["conv_hr"].cin == 64
            # conv_last + bilinear base in one tcgen05 launch (fp32 accumulators straight to the fp32 NCHW frame, no bf16 round trip)
            fused_last = (hr_planar and x.dtype == torch.float32 and convs["conv_last"].tcgen05_ok() and convs["conv_last"].cout == 3
                          and convs["conv_hr"].cout == 64)
            o = convs["upconv2"](o, p, ACT_LRELU, shuffle=2, y_planar=hr_planar)     # 720p tensors planar-8 between the tcgen05 convs
            o = convs["conv_hr"](o, p, ACT_LRELU, x_planar=hr_planar, y_planar=fused_last)
            direct = (height, weight) == (4 * h, 4 * w)
            hr = out[:, i] if direct else torch.empty((b, 3, 4 * h, 4 * w), dtype=torch.float32, device=dev)
            xi = x[:, i]
            with torch.cuda.device(dev):
                if fused_last:
                    _lib.check(L.b200sr_vsr_conv_last_base(convs["conv_last"]._h, _ptr(o), 1, 64, 0, _ptr(xi), x.stride(0), _ptr(hr), hr.stride(0),
                                                           b, 4 * h, 4 * w, st))
                else:
                    o = convs["conv_last"](o, p, ACT_NONE)
                    _lib.check(L.b200sr_vsr_base_add(_ptr(o), _lib.dtype_code(o.dtype), o.shape[-1], _ptr(xi), _lib.dtype_code(x.dtype), x.stride(0),
                                                     _ptr(hr), hr.stride(0), b, h, w, st))
                if not direct:   # F.interpolate(out, size=(height, weight), mode='bilinear'), :93
                    res = torch.empty((b, 3, height, weight), dtype=torch.float32, device=dev)
                    _lib.check(L.b200sr_resize_bilinear_nchw(_ptr(hr), _lib.F32, _ptr(res), b, 3, 4 * h, 4 * w, height, weight, 0, None, None, st))
                    out[:, i] = res
        return out


class BasicVSR(_VsrBase):
    """The fork's light BasicVSR (models/basicvsr_arch.py:10-105): same constructor and state_dict; ``get_flow`` and the
    propagation loops and the ConvTranspose2d tail run on the B200 path.  Its ``forward`` is broken as committed for
    ``num_feat != 3`` -- ``conv_last`` yields ``num_feat`` channels that are added to a 3-channel bilinear base (:96-100) -- and that
    ``RuntimeError`` is reproduced rather than "fixed" (SURVEY.md 0-3); with ``num_feat == 3`` it runs, here as there."""

    def __init__(self, num_feat=64, num_block=15, spynet_path=None):
        super().__init__()
        self.num_feat = num_feat
        self.spynet = SpyNet(spynet_path)
        self.scale = 4
        self.backward_trunk = ConvResidualBlocks(num_feat + 3, num_feat, num_block)
        self.forward_trunk = ConvResidualBlocks(num_feat + 3, num_feat, num_block)
        self.fusion = nn.Conv2d(num_feat * 2, num_feat * 2, 1, 1, 0, bias=True)
        self.upconv1 = nn.Conv2d(num_feat, num_feat * 4, 3, 1, 1, bias=True)
        self.upconv2 = nn.Conv2d(num_feat, num_feat * 4, 3, 1, 1, bias=True)
        self.conv_last = nn.ConvTranspose2d(num_feat * 2, num_feat, 5, stride=self.scale)
        self.conv_hr = nn.Conv2d(num_feat, 3, 3, 1, 1)
        self.pixel_shuffle = nn.PixelShuffle(2)
        self.lrelu = nn.LeakyReLU(negative_slope=0.1, inplace=True)

    def forward(self, x: torch.Tensor, height: int = 1080, weight: int = 1920) -> torch.Tensor:
        """models/basicvsr_arch.py:56-105.  As committed the reference only runs for ``num_feat == 3`` (``conv_last`` yields ``num_feat``
        channels that are added to the 3-channel bilinear base, :96-101); for any other width its ``out += base`` raises, and so does this."""
        if self.num_feat != 3:
            raise RuntimeError(f"The size of tensor a ({self.num_feat}) must match the size of tensor b (3) at non-singleton dimension 1")
        _lib.require_cuda_tensor(x, "x")
        flows_forward, flows_backward = self.get_flow(x)
        back, fwd = self.propagate(x.contiguous(), flows_forward, flows_backward)
        return self._deconv_tail(x.contiguous(), back, fwd, height, weight)


def _transposed_s4k5_as_conv3x3(deconv: nn.ConvTranspose2d) -> nn.Conv2d:
    """ConvTranspose2d(cin, cout, 5, stride=4) (models/mvvsr_arch.py:38) as an ordinary 3x3 convolution + PixelShuffle(4):
    out[4y+i, 4x+j] = sum_ci x[y, x] w[ci, c, i, j]  (+ x[y-1, x] w[.., 4, j] if i == 0) (+ x[y, x-1] w[.., i, 4] if j == 0)
    (+ x[y-1, x-1] w[.., 4, 4] if i == j == 0) -- the fifth tap of a stride-4 kernel only overlaps the next cell's first row / column.
    Evaluated on the input zero-padded by one row and column (bottom / right) it yields all (4h+1) x (4w+1) outputs."""
    wt = deconv.weight.detach().float().cpu()                      # (cin, cout, 5, 5)
    cin, cout = int(wt.shape[0]), int(wt.shape[1])
    conv = nn.Conv2d(cin, cout * 16, 3, 1, 1, bias=True)
    w = torch.zeros(cout, 4, 4, cin, 3, 3)
    core = wt.permute(1, 2, 3, 0)                                   # (cout, 5, 5, cin)
    w[:, :, :, :, 1, 1] = core[:, :4, :4]
    w[:, 0, :, :, 0, 1] = core[:, 4, :4]
    w[:, :, 0, :, 1, 0] = core[:, :4, 4]
    w[:, 0, 0, :, 0, 0] = core[:, 4, 4]
    with torch.no_grad():
        conv.weight.copy_(w.reshape(cout * 16, cin, 3, 3))
        b = deconv.bias.detach().float().cpu() if deconv.bias is not None else torch.zeros(cout)
        conv.bias.copy_(b.view(cout, 1).expand(cout, 16).reshape(-1))
    return conv


class MotionVectorVSR(_VsrBase):
    """models/mvvsr_arch.py:10-109: BasicVSR whose flows are the codec's motion vectors carried in input channels 3:5 (SPyNet is
    constructed -- its parameters are in the state_dict -- but never run).  ``forward(x_, height, weight)``: x_ (b,n,5,h,w) ->
    (b,n,3,height,weight) float32.  Propagation = the BasicVSR trunks (tcgen05 convs in bf16); the tail ConvTranspose2d(2nf, 3, 5,
    stride 4) runs as a 3x3 convolution + PixelShuffle(4) (``_transposed_s4k5_as_conv3x3``)."""

    def __init__(self, num_feat=64, num_block=15, spynet_path=None):
        super().__init__()
        self.num_feat = num_feat
        self.spynet = SpyNet(spynet_path)
        self.scale = 4
        self.backward_trunk = ConvResidualBlocks(num_feat + 3, num_feat, num_block)
        self.forward_trunk = ConvResidualBlocks(num_feat + 3, num_feat, num_block)
        self.fusion = nn.Conv2d(num_feat * 2, num_feat * 2, 1, 1, 0, bias=True)
        self.upconv1 = nn.Conv2d(num_feat, num_feat * 4, 3, 1, 1, bias=True)     # unused by forward (as in the reference)
        self.upconv2 = nn.Conv2d(num_feat, num_feat * 4, 3, 1, 1, bias=True)
        self.conv_hr = nn.Conv2d(num_feat, num_feat, 3, 1, 1)
        self.conv_last = nn.ConvTranspose2d(num_feat * 2, 3, 5, stride=self.scale)
        self.pixel_shuffle = nn.PixelShuffle(2)
        self.lrelu = nn.LeakyReLU(negative_slope=0.1, inplace=True)

    def forward(self, x_: torch.Tensor, height: int = 1080, weight: int = 1920) -> torch.Tensor:
        _lib.require_cuda_tensor(x_, "x_")
        x = x_[:, :, :3].contiguous()
        flows_forward = x_[:, 1:, 3:].float().contiguous()                    # mv[:, 1:]            (:65-66)
        flows_backward = flows_forward * (-1)
        back, fwd = self.propagate(x, flows_forward, flows_backward)
        return self._deconv_tail(x, back, fwd, height, weight)
