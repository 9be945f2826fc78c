#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native super-resolution forward path.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun, one rank per GPU)
    python bench.py --impl reference ...                      (the reference's CPU forward on the host cores)

Workload (BASELINE.json configs[1], the config the metric is quoted on that fits one GPU):
    WDSR-B x4, num_blocks=16, num_residual_units=24, reference seeded init, batch 64 of 96x96 LR patches
    -> 64 x 3 x 384 x 384, bf16 arithmetic.  One "step" = one forward over one batch, per GPU (weak scaling:
    every rank runs a full batch, no data-path collective).
Metric: output Mpixels/s (spatial output pixels N*sH*sW, not x3), whole job.

Rank 0 prints ONE JSON line (keys described in DESIGN.md "Measurement").
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time
import types

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SCALE, NB, NRU = 4, 16, 24
BATCH, LR = 64, 96
FLOP_PER_LR_PX_BLOCK = 2 * (24 * 144 + 144 * 20 + 9 * 20 * 24)      # 21,312
BYTES_PER_LR_PX_BLOCK = 2 * 24 * 2                                    # read + write the bf16 trunk once = 96
# measured DRAM bytes of ONE block launch at this workload (ncu --set full, profiles/r02_block_final_ncu.md; round 1's capture,
# profiles/r01_block_tcgen05_v2_ncu.md, gave 28,391,936 + 33,792): the output stays in L2 for the next block
NCU_DRAM_BYTES_PER_BLOCK_LAUNCH = 28_395_008 + 85_504   # dram__bytes_read.sum + dram__bytes_write.sum of one launch (profiles/r02_block_final_ncu.md)
WORKLOAD = "cfg2: WDSR-B x4 nb16 nru24 (reference seeded init), batch 64 x 3x96x96 LR -> 3x384x384, bf16"


def params():
    return types.SimpleNamespace(image_mean=0.5, num_channels=3, scale=SCALE, num_blocks=NB, num_residual_units=NRU,
                                 width_search=False, pretrained=False)


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            p = json.load(fh)
        return {"hbm_gbs": float(p["hbm_gbs"]), "bf16_tflops": float(p["bf16_tflops"]),
                "bf16_tflops_sustained": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), "source": "measured"}
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons sampled through NVML during the timed region (same source as nvidia-smi)."""

    def __init__(self, index: int, period_s: float = 0.004):
        super().__init__(daemon=True)
        self.index, self.period = index, period_s
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.nv = None

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {getattr(nv, n): n for n in dir(nv) if n.startswith("nvmlClocksEventReason") or n.startswith("nvmlClocksThrottleReason")}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if isinstance(bit, int) and bit and (mask & bit) == bit and bin(bit).count("1") == 1:
                        self.reasons.add(nm.replace("nvmlClocksEventReason", "").replace("nvmlClocksThrottleReason", ""))
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        norm = set()
        for r in self.reasons:
            r = r.lower()
            if "none" in r or "all" in r or "idle" in r:
                continue
            norm.add({"swpowercap": "sw_power_cap", "hwslowdown": "hw_slowdown", "hwthermalslowdown": "hw_thermal_slowdown",
                      "swthermalslowdown": "sw_thermal_slowdown", "hwpowerbrakeslowdown": "hw_power_brake_slowdown",
                      "applicationsclockssetting": "applications_clocks_setting", "syncboost": "sync_boost",
                      "displayclocksetting": "display_clock_setting"}.get(r, r))
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(norm), "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the reference's own CPU forward (torch CPU fp32, oneDNN) restated in oracle/port.py
# ---------------------------------------------------------------------------------------------------------
def cpu_forward_rate(sample_patches: int, reps: int, warmup: int):
    """Output Mpix/s of the reference's CPU forward on `sample_patches` patches of the workload."""
    from oracle import port
    import mobilesuperresolution_b200 as sr
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    sd = {k: v.clone() for k, v in sr.BASIC_MODEL(params()).state_dict().items()}
    x = torch.rand(BATCH, 3, LR, LR, generator=torch.Generator().manual_seed(1234))[:sample_patches]
    times = []
    with torch.no_grad():
        for i in range(warmup + reps):
            t0 = time.perf_counter()
            port.basic_model_forward(sd, x, SCALE)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    out_mpix = sample_patches * (LR * SCALE) ** 2 / 1e6
    return out_mpix / statistics.median(times), statistics.median(times), cores


def run_reference(args, rank):
    if rank != 0:
        return
    sample = 16
    # each "step" = the CPU forward over a 16-patch sample of the 64-patch batch (bounded so K steps end in minutes)
    steps, warmup = max(1, min(args.steps, 5)), max(1, min(args.warmup, 1))
    rate, sec, cores = cpu_forward_rate(sample, steps, warmup)
    line = {"impl": "reference", "metric": "output Mpixels/sec", "value": rate, "unit": "Mpix/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": sec * 1e3 * (BATCH / sample), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "note": "reference CPU path is fp32 (torch/oneDNN); ms_per_step scaled to the full 64-patch batch"},
            "cpu_baseline": {"value": rate, "unit": "Mpix/s", "cores": cores, "kind": "port",
                             "sample": f"{sample} of the {BATCH} patches per step, median of {steps} steps"},
            "e2e": {"value": rate, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------
def l2_flusher(device):
    buf = torch.empty(256 << 20, dtype=torch.uint8, device=device)

    def flush():
        buf.fill_(0)          # writes 256 MiB > 126 MB L2
    return flush


P1_WIDTHS = [(9, 91, 14), (9, 94, 10), (9, 107, 12), (9, 110, 13), (9, 115, 12), (9, 94, 12), (9, 115, 17), (9, 116, 16)]


def _graph_us(module, x, *fargs, reps=10, warm=2):
    """One forward as a CUDA graph (the serving form, mobilesuperresolution_b200.Graphed): microseconds per replay, CUDA events
    on the replaying stream."""
    import mobilesuperresolution_b200 as sr
    g = sr.Graphed(module, x, *fargs, warmup=warm)
    st = torch.cuda.current_stream()
    for _ in range(2):
        g(x)
    st.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st)
    for _ in range(reps):
        g.graph.replay()
    b.record(st)
    st.synchronize()
    return a.elapsed_time(b) / reps * 1e3, g


def north_star_configs(dev, rank, world, peaks, shard):
    """The other configurations BASELINE.json's north star is judged on (SURVEY.md 8d ceilings), each through the public module as
    one CUDA graph, inputs resident, CUDA-event timed; per-rank work is fixed (weak), times are the max over ranks.  Side numbers:
    cfg2 stays the headline."""
    import tempfile
    import mobilesuperresolution_b200 as sr
    from mobilesuperresolution_b200 import video
    out = {}
    dv = str(dev) if world > 1 else "cpu"

    def P(scale):
        return types.SimpleNamespace(image_mean=0.5, num_channels=3, scale=scale, num_blocks=NB, num_residual_units=NRU,
                                     width_search=False, pretrained=False)

    def reduce_us(us):
        return shard.max_over_ranks(us, dv)

    with torch.no_grad():
        # -- dense WDSR-B x4 360p -> 1440p (the north-star target: >= 60 % of the 16.3 k frames/s ceiling)
        torch.manual_seed(0)
        m = sr.BASIC_MODEL(P(4)).eval().to(dev).set_precision("bf16")
        ceil_fps = 1.0 / max(370224 * 230400 / (peaks["bf16_tflops_sustained"] * 1e12), 1740 * 230400 / (peaks["hbm_gbs"] * 1e9))
        for b in (1, 8):
            x = torch.rand(b, 3, 360, 640, device=dev).bfloat16()
            us, g = _graph_us(m, x)
            us = reduce_us(us)
            out[f"dense_360p_b{b}"] = {"us_per_frame": us / b, "frames_per_s": world * b / us * 1e6, "frames_per_s_per_gpu": b / us * 1e6,
                                       "ceiling_frames_per_s_per_gpu": ceil_fps, "frac_of_ceiling": b / us * 1e6 / ceil_fps,
                                       "out_mpix_s": world * b * 1440 * 2560 / us}
            del g
        # -- cfg3: searched widths P1 (pixelshuffle.onnx) at x4, 360p frames
        f = tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False)
        f.write(repr((list(range(len(P1_WIDTHS))), [list(w) for w in P1_WIDTHS])) + "\n")
        f.close()
        mp = sr.Model(4, f.name).eval().to(dev).set_precision("bf16")
        os.unlink(f.name)
        ceil_p1 = 1.0 / max(70284 * 230400 / (peaks["bf16_tflops_sustained"] * 1e12), 432 * 230400 / (peaks["hbm_gbs"] * 1e9))
        for b in (1, 8):
            x = torch.rand(b, 3, 360, 640, device=dev).bfloat16()
            us, g = _graph_us(mp, x)
            us = reduce_us(us)
            out[f"cfg3_P1_b{b}"] = {"us_per_frame": us / b, "frames_per_s": world * b / us * 1e6, "frames_per_s_per_gpu": b / us * 1e6,
                                    "ceiling_frames_per_s_per_gpu": ceil_p1, "frac_of_ceiling": b / us * 1e6 / ceil_p1}
            del g
        # -- the fork's NAS_MODEL as committed (models/wdsr_b.py:30-137: head -> 16 x Split_Block -> tail, all blocks kept) at 360p x4
        torch.manual_seed(0)
        pn = P(4)
        pn.width_search = True
        mn = sr.NAS_MODEL(pn).eval().to(dev).set_precision("bf16")
        for b in (1, 8):
            x = torch.rand(b, 3, 360, 640, device=dev).bfloat16()
            us, g = _graph_us(mn, x)
            us = reduce_us(us)
            out[f"nas_fork_360p_b{b}"] = {"us_per_frame": us / b, "frames_per_s": world * b / us * 1e6, "frames_per_s_per_gpu": b / us * 1e6,
                                          "kept_blocks": mn.get_current_blocks(), "out_mpix_s": world * b * 1440 * 2560 / us}
            del g
        del mn
        # -- cfg5: WDSR-B x2 1080p -> 2160p, frames sharded over the ranks (4 frames per rank, no collective)
        torch.manual_seed(0)
        m2 = sr.BASIC_MODEL(P(2)).eval().to(dev).set_precision("bf16")
        lo, hi = shard.shard_slice(4 * world, rank, world)
        x = torch.rand(hi - lo, 3, 1080, 1920, device=dev).bfloat16()
        us, g = _graph_us(m2, x, reps=5)
        us = reduce_us(us)
        ceil5 = 1.0 / max(349272 * 2073600 / (peaks["bf16_tflops_sustained"] * 1e12), 1668 * 2073600 / (peaks["hbm_gbs"] * 1e9))
        out["cfg5_1080p_x2"] = {"frames_total": 4 * world, "frames_per_rank": hi - lo, "ms_per_frame_per_gpu": us / (hi - lo) / 1e3,
                                "frames_per_s": 4 * world / us * 1e6, "frames_per_s_per_gpu": (hi - lo) / us * 1e6,
                                "ceiling_frames_per_s_per_gpu": ceil5, "frac_of_ceiling": (hi - lo) / us * 1e6 / ceil5,
                                "out_mpix_s": 4 * world * 2160 * 3840 / us}
        del g, x
        # -- cfg4: BasicVSR_origin(64, 30), one 15-frame 180x320 clip per rank -> 720x1280 (11.23 TFLOP per clip)
        mv = video.BasicVSR_origin(64, 30).to(dev).eval().set_precision("bf16")
        clip = torch.rand(1, 15, 3, 180, 320, device=dev)
        us, g = _graph_us(mv, clip, 720, 1280, reps=5, warm=1)
        us = reduce_us(us)
        roof_ms = 11.23e12 / (peaks["bf16_tflops_sustained"] * 1e12) * 1e3
        out["cfg4_clip15"] = {"ms_per_clip": us / 1e3, "frames_per_s": world * 15 / us * 1e6, "tflops_per_gpu": 11.23e12 / us / 1e6,
                              "roofline_ms_per_clip": roof_ms, "frac_of_roofline": roof_ms / (us / 1e3)}
        del g
        # -- the fork's video model: MotionVectorVSR(64, 15), motion vectors in input channels 3:5, same clip size (models/mvvsr_arch.py:10-109)
        mm = video.MotionVectorVSR(64, 15).to(dev).eval().set_precision("bf16")
        xm = torch.rand(1, 15, 5, 180, 320, device=dev)
        xm[:, :, 3:] = (xm[:, :, 3:] - 0.5) * 8
        us, g = _graph_us(mm, xm, 720, 1280, reps=5, warm=1)
        us = reduce_us(us)
        out["mvvsr_clip15"] = {"ms_per_clip": us / 1e3, "frames_per_s": world * 15 / us * 1e6}
        del g, mm
        # -- sustained leg: the headline forward replayed back to back for >= 3 s (power-limited steady state), clocks sampled
        torch.manual_seed(0)
        mh = sr.BASIC_MODEL(params()).eval().to(dev).set_precision("bf16")
        xh = torch.rand(BATCH, 3, LR, LR, device=dev).bfloat16()
        us1, g = _graph_us(mh, xh, reps=20)
        n_rep = max(100, int(3.2e6 / us1))
        idx = int(os.environ["CUDA_VISIBLE_DEVICES"].split(",")[dev.index]) if "CUDA_VISIBLE_DEVICES" in os.environ else dev.index
        sampler = ClockSampler(idx, period_s=0.02)
        st = torch.cuda.current_stream()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        shard.barrier()
        st.synchronize()
        sampler.start()
        a.record(st)
        for _ in range(n_rep):
            g.graph.replay()
        b.record(st)
        st.synchronize()
        clk = sampler.stop()
        ms = shard.max_over_ranks(a.elapsed_time(b), dv)
        out["sustained_cfg2"] = {"seconds": ms / 1e3, "replays": n_rep, "ms_per_step": ms / n_rep, "ms_per_step_burst": us1 / 1e3,
                                 "value": world * n_rep * BATCH * (LR * SCALE) ** 2 / 1e6 / (ms / 1e3), "unit": "Mpix/s",
                                 "l2": "not flushed (back-to-back replays; the 3.5 MB input and 57 MB output are overwritten every step)",
                                 "clocks": clk}
        del g
    out["how"] = "each configuration as one CUDA graph of the public module's forward (Graphed), resident inputs, CUDA events, max over ranks"
    return out


def run_b200(args, rank, local_rank, world):
    import mobilesuperresolution_b200 as sr
    from mobilesuperresolution_b200 import shard
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    shard.init_distributed("nccl")
    K, W = args.steps, max(args.warmup, 3)
    peaks = load_peaks()

    torch.manual_seed(0)
    model = sr.BASIC_MODEL(params()).eval().to(dev).set_precision("bf16")
    plan = model.prepare(dev)
    x_cpu = torch.rand(BATCH, 3, LR, LR, generator=torch.Generator().manual_seed(1234 + rank))
    x_dev = x_cpu.to(dev).bfloat16().contiguous()
    y_dev = torch.empty(BATCH, 3, LR * SCALE, LR * SCALE, dtype=torch.bfloat16, device=dev)
    flush = l2_flusher(dev)
    out_mpix_step = BATCH * (LR * SCALE) ** 2 / 1e6

    def step():
        plan.forward(x_dev, "bf16", out=y_dev)

    # ---- device-resident timing: K steps, L2 flushed before each, CUDA events on the launching stream
    for _ in range(W):
        flush()
        step()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank if "CUDA_VISIBLE_DEVICES" not in os.environ else
                           int(os.environ["CUDA_VISIBLE_DEVICES"].split(",")[local_rank]))
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    shard.barrier()
    torch.cuda.synchronize()
    sampler.start()
    t_wall0 = time.perf_counter()
    for a, b in ev:
        flush()
        a.record()
        step()
        b.record()
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t_wall0
    shard.barrier()
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    launches = K * plan.launches_per_forward()
    total_ms = shard.max_over_ranks(dev_ms, str(dev) if world > 1 else "cpu")
    ms_per_step = total_ms / K
    value = world * out_mpix_step / (ms_per_step / 1e3)

    # ---- dominant kernel (fused residual block) timed alone with events: 16 launches per step on real trunk data
    trunk = plan.head_internal(x_dev, "bf16")   # the kernels' own trunk layout (planar-8 on the tcgen05 path)
    tb = torch.empty_like(trunk)
    for _ in range(3):
        plan.block_internal(0, trunk, "bf16", out=tb)
    torch.cuda.synchronize()
    kev = []
    for _ in range(min(K, 20)):
        flush()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        src, dst = trunk, tb
        for i in range(NB):
            from mobilesuperresolution_b200 import _lib
            _lib.check(_lib.lib().b200sr_wdsr_block(plan.handle, i, src.data_ptr(), dst.data_ptr(), BATCH, LR, LR, _lib.BF16,
                                                    _lib.current_stream_ptr(dev)))
            src, dst = dst, src
        b.record()
        kev.append((a, b))
    torch.cuda.synchronize()
    clocks = sampler.stop()
    blk_ms = statistics.median(a.elapsed_time(b) for a, b in kev) / NB
    lr_px = BATCH * LR * LR
    ach_tflops = lr_px * FLOP_PER_LR_PX_BLOCK / (blk_ms * 1e-3) / 1e12
    ach_gbs = lr_px * BYTES_PER_LR_PX_BLOCK / (blk_ms * 1e-3) / 1e9
    # The kernel is timed ALONE (16 launches, ~0.4 ms of GPU work at boost clocks): the peak that applies is the BURST bf16
    # figure of MEASURED_PEAKS.json; the sustained one is carried next to it.  The block sits on the ridge, so the roofline time
    # is max(FLOP / peak, bytes / HBM) and `roof_frac` = T_roof / T_measured.
    t_tensor = lr_px * FLOP_PER_LR_PX_BLOCK / (peaks["bf16_tflops"] * 1e12)
    t_hbm = lr_px * BYTES_PER_LR_PX_BLOCK / (peaks["hbm_gbs"] * 1e9)
    impl = os.environ.get("B200SR_BLOCK_IMPL", "tc5")
    roofline = {"kernel": f"fused residual block ({impl}: "
                          f"{ {'rs': 'wdsr_block_rs_kernel', 'rh': 'wdsr_block_rh_kernel'}.get(impl, 'wdsr_block_tc5p_kernel') }, tcgen05)",
                "bound": "tensor", "achieved": ach_tflops,
                "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": ach_tflops / peaks["bf16_tflops"],
                "peak_sustained": peaks["bf16_tflops_sustained"], "frac_sustained": ach_tflops / peaks["bf16_tflops_sustained"],
                "roof_frac": max(t_tensor, t_hbm) / (blk_ms * 1e-3),
                "traffic": NCU_DRAM_BYTES_PER_BLOCK_LAUNCH, "traffic_source": "profiles/r02_block_final_ncu.md (dram__bytes_read.sum + dram__bytes_write.sum, one ncu --set full capture of the final kernel; r01_block_tcgen05_v2_ncu.md gave the same 28.4 MB)",
                "peak_source": peaks["source"] + " (burst bf16 GEMM: the kernel is timed alone, 16 launches between two events; the sustained figure is peak_sustained)",
                "us_per_launch": blk_ms * 1e3, "algorithmic_flop_per_launch": lr_px * FLOP_PER_LR_PX_BLOCK,
                "algorithmic_bytes_per_launch": lr_px * BYTES_PER_LR_PX_BLOCK,
                "hbm": {"achieved": ach_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach_gbs / peaks["hbm_gbs"]},
                "block_share_of_step": NB * blk_ms / ms_per_step}

    # ---- end to end through the host-buffer C-ABI entry: pinned host in -> H2D -> forward -> D2H -> pinned host out
    depth = 2
    xs_h = [x_cpu.bfloat16().pin_memory() for _ in range(depth)]
    ys_h = [torch.empty(BATCH, 3, LR * SCALE, LR * SCALE, dtype=torch.bfloat16).pin_memory() for _ in range(depth)]
    plans = [plan] + [model._build_plan(dev) for _ in range(depth - 1)]           # one workspace per in-flight step
    xs_d = [torch.empty_like(x_dev) for _ in range(depth)]
    ys_d = [torch.empty_like(y_dev) for _ in range(depth)]
    streams = [torch.cuda.Stream(dev) for _ in range(depth)]

    def e2e_steps(n):
        for i in range(n):
            j = i % depth
            with torch.cuda.stream(streams[j]):
                plans[j].forward_host(xs_h[j], ys_h[j], "bf16", xs_d[j], ys_d[j])

    e2e_steps(W)
    torch.cuda.synchronize()
    shard.barrier()
    t0 = time.perf_counter()
    e2e_steps(K)
    torch.cuda.synchronize()
    e2e_s = shard.max_over_ranks(time.perf_counter() - t0, str(dev) if world > 1 else "cpu")
    e2e = {"value": world * K * out_mpix_step / e2e_s, "unit": "Mpix/s", "h2d_bytes_per_step": xs_h[0].numel() * 2,
           "d2h_bytes_per_step": ys_h[0].numel() * 2, "ms_per_step": e2e_s / K * 1e3,
           "how": f"b200sr_wdsr_forward_host, pinned host buffers, {depth} steps in flight on {depth} streams"}
    checksum = float(ys_h[0][:1].float().sum())           # device->host result actually read

    # ---- the same end-to-end call returning 8-bit frames ((sr*255).round().clamp(0,255), what the reference's evaluation makes of
    #      every output, common/metrics.py:12) written by the tail epilogue: half the bf16 bytes over PCIe.  Reported NEXT TO e2e,
    #      not instead of it: e2e keeps the float (bf16) output of the headline configuration.
    yu_h = [torch.empty(BATCH, 3, LR * SCALE, LR * SCALE, dtype=torch.uint8).pin_memory() for _ in range(depth)]
    yu_d = [torch.empty(BATCH, 3, LR * SCALE, LR * SCALE, dtype=torch.uint8, device=dev) for _ in range(depth)]

    def e2e_u8_steps(n):
        for i in range(n):
            j = i % depth
            with torch.cuda.stream(streams[j]):
                plans[j].forward_host(xs_h[j], yu_h[j], "bf16", xs_d[j], yu_d[j])

    e2e_u8_steps(W)
    torch.cuda.synchronize()
    shard.barrier()
    t0 = time.perf_counter()
    e2e_u8_steps(K)
    torch.cuda.synchronize()
    u8_s = shard.max_over_ranks(time.perf_counter() - t0, str(dev) if world > 1 else "cpu")
    e2e_u8 = {"value": world * K * out_mpix_step / u8_s, "unit": "Mpix/s", "h2d_bytes_per_step": xs_h[0].numel() * 2,
              "d2h_bytes_per_step": yu_h[0].numel(), "ms_per_step": u8_s / K * 1e3,
              "how": "same call with y_dtype = B200SR_U8: 8-bit frames from the tail epilogue (extra information, not the headline e2e)",
              "checksum": float(yu_h[0][:1].float().sum())}

    # ---- host ceiling of the end-to-end number: every rank copies its step's output bytes device -> pinned host, all ranks at once
    #      (plain cudaMemcpyAsync, what forward_host issues); the aggregate GB/s is what e2e can reach at most at this N.
    shard.barrier()
    torch.cuda.synchronize()
    reps_c = 10
    ys_h[0].copy_(ys_d[0], non_blocking=True)
    torch.cuda.synchronize()
    shard.barrier()
    t0 = time.perf_counter()
    for i in range(reps_c):
        ys_h[i % depth].copy_(ys_d[i % depth], non_blocking=True)
    torch.cuda.synchronize()
    d2h_s = shard.max_over_ranks(time.perf_counter() - t0, str(dev) if world > 1 else "cpu")
    d2h_gbs = world * reps_c * ys_h[0].numel() * 2 / d2h_s / 1e9
    e2e["host_ceiling_gbs"] = d2h_gbs
    e2e["host_ceiling_mpix_s"] = d2h_gbs * 1e9 / (3 * 2) / 1e6                 # 3 channels x 2 bytes per output pixel
    e2e["frac_of_host_ceiling"] = e2e["value"] / e2e["host_ceiling_mpix_s"]
    e2e["limiter"] = (f"device->host copy of the bf16 output ({ys_h[0].numel() * 2 / 1e6:.1f} MB per step and GPU): {world} concurrent pinned "
                      f"cudaMemcpyAsync reach {d2h_gbs:.1f} GB/s in aggregate on this box")
    try:
        e2e["cpu_affinity"] = sorted(os.sched_getaffinity(0))[:1] + [len(os.sched_getaffinity(0))]
    except Exception:
        pass

    extras = north_star_configs(dev, rank, world, peaks, shard) if not args.no_extras else None

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, sec, cores = cpu_forward_rate(16, 3, 1)
        cpu_baseline = {"value": rate, "unit": "Mpix/s", "cores": cores, "kind": "port",
                        "sample": "16 of the 64 patches, fp32 torch-CPU restatement of the reference forward (oracle/port.py), "
                                  "1 warm-up + median of 3"}
    if rank == 0:
        line = {"metric": "output Mpixels/sec", "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
                "data": "synthetic",
                "config": {"workload": WORKLOAD, "sharding": f"{world} x full batch, no collective", "l2": "flushed before every timed step (256 MiB fill)",
                           "timing": "sum of per-step CUDA-event intervals on the launch stream, max over ranks"},
                "frames_per_s": world * BATCH / (ms_per_step / 1e3), "wall_ms_per_step_incl_flush": t_wall / K * 1e3,
                "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
                "e2e_checksum": checksum, "e2e_u8_frames": e2e_u8, "north_star": extras}
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the north-star side configs (dense 360p, cfg3, cfg4, cfg5, sustained leg)")
    args = ap.parse_args()
    rank, local_rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    run_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
