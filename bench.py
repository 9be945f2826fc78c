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
# measured DRAM bytes of ONE block launch at this workload (ncu --set full, profiles/r01_block_tcgen05_v2_ncu.md):
# 28,391,936 read + 33,792 written -- the output stays in L2 for the next block
NCU_DRAM_BYTES_PER_BLOCK_LAUNCH = 28_391_936 + 33_792
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
    roofline = {"kernel": "wdsr_block_tc5p_kernel (tcgen05 fused residual block)", "bound": "tensor", "achieved": ach_tflops,
                "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": ach_tflops / peaks["bf16_tflops_sustained"],
                "traffic": NCU_DRAM_BYTES_PER_BLOCK_LAUNCH, "traffic_source": "profiles/r01_block_tcgen05_v2_ncu.md (dram__bytes_read.sum + dram__bytes_write.sum, one ncu --set full capture)", "peak_source": peaks["source"] + " (sustained bf16 GEMM; kernel timed inside a 16-launch loop)",
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
                "e2e_checksum": checksum, "e2e_u8_frames": e2e_u8}
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
    args = ap.parse_args()
    rank, local_rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    run_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
