#!/usr/bin/env python
"""bench.py -- the reference's headline workload on B200 (BASELINE.json):
"UPOLS channel-Msamples/s (B=1024, 2^20 taps); batched FFT GB/s vs HBM peak".

    python bench.py --gpus N --steps K --warmup W [--frame T | --frame 0 --blocks T] [--impl reference]

A step is one call of the convolver bank: T consecutive blocks of 1024 samples for each of the 1024 channels, each
channel convolved with its own 2^20-tap impulse response (P = 1024 partitions) -- BASELINE config 5. T = 1 is the
reference's streaming call (one block per call); T > 1 hands the bank T blocks at once (the CLI / offline case,
extra/cli/src/convolver.cpp:42-55). Two forms of T > 1: the direct form (--frame 0 --blocks T) reuses every filter
partition T times in the MAC kernel (FP32-bound from T = 16); frame mode (--frame T, the default, T = 256) evaluates the
sum over partitions by a second overlap-save level along block time (neo-dsp_b200/csrc/conv_frame.cuh), HBM-bound again.
Same results within the float32 tolerance either way (tests/test_conv_frame_gpu.py); every number states its mode.

  value        whole-job channel-Msamples/s, inputs resident in HBM, CUDA events on the launching stream, max over ranks
  e2e          the same through the C-ABI call with HOST (pinned) buffers: H2D and D2H inside the timed region
  roofline     the dominant kernel (spectral MAC): algorithmic bytes per launch / event-timed duration vs measured HBM peak
  cpu_baseline the reference's own CPU convolver (oracle/_ref) on this box's host cores, bounded sample (N=1, rank 0)
  fft_sweep    batched rfft/irfft float32 N=2^10..2^16 (BASELINE config 2), GB/s against the same HBM peak

N > 1 (torchrun, one rank per GPU): the partitions of every impulse response are sharded across ranks, each rank
produces partial spectra for all channels, one NCCL reduce-scatter over NVLink sums them and leaves every rank with
the channels whose c2r it runs. Total work is fixed: "scaling": "strong".
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CHANNELS, BLOCK, TAPS = 1024, 1024, 1 << 20
PARTS = TAPS // BLOCK
METRIC = "UPOLS channel-Msamples/s (B=1024, 2^20 taps); batched FFT GB/s vs HBM peak"
UNIT = "channel-Msamples/s"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler:
    """SM clock + throttle reasons sampled through NVML while the timed region runs."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.02)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        med = int(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---- reference arm ---------------------------------------------------------------------------------------------------------
_CPU_INPUTS: dict = {}


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_reference(sample_channels: int, blocks: int, threads: int, kind: int):
    """The reference's own convolver (oracle/_ref when it travelled, else the C restatement) on the host cores:
    `sample_channels` independent convolvers, own random filter each, P = 1024, B = 1024, `blocks` blocks."""
    from oracle import pyoracle

    bins = BLOCK + 1
    key = (sample_channels, blocks)
    if key not in _CPU_INPUTS:
        rng = np.random.default_rng(11)
        one = (rng.uniform(-1, 1, size=(PARTS, 2 * bins)).astype(np.float32) * np.float32(1e-3)).view(np.complex64)
        filt = np.empty((sample_channels, PARTS, bins), dtype=np.complex64)  # own memory per channel, as in the real workload
        filt[:] = one[None]
        sig = rng.uniform(-1, 1, size=(sample_channels, blocks * BLOCK)).astype(np.float32)
        _CPU_INPUTS.clear()
        _CPU_INPUTS[key] = (filt, sig)
    filt, sig0 = _CPU_INPUTS[key]
    sig = sig0.copy()
    ref = pyoracle.ref()
    if ref is not None:
        seconds = ref.conv_bench(kind, filt, PARTS * bins, sig, threads)
        how = "reference"
    else:
        orc = pyoracle.oracle()
        start = time.perf_counter()
        orc.convolve_blocks(0, filt, sig)
        seconds = time.perf_counter() - start
        how, threads = "port", 1
    return sample_channels * blocks * BLOCK / seconds / 1e6, how, threads


def run_reference(args, rank: int):
    if rank != 0:
        return
    cores = host_cores()
    sample = max(cores, min(2 * cores, 128))
    blocks = 32
    vals = []
    for _ in range(args.warmup):
        cpu_reference(min(sample, cores), 4, cores, 2)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        v, how, used = cpu_reference(sample, blocks, cores, 2)
        vals.append(v)
    wall = time.perf_counter() - t0
    value = float(np.mean(vals))
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": value,
        "unit": UNIT,
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": 1e3 * wall / max(1, args.steps),
        "higher_is_better": True,
        "scaling": "strong" if args.gpus > 1 else "weak",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "C5: 1024 channels x 2^20-tap IR, UPOLS B=1024 (P=1024)", "blocks_per_call": 1},
        "cpu_baseline": {
            "value": value,
            "unit": UNIT,
            "cores": used,
            "kind": how,
            "sample": f"{sample} of 1024 channels x {blocks} blocks per step, own random filter per channel, "
                      "neo::split_upols_convolver (the reference's faster dense form), g++ -O3 -march=x86-64-v3, no xsimd",
        },
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---- our arm -------------------------------------------------------------------------------------------------------------------
def fft_sweep(pkg, torch, peak):
    """BASELINE config 2: batched rfft / irfft float32, N = 2^10..2^16, batch = 2^29/N (2 GiB in per pass)."""
    out = []
    x = torch.rand(1 << 29, device="cuda", dtype=torch.float32) * 2 - 1
    for order in range(10, 17):
        n = 1 << order
        batch = (1 << 29) // n
        plan = pkg.RFFTPlan(order, "float32")
        plan.set_stream(torch.cuda.current_stream())
        xin = x.view(batch, n)
        spec = torch.empty((batch, n // 2 + 1), dtype=torch.complex64, device="cuda")
        back = torch.empty_like(xin)
        bytes_pass = batch * (4 * n + 8 * (n // 2 + 1))
        res = {"n": n, "batch": batch}
        for name, fn in (("r2c", lambda: plan.rfft(xin, out=spec)), ("c2r", lambda: plan.irfft(spec, out=back))):
            for _ in range(3):
                fn()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
            ev[0].record()
            for i in range(10):
                fn()
                ev[i + 1].record()
            torch.cuda.synchronize()
            ms = float(np.median([ev[i].elapsed_time(ev[i + 1]) for i in range(10)]))
            res[name + "_gbs"] = bytes_pass / ms / 1e6
            res[name + "_frac"] = res[name + "_gbs"] / peak
        out.append(res)
        plan.close()
        del spec, back
    del x
    torch.cuda.empty_cache()
    return out


def mode_name(frame: int, T: int) -> str:
    if frame > 0:
        return f"frame mode: {T} blocks per call, sum over partitions by overlap-save along block time (frame transforms of length {2 * T})"
    return "streaming (reference call shape)" if T == 1 else f"time-batched direct form, {T} blocks per call"


def measure_channel_sharded(args, torch, dist, pkg, rank, world, local, steps):
    """Alternative multi-GPU layout (north_star: "independent channels ... are sharded with no communication"):
    channels [r*C/G, (r+1)*C/G) with their whole filters on rank r, no data-path collective at all. Returns the result dict on
    every rank (value = whole-job throughput, max over ranks)."""
    T = args.frame if args.frame > 0 else args.blocks
    ch = CHANNELS // world
    stream = torch.cuda.current_stream()
    conv = pkg.Convolver(pkg.UPOLS, "float32", pkg.DIAGONAL, max_blocks=T, frame_blocks=args.frame)
    conv.set_stream(stream)
    gen = torch.Generator(device="cuda").manual_seed(11 + rank)
    ir = torch.rand((ch, TAPS), device="cuda", dtype=torch.float32, generator=gen) * 2 - 1
    ir *= 1.0 / ir.square().sum(dim=1).max().sqrt()
    conv.impulse(ir, BLOCK)
    del ir
    xs = [torch.rand((ch, T * BLOCK), device="cuda", dtype=torch.float32) * 2 - 1 for _ in range(4)]
    ys = torch.empty((ch, T * BLOCK), device="cuda", dtype=torch.float32)
    for i in range(max(3, args.warmup)):
        conv(xs[i % 4], out=ys)
    dist.barrier()
    torch.cuda.synchronize()
    launches0 = pkg.kernel_launches()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        start.record()
        for i in range(steps):
            conv(xs[i % 4], out=ys)
        stop.record()
        dist.barrier()
        torch.cuda.synchronize()
    t = torch.tensor([start.elapsed_time(stop)], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = CHANNELS * BLOCK * T * steps / (ms_total * 1e-3) / 1e6
    launches = pkg.kernel_launches() - launches0
    conv.close()
    del conv, xs, ys
    torch.cuda.empty_cache()
    return {"value": value, "unit": UNIT, "steps": steps, "ms_per_step": ms_total / steps, "blocks_per_call": T,
            "mode": mode_name(args.frame, T), "sharding": f"channels sharded {world}-way, no collective",
            "realtime_x_wall_1024ch_48k": value * 1e6 / CHANNELS / 48000.0, "clocks": clocks.summary(), "gpu_launches": launches}


def run_channel_sharded(args, torch, dist, pkg, emit, rank, world, local):
    """--shard channels: the no-collective layout as the headline line (not BASELINE config 5's prescribed sharding)."""
    r = measure_channel_sharded(args, torch, dist, pkg, rank, world, local, args.steps)
    if rank == 0:
        emit({
            "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "C5: 1024 channels x 2^20-tap IR each, UPOLS B=1024 (P=1024, K=1025), white-noise input",
                       "blocks_per_call": r["blocks_per_call"], "mode": r["mode"],
                       "sharding": f"channels sharded {world}-way, no collective (alternative layout)",
                       "realtime_x_wall_1024ch_48k": r["realtime_x_wall_1024ch_48k"]},
            "clocks": r["clocks"], "gpu_launches": r["gpu_launches"],
        })
    dist.barrier()
    dist.destroy_process_group()


def run_ours(args):
    # stdout carries exactly ONE JSON line: native libraries (NCCL's version banner) write to fd 1 too, so fd 1 is pointed at
    # stderr for the duration of the run and the line goes to the saved descriptor
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())

    import torch
    import torch.distributed as dist

    import __graft_entry__ as entry

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local)
    pkg = entry.load_package()
    if pkg.device_count() < 1:
        raise SystemExit("no CUDA device: neo_b200 has no CPU fallback")
    pkg.set_device(local)
    if world > 1:
        # NCCL prints its version banner on stdout at NCCL_DEBUG=VERSION/INFO; stdout carries exactly one JSON line
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    frame = args.frame
    if frame > 0 and world > 1 and args.shard == "partitions":
        frame = min(frame, PARTS // world)  # a shard starts on a frame boundary of the partition axis
    T = frame if frame > 0 else args.blocks
    peak, peak_src = measured_peaks()
    stream = torch.cuda.current_stream()
    if args.fft_only:  # development aid: just the BASELINE config 2 sweep
        emit({"fft_sweep": fft_sweep(pkg, torch, peak)})
        return
    # BASELINE config 2 runs first: its passes are short bursts (10 launches each) and are compared with the burst copy
    # bandwidth, so they are taken before the sustained convolver loop pulls the part into its power cap
    sweep = None
    if world == 1 and not args.no_fft_sweep:
        sweep = fft_sweep(pkg, torch, peak)

    # ---- state: random impulse responses (unit energy like normalize_impulse), partitioned on the device ----
    by_channel = args.shard == "channels" and world > 1
    if by_channel:
        run_channel_sharded(args, torch, dist, pkg, emit, rank, world, local)
        return
    lo, hi = rank * PARTS // world, (rank + 1) * PARTS // world
    conv = pkg.Convolver(pkg.UPOLS, "float32", pkg.DIAGONAL, max_blocks=T, partition_range=(lo, hi) if world > 1 else None,
                         frame_blocks=frame)
    conv.set_stream(stream)
    gen = torch.Generator(device="cuda").manual_seed(11)
    ir = torch.rand((CHANNELS, TAPS), device="cuda", dtype=torch.float32, generator=gen) * 2 - 1
    ir *= 1.0 / ir.square().sum(dim=1).max().sqrt()
    conv.impulse(ir, BLOCK)
    del ir
    torch.cuda.empty_cache()

    gen = torch.Generator(device="cuda").manual_seed(13)  # same white noise on every rank
    nbuf = 4 if T <= 256 else 2
    xs = [torch.rand((CHANNELS, T * BLOCK), device="cuda", dtype=torch.float32, generator=gen) * 2 - 1 for _ in range(nbuf)]
    ys = torch.empty((CHANNELS, T * BLOCK), device="cuda", dtype=torch.float32)
    shard = CHANNELS // world
    # sharded runs walk the bank in channel groups so that the NCCL reduce-scatter of one group's partial spectra overlaps the
    # MAC of the next; rank r ends up owning the r-th slice of every group
    # (measured at 8 GPUs: 4 groups LOSE, 20.8k vs 22.8k channel-Msamples/s -- a rank holds only 128 partitions, so quartering the
    # channels makes every MAC launch too short; the default is therefore one group, --groups overrides)
    groups = max(1, args.groups) if world > 1 else 1
    gch = CHANNELS // groups          # channels per group
    gsh = gch // world                # of which this rank keeps gsh after the reduce-scatter
    spectra_shard = [torch.empty((gsh, T, 2 * BLOCK), device="cuda", dtype=torch.float32) for _ in range(groups)] if world > 1 else None
    ys_shard = torch.empty((groups, gsh, T * BLOCK), device="cuda", dtype=torch.float32) if world > 1 else None

    def sharded_step(x):
        works = []
        for gi in range(groups):
            conv.forward_range(x, gi * gch, gch, gi == groups - 1)
            spectra = conv.spectra_tensor(T)  # the buffer of this call (the handle alternates between two)
            works.append(dist.reduce_scatter_tensor(spectra_shard[gi], spectra[gi * gch : (gi + 1) * gch], async_op=True))
        for gi in range(groups):
            works[gi].wait()
            conv.inverse(spectra_shard[gi], ys_shard[gi], gi * gch + rank * gsh, gsh, T)

    # software pipeline across steps (throughput mode): the reduce-scatter of step i runs on NCCL's stream while step i+1's
    # r2c + MAC run; step i's c2r follows one step later (the handle double-buffers its partial spectra for exactly this).
    # drain() finishes the last step inside the timed region.
    pipelined = world > 1 and groups == 1 and not args.no_pipeline
    pending = []
    if pipelined:
        shard2 = [torch.empty((shard, T, 2 * BLOCK), device="cuda", dtype=torch.float32) for _ in range(2)]

    def finish(slot, work):
        work.wait()
        conv.inverse(shard2[slot], ys_shard.view(shard, T * BLOCK), rank * shard, shard, T)

    def pipelined_step(i, x):
        slot = i % 2
        conv.forward(x)
        # a sharded handle alternates between two partial-spectra buffers: NCCL reads this one while the next forward fills the other
        work = dist.reduce_scatter_tensor(shard2[slot], conv.spectra_tensor(T), async_op=True)
        if pending:
            finish(*pending.pop())
        pending.append((slot, work))

    def drain():
        while pending:
            finish(*pending.pop())

    def step(i):
        x = xs[i % nbuf]
        if world == 1:
            conv(x, out=ys)
        elif pipelined:
            pipelined_step(i, x)
        else:
            sharded_step(x)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(3, args.warmup)):
        step(i)
    drain()
    barrier()
    launches0 = pkg.kernel_launches()
    conv.profile(True)
    conv.profile_read()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        barrier()
        start.record()
        for i in range(args.steps):
            step(i)
        drain()
        stop.record()
        barrier()
    ms_total = start.elapsed_time(stop)
    ms_r2c, ms_mac, ms_c2r, mac_launches = conv.profile_read()
    conv.profile(False)
    launches = pkg.kernel_launches() - launches0
    if world > 1:
        t = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    samples = CHANNELS * BLOCK * T * args.steps
    value = samples / (ms_total * 1e-3) / 1e6

    # ---- e2e: the C-ABI call a reference-side caller makes, HOST buffers in pinned memory ----
    e2e = None
    if world == 1:
        hx = torch.rand((CHANNELS, T * BLOCK), dtype=torch.float32).pin_memory()
        hy = torch.empty((CHANNELS, T * BLOCK), dtype=torch.float32).pin_memory()
        hxn, hyn = hx.numpy(), hy.numpy()
        for _ in range(3):
            conv(hxn, out=hyn)
        torch.cuda.synchronize()
        n_e2e = max(3, min(args.steps, 20))
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            conv(hxn, out=hyn)  # H2D + r2c + MAC + c2r + D2H, returns when hy holds the result
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        e2e = {
            "value": CHANNELS * BLOCK * T * n_e2e / dt / 1e6,
            "unit": UNIT,
            "h2d_bytes_per_step": CHANNELS * T * BLOCK * 4,
            "d2h_bytes_per_step": CHANNELS * T * BLOCK * 4,
            "steps": n_e2e,
        }
    else:
        # sharded: every rank receives the same host block and returns its channel shard
        hx = torch.rand((CHANNELS, T * BLOCK), dtype=torch.float32).pin_memory()
        hy = torch.empty((shard, T * BLOCK), dtype=torch.float32).pin_memory()
        dx = torch.empty((CHANNELS, T * BLOCK), device="cuda", dtype=torch.float32)

        def e2e_step():  # one call at a time, fully synchronous: no cross-step pipelining here
            dx.copy_(hx, non_blocking=True)
            sharded_step(dx)
            hy.copy_(ys_shard.view(shard, T * BLOCK), non_blocking=True)
            torch.cuda.synchronize()

        for _ in range(3):
            e2e_step()
        barrier()
        n_e2e = max(3, min(args.steps, 20))
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            e2e_step()
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {
            "value": CHANNELS * BLOCK * T * n_e2e / float(dt.item()) / 1e6,
            "unit": UNIT,
            "h2d_bytes_per_step": CHANNELS * T * BLOCK * 4,
            "d2h_bytes_per_step": shard * T * BLOCK * 4,
            "steps": n_e2e,
        }

    # the no-collective layout of the same workload, measured in the same job (all ranks take part)
    alt = None
    if world > 1 and not args.no_modes:
        conv_bytes = conv.device_bytes()
        conv.close()
        del xs, ys
        if pipelined:
            del shard2
        torch.cuda.empty_cache()
        alt = measure_channel_sharded(args, torch, dist, pkg, rank, world, local, max(5, min(args.steps, 30)))
        alt["partition_sharded_device_bytes_per_rank"] = conv_bytes

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (spectral MAC), from the event-timed launches inside the timed region ----
    parts_local = hi - lo
    bins = BLOCK + 1
    if frame > 0:
        # fused frame kernel, per channel: filter rows Q*L*K + ring slots of older frames (Q-1 if this handle holds partition 0,
        # else Q) * L*K read, the two frames' level-1 spectra 2T*B read, the new ring slot L*K and the T result rows T*B written
        q_local = (parts_local + T - 1) // T
        ring_rows = q_local - 1 if lo == 0 else q_local
        alg_bytes_launch = (CHANNELS // groups) * 8 * (bins * 2 * T * (q_local + ring_rows + 1) + BLOCK * 2 * T + BLOCK * T)
        kernel_name = f"frame_fused_kernel<float, LOGL={(2 * T).bit_length() - 1}> (frame transform + ring insert + MAC + inverse frame transform)"
        fp32 = None
    else:
        # 16*K*P bytes per channel-block at T=1 (FDL row + filter row per partition); with T blocks per launch the
        # filter is read once and P+T-1 FDL rows serve all T blocks; plus the T accumulator rows written
        alg_bytes_launch = (CHANNELS // groups) * 8 * bins * (parts_local + (parts_local + T - 1) + T)  # one launch = one channel group
        kernel_name = "fdl_mac_stream_kernel<float>" if T == 1 else ("fdl_mac_tma_kernel<16,16,2>" if T == 16 else "fdl_mac_tma_kernel<32,8,3>" if T == 32 else f"fdl_mac (T={T})")
        fp32 = None
    mac_ms_avg = ms_mac / max(1, mac_launches)
    achieved = alg_bytes_launch / (mac_ms_avg * 1e-3) / 1e9 if mac_ms_avg > 0 else 0.0
    if frame == 0 and mac_ms_avg > 0:
        fp32 = (CHANNELS // groups) * 8.0 * bins * parts_local * T / (mac_ms_avg * 1e-3) / 1e12
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(f"frame_fused_T{T}_G{world}" if frame > 0 else f"fdl_mac_T{T}_G{world}")
    roofline = {
        "kernel": kernel_name,
        "bound": "hbm",
        "achieved": achieved,
        "peak": peak,
        "peak_source": peak_src,
        "unit": "GB/s",
        "frac": achieved / peak,
        "traffic": traffic,
        "algorithmic_bytes_per_launch": alg_bytes_launch,
        "launch_ms": mac_ms_avg,
        "share_of_step": ms_mac / ms_total if ms_total > 0 else None,
        "fp32_tflops": fp32,
        "phases_ms_per_step": {"r2c_fdl_insert": ms_r2c / args.steps, "mac": ms_mac / args.steps, "c2r_discard": ms_c2r / args.steps},
    }

    cpu_baseline = None
    modes = None
    if world == 1 and not args.no_modes:
        # the other call shapes, measured the same way (fewer steps), so every number on the line states its T
        del conv
        torch.cuda.empty_cache()
        tag = f"frame{T}" if frame > 0 else f"T{T}"
        modes = {tag: {"value": value, "unit": UNIT, "mac_algorithmic_gbs": achieved, "mac_fp32_tflops": fp32}}
        others = [(0, 1), (0, 16), (64, 64), (512, 512)]  # (frame, blocks per call)
        for f_other, t_other in others:
            if (f_other, t_other) == (frame, T):
                continue
            c2 = pkg.Convolver(pkg.UPOLS, "float32", pkg.DIAGONAL, max_blocks=t_other, frame_blocks=f_other)
            c2.set_stream(stream)
            gen2 = torch.Generator(device="cuda").manual_seed(11)
            ir2 = torch.rand((CHANNELS, TAPS), device="cuda", dtype=torch.float32, generator=gen2) * 2 - 1
            ir2 *= 1.0 / ir2.square().sum(dim=1).max().sqrt()
            c2.impulse(ir2, BLOCK)
            del ir2
            x2 = torch.rand((CHANNELS, t_other * BLOCK), device="cuda", dtype=torch.float32) * 2 - 1
            y2 = torch.empty_like(x2)
            for _ in range(3):
                c2(x2, out=y2)
            c2.profile(True)
            c2.profile_read()
            n2 = 30 if t_other == 1 else 10
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(n2):
                c2(x2, out=y2)
            e1.record()
            torch.cuda.synchronize()
            _, mac2, _, nl2 = c2.profile_read()
            ms2 = e0.elapsed_time(e1)
            if f_other > 0:
                q2 = (PARTS + t_other - 1) // t_other
                alg2 = CHANNELS * 8 * (bins * 2 * t_other * (2 * q2) + BLOCK * 2 * t_other + BLOCK * t_other)
                flops2 = None
            else:
                alg2 = CHANNELS * 8 * bins * (PARTS + (PARTS + t_other - 1) + t_other)
                flops2 = CHANNELS * 8.0 * bins * PARTS * t_other / (mac2 / max(1, nl2) * 1e-3) / 1e12
            modes[f"frame{t_other}" if f_other > 0 else f"T{t_other}"] = {
                "value": CHANNELS * BLOCK * t_other * n2 / (ms2 * 1e-3) / 1e6,
                "unit": UNIT,
                "mac_algorithmic_gbs": alg2 / (mac2 / max(1, nl2) * 1e-3) / 1e9,
                "mac_fp32_tflops": flops2,
                "steps": n2,
            }
            c2.close()
            del c2, x2, y2
            torch.cuda.empty_cache()
    if world == 1:
        cores = host_cores()
        sample = max(cores, min(2 * cores, 128))
        v_split, how, used = cpu_reference(sample, 16, cores, 2)
        v_aos, _, _ = cpu_reference(sample, 16, cores, 0)
        cpu_baseline = {
            "value": max(v_split, v_aos),
            "unit": UNIT,
            "cores": used,
            "kind": how,
            "sample": f"{sample} of 1024 channels x 16 blocks, own random filter per channel; split_upols_convolver {v_split:.2f}, "
                      f"upols_convolver {v_aos:.2f} {UNIT} on {used} threads (g++ -O3 -march=x86-64-v3, no xsimd)",
        }

    line = {
        "metric": METRIC,
        "value": value,
        "unit": UNIT,
        "n_gpus": world,
        "steps": args.steps,
        "warmup": max(3, args.warmup),
        "ms_per_step": ms_total / args.steps,
        "higher_is_better": True,
        "scaling": "strong" if world > 1 else "weak",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic",
        "config": {
            "workload": "C5: 1024 channels x 2^20-tap IR each, UPOLS B=1024 (P=1024, K=1025), white-noise input",
            "blocks_per_call": T,
            "mode": mode_name(frame, T),
            "sharding": "none" if world == 1 else f"partitions sharded {world}-way + NCCL reduce-scatter of partial spectra, "
                        + ("the reduce-scatter of step i overlaps the r2c+MAC of step i+1 (c2r one step later, drained inside the timed region)"
                           if pipelined else f"{groups} channel group(s) per step"),
            "l2": "working set per step (filter+FDL, 17 GB at 1 GPU) exceeds the 126 MB L2; 4 rotating input buffers",
            "realtime_x_aggregate_48k": value * 1e6 / 48000.0,
            "realtime_x_wall_1024ch_48k": value * 1e6 / CHANNELS / 48000.0,
        },
        "clocks": clocks.summary(),
        "e2e": e2e,
        "gpu_launches": launches,
        "roofline": roofline,
    }
    if cpu_baseline is not None:
        line["cpu_baseline"] = cpu_baseline
    if modes is not None:
        line["modes"] = modes
    if alt is not None:
        line["channel_sharded"] = alt
    if sweep is not None:
        line["fft_sweep"] = sweep
    emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--blocks", type=int, default=16, help="direct form (--frame 0): blocks per call T (1 = the reference's streaming call)")
    ap.add_argument("--frame", type=int, default=256,
                    help="frame mode: blocks per call T (power of two, 2..512), second overlap-save level along block time; 0 = direct form")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-fft-sweep", action="store_true")
    ap.add_argument("--fft-only", action="store_true")
    ap.add_argument("--no-modes", action="store_true", help="skip the extra T=1 / T=32 measurements")
    ap.add_argument("--shard", default="partitions", choices=["partitions", "channels"],
                    help="multi-GPU layout: BASELINE config 5's partition sharding + NCCL reduce (default) or plain channel sharding")
    ap.add_argument("--no-pipeline", action="store_true", help="sharded runs: do not overlap step i's reduce-scatter with step i+1")
    ap.add_argument("--groups", type=int, default=1, help="channel groups per sharded step (overlap of reduce-scatter and MAC)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args, int(os.environ.get("RANK", "0")))
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
