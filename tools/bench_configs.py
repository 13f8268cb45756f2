#!/usr/bin/env python
"""Measures the BASELINE.json configs that are not bench.py's headline line (C1, C3, C4) on one GPU and prints one JSON object.

C1  neo::fft c2c complex<float> N=1024 batch=1 forward+inverse (+1/N scale): reference CPU (oracle/_ref) vs the CUDA plan through
    the C ABI with host buffers (latency-bound by construction: one tiny transform per call) and batched on the device.
C3  UPOLS stereo, B=512, 2^17-tap IR (P=256): microseconds per block (T=1 streaming) and x real-time at 48 kHz.
C4  64-in x 64-out convolution matrix, B=256, 2^16-tap IRs (P=256): ms per block-step, T=1 (HBM stream of the 2.16 GB filter
    set) and T=16.
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry
from oracle import pyoracle

pkg = entry.load_package()
pkg.set_device(0)
orc = pyoracle.oracle()
ref = pyoracle.ref()
stream = torch.cuda.current_stream()
out = {}


def gpu_time(fn, reps=50, warm=5):
    for _ in range(warm):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps  # ms


# ---- C1 -----------------------------------------------------------------------------------------------------------------
x = orc.noise(1024, 1, np.complex64)
chk = ref if ref is not None else orc
t0 = time.perf_counter()
reps = 2000
for _ in range(reps):
    y = chk.fft(chk.fft(x, -1), +1) / 1024
cpu_us = (time.perf_counter() - t0) / reps * 1e6  # includes ctypes overhead of two calls
plan = pkg.FFTPlan(10, np.complex64)
buf = x.copy()
t0 = time.perf_counter()
for _ in range(200):
    plan(buf, pkg.FORWARD)
    plan(buf, pkg.BACKWARD)
host_us = (time.perf_counter() - t0) / 200 * 1e6
plan.set_stream(stream)
batch = 1 << 18
dx = torch.randn((batch, 1024), dtype=torch.complex64, device="cuda")
ms = gpu_time(lambda: (plan(dx, pkg.FORWARD), plan(dx, pkg.BACKWARD)))
out["C1"] = {
    "reference_cpu_us_per_roundtrip_incl_plan_build": cpu_us,
    "b200_host_call_us_per_roundtrip_batch1": host_us,
    "b200_device_batched_ns_per_roundtrip": ms * 1e6 / batch,
    "b200_device_batched_gbs": 2 * 2 * batch * 1024 * 8 / ms / 1e6,
    "note": "batch=1 through host buffers is PCIe/launch latency; the batched figure is the one comparable with the HBM roofline",
}

# ---- C3 -----------------------------------------------------------------------------------------------------------------
B, L, C = 512, 1 << 17, 2
ir = torch.rand((C, L), device="cuda") * 2 - 1
ir *= 1.0 / ir.square().sum(dim=1).max().sqrt()
res = {}
for T in (1, 16):
    conv = pkg.Convolver(pkg.UPOLS, "float32", pkg.DIAGONAL, max_blocks=T)
    conv.set_stream(stream)
    conv.impulse(ir, B)
    xin = torch.rand((C, T * B), device="cuda") * 2 - 1
    yout = torch.empty_like(xin)
    ms = gpu_time(lambda: conv(xin, out=yout), reps=200, warm=20)
    res[f"T{T}"] = {"us_per_block": ms * 1e3 / T, "realtime_x_48k": (B * T / 48000.0) / (ms * 1e-3),
                    "channel_msamples_s": C * B * T / ms / 1e3}
    conv.close()
for T in (64,):  # frame mode: second overlap-save level along block time
    conv = pkg.Convolver(pkg.UPOLS, "float32", pkg.DIAGONAL, frame_blocks=T)
    conv.set_stream(stream)
    conv.impulse(ir, B)
    xin = torch.rand((C, T * B), device="cuda") * 2 - 1
    yout = torch.empty_like(xin)
    ms = gpu_time(lambda: conv(xin, out=yout), reps=100, warm=10)
    res[f"frame{T}"] = {"us_per_block": ms * 1e3 / T, "realtime_x_48k": (B * T / 48000.0) / (ms * 1e-3),
                        "channel_msamples_s": C * B * T / ms / 1e3}
    conv.close()
res["bytes_per_channel_block_T1"] = 8 * B + 8 * (B + 1) + 16 * (B + 1) * (L // B)
res["note"] = "2 channels x 2.1 MB of state: L2-resident, bound by launch latency of three small kernels, not by HBM"
out["C3"] = res

# ---- C4 -----------------------------------------------------------------------------------------------------------------
B, L, O, I = 256, 1 << 16, 64, 64
ir = torch.rand((O, I, L), device="cuda") * 2 - 1
ir *= 1.0 / ir.square().sum(dim=2).max().sqrt()
res = {}
P, K = L // B, B + 1
for T in (1, 16):
    conv = pkg.Convolver(pkg.UPOLS, "float32", pkg.MATRIX, max_blocks=T)
    conv.set_stream(stream)
    conv.impulse(ir, B)
    xin = torch.rand((I, T * B), device="cuda") * 2 - 1
    yout = torch.empty((O, T * B), device="cuda")
    conv.profile(True)
    ms = gpu_time(lambda: conv(xin, out=yout), reps=30, warm=5)
    r2c_ms, mac_ms, c2r_ms, launches = conv.profile_read()
    filt_bytes = 8 * K * P * O * I
    res[f"T{T}"] = {
        "ms_per_block_step": ms / T,
        "realtime_x_48k": (B * T / 48000.0) / (ms * 1e-3),
        "mac_ms_per_call": mac_ms / 35,
        "mac_filter_stream_gbs": filt_bytes / (mac_ms / 35) / 1e6,
        "mac_fp32_tflops": 8.0 * K * P * O * I * T / (mac_ms / 35) / 1e9,
    }
    conv.close()
for T in (32, 64):  # frame mode: the contraction becomes Q = P/T streamed rows of 2T*B bins per (output, input) pair
    conv = pkg.Convolver(pkg.UPOLS, "float32", pkg.MATRIX, frame_blocks=T)
    conv.set_stream(stream)
    conv.impulse(ir, B)
    xin = torch.rand((I, T * B), device="cuda") * 2 - 1
    yout = torch.empty((O, T * B), device="cuda")
    conv.profile(True)
    ms = gpu_time(lambda: conv(xin, out=yout), reps=30, warm=5)
    r2c_ms, mac_ms, c2r_ms, ff_ms, fi_ms, launches = conv.profile_read(frame_phases=True)
    filt2_bytes = 8 * K * 2 * T * ((P + T - 1) // T) * O * I
    res[f"frame{T}"] = {
        "ms_per_block_step": ms / T,
        "realtime_x_48k": (B * T / 48000.0) / (ms * 1e-3),
        "mac_ms_per_call": mac_ms / 35,
        "mac_filter_stream_gbs": filt2_bytes / (mac_ms / 35) / 1e6,
        "frame_transform_ms_per_call": (ff_ms + fi_ms) / 35,
        "device_gb": conv.device_bytes() / 1e9,
    }
    conv.close()
res["filter_bytes"] = 8 * K * P * O * I
res["note"] = "T=1: the 2.16 GB filter set streams once per block-step (HBM bound); FDL (34 MB) is L2-resident"
out["C4"] = res
# ---- float64: the same kernels instantiated for double (8 points per thread, 64-element tiles) ------------------------------------
res = {}
peak = None
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
for order in (10, 12, 14):
    n = 1 << order
    batch = (1 << 27) // n
    xr = torch.rand((batch, n), device="cuda", dtype=torch.float64) * 2 - 1
    rp = pkg.RFFTPlan(order, "float64")
    rp.set_stream(stream)
    spec = torch.empty((batch, n // 2 + 1), dtype=torch.complex128, device="cuda")
    back = torch.empty_like(xr)
    ms_f = gpu_time(lambda: rp.rfft(xr, out=spec), reps=10, warm=3)
    ms_b = gpu_time(lambda: rp.irfft(spec, out=back), reps=10, warm=3)
    bytes_ = batch * (8 * n + 16 * (n // 2 + 1))
    res[f"rfft_n{n}"] = {"r2c_gbs": bytes_ / ms_f / 1e6, "c2r_gbs": bytes_ / ms_b / 1e6,
                         "r2c_frac": bytes_ / ms_f / 1e6 / peak if peak else None, "c2r_frac": bytes_ / ms_b / 1e6 / peak if peak else None}
    rp.close()
    del xr, spec, back
B, L, C = 1024, 1 << 18, 256
ir = torch.rand((C, L), device="cuda", dtype=torch.float64) * 2 - 1
ir *= 1.0 / ir.square().sum(dim=1).max().sqrt()
for frame, T in ((0, 1), (0, 8), (64, 64)):
    conv = pkg.Convolver(pkg.UPOLS, "float64", pkg.DIAGONAL, max_blocks=T, frame_blocks=frame)
    conv.set_stream(stream)
    conv.impulse(ir, B)
    xin = torch.rand((C, T * B), device="cuda", dtype=torch.float64) * 2 - 1
    yout = torch.empty_like(xin)
    ms = gpu_time(lambda: conv(xin, out=yout), reps=20, warm=3)
    res[("frame" if frame else "T") + str(T)] = {"channel_msamples_s": C * B * T / ms / 1e3, "ms_per_call": ms}
    conv.close()
res["bank"] = "256 channels x 2^18-tap IRs, B = 1024 (P = 256), float64"
out["F64"] = res
print(json.dumps(out, indent=1))
