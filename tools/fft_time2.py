"""times batched rfft and irfft of one order (device buffers, 2^28 samples per call): python tools/fft_time2.py ORDER [r2c|c2r|both]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry
order = int(sys.argv[1]); which = sys.argv[2] if len(sys.argv) > 2 else "both"
pkg = entry.load_package(); pkg.set_device(0)
n = 1 << order; batch = (1 << 28) // n
x = torch.rand((batch, n), device="cuda") * 2 - 1
plan = pkg.RFFTPlan(order, "float32"); plan.set_stream(torch.cuda.current_stream())
spec = torch.empty((batch, n // 2 + 1), dtype=torch.complex64, device="cuda")
back = torch.empty_like(x)
plan.rfft(x, out=spec)
def timed(fn):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / 10
gb = batch * (4 * n + 8 * (n // 2 + 1)) / 1e6
if which in ("r2c", "both"):
    ms = timed(lambda: plan.rfft(x, out=spec)); print(f"order {order} r2c: {ms*1e3:.0f} us, {gb/ms:.0f} GB/s")
if which in ("c2r", "both"):
    ms = timed(lambda: plan.irfft(spec, out=back)); print(f"order {order} c2r: {ms*1e3:.0f} us, {gb/ms:.0f} GB/s")
    err = float((back / n - x).norm() / x.norm()); print(f"round trip rel l2 {err:.2e}")
