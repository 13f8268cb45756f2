"""Runs batched rfft/irfft of one size a few times (profiling target): python tools/fft_run.py ORDER [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry

order = int(sys.argv[1])
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
pkg = entry.load_package()
pkg.set_device(0)
n = 1 << order
batch = (1 << 28) // n
x = torch.rand((batch, n), device="cuda") * 2 - 1
plan = pkg.RFFTPlan(order, "float32")
plan.set_stream(torch.cuda.current_stream())
spec = torch.empty((batch, n // 2 + 1), dtype=torch.complex64, device="cuda")
back = torch.empty_like(x)
for _ in range(reps):
    plan.rfft(x, out=spec)
    plan.irfft(spec, out=back)
torch.cuda.synchronize()
print("ok", order, batch)
