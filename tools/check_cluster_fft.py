"""Quick parity check of the cluster rfft/irfft (orders 14-16) against numpy float64, with batches below and above the number of
resident clusters (exercises the persistent loop and the double-buffered scratch)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry

pkg = entry.load_package()
pkg.set_device(0)
rng = np.random.default_rng(3)
ok = True
for order in (14, 15, 16):
    n = 1 << order
    for batch in (1, 3, 200, 1000):
        x = rng.uniform(-1, 1, size=(batch, n)).astype(np.float32)
        plan = pkg.RFFTPlan(order, "float32")
        spec = plan.rfft(x)
        ref = np.fft.rfft(x.astype(np.float64), axis=1)
        e1 = np.linalg.norm(spec - ref) / np.linalg.norm(ref)
        back = plan.irfft(ref.astype(np.complex64))
        e2 = np.linalg.norm(back / n - x) / np.linalg.norm(x)
        worst = max(np.linalg.norm(spec[i] - ref[i]) / np.linalg.norm(ref[i]) for i in range(batch))
        print(f"order {order} batch {batch}: rfft {e1:.2e} (worst row {worst:.2e}) irfft {e2:.2e}", flush=True)
        ok &= e1 < 1e-5 and e2 < 1e-5 and worst < 1e-5
        plan.close()
print("CLUSTER FFT OK" if ok else "CLUSTER FFT FAILED")
sys.exit(0 if ok else 1)
