"""development aid: rfft/irfft of one order against the oracle:  python tools/check_rfft_order.py ORDER"""
import os, sys, numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import __graft_entry__ as entry
from oracle import pyoracle
order = int(sys.argv[1]); pkg = entry.load_package(); orc = pyoracle.oracle()
n = 1 << order
x = np.stack([orc.noise(n, 40 + b, np.float32) for b in range(3)])
rp = pkg.RFFTPlan(order, np.float32)
S = rp.rfft(x); want = orc.rfft(x)
e1 = np.linalg.norm(S - want) / np.linalg.norm(want)
e2 = np.linalg.norm(rp.irfft(S) / n - x) / np.linalg.norm(x)
print(f"order {order} rel err", e1, e2)
assert e1 < 1e-5 and e2 < 1e-5
