"""Development aid: the streaming (T = 1) MAC of a bank with sparse filters (neo_b200_conv_set_filter_csr) against the dense bank, by
density and pattern. usage: python tools/sparse_time.py   (needs a B200; one JSON line per case)"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

import __graft_entry__ as entry

CHANNELS, BLOCK, PARTS = 64, 1024, 1024


def run(pkg, conv, x, y, steps=40):
    conv.set_stream(torch.cuda.current_stream())
    for _ in range(3):
        conv(x, out=y)
    conv.profile(True)
    conv.profile_read()
    for _ in range(steps):
        conv(x, out=y)
    r2c, mac, c2r, n = conv.profile_read()
    conv.profile(False)
    return mac / steps


def main():
    pkg = entry.load_package()
    rng = np.random.default_rng(5)
    H = (rng.standard_normal((CHANNELS, PARTS, BLOCK + 1), dtype=np.float32) + 1j * rng.standard_normal((CHANNELS, PARTS, BLOCK + 1), dtype=np.float32)).astype(np.complex64)
    H[:, :, 0] = H[:, :, 0].real
    H[:, :, BLOCK] = H[:, :, BLOCK].real
    x = torch.rand((CHANNELS, BLOCK), device="cuda") * 2 - 1
    y = torch.empty_like(x)
    dense = pkg.Convolver(pkg.UPOLS, "float32", pkg.DIAGONAL)
    dense.filter(H)
    t_dense = run(pkg, dense, x, y)
    dense_bytes = dense.device_bytes()
    want = y.clone()
    dense.close()
    alg_dense = CHANNELS * 8 * (BLOCK + 1) * (2 * PARTS + 1)
    print(json.dumps({"case": "dense", "mac_ms": round(t_dense, 4), "gbs": round(alg_dense / t_dense / 1e6, 1), "device_gb": round(dense_bytes / 1e9, 3)}), flush=True)
    mag = np.abs(H)
    cases = [("all kept", np.ones(H.shape, dtype=bool))]
    for q in (0.5, 0.9):
        cases.append((f"random, {int((1 - q) * 100)} % kept", mag > np.quantile(mag[0], q)))
    band = np.ones(H.shape, dtype=bool)
    band[:, PARTS // 8:, BLOCK // 2:] = False  # the upper half of the band decays after an eighth of the response
    cases.append(("upper half of the band dropped after P/8 partitions", band))
    for name, keep in cases:
        conv = pkg.Convolver(pkg.UPOLS, "float32", pkg.DIAGONAL)
        conv.filter_sparse(H, keep)
        t = run(pkg, conv, x, y)
        line = {"case": name, "density": round(float(keep.mean()), 3), "mac_ms": round(t, 4), "vs_dense": round(t_dense / t, 2),
                "device_gb": round(conv.device_bytes() / 1e9, 3)}
        if keep.all():
            line["max_abs_diff_vs_dense"] = float((y - want).abs().max())
        print(json.dumps(line), flush=True)
        conv.close()


if __name__ == "__main__":
    main()
