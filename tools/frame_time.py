"""Development aid: device time per phase of the BASELINE config 5 bank (1024 ch x 2^20 taps, B = 1024) in frame mode.
usage: python tools/frame_time.py [T ...]   (needs a B200; prints one JSON line per T)"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch

import __graft_entry__ as entry

CHANNELS, BLOCK = 1024, 1024
TAPS = int(os.environ.get("FRAME_TAPS", 1 << 20))  # 2^19 taps at T = 256 gives Q = 2 frame partitions: one shard of a 2-way partition split


def main():
    pkg = entry.load_package()
    frames = [int(a) for a in sys.argv[1:]] or [64, 128, 256]
    channels = int(os.environ.get("FRAME_CHANNELS", CHANNELS))
    gen = torch.Generator(device="cuda").manual_seed(11)
    ir = torch.rand((channels, TAPS), device="cuda", dtype=torch.float32, generator=gen) * 2 - 1
    ir *= 1.0 / ir.square().sum(dim=1).max().sqrt()
    for T in frames:
        conv = pkg.Convolver(pkg.UPOLS, "float32", pkg.DIAGONAL, frame_blocks=T)
        conv.set_stream(torch.cuda.current_stream())
        conv.impulse(ir, BLOCK)
        x = torch.rand((channels, T * BLOCK), device="cuda", dtype=torch.float32, generator=gen) * 2 - 1
        y = torch.empty_like(x)
        steps = max(3, 2048 // T)
        for _ in range(2):
            conv(x, out=y)
        conv.profile(True)
        conv.profile_read()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            conv(x, out=y)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        r2c, mac, c2r, ff, fi, n = conv.profile_read(frame_phases=True)
        conv.profile(False)
        print(json.dumps({
            "T": T, "channels": channels, "ms_per_call": round(ms, 4),
            "channel_msamples_per_s": round(channels * T * BLOCK / ms / 1e3, 1),
            "phase_ms": {k: round(v / steps, 4) for k, v in
                         dict(r2c=r2c, frame_fwd=ff, mac=mac, frame_inv=fi, c2r=c2r).items()},
            "device_gb": round(conv.device_bytes() / 1e9, 2),
        }), flush=True)
        conv.close()
        del x, y
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
