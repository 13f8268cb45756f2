"""Development aid: what the host link gives -- pinned H2D, D2H and both at once (GB/s per direction), by chunk size.
The e2e numbers of bench.py are bound by this."""
import torch

def run(nbytes, chunk, both, iters=4):
    h_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    res = {}
    for mode in (["h2d", "d2h", "both"] if both else ["h2d"]):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s1.wait_event(e0); s2.wait_event(e0)
        for _ in range(iters):
            for off in range(0, nbytes, chunk):
                if mode in ("h2d", "both"):
                    with torch.cuda.stream(s1):
                        d_in[off:off + chunk].copy_(h_in[off:off + chunk], non_blocking=True)
                if mode in ("d2h", "both"):
                    with torch.cuda.stream(s2):
                        h_out[off:off + chunk].copy_(d_out[off:off + chunk], non_blocking=True)
        torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
        e1.record()
        torch.cuda.synchronize()
        res[mode] = round(nbytes * iters / (e0.elapsed_time(e1) * 1e-3) / 1e9, 1)
    return res

for chunk in (1 << 30, 1 << 26, 1 << 24):
    print("chunk MiB", chunk >> 20, run(1 << 30, chunk, True), flush=True)
