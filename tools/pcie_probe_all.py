"""Development aid: the host link of EVERY GPU of the box at the same time (torchrun, one rank per GPU): pinned H2D + D2H running together
on each rank, all ranks started by a barrier. Prints per-rank GB/s per direction and the aggregate -- the bound of the N-GPU e2e numbers.
usage: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_probe_all.py"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import torch.distributed as dist


def main():
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl")
    world = dist.get_world_size()
    try:
        import bench
        bench.bind_to_gpu_numa_node(local)  # the same NUMA binding bench.py applies before it allocates pinned memory
    except Exception as exc:  # pragma: no cover
        print("no NUMA binding:", exc, file=sys.stderr)
    nbytes = 256 << 20
    h_in, h_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory(), torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d_in, d_out = torch.empty(nbytes, dtype=torch.uint8, device="cuda"), torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    res = {}
    for mode in ("h2d", "d2h", "both"):
        best = 0.0
        for _ in range(3):
            torch.cuda.synchronize()
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            s1.wait_event(e0)
            s2.wait_event(e0)
            for _ in range(8):
                if mode in ("h2d", "both"):
                    with torch.cuda.stream(s1):
                        d_in.copy_(h_in, non_blocking=True)
                if mode in ("d2h", "both"):
                    with torch.cuda.stream(s2):
                        h_out.copy_(d_out, non_blocking=True)
            torch.cuda.current_stream().wait_stream(s1)
            torch.cuda.current_stream().wait_stream(s2)
            e1.record()
            torch.cuda.synchronize()
            best = max(best, 8 * nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9)
        res[mode] = best
    t = torch.tensor([res["h2d"], res["d2h"], res["both"]], device="cuda", dtype=torch.float64)
    allv = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allv, t)
    if dist.get_rank() == 0:
        rows = [[round(float(v), 1) for v in a.tolist()] for a in allv]
        print({"gpus": world, "per_rank_gbs [h2d alone, d2h alone, each direction with both]": rows,
               "aggregate_each_direction_with_both": round(sum(r[2] for r in rows), 1)}, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
