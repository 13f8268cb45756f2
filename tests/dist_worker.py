"""Worker for the multi-process tests (launched with torch.distributed.run).

    --backend gloo : CPU, world_size 2. Exercises the host-side sharding logic of the partition-sharded path (partition ranges,
                     channel shards, reduce-scatter layout) with the ORACLE standing in for the per-rank CUDA kernels --
                     test infrastructure only; it proves the decomposition bench.py uses is exact, not the kernels.
    --backend nccl : one GPU per rank, the real path: conv_forward -> NCCL reduce-scatter of partial spectra -> conv_inverse.
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402
from oracle import pyoracle  # noqa: E402


def shard_ranges(parts: int, channels: int, world: int, rank: int):
    """the decomposition bench.py uses: partitions [lo, hi) of every filter, channels [c0, c1) for the inverse"""
    return (rank * parts // world, (rank + 1) * parts // world), (rank * channels // world, (rank + 1) * channels // world)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", default="gloo")
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    orc = pyoracle.oracle()
    C, B, P, NB, T = 4, 64, 6, 12, 3
    ir = orc.normalize_impulse(np.stack([orc.noise(B * P - 9, 11 + c, np.float32) for c in range(C)]))
    sig = np.stack([orc.noise(B * NB, 13 + c, np.float32) for c in range(C)])
    H = orc.uniform_partition(ir, B)
    want = orc.convolve_blocks(0, H, sig)
    (lo, hi), (c0, c1) = shard_ranges(P, C, world, rank)

    if args.backend == "gloo":
        dist.init_process_group("gloo")
        # rank-local partial result: this rank's partitions only, on an input delayed by `lo` blocks (linearity of the FDL)
        Hl = np.ascontiguousarray(H[:, lo:hi])
        delayed = np.concatenate([np.zeros((C, lo * B), np.float32), sig], axis=1)[:, : B * NB]
        part = torch.from_numpy(orc.convolve_blocks(0, Hl, delayed).astype(np.float64))
        # gloo has no reduce_scatter: all_reduce then slice is the same reduction
        dist.all_reduce(part)
        out = part[c0:c1]
        err = np.linalg.norm(out.numpy() - want[c0:c1]) / np.linalg.norm(want[c0:c1])
        assert err < 1e-6, err
    else:
        local = int(os.environ.get("LOCAL_RANK", rank))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        pkg = entry.load_package()
        pkg.set_device(local)
        conv = pkg.Convolver(pkg.UPOLS, "float32", pkg.DIAGONAL, max_blocks=T, partition_range=(lo, hi))
        conv.set_stream(torch.cuda.current_stream())
        conv.filter(H)
        got = np.zeros((c1 - c0, B * NB), dtype=np.float32)
        shard = torch.empty((c1 - c0, T, 2 * B), device="cuda", dtype=torch.float32)
        for pos in range(0, NB, T):
            x = torch.from_numpy(np.ascontiguousarray(sig[:, pos * B : (pos + T) * B])).cuda()
            y = torch.empty((c1 - c0, T * B), device="cuda", dtype=torch.float32)
            conv.forward(x)
            dist.reduce_scatter_tensor(shard, conv.spectra_tensor(T))
            conv.inverse(shard, y, c0, c1 - c0, T)
            torch.cuda.synchronize()
            got[:, pos * B : (pos + T) * B] = y.cpu().numpy()
        err = np.linalg.norm(got - want[c0:c1]) / np.linalg.norm(want[c0:c1])
        assert err < 1e-5, err

        # the grouped form bench.py uses: channel groups, async reduce-scatter of one group while the next group's MAC runs;
        # rank r then owns the r-th slice of every group
        conv.reset()
        groups, gch = 2, C // 2
        gsh = gch // world
        mine = [g * gch + rank * gsh + i for g in range(groups) for i in range(gsh)]
        got2 = np.zeros((len(mine), B * NB), dtype=np.float32)
        shards = [torch.empty((gsh, T, 2 * B), device="cuda", dtype=torch.float32) for _ in range(groups)]
        for pos in range(0, NB, T):
            x = torch.from_numpy(np.ascontiguousarray(sig[:, pos * B : (pos + T) * B])).cuda()
            y = torch.empty((groups, gsh, T * B), device="cuda", dtype=torch.float32)
            works = []
            for g in range(groups):
                conv.forward_range(x, g * gch, gch, g == groups - 1)
                spectra = conv.spectra_tensor(T)  # after the forward: a sharded handle alternates between two buffers
                works.append(dist.reduce_scatter_tensor(shards[g], spectra[g * gch : (g + 1) * gch], async_op=True))
            for g in range(groups):
                works[g].wait()
                conv.inverse(shards[g], y[g], g * gch + rank * gsh, gsh, T)
            torch.cuda.synchronize()
            got2[:, pos * B : (pos + T) * B] = y.view(-1, T * B).cpu().numpy()
        err2 = np.linalg.norm(got2 - want[mine]) / np.linalg.norm(want[mine])
        assert err2 < 1e-5, err2
        conv.close()

        # frame mode (calls of exactly TF blocks, second overlap-save level along block time): a shard starts on a frame boundary
        # of the partition axis; the partial LEVEL-1 spectra are reduced exactly as above
        P2, TF = 8, 2
        ir2 = orc.normalize_impulse(np.stack([orc.noise(B * P2 - 9, 21 + c, np.float32) for c in range(C)]))
        H2 = orc.uniform_partition(ir2, B)
        want3 = orc.convolve_blocks(0, H2, sig)
        (lo2, hi2), _ = shard_ranges(P2, C, world, rank)
        assert lo2 % TF == 0
        conv = pkg.Convolver(pkg.UPOLS, "float32", pkg.DIAGONAL, partition_range=(lo2, hi2), frame_blocks=TF)
        conv.set_stream(torch.cuda.current_stream())
        conv.filter(H2)
        got3 = np.zeros((c1 - c0, B * NB), dtype=np.float32)
        shard = torch.empty((c1 - c0, TF, 2 * B), device="cuda", dtype=torch.float32)
        for pos in range(0, NB, TF):
            x = torch.from_numpy(np.ascontiguousarray(sig[:, pos * B : (pos + TF) * B])).cuda()
            y = torch.empty((c1 - c0, TF * B), device="cuda", dtype=torch.float32)
            conv.forward(x)
            dist.reduce_scatter_tensor(shard, conv.spectra_tensor(TF))
            conv.inverse(shard, y, c0, c1 - c0, TF)
            torch.cuda.synchronize()
            got3[:, pos * B : (pos + TF) * B] = y.cpu().numpy()
        err3 = np.linalg.norm(got3 - want3[c0:c1]) / np.linalg.norm(want3[c0:c1])
        assert err3 < 1e-5, err3
    dist.barrier()
    if rank == 0:
        print(f"dist_worker ok backend={args.backend} world={world} err={err:.2e}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
