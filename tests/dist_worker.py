"""Worker for the multi-process tests (launched with torch.distributed.run).

    --backend gloo : CPU, world_size 2. The library's layout arithmetic (neo_b200_bank_layout_info: which rows, channel group and
                     partition range each rank holds, and how late a shard sees its input) with the ORACLE standing in for the
                     per-rank CUDA kernels -- test infrastructure only; it proves the decomposition the bank uses is exact.
    --backend nccl : one GPU per rank, the real path: the library bank in rank-per-process mode (neo_b200_bank_create_rank):
                     ncclAllGather of input rows, forward, ncclReduceScatter of partial spectra, inverse; pipelined submits.
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


def rel(got, want):
    return float(np.linalg.norm(got.astype(np.float64) - want) / np.linalg.norm(want))


def gloo_layouts(pkg, orc, rank, world):
    C, B, P, NB = 4, 64, 8, 16
    ir = orc.normalize_impulse(np.stack([orc.noise(B * P - 9, 11 + c, np.float32) for c in range(C)]))
    sig = np.stack([orc.noise(B * NB, 13 + c, np.float32) for c in range(C)])
    H = orc.uniform_partition(ir, B)
    want = orc.convolve_blocks(0, H, sig)
    worst = 0.0
    for layout in ((1, 2), (2, 1)):
        for frame in (0, 4):
            info = pkg.Bank.layout_info(pkg.UPOLS, "float32", pkg.DIAGONAL, C, C, B, P, frame or 2, frame, layout, rank)
            lo, hi = info["partition_begin"], info["partition_end"]
            g0, gn = info["group_first"], info["group_count"]
            if frame:
                assert lo % frame == 0 and info["delay_blocks"] == lo
            # the rank's partial result for its channel group: its partitions only, on an input delayed by `lo` blocks (linearity of
            # the delay line, fdl_index.hpp:24-36) -- what the delayed-input shard computes, in the time domain
            Hl = np.ascontiguousarray(H[g0 : g0 + gn, lo:hi])
            delayed = np.concatenate([np.zeros((gn, lo * B), np.float32), sig[g0 : g0 + gn]], axis=1)[:, : B * NB]
            part = np.zeros((C, B * NB), dtype=np.float64)
            part[g0 : g0 + gn] = orc.convolve_blocks(0, Hl, delayed).astype(np.float64)
            total = torch.from_numpy(part)
            dist.all_reduce(total)  # gloo has no reduce_scatter: all_reduce then slice is the same reduction
            o0, on = info["out_first"], info["out_count"]
            err = rel(total[o0 : o0 + on].numpy(), want[o0 : o0 + on])
            assert err < 1e-6, (layout, frame, err)
            worst = max(worst, err)
            covered = torch.zeros(C)
            covered[o0 : o0 + on] = 1
            dist.all_reduce(covered)
            assert torch.all(covered == 1), "every output row belongs to exactly one rank"
    return worst


def note(rank, *what):
    if rank == 0:
        print("dist_worker:", *what, file=sys.stderr, flush=True)


def nccl_bank(pkg, orc, rank, world, local):
    worst = 0.0
    uid_box = [pkg.bank_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid_box, src=0)
    # diagonal bank: both layouts of two ranks, direct form and frame mode, overlap-save and overlap-add, two steps in flight
    C, B, P, T, steps = 8, 64, 16, 4, 6
    ir = orc.normalize_impulse(np.stack([orc.noise(B * P - 9, 11 + c, np.float32) for c in range(C)]))
    sig = np.stack([orc.noise(B * T * steps, 13 + c, np.float32) for c in range(C)])
    H = orc.uniform_partition(ir, B)
    for kind in (pkg.UPOLS, pkg.UPOLA):
        want = orc.convolve_blocks(kind, H, sig)
        for layout in ((1, 2), (2, 1)):
            for frame in (0, T):
                note(rank, "diagonal", kind, layout, frame)
                uid = [pkg.bank_unique_id() if rank == 0 else None]
                dist.broadcast_object_list(uid, src=0)
                bank = pkg.Bank(kind, "float32", pkg.DIAGONAL, C, C, B, P, max_blocks=T, frame_blocks=frame, layout=layout, rank=rank,
                                world=world, unique_id=uid[0], device=local)
                info = bank.ranks[0]
                bank.impulse([np.ascontiguousarray(ir[info["group_first"] : info["group_first"] + info["group_count"]])])
                i0, n_in, o0, n_out = info["in_first"], info["in_count"], info["out_first"], info["out_count"]
                ins = [np.ascontiguousarray(sig[i0 : i0 + n_in, s * T * B : (s + 1) * T * B]) for s in range(steps)]
                outs = [np.zeros((n_out, T * B), np.float32) for _ in range(steps)]
                for s in range(steps):
                    bank.submit([ins[s]], [outs[s]])
                    if s >= 1:
                        bank.wait()
                bank.wait()
                got = np.concatenate(outs, axis=1)
                err = rel(got, want[o0 : o0 + n_out])
                assert err < 1e-5, (kind, layout, frame, err)
                worst = max(worst, err)
                # device buffers after a reset
                note(rank, "  reset + device buffers")
                bank.reset()
                got2 = []
                for s in range(steps):
                    x = torch.from_numpy(ins[s]).cuda()
                    y = torch.empty((n_out, T * B), device="cuda", dtype=torch.float32)
                    torch.cuda.synchronize()
                    bank([x], [y])
                    got2.append(y.cpu().numpy())
                assert rel(np.concatenate(got2, axis=1), want[o0 : o0 + n_out]) < 1e-5, (kind, layout, frame)
                bank.close()
    # a bank wide enough for the fused frame kernel, every exchange form between processes (NEO_B200_BANK_EXCHANGE): "dma" (copy engines
    # write into the owner's inbox through CUDA IPC mappings, a one-word all-gather as the cross-process gate), "kernel" (the fused frame
    # kernel stores into those mappings itself) and "collective" (ncclReduceScatter), against the oracle and each other
    C2, B2, P2, T2, steps2 = 64, 128, 128, 32, 5
    ir2 = orc.normalize_impulse(np.stack([orc.noise(B2 * P2 - 3, 11 + c, np.float32) for c in range(C2)]))
    sig2 = np.stack([orc.noise(B2 * T2 * steps2, 13 + c, np.float32) for c in range(C2)])
    want2 = orc.convolve_blocks(0, orc.uniform_partition(ir2, B2), sig2)
    forms = {}
    for form in ("dma", "kernel", "collective"):
        os.environ["NEO_B200_BANK_EXCHANGE"] = form
        note(rank, "exchange form", form)
        uid = [pkg.bank_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        bank = pkg.Bank(pkg.UPOLS, "float32", pkg.DIAGONAL, C2, C2, B2, P2, frame_blocks=T2, layout=(1, 2), rank=rank, world=world,
                        unique_id=uid[0], device=local)
        info = bank.ranks[0]
        bank.impulse([ir2])
        i0, n_in, o0, n_out = info["in_first"], info["in_count"], info["out_first"], info["out_count"]
        ins = [np.ascontiguousarray(sig2[i0 : i0 + n_in, s * T2 * B2 : (s + 1) * T2 * B2]) for s in range(steps2)]
        outs = [np.zeros((n_out, T2 * B2), np.float32) for _ in range(steps2)]
        for s in range(steps2):
            bank.submit([ins[s]], [outs[s]])
            if s >= 2:
                bank.wait()
        bank.wait()
        bank.wait()
        forms[form] = np.concatenate(outs, axis=1)
        err = rel(forms[form], want2[o0 : o0 + n_out])
        assert err < 1e-5, (form, err)
        worst = max(worst, err)
        forms[form + "_bytes"] = bank.device_bytes(0)
        bank.close()
    os.environ.pop("NEO_B200_BANK_EXCHANGE", None)
    assert forms["dma_bytes"] != forms["collective_bytes"] and forms["kernel_bytes"] != forms["collective_bytes"], "form not taken"
    assert np.array_equal(forms["dma"], forms["kernel"])
    # matrix topology sharded by output channel (BASELINE config 4's sharding) and by partitions
    O, I, L = 4, 2, B * 6
    irm = np.stack([np.stack([orc.noise(L, 100 + 10 * o + i, np.float32) for i in range(I)]) for o in range(O)])
    irm /= np.sqrt((irm**2).sum(axis=2).max())
    sigm = np.stack([orc.noise(B * T * steps, 13 + i, np.float32) for i in range(I)])
    wantm = np.zeros((O, B * T * steps))
    for o in range(O):
        wantm[o] = orc.convolve_blocks(0, orc.uniform_partition(irm[o], B), sigm).astype(np.float64).sum(axis=0)
    for layout in ((2, 1), (1, 2)):
        for frame in (0, 2):
            Tm = 2
            note(rank, "matrix", layout, frame)
            uid = [pkg.bank_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            bank = pkg.Bank(pkg.UPOLS, "float32", pkg.MATRIX, O, I, B, L // B, max_blocks=Tm, frame_blocks=frame, layout=layout, rank=rank,
                            world=world, unique_id=uid[0], device=local)
            info = bank.ranks[0]
            bank.impulse([np.ascontiguousarray(irm[info["group_first"] : info["group_first"] + info["group_count"]])])
            i0, n_in, o0, n_out = info["in_first"], info["in_count"], info["out_first"], info["out_count"]
            got = []
            for s in range(T * steps // Tm):
                x = np.ascontiguousarray(sigm[i0 : i0 + n_in, s * Tm * B : (s + 1) * Tm * B])
                y = np.zeros((n_out, Tm * B), np.float32)
                bank([x], [y])
                got.append(y)
            err = rel(np.concatenate(got, axis=1), wantm[o0 : o0 + n_out])
            assert err < 1e-5, ("matrix", layout, frame, err)
            worst = max(worst, err)
            bank.close()
    return worst


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", default="gloo")
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    orc = pyoracle.oracle()
    pkg = entry.load_package()
    if args.backend == "gloo":
        dist.init_process_group("gloo")
        err = gloo_layouts(pkg, orc, rank, world)
    else:
        local = int(os.environ.get("LOCAL_RANK", rank))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        pkg.set_device(local)
        err = nccl_bank(pkg, orc, rank, world, local)
    dist.barrier()
    if rank == 0:
        print(f"dist_worker ok backend={args.backend} world={world} err={err:.2e}")
    dist.destroy_process_group()


if __name__ == "__main__":
    try:
        main()
    except BaseException:  # a failed rank must not sit in a destructor waiting for a collective its peer will never join
        import traceback

        traceback.print_exc()
        sys.stderr.flush()
        os._exit(1)
