"""GPU: the multi-device bank (neo_b200_bank_*) against the oracle (= one reference convolver per channel,
uniform_partitioned_convolver.hpp:48-65). On a one-GPU box every rank is placed on device 0 (`devices=[0, 0, ...]`): the layout
arithmetic, the input ring with delayed slots, the row exchange, the reduction fused into the c2r kernel and the software pipeline
are exactly what runs on N GPUs -- only the peer pointers happen to be local. With more GPUs the ranks spread over them."""
import numpy as np
import pytest

from conftest import TOL, rel_l2
from test_conv_gpu import make_case

pytestmark = pytest.mark.gpu


def devices_for(gpu, n):
    have = gpu.device_count()
    return [i % have for i in range(n)]


def run_bank_steps(bank, sig, block, T, pipelined=False):
    """feed sig[rows][n] through the bank T blocks per call; pipelined: three submits in flight, separate buffers per step"""
    n_in = sig.shape[0]
    n_out = sum(r["out_count"] for r in bank.ranks)
    steps = sig.shape[1] // (T * block)
    outs = [np.zeros((n_out, T * block), dtype=sig.dtype) for _ in range(steps)]
    ins = [np.ascontiguousarray(sig[:, s * T * block : (s + 1) * T * block]) for s in range(steps)]
    if pipelined:
        for s in range(steps):
            bank.submit(ins[s], outs[s])
            if s >= 2:
                bank.wait()
        bank.wait()
        bank.wait()
    else:
        for s in range(steps):
            bank(ins[s], outs[s])
    return np.concatenate(outs, axis=1)


@pytest.mark.parametrize("layout", [(1, 1), (2, 1), (1, 2), (2, 2), (1, 4), (4, 2)])
@pytest.mark.parametrize("frame", [0, 4])
def test_bank_layouts_match_oracle(gpu, orc, layout, frame):
    C, B, P, T, steps = 8, 64, 16, 4, 7
    ir, sig = make_case(orc, C, B * P - 9, B, T * steps)
    want = orc.convolve_blocks(0, orc.uniform_partition(ir, B), sig)
    n = layout[0] * layout[1]
    bank = gpu.Bank(gpu.UPOLS, np.float32, gpu.DIAGONAL, C, C, B, P, max_blocks=T, frame_blocks=frame, layout=layout,
                    devices=devices_for(gpu, n))
    assert [r["rank"] for r in bank.ranks] == list(range(n))
    assert sum(r["out_count"] for r in bank.ranks) == C
    if frame and layout[1] > 1:
        assert bank.ranks[1]["delay_blocks"] == bank.ranks[1]["partition_begin"] > 0  # shard 1 reads a delayed input slot
    bank.impulse_global(ir)
    got = run_bank_steps(bank, sig, B, T)
    assert rel_l2(got, want) <= 1e-5, rel_l2(got, want)
    for s in range(steps):  # every step on its own: a wrong slot / ring pairing must not hide behind the whole-signal norm
        sl = slice(s * T * B, (s + 1) * T * B)
        assert rel_l2(got[:, sl], want[:, sl]) <= 2e-5, s
    # same state machine after reset, now software-pipelined (three steps in flight)
    bank.reset()
    got2 = run_bank_steps(bank, sig, B, T, pipelined=True)
    assert rel_l2(got2, want) <= 1e-5
    bank.close()


@pytest.mark.parametrize("exchange", ["dma", "kernel"])
def test_bank_long_frames_partition_shards_take_the_32_point_fused_kernel(gpu, orc, monkeypatch, exchange):
    # BASELINE config 5's sharded geometry in small: T = 256 (L = 512 frame transforms), P = 1024 over two partition shards = two
    # second-level partitions per shard, shard 1 on a delayed input -- the case the 32-points-per-thread fused frame kernel is taken
    # for (conv_frame.cuh); enough steps for shard 1's delayed history to reach the output
    monkeypatch.setenv("NEO_B200_BANK_EXCHANGE", exchange)
    C, B, P, T, steps = 4, 16, 1024, 256, 6
    ir, sig = make_case(orc, C, B * P - 3, B, T * steps)
    want = orc.convolve_blocks(0, orc.uniform_partition(ir, B), sig)
    bank = gpu.Bank(gpu.UPOLS, np.float32, gpu.DIAGONAL, C, C, B, P, max_blocks=T, frame_blocks=T, layout=(2, 2), devices=devices_for(gpu, 4))
    assert bank.ranks[1]["delay_blocks"] == 512
    bank.impulse_global(ir)
    got = run_bank_steps(bank, sig, B, T, pipelined=True)
    bank.close()
    for s in range(steps):
        sl = slice(s * T * B, (s + 1) * T * B)
        assert rel_l2(got[:, sl], want[:, sl]) <= 2e-5, s


@pytest.mark.parametrize("real", [np.float32, np.float64])
@pytest.mark.parametrize("kind", [0, 1])
def test_bank_upola_and_f64(gpu, orc, real, kind):
    C, B, P, T, steps = 4, 32, 12, 2, 9
    ir, sig = make_case(orc, C, B * P - 5, B, T * steps, real)
    want = orc.convolve_blocks(kind, orc.uniform_partition(ir, B), sig)
    for frame, layout in ((0, (2, 2)), (2, (1, 2)), (2, (2, 2))):
        bank = gpu.Bank(kind, real, gpu.DIAGONAL, C, C, B, P, max_blocks=T, frame_blocks=frame, layout=layout, devices=devices_for(gpu, 4 if layout == (2, 2) else 2))
        bank.impulse_global(ir)
        got = run_bank_steps(bank, sig, B, T, pipelined=True)
        assert rel_l2(got, want) <= TOL[np.dtype(real).name], (frame, layout, rel_l2(got, want))
        bank.close()


def test_bank_ragged_partition_shards_and_variable_call_lengths(gpu, orc):
    # P = 11 over 3 shards (direct form: 3/4/4 partitions), calls of 1, 3 and 2 blocks
    C, B, P = 6, 128, 11
    pattern = [1, 3, 2, 3, 1, 2, 3]
    ir, sig = make_case(orc, C, B * P - 40, B, sum(pattern))
    want = orc.convolve_blocks(0, orc.uniform_partition(ir, B), sig)
    bank = gpu.Bank(gpu.UPOLS, np.float32, gpu.DIAGONAL, C, C, B, P, max_blocks=3, layout=(2, 3), devices=devices_for(gpu, 6))
    bank.impulse_global(ir)
    got = np.zeros_like(sig)
    pos = 0
    for t in pattern:
        x = np.ascontiguousarray(sig[:, pos * B : (pos + t) * B])
        y = np.zeros_like(x)
        bank(x, y)
        got[:, pos * B : (pos + t) * B] = y
        pos += t
    assert rel_l2(got, want) <= 1e-5
    bank.close()


def test_bank_matrix_topology_sharded_by_output_channel(gpu, orc):
    # BASELINE config 4's sharding on a slice: every rank holds H[o in its group][all inputs]
    O, I, B, L, T, steps = 8, 4, 64, 64 * 6, 2, 6
    ir = np.stack([np.stack([orc.noise(L, 100 + 10 * o + i, np.float32) for i in range(I)]) for o in range(O)])
    ir /= np.sqrt((ir**2).sum(axis=2).max())
    sig = np.stack([orc.noise(B * T * steps, 13 + i, np.float32) for i in range(I)])
    want = np.zeros((O, B * T * steps), dtype=np.float64)
    for o in range(O):
        want[o] = orc.convolve_blocks(0, orc.uniform_partition(ir[o], B), sig).astype(np.float64).sum(axis=0)
    for frame, layout in ((0, (4, 1)), (0, (2, 2)), (2, (4, 1)), (2, (2, 2))):
        bank = gpu.Bank(gpu.UPOLS, np.float32, gpu.MATRIX, O, I, B, L // B, max_blocks=T, frame_blocks=frame, layout=layout, devices=devices_for(gpu, 4))
        bank.impulse_global(ir)
        got = run_bank_steps(bank, sig, B, T, pipelined=True)
        assert rel_l2(got, want) <= 1e-5, (frame, layout, rel_l2(got, want))
        bank.close()


def test_bank_device_buffers_and_profile(gpu, orc):
    import torch

    C, B, P, T, steps = 4, 256, 8, 4, 5
    ir, sig = make_case(orc, C, B * P, B, T * steps)
    want = orc.convolve_blocks(0, orc.uniform_partition(ir, B), sig)
    bank = gpu.Bank(gpu.UPOLS, np.float32, gpu.DIAGONAL, C, C, B, P, frame_blocks=T, layout=(1, 2), devices=devices_for(gpu, 2))
    bank.impulse_global(ir)
    bank.profile(True)
    got = np.zeros_like(sig)
    for s in range(steps):
        xs, ys = [], []
        for r in bank.ranks:
            with torch.cuda.device(r["device"]):
                rows = sig[r["in_first"] : r["in_first"] + r["in_count"], s * T * B : (s + 1) * T * B]
                xs.append(torch.from_numpy(np.ascontiguousarray(rows)).cuda())
                ys.append(torch.empty((r["out_count"], T * B), device="cuda", dtype=torch.float32))
        torch.cuda.synchronize()
        bank(xs, ys)
        for r, y in zip(bank.ranks, ys):
            got[r["out_first"] : r["out_first"] + r["out_count"], s * T * B : (s + 1) * T * B] = y.cpu().numpy()
    assert rel_l2(got, want) <= 1e-5
    ms = bank.profile_read(0)
    assert ms[0] > 0 and ms[1] > 0 and ms[2] > 0 and ms[5] == steps
    assert bank.device_bytes(0) > 0
    bank.close()


@pytest.mark.parametrize("block,T,P,channels,layout", [(256, 8, 32, 64, (1, 2)), (128, 32, 128, 64, (2, 2)), (128, 32, 128, 32, (1, 4))])
def test_bank_exchange_forms_agree(gpu, orc, monkeypatch, block, T, P, channels, layout):
    # three ways for the partial spectra of a group to reach the rank that finishes a channel (NEO_B200_BANK_EXCHANGE), on banks wide
    # enough for the fused frame kernel: "dma" (default: copy engines move each owner's rows into its inbox, the c2r kernel sums its own
    # rows and the inbox slots), "kernel" (frame_fused_kernel stores the rows into the owners' inboxes itself: peer stores) and
    # "collective" (the c2r kernel loads the shards' buffers through peer pointers). Same bits, and the oracle's answer.
    steps = 5
    ir, sig = make_case(orc, channels, block * P - 3, block, T * steps)
    want = orc.convolve_blocks(0, orc.uniform_partition(ir, block), sig)
    n = layout[0] * layout[1]
    results, sizes = {}, {}
    for form in ("dma", "kernel", "collective"):
        monkeypatch.setenv("NEO_B200_BANK_EXCHANGE", form)
        bank = gpu.Bank(gpu.UPOLS, np.float32, gpu.DIAGONAL, channels, channels, block, P, frame_blocks=T, layout=layout, devices=devices_for(gpu, n))
        bank.impulse_global(ir)
        results[form] = run_bank_steps(bank, sig, block, T, pipelined=True)
        sizes[form] = bank.device_bytes(0)
        assert rel_l2(results[form], want) <= 1e-5, (form, rel_l2(results[form], want))
        for st in range(steps):
            sl = slice(st * T * block, (st + 1) * T * block)
            assert rel_l2(results[form][:, sl], want[:, sl]) <= 2e-5, (form, st)
        bank.close()
    monkeypatch.delenv("NEO_B200_BANK_EXCHANGE")
    assert sizes["dma"] > sizes["collective"] and sizes["kernel"] > sizes["collective"], "inboxes were not allocated: form not taken"
    assert np.array_equal(results["dma"], results["collective"]) and np.array_equal(results["kernel"], results["collective"])


def test_bank_dma_exchange_direct_form_and_upola(gpu, orc):
    # the default exchange works on any kernel path: direct form with ragged call lengths, overlap-add, float64
    C, B, P = 8, 64, 12
    pattern = [2, 3, 1, 3, 3, 2]
    for kind, real in ((gpu.UPOLS, np.float32), (gpu.UPOLA, np.float64)):
        ir, sig = make_case(orc, C, B * P - 7, B, sum(pattern), real)
        want = orc.convolve_blocks(kind, orc.uniform_partition(ir, B), sig)
        bank = gpu.Bank(kind, real, gpu.DIAGONAL, C, C, B, P, max_blocks=3, layout=(2, 2), devices=devices_for(gpu, 4))
        bank.impulse_global(ir)
        got = np.zeros_like(sig)
        pos = 0
        for t in pattern:
            x = np.ascontiguousarray(sig[:, pos * B : (pos + t) * B])
            y = np.zeros_like(x)
            bank(x, y)
            got[:, pos * B : (pos + t) * B] = y
            pos += t
        assert rel_l2(got, want) <= TOL[np.dtype(real).name], (kind, rel_l2(got, want))
        bank.close()


def test_bank_error_contract(gpu):
    with pytest.raises(RuntimeError):  # layout does not match the number of devices
        gpu.Bank(gpu.UPOLS, np.float32, gpu.DIAGONAL, 8, 8, 64, 8, layout=(2, 2), devices=[0, 0])
    with pytest.raises(RuntimeError):  # channels not a multiple of the ranks
        gpu.Bank(gpu.UPOLS, np.float32, gpu.DIAGONAL, 6, 6, 64, 8, layout=(4, 1), devices=[0, 0, 0, 0])
    with pytest.raises(RuntimeError):  # more shards than frames of partitions
        gpu.Bank(gpu.UPOLS, np.float32, gpu.DIAGONAL, 8, 8, 64, 8, frame_blocks=4, layout=(1, 4), devices=[0, 0, 0, 0])
    bank = gpu.Bank(gpu.UPOLS, np.float32, gpu.DIAGONAL, 4, 4, 64, 8, layout=(2, 1), devices=[0, 0])
    with pytest.raises(RuntimeError):  # no filter yet
        bank(np.zeros((4, 64), np.float32), np.zeros((4, 64), np.float32))
    bank.close()
