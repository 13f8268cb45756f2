"""GPU: frame mode (neo_b200_conv_config::frame_blocks = T) -- the sum over partitions evaluated by a second overlap-save level
along block time -- must reproduce the reference's block-by-block upols/upola convolvers (the oracle) to the same tolerance as
the direct form: different operation order, same mathematics."""
import numpy as np
import pytest

from conftest import TOL, rel_l2
from test_conv_gpu import make_case, run_bank

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,kind", [("upols", 0), ("upola", 1)])
@pytest.mark.parametrize("tag,real", [("f32", np.float32), ("f64", np.float64)])
def test_golden_vectors_from_the_reference(gpu, golden, name, kind, tag, real):
    B, L, NB = (int(v) for v in golden["conv/block"])
    H, sig, want = golden[f"conv/{tag}/H"], golden[f"conv/{tag}/signal"], golden[f"conv/{tag}/{name}"]
    for T in (2, 4):
        if NB % T:
            continue
        conv = gpu.Convolver(kind, real, gpu.DIAGONAL, frame_blocks=T)
        conv.filter(H)
        got = run_bank(conv, sig, B, [T])
        assert rel_l2(got, want) <= TOL[np.dtype(real).name], T
        conv.close()


@pytest.mark.parametrize("block,taps,channels,T,frames", [
    (2, 7, 3, 2, 5),                  # P = 4, L = 4: smallest of everything
    (16, 100, 3, 4, 6),               # P = 7 -> Q = 2, last second-level partition ragged
    (128, 128 * 9 - 3, 3, 4, 5),      # P = 9 -> Q = 3
    (128, 1000, 3, 8, 4),             # P = 8 = T: one second-level partition
    (512, 512 * 17, 2, 16, 3),        # P = 17 -> Q = 2
    (1024, 1024 * 33 - 5, 2, 32, 3),  # BASELINE config 5's block size
    (64, 64 * 5, 2, 64, 3),           # T > P: L = 128 frame transform, Nyquist tile wider than one tile of bins
    (256, 256 * 3, 1, 512, 2),        # largest frame: L = 1024
    (4096, 4096 * 3, 1, 2, 3),
])
def test_frame_mode_matches_oracle(gpu, orc, block, taps, channels, T, frames):
    ir, sig = make_case(orc, channels, taps, block, T * frames)
    H = orc.uniform_partition(ir, block)
    want = orc.convolve_blocks(0, H, sig)
    conv = gpu.Convolver(gpu.UPOLS, np.float32, gpu.DIAGONAL, frame_blocks=T)
    conv.filter(H)
    got = run_bank(conv, sig, block, [T])
    assert rel_l2(got, want) <= 1e-5, rel_l2(got, want)
    # every frame on its own is right too (a wrong ring pairing would still pass a whole-signal norm if late frames dominate)
    for f in range(frames):
        sl = slice(f * T * block, (f + 1) * T * block)
        assert rel_l2(got[:, sl], want[:, sl]) <= 2e-5, f
    # same state machine from time-domain impulse responses, and after reset()
    conv.impulse(ir, block)
    assert rel_l2(run_bank(conv, sig, block, [T]), want) <= 1e-5
    conv.reset()
    assert rel_l2(run_bank(conv, sig, block, [T]), want) <= 1e-5
    conv.close()


@pytest.mark.parametrize("block,T,P,channels", [(1024, 64, 128, 16), (512, 128, 200, 12), (256, 512, 700, 8)])
def test_frame_mode_persistent_pipelined_kernel_many_units(gpu, orc, monkeypatch, block, T, P, channels):
    # banks with more units (16 adjacent bins of a channel) than resident CTAs: the persistent software-pipelined fused kernel walks
    # several units per CTA, prefetching the next unit's spectra and first MAC chunks (taken with at most two second-level partitions:
    # P <= 2T). Same bits as the one-CTA-per-unit form
    # (NEO_B200_FRAME_NO_PIPELINE), and the oracle's answer.
    steps = 3
    ir, sig = make_case(orc, channels, block * P - 11, block, T * steps)
    H = orc.uniform_partition(ir, block)
    want = orc.convolve_blocks(0, H, sig)
    got = {}
    for form in ("pipelined", "plain"):
        if form == "plain":
            monkeypatch.setenv("NEO_B200_FRAME_NO_PIPELINE", "1")
        conv = gpu.Convolver(gpu.UPOLS, np.float32, gpu.DIAGONAL, frame_blocks=T)
        conv.filter(H)
        got[form] = run_bank(conv, sig, block, [T])
        conv.close()
        assert rel_l2(got[form], want) <= 1e-5, (form, rel_l2(got[form], want))
    monkeypatch.delenv("NEO_B200_FRAME_NO_PIPELINE")
    assert np.array_equal(got["pipelined"], got["plain"])


@pytest.mark.parametrize("block,taps,T", [(16, 100, 4), (256, 256 * 9 - 3, 8)])
def test_frame_mode_upola_and_f64(gpu, orc, block, taps, T):
    for real in (np.float32, np.float64):
        ir, sig = make_case(orc, 2, taps, block, T * 4, real)
        H = orc.uniform_partition(ir, block)
        for kind in (gpu.UPOLS, gpu.UPOLA):
            conv = gpu.Convolver(kind, real, gpu.DIAGONAL, frame_blocks=T)
            conv.filter(H)
            got = run_bank(conv, sig, block, [T])
            assert rel_l2(got, orc.convolve_blocks(kind, H, sig)) <= TOL[np.dtype(real).name], (kind, real)
            conv.close()


@pytest.mark.parametrize("outputs", [2, 4])  # 4: the four-outputs-per-thread streaming kernel
def test_frame_mode_matrix_topology(gpu, orc, outputs):
    O, I, B, L, T, frames = outputs, 3, 64, 64 * 9, 4, 4
    NB = T * frames
    ir = np.stack([np.stack([orc.noise(L, 100 + 10 * o + i, np.float32) for i in range(I)]) for o in range(O)])
    ir /= np.sqrt((ir**2).sum(axis=2).max())
    sig = np.stack([orc.noise(B * NB, 13 + i, np.float32) for i in range(I)])
    want = np.zeros((O, B * NB), dtype=np.float64)
    for o in range(O):
        want[o] = orc.convolve_blocks(0, orc.uniform_partition(ir[o], B), sig).astype(np.float64).sum(axis=0)
    Hm = np.stack([orc.uniform_partition(ir[o], B) for o in range(O)])
    conv = gpu.Convolver(gpu.UPOLS, np.float32, gpu.MATRIX, frame_blocks=T)
    conv.filter(Hm)
    got = np.zeros((O, B * NB), dtype=np.float32)
    for f in range(frames):
        got[:, f * T * B : (f + 1) * T * B] = conv(np.ascontiguousarray(sig[:, f * T * B : (f + 1) * T * B]))
    assert rel_l2(got, want) <= 1e-5
    conv.close()


def test_frame_mode_partition_sharded_handles_sum_to_the_whole(gpu, orc):
    import torch

    C, B, P, T, frames = 4, 128, 12, 2, 9
    ir, sig = make_case(orc, C, B * P - 17, B, T * frames)
    H = orc.uniform_partition(ir, B)
    want = orc.convolve_blocks(0, H, sig)
    convs = []
    for lo, hi in [(0, 4), (4, 6), (6, 12)]:
        c = gpu.Convolver(gpu.UPOLS, np.float32, gpu.DIAGONAL, partition_range=(lo, hi), frame_blocks=T)
        c.filter(H)
        convs.append(c)
    got = np.zeros_like(sig)
    for pos in range(0, T * frames, T):
        x = torch.from_numpy(np.ascontiguousarray(sig[:, pos * B : (pos + T) * B])).cuda()
        total = None
        for c in convs:
            c.forward(x)
            c.synchronize()
            part = c.spectra_tensor(T).clone()
            total = part if total is None else total + part
        y = torch.empty_like(x)
        convs[1].inverse(total.contiguous(), y, 0, C, T)
        convs[1].synchronize()
        got[:, pos * B : (pos + T) * B] = y.cpu().numpy()
    assert rel_l2(got, want) <= 1e-5
    with pytest.raises(RuntimeError):  # a shard must start on a frame boundary of the partition axis
        c = gpu.Convolver(gpu.UPOLS, np.float32, gpu.DIAGONAL, partition_range=(3, 12), frame_blocks=T)
        c.filter(H)


def test_frame_mode_wide_bank_host_pipeline_and_device_path(gpu, orc):
    import torch

    C, B, P, T, frames = 160, 64, 5, 4, 4
    ir, sig = make_case(orc, C, B * P - 3, B, T * frames)
    H = orc.uniform_partition(ir, B)
    want = orc.convolve_blocks(0, H, sig)
    conv = gpu.Convolver(gpu.UPOLS, np.float32, gpu.DIAGONAL, frame_blocks=T)
    conv.filter(H)
    assert rel_l2(run_bank(conv, sig, B, [T]), want) <= 1e-5  # host buffers: channel groups over three streams
    conv.close()
    conv = gpu.Convolver(gpu.UPOLS, np.float32, gpu.DIAGONAL, frame_blocks=T)
    conv.set_stream(torch.cuda.current_stream())
    conv.filter(torch.from_numpy(H).cuda())
    dx = torch.from_numpy(sig).cuda()
    dy = torch.empty_like(dx)
    for pos in range(0, T * frames, T):
        chunk = dx[:, pos * B : (pos + T) * B].contiguous()
        out = torch.empty_like(chunk)
        conv(chunk, out=out)
        dy[:, pos * B : (pos + T) * B] = out
    torch.cuda.synchronize()
    assert rel_l2(dy.cpu().numpy(), want) <= 1e-5
    conv.close()


def test_frame_mode_small_bank_splits_rows_across_ctas(gpu, orc):
    # one channel, many second-level partitions: the streamed rows are split over CTAs and folded by the last one
    B, P, T, frames = 32, 96, 2, 6
    ir, sig = make_case(orc, 1, B * P, B, T * frames)
    H = orc.uniform_partition(ir, B)
    conv = gpu.Convolver(gpu.UPOLS, np.float32, gpu.DIAGONAL, frame_blocks=T)
    conv.filter(H)
    assert rel_l2(run_bank(conv, sig, B, [T]), orc.convolve_blocks(0, H, sig)) <= 1e-5
    conv.close()


def test_frame_mode_agrees_with_direct_form_over_a_grid_of_shapes(gpu, orc):
    # geometry sweep: every combination of block size (tile narrower / as wide as / wider than 16 bins, several tiles), frame length
    # (L below, at and above the tile width), partition count (fewer / more than a frame) and channel count (ragged bin groups):
    # frame mode against the direct form on the same device (both against the oracle for the smallest block sizes)
    rng = np.random.default_rng(7)
    for kind in (gpu.UPOLS, gpu.UPOLA):
        for B in (2, 4, 32, 128, 256):
            for T in (2, 8, 64):
                for P in (1, 3, 10):
                    for C in (1, 5):
                        if kind == gpu.UPOLA and (B, T) not in ((4, 8), (128, 2), (256, 64)):
                            continue
                        taps, frames = B * P - (B // 2 if P > 1 else 0), 3
                        ir = rng.uniform(-1, 1, size=(C, taps)).astype(np.float32)
                        ir /= np.sqrt((ir.astype(np.float64) ** 2).sum(axis=1).max())
                        sig = rng.uniform(-1, 1, size=(C, B * T * frames)).astype(np.float32)
                        direct = gpu.Convolver(kind, np.float32, gpu.DIAGONAL, max_blocks=T)
                        direct.impulse(ir, B)
                        want = run_bank(direct, sig, B, [T])
                        direct.close()
                        framed = gpu.Convolver(kind, np.float32, gpu.DIAGONAL, frame_blocks=T)
                        framed.impulse(ir, B)
                        got = run_bank(framed, sig, B, [T])
                        framed.close()
                        assert rel_l2(got, want) <= 5e-6, (kind, B, T, P, C, rel_l2(got, want))
                        if B <= 4:
                            H = orc.uniform_partition(ir, B)
                            assert rel_l2(got, orc.convolve_blocks(kind, H, sig)) <= 1e-5, (kind, B, T, P, C)


def test_frame_mode_full_size_north_star_shape(gpu, orc):
    # BASELINE config 5 geometry (B = 1024, 2^20 taps -> P = 1024) at bench.py's default frame length T = 256 (L = 512, Q = 4,
    # the cp.async-staged fused kernel). The oracle would need minutes here, so the check is the size-independent property:
    # an impulse response that is a unit impulse at tap d delays the input by d samples -- delays reach across partitions,
    # second-level partitions and frames (three frames = 768 blocks are streamed).
    B, L, T, frames, C = 1024, 1 << 20, 256, 3, 16
    n = B * T * frames
    delays = [0, 1, B - 1, B, B * T - 1, B * T, B * T + 5, 2 * B * T - 7, n - 1] + [(c * 1000003 + 11) % (n // 2) for c in range(C - 9)]
    ird = np.zeros((C, L), dtype=np.float32)
    for c, d in enumerate(delays):
        ird[c, d] = 1
    rng = np.random.default_rng(5)
    sigd = rng.uniform(-1, 1, size=(C, n)).astype(np.float32)
    conv = gpu.Convolver(gpu.UPOLS, np.float32, gpu.DIAGONAL, frame_blocks=T)
    conv.impulse(ird, B)
    out = run_bank(conv, sigd, B, [T])
    for c, d in enumerate(delays):
        expect = np.concatenate([np.zeros(d, dtype=np.float32), sigd[c, : n - d]])
        assert np.allclose(out[c], expect, atol=5e-5), (c, d, float(np.abs(out[c] - expect).max()))
    conv.close()

    # and full oracle parity at the same geometry on a 2-channel slice with random impulse responses (2 frames = 512 blocks: the
    # reference's block-by-block convolver needs a few seconds for that on the CPU)
    ir, sig = make_case(orc, 2, L, B, 2 * T)
    conv = gpu.Convolver(gpu.UPOLS, np.float32, gpu.DIAGONAL, frame_blocks=T)
    conv.impulse(ir, B)
    got = run_bank(conv, sig, B, [T])
    conv.close()
    want = orc.convolve_blocks(0, orc.uniform_partition(ir, B), sig)
    assert rel_l2(got, want) <= 1e-5, rel_l2(got, want)


@pytest.mark.parametrize("knob,value", [(None, None), ("NEO_B200_FRAME_NO_ASYNC", "1"), ("NEO_B200_FRAME_VARIANT", "1"),
                                        ("NEO_B200_FRAME_VARIANT", "2"), ("NEO_B200_FRAME_VARIANT", "3"), ("NEO_B200_FRAME_VARIANT", "4"),
                                        ("NEO_B200_FRAME_VARIANT", "5")])
def test_fused_frame_kernel_geometries_at_long_frames(gpu, orc, monkeypatch, knob, value):
    # L = 256 and L = 512 frame transforms: the shipped geometry (16 points per thread, cp.async-staged MAC) and the alternatives the
    # knobs select when the handle is created (tuning knobs) must all reproduce the reference
    if knob is not None:
        monkeypatch.setenv(knob, value)
    for B, P, T, frames in ((128, 20, 128, 2), (128, 5, 256, 2)):
        ir, sig = make_case(orc, 2, B * P - 9, B, T * frames)
        H = orc.uniform_partition(ir, B)
        conv = gpu.Convolver(gpu.UPOLS, np.float32, gpu.DIAGONAL, frame_blocks=T)
        conv.filter(H)
        got = run_bank(conv, sig, B, [T])
        assert rel_l2(got, orc.convolve_blocks(0, H, sig)) <= 1e-5, (knob, value, T)
        conv.close()


@pytest.mark.parametrize("variant", ["0", "5"])
@pytest.mark.parametrize("B,P,T,frames,channels", [(128, 600, 256, 3, 2), (128, 1100, 512, 2, 2), (16, 1500, 256, 7, 3)])
def test_fused_frame_kernel_32_points_per_thread_several_partitions(gpu, orc, monkeypatch, variant, B, P, T, frames, channels):
    # L = 512 / 1024 frame transforms with 32 points per thread (two register stages, one exchange, the sum over the second-level
    # partitions formed in place): Q = 3, 3 and 6 second-level partitions, ring wrap, ragged last partition
    monkeypatch.setenv("NEO_B200_FRAME_VARIANT", variant)
    ir, sig = make_case(orc, channels, B * P - 5, B, T * frames)
    H = orc.uniform_partition(ir, B)
    conv = gpu.Convolver(gpu.UPOLS, np.float32, gpu.DIAGONAL, frame_blocks=T)
    conv.filter(H)
    got = run_bank(conv, sig, B, [T])
    conv.close()
    assert rel_l2(got, orc.convolve_blocks(0, H, sig)) <= 1e-5, (variant, T)


def test_frame_mode_three_kernel_form_for_banks(gpu, orc, monkeypatch):
    # banks normally take the fused kernel; NEO_B200_FRAME_UNFUSED (read when the handle is created) selects the three-kernel form
    # (frame transform, frame_mac_kernel over several columns per thread, inverse transform) -- same results
    monkeypatch.setenv("NEO_B200_FRAME_UNFUSED", "1")
    for block, taps, T, frames in ((128, 128 * 9 - 3, 4, 5), (1024, 1024 * 33 - 5, 32, 3), (64, 64 * 5, 64, 3)):
        ir, sig = make_case(orc, 3, taps, block, T * frames)
        H = orc.uniform_partition(ir, block)
        conv = gpu.Convolver(gpu.UPOLS, np.float32, gpu.DIAGONAL, frame_blocks=T)
        conv.filter(H)
        assert rel_l2(run_bank(conv, sig, block, [T]), orc.convolve_blocks(0, H, sig)) <= 1e-5, (block, T)
        conv.close()


def test_frame_mode_error_contract(gpu, orc):
    H = np.zeros((2, 3, 65), dtype=np.complex64)
    for bad in (3, 1, 1024):
        with pytest.raises(RuntimeError):
            gpu.Convolver(gpu.UPOLS, np.float32, frame_blocks=bad).filter(H)
    conv = gpu.Convolver(gpu.UPOLS, np.float32, frame_blocks=4)
    conv.filter(H)
    with pytest.raises(RuntimeError):
        conv(np.zeros((2, 64 * 2), dtype=np.float32))  # not a whole frame
    with pytest.raises(RuntimeError):
        conv(np.zeros((2, 64 * 8), dtype=np.float32))
    conv(np.zeros((2, 64 * 4), dtype=np.float32))
    conv.close()
