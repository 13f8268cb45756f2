"""TEST INFRASTRUCTURE ONLY: writes tests/golden/neo_ref_golden.npz from oracle/_ref/libneo_ref.so,
i.e. from the UNMODIFIED reference headers compiled in place (see oracle/Makefile, oracle/ref_wrapper.cpp).

Run in the build container (needs /root/reference):   python oracle/make_golden.py
The .npz is committed; the GPU box has no /root/reference and only reads the fixture.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pyoracle as po  # noqa: E402

po.build()
r = po.ref()
assert r is not None, "oracle/_ref/libneo_ref.so missing: run `make -C oracle ref` where /root/reference exists"

g = {}

# integer tables -- bit-exact contract (SURVEY 8a rows a3, a3', a11; stft/uniform_partition frame counts)
for order in range(0, 11):
    g[f"bitrev/{order}"] = r.bitrev_table(order)
for radix, size in [(2, 16), (3, 9), (3, 27), (4, 16), (4, 64), (4, 256), (5, 25), (8, 64), (8, 512)]:
    g[f"digitrev/{radix}/{size}"] = r.digitrev_perm(radix, size)
for parts in (1, 2, 3, 4, 7):
    wp, pairs = r.fdl_index_sequence(parts, 2 * parts + 3)
    g[f"fdl_index/{parts}/write_pos"] = wp
    g[f"fdl_index/{parts}/pairs"] = pairs
frames = [(1024, 128, 0), (1024, 256, 128), (4096, 128, 0), (4095, 128, 0), (1000, 128, 0), (128, 128, 0), (1 << 20, 1024, 0)]
g["stft_frames/args"] = np.array(frames, dtype=np.int64)
g["stft_frames/out"] = np.array([r.num_stft_frames(*a) for a in frames], dtype=np.int64)
g["next_order/args"] = np.array([1, 2, 3, 4, 5, 1023, 1024, 1025, 4095, 65536, 65537], dtype=np.int64)
g["next_order/out"] = np.array([r.next_order(int(n)) for n in g["next_order/args"]], dtype=np.int64)
g["fft_max_order"] = np.array([r.fft_max_order()], dtype=np.int64)
g["fft_status_28"] = np.array([r.fft_status(28)], dtype=np.int64)

# input distribution (testing/testing.hpp:37-72)
for seed in (1, 2, 11, 13):
    g[f"noise/f32/{seed}"] = r.noise(96, seed, np.float32)
    g[f"noise/f64/{seed}"] = r.noise(96, seed, np.float64)
g["noise/c64/1"] = r.noise(48, 1, np.complex64)

# twiddle LUTs (fft/twiddle.hpp:47-52)
for size in (2, 4, 16, 256):
    for d, name in ((-1, "fwd"), (1, "bwd")):
        g[f"twiddle/f32/{size}/{name}"] = r.twiddle_lut(size, d, np.float32)
        g[f"twiddle/f64/{size}/{name}"] = r.twiddle_lut(size, d, np.float64)

# c2c / r2c / c2r on the noise inputs the benchmarks use (seed 1 complex, seed 2 real)
for real, cplx, tag in ((np.float32, np.complex64, "f32"), (np.float64, np.complex128, "f64")):
    for order in (1, 2, 3, 4, 5, 8, 10, 11):
        n = 1 << order
        x = r.noise(n, 1, cplx)
        g[f"c2c/{tag}/{order}/fwd"] = r.fft(x, -1)
        g[f"c2c/{tag}/{order}/bwd"] = r.fft(x, +1)
        xr = r.noise(n, 2, real)
        spec = r.rfft(xr)
        g[f"r2c/{tag}/{order}"] = spec
        g[f"c2r/{tag}/{order}"] = r.irfft(spec, n)
    # c2r semantics on a NON-Hermitian-consistent spectrum (imag of DC/Nyquist must be ignored, full-length input accepted)
    n = 64
    junk = r.noise(n, 5, cplx)
    g[f"c2r_junk/{tag}/half"] = r.irfft(junk[: n // 2 + 1], n)
    g[f"c2r_junk/{tag}/full"] = r.irfft(junk, n)

# split-complex plan (fft/fallback/fallback_split_fft_plan.hpp) on the same noise: planes [re, im]
for real, cplx, tag in ((np.float32, np.complex64, "f32"), (np.float64, np.complex128, "f64")):
    for order in (1, 4, 8, 11):
        x = r.noise(1 << order, 1, cplx)
        for d, name in ((-1, "fwd"), (1, "bwd")):
            re, im = r.split_fft(x.real.copy(), x.imag.copy(), d)
            g[f"split/{tag}/{order}/{name}"] = np.stack([re, im])

# dft_plan (Bluestein, fft/fallback/fallback_dft_plan.hpp): sizes that are NOT powers of two, and one that is
for cplx, tag in ((np.complex64, "c64"), (np.complex128, "c128")):
    for n in (2, 3, 5, 12, 21, 100, 127, 128):
        x = r.noise(n, 70 + n, cplx)
        g[f"dft/{tag}/{n}/x"] = x
        g[f"dft/{tag}/{n}/fwd"] = r.dft(x, -1)
        g[f"dft/{tag}/{n}/bwd"] = r.dft(x, +1)

# stft_plan (fft/stft.hpp:39-109): (frame, transform, overlap, window kind 0 rect / 1 hann / 2 hamming), 2 channels x 1000 samples
for real, tag in ((np.float32, "f32"), (np.float64, "f64")):
    x = np.stack([r.noise(1000, 5 + c, real) for c in range(2)])
    g[f"stft/{tag}/x"] = x
    for frame, transform, overlap, win in ((256, 256, 128, 1), (128, 256, 0, 0), (100, 128, 30, 2), (64, 64, 48, 1), (256, 300, 17, 1)):
        g[f"stft/{tag}/{frame}_{transform}_{overlap}_{win}"] = r.stft(x, frame, transform, overlap, win)

# fft_convolve (convolution/fft_convolver.hpp:18-93), mode::full
for real, tag in ((np.float32, "f32"), (np.float64, "f64")):
    for n, m in ((2, 2), (7, 3), (100, 31), (513, 512)):
        x, h = r.noise(n, 5, real), r.noise(m, 6, real)
        g[f"fft_convolve/{tag}/{n}_{m}/signal"] = x
        g[f"fft_convolve/{tag}/{n}_{m}/patch"] = h
        g[f"fft_convolve/{tag}/{n}_{m}/out"] = r.fft_convolve(x, h)

# fallback_dct2_plan (fft/dct.hpp:24-68)
for real, tag in ((np.float32, "f32"), (np.float64, "f64")):
    for order in (1, 3, 6, 10):  # order 0 is undefined in the reference (order-0 c2c reads past its buffer)
        x = r.noise(1 << order, 90 + order, real)
        g[f"dct2/{tag}/{order}/x"] = x
        g[f"dct2/{tag}/{order}/out"] = r.dct2(x)
g["kat/dct2_1to8"] = r.dct2(np.arange(1, 9, dtype=np.float64))  # fft/dct_test.cpp:24-39

# known-answer inputs of the reference's own tests, run through the reference
g["kat/c2c_1234"] = r.fft(np.array([1, 2, 3, 4], dtype=np.complex64), -1)  # fft/rfft_test.cpp:170-186
delta = np.zeros(16, dtype=np.complex64)
delta[0] = 1
g["kat/c2c_delta16"] = r.fft(delta, -1)  # fft/rfft_test.cpp:132-168
x, y, z = (np.full(33, v, dtype=np.complex64) for v in (1 + 2j, 3 + 4j, 5 + 6j))
g["kat/multiply_add"] = r.multiply_add(x, y, z)  # algorithm/multiply_add_test.cpp:52-95 -> 0+16i

# filter preparation + convolvers: 2 channels, B=32, L=150 (ragged last partition, P=5), 12 blocks
B, L, NB = 32, 150, 12
for real, tag in ((np.float32, "f32"), (np.float64, "f64")):
    ir = np.stack([r.noise(L, 11 + c, real) for c in range(2)])
    if real == np.float32:
        ir = r.normalize_impulse(ir)
    else:
        ir = ir / np.sqrt((ir.astype(np.float64) ** 2).sum(axis=1).max())
    sig = np.stack([r.noise(B * NB, 13 + c, real) for c in range(2)])
    H = r.uniform_partition(ir, B)
    g[f"conv/{tag}/ir"] = ir
    g[f"conv/{tag}/signal"] = sig
    g[f"conv/{tag}/H"] = H
    for kind, name in enumerate(("upols", "upola", "split_upols", "split_upola", "upola_v2")):
        g[f"conv/{tag}/{name}"] = r.convolve_blocks(kind, H, sig)
    g[f"conv/{tag}/upola_v2_chunk96"] = r.convolve_blocks(4, H, sig, chunk=96)
g["conv/block"] = np.array([B, L, NB], dtype=np.int64)

# sparse convolvers (convolution/sparse_convolver.hpp:14-22) on the f32 case above: csr_matrix containers of channel 0 and the outputs,
# predicate |re| > threshold or |im| > threshold
THRESH = 0.25
H32 = g["conv/f32/H"]
rows, cols, vals = r.csr_build(H32[0], THRESH)
g["sparse/threshold"] = np.array([THRESH], dtype=np.float32)
g["sparse/csr_rows"], g["sparse/csr_cols"], g["sparse/csr_vals"] = rows, cols, vals
g["sparse/upols"] = r.convolve_blocks_sparse(5, H32, g["conv/f32/signal"], THRESH)
g["sparse/upola"] = r.convolve_blocks_sparse(6, H32, g["conv/f32/signal"], THRESH)

# compressed_fdl (convolution/compressed_fdl.hpp:17-52): the known-answer row of compressed_fdl_test.cpp:31-41 and noise in (-1, 1),
# inserted and read back through the reference's accessor, int8 and int16
kat = np.array([0 + 0.125j, 0.25 + 0.333j, 0.5 + 0.666j, 0.75 + 1j, -0.0 - 0.125j, -0.25 - 0.333j, -0.5 - 0.666j, -0.75 - 1j], dtype=np.complex64)
cn = r.noise(257, 77, np.complex64)
g["cfdl/kat_in"], g["cfdl/noise_in"] = kat, cn
for bits in (8, 16):
    g[f"cfdl/{bits}/kat_out"] = r.compressed_fdl_roundtrip(kat, bits)
    g[f"cfdl/{bits}/noise_out"] = r.compressed_fdl_roundtrip(cn, bits)

# overlap policies with an identity callback (convolution/overlap_test.cpp:21-64)
sig = r.noise(128 * 6, 3, np.float32)
g["overlap/signal"] = sig
g["overlap/save"] = r.overlap_identity(False, 128, 64, sig)
g["overlap/add"] = r.overlap_identity(True, 128, 64, sig)

out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "neo_ref_golden.npz")
np.savez_compressed(out, **g)
print(f"wrote {out}: {len(g)} arrays, {os.path.getsize(out)/1024:.1f} KiB")
