"""TEST INFRASTRUCTURE ONLY: ctypes access to the two CPU checkers.

* ``oracle()``  -> liboracle.so       plain-C restatement (oracle/neo_oracle.c), built by ``make -C oracle oracle``
* ``ref()``     -> _ref/libneo_ref.so  the unmodified reference headers compiled in place (oracle/ref_wrapper.cpp);
                                       ``None`` when the prebuilt library is absent (it cannot be rebuilt on the GPU box).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg import this module.
"""
from __future__ import annotations

import ctypes as C
import functools
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "liboracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libneo_ref.so")
REF_V4_SO = os.path.join(_HERE, "_ref", "libneo_ref_v4.so")  # the same translation unit built -march=x86-64-v4 (AVX-512 hosts)

_vp, _sz, _i, _u32 = C.c_void_p, C.c_size_t, C.c_int, C.c_uint32


def build(force: bool = False) -> None:
    """(Re)build the checkers. `_ref` only where /root/reference exists."""
    if force or not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < max(
        os.path.getmtime(os.path.join(_HERE, f)) for f in ("neo_oracle.c", "neo_oracle_impl.inc", "neo_oracle.h")
    ):
        subprocess.run(["make", "-C", _HERE, "-B", "oracle"], check=True, capture_output=True)
    if os.path.isdir("/root/reference/src/neo"):
        stale = (not os.path.exists(REF_SO)) or os.path.getmtime(REF_SO) < os.path.getmtime(
            os.path.join(_HERE, "ref_wrapper.cpp")
        )
        if force or stale:
            subprocess.run(["make", "-C", _HERE, "-B", "_ref/libneo_ref.so", "_ref/libneo_ref_v4.so"], check=True, capture_output=True)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(_vp)


_SUF = {np.dtype(np.float32): "f32", np.dtype(np.float64): "f64"}
_CPLX = {np.dtype(np.float32): np.complex64, np.dtype(np.float64): np.complex128}
_REAL = {np.dtype(np.complex64): np.float32, np.dtype(np.complex128): np.float64}


class _Checker:
    """Common numpy-level API over either library (prefix 'oracle_' or 'ref_')."""

    def __init__(self, path: str, prefix: str):
        self.lib = C.CDLL(path)
        self.prefix = prefix
        self.path = path

    def _fn(self, name, restype=None, argtypes=None):
        f = getattr(self.lib, self.prefix + name)
        f.restype = restype
        if argtypes is not None:
            f.argtypes = argtypes
        return f

    # ---- integer tables -------------------------------------------------------------------------
    def bitrev_table(self, order: int) -> np.ndarray:
        out = np.zeros(1 << order, dtype=np.uint32)
        self._fn("bitrev_table", None, [_sz, _vp])(order, _ptr(out))
        return out

    def digitrev_perm(self, radix: int, size: int) -> np.ndarray:
        out = np.zeros(size, dtype=np.uint32)
        self._fn("digitrev_perm", None, [_sz, _sz, _vp])(radix, size, _ptr(out))
        return out

    def fdl_index_sequence(self, parts: int, calls: int):
        wp = np.zeros(calls, dtype=np.uint32)
        pairs = np.zeros((calls, parts, 2), dtype=np.uint32)
        self._fn("fdl_index_sequence", None, [_sz, _sz, _vp, _vp])(parts, calls, _ptr(wp), _ptr(pairs))
        return wp, pairs

    def num_stft_frames(self, signal: int, frame: int, overlap: int) -> int:
        return int(self._fn("num_stft_frames", _sz, [_sz, _sz, _sz])(signal, frame, overlap))

    def next_order(self, n: int) -> int:
        return int(self._fn("next_order", _sz, [_sz])(n))

    # ---- plans ------------------------------------------------------------------------------------
    def twiddle_lut(self, size: int, direction: int, dtype=np.float32) -> np.ndarray:
        dtype = np.dtype(dtype)
        out = np.zeros(size, dtype=dtype)  # size/2 complex
        self._fn("twiddle_lut_" + _SUF[dtype], None, [_sz, _i, _vp])(size, direction, _ptr(out))
        return out.view(_CPLX[dtype])

    def fft(self, x: np.ndarray, direction: int = -1) -> np.ndarray:
        """c2c of the last axis (any batch shape), unnormalised; returns a new array. Raises on order > 27."""
        x = np.ascontiguousarray(x)
        real = np.dtype(_REAL[x.dtype])
        n = x.shape[-1]
        order = int(n).bit_length() - 1
        assert 1 << order == n
        out = x.reshape(-1, n).copy()
        fn = self._fn("fft_c2c_" + _SUF[real], _i, [_sz, _vp, _i])
        for row in out:
            if fn(order, _ptr(row), direction) != 0:
                raise RuntimeError(f"unsupported order {order}")
        return out.reshape(x.shape)

    def dft(self, x: np.ndarray, direction: int = -1) -> np.ndarray:
        """dft_plan (Bluestein, fft/fallback/fallback_dft_plan.hpp:24-96) of the last axis, ANY size, unnormalised; new array."""
        x = np.ascontiguousarray(x)
        real = np.dtype(_REAL[x.dtype])
        n = x.shape[-1]
        out = x.reshape(-1, n).copy()
        fn = self._fn("dft_c2c_" + _SUF[real], _i, [_sz, _vp, _i])
        for row in out:
            if fn(n, _ptr(row), direction) != 0:
                raise RuntimeError(f"unsupported size {n}")
        return out.reshape(x.shape)

    def dct2(self, x: np.ndarray) -> np.ndarray:
        """fallback_dct2_plan (fft/dct.hpp:24-68): unnormalised type-2 DCT of the last axis (power-of-two length); new array."""
        x = np.ascontiguousarray(x)
        real = np.dtype(x.dtype)
        n = x.shape[-1]
        order = int(n).bit_length() - 1
        assert 1 << order == n
        out = x.reshape(-1, n).copy()
        fn = self._fn("dct2_" + _SUF[real], _i, [_sz, _vp])
        for row in out:
            if fn(order, _ptr(row)) != 0:
                raise RuntimeError(f"unsupported order {order}")
        return out.reshape(x.shape)

    def stft(self, x: np.ndarray, frame: int, transform: int, overlap: int, window: int = 1) -> np.ndarray:
        """stft_plan (fft/stft.hpp:39-109) of x[C][L]: window 0 rectangular / 1 hann / 2 hamming; returns [C][frames][bins]."""
        x = np.ascontiguousarray(x)
        real = np.dtype(x.dtype)
        ch, length = x.shape
        n = 1 << max(0, int(transform - 1).bit_length())
        frames = self.num_stft_frames(length, frame, overlap)
        out = np.zeros((ch, frames, n // 2 + 1), dtype=_CPLX[real])
        fn = self._fn("stft_" + _SUF[real], _sz, [_vp, _sz, _sz, _sz, _sz, _sz, _i, _vp])
        got = fn(_ptr(x), ch, length, frame, transform, overlap, window, _ptr(out))
        assert got == frames
        return out

    def fft_convolve(self, signal: np.ndarray, patch: np.ndarray) -> np.ndarray:
        """fft_convolve (convolution/fft_convolver.hpp:18-93), mode::full: signal[n] * patch[m] -> [n + m - 1]."""
        signal = np.ascontiguousarray(signal)
        patch = np.ascontiguousarray(patch, dtype=signal.dtype)
        out = np.zeros(signal.size + patch.size - 1, dtype=signal.dtype)
        self._fn("fft_convolve_" + _SUF[signal.dtype], None, [_vp, _sz, _vp, _sz, _vp])(_ptr(signal), signal.size, _ptr(patch), patch.size, _ptr(out))
        return out

    def fft_status(self, order: int) -> int:
        """0 if a c2c plan of this order can be built (runs a transform only for small orders)."""
        if order > 27:
            buf = np.zeros(4, dtype=np.float32)
            return int(self._fn("fft_c2c_f32", _i, [_sz, _vp, _i])(order, _ptr(buf), -1))
        return 0

    def rfft(self, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x)
        n = x.shape[-1]
        order = int(n).bit_length() - 1
        assert 1 << order == n
        rows = x.reshape(-1, n)
        out = np.zeros((rows.shape[0], n // 2 + 1), dtype=_CPLX[x.dtype])
        fn = self._fn("rfft_" + _SUF[x.dtype], None, [_sz, _vp, _vp])
        for r, o in zip(rows, out):
            fn(order, _ptr(r), _ptr(o))
        return out.reshape(x.shape[:-1] + (n // 2 + 1,))

    def irfft(self, x: np.ndarray, n: int) -> np.ndarray:
        """c2r, UNNORMALISED (fallback_rfft_plan.hpp:39-55). x[..., n/2+1] (or [..., n])."""
        x = np.ascontiguousarray(x)
        real = np.dtype(_REAL[x.dtype])
        order = int(n).bit_length() - 1
        rows = x.reshape(-1, x.shape[-1])
        out = np.zeros((rows.shape[0], n), dtype=real)
        fn = self._fn("irfft_" + _SUF[real], None, [_sz, _vp, _sz, _vp])
        for r, o in zip(rows, out):
            fn(order, _ptr(r), x.shape[-1], _ptr(o))
        return out.reshape(x.shape[:-1] + (n,))

    def multiply_add(self, x, y, z):
        x, y, z = (np.ascontiguousarray(a) for a in (x, y, z))
        out = np.zeros_like(x)
        name = "multiply_add_" + ("c64" if self.prefix == "ref_" else "f32")
        assert x.dtype == np.complex64
        self._fn(name, None, [_vp, _vp, _vp, _vp, _sz])(_ptr(x), _ptr(y), _ptr(z), _ptr(out), x.size)
        return out

    # ---- inputs -------------------------------------------------------------------------------------
    def noise(self, n: int, seed: int, dtype=np.float32) -> np.ndarray:
        """generate_noise_signal (testing/testing.hpp:37-72). Complex dtypes draw re then im."""
        dtype = np.dtype(dtype)
        if dtype.kind == "c":
            return self.noise(2 * n, seed, _REAL[dtype]).view(dtype)
        out = np.zeros(n, dtype=dtype)
        self._fn("noise_" + _SUF[dtype], None, [_sz, _u32, _vp])(n, seed, _ptr(out))
        return out

    def normalize_impulse(self, ir: np.ndarray) -> np.ndarray:
        ir = np.ascontiguousarray(ir, dtype=np.float32).copy()
        ch, ln = ir.shape
        self._fn("normalize_impulse_f32", None, [_vp, _sz, _sz])(_ptr(ir), ch, ln)
        return ir

    # ---- filter preparation + convolvers ------------------------------------------------------------
    def uniform_partition(self, ir: np.ndarray, block: int) -> np.ndarray:
        """ir[C][L] -> H[C][P][B+1] (convolution/uniform_partition.hpp:13-26)."""
        ir = np.ascontiguousarray(ir)
        ch, ln = ir.shape
        fn = self._fn("uniform_partition_" + _SUF[ir.dtype], _sz, [_vp, _sz, _sz, _sz, _vp])
        parts = int(fn(_ptr(ir), ch, ln, block, None))
        out = np.zeros((ch, parts, block + 1), dtype=_CPLX[ir.dtype])
        fn(_ptr(ir), ch, ln, block, _ptr(out))
        return out

    def convolver(self, kind: int, dtype=np.float32) -> "_Convolver":
        return _Convolver(self, kind, np.dtype(dtype))

    def compressed_fdl_roundtrip(self, row: np.ndarray, bits: int) -> np.ndarray:
        """compressed_fdl<complex<float>, scalar_complex<intN>>: insert(row, 0) then operator[](0) (compressed_fdl.hpp:26-48)"""
        row = np.ascontiguousarray(row, dtype=np.complex64)
        out = np.zeros_like(row)
        if self.prefix == "ref_":
            self._fn("compressed_fdl_roundtrip_f32", None, [_vp, _sz, _i, _vp])(_ptr(row), row.size, bits, _ptr(out))
        else:
            q = self.compress_row(row, bits)
            self._fn("decompress_row_f32", None, [_vp, _sz, _i, _vp])(_ptr(q), row.size, bits, _ptr(out))
        return out

    def compress_row(self, row: np.ndarray, bits: int) -> np.ndarray:
        """the stored integers of compressed_fdl::insert: int16 [n][2] (int8 values sign-extended). Oracle only."""
        row = np.ascontiguousarray(row)
        q = np.zeros((row.size, 2), dtype=np.int16)
        self._fn("compress_row_" + _SUF[np.dtype(_REAL[row.dtype])], None, [_vp, _sz, _i, _vp])(_ptr(row), row.size, bits, _ptr(q))
        return q

    def csr_build(self, H: np.ndarray, threshold: float):
        """neo::csr_matrix(H, predicate) (container/csr_matrix.hpp:64-98), predicate |re| > threshold or |im| > threshold:
        (row_ptr [P+1] uint64, cols [nnz] uint64, values [nnz] complex). float32 only."""
        H = np.ascontiguousarray(H, dtype=np.complex64)
        fn = self._fn("csr_build_f32", _sz, [_vp, _sz, _sz, C.c_float, _vp, _vp, _vp])
        rows = np.zeros(H.shape[0] + 1, dtype=np.uint64)
        nnz = int(fn(_ptr(H), H.shape[0], H.shape[1], threshold, _ptr(rows), None, None))
        cols, vals = np.zeros(max(nnz, 1), dtype=np.uint64), np.zeros(max(nnz, 1), dtype=np.complex64)
        fn(_ptr(H), H.shape[0], H.shape[1], threshold, _ptr(rows), _ptr(cols), _ptr(vals))
        return rows, cols[:nnz], vals[:nnz]

    def convolve_blocks_sparse(self, kind: int, H: np.ndarray, signal: np.ndarray, threshold: float) -> np.ndarray:
        """convolve_blocks with sparse_upols (kind 5) / sparse_upola (kind 6) convolvers"""
        real = np.dtype(_REAL[H.dtype])
        out = np.ascontiguousarray(signal, dtype=real).copy()
        block = H.shape[-1] - 1
        for c in range(out.shape[0]):
            conv = self.convolver(kind, real)
            conv.filter_sparse(H[c], threshold)
            for s in range(0, out.shape[1], block):
                conv.process(out[c, s : s + block])
            conv.close()
        return out

    def convolve_blocks(self, kind: int, H: np.ndarray, signal: np.ndarray, chunk: int | None = None) -> np.ndarray:
        """Run one convolver per channel over signal[C][n] (n multiple of B) and return the output.
        H: [C][P][K] (own filter per channel)."""
        real = np.dtype(_REAL[H.dtype])
        signal = np.ascontiguousarray(signal, dtype=real)
        out = signal.copy()
        block = H.shape[-1] - 1
        step = chunk or block
        for c in range(signal.shape[0]):
            conv = self.convolver(kind, real)
            conv.filter(H[c])
            for s in range(0, signal.shape[1], step):
                conv.process(out[c, s : s + step])
            conv.close()
        return out


class _Convolver:
    def __init__(self, chk: _Checker, kind: int, dtype: np.dtype):
        self.chk, self.suf = chk, _SUF[dtype]
        self.dtype = dtype
        self.h = chk._fn("conv_create_" + self.suf, _vp, [_i])(kind)

    def filter(self, H: np.ndarray) -> None:
        H = np.ascontiguousarray(H, dtype=_CPLX[self.dtype])
        self.chk._fn("conv_filter_" + self.suf, None, [_vp, _vp, _sz, _sz])(self.h, _ptr(H), H.shape[0], H.shape[1])

    def filter_sparse(self, H: np.ndarray, threshold: float) -> None:
        """sparse_filter::filter(H, sparsity) with the predicate |re| > threshold or |im| > threshold (kinds 5, 6)"""
        H = np.ascontiguousarray(H, dtype=_CPLX[self.dtype])
        thr = C.c_float if self.dtype == np.float32 else C.c_double
        self.chk._fn("conv_filter_sparse_" + self.suf, None, [_vp, _vp, _sz, _sz, thr])(self.h, _ptr(H), H.shape[0], H.shape[1], threshold)

    def process(self, block: np.ndarray) -> None:
        assert block.dtype == self.dtype and block.flags.c_contiguous
        self.chk._fn("conv_process_" + self.suf, None, [_vp, _vp, _sz])(self.h, _ptr(block), block.size)

    def close(self) -> None:
        if self.h:
            self.chk._fn("conv_destroy_" + self.suf, None, [_vp])(self.h)
            self.h = None


class _Oracle(_Checker):
    def direct_convolve(self, signal: np.ndarray, ir: np.ndarray, out_len: int) -> np.ndarray:
        signal, ir = np.ascontiguousarray(signal), np.ascontiguousarray(ir, dtype=signal.dtype)
        out = np.zeros(out_len, dtype=signal.dtype)
        self._fn("direct_convolve_" + _SUF[signal.dtype], None, [_vp, _sz, _vp, _sz, _vp, _sz])(
            _ptr(signal), signal.size, _ptr(ir), ir.size, _ptr(out), out_len
        )
        return out

    def digitrev_lut(self, radix: int, size: int) -> np.ndarray:
        out = np.zeros(size, dtype=np.uint32)
        self._fn("digitrev_lut", None, [_sz, _sz, _vp])(radix, size, _ptr(out))
        return out


class _Ref(_Checker):
    def fft_max_order(self) -> int:
        return int(self._fn("fft_max_order", _sz, [])())

    def split_fft(self, re: np.ndarray, im: np.ndarray, direction: int = -1):
        """split_fft_plan in place on copies of the two planes (1-D)."""
        re, im = np.ascontiguousarray(re).copy(), np.ascontiguousarray(im).copy()
        order = int(re.size).bit_length() - 1
        self._fn("split_fft_" + _SUF[re.dtype], None, [_sz, _vp, _vp, _i])(order, _ptr(re), _ptr(im), direction)
        return re, im

    def overlap_identity(self, add: bool, block: int, filter_size: int, signal: np.ndarray) -> np.ndarray:
        sig = np.ascontiguousarray(signal, dtype=np.float32).copy()
        self._fn("overlap_identity_f32", None, [_i, _sz, _sz, _vp, _sz])(
            int(add), block, filter_size, _ptr(sig), sig.size // block
        )
        return sig

    def conv_bench(self, kind, filters, filter_stride, signal, threads) -> float:
        """signal[C][nblocks*B] processed in place; returns seconds of the block loop."""
        parts, bins = filters.shape[-2], filters.shape[-1]
        ch = signal.shape[0]
        nblocks = signal.shape[1] // (bins - 1)
        fn = self._fn("conv_bench_f32", C.c_double, [_i, _vp, _sz, _sz, _sz, _vp, _sz, _sz, _sz])
        return float(fn(kind, _ptr(filters), filter_stride, parts, bins, _ptr(signal), ch, nblocks, threads))

    def rfft_bench(self, order: int, data: np.ndarray, threads: int) -> float:
        fn = self._fn("rfft_bench_f32", C.c_double, [_sz, _vp, _sz, _sz])
        return float(fn(order, _ptr(data), data.shape[0], threads))

    def c2c_bench(self, order: int, data: np.ndarray, reps: int) -> float:
        """`reps` forward + inverse + 1/N round trips of one complex64 plan in place on one thread; returns seconds."""
        fn = self._fn("c2c_bench_f32", C.c_double, [_sz, _vp, _sz])
        return float(fn(order, _ptr(data), reps))


@functools.lru_cache(maxsize=None)
def oracle() -> _Oracle:
    if not os.path.exists(ORACLE_SO):
        build()
    return _Oracle(ORACLE_SO, "oracle_")


def _host_has_avx512() -> bool:
    try:
        with open("/proc/cpuinfo") as f:
            flags = next((line for line in f if line.startswith("flags")), "")
    except OSError:
        return False
    return all(f" {name}" in flags for name in ("avx512f", "avx512bw", "avx512cd", "avx512dq", "avx512vl"))


@functools.lru_cache(maxsize=None)
def ref() -> _Ref | None:
    if not os.path.exists(REF_SO):
        return None
    return _Ref(REF_SO, "ref_")


@functools.lru_cache(maxsize=None)
def ref_timing() -> tuple[_Ref, str] | None:
    """The reference build bench.py TIMES: the widest instruction set this host runs (the library cannot be rebuilt -march=native on
    the GPU box, so two builds travel). Returns (library, compiler flags it was built with)."""
    if os.path.exists(REF_V4_SO) and _host_has_avx512():
        return _Ref(REF_V4_SO, "ref_"), "g++ -O3 -march=x86-64-v4"
    r = ref()
    return (r, "g++ -O3 -march=x86-64-v3") if r is not None else None
