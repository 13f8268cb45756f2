"""Index-math model of csrc/fft_wide.cuh (CPU, numpy): replays the kernels thread by thread -- who loads what, where a value
sits in the swizzled tile, which Hermitian partners meet in one thread -- and checks the result against numpy's FFT and the bank
pattern of every shared-memory access. Development aid; not imported by the product or the tests.

    python tools/wide_fft_model.py            # all shipped geometries, forward + backward, split parity 0/1
"""
import sys

import numpy as np


def elem(k1, x, a, R2):
    """tile position (complex elements) of (k1, x, a): row = k1*R2 + x holds 16 values, 16-byte chunks XOR-swizzled"""
    return (k1 * R2 + x) * 16 + ((((a >> 1) ^ (k1 & 7)) << 1) | (a & 1))


def conflicts(addr_bytes, width):
    """wavefronts of one warp-wide shared-memory access (32 lanes, `width` bytes each) beyond the minimum"""
    lanes_per_wave = 128 // width
    extra = 0
    for g in range(0, 32, lanes_per_wave):
        lanes = addr_bytes[g : g + lanes_per_wave]
        banks = {}
        for a in lanes:
            for w in range(width // 4):
                banks.setdefault(((a // 4) + w) % 32, set()).add((a // 4) + w)
        extra += max(len(v) for v in banks.values()) - 1
    return extra


class Tile:
    def __init__(self, m):
        self.d = np.zeros(m, np.complex128)
        self.extra = 0
        self.log = {}

    def access(self, key, t, pos, width):
        self.log.setdefault(key, {})[t] = pos * 8

    def flush(self, width_of):
        for key, lanes in self.log.items():
            ts = sorted(lanes)
            for w0 in range(0, len(ts), 32):
                warp = [lanes[t] for t in ts[w0 : w0 + 32]]
                if len(warp) == 32:
                    self.extra += conflicts(warp, width_of(key))
        self.log = {}


def forward(z, logm, logr1, logr2, shift=0):
    """Z = FFT_M(z), returned as the per-thread stage-3 registers: dict u -> (jA, jB, ZA[16], ZB[16]).
    shift: Hermitian pairing of the split kernels (rows j and (J - shift - j) mod J meet in one thread)."""
    M, R1, R2 = 1 << logm, 1 << logr1, 1 << logr2
    assert logr1 + logr2 + 4 == logm
    NT, S1, J = M // 32, M // R1, M // 16
    BF1, BF2 = 32 // R1, 32 // R2
    tile = Tile(M)
    # stage 1
    for t in range(NT):
        if BF1 >= 2:
            for m in range(BF1 // 2):
                n0 = 2 * t + 2 * NT * m
                for h in range(2):
                    n = n0 + h
                    u = np.array([z[n + S1 * n1] for n1 in range(R1)])
                    U = np.fft.fft(u) * np.exp(-2j * np.pi * n * np.arange(R1) / M)
                    a, a2 = n & 15, n >> 4
                    for k1 in range(R1):
                        tile.d[elem(k1, a2, a, R2)] = U[k1]
                        if h == 0:
                            tile.access(("s1", m, k1), t, elem(k1, a2, a, R2), 16)
        else:
            n = t
            u = np.array([z[n + S1 * n1] for n1 in range(R1)])
            U = np.fft.fft(u) * np.exp(-2j * np.pi * n * np.arange(R1) / M)
            a, a2 = n & 15, n >> 4
            for k1 in range(R1):
                tile.d[elem(k1, a2, a, R2)] = U[k1]
                tile.access(("s1", 0, k1), t, elem(k1, a2, a, R2), 8)
    tile.flush(lambda key: 16 if BF1 >= 2 else 8)
    # stage 2, in place
    for t in range(NT):
        for m in range(BF2):
            beta = t + NT * m
            a, k1 = beta & 15, beta >> 4
            u = np.array([tile.d[elem(k1, a2, a, R2)] for a2 in range(R2)])
            for a2 in range(R2):
                tile.access(("s2", m, a2), t, elem(k1, a2, a, R2), 8)
            U = np.fft.fft(u) * np.exp(-2j * np.pi * a * np.arange(R2) / (R2 * 16))
            for k2 in range(R2):
                tile.d[elem(k1, k2, a, R2)] = U[k2]
    tile.flush(lambda key: 8)
    # stage 3
    regs = {}
    for u_ in range(NT):
        if shift == 0:
            jA, jB = (0, J // 2) if u_ == 0 else (u_, J - u_)
        else:
            jA, jB = u_, J - 1 - u_
        out = []
        for which, j in (("A", jA), ("B", jB)):
            k1, k2 = j & (R1 - 1), j >> logr1
            v = np.zeros(16, np.complex128)
            for c in range(8):
                pc = c ^ (k1 & 7)
                pos = (k1 * R2 + k2) * 16 + 2 * pc
                v[2 * c], v[2 * c + 1] = tile.d[pos], tile.d[pos + 1]
                tile.access(("s3", which, c), u_, pos, 16)
            out.append(np.fft.fft(v))
        regs[u_] = (jA, jB, out[0], out[1])
    tile.flush(lambda key: 16)
    return regs, tile.extra


def rfft_wide(x, logm, logr1, logr2):
    M = 1 << logm
    J = M // 16
    z = x[0::2] + 1j * x[1::2]
    regs, extra = forward(z, logm, logr1, logr2)
    X = np.zeros(M + 1, np.complex128)

    def pair(zk, zmk, w):
        zc = np.conj(zmk)
        s, d = zk + zc, zk - zc
        return 0.5 * (s - 1j * w * d), 0.5 * np.conj(s + 1j * w * d)

    for u_, (jA, jB, ZA, ZB) in regs.items():
        if u_ == 0:
            X[0] = ZA[0].real + ZA[0].imag
            X[M] = ZA[0].real - ZA[0].imag
            for k3 in range(1, 8):
                X[J * k3], X[M - J * k3] = pair(ZA[k3], ZA[16 - k3], np.exp(-2j * np.pi * k3 / 32))
            X[M // 2] = np.conj(ZA[8])
            for k3 in range(8):
                k = J // 2 + J * k3
                X[k], X[M - k] = pair(ZB[k3], ZB[15 - k3], np.exp(-2j * np.pi * (1 + 2 * k3) / 64))
        else:
            wu = np.exp(-1j * np.pi * u_ / M)
            for k3 in range(16):
                k = u_ + J * k3
                X[k], X[M - k] = pair(ZA[k3], ZB[15 - k3], wu * np.exp(-2j * np.pi * k3 / 32))
    return X, extra


def irfft_wide(X, logm, logr1, logr2):
    """unnormalised backward transform (2M * numpy's irfft), the transposed flow"""
    M, R1, R2 = 1 << logm, 1 << logr1, 1 << logr2
    NT, S1, J = M // 32, M // R1, M // 16
    BF1, BF2 = 32 // R1, 32 // R2
    tile = Tile(M)

    def pre(xk, xmk, w):
        xc = np.conj(xmk)
        s, d = xk + xc, xk - xc
        wd = d * np.conj(w)
        return s + 1j * wd, np.conj(s - 1j * wd)

    for u_ in range(NT):
        jA, jB = (0, J // 2) if u_ == 0 else (u_, J - u_)
        ZA, ZB = np.zeros(16, np.complex128), np.zeros(16, np.complex128)
        if u_ == 0:
            ZA[0] = (X[0].real + X[M].real) + 1j * (X[0].real - X[M].real)
            ZA[8] = 2 * np.conj(X[M // 2])
            for k3 in range(1, 8):
                ZA[k3], ZA[16 - k3] = pre(X[J * k3], X[M - J * k3], np.exp(-2j * np.pi * k3 / 32))
            for k3 in range(8):
                k = J // 2 + J * k3
                ZB[k3], ZB[15 - k3] = pre(X[k], X[M - k], np.exp(-2j * np.pi * (1 + 2 * k3) / 64))
        else:
            wu = np.exp(-1j * np.pi * u_ / M)
            for k3 in range(16):
                k = u_ + J * k3
                ZA[k3], ZB[15 - k3] = pre(X[k], X[M - k], wu * np.exp(-2j * np.pi * k3 / 32))
        for which, j, Z in (("A", jA, ZA), ("B", jB, ZB)):
            Bv = np.fft.ifft(Z) * 16 * np.exp(2j * np.pi * np.arange(16) * j / M)
            k1, k2 = j & (R1 - 1), j >> logr1
            for c in range(8):
                pc = c ^ (k1 & 7)
                pos = (k1 * R2 + k2) * 16 + 2 * pc
                tile.d[pos], tile.d[pos + 1] = Bv[2 * c], Bv[2 * c + 1]
                tile.access(("s3", which, c), u_, pos, 16)
    tile.flush(lambda key: 16)
    for t in range(NT):
        for m in range(BF2):
            beta = t + NT * m
            a, k1 = beta & 15, beta >> 4
            u = np.array([tile.d[elem(k1, k2, a, R2)] for k2 in range(R2)])
            U = np.fft.ifft(u) * R2 * np.exp(2j * np.pi * np.arange(R2) * k1 / (R1 * R2))
            for a2 in range(R2):
                tile.d[elem(k1, a2, a, R2)] = U[a2]
                tile.access(("s2", m, a2), t, elem(k1, a2, a, R2), 8)
    tile.flush(lambda key: 8)
    z = np.zeros(M, np.complex128)
    for t in range(NT):
        ns = [2 * t + 2 * NT * m + h for m in range(BF1 // 2) for h in range(2)] if BF1 >= 2 else [t]
        for n in ns:
            a, a2 = n & 15, n >> 4
            u = np.array([tile.d[elem(k1, a2, a, R2)] for k1 in range(R1)])
            U = np.fft.ifft(u) * R1
            for b2 in range(R1):
                z[n + S1 * b2] = U[b2]
    x = np.zeros(2 * M)
    x[0::2], x[1::2] = z.real, z.imag
    return x, tile.extra


def main():
    rng = np.random.default_rng(1)
    geoms = [(13, 4, 5), (14, 5, 5), (12, 4, 4), (11, 4, 3), (10, 3, 3)]
    for logm, l1, l2 in geoms:
        M = 1 << logm
        x = rng.standard_normal(2 * M)
        X, extra_f = rfft_wide(x, logm, l1, l2)
        want = np.fft.rfft(x)
        ef = np.linalg.norm(X - want) / np.linalg.norm(want)
        y, extra_b = irfft_wide(want, logm, l1, l2)
        eb = np.linalg.norm(y - 2 * M * x) / np.linalg.norm(2 * M * x)
        print(f"M=2^{logm} R=({1 << l1},{1 << l2},16): rfft {ef:.1e} irfft {eb:.1e} extra wavefronts fwd {extra_f} bwd {extra_b}")
        assert ef < 1e-12 and eb < 1e-12
    # the shifted pairing of the odd-parity split CTA: rows j and J-1-j meet in one thread
    logm, l1, l2 = 13, 4, 5
    M = 1 << logm
    J = M // 16
    z = rng.standard_normal(M) + 1j * rng.standard_normal(M)
    regs, _ = forward(z, logm, l1, l2, shift=1)
    Z = np.fft.fft(z)
    for u_, (jA, jB, ZA, ZB) in regs.items():
        for k3 in range(16):
            assert abs(ZA[k3] - Z[jA + J * k3]) < 1e-9 and abs(ZB[k3] - Z[jB + J * k3]) < 1e-9
            assert (jA + J * k3) + (jB + J * (15 - k3)) == M - 1
    print("shifted pairing ok")


if __name__ == "__main__":
    sys.exit(main())
