#!/usr/bin/env python
"""Shared-memory bank-conflict check of cta_fft's stage exchanges (neo-dsp_b200/csrc/fft_core.cuh), mirroring its index
math: writes sm[pad(base + q*Ns)], reads sm[pad(t + e*TN)], pad(i) = i + (i >> shift). 8-byte elements are served per
half-warp (16 lanes x 8 B = 128 B per wavefront), 16-byte elements per quarter-warp."""
import sys


def pick_loge(logm, f32):
    if logm <= 2:
        return logm
    return (4 if logm >= 8 else 3) if f32 else 3


def wavefronts(addrs, elem_bytes):
    """addrs: element indices of the 32 lanes; returns wavefronts needed (ideal = 32*elem_bytes/128)"""
    lanes_per = 128 // elem_bytes
    banks_per = 32 * 4 // elem_bytes
    total = 0
    for h in range(0, 32, lanes_per):
        group = addrs[h:h + lanes_per]
        per_bank = {}
        for a in set(group):
            per_bank.setdefault(a % banks_per, set()).add(a)
        total += max(len(v) for v in per_bank.values())
    return total


def check(logm, f32=True):
    loge = pick_loge(logm, f32)
    if loge == 0:
        return []
    m, e = 1 << logm, 1 << loge
    tn = m // e
    shift = 4 if f32 else 3
    eb = 8 if f32 else 16
    pad = lambda i: i + (i >> shift)
    tile = pad(m) + 1
    g = 1 if tn >= 64 else 64 // tn
    threads = tn * g
    r0 = logm % loge
    stages = ([(0, r0)] if r0 else []) + [(ns, loge) for ns in range(r0, logm, loge)]
    out = []
    ideal = 32 * eb // 128
    for logns, logr in stages:
        if logns + logr == logm:
            continue
        ns, r = 1 << logns, 1 << logr
        bf = e // r
        worst_w = worst_r = 0
        for warp in range(0, threads, 32):
            lanes = range(warp, min(warp + 32, threads))
            for mm in range(bf):
                for q in range(r):
                    addrs = []
                    for tid in lanes:
                        grp, t = divmod(tid, tn)
                        j = t + mm * tn
                        k = j & (ns - 1)
                        base = ((j >> logns) << (logns + logr)) + k
                        addrs.append(grp * tile + pad(base + (q << logns)))
                    addrs += [addrs[-1]] * (32 - len(addrs))
                    worst_w = max(worst_w, wavefronts(addrs, eb))
            for ee in range(e):
                addrs = []
                for tid in lanes:
                    grp, t = divmod(tid, tn)
                    addrs.append(grp * tile + pad(t + ee * tn))
                addrs += [addrs[-1]] * (32 - len(addrs))
                worst_r = max(worst_r, wavefronts(addrs, eb))
        out.append((logns, logr, worst_w, worst_r, ideal))
    return out


if __name__ == "__main__":
    for f32 in (True, False):
        print("float32" if f32 else "float64")
        for logm in range(3, 14 if f32 else 13):
            res = check(logm, f32)
            bad = [r for r in res if r[2] > r[4] or r[3] > r[4]]
            print(f"  M=2^{logm:<2d} stages(logNs,logR,write wf,read wf,ideal)={res} {'CONFLICTS' if bad else 'ok'}")
