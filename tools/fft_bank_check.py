#!/usr/bin/env python
"""Brute-force check that every radix pass of every plan in sx_fft.cuh is shared-memory bank-conflict
free under the XOR swizzle (8-byte elements: a half-warp of 16 lanes must hit 16 distinct slots mod 16)."""
PLANS = {11: (16, 8, 8), 12: (16, 16, 8), 13: (16, 16, 16), 14: (16, 8, 8, 8), 15: (16, 16, 8, 8)}


def swz_f(l):
    return ((l >> 4) & 7) ^ ((((l >> 6) ^ (l >> 7)) & 1) << 3)


def swz(l):
    return l ^ swz_f(l)


def check(log2n, radices, nt):
    n = 1 << log2n
    h = n // 2
    L = h
    worst = 1
    for R in radices:
        S = L // R
        for b0 in range(0, n // R, 16):  # one half-warp of butterflies
            for m in range(R):
                slots = set()
                for b in range(b0, b0 + 16):
                    j, blk = b % S, b // S
                    slots.add(swz(blk * L + j + m * S) % 16)
                worst = max(worst, 16 // len(slots) if len(slots) else 1)
                assert len(slots) == 16, (log2n, R, L, b0, m, sorted(slots))
        L = S
    return worst


if __name__ == "__main__":
    for l2, rad in PLANS.items():
        print(l2, rad, "conflict-free" if check(l2, rad, 256) == 1 else "CONFLICTS")
