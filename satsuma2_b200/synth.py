"""Synthetic workloads of the shapes BASELINE.json names (no network, no datasets).

config 2: independent target/query chunk pairs of uniform random A/C/G/T with one planted
homologous segment per pair (length U[60,500], identity U[0.70,0.95], substitutions only, random
offsets, half of them on the reverse strand).  Generation is blocked and seeded per block so that
any rank can produce its own shard deterministically.
"""
from __future__ import annotations

import numpy as np

_ASCII = np.frombuffer(b"ACGT", dtype=np.uint8)


def random_pairs(n: int, chunk: int = 4096, seed: int = 1, block: int = 16384, plant: bool = True,
                 out_t: np.ndarray | None = None, out_q: np.ndarray | None = None):
    """-> (targets[n, chunk] uint8 ASCII, queries[n, chunk] uint8 ASCII, truth[n] structured)."""
    T = out_t if out_t is not None else np.empty((n, chunk), dtype=np.uint8)
    Q = out_q if out_q is not None else np.empty((n, chunk), dtype=np.uint8)
    truth = np.zeros(n, dtype=[("tpos", "<i4"), ("qpos", "<i4"), ("len", "<i4"), ("ident", "<f4"), ("reverse", "u1")])
    for b0 in range(0, n, block):
        b1 = min(n, b0 + block)
        m = b1 - b0
        rng = np.random.default_rng([seed, b0 // block])
        t = rng.integers(0, 4, size=(m, chunk), dtype=np.uint8)
        q = rng.integers(0, 4, size=(m, chunk), dtype=np.uint8)
        if plant:
            hi = min(500, chunk - 1)
            lo = min(60, hi)
            seglen = rng.integers(lo, hi + 1, size=m)
            ident = rng.uniform(0.70, 0.95, size=m)
            tpos = (rng.random(m) * (chunk - seglen + 1)).astype(np.int64)
            qpos = (rng.random(m) * (chunk - seglen + 1)).astype(np.int64)
            rev = rng.random(m) < 0.5
            # flat index arrays over all planted bases of the block
            pair = np.repeat(np.arange(m), seglen)
            k = np.arange(seglen.sum()) - np.repeat(np.cumsum(seglen) - seglen, seglen)
            src = t[pair, tpos[pair] + k]
            mut = rng.random(src.shape[0]) >= ident[pair]
            sub = rng.integers(1, 4, size=src.shape[0], dtype=np.uint8)
            val = np.where(mut, (src + sub) & 3, src).astype(np.uint8)
            r = rev[pair]
            qidx = np.where(r, qpos[pair] + (seglen[pair] - 1 - k), qpos[pair] + k)
            q[pair, qidx] = np.where(r, 3 - val, val)
            truth["tpos"][b0:b1] = tpos
            truth["qpos"][b0:b1] = qpos
            truth["len"][b0:b1] = seglen
            truth["ident"][b0:b1] = ident
            truth["reverse"][b0:b1] = rev
        T[b0:b1] = _ASCII[t]
        Q[b0:b1] = _ASCII[q]
    return T, Q, truth


def repeat_rich_pair(length: int, seed: int = 7, families: int = 6, divergence: float = 0.10):
    """Two related sequences full of tandem repeats and interspersed repeat families (config 5)."""
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 4, size=length, dtype=np.uint8)
    fams = [rng.integers(0, 4, size=int(rng.integers(150, 600)), dtype=np.uint8) for _ in range(families)]
    pos = 0
    while pos < length - 700:
        kind = rng.random()
        if kind < 0.35:  # interspersed family copy, diverged
            f = fams[int(rng.integers(0, families))].copy()
            mut = rng.random(f.shape[0]) < divergence
            f[mut] = (f[mut] + rng.integers(1, 4, size=int(mut.sum()), dtype=np.uint8)) & 3
            base[pos:pos + f.shape[0]] = f
            pos += f.shape[0]
        elif kind < 0.55:  # tandem repeat
            unit = rng.integers(0, 4, size=int(rng.integers(2, 40)), dtype=np.uint8)
            reps = int(rng.integers(5, 30))
            tr = np.tile(unit, reps)[: length - pos]
            base[pos:pos + tr.shape[0]] = tr
            pos += tr.shape[0]
        elif kind < 0.62:  # low-complexity tract
            n = int(rng.integers(40, 200))
            base[pos:pos + n] = rng.integers(0, 4)
            pos += n
        pos += int(rng.integers(50, 400))
    other = base.copy()
    mut = rng.random(length) < 0.08
    other[mut] = (other[mut] + rng.integers(1, 4, size=int(mut.sum()), dtype=np.uint8)) & 3
    return _ASCII[base], _ASCII[other]


def chunk_sequence(seq: np.ndarray, size: int, overlap: int, seq_id: int = 0):
    """ChunkManager::ChunkItSelect arithmetic (analysis/SeqChunk.cc:72-165) for ONE sequence:
    1 + l/(size-overlap) chunks, chunk j = [j*stride, min((j+1)*stride + overlap, l)); all-N/X chunks
    are emptied but keep their index; sequences shorter than 6 bases produce no chunks.
    -> (offsets, lens, starts) relative to seq."""
    l = int(seq.shape[0])
    if l < 6:
        return np.zeros(0, np.int64), np.zeros(0, np.int32), np.zeros(0, np.int32)
    stride = size - overlap
    n = 1 + l // stride
    starts = np.arange(n, dtype=np.int64) * stride
    ends = np.minimum(starts + stride + overlap, l)
    lens = np.maximum(ends - starts, 0).astype(np.int32)
    starts = np.minimum(starts, l)
    for j in range(n):
        c = seq[starts[j]:starts[j] + lens[j]]
        if lens[j] and np.all((c == ord("N")) | (c == ord("X"))):
            lens[j] = 0
    return starts.astype(np.int64), lens, starts.astype(np.int32)


def genome_pair(length: int, seed: int = 11, divergence: float = 0.12, inversions: int = 4):
    """config 4: a random target genome and a query that is a diverged copy of it with a few inverted
    (reverse-complemented) stretches -- syntenic along the main diagonal.  -> (target, query) uint8 ASCII."""
    rng = np.random.default_rng(seed)
    t = rng.integers(0, 4, size=length, dtype=np.uint8)
    q = t.copy()
    mut = rng.random(length) < divergence
    q[mut] = (q[mut] + rng.integers(1, 4, size=int(mut.sum()), dtype=np.uint8)) & 3
    for _ in range(inversions):
        n = int(rng.integers(length // 50, length // 20))
        a = int(rng.integers(0, length - n))
        q[a:a + n] = 3 - q[a:a + n][::-1]
    return _ASCII[t], _ASCII[q]


def diagonal_blocks(n_target_chunks: int, n_query_chunks: int, t_stride: int, q_stride: int, pixel: int = 24):
    """t_pair blocks of pixel x pixel chunks along the syntenic diagonal, the way GridSearch issues them
    (analysis/GridSearch.cc): query chunks [qb, qb + pixel) against the target chunks that cover the same bases."""
    blocks = []
    for qb in range(0, n_query_chunks, pixel):
        q_hi = min(n_query_chunks - 1, qb + pixel - 1)
        t_lo = max(0, (qb * q_stride) // t_stride - 1)
        t_hi = min(n_target_chunks - 1, t_lo + pixel - 1)
        if t_lo <= t_hi:
            blocks.append((t_lo, t_hi, qb, q_hi, 0))
    return blocks
