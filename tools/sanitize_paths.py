#!/usr/bin/env python
"""Small driver for compute-sanitizer (tools/sanitize.sh): every kernel route added in the second session of round 2,
on inputs small enough for racecheck -- three-channel form (pair mode and grid with cached targets), the fused
transform + correlation kernel with a pair it hands back (IUPAC letter -> list-mode fall-back kernels), explicit
reverse-strand signals through the preparation kernel, the N = 32768 cluster kernel."""
import sys

import numpy as np

sys.path.insert(0, "/root/repo")
import satsuma2_b200 as sx  # noqa: E402
from satsuma2_b200 import synth  # noqa: E402

n = 48
T, Q, _ = synth.random_pairs(n, 4096, seed=11)
tl = [(T[i, : (4096 if i % 3 else 3000 + 7 * i)].tobytes(), 0, i, 4096) for i in range(n)]
ql = [(Q[i, : (4096 if i % 4 else 2500 + 13 * i)].tobytes(), 0, i, 4096) for i in range(n)]
t5 = bytearray(tl[5][0]); t5[100] = ord("R"); tl[5] = (bytes(t5), 0, 5, 4096)
q9 = bytearray(ql[9][0]); q9[200] = ord("N"); ql[9] = (bytes(q9), 0, 9, 4096)
pairs = [(i, i) for i in range(n)]
grid = [(t, q) for t in range(8) for q in range(8, 16)]
tot = 0
for kw in (dict(), dict(fuse_pairs=1), dict(fuse_pairs=1, spectra_cache_bytes=-1), dict(debug_flags=4)):
    with sx.XCorrEngine(target_total=50000.0, max_batch_pairs=20, **kw) as eng:
        eng.set_targets(sx.ChunkSet.from_list(tl))
        eng.set_queries(sx.ChunkSet.from_list(ql))
        tot += len(eng.align_pairs(pairs)) + len(eng.align_pairs(grid))
n2, chunk = 3, 16384
T2, Q2, _ = synth.random_pairs(n2, chunk, seed=12)
for flags in (0, 8):
    with sx.XCorrEngine(t_chunk=chunk, q_chunk=chunk, target_total=float(n2 * chunk), debug_flags=flags) as eng:
        eng.set_targets(sx.ChunkSet.independent(T2))
        eng.set_queries(sx.ChunkSet.independent(Q2))
        tot += len(eng.align_pairs([(i, i) for i in range(n2)]))
print("records", tot)
