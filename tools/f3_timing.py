#!/usr/bin/env python
"""the master-side list operations (MultiMatches::Sort / Collapse, RunMatchDynProg) on >= 10^6
matches -- this repository's host code (XCorrMatchTool) timed beside the compiled reference's own objects
(oracle/_ref/libsatsuma_ref.so, ref_sort_collapse_chain_timed), one CPU thread each, same records, same output.
CPU only.   python tools/f3_timing.py [n_matches]  -> one JSON line (kept under profiles/)."""
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
from satsuma2_b200 import build as sxbuild  # noqa: E402
from test_host_tools import _write_match_file  # noqa: E402


def synth_matches(n, n_seq, size, seed=3):
    """Synteny along the diagonals of n_seq sequence pairs with indel drift, duplicates from overlapping target chunks
    (SURVEY Q14), off-diagonal noise and a repeat pile-up -- what the slaves deliver cycle after cycle."""
    rng = np.random.default_rng(seed)
    k = int(n * 0.55)
    recs = np.zeros((k, 10))
    recs[:, 0] = rng.integers(0, n_seq, k)
    recs[:, 1] = recs[:, 0]
    recs[:, 2] = size
    recs[:, 3] = rng.integers(1, size - 2000, k)
    recs[:, 4] = np.clip(recs[:, 3] + rng.integers(-300, 300, k), 1, size - 1000)
    recs[:, 5] = rng.integers(46, 600, k)
    recs[:, 9] = rng.uniform(0.6, 0.98, k)
    dup = recs[rng.integers(0, k, int(n * 0.25))].copy()
    dup[:, 3] += rng.integers(0, 4, len(dup))
    dup[:, 4] += rng.integers(-3, 4, len(dup))
    m = n - k - len(dup)
    noise = np.zeros((m, 10))
    noise[:, 0] = rng.integers(0, n_seq, m)
    noise[:, 1] = rng.integers(0, n_seq, m)
    noise[:, 2] = size
    noise[:, 3] = rng.integers(1, size - 2000, m)
    noise[:, 4] = rng.integers(1, size - 2000, m)
    noise[:, 5] = rng.integers(46, 150, m)
    noise[:, 6] = rng.integers(0, 2, m)
    noise[:, 9] = rng.uniform(0.5, 0.8, m)
    allr = np.concatenate([recs, dup, noise])
    allr[:, 7] = allr[:, 9] * allr[:, 5]
    allr[:, 8] = rng.uniform(0.99, 1.0, len(allr))
    rng.shuffle(allr)
    return allr


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1200000
    n_seq, size = 4, 40000000
    recs = synth_matches(n, n_seq, size)
    exe = [e for e in sxbuild.build_host() if e.endswith("XCorrMatchTool")][0]
    tmp = tempfile.mkdtemp()
    src, dst = os.path.join(tmp, "in.match"), os.path.join(tmp, "out.match")
    _write_match_file(src, recs, n_seq, n_seq, size)
    t0 = time.perf_counter()
    out = subprocess.run([exe, "-i", src, "-o", dst, "-sort", "1", "-collapse", "1", "-chain", "1"], check=True,
                         capture_output=True, text=True).stdout
    wall = time.perf_counter() - t0
    mine = [float(x) for x in out.split("seconds: sort ")[1].replace("collapse", "").replace("chain", "").split()]
    n_chain = int(out.split("Matches in the chain:")[1].split()[0])
    n_coll = int(out.split("Matches after collapse:")[1].split()[0])
    R = oracle.Reference()
    L = R.lib
    L.ref_sort_collapse_chain_timed.restype = C.c_long
    L.ref_sort_collapse_chain_timed.argtypes = [C.c_void_p, C.c_long, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                                C.c_void_p, C.c_void_p]
    arr = np.ascontiguousarray(recs, dtype=np.float64)
    sizes = np.full(n_seq, size, dtype=np.int32)
    secs = np.zeros(3)
    ncol = C.c_long()
    k = L.ref_sort_collapse_chain_timed(arr.ctypes.data, len(arr), n_seq, n_seq, sizes.ctypes.data, sizes.ctypes.data, 0,
                                        secs.ctypes.data, C.byref(ncol))
    assert (k, ncol.value) == (n_chain, n_coll), (k, ncol.value, n_chain, n_coll)
    line = {"what": "MultiMatches::Sort / Collapse / RunMatchDynProg on synthetic matches; the reference on its one thread, the "
                    "host code with one chain per target sequence in parallel (identical output)", "targets": n_seq,
            "host_cores": os.cpu_count(),
            "matches": n, "after_collapse": n_coll, "chain": n_chain,
            "reference_s": {"sort": secs[0], "collapse": secs[1], "chain": secs[2], "total": float(secs.sum())},
            "b200_host_s": {"sort": mine[0], "collapse": mine[1], "chain": mine[2], "total": float(sum(mine)),
                            "tool_wall_incl_file_io": wall},
            "speedup_total": float(secs.sum()) / max(sum(mine), 1e-9)}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
