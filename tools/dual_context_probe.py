#!/usr/bin/env python
"""Does running the scan of one batch beside the transforms of another pay?  
Two INDEPENDENT contexts on the same GPU, each on its own stream and host thread, each with half of the pairs: their
kernels overlap at arbitrary phases (scan of one beside encode / correlate of the other).  Aggregate device-resident
throughput against one context doing all pairs.  python tools/dual_context_probe.py [pairs]"""
import json
import sys
import threading
import time

import numpy as np
import torch

sys.path.insert(0, "/root/repo")
import satsuma2_b200 as sx  # noqa: E402
from satsuma2_b200 import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
T, Q, _ = synth.random_pairs(n, 4096, seed=1)
res = {}
for parts in (1, 2, 3):
    per = n // parts
    engs, sets, bufs = [], [], []
    for p in range(parts):
        e = sx.XCorrEngine(target_total=4294967296.0, spectra_cache_bytes=-1)
        e.set_targets(sx.ChunkSet.independent(T[p * per:(p + 1) * per]))
        e.set_queries(sx.ChunkSet.independent(Q[p * per:(p + 1) * per]))
        engs.append(e)
        sets.append(np.ascontiguousarray(np.stack([np.arange(per), np.arange(per)], axis=1), dtype=np.int32))
        bufs.append(np.zeros(2 * per, dtype=sx.RESULT_DTYPE))

    def run(i, reps):
        for _ in range(reps):
            engs[i].align_pairs(sets[i], out=bufs[i])

    for reps, timed in ((1, False), (3, True)):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        th = [threading.Thread(target=run, args=(i, reps)) for i in range(parts)]
        [t.start() for t in th]
        [t.join() for t in th]
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if timed:
            res[parts] = per * parts * reps / dt
    for e in engs:
        e.close()
print(json.dumps({"pairs": n, "chunk_pairs_per_s": res, "gain_2": res[2] / res[1], "gain_3": res[3] / res[1]}))
