"""Do two engines (two CUDA streams) on one GPU overlap productively?  Aggregate throughput of
1 vs 2 vs 3 concurrent engines, each on its own shard."""
import sys, time, threading, numpy as np, torch
sys.path.insert(0, '/root/repo')
import satsuma2_b200 as sx
from satsuma2_b200 import synth
n = 98304
T, Q, _ = synth.random_pairs(n, 4096, seed=1)
def make(k, parts):
    lo, hi = k * n // parts, (k + 1) * n // parts
    e = sx.XCorrEngine(target_total=float(n) * 4096, spectra_cache_bytes=-1, max_batch_pairs=int(sys.argv[1]) if len(sys.argv) > 1 else 16384)
    e.set_targets(sx.ChunkSet.independent(T[lo:hi])); e.set_queries(sx.ChunkSet.independent(Q[lo:hi]))
    m = hi - lo
    return e, np.ascontiguousarray(np.stack([np.arange(m), np.arange(m)], axis=1), dtype=np.int32)
for parts in (1, 2, 3):
    engs = [make(k, parts) for k in range(parts)]
    def run(e, p, out, i): out[i] = len(e.align_pairs(p, cap_hint=2 * len(p)))
    for rep in range(3):
        out = [0] * parts
        th = [threading.Thread(target=run, args=(e, p, out, i)) for i, (e, p) in enumerate(engs)]
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for t in th: t.start()
        for t in th: t.join()
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"{parts} engine(s): {n / dt / 1e6:.3f} M pairs/s  ({dt * 1e3:.1f} ms, records {sum(out)})")
    for e, _ in engs: e.close()
