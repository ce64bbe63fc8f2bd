import sys, time, numpy as np, torch
sys.path.insert(0, "/root/repo")
import satsuma2_b200 as sx
from satsuma2_b200 import synth
n=1<<20; chunk=4096
tt = torch.empty((n, chunk), dtype=torch.uint8, pin_memory=True); tq = torch.empty((n, chunk), dtype=torch.uint8, pin_memory=True)
T,Q=tt.numpy(),tq.numpy()
synth.random_pairs(n, chunk, seed=1, out_t=T, out_q=Q)
cs_t, cs_q = sx.ChunkSet.independent(T), sx.ChunkSet.independent(Q)
pairs = np.ascontiguousarray(np.stack([np.arange(n), np.arange(n)], axis=1), dtype=np.int32)
eng = sx.XCorrEngine(target_total=float(n)*chunk, spectra_cache_bytes=-1, async_upload=1)
rec = np.zeros(2*n, dtype=sx.RESULT_DTYPE)
for it in range(4):
    t0=time.perf_counter(); eng.set_targets_raw(T.ctypes.data, cs_t); t1=time.perf_counter(); eng.set_queries_raw(Q.ctypes.data, cs_q); t2=time.perf_counter()
    r=eng.align_pairs(pairs, out=rec); t3=time.perf_counter()
    print(f"set_t {1e3*(t1-t0):.1f} ms set_q {1e3*(t2-t1):.1f} ms align {1e3*(t3-t2):.1f} ms total {1e3*(t3-t0):.1f}")
for it in range(2):
    t2=time.perf_counter(); r=eng.align_pairs(pairs, out=rec); t3=time.perf_counter()
    print(f"device-resident align {1e3*(t3-t2):.1f} ms")
