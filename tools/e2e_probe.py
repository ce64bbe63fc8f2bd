import sys, time, numpy as np, torch
sys.path.insert(0,'/root/repo')
import satsuma2_b200 as sx
from satsuma2_b200 import synth
n=131072
tt=torch.empty((n,4096),dtype=torch.uint8,pin_memory=True); tq=torch.empty((n,4096),dtype=torch.uint8,pin_memory=True)
T,Q=tt.numpy(),tq.numpy()
synth.random_pairs(n,4096,seed=1,out_t=T,out_q=Q)
cs_t,cs_q=sx.ChunkSet.independent(T),sx.ChunkSet.independent(Q)
eng=sx.XCorrEngine(target_total=float(n)*4096,spectra_cache_bytes=-1)
pairs=np.ascontiguousarray(np.stack([np.arange(n),np.arange(n)],axis=1),dtype=np.int32)
for it in range(3):
    t0=time.perf_counter(); eng.set_targets_raw(T.ctypes.data,cs_t); t1=time.perf_counter(); eng.set_queries_raw(Q.ctypes.data,cs_q); t2=time.perf_counter()
    r=eng.align_pairs(pairs,cap_hint=2*n); t3=time.perf_counter()
    print(f"set_targets {1e3*(t1-t0):.1f} ms  set_queries {1e3*(t2-t1):.1f} ms  align {1e3*(t3-t2):.1f} ms  records {len(r)}")
eng.set_profiling(True); eng.reset_stats()
t0=time.perf_counter(); r=eng.align_pairs(pairs,cap_hint=2*n); t1=time.perf_counter()
st=eng.stats(); print("align wall", 1e3*(t1-t0), "gpu ms_total", st['ms_total'], {k:round(st[k],1) for k in ('ms_encode_fft','ms_xcorr','ms_scan_score')}, "batches", st['batches'])
