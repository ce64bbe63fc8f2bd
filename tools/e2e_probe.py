"""Where does the end-to-end step spend its time?  Wall-clock split of sx_set_targets / sx_set_queries /
sx_align_pairs with blocking and asynchronous uploads, plus the per-kernel CUDA-event times."""
import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
import satsuma2_b200 as sx
from satsuma2_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
tt = torch.empty((n, 4096), dtype=torch.uint8, pin_memory=True); tq = torch.empty((n, 4096), dtype=torch.uint8, pin_memory=True)
T, Q = tt.numpy(), tq.numpy()
synth.random_pairs(n, 4096, seed=1, out_t=T, out_q=Q)
cs_t, cs_q = sx.ChunkSet.independent(T), sx.ChunkSet.independent(Q)
pairs = np.ascontiguousarray(np.stack([np.arange(n), np.arange(n)], axis=1), dtype=np.int32)
for mode in (0, 1):
    eng = sx.XCorrEngine(target_total=float(n) * 4096, spectra_cache_bytes=-1, async_upload=mode)
    for it in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter(); eng.set_targets_raw(T.ctypes.data, cs_t); t1 = time.perf_counter()
        eng.set_queries_raw(Q.ctypes.data, cs_q); t2 = time.perf_counter()
        r = eng.align_pairs(pairs, cap_hint=2 * n); t3 = time.perf_counter()
        print(f"async={mode} set_targets {1e3*(t1-t0):.1f} ms  set_queries {1e3*(t2-t1):.1f} ms  align {1e3*(t3-t2):.1f} ms  total {1e3*(t3-t0):.1f}  records {len(r)}")
    t0 = time.perf_counter(); r = eng.align_pairs(pairs, cap_hint=2 * n); t1 = time.perf_counter()
    print(f"async={mode} resident align {1e3*(t1-t0):.1f} ms")
    eng.set_profiling(True); eng.reset_stats()
    t0 = time.perf_counter(); r = eng.align_pairs(pairs, cap_hint=2 * n); t1 = time.perf_counter()
    st = eng.stats(); print("  profiled align wall", round(1e3*(t1-t0),1), "gpu ms_total", round(st['ms_total'],1), {k: round(st[k], 1) for k in ('ms_encode_fft', 'ms_xcorr', 'ms_scan_score')}, "batches", st['batches'])
    eng.close()
print("---- e2e with profiling (kernel event times while uploads are in flight)")
eng = sx.XCorrEngine(target_total=float(n) * 4096, spectra_cache_bytes=-1, async_upload=1)
eng.set_profiling(True)
for it in range(3):
    eng.reset_stats()
    t0 = time.perf_counter(); eng.set_targets_raw(T.ctypes.data, cs_t); eng.set_queries_raw(Q.ctypes.data, cs_q)
    r = eng.align_pairs(pairs, cap_hint=2 * n); t1 = time.perf_counter()
    st = eng.stats(); print("  e2e wall", round(1e3*(t1-t0),1), "gpu ms_total", round(st['ms_total'],1), {k: round(st[k], 1) for k in ('ms_encode_fft', 'ms_xcorr', 'ms_scan_score')})
eng.close()
