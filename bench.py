#!/usr/bin/env python
"""Benchmark of the chunk-pair cross-correlation hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repository's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU path

One "step" = one pass of the whole hot path (encode -> FFT -> product/inverse -> FindTop ->
diagonal scan -> scoring -> match records) over the workload: `--pairs` independent random
4096x4096 chunk pairs per GPU with one planted homologous segment each (BASELINE.json configs[1];
default 1,048,576 pairs).  Under torchrun every rank owns its own shard (weak scaling, no
collective on the data path); timing is CUDA events on the library's stream, max over ranks.

`value`  : device-resident -- chunk bases already in HBM, spectra recomputed every step.
`e2e`    : through the C ABI with HOST (pinned) buffers: sx_set_targets + sx_set_queries (H2D of all
           bases) + sx_align_pairs (records back on the host) inside the timed region.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "chunk_pair_xcorrs_per_sec"
UNIT = "chunk-pairs/s"
CHUNK = 4096
FFT_N = 8192
# Algorithmic work per chunk pair (SURVEY 8d / DESIGN.md): real N-point FFT = 2.5 N log2 N flop.
FLOP_FWD_PER_SIGNAL = 4 * 2.5 * FFT_N * 13            # 4 channels
FLOP_XCORR_PER_STRAND = 2.5 * FFT_N * 13 + 4 * 4097 * 8  # one inverse + spectral MAC over 4 channels
SPECTRUM_BYTES = FFT_N * 8                             # one packed complex spectrum (two real channels), fp32
# Three-channel form: a pure A/C/G/T chunk stores (A + iC) of its own and shares one (G + iG') spectrum
# with its partner -- in this workload the other chunk of its pair -- so a chunk pair writes and reads THREE spectra
# (the four-channel form of round 1: four).  The flop figures above stay the reference's algorithmic work.
SPECTRA_PER_CHUNK_PAIR = 3
SCAN_ALU_OPS_PER_POSITION = 88.0 / 32.0                # bit-sliced window count: 57 LOP3 + 24 SHF + 7 match per 32 positions


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured", d
        except Exception:
            pass
    return 6650.0, "fallback", {}


def load_traffic(key="dram_bytes_per_launch"):
    """dram__bytes_read+write per launch (device batch of 16384 pairs) and the pipe utilisation figures of the same
    `ncu --set full` captures, from profiles/*_traffic.json (tools/make_profiles.py)."""
    for name in ("r2_traffic.json", "r1_traffic.json"):
        try:
            d = json.load(open(os.path.join(ROOT, "profiles", name)))
            if key in d:
                return d[key]
        except Exception:
            continue
    return {}


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.idx = device_index
        self.proc = None
        self.lines = []
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
                power.append(float(parts[2]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        busy = [s for s, p in zip(sm, power) if p > 250.0] or sm
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def make_workload(n_pairs: int, seed: int, pinned: bool):
    """-> (T, Q) uint8 [n, 4096] ASCII arrays (numpy views of pinned torch tensors when pinned)."""
    from satsuma2_b200 import synth

    keep = None
    if pinned:
        import torch

        tt = torch.empty((n_pairs, CHUNK), dtype=torch.uint8, pin_memory=True)
        tq = torch.empty((n_pairs, CHUNK), dtype=torch.uint8, pin_memory=True)
        T, Q = tt.numpy(), tq.numpy()
        keep = (tt, tq)
    else:
        T = np.empty((n_pairs, CHUNK), dtype=np.uint8)
        Q = np.empty((n_pairs, CHUNK), dtype=np.uint8)
    synth.random_pairs(n_pairs, CHUNK, seed=seed, out_t=T, out_q=Q)
    return T, Q, keep


# ------------------------------------------------------------------------------------------------
def cpu_reference_rate(T, Q, n_sample: int, threads: int):
    """Times the reference's own CPU implementation (oracle/_ref, unmodified Satsuma2 compiled from
    source) on the first n_sample pairs; falls back to the C port (oracle/) when it is not built."""
    import oracle

    n_sample = min(n_sample, T.shape[0])
    total = float(T.shape[0]) * CHUNK
    if oracle.have_reference():
        R = oracle.Reference()
        R.configure()
        tl = [(T[i].tobytes(), 0, i, CHUNK) for i in range(n_sample)]
        ql = [(Q[i].tobytes(), 0, i, CHUNK) for i in range(n_sample)]
        R.set_chunks(True, tl, [CHUNK] * n_sample)
        R.set_chunks(False, ql, [CHUNK] * n_sample)
        R.lib.ref_set_target_total(total)
        tp = np.array([[i, i, i, i, 0] for i in range(n_sample)], dtype=np.int32)
        recs, secs, nrec = R.align_pairs_mt(tp, threads)
        assert nrec == len(recs)
        return n_sample / secs, "reference", secs, recs
    O = oracle.Oracle()
    tl = [(T[i].tobytes(), 0, i, CHUNK) for i in range(n_sample)]
    ql = [(Q[i].tobytes(), 0, i, CHUNK) for i in range(n_sample)]
    params = O.make_params(target_total=total)
    pairs = np.stack([np.arange(n_sample), np.arange(n_sample)], axis=1)
    t0 = time.perf_counter()
    rec = O.align_pairs(params, tl, ql, pairs, threads=threads)
    secs = time.perf_counter() - t0
    return n_sample / secs, "port", secs, rec


def parity_check(T, Q, n_sample, gpu_recs, cpu_recs, kind, target_total):
    """The GPU's records of the first n_sample pairs of the timed step against the CPU reference's records of the
    same pairs (each pair is its own sequence: query_id = pair index).  Every difference must be explained by a
    candidate lag within 1e-4 of the FindTop threshold or a probability within 1e-4 of min_prob (tests/parity.py);
    explained ones are listed, anything else counts as unexplained."""
    import oracle

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from parity import compare_pair_records, rec_key

    O = oracle.Oracle()
    g = gpu_recs[gpu_recs["query_id"] < n_sample]
    gk = {rec_key(r) for r in g}
    ck = {rec_key(r) for r in cpu_recs}
    listed, unexplained = [], []
    for i in sorted({k[0] for k in gk ^ ck}):
        try:
            compare_pair_records(O, g[g["query_id"] == i], cpu_recs[cpu_recs["query_id"] == i], T[i].tobytes(),
                                 Q[i].tobytes(), 0, 0, CHUNK, CHUNK, FFT_N, 1.8, 0.99, target_total, listed)
        except AssertionError as e:
            unexplained.append(str(e)[:200])
    ident_ok = True
    cmap = {rec_key(r): r for r in cpu_recs}
    for r in g:
        c = cmap.get(rec_key(r))
        if c is not None and (r["ident"] != c["ident"] or abs(float(r["prob"]) - float(c["prob"])) > 1e-6 * abs(float(c["prob"]))):
            ident_ok = False
    return {"against": kind, "pairs": int(n_sample), "records_gpu": int(len(g)), "records_cpu": int(len(cpu_recs)),
            "differing": len(gk ^ ck), "unexplained": len(unexplained) + (0 if ident_ok else 1),
            "listed": [dict(x, key=[int(v) for v in x["key"]]) for x in listed], "prob_ident_equal": ident_ok,
            "notes": unexplained}


def pairs_workload(grp, local_rank, chunk, n, steps, warmup, batch, seed=7, engine_kw=None, e2e=True):
    """configs[2]: independent random chunk pairs of a larger chunk size (FFT length 2 x chunk), one planted segment
    each -- the headline workload's shape at N = 16384 / 32768.  -> dict (device-resident value, e2e, kernel ms)."""
    import torch

    import satsuma2_b200 as sx
    from satsuma2_b200 import synth

    dev = torch.device("cuda", local_rank)
    tt = torch.empty((n, chunk), dtype=torch.uint8, pin_memory=True)
    tq = torch.empty((n, chunk), dtype=torch.uint8, pin_memory=True)
    T, Q = tt.numpy(), tq.numpy()
    synth.random_pairs(n, chunk, seed=seed, out_t=T, out_q=Q)
    cs_t, cs_q = sx.ChunkSet.independent(T), sx.ChunkSet.independent(Q)
    pairs = np.ascontiguousarray(np.stack([np.arange(n), np.arange(n)], axis=1), dtype=np.int32)
    kw = dict(device=local_rank, t_chunk=chunk, q_chunk=chunk, target_total=float(n) * chunk, max_batch_pairs=batch,
              spectra_cache_bytes=-1, async_upload=1)
    kw.update(engine_kw or {})
    eng = sx.XCorrEngine(**kw)
    stream = torch.cuda.ExternalStream(eng.stream_handle(), device=dev)
    rec_buf = np.zeros(4 * n, dtype=sx.RESULT_DTYPE)

    def step_device():
        return eng.align_pairs(pairs, out=rec_buf)

    def step_e2e():
        eng.set_targets_raw(T.ctypes.data, cs_t)
        eng.set_queries_raw(Q.ctypes.data, cs_q)
        return eng.align_pairs(pairs, out=rec_buf)

    eng.set_targets_raw(T.ctypes.data, cs_t)
    eng.set_queries_raw(Q.ctypes.data, cs_q)
    eng.set_profiling(True)
    ms_dev, clocks, rec, n_dev = _timed_steps(step_device, steps, warmup, stream, grp, local_rank, eng.reset_stats, 1.0)
    st = eng.stats()
    eng.set_profiling(False)
    b = max(st["batches"], 1)
    if not e2e:  # A/B leg: device-resident figures only
        eng.close()
        return {"value": n / (ms_dev / 1e3), "ms_per_step": ms_dev, "steps": n_dev,
                "fused_pairs_per_step": int(st["fused_pairs"] / max(n_dev, 1)),
                "kernel_ms_per_batch": {"encode_fft": st["ms_encode_fft"] / b, "xcorr_findtop": st["ms_xcorr"] / b,
                                        "scan_score": st["ms_scan_score"] / b}, "records_per_step": int(len(rec))}
    ms_e2e, _, rec, n_e2e = _timed_steps(step_e2e, steps, 1, stream, grp, local_rank, eng.reset_stats, 1.0)
    st_e2e = eng.stats()
    out = {"workload": f"configs[2]: {n} random {chunk}x{chunk} chunk pairs (FFT length {2 * chunk}), one planted "
                       "segment each", "chunk": chunk, "fft_n": 2 * chunk, "pairs_per_step": n, "device_batch_pairs": batch,
           "metric": METRIC, "unit": UNIT, "value": n / (ms_dev / 1e3), "ms_per_step": ms_dev, "steps": n_dev,
           "e2e": {"value": n / (ms_e2e / 1e3), "unit": UNIT, "ms_per_step": ms_e2e, "steps": n_e2e,
                   "h2d_bytes_per_step": int(st_e2e["h2d_bytes"] / n_e2e), "d2h_bytes_per_step": int(st_e2e["d2h_bytes"] / n_e2e)},
           "kernel_ms_per_batch": {"encode_fft": st["ms_encode_fft"] / b, "xcorr_findtop": st["ms_xcorr"] / b,
                                   "scan_score": st["ms_scan_score"] / b},
           "records_per_step": int(len(rec)), "gpu_launches": int(st["kernel_launches"] / max(n_dev, 1)),
           "per_pair": {"candidates": st["candidates"] / max(st["chunk_pairs"], 1),
                        "matches": st["matches"] / max(st["chunk_pairs"], 1)}, "clocks": clocks}
    eng.close()
    return out


def repeats_workload(grp, local_rank, genome_mb, steps, warmup, batch):
    """configs[4]: repeat-rich synthetic genome pair (tandem + interspersed repeat families, low-complexity tracts)
    in -prob_table 1 mode (slave semantics; -dups 1 only changes the master's chaining): high candidate density,
    thousands of kept records per block.  Blocks of 24x24 chunks along the diagonal, one GPU."""
    import torch

    import satsuma2_b200 as sx
    from satsuma2_b200 import synth

    dev = torch.device("cuda", local_rank)
    L = int(genome_mb * 1e6)
    a, b = synth.repeat_rich_pair(L, seed=21)
    to, tl, ts = synth.chunk_sequence(a, CHUNK, CHUNK // 4)
    qo, ql, qs = synth.chunk_sequence(b, CHUNK, 0)
    ta, tb = torch.from_numpy(a).pin_memory(), torch.from_numpy(b).pin_memory()
    cs_t = sx.ChunkSet(ta.numpy(), to, tl, ts, np.zeros(len(tl), np.int32), [L])
    cs_q = sx.ChunkSet(tb.numpy(), qo, ql, qs, np.zeros(len(ql), np.int32), [L])
    blocks = sx.make_blocks(synth.diagonal_blocks(len(tl), len(ql), CHUNK - CHUNK // 4, CHUNK, pixel=24))
    t0 = time.perf_counter()
    tab = sx.build_prob_table(float(L))
    table_s = time.perf_counter() - t0
    eng = sx.MultiEngine(devices=[local_rank], target_total=float(L), use_prob_table=1, prob_table_value=0.9999,
                         max_batch_pairs=batch)
    eng.set_prob_table(tab)
    stream = torch.cuda.ExternalStream(eng.stream_handle(0), device=dev)
    rec_buf = np.zeros(1 << 22, dtype=sx.RESULT_DTYPE)

    def step_device():
        eng.invalidate_spectra()
        return eng.align_blocks(blocks, out=rec_buf)

    def step_e2e():
        eng.set_targets(cs_t)
        eng.set_queries(cs_q)
        return eng.align_blocks(blocks, out=rec_buf)

    eng.set_targets(cs_t)
    eng.set_queries(cs_q)
    ms_dev, clocks, rec, n_dev = _timed_steps(step_device, steps, warmup, stream, grp, local_rank, eng.reset_stats, 1.0)
    st = eng.stats()
    ms_e2e, _, rec, n_e2e = _timed_steps(step_e2e, steps, 1, stream, grp, local_rank, eng.reset_stats, 1.0)
    st_e2e = eng.stats()
    n = st["chunk_pairs"] / n_dev
    out = {"workload": f"configs[4]: repeat-rich {genome_mb:g} Mb synthetic genome pair, -prob_table 1, 24x24-chunk blocks "
                       "along the diagonal", "pairs_per_step": int(n), "metric": METRIC, "unit": UNIT,
           "value": n / (ms_dev / 1e3), "ms_per_step": ms_dev, "steps": n_dev,
           "e2e": {"value": n / (ms_e2e / 1e3), "unit": UNIT, "ms_per_step": ms_e2e, "steps": n_e2e,
                   "h2d_bytes_per_step": int(st_e2e["h2d_bytes"] / n_e2e), "d2h_bytes_per_step": int(st_e2e["d2h_bytes"] / n_e2e)},
           "per_pair": {"candidates": st["candidates"] / max(st["chunk_pairs"], 1),
                        "segments": st["segments"] / max(st["chunk_pairs"], 1),
                        "matches": st["matches"] / max(st["chunk_pairs"], 1)},
           "records_per_step": int(len(rec)), "prob_table_build_s": table_s,
           "gpu_launches": int(st["kernel_launches"] / max(n_dev, 1)),
           "clocks": clocks}
    eng.close()
    return out


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    # each step: a bounded sample (~3 s of all-core CPU work at ~36 pairs/s/thread)
    n_step = max(16, min(args.pairs, args.ref_pairs_per_core * cores))
    T, Q, _ = make_workload(n_step, seed=1000 + 0, pinned=False)
    times, kind, nrec = [], "reference", 0
    for it in range(args.warmup + args.steps):
        rate, kind, secs, recs = cpu_reference_rate(T, Q, n_step, cores)
        nrec = len(recs)
        if it >= args.warmup:
            times.append(secs)
    ms = 1e3 * sum(times) / max(len(times), 1)
    value = n_step / (ms / 1e3)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[1]: random 4096x4096 chunk pairs, one planted 60-500 bp segment at 70-95% "
                               "identity per pair (bounded sample per step)", "pairs_per_step": n_step, "chunk": CHUNK,
                   "fft_n": FFT_N, "cutoff": 1.8, "min_prob": 0.99},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{n_step} pairs per step through HomologyByXCorr::align_target on {cores} threads"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "records_per_step": nrec,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------
def _timed_steps(fn, steps, warmup, stream, grp, local_rank, reset=None, min_seconds=0.0):
    """warm-up, then `steps` calls bracketed by barrier + synchronize, CUDA events on the library's stream,
    max over ranks; nvidia-smi sampled meanwhile.  -> (ms PER STEP, clocks, last result, steps timed).
    min_seconds: short steps are repeated until the timed region is long enough for the clock sampler
    (nvidia-smi reports every 200 ms); every rank derives the same count from the slowest rank's warm-up step."""
    import torch

    for _ in range(warmup):
        res = fn()
    if min_seconds > 0:
        torch.cuda.synchronize()
        grp.barrier()
        t0 = time.perf_counter()
        res = fn()
        torch.cuda.synchronize()
        one = grp.max(time.perf_counter() - t0)
        steps = int(max(steps, min(200, np.ceil(min_seconds / max(one, 1e-4)))))
    if reset is not None:
        reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    grp.barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0.record(stream)
    for _ in range(steps):
        res = fn()
    e1.record(stream)
    e1.synchronize()
    torch.cuda.synchronize()
    grp.barrier()
    return grp.max(e0.elapsed_time(e1)) / steps, sampler.stop(), res, steps


def grid_workload(args, grp, rank, local_rank, world, genome_mb, steps, warmup):
    """configs[3]: GridSearch-style 24x24-chunk blocks along the syntenic diagonal of a synthetic genome pair.
    STRONG scaling: the target chunk list is cut into `world` contiguous ranges (sx_multi, shard = rank); a rank is
    sent only the bases of its target range and of the query chunks its blocks touch, keeps its target spectra in
    HBM (cold at the start of every step) and returns its own records -- no collective.  -> dict for the JSON line."""
    import torch

    import satsuma2_b200 as sx
    from satsuma2_b200 import synth

    dev = torch.device("cuda", local_rank)
    L = int(genome_mb * 1e6)
    tgt, qry = synth.genome_pair(L, seed=11)
    to, tl, ts = synth.chunk_sequence(tgt, CHUNK, CHUNK // 4)  # slave / grid overlap = size / 4 (Slave.cc:400)
    qo, ql, qs = synth.chunk_sequence(qry, CHUNK, 0)
    # pinned host copies: the e2e leg uploads from them every step
    tt, tq = torch.from_numpy(tgt).pin_memory(), torch.from_numpy(qry).pin_memory()
    cs_t = sx.ChunkSet(tt.numpy(), to, tl, ts, np.zeros(len(tl), np.int32), [L])
    cs_q = sx.ChunkSet(tq.numpy(), qo, ql, qs, np.zeros(len(ql), np.int32), [L])
    blocks = sx.make_blocks(synth.diagonal_blocks(len(tl), len(ql), CHUNK - CHUNK // 4, CHUNK, pixel=24))
    # one process per GPU (torchrun: shard = rank) or, with --inproc N, ONE process driving N GPUs through the same handle
    devices = list(range(args.inproc)) if args.inproc > 0 else [local_rank]
    eng = sx.MultiEngine(devices=devices, shard_rank=rank, shard_world=world, target_total=float(L),
                         max_batch_pairs=args.batch)
    stream = torch.cuda.ExternalStream(eng.stream_handle(0), device=dev)
    n_all = int(((blocks["target_to"] - blocks["target_from"] + 1).astype(np.int64) *
                 (blocks["query_to"] - blocks["query_from"] + 1)).sum())
    rec_buf = np.zeros(max(1 << 16, 4 * n_all // world + (1 << 16)), dtype=sx.RESULT_DTYPE)
    if len(devices) > 1:
        torch.cuda.synchronize()

    def step_device():
        eng.invalidate_spectra()
        return eng.align_blocks(blocks, out=rec_buf)

    def step_e2e():
        eng.set_targets(cs_t)
        eng.set_queries(cs_q)
        return eng.align_blocks(blocks, out=rec_buf)

    eng.set_targets(cs_t)
    eng.set_queries(cs_q)
    ms_dev, clocks, rec, n_dev = _timed_steps(step_device, steps, warmup, stream, grp, local_rank, eng.reset_stats, 1.5)
    st = eng.stats()
    ms_e2e, clocks_e2e, rec, n_e2e = _timed_steps(step_e2e, steps, max(1, warmup - 2), stream, grp, local_rank,
                                                 eng.reset_stats, 1.5)
    st_e2e = eng.stats()
    mine = st["chunk_pairs"] / max(n_dev, 1)
    total_pairs = grp.sum(float(mine))
    h2d_max = grp.max(float(st_e2e["h2d_bytes"]) / max(n_e2e, 1))
    h2d_sum = grp.sum(float(st_e2e["h2d_bytes"]) / max(n_e2e, 1))
    d2h_sum = grp.sum(float(st_e2e["d2h_bytes"]) / max(n_e2e, 1))
    recs = grp.sum(float(len(rec)))
    out = {
        "workload": f"configs[3]: {genome_mb:g} Mb x {genome_mb:g} Mb synthetic genome pair, 24x24-chunk blocks along "
                    "the diagonal, target spectra cached in HBM (cold at the start of every step), target chunk "
                    "list split by range over the ranks (sx_multi), no collective",
        "scaling": "strong", "metric": METRIC, "unit": UNIT, "n_gpus": world * len(devices),
        "processes": world, "gpus_per_process": len(devices),
        "value": total_pairs / (ms_dev / 1e3), "ms_per_step": ms_dev, "steps": n_dev,
        "e2e": {"value": total_pairs / (ms_e2e / 1e3), "unit": UNIT, "ms_per_step": ms_e2e, "steps": n_e2e,
                "h2d_bytes_per_step": int(h2d_sum), "h2d_bytes_per_step_max_rank": int(h2d_max),
                "d2h_bytes_per_step": int(d2h_sum)},
        "target_chunks": int(len(tl)), "query_chunks": int(len(ql)), "blocks": len(blocks),
        "pairs_per_step_all_ranks": int(total_pairs), "target_total": float(L),
        "signals_per_pair": st["signals"] / max(st["chunk_pairs"], 1), "records_per_step": int(recs),
        "gpu_launches": int(st["kernel_launches"] / max(n_dev, 1)), "clocks": clocks, "clocks_e2e": clocks_e2e,
    }
    assert int(total_pairs) == n_all, (total_pairs, n_all)
    eng.close()
    del tt, tq
    return out


def run_grid(args):
    """--workload grid: the configs[3] line on its own (secondary workload, not the headline metric)."""
    import torch

    from satsuma2_b200 import build as sxbuild
    from satsuma2_b200.dist import Group

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback")
    sxbuild.build()
    rank, local_rank, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
    torch.cuda.set_device(local_rank)
    grp = Group("nccl", torch.device("cuda", local_rank))
    g = grid_workload(args, grp, rank, local_rank, world, args.genome_mb, args.steps, args.warmup)
    if rank == 0:
        line = {"metric": METRIC, "value": g["value"], "unit": UNIT, "n_gpus": g["n_gpus"], "steps": g["steps"],
                "warmup": args.warmup, "ms_per_step": g["ms_per_step"], "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": g["workload"], "chunk": CHUNK, "fft_n": FFT_N, "device_batch_pairs": args.batch},
                "clocks": g["clocks"], "e2e": g["e2e"], "gpu_launches": g["gpu_launches"], "grid": g}
        print(json.dumps(line), flush=True)
    grp.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs", type=int, default=int(os.environ.get("SX_BENCH_PAIRS", 1 << 20)),
                    help="chunk pairs per GPU per step (configs[1] = 1M)")
    ap.add_argument("--batch", type=int, default=int(os.environ.get("SX_BENCH_BATCH", 16384)))
    ap.add_argument("--cpu-sample-per-core", type=int, default=400)
    ap.add_argument("--ref-pairs-per-core", type=int, default=100)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--debug-flags", type=int, default=0,
                    help="sx_config::debug_flags of the headline engine (A/B runs: 4 = four channels per chunk)")
    ap.add_argument("--fuse-pairs", type=int, default=0,
                    help="sx_config::fuse_pairs of the headline engine (A/B runs: 1 = fused transform + correlation kernel)")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary workloads (grid / chunk sizes / repeats)")
    ap.add_argument("--workload", default="pairs", choices=["pairs", "grid"],
                    help="pairs = configs[1] (the headline metric); grid = configs[3]-style block search of a synthetic "
                         "genome pair with target spectra cached in HBM, sharded by target range (strong scaling)")
    ap.add_argument("--genome-mb", type=float, default=150.0,
                    help="grid workload (configs[3]): bases per genome, in millions")
    ap.add_argument("--inproc", type=int, default=0,
                    help="--workload grid: ONE process drives this many GPUs through sx_multi (host threads + host gather) "
                         "instead of one torchrun rank per GPU")
    ap.add_argument("--target-total", type=float, default=0.0,
                    help="targetTotal of the probability filter (default: pairs x chunk, i.e. every target chunk of the "
                         "step); profiling runs with fewer pairs pass the full step's 4294967296")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3  # timing hygiene: at least 3 warm-up steps
    # stdout carries ONE JSON line: NCCL's "NCCL version ..." banner (NCCL_DEBUG=VERSION) would precede it
    # (any level from VERSION up prints it; INFO and above are left alone, whoever set them wants the log)
    if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
        del os.environ["NCCL_DEBUG"]

    if args.impl == "reference":
        return run_reference_arm(args)
    if args.workload == "grid":
        return run_grid(args)

    import torch

    import satsuma2_b200 as sx
    from satsuma2_b200 import build as sxbuild
    from satsuma2_b200.dist import Group, shard_seed

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback")
    sxbuild.build()
    rank, local_rank, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    grp = Group("nccl", dev)
    n = args.pairs

    T, Q, keep = make_workload(n, seed=shard_seed(1, rank), pinned=True)
    cs_t, cs_q = sx.ChunkSet.independent(T), sx.ChunkSet.independent(Q)
    pairs = np.ascontiguousarray(np.stack([np.arange(n), np.arange(n)], axis=1), dtype=np.int32)

    # spectra are never kept across steps (cache disabled): every step redoes the whole path
    target_total = args.target_total if args.target_total > 0 else float(n) * CHUNK
    eng = sx.XCorrEngine(device=local_rank, target_total=target_total, max_batch_pairs=args.batch,
                         spectra_cache_bytes=-1, async_upload=1, debug_flags=args.debug_flags, fuse_pairs=args.fuse_pairs)
    stream = torch.cuda.ExternalStream(eng.stream_handle(), device=dev)
    t_ptr, q_ptr = T.ctypes.data, Q.ctypes.data
    h2d_per_step = int(T.nbytes + Q.nbytes)

    # caller-owned host buffer for the t_result records, reused by every step (as a slave would)
    rec_buf = np.zeros(2 * n, dtype=sx.RESULT_DTYPE)

    def step_device():
        return eng.align_pairs(pairs, out=rec_buf)

    def step_e2e():
        eng.set_targets_raw(t_ptr, cs_t)
        eng.set_queries_raw(q_ptr, cs_q)
        return eng.align_pairs(pairs, out=rec_buf)

    def timed(fn, steps, warmup, profile):
        for _ in range(warmup):
            rec = fn()
        eng.reset_stats()
        eng.set_profiling(profile)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        grp.barrier()
        sampler = ClockSampler(local_rank)
        sampler.start()
        e0.record(stream)
        for _ in range(steps):
            rec = fn()
        e1.record(stream)
        e1.synchronize()
        torch.cuda.synchronize()
        grp.barrier()
        clocks = sampler.stop()
        ms = grp.max(e0.elapsed_time(e1))
        st = eng.stats()
        eng.set_profiling(False)
        return ms, st, clocks, rec

    # ---- device-resident: bases uploaded once, outside the timed region
    eng.set_targets_raw(t_ptr, cs_t)
    eng.set_queries_raw(q_ptr, cs_q)
    ms_dev, st_dev, clocks, rec = timed(step_device, args.steps, args.warmup, True)
    # ---- end to end: host buffers in, host records out, every step
    ms_e2e, st_e2e, clocks_e2e, rec_e2e = timed(step_e2e, args.steps, max(1, args.warmup - 2), False)

    total_pairs = grp.sum(float(n)) * args.steps
    n_records = int(len(rec))
    value = total_pairs / (ms_dev / 1e3)
    e2e_value = total_pairs / (ms_e2e / 1e3)
    d2h_per_step = int(st_e2e["d2h_bytes"] / max(args.steps, 1))

    # ---- rooflines (per-kernel CUDA-event time from the library, measured in THIS run)
    hbm_peak, peak_kind, peaks = load_peaks()
    batches = max(st_dev["batches"], 1)
    traffic = load_traffic()  # dram bytes per launch from the committed ncu --set full captures
    ncu = load_traffic("ncu_per_launch")  # pipe utilisation of the same captures (counters, not a model)
    # with --fuse-pairs 1 the fused kernel does the transforms of its pairs too (its time is under ms_xcorr) and their
    # spectra never reach HBM: only the bases are read
    ff = st_dev["fused_pairs"] / max(st_dev["chunk_pairs"], 1)
    kern = {
        "encode_fft": {"ms": st_dev["ms_encode_fft"], "launches": batches,
                       "flop": FLOP_FWD_PER_SIGNAL * st_dev["signals"] * (1 - ff),
                       "bytes": (SPECTRA_PER_CHUNK_PAIR * SPECTRUM_BYTES / 2 * (1 - ff) + CHUNK) * st_dev["signals"]},
        "xcorr_findtop": {"ms": st_dev["ms_xcorr"], "launches": batches,
                          "flop": FLOP_XCORR_PER_STRAND * st_dev["strand_pairs"] + FLOP_FWD_PER_SIGNAL * st_dev["signals"] * ff,
                          # the spectra of a chunk pair read once (both strands derived from them)
                          "bytes": SPECTRA_PER_CHUNK_PAIR * SPECTRUM_BYTES * st_dev["chunk_pairs"] * (1 - ff)},
        "scan_score": {"ms": st_dev["ms_scan_score"], "launches": batches, "flop": 0.0,
                       "bytes": (4 * (FFT_N // 32) * 4) * st_dev["strand_pairs"] + 2 * st_dev["candidates"],
                       "positions": float(st_dev["positions"])},
    }
    for k, v in kern.items():
        secs = max(v["ms"], 1e-9) / 1e3
        v["gbs"] = v["bytes"] / secs / 1e9
        v["tflops"] = v["flop"] / secs / 1e12
        v["share"] = v["ms"] / max(st_dev["ms_total"], 1e-9)
        v["avg_launch_ms"] = v["ms"] / v["launches"]
    dom = max(kern, key=lambda k: kern[k]["ms"])
    sm_mhz = clocks.get("sm_mhz") or 1965.0
    fp32_peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12  # TFLOP/s at the clock seen under load
    # Integer roofline of the diagonal scan: LOP3/SHF issue on the ALU pipe, 64 lanes/clk/SM (half rate).
    # Algorithmic cost = the bit-sliced 46-window count: 88 logic/shift ops per 32 positions.
    alu_peak = 148 * 64 * sm_mhz * 1e6 / 1e12  # T lane-ops/s
    scan_ops = kern["scan_score"]["positions"] * SCAN_ALU_OPS_PER_POSITION
    scan_tops = scan_ops / max(kern["scan_score"]["ms"] / 1e3, 1e-9) / 1e12
    per_kernel = {k: {"ms": round(v["ms"], 3), "share": round(v["share"], 4), "avg_launch_ms": round(v["avg_launch_ms"], 4),
                      "GB/s": round(v["gbs"], 1), "hbm_frac": round(v["gbs"] / hbm_peak, 4),
                      "TFLOP/s": round(v["tflops"], 3), "fp32_frac": round(v["tflops"] / fp32_peak, 4),
                      "traffic": traffic.get(k), "ncu": ncu.get(k)} for k, v in kern.items()}
    per_kernel["scan_score"].update({"Tintop/s": round(scan_tops, 3), "alu_frac": round(scan_tops / alu_peak, 4),
                                     "positions_per_s": kern["scan_score"]["positions"] /
                                     max(kern["scan_score"]["ms"] / 1e3, 1e-9)})
    if dom == "scan_score":
        # the dominant kernel is bound by ALU-pipe instruction issue, neither by HBM (0.3 %) nor tensor cores
        roofline = {"kernel": dom, "bound": "alu", "achieved": scan_tops, "peak": alu_peak, "unit": "Tintop/s",
                    "frac": scan_tops / alu_peak, "traffic": traffic.get(dom), "peak_source": "148 SM x 64 lanes x clock",
                    "model": "algorithmic cost = the brute-force bit-sliced 46-window count, 88 logic/shift ops per 32 "
                             "positions x positions scanned (counted by the kernel); the kernel itself filters and "
                             "prunes, so frac can exceed what its own instruction stream issues -- the counters of "
                             "the ncu capture are under kernels.scan_score.ncu",
                    "hbm_frac": kern[dom]["gbs"] / hbm_peak}
    elif dom == "xcorr_findtop":
        roofline = {"kernel": dom, "bound": "hbm", "achieved": kern[dom]["gbs"], "peak": hbm_peak, "unit": "GB/s",
                    "frac": kern[dom]["gbs"] / hbm_peak, "traffic": traffic.get(dom), "peak_source": peak_kind}
    else:
        roofline = {"kernel": dom, "bound": "fp32", "achieved": kern[dom]["tflops"], "peak": fp32_peak,
                    "unit": "TFLOP/s", "frac": kern[dom]["tflops"] / fp32_peak, "traffic": traffic.get(dom),
                    "peak_source": "148 SM x 128 lanes x 2 x clock", "hbm_frac": kern[dom]["gbs"] / hbm_peak}
    roofline.update({"avg_launch_ms": kern[dom]["avg_launch_ms"], "share_of_step": kern[dom]["share"],
                     "sm_mhz": sm_mhz, "kernels": per_kernel,
                     # the HBM-bound kernel of the path, against the measured copy bandwidth
                     "hbm_bound_kernel": {"kernel": "xcorr_findtop", "bound": "hbm",
                                          "achieved": kern["xcorr_findtop"]["gbs"], "peak": hbm_peak, "unit": "GB/s",
                                          "frac": kern["xcorr_findtop"]["gbs"] / hbm_peak,
                                          "traffic": traffic.get("xcorr_findtop"), "peak_source": peak_kind}})

    cpu_baseline, parity = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        ns = min(n, args.cpu_sample_per_core * cores)
        rate, kind, secs, cpu_recs = cpu_reference_rate(T, Q, ns, cores)
        cpu_baseline = {"value": rate, "unit": UNIT, "cores": cores, "kind": kind,
                        "sample": f"first {ns} pairs of the same workload, {secs:.1f} s on {cores} threads"}
        # the timed step's own output, checked: the reference's records of those pairs vs the GPU's
        parity = parity_check(T, Q, ns, rec, cpu_recs, kind, target_total)

    # ---- secondary workloads (sub-objects of the same line; the headline stays configs[1])
    eng.close()
    eng = None
    del T, Q, keep, cs_t, cs_q, rec_buf
    extras = {}
    if not args.no_extras:
        extras["grid"] = grid_workload(args, grp, rank, local_rank, world, args.genome_mb, args.steps, args.warmup)
        if world == 1:
            extras["chunk8192"] = pairs_workload(grp, local_rank, 8192, 32768, args.steps, args.warmup, 8192)
            extras["chunk16384"] = pairs_workload(grp, local_rank, 16384, 8192, args.steps, args.warmup, 4096)
            # A/B legs (device-resident, records must agree): N = 32768 through the round-1 route (two half kernels,
            # HBM scratch, combine kernel) instead of the two-CTA cluster kernel; the headline shape through the fused
            # transform + correlation kernel (sx_config::fuse_pairs) instead of the separate kernels
            ab = pairs_workload(grp, local_rank, 16384, 8192, args.steps, args.warmup, 4096, engine_kw={"debug_flags": 8}, e2e=False)
            ab["same_records"] = ab["records_per_step"] == extras["chunk16384"]["records_per_step"]
            extras["chunk16384"]["three_kernel_route"] = ab
            nf = min(n, 262144)
            legs = {}
            for name, fp in (("separate_kernels", 0), ("fused_kernel", 1)):
                legs[name] = pairs_workload(grp, local_rank, CHUNK, nf, args.steps, args.warmup, args.batch, seed=1,
                                            engine_kw={"fuse_pairs": fp, "target_total": target_total}, e2e=False)
            legs["same_records"] = legs["separate_kernels"]["records_per_step"] == legs["fused_kernel"]["records_per_step"]
            legs["note"] = ("sx_config::fuse_pairs = 1: transforms, product, inverse and FindTop of a chunk pair in one "
                            "kernel, spectra never in HBM (its time is under xcorr_findtop); default 0, the faster one")
            extras["fused_ab"] = legs
            extras["repeats"] = repeats_workload(grp, local_rank, 8.0, args.steps, args.warmup, args.batch)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[1]: random 4096x4096 chunk pairs, one planted 60-500 bp segment at "
                                   "70-95% identity per pair, half reverse strand",
                       "pairs_per_gpu_per_step": n, "chunk": CHUNK, "fft_n": FFT_N, "cutoff": 1.8, "min_prob": 0.99,
                       "device_batch_pairs": args.batch, "parallelism": f"pairs sharded over {world} GPU(s), no collective",
                       "l2": f"inputs larger than L2: {h2d_per_step >> 20} MiB of bases and "
                             f"{(args.batch * SPECTRA_PER_CHUNK_PAIR * SPECTRUM_BYTES) >> 20} MiB of spectra per device batch"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_per_step,
                    "d2h_bytes_per_step": d2h_per_step, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(st_dev["kernel_launches"]),
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "parity_check": parity,
            "records_per_step": n_records,
            "per_pair": {"candidates": st_dev["candidates"] / max(st_dev["chunk_pairs"], 1),
                         "segments": st_dev["segments"] / max(st_dev["chunk_pairs"], 1),
                         "matches": st_dev["matches"] / max(st_dev["chunk_pairs"], 1)},
        }
        line.update(extras)
        print(json.dumps(line), flush=True)
    if eng is not None:
        eng.close()
    grp.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
