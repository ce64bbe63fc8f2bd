#!/usr/bin/env python
"""Regenerates profiles/<round>_* from the ncu captures of one gpurun call (tools/ncu_capture.sh <tag> + launch list).

    python tools/make_profiles.py <tag> [bench-json ...]      (round prefix = first two characters of the tag, e.g. r2)
"""
import collections
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1]
RND = tag[:2]
kernels = {"encode_fft": "encode_fft_kernel", "xcorr_findtop": "xcorr_pair_kernel", "scan_score": "scan_score_kernel"}
traffic = {}
metrics = {}
for key, k in kernels.items():
    rep = os.path.join(G, f"prof_{k}_{tag}.ncu-rep")
    txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep, "30"],
                         capture_output=True, text=True).stdout
    open(os.path.join(P, f"{RND}_ncu_{k}.txt"), "w").write(
        f"# ncu --set full --clock-control none --import-source on -k regex:{k} -s 4 -c 1  python bench.py --pairs 32768 "
        f"--target-total 4294967296 --steps 1 --warmup 3 --no-cpu-baseline --no-extras   (capture tag {tag}; tools/ncu_summary.py)\n" + txt)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, u, v = rows[0], rows[1], rows[2]

    def val(name):
        i = h.index(name)
        x = float(v[i].replace(",", ""))
        return x * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(u[i], 1)

    traffic[key] = int(val("dram__bytes_read.sum") + val("dram__bytes_write.sum"))
    metrics[key] = {m.split(".")[0].replace("sm__inst_executed_", "").replace("smsp__", "").replace("sm__", ""): round(val(m), 3)
                    for m in ("gpu__time_duration.sum", "launch__registers_per_thread",
                              "sm__warps_active.avg.pct_of_peak_sustained_active",
                              "smsp__issue_active.avg.pct_of_peak_sustained_active",
                              "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
                              "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
                              "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
                              "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
                              "dram__throughput.avg.pct_of_peak_sustained_elapsed") if m in h}
    if "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum" in h:
        metrics[key]["smem_conflict_wavefront_frac"] = round(
            val("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum") / max(val("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"), 1), 4)
json.dump({
    "source": f"ncu --set full --clock-control none, one launch each of a 16384-pair device batch (32768 strand-pairs, "
              f"32768 signals: target + forward query per pair); capture tag {tag}; profiles/{RND}_ncu_*.txt",
    "dram_bytes_per_launch": traffic,
    "ncu_per_launch": metrics,
    "algorithmic_bytes_per_launch": {"scan_score": 32768 * 4224 + 2 * 581 * 16384,
                                     "xcorr_findtop": 16384 * 3 * 65536, "encode_fft": 16384 * 3 * 65536 + 32768 * 4096},
    "note": "three-channel form: a chunk pair owns three packed spectra of 64 KiB ((A + iC) of the target, (A + iC) of the "
            "query, (G_target + i G_query)); xcorr_findtop = xcorr_pair_kernel: one CTA per chunk pair reads them once and "
            "derives both strands; encode_fft writes them (+ planes) and reads the bases",
}, open(os.path.join(P, f"{RND}_traffic.json"), "w"), indent=2)
src = os.path.join(G, f"launches_{tag}.csv")
shutil.copy(src, os.path.join(P, f"{RND}_launches.csv"))
rows = [r for r in csv.reader(open(src)) if len(r) > 10 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    a = agg.setdefault(r[4].split("(")[0], [0, 0.0])
    a[0] += 1
    a[1] += float(r[-1]) / 1e6
tot = sum(x[1] for x in agg.values())
with open(os.path.join(P, f"{RND}_launch_summary.txt"), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 400  python bench.py --pairs 65536 --target-total "
            "4294967296 --steps 2 --warmup 3 --no-cpu-baseline --no-extras\n# per-launch times are cold-cache and serialised: compare SHARES with "
            "bench.py's live CUDA-event shares (roofline.kernels.*.share)\n")
    for k, x in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k:62s} launches={x[0]:4d} total_ms={x[1]:9.3f} share={x[1] / tot:5.3f} avg_ms={x[1] / x[0]:.3f}\n")
for b in sys.argv[2:]:
    shutil.copy(b, os.path.join(P, RND + "_" + os.path.basename(b)))
print(open(os.path.join(P, f"{RND}_launch_summary.txt")).read(), json.dumps(traffic))
