#!/usr/bin/env python
"""KMatch seeding on a synthetic genome pair -- this repository's GPU program beside the
reference's own KMatch (oracle/_ref/KMatch_ref, std::sort on one thread per genome), same files, same arguments,
same records.  Run on the GPU box:  python tools/kmatch_timing.py [megabases]  -> one JSON line."""
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
from satsuma2_b200 import build as sxbuild, synth  # noqa: E402
from test_kmatch import DT, _fa, _keys  # noqa: E402


def main():
    mb = float(sys.argv[1]) if len(sys.argv) > 1 else 20.0
    L = int(mb * 1e6)
    tgt, qry = synth.genome_pair(L, seed=5, divergence=0.03)
    qry = np.concatenate([np.frombuffer(b"ACGTTGCA" * 40, np.uint8), qry])  # shift the query coordinates
    tmp = tempfile.mkdtemp()
    q, t = os.path.join(tmp, "q.fa"), os.path.join(tmp, "t.fa")
    _fa(q, [("q", qry)])
    _fa(t, [("t", tgt)])
    mine = [e for e in sxbuild.build_host() if e.endswith("KMatch")][0]
    K = 31
    res = {}
    for name, exe in (("b200", mine), ("reference", oracle.REF_KMATCH)):
        out = os.path.join(tmp, name + ".k")
        t0 = time.perf_counter()
        r = subprocess.run([exe, q, t, str(K), out, str(K), str(K - 1), "1"], check=True, capture_output=True, text=True)
        res[name] = (time.perf_counter() - t0, np.fromfile(out, dtype=DT))
        if name == "b200":
            dev_ms = float(r.stdout.split("device work (upload, kernels, download):")[1].split()[0])
    same = _keys(res["b200"][1]) == _keys(res["reference"][1])
    print(json.dumps({"what": f"KMatch K={K} on a synthetic {mb:g} Mb x {mb:g} Mb genome pair (3 % divergence), wall clock of "
                              "the whole program incl. reading the FASTA files", "records": int(len(res["reference"][1])),
                      "identical_record_sets": bool(same), "reference_s": res["reference"][0], "b200_s": res["b200"][0], "b200_device_work_s": dev_ms / 1e3,
                      "speedup": res["reference"][0] / res["b200"][0], "host_cores": os.cpu_count()}))


if __name__ == "__main__":
    main()
