"""Generates tests/golden/large_n.npz from the UNMODIFIED reference (oracle/_ref/libsatsuma_ref.so):
config 3 of BASELINE.json -- 8192-bp chunks (N = 16384) and 16384-bp chunks (N = 32768).

    python tests/golden/make_golden_large.py      (development container only)

Per transform size: one synthetic chunk pair (satsuma2_b200.synth.random_pairs, seed = chunk size, pair 0),
the reference's correlation vector and FindTop list for both strands and the t_result records of
HomologyByXCorr::align_target for that pair.  At these sizes the reference's own float FFT drifts
from the exact transform (float rotation recurrence for the twiddles of passes > 12):
measured here 1.6e-5 (N = 16384) and 2.0e-4 (N = 32768) of max|xc|.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from satsuma2_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    R = oracle.Reference()
    data = {}
    for chunk in (8192, 16384):
        N = 2 * chunk
        T, Q, _ = synth.random_pairs(1, chunk, seed=chunk)
        t, q = T[0].tobytes(), Q[0].tobytes()
        data[f"t_{N}"] = T[0].copy()
        data[f"q_{N}"] = Q[0].copy()
        for strand in (0, 1):
            qs = R.revcomp(q) if strand else q
            xc = R.xcorr(t, qs, N)
            data[f"xc_{N}_{strand}"] = xc.astype(np.float32)
            data[f"cands_{N}_{strand}"] = R.findtop(xc, 1.8).astype(np.int32)
        R.configure(t_chunk=chunk, q_chunk=chunk)
        R.set_chunks(True, [(t, 0, 0, chunk)], [chunk])
        R.set_chunks(False, [(q, 0, 0, chunk)], [chunk])
        R.lib.ref_set_target_total(1e6)
        data[f"records_{N}"] = R.align_block(0, 0, 0, 0)
        print(N, "candidates", len(data[f"cands_{N}_0"]), len(data[f"cands_{N}_1"]), "records", len(data[f"records_{N}"]))
    np.savez_compressed(os.path.join(HERE, "large_n.npz"), **data)


if __name__ == "__main__":
    main()
