"""Generates tests/golden/samples_full.npz: the WHOLE of config 1 (BASELINE.json configs[0]).

Run in the development container only (needs /root/reference for the sample FASTA files and the
compiled reference):   python tests/golden/make_samples_full.py

Holds the reference's own chunking of samples/dog.X.part.fasta (target, 261 chunks, overlap 1024) and
samples/human.X.part.fasta (query, 245 chunks) as the two upper-cased sequences plus the chunk
geometry, and a digest of what the unmodified reference (oracle/_ref/libsatsuma_ref.so,
HomologyByXCorr::align_target, slave semantics) emits for all 261 x 245 chunk pairs: the record
count and a SHA-256 over the sorted record keys + ident bits.  The GPU test runs the reference live
on the GPU box (the .so travels there) and checks it against this digest before comparing the CUDA
path with it -- so the comparison is pinned to what the reference produced HERE.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
STRIPE_Q = (100, 106)


def record_digest(recs: np.ndarray) -> str:
    """SHA-256 over the records sorted by key, prob excluded (libm-dependent in the last ulp)."""
    keys = np.zeros(len(recs), dtype=[("query_id", "<u8"), ("target_id", "<u8"), ("query_size", "<u8"),
                                      ("qstart", "<u8"), ("tstart", "<u8"), ("len", "<u8"), ("reverse", "u1"),
                                      ("ident", "<f8")])
    for f in keys.dtype.names:
        keys[f] = recs[f]
    keys = np.sort(keys, order=list(keys.dtype.names))
    return hashlib.sha256(keys.tobytes()).hexdigest()


def main():
    R = oracle.Reference()
    R.configure()
    R.load_fasta("/root/reference/samples/dog.X.part.fasta", "/root/reference/samples/human.X.part.fasta")
    T, Q = R.chunks(True), R.chunks(False)
    total = R.target_total()

    def rebuild(chunks):  # the sequence from its chunks (they tile it, overlapping or not)
        size = chunks[0][3]
        seq = np.zeros(size, np.uint8)
        for b, start, sid, ssz in chunks:
            assert sid == 0 and ssz == size
            seq[start:start + len(b)] = np.frombuffer(b, np.uint8)
        return seq

    tseq, qseq = rebuild(T), rebuild(Q)
    for chunks, seq in ((T, tseq), (Q, qseq)):
        for b, start, _, _ in chunks:
            assert seq[start:start + len(b)].tobytes() == b
    tp = np.array([[t, t, q, q, 0] for q in range(len(Q)) for t in range(len(T))], dtype=np.int32)
    recs, secs, n = R.align_pairs_mt(tp, os.cpu_count() or 1)
    assert n == len(recs)
    print(f"{len(tp)} chunk pairs, {len(recs)} records, {secs:.1f} s on {os.cpu_count()} threads")
    # a stripe of the same grid (every target chunk x query chunks STRIPE_Q) for the CPU suite: the C restatement
    # is checked against this digest in seconds
    sp = np.array([[t, t, q, q, 0] for q in range(*STRIPE_Q) for t in range(len(T))], dtype=np.int32)
    srecs, _, sn = R.align_pairs_mt(sp, os.cpu_count() or 1)
    assert sn == len(srecs)
    np.savez_compressed(
        os.path.join(HERE, "samples_full.npz"), target_total=total, t_seq=tseq, q_seq=qseq,
        t_starts=np.array([c[1] for c in T], np.int32), t_lens=np.array([len(c[0]) for c in T], np.int32),
        q_starts=np.array([c[1] for c in Q], np.int32), q_lens=np.array([len(c[0]) for c in Q], np.int32),
        n_records=np.int64(len(recs)), n_reverse=np.int64(int(recs["reverse"].sum())),
        digest=np.array(record_digest(recs)), stripe_q=np.array(STRIPE_Q, np.int32), stripe_n=np.int64(len(srecs)),
        stripe_digest=np.array(record_digest(srecs)))


if __name__ == "__main__":
    main()
