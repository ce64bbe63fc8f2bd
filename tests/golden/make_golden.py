"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libsatsuma_ref.so).

Run in the development container only (needs /root/reference for the sample FASTA files and
the compiled reference):   python tests/golden/make_golden.py
The fixtures are small, committed, and are what pins the oracle and the CUDA path when the
reference itself is not available (e.g. on the GPU box).

samples.npz -- config 1 (samples/dog.X.part.fasta = target, samples/human.X.part.fasta = query,
               4096-bp chunks, target overlap 1024, -cutoff 1.8, slave semantics):
   a subset of the reference's own chunks + per-stage reference outputs for a few pairs +
   t_result records of whole blocks (HomologyByXCorr::align_target).
synthetic.npz -- IUPAC / N / gap / short-chunk edge cases on synthetic sequences.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
N = 8192


def pack_chunks(chunks):
    blob = b"".join(c[0] for c in chunks)
    lens = np.array([len(c[0]) for c in chunks], dtype=np.int32)
    return (np.frombuffer(blob, dtype=np.uint8).copy(), lens, np.array([c[1] for c in chunks], np.int32),
            np.array([c[2] for c in chunks], np.int32), np.array([c[3] for c in chunks], np.int32))


def stage_outputs(R, O, t, q, cutoff, target_total):
    out = {}
    for strand in (0, 1):
        qs = R.revcomp(q) if strand else q
        xc = R.xcorr(t, qs, N)
        cands = R.findtop(xc, cutoff)
        segs = R.matchup(qs, t, xc, cutoff)
        probs = np.zeros((len(segs), 2))
        for i, s in enumerate(segs):
            probs[i] = R.match_prob(t, qs, int(s["start_target"]), int(s["start_query"]), int(s["len"]), target_total)
        out[strand] = (xc, cands, segs, probs)
    return out


def make_samples():
    R = oracle.Reference()
    O = oracle.Oracle()
    R.configure()
    R.load_fasta("/root/reference/samples/dog.X.part.fasta", "/root/reference/samples/human.X.part.fasta")
    T, Q = R.chunks(True), R.chunks(False)
    total = R.target_total()
    # subset: first 8 and last 5 target chunks, first 8 and last 4 query chunks, plus a mid block
    tsel = list(range(0, 8)) + list(range(100, 104)) + list(range(256, 261))
    qsel = list(range(0, 8)) + list(range(60, 64)) + list(range(241, 245))
    tb, tl, ts, tid, tsz = pack_chunks([T[i] for i in tsel])
    qb, ql, qs, qid, qsz = pack_chunks([Q[i] for i in qsel])
    data = dict(target_total=total, n_target_chunks=len(T), n_query_chunks=len(Q),
                t_index=np.array(tsel, np.int32), q_index=np.array(qsel, np.int32),
                t_bases=tb, t_lens=tl, t_starts=ts, t_seq=tid, t_seqsize=tsz,
                q_bases=qb, q_lens=ql, q_starts=qs, q_seq=qid, q_seqsize=qsz)
    # per-stage taps on 4 pairs (indices into the subset)
    tap_pairs = [(3, 5), (0, 0), (16, 15), (9, 10)]
    data["tap_pairs"] = np.array(tap_pairs, np.int32)
    for k, (ti, qi) in enumerate(tap_pairs):
        t, q = T[tsel[ti]][0], Q[qsel[qi]][0]
        data[f"sig_t_{k}"] = R.signal(t, N)
        data[f"sig_q_{k}"] = R.signal(q, N)
        data[f"sig_qrc_{k}"] = R.signal(R.revcomp(q), N)
        so = stage_outputs(R, O, t, q, 1.8, total)
        for strand in (0, 1):
            xc, cands, segs, probs = so[strand]
            data[f"xc_{k}_{strand}"] = xc
            data[f"cand_{k}_{strand}"] = cands
            data[f"segs_{k}_{strand}"] = segs
            data[f"probs_{k}_{strand}"] = probs
    # whole blocks through align_target, on the subset re-injected as the chunk lists
    R.set_chunks(True, [T[i] for i in tsel], R.seq_sizes(True))
    R.set_chunks(False, [Q[i] for i in qsel], R.seq_sizes(False))
    blocks = [(0, 7, 0, 7, 0), (12, 16, 12, 15, 0), (8, 11, 8, 11, 0), (0, 7, 0, 7, 1)]
    data["blocks"] = np.array(blocks, np.int32)
    for k, b in enumerate(blocks):
        data[f"block_{k}"] = R.align_block(*b[:4], fast=bool(b[4]))
        print("block", b, len(data[f"block_{k}"]))
    # prob-table mode (-prob_table 1 -min_prob 0.9999), a small block
    R.configure(use_prob_table=True, min_prob_flag=0.9999)
    R.build_prob_table()
    data["block_table"] = np.array([0, 2, 0, 2, 0], np.int32)
    data["block_table_records"] = R.align_block(0, 2, 0, 2)
    print("table block", len(data["block_table_records"]))
    tab = R.prob_table(total, 0.9999)
    data["prob_table_rows"] = np.array([1, 64, 128, 255, 400, 511], np.int32)
    data["prob_table_vals"] = tab[[1, 64, 128, 255, 400, 511]]
    np.savez_compressed(os.path.join(HERE, "samples.npz"), **data)


def make_synthetic():
    R = oracle.Reference()
    O = oracle.Oracle()
    rng = np.random.default_rng(11)
    L = 4096

    def rnd(n, alphabet=b"ACGT"):
        return bytes(rng.choice(list(alphabet), n).astype(np.uint8))

    base = bytearray(rnd(L))
    cases_t, cases_q = [], []
    # 0: IUPAC-rich related pair
    t0 = bytearray(base)
    q0 = bytearray(base)
    for i in rng.choice(L, 600, replace=False):
        q0[i] = rng.choice(list(b"ACGT"))
    for i in rng.choice(L, 200, replace=False):
        t0[i] = rng.choice(list(b"KMRYSWBVHDN"))
    for i in rng.choice(L, 200, replace=False):
        q0[i] = rng.choice(list(b"KMRYSWBVHDNX"))
    cases_t.append(bytes(t0)); cases_q.append(bytes(q0))
    # 1: N runs + gaps + unknown letters
    t1 = bytearray(base); q1 = bytearray(base)
    t1[500:900] = b"N" * 400
    q1[2000:2100] = b"-" * 100
    for i in rng.choice(L, 50, replace=False):
        q1[i] = ord("Z")
    cases_t.append(bytes(t1)); cases_q.append(bytes(q1))
    # 2: short chunks (< 1024: entropy weight 1) with shared segment
    t2 = rnd(700); q2 = rnd(300) + t2[100:400] + rnd(177)
    cases_t.append(t2); cases_q.append(q2)
    # 3: reverse-strand homology, length not a multiple of the entropy window
    t3 = rnd(4090)
    q3 = R.revcomp(t3[1000:3000]) + rnd(1531)
    cases_t.append(t3); cases_q.append(q3)
    # 4: low-complexity / tandem repeats (many candidates)
    unit = rnd(7)
    t4 = (unit * 600)[:L]
    q4 = bytearray((unit * 600)[3:L + 3])
    for i in rng.choice(L, 300, replace=False):
        q4[i] = rng.choice(list(b"ACGT"))
    cases_t.append(bytes(t4)); cases_q.append(bytes(q4))
    # 5: empty target chunk (all-N chunks are emptied by ChunkManager) vs normal query
    cases_t.append(b""); cases_q.append(rnd(L))
    # 6: tiny chunks
    cases_t.append(rnd(50)); cases_q.append(rnd(47))
    # 7: identical chunks (perfect diagonal)
    cases_t.append(bytes(base)); cases_q.append(bytes(base))

    data = {"n_cases": len(cases_t)}
    total = 1.0e6
    R.configure()
    chunks_t = [(c, 1000 * i, i, len(c) + 1000 * i + 77) for i, c in enumerate(cases_t)]
    chunks_q = [(c, 500 * i, i, len(c) + 500 * i + 33) for i, c in enumerate(cases_q)]
    R.set_chunks(True, chunks_t, [c[3] for c in chunks_t])
    R.set_chunks(False, chunks_q, [c[3] for c in chunks_q])
    R.lib.ref_set_target_total(total)
    for i, (t, q) in enumerate(zip(cases_t, cases_q)):
        data[f"t_{i}"] = np.frombuffer(t, dtype=np.uint8)
        data[f"q_{i}"] = np.frombuffer(q, dtype=np.uint8)
        data[f"sig_t_{i}"] = R.signal(t, N)
        data[f"sig_q_{i}"] = R.signal(q, N)
        data[f"sig_qrc_{i}"] = R.signal(R.revcomp(q), N)
        so = stage_outputs(R, O, t, q, 1.8, total)
        for strand in (0, 1):
            xc, cands, segs, probs = so[strand]
            data[f"xc_{i}_{strand}"] = xc
            data[f"cand_{i}_{strand}"] = cands
            data[f"segs_{i}_{strand}"] = segs
            data[f"probs_{i}_{strand}"] = probs
        data[f"records_{i}"] = R.align_block(i, i, i, i)
        print("case", i, len(t), len(q), "cands", len(so[0][1]), len(so[1][1]), "segs", len(so[0][2]), len(so[1][2]),
              "records", len(data[f"records_{i}"]))
    data["t_starts"] = np.array([c[1] for c in chunks_t], np.int32)
    data["t_seqsize"] = np.array([c[3] for c in chunks_t], np.int32)
    data["q_starts"] = np.array([c[1] for c in chunks_q], np.int32)
    data["q_seqsize"] = np.array([c[3] for c in chunks_q], np.int32)
    data["target_total"] = total
    np.savez_compressed(os.path.join(HERE, "synthetic.npz"), **data)


if __name__ == "__main__":
    make_samples()
    make_synthetic()
    for f in ("samples.npz", "synthetic.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")
