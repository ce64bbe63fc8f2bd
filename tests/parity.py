"""Shared comparison helpers for the parity tests (CUDA path vs oracle / reference).

Parity contract (BASELINE.json north_star):
  * correlation values within XC_TOL = 1e-4 of max|xc| of that strand-pair;
  * emitted matches: bit-identical set (coordinates, strand, length) except matches that hinge on
    a candidate lag whose correlation lies within BORDER = 1e-4 (relative) of the FindTop
    threshold, or whose probability lies within 1e-4 of min_prob -- those are LISTED, not hidden;
  * prob within PROB_TOL (CUDA erf/exp vs glibc: a few ulp, amplified by exp(-cdf*T)), ident exact.
"""
from __future__ import annotations

import numpy as np

XC_TOL = 1e-4
BORDER = 1e-4
PROB_TOL = 1e-6


def chunk_list(bases, lens, starts, seq, seqsize):
    offs = np.zeros(len(lens), dtype=np.int64)
    if len(lens) > 1:
        offs[1:] = np.cumsum(lens[:-1])
    return [(bytes(bases[offs[i]:offs[i] + lens[i]]), int(starts[i]), int(seq[i]), int(seqsize[i]))
            for i in range(len(lens))]


def rec_key(r):
    return (int(r["query_id"]), int(r["target_id"]), int(r["query_size"]), int(r["qstart"]), int(r["tstart"]),
            int(r["len"]), int(r["reverse"]))


def xc_rel_err(got, ref):
    scale = float(np.abs(ref).max())
    if scale == 0.0:
        return float(np.abs(got).max())
    return float(np.abs(got.astype(np.float64) - ref.astype(np.float64)).max() / scale)


def borderline_lags(oracle, xc, cutoff, border=BORDER):
    """Lag indices whose correlation is within `border` (relative; the contract's 1e-4 unless a test documents a wider
    allowance, as for N = 32768 against the reference's own drifting FFT) of the FindTop threshold."""
    _, env = oracle.findtop(xc, cutoff, with_env=True)
    N = xc.shape[0]
    thr = env[np.arange(N) // 256] * cutoff + 1.0 if N // 256 > 8 else np.ones(N)
    return set(np.nonzero(np.abs(xc.astype(np.float64) - thr) <= border * np.abs(thr))[0].tolist())


def compare_candidates(oracle, got_idx, ref_xc, cutoff, border=BORDER):
    """Candidate sets must agree except for borderline lags. Returns the listed borderline lags."""
    exp = set(oracle.findtop(ref_xc, cutoff).tolist())
    got = set(int(i) for i in got_idx)
    diff = exp ^ got
    unexplained = diff - borderline_lags(oracle, ref_xc, cutoff, border)
    assert not unexplained, f"candidate lags differ away from the threshold: {sorted(unexplained)[:10]}"
    return sorted(diff)


def compare_pair_records(oracle, got, exp, tchunk, qchunk, t_start, q_start, q_seqsize, q_chunk_flag, N, cutoff,
                         min_prob, target_total, listed, border_tol=BORDER):
    """Records of ONE chunk pair (both strands).  got/exp: structured arrays (t_result layout).
    Differences must be explained by a borderline candidate lag or borderline probability."""
    g = {rec_key(r): r for r in got}
    e = {rec_key(r): r for r in exp}
    for k in g.keys() & e.keys():
        assert g[k]["ident"] == e[k]["ident"], (k, g[k]["ident"], e[k]["ident"])
        pe, pg = float(e[k]["prob"]), float(g[k]["prob"])
        assert pg == pe or abs(pg - pe) <= PROB_TOL * max(abs(pe), 1e-300) or (np.isnan(pg) and np.isnan(pe)), (k, pg, pe)
    diff = set(g.keys()) ^ set(e.keys())
    if not diff:
        return
    # explain: recompute the oracle's correlation for the strand and look at the lag of each differing record
    tb, qb = tchunk, qchunk
    for k in sorted(diff):
        reverse = k[6]
        qs = oracle.revcomp(qb) if reverse else qb
        xc = oracle.xcorr(tb, qs, N)
        start_t = k[4] - t_start
        if not reverse:
            start_q = k[3] - q_start
        else:
            # invert RCQuery: qStart = startQ + size - start - q_chunk  (int wrapped into u64)
            qstart = k[3] if k[3] < (1 << 63) else k[3] - (1 << 64)
            start_q = qstart - q_seqsize + q_start + q_chunk_flag
        lag = start_q - start_t + N // 2
        border = borderline_lags(oracle, xc, cutoff, border_tol)
        rec = g.get(k, e.get(k))
        near_prob = abs(float(rec["prob"]) - min_prob) <= 1e-4
        assert (lag in border) or near_prob, f"unexplained match difference {k} (lag {lag}, prob {rec['prob']})"
        listed.append(dict(key=k, lag=int(lag), where="gpu-only" if k in g else "oracle-only",
                           prob=float(rec["prob"]), reason="threshold" if lag in border else "min_prob"))


def candidate_pairs(key, T, Q, q_chunk_flag):
    """Chunk pairs (t, q) a record key can have come from.  T / Q: lists of (bases, start, seq_id, seq_size).
    Target chunks overlap, so up to two targets qualify; the reverse-strand query coordinate is inverted
    with the slave's RCQuery formula (Slave.cc:56-60, 180)."""
    qid, tid, qsize, qstart, tstart, _, reverse = key
    ts = [i for i, c in enumerate(T) if c[2] == tid and c[1] <= tstart < c[1] + len(c[0])]
    qs = []
    for j, c in enumerate(Q):
        if c[2] != qid:
            continue
        if not reverse:
            sq = qstart - c[1]
        else:
            v = qstart if qstart < (1 << 63) else qstart - (1 << 64)
            sq = v - qsize + c[1] + q_chunk_flag
        if 0 <= sq < len(c[0]):
            qs.append(j)
    return [(t, q) for t in ts for q in qs]


def explain_set_difference(oracle, got, exp, T, Q, gpu_pair, exp_pair, q_chunk_flag, N, cutoff, min_prob, target_total,
                           listed, max_pairs=64):
    """Whole record sets of many chunk pairs (a block, a grid, a genome).  Identical sets pass at once;
    otherwise every record of the symmetric difference is traced back to the chunk pair(s) it can have come
    from, those pairs are re-run one by one on both sides (gpu_pair(t, q) / exp_pair(t, q) -> records) and
    compare_pair_records must explain each difference by a borderline candidate lag or probability (it appends
    to `listed`, it asserts otherwise).  A differing record that no re-run pair reproduces is unexplained."""
    gk = {rec_key(r) for r in got}
    ek = {rec_key(r) for r in exp}
    diff = gk ^ ek
    if not diff:
        return 0
    pairs = []
    for k in sorted(diff):
        cp = candidate_pairs(k, T, Q, q_chunk_flag)
        assert cp, f"record {k} maps to no chunk pair"
        for p in cp:
            if p not in pairs:
                pairs.append(p)
    assert len(pairs) <= max_pairs, f"{len(diff)} differing records over {len(pairs)} chunk pairs: not borderline noise"
    before = len(listed)
    for (t, q) in pairs:
        gp, ep = gpu_pair(t, q), exp_pair(t, q)
        compare_pair_records(oracle, gp, ep, T[t][0], Q[q][0], T[t][1], Q[q][1], Q[q][3], q_chunk_flag, N, cutoff,
                             min_prob, target_total, listed)
    accounted = {tuple(item["key"]) for item in listed[before:]}
    missing = diff - accounted
    assert not missing, f"differences not reproduced pair by pair: {sorted(missing)[:5]}"
    return len(diff)
