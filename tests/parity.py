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


def borderline_lags(oracle, xc, cutoff):
    """Lag indices whose correlation is within BORDER (relative) of the FindTop threshold."""
    _, env = oracle.findtop(xc, cutoff, with_env=True)
    N = xc.shape[0]
    thr = env[np.arange(N) // 256] * cutoff + 1.0 if N // 256 > 8 else np.ones(N)
    return set(np.nonzero(np.abs(xc.astype(np.float64) - thr) <= BORDER * np.abs(thr))[0].tolist())


def compare_candidates(oracle, got_idx, ref_xc, cutoff):
    """Candidate sets must agree except for borderline lags. Returns the listed borderline lags."""
    exp = set(oracle.findtop(ref_xc, cutoff).tolist())
    got = set(int(i) for i in got_idx)
    diff = exp ^ got
    border = borderline_lags(oracle, ref_xc, cutoff)
    unexplained = diff - border
    assert not unexplained, f"candidate lags differ away from the threshold: {sorted(unexplained)[:10]}"
    return sorted(diff)


def compare_pair_records(oracle, got, exp, tchunk, qchunk, t_start, q_start, q_seqsize, q_chunk_flag, N, cutoff,
                         min_prob, target_total, listed):
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
        border = borderline_lags(oracle, xc, cutoff)
        rec = g.get(k, e.get(k))
        near_prob = abs(float(rec["prob"]) - min_prob) <= 1e-4
        assert (lag in border) or near_prob, f"unexplained match difference {k} (lag {lag}, prob {rec['prob']})"
        listed.append(dict(key=k, lag=int(lag), where="gpu-only" if k in g else "oracle-only",
                           prob=float(rec["prob"]), reason="threshold" if lag in border else "min_prob"))
