"""C++ host side above the C ABI (satsuma2_b200/host): FASTA loading, chunk arithmetic, the binary
match file, and the two drop-in executables (standalone HomologyByXCorr, TCP HomologyByXCorrSlave).
CPU tests cover parsing / chunking; GPU tests run the executables end to end."""
import os
import socket
import struct
import subprocess
import threading
import time

import numpy as np
import pytest

from parity import rec_key

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host_bins(sx):
    from satsuma2_b200 import build as sxbuild

    exes = sxbuild.build_host()
    return {os.path.basename(e): e for e in exes}


def _write_fasta(path, records, width=60, lower=False):
    with open(path, "w") as f:
        for name, seq in records:
            f.write(">" + name + "\n")
            s = seq.decode() if isinstance(seq, bytes) else seq
            if lower:
                s = s.lower()
            for i in range(0, len(s), width):
                f.write(s[i:i + width] + "\n")


def _dump_chunks(exe, q, t, *extra):
    out = subprocess.run([exe, "-q", q, "-t", t, "-dump_chunks", "1", *extra], capture_output=True, text=True, check=True)
    T, Q = [], []
    for ln in out.stdout.splitlines():
        p = ln.split()
        if len(p) == 5 and p[0] in ("T", "Q"):
            rec = (int(p[1]), int(p[2][4:]), int(p[3][6:]), int(p[4][4:]))
            (T if p[0] == "T" else Q).append(rec)
    return T, Q


def test_fasta_and_chunking(host_bins, tmp_path):
    """Header tokens joined by '_', only the first token of sequence lines, soft-masking ignored,
    all-N chunks emptied, < 6 bp sequences skipped, trailing empty chunk when l % stride == 0."""
    rng = np.random.default_rng(1)
    s1 = bytes(rng.choice(list(b"ACGT"), 10000).astype(np.uint8))
    s2 = b"N" * 4096 + bytes(rng.choice(list(b"ACGT"), 4096).astype(np.uint8))
    q = tmp_path / "q.fa"
    t = tmp_path / "t.fa"
    with open(q, "w") as f:
        f.write(">chr1 some description here\n")
        for i in range(0, len(s1), 70):
            f.write(s1[i:i + 70].decode().lower() + "  trailing tokens are ignored\n")
        f.write("\n>tiny\nACGT\n>chr2\n" + s2.decode() + "\n")
    _write_fasta(t, [("t1", s1[:8192])])
    T, Q = _dump_chunks(host_bins["HomologyByXCorr"], str(q), str(t))
    # query: chr1 (10000 bp -> 3 chunks), tiny (< 6 bp: none), chr2 (8192 bp -> 3 chunks, first all-N, last empty)
    assert [c[3] for c in Q] == [4096, 4096, 1808, 0, 4096, 0]
    assert [c[1] for c in Q] == [0, 0, 0, 2, 2, 2]
    assert [c[2] for c in Q] == [0, 4096, 8192, 0, 4096, 8192]
    # target overlap t_chunk/2 in the standalone tool: stride 2048, 1 + 8192/2048 = 5 chunks
    assert [(c[2], c[3]) for c in T] == [(0, 4096), (2048, 4096), (4096, 4096), (6144, 2048), (8192, 0)]
    # -nblocks/-block: only the block's chunks keep their bases
    T2, _ = _dump_chunks(host_bins["HomologyByXCorr"], str(q), str(t), "-nblocks", "2", "-block", "1")
    assert [c[3] for c in T2] == [0, 0, 0, 2048, 0]


def test_fastq_input(host_bins, tmp_path):
    """A file whose first record line starts with '@' is read as FASTQ (vecDNAVector::ReadQ,
    analysis/DNAVector.cc:1150-1173, 1223-1228): four lines per record, bases = the whole second line."""
    rng = np.random.default_rng(2)
    r1 = bytes(rng.choice(list(b"ACGT"), 5000).astype(np.uint8)).decode()
    r2 = bytes(rng.choice(list(b"ACGT"), 100).astype(np.uint8)).decode()
    q = tmp_path / "q.fastq"
    with open(q, "w") as f:
        f.write(f"@read1 some text\n{r1}\n+\n{'I' * 5000}\n\n@read2\n{r2}\n+\n{'I' * 100}\n")
    t = tmp_path / "t.fa"
    _write_fasta(t, [("t1", r1.encode())])
    T, Q = _dump_chunks(host_bins["HomologyByXCorr"], str(q), str(t))
    assert [(c[1], c[2], c[3]) for c in Q] == [(0, 0, 4096), (0, 4096, 904), (1, 0, 100)]
    assert [(c[2], c[3]) for c in T] == [(0, 4096), (2048, 2952), (4096, 904)]


def _equal_length_case(tmp_path, n_t=24, n_q=20, L=2000, seed=12):
    """Many sequences of ONE length below the chunk stride: every sequence is one chunk and all chunks are equally
    long, so the reference tool's re-used CCSignal objects never carry stale samples and its output is
    a valid parity target.  Queries are diverged copies of the targets, every third one reverse-complemented,
    every fourth unrelated."""
    rng = np.random.default_rng(seed)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    comp = np.zeros(256, np.uint8)
    comp[list(b"ACGT")] = list(b"TGCA")
    T = [rng.choice(acgt, L) for _ in range(n_t)]
    Q = []
    for i in range(n_q):
        q = T[i].copy()
        mut = rng.random(L) < rng.uniform(0.05, 0.15)
        q[mut] = rng.choice(acgt, int(mut.sum()))
        if i % 3 == 1:
            q = comp[q[::-1]]
        if i % 4 == 3:
            q = rng.choice(acgt, L)
        Q.append(q)
    t, q = tmp_path / "t.fa", tmp_path / "q.fa"
    _write_fasta(t, [(f"t{i}", T[i].tobytes()) for i in range(n_t)])
    _write_fasta(q, [(f"q{i}", Q[i].tobytes()) for i in range(n_q)])
    return T, Q, str(t), str(q)


def _rows(arr):
    return sorted(tuple(float(x) for x in row) for row in arr)


def test_oracle_matches_the_reference_standalone_tool(oracle_lib, reference_lib, tmp_path):
    """The UNMODIFIED reference executable (tools/analysis/HomologyByXCorr, built by `make -C oracle reftool`) run
    end to end on a FASTA pair: its match file equals what the oracle predicts under the standalone semantics
    (-min_prob applied, RC coordinate from the real chunk length, target_total = sum of target lengths)."""
    import oracle

    if not os.path.exists(oracle.REF_TOOL):
        pytest.skip("oracle/_ref/HomologyByXCorr_ref not built (needs /root/reference)")
    L = 2000
    T, Q, t, q = _equal_length_case(tmp_path, L=L)
    out = tmp_path / "ref.match"
    subprocess.run([oracle.REF_TOOL, "-q", q, "-t", t, "-o", str(out)], check=True, capture_output=True)
    arr, nt, nq = reference_lib.read_match_file(str(out))
    assert (nt, nq) == (len(T), len(Q)) and len(arr) >= 15 and arr[:, 6].sum() >= 3
    tl = [(T[i].tobytes(), 0, i, L) for i in range(len(T))]
    ql = [(Q[i].tobytes(), 0, i, L) for i in range(len(Q))]
    params = oracle_lib.make_params(min_prob=0.9999, target_total=float(len(T) * L))
    exp = oracle_lib.align_pairs(params, tl, ql, [(a, b) for a in range(len(T)) for b in range(len(Q))],
                                 threads=os.cpu_count() or 1)
    rows = []
    for rr in exp:
        k = list(rec_key(rr))
        if k[6]:  # standalone RC coordinate uses the real chunk length (tools/...:173,799)
            qs = k[3] if k[3] < (1 << 63) else k[3] - (1 << 64)
            k[3] = qs + 4096 - L
        rows.append((k[1], k[0], k[2], k[4], k[3], k[5], k[6], float(rr["ident"]) * k[5], float(rr["prob"]), float(rr["ident"])))
    assert _rows(arr) == _rows(rows)


def _write_match_file(path, recs, n_t, n_q, size=1000000):
    """Version-3 MultiMatches file (analysis/SequenceMatch.cc:320-360) from n x 10 records."""
    with open(path, "wb") as f:
        f.write(struct.pack("<ii", 3, n_t))
        for i in range(n_t):
            name = f"t{i}".encode() + b"\0"
            f.write(struct.pack("<q", len(name)) + name)
        f.write(struct.pack("<i", n_q))
        for i in range(n_q):
            name = f"q{i}".encode() + b"\0"
            f.write(struct.pack("<q", len(name)) + name)
        f.write(struct.pack("<i", len(recs)))
        for r in recs:
            f.write(struct.pack("<iiiiiiiddd", int(r[0]), int(r[1]), int(r[2]), int(r[3]), int(r[4]), int(r[5]), int(r[6]),
                                float(r[7]), float(r[8]), float(r[9])))
        f.write(struct.pack(f"<{n_t}i", *([size] * n_t)))
        f.write(struct.pack(f"<{n_q}i", *([size] * n_q)))


def test_sort_and_collapse_match_the_reference(host_bins, reference_lib, tmp_path):
    """MultiMatches::Sort / Collapse (analysis/SequenceMatch.h:211-215, SequenceMatch.cc:418-469)
    as the master applies them to the slaves' records every cycle -- same order, same fused matches, quirks
    included (query id ignored when fusing, last group dropped), against the compiled reference."""
    rng = np.random.default_rng(4)
    n = 5000
    recs = np.zeros((n, 10))
    recs[:, 0] = rng.integers(0, 3, n)            # target id
    recs[:, 1] = rng.integers(0, 3, n)            # query id
    recs[:, 2] = 1000000
    recs[:, 3] = rng.integers(0, 20000, n)        # start in target
    recs[:, 4] = recs[:, 3] + rng.integers(-40, 40, n)  # start in query: near the diagonal
    recs[:, 5] = rng.integers(46, 400, n)
    recs[:, 6] = rng.integers(0, 2, n)
    recs[:, 9] = rng.uniform(0.5, 1.0, n)
    recs[:, 7] = recs[:, 9] * recs[:, 5]
    recs[:, 8] = rng.uniform(0.99, 1.0, n)
    # duplicates as overlapping target chunks produce them, and near-duplicates 1-3 bases apart
    dup = recs[rng.integers(0, n, 1500)].copy()
    dup[:, 3] += rng.integers(0, 4, len(dup))
    dup[:, 4] += rng.integers(-3, 4, len(dup))
    dup[:, 5] += rng.integers(-10, 30, len(dup))
    recs = np.concatenate([recs, dup])
    rng.shuffle(recs)
    src = tmp_path / "in.match"
    _write_match_file(src, recs, 3, 3)
    for collapse in (0, 1):
        dst = tmp_path / f"out{collapse}.match"
        subprocess.run([host_bins["XCorrMatchTool"], "-i", str(src), "-o", str(dst), "-sort", "1", "-collapse", str(collapse)],
                       check=True, capture_output=True)
        got, nt, nq = reference_lib.read_match_file(str(dst))
        exp = reference_lib.sort_collapse(recs, bool(collapse))
        assert (nt, nq) == (3, 3)
        assert got.shape == exp.shape and len(exp) > 1000
        assert np.array_equal(got, exp)
    assert len(exp) < len(recs) - 500  # collapse fused the planted duplicates


def test_chain_matches_the_reference(host_bins, reference_lib, tmp_path):
    """SURVEY 8f rank 3, second half: ChainMatches = Sort, Collapse, RunMatchDynProg (tools/analysis/ChainMatches.cc:
    60-75, analysis/MatchDynProg.cc:199-243, 401-561) on a synthetic synteny: true matches along a diagonal with
    an inversion and indel drift, noise matches off the diagonal, and a repeat that piles matches onto one place."""
    rng = np.random.default_rng(6)
    size = 400000
    recs = []
    for tid, qid in ((0, 0), (1, 1), (1, 0)):
        pos_t, drift = 500, 0
        while pos_t < size - 2000:
            ln = int(rng.integers(46, 600))
            rc = 150000 < pos_t < 200000 and tid == 0
            sq = (size - (pos_t + drift) - ln) if rc else pos_t + drift
            ident = float(rng.uniform(0.6, 0.98))
            if 0 < sq < size - ln:
                recs.append([tid, qid, size, pos_t, sq, ln, int(rc), ident * ln, float(rng.uniform(0.99, 1.0)), ident])
            pos_t += ln + int(rng.integers(50, 3000))
            drift += int(rng.integers(-20, 21))
    for _ in range(3000):  # noise
        ln = int(rng.integers(46, 120))
        ident = float(rng.uniform(0.5, 0.8))
        recs.append([int(rng.integers(0, 2)), int(rng.integers(0, 2)), size, int(rng.integers(1, size - 200)),
                     int(rng.integers(1, size - 200)), ln, int(rng.integers(0, 2)), ident * ln, float(rng.uniform(0.99, 1.0)),
                     ident])
    for _ in range(400):  # a repeat: many query places hit the same target place
        ln = int(rng.integers(100, 300))
        recs.append([0, 0, size, 300000 + int(rng.integers(0, 50)), int(rng.integers(1, size - 400)), ln, 0, 0.8 * ln,
                     0.995, 0.8])
    recs = np.array(recs, dtype=np.float64)
    rng.shuffle(recs)
    src, dst = tmp_path / "in.match", tmp_path / "chained.match"
    _write_match_file(src, recs, 2, 2, size)
    for dups in (0, 1):  # -dups 1 = RunMatchDynProgMult (config 5's master-side flag)
        subprocess.run([host_bins["XCorrMatchTool"], "-i", str(src), "-o", str(dst), "-sort", "1", "-collapse", "1",
                        "-chain", "1", "-dups", str(dups)], check=True, capture_output=True)
        got, nt, nq = reference_lib.read_match_file(str(dst))
        exp = reference_lib.chain(reference_lib.sort_collapse(recs, True), [size, size], [size, size], bool(dups))
        assert got.shape == exp.shape and 200 < len(exp) < len(recs) // 4
        assert np.array_equal(got, exp)
        if not dups:
            on_diag = np.abs(exp[:, 3] - exp[:, 4]) < 2000
            assert on_diag[exp[:, 6] == 0].mean() > 0.9  # the chain follows the planted synteny
            n_single = len(exp)
    assert len(exp) > n_single  # the second pass added the secondary (duplicated) alignments


def test_sample_chunk_counts(host_bins):
    ref = "/root/reference/samples"
    if not os.path.isdir(ref):
        pytest.skip("reference samples not present")
    T, Q = _dump_chunks(host_bins["HomologyByXCorr"], f"{ref}/human.X.part.fasta", f"{ref}/dog.X.part.fasta")
    assert len(Q) == 245 and Q[-1][3] == 577
    assert len(T) == 1 + 800001 // 2048 and T[-1][3] == 800001 - 390 * 2048


def _parse_match_file(path):
    b = open(path, "rb").read()
    off = 0

    def rd(fmt):
        nonlocal off
        v = struct.unpack_from("<" + fmt, b, off)
        off += struct.calcsize("<" + fmt)
        return v

    ver, nt = rd("ii")
    assert ver == 3
    tn = []
    for _ in range(nt):
        (ln,) = rd("q")
        tn.append(b[off:off + ln - 1].decode())
        off += ln
    (nq,) = rd("i")
    qn = []
    for _ in range(nq):
        (ln,) = rd("q")
        qn.append(b[off:off + ln - 1].decode())
        off += ln
    (n,) = rd("i")
    recs = [rd("iiiiiiiddd") for _ in range(n)]
    ts = rd(f"{nt}i")
    qs = rd(f"{nq}i")
    assert off == len(b)
    return tn, qn, recs, ts, qs


@pytest.mark.gpu
def test_standalone_tool_writes_reference_compatible_match_file(host_bins, sx, oracle_lib, tmp_path):
    """Runs the B200 HomologyByXCorr on a small FASTA pair; the v3 match file is parsed here, read
    back by the reference's own MultiMatches::Read, and its records equal the oracle under the
    standalone tool's semantics (target overlap size/2, -min_prob applied, RC coordinate by chunk length)."""
    rng = np.random.default_rng(3)
    base = rng.choice(list(b"ACGT"), 14000).astype(np.uint8)
    tgt = base.copy()
    qry = base.copy()
    mut = rng.random(14000) < 0.12
    qry[mut] = rng.choice(list(b"ACGT"), int(mut.sum()))
    comp = np.zeros(256, np.uint8)
    comp[list(b"ACGT")] = list(b"TGCA")
    qry[6000:9000] = comp[qry[6000:9000][::-1]]  # an inverted block -> reverse-strand matches
    q, t, o = tmp_path / "q.fa", tmp_path / "t.fa", tmp_path / "out.match"
    _write_fasta(q, [("qseq extra", qry.tobytes())], lower=True)
    _write_fasta(t, [("tseq", tgt.tobytes())])
    r = subprocess.run([host_bins["HomologyByXCorr"], "-q", str(q), "-t", str(t), "-o", str(o)], capture_output=True,
                       text=True)
    assert r.returncode == 0, r.stderr
    tn, qn, recs, ts, qs = _parse_match_file(str(o))
    assert tn == ["tseq"] and qn == ["qseq_extra"] and ts == (14000,) and qs == (14000,)
    assert len(recs) > 3 and any(rc[6] for rc in recs) and any(not rc[6] for rc in recs)
    for rc in recs:
        assert abs(rc[7] - rc[9] * rc[5]) < 1e-9 and rc[8] >= 0.9999
    # expected set from the oracle with the same chunking and semantics
    from satsuma2_b200 import synth

    to, tl, tst = synth.chunk_sequence(tgt, 4096, 2048)
    qo, ql, qst = synth.chunk_sequence(qry, 4096, 0)
    T = [(tgt[a:a + n].tobytes(), int(s), 0, 14000) for a, n, s in zip(to, tl, tst)]
    Q = [(qry[a:a + n].tobytes(), int(s), 0, 14000) for a, n, s in zip(qo, ql, qst)]
    params = oracle_lib.make_params(min_prob=0.9999, target_total=14000.0)
    exp = set()
    for ti, tc in enumerate(T):
        for qi, qc in enumerate(Q):
            if not tc[0] or not qc[0]:
                continue
            for rr in oracle_lib.align_pairs(params, T, Q, [(ti, qi)]):
                k = list(rec_key(rr))
                if k[6]:  # standalone RC coordinate uses the real chunk length (tools/...:173,799)
                    qstart = k[3] if k[3] < (1 << 63) else k[3] - (1 << 64)
                    k[3] = qstart + 4096 - len(qc[0])
                exp.add((k[1], k[0], k[2], k[4], k[3], k[5], k[6]))
    got = set((rc[0], rc[1], rc[2], rc[3], rc[4], rc[5], rc[6]) for rc in recs)
    assert got == exp
    # the reference's own reader accepts the file
    import oracle

    if oracle.have_reference():
        arr, nt, nq = oracle.Reference().read_match_file(str(o))
        assert (nt, nq) == (1, 1) and len(arr) == len(recs)
        assert np.allclose(arr, np.array(recs, dtype=np.float64), rtol=0, atol=0)


@pytest.mark.gpu
def test_standalone_tool_equals_the_reference_tool(host_bins, sx, reference_lib, tmp_path):
    """Executable against executable: the B200 HomologyByXCorr and the unmodified reference tool on the same FASTA
    pair write the same match file (names, sizes, every record; probabilities to 1e-6 relative)."""
    import oracle

    if not os.path.exists(oracle.REF_TOOL):
        pytest.skip("oracle/_ref/HomologyByXCorr_ref not built")
    T, Q, t, q = _equal_length_case(tmp_path)
    ref_out, my_out = tmp_path / "ref.match", tmp_path / "b200.match"
    subprocess.run([oracle.REF_TOOL, "-q", q, "-t", t, "-o", str(ref_out)], check=True, capture_output=True)
    r = subprocess.run([host_bins["HomologyByXCorr"], "-q", q, "-t", t, "-o", str(my_out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    a = _parse_match_file(str(ref_out))
    b = _parse_match_file(str(my_out))
    assert a[0] == b[0] and a[1] == b[1] and a[3] == b[3] and a[4] == b[4]  # names and sizes
    ra, rb = sorted(a[2]), sorted(b[2])
    assert len(ra) == len(rb) >= 15
    for x, y in zip(ra, rb):
        assert x[:7] == y[:7] and x[9] == y[9], (x, y)          # coordinates, strand, identity
        assert abs(x[8] - y[8]) <= 1e-6 * abs(x[8]) and abs(x[7] - y[7]) <= 1e-9


@pytest.mark.gpu
def test_guided_refinement_equals_the_reference_tool(host_bins, sx, reference_lib, tmp_path):
    """`-guide` (the refinement pass behind SatsumaSynteny2 -do_refine): executable against executable.  Chained
    guide matches every 1500 bases make every gap -- hence every piece -- equally long, so the reference tool's
    re-used signal objects carry no stale samples and its output is a valid target.  One sequence
    pair is forward, the other reverse (forced orientation -1, query window taken from the far end)."""
    import oracle

    if not os.path.exists(oracle.REF_TOOL):
        pytest.skip("oracle/_ref/HomologyByXCorr_ref not built")
    rng = np.random.default_rng(8)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    comp = np.zeros(256, np.uint8)
    comp[list(b"ACGT")] = list(b"TGCA")
    t0, t1 = rng.choice(acgt, 60000), rng.choice(acgt, 40000)

    def diverged(x):
        y = x.copy()
        mut = rng.random(len(y)) < 0.10
        y[mut] = rng.choice(acgt, int(mut.sum()))
        return y

    q0, q1 = diverged(t0), comp[diverged(t1)[::-1]]
    t, q = tmp_path / "t.fa", tmp_path / "q.fa"
    _write_fasta(t, [("t0", t0.tobytes()), ("t1", t1.tobytes())])
    _write_fasta(q, [("q0", q0.tobytes()), ("q1", q1.tobytes())])
    recs = []
    for k in range(36):
        recs.append([0, 0, len(q0), 1000 + 1500 * k, 1000 + 1500 * k, 100, 0, 90.0, 0.99995, 0.9])
    for k in range(22):  # reverse strand: query coordinates are on the reverse complement, colinear with the target
        recs.append([1, 1, len(q1), 2000 + 1500 * k, 2000 + 1500 * k, 100, 1, 88.0, 0.99991, 0.88])
    guide = tmp_path / "chained.match"
    with open(guide, "wb") as f:
        f.write(struct.pack("<ii", 3, 2))
        for name in (b"t0\0", b"t1\0"):
            f.write(struct.pack("<q", len(name)) + name)
        f.write(struct.pack("<i", 2))
        for name in (b"q0\0", b"q1\0"):
            f.write(struct.pack("<q", len(name)) + name)
        f.write(struct.pack("<i", len(recs)))
        for r in recs:
            f.write(struct.pack("<iiiiiiiddd", *[int(x) for x in r[:7]], *[float(x) for x in r[7:]]))
        f.write(struct.pack("<2i", len(t0), len(t1)) + struct.pack("<2i", len(q0), len(q1)))
    ref_out, my_out = tmp_path / "ref.match", tmp_path / "b200.match"
    common = ["-q", str(q), "-t", str(t), "-guide", str(guide), "-cutoff", "1.2"]
    subprocess.run([oracle.REF_TOOL, *common, "-o", str(ref_out)], check=True, capture_output=True)
    r = subprocess.run([host_bins["HomologyByXCorr"], *common, "-o", str(my_out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    a, b = _parse_match_file(str(ref_out)), _parse_match_file(str(my_out))
    assert a[0] == b[0] and a[1] == b[1] and a[3] == b[3] and a[4] == b[4]
    ra, rb = sorted(a[2]), sorted(b[2])
    new_fwd = [x for x in ra if x[6] == 0 and x[5] != 100]
    new_rev = [x for x in ra if x[6] == 1 and x[5] != 100]
    assert len(new_fwd) > 20 and len(new_rev) > 10      # the pass found matches inside the gaps, on both strands
    assert len(ra) == len(rb)
    for x, y in zip(ra, rb):
        assert x[:7] == y[:7] and x[9] == y[9], (x, y)
        assert abs(x[8] - y[8]) <= 1e-6 * abs(x[8]) and abs(x[7] - y[7]) <= 1e-9


class _ScriptedMaster:
    """The master's side of the slave protocol (analysis/WorkQueue.cc:71-166) on loopback TCP:
    per connection  <- uint32 slave_id, uint32 n, n x t_result ; -> int32 count, count x t_pair (one send).
    Hands out `per_exchange` blocks per connection (the reference master: 2 x its thread count), answers 0 while the
    slave still owes results and -1 once every block's results can have arrived."""

    def __init__(self, sx, blocks, sid, per_exchange=2, quiet_s=1.5):
        self.sx, self.blocks, self.sid, self.per, self.quiet_s = sx, list(blocks), sid, per_exchange, quiet_s
        self.last_data = time.monotonic()
        self.srv = socket.socket()
        self.srv.bind(("127.0.0.1", 0))
        self.srv.listen(16)
        self.port = self.srv.getsockname()[1]
        self.received, self.connections, self.sent = [], 0, 0
        self.done = False
        self.thread = threading.Thread(target=self._serve, daemon=True)
        self.thread.start()

    @staticmethod
    def _recv_all(c, n):
        buf = b""
        while len(buf) < n:
            part = c.recv(n - len(buf))
            if not part:
                raise IOError("short read")
            buf += part
        return buf

    def _serve(self):
        sx = self.sx
        while not self.done:
            c, _ = self.srv.accept()
            self.connections += 1
            sid, n = struct.unpack("<II", self._recv_all(c, 8))
            assert sid == self.sid
            if n:
                self.received.append(np.frombuffer(self._recv_all(c, 72 * n), dtype=sx.RESULT_DTYPE).copy())
            now = time.monotonic()
            if n:
                self.last_data = now
            if self.sent < len(self.blocks):
                part = self.blocks[self.sent:self.sent + self.per]
                self.sent += len(part)
                self.last_data = now
                c.sendall(struct.pack("<i", len(part)) + sx.make_blocks(part).tobytes())
            elif now - self.last_data < self.quiet_s:
                # everything handed out: keep answering "nothing now" until no records have arrived for a while
                # (the reference slave's workers poll their queue once a second)
                time.sleep(0.2)
                c.sendall(struct.pack("<i", 0))
            else:
                c.sendall(struct.pack("<i", -1))
                self.done = True
            c.close()

    def records(self):
        self.thread.join(timeout=60)
        self.srv.close()
        return np.concatenate(self.received) if self.received else np.zeros(0, dtype=self.sx.RESULT_DTYPE)


def _slave_case(tmp_path):
    rng = np.random.default_rng(5)
    base = rng.choice(list(b"ACGT"), 60000).astype(np.uint8)
    qry = base.copy()
    mut = rng.random(len(base)) < 0.1
    qry[mut] = rng.choice(list(b"ACGT"), int(mut.sum()))
    qry[30000:40000] = np.frombuffer(bytes(qry[30000:40000][::-1]).translate(bytes.maketrans(b"ACGT", b"TGCA")), np.uint8)
    q, t = tmp_path / "q.fa", tmp_path / "t.fa"
    _write_fasta(q, [("q", qry.tobytes())])
    _write_fasta(t, [("t", base.tobytes())])
    # 20 target chunks (overlap 1024), 15 query chunks; blocks tile the grid, one of them in "fast" mode
    blocks = [(t0, min(t0 + 4, 19), q0, min(q0 + 3, 14), 0) for t0 in range(0, 20, 5) for q0 in range(0, 15, 4)]
    blocks.append((0, 19, 0, 14, 1))
    return str(q), str(t), blocks


def _reference_block_records(q, t, blocks):
    """What the UNMODIFIED reference emits for these t_pairs: its own FASTA loader + ChunkManager +
    HomologyByXCorr::align_target (oracle/_ref/libsatsuma_ref.so)."""
    import oracle

    R = oracle.Reference()
    R.configure()
    R.load_fasta(t, q)
    return np.concatenate([R.align_block(*b[:4], fast=bool(b[4])) for b in blocks])


@pytest.mark.gpu
def test_slave_speaks_the_reference_wire_protocol(host_bins, sx, reference_lib, tmp_path):
    """A scripted master hands t_pairs to the B200 slave over loopback TCP, two per connection like the reference's
    WorkQueue; the slave queues them over several connections, aligns them in batches on the GPU while it keeps
    exchanging, and delivers every record.  Expected: the unmodified reference's align_target on the same FASTA."""
    q, t, blocks = _slave_case(tmp_path)
    exp = _reference_block_records(q, t, blocks)
    m = _ScriptedMaster(sx, blocks, sid=7)
    r = subprocess.run([host_bins["HomologyByXCorrSlave"], "-master", "127.0.0.1", "-port", str(m.port), "-sid", "7", "-q",
                        q, "-t", t, "-device", "0"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    got = m.records()
    assert len(exp) > 100
    assert sorted(map(rec_key, got)) == sorted(map(rec_key, exp))
    ge = {rec_key(x): x for x in got}
    for x in exp:
        assert ge[rec_key(x)]["ident"] == x["ident"]
        assert abs(float(ge[rec_key(x)]["prob"]) - float(x["prob"])) <= 1e-6 * abs(float(x["prob"]))
    # several blocks per GPU batch although the master hands out two per connection
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("exchanges")][0]
    nb = float(line.split("(")[1].split()[0])
    assert nb > 2.0, line


@pytest.mark.gpu
def test_reference_slave_with_the_binding_compiled_in(sx, reference_lib, tmp_path):
    """INTEGRATION.md section 2, built for real: the reference's own HomologyByXCorrSlave.cc with
    HomologyByXCorr::align_target calling sx_align_blocks (oracle/make_refslave_b200.py patches the translation unit
    where it lies under /root/reference; flag parsing, FASTA loader, ChunkManager, worker threads and the TCP loop
    stay the reference's).  Driven by the scripted master; records identical to the unmodified reference's."""
    import oracle

    exe = os.path.join(os.path.dirname(oracle.REF_SO), "HomologyByXCorrSlave_b200bind")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/HomologyByXCorrSlave_b200bind not built (make -C oracle refslave_b200, needs /root/reference)")
    q, t, blocks = _slave_case(tmp_path)
    exp = _reference_block_records(q, t, blocks)
    # the reference slave asks for more whenever fewer than 2 x p blocks are queued and otherwise sleeps 10 s
    # (Slave.cc:451, 524); its workers poll once a second: a generous quiet period before "terminate"
    m = _ScriptedMaster(sx, blocks, sid=1, per_exchange=8, quiet_s=4.0)
    env = dict(os.environ, SX_DEVICE="0")
    p = subprocess.Popen([exe, "-master", "127.0.0.1", "-port", str(m.port), "-sid", "1", "-p", "8", "-q", q, "-t", t],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env)
    try:
        out, err = p.communicate(timeout=300)
    except subprocess.TimeoutExpired:
        p.kill()
        raise
    got = m.records()
    assert p.returncode == 0, err[-2000:]
    assert sorted(map(rec_key, got)) == sorted(map(rec_key, exp))
    ge = {rec_key(x): x for x in got}
    for x in exp:
        assert ge[rec_key(x)]["ident"] == x["ident"]


@pytest.mark.gpu
def test_class_level_shims_match_the_reference_classes(sx, reference_lib, tmp_path):
    """sx_shim::CCSignal / CrossCorrelation / SeqAnalyzer (host/binding/crosscorr_shim.h), the
    reference's class interfaces on top of the C ABI, in one program with the reference's own classes
    (oracle/shim_check.cc, compiled against the reference headers): signals bit-equal, correlation within 1e-4,
    MatchUp on the reference's correlation vector gives identical segment lists, both strands."""
    import oracle

    exe = os.path.join(os.path.dirname(oracle.REF_SO), "shim_check")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/shim_check not built (python oracle/make_refslave_b200.py, needs /root/reference)")
    rng = np.random.default_rng(8)
    t = rng.choice(list(b"ACGT"), 4096).astype(np.uint8)
    q = np.concatenate([rng.choice(list(b"ACGT"), 700).astype(np.uint8), t[500:3000]])
    mut = rng.random(len(q)) < 0.08
    q[mut] = rng.choice(list(b"ACGT"), int(mut.sum()))
    q[100] = ord("N")
    tf, qf = tmp_path / "t.fa", tmp_path / "q.fa"
    _write_fasta(tf, [("t", t.tobytes())])
    _write_fasta(qf, [("q", q.tobytes())])
    r = subprocess.run([exe, str(tf), str(qf)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "SHIM OK" in r.stdout
    n_seg = int(r.stdout.split("err")[1].split(",")[1].split()[0])
    assert n_seg > 1000
