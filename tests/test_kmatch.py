"""KMatch seeding on the GPU (include/satsuma_kmatch.h) against the reference's own KMatch
program compiled unmodified (oracle/_ref/KMatch_ref, `make -C oracle refkmatch`): same FASTA pair, same command line,
same set of t_result records.  -m gpu."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DT = np.dtype([("query_id", "<u8"), ("target_id", "<u8"), ("query_size", "<u8"), ("qstart", "<u8"), ("tstart", "<u8"),
               ("len", "<u8"), ("reverse", "u1"), ("pad", "u1", (7,)), ("prob", "<f8"), ("ident", "<f8")])
ACGT = np.frombuffer(b"ACGT", np.uint8)
COMP = np.zeros(256, np.uint8)
COMP[list(b"ACGT")] = list(b"TGCA")


def _fa(path, recs, width=70):
    with open(path, "w") as f:
        for name, s in recs:
            f.write(">" + name + "\n")
            s = s.tobytes().decode() if isinstance(s, np.ndarray) else s
            for i in range(0, len(s), width):
                f.write(s[i:i + width] + "\n")


def _keys(recs):
    return sorted((int(r["query_id"]), int(r["target_id"]), int(r["query_size"]), int(r["qstart"]), int(r["tstart"]),
                   int(r["len"]), int(r["reverse"]), float(r["prob"]), float(r["ident"])) for r in recs)


def _case(tmp_path, seed, n_t=3, n_q=3, size=60000, lower=False, with_n=False):
    rng = np.random.default_rng(seed)
    T = [rng.choice(ACGT, int(size * rng.uniform(0.6, 1.0))) for _ in range(n_t)]
    Q = []
    for i in range(n_q):
        parts = [rng.choice(ACGT, int(rng.integers(50, 800)))]  # query coordinates never equal target coordinates
        for _ in range(int(rng.integers(2, 6))):
            t = T[int(rng.integers(0, n_t))]
            a = int(rng.integers(0, len(t) - 3000))
            seg = t[a:a + int(rng.integers(500, 9000))].copy()
            mut = rng.random(len(seg)) < rng.uniform(0.0, 0.06)
            seg[mut] = rng.choice(ACGT, int(mut.sum()))
            if rng.random() < 0.5:
                seg = COMP[seg[::-1]]
            parts += [seg, rng.choice(ACGT, int(rng.integers(20, 2000)))]
        Q.append(np.concatenate(parts))
    if with_n:
        T[0][1000:1040] = ord("N")
        Q[0][700] = ord("N")
        Q[-1][300:330] = ord("n")
    # a repeated stretch inside one target: its k-mers occur twice and fall to max_freq = 1
    m = len(T[-1])
    T[-1][m // 8:m // 8 + 600] = T[-1][m // 2:m // 2 + 600]
    tq, tt = tmp_path / "q.fa", tmp_path / "t.fa"

    def text(a):
        s = a.tobytes().decode()
        return s.lower() if lower else s

    _fa(tt, [(f"t{i}", text(t)) for i, t in enumerate(T)])
    _fa(tq, [(f"q{i} some description", text(q)) for i, q in enumerate(Q)])
    return str(tq), str(tt)


@pytest.fixture(scope="module")
def kmatch_bins(sx):
    import oracle
    from satsuma2_b200 import build as sxbuild

    exes = {os.path.basename(e): e for e in sxbuild.build_host()}
    if not os.path.exists(oracle.REF_KMATCH):
        pytest.skip("oracle/_ref/KMatch_ref not built (make -C oracle refkmatch, needs /root/reference)")
    return exes["KMatch"], oracle.REF_KMATCH


@pytest.mark.gpu
@pytest.mark.parametrize("K,seed,kw", [(31, 1, {}), (15, 2, {}), (21, 3, dict(lower=True)), (27, 4, dict(with_n=True)),
                                       (11, 5, dict(size=20000))])
def test_kmatch_equals_the_reference_program(kmatch_bins, tmp_path, K, seed, kw):
    mine, ref = kmatch_bins
    q, t = _case(tmp_path, seed, **kw)
    out_r, out_m = str(tmp_path / "ref.k"), str(tmp_path / "b200.k")
    args = [q, t, str(K), None, str(K), str(K - 1), "1"]  # as SatsumaSynteny2 calls it (SatsumaSynteny2.cc:419-420)
    subprocess.run([ref] + args[:3] + [out_r] + args[4:], check=True, capture_output=True)
    r = subprocess.run([mine] + args[:3] + [out_m] + args[4:], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    a, b = np.fromfile(out_r, dtype=DT), np.fromfile(out_m, dtype=DT)
    assert len(a) > 20
    assert _keys(b) == _keys(a)
    assert a["reverse"].any() and not a["reverse"].all()


@pytest.mark.gpu
def test_kmatch_min_length_and_jump(kmatch_bins, tmp_path):
    mine, ref = kmatch_bins
    q, t = _case(tmp_path, 9)
    for K, min_len, jump in ((19, 60, 5), (19, 19, 40), (25, 200, 24)):
        out_r, out_m = str(tmp_path / "ref.k"), str(tmp_path / "b200.k")
        subprocess.run([ref, q, t, str(K), out_r, str(min_len), str(jump), "1"], check=True, capture_output=True)
        r = subprocess.run([mine, q, t, str(K), out_m, str(min_len), str(jump), "1"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        a, b = np.fromfile(out_r, dtype=DT), np.fromfile(out_m, dtype=DT)
        assert len(a) > 5 and _keys(b) == _keys(a)


@pytest.mark.gpu
def test_kmatch_rejects_what_the_reference_cannot_run(kmatch_bins, tmp_path):
    mine, _ = kmatch_bins
    q, t = _case(tmp_path, 11)
    r = subprocess.run([mine, q, t, "16", str(tmp_path / "o"), "16", "15", "1"], capture_output=True, text=True)
    assert r.returncode == 1 and "odd" in r.stdout


@pytest.mark.gpu
def test_kmatch_on_the_sample_genomes(kmatch_bins, tmp_path):
    """configs[0]'s sequences (samples/dog.X.part.fasta as target, human.X.part.fasta as query, from
    tests/golden/samples_full.npz) through both programs with the master's first seeding round (K = 31)."""
    mine, ref = kmatch_bins
    g = np.load(os.path.join(ROOT, "tests", "golden", "samples_full.npz"))
    q, t = str(tmp_path / "q.fa"), str(tmp_path / "t.fa")
    _fa(q, [("human.X.part", g["q_seq"])])
    _fa(t, [("dog.X.part", g["t_seq"])])
    for K in (31, 21):
        out_r, out_m = str(tmp_path / "ref.k"), str(tmp_path / "b200.k")
        subprocess.run([ref, q, t, str(K), out_r, str(K), str(K - 1), "1"], check=True, capture_output=True)
        r = subprocess.run([mine, q, t, str(K), out_m, str(K), str(K - 1), "1"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        a, b = np.fromfile(out_r, dtype=DT), np.fromfile(out_m, dtype=DT)
        assert _keys(b) == _keys(a)
    assert len(a) > 0
