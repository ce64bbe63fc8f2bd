"""N > 1 path on CPU: world_size-2 gloo run of the benchmark's rank plumbing (shard seeds / ranges,
barrier, max- and sum-reductions that turn per-rank timings into the whole-job number).  The data
path itself has no collective: every rank owns its chunk pairs and its match list."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


WORKER = textwrap.dedent(
    """
    import os, sys, json
    sys.path.insert(0, %r)
    import numpy as np
    from satsuma2_b200.dist import Group, shard_seed, shard_range
    from satsuma2_b200 import synth
    g = Group("gloo")
    assert g.world == 2
    # every rank generates ITS shard (weak scaling) from its own seed
    T, Q, truth = synth.random_pairs(8, 256, seed=shard_seed(1, g.rank))
    digest = int(T.astype(np.int64).sum() * 31 + Q.astype(np.int64).sum())
    lo, hi = shard_range(101, g.rank, g.world)
    g.barrier()
    ms = g.max(10.0 + 5.0 * g.rank)          # slowest rank defines the step time
    pairs = g.sum(float(T.shape[0]))          # whole-job units = sum over ranks
    covered = g.sum(float(hi - lo))
    print(json.dumps({"rank": g.rank, "digest": digest, "ms": ms, "pairs": pairs, "range": [lo, hi],
                      "covered": covered}), flush=True)
    g.close()
    """
)


def test_two_rank_gloo_plumbing(tmp_path):
    port = _free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, "-c", WORKER % ROOT], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    outs = []
    for p in procs:
        out, err = p.communicate(timeout=180)
        assert p.returncode == 0, err[-2000:]
        outs.append(eval(out.strip().splitlines()[-1].replace("true", "True").replace("false", "False")))
    outs.sort(key=lambda d: d["rank"])
    assert outs[0]["ms"] == outs[1]["ms"] == 15.0          # max over ranks
    assert outs[0]["pairs"] == outs[1]["pairs"] == 16.0    # sum over ranks
    assert outs[0]["digest"] != outs[1]["digest"]          # different shards
    assert outs[0]["range"] == [0, 51] and outs[1]["range"] == [51, 101]
    assert outs[0]["covered"] == 101.0


def test_shard_helpers():
    from satsuma2_b200.dist import shard_range, shard_seed

    for n, w in ((10, 3), (1, 8), (0, 2), (1 << 20, 8)):
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    assert len({shard_seed(1, r) for r in range(8)}) == 8


def test_block_sharding_by_target_range_partitions_the_work_list():
    """Block mode: the t_pair work list clipped to per-rank target ranges covers every chunk pair exactly once."""
    from satsuma2_b200 import synth
    from satsuma2_b200.dist import shard_blocks_by_target

    n_t, n_q = 431, 330
    blocks = synth.diagonal_blocks(n_t, n_q, 3072, 4096, pixel=24)
    assert len(blocks) == 14

    def pairs(bl):
        return [(t, q) for b in bl for q in range(b[2], b[3] + 1) for t in range(b[0], b[1] + 1)]

    want = sorted(pairs(blocks))
    for world in (1, 2, 4, 8):
        got = []
        for rank in range(world):
            mine = shard_blocks_by_target(blocks, n_t, rank, world)
            lo = (n_t + world - 1) // world * rank
            assert all(lo <= b[0] <= b[1] < lo + (n_t + world - 1) // world for b in mine)
            got += pairs(mine)
        assert sorted(got) == want and len(set(got)) == len(got)


def test_synth_generator_is_deterministic_and_plants_segments():
    from satsuma2_b200 import synth

    T1, Q1, tr1 = synth.random_pairs(40, 4096, seed=7)
    T2, Q2, tr2 = synth.random_pairs(40, 4096, seed=7)
    assert np.array_equal(T1, T2) and np.array_equal(Q1, Q2) and np.array_equal(tr1, tr2)
    assert set(np.unique(T1)) <= set(b"ACGT") and set(np.unique(Q1)) <= set(b"ACGT")
    comp = np.zeros(256, np.uint8)
    comp[list(b"ACGT")] = list(b"TGCA")
    for i in range(40):
        t, q, n, rev = int(tr1["tpos"][i]), int(tr1["qpos"][i]), int(tr1["len"][i]), bool(tr1["reverse"][i])
        assert 60 <= n <= 500
        a = T1[i, t:t + n]
        b = Q1[i, q:q + n]
        if rev:
            b = comp[b[::-1]]
        ident = float((a == b).mean())
        assert ident > 0.6, (i, ident)


def test_chunk_arithmetic_matches_reference_counts():
    """ChunkManager arithmetic (analysis/SeqChunk.cc:72-165): sample sizes give 261 / 245 chunks."""
    from satsuma2_b200 import synth

    dog = np.full(800001, ord("A"), np.uint8)
    human = np.full(1000001, ord("A"), np.uint8)
    o, l, s = synth.chunk_sequence(dog, 4096, 1024)
    assert len(l) == 261 and l[-1] == 1281 and o[1] == 3072
    o, l, s = synth.chunk_sequence(human, 4096, 0)
    assert len(l) == 245 and l[-1] == 577
    # a length that is a multiple of the stride yields a trailing empty chunk
    o, l, s = synth.chunk_sequence(np.full(8192, ord("C"), np.uint8), 4096, 0)
    assert list(l) == [4096, 4096, 0]
    # all-N chunks are emptied but keep their index; sequences < 6 bp produce nothing
    seq = np.concatenate([np.full(4096, ord("N"), np.uint8), np.full(100, ord("G"), np.uint8)])
    o, l, s = synth.chunk_sequence(seq, 4096, 0)
    assert list(l) == [0, 100]
    assert len(synth.chunk_sequence(np.full(5, ord("A"), np.uint8), 4096, 0)[1]) == 0
