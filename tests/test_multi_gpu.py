"""sx_multi: the chunk-pair grid split by target range over GPUs (and ranks).  -m gpu.
One-GPU box: the split is exercised with shard_world > 1 (several handles on the same device, each owning a
part of the target list and fetching only the query range its blocks touch); the union must be the unsharded
result pair for pair.  With >= 2 GPUs the same through one handle with several devices (host gather)."""
import numpy as np
import pytest

from parity import rec_key

pytestmark = pytest.mark.gpu


def _genome_case(sx):
    from satsuma2_b200 import synth

    tgt, qry = synth.genome_pair(400000, seed=17)
    to, tl, ts = synth.chunk_sequence(tgt, 4096, 1024)
    qo, ql, qs = synth.chunk_sequence(qry, 4096, 0)
    tcs = sx.ChunkSet(tgt, to, tl, ts, np.zeros(len(tl), np.int32), [len(tgt)])
    qcs = sx.ChunkSet(qry, qo, ql, qs, np.zeros(len(ql), np.int32), [len(qry)])
    blocks = synth.diagonal_blocks(len(tl), len(ql), 3072, 4096, pixel=12)
    return tgt, tcs, qcs, blocks


def _sorted(recs):
    return sorted((rec_key(x), float(x["prob"]), float(x["ident"])) for x in recs)


def test_sharded_handles_reproduce_the_unsharded_result(sx):
    tgt, tcs, qcs, blocks = _genome_case(sx)
    with sx.XCorrEngine(target_total=float(len(tgt))) as eng:
        eng.set_targets(tcs)
        eng.set_queries(qcs)
        whole = _sorted(eng.align_blocks(blocks))
        n_pairs = eng.stats()["chunk_pairs"]
    assert len(whole) > 100
    for world in (1, 3):
        parts, pairs, h2d = [], 0, []
        for rank in range(world):
            with sx.MultiEngine(devices=[0], shard_rank=rank, shard_world=world) as m:
                m.set_targets(tcs)
                m.set_queries(qcs)
                lo, hi = m.target_range(0)
                assert hi - lo <= -(-len(tcs) // world)
                r1 = m.align_blocks(blocks)
                r2 = m.align_blocks(blocks)  # resident query range and cached target spectra are reused
                assert _sorted(r1) == _sorted(r2)
                parts += list(_sorted(r1))
                st = m.stats()
                pairs += st["chunk_pairs"] // 2
                h2d.append(st["h2d_bytes"])
        assert sorted(parts) == whole
        assert pairs == n_pairs
        if world > 1:
            # every shard was sent its own part of the genomes only (plus descriptors), not both whole genomes
            assert max(h2d) < 0.6 * (tcs.bases.nbytes + qcs.bases.nbytes)


def test_one_handle_many_gpus_gathers_one_list(sx):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    tgt, tcs, qcs, blocks = _genome_case(sx)
    with sx.XCorrEngine(target_total=float(len(tgt))) as eng:
        eng.set_targets(tcs)
        eng.set_queries(qcs)
        whole = _sorted(eng.align_blocks(blocks))
    with sx.MultiEngine() as m:  # every visible GPU
        assert m.n_devices == torch.cuda.device_count()
        m.set_targets(tcs)
        m.set_queries(qcs)
        got = m.align_blocks(blocks)
        per_dev = [m.stats(i)["chunk_pairs"] for i in range(m.n_devices)]
        spans = [m.target_range(i) for i in range(m.n_devices)]
    assert _sorted(got) == whole
    assert all(p > 0 for p in per_dev)
    assert spans[0][0] == 0 and spans[-1][1] == len(tcs) and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
