"""world_size-2 gloo test of the multi-GPU host logic (CPU): replicate sharding by global id and
the single all-gather of result rows.  The per-rank rows come from the oracle (tests only)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, total, tmp):
    for p in (ROOT, os.path.join(ROOT, "plspm-python_b200")):
        sys.path.insert(0, p)
    import torch.distributed as dist
    from oracle import plspm_oracle as orc
    from plspm_b200 import distributed as pdist
    from plspm_b200.synth import make_synthetic
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert pdist.rank_world() == (rank, world)
        seed = pdist.broadcast_int(1234 + rank)  # every rank ends up with rank 0's seed
        assert seed == 1234
        X, path = make_synthetic(300, 4, 3, seed=2)
        begin, count = pdist.shard_range(total, rank, world)
        idx = np.stack([orc.philox_indices(seed, begin + b, 300) for b in range(count)]) if count else \
            np.zeros((0, 300), dtype=np.int32)
        rows, iters, status = orc.bootstrap(X, idx, [3] * 4, [0] * 4, path, "centroid", True) if count else \
            (np.zeros((0, 1)), np.zeros(0, np.int32), np.zeros(0, np.int32))
        width = 2 * 12 + 4 + 2 * len(orc.effect_pairs(path))
        rows = rows.reshape(count, width)
        allr, alls, alli = pdist.allgather_rows(rows, status, iters, total, width)
        np.savez(os.path.join(tmp, "rank%d.npz" % rank), rows=allr, status=alls, iters=alli, begin=begin, count=count)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", (7, 8))
def test_two_rank_shard_and_allgather(tmp_path, total):
    import torch.multiprocessing as mp
    from oracle import plspm_oracle as orc
    from plspm_b200.synth import make_synthetic
    port = _free_port()
    mp.spawn(_worker, args=(2, port, total, str(tmp_path)), nprocs=2, join=True)
    X, path = make_synthetic(300, 4, 3, seed=2)
    idx = np.stack([orc.philox_indices(1234, b, 300) for b in range(total)])
    rows, iters, status = orc.bootstrap(X, idx, [3] * 4, [0] * 4, path, "centroid", True)
    got = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(2)]
    assert int(got[0]["count"]) + int(got[1]["count"]) == total and int(got[1]["begin"]) == int(got[0]["count"])
    for g in got:  # every rank holds every replicate, in global replicate order, independent of world size
        np.testing.assert_array_equal(g["rows"], rows)
        np.testing.assert_array_equal(g["iters"], iters)
        np.testing.assert_array_equal(g["status"], status)


def test_shard_range_partitions_exactly():
    from plspm_b200.distributed import shard_range
    for total in (0, 1, 7, 10000, 10001):
        for world in (1, 2, 4, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (b0, c0), (b1, _) in zip(spans, spans[1:]):
                assert b1 == b0 + c0
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
            assert spans[0][1] == max(c for _, c in spans)
