"""N>1 path on CPU: two gloo ranks shard the trial index range, each computes its packed statistics, and ONE all-reduce
(sum, plus a max for x_max) must reproduce the single-process statistics exactly.  The per-rank compute
is the numpy oracle here; on the GPU box bench.py runs the same sharding + all-reduce with the CUDA engine over NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import BASIC_MEAN, BASIC_SIGMA, THERMAL_COEF, ROOT
from startrackermpm_b200 import sharding

N, SEED, F = 900, 17, 3500.0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rank_stats(begin, end):
    from oracle import stats as ostats
    from oracle import trial
    from startrackermpm_b200.catalog import load_catalog
    xyz, mag = load_catalog()
    cols = trial.sample_params(BASIC_MEAN, BASIC_SIGMA, THERMAL_COEF, SEED, begin, end - begin)
    e, d = trial.run_trials(cols, xyz, mag, F, SEED, begin)
    return ostats.packed_stats(e, d)


def _worker(rank, world, port, out_dir):
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        b, e = sharding.shard_range(N, rank, world)
        st = torch.from_numpy(_rank_stats(b, e))
        sharding.allreduce_stats(st)
        np.save(os.path.join(out_dir, f"stats_{rank}.npy"), st.numpy())
    finally:
        dist.destroy_process_group()


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 100, 10**10):
        for w in (1, 2, 3, 8):
            spans = [sharding.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    assert [list(sharding.shard_segments(10, r, 4)) for r in range(4)] == [[0, 4, 8], [1, 5, 9], [2, 6], [3, 7]]


@pytest.mark.timeout(300)
def test_two_rank_allreduce_equals_single_process(tmp_path):
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    s0, s1 = np.load(tmp_path / "stats_0.npy"), np.load(tmp_path / "stats_1.npy")
    assert np.array_equal(s0, s1)
    whole = _rank_stats(0, N)
    head = [0, 1, 4, 5, 6, 7]                                  # counts and x_max: exact
    assert np.array_equal(s0[head], whole[head])
    assert np.array_equal(s0[8:], whole[8:])                   # histogram: exact
    assert np.allclose(s0[[2, 3]], whole[[2, 3]], rtol=1e-13)  # float sums: order of addition differs
