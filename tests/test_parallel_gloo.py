"""The N > 1 path on CPU: world_size 2 over gloo."""
import json
import os
import subprocess
import sys

import numpy as np

from covest_b200 import parallel, workload
from oracle import covest_oracle as orc
from tests.helpers import case_hist, load_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharding_helpers():
    assert parallel.shard_strided(10, 0, 4) == (0, 4, 3)
    assert parallel.shard_strided(10, 3, 4) == (3, 4, 2)
    assert parallel.shard_strided(2, 3, 4) == (3, 4, 0)
    assert sum(parallel.shard_strided(1000003, r, 8)[2] for r in range(8)) == 1000003
    spans = [parallel.shard_contiguous(10, r, 4) for r in range(4)]
    assert spans == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert parallel.world() == (0, 1)
    rows = np.array([[-5.0, 1], [np.nan, 2], [-1.0, 3], [-1.0, 4], [-np.inf, 5]])
    got = parallel.merge_topk(rows, 3).numpy()
    assert got[:, 1].tolist() == [3, 4, 1]  # ties: lower parameter first; NaN last


def test_lattice_points_order_is_itertools_product():
    import itertools
    axes = [np.array([1., 2.]), np.array([10., 20., 30.]), np.array([.1, .2])]
    want = np.array(list(itertools.product(*axes)))
    assert np.array_equal(workload.lattice_points(axes), want)
    assert np.array_equal(workload.lattice_points(axes, first=1, stride=5), want[1::5])
    # runs of `block` consecutive indices, dealt round-robin: the slicing of cvb_lattice_eval
    n = len(want)
    for w in (2, 3):
        seen = []
        for r in range(w):
            first, stride, block, count = parallel.shard_blocked(n, 4, r, w)
            part = workload.lattice_points(axes, first=first, stride=stride, count=count, block=block)
            idx = [(first + (i // block) * stride) * block + i % block for i in range(count)]
            assert np.array_equal(part, want[idx])
            seen += idx
        assert sorted(seen) == list(range(n))


def test_world_size_two_equals_single_process(tmp_path):
    out = str(tmp_path / 'out.json')
    env = dict(os.environ, OMP_NUM_THREADS='1')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2',
           '--master-addr', '127.0.0.1', '--master-port', '29533',
           os.path.join(ROOT, 'tests', 'dist_worker.py'), out]
    subprocess.run(cmd, check=True, env=env, timeout=600, cwd=ROOT,
                   stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
    with open(out) as f:
        res = json.load(f)
    assert res['world'] == 2
    # single-process answer
    case = load_case('e05_repeats')
    model = orc.Model('repeats', 21, 100, case_hist(case), 0, max_error=8)
    axes = [np.geomspace(5, 20, 6), np.geomspace(.01, .2, 5), np.linspace(.3, 1, 3),
            np.linspace(0, 1, 3), np.linspace(.05, 1, 3)]
    pts = workload.lattice_points(axes)
    ll = model.loglik_batch(pts, threads=4)
    order = np.argsort(-np.where(np.isnan(ll), -np.inf, ll), kind='stable')[:8]
    rows = np.array(res['rows'])
    assert np.array_equal(rows[:, 0], ll[order])
    assert np.array_equal(rows[:, 1:], pts[order])  # ties resolved as a single GPU would
    # the refinement starts are dealt round-robin and cover the best rows exactly once
    starts = [np.array(s) for s in res['starts']]
    assert np.array_equal(starts[0], rows[0::2]) and np.array_equal(starts[1], rows[1::2])
    # refinement of the best rows, sharded: the same optima as one process refining all of them
    from covest_b200.covest import CoverageEstimator
    from tests.test_host_logic import ORepeats
    est = CoverageEstimator(ORepeats(21, 100, case_hist(case), 0, max_error=8), optimizer='lockstep')
    x, fun, ok, table = est.refine_starts(rows[:4, 1:])
    assert res['tables'][0] == res['tables'][1]                  # every rank holds the same table
    assert np.array_equal(np.array(res['tables'][0]), table)    # ... and it is the single-process one
    assert res['refined_x'] == [float(v) for v in x] and res['refined_fun'] == fun and res['refined_ok'] == ok
    assert fun <= -rows[0, 0]                                    # refinement improves on the best lattice point
    assert res['broadcast'] == [[5.0] * 3] * 2                   # rank 0's block
