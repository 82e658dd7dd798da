"""Worker of tests/test_parallel_gloo.py: run under torch.distributed.run with the gloo backend on
CPU.  The device evaluator is replaced by the CPU oracle (tests only); everything else -- strided
sharding, the all-gather of per-rank best rows, the merge, the split of refinement starts -- is the
product's covest_b200.parallel."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from covest_b200 import parallel, workload  # noqa: E402
from oracle import covest_oracle as orc  # noqa: E402
from tests.helpers import case_hist, load_case  # noqa: E402


def main():
    out_path = sys.argv[1]
    dist.init_process_group('gloo')
    rank, world = parallel.world()
    case = load_case('e05_repeats')
    model = orc.Model('repeats', 21, 100, case_hist(case), 0, max_error=8)
    axes = [np.geomspace(5, 20, 6), np.geomspace(.01, .2, 5), np.linspace(.3, 1, 3),
            np.linspace(0, 1, 3), np.linspace(.05, 1, 3)]
    total = int(np.prod([len(a) for a in axes]))
    k_best = 8

    block = 27  # the (q1, q2, q) combinations: whole (coverage, error_rate) groups per rank

    def evaluate_slice(first, stride, block, count):
        pts = workload.lattice_points(axes, first=first, stride=stride, count=count, block=block)
        ll = model.loglik_batch(pts, threads=2)
        key = np.where(np.isnan(ll), -np.inf, ll)
        order = np.lexsort((np.arange(len(key)), -key))[:k_best]
        rows = np.column_stack([ll[order], pts[order]])
        if len(rows) < k_best:
            pad = np.full((k_best - len(rows), rows.shape[1]), np.nan)
            pad[:, 0] = -np.inf
            rows = np.vstack([rows, pad])
        return torch.from_numpy(rows)

    rows = parallel.sharded_best_rows(evaluate_slice, total, k_best, block=block)
    mine = parallel.split_starts(rows)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine.numpy().tolist())

    # the refinement after the all-gather (CoverageEstimator.refine_starts): every rank refines its
    # share of the global best rows with the lock-step optimiser, the optima are all-gathered
    from covest_b200.covest import CoverageEstimator
    from tests.test_host_logic import ORepeats
    est = CoverageEstimator(ORepeats(21, 100, case_hist(case), 0, max_error=8), optimizer='lockstep')
    x, fun, ok, table = est.refine_starts(rows.numpy()[:4, 1:])
    tables = [None] * world
    dist.all_gather_object(tables, table.tolist())
    drawn = parallel.broadcast_rows(np.full((2, 3), float(rank + 5)))
    if rank == 0:
        with open(out_path, 'w') as f:
            json.dump({'world': world, 'rows': rows.numpy().tolist(), 'starts': gathered,
                       'slice': list(parallel.shard_blocked(total, block, rank, world)),
                       'refined_x': [float(v) for v in x], 'refined_fun': fun, 'refined_ok': ok,
                       'tables': tables, 'broadcast': drawn.tolist(), 'launches': est.launches}, f)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
