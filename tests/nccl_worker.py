"""Worker of tests/test_gpu_nccl.py: one process per GPU under torch.distributed.run, NCCL backend.
The flow of SURVEY.md section 8(e) on the device: the candidate lattice sharded over the ranks
(whole (coverage, error_rate) groups, points generated on the device), per-rank top-K, NCCL
all-gather, merge, refinement starts dealt round-robin, lock-step refinement, second all-gather."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from covest_b200 import constants, grid, parallel  # noqa: E402
from covest_b200.covest import CoverageEstimator  # noqa: E402
from covest_b200.models import RepeatsModel  # noqa: E402
from covest_b200.optimizer import lockstep_minimize  # noqa: E402
from tests.helpers import case_hist, load_case  # noqa: E402


def main():
    out_path = sys.argv[1]
    constants.VERBOSE = False
    local = int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    os.environ['COVEST_B200_DEVICE'] = str(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    rank, world = parallel.world()
    case = load_case('cfg2_repeats')
    model = RepeatsModel(21, 100, case_hist(case), 0, max_error=8)
    axes = [np.geomspace(10, 90, 12), np.geomspace(.01, .09, 8), np.linspace(.3, 1, 6),
            np.linspace(0, 1, 5), np.linspace(.05, 1, 8)]
    k_best = 16
    rows = grid.lattice_search(model, axes, k_best=k_best)               # sharded, all-gathered, merged
    ll_all, rows_one = model.device_context.lattice_eval(axes, k_best=k_best)  # the whole lattice on this rank
    est = CoverageEstimator(model, optimizer='lockstep')
    x, ok, rows2 = est.compute_coverage_from_lattice(axes, k_best=k_best)
    # what one process refining every start alone obtains
    starts = rows[np.isfinite(rows[:, 0]), 1:]
    single, _ = lockstep_minimize(est.likelihood_batch, starts, est.bounds)
    best = min(range(len(single)), key=lambda i: (single[i].fun, i))
    # multi-start through the reference's entry point: the random starts are rank 0's on every rank
    import random
    random.seed(100 + rank)  # different streams per rank on purpose
    x_sp, ok_sp = est.compute_coverage([30.0, .03, .65, .5, .5], starting_points=6)
    gathered = [None] * world
    dist.all_gather_object(gathered, {'x': [float(v) for v in x], 'x_sp': [float(v) for v in x_sp]})
    if rank == 0:
        with open(out_path, 'w') as f:
            json.dump({'world': world, 'backend': dist.get_backend(),
                       'rows_equal_single_rank': bool(np.array_equal(rows, rows_one)),
                       'rows2_equal': bool(np.array_equal(rows, rows2)),
                       'x': [float(v) for v in x], 'single_x': [float(v) for v in single[best].x],
                       'single_fun': single[best].fun, 'ok': bool(ok),
                       'per_rank': gathered, 'best_lattice_ll': float(rows[0, 0]),
                       'refined_ll': -float(est.likelihood_f(list(x)))}, f)
    dist.barrier()
    model.close()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
