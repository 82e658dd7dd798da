#!/usr/bin/env python
"""bench.py -- throughput of the batched log-likelihood path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload cfg3|cfg5]

Metric (BASELINE.json): log-likelihood evaluations per second counted as parameter point x
histogram bin.  A step is one pass of the hot path over one batch of candidate points: every rank
evaluates its slice of the candidate lattice (points resident in HBM), keeps its 64 best rows on
the device and -- for N > 1 -- all-gathers them (NCCL) so that every rank holds the global best
rows.  Weak scaling: the lattice is refined with N so that every rank always evaluates the same
number of points.

  value      device-resident inputs, CUDA-event timed, max over ranks
  e2e        the same batch through the C ABI with HOST buffers (pinned), copies in the timed region
  roofline   FP64 pipe, dominant kernel (cvf_prefix_kernel: running sums over the copy numbers per
             q-run, three-term combination and log-likelihood epilogue per point; with
             --path gemm cvf_gemm_kernel: copy weights x bin profiles on the FP64 tensor cores): its
             algorithmic flop per launch / its CUDA-event duration / the DFMA peak measured in this
             run (MEASURED_PEAKS.json has no FP64 figure).  `phases` breaks the
             whole evaluation down.
  cpu_baseline / --impl reference: the reference's own C module (oracle/_ref) driven by a Python
             restatement of models.py with a fork pool over all host cores, on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K_BEST = 64
METRIC = 'loglik_point_bin_evals_per_sec'
UNIT = 'point*bin/s'
POINTS_PER_RANK = {'cfg3': 1000000, 'cfg4': 100000, 'cfg5': 12500000}


def workload_axes(name, world):
    """cfg3: 40 x 25 x 10 x 10 x 10 = 1e6 points; with N ranks the
    coverage axis is refined N-fold (weak scaling).  cfg5: 1e8 points over 8 ranks."""
    from covest_b200 import workload
    theta = workload.CONFIGS[name]['theta']
    if name == 'cfg3':
        return workload.lattice_axes(theta, n_c=40 * world, n_e=25)
    if name == 'cfg4':  # k = 31, r = 150, 5000 bins, coverage 200: 10^5 points per rank
        return workload.lattice_axes(theta, n_c=10 * world, n_e=10)
    return workload.lattice_axes(theta, n_c=25 * world, n_e=50, n_q1=10, n_q2=10, n_q=100)


class ClockSampler:
    """SM clock and throttle reasons of one GPU while the timed region runs: NVML every 5 ms
    (pynvml), or nvidia-smi every 200 ms when NVML cannot be loaded."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')
    NAMES = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']

    def __init__(self, index):
        self.index = index
        self.sm, self.mx, self.reasons = [], [], set()
        self.source = None
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run_nvml(self):
        import pynvml as nv
        nv.nvmlInit()
        try:
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)))
            bits = {'hw_slowdown': getattr(nv, 'nvmlClocksEventReasonHwSlowdown', 0x8),
                    'hw_thermal_slowdown': getattr(nv, 'nvmlClocksEventReasonHwThermalSlowdown', 0x40),
                    'sw_thermal_slowdown': getattr(nv, 'nvmlClocksEventReasonSwThermalSlowdown', 0x20),
                    'sw_power_cap': getattr(nv, 'nvmlClocksEventReasonSwPowerCap', 0x4)}
            get_reasons = getattr(nv, 'nvmlDeviceGetCurrentClocksEventReasons', None) or \
                nv.nvmlDeviceGetCurrentClocksThrottleReasons
            self.source = 'nvml'
            while not self._stop.is_set():
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                mask = int(get_reasons(h))
                for name, bit in bits.items():
                    if mask & bit:
                        self.reasons.add(name)
                self._stop.wait(0.005)
        finally:
            nv.nvmlShutdown()

    def _run_smi(self):
        self.source = 'nvidia-smi'
        while not self._stop.is_set():
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                      '--format=csv,noheader,nounits'], capture_output=True,
                                     text=True, timeout=5).stdout.strip()
                r = [x.strip() for x in out.split(',')] if out else []
                if r and r[0].replace('.', '').isdigit():
                    self.sm.append(float(r[0]))
                if len(r) > 1 and r[1].replace('.', '').isdigit():
                    self.mx.append(float(r[1]))
                for i, name in enumerate(self.NAMES):
                    if len(r) > 3 + i and r[3 + i].lower().startswith('active'):
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.2)

    def _run(self):
        try:
            self._run_nvml()
        except Exception:
            if not self.sm:
                self._run_smi()

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        return {'sm_mhz': float(np.median(self.sm)) if self.sm else None,
                'sm_max_mhz': max(self.mx) if self.mx else None,
                'reasons': [n for n in self.NAMES if n in self.reasons], 'samples': len(self.sm),
                'source': self.source}


# ---------------------------------------------------------------------------------------------
# reference arm / cpu baseline
# ---------------------------------------------------------------------------------------------
def cpu_sample_points(axes, n, seed=4242):
    from covest_b200 import workload
    total = int(np.prod([len(a) for a in axes]))
    rng = np.random.default_rng(seed)
    idx = np.sort(rng.choice(total, size=n, replace=False))
    return np.vstack([workload.lattice_points(axes, first=int(i), count=1) for i in idx])


def cpu_reference_rate(hist, cfg, axes, n_points, cores):
    """point*bin/s of the reference's C module + models.py restatement over a fork pool."""
    from oracle import covest_oracle as orc
    kind = 'reference' if orc.ref_module() is not None else 'port'
    m = orc.Model(cfg['model'], cfg['k'], cfg['r'], hist, 0, max_error=8)
    pts = cpu_sample_points(axes, n_points)
    t0 = time.perf_counter()
    if kind == 'reference':
        orc.ref_loglik_batch(m, pts, processes=cores)
    else:
        m.loglik_batch(pts, mode=orc.FAITHFUL, threads=cores)
    dt = time.perf_counter() - t0
    return n_points * len(hist) / dt, dt, kind


def covest_end_to_end(cores):
    """Wall time of the estimator flow (process_histogram -> model -> CoverageEstimator, the body of
    `covest cfg2.hist -m repeat -k 21 -r 100 -sf 1`) on the synthetic cfg2 histogram: on the device,
    and with the likelihood routed to the reference's own C module over a fork pool of all host
    cores (the same batches, so both arms follow the same optimiser path)."""
    from covest_b200 import constants
    from covest_b200.covest import CoverageEstimator
    from covest_b200.histogram import process_histogram
    from covest_b200.models import RepeatsModel
    from oracle import covest_oracle as orc
    constants.VERBOSE = False
    with open(os.path.join(ROOT, 'tests', 'golden', 'e2e_golden.json')) as f:
        g = json.load(f)['cfg2_repeats']
    hist = {int(j): int(h) for j, h in g['hist']}

    def flow(cls):
        t0 = time.perf_counter()
        h2, tail, sf, gc, ge = process_histogram(hist, g['k'], g['r'], sample_factor=1)
        model = cls(g['k'], g['r'], h2, tail, max_error=8)
        guess = list(model.defaults)
        guess[:2] = gc, ge
        est = CoverageEstimator(model)
        x, ok = est.compute_coverage(guess)
        return time.perf_counter() - t0, [float(v) for v in x], est.evaluations

    class CpuRepeats(RepeatsModel):  # likelihood on the host cores, everything else unchanged
        def loglikelihood_batch(self, points):
            m = orc.Model('repeats', self.k, self.r, self.hist, self.tail, max_error=self.max_error)
            pts = np.asarray(points, dtype=np.float64).reshape(-1, self.param_count)
            if orc.ref_module() is not None:
                return orc.ref_loglik_batch(m, pts, processes=min(cores, len(pts)))
            return m.loglik_batch(pts, mode=orc.FAITHFUL, threads=cores)

    def flow_device(**kw):  # the reference's experiment invocations (tools/run_covest.py: -sp 16; -g)
        import random
        random.seed(7)
        t0 = time.perf_counter()
        h2, tail, sf, gc, ge = process_histogram(hist, g['k'], g['r'], sample_factor=1)
        model = RepeatsModel(g['k'], g['r'], h2, tail, max_error=8)
        guess = list(model.defaults)
        guess[:2] = gc, ge
        est = CoverageEstimator(model)
        x, ok = est.compute_coverage(guess, **kw)
        return time.perf_counter() - t0, float(x[0]), est.evaluations

    flow(RepeatsModel)  # warm-up: context creation, first launches
    dev_s, dev_x, dev_n = flow(RepeatsModel)
    cpu_s, cpu_x, cpu_n = flow(CpuRepeats)
    extra = {}
    try:
        s16, c16, n16 = flow_device(starting_points=16)
        sg, cg, ng = flow_device(starting_points=1, use_grid_search=True)
        extra = {'device_multi_start_16': {'seconds': s16, 'coverage': c16, 'evaluations': n16},
                 'device_grid_search': {'seconds': sg, 'coverage': cg, 'evaluations': ng}}
    except Exception as exc:
        extra = {'extra_error': repr(exc)}
    return {'workload': 'cfg2: repeats model k=21 r=100, %d bins, single start, L-BFGS-B' % len(hist),
            'device_s': dev_s, 'device_coverage': dev_x[0], 'device_evaluations': dev_n,
            'cpu_s': cpu_s, 'cpu_coverage': cpu_x[0], 'cpu_evaluations': cpu_n, 'cpu_cores': cores,
            'cpu_kind': 'reference' if orc.ref_module() is not None else 'port', **extra}


def reference_histogram(name):
    """The workload histogram for the CPU-only reference arm: drawn from the oracle's p_j (the
    product arm draws it from the device's, the same numbers to ~1e-15)."""
    from covest_b200 import workload
    from oracle import covest_oracle as orc
    c = workload.CONFIGS[name]
    probe = orc.Model(c['model'], c['k'], c['r'], {j: 1 for j in range(1, c['bins'] + 1)}, 0, max_error=8)
    p = probe.probs(list(c['theta']))
    rng = np.random.default_rng(c['seed'])
    h = rng.poisson(c['kmers'] * np.maximum(p, 0.0))
    return {j: int(v) for j, v in zip(range(1, c['bins'] + 1), h)}


def run_reference(args, rank, world):
    if rank != 0:
        return
    from covest_b200 import workload
    cfg = workload.CONFIGS[args.workload]
    axes = workload_axes(args.workload, world)
    hist = reference_histogram(args.workload)
    cores = os.cpu_count() or 1
    n = max(cores, 2 * cores if args.workload == 'cfg3' else cores)
    rates, secs, kind = [], [], 'reference'
    for step in range(args.warmup + args.steps):
        rate, dt, kind = cpu_reference_rate(hist, cfg, axes, n, cores)
        if step >= args.warmup:
            rates.append(rate)
            secs.append(dt)
    value = float(np.mean(rates))
    sample = '%d lattice points (seeded) x %d bins per step, fork pool of %d' % (n, len(hist), cores)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': float(np.mean(secs) * 1e3),
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64 (x87 f80 inside truncated_poisson)',
        'data': 'synthetic', 'config': workload_config(args.workload, world, len(hist), axes),
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': kind, 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def committed_traffic(kernel, workload_name, n_points):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from the committed
    `ncu --set full` capture of this same command (profiles/, per launch) -- only when this run
    launches the kernel on the batch the capture was taken on; else None."""
    path = os.path.join(ROOT, 'profiles', 'r01_ncu_traffic.json')
    try:
        with open(path) as f:
            for row in json.load(f):
                if row['kernel'] == kernel and row['workload'] == workload_name and row['points'] == n_points:
                    return {'bytes': row['dram_bytes'], 'source': row['source']}
    except (OSError, ValueError, KeyError):
        pass
    return None


def workload_config(name, world, n_bins, axes):
    lens = [len(a) for a in axes]
    from covest_b200 import workload
    c = workload.CONFIGS[name]
    return {'workload': '%s: repeats model k=%d r=%d, lattice %s = %d points x %d bins, %d per rank' % (
        name, c['k'], c['r'], 'x'.join(map(str, lens)), int(np.prod(lens)), n_bins, int(np.prod(lens)) // world),
        'bins': n_bins, 'points_per_rank': int(np.prod(lens)) // world, 'lattice': lens,
        'max_error': 8, 'k_best': K_BEST,
        'sharding': 'whole (coverage, error_rate) groups of the lattice dealt round-robin: rank r takes groups r, r+N, ...',
        'l2': 'flushed between timed steps (256 MiB write)'}


# ---------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------
def run_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from covest_b200 import parallel, workload
    from covest_b200.models import RepeatsModel

    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (the library has no CPU path)')
    torch.cuda.set_device(local_rank)
    os.environ['COVEST_B200_DEVICE'] = str(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    dev = torch.device('cuda', local_rank)

    cfg = workload.CONFIGS[args.workload]
    hist = workload.synthetic_histogram(args.workload)
    model = RepeatsModel(cfg['k'], cfg['r'], hist, 0, max_error=8)
    ctx = model.device_context
    if args.path != 'auto':
        ctx.set_path({'prefix': ctx.PATH_FACTORED_PREFIX, 'gemm': ctx.PATH_FACTORED_GEMM,
                      'direct': ctx.PATH_PER_POINT}[args.path])
    n_bins = len(hist)
    axes = workload_axes(args.workload, world)
    total = int(np.prod([len(a) for a in axes]))
    count = total // world
    if args.points:
        count = min(count, args.points)

    # the rank's candidate points, resident in HBM: whole (c, e) groups, dealt round-robin
    block = int(np.prod([len(a) for a in axes[2:]]))
    count -= count % block
    host_pts = workload.lattice_points(axes, first=rank, stride=world, count=count, block=block)
    dev_pts = torch.from_numpy(host_pts).to(dev)
    dev_ll = torch.empty(count, dtype=torch.float64, device=dev)
    pin_pts = torch.from_numpy(host_pts).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()

    terms = workload.term_counts(model, host_pts)
    counted_bins = int(sum(1 for v in hist.values() if v))
    work = workload.factored_flop(model, host_pts, n_bins, counted_bins)

    peak_tflops = ctx.fp64_peak(0) if rank == 0 else None
    peak_dmma = ctx.fp64_peak(1) if rank == 0 else None

    step_launches = [0]  # kernels of this library launched by the last step (counted by the library)

    def step_device():
        ctx.loglik(dev_pts, out=dev_ll, stream=stream)
        n_eval = ctx.last_launches()
        rows = ctx.topk(dev_ll, dev_pts, K_BEST, stream=stream)
        step_launches[0] = n_eval + ctx.last_launches() + (1 if world > 1 else 0)  # + the merge
        if world > 1:
            rows = parallel.merge_topk(parallel.allgather_rows(rows), K_BEST)
        return rows

    def step_host():
        ll, rows = ctx.loglik_topk(pin_pts.numpy(), K_BEST)
        if world > 1:
            rows = parallel.merge_topk(parallel.allgather_rows(torch.from_numpy(rows).to(dev)), K_BEST)
            rows = rows.cpu().numpy()
        return rows

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    barrier()

    # ---- timed: device-resident inputs ----
    ctx.set_timing(True)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kernel_ms, launches = [], 0
    with ClockSampler(local_rank) as clocks:
        for a, b in ev:
            flush.fill_(1)
            barrier()
            a.record(stream)
            rows = step_device()
            b.record(stream)
            barrier()
            launches += step_launches[0]
        step_ms = [a.elapsed_time(b) for a, b in ev]
        # the kernels of the evaluation alone, CUDA events around the launches on the launching stream
        phases = []
        for _ in range(args.steps):
            flush.fill_(1)
            torch.cuda.synchronize()
            ctx.loglik(dev_pts, out=dev_ll, stream=stream)
            torch.cuda.synchronize()
            kernel_ms.append(ctx.last_kernel_ms()[0])
            phases.append(ctx.last_path_info())
        # ---- timed: end to end with host buffers ----
        for _ in range(min(args.warmup, 2)):
            step_host()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            rows_host = step_host()
        barrier()
        e2e_s = time.perf_counter() - t0
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_value = world * count * n_bins * args.steps / float(t.item())
        # the same batch described by its axes (what grid.py's callers hand over): points generated
        # on the device, only the values and the best rows cross the bus
        lat_value = None
        if not args.points:
            def step_lattice():
                ll, rows = ctx.lattice_eval(axes, first=rank, stride=world, count=count, block=block, k_best=K_BEST)
                if world > 1:
                    rows = parallel.merge_topk(parallel.allgather_rows(torch.from_numpy(rows).to(dev)), K_BEST)
                return ll
            ll_lat = step_lattice()   # warm-up, holding a result as the timed loop does (the pinned
            ll_lat = step_lattice()   # output buffers are allocated once and then recycled)
            ll_lat = step_lattice()
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                ll_lat = step_lattice()
            barrier()
            t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            lat_value = world * count * n_bins * args.steps / float(t.item())

    total_ms = float(sum(step_ms))
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = world * count * n_bins * args.steps / (total_ms * 1e-3)

    best = rows if isinstance(rows, np.ndarray) else rows.cpu().numpy()
    if rank == 0:
        km = float(np.mean(kernel_ms))
        info = phases[-1]
        factored = info['path'] == 'factored'
        if factored:
            gemm_ms = float(np.mean([p['gemm_ms'] for p in phases]))
            dominant, dom_ms = info['kernel'], gemm_ms
            flop_launch = work['prefix_flop'] if dominant == 'cvf_prefix_kernel' else work['gemm_flop']
        else:
            flop_launch = workload.algorithmic_flop(n_bins, terms)
            dominant, dom_ms = 'cv_loglik_kernel', km
        achieved = flop_launch / (dom_ms * 1e-3) / 1e12
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': total_ms / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': workload_config(args.workload, world, n_bins, axes),
            'e2e': {'value': e2e_value, 'unit': UNIT,
                    'h2d_bytes_per_step': int(count * 5 * 8),
                    'd2h_bytes_per_step': int(count * 8 + K_BEST * 6 * 8)},
            'e2e_lattice': {'value': lat_value, 'unit': UNIT, 'h2d_bytes_per_step': int(8 * sum(len(a) for a in axes)),
                            'd2h_bytes_per_step': int(count * 8 + K_BEST * 6 * 8),
                            'note': 'cvb_lattice_eval: the batch handed over as its axes, values to a host buffer'},
            'gpu_launches': launches,
            'clocks': clocks.summary(),
            'roofline': {'bound': 'fp64', 'achieved': achieved, 'peak': peak_tflops, 'unit': 'TFLOP/s',
                         'frac': achieved / peak_tflops if peak_tflops else None,
                         'traffic': (committed_traffic(dominant, args.workload, count) or {}).get('bytes'),
                         'traffic_source': (committed_traffic(dominant, args.workload, count) or {}).get('source'),
                         'kernel': dominant, 'kernel_ms': dom_ms,
                         'flop_per_launch': flop_launch, 'mean_terms_per_point': float(terms.mean()),
                         'peak_source': 'register-resident DFMA chain measured in this run (cvb_fp64_peak); '
                                        'DMMA m8n8k4 chain: %.2f TFLOP/s' % peak_dmma,
                         'kernel_share_of_step': dom_ms * args.steps / total_ms if world == 1 else None,
                         # SURVEY.md section 8(d) for comparison: the work of the reference's own
                         # formulation of the same batch (an exp per mixture term and bin: 40 flop per
                         # (term, bin) + 64 per bin -- "primary"; a recurrence, 4 flop per (term, bin)
                         # + 64 per bin -- "secondary") over the time of the whole evaluation here
                         'survey_accounting': {
                             'primary_flop': float(n_bins * (40.0 * terms.sum() + 64.0 * len(terms))),
                             'secondary_flop': float(n_bins * (4.0 * terms.sum() + 64.0 * len(terms))),
                             'primary_equivalent_tflops': float(n_bins * (40.0 * terms.sum() + 64.0 * len(terms))) / (km * 1e-3) / 1e12,
                             'secondary_equivalent_tflops': float(n_bins * (4.0 * terms.sum() + 64.0 * len(terms))) / (km * 1e-3) / 1e12,
                             'gemm_formulation_flop': work['gemm_flop'] if factored else None,
                             'note': 'equivalent rates exceed the FP64 peak because the factored formulations do less work for the same values'}},
            'phases': ({'path': 'factored', 'kernel': info['kernel'], 'groups': info['groups'],
                        'q_runs': info['q_runs'], 'tiles': info['tiles'],
                        'profile_workspace_bytes': 8 * info['profile_doubles'],
                        'plan_ms': float(np.mean([p['plan_ms'] for p in phases])),
                        'profile_ms': float(np.mean([p['profile_ms'] for p in phases])),
                        'kernel_ms': gemm_ms, 'evaluation_ms': km,
                        'profile_flop': work['profile_flop'], 'kernel_flop': flop_launch,
                        'counted_bins': counted_bins, 'mean_copies_per_point': work['mean_copies'],
                        'evaluation_tflops': (work['profile_flop'] + flop_launch) / (km * 1e-3) / 1e12}
                       if factored else {'path': info['path'], 'evaluation_ms': km}),
            'best_row': [float(x) for x in best[0]],
        }
        if world == 1 and not args.no_cpu:
            cores = os.cpu_count() or 1
            ref_hist = {j: int(v) for j, v in hist.items()}
            n_cpu = 2 * cores
            rate, dt, kind = cpu_reference_rate(ref_hist, cfg, axes, n_cpu, cores)
            if dt < 10.0:  # aim for 10-30 s of CPU work
                n_cpu = int(min(128 * cores, max(n_cpu, n_cpu * 20.0 / max(dt, 1e-3))))
                rate, dt, kind = cpu_reference_rate(ref_hist, cfg, axes, n_cpu, cores)
            line['cpu_baseline'] = {'value': rate, 'unit': UNIT, 'cores': cores, 'kind': kind,
                                    'sample': '%d seeded lattice points x %d bins in %.1f s' % (n_cpu, n_bins, dt)}
            try:  # BASELINE.json's second figure: covest end-to-end wall time
                line['covest_e2e'] = covest_end_to_end(cores)
            except Exception as exc:  # never lose the throughput line over the extra figure
                line['covest_e2e'] = {'error': repr(exc)}
        print(json.dumps(line), flush=True)
    barrier()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='cfg3', choices=['cfg3', 'cfg4', 'cfg5'])
    ap.add_argument('--points', type=int, default=0, help='cap the points per rank (development)')
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--path', default='auto', choices=['auto', 'prefix', 'gemm', 'direct'],
                    help='evaluation path of the device arm (development; auto = what a user gets)')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, world, local_rank)


if __name__ == '__main__':
    main()
