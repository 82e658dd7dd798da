#!/bin/bash
# round-end measurement cycle on one B200: parity tests, both bench arms, other shapes, ncu launch list and captures
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest.log
PARITY_EXTRA=--lattice timeout 900 python tools/parity_report.py > gpurun_out/${TAG}_parity.json 2> gpurun_out/${TAG}_parity.err
echo "parity rc=$?"; grep -c '"bad": 0' gpurun_out/${TAG}_parity.json
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; tail -c 300 gpurun_out/${TAG}_bench.json
timeout 900 python bench.py --impl reference > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
echo "reference rc=$?"; cut -c1-200 gpurun_out/${TAG}_bench_reference.json
timeout 600 python bench.py --workload cfg4 --no-cpu > gpurun_out/${TAG}_bench_cfg4.json 2>> gpurun_out/${TAG}_bench.err
echo "cfg4 rc=$?"
timeout 600 python bench.py --workload cfg5 --no-cpu > gpurun_out/${TAG}_bench_cfg5_1gpu.json 2>> gpurun_out/${TAG}_bench.err
echo "cfg5 rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cvf_prefix_kernel -s 2 -c 1 -o gpurun_out/${TAG}_prof_prefix -f python tools/prof_lattice.py cfg3 3 > gpurun_out/${TAG}_ncu_prefix.log 2>&1
echo "ncu prefix rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cvf_profile_kernel -s 2 -c 1 -o gpurun_out/${TAG}_prof_k1 -f python tools/prof_lattice.py cfg3 3 > gpurun_out/${TAG}_ncu_k1.log 2>&1
echo "ncu k1 rc=$?"
M=gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor_subpipe_dmma.sum,sm__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum
COVEST_B200_PATH=gemm timeout 600 ncu --metrics $M --clock-control none -k regex:'cvf_gemm|cvf_profile|cvf_weights' -s 3 -c 3 --csv --log-file gpurun_out/${TAG}_dmma_metrics.csv python tools/prof_lattice.py cfg3 2 > gpurun_out/${TAG}_ncu_dmma.log 2>&1
echo "ncu dmma rc=$?"
timeout 600 ncu --metrics $M --clock-control none -k regex:'cv_loglik_kernel' -c 1 --csv --log-file gpurun_out/${TAG}_dmma_metrics_perpoint.csv python -c "
import numpy as np, sys
sys.path.insert(0, '.')
from covest_b200 import workload
from covest_b200.models import RepeatsModel
cfg = workload.CONFIGS['cfg3']; hist = workload.synthetic_histogram('cfg3')
m = RepeatsModel(cfg['k'], cfg['r'], hist, 0, max_error=8)
m.loglikelihood_batch(workload.random_box_points(cfg['theta'], 200000, 1))
" > gpurun_out/${TAG}_ncu_dmma2.log 2>&1
echo "ncu per-point rc=$?"
