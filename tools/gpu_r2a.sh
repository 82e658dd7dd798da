#!/bin/bash
# round 2, first call: parity statistics on the 10^4-point goldens, the GPU suite, a short bench
TAG=${1:-r2a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${TAG}_smi.txt
nproc >> gpurun_out/${TAG}_smi.txt
PARITY_EXTRA=--lattice timeout 900 python tools/parity_report.py > gpurun_out/${TAG}_parity.json 2> gpurun_out/${TAG}_parity.err
echo "parity rc=$?"; cut -c1-400 gpurun_out/${TAG}_parity.json | head -30
timeout 1200 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_big_golden.py > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; tail -c 2500 gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
