#!/bin/bash
TAG=${1:-r2b}
mkdir -p gpurun_out
PARITY_EXTRA=--lattice timeout 1200 python tools/parity_report.py > gpurun_out/${TAG}_parity.json 2> gpurun_out/${TAG}_parity.err
echo "parity rc=$?"; cut -c1-330 gpurun_out/${TAG}_parity.json | head -40; tail -3 gpurun_out/${TAG}_parity.err
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
