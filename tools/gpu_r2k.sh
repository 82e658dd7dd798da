#!/bin/bash
TAG=${1:-r2k}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/${TAG}_pytest.log
timeout 300 python tools/prof_lattice.py cfg3 4 > gpurun_out/${TAG}_lattice.log 2>&1; tail -1 gpurun_out/${TAG}_lattice.log | cut -c1-320
timeout 600 python bench.py --workload cfg4 --no-cpu > gpurun_out/${TAG}_bench_cfg4.json 2> gpurun_out/${TAG}_bench.err
echo "cfg4 rc=$?"; python - <<PY
import json
b=json.loads([l for l in open("gpurun_out/${TAG}_bench_cfg4.json") if l.startswith("{")][-1]); print(b["value"], b["ms_per_step"], b["phases"])
PY
PARITY_EXTRA= timeout 900 python tools/parity_report.py cfg4 cfg2 > gpurun_out/${TAG}_parity.json 2> gpurun_out/${TAG}_parity.err
echo "parity rc=$?"; cut -c1-260 gpurun_out/${TAG}_parity.json
