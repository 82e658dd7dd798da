#!/bin/bash
TAG=${1:-r2f}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/${TAG}_pytest.log
bash tools/gpu_sanitize.sh ${TAG}
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; python - <<PY
import json
b=json.load(open("gpurun_out/${TAG}_bench.json")); print(b["value"], b["ms_per_step"], b["e2e"]["value"], b["roofline"]["frac"], b["phases"])
PY
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
echo "reference rc=$?"; cut -c1-400 gpurun_out/${TAG}_bench_reference.json
