#!/bin/bash
TAG=${1:-r2c}
mkdir -p gpurun_out
# smoke of the new prefix kernel first, under a short timeout (an mbarrier bug would hang)
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1
echo "smoke rc=$?"; tail -3 gpurun_out/${TAG}_smoke.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -25 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench_v2.json 2> gpurun_out/${TAG}_bench.err
echo "bench v2 rc=$?"; python - <<PY
import json
for name in ("gpurun_out/${TAG}_bench_v2.json",):
    b=json.load(open(name)); print(b["value"], b["ms_per_step"], b["phases"], b["e2e"]["value"], b["e2e_lattice"]["value"])
PY
COVEST_B200_PREFIX_KERNEL=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench_v1.json 2>> gpurun_out/${TAG}_bench.err
echo "bench v1 rc=$?"; python - <<PY
import json
b=json.load(open("gpurun_out/${TAG}_bench_v1.json")); print(b["value"], b["ms_per_step"], b["phases"], b["e2e"]["value"], b["e2e_lattice"]["value"])
PY
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --workload cfg5 > gpurun_out/${TAG}_bench_cfg5_v2.json 2>> gpurun_out/${TAG}_bench.err
echo "bench cfg5 v2 rc=$?"; python - <<PY
import json
b=json.load(open("gpurun_out/${TAG}_bench_cfg5_v2.json")); print(b["value"], b["ms_per_step"], b["phases"])
PY
tail -5 gpurun_out/${TAG}_bench.err
