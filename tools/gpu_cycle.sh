#!/bin/bash
# One measurement cycle on the GPU box: parity tests, a bench line, an ncu capture of the likelihood
# kernel.  usage: tools/gpu_cycle.sh TAG [points_for_ncu]
TAG=${1:-t}
NP=${2:-200000}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" 
tail -3 gpurun_out/${TAG}_pytest.log
python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}.err
echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
r=d['roofline']
print('ms_per_step',d['ms_per_step'],'kernel_ms',r['kernel_ms'],'frac',r['frac'],'peak',r['peak'],'e2e',d['e2e']['value'],'value',d['value'])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cv_loglik -s 3 -c 1 -o gpurun_out/${TAG}_prof_loglik -f python bench.py --steps 1 --warmup 3 --no-cpu --points $NP > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu rc=$?"
