#!/bin/bash
# ncu captures of the factored path's kernels at the bench size.  usage: tools/gpu_prof.sh TAG [points]
TAG=${1:-t}
NP=${2:-1000000}
mkdir -p gpurun_out
for K in cvf_prefix cvf_profile; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -o gpurun_out/${TAG}_prof_$K -f python bench.py --steps 1 --warmup 3 --no-cpu --points $NP > gpurun_out/${TAG}_ncu_$K.log 2>&1
echo "ncu $K rc=$?"
done
