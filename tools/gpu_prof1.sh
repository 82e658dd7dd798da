#!/bin/bash
# ncu capture of one kernel of the bench.  usage: tools/gpu_prof1.sh TAG KERNEL_REGEX [points]
TAG=${1:-t}
K=${2:-cvf_gemm}
NP=${3:-200000}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -o gpurun_out/${TAG}_prof_$K -f python bench.py --steps 1 --warmup 3 --no-cpu --points $NP > gpurun_out/${TAG}_ncu_$K.log 2>&1
echo "ncu $K rc=$?"
