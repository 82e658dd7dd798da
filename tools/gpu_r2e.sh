#!/bin/bash
TAG=${1:-r2e}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_api.py tests/test_gpu_factored.py tests/test_gpu_fullsize.py -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/${TAG}_pytest.log
for V in 3 2 1; do
COVEST_B200_PREFIX_KERNEL=$V timeout 300 python tools/prof_lattice.py cfg3 4 > gpurun_out/${TAG}_lattice_v$V.log 2>&1
echo "cfg3 v$V rc=$?"; tail -1 gpurun_out/${TAG}_lattice_v$V.log | cut -c1-330
COVEST_B200_PREFIX_KERNEL=$V timeout 300 python tools/prof_lattice.py cfg5 3 > gpurun_out/${TAG}_lattice5_v$V.log 2>&1
echo "cfg5 v$V rc=$?"; tail -1 gpurun_out/${TAG}_lattice5_v$V.log | cut -c1-330
done
COVEST_B200_PREFIX_KERNEL=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:cvf_prefix_kernel -s 2 -c 1 -o gpurun_out/${TAG}_prof_prefix3 -f python tools/prof_lattice.py cfg3 3 > gpurun_out/${TAG}_ncu_prefix3.log 2>&1
echo "ncu prefix3 rc=$?"
COVEST_B200_PREFIX_KERNEL=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:cvf_prefix2 -s 2 -c 1 -o gpurun_out/${TAG}_prof_prefix2 -f python tools/prof_lattice.py cfg3 3 > gpurun_out/${TAG}_ncu_prefix2.log 2>&1
echo "ncu prefix2 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cvf_profile -s 2 -c 1 -o gpurun_out/${TAG}_prof_profile -f python tools/prof_lattice.py cfg3 3 > gpurun_out/${TAG}_ncu_profile.log 2>&1
echo "ncu profile rc=$?"
