#!/bin/bash
TAG=${1:-r2d}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -12 gpurun_out/${TAG}_pytest.log
for V in 2 1; do
COVEST_B200_PREFIX_KERNEL=$V timeout 300 python tools/prof_lattice.py cfg3 4 > gpurun_out/${TAG}_lattice_v$V.log 2>&1
echo "lattice v$V rc=$?"; tail -2 gpurun_out/${TAG}_lattice_v$V.log
done
COVEST_B200_PREFIX_KERNEL=2 timeout 300 python tools/prof_lattice.py cfg5 3 > gpurun_out/${TAG}_lattice5_v2.log 2>&1; tail -1 gpurun_out/${TAG}_lattice5_v2.log
COVEST_B200_PREFIX_KERNEL=1 timeout 300 python tools/prof_lattice.py cfg5 3 > gpurun_out/${TAG}_lattice5_v1.log 2>&1; tail -1 gpurun_out/${TAG}_lattice5_v1.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; tail -c 600 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
COVEST_B200_PREFIX_KERNEL=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:cvf_prefix2 -s 2 -c 1 -o gpurun_out/${TAG}_prof_prefix2 -f python tools/prof_lattice.py cfg3 3 > gpurun_out/${TAG}_ncu_prefix2.log 2>&1
echo "ncu prefix2 rc=$?"
COVEST_B200_PREFIX_KERNEL=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:cvf_prefix_kernel -s 2 -c 1 -o gpurun_out/${TAG}_prof_prefix1 -f python tools/prof_lattice.py cfg3 3 > gpurun_out/${TAG}_ncu_prefix1.log 2>&1
echo "ncu prefix1 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cv_faithful -s 2 -c 1 -o gpurun_out/${TAG}_prof_faithful -f python tools/prof_lattice.py cfg3 3 > gpurun_out/${TAG}_ncu_faithful.log 2>&1
echo "ncu faithful rc=$?"
