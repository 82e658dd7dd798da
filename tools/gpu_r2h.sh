#!/bin/bash
TAG=${1:-r2h}
mkdir -p gpurun_out
for V in 1 2; do for O in 0 1; do
COVEST_B200_PREFIX_KERNEL=$V COVEST_B200_TILE_ORDER=$O timeout 300 python tools/prof_lattice.py cfg3 4 > gpurun_out/${TAG}_lattice_v${V}_o$O.log 2>&1
echo "cfg3 kernel $V order $O rc=$?"; tail -1 gpurun_out/${TAG}_lattice_v${V}_o$O.log | cut -c1-300
COVEST_B200_PREFIX_KERNEL=$V COVEST_B200_TILE_ORDER=$O timeout 300 python tools/prof_lattice.py cfg5 3 > gpurun_out/${TAG}_lattice5_v${V}_o$O.log 2>&1
echo "cfg5 kernel $V order $O rc=$?"; tail -1 gpurun_out/${TAG}_lattice5_v${V}_o$O.log | cut -c1-300
done; done
timeout 600 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_api.py tests/test_gpu_factored.py -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest.log
