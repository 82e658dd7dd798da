#!/bin/bash
# tuning experiment: the cfg3 / cfg5 lattices with the library rebuilt under extra -D flags
mkdir -p gpurun_out
for FLAGS in "-DCVF_SL=3" "-DCVF_SL=3 -DCVF_WARPS_SM=18" "-DCVF_WARPS_SM=18" "-DCVF_LOG_REP=4" "-DCVF_SL=3 -DCVF_WARPS_SM=18 -DCVF_LOG_REP=2"; do
python -m covest_b200.build --force $FLAGS > /dev/null 2>&1
echo "== flags: [$FLAGS]"
python tools/prof_lattice.py cfg3 4 2>&1 | tail -1 | cut -c1-60,200-300
python tools/prof_lattice.py cfg5 3 2>&1 | tail -1 | cut -c1-60,200-300
done
