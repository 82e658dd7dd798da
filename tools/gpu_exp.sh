#!/bin/bash
# tuning experiment: the cfg3 / cfg5 lattices with the library rebuilt under extra -D flags
mkdir -p gpurun_out
for FLAGS in "-DCVF_PNQ=3" "-DCVF_PNQ=2"; do
python -m covest_b200.build --force $FLAGS > /dev/null 2>&1
echo "== flags: [$FLAGS]"
python tools/prof_lattice.py cfg3 4 2>&1 | tail -1 | cut -c1-260
python tools/prof_lattice.py cfg5 3 2>&1 | tail -1 | cut -c1-260
done
