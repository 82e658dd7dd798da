#!/bin/bash
# Full measurement cycle on the GPU box: parity tests, both bench arms, the ncu launch list of the
# bench command and one `--set full` capture of the dominant kernel at the bench's own size.
# usage: tools/gpu_round.sh TAG
TAG=${1:-t}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/${TAG}_bench.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}.err
echo "reference rc=$?"; cat gpurun_out/${TAG}_bench_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cvf_prefix -s 3 -c 1 -o gpurun_out/${TAG}_prof_cvf_prefix -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_cvf_prefix.log 2>&1
echo "ncu prefix rc=$?"
