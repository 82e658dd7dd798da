#!/bin/bash
# two GPUs: the NCCL test, the bench at N = 2 with the cfg5 sub-record forced (its code path before the 8-GPU run)
TAG=${1:-r2g}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_gpus.txt
timeout 900 python -m pytest tests/test_gpu_nccl.py tests/test_gpu_api.py -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/${TAG}_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --cfg5 > gpurun_out/${TAG}_bench2.json 2> gpurun_out/${TAG}_bench2.err
echo "bench2 rc=$?"; tail -c 1800 gpurun_out/${TAG}_bench2.json; tail -5 gpurun_out/${TAG}_bench2.err
timeout 300 python tools/prof_lattice.py cfg3 4 > gpurun_out/${TAG}_lattice.log 2>&1; tail -1 gpurun_out/${TAG}_lattice.log | cut -c1-300
