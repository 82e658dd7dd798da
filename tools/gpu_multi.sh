#!/bin/bash
# N GPUs: the driver's bench command under torchrun
TAG=${1:-r02}; N=${2:-2}; EXTRA=${3:-}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_gpus.txt
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 $EXTRA > gpurun_out/${TAG}_bench${N}.json 2> gpurun_out/${TAG}_bench${N}.err
echo "bench$N rc=$?"; tail -c 2500 gpurun_out/${TAG}_bench${N}.json; grep -v "^$" gpurun_out/${TAG}_bench${N}.err | tail -5 | cut -c1-300
