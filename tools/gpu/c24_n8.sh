#!/bin/bash
# call 24 (8 GPUs): the target configuration (cfg 4, one EAGLE-shaped snapshot dealt to 8 ranks) with the final library
N=${1:-8}
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --workload cfg4 --steps 3 --warmup 3 \
  > gpurun_out/r02_n${N}_weak_cfg4_v8.json 2> gpurun_out/r02_n${N}_weak_cfg4_v8.err
echo "rc=$?"; tail -c 1500 gpurun_out/r02_n${N}_weak_cfg4_v8.json; tail -3 gpurun_out/r02_n${N}_weak_cfg4_v8.err
