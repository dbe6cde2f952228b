#!/bin/bash
# multi-GPU call: N = number of GPUs of the box (passed as $1).  Strong scaling of ONE AqA2 snapshot (walk target split + NCCL
# all-reduce), the default weak/replica mode, and the EAGLE-shaped LPT-sharded snapshot (cfg 4) at 1.7e8 particles per GPU.
N=${1:-2}
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
run() { # tag, extra args
  tag=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@" \
    > gpurun_out/r02_n${N}_${tag}.json 2> gpurun_out/r02_n${N}_${tag}.err
  echo "== $tag rc=$?"; tail -c 900 gpurun_out/r02_n${N}_${tag}.json; tail -3 gpurun_out/r02_n${N}_${tag}.err
}
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
run strong_cfg2 --scaling strong --steps 3 --warmup 2
run weak_cfg4 --workload cfg4 --steps 3 --warmup 2
run weak_cfg2 --steps 3 --warmup 2
