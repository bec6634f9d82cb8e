#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; OUT=gpurun_out
N=${1:-8}
nvidia-smi -L > $OUT/n${N}_gpus.txt
timeout -s KILL 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 > $OUT/bench_n$N.log 2>&1; tail -1 $OUT/bench_n$N.log | cut -c1-330
timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 tools/run_wsi.py --size 16384 > $OUT/wsi_16384_n$N.log 2>&1; tail -1 $OUT/wsi_16384_n$N.log
