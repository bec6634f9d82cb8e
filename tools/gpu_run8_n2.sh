#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; OUT=gpurun_out
nvidia-smi -L > $OUT/n2_gpus.txt
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $OUT/bench_n2.log 2>&1; tail -1 $OUT/bench_n2.log | cut -c1-400
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/run_wsi.py --size 8192 > $OUT/wsi_8192_n2.log 2>&1; tail -1 $OUT/wsi_8192_n2.log
timeout -s KILL 900 python tools/run_wsi.py --size 8192 > $OUT/wsi_8192_n1b.log 2>&1; tail -1 $OUT/wsi_8192_n1b.log
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 tools/run_wsi.py --size 16384 > $OUT/wsi_16384_n2.log 2>&1; tail -1 $OUT/wsi_16384_n2.log
