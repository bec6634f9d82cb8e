#!/bin/bash
# Round 2, call z (2 GPUs): NCCL 1-GPU == 2-GPU bit identity after the apply_model changes, C4 on the bench clock at N = 2.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; OUT=gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_model.py -m gpu -q -p no:cacheprovider -k "sharded_slide" > $OUT/r02z_pytest_nccl.log 2>&1; tail -2 $OUT/r02z_pytest_nccl.log
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus 2 --steps 3 --warmup 3 > $OUT/r02z_bench_n2.log 2>&1; grep '^{' $OUT/r02z_bench_n2.log | tail -1 | cut -c1-2200
