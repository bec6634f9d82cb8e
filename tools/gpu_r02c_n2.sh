#!/bin/bash
# Round 2, call c (2 GPUs): the driver's N > 1 launch of bench.py (C4 slide on the clock), the NCCL bit-identity test,
# the reference arm of the same config.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; OUT=gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > $OUT/n2_gpus.txt 2>&1
timeout -s KILL 900 python -m pytest tests/test_gpu_model.py -m gpu -q -p no:cacheprovider -k "sharded_slide" > $OUT/r02c_pytest_nccl.log 2>&1; tail -3 $OUT/r02c_pytest_nccl.log
timeout -s KILL 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus 2 > $OUT/r02c_bench_n2.log 2>&1; grep '^{' $OUT/r02c_bench_n2.log | tail -1 | cut -c1-3500; tail -3 $OUT/r02c_bench_n2.log | cut -c1-300
timeout -s KILL 900 python bench.py --workload c4 --steps 2 --warmup 3 --quick --no-cpu-baseline > $OUT/r02c_bench_c4_n1.log 2>&1; grep '^{' $OUT/r02c_bench_c4_n1.log | tail -1 | cut -c1-1800
timeout -s KILL 300 python tools/bench_decode.py > $OUT/r02c_bench_decode.log 2>&1; cat $OUT/r02c_bench_decode.log
