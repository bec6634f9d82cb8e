#!/bin/bash
# Round 2, call l (2 GPUs): final bench of both arms at N = 1 and N = 2, NCCL bit-identity test.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; OUT=gpurun_out
timeout -s KILL 900 python bench.py > $OUT/r02q_bench_n1.log 2>&1; tail -1 $OUT/r02q_bench_n1.log | cut -c1-700
timeout -s KILL 900 python -m pytest tests/test_gpu_model.py -m gpu -q -p no:cacheprovider -k "sharded_slide" > $OUT/r02q_pytest_nccl.log 2>&1; tail -2 $OUT/r02q_pytest_nccl.log
timeout -s KILL 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus 2 > $OUT/r02q_bench_n2.log 2>&1; grep '^{' $OUT/r02q_bench_n2.log | tail -1 | cut -c1-2600
timeout -s KILL 900 python bench.py --workload c4 --steps 2 --warmup 3 --quick --no-cpu-baseline > $OUT/r02q_bench_c4_n1.log 2>&1; grep '^{' $OUT/r02q_bench_c4_n1.log | tail -1 | cut -c1-1500
timeout -s KILL 600 python bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > $OUT/r02q_bench_reference_n2.log 2>&1; tail -1 $OUT/r02q_bench_reference_n2.log | cut -c1-300
