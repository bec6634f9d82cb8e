#!/bin/bash
# First-contact diagnostics on the GPU box: everything lands in gpurun_out/ (merged back by gpurun).
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
OUT=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/nvidia_smi.txt 2>&1
echo "== conv simt32" | tee $OUT/conv_checks.log
timeout -s KILL 300 python tests/gpu_conv_check.py simt32 >> $OUT/conv_checks.log 2>&1
echo "== conv tcgen05 (dump)" | tee -a $OUT/conv_checks.log
CPN_DUMP=$OUT timeout -s KILL 180 python tests/gpu_conv_check.py tcgen05 0 1 >> $OUT/conv_checks.log 2>&1
echo "rc=$?" >> $OUT/conv_checks.log
echo "== conv tcgen05 (all)" | tee -a $OUT/conv_checks.log
timeout -s KILL 300 python tests/gpu_conv_check.py tcgen05 >> $OUT/conv_checks.log 2>&1
echo "rc=$?" >> $OUT/conv_checks.log
echo "== pytest" 
timeout -s KILL 2400 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > $OUT/pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest.log
tail -40 $OUT/pytest.log
echo "== smoke"
timeout -s KILL 600 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "rc=$?" >> $OUT/smoke.log
echo "== bench fp32 (strict engine, 2 steps)"
timeout -s KILL 900 python bench.py --steps 2 --warmup 3 --precision fp32 --no-cpu-baseline > $OUT/bench_fp32.log 2>&1; echo "rc=$?" >> $OUT/bench_fp32.log
echo "== bench fp16"
timeout -s KILL 900 python bench.py --steps 5 --warmup 3 > $OUT/bench_fp16.log 2>&1; echo "rc=$?" >> $OUT/bench_fp16.log
tail -3 $OUT/bench_fp16.log
echo "== decode microbench"
timeout -s KILL 300 python tools/bench_decode.py > $OUT/bench_decode.log 2>&1; echo "rc=$?" >> $OUT/bench_decode.log
cat $OUT/bench_decode.log
cat $OUT/conv_checks.log | cut -c1-220
