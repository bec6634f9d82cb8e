#!/bin/bash
# Round 2, call t: sparse-heads row trimming A/B (same box), sparse-heads parity tests.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; OUT=gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_model.py -m gpu -q --timeout 600 -p no:cacheprovider -k "sparse" > $OUT/r02t_pytest_sparse.log 2>&1
echo "pytest rc=$?" >> $OUT/r02t_pytest_sparse.log; tail -5 $OUT/r02t_pytest_sparse.log
for rep in 1 2; do
  CPN_SPARSE_TRIM=0 timeout -s KILL 600 python bench.py --quick --no-cpu-baseline --steps 20 > $OUT/r02t_bench_notrim$rep.log 2>&1; tail -1 $OUT/r02t_bench_notrim$rep.log | cut -c1-260
  CPN_SPARSE_TRIM=1 timeout -s KILL 600 python bench.py --quick --no-cpu-baseline --steps 20 > $OUT/r02t_bench_trim$rep.log 2>&1; tail -1 $OUT/r02t_bench_trim$rep.log | cut -c1-260
done
