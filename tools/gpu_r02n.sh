#!/bin/bash
# Round 2, call n: staged plan (score head early, count read-back overlapped with the full-resolution branch) -- parity + A/B.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; OUT=gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -x -k "gate_passing or batch_invariance or variant or sparse or input_contract or cuda_graph or apply_model or full_size_c3 or WideU22 or default_init" > $OUT/r02n_pytest.log 2>&1
rc=$?; echo "pytest rc=$rc" >> $OUT/r02n_pytest.log; tail -6 $OUT/r02n_pytest.log
if [ $rc -ne 0 ]; then exit 0; fi
env CPN_STAGED=0 timeout -s KILL 600 python bench.py --quick --no-cpu-baseline > $OUT/r02n_bench_unstaged.log 2>&1; tail -1 $OUT/r02n_bench_unstaged.log | cut -c1-230
timeout -s KILL 600 python bench.py --quick --no-cpu-baseline > $OUT/r02n_bench_staged.log 2>&1; tail -1 $OUT/r02n_bench_staged.log | cut -c1-230
env CPN_STAGED=0 timeout -s KILL 600 python bench.py --quick --no-cpu-baseline > $OUT/r02n_bench_unstaged2.log 2>&1; tail -1 $OUT/r02n_bench_unstaged2.log | cut -c1-230
timeout -s KILL 600 python bench.py --quick --no-cpu-baseline > $OUT/r02n_bench_staged2.log 2>&1; tail -1 $OUT/r02n_bench_staged2.log | cut -c1-230
grep -o '"e2e": {[^}]*}' $OUT/r02n_bench_unstaged.log $OUT/r02n_bench_staged.log $OUT/r02n_bench_unstaged2.log $OUT/r02n_bench_staged2.log | cut -c1-200
