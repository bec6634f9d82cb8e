#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; OUT=gpurun_out
for i in 1 2; do timeout -s KILL 300 python tools/run_wsi.py --size 8192 > $OUT/det_a$i.log 2>&1; tail -1 $OUT/det_a$i.log; done
CPN_COALESCE=0 timeout -s KILL 300 python tools/run_wsi.py --size 8192 > $OUT/det_nocoal.log 2>&1; tail -1 $OUT/det_nocoal.log
timeout -s KILL 300 python tools/run_wsi.py --size 8192 --batch 8 > $OUT/det_b8.log 2>&1; tail -1 $OUT/det_b8.log
