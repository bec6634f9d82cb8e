#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; OUT=gpurun_out
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/run_wsi.py --size 8192 > $OUT/det2_n2.log 2>&1; tail -1 $OUT/det2_n2.log
timeout -s KILL 300 python tools/run_wsi.py --size 8192 > $OUT/det2_default.log 2>&1; tail -1 $OUT/det2_default.log
