#!/bin/bash
set -u
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
OUT=gpurun_out/r2u
mkdir -p $OUT
timeout 130 python -m pytest -q -x -m gpu tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_sharded_gpu.py -k "(test_gpu_vs_oracle_random and group and (Countgraph or Nodetable)) or more_than_ten or first_touch_log_replicas or (sharded_equals and (Countgraph or SmallCounttable))" > $OUT/tests.log 2>&1; echo "tests rc=$?" | tee $OUT/progress.txt
tail -3 $OUT/tests.log | cut -c1-200 | tee -a $OUT/progress.txt
