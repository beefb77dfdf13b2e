#!/bin/bash
set -u
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
OUT=gpurun_out/r2r
mkdir -p $OUT
timeout 100 python -m pytest -q -x -m gpu tests/test_gpu_benchscale.py -k "group" > $OUT/tests.log 2>&1; echo "tests rc=$?" | tee $OUT/progress.txt
tail -3 $OUT/tests.log | cut -c1-300 | tee -a $OUT/progress.txt
