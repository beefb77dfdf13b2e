#!/bin/bash
set -u
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
OUT=gpurun_out/r2s
mkdir -p $OUT
timeout 70 python -m pytest -q -x -m gpu tests/test_host_gpu.py -k "trim_low_abund" > $OUT/tests.log 2>&1; echo "tests rc=$?" | tee $OUT/progress.txt
tail -4 $OUT/tests.log | cut -c1-300 | tee -a $OUT/progress.txt
