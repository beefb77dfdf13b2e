#!/bin/bash
set -u
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
OUT=gpurun_out/r2t
mkdir -p $OUT
KMGPU_DEBUG=1 timeout 90 python tools/bench_configs.py NORM 2> $OUT/norm.err | cut -c1-300 | tee $OUT/progress.txt
grep "4000000 reads" $OUT/norm.err | tail -1 | tee -a $OUT/progress.txt
