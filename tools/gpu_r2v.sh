#!/bin/bash
set -u
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
OUT=gpurun_out/r2v
mkdir -p $OUT
timeout 100 python tools/bench_configs.py --no-queries C4 C4S 2> $OUT/configs.err | cut -c1-300 | tee $OUT/progress.txt
timeout 100 python bench.py --mode sharded --steps 4 --warmup 2 2> $OUT/sh.err | cut -c1-1400 | tee -a $OUT/progress.txt
