#!/bin/bash
# round-2 GPU session O: where the normalization's time goes (debug timers), twice
set -u
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
OUT=gpurun_out/r2o
mkdir -p $OUT
echo "== normalization, phases" | tee $OUT/progress.txt
for i in 1 2; do
KMGPU_DEBUG=1 timeout 600 python tools/bench_configs.py NORM > $OUT/norm$i.json 2> $OUT/norm$i.err
cut -c1-400 $OUT/norm$i.json | tee -a $OUT/progress.txt
grep normalize_batch $OUT/norm$i.err | tail -2 | tee -a $OUT/progress.txt
done




