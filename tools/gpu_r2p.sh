#!/bin/bash
# round-2 GPU session P: normalization with pre-sized in-between buffers
set -u
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
OUT=gpurun_out/r2p
mkdir -p $OUT
echo "== normalization tests" | tee $OUT/progress.txt
timeout 900 python -m pytest -q -x -s -m gpu tests/test_gpu_normalize.py > $OUT/tests_norm.log 2>&1; echo "norm tests rc=$?" | tee -a $OUT/progress.txt
tail -3 $OUT/tests_norm.log | cut -c1-300 | tee -a $OUT/progress.txt
grep "normalize 1M" $OUT/tests_norm.log | tee -a $OUT/progress.txt
for i in 1 2; do
KMGPU_DEBUG=1 timeout 600 python tools/bench_configs.py NORM > $OUT/norm$i.json 2> $OUT/norm$i.err
cut -c1-400 $OUT/norm$i.json | tee -a $OUT/progress.txt
grep "4000000 reads" $OUT/norm$i.err | tail -1 | tee -a $OUT/progress.txt
done
